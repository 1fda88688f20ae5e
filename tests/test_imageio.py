"""host/imageio.cpp: the texture-file decoder that stands in for the reference's stb_image (texture2d.cpp:31-52).  Fixtures and the
expected texels -- produced by the reference's own vendored stb_image.h, compiled where it lies -- are in tests/golden/images (make_golden_images.py)."""
import json
import os

import numpy as np
import pytest

from oracle import pyref

IMAGES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "images")
EXPECTED = np.load(os.path.join(IMAGES, "expected.npz"))


@pytest.mark.parametrize("name", sorted(EXPECTED.files))
def test_decoder_matches_stb_image(g, name):
    got = g.load_image(os.path.join(IMAGES, name))
    want = EXPECTED[name]
    assert got.shape == want.shape
    if name.endswith(".jpg"):
        # JPEG decoders may differ in the last bits of the inverse DCT / chroma upsampling / colour conversion (T.81 allows +-1 per stage)
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 3 and d.mean() < 0.1, (d.max(), d.mean())
        assert np.array_equal(got[..., 3], want[..., 3])
    else:
        assert np.array_equal(got, want)


def test_unsupported_and_corrupt_files_are_refused(g, tmp_path):
    with pytest.raises(g.GpugiError):
        g.load_image(str(tmp_path / "missing.png"))
    blob = open(os.path.join(IMAGES, "rgb8.png"), "rb").read()
    (tmp_path / "cut.png").write_bytes(blob[: len(blob) // 2])
    with pytest.raises(g.GpugiError):
        g.load_image(str(tmp_path / "cut.png"))
    (tmp_path / "noise.bin").write_bytes(bytes(range(7, 200)))
    with pytest.raises(g.GpugiError):
        g.load_image(str(tmp_path / "noise.bin"))
    jpg = bytearray(open(os.path.join(IMAGES, "q90_444.jpg"), "rb").read())
    i = jpg.find(b"\xff\xc0")
    jpg[i + 1] = 0xC2   # pretend the frame is progressive
    (tmp_path / "prog.jpg").write_bytes(bytes(jpg))
    with pytest.raises(g.GpugiError) as e:
        g.load_image(str(tmp_path / "prog.jpg"))
    assert "progressive" in str(e.value)


def test_mutated_files_never_crash_and_huge_headers_are_refused(g, tmp_path):
    """Truncations and byte flips of every fixture either decode to width x height x 4 texels or are refused with an error (the same
    mutations ran 60 000 times under ASan / UBSan while this was written); a header that claims more texels than stb_image's own limit
    (width * height * 4 must fit an int) or than its compressed bytes can hold is refused before anything is allocated."""
    import struct
    rng = np.random.RandomState(5)
    path = str(tmp_path / "m.bin")
    decoded = refused = 0
    for name in sorted(EXPECTED.files):
        blob = open(os.path.join(IMAGES, name), "rb").read()
        for k in range(40):
            d = bytearray(blob)
            if k % 4 == 0:
                d = d[: int(rng.randint(0, len(d)))]
            for _ in range(int(rng.randint(0, 6)) + (k % 4 != 0)):
                if d:
                    d[int(rng.randint(0, min(len(d), 64) if k % 4 == 2 else len(d)))] = int(rng.randint(0, 256))
            open(path, "wb").write(bytes(d))
            try:
                img = g.load_image(path)
                assert img.ndim == 3 and img.shape[2] == 4 and img.size > 0
                decoded += 1
            except g.GpugiError:
                refused += 1
    assert decoded > 100 and refused > 100, (decoded, refused)
    png = bytearray(open(os.path.join(IMAGES, "rgb8.png"), "rb").read())
    for w, h in ((1 << 24, 1 << 24), (40000, 40000), (20000, 20000)):   # the last one passes stb's limit but not the IDAT bound
        png[16:24] = struct.pack(">II", w, h)
        open(path, "wb").write(bytes(png))
        with pytest.raises(g.GpugiError):
            g.load_image(path)
    # header fields the pixel loops index with (found by 400 000 header-only mutations under ASan / UBSan): a BMP bit depth that is not 1 / 2 / 4 / 8 /
    # 16 / 24 / 32, a JPEG component sampled 2 : 3 against the largest one (its plane is smaller than the 1 : 1 path assumed)
    bmp = bytearray(open(os.path.join(IMAGES, "pal8.bmp"), "rb").read())
    bmp[28:30] = struct.pack("<H", 3)
    open(path, "wb").write(bytes(bmp))
    with pytest.raises(g.GpugiError):
        g.load_image(path)
    jpg = bytearray(open(os.path.join(IMAGES, "q90_444.jpg"), "rb").read())
    i = jpg.find(b"\xff\xc0")
    assert jpg[i + 11] == 0x11 and jpg[i + 14] == 0x11     # Y and Cb: 1 x 1
    jpg[i + 11], jpg[i + 14] = 0x33, 0x22                   # Y 3 x 3, Cb 2 x 2
    open(path, "wb").write(bytes(jpg))
    with pytest.raises(g.GpugiError):
        g.load_image(path)
    tga = bytearray(open(os.path.join(IMAGES, "rgb24.tga"), "rb").read())
    tga[12:16] = struct.pack("<HH", 65535, 65535)
    open(path, "wb").write(bytes(tga))
    with pytest.raises(g.GpugiError):
        g.load_image(path)


@pytest.mark.skipif(not pyref.stb_available(), reason="oracle/_ref/libref_stb.so not built here (needs /root/reference)")
def test_decoder_vs_stb_image_live(g, tmp_path):
    """Fresh images (other sizes, odd widths, 16-bit RGB(A), alpha in palettes, bottom-up / top-down TGA, RLE runs across rows) through both decoders."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(9)
    for k in range(12):
        w, h = int(rng.randint(1, 70)), int(rng.randint(1, 50))
        a = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
        a[:, : w // 2] = a[0, 0]    # flat areas: RLE runs, PNG filters that pay
        variants = [("x.png", Image.fromarray(a), {}), ("y.png", Image.fromarray(a[..., :3]), {}),
                    ("x.bmp", Image.fromarray(a[..., :3]), {}), ("x.tga", Image.fromarray(a), {"compression": "tga_rle"}), ("y.tga", Image.fromarray(a[..., 0]), {}),
                    ("z.png", Image.fromarray(a[..., 0]), {})]
        for name, img, kw in variants:
            path = str(tmp_path / name)
            img.save(path, **kw)
            ref = pyref.stbi_load(path)
            assert ref is not None
            assert np.array_equal(g.load_image(path), ref[0]), (k, name, w, h)
        if w >= 8 and h >= 8:
            for sub in (0, 1, 2):
                path = str(tmp_path / "x.jpg")
                smooth = np.clip(np.cumsum(np.cumsum(rng.randn(h, w, 3), 0), 1) * 4 + 128, 0, 255).astype(np.uint8)
                Image.fromarray(smooth).save(path, quality=int(rng.randint(40, 96)), subsampling=sub)
                d = np.abs(g.load_image(path).astype(int) - pyref.stbi_load(path)[0].astype(int))
                assert d.max() <= 4 and d.mean() < 0.2, (k, sub, d.max(), d.mean())


def test_scene_textures_decode_through_imageio(g, scenes, tmp_path):
    """Scene::GetBindlessHandle (scene.cpp:182-199): a material whose albedo names a PNG / JPEG / TGA file gets the file's texels as an
    SRGB8_ALPHA8 texture -- the same GiTexture as the equivalent PPM, which the synthetic scenes used so far."""
    import ctypes as C
    import shutil
    for ext in ("png", "tga", "bmp", "ppm"):
        src = {"png": "rgb8.png", "tga": "rgb24.tga", "bmp": "rgb24.bmp", "ppm": "rgb.ppm"}[ext]
        shutil.copy(os.path.join(IMAGES, src), str(tmp_path / ("albedo." + ext)))
    pos, nrm, uv, tri, mat = scenes.box1_mesh(1)
    want = EXPECTED["rgb8.png"]
    for ext in ("png", "tga", "bmp", "ppm"):
        mats = {n: scenes.legacy_material((0.5, 0.5, 0.5)) for n in scenes.BOX1_MATERIAL_NAMES}
        mats["Floor"] = scenes.legacy_material(albedo_texture="albedo." + ext)
        sc = g.Scene.from_mesh(pos, nrm, uv, tri, mat, mats, str(tmp_path) + "/")
        d = sc.desc
        tex = [d.textures[i] for i in range(d.num_textures)]
        floor = d.materials[scenes.BOX1_MATERIAL_NAMES.index("Floor")]
        t = tex[floor.diffuse_tex]
        assert (t.width, t.height, t.format) == (want.shape[1], want.shape[0], 0)   # GI_TEX_SRGB8_ALPHA8
        px = np.ctypeslib.as_array(C.cast(t.pixels, C.POINTER(C.c_uint8)), shape=(t.height, t.width, 4))
        assert np.array_equal(px, want), ext
