"""Host layer (C++ mirror of Scene / LightSampler / Camera / PFM / bim) -- no GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_hierarchy(scene):
    """Invariants of the firstChild/escape preorder layout (bim_hierarchy.cpp:238-252, chunk.hpp:39-44)."""
    nodes = scene.nodes()
    lt = scene.leaf_triangles()
    P = scene.positions()
    parents = scene.parents()
    n = len(nodes)
    tris = scene.triangles()
    # every input triangle appears exactly once among the valid leaf slots
    valid = lt["v"][:, 0] != lt["v"][:, 1]
    got = sorted(map(tuple, lt["v"][valid].tolist()))
    want = sorted(map(tuple, tris[tris[:, 0] != tris[:, 1]].tolist()))
    assert got == want
    # unused slots are all-zero and only ever the second slot of a leaf
    inval = np.nonzero(~valid)[0]
    assert np.all(inval % 2 == 1) and np.all(lt["v"][inval] == 0) and np.all(lt["material"][inval] == 0)
    assert nodes["escape"][0] == 0
    seen_leaves = set()
    for i in range(n):
        fc, esc = int(nodes["first_child"][i]), int(nodes["escape"][i])
        assert esc == 0 or esc > i  # escape always points forward in preorder
        if fc & 0x80000000:
            leaf = fc & 0x7FFFFFFF
            assert leaf not in seen_leaves
            seen_leaves.add(leaf)
            idx = [2 * leaf] + ([2 * leaf + 1] if valid[2 * leaf + 1] else [])
            pts = P[lt["v"][idx].reshape(-1)]
            assert np.array_equal(nodes["min"][i], pts.min(axis=0)) and np.array_equal(nodes["max"][i], pts.max(axis=0))
        else:
            assert fc == i + 1  # first child follows its parent in preorder
            right = int(nodes["escape"][fc])
            assert parents[fc] == i and parents[right] == i and int(nodes["escape"][right]) == esc
            lo = np.minimum(nodes["min"][fc], nodes["min"][right])
            hi = np.maximum(nodes["max"][fc], nodes["max"][right])
            assert np.array_equal(nodes["min"][i], lo) and np.array_equal(nodes["max"][i], hi)
    assert len(seen_leaves) == len(lt) // 2 and n == 2 * len(seen_leaves) - 1


@pytest.mark.parametrize("method", ["sah", "kd"])
def test_hierarchy_invariants(scenes, method):
    _check_hierarchy(scenes.make_random_triangles(500, seed=3, build_method=method))
    _check_hierarchy(scenes.make_box1(tess=2, build_method=method))


def test_single_triangle_and_odd_counts(scenes, g, orc):
    """Ragged inputs: 1 triangle (root is a leaf), 3 triangles (a half-filled leaf)."""
    for n in (1, 2, 3, 5):
        sc = scenes.make_random_triangles(n, seed=n)
        info = sc.info
        assert info.num_triangles == n and info.num_leaf_triangles % 2 == 0
        rays = np.zeros(64, dtype=g.RAY_DTYPE)
        rng = np.random.RandomState(n)
        rays["origin"] = rng.uniform(-12, 12, size=(64, 3)).astype(np.float32)
        d = -rays["origin"] + rng.uniform(-1, 1, size=(64, 3))
        rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        rays["t_max"] = np.float32(g.RAY_MAX)
        orc.trace_batch(sc.desc, rays)  # must terminate


def test_bim_file_round_trip(scenes, g, tmp_path):
    """store (.json + .bim, zlib sections) -> Scene(file) reproduces every uploaded array bit for bit."""
    sc = scenes.make_box1(tess=2)
    sc.add_point_light((0.1, 2.0, -1.0), (5.0, 6.0, 7.0))
    json_path = str(tmp_path / "box1.json")
    sc.store(json_path, "box1.bim")
    # section header: type u32 (+4 pad), size u64, uncompressedSize u64 = 24 bytes; META_SECTION first with size 40 (bim_io.cpp:36-56)
    raw = open(str(tmp_path / "box1.bim"), "rb").read()
    typ, _, size, usize = np.frombuffer(raw[:24], dtype=[("t", "<u4"), ("p", "<u4"), ("s", "<u8"), ("u", "<u8")])[0]
    assert (typ, size, usize) == (0, 40, 0)
    nchunks = np.frombuffer(raw[24:36], dtype="<i4")
    assert list(nchunks) == [1, 1, 1] and np.frombuffer(raw[60:64], dtype="<u4")[0] == 2  # numTrianglesPerLeaf
    mtyp = np.frombuffer(raw[64:68], dtype="<u4")[0]
    assert mtyp == 0x5  # MATERIAL_REFERENCE
    sc2 = g.Scene.load(json_path)
    for name in ("nodes", "leaf_triangles", "positions", "vertex_info", "triangles", "triangle_materials", "parents", "emissivity"):
        a, b = getattr(sc, name)(), getattr(sc2, name)()
        assert a.tobytes() == b.tobytes(), name
    assert sc2.info.num_tree_levels == sc.info.num_tree_levels
    d1, d2 = sc.desc, sc2.desc
    assert d1.num_materials == d2.num_materials == 9 and d1.num_textures == d2.num_textures
    m1 = np.ctypeslib.as_array(C.cast(d1.materials, C.POINTER(C.c_uint8)), shape=(d1.num_materials * 48,))
    m2 = np.ctypeslib.as_array(C.cast(d2.materials, C.POINTER(C.c_uint8)), shape=(d2.num_materials * 48,))
    assert m1.tobytes() == m2.tobytes()


def test_scene_load_errors(g, tmp_path):
    with pytest.raises(g.GpugiError):
        g.Scene.load(str(tmp_path / "does_not_exist.json"))
    bad = tmp_path / "bad.json"
    bad.write_text('{"materials": {}, "scene": "missing.bim"}')
    with pytest.raises(g.GpugiError):
        g.Scene.load(str(bad))


def test_reference_material_json_is_readable(g, scenes, tmp_path):
    """The material table of scenes/box1/box1.json is embedded verbatim in scenes.box1_materials(): check the
    derived GPU material values (scene.cpp:147-180): exponent = 1/(roughness^2+1e-20), fresnel1 = 1 - fresnel0."""
    sc = scenes.make_box1(tess=1)
    d = sc.desc
    mats = [d.materials[i] for i in range(d.num_materials)]
    tex = [d.textures[i] for i in range(d.num_textures)]
    glass = mats[scenes.BOX1_MATERIAL_NAMES.index("Glass")]
    assert np.allclose(list(glass.fresnel0), [0.035] * 3) and np.allclose(list(glass.fresnel1), [0.965] * 3) and glass.refraction_index_avg == 1.5
    refl = tex[glass.reflectiveness_tex]
    assert (refl.width, refl.height, refl.format) == (1, 1, 2)
    vals = np.ctypeslib.as_array(C.cast(refl.pixels, C.POINTER(C.c_float)), shape=(4,))
    assert np.allclose(vals, [1, 1, 1, 1.0 / (0.01 * 0.01 + 1e-20)])
    # albedo constants are RGB8: 0.5882 -> 150/255
    dif = tex[glass.diffuse_tex]
    assert dif.format == 1 and list(np.ctypeslib.as_array(C.cast(dif.pixels, C.POINTER(C.c_uint8)), shape=(4,))) == [150, 150, 150, 255]
    lightr = mats[scenes.BOX1_MATERIAL_NAMES.index("LightR")]
    em = tex[lightr.emissivity_tex]
    assert np.allclose(np.ctypeslib.as_array(C.cast(em.pixels, C.POINTER(C.c_float)), shape=(4,))[:3], [803.9, 192.2, 19.6])
    assert np.allclose(sc.emissivity()[scenes.BOX1_MATERIAL_NAMES.index("LightL")], [98.0, 392.2, 788.2])


def test_light_sampler_matches_oracle(g, orc, scenes):
    """LightSampler::GenerateRandomSamples + light tables: host C++ vs the oracle's restatement, bit for bit."""
    sc = scenes.make_box1(tess=3)
    for pl in ([], [((0.2, 1.0, -1.0), (3.0, 4.0, 5.0)), ((-0.5, 2.0, 0.5), (50.0, 10.0, 1.0))]):
        for p, i in pl:
            sc.add_point_light(p, i)
        a = g.LightSampler(sc)
        b = orc.LightSampler(sc.positions(), sc.triangles(), sc.triangle_materials(), sc.emissivity(), sc.point_lights())
        assert b.num_light_triangles() == sc.info.num_light_triangles == 16
        for n in (128, 256, 7):
            assert a.generate(n).tobytes() == b.generate(n).tobytes()
        a.seed(99); b.seed(99)
        assert a.generate(64).tobytes() == b.generate(64).tobytes()
    s = a.generate(128)
    assert (s["normal_theta_cos"] > 1.5).sum() > 0 and (s["normal_theta_cos"] < 1.5).sum() > 0  # both kinds present
    # scene without any light: black dummy samples (lightsampler.cpp:21-31)
    dark = scenes.make_random_triangles(10, seed=5)
    assert not g.LightSampler(dark).generate(16).view(np.uint8).any()


def test_camera_matches_oracle(g, orc):
    for pos, look, fov, w, h in [((0.0, 1.5, 2.27), (0.0, 1.5, 2.0), 67.0, 512, 512), ((784.069214, 525.171387, -123.871895), (-784.925598, -525.603210, 124.154999), 70.0, 1920, 1080)]:
        a = np.frombuffer(g.camera_params(pos, look, fov, w, h), dtype=np.float32)
        assert np.array_equal(a, orc.camera(pos, look, fov, w, h))
        U, V, W = a[3:6], a[6:9], a[9:12]
        assert abs(np.dot(U, W)) < 1e-5 and abs(np.dot(V, W)) < 1e-5 and abs(np.linalg.norm(W) - 1) < 1e-6
        assert np.isclose(np.linalg.norm(U) / np.linalg.norm(V), w / h, rtol=1e-5)


def test_pfm_round_trip(g, tmp_path):
    rng = np.random.RandomState(0)
    img = rng.rand(5, 7, 4).astype(np.float32) * 100
    path = str(tmp_path / "x.pfm")
    g.write_pfm(path, img, divisor=4)
    raw = open(path, "rb").read()
    assert raw.startswith(b"PF\n7 5\n-1.000000\n") and len(raw) == len(b"PF\n7 5\n-1.000000\n") + 5 * 7 * 12
    back = g.load_pfm(path)
    assert np.array_equal(back, img[..., :3] / np.float32(4))
    with pytest.raises(g.GpugiError):
        g.load_pfm(str(tmp_path / "nope.pfm"))


def test_pfm_matches_reference_golden_file(g, tmp_path):
    """tests/golden/ref_pfm_7x5_div3.pfm was written by the REFERENCE'S imageut/hdrimage.cpp (compiled where it lies into oracle/_ref, see
    tests/golden/make_golden_pfm.py) from golden_input(): the host layer's WritePfm must produce the same bytes and its LoadPfm the same pixels."""
    from tests.golden.make_golden_pfm import golden_input
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_pfm_7x5_div3.pfm")
    img = golden_input()
    path = str(tmp_path / "ours.pfm")
    g.write_pfm(path, img, divisor=3)
    assert open(path, "rb").read() == open(golden, "rb").read()
    back = g.load_pfm(golden)
    assert back.shape == (5, 7, 3)
    assert np.array_equal(back.view(np.uint32), (img[..., :3] / np.float32(3)).view(np.uint32))


def test_pfm_matches_reference_object_live(g, tmp_path):
    """Where oracle/_ref/libref_hdrimage.so exists: random sizes and divisors through the reference's WritePfm / LoadPfm and ours, both ways."""
    ref_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_hdrimage.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref not built here (needs /root/reference); the committed golden file covers the same functions")
    import ctypes as C
    from tests.golden.make_golden_pfm import load_ref
    ref = load_ref()
    rng = np.random.RandomState(8)
    for w, h, div in [(1, 1, 1), (8, 8, 64), (33, 17, 7), (64, 3, 1000)]:
        img = (rng.rand(h, w, 4) * 1000).astype(np.float32)
        a, b = str(tmp_path / "ref.pfm"), str(tmp_path / "ours.pfm")
        assert ref.ref_write_pfm(os.fsencode(a), img.ctypes.data, w, h, div) == 0
        g.write_pfm(b, img, divisor=div)
        assert open(a, "rb").read() == open(b, "rb").read()
        out = np.zeros((h, w, 3), dtype=np.float32)
        rw, rh = C.c_int(), C.c_int()
        assert ref.ref_load_pfm(os.fsencode(b), out.ctypes.data, w * h, C.byref(rw), C.byref(rh)) == 0 and (rw.value, rh.value) == (w, h)
        assert np.array_equal(out.view(np.uint32), g.load_pfm(a).view(np.uint32))


def test_c_abi_exports_every_declared_symbol(g):
    """Every function declared in include/*.h is exported by the library that implements it."""
    from gpugi_b200 import _capi
    for header, lib, table in (("gpugi_b200.h", _capi.dev, _capi.DEVICE_API), ("gpugi_b200_host.h", _capi.host, _capi.HOST_API)):
        text = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"GI_API\s+[\w\s\*]+?\b(gi[h]?_\w+)\s*\(", text))
        assert declared, header
        for name in declared:
            assert hasattr(lib, name), "%s declares %s but the library does not export it" % (header, name)
        assert declared == set(table.keys()), (declared ^ set(table.keys()))
    assert _capi.dev.gi_version() == 100
    # POD sizes the ABI promises
    assert C.sizeof(_capi.GiNode) == 32 and C.sizeof(_capi.GiTriangle) == 16 and C.sizeof(_capi.GiVertexInfo) == 16
    assert C.sizeof(_capi.GiLightSample) == 32 and C.sizeof(_capi.GiRay) == 32 and C.sizeof(_capi.GiHit) == 16


@pytest.mark.skipif(__import__("tests.conftest", fromlist=["HAVE_GPU"]).HAVE_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback(g):
    """Without a CUDA device the product fails loudly instead of computing anything on the CPU."""
    with pytest.raises(g.GpugiError) as e:
        g.Device(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(g.GpugiError):
        g.RendererSystem(0)


def test_script_syntax_and_global_config(g, capfd):
    """ScriptProcessing::ParseCommand (scriptprocessing.cpp:66-158) + GlobalConfig (globalconfig.cpp): comments, blanks, value lists,
    getters, value-count enforcement, unknown names, the two script-only wait commands."""
    from gpugi_b200 import _capi as c
    for name in (b"t_vec", b"t_event", b"t_str"):
        c.host.gih_config_remove(name)
    assert c.host.gih_config_add(b"t_vec", b"1,2,3") == 0
    assert c.host.gih_config_add(b"t_vec", b"1") == 1                      # already exists (globalconfig.cpp:20-27)
    assert c.host.gih_config_add(b"t_event", b"") == 0 and c.host.gih_config_add(b"t_str", b"x") == 0
    assert g.config_get("t_vec") == ["1", "2", "3"] and g.config_get("t_event") == [] and g.config_get("nope") is None
    assert not g.config_parse("t_vec = 0.0,\t1.5 ,  2.27   // camera position")
    assert g.config_get("t_vec") == ["0.0", "1.5", "2.27"]
    assert not g.config_parse("   # a comment line") and not g.config_parse("") and not g.config_parse("//t_vec = 9,9,9")
    assert g.config_get("t_vec") == ["0.0", "1.5", "2.27"]
    assert not g.config_parse("t_str = ../scenes/box1/box1.json")
    assert g.config_get("t_str") == ["../scenes/box1/box1.json"]
    assert not g.config_parse("t_vec = 1, 2")                             # wrong value count: error, value kept
    assert g.config_get("t_vec") == ["0.0", "1.5", "2.27"]
    assert not g.config_parse("doesNotExist = 3")
    err = capfd.readouterr().err
    assert "expects 3 value(s)" in err and "does not exist" in err
    assert not g.config_parse("t_vec")                                    # getter
    assert "{0.0, 1.5, 2.27 }" in capfd.readouterr().out
    assert not g.config_parse("t_event")                                  # an event: empty value list, fires listeners
    assert g.config_parse("waitIterations = 50") and g.config_parse("waitSeconds = 0.5")  # script-only commands stop the processor
    assert g.config_get("pathLength") == ["16"]
    for name in (b"t_vec", b"t_event", b"t_str"):
        c.host.gih_config_remove(name)
