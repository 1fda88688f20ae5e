#!/usr/bin/env python
"""Texture-file fixtures for host/imageio.cpp: small images written with Pillow in every format the host decodes, and what the
REFERENCE'S decoder -- stb_image.h as vendored in dependencies/stb, compiled where it lies into oracle/_ref/libref_stb.so -- makes of them
(stbi_load(file, &w, &h, &n, 4), the call of glhelper's Texture2D::LoadFromFile).  Run in the build container; outputs are committed:
tests/golden/images/*.{png,bmp,tga,jpg,ppm} + tests/golden/images/expected.npz."""
import os
import sys
import warnings

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyref  # noqa: E402

OUT = os.path.join(HERE, "images")


def smooth(w, h, c, seed):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(np.sin(x / 7.0 + k) + np.cos(y / 5.0 - k)) * 60 + 128 + rng.randn(h, w) * 6 for k in range(c)], -1)
    return np.clip(img, 0, 255).astype(np.uint8)


def write_adam7_png(path, img):
    import struct
    import zlib
    h, w, c = img.shape
    raw = b""
    for x0, y0, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
        sub = img[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        for row in sub:
            raw += b"\0" + row.tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    ihdr = struct.pack(">IIBBBBB", w, h, 8, {3: 2, 4: 6}[c], 0, 0, 1)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


def main():
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    rgb, rgba, gray = smooth(37, 25, 3, 1), smooth(23, 17, 4, 2), smooth(24, 18, 1, 3)[..., 0]
    files = []

    def save(name, img, **kw):
        img.save(os.path.join(OUT, name), **kw)
        files.append(name)
    save("rgb8.png", Image.fromarray(rgb))
    save("rgba8.png", Image.fromarray(rgba))
    save("gray8.png", Image.fromarray(gray))
    save("pal4.png", Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=16), bits=4)
    save("bw1.png", Image.fromarray((gray > 128).astype(np.uint8) * 255).convert("1"))
    save("gray16.png", Image.fromarray(gray.astype(np.uint16) * 257 + 13))
    # Pillow cannot write interlaced files: a minimal Adam7 encoder (7 reduced images, filter 0, one zlib stream)
    write_adam7_png(os.path.join(OUT, "rgba8_adam7.png"), rgba)
    files.append("rgba8_adam7.png")
    pal = Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=8)
    pal.info["transparency"] = 3
    save("pal_trns.png", pal, transparency=3)
    save("rgb24.bmp", Image.fromarray(rgb))
    save("rgba32.bmp", Image.fromarray(rgba))
    save("pal8.bmp", Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=200))
    save("rgb24.tga", Image.fromarray(rgb))
    save("rgba32_rle.tga", Image.fromarray(rgba), compression="tga_rle")
    save("gray8.tga", Image.fromarray(gray))
    save("q90_444.jpg", Image.fromarray(rgb), quality=90, subsampling=0)
    save("q75_420.jpg", Image.fromarray(rgb), quality=75, subsampling=2)
    save("q60_422.jpg", Image.fromarray(rgb), quality=60, subsampling=1)
    save("gray.jpg", Image.fromarray(gray), quality=85)
    save("restart.jpg", Image.fromarray(smooth(72, 40, 3, 4)), quality=80, subsampling=2, restart_marker_blocks=2)
    save("rgb.ppm", Image.fromarray(rgb))
    expected = {}
    for f in files:
        r = pyref.stbi_load(os.path.join(OUT, f))
        assert r is not None, f
        expected[f] = r[0]
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
    print(len(files), "files,", sum(os.path.getsize(os.path.join(OUT, f)) for f in files), "bytes")


if __name__ == "__main__":
    main()
