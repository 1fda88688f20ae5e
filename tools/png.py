"""Minimal PNG writer (zlib only) for eyeballing HDR renders: tools/png.py is dev tooling, not product."""
import struct
import zlib

import numpy as np


def write_png(path, rgb8):
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def tonemap(img, exposure=None):
    img = np.asarray(img, dtype=np.float32)[..., :3]
    if exposure is None:
        exposure = 1.0 / (4.0 * max(float(img.mean()), 1e-8))
    x = np.clip(img * exposure, 0, 1) ** (1 / 2.2)
    return (x[::-1] * 255).astype(np.uint8)  # image row 0 is the bottom scanline (screenCoord.y = -1)
