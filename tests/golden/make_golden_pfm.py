"""Golden PFM FROM THE REFERENCE ITSELF (run once, in the build container: python tests/golden/make_golden_pfm.py).

oracle/_ref/libref_hdrimage.so is the reference's own gpugi/imageut/hdrimage.cpp compiled where it lies (oracle/Makefile; stand-in
<ei/vector.hpp> with epsilon's layout).  Its WritePfm turns a fixed 7x5 RGBA accumulator (sum image, divisor 3 = iteration count) into
ref_pfm_7x5_div3.pfm; the host layer's WritePfm must produce the same bytes and its LoadPfm the same pixels (tests/test_host.py)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_ref():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_hdrimage.so"))
    ref.ref_write_pfm.restype = C.c_int
    ref.ref_write_pfm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
    ref.ref_load_pfm.restype = C.c_int
    ref.ref_load_pfm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return ref


def golden_input():
    rng = np.random.RandomState(31)
    img = (rng.rand(5, 7, 4) * 100).astype(np.float32)
    img[0, 0] = [0.0, 1.0, 3.0, 7.0]
    img[4, 6] = [1e-30, 1e30, 0.1, 0.0]
    return img


def main():
    ref = load_ref()
    img = golden_input()
    path = os.path.join(HERE, "ref_pfm_7x5_div3.pfm")
    assert ref.ref_write_pfm(os.fsencode(path), img.ctypes.data, 7, 5, 3) == 0
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
