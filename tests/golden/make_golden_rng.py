"""Golden vectors FROM THE REFERENCE ITSELF (run once, in the build container: python tests/golden/make_golden_rng.py).

oracle/_ref/libref_random.so is the reference's own gpugi/utilities/random.cpp, compiled where it lies by oracle/Makefile (the only file of
the reference that g++ accepts unmodified).  This script runs it on fixed inputs and stores inputs + outputs in
ref_rng.npz, which travels with the repository: the oracle (or_*) and the host layer (gih_*) are checked against it on every box.
  wang_in / wang_out          WangHash                      (utilities/random.cpp:5-13; == random.glsl:30-36, the per-thread seed hash)
  xs32_in / xs32_out          Xorshift(uint32)              (:15-22; == random.glsl:11-19 RandomUInt)
  xs64_in / xs64_out          Xorshift(uint64) xorshift64*  (:25-31)
  f_seed / f_out / f_state    XorshiftF stream of 4096 draws from seed 123 (lightsampler.cpp:8), values + final state
  d_out                       XorshiftD stream of 256 draws from seed 123
  frame_seeds                 WangHash(i), i < 1024: PerIteration.FrameSeed of the first 1024 iterations (renderersystem.cpp:101)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_ref():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_random.so"))
    ref.ref_wanghash.restype = C.c_uint32
    ref.ref_wanghash.argtypes = [C.c_uint32]
    ref.ref_xorshift32.restype = C.c_uint32
    ref.ref_xorshift32.argtypes = [C.c_uint32]
    ref.ref_xorshift64.restype = C.c_uint64
    ref.ref_xorshift64.argtypes = [C.c_uint64]
    ref.ref_xorshift_f.restype = C.c_float
    ref.ref_xorshift_f.argtypes = [C.POINTER(C.c_uint64)]
    ref.ref_xorshift_d.restype = C.c_double
    ref.ref_xorshift_d.argtypes = [C.POINTER(C.c_uint64)]
    return ref


def main():
    ref = load_ref()
    rng = np.random.RandomState(20261017)
    edge32 = np.array([0, 1, 2, 61, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF, 123, 123456789], dtype=np.uint32)
    wang_in = np.concatenate([edge32, rng.randint(0, 2**32, 2038, dtype=np.uint64).astype(np.uint32)])
    xs64_in = np.concatenate([np.array([1, 123, 2**63, 2**64 - 1], dtype=np.uint64), rng.randint(0, 2**63, 1020, dtype=np.uint64) * np.uint64(2) + np.uint64(1)])
    out = {
        "wang_in": wang_in, "wang_out": np.array([ref.ref_wanghash(int(x)) for x in wang_in], dtype=np.uint32),
        "xs32_in": wang_in, "xs32_out": np.array([ref.ref_xorshift32(int(x)) for x in wang_in], dtype=np.uint32),
        "xs64_in": xs64_in, "xs64_out": np.array([ref.ref_xorshift64(int(x)) for x in xs64_in], dtype=np.uint64),
        "frame_seeds": np.array([ref.ref_wanghash(i) for i in range(1024)], dtype=np.uint32),
    }
    st = C.c_uint64(123)
    out["f_seed"] = np.uint64(123)
    out["f_out"] = np.array([ref.ref_xorshift_f(C.byref(st)) for _ in range(4096)], dtype=np.float32)
    out["f_state"] = np.uint64(st.value)
    st = C.c_uint64(123)
    out["d_out"] = np.array([ref.ref_xorshift_d(C.byref(st)) for _ in range(256)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_rng.npz"), **out)
    print("wrote ref_rng.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
