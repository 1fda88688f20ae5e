"""Generates the committed golden vectors from the CPU oracle (run once: python tests/golden/make_golden.py).

PARITY UNPINNED by the reference (no golden vectors / images ship with gpugi): these
files pin the ORACLE against drift; the epsilon known-answer tests in test_oracle_kat.py are the only
vectors that come from the reference tree itself.
  box1_rays.npz      8192 rays (primary / secondary / shadow mix + random) on box1(tess=4) with the
                     oracle's {t, leaf_tri, bary} and per-ray visit counts
  box1_pt_64.npy     64x64 PT, 8 iterations (sum image), light sampler seed 123 free-running
  box1_lpt_64.npy    64x64 LPT, 4 iterations      box1_bpt_64.npy  64x64 BPT, 4 iterations (1 thread)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ["OMP_NUM_THREADS"] = "1"  # BPT light-cache order and splat order are then deterministic

import gpugi_b200 as g  # noqa: E402
from gpugi_b200 import scenes  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402
from tests.conftest import camera_ray_batch, make_ray_batch  # noqa: E402


def main():
    sc = scenes.make_box1(tess=4)
    desc = sc.desc
    C = scenes.BOX1_CAMERA
    rays = np.concatenate([camera_ray_batch(g, orc, sc, C["pos"], C["lookat"], C["fov"], 48, 48, 3), make_ray_batch(g, sc, 2048, 9)])[:8192]
    hits, nodes, tris = orc.trace_batch(desc, rays, want_counts=True)
    np.savez_compressed(os.path.join(HERE, "box1_rays.npz"), rays=rays, hits=hits, nodes=nodes, tris=tris)
    W = H = 64
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)

    def sampler():
        return orc.LightSampler(sc.positions(), sc.triangles(), sc.triangle_materials(), sc.emissivity(), sc.point_lights())

    ls = sampler()
    acc = np.zeros((H, W, 4), np.float32)
    for it in range(8):
        acc, _ = orc.render_pt(desc, cam, W, H, ls.generate(128), orc.wanghash(it), 16, accum=acc)
    np.save(os.path.join(HERE, "box1_pt_64.npy"), acc)
    ls = sampler()
    acc = np.zeros((H, W, 4), np.float32)
    nr = max(64, (W * H // 256 // 64) * 64)
    for it in range(4):
        acc, _ = orc.render_lpt(desc, cam, W, H, ls.generate(256), orc.wanghash(it), nr, 16, accum=acc)
    np.save(os.path.join(HERE, "box1_lpt_64.npy"), acc)
    ls = sampler()
    acc = np.zeros((H, W, 4), np.float32)
    nr = max(64, (W * H // 128 // 64) * 64)
    for it in range(4):
        acc, _, _ = orc.render_bpt(desc, cam, W, H, ls.generate(128), orc.wanghash(it), nr, 40000, 16, accum=acc)
    np.save(os.path.join(HERE, "box1_bpt_64.npy"), acc)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
