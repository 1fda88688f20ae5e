#!/usr/bin/env python
"""Golden vectors written by the REFERENCE'S OWN code (oracle/_ref/libref_bim.so = dependencies/bim + dependencies/epsilon compiled where
they lie, oracle/Makefile).  Run in the build container (needs /root/reference); the outputs are committed:

  ref_box1_t2.json / ref_box1_t2.bim   the box1 mesh (tess 2) through tobim's call sequence: SAH hierarchy, AABoxes, stored by
                                       BinaryModel::storeEnvironmentFile / storeBinaryHeader / storeChunk
  ref_bim_arrays.npz                   what the reference's chunk holds for that mesh after buildHierarchy(SAH) and, separately,
                                       buildHierarchy(KD_TREE): hierarchy, parents, leaves, AABoxes, tree levels (+ the input mesh);
                                       plus a 500-triangle random soup through both builders
  ref_bvhmake_arrays.npz               node pool {left, right}, boxes and 8-triangle leaves of bvhmake's own kdtree and sweep build methods
  ref_ei_kat.npz                       ei::intersects(Ray, Box) / (Ray, Triangle) on 2048 + 2048 seeded random cases: hit flag, distances,
                                       barycentrics
Usage: python tests/golden/make_golden_bim.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from gpugi_b200 import scenes  # noqa: E402  (mesh generators only)
from oracle import pyref  # noqa: E402


def soup(n, seed):
    rng = np.random.RandomState(seed)
    c = rng.uniform(-10, 10, size=(n, 1, 3))
    p = (c + rng.uniform(-1, 1, size=(n, 3, 3))).astype(np.float32).reshape(-1, 3)
    nrm = np.tile(np.array([0, 1, 0], dtype=np.float32), (3 * n, 1))
    uv = rng.rand(3 * n, 2).astype(np.float32)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mat = (np.arange(n) % 2).astype(np.uint32)
    return p, nrm, uv, tri, mat


def ei_cases(seed, n):
    rng = np.random.RandomState(seed)
    o = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    d = rng.randn(n, 3)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return rng, o, d


def main():
    assert pyref.available(), "oracle/_ref/libref_bim.so is missing: run `make -C oracle` where /root/reference exists"
    out = {}
    pos, nrm, uv, tri, mat = scenes.box1_mesh(2)
    env_in = os.path.join(HERE, "_env_in.json")
    json.dump({"scene": "", "accelerator": "aabox", "materials": scenes.box1_materials(), "lights": {}, "cameras": {}}, open(env_in, "w"))
    ref = pyref.RefModel.build(pos, nrm, uv, tri, mat, scenes.BOX1_MATERIAL_NAMES, "sah", env_json=env_in)
    cwd = os.getcwd()
    os.chdir(HERE)  # the environment file stores the bim name as given: keep it relative
    ref.store("ref_box1_t2.json", "ref_box1_t2.bim")
    os.chdir(cwd)
    os.unlink(env_in)
    for tag, mesh, names in [("box1", (pos, nrm, uv, tri, mat), scenes.BOX1_MATERIAL_NAMES), ("soup", soup(500, 5), ["a", "b"])]:
        for k, v in zip(["positions", "normals", "texcoords", "triangles", "triangle_materials"], mesh):
            out["%s_in_%s" % (tag, k)] = v
        for method in ["sah", "kd"]:
            m = pyref.RefModel.build(*mesh, names, method)
            for k, v in m.arrays().items():
                if k in ("hierarchy", "parents", "leaves", "aaboxes", "tree_levels", "bounding_box"):
                    out["%s_%s_%s" % (tag, method, k)] = v
    np.savez_compressed(os.path.join(HERE, "ref_bim_arrays.npz"), **out)

    # bvhmake's own build methods (oracle/_ref/libref_bvhmake.so) on the box1 mesh (tess 4) and a 700-triangle soup
    bv = {}
    for tag, mesh in [("box1", scenes.box1_mesh(4)), ("soup", soup(700, 6))]:
        mpos, _, _, mtri, mmat = mesh
        bv[tag + "_in_positions"], bv[tag + "_in_triangles"], bv[tag + "_in_materials"] = mpos, mtri, mmat
        for method in ["kdtree", "sweep"]:
            lr, boxes, leaves = pyref.bvhmake_build(mpos, mtri, mmat, method)
            bv["%s_%s_children" % (tag, method)], bv["%s_%s_boxes" % (tag, method)], bv["%s_%s_leaves" % (tag, method)] = lr, boxes, leaves
    np.savez_compressed(os.path.join(HERE, "ref_bvhmake_arrays.npz"), **bv)

    # epsilon's routines
    n = 2048
    rng, o, d = ei_cases(77, n)
    lo = rng.uniform(-3, 2, size=(n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0.05, 3, size=(n, 3))).astype(np.float32)
    o[: n // 8] = (lo[: n // 8] + (hi[: n // 8] - lo[: n // 8]) * rng.rand(n // 8, 3)).astype(np.float32)  # origins inside the box
    box = np.zeros(n, dtype=[("hit", "u1"), ("t0", "<f4"), ("t1", "<f4"), ("hit_distance_overload", "u1")])
    for i in range(n):
        box[i] = pyref.ei_ray_box(o[i], d[i], lo[i], hi[i])
    rng2, o2, d2 = ei_cases(78, n)
    c = o2 + d2 * rng2.uniform(0.5, 6, size=(n, 1)).astype(np.float32)
    p = (c[:, None, :] + rng2.uniform(-1.5, 1.5, size=(n, 3, 3))).astype(np.float32)
    tr = np.zeros(n, dtype=[("hit", "u1"), ("t", "<f4"), ("bary", "<f4", 3)])
    for i in range(n):
        tr[i] = pyref.ei_ray_triangle(o2[i], d2[i], p[i, 0], p[i, 1], p[i, 2])
    # Camera::ComputeCameraParams and the light-sample point on epsilon's vector arithmetic
    rng3 = np.random.RandomState(79)
    m = 512
    cam_in = np.concatenate([rng3.uniform(-100, 100, size=(m, 6)), rng3.uniform(20, 100, size=(m, 1)), rng3.choice([1.0, 16.0 / 9.0, 4.0 / 3.0], size=(m, 1))], 1).astype(np.float32)
    cam_out = np.zeros((m, 9), dtype=np.float32)
    for i in range(m):
        cam_out[i] = np.concatenate(pyref.ei_camera_params(cam_in[i, :3], cam_in[i, 3:6], (0, 1, 0), cam_in[i, 6], cam_in[i, 7]))
    ls_tri = rng3.uniform(-5, 5, size=(m, 3, 3)).astype(np.float32)
    xi = rng3.rand(m, 2).astype(np.float32)
    ls_out = np.zeros((m, 6), dtype=np.float32)
    for i in range(m):
        x0 = np.sqrt(xi[i, 0])
        a, b, c2 = np.float32(1.0) - x0, x0 * (np.float32(1.0) - xi[i, 1]), x0 * xi[i, 1]
        ls_out[i] = np.concatenate(pyref.ei_light_sample_point(ls_tri[i, 0], ls_tri[i, 1], ls_tri[i, 2], a, b, c2, 1.0e-4))
    np.savez_compressed(os.path.join(HERE, "ref_ei_kat.npz"), cam_in=cam_in, cam_out=cam_out, ls_tri=ls_tri, ls_xi=xi, ls_out=ls_out, box_o=o, box_d=d, box_lo=lo, box_hi=hi, box_out=box, tri_o=o2, tri_d=d2, tri_p=p, tri_out=tr)
    print("box hits %d / %d, triangle hits %d / %d" % (box["hit"].sum(), n, tr["hit"].sum(), n))
    for f in ["ref_box1_t2.json", "ref_box1_t2.bim", "ref_bim_arrays.npz", "ref_ei_kat.npz"]:
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
