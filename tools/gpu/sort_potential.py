#!/usr/bin/env python
"""How much would ray sorting buy?  Secondary (cosine-hemisphere) rays from the first hits of the bench frame, traced with gi_trace_device
(a) in the order the shading wave produces them (pixel tiles) and (b) sorted by (origin cell Morton code, direction octant); sorting cost not
included.  Usage: python tools/gpu/sort_potential.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import gpugi_b200 as g
    from gpugi_b200 import scenes
    W, H = 1920, 1080
    sc = scenes.make_sponza_like(tempfile.mkdtemp(), texture_size=64)
    C = scenes.SPONZA_CAMERA
    cam = g.camera_params(C["pos"], C["lookat"], C["fov"], W, H)
    dev = g.Device(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); dev.set_stream(stream.cuda_stream)
    dev.upload_scene(sc)
    # primary rays in 8x8 tile order
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    tiles = ((ys // 8) * (W // 8) + (xs // 8)) * 64 + (ys % 8) * 8 + (xs % 8)
    order = np.argsort(tiles.ravel())
    px, py = xs.ravel()[order], ys.ravel()[order]
    rng = np.random.RandomState(1)
    sx = (px + rng.rand(px.size)) / W * 2 - 1
    sy = (py + rng.rand(px.size)) / H * 2 - 1
    U, V, Wv = np.array(cam.u[:]), np.array(cam.v[:]), np.array(cam.w[:])
    d = sx[:, None] * U + sy[:, None] * V + Wv
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(px.size, dtype=g.RAY_DTYPE)
    rays["origin"] = np.array(cam.position[:], dtype=np.float32); rays["direction"] = d.astype(np.float32); rays["t_max"] = g.RAY_MAX
    hits = dev.trace(rays)
    ok = hits["leaf_tri"] >= 0
    lt, P = sc.leaf_triangles(), sc.positions()
    tri = lt["v"][hits["leaf_tri"][ok]]
    p0, p1, p2 = P[tri[:, 0]], P[tri[:, 1]], P[tri[:, 2]]
    n = np.cross(p1 - p0, p2 - p0); n /= np.linalg.norm(n, axis=1, keepdims=True) + 1e-30
    din = rays["direction"][ok]
    n = np.where((np.einsum("ij,ij->i", n, din) > 0)[:, None], -n, n)
    o = rays["origin"][ok] + hits["t"][ok][:, None] * din + 1e-3 * n
    r1, r2 = rng.rand(o.shape[0]), rng.rand(o.shape[0])
    a = np.where(np.abs(n[:, 0:1]) > 0.9, np.array([[0, 1.0, 0]]), np.array([[1.0, 0, 0]]))
    t1 = np.cross(n, a); t1 /= np.linalg.norm(t1, axis=1, keepdims=True); t2 = np.cross(n, t1)
    phi, ct, st = 2 * np.pi * r1, np.sqrt(1 - r2), np.sqrt(r2)
    d2 = (st * np.cos(phi))[:, None] * t1 + (st * np.sin(phi))[:, None] * t2 + ct[:, None] * n
    sec = np.zeros(o.shape[0], dtype=g.RAY_DTYPE)
    sec["origin"] = o.astype(np.float32); sec["direction"] = d2.astype(np.float32); sec["t_max"] = g.RAY_MAX
    lo, hi = np.array(sc.info.bbox_min), np.array(sc.info.bbox_max)

    def morton(o, bits):
        q = np.clip(((o - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
        code = np.zeros(o.shape[0], np.int64)
        for b in range(bits):
            for k in range(3):
                code |= ((q[:, k] >> b) & 1) << (3 * b + k)
        return code
    octant = (sec["direction"][:, 0] < 0) * 1 + (sec["direction"][:, 1] < 0) * 2 + (sec["direction"][:, 2] < 0) * 4
    orders = {"shading order (pixel tiles)": np.arange(sec.shape[0]),
              "sorted: octant, then origin cell (8 bits/axis)": np.lexsort((morton(sec["origin"], 8), octant)),
              "sorted: origin cell (6 bits/axis), then octant": np.lexsort((octant, morton(sec["origin"], 6))),
              "random permutation": rng.permutation(sec.shape[0])}
    d_hits = torch.empty(sec.shape[0] * 16, dtype=torch.uint8, device="cuda")
    for name, perm in orders.items():
        batch = np.ascontiguousarray(sec[perm])
        d_rays = torch.from_numpy(batch.view(np.uint8).reshape(-1)).cuda()
        for _ in range(3):
            dev.trace_device(d_rays.data_ptr(), batch.shape[0], d_hits.data_ptr())
        dev.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            dev.trace_device(d_rays.data_ptr(), batch.shape[0], d_hits.data_ptr())
        e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%-50s %8d rays  %.3f ms  %.0f Mrays/s" % (name, batch.shape[0], ms, batch.shape[0] / ms / 1e3), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
