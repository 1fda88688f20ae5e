#!/usr/bin/env python
"""Small tour of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu/sanitize.py
Tiny workloads (box1 tess 4, 64x48) so that the instrumented run takes seconds: traversal (ordered + reference-order, both upload
topologies), PT / LPT / BPT / Whitted single iterations in both schedulers, overlapped batches, streaming Draw with asynchronous
read-out, device light sampling, environment cube, MSE, device hierarchy builders."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import gpugi_b200 as g
    from gpugi_b200 import scenes

    W, H = 64, 48
    sc = scenes.make_box1(tess=4)
    C = scenes.BOX1_CAMERA
    info = sc.info
    lo, hi = np.array(info.bbox_min), np.array(info.bbox_max)
    rng = np.random.RandomState(3)
    rays = np.zeros(4096, dtype=g.RAY_DTYPE)
    rays["origin"] = (lo + rng.rand(len(rays), 3) * (hi - lo)).astype(np.float32)
    d = rng.randn(len(rays), 3)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    rays["flags"][:1024] = 1
    rays["t_max"][:1024] = rng.rand(1024).astype(np.float32) * 4
    ref_hits = None
    for topo in (g.GI_TOPOLOGY_PLOC, g.GI_TOPOLOGY_REFERENCE):
        dev = g.Device(0)
        dev.set_option(g.GI_OPT_UPLOAD_TOPOLOGY, topo)
        dev.upload_scene(sc)
        for mode in (g.GI_TRAVERSAL_FAST, g.GI_TRAVERSAL_REFERENCE):
            dev.set_option(g.GI_OPT_TRAVERSAL, mode)
            hits = dev.trace(rays)
            if ref_hits is None:
                ref_hits = hits.copy()
            assert np.array_equal(hits["leaf_tri"], ref_hits["leaf_tri"]) and np.array_equal(hits["t"].view(np.uint32), ref_hits["t"].view(np.uint32))
        dev.set_option(g.GI_OPT_TRAVERSAL, g.GI_TRAVERSAL_FAST)
        dev.set_screen_size(W, H)
        dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
        ls = g.LightSampler(sc)
        for sched in (g.GI_SCHEDULER_WAVEFRONT, g.GI_SCHEDULER_MEGAKERNEL):
            dev.set_option(g.GI_OPT_SCHEDULER, sched)
            dev.set_light_samples(ls.generate(128))
            dev.render_iteration(g.GI_RENDERER_PT, g.wang_hash(0), 16)
            dev.render_iteration(g.GI_RENDERER_WHITTED, g.wang_hash(1), 16)
            dev.set_light_samples(ls.generate(256))
            dev.render_iteration(g.GI_RENDERER_LPT, g.wang_hash(2), 16, num_rays_per_light_sample=64)
            dev.show_light_caches()
            dev.set_light_samples(ls.generate(128))
            dev.render_iteration(g.GI_RENDERER_BPT, g.wang_hash(3), 16, num_rays_per_light_sample=64, warmup=1)
            n0 = dev.counters()["light_cache_entries"]
            dev.alloc_light_cache(n0 + n0 // 4 + 16)
            dev.render_iteration(g.GI_RENDERER_BPT, g.wang_hash(3), 16, num_rays_per_light_sample=64, light_cache_capacity=n0 + n0 // 4 + 16)
        img = dev.read_accumulator(W, H)
        assert np.isfinite(img).all()
        dev.compute_mse(np.ascontiguousarray(img[..., :3]), 1)
        dev.close()

    # Renderer API: batches through Draw in streaming mode, asynchronous read-out, device light sampling, environment cube
    for name in ("PT", "BPT", "LPT"):
        rs = g.RendererSystem(0)
        rs.SetScene(sc)
        rs.SetScreenSize(W, H)
        rs.SetCamera(C["pos"], C["lookat"], C["fov"])
        rs.SetRenderer(name)
        if name == "PT":
            faces = [np.full((4, 4, 4), 40 * (i + 1), dtype=np.uint8) for i in range(6)]
            rs.device.set_environment_cube(faces)
            rs.SetDeviceLightSampling(True)
        rs.SetStreaming(True)
        out = np.zeros((H, W, 4), dtype=np.float32)
        for k in range(6):
            rs.Draw()
            if k % 2 == 1:
                rs.ReadBackbufferAsync(out)
        rs.WaitReads()
        img = rs.ReadBackbuffer()
        assert np.isfinite(img).all()
        rs.close()

    # overlapped batch through the C ABI
    import torch
    dev = g.Device(0)
    dev.upload_scene(sc)
    dev.set_screen_size(W, H)
    dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
    ls = g.LightSampler(sc)
    K = 5
    samples = np.concatenate([ls.generate(128) for _ in range(K)])
    d_all = torch.from_numpy(samples.view(np.uint8).reshape(-1)).cuda()
    dev.render_iterations(g.GI_RENDERER_PT, [g.wang_hash(k) for k in range(K)], 16, d_all.data_ptr(), 128)
    dev.render_iterations(g.GI_RENDERER_LPT, [g.wang_hash(k) for k in range(K)], 16, d_all.data_ptr(), 128, num_rays_per_light_sample=64)
    dev.synchronize()
    # device hierarchy builders
    pos, tri, mat = sc.positions(), sc.triangles(), sc.triangle_materials()
    for b in (g.GI_BUILDER_LBVH, g.GI_BUILDER_PLOC):
        dev.set_option(g.GI_OPT_DEVICE_BUILDER, b)
        nodes, leaves, parents, levels = dev.build_hierarchy(pos, tri, mat)
        assert len(nodes) >= 1 and levels >= 1
    dev.close()
    print("sanitize tour ok")


if __name__ == "__main__":
    main()
