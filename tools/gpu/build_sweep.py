#!/usr/bin/env python
"""Host SAH sweep vs device LBVH (gi_build_hierarchy): build time and path-tracer speed on the resulting trees.
Usage: python tools/gpu/build_sweep.py [--out gpurun_out/build_sweep.json]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=24)
    ap.add_argument("--out", default="gpurun_out/build_sweep.json")
    args = ap.parse_args()
    import torch
    import gpugi_b200 as g
    from gpugi_b200 import scenes
    tmp = tempfile.mkdtemp(prefix="gpugi_build_")
    cases = {
        "sponza-like 262k": (lambda m: scenes.make_sponza_like(tmp, texture_size=256, build_method=m), scenes.SPONZA_CAMERA, 1920, 1080),
        "village-like 2.4M": (lambda m: scenes.make_village_like(4000000, build_method=m), scenes.VILLAGE_CAMERA, 1920, 1080),
    }
    results = []
    dev0 = g.Device(0)  # context creation out of the timings
    for label, (mk, C, W, H) in cases.items():
        for method in ("sah", "lbvh", "ploc"):
            t0 = time.time()
            sc = mk(method)
            total_s = time.time() - t0
            # the build alone, on the final mesh
            pos, tri, mat = sc.positions(), sc.triangles(), sc.triangle_materials()
            build_ms = None
            if method != "sah":
                dev0.set_option(g.GI_OPT_DEVICE_BUILDER, g.GI_BUILDER_LBVH if method == "lbvh" else g.GI_BUILDER_PLOC)
                dev0.build_hierarchy(pos, tri, mat)
                t1 = time.time(); dev0.build_hierarchy(pos, tri, mat); build_ms = (time.time() - t1) * 1e3
            dev = g.Device(0)
            stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); dev.set_stream(stream.cuda_stream)
            dev.upload_scene(sc); dev.set_screen_size(W, H)
            dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
            ls = g.LightSampler(sc)
            samples = np.concatenate([ls.generate(128) for _ in range(args.iters)])
            d_all = torch.from_numpy(samples.view(np.uint8).reshape(-1)).cuda()
            seeds = [g.wang_hash(k) for k in range(args.iters)]
            run = lambda: dev.render_iterations(g.GI_RENDERER_PT, seeds, 16, d_all.data_ptr(), 128)
            dev.clear(); run(); dev.synchronize(); dev.clear(); dev.reset_counters()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); run(); e1.record(stream); e1.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            cnt = dev.counters()
            r = {"scene": label, "build": method, "triangles": int(tri.shape[0]), "nodes": int(sc.info.num_nodes), "tree_levels": int(sc.info.num_tree_levels),
                 "scene_from_mesh_s": total_s, "device_build_ms_incl_h2d_d2h": build_ms, "pt_ms_per_iteration": ms,
                 "mrays_per_s": (cnt["closest_rays"] + cnt["any_rays"]) / (ms * args.iters) / 1e3, "mean_radiance": float(dev.read_accumulator(W, H)[..., :3].mean() / args.iters)}
            results.append(r)
            print(json.dumps(r), flush=True)
            dev.close()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
