#!/usr/bin/env python
"""GPU driver for the other BASELINE.json configs on their stand-in scenes (gpugi_b200/scenes.py):
  C3 bathroom-like  BidirectionalPathtracer 1920x1080      C4 villa-like  LightPathtracer 3840x2160      C5 village-like  Pathtracer 3840x2160
Each is rendered as one gi_render_iterations batch (the schedule the bench uses for C2) and, for comparison, one iteration per call.
Prints one JSON line per measurement and writes them to --out.   Usage: python tools/gpu/config_sweep.py [--iters 8] [--configs C3,C4,C5]"""
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
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--configs", default="C3,C4,C5")
    ap.add_argument("--village-triangles", type=int, default=4000000)
    ap.add_argument("--out", default="gpurun_out/config_sweep.json")
    ap.add_argument("--topologies", default="ploc", help="comma list of ploc,reference (GI_OPT_UPLOAD_TOPOLOGY): every config is measured on each")
    ap.add_argument("--modes", default="sequential,batch")
    ap.add_argument("--radii", default="4", help="comma list of GI_OPT_PLOC_RADIUS values tried for the ploc topology")
    args = ap.parse_args()
    import torch
    import gpugi_b200 as g
    from gpugi_b200 import scenes

    tmp = tempfile.mkdtemp(prefix="gpugi_cfg_")
    configs = {
        "C2": ("sponza-shaped Pathtracer 1920x1080", lambda: scenes.make_sponza_like(tmp, 262144, 1024), scenes.SPONZA_CAMERA, 1920, 1080, "PT"),
        "C3": ("bathroom-like BidirectionalPathtracer 1920x1080", lambda: scenes.make_bathroom_like(tmp, 500000, 512), scenes.BATHROOM_CAMERA, 1920, 1080, "BPT"),
        "C4": ("villa-like LightPathtracer 3840x2160", lambda: scenes.make_villa_like(tmp, 1000000, 512), scenes.VILLA_CAMERA, 3840, 2160, "LPT"),
        "C5": ("village-like Pathtracer 3840x2160", lambda: scenes.make_village_like(args.village_triangles), scenes.VILLAGE_CAMERA, 3840, 2160, "PT"),
    }
    results = []
    for key in args.configs.split(","):
        label, build, C, W, H, name = configs[key]
        t0 = time.time()
        sc = build()
        build_s = time.time() - t0
        for topo, radius in [(t, int(r)) for t in args.topologies.split(",") for r in (args.radii.split(",") if t == "ploc" else ["4"])]:
            dev = g.Device(0)
            stream = torch.cuda.Stream()
            torch.cuda.set_stream(stream)
            dev.set_stream(stream.cuda_stream)
            dev.set_option(g.GI_OPT_UPLOAD_TOPOLOGY, g.GI_TOPOLOGY_PLOC if topo == "ploc" else g.GI_TOPOLOGY_REFERENCE)
            dev.set_option(g.GI_OPT_PLOC_RADIUS, radius)
            t1 = time.time()
            dev.upload_scene(sc)
            upload_s = time.time() - t1
            active, topo_ms = dev.scene_topology()
            dev.set_screen_size(W, H)
            dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
            rid = {"PT": g.GI_RENDERER_PT, "LPT": g.GI_RENDERER_LPT, "BPT": g.GI_RENDERER_BPT}[name]
            nls = 256 if name == "LPT" else 128
            nr = max(64, (W * H // nls // 64) * 64) if name != "PT" else 0
            ls = g.LightSampler(sc)
            samples = [ls.generate(nls) for _ in range(args.iters)]
            d_all = torch.from_numpy(np.concatenate(samples).view(np.uint8).reshape(-1)).cuda()
            seeds = [g.wang_hash(k) for k in range(args.iters)]
            cap = 0
            if name == "BPT":
                dev.set_light_samples(samples[0])
                dev.render_iteration(rid, seeds[0], 16, num_rays_per_light_sample=nr, warmup=1)
                n0 = dev.counters()["light_cache_entries"]
                cap = n0 + n0 // 4
                dev.alloc_light_cache(cap)
            for mode in args.modes.split(","):
                def run():
                    if mode == "batch":
                        dev.render_iterations(rid, seeds, 16, d_all.data_ptr(), nls, num_rays_per_light_sample=nr, light_cache_capacity=cap)
                    else:
                        for k in range(args.iters):
                            dev.set_light_samples_device(d_all.data_ptr() + k * nls * 32, nls)
                            dev.render_iteration(rid, seeds[k], 16, num_rays_per_light_sample=nr, light_cache_capacity=cap)
                dev.clear(); run(); dev.synchronize()
                dev.clear(); dev.reset_counters()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dev.synchronize()
                e0.record(stream); run(); e1.record(stream); e1.synchronize()
                ms = e0.elapsed_time(e1)
                cnt = dev.counters()
                img = dev.read_accumulator(W, H)
                paths = cnt["camera_paths"] + cnt["light_paths"]
                rays = cnt["closest_rays"] + cnt["any_rays"]
                r = {"config": key, "workload": label, "mode": mode, "topology": topo, "ploc_radius": radius, "topology_active": active, "topology_build_ms": topo_ms, "upload_s": upload_s, "iterations": args.iters, "triangles": int(sc.info.num_leaf_triangles), "scene_build_s": build_s,
                     "ms_per_iteration": ms / args.iters, "mpaths_per_s": paths / ms / 1e3, "mrays_per_s": rays / ms / 1e3, "rays_per_path": rays / max(paths, 1),
                     "mean_radiance": float(img[..., :3].mean() / args.iters), "finite": bool(np.isfinite(img).all()), "light_cache_capacity": cap,
                     "num_rays_per_light_sample": nr, "free_gb_after": torch.cuda.mem_get_info()[0] / 2**30}
                results.append(r)
                print(json.dumps(r), flush=True)
            dev.close()
            del dev
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
