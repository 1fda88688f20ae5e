#!/usr/bin/env python
"""GPU experiment driver: times PT / LPT / BPT iterations of the Sponza-shaped 1080p workload under both schedulers
(one thread per path = the reference's launch shape, and the wavefront schedule) and, where iterations are independent,
as one overlapped batch.  Prints one JSON line per variant and writes them to --out.
Usage: python tools/gpu/renderer_sweep.py [--iters 10] [--width 1920 --height 1080]"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--texture-size", type=int, default=1024)
    ap.add_argument("--renderers", default="PT,LPT,BPT")
    ap.add_argument("--out", default="gpurun_out/renderer_sweep.json")
    args = ap.parse_args()
    import torch
    import gpugi_b200 as g
    from gpugi_b200 import scenes

    W, H = args.width, args.height
    d = tempfile.mkdtemp(prefix="gpugi_sponza_")
    sc = scenes.make_sponza_like(d, target_triangles=262144, texture_size=args.texture_size, num_materials=25, seed=7)
    C = scenes.SPONZA_CAMERA
    dev = g.Device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    dev.upload_scene(sc)
    dev.set_screen_size(W, H)
    dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
    ls = g.LightSampler(sc)
    ls.seed(4242)
    results = []
    # trace-loop tuning from the environment (tools/gpu/run_variants.sh): idle lanes that trigger a refill, inner-step vote threshold
    if os.environ.get("GPUGI_WF_REFILL"):
        dev.set_option(g.GI_OPT_WF_REFILL_THRESHOLD, int(os.environ["GPUGI_WF_REFILL"]))
    if os.environ.get("GPUGI_WF_MIN_INNER"):
        dev.set_option(g.GI_OPT_WF_MIN_INNER, int(os.environ["GPUGI_WF_MIN_INNER"]))

    def timed(fn):
        dev.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

    for name in args.renderers.split(","):
        rid = {"PT": g.GI_RENDERER_PT, "LPT": g.GI_RENDERER_LPT, "BPT": g.GI_RENDERER_BPT}[name]
        nls = 256 if name == "LPT" else 128
        nr = max(64, (W * H // nls // 64) * 64) if name != "PT" else 0  # lightpathtracer.cpp:34
        samples = [ls.generate(nls) for _ in range(args.iters)]
        d_all = torch.from_numpy(np.concatenate(samples).view(np.uint8).reshape(-1)).cuda()
        seeds = [g.wang_hash(k) for k in range(args.iters)]
        cap = 0
        if name == "BPT":
            dev.set_option(g.GI_OPT_SCHEDULER, g.GI_SCHEDULER_WAVEFRONT)
            dev.set_light_samples(samples[0])
            dev.render_iteration(rid, seeds[0], 16, num_rays_per_light_sample=nr, warmup=1)
            n0 = dev.counters()["light_cache_entries"]
            cap = n0 + n0 // 4
            dev.alloc_light_cache(cap)
        images = {}
        for sched_name, sched in (("megakernel", g.GI_SCHEDULER_MEGAKERNEL), ("wavefront", g.GI_SCHEDULER_WAVEFRONT)):
            dev.set_option(g.GI_OPT_SCHEDULER, sched)
            modes = ["sequential"] + (["batch"] if sched == g.GI_SCHEDULER_WAVEFRONT else [])
            for mode in modes:
                def run():
                    if mode == "batch":
                        dev.render_iterations(rid, seeds, 16, d_all.data_ptr(), nls, num_rays_per_light_sample=nr, light_cache_capacity=cap)
                    else:
                        for k in range(args.iters):
                            dev.set_light_samples_device(d_all.data_ptr() + k * nls * 32, nls)
                            dev.render_iteration(rid, seeds[k], 16, num_rays_per_light_sample=nr, light_cache_capacity=cap)
                dev.clear(); run(); dev.synchronize()  # warm-up (allocations, caches)
                dev.clear(); dev.reset_counters()
                ms = timed(run)
                cnt = dev.counters()
                img = dev.read_accumulator(W, H)
                images[(sched_name, mode)] = img
                paths = cnt["camera_paths"] + cnt["light_paths"]
                rays = cnt["closest_rays"] + cnt["any_rays"]
                r = {"renderer": name, "scheduler": sched_name, "mode": mode, "iterations": args.iters, "ms_per_iteration": ms / args.iters,
                     "mpaths_per_s": paths / ms / 1e3, "mrays_per_s": rays / ms / 1e3, "rays_per_path": rays / max(paths, 1),
                     "mean_radiance": float(img[..., :3].mean() / args.iters), "light_cache_capacity": cap, "num_rays_per_light_sample": nr,
                     "counters": {k: int(v) for k, v in cnt.items()}}
                results.append(r)
                print(json.dumps(r), flush=True)
        ref = images[("megakernel", "sequential")]
        for key, img in images.items():
            diff = float(np.abs(img[..., :3] - ref[..., :3]).sum() / max(np.abs(ref[..., :3]).sum(), 1e-30))
            print("# %s %s/%s: sum|a-b|/sum|b| vs megakernel = %.3e" % (name, key[0], key[1], diff), flush=True)
            results.append({"renderer": name, "variant": "%s/%s" % key, "l1_rel_diff_vs_megakernel": diff})
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    dev.close()


if __name__ == "__main__":
    main()
