#!/usr/bin/env python
"""Profiling driver: ONE gi_render_iterations batch of a BASELINE config (the schedule bench.py times), short enough to sit under ncu.
    ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 20 -c 1 -o gpurun_out/x python tools/gpu/prof_batch.py --config C2 --iters 24
In a batch of N iterations, wave w >= 16 carries all 16 bounces (steady state); per wave the launches are generate, trace, shade.
Prints the event-timed ms per iteration and the kernel-class times."""
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
    ap.add_argument("--config", default="C2")
    ap.add_argument("--iters", type=int, default=24)
    ap.add_argument("--texture-size", type=int, default=1024)
    ap.add_argument("--village-triangles", type=int, default=4000000)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--kernel-timing", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    import gpugi_b200 as g
    cfg = bench.configs(args)[args.config]
    sc, _ = bench.build_scene(cfg)
    W, H, name = cfg["W"], cfg["H"], cfg["renderer"]
    C = cfg["cam"]
    dev = g.Device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    if os.environ.get("GPUGI_WF_REFILL"):
        dev.set_option(g.GI_OPT_WF_REFILL_THRESHOLD, int(os.environ["GPUGI_WF_REFILL"]))
    if os.environ.get("GPUGI_WF_MIN_INNER"):
        dev.set_option(g.GI_OPT_WF_MIN_INNER, int(os.environ["GPUGI_WF_MIN_INNER"]))
    dev.upload_scene(sc)
    dev.set_screen_size(W, H)
    dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
    rid = g.RENDERERS[name]
    nls = cfg["nls"]
    ppi, nr = bench.paths_per_iteration(cfg)
    ls = g.LightSampler(sc)
    ls.seed(4242)
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
    extra = dict(num_rays_per_light_sample=nr if name != "PT" else 0, light_cache_capacity=cap)
    out = []
    for rep in range(args.repeat):
        dev.clear(); dev.reset_counters()
        if args.kernel_timing:
            dev.set_option(g.GI_OPT_KERNEL_TIMING, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev.synchronize()
        e0.record(stream)
        dev.render_iterations(rid, seeds, 16, d_all.data_ptr(), nls, **extra)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        r = {"config": args.config, "iters": args.iters, "ms_per_iteration": ms / args.iters, "mpaths_per_s": ppi * args.iters / ms / 1e3}
        if args.kernel_timing:
            kt = dev.kernel_times()
            r["kernel_ms_per_iteration"] = {k: v[0] / args.iters for k, v in kt.items() if v[1]}
            dev.set_option(g.GI_OPT_KERNEL_TIMING, 0)
        cn = dev.counters()
        r["shaded_per_closest_ray"] = cn["shaded_vertices"] / max(cn["closest_rays"], 1)
        r["splats_per_any_ray"] = cn["splats"] / max(cn["any_rays"], 1)
        out.append(r)
        print(json.dumps(r), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
