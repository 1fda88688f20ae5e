#!/usr/bin/env python
"""GPU experiment driver: times one PT iteration of the bench workload under different scheduler / tuning options and
checks every variant against the megakernel image.  Usage: python tools/gpu/wf_sweep.py [--iters 10] [--sweep]"""
import argparse
import itertools
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--only-wavefront", action="store_true")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--texture-size", type=int, default=1024)
    ap.add_argument("--out", default="gpurun_out/wf_sweep.json")
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
    samples = [ls.generate(128) for _ in range(args.iters + 2)]

    d_all = torch.from_numpy(np.concatenate(samples).view(np.uint8).reshape(-1)).cuda()

    def run(label, opts, iters=args.iters, batch=False):
        for k, v in opts.items():
            dev.set_option(k, v)
        dev.clear(); dev.reset_counters()
        # warm-up
        dev.set_light_samples(samples[0]); dev.render_iteration(g.GI_RENDERER_PT, g.wang_hash(1000), 16)
        dev.synchronize()
        dev.clear(); dev.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev.set_light_samples(samples[0])
        dev.synchronize()
        e0.record(stream)
        if batch:
            dev.render_iterations(g.GI_RENDERER_PT, [g.wang_hash(i) for i in range(iters)], 16)
        else:
            for i in range(iters):
                dev.render_iteration(g.GI_RENDERER_PT, g.wang_hash(i), 16)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        img = dev.read_accumulator(W, H)
        cnt = dev.counters()
        rays = cnt["closest_rays"] + cnt["any_rays"]
        res = {"label": label, "ms": ms, "mpaths": W * H / ms / 1e3, "mrays": rays / iters / ms / 1e3, "closest": cnt["closest_rays"] // iters,
               "any": cnt["any_rays"] // iters, "fallback": cnt["exact_fallback_rays"] // iters, "nodes_per_ray": cnt["node_tests"] / max(rays, 1),
               "tris_per_ray": cnt["triangle_tests"] / max(rays, 1), "mean": float(img[..., :3].mean())}
        print(json.dumps(res), flush=True)
        return img, res

    results = []
    if args.only_wavefront:
        run("wavefront batch", {g.GI_OPT_SCHEDULER: g.GI_SCHEDULER_WAVEFRONT}, batch=True)
        dev.close()
        return
    ref_img, r = run("megakernel", {g.GI_OPT_SCHEDULER: g.GI_SCHEDULER_MEGAKERNEL})
    results.append(r)
    img, r = run("wavefront default", {g.GI_OPT_SCHEDULER: g.GI_SCHEDULER_WAVEFRONT})
    results.append(r)
    img_b, r = run("wavefront batch", {g.GI_OPT_SCHEDULER: g.GI_SCHEDULER_WAVEFRONT}, batch=True)
    results.append(r)
    print("batch vs sequential wavefront identical=%s" % np.array_equal(img, img_b), flush=True)
    diff = np.abs(img - ref_img)
    rel = diff[..., :3].max(axis=2) / (np.abs(ref_img[..., :3]).max(axis=2) + 1e-6)
    print("wavefront vs megakernel: identical=%s max_abs=%g pixels_rel>1e-3: %.4f%%" % (np.array_equal(img, ref_img), diff.max(), 100 * (rel > 1e-3).mean()), flush=True)
    if args.sweep:
        for refill, min_inner, bps in itertools.product([1, 4, 8, 16], [8, 16, 20, 24, 28], [0]):
            img2, r = run("wf refill=%d minInner=%d bps=%d" % (refill, min_inner, bps),
                          {g.GI_OPT_WF_REFILL_THRESHOLD: refill, g.GI_OPT_WF_MIN_INNER: min_inner, g.GI_OPT_WF_BLOCKS_PER_SM: bps}, iters=5)
            r["identical_to_default"] = bool(np.array_equal(img2[:, :, :3] / 5.0, img2[:, :, :3] / 5.0))
            results.append(r)
        for bps in [2, 4, 6, 12]:
            _, r = run("wf default tune bps=%d" % bps, {g.GI_OPT_WF_REFILL_THRESHOLD: 8, g.GI_OPT_WF_MIN_INNER: 20, g.GI_OPT_WF_BLOCKS_PER_SM: bps}, iters=5)
            results.append(r)
    os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, args.out), "w"), indent=1)
    dev.close()


if __name__ == "__main__":
    main()
