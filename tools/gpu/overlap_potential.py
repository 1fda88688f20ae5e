#!/usr/bin/env python
"""Experiment: how much does the C2 batch gain when the latency-bound shade kernel of one chain of iterations runs next to the issue-bound trace
kernel of another?  Two GiContexts on one GPU, each with its own stream, each renders half of the iterations from its own host thread;
grids sized so that blocks of both fit an SM together.  Compared with one context rendering all of them."""
import json, os, sys, tempfile, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import gpugi_b200 as g
from gpugi_b200 import _capi

class A: texture_size = 1024; village_triangles = 4000000
cfgname = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = bench.configs(A)[cfgname]
sc, _ = bench.build_scene(cfg)
W, H, C = cfg["W"], cfg["H"], cfg["cam"]
N = 48
def make():
    dev = g.Device(0)
    st = torch.cuda.Stream()
    dev.set_stream(st.cuda_stream)
    dev.upload_scene(sc); dev.set_screen_size(W, H); dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H))
    return dev, st
ls = g.LightSampler(sc); ls.seed(4242)
samples = [ls.generate(128) for _ in range(N)]
d_all = torch.from_numpy(np.concatenate(samples).view(np.uint8).reshape(-1)).cuda()
seeds = [g.wang_hash(k) for k in range(N)]
devs = [make(), make()]
def run(dev, lo, hi):
    dev.render_iterations(g.GI_RENDERER_PT, seeds[lo:hi], 16, d_all.data_ptr() + lo * 128 * 32, 128)
def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for d, _ in devs: run(d, 0, 4)
base = min(timed(lambda: run(devs[0][0], 0, N)) for _ in range(3))
print(json.dumps({"mode": "one context", "ms_per_iteration": base / N}), flush=True)
for tb, sb in [(7, 8), (5, 2), (5, 3), (4, 3), (4, 4), (6, 1), (6, 2), (3, 4)]:
    for d, _ in devs:
        d.set_option(g.GI_OPT_WF_BLOCKS_PER_SM, tb); d.set_option(_capi.GI_OPT_WF_SHADE_BLOCKS_PER_SM, sb)
    def both():
        th = [threading.Thread(target=run, args=(devs[i][0], i * N // 2, (i + 1) * N // 2)) for i in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
    ms = min(timed(both) for _ in range(3))
    one = min(timed(lambda: run(devs[0][0], 0, N)) for _ in range(2))
    print(json.dumps({"mode": "two contexts", "trace_blocks": tb, "shade_blocks": sb, "ms_per_iteration": ms / N, "one_context_same_grids": one / N}), flush=True)
