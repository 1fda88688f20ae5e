#!/usr/bin/env python
"""bench.py -- headline benchmark of the light-transport hot path (BASELINE.json: "Mpaths/s & Mrays/s, Sponza
1080p PT at 1/2/4/8 B200; relMSE vs ref").

A "step" is one PathTracer sample pass (iteration) over the Sponza-shaped scene at 1920x1080 (BASELINE config 2:
262 144 triangles, 25 textured materials, camera + 4 point lights of scenes/cryteksponza.txt, pathLength 16).
  value  whole-job Mpaths/s with every input resident in HBM before the timed region (scene, textures, the light
         samples of all iterations, accumulator); K iterations timed with CUDA events on the launching stream.
  e2e    the same metric through the Renderer API with HOST buffers: per step RendererSystem.Draw() (host light sampling +
         4 KB H2D + the iteration's kernels, streaming mode: one wave per call) and a D2H read of the full RGBA32F
         progressive image into pinned memory (ReadBackbufferAsync), final draining ReadBackbuffer inside the timed
         region; e2e_sync: plain Draw() + blocking ReadBackbuffer every step.
  --impl reference   the reference's CPU side of this path: the oracle (CPU restatement of the reference shaders,
         kind "port" -- the GLSL/MSVC reference cannot be built here) on all host threads, each step a bounded
         sample (evenly spaced 8-pixel stripes) of the same iteration.
Multi-GPU: one process per GPU (torchrun), iterations sharded round-robin over ranks, no data-path collective
except one NCCL sum-reduce of the accumulators at read-out (inside the timed region); scaling "weak".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
PATH_LENGTH = 16
METRIC = "Mpaths/s, Sponza-shaped 1080p PathTracer (paths started per second; Mrays/s alongside)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line (the JSON result).  Libraries write there too (NCCL prints its version banner on
# stdout when NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the whole run and the result line goes
# to the saved descriptor.
_RESULT_FD = None


def capture_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            sm, smax, reasons = [], [], set()
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ----------------------------------------------------------------------------------------------- scene
def build_scene(texture_size):
    from gpugi_b200 import scenes
    d = tempfile.mkdtemp(prefix="gpugi_sponza_")
    t = time.time()
    sc = scenes.make_sponza_like(d, target_triangles=262144, texture_size=texture_size, num_materials=25, seed=7)
    return sc, time.time() - t


def algorithmic_bytes(c):
    """bytes the REFERENCE layout must touch, from reference-order visit counts.
    per traversal 32*n_nodes + 52*n_tris + 32 (ray in) + 16 (hit out); per shaded vertex 176; per NEE 32;
    per finished camera path 32 (accumulator read + write)."""
    rays = c["closest_rays"] + c["any_rays"]
    return 32 * c["node_tests"] + 52 * c["triangle_tests"] + 48 * rays + 176 * c["shaded_vertices"] + 32 * c["shaded_vertices"] + 32 * c["camera_paths"]


# ----------------------------------------------------------------------------------------------- oracle legs
def stripe_regions(num_stripes):
    ys = [int(HEIGHT * (i + 0.5) / num_stripes) // 8 * 8 for i in range(num_stripes)]
    return [(0, y, WIDTH, y + 8) for y in ys]


def oracle_pass(orc, scene, cam, sampler, iteration, regions):
    """One bounded sample of iteration `iteration` on all host threads; returns (paths, rays, seconds, counters)."""
    samples = sampler.generate(128)
    acc = np.zeros((HEIGHT, WIDTH, 4), dtype=np.float32)
    tot = {}
    t0 = time.perf_counter()
    for r in regions:
        acc, cnt = orc.render_pt(scene.desc, cam, WIDTH, HEIGHT, samples, orc.wanghash(iteration), PATH_LENGTH, region=r, accum=acc)
        for k, v in cnt.items():
            tot[k] = tot.get(k, 0) + v
    dt = time.perf_counter() - t0
    return tot["camera_paths"], tot["closest_rays"] + tot["any_rays"], dt, tot


def cpu_baseline(scene, budget_s=12.0):
    from gpugi_b200 import scenes
    from oracle import pyoracle as orc
    C = scenes.SPONZA_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], WIDTH, HEIGHT)
    sampler = orc.LightSampler(scene.positions(), scene.triangles(), scene.triangle_materials(), scene.emissivity(), scene.point_lights())
    regions = stripe_regions(9)
    paths = rays = 0
    secs = 0.0
    it = 0
    tot = {}
    while secs < budget_s and it < 400:
        p, r, dt, cnt = oracle_pass(orc, scene, cam, sampler, it, regions)
        paths += p; rays += r; secs += dt; it += 1
        for k, v in cnt.items():
            tot[k] = tot.get(k, 0) + v
    return {"value": paths / secs / 1e6, "unit": "Mpaths/s", "mrays_per_s": rays / secs / 1e6, "cores": orc.num_threads(), "kind": "port",
            "sample": "%d iterations x 9 evenly spaced 8-pixel stripes of the 1920x1080 frame (%d paths, %.1f s)" % (it, paths, secs)}, tot


def run_reference(args, rank, world):
    """--impl reference: the oracle port on this box's host cores, same config / metric."""
    if rank != 0:
        return
    from gpugi_b200 import scenes
    from oracle import pyoracle as orc
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 alone runs this arm and is meant to use every host core it may run on
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    scene, build_s = build_scene(args.texture_size)
    C = scenes.SPONZA_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], WIDTH, HEIGHT)
    sampler = orc.LightSampler(scene.positions(), scene.triangles(), scene.triangle_materials(), scene.emissivity(), scene.point_lights())
    # calibrate the stripe count so that one step takes ~1.5 s
    oracle_pass(orc, scene, cam, sampler, 0, stripe_regions(3))  # cold pass (page faults, sRGB table)
    p, r, dt, _ = oracle_pass(orc, scene, cam, sampler, 0, stripe_regions(3))
    stripes = int(max(1, min(HEIGHT // 8, round(3 * 1.5 / max(dt, 1e-3)))))
    regions = stripe_regions(stripes)
    for w in range(args.warmup):
        oracle_pass(orc, scene, cam, sampler, w, regions)
    paths = rays = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        p, r, _, _ = oracle_pass(orc, scene, cam, sampler, args.warmup + k, regions)
        paths += p; rays += r
    secs = time.perf_counter() - t0
    value = paths / secs / 1e6
    sample = "each step = %d evenly spaced 8-pixel stripes (%d paths) of one 1920x1080 PT iteration" % (stripes, paths // max(args.steps, 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mrays_per_s": rays / secs / 1e6,
            "config": {"workload": "sponza-shaped atrium PathTracer 1920x1080 (BASELINE config 2), pathLength 16, 262144 triangles", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": orc.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import gpugi_b200 as g
    from gpugi_b200 import parallel, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gpugi_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = parallel.init_process_group("nccl") if world > 1 else None
    K, Wm = args.steps, args.warmup
    scene, build_s = build_scene(args.texture_size)
    info = scene.info
    C = scenes.SPONZA_CAMERA

    rs = g.RendererSystem(local_rank)
    dev = rs.device
    # a non-default torch stream carries every launch, so torch.cuda.Event times exactly our kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)
    if args.topology == "reference":
        dev.set_option(g.GI_OPT_UPLOAD_TOPOLOGY, g.GI_TOPOLOGY_REFERENCE)
    rs.SetScene(scene)
    topo_active, topo_ms = dev.scene_topology()
    rs.SetScreenSize(WIDTH, HEIGHT)
    rs.SetCamera(C["pos"], C["lookat"], C["fov"])
    rs.SetRenderer("PT")
    rs.SetIterationShard(rank, world)
    rs.SetSynchronousDraw(False)
    paths_per_iter = (WIDTH // 8) * (HEIGHT // 8) * 64
    acc_t = parallel.accumulator_as_tensor(dev, WIDTH, HEIGHT)

    # ---- algorithmic bytes of one iteration: reference-order visit counts from the device counters
    ls_host = g.LightSampler(scene)
    ls_host.seed(4242)
    probe = ls_host.generate(128)
    dev.set_option(g.GI_OPT_TRAVERSAL, g.GI_TRAVERSAL_REFERENCE)
    dev.set_light_samples(probe); dev.reset_counters()
    dev.render_iteration(g.GI_RENDERER_PT, g.wang_hash(rank), PATH_LENGTH)
    ref_counts = dev.counters()
    dev.set_option(g.GI_OPT_TRAVERSAL, g.GI_TRAVERSAL_FAST)
    bytes_per_iter = algorithmic_bytes(ref_counts)
    rays_per_iter = ref_counts["closest_rays"] + ref_counts["any_rays"]

    # ---- value: inputs resident in HBM.  Light samples of all iterations are generated up front and uploaded once; the K
    # iterations are handed to the device as one batch (gi_render_iterations: waves overlap consecutive iterations, result
    # bit-identical to K single calls -- tests/test_gpu_renderers.py).
    n_iter = Wm + K
    all_samples = np.concatenate([ls_host.generate(128) for _ in range(n_iter)])
    d_samples = torch.from_numpy(all_samples.view(np.uint8).reshape(-1)).cuda()
    seeds = [g.wang_hash(parallel.global_iteration(j, rank, world)) for j in range(n_iter)]
    dev.clear(); dev.reset_counters()
    dev.render_iterations(g.GI_RENDERER_PT, seeds[:Wm], PATH_LENGTH, d_samples.data_ptr(), 128)
    if dist is not None:
        # the warm-up covers the read-out collective as well: NCCL sets up its channels / NVLS buffers for a message size at first use
        # (8 ranks: tens of ms), which is not part of a steady-state step.  Same call, same shape, on a scratch image.
        scratch = torch.zeros_like(acc_t)
        for _ in range(2):
            parallel.reduce_accumulator(scratch, Wm, dist)
        del scratch
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = dev.kernel_launches()
    dev.reset_counters()
    clocks = ClockSampler(local_rank)
    clocks.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    dev.render_iterations(g.GI_RENDERER_PT, seeds[Wm:], PATH_LENGTH, d_samples.data_ptr() + Wm * 128 * 32, 128)
    local_iters = Wm + K
    if dist is not None:
        parallel.reduce_accumulator(acc_t, local_iters, dist)  # the read-out reduce over NVLink, once per job
    end.record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = start.elapsed_time(end)
    launches = dev.kernel_launches() - launches0
    cnt = dev.counters()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_paths = paths_per_iter * K * world
    rays_timed = cnt["closest_rays"] + cnt["any_rays"]
    value = total_paths / (ms_max * 1e-3) / 1e6
    mrays = rays_timed * world / (ms_max * 1e-3) / 1e6

    # ---- the dominant kernel (k_wf_trace), timed live: the same K-iteration batch once more with every launch bracketed by CUDA
    # events on the launching stream (GI_OPT_KERNEL_TIMING); scratch accumulator state is irrelevant here
    if rank == 0:
        dev.set_option(g.GI_OPT_KERNEL_TIMING, 1)
        dev.render_iterations(g.GI_RENDERER_PT, seeds[Wm:], PATH_LENGTH, d_samples.data_ptr() + Wm * 128 * 32, 128)
        dev.synchronize()
        ktimes = dev.kernel_times()
        dev.set_option(g.GI_OPT_KERNEL_TIMING, 0)
    if dist is not None:
        dist.barrier()

    # ---- relMSE of the rendered image vs the oracle is a test-suite matter (tests/test_gpu_renderers.py); here only sanity
    mean_lum = float(acc_t[..., :3].mean().item()) / max(1, local_iters * (world if dist is not None else 1))

    # ---- e2e: Renderer API with host buffers, the reference's application loop (application.cpp:287-344): per step one
    # RendererSystem.Draw() (host light sampling + 4 KB H2D + the iteration's kernels) and one read of the progressive image
    # into pinned host memory (33 MB D2H).  (a) synchronous: Draw() + ReadBackbuffer(), each read waits for the iteration's
    # last bounce.  (b) streaming (GI_OPT_STREAMING, the headline): Draw() submits one wave of a run that stays open, so
    # consecutive iterations overlap as in a batch; ReadBackbufferAsync() snapshots the accumulator in stream order
    # (complete iterations only) and copies it out on a second stream while the next wave runs; the final ReadBackbuffer()
    # drains the stream inside the timed region, so all K iterations are rendered, read, and on the host when the clock stops.
    host_imgs = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]

    def e2e_run(streaming):
        rs.SetStreaming(streaming)
        rs.SetSynchronousDraw(False)
        rs.SetIterationShard(rank, world)  # resets the iteration count, regenerates light samples
        dev.clear()
        for j in range(Wm):
            rs.Draw()
            if streaming:
                rs.ReadBackbufferAsync(host_imgs[j & 1])
            else:
                rs.ReadBackbuffer(out=host_imgs[0])
        rs.ReadBackbuffer(out=host_imgs[0])
        rs.WaitReads()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for j in range(K):
            rs.Draw()
            if streaming:
                rs.ReadBackbufferAsync(host_imgs[j & 1])
            else:
                rs.ReadBackbuffer(out=host_imgs[0])  # synchronises: the step's result is on the host
        if streaming:
            rs.ReadBackbuffer(out=host_imgs[K & 1])  # drains the stream: the image of all K iterations, on the host
            rs.WaitReads()
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        tt = torch.tensor([secs], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return total_paths / float(tt.item()) / 1e6

    e2e_sync_value = e2e_run(False)
    e2e_value = e2e_run(True)
    rs.SetStreaming(False)

    if rank != 0:
        rs.close()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # roofline of k_wf_trace: algorithmic traversal bytes (reference layout, reference-order visit counts, SURVEY 8d:
    # 32 B per box test + 52 B per triangle test + 32 B ray in + 16 B hit out) per launch / mean launch duration
    trace_ms, trace_launches = ktimes["trace"]
    trace_bytes_iter = 32 * ref_counts["node_tests"] + 52 * ref_counts["triangle_tests"] + 48 * rays_per_iter
    bytes_per_launch = trace_bytes_iter * K / max(trace_launches, 1)
    k_ms = trace_ms / max(trace_launches, 1)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    step_ms = sum(v[0] for v in ktimes.values()) / K
    traffic = l2_bytes = l2_gbs = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("k_wf_trace_dram_bytes_per_launch")
        l2_bytes = tj.get("k_wf_trace_l2_bytes_per_launch") or None
        l2_gbs = tj.get("k_wf_trace_l2_gbs_in_capture")
    except Exception:
        pass
    scene_mb = (info.num_nodes * 32 + (info.num_nodes // 2) * 64 + info.num_leaf_triangles * 64 + info.num_vertices * 28 + 25 * args.texture_size ** 2 * 8
                + WIDTH * HEIGHT * 16) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mrays_per_s": mrays, "rays_per_path": rays_timed / (paths_per_iter * K),
        "config": {"workload": "sponza-shaped atrium PathTracer 1920x1080 (BASELINE config 2), pathLength 16, 262144 triangles, 25 textured materials "
                               "(%d^2 sRGB), 4 point lights, 128 light samples/iteration" % args.texture_size,
                   "parallelism": "iterations sharded round-robin over %d rank(s); 1 NCCL sum-reduce of the accumulator at read-out" % world,
                   "schedule": "value: one gi_render_iterations batch of K iterations (waves overlap consecutive iterations, bit-identical to K single "
                               "calls); e2e: per step RendererSystem.Draw() in streaming mode (GI_OPT_STREAMING: one wave per call, iterations overlap "
                               "as in the batch, bit-identical after the drain) + ReadBackbufferAsync of the progressive image (33 MB snapshot of the "
                               "complete iterations, D2H on a second stream), final draining ReadBackbuffer inside the timed region; e2e_sync: Draw() + "
                               "blocking ReadBackbuffer per step",
                   "l2": "no flush between iterations: resident inputs ~%.0f MB (BVH + leaf block + textures + accumulator) exceed the 126 MB L2" % scene_mb,
                   "traversal": "fast (ordered short stack + exact fallback)",
                   "upload_topology": ("ploc: the bim hierarchy's leaves re-clustered on the device at upload (%.0f ms), hit ids / images bit-identical to the bim topology" % topo_ms)
                   if topo_active == g.GI_TOPOLOGY_PLOC else "reference: the bim hierarchy as stored", "scene_build_s": build_s, "mean_radiance": mean_lum},
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": 128 * 32, "d2h_bytes_per_step": WIDTH * HEIGHT * 16},
        "e2e_sync": {"value": e2e_sync_value, "unit": "Mpaths/s", "h2d_bytes_per_step": 128 * 32, "d2h_bytes_per_step": WIDTH * HEIGHT * 16},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "l2_traffic": l2_bytes, "l2_gbs_in_capture": l2_gbs,  # L2 bytes and L2 GB/s of the captured steady-state launch (ncu; SURVEY 8d: the scene is L2-resident)
                     "kernel": "k_wf_trace", "kernel_ms": k_ms, "launches_timed": trace_launches, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "peak_source": peak_src, "share_of_step": trace_ms / K / step_ms,
                     "kernel_ms_per_step": {k: v[0] / K for k, v in ktimes.items() if v[1]},
                     "per_ray": {"nodes": ref_counts["node_tests"] / rays_per_iter, "tris": ref_counts["triangle_tests"] / rays_per_iter},
                     "whole_step": {"algorithmic_bytes": bytes_per_iter, "achieved": bytes_per_iter / (ms_max / K * 1e-3) / 1e9,
                                    "frac": bytes_per_iter / (ms_max / K * 1e-3) / 1e9 / peak},
                     "note": "algorithmic bytes = reference-layout bytes (SURVEY 8d) from reference-order visit counts; the scene is L2-resident, so this "
                             "is not DRAM traffic (see traffic)"},
        "counters": {"exact_fallback_rays": cnt["exact_fallback_rays"], "closest_rays": cnt["closest_rays"], "any_rays": cnt["any_rays"],
                     "device_node_tests_per_ray": cnt["node_tests"] / max(rays_timed, 1), "device_tri_tests_per_ray": cnt["triangle_tests"] / max(rays_timed, 1)},
    }
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(scene)
        line["cpu_baseline"] = base
    emit(line)
    rs.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--texture-size", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--topology", default="ploc", choices=["ploc", "reference"], help="GI_OPT_UPLOAD_TOPOLOGY of the ordered walks (results identical)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    capture_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=_RESULT_FD))  # the ranks' stdout is the real one again
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
