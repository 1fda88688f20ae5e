#!/usr/bin/env python
"""bench.py -- headline benchmark of the light-transport hot path (BASELINE.json: "Mpaths/s & Mrays/s, Sponza
1080p PT at 1/2/4/8 B200; relMSE vs ref").

Workload (--config, default C2 = the configuration the metric is quoted on):
  C2  Sponza-shaped atrium, PathTracer 1920x1080 (262 144 triangles, 25 textured materials, camera + 4 point lights of scenes/cryteksponza.txt)
  C3  bathroom-like interior, BidirectionalPathTracer 1920x1080      C4  villa-like, LightPathTracer 3840x2160
  C5  village-like (4 M triangles), PathTracer 3840x2160              (C1, box1 512^2 x 64 spp, is a parity-test case: tests/test_gpu_configs.py)
A "step" is one progressive refresh: --iters-per-step (default 16) sample passes (iterations, pathLength 16) + one read-out of the image.
  value  whole-job Mpaths/s with every input resident in HBM before the timed region (scene, textures, the light samples of all
         iterations, accumulator); the K steps are ONE gi_render_iterations batch, timed with CUDA events on the launching stream;
         N > 1: + one reduced read-out (gi_read_accumulator_reduced: snapshot + reduce-scatter over NVLink peer memory) inside the region.
  e2e    the same metric through the reference-facing Renderer API with HOST buffers: per iteration RendererSystem.Draw() (host light
         sampling + 4-8 KB H2D + the iteration's kernels, streaming mode: one wave per call), per step one D2H read of the progressive
         RGBA32F image into pinned host memory (N > 1: every rank copies out its rows of the SUMMED image into one shared host image);
         final draining read inside the timed region.
  relmse the iterations the CPU baseline (the reference's shaders) renders for cpu_baseline, rendered on the GPU with the same light samples and frame seeds, compared
         on the oracle's pixels: mean_px sum_rgb (x - r)^2 / (r^2 + 1e-2)   (N = 1).
  --impl reference   the reference's CPU side of this path: the reference's OWN shaders (gpugi/shader/*.comp, read where they lie and compiled as
         C++ against a GLSL shim: oracle/ref_glsl, oracle/_ref/libref_glsl_*.so, kind "reference" -- the GL 4.5 / MSVC build itself cannot run here) on
         all host threads, each step a bounded sample of the same iteration; where those libraries were not built, the oracle port (kind "port").
Multi-GPU: one process per GPU (torchrun), iterations sharded round-robin over ranks, no data-path collective; scaling "weak".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATH_LENGTH = 16


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

    def __init__(self, gpu_index, period_ms=20):
        self.gpu = gpu_index
        self.period = period_ms
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.15)  # the first sample needs ~100 ms to appear
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


# ----------------------------------------------------------------------------------------------- workloads
def configs(args):
    from gpugi_b200 import scenes
    ts = args.texture_size
    return {
        "C2": {"label": "sponza-shaped atrium PathTracer 1920x1080 (BASELINE config 2), pathLength 16, 262144 triangles, 25 textured materials (%d^2 sRGB), 4 point lights" % ts,
               "short": "sponza-shaped atrium PathTracer 1920x1080 (BASELINE config 2), pathLength 16, 262144 triangles",
               "build": lambda d: scenes.make_sponza_like(d, target_triangles=262144, texture_size=ts, num_materials=25, seed=7),
               "cam": scenes.SPONZA_CAMERA, "W": 1920, "H": 1080, "renderer": "PT", "nls": 128},
        "C3": {"label": "bathroom-like interior BidirectionalPathTracer 1920x1080 (BASELINE config 3), pathLength 16, 500000 triangles, 39 materials, 2 area lights",
               "short": "bathroom-like interior BidirectionalPathTracer 1920x1080 (BASELINE config 3), pathLength 16, 500000 triangles",
               "build": lambda d: scenes.make_bathroom_like(d, 500000, min(ts, 512)), "cam": scenes.BATHROOM_CAMERA, "W": 1920, "H": 1080, "renderer": "BPT", "nls": 128},
        "C4": {"label": "villa-like LightPathTracer 3840x2160 (BASELINE config 4), pathLength 16, 1000000 triangles, glass + water, 6 point lights",
               "short": "villa-like LightPathTracer 3840x2160 (BASELINE config 4), pathLength 16, 1000000 triangles",
               "build": lambda d: scenes.make_villa_like(d, 1000000, min(ts, 512)), "cam": scenes.VILLA_CAMERA, "W": 3840, "H": 2160, "renderer": "LPT", "nls": 256},
        "C5": {"label": "village-like PathTracer 3840x2160 (BASELINE config 5), pathLength 16, 4000000 triangles (deep BVH), 48 point lights",
               "short": "village-like PathTracer 3840x2160 (BASELINE config 5), pathLength 16, 4000000 triangles",
               "build": lambda d: scenes.make_village_like(args.village_triangles), "cam": scenes.VILLAGE_CAMERA, "W": 3840, "H": 2160, "renderer": "PT", "nls": 128},
    }


def build_scene(cfg):
    d = tempfile.mkdtemp(prefix="gpugi_bench_")
    t = time.time()
    sc = cfg["build"](d)
    return sc, time.time() - t


def paths_per_iteration(cfg):
    W, H = cfg["W"], cfg["H"]
    cam = (W // 8) * (H // 8) * 64
    nr = max(64, (W * H // cfg["nls"] // 64) * 64)      # lightpathtracer.cpp:34 / bidirectionalpathtracer.cpp (NumRaysPerLightSample)
    light = cfg["nls"] * nr // 64 * 64
    return {"PT": cam, "LPT": light, "BPT": cam + light}[cfg["renderer"]], nr


def device_bytes(c):
    """Bytes of the wavefront walk's OWN layout behind the records it visits (DESIGN.md "Data layout"), from the device's visit counters
    (GI_OPT_COUNT_STEPS): 96 B per 4-wide inner-node record + 96 B per leaf record + 32 B root box per ray; exact re-traces
    (reference-order walk, 32 B per box test) are inside node_tests.  Returns (record bytes, queue bytes): the queue bytes -- 32 B ray in +
    8 B hit out per closest-hit ray, 48 B shadow record in per any-hit ray -- cross HBM once per launch whatever the caches do; the record
    bytes are re-reads of a scene that stays in L2 / L1."""
    rays = c["closest_rays"] + c["any_rays"]
    refwalk_nodes = max(0, c["node_tests"] - 4 * c["inner_node_steps"] - c["leaf_steps"] - rays)
    records = 96 * c["inner_node_steps"] + 96 * c["leaf_steps"] + 32 * rays + 32 * refwalk_nodes
    queues = 40 * c["closest_rays"] + 48 * c["any_rays"]
    return records, queues


def reference_layout_bytes(c):
    """bytes the REFERENCE layout must touch, from reference-order visit counts: per traversal 32*n_nodes + 52*n_tris + 32 (ray in) + 16 (hit out)."""
    return 32 * c["node_tests"] + 52 * c["triangle_tests"] + 48 * (c["closest_rays"] + c["any_rays"])


# ----------------------------------------------------------------------------------------------- oracle legs
def stripe_regions(cfg, num_stripes):
    W, H = cfg["W"], cfg["H"]
    ys = [int(H * (i + 0.5) / num_stripes) // 8 * 8 for i in range(num_stripes)]
    return [(0, y, W, y + 8) for y in ys]


class ReferenceShaders:
    """The reference's OWN shaders on the host cores (oracle/_ref/libref_glsl_*.so: gpugi/shader/*.comp compiled as C++ where they lie, oracle/ref_glsl/):
    cpu_baseline.kind = "reference".  None when the libraries were not built (no /root/reference at build time): the oracle port is timed instead."""

    def __init__(self, scene):
        from oracle import pyrefglsl as rg
        self.pt, self.lpt, self.show, self.lpt_save, self.eye = rg.PathTracer(), rg.LightTracer(), rg.ShowLightCaches(), rg.LightTracer(save=True), rg.BidirEyePass()
        assert self.pt.max_path_length == PATH_LENGTH
        for lib in (self.pt, self.lpt, self.show, self.lpt_save, self.eye):
            lib.set_scene(scene)

    @staticmethod
    def create(scene):
        try:
            from oracle import pyrefglsl as rg
            return ReferenceShaders(scene) if rg.available() else None
        except Exception as e:  # a library that does not load is not a reason to lose the baseline
            log("reference shaders unavailable (%s): timing the oracle port" % e)
            return None


def oracle_pass(orc, cfg, scene, cam, samples, iteration, regions, acc, light_fraction=1.0, ref=None):
    """One bounded sample of iteration `iteration` on all host threads; returns (paths, rays, seconds, counters).  ref: ReferenceShaders (the
    reference's own shaders; rays are not counted there) or None (the oracle port)."""
    W, H = cfg["W"], cfg["H"]
    tot = {}
    seed = orc.wanghash(iteration)
    _, nr = paths_per_iteration(cfg)
    gw, gh = W // 8 * 8, H // 8 * 8
    region_paths = lambda r: max(0, min(r[2], gw) - r[0]) * max(0, min(r[3], gh) - r[1])
    t0 = time.perf_counter()
    if ref is not None:
        if cfg["renderer"] == "PT":
            for r in regions:
                ref.pt.render(cam, W, H, samples, seed, accum=acc, region=r)
            tot = {"camera_paths": sum(region_paths(r) for r in regions)}
        elif cfg["renderer"] == "LPT":
            n = max(64, int(cfg["nls"] * nr * light_fraction) // 64 * 64)
            ref.lpt.render(cam, W, H, samples, seed, nr, accum=acc, first=0, count=n, parallel=True)
            if light_fraction >= 1.0:
                ref.show.render(cam, W, H, samples, acc)
            tot = {"light_paths": n}
        else:
            cap = int(cfg.get("cap") or cfg["nls"] * nr * 3)
            n = cfg["nls"] * nr // 64 * 64
            _, cache, nent, _ = ref.lpt_save.render(cam, W, H, samples, seed, nr, accum=acc, capacity=cap, parallel=True)
            ref.eye.render(cam, W, H, samples, seed, nr, cache, cap, nent, accum=acc, region=regions[0])
            tot = {"light_paths": n, "camera_paths": region_paths(regions[0])}
    elif cfg["renderer"] == "PT":
        for r in regions:
            _, cnt = orc.render_pt(scene.desc, cam, W, H, samples, seed, PATH_LENGTH, region=r, accum=acc)
            for k, v in cnt.items():
                tot[k] = tot.get(k, 0) + v
    elif cfg["renderer"] == "LPT":
        n = max(64, int(cfg["nls"] * nr * light_fraction) // 64 * 64)
        _, tot = orc.render_lpt(scene.desc, cam, W, H, samples, seed, nr, PATH_LENGTH, first_path=0, num_paths=n, show_light_caches=light_fraction >= 1.0, accum=acc)
    else:
        # the light pass is whole (the eye pass of ANY pixel needs the complete cache); the eye pass runs on one band of pixels
        cap = int(cfg.get("cap") or cfg["nls"] * nr * 3)
        _, tot, _ = orc.render_bpt(scene.desc, cam, W, H, samples, seed, nr, cap, PATH_LENGTH, region=regions[0], accum=acc)
    dt = time.perf_counter() - t0
    return tot.get("camera_paths", 0) + tot.get("light_paths", 0), tot.get("closest_rays", 0) + tot.get("any_rays", 0), dt, tot


REFERENCE_WHAT = ("the reference's own shaders (gpugi/shader/*.comp + *.glsl, read where they lie and compiled as C++: oracle/ref_glsl, oracle/_ref/libref_glsl_*.so) "
                  "on the host cores with OpenMP over invocations")
PORT_WHAT = "the oracle: a CPU restatement of the reference's shaders (the reference's GLSL / MSVC build cannot run here; oracle/_ref/libref_glsl_*.so not built)"


def oracle_setup(cfg, scene):
    from oracle import pyoracle as orc
    C = cfg["cam"]
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], cfg["W"], cfg["H"])
    return orc, cam


def cpu_baseline(cfg, scene, sample_sets, budget_s=12.0):
    """The reference's own shaders (else the oracle port) on this box's host cores over a bounded sample; also returns what was rendered (for relmse)."""
    orc, cam = oracle_setup(cfg, scene)
    ref = ReferenceShaders.create(scene)
    W, H = cfg["W"], cfg["H"]
    if cfg["renderer"] == "PT":
        regions, frac, what = stripe_regions(cfg, 9), 1.0, "9 evenly spaced 8-pixel stripes of the %dx%d frame" % (W, H)
    elif cfg["renderer"] == "LPT":
        regions, frac, what = None, 1.0, "all light paths of the iteration (splats land anywhere on the %dx%d frame)" % (W, H)
    else:
        regions, frac, what = [(0, H // 2 - 8, W, H // 2 + 8)], 1.0, "whole light pass + the eye pass of a 16-pixel band of the %dx%d frame" % (W, H)
    acc = np.zeros((H, W, 4), dtype=np.float32)
    paths = rays = it = 0
    secs = 0.0
    tot = {}
    # relmse is quoted at 64 iterations: keep going past the budget (up to 5 x) until they are there
    while (secs < budget_s or (it < 64 and secs < 5.0 * budget_s)) and it < len(sample_sets):
        p, r, dt, cnt = oracle_pass(orc, cfg, scene, cam, sample_sets[it], it, regions, acc, frac, ref)
        paths += p; rays += r; secs += dt; it += 1
        for k, v in cnt.items():
            tot[k] = tot.get(k, 0) + v
    info = {"value": paths / secs / 1e6, "unit": "Mpaths/s", "mrays_per_s": rays / secs / 1e6 if ref is None else None, "cores": orc.num_threads(),
            "kind": "reference" if ref is not None else "port",
            "what": REFERENCE_WHAT if ref is not None else PORT_WHAT,
            "sample": "%d iteration(s) x %s (%d paths, %.1f s)" % (it, what, paths, secs)}
    return info, acc, it, regions


def relmse_vs_oracle(cfg, dev, g, sample_sets, oracle_acc, iterations, regions, nr, cap, baseline_what=PORT_WHAT):
    """The oracle's iterations again on the GPU (same light samples, same FrameSeeds), compared on the pixels the oracle rendered."""
    from gpugi_b200 import parallel
    W, H = cfg["W"], cfg["H"]
    rid = g.RENDERERS[cfg["renderer"]]
    dev.set_option(g.GI_OPT_STREAMING, 0)
    dev.clear()
    for k in range(iterations):
        dev.set_light_samples(sample_sets[k])
        dev.render_iteration(rid, g.wang_hash(k), PATH_LENGTH, num_rays_per_light_sample=nr if cfg["renderer"] != "PT" else 0, light_cache_capacity=cap)
        if cfg["renderer"] == "LPT":
            dev.show_light_caches()
    img = dev.read_accumulator(W, H)
    mask = np.zeros((H, W), dtype=bool)
    if regions is None:
        mask[:] = True
    else:
        for (x0, y0, x1, y1) in regions:
            mask[y0:y1, x0:x1] = True
    x = img[..., :3] / np.float32(iterations)
    r = oracle_acc[..., :3] / np.float32(iterations)
    out = {"iterations": iterations, "pixels": int(mask.sum()), "vs": baseline_what + ", same light samples and FrameSeeds",
           "definition": "mean_px sum_rgb (x - r)^2 / (r^2 + 1e-2)"}
    if cfg["renderer"] == "BPT":
        # the eye pass picks light-cache entries by slot and the slot order is scheduling dependent: the two images are two unbiased
        # estimates, not the same numbers -> compared on 8x8 block means, tolerance statistical
        y0, y1 = regions[0][1], regions[0][3]
        bx = x[y0:y1].reshape((y1 - y0) // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))
        br = r[y0:y1].reshape((y1 - y0) // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))
        out.update({"value": parallel.rel_mse(bx, br), "tolerance": 0.5, "note": "8x8 block means; statistical (light-cache slot order is scheduling dependent)",
                    "mean_ratio": float(bx.mean() / max(br.mean(), 1e-20))})
    else:
        # same paths, same random numbers: what differs is a discrete decision flipped by a transcendental's last bit on a few paths; the number of
        # such paths grows with the iteration count N and each one's weight in the mean image falls with 1 / N, so relMSE ~ 1 / N: the
        # tolerance is stated at 64 iterations and scaled when the CPU budget allowed fewer
        out.update({"value": parallel.rel_mse(x[mask], r[mask]), "tolerance": 1e-3 * max(1.0, 64.0 / iterations), "tolerance_at_64_iterations": 1e-3,
                    "note": "same paths, same random numbers; differences are transcendental-function rounding (device vs libm) on a few pixels"
                            + (" and the order of the float atomics" if cfg["renderer"] == "LPT" else "")})
    out["ok"] = bool(out["value"] <= out["tolerance"])
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own shaders (else the oracle port) on this box's host cores, same config / metric."""
    if rank != 0:
        return
    import gpugi_b200 as g
    from oracle import pyoracle as orc
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 alone runs this arm and is meant to use every host core it may run on
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    cfg = configs(args)[args.config]
    scene, build_s = build_scene(cfg)
    orc, cam = oracle_setup(cfg, scene)
    ref = ReferenceShaders.create(scene)
    W, H = cfg["W"], cfg["H"]
    ls = g.LightSampler(scene)
    acc = np.zeros((H, W, 4), dtype=np.float32)
    # calibrate the sample so that one step takes ~1.5 s
    if cfg["renderer"] == "PT":
        oracle_pass(orc, cfg, scene, cam, ls.generate(cfg["nls"]), 0, stripe_regions(cfg, 3), acc, 1.0, ref)  # cold pass (page faults, sRGB table)
        _, _, dt, _ = oracle_pass(orc, cfg, scene, cam, ls.generate(cfg["nls"]), 0, stripe_regions(cfg, 3), acc, 1.0, ref)
        stripes = int(max(1, min(H // 8, round(3 * 1.5 / max(dt, 1e-3)))))
        regions, frac = stripe_regions(cfg, stripes), 1.0
        what = "%d evenly spaced 8-pixel stripes of one %dx%d iteration" % (stripes, W, H)
    elif cfg["renderer"] == "LPT":
        _, _, dt, _ = oracle_pass(orc, cfg, scene, cam, ls.generate(cfg["nls"]), 0, None, acc, 0.02, ref)
        frac = float(min(1.0, max(0.002, 0.02 * 1.5 / max(dt, 1e-3))))
        regions = None
        what = "%.1f %% of the light paths of one %dx%d iteration" % (100 * frac, W, H)
    else:
        regions, frac = [(0, H // 2 - 8, W, H // 2 + 8)], 1.0
        what = "the whole light pass + the eye pass of a 16-pixel band of one %dx%d iteration" % (W, H)
    for w in range(args.warmup if cfg["renderer"] == "PT" else 1):
        oracle_pass(orc, cfg, scene, cam, ls.generate(cfg["nls"]), w, regions, acc, frac, ref)
    paths = rays = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        p, r, _, _ = oracle_pass(orc, cfg, scene, cam, ls.generate(cfg["nls"]), args.warmup + k, regions, acc, frac, ref)
        paths += p; rays += r
    secs = time.perf_counter() - t0
    value = paths / secs / 1e6
    sample = "each step = %s (%d paths)" % (what, paths // max(args.steps, 1))
    line = {"impl": "reference", "metric": metric_name(cfg), "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mrays_per_s": rays / secs / 1e6 if ref is None else None,
            "config": {"workload": cfg["short"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": orc.num_threads(), "kind": "reference" if ref is not None else "port",
                             "what": REFERENCE_WHAT if ref is not None else PORT_WHAT, "sample": sample},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def metric_name(cfg):
    if cfg["renderer"] == "PT" and cfg["W"] == 1920:
        return "Mpaths/s, Sponza-shaped 1080p PathTracer (paths started per second; Mrays/s alongside)"
    return "Mpaths/s, %s (paths started per second; Mrays/s alongside)" % cfg["short"].split(",")[0]


# ----------------------------------------------------------------------------------------------- our arm
def shared_host_image(W, H, world, rank, dist, tag):
    """One RGBA32F image in host memory that every rank can write its rows into (POSIX shared memory, page-locked in every process)."""
    import torch
    nbytes = W * H * 16
    path = "/dev/shm/gpugi_b200_%s.img" % tag
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(nbytes)
    if dist is not None:
        dist.barrier()
    img = np.memmap(path, dtype=np.float32, mode="r+", shape=(H, W, 4))
    pinned = False
    try:
        pinned = int(torch.cuda.cudart().cudaHostRegister(img.ctypes.data, nbytes, 0)) == 0
    except Exception:
        pinned = False
    return img, path, pinned


def run_ours(args, rank, world, local_rank):
    import torch
    import gpugi_b200 as g
    from gpugi_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gpugi_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = parallel.init_process_group("nccl") if world > 1 else None
    K, Wm, IPS = args.steps, args.warmup, args.iters_per_step
    cfg = configs(args)[args.config]
    W, H, name = cfg["W"], cfg["H"], cfg["renderer"]
    rid = g.RENDERERS[name]
    scene, build_s = build_scene(cfg)
    info = scene.info
    C = cfg["cam"]
    ppi, nr = paths_per_iteration(cfg)

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
    rs.SetScreenSize(W, H)
    rs.SetCamera(C["pos"], C["lookat"], C["fov"])
    rs.SetRenderer(name)
    rs.SetIterationShard(rank, world)
    rs.SetSynchronousDraw(False)
    cap = 0
    if name == "BPT":
        rs.Draw()  # the first Draw runs the warm-up passes that size the light cache (bidirectionalpathtracer.cpp:76-108)
        cap = rs.GetRendererInfo()["light_cache_capacity"]
        cfg["cap"] = cap
    comm = "none"
    if world > 1:
        try:
            parallel.connect_ranks(dev, dist)
            comm = "gi_comm (CUDA IPC exchange blocks, reduce-scatter over NVLink peer loads)"
        except g.GpugiError as e:
            comm = "nccl all_reduce fallback (gi_comm unavailable: %s)" % str(e)[:120]
        flags = [None] * world
        dist.all_gather_object(flags, comm.startswith("gi_comm"))
        if not all(flags):
            comm = comm if not comm.startswith("gi_comm") else "nccl all_reduce fallback (a peer could not map the exchange blocks)"
    use_comm = comm.startswith("gi_comm")
    acc_t = parallel.accumulator_as_tensor(dev, W, H)
    first_row, num_rows = g.comm_slice(H, rank, world)
    rows_pinned = torch.empty((max(num_rows, 1), W, 4), dtype=torch.float32, pin_memory=True).numpy()

    ls_host = g.LightSampler(scene)
    ls_host.seed(4242)
    nls = cfg["nls"]
    extra = dict(num_rays_per_light_sample=nr if name != "PT" else 0, light_cache_capacity=cap)

    # ---- algorithmic bytes of one iteration: (a) what the kernel's own layout touches (device visit counters), (b) SURVEY 8d: the
    # reference layout in reference visit order
    probe = ls_host.generate(nls)
    counts = {}
    for key, opt, val in (("device", g.GI_OPT_COUNT_STEPS, 1), ("reference", g.GI_OPT_TRAVERSAL, g.GI_TRAVERSAL_REFERENCE)):
        dev.set_option(opt, val)
        dev.set_light_samples(probe); dev.reset_counters()
        dev.render_iteration(rid, g.wang_hash(rank), PATH_LENGTH, **extra)
        counts[key] = dev.counters()
        dev.set_option(opt, 0)
    rays_per_iter = counts["device"]["closest_rays"] + counts["device"]["any_rays"]
    rec_bytes_iter, queue_bytes_iter = device_bytes(counts["device"])
    dev_bytes_iter = rec_bytes_iter + queue_bytes_iter
    resident = dev.scene_resident_bytes()
    ref_bytes_iter = reference_layout_bytes(counts["reference"])

    # ---- value: inputs resident in HBM.  Light samples of all iterations are generated up front and uploaded once; the K steps are handed
    # to the device as one batch (gi_render_iterations: waves overlap consecutive iterations, result identical to single calls --
    # tests/test_gpu_renderers.py).
    n_iter = (Wm + K) * IPS
    all_samples = np.concatenate([ls_host.generate(nls) for _ in range(n_iter)])
    d_samples = torch.from_numpy(all_samples.view(np.uint8).reshape(-1)).cuda()
    seeds = [g.wang_hash(parallel.global_iteration(j, rank, world)) for j in range(n_iter)]
    dev.clear(); dev.reset_counters()
    dev.render_iterations(rid, seeds[:Wm * IPS], PATH_LENGTH, d_samples.data_ptr(), nls, **extra)
    if world > 1:
        # the warm-up covers the read-out collective as well (first use: IPC mappings are touched / NCCL sets up its channels)
        for _ in range(2):
            if use_comm:
                dev.read_accumulator_reduced(rows_pinned)
            else:
                scratch = torch.zeros_like(acc_t)
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
    dev.render_iterations(rid, seeds[Wm * IPS:], PATH_LENGTH, d_samples.data_ptr() + Wm * IPS * nls * 32, nls, **extra)
    local_iters = (Wm + K) * IPS
    if world > 1:
        if use_comm:
            dev.read_accumulator_reduced(rows_pinned)          # the read-out reduce over NVLink, once per job
        else:
            parallel.reduce_accumulator(acc_t, local_iters, dist)
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
    total_paths = ppi * K * IPS * world
    rays_timed = cnt["closest_rays"] + cnt["any_rays"]
    value = total_paths / (ms_max * 1e-3) / 1e6
    mrays = rays_timed * world / (ms_max * 1e-3) / 1e6

    # ---- the dominant kernel (k_wf_trace), timed live: the same batch once more with every launch bracketed by CUDA events on the
    # launching stream (GI_OPT_KERNEL_TIMING); scratch accumulator state is irrelevant here
    ktimes = None
    if rank == 0:
        dev.set_option(g.GI_OPT_KERNEL_TIMING, 1)
        dev.render_iterations(rid, seeds[Wm * IPS:], PATH_LENGTH, d_samples.data_ptr() + Wm * IPS * nls * 32, nls, **extra)
        dev.synchronize()
        ktimes = dev.kernel_times()
        dev.set_option(g.GI_OPT_KERNEL_TIMING, 0)
    if dist is not None:
        dist.barrier()

    # ---- e2e: Renderer API with host buffers, the reference's application loop (application.cpp:287-344): per iteration one
    # RendererSystem.Draw() (host light sampling + H2D + the iteration's kernels, GI_OPT_STREAMING: one wave per call, iterations overlap
    # as in a batch), per step one read of the progressive image into page-locked host memory that does not stall the renderer
    # (snapshot in stream order, complete iterations only, copy on a second stream).  N > 1: the read is the reduced one -- every rank
    # copies its rows of the SUMMED image into one shared host image.  The final read drains the stream inside the timed region: all
    # iterations are rendered, summed, and on the host when the clock stops.
    shared, shm_path, shm_pinned = (None, None, False)
    if world > 1:
        shared, shm_path, shm_pinned = shared_host_image(W, H, world, rank, dist, os.environ.get("MASTER_PORT", "0"))
    host_imgs = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)] if world == 1 else None

    def read_step(j, final):
        if world == 1:
            if final:
                rs.ReadBackbuffer(out=host_imgs[j & 1])
            else:
                rs.ReadBackbufferAsync(host_imgs[j & 1])
        elif use_comm:
            rows = shared[first_row:first_row + num_rows]
            if final:
                rs.ReadBackbufferReduced(rows)
            else:
                rs.ReadBackbufferReducedAsync(rows)
        else:
            # fallback: NCCL all-reduce of a snapshot, then this rank's rows to the host
            snap = acc_t.clone() if not final else acc_t
            if final:
                dev.synchronize()
            parallel.reduce_accumulator(snap, 0, dist)
            shared[first_row:first_row + num_rows] = snap[first_row:first_row + num_rows].cpu().numpy()

    def e2e_run(streaming):
        rs.SetStreaming(streaming)
        rs.SetSynchronousDraw(False)
        rs.SetIterationShard(rank, world)  # resets the iteration count, regenerates light samples
        dev.clear()
        for j in range(Wm):
            for _ in range(IPS):
                rs.Draw()
            read_step(j, not streaming)
        read_step(Wm, True)
        rs.WaitReads()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for j in range(K):
            for _ in range(IPS):
                rs.Draw()
            read_step(j, (not streaming) or j == K - 1)
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
    e2e_iters = (Wm + K) * IPS
    if world > 1:
        dist.barrier()
        mean_lum = float(np.asarray(shared[..., :3]).mean()) / max(1, e2e_iters * world) if rank == 0 else 0.0
    else:
        mean_lum = float(host_imgs[(K - 1) & 1][..., :3].mean()) / max(1, e2e_iters)
    d2h_step = W * H * 16 if world == 1 else num_rows * W * 16

    if rank != 0:
        rs.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    if shm_path:
        try:
            os.unlink(shm_path)
        except OSError:
            pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # roofline of k_wf_trace against HBM.  achieved = ALGORITHMIC HBM bytes per launch / mean launch duration: what one launch has to move
    # through HBM however well the caches work -- its queue records (rays in, hits out, shadow records in) plus the traversal records it
    # visits, each counted ONCE per launch (the inner-node and leaf records are 20-700 MB, re-read from L2 / L1 thousands of times; only the
    # first touch can come from HBM).  Both factors are totals over the SAME launches of the timed batch (steady-state waves and drain tail).
    trace_ms, trace_launches = ktimes["trace"]
    total_iters = K * IPS
    records_per_launch = rec_bytes_iter * total_iters / max(trace_launches, 1)
    queues_per_launch = queue_bytes_iter * total_iters / max(trace_launches, 1)
    resident_walk = resident["inner_records"] + resident["leaf_records"]
    bytes_per_launch = queues_per_launch + min(records_per_launch, resident_walk)
    k_ms = trace_ms / max(trace_launches, 1)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    records_gbs = records_per_launch / (k_ms * 1e-3) / 1e9
    step_ms = sum(v[0] for v in ktimes.values()) / K
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    pk = prof.get(args.config, prof if args.config == "C2" else {})
    traffic = pk.get("k_wf_trace_dram_bytes_per_launch")
    steady_ms = pk.get("k_wf_trace_duration_ms_in_capture")
    scene_mb = (info.num_nodes * 32 + (info.num_nodes // 2) * 64 + info.num_leaf_triangles * 64 + info.num_vertices * 28 + W * H * 16) / 1e6
    dc, rc = counts["device"], counts["reference"]
    line = {
        "metric": metric_name(cfg), "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mrays_per_s": mrays, "rays_per_path": rays_timed / (ppi * K * IPS), "ms_per_iteration": ms_max / (K * IPS),
        "config": {"workload": cfg["label"] + ", %d light samples/iteration" % nls, "config": args.config, "iterations_per_step": IPS,
                   "parallelism": "iterations sharded round-robin over %d rank(s); read-out: %s" % (world, comm if world > 1 else "single GPU"),
                   "schedule": "value: one gi_render_iterations batch of steps x iterations_per_step iterations (waves overlap consecutive iterations, result "
                               "identical to single calls); e2e: per iteration RendererSystem.Draw() in streaming mode (GI_OPT_STREAMING: one wave per call) "
                               "with host light sampling, per step one asynchronous read of the progressive image (snapshot of the complete iterations, D2H "
                               "on a second stream; N > 1: reduce-scatter first, every rank copies its rows of the summed image into one shared host "
                               "image), final draining read inside the timed region; e2e_sync: Draw() x iterations_per_step + blocking read per step",
                   "l2": "no flush between iterations: resident inputs ~%.0f MB (BVH + leaf block + accumulator, + textures) and %.1f GB of ray queues / path state "
                         "exceed the 126 MB L2" % (scene_mb, ppi * PATH_LENGTH * 156 / 1e9),
                   "traversal": "fast (ordered short stack + exact fallback)",
                   "upload_topology": ("ploc: the bim hierarchy's leaves re-clustered on the device at upload (%.0f ms), hit ids / images bit-identical to the bim topology" % topo_ms)
                   if topo_active == g.GI_TOPOLOGY_PLOC else "reference: the bim hierarchy as stored", "scene_build_s": build_s, "mean_radiance": mean_lum,
                   "shared_host_image_pinned": shm_pinned if world > 1 else None},
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": nls * 32 * IPS, "d2h_bytes_per_step": d2h_step},
        "e2e_sync": {"value": e2e_sync_value, "unit": "Mpaths/s", "h2d_bytes_per_step": nls * 32 * IPS, "d2h_bytes_per_step": d2h_step},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "samples": clk["samples"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_wf_trace", "kernel_ms": k_ms, "launches_timed": trace_launches, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "peak_source": peak_src, "share_of_step": trace_ms / K / step_ms,
                     "kernel_ms_per_iteration": {k: v[0] / total_iters for k, v in ktimes.items() if v[1]},
                     "definition": "achieved = algorithmic HBM bytes per k_wf_trace launch / its mean launch duration (CUDA events on the launching stream). "
                                   "Algorithmic HBM bytes = queue records (32 B ray in + 8 B hit out per closest-hit ray, 48 B shadow record in per any-hit ray) + the "
                                   "traversal records the launch visits counted once (min(visited record bytes, resident inner + leaf records)): the scene stays in "
                                   "L2 / L1, so the kernel is NOT HBM-bound and the fraction is low by design -- `traffic` (ncu dram bytes of one steady-state launch) "
                                   "agrees with the algorithmic figure, `records` is what L2 / L1 serve, `issue` (ncu) is the roof that binds: issue slots / ALU pipe "
                                   "at the measured lanes per instruction",
                     "hbm_bytes_per_launch": {"queues": queues_per_launch, "records_once": min(records_per_launch, resident_walk)},
                     "records": {"bytes_per_launch": records_per_launch, "gbs": records_gbs, "ratio_to_hbm_peak": records_gbs / peak, "served_by": "L2 / L1",
                                 "resident_inner_record_bytes": resident["inner_records"], "resident_leaf_record_bytes": resident["leaf_records"],
                                 "note": "visit counters of the device (GI_OPT_COUNT_STEPS pass) x record size (96 B per 4-wide inner-node record, 96 B per leaf record, "
                                         "32 B root box per ray, 32 B per box test of an exact reference-order re-trace); a bandwidth of the cache hierarchy, not of HBM"},
                     "per_ray": {"inner_node_records": dc["inner_node_steps"] / max(rays_per_iter, 1), "leaf_records": dc["leaf_steps"] / max(rays_per_iter, 1),
                                 "box_tests": dc["node_tests"] / max(rays_per_iter, 1), "triangle_tests": dc["triangle_tests"] / max(rays_per_iter, 1),
                                 "bytes": dev_bytes_iter / max(rays_per_iter, 1)},
                     "dram": {"bytes_per_steady_launch": traffic, "steady_launch_ms_in_capture": steady_ms,
                              "gbs": (traffic / (steady_ms * 1e-3) / 1e9) if traffic and steady_ms else None,
                              "frac": (traffic / (steady_ms * 1e-3) / 1e9 / peak) if traffic and steady_ms else None,
                              "algorithmic_bytes_steady_launch": dev_bytes_iter},
                     "l2": {"bytes_per_steady_launch": pk.get("k_wf_trace_l2_bytes_per_launch"), "gbs_in_capture": pk.get("k_wf_trace_l2_gbs_in_capture")},
                     "issue": pk.get("k_wf_trace_issue"),
                     "reference_layout": {"bytes_per_iteration": ref_bytes_iter, "equivalent_gbs": ref_bytes_iter * total_iters / max(trace_ms, 1e-9) / 1e6,
                                          "ratio_to_hbm_peak": ref_bytes_iter * total_iters / max(trace_ms, 1e-9) / 1e6 / peak,
                                          "per_ray": {"nodes": rc["node_tests"] / max(rays_per_iter, 1), "tris": rc["triangle_tests"] / max(rays_per_iter, 1)},
                                          "note": "SURVEY 8d figure of merit: bytes the REFERENCE layout would touch in reference visit order / the trace time. "
                                                  "Not a bandwidth of this kernel (it walks fewer, smaller records, nearest first)"}},
        "counters": {"exact_fallback_rays": cnt["exact_fallback_rays"], "closest_rays": cnt["closest_rays"], "any_rays": cnt["any_rays"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        sets = [ls_host.generate(nls) for _ in range(64)]
        base, oracle_acc, its, regions = cpu_baseline(cfg, scene, sets, budget_s=args.cpu_budget)
        line["cpu_baseline"] = base
        line["relmse"] = relmse_vs_oracle(cfg, dev, g, sets, oracle_acc, its, regions, nr, cap, base.get("what", PORT_WHAT))
    emit(line)
    rs.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"], help="BASELINE.json config (default C2: the one the metric is quoted on)")
    ap.add_argument("--iters-per-step", type=int, default=16, help="sample passes per step (one progressive refresh + read-out)")
    ap.add_argument("--texture-size", type=int, default=1024)
    ap.add_argument("--village-triangles", type=int, default=4000000)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of oracle work for cpu_baseline / relmse")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--topology", default="ploc", choices=["ploc", "reference"], help="GI_OPT_UPLOAD_TOPOLOGY of the ordered walks (results identical)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    args.iters_per_step = max(args.iters_per_step, 1)
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
