"""Golden vectors FROM THE REFERENCE'S OWN SHADER SOURCE (run once, in the build container: python tests/golden/make_golden_glsl.py).

oracle/_ref/libref_glsl_trace.so and libref_glsl_pt.so are gpugi/shader/traceray.glsl, intersectiontests.glsl, materials.glsl, lightsource.glsl,
random.glsl, helper.glsl, scenedata.glsl and pathtracer.comp themselves -- read where they lie, made C++ by the token-level rewrite of
oracle/ref_glsl/glsl2cpp.py and compiled against oracle/ref_glsl/glsl_shim.hpp (recipe: oracle/Makefile).  This script runs them on fixed
inputs and stores inputs + outputs in ref_glsl.npz, which travels with the repository: the oracle is checked against it on every box
(tests/test_ref_glsl.py), where the libraries are present (here, and prebuilt on the GPU box) fresh inputs go through both and through the CUDA path.
  <scene>_rays / <scene>_hits    GiRay batches and the GiHit records of TraceRay (HIT_INDEX_OUTPUT) / TraceRayAnyHit (traceray.glsl:25-180):
                                 box1 (tess 4), box1 with coincident faces (exact ties: the first triangle in the reference's visit order wins),
                                 3000 random triangles (overlaps, slivers)
  pt_samples, pt_seeds, pt_cam   light samples, FrameSeeds and camera of ...
  pt_box1_64                     ... the SUM image of 3 iterations of pathtracer.comp's main() on box1 at 64 x 64, MAX_PATHLENGTH 16
  lpt_box1_64, whitted_box1_64,  one iteration of lighttracer.comp + showlightcaches.comp, of whittedraytracer.comp, and of the bidirectional path tracer
  bpt_box1_64, bpt_cache         (lighttracer.comp with SAVE_LIGHT_CACHE, then bidirpathtracer/pathtracer_bidir.comp) on box1 at 64 x 64 with 64 light
                                 samples x 64 rays (x_*), invocations in ascending order on one thread; bpt_cache = the light cache the light pass wrote
  lpt_f<k>_box1_64               lighttracer.comp compiled with RECONSTRUCTION_FILTER = 1, 2, 3;  cmis_bpt_box1_64 / cmis_cache: both BPT shaders with CHAINED_MIS
  pt_sponza_*                    the same for one iteration on the textured Sponza-shaped scene (30 k triangles, 64 x 36): textureLod with the
                                 sampler of scene.cpp:16-21 as defined in glsl_shim.hpp (GL 4.5 section 8.14.2, the oracle's definition)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def scenes_for_goldens(scenes):
    return {"box1": scenes.make_box1(tess=4), "box1_coincident": scenes.make_box1(tess=2, coincident_faces=True), "random_tris": scenes.make_random_triangles(3000, seed=1)}


def tie_rays(g, n=4000, seed=5):
    """rays that start inside the glass / white box of box1 and hit the floor where the box bottoms lie exactly in it"""
    rng = np.random.RandomState(seed)
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    lo, hi = np.array([0.25, 0.05, -1.3]), np.array([1.05, 0.7, -0.5])
    rays["origin"] = (lo + rng.rand(n, 3) * (hi - lo)).astype(np.float32)
    d = rng.randn(n, 3); d[:, 1] = -np.abs(d[:, 1]) - 0.2
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    return rays


def sponza_small(scenes, d):
    return scenes.make_sponza_like(d, target_triangles=30000, texture_size=64, seed=7, detail=0.3)   # = the sponza_small fixture of tests/conftest.py


def main():
    import gpugi_b200 as g
    from gpugi_b200 import scenes
    from oracle import pyoracle as orc
    from tests import ref_glsl
    from tests.conftest import make_ray_batch
    out = {}
    tr = ref_glsl.Trace()
    for name, sc in scenes_for_goldens(scenes).items():
        rays = tie_rays(g) if name == "box1_coincident" else make_ray_batch(g, sc, 20000, 3)
        tr.set_scene(sc)
        out[name + "_rays"] = rays
        out[name + "_hits"] = tr.trace(g, rays)
        print(name, "hit fraction", float((out[name + "_hits"]["leaf_tri"][rays["flags"] == 0] >= 0).mean()))
    pt = ref_glsl.PathTracer()
    assert pt.max_path_length == 16
    box1 = scenes.make_box1(tess=4)
    pt.set_scene(box1)
    W = H = 64
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    ls = g.LightSampler(box1); ls.seed(7)
    samples = np.stack([ls.generate(128) for _ in range(3)])
    seeds = np.array([orc.wanghash(k) for k in range(3)], dtype=np.uint32)
    acc = np.zeros((H, W, 4), dtype=np.float32)
    for k in range(3):
        pt.render(cam, W, H, samples[k], int(seeds[k]), accum=acc)
    out.update(pt_cam=cam, pt_samples=samples, pt_seeds=seeds, pt_box1_64=acc)
    with tempfile.TemporaryDirectory() as d:
        sp = sponza_small(scenes, d)
        pt.set_scene(sp)
        W2, H2 = 64, 40
        C2 = scenes.SPONZA_CAMERA
        cam2 = orc.camera(C2["pos"], C2["lookat"], C2["fov"], W2, H2)
        ls2 = g.LightSampler(sp); ls2.seed(9)
        s2 = ls2.generate(128)
        out.update(pt_sponza_cam=cam2, pt_sponza_samples=s2, pt_sponza_seed=np.uint32(orc.wanghash(5)),
                   pt_sponza=pt.render(cam2, W2, H2, s2, orc.wanghash(5)))
        print("sponza mean", float(out["pt_sponza"][..., :3].mean()))
    # the other renderers (one thread, invocations in ascending order: the order of the float atomics and of the light-cache slots is part of the result)
    os.environ["OMP_NUM_THREADS"] = "1"
    W = H = 64
    nr = 64
    s64 = samples[0][:64].copy()
    seed = int(orc.wanghash(2))
    lt = ref_glsl.LightTracer(); lt.set_scene(box1)
    lpt = lt.render(cam, W, H, s64, seed, nr)
    show = ref_glsl.ShowLightCaches(); show.set_scene(box1)
    show.render(cam, W, H, s64, lpt)
    wt = ref_glsl.WhittedRayTracer(); wt.set_scene(box1)
    cap = 64 * nr * 4
    lts = ref_glsl.LightTracer(save=True); lts.set_scene(box1)
    bpt, cache, nent, _ = lts.render(cam, W, H, s64, seed, nr, capacity=cap)
    eye = ref_glsl.BidirEyePass(); eye.set_scene(box1)
    eye.render(cam, W, H, s64, seed, nr, cache, cap, nent, accum=bpt)
    out.update(x_samples=s64, x_seed=np.uint32(seed), x_nr=np.int32(nr), x_capacity=np.int32(cap), lpt_box1_64=lpt, whitted_box1_64=wt.render(cam, W, H, s64, seed),
               bpt_box1_64=bpt, bpt_cache=cache[:nent].copy(), bpt_entries=np.int32(nent))
    # shader variants
    for k in (1, 2, 3):
        ltf = ref_glsl.LightTracer(reconstruction_filter=k); ltf.set_scene(box1)
        out["lpt_f%d_box1_64" % k] = ltf.render(cam, W, H, s64, seed, nr)
    ltc = ref_glsl.LightTracer(chained_mis=True); ltc.set_scene(box1)
    cb, ccache, cn, _ = ltc.render(cam, W, H, s64, seed, nr, capacity=cap)
    eyec = ref_glsl.BidirEyePass(chained_mis=True); eyec.set_scene(box1)
    eyec.render(cam, W, H, s64, seed, nr, ccache, cap, cn, accum=cb)
    out.update(cmis_bpt_box1_64=cb, cmis_cache=ccache[:cn].copy())
    print("lpt sum", float(lpt[..., :3].sum()), "bpt sum", float(bpt[..., :3].sum()), "cache entries", nent)
    np.savez_compressed(os.path.join(HERE, "ref_glsl.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_glsl.npz"), os.path.getsize(os.path.join(HERE, "ref_glsl.npz")), "bytes")


if __name__ == "__main__":
    main()
