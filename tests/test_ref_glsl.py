"""The oracle against the REFERENCE'S OWN SHADER SOURCE (rows a-1 .. a-5, a-9 .. a-11, a-13).

oracle/_ref/libref_glsl_*.so are gpugi/shader/{traceray, intersectiontests, materials, lightsource, random, helper, scenedata}.glsl and
pathtracer.comp themselves, compiled as C++ (oracle/ref_glsl/: a token-level rewrite + a GLSL vector shim that states the same arithmetic
definition the oracle and the CUDA path are held to).  tests/golden/ref_glsl.npz holds what they return for fixed inputs
(tests/golden/make_golden_glsl.py) and travels with the repository; the live tests run where the libraries exist (the build container and,
prebuilt, the GPU box).  Bar: bit-exact -- hit flag, triangle index, distance, barycentrics; every float of the path tracer's sum image."""
import os

import numpy as np
import pytest

from tests import ref_glsl
from tests.conftest import make_ray_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_glsl.npz")
live = pytest.mark.skipif(not ref_glsl.available(), reason="oracle/_ref/libref_glsl_*.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


@pytest.fixture
def one_thread(orc):
    """the oracle on one thread: float atomics and light-cache slots in invocation order, as in the single-threaded run of the shaders"""
    n = orc.num_threads()
    orc.set_num_threads(1)
    yield
    orc.set_num_threads(n)


def assert_same_hits(got, want, rays, what):
    anyhit = rays["flags"] == 1
    assert np.array_equal(got["leaf_tri"], want["leaf_tri"]), what + ": hit flag / triangle index"
    c = ~anyhit
    for f in ("t", "bary_y", "bary_z"):
        assert np.array_equal(got[f][c].view(np.uint32), want[f][c].view(np.uint32)), what + ": " + f


@pytest.mark.parametrize("name", ["box1", "box1_coincident", "random_tris"])
def test_oracle_traversal_equals_reference_glsl_golden(g, orc, scenes, gold, name):
    """TraceRay / TraceRayAnyHit of the oracle == the reference's traceray.glsl on the committed batches (incl. exact ties on coincident faces)."""
    from tests.golden.make_golden_glsl import scenes_for_goldens
    sc = scenes_for_goldens(scenes)[name]
    rays, want = gold[name + "_rays"], gold[name + "_hits"]
    assert (want["leaf_tri"][rays["flags"] == 0] >= 0).mean() > 0.3
    assert_same_hits(orc.trace_batch(sc.desc, rays), want, rays, name)


def test_oracle_pathtracer_equals_reference_glsl_golden(g, orc, scenes, gold):
    """Three iterations of the oracle's path tracer == main() of the reference's pathtracer.comp, every float of the sum image (box1, 64 x 64)."""
    sc = scenes.make_box1(tess=4)
    W = H = 64
    acc = np.zeros((H, W, 4), dtype=np.float32)
    for k in range(3):
        orc.render_pt(sc.desc, gold["pt_cam"], W, H, gold["pt_samples"][k], int(gold["pt_seeds"][k]), 16, accum=acc)
    assert acc[..., :3].sum() > 0
    assert np.array_equal(acc.view(np.uint32), gold["pt_box1_64"].view(np.uint32))


def test_oracle_pathtracer_textured_equals_reference_glsl_golden(g, orc, scenes, gold, sponza_small):
    """... and on the textured Sponza-shaped scene (SampleMaterialData through textureLod, 25 materials, point lights)."""
    W, H = 64, 40
    img, _ = orc.render_pt(sponza_small.desc, gold["pt_sponza_cam"], W, H, gold["pt_sponza_samples"], int(gold["pt_sponza_seed"]), 16)
    assert img[..., :3].sum() > 0
    assert np.array_equal(img.view(np.uint32), gold["pt_sponza"].view(np.uint32))


def test_oracle_other_renderers_equal_reference_glsl_golden(g, orc, scenes, gold, one_thread):
    """lighttracer.comp + showlightcaches.comp, whittedraytracer.comp, and the bidirectional path tracer (lighttracer.comp with SAVE_LIGHT_CACHE:
    the light cache entry for entry; pathtracer_bidir.comp: the image) -- the oracle on one thread equals the reference's shaders in every float."""
    sc = scenes.make_box1(tess=4)
    W = H = 64
    cam, s, seed, nr, cap = gold["pt_cam"], gold["x_samples"], int(gold["x_seed"]), int(gold["x_nr"]), int(gold["x_capacity"])
    lpt, _ = orc.render_lpt(sc.desc, cam, W, H, s, seed, nr, 16, show_light_caches=True)
    assert lpt[..., :3].sum() > 0 and np.array_equal(lpt.view(np.uint32), gold["lpt_box1_64"].view(np.uint32))
    wh, _ = orc.render_whitted(sc.desc, cam, W, H, s, seed, 16)
    assert np.array_equal(wh.view(np.uint32), gold["whitted_box1_64"].view(np.uint32))
    bpt, _, counts, cache = orc.render_bpt(sc.desc, cam, W, H, s, seed, nr, cap, 16, return_cache=True)
    n = int(gold["bpt_entries"])
    assert counts[0] == n and np.array_equal(cache[:n].view(np.uint32), gold["bpt_cache"].view(np.uint32))
    assert np.array_equal(bpt.view(np.uint32), gold["bpt_box1_64"].view(np.uint32))


def test_oracle_shader_variants_equal_reference_glsl_golden(g, orc, scenes, gold, one_thread):
    """RECONSTRUCTION_FILTER 1-3 (screenreproject.glsl:186-223) and CHAINED_MIS (chained_mis.glsl + its hooks in lighttracer.comp, screenreproject.glsl
    and pathtracer_bidir.comp): the oracle's variants equal the reference's shaders compiled with those defines, in every float."""
    sc = scenes.make_box1(tess=4)
    W = H = 64
    cam, s, seed, nr, cap = gold["pt_cam"], gold["x_samples"], int(gold["x_seed"]), int(gold["x_nr"]), int(gold["x_capacity"])
    try:
        for k in (1, 2, 3):
            orc.set_variant(reconstruction_filter=k)
            img, _ = orc.render_lpt(sc.desc, cam, W, H, s, seed, nr, 16, show_light_caches=False)
            assert np.array_equal(img.view(np.uint32), gold["lpt_f%d_box1_64" % k].view(np.uint32)), k
        orc.set_variant(reconstruction_filter=0, chained_mis=True)
        img, _, counts, cache = orc.render_bpt(sc.desc, cam, W, H, s, seed, nr, cap, 16, return_cache=True)
        n = gold["cmis_cache"].shape[0]
        assert counts[0] == n and np.array_equal(cache[:n].view(np.uint32), gold["cmis_cache"].view(np.uint32))
        assert np.array_equal(img.view(np.uint32), gold["cmis_bpt_box1_64"].view(np.uint32))
        assert not np.array_equal(img, gold["bpt_box1_64"])   # (the variant does change the image)
    finally:
        orc.set_variant(reconstruction_filter=0, chained_mis=False)


def test_product_device_logic_equals_reference_glsl_golden(g, emul, scenes, gold):
    """The PRODUCT's device code (gi_core.h / gi_lane.h / gi_shade.h compiled for the host, tests/emul) against the shader-source goldens directly, no oracle in
    between: the lane walk of the trace kernel (hit records), the path tracer, the light tracer + showlightcaches, the Whitted tracer and the bidirectional
    path tracer incl. its light cache -- one thread, bit for bit."""
    from tests.golden.make_golden_glsl import scenes_for_goldens
    os.environ["OMP_NUM_THREADS"] = "1"
    for name, sc in scenes_for_goldens(scenes).items():
        rays, want = gold[name + "_rays"], gold[name + "_hits"]
        for mode in (0, 2):   # ordered walk of the one-thread-per-path kernels, lane state machine of k_wf_trace
            got, _ = emul.trace(sc.desc, rays, mode=mode)
            assert_same_hits(got, want, rays, "%s mode %d" % (name, mode))
    sc = scenes.make_box1(tess=4)
    W = H = 64
    cam, s, seed, nr, cap = gold["pt_cam"], gold["x_samples"], int(gold["x_seed"]), int(gold["x_nr"]), int(gold["x_capacity"])
    acc = np.zeros((H, W, 4), dtype=np.float32)
    for k in range(3):
        emul.render(sc.desc, cam, W, H, gold["pt_samples"][k], 0, int(gold["pt_seeds"][k]), 16, accum=acc)
    assert np.array_equal(acc.view(np.uint32), gold["pt_box1_64"].view(np.uint32)), "PT"
    img, _ = emul.render(sc.desc, cam, W, H, s, 3, seed, 16, mode=1)
    assert np.array_equal(img.view(np.uint32), gold["whitted_box1_64"].view(np.uint32)), "Whitted"
    img, _ = emul.render(sc.desc, cam, W, H, s, 1, seed, 16, num_rays_per_light_sample=nr, show_light_caches=True)
    assert np.array_equal(img.view(np.uint32), gold["lpt_box1_64"].view(np.uint32)), "LPT"
    cache = np.zeros(cap * 16, dtype=np.float32)
    img, cnt = emul.render(sc.desc, cam, W, H, s, 2, seed, 16, num_rays_per_light_sample=nr, capacity=cap, cache=cache)
    n = int(gold["bpt_entries"])
    assert cnt["light_cache_entries"] == n and np.array_equal(cache.reshape(-1, 16)[:n].view(np.uint32), gold["bpt_cache"].view(np.uint32)), "light cache"
    assert np.array_equal(img.view(np.uint32), gold["bpt_box1_64"].view(np.uint32)), "BPT"


def test_product_device_logic_variants_equal_reference_glsl_golden(g, emul, scenes, gold):
    """The same for the shader variants: the product's RECONSTRUCTION_FILTER 1-3 splats (one thread per path and the wavefront lane functions) and its
    CHAINED_MIS light pass + eye pass against the goldens of the reference's shaders compiled with those defines -- image and light cache, bit for bit."""
    os.environ["OMP_NUM_THREADS"] = "1"
    sc = scenes.make_box1(tess=4)
    W = H = 64
    cam, s, seed, nr, cap = gold["pt_cam"], gold["x_samples"], int(gold["x_seed"]), int(gold["x_nr"]), int(gold["x_capacity"])
    try:
        for k in (1, 2, 3):
            emul.set_variant(reconstruction_filter=k)
            img, _ = emul.render(sc.desc, cam, W, H, s, 1, seed, 16, num_rays_per_light_sample=nr)
            assert np.array_equal(img.view(np.uint32), gold["lpt_f%d_box1_64" % k].view(np.uint32)), "filter %d" % k
            wf, _, _, _ = emul.render_wavefront(sc.desc, cam, W, H, s, 1, seed, 16, num_rays_per_light_sample=nr)
            d = np.abs(wf - gold["lpt_f%d_box1_64" % k])       # the waves splat in another order: same terms, fp32 sums reordered
            assert d.max() <= 1e-5 * gold["lpt_f%d_box1_64" % k].max(), "filter %d, waves" % k
        emul.set_variant(reconstruction_filter=0, chained_mis=True)
        cache = np.zeros(cap * 16, dtype=np.float32)
        img, cnt = emul.render(sc.desc, cam, W, H, s, 2, seed, 16, num_rays_per_light_sample=nr, capacity=cap, cache=cache)
        n = gold["cmis_cache"].shape[0]
        assert cnt["light_cache_entries"] == n and np.array_equal(cache.reshape(-1, 16)[:n].view(np.uint32), gold["cmis_cache"].view(np.uint32)), "CHAINED_MIS light cache"
        assert np.array_equal(img.view(np.uint32), gold["cmis_bpt_box1_64"].view(np.uint32)), "CHAINED_MIS BPT"
    finally:
        emul.set_variant()


@live
def test_oracle_other_renderers_equal_reference_glsl_live(g, orc, scenes, sponza_small, one_thread):
    """Fresh seeds, another resolution, the textured scene with its point lights (omnidirectional light samples), a light cache that overflows."""
    lt, show, wt, lts, eye = ref_glsl.LightTracer(), ref_glsl.ShowLightCaches(), ref_glsl.WhittedRayTracer(), ref_glsl.LightTracer(save=True), ref_glsl.BidirEyePass()
    for sc, C, W, H, nls, nr, it, capf in ((scenes.make_box1(tess=2), scenes.BOX1_CAMERA, 56, 40, 32, 64, 4, 4.0), (sponza_small, scenes.SPONZA_CAMERA, 64, 40, 32, 64, 9, 0.6)):
        cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
        ls = g.LightSampler(sc); ls.seed(50 + it)
        s = ls.generate(nls)
        seed = orc.wanghash(it)
        for lib in (lt, show, wt, lts, eye):
            lib.set_scene(sc)
        want = show.render(cam, W, H, s, lt.render(cam, W, H, s, seed, nr))
        got, _ = orc.render_lpt(sc.desc, cam, W, H, s, seed, nr, 16, show_light_caches=True)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "LPT"
        got, _ = orc.render_whitted(sc.desc, cam, W, H, s, seed, 16)
        assert np.array_equal(got.view(np.uint32), wt.render(cam, W, H, s, seed).view(np.uint32)), "Whitted"
        cap = int(nls * nr * capf)
        want, cache, nent, _ = lts.render(cam, W, H, s, seed, nr, capacity=cap)
        eye.render(cam, W, H, s, seed, nr, cache, cap, nent, accum=want)
        got, _, counts, ocache = orc.render_bpt(sc.desc, cam, W, H, s, seed, nr, cap, 16, return_cache=True)
        assert counts[0] == nent and np.array_equal(ocache[:min(nent, cap)].view(np.uint32), cache[:min(nent, cap)].view(np.uint32)), "light cache"
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "BPT"


@live
def test_extreme_materials_equal_reference_glsl_live(g, orc, emul, scenes, one_thread):
    """Material corners the BASELINE scenes do not reach -- refraction index 1 and < 1 (total internal reflection), Fresnel0 = 1, albedo > 1, roughness 0 /
    1e-4 / 1000 (Phong exponents 1e20 / 1e8 / 1e-6), half-transparent glossy and emissive surfaces, emissivity 1e6 beside 1e-6 -- through every renderer:
    the reference's shaders, the oracle and the PRODUCT's device logic (host build) agree in every float of the image and of the light cache."""
    from gpugi_b200.scenes import legacy_material as lm
    os.environ["OMP_NUM_THREADS"] = "1"
    materials = {
        "DefaultMaterial": lm((0.6, 0.6, 0.6), (0, 0, 0), 1.0, 1.0),
        "Glass": lm((1, 1, 1), (1, 1, 1), 0.0, 0.0, refraction_index=1.0),
        "Floor": lm((0, 0, 0), (0, 0, 0), 1e-4, 1.0),
        "RightWall": lm((2.0, 2.0, 2.0), (0.5, 0.5, 0.5), 0.3, 0.5),
        "Wall": lm((0.5, 0.5, 0.5), (0, 0, 0), 1e3, 1.0),
        "LeftWall": lm((0.02, 0.02, 0.33), (0.9, 0.9, 0.9), 0.05, 1.0, refraction_index=2.4),
        "White": lm((1, 1, 1), (0, 0, 0), 0.9, 0.0, refraction_index=0.5),
        "LightR": lm((0.5, 0.5, 0.5), (0, 0, 0), 1.0, 1.0, (1e6, 0, 1e-6)),
        "LightL": lm((0.5, 0.5, 0.5), (0.2, 0.2, 0.2), 0.01, 0.3, (98.0, 392.2, 788.2)),
    }
    sc = g.Scene.from_mesh(*scenes.box1_mesh(2), materials, "", "sah")
    W, H, nls, nr = 56, 40, 32, 64
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    pt, lt, show, wt, lts, eye = (ref_glsl.PathTracer(), ref_glsl.LightTracer(), ref_glsl.ShowLightCaches(), ref_glsl.WhittedRayTracer(), ref_glsl.LightTracer(save=True),
                                  ref_glsl.BidirEyePass())
    for lib in (pt, lt, show, wt, lts, eye):
        lib.set_scene(sc)
    same = lambda a, b: np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for it in (2, 5):
        ls = g.LightSampler(sc); ls.seed(70 + it)
        s, seed = ls.generate(nls), orc.wanghash(it)
        want = pt.render(cam, W, H, s, seed)
        assert want[..., :3].mean() > 1 and np.isfinite(want).all()
        assert same(orc.render_pt(sc.desc, cam, W, H, s, seed, 16)[0], want) and same(emul.render(sc.desc, cam, W, H, s, 0, seed, 16)[0], want), "PT"
        want = show.render(cam, W, H, s, lt.render(cam, W, H, s, seed, nr))
        assert same(orc.render_lpt(sc.desc, cam, W, H, s, seed, nr, 16, show_light_caches=True)[0], want), "LPT, oracle"
        assert same(emul.render(sc.desc, cam, W, H, s, 1, seed, 16, num_rays_per_light_sample=nr, show_light_caches=True)[0], want), "LPT, product"
        want = wt.render(cam, W, H, s, seed)
        assert same(orc.render_whitted(sc.desc, cam, W, H, s, seed, 16)[0], want) and same(emul.render(sc.desc, cam, W, H, s, 3, seed, 16, mode=1)[0], want), "Whitted"
        cap = nls * nr * 4
        want, cache, nent, _ = lts.render(cam, W, H, s, seed, nr, capacity=cap)
        eye.render(cam, W, H, s, seed, nr, cache, cap, nent, accum=want)
        got, _, counts, ocache = orc.render_bpt(sc.desc, cam, W, H, s, seed, nr, cap, 16, return_cache=True)
        assert counts[0] == nent and same(ocache[:nent], cache[:nent]) and same(got, want), "BPT, oracle"
        pcache = np.zeros(cap * 16, dtype=np.float32)
        got, cnt = emul.render(sc.desc, cam, W, H, s, 2, seed, 16, num_rays_per_light_sample=nr, capacity=cap, cache=pcache)
        assert cnt["light_cache_entries"] == nent and same(pcache.reshape(-1, 16)[:nent], cache[:nent]) and same(got, want), "BPT, product"


@live
def test_degenerate_triangles_equal_reference_glsl_live(g, orc, emul, scenes):
    """Degenerate geometry: triangles whose first two INDICES coincide (the reference's marker of an unused leaf slot, traceray.glsl:128 -- on a real
    triangle it ends the leaf), zero area by index (y == z, x == z), coincident and collinear positions.  traceray.glsl, the oracle and the three walks of
    the product agree on hit flag, triangle and distance bits, on the SAH and the kd hierarchy."""
    from gpugi_b200.scenes import legacy_material as lm
    rng = np.random.RandomState(5)
    n = 1500
    pos = ((rng.rand(n, 3) * 10)[:, None, :] + rng.randn(n, 3, 3) * 0.5).reshape(-1, 3).astype(np.float32)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    k = n // 10
    tri[0:k, 1] = tri[0:k, 0]
    tri[k:2 * k, 2] = tri[k:2 * k, 1]
    tri[2 * k:3 * k, 2] = tri[2 * k:3 * k, 0]
    pos[tri[3 * k:4 * k, 1]] = pos[tri[3 * k:4 * k, 0]]
    pos[tri[4 * k:5 * k, 2]] = (pos[tri[4 * k:5 * k, 0]] + pos[tri[4 * k:5 * k, 1]]) / 2
    nrm, uv = np.tile(np.float32([0, 0, 1]), (pos.shape[0], 1)), np.zeros((pos.shape[0], 2), np.float32)
    tr = ref_glsl.Trace()
    for build in ("sah", "kd"):
        sc = g.Scene.from_mesh(pos, nrm, uv, tri, np.zeros(n, np.uint32), {"M": lm((0.5, 0.5, 0.5))}, "", build)
        rays = make_ray_batch(g, sc, 20000, 3)
        tr.set_scene(sc)
        want = tr.trace(g, rays)
        assert (want["leaf_tri"] >= 0).sum() > 5000
        assert_same_hits(orc.trace_batch(sc.desc, rays), want, rays, "oracle, " + build)
        for mode in (0, 1, 2):
            assert_same_hits(emul.trace(sc.desc, rays, mode)[0], want, rays, "product mode %d, %s" % (mode, build))


@live
def test_oracle_traversal_equals_reference_glsl_live(g, orc, scenes, sponza_small):
    """Fresh batches (other seeds, the Sponza-shaped scene, finite ray lengths, axis-aligned rays) through both."""
    tr = ref_glsl.Trace()
    for sc, n, seed in ((scenes.make_box1(tess=3), 30000, 21), (scenes.make_random_triangles(1200, seed=8), 30000, 22), (sponza_small, 60000, 23)):
        rays = make_ray_batch(g, sc, n, seed)
        tr.set_scene(sc)
        assert_same_hits(orc.trace_batch(sc.desc, rays), tr.trace(g, rays), rays, "live seed %d" % seed)


@live
def test_oracle_pathtracer_equals_reference_glsl_live(g, orc, scenes):
    """Other resolutions (not a multiple of 8: the work groups of pathtracer.cpp:50 leave a border), seeds and light-sample counts."""
    pt = ref_glsl.PathTracer()
    sc = scenes.make_box1(tess=2)
    pt.set_scene(sc)
    C = scenes.BOX1_CAMERA
    for W, H, nls, it in ((52, 44, 32, 3), (96, 64, 256, 11)):
        cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
        ls = g.LightSampler(sc); ls.seed(100 + it)
        s = ls.generate(nls)
        want = pt.render(cam, W, H, s, orc.wanghash(it))
        got, _ = orc.render_pt(sc.desc, cam, W, H, s, orc.wanghash(it), 16)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (W, H)
        assert (want[H // 8 * 8:] == 0).all() and (want[:, W // 8 * 8:] == 0).all()


@live
@pytest.mark.gpu
def test_cuda_traversal_equals_reference_glsl(g, scenes, sponza_small):
    """The CUDA path (gi_trace, both walks) directly against the reference's traceray.glsl -- no oracle in between."""
    tr = ref_glsl.Trace()
    dev = g.Device(0)
    try:
        for sc, n, seed in ((scenes.make_box1(tess=4), 100000, 31), (scenes.make_box1(tess=2, coincident_faces=True), 50000, 32), (sponza_small, 200000, 33)):
            rays = make_ray_batch(g, sc, n, seed)
            tr.set_scene(sc)
            want = tr.trace(g, rays)
            dev.upload_scene(sc)
            for mode in (g.GI_TRAVERSAL_FAST, g.GI_TRAVERSAL_REFERENCE):
                dev.set_option(g.GI_OPT_TRAVERSAL, mode)
                assert_same_hits(dev.trace(rays), want, rays, "cuda mode %d seed %d" % (mode, seed))
    finally:
        dev.close()


@live
@pytest.mark.gpu
def test_cuda_pathtracer_vs_reference_glsl(g, orc, scenes):
    """One path-tracer iteration of the CUDA path against pathtracer.comp itself: same paths and random numbers; what differs is the device's
    transcendental functions and FMA contraction in shading -- relMSE <= 1e-3 (the tolerance of bench.py's relmse), <= 2 % of the pixels off by > 1 %."""
    pt = ref_glsl.PathTracer()
    sc = scenes.make_box1(tess=4)
    pt.set_scene(sc)
    W = H = 128
    rs = g.RendererSystem(0)
    try:
        rs.SetScene(sc); rs.SetScreenSize(W, H)
        rs.SetCamera(scenes.BOX1_CAMERA["pos"], scenes.BOX1_CAMERA["lookat"], scenes.BOX1_CAMERA["fov"])
        rs.SetRenderer("PT")
        samples, seed, cam = rs.GetInitialLightSamples(), rs.GetFrameSeed(), rs.GetCamera()
        rs.Draw()
        img = rs.ReadBackbuffer()
    finally:
        rs.close()
    want = pt.render(np.frombuffer(bytes(cam), dtype=np.float32) if not isinstance(cam, np.ndarray) else cam, W, H, samples, seed)
    x, r = img[..., :3].astype(np.float64), want[..., :3].astype(np.float64)
    relmse = float((((x - r) ** 2) / (r ** 2 + 1e-2)).sum(axis=2).mean())
    off = float((np.abs(x - r) / (np.abs(r) + 1e-3)).max(axis=2).__gt__(1e-2).mean())
    assert relmse <= 1e-3 and off < 0.02, (relmse, off)


@live
@pytest.mark.gpu
def test_cuda_bidirectional_and_lighttracer_vs_reference_glsl(g, orc, scenes, box1):
    """The CUDA light tracer against lighttracer.comp + showlightcaches.comp, and one CUDA BPT iteration against the reference's two BPT shaders run
    on the GPU's own light cache (the slot order of the cache is scheduling dependent; everything else is a deterministic function of it)."""
    from tests.conftest import pixel_mismatch_fraction
    from tests.test_gpu_configs import _bpt_iteration_vs_oracle_on_the_same_cache, _device
    W = H = 96
    C = scenes.BOX1_CAMERA
    camp = np.frombuffer(g.camera_params(C["pos"], C["lookat"], C["fov"], W, H), dtype=np.float32)
    # BPT
    img, _, cache, entries = _bpt_iteration_vs_oracle_on_the_same_cache(g, orc, box1, C, W, H, 3)
    nls = 128
    nr = max(64, (W * H // nls // 64) * 64)
    samples = g.LightSampler(box1).generate(nls)   # (the helper draws the same first set from a fresh sampler)
    seed = g.wang_hash(3)
    cap = cache.shape[0]
    lts, eye = ref_glsl.LightTracer(save=True), ref_glsl.BidirEyePass()
    lts.set_scene(box1); eye.set_scene(box1)
    want, _, n_ref, _ = lts.render(camp, W, H, samples, seed, nr, capacity=cap)
    assert n_ref == entries
    eye.render(camp, W, H, samples, seed, nr, cache, cap, entries, accum=want)
    assert pixel_mismatch_fraction(img, want, 1e-2) < 0.02
    assert abs(float(img[..., :3].sum()) - float(want[..., :3].sum())) <= 2e-3 * float(want[..., :3].sum())
    # LPT
    dev = _device(g, box1, W, H, C)
    try:
        dev.set_light_samples(samples)
        dev.render_iteration(g.GI_RENDERER_LPT, seed, 16, num_rays_per_light_sample=nr)
        dev.show_light_caches()
        got = dev.read_accumulator(W, H)
    finally:
        dev.close()
    lt, show = ref_glsl.LightTracer(), ref_glsl.ShowLightCaches()
    lt.set_scene(box1); show.set_scene(box1)
    want = show.render(camp, W, H, samples, lt.render(camp, W, H, samples, seed, nr))
    x, r = got[..., :3].astype(np.float64), want[..., :3].astype(np.float64)
    assert float((((x - r) ** 2) / (r ** 2 + 1e-2)).sum(axis=2).mean()) <= 1e-2   # (splats: a flipped decision moves a whole contribution to another pixel)
    assert abs(x.sum() - r.sum()) <= 2e-3 * r.sum()
