"""Parity tests proper: the CUDA kernels, called through the C ABI (gpugi_b200.Device == include/gpugi_b200.h) and
through the Renderer API mirror (RendererSystem), against the CPU oracle on the same seeded inputs.

Bar: bit-exact for hit flag / triangle id / distance / barycentrics; images: identical RNG streams and operation
order, and the device code is built without FMA contraction (gpugi_b200/Makefile, FMAD=false: every float operation is the one
IEEE binary32 operation of the GLSL source, as in the oracle), so a pixel only differs where CUDA's sinf / cosf / powf / atan2f
(<= 2 ulp against libm) flips a discrete decision -- tolerance, stated per test: <= 0.2 - 0.5 % of the pixels off by more than 1 %
at one iteration (measured on B200: 0 on box1 and the stand-in scenes, 0.08 % on the 1080p Sponza-shaped scene; the contracted
build had needed 2 - 3 %) and relMSE <= 1e-6 on box1 at 64 iterations (measured 2.5e-9).  profiles/r02_parity_stats_fmadoff.jsonl
holds every measured statistic (GPUGI_B200_PARITY_LOG)."""
import os

import numpy as np
import pytest

from tests.conftest import assert_hits_equal, camera_ray_batch, make_ray_batch, pixel_mismatch_fraction

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAST, REFERENCE = 0, 1


@pytest.fixture(scope="module")
def dev(g):
    d = g.Device(0)
    yield d
    d.close()


def _trace(g, dev, scene, rays, mode):
    dev.upload_scene(scene)
    dev.set_option(g.GI_OPT_TRAVERSAL, mode)
    dev.reset_counters()
    hits = dev.trace(rays)
    return hits, dev.counters()


@pytest.mark.parametrize("mode", [FAST, REFERENCE])
def test_trace_parity_box1(g, orc, dev, scenes, box1, mode):
    C = scenes.BOX1_CAMERA
    rays = np.concatenate([camera_ray_batch(g, orc, box1, C["pos"], C["lookat"], C["fov"], 128, 128, 1), make_ray_batch(g, box1, 200000, 2)])
    hits, cnt = _trace(g, dev, box1, rays, mode)
    assert_hits_equal(hits, orc.trace_batch(box1.desc, rays), "box1 mode %d" % mode)
    assert cnt["closest_rays"] + cnt["any_rays"] == len(rays)
    if mode == REFERENCE:
        assert cnt["exact_fallback_rays"] == 0


@pytest.mark.parametrize("mode", [FAST, REFERENCE])
def test_trace_parity_golden_file(g, dev, box1, mode):
    z = np.load(os.path.join(GOLDEN, "box1_rays.npz"))
    hits, _ = _trace(g, dev, box1, z["rays"], mode)
    assert_hits_equal(hits, z["hits"], "golden mode %d" % mode)


@pytest.mark.parametrize("mode", [FAST, REFERENCE])
@pytest.mark.parametrize("build", ["sah", "kd"])
def test_trace_parity_random_triangles(g, orc, dev, scenes, mode, build):
    sc = scenes.make_random_triangles(20000, seed=7, build_method=build)
    rays = make_ray_batch(g, sc, 300000, 3)
    hits, _ = _trace(g, dev, sc, rays, mode)
    assert_hits_equal(hits, orc.trace_batch(sc.desc, rays), "random %s mode %d" % (build, mode))


def test_trace_parity_sponza_mix(g, orc, dev, scenes, tmp_path):
    """The SURVEY 8d mix (primary / secondary / shadow rays) on the Sponza-shaped scene, 2^20 rays."""
    sc = scenes.make_sponza_like(str(tmp_path), texture_size=32)
    C = scenes.SPONZA_CAMERA
    rays = camera_ray_batch(g, orc, sc, C["pos"], C["lookat"], C["fov"], 640, 360, 4)
    rays = np.concatenate([rays, make_ray_batch(g, sc, (1 << 20) - len(rays), 5)]) if len(rays) < (1 << 20) else rays[: 1 << 20]
    ref = orc.trace_batch(sc.desc, rays)
    for mode in (FAST, REFERENCE):
        hits, cnt = _trace(g, dev, sc, rays, mode)
        assert_hits_equal(hits, ref, "sponza mode %d" % mode)
    assert (ref["leaf_tri"][rays["flags"] == 0] >= 0).mean() > 0.5


def test_trace_parity_exact_ties_edges_and_limits(g, orc, dev, box1, box1_coincident):
    rng = np.random.RandomState(5)
    n = 20000
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    lo, hi = np.array([0.25, 0.05, -1.3]), np.array([1.05, 0.7, -0.5])
    rays["origin"] = (lo + rng.rand(n, 3) * (hi - lo)).astype(np.float32)
    d = rng.randn(n, 3); d[:, 1] = -np.abs(d[:, 1]) - 0.2
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    hits, cnt = _trace(g, dev, box1_coincident, rays, FAST)
    assert_hits_equal(hits, orc.trace_batch(box1_coincident.desc, rays), "ties")
    assert cnt["exact_fallback_rays"] > n // 4
    # shared edges / vertices / in-plane rays
    P = box1.positions(); lt = box1.leaf_triangles()
    v = lt["v"][lt["v"][:, 0] != lt["v"][:, 1]]
    m = 30000
    pick = rng.randint(0, len(v), size=m)
    a, b = P[v[pick, 0]], P[v[pick, 1]]
    w = rng.rand(m, 1).astype(np.float32); w[: m // 3] = 0.0
    origin = np.array([0.0, 1.5, 1.0], dtype=np.float32) + rng.uniform(-0.3, 0.3, size=(m, 3)).astype(np.float32)
    dd = a * (1 - w) + b * w - origin
    e = np.zeros(m, dtype=g.RAY_DTYPE)
    e["origin"] = origin; e["direction"] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32); e["t_max"] = np.float32(g.RAY_MAX)
    flat = np.zeros(5000, dtype=g.RAY_DTYPE)
    flat["origin"] = np.stack([rng.uniform(-1.4, 1.4, 5000), np.zeros(5000), rng.uniform(-2.9, 2.2, 5000)], 1).astype(np.float32)
    fd = np.stack([rng.randn(5000), np.zeros(5000), rng.randn(5000)], 1)
    flat["direction"] = (fd / np.linalg.norm(fd, axis=1, keepdims=True)).astype(np.float32); flat["t_max"] = np.float32(g.RAY_MAX)
    # finite closest-hit limits at / next to the hit distance
    lim = make_ray_batch(g, box1, 30000, 12, any_fraction=0.0)
    t = orc.trace_batch(box1.desc, lim)["t"].copy()
    lim["t_max"][:10000] = t[:10000]
    with np.errstate(over="ignore"):
        lim["t_max"][10000:20000] = np.nextafter(t[10000:20000], np.float32(np.inf))
    lim["t_max"][20000:] = t[20000:] * np.float32(0.5)
    lim["t_max"][~np.isfinite(lim["t_max"])] = np.float32(10.0)
    batch = np.concatenate([e, flat, lim])
    ref = orc.trace_batch(box1.desc, batch)
    for mode in (FAST, REFERENCE):
        hits, _ = _trace(g, dev, box1, batch, mode)
        assert_hits_equal(hits, ref, "edges/limits mode %d" % mode)


def test_trace_ragged_and_deep(g, orc, dev, scenes):
    """Empty batch, 1-triangle scene (root is a leaf), half-filled leaves, and a list-like deep tree (stack overflow)."""
    sc = scenes.make_random_triangles(50, seed=1)
    dev.upload_scene(sc)
    assert dev.trace(np.zeros(0, dtype=g.RAY_DTYPE)).shape == (0,)
    for n in (1, 2, 3, 5):
        s = scenes.make_random_triangles(n, seed=n)
        rng = np.random.RandomState(n)
        rays = np.zeros(4096, dtype=g.RAY_DTYPE)
        rays["origin"] = rng.uniform(-12, 12, size=(4096, 3)).astype(np.float32)
        c = s.positions().mean(axis=0)
        d = c - rays["origin"] + rng.uniform(-1.5, 1.5, size=(4096, 3))
        rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        rays["t_max"] = np.float32(g.RAY_MAX)
        for mode in (FAST, REFERENCE):
            hits, _ = _trace(g, dev, s, rays, mode)
            assert_hits_equal(hits, orc.trace_batch(s.desc, rays), "%d triangles mode %d" % (n, mode))
    n = 4000
    x = np.arange(n, dtype=np.float32)[:, None] * (1.5 ** np.linspace(0, 8, n))[:, None].astype(np.float32) * 0.001
    p = np.zeros((n, 3, 3), dtype=np.float32)
    p[:, :, 0] = x + np.array([0, 0.4, 0.2], np.float32); p[:, 1, 1] = 1.0; p[:, 2, 2] = 1.0
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    deep = g.Scene.from_mesh(p.reshape(-1, 3), np.tile(np.float32([0, 1, 0]), (3 * n, 1)), np.zeros((3 * n, 2), np.float32), tri, np.zeros(n, np.uint32),
                             {"m": scenes.legacy_material()})
    rays = make_ray_batch(g, deep, 50000, 6)
    hits, cnt = _trace(g, dev, deep, rays, FAST)
    assert_hits_equal(hits, orc.trace_batch(deep.desc, rays), "deep tree")


def test_far_origins_and_offset_scenes_stay_exact(g, orc, dev, scenes):
    """The trace kernel at and beyond the domain of its conservative inner-node test (gi_lane.h: wfSlabSetup), and on a scene 30 000 units away
    from the coordinate origin (the node records are relative to the root box's centre): ray origins 40 x (inside the domain), 80 x and
    5000 x (outside: answered by the exact reference walk) the scene's half extent away, closest-hit and any-hit -- bit-exact; the offset
    scene does as many box tests per ray as the centred one (same triangles, same rays up to the shift)."""
    from tests.test_emul_parity import _far_and_offset_cases
    tests_per_ray = []
    for sc, rays in _far_and_offset_cases(g, scenes):
        ref = orc.trace_batch(sc.desc, rays)
        for mode in (FAST, REFERENCE):
            hits, cnt = _trace(g, dev, sc, rays, mode)
            assert_hits_equal(hits, ref, "far / offset mode %d" % mode)
            if mode == FAST:
                assert cnt["exact_fallback_rays"] >= 0
        near = make_ray_batch(g, sc, 20000, 3, axis_aligned=False)
        dev.set_option(g.GI_OPT_COUNT_STEPS, 1)
        hits, cnt = _trace(g, dev, sc, near, FAST)
        dev.set_option(g.GI_OPT_COUNT_STEPS, 0)
        assert_hits_equal(hits, orc.trace_batch(sc.desc, near), "offset scene, rays inside")
        tests_per_ray.append(cnt["node_tests"] / len(near))
    assert abs(tests_per_ray[1] / tests_per_ray[0] - 1.0) < 0.03, tests_per_ray


def test_scene_resident_bytes(g, dev, box1):
    """gi_scene_resident_bytes: what bench.py's roofline counts once per launch -- 96-byte inner-node and leaf records, 32-byte reference nodes."""
    dev.upload_scene(box1)
    r = dev.scene_resident_bytes()
    leaves = box1.info.num_leaf_triangles // 2
    assert r["leaf_records"] == 96 * leaves and r["reference_nodes"] == 32 * box1.info.num_nodes
    assert r["inner_records"] % 96 == 0 and 0 < r["inner_records"] <= 96 * (leaves - 1)
    assert r["shading_records"] > 0


def test_trace_device_buffers(g, orc, dev, box1):
    """gi_trace_device on caller-owned device memory (torch is only the allocator here)."""
    import torch
    rays = make_ray_batch(g, box1, 100000, 21)
    dev.upload_scene(box1)
    dev.set_option(g.GI_OPT_TRAVERSAL, FAST)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.zeros(len(rays) * 16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    dev.trace_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
    dev.synchronize()
    hits = d_hits.cpu().numpy().view(g.HIT_DTYPE)
    assert_hits_equal(hits, orc.trace_batch(box1.desc, rays), "device buffers")


def test_error_behaviour(g, dev, box1):
    d = g.Device(0)
    with pytest.raises(g.GpugiError, match="no scene"):
        d.trace(np.zeros(4, dtype=g.RAY_DTYPE))
    with pytest.raises(g.GpugiError, match="no scene"):
        d.render_iteration(g.GI_RENDERER_PT, 1)
    d.upload_scene(box1)
    with pytest.raises(g.GpugiError, match="gi_set_screen_size"):
        d.render_iteration(g.GI_RENDERER_PT, 1)
    d.set_screen_size(64, 64)
    with pytest.raises(g.GpugiError, match="gi_set_camera"):
        d.render_iteration(g.GI_RENDERER_PT, 1)
    with pytest.raises(g.GpugiError):
        g.Device(99)
    # the other renderers' preconditions, option values, the device builders' inputs
    from gpugi_b200 import scenes as _scenes
    d.set_camera(g.camera_params(_scenes.BOX1_CAMERA["pos"], _scenes.BOX1_CAMERA["lookat"], _scenes.BOX1_CAMERA["fov"], 64, 64))
    with pytest.raises(g.GpugiError, match="gi_set_light_samples"):
        d.render_iteration(g.GI_RENDERER_PT, 1)
    d.set_light_samples(g.LightSampler(box1).generate(128))
    with pytest.raises(g.GpugiError, match="max_path_length"):
        d.render_iteration(g.GI_RENDERER_PT, 1, 0)
    with pytest.raises(g.GpugiError, match="num_rays_per_light_sample"):
        d.render_iteration(g.GI_RENDERER_LPT, 1)
    with pytest.raises(g.GpugiError, match="gi_alloc_light_cache"):
        d.render_iteration(g.GI_RENDERER_BPT, 1, 16, num_rays_per_light_sample=64, light_cache_capacity=100000)
    d.alloc_light_cache(100000)
    with pytest.raises(g.GpugiError, match="smaller than the number of light paths"):
        d.render_iteration(g.GI_RENDERER_BPT, 1, 16, num_rays_per_light_sample=64, light_cache_capacity=128 * 64 - 1)
    with pytest.raises(g.GpugiError, match="exceeds the allocated"):
        d.render_iteration(g.GI_RENDERER_BPT, 1, 16, num_rays_per_light_sample=64, light_cache_capacity=100001)
    with pytest.raises(g.GpugiError, match="unknown renderer"):
        d.render_iteration(7, 1)
    for opt, val in ((g.GI_OPT_TRAVERSAL, 5), (g.GI_OPT_SCHEDULER, 9), (1234, 0)):
        with pytest.raises(g.GpugiError):
            d.set_option(opt, val)
    d.render_iterations(g.GI_RENDERER_PT, [], 16)                      # an empty batch is a no-op
    with pytest.raises(g.GpugiError, match="out of range"):
        d.build_hierarchy(np.zeros((3, 3), np.float32), np.array([[0, 1, 3]], np.uint32))
    with pytest.raises(g.GpugiError, match="empty mesh"):
        d.build_hierarchy(np.zeros((3, 3), np.float32), np.zeros((0, 3), np.uint32))
    # after all these refusals the context still renders
    d.render_iteration(g.GI_RENDERER_PT, g.wang_hash(0), 16)
    assert d.read_accumulator(64, 64)[..., :3].sum() > 0
    d.close()


def test_obox_hierarchy_on_device(g, orc, scenes):
    """bvhType = 1: the oriented-box walk on the GPU (gi_trace, both schedulers) equals the oracle bit for bit, and a PT iteration on the
    OBox hierarchy matches the oracle's and -- the boxes being conservative -- the AABox image of the same tree."""
    sc = scenes.make_random_triangles(3000, seed=9, extent=6.0, size=1.2, bvh_type=1)
    rng = np.random.RandomState(4)
    n = 200000
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    rays["origin"] = rng.uniform(-8, 8, (n, 3))
    d = rng.randn(n, 3)
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["t_max"] = g.RAY_MAX
    rays["flags"][: n // 4] = 1
    rays["t_max"][: n // 4] = rng.rand(n // 4) * 12
    ref = orc.trace_batch(sc.desc, rays)
    dev = g.Device(0)
    dev.upload_scene(sc)
    tests = {}
    for mode in (g.GI_TRAVERSAL_REFERENCE, g.GI_TRAVERSAL_FAST):
        # GI_TRAVERSAL_FAST on an oriented-box hierarchy = the ordered walk on the reference's own tree and boxes (traceOrderedOBox): nearest
        # child first, near ties re-traced in reference order -- the same answers from fewer box tests
        dev.set_option(g.GI_OPT_TRAVERSAL, mode)
        for sched in (g.GI_SCHEDULER_WAVEFRONT, g.GI_SCHEDULER_MEGAKERNEL):
            dev.set_option(g.GI_OPT_SCHEDULER, sched)
            dev.reset_counters()
            hits = dev.trace(rays)
            assert np.array_equal(hits["leaf_tri"], ref["leaf_tri"]), (mode, sched)
            assert np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32))
            tests[(mode, sched)] = dev.counters()
    slow, fast = tests[(g.GI_TRAVERSAL_REFERENCE, g.GI_SCHEDULER_MEGAKERNEL)], tests[(g.GI_TRAVERSAL_FAST, g.GI_SCHEDULER_MEGAKERNEL)]
    assert fast["node_tests"] < 0.8 * slow["node_tests"] and fast["triangle_tests"] < slow["triangle_tests"]
    assert 0 < fast["exact_fallback_rays"] < n // 20
    dev.set_option(g.GI_OPT_TRAVERSAL, g.GI_TRAVERSAL_FAST)
    dev.close()
    # exact ties on an oriented-box hierarchy (coincident faces) and a whole BPT / LPT iteration in both walks
    tie = scenes.make_box1(tess=4, coincident_faces=True, bvh_type=1)
    from tests.conftest import make_ray_batch
    tr = make_ray_batch(g, tie, 100000, 17, any_fraction=0.25)
    want = orc.trace_batch(tie.desc, tr)
    dev = g.Device(0)
    dev.upload_scene(tie)
    got = dev.trace(tr)
    assert np.array_equal(got["leaf_tri"], want["leaf_tri"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    dev.set_screen_size(64, 64)
    C = scenes.BOX1_CAMERA
    dev.set_camera(g.camera_params(C["pos"], C["lookat"], C["fov"], 64, 64))
    dev.set_light_samples(g.LightSampler(tie).generate(256))
    lpt = {}
    for mode in (g.GI_TRAVERSAL_REFERENCE, g.GI_TRAVERSAL_FAST):
        dev.set_option(g.GI_OPT_TRAVERSAL, mode)
        dev.clear()
        dev.render_iteration(g.GI_RENDERER_LPT, g.wang_hash(1), 16, num_rays_per_light_sample=64)
        lpt[mode] = dev.read_accumulator(64, 64)
    assert lpt[g.GI_TRAVERSAL_FAST][..., :3].sum() > 0
    assert np.abs(lpt[g.GI_TRAVERSAL_FAST] - lpt[g.GI_TRAVERSAL_REFERENCE]).max() <= 1e-5 * np.abs(lpt[g.GI_TRAVERSAL_REFERENCE]).max()
    dev.close()
    from tests.conftest import pixel_mismatch_fraction
    imgs = {}
    for bvh in (0, 1):
        box = scenes.make_box1(tess=4, bvh_type=bvh)
        W = H = 128
        rs = g.RendererSystem(0)
        rs.SetScene(box); rs.SetScreenSize(W, H)
        C = scenes.BOX1_CAMERA
        rs.SetCamera(C["pos"], C["lookat"], C["fov"]); rs.SetRenderer("PT")
        rs.device.set_option(g.GI_OPT_TEXTURE_FILTER, 1)
        samples, seed, cam = rs.GetInitialLightSamples(), rs.GetFrameSeed(), rs.GetCamera()
        rs.Draw()
        imgs[bvh] = rs.ReadBackbuffer()
        if bvh == 1:
            ref_img, _ = orc.render_pt(box.desc, cam, W, H, samples, seed, 16)
            assert pixel_mismatch_fraction(imgs[1], ref_img, 1e-2) < 0.002
        rs.close()
    assert pixel_mismatch_fraction(imgs[1], imgs[0], 1e-2) < 0.002


@pytest.mark.parametrize("method", ["lbvh", "ploc"])
@pytest.mark.parametrize("which", ["random", "box1", "odd", "tiny"])
def test_device_hierarchy_build(g, orc, scenes, which, method):
    """gi_build_hierarchy: the hierarchy built on the GPU (Karras radix tree / PLOC clustering) is a valid hierarchy in the reference's layout (every triangle once,
    preorder child / escape links, exact boxes, parents), the GPU walks on it equal the oracle's walk on the same nodes bit for bit, and
    -- the closest hit being a property of the geometry, not of the tree -- the hit distances equal those found through the host-built
    SAH tree."""
    from tests.test_host import _check_hierarchy
    if which == "random":
        mk = lambda m: scenes.make_random_triangles(5000, seed=12, extent=6.0, size=0.8, build_method=m)
    elif which == "box1":
        mk = lambda m: scenes.make_box1(tess=6, build_method=m)
    elif which == "odd":
        mk = lambda m: scenes.make_random_triangles(333, seed=2, build_method=m)
    else:
        mk = lambda m: scenes.make_random_triangles(2, seed=1, build_method=m)
    dev_sc, sah_sc = mk(method), mk("sah")
    _check_hierarchy(dev_sc)
    info = dev_sc.info
    assert info.num_nodes == info.num_leaf_triangles - 1    # N nodes, (N + 1) / 2 leaves, 2 slots each
    lo, hi = np.array(info.bbox_min), np.array(info.bbox_max)
    rng = np.random.RandomState(6)
    n = 100000
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    rays["origin"] = (lo - 1 + rng.rand(n, 3) * (hi - lo + 2)).astype(np.float32)
    d = rng.randn(n, 3)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = g.RAY_MAX
    rays["flags"][: n // 4] = 1
    rays["t_max"][: n // 4] = (rng.rand(n // 4) * np.linalg.norm(hi - lo)).astype(np.float32)
    ref = orc.trace_batch(dev_sc.desc, rays)
    dev = g.Device(0)
    dev.upload_scene(dev_sc)
    hits = dev.trace(rays)
    assert np.array_equal(hits["leaf_tri"], ref["leaf_tri"]) and np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32))
    dev.close()
    sah = orc.trace_batch(sah_sc.desc, rays)
    same_t = hits["t"].view(np.uint32) == sah["t"].view(np.uint32)
    assert same_t[n // 4:].mean() > 0.9999                       # closest hits: the same distance bits (ties aside)
    assert np.array_equal(hits["leaf_tri"][: n // 4], sah["leaf_tri"][: n // 4])   # any-hit: occluded or not


def test_upload_topology_is_exact(g, orc, scenes, box1, box1_coincident, tmp_path):
    """GI_OPT_UPLOAD_TOPOLOGY (default GI_TOPOLOGY_PLOC): the ordered walks run on a tree the device re-clustered over the REFERENCE'S
    leaves at upload.  Hit flag, triangle id, distance and barycentrics must equal the oracle's bit for bit on both topologies (camera
    rays, secondaries, finite shadow rays, overlapping slivers, exact ties), and a PT iteration must be bit-identical between them."""
    C = scenes.BOX1_CAMERA
    sponza = scenes.make_sponza_like(str(tmp_path), texture_size=32)
    S = scenes.SPONZA_CAMERA
    sliver = scenes.make_random_triangles(20000, seed=7, build_method="sah")
    rng = np.random.RandomState(5)
    n = 50000
    ties = np.zeros(n, dtype=g.RAY_DTYPE)
    lo, hi = np.array([0.25, 0.05, -1.3]), np.array([1.05, 0.7, -0.5])
    ties["origin"] = (lo + rng.rand(n, 3) * (hi - lo)).astype(np.float32)
    d = rng.randn(n, 3); d[:, 1] = -np.abs(d[:, 1]) - 0.2
    ties["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ties["t_max"] = np.float32(g.RAY_MAX)
    cases = [
        (box1, np.concatenate([camera_ray_batch(g, orc, box1, C["pos"], C["lookat"], C["fov"], 128, 128, 1), make_ray_batch(g, box1, 200000, 2)])),
        (sponza, np.concatenate([camera_ray_batch(g, orc, sponza, S["pos"], S["lookat"], S["fov"], 320, 180, 4), make_ray_batch(g, sponza, 300000, 5)])),
        (sliver, make_ray_batch(g, sliver, 300000, 3)),
        (box1_coincident, ties),
    ]
    for topo in (g.GI_TOPOLOGY_PLOC, g.GI_TOPOLOGY_REFERENCE):
        dev = g.Device(0)
        dev.set_option(g.GI_OPT_UPLOAD_TOPOLOGY, topo)
        for sc, rays in cases:
            dev.upload_scene(sc)
            active, ms = dev.scene_topology()
            assert active == topo, "the re-clustered tree was rejected (boxes not nested?)"
            dev.reset_counters()
            hits = dev.trace(rays)
            assert_hits_equal(hits, orc.trace_batch(sc.desc, rays), "topology %d" % topo)
            assert dev.counters()["exact_fallback_rays"] < len(rays) // 2 or sc is box1_coincident
        dev.close()
    imgs = []
    for topo in (g.GI_TOPOLOGY_PLOC, g.GI_TOPOLOGY_REFERENCE):
        rs = g.RendererSystem(0)
        rs.device.set_option(g.GI_OPT_UPLOAD_TOPOLOGY, topo)
        rs.SetScene(sponza)
        rs.SetScreenSize(320, 184)
        rs.SetCamera(S["pos"], S["lookat"], S["fov"])
        rs.SetRenderer("PT")
        for _ in range(3):
            rs.Draw()
        imgs.append(rs.ReadBackbuffer().copy())
        assert rs.device.scene_topology()[0] == topo
        rs.close()
    assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32)), "PT images differ between the two topologies"


@pytest.mark.parametrize("ntri", [2, 3, 4, 5, 7, 9, 33])
def test_upload_topology_small_scenes(g, orc, scenes, ntri):
    """1 .. 17 leaves through the device re-clustering: single-leaf scenes (nothing to cluster), one merge, wide nodes with unused slots."""
    sc = scenes.make_random_triangles(ntri, seed=ntri, build_method="sah")
    rays = make_ray_batch(g, sc, 50000, 11)
    ref = orc.trace_batch(sc.desc, rays)
    for topo in (g.GI_TOPOLOGY_PLOC, g.GI_TOPOLOGY_REFERENCE):
        dev = g.Device(0)
        dev.set_option(g.GI_OPT_UPLOAD_TOPOLOGY, topo)
        dev.upload_scene(sc)
        nleaves = sc.info.num_leaf_triangles // 2
        assert dev.scene_topology()[0] == (topo if nleaves >= 2 else g.GI_TOPOLOGY_REFERENCE)
        assert_hits_equal(dev.trace(rays), ref, "%d triangles topology %d" % (ntri, topo))
        dev.close()


def test_upload_topology_coincident_geometry(g, orc, scenes):
    """4000 copies of the same two triangles: every leaf box is identical, so the clustering finds one mutual pair per pass -- the forced
    pairing must keep the number of passes logarithmic (the upload finishes at once) and the walk exact (every hit is a 4000-fold tie)."""
    import time
    base_pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float32)
    n = 4000
    pos = np.tile(base_pos, (n, 1))
    tri = np.concatenate([np.array([[0, 1, 2], [1, 3, 2]], dtype=np.uint32) + 4 * i for i in range(n)])
    nrm = np.tile(np.array([[0, 0, 1]], dtype=np.float32), (4 * n, 1))
    uv = np.zeros((4 * n, 2), dtype=np.float32)
    sc = g.Scene.from_mesh(pos, nrm, uv, tri, np.zeros(len(tri), dtype=np.uint32), {"m": {"albedo": [0.5, 0.5, 0.5]}}, build_method="kd")
    rng = np.random.RandomState(3)
    rays = np.zeros(20000, dtype=g.RAY_DTYPE)
    rays["origin"] = np.concatenate([rng.rand(len(rays), 2) * 1.2 - 0.1, np.full((len(rays), 1), 2.0)], axis=1).astype(np.float32)
    d = rng.randn(len(rays), 3) * 0.2; d[:, 2] = -1.0
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    dev = g.Device(0)
    t0 = time.time()
    dev.upload_scene(sc)
    assert time.time() - t0 < 5.0
    assert dev.scene_topology()[0] == g.GI_TOPOLOGY_PLOC
    assert_hits_equal(dev.trace(rays), orc.trace_batch(sc.desc, rays), "coincident geometry")
    dev.close()
