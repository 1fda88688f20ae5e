"""The product's device logic (gpugi_b200/csrc/gi_core.h + gi_layout.h), compiled for the host by tests/emul,
against the CPU oracle.  Runs without a GPU; the `-m gpu` tests repeat the same comparisons on the real kernels."""
import os

import numpy as np
import pytest

from tests.conftest import assert_hits_equal, camera_ray_batch, make_ray_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAST, REFERENCE, LANE = 0, 1, 2  # LANE: the per-lane state machine of the wavefront trace kernel (gi_lane.h)


@pytest.mark.parametrize("mode", [FAST, REFERENCE, LANE])
def test_trace_parity_box1(g, orc, emul, scenes, box1, mode):
    C = scenes.BOX1_CAMERA
    rays = np.concatenate([camera_ray_batch(g, orc, box1, C["pos"], C["lookat"], C["fov"], 64, 64, 1), make_ray_batch(g, box1, 20000, 2)])
    hits, cnt = emul.trace(box1.desc, rays, mode)
    assert cnt["fast_valid"] == 1
    assert_hits_equal(hits, orc.trace_batch(box1.desc, rays), "box1 mode %d" % mode)
    if mode == REFERENCE:
        assert cnt["fallback_rays"] == 0


@pytest.mark.parametrize("mode", [FAST, REFERENCE, LANE])
def test_trace_parity_golden_file(emul, box1, mode):
    z = np.load(os.path.join(GOLDEN, "box1_rays.npz"))
    hits, _ = emul.trace(box1.desc, z["rays"], mode)
    assert_hits_equal(hits, z["hits"], "golden mode %d" % mode)


@pytest.mark.parametrize("mode", [FAST, REFERENCE, LANE])
@pytest.mark.parametrize("build", ["sah", "kd"])
def test_trace_parity_random_triangles(g, orc, emul, scenes, mode, build):
    """Overlapping slivers: the adversarial case for visit-order dependence."""
    sc = scenes.make_random_triangles(2500, seed=7, build_method=build)
    rays = make_ray_batch(g, sc, 40000, 3)
    hits, cnt = emul.trace(sc.desc, rays, mode)
    assert_hits_equal(hits, orc.trace_batch(sc.desc, rays), "random %s mode %d" % (build, mode))


def test_trace_parity_exact_ties(g, orc, emul, box1_coincident):
    """Box bottoms lying exactly in the floor plane: two triangles at the same distance.  The reference keeps the
    first one found in ITS visit order (traceray.glsl:145); the fast traversal must detect the tie and agree."""
    rng = np.random.RandomState(5)
    n = 4000
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    # start inside the glass box / white box, shoot downwards at the coincident floor + box bottom
    lo, hi = np.array([0.25, 0.05, -1.3]), np.array([1.05, 0.7, -0.5])
    rays["origin"] = (lo + rng.rand(n, 3) * (hi - lo)).astype(np.float32)
    d = rng.randn(n, 3); d[:, 1] = -np.abs(d[:, 1]) - 0.2
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    ref = orc.trace_batch(box1_coincident.desc, rays)
    for mode in (FAST, LANE):
        hits, cnt = emul.trace(box1_coincident.desc, rays, mode)
        assert_hits_equal(hits, ref, "ties mode %d" % mode)
        assert cnt["fallback_rays"] > n // 4  # the ties really occur and are resolved by the exact re-trace / the collecting pass


def test_trace_edges_and_vertices(g, orc, emul, box1):
    """Rays aimed exactly at shared edges / vertices of the tessellated walls, and rays inside face planes."""
    P = box1.positions()
    lt = box1.leaf_triangles()
    valid = lt["v"][:, 0] != lt["v"][:, 1]
    v = lt["v"][valid]
    rng = np.random.RandomState(8)
    pick = rng.randint(0, len(v), size=3000)
    a, b = P[v[pick, 0]], P[v[pick, 1]]
    w = rng.rand(3000, 1).astype(np.float32)
    w[:1000] = 0.0  # exactly a vertex
    target = a * (1 - w) + b * w
    origin = np.array([0.0, 1.5, 1.0], dtype=np.float32) + rng.uniform(-0.3, 0.3, size=(3000, 3)).astype(np.float32)
    d = target - origin
    rays = np.zeros(3000, dtype=g.RAY_DTYPE)
    rays["origin"] = origin
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    # in-plane rays: origin on the floor plane y = 0, direction.y == 0
    flat = np.zeros(500, dtype=g.RAY_DTYPE)
    flat["origin"] = np.stack([rng.uniform(-1.4, 1.4, 500), np.zeros(500), rng.uniform(-2.9, 2.2, 500)], 1).astype(np.float32)
    dd = np.stack([rng.randn(500), np.zeros(500), rng.randn(500)], 1)
    flat["direction"] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    flat["t_max"] = np.float32(g.RAY_MAX)
    rays = np.concatenate([rays, flat])
    for mode in (FAST, REFERENCE, LANE):
        hits, _ = emul.trace(box1.desc, rays, mode)
        assert_hits_equal(hits, orc.trace_batch(box1.desc, rays), "edges mode %d" % mode)


def test_trace_finite_closest_length(g, orc, emul, box1):
    """Closest-hit queries with a finite initial rayLength (hits exactly at / just beyond the limit)."""
    rays = make_ray_batch(g, box1, 6000, 12, any_fraction=0.0)
    full = orc.trace_batch(box1.desc, rays)
    t = full["t"].copy()
    rays["t_max"][:2000] = t[:2000]                      # limit == hit distance: `newHit < rayLength` must fail
    with np.errstate(over="ignore"):
        rays["t_max"][2000:4000] = np.nextafter(t[2000:4000], np.float32(np.inf))
    rays["t_max"][4000:] = t[4000:] * np.float32(0.5)
    rays["t_max"][~np.isfinite(rays["t_max"])] = np.float32(10.0)
    for mode in (FAST, REFERENCE, LANE):
        hits, _ = emul.trace(box1.desc, rays, mode)
        assert_hits_equal(hits, orc.trace_batch(box1.desc, rays), "finite length mode %d" % mode)


def test_trace_empty_batch_and_deep_tree(g, orc, emul, scenes):
    sc = scenes.make_random_triangles(50, seed=1)
    hits, _ = emul.trace(sc.desc, np.zeros(0, dtype=g.RAY_DTYPE), FAST)
    assert hits.shape == (0,)
    # a long thin strip of triangles sorted along x gives the SAH builder a deep, list-like tree: the short stack
    # overflows and the exact re-trace has to take over
    n = 4000
    x = np.arange(n, dtype=np.float32)[:, None] * (1.5 ** np.linspace(0, 8, n))[:, None].astype(np.float32) * 0.001
    p = np.zeros((n, 3, 3), dtype=np.float32)
    p[:, :, 0] = x + np.array([0, 0.4, 0.2], np.float32)
    p[:, 1, 1] = 1.0; p[:, 2, 2] = 1.0
    pos = p.reshape(-1, 3)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    deep = g.Scene.from_mesh(pos, np.tile(np.float32([0, 1, 0]), (3 * n, 1)), np.zeros((3 * n, 2), np.float32), tri, np.zeros(n, np.uint32),
                             {"m": scenes.legacy_material()})
    rays = make_ray_batch(g, deep, 3000, 6)
    for mode in (FAST, LANE):
        hits, cnt = emul.trace(deep.desc, rays, mode)
        assert_hits_equal(hits, orc.trace_batch(deep.desc, rays), "deep tree (levels=%d) mode %d" % (deep.info.num_tree_levels, mode))


def test_lane_parity_sponza_mix(g, orc, emul, scenes, tmp_path):
    """The wavefront lane logic on the Sponza-shaped scene: primary / secondary / shadow rays, 2^17 rays.  This mix holds
    leaves whose two triangles are hit at (almost) the same distance with a box entry distance rounded beyond them."""
    sc = scenes.make_sponza_like(str(tmp_path), texture_size=32)
    C = scenes.SPONZA_CAMERA
    rays = camera_ray_batch(g, orc, sc, C["pos"], C["lookat"], C["fov"], 224, 126, 4)
    rays = np.concatenate([rays, make_ray_batch(g, sc, (1 << 17) - len(rays), 5)])
    hits, cnt = emul.trace(sc.desc, rays, LANE)
    assert_hits_equal(hits, orc.trace_batch(sc.desc, rays), "sponza lane")
    assert cnt["fallback_rays"] > 0


def _light_sampler(orc, sc):
    return orc.LightSampler(sc.positions(), sc.triangles(), sc.triangle_materials(), sc.emissivity(), sc.point_lights())


@pytest.mark.parametrize("mode", [FAST, REFERENCE, LANE])
def test_pathtracer_matches_oracle(orc, emul, scenes, box1, mode):
    W = H = 48
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    ls = _light_sampler(orc, box1)
    for it in range(2):
        s = ls.generate(128)
        ref, rc = orc.render_pt(box1.desc, cam, W, H, s, orc.wanghash(it), 16)
        img, ec = emul.render(box1.desc, cam, W, H, s, 0, orc.wanghash(it), 16, mode=mode)
        assert np.array_equal(img, ref)  # same libm, same operation order => bit-identical
        for k in ("closest_rays", "any_rays", "camera_paths", "shaded_vertices"):
            assert ec[k] == rc[k]


def test_pathtracer_textured_scene_and_env(orc, emul, scenes, sponza_small):
    W, H = 64, 40
    C = scenes.SPONZA_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, sponza_small).generate(128)
    assert (s["normal_theta_cos"] > 1.5).all()  # point lights only
    ref, _ = orc.render_pt(sponza_small.desc, cam, W, H, s, orc.wanghash(1), 16, env_rgb=(0.2, 0.3, 0.4))
    img, _ = emul.render(sponza_small.desc, cam, W, H, s, 0, orc.wanghash(1), 16, env_rgb=(0.2, 0.3, 0.4))
    assert np.array_equal(img, ref) and ref[..., :3].max() > 0


def test_pathtracer_path_length_and_odd_resolution(orc, emul, scenes, box1):
    """MAX_PATHLENGTH as a parameter; W,H not multiples of 8 leave the right/top pixels unrendered (pathtracer.cpp:50)."""
    W, H = 44, 29
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(128)
    for max_len in (1, 3):
        ref, _ = orc.render_pt(box1.desc, cam, W, H, s, 77, max_len)
        img, _ = emul.render(box1.desc, cam, W, H, s, 0, 77, max_len)
        assert np.array_equal(img, ref)
        assert not img[:, 40:].any() and not img[24:].any() and img[:24, :40, :3].any()


def test_lighttracer_matches_oracle(orc, emul, scenes, box1):
    os.environ["OMP_NUM_THREADS"] = "1"
    W = H = 48
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(256)
    nr = 64
    ref, rc = orc.render_lpt(box1.desc, cam, W, H, s, 5, nr, 16, show_light_caches=True)
    img, ec = emul.render(box1.desc, cam, W, H, s, 1, 5, 16, num_rays_per_light_sample=nr, show_light_caches=True)
    # splats land in the same pixels with the same values; only the summation order may differ between the
    # oracle's OpenMP loop and the sequential emulation
    assert np.allclose(img, ref, rtol=1e-5, atol=1e-6)
    assert ec["splats"] == rc["splats"] and ec["light_paths"] == rc["light_paths"] == 256 * nr


def test_bidirectional_matches_oracle(orc, emul, scenes, box1):
    W = H = 40
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(128)
    nr = 64
    # warm-up pass (SAVE_LIGHT_CACHE_WARMUP): entry count and path length sum are order independent
    _, _, counts = orc.render_bpt(box1.desc, cam, W, H, s, 9, nr, 0, 16, warmup=True)
    _, wc = emul.render(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, warmup=True)
    assert (wc["light_cache_entries"], wc["light_path_length_sum"]) == counts
    assert counts[0] - 128 * nr == counts[1]  # every stored vertex beyond the light origins was counted once
    cap = counts[0] + counts[0] // 10


def test_bidirectional_image_single_thread(orc, emul, scenes, box1):
    """With one oracle thread the light-cache slot order equals the sequential emulation's => identical image."""
    import subprocess, sys
    code = r'''
import os, sys
os.environ["OMP_NUM_THREADS"] = "1"
sys.path.insert(0, %r)
import numpy as np
from gpugi_b200 import scenes
from oracle import pyoracle as orc
from tests.emul import pyemul as emul
box1 = scenes.make_box1(tess=4)
W = H = 40
C = scenes.BOX1_CAMERA
cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
s = orc.LightSampler(box1.positions(), box1.triangles(), box1.triangle_materials(), box1.emissivity(), box1.point_lights()).generate(128)
nr = 64
_, _, counts = orc.render_bpt(box1.desc, cam, W, H, s, 9, nr, 0, 16, warmup=True)
for cap in (counts[0] + counts[0] // 10, counts[0] - 500):   # roomy cache, and one that overflows (entries dropped, :27)
    ref, rc, c2 = orc.render_bpt(box1.desc, cam, W, H, s, 9, nr, cap, 16)
    img, ec = emul.render(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, capacity=cap)
    assert np.array_equal(img, ref), "BPT image differs (capacity %%d)" %% cap
    assert ec["light_cache_entries"] == c2[0] == counts[0]
    assert np.all(img[..., 3] == 1.0)   # BPT adds w = 1 (pathtracer_bidir.comp:248)
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_wavefront_lighttracer_equals_per_path_form(orc, emul, scenes, box1):
    """The light tracer as waves (gi_shade.h lane functions, the schedule of gi_wavefront.cuh run sequentially) produces the
    same splats as one invocation per path (lighttracer.comp form); only the summation order inside a pixel differs."""
    W = H = 48
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(256)
    for nr, max_len in ((64, 16), (64, 3), (128, 1)):
        img, ec = emul.render(box1.desc, cam, W, H, s, 1, 5, max_len, num_rays_per_light_sample=nr)
        wimg, wc, _, _ = emul.render_wavefront(box1.desc, cam, W, H, s, 1, 5, max_len, num_rays_per_light_sample=nr)
        assert np.allclose(wimg, img, rtol=1e-5, atol=1e-6)
        for k in ("splats", "light_paths", "shaded_vertices", "closest_rays"):
            assert wc[k] == ec[k], k
        # a splat whose pixel rounds to W or H is dropped before its any-hit query instead of after it
        assert 0 <= ec["any_rays"] - wc["any_rays"] <= ec["any_rays"] // 100 + 2


def test_wavefront_bidirectional_equals_per_path_form(orc, emul, scenes, box1):
    """BPT as waves: (1) warm-up counts, (2) the light cache holds the same entries as the per-path form (as a multiset:
    the slot order is scheduling dependent in the reference too), (3) the eye waves on a given cache are bit-identical
    to pathtracer_bidir.comp's one-invocation-per-pixel form on the same cache."""
    W = H = 40
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(128)
    nr = 64
    _, wc = emul.render(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, warmup=True)
    _, ww, _, _ = emul.render_wavefront(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, warmup=True)
    assert (ww["light_cache_entries"], ww["light_path_length_sum"]) == (wc["light_cache_entries"], wc["light_path_length_sum"])
    n = wc["light_cache_entries"]
    cap = n + n // 10
    # light pass only, as waves
    acc_l, lc, cache, count = emul.render_wavefront(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, capacity=cap, phase=1)
    assert count == n
    # per-path form of the whole iteration and the wavefront form of the whole iteration agree statistically ...
    pcache = np.zeros((cap, 16), dtype=np.float32)
    img, ec = emul.render(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, capacity=cap, cache=pcache)
    wimg, wcnt, wcache, wcount = emul.render_wavefront(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, capacity=cap)
    assert wcount == n == ec["light_cache_entries"]
    for k in ("light_paths", "camera_paths", "splats"):
        assert wcnt[k] == ec[k], k
    assert abs(wcnt["shaded_vertices"] - ec["shaded_vertices"]) == 0  # eye + light vertices do not depend on the connections
    assert np.all(wimg[..., 3] == 1.0)
    assert abs(wimg[..., :3].sum() - img[..., :3].sum()) / img[..., :3].sum() < 0.05
    # ... the cache entries of the per-path form and of the waves are the same multiset
    assert np.array_equal(cache.view(np.uint32), wcache.view(np.uint32))  # (bit patterns: the material index -1 reads as NaN)
    rows = np.ascontiguousarray(pcache[:n]).view(np.uint32).reshape(n, 16)
    wrows = np.ascontiguousarray(wcache[:n]).view(np.uint32).reshape(n, 16)
    key = lambda r: r[np.lexsort(r.T[::-1])]
    assert np.array_equal(key(rows), key(wrows))
    assert np.array_equal(rows[:128 * nr], wrows[:128 * nr])  # light origins sit in their own slots (lighttracer.comp:56-61)
    # ... and the eye pass alone, on the same cache, is bit-identical in both forms
    eye_ref, rc = emul.bidir_eye(box1.desc, cam, W, H, 9, cache, count, cap, 16, nr)
    eye_wf, ecn, _, _ = emul.render_wavefront(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, capacity=cap, phase=2, cache=cache, cache_count=count)
    assert np.array_equal(eye_wf, eye_ref)
    for k in ("camera_paths", "closest_rays", "any_rays", "shaded_vertices"):
        assert ecn[k] == rc[k], k
    # a short path length and a cache that overflows
    small = n - 500
    _, _, c2, n2 = emul.render_wavefront(box1.desc, cam, W, H, s, 2, 9, 16, num_rays_per_light_sample=nr, capacity=small, phase=1)
    assert n2 == n
    a, _ = emul.bidir_eye(box1.desc, cam, W, H, 9, c2, n2, small, 4, nr)
    b, _, _, _ = emul.render_wavefront(box1.desc, cam, W, H, s, 2, 9, 4, num_rays_per_light_sample=nr, capacity=small, phase=2, cache=c2, cache_count=n2)
    assert np.array_equal(a, b)


def test_layout_rejects_bad_scenes(g, emul, box1):
    """Scene::SanityCheck equivalents: out-of-range vertex / material / child / escape indices are refused."""
    import ctypes as C
    d = box1.desc
    nodes = box1.nodes(); lt = box1.leaf_triangles()
    rays = np.zeros(1, dtype=g.RAY_DTYPE)

    def with_arrays(nodes_arr, lt_arr):
        d2 = type(d)()
        C.memmove(C.addressof(d2), C.addressof(d), C.sizeof(d))
        d2.nodes = C.cast(nodes_arr.ctypes.data, type(d.nodes))
        d2.leaf_triangles = C.cast(lt_arr.ctypes.data, type(d.leaf_triangles))
        return d2
    bad = lt.copy(); bad["v"][0, 2] = d.num_vertices
    with pytest.raises(RuntimeError):
        emul.trace(with_arrays(nodes, bad), rays)
    bad = lt.copy(); bad["material"][0] = d.num_materials
    with pytest.raises(RuntimeError):
        emul.trace(with_arrays(nodes, bad), rays)
    badn = nodes.copy(); badn["escape"][1] = d.num_nodes
    with pytest.raises(RuntimeError):
        emul.trace(with_arrays(badn, lt), rays)
    badn = nodes.copy(); badn["first_child"][0] = d.num_nodes + 5
    with pytest.raises(RuntimeError):
        emul.trace(with_arrays(badn, lt), rays)


def test_obox_hierarchy_traversal(g, orc, emul, scenes, tmp_path):
    """bvhType = 1: FetchIntersectOBoxNode (intersectiontests.glsl:114-128) in the reference-order walk.  The product's
    device logic (host build) equals the oracle bit for bit on an oriented-box hierarchy; the boxes contain their geometry, so the hits
    equal those of the AABox hierarchy of the same tree; the OBOX_BVH section survives a .bim round trip."""
    sc0 = scenes.make_random_triangles(400, seed=5)
    sc1 = scenes.make_random_triangles(400, seed=5, bvh_type=1)
    ob, nodes = sc1.obox_nodes(), sc1.nodes()
    assert sc0.obox_nodes() is None and ob.shape[0] == nodes.shape[0]
    assert np.array_equal(ob["first_child"], nodes["first_child"]) and np.array_equal(ob["escape"], nodes["escape"])
    assert np.allclose(np.linalg.norm(ob["rotation_inv"], axis=1), 1.0, atol=1e-5) and (ob["sides_half"] >= 0).all()

    def rotate(x, q):  # intersectiontests.glsl:82-88
        t = np.cross(q[:3], x)
        return x + 2 * np.stack([q[3] * t[:, 0] + q[1] * t[:, 2] - q[2] * t[:, 1], q[3] * t[:, 1] + q[2] * t[:, 0] - q[0] * t[:, 2],
                                 q[3] * t[:, 2] + q[0] * t[:, 1] - q[1] * t[:, 0]], 1)
    P = sc1.positions().astype(np.float64)
    # the root box holds every vertex; every leaf box holds the vertices of its triangles
    assert (np.abs(rotate(P - ob[0]["center"], ob[0]["rotation_inv"].astype(np.float64))) <= ob[0]["sides_half"] * (1 + 1e-6) + 1e-6).all()
    lt = sc1.leaf_triangles()
    for i in np.nonzero(nodes["first_child"] & 0x80000000)[0][:200]:
        leaf = int(nodes["first_child"][i] & 0x7fffffff)
        for t in lt[2 * leaf: 2 * leaf + 2]:
            if t["v"][0] == t["v"][1]:
                break
            local = rotate(P[t["v"]] - ob[i]["center"], ob[i]["rotation_inv"].astype(np.float64))
            assert (np.abs(local) <= ob[i]["sides_half"] * (1 + 1e-6) + 1e-6).all()
    rng = np.random.RandomState(2)
    rays = np.zeros(30000, dtype=g.RAY_DTYPE)
    rays["origin"] = rng.uniform(-12, 12, (30000, 3))
    d = rng.randn(30000, 3)
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["t_max"] = g.RAY_MAX
    rays["flags"][:8000] = 1
    rays["t_max"][:8000] = rng.rand(8000) * 20
    ref = orc.trace_batch(sc1.desc, rays)
    work = {}
    for mode in (0, 1, 2):  # 0: the ordered walk on the oriented boxes (traceOrderedOBox); 1: reference order; 2 (lane walk, built on AABoxes): falls back to 1
        hits, cnt = emul.trace(sc1.desc, rays, mode=mode)
        assert cnt["fast_valid"] == 0
        assert np.array_equal(hits["leaf_tri"], ref["leaf_tri"]) and np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32))
        work[mode] = cnt
    assert work[0]["node_tests"] < work[1]["node_tests"] and work[2]["node_tests"] == work[1]["node_tests"]   # (a sparse soup seen from outside: little to cull)
    assert work[0]["fallback_rays"] < 1000 and work[1]["fallback_rays"] == 0   # (near ties are rare in a random soup: 0 with this seed; the coincident-face case below forces them)
    aabox = orc.trace_batch(sc0.desc, rays)
    assert (aabox["leaf_tri"] == ref["leaf_tri"]).mean() > 0.9995
    # forced ties: box bottoms lying exactly in the floor plane, on an oriented-box hierarchy -- the ordered walk must notice and re-trace in reference order
    scc = scenes.make_box1(tess=2, coincident_faces=True, bvh_type=1)
    n = 3000
    tr = np.zeros(n, dtype=g.RAY_DTYPE)
    lo, hi = np.array([0.25, 0.05, -1.3]), np.array([1.05, 0.7, -0.5])
    tr["origin"] = (lo + rng.rand(n, 3) * (hi - lo)).astype(np.float32)
    dd = rng.randn(n, 3); dd[:, 1] = -np.abs(dd[:, 1]) - 0.2
    tr["direction"] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    tr["t_max"] = np.float32(g.RAY_MAX)
    tref = orc.trace_batch(scc.desc, tr)
    thits, tcnt = emul.trace(scc.desc, tr, mode=0)
    assert np.array_equal(thits["leaf_tri"], tref["leaf_tri"]) and np.array_equal(thits["t"].view(np.uint32), tref["t"].view(np.uint32))
    assert tcnt["fallback_rays"] > n // 4
    # .bim round trip (OBOX_BVH = 0x02000000, 40-byte ei::OBox elements)
    sc1.store(str(tmp_path / "ob.json"), "ob.bim")
    sc2 = g.Scene.load(str(tmp_path / "ob.json"), bvh_type=1)
    assert sc2.obox_nodes().tobytes() == ob.tobytes()
    assert g.Scene.load(str(tmp_path / "ob.json"), bvh_type=0).obox_nodes() is None


def test_whitted_matches_oracle(orc, emul, scenes, box1):
    """whittedraytracer.comp: STOP_ON_DIFFUSE_BOUNCE sampling, coherent seed, two light samples at the diffuse stop,
    alpha += 1 -- the device logic (host build) against the oracle, bit for bit with one thread."""
    os.environ["OMP_NUM_THREADS"] = "1"
    W = H = 48
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(128)
    for seed, max_len in ((3, 16), (11, 2)):
        ref, rc = orc.render_whitted(box1.desc, cam, W, H, s, seed, max_len)
        img, ec = emul.render(box1.desc, cam, W, H, s, 3, seed, max_len, mode=1)
        assert np.array_equal(img, ref)
        assert np.all(img[..., 3] == 1.0) and img[..., :3].sum() > 0
        for k in ("camera_paths", "closest_rays", "any_rays", "shaded_vertices"):
            assert ec[k] == rc[k], k
        fast, _ = emul.render(box1.desc, cam, W, H, s, 3, seed, max_len, mode=0)
        assert np.array_equal(fast, ref)


# ---------------------------------------------------------------- GI_OPT_UPLOAD_TOPOLOGY: another tree over the reference's leaves
@pytest.fixture
def alt_topology(emul):
    yield emul
    emul.set_topology(emul.TOPOLOGY_REFERENCE)


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("mode", [FAST, LANE])
def test_retopologised_trace_is_exact(g, orc, alt_topology, scenes, box1, box1_coincident, kind, mode):
    """The ordered walks on a DIFFERENT binary tree over the same leaves (median split; random pairing = a terrible tree) return the
    reference's hit flag, triangle id, t and barycentrics bit for bit -- overlapping slivers, exact ties, finite lengths, camera rays.
    This is the property gi_upload_scene's device-side PLOC re-clustering rests on (gi_layout.h: applyLeafTree)."""
    emul = alt_topology
    emul.set_topology(kind, seed=3 + kind)
    C = scenes.BOX1_CAMERA
    cases = []
    cases.append((box1, np.concatenate([camera_ray_batch(g, orc, box1, C["pos"], C["lookat"], C["fov"], 48, 48, 1), make_ray_batch(g, box1, 12000, 2)])))
    sliver = scenes.make_random_triangles(1500, seed=7, build_method="sah")
    cases.append((sliver, make_ray_batch(g, sliver, 15000, 3)))
    odd = scenes.make_random_triangles(333, seed=2, build_method="kd")
    cases.append((odd, make_ray_batch(g, odd, 8000, 4)))
    rng = np.random.RandomState(5)
    n = 3000
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    lo, hi = np.array([0.25, 0.05, -1.3]), np.array([1.05, 0.7, -0.5])
    rays["origin"] = (lo + rng.rand(n, 3) * (hi - lo)).astype(np.float32)
    d = rng.randn(n, 3); d[:, 1] = -np.abs(d[:, 1]) - 0.2
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    cases.append((box1_coincident, rays))
    for sc, batch in cases:
        hits, cnt = emul.trace(sc.desc, batch, mode)
        assert emul.topology_applied() and cnt["fast_valid"] == 1
        assert_hits_equal(hits, orc.trace_batch(sc.desc, batch), "topology %d mode %d" % (kind, mode))


@pytest.mark.parametrize("mode", [FAST, LANE])
def test_retopologised_pathtracer_matches_oracle(orc, alt_topology, scenes, box1, mode):
    """A whole PT iteration on the median-split topology equals the oracle's bit for bit (as on the reference topology)."""
    emul = alt_topology
    emul.set_topology(emul.TOPOLOGY_MEDIAN, seed=9)
    W = H = 40
    C = scenes.BOX1_CAMERA
    cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)
    s = _light_sampler(orc, box1).generate(128)
    ref, rc = orc.render_pt(box1.desc, cam, W, H, s, orc.wanghash(3), 16)
    img, ec = emul.render(box1.desc, cam, W, H, s, 0, orc.wanghash(3), 16, mode=mode)
    assert emul.topology_applied()
    assert np.array_equal(img, ref)
    for k in ("closest_rays", "any_rays", "camera_paths", "shaded_vertices"):
        assert ec[k] == rc[k]


def test_quantised_child_boxes_are_conservative(g, emul, scenes, box1):
    """Every child box the layout encodes (4-wide nodes; reference and re-clustered topologies) against random rays: whenever the exact
    slab test with first <= limit accepts, the quantised test accepts too and its entry distance is a lower bound (em_slab_check)."""
    sliver = scenes.make_random_triangles(1500, seed=7, build_method="sah")
    try:
        for kind in (emul.TOPOLOGY_REFERENCE, emul.TOPOLOGY_MEDIAN, emul.TOPOLOGY_RANDOM):
            emul.set_topology(kind, seed=5)
            for sc, seed in ((box1, 2), (sliver, 3)):
                r = emul.slab_check(sc.desc, make_ray_batch(g, sc, 3000, seed), stride=7)
                assert r["tested"] > 50000 and r["exact_accepts"] > 1000
                assert r["violations"] == 0, r
                assert r["conservative_accepts"] >= r["exact_accepts"]
    finally:
        emul.set_topology(emul.TOPOLOGY_REFERENCE)


def _far_and_offset_cases(g, scenes):
    """Scenes and rays at the edge of / outside the domain of the conservative inner-node test (gi_lane.h: wfSlabSetup): a scene far from
    the coordinate origin (the records are relative to the root box's centre, so nothing may depend on the offset), ray origins 40 x (inside
    the domain), 80 x and 5000 x (outside: answered by the reference walk) the scene's half extent away from its centre, aimed at the scene."""
    cases = []
    for off in ((0.0, 0.0, 0.0), (30000.0, -20000.0, 10000.0)):
        sc = scenes.make_random_triangles(600, seed=9, offset=off)
        lo, hi = np.array(sc.info.bbox_min, dtype=np.float64), np.array(sc.info.bbox_max, dtype=np.float64)
        centre, half = 0.5 * (lo + hi), 0.5 * (hi - lo).max()
        rng = np.random.RandomState(4)
        n = 3000
        rays = np.zeros(n, dtype=g.RAY_DTYPE)
        dirs = rng.randn(n, 3); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        dist = np.choose(np.arange(n) % 3, [40.0, 80.0, 5000.0]) * half
        origin = centre + dirs * dist[:, None]
        target = lo + rng.rand(n, 3) * (hi - lo)
        d = target - origin; d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays["origin"] = origin.astype(np.float32); rays["direction"] = d.astype(np.float32)
        rays["t_max"] = np.float32(g.RAY_MAX)
        rays["flags"][: n // 4] = 1
        rays["t_max"][: n // 4] = (dist[: n // 4] * (0.9 + 0.2 * rng.rand(n // 4))).astype(np.float32)
        cases.append((sc, rays))
    return cases


def test_far_origins_and_offset_scenes_stay_exact(g, orc, emul, scenes):
    """The lane walk at and beyond the domain of its conservative node test, and on a scene 30 000 units away from the origin."""
    for sc, rays in _far_and_offset_cases(g, scenes):
        ref = orc.trace_batch(sc.desc, rays)
        assert (ref["leaf_tri"][rays["flags"] == 0] >= 0).mean() > 0.3   # the rays do reach the scene
        for mode in (FAST, LANE):
            hits, cnt = emul.trace(sc.desc, rays, mode)
            assert_hits_equal(hits, ref, "far / offset mode %d" % mode)
        # the offset scene's node test is as tight as the centred one's: relative coordinates (same triangles, same rays up to the shift)
        r = emul.slab_check(sc.desc, make_ray_batch(g, sc, 2000, 3, axis_aligned=False), stride=5)
        assert r["violations"] == 0 and r["out_of_domain"] == 0, r
        assert r["conservative_accepts"] <= 1.05 * r["exact_accepts"] + 10, r


@pytest.mark.parametrize("variant", ["bin2", "sort3", "grid16"])
def test_build_switches_of_the_wavefront_walk(variant):
    """The measured-and-switched-off alternatives of the wavefront walk stay exact: binary inner nodes (WF_WIDE4=0), three sort exchanges
    (WF_SORT_EXCHANGES=3), 16-bit grid records (WF_CENTRE_EXTENT=0) -- tests/emul/libgi_emul_<variant>.so passes the same traversal parity tests;
    run in a child process because the emulator library is chosen at import."""
    import subprocess
    import sys
    env = dict(os.environ, GI_EMUL_VARIANT=variant)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_emul_parity.py"), "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "trace_parity or retopologised or exact_ties or conservative or small_trees"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("ntri", [2, 3, 4, 5, 7, 9])
def test_small_trees_all_layouts(g, orc, alt_topology, scenes, ntri):
    """1 .. 5 leaves: wide nodes with two or three children and unused slots, single-leaf scenes without any inner node; on the
    reference's topology and on the two alternative ones."""
    emul = alt_topology
    sc = scenes.make_random_triangles(ntri, seed=ntri, build_method="sah")
    rays = make_ray_batch(g, sc, 4000, 11)
    ref = orc.trace_batch(sc.desc, rays)
    for kind in (emul.TOPOLOGY_REFERENCE, emul.TOPOLOGY_MEDIAN, emul.TOPOLOGY_RANDOM):
        emul.set_topology(kind, seed=2)
        for mode in (FAST, LANE):
            hits, cnt = emul.trace(sc.desc, rays, mode)
            assert_hits_equal(hits, ref, "%d triangles topology %d mode %d" % (ntri, kind, mode))


def test_invalid_leaf_trees_are_rejected(g, orc, alt_topology, scenes):
    """applyLeafTree must refuse anything that is not a binary tree over exactly the reference's leaves (the device clustering is checked, not
    trusted): a leaf reached twice, a cycle, a leaf list that is not a permutation, an out-of-range link.  The layout then keeps the
    reference's topology and the walk stays exact."""
    emul = alt_topology
    sc = scenes.make_random_triangles(40, seed=4, build_method="sah")
    nl = sc.info.num_leaf_triangles // 2
    assert nl >= 8
    N = 2 * nl - 1
    rays = make_ray_batch(g, sc, 3000, 5)
    ref = orc.trace_batch(sc.desc, rays)

    def chain():
        # a valid (list-like) tree: inner node nl + k joins leaf k + 1 with the subtree so far
        left = np.zeros(N, np.uint32); right = np.zeros(N, np.uint32)
        left[nl], right[nl] = 0, 1
        for k in range(1, nl - 1):
            left[nl + k], right[nl + k] = nl + k - 1, k + 1
        return left, right, np.arange(nl, dtype=np.uint32), N - 1

    left, right, leaf_of, root = chain()
    emul.set_tree(left, right, leaf_of, root)
    hits, _ = emul.trace(sc.desc, rays, LANE)
    assert emul.topology_applied()
    assert_hits_equal(hits, ref, "valid chain")

    broken = []
    l, r, lo, rt = chain(); r[nl + 3] = 0; broken.append(("leaf reached twice", l, r, lo, rt))
    l, r, lo, rt = chain(); l[nl + 2] = nl + 5; broken.append(("cycle", l, r, lo, rt))
    l, r, lo, rt = chain(); lo[1] = lo[0]; broken.append(("not a permutation", l, r, lo, rt))
    l, r, lo, rt = chain(); r[nl + 1] = N + 7; broken.append(("link out of range", l, r, lo, rt))
    l, r, lo, rt = chain(); broken.append(("root is a leaf", l, r, lo, 0))
    for name, l, r, lo, rt in broken:
        emul.set_tree(l, r, lo, rt)
        hits, cnt = emul.trace(sc.desc, rays, LANE)
        assert not emul.topology_applied(), name
        assert cnt["fast_valid"] == 1
        assert_hits_equal(hits, ref, name)


def test_axis_aligned_rays_stay_exact(g, orc, emul, scenes):
    """Rays with one or two zero direction components (+0 and -0: 1 / d = +-inf, and inf * 0 = NaN in the slab test where the origin lies in
    a box plane -- a third of the origins are snapped to vertex coordinates) through every walk of the product: ordered one-thread walk,
    reference-order walk and the lane state machine of k_wf_trace (which hands such rays to its exact re-trace), on axis-aligned geometry,
    coincident faces, a random soup and an oriented-box hierarchy -- hit flag, triangle, distance and barycentrics bit-exact vs the oracle."""
    rng = np.random.RandomState(3)

    def batch(sc, n=3000):
        lo, hi = np.array(sc.info.bbox_min[:]), np.array(sc.info.bbox_max[:])
        pos = sc.positions()
        o = lo + (hi - lo) * rng.rand(n, 3)
        o = np.where(rng.rand(n, 3) < 0.3, pos[rng.randint(0, pos.shape[0], n)], o)
        d = np.zeros((n, 3), np.float32)
        for i in range(n):
            kind = i % 3
            if kind == 0:
                d[i, rng.randint(3)] = rng.choice([-1.0, 1.0])
            elif kind == 1:
                d[i, rng.permutation(3)[:2]] = rng.randn(2)
            else:
                v = rng.randn(3)
                v[rng.randint(3)] = rng.choice([0.0, -0.0])
                d[i] = v
            d[i] /= np.linalg.norm(d[i])
        rays = np.zeros(n, dtype=g.RAY_DTYPE)
        rays["origin"], rays["direction"], rays["t_max"] = o.astype(np.float32), d, g.RAY_MAX
        sh = rays[: n // 3].copy()
        sh["flags"] = 1
        sh["t_max"] = (rng.rand(n // 3) * 4).astype(np.float32)
        return np.concatenate([rays, sh])

    cases = [("box1", scenes.make_box1(tess=4)), ("coincident", scenes.make_box1(tess=2, coincident_faces=True)),
             ("soup", scenes.make_random_triangles(3000, seed=1)), ("obox", scenes.make_random_triangles(2000, seed=2, bvh_type=1))]
    for name, sc in cases:
        rays = batch(sc)
        ref = orc.trace_batch(sc.desc, rays)
        assert (ref["leaf_tri"] >= 0).sum() > 1000
        for mode in (FAST, REFERENCE, LANE):
            hits, _ = emul.trace(sc.desc, rays, mode)
            assert_hits_equal(hits, ref, "%s mode %d" % (name, mode))


def test_non_finite_and_degenerate_rays_stay_exact(g, orc, emul, box1):
    """Rays no renderer produces but a gi_trace_rays caller can: NaN / +-inf / denormal / 3.4e38 components in origin or direction, the zero
    direction, NaN / 0 / negative / infinite lengths, closest- and any-hit.  Whatever traceray.glsl makes of them (NaN comparisons are false,
    min / max drop NaN) every walk of the product makes too -- same hit flag, triangle and distance bits -- and every walk terminates."""
    vals = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, 1e38, -1e38, 1.0, -1.0, 3.4e38]
    rng = np.random.RandomState(1)
    n = 4000
    o, d = rng.randn(n, 3).astype(np.float32), rng.randn(n, 3).astype(np.float32)
    for i in range(n):
        for _ in range(rng.randint(1, 3)):
            (o if i % 2 == 0 else d)[i, rng.randint(3)] = rng.choice(vals)
        if i % 50 == 0:
            d[i] = 0
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    rays["origin"], rays["direction"] = o, d
    tm = np.full(n, g.RAY_MAX, np.float32)
    tm[::7] = rng.choice(vals, size=len(tm[::7]))
    rays["t_max"] = tm
    rays["flags"][::3] = 1
    with np.errstate(all="ignore"):
        ref = orc.trace_batch(box1.desc, rays)
        assert (ref["leaf_tri"] >= 0).sum() > 500
        for mode in (FAST, REFERENCE, LANE):
            hits, _ = emul.trace(box1.desc, rays, mode)
            assert np.array_equal(hits["leaf_tri"], ref["leaf_tri"]), mode
            assert np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32)), mode
