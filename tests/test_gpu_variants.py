"""SURVEY 8 row f-4, the shader #define variants the reference's build leaves off: RECONSTRUCTION_FILTER 1-3 (screenreproject.glsl:14-17,186-223)
and CHAINED_MIS (bidirpathtracer/chained_mis.glsl + its hooks in lighttracer.comp, screenreproject.glsl, pathtracer_bidir.comp).  GPU vs oracle."""
import numpy as np
import pytest

from tests.conftest import pixel_mismatch_fraction

pytestmark = pytest.mark.gpu


def _device(g, scene, W, H, cam):
    dev = g.Device(0)
    dev.upload_scene(scene)
    dev.set_screen_size(W, H)
    dev.set_camera(g.camera_params(cam["pos"], cam["lookat"], cam["fov"], W, H))
    dev.set_option(g.GI_OPT_TEXTURE_FILTER, 1)
    return dev


@pytest.mark.parametrize("flt", [1, 2, 3])
def test_reconstruction_filters_match_oracle(g, orc, scenes, box1, flt):
    """Every camera-connection splat of the light tracer becomes 5 x 5 weighted splats: waves == one thread per path == oracle."""
    W, H = 120, 72
    cam = scenes.BOX1_CAMERA
    dev = _device(g, box1, W, H, cam)
    dev.set_option(g.GI_OPT_RECONSTRUCTION_FILTER, flt)
    nr = 64
    samples = g.LightSampler(box1).generate(256)
    dev.set_light_samples(samples)
    seed = g.wang_hash(2)
    out = {}
    for sched in (g.GI_SCHEDULER_MEGAKERNEL, g.GI_SCHEDULER_WAVEFRONT):
        dev.set_option(g.GI_OPT_SCHEDULER, sched)
        dev.clear(); dev.reset_counters()
        dev.render_iteration(g.GI_RENDERER_LPT, seed, 16, num_rays_per_light_sample=nr)
        out[sched] = dev.read_accumulator(W, H)
    # and a streamed run (splat ring) of three iterations against three plain ones
    dev.set_option(g.GI_OPT_SCHEDULER, g.GI_SCHEDULER_WAVEFRONT)
    dev.clear()
    for k in range(3):
        dev.render_iteration(g.GI_RENDERER_LPT, g.wang_hash(k), 16, num_rays_per_light_sample=nr)
    plain3 = dev.read_accumulator(W, H)
    dev.set_option(g.GI_OPT_STREAMING, 1)
    dev.clear()
    for k in range(3):
        dev.render_iteration(g.GI_RENDERER_LPT, g.wang_hash(k), 16, num_rays_per_light_sample=nr)
    stream3 = dev.read_accumulator(W, H)
    dev.close()
    camp = np.frombuffer(g.camera_params(cam["pos"], cam["lookat"], cam["fov"], W, H), dtype=np.float32)
    orc.set_variant(flt, False)
    try:
        ref, _ = orc.render_lpt(box1.desc, camp, W, H, samples, seed, nr, 16, show_light_caches=False)
        orc.set_variant(0, False)
        box, _ = orc.render_lpt(box1.desc, camp, W, H, samples, seed, nr, 16, show_light_caches=False)
    finally:
        orc.set_variant(0, False)
    a, b = out[g.GI_SCHEDULER_MEGAKERNEL], out[g.GI_SCHEDULER_WAVEFRONT]
    # (the Mitchell kernel has negative lobes: pixels where contributions cancel are compared on the scale of the image, not of the pixel)
    assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max() + 1e-6
    assert np.abs(plain3 - stream3).max() <= 1e-5 * np.abs(plain3).max() + 1e-6
    e, er = float(b[..., :3].sum()), float(ref[..., :3].sum())
    assert er > 0 and abs(e - er) <= 5e-3 * er
    assert pixel_mismatch_fraction(b, ref, 1e-2) < 0.005
    # the filter does something: the image is not the box-filtered one, and it is smoother (more pixels receive light)
    assert not np.allclose(ref, box, rtol=1e-3, atol=1e-6)
    assert (ref[..., :3].sum(axis=2) != 0).sum() > (box[..., :3].sum(axis=2) != 0).sum()


def test_filter_option_is_validated(g, scenes, box1):
    dev = _device(g, box1, 32, 32, scenes.BOX1_CAMERA)
    with pytest.raises(g.GpugiError):
        dev.set_option(g.GI_OPT_RECONSTRUCTION_FILTER, 4)
    dev.close()


def test_chained_mis_matches_oracle_on_the_same_cache(g, orc, scenes, box1):
    """CHAINED_MIS: the light pass stores two quality weights per light vertex and weights its camera connections with them, the eye pass chains
    the per-segment MIS weights (g_mis).  As for the plain BPT, the eye pass is deterministic given the cache: GPU vs oracle on the GPU's cache."""
    W = H = 96
    cam = scenes.BOX1_CAMERA
    nls = 128
    nr = max(64, (W * H // nls // 64) * 64)
    samples = g.LightSampler(box1).generate(nls)
    camp = np.frombuffer(g.camera_params(cam["pos"], cam["lookat"], cam["fov"], W, H), dtype=np.float32)
    imgs = {}
    for chained in (0, 1):
        dev = _device(g, box1, W, H, cam)
        dev.set_option(g.GI_OPT_CHAINED_MIS, chained)
        dev.set_light_samples(samples)
        cap = nls * nr * 4
        dev.alloc_light_cache(cap)
        seed = g.wang_hash(4)
        launches = dev.kernel_launches()
        dev.render_iteration(g.GI_RENDERER_BPT, seed, 16, num_rays_per_light_sample=nr, light_cache_capacity=cap)
        if chained:
            assert dev.kernel_launches() - launches == 3   # resetcounter + light pass + eye pass: one thread per path
        img = dev.read_accumulator(W, H)
        cache, entries = dev.get_light_cache(cap)
        dev.close()
        orc.set_variant(0, bool(chained))
        try:
            ref, _, (ref_entries, _) = orc.render_bpt_with_cache(box1.desc, camp, W, H, samples, seed, nr, cap, cache, entries, 16)
            _, _, _, own = orc.render_bpt(box1.desc, camp, W, H, samples, seed, nr, cap, 16, return_cache=True)
        finally:
            orc.set_variant(0, False)
        assert ref_entries == entries
        bad = pixel_mismatch_fraction(img, ref, 1e-2)
        assert bad < 0.002, "chained=%d: %.2f %% of the pixels differ by more than 1 %%" % (chained, 100 * bad)
        e, er = float(img[..., :3].sum()), float(ref[..., :3].sum())
        assert abs(e - er) <= 3e-3 * er
        # the quality weights in the cache (columns 14, 15): 0 without CHAINED_MIS; with it, (1, 1) at the light origins and the same multiset as the oracle's
        n0 = nls * nr
        if chained:
            assert np.all(cache[:n0, 14:16] == 1.0)
            a, b = np.sort(cache[n0:entries, 14:16], axis=0), np.sort(own[n0:entries, 14:16], axis=0)
            assert (np.abs(a - b) > 3e-3 * np.abs(b) + 1e-6).any(axis=1).mean() < 0.01   # (products over up to 16 segments: rounding accumulates)
            assert cache[n0:entries, 14].min() >= 1.0 and cache[n0:entries, 14:16].max() > 1.0
        else:
            assert not cache[:entries, 14:16].any()
        imgs[chained] = img
    assert not np.allclose(imgs[0], imgs[1], rtol=1e-2, atol=1e-4)   # the variant changes the weights


def test_chained_mis_path_length_limit(g, scenes, box1):
    dev = _device(g, box1, 32, 32, scenes.BOX1_CAMERA)
    dev.set_option(g.GI_OPT_CHAINED_MIS, 1)
    dev.set_light_samples(g.LightSampler(box1).generate(128))
    dev.alloc_light_cache(128 * 64 * 4)
    with pytest.raises(g.GpugiError):
        dev.render_iteration(g.GI_RENDERER_BPT, 1, 65, num_rays_per_light_sample=64, light_cache_capacity=128 * 64 * 4)
    dev.close()
