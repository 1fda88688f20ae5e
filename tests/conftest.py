import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def g():
    import gpugi_b200
    return gpugi_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def emul():
    from tests.emul import pyemul
    return pyemul


@pytest.fixture(scope="session")
def scenes():
    from gpugi_b200 import scenes
    return scenes


@pytest.fixture(scope="session")
def box1(scenes):
    return scenes.make_box1(tess=4)


@pytest.fixture(scope="session")
def box1_coincident(scenes):
    return scenes.make_box1(tess=2, coincident_faces=True)


@pytest.fixture(scope="session")
def random_tris(scenes):
    return scenes.make_random_triangles(3000, seed=1)


@pytest.fixture(scope="session")
def sponza_small(scenes, tmp_path_factory):
    d = tmp_path_factory.mktemp("sponza_tex")
    return scenes.make_sponza_like(str(d), target_triangles=30000, texture_size=64, seed=7, detail=0.3)


def make_ray_batch(g, scene, n, seed, any_fraction=0.25, axis_aligned=True):
    """Mixed batch: closest-hit rays with RAY_MAX, any-hit rays with finite length, some axis-aligned rays."""
    rng = np.random.RandomState(seed)
    info = scene.info
    lo, hi = np.array(info.bbox_min), np.array(info.bbox_max)
    ext = hi - lo
    rays = np.zeros(n, dtype=g.RAY_DTYPE)
    rays["origin"] = (lo - 0.1 * ext + rng.rand(n, 3) * 1.2 * ext).astype(np.float32)
    d = rng.randn(n, 3)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["t_max"] = np.float32(g.RAY_MAX)
    k = int(n * any_fraction)
    rays["flags"][:k] = 1
    rays["t_max"][:k] = (rng.rand(k) * np.linalg.norm(ext)).astype(np.float32)
    if axis_aligned and n >= 1000:
        m = min(64, (n - k) // 8)
        axes = [(0, 0, 1), (1, 0, 0), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (0, 0, -1)]
        for j, a in enumerate(axes):
            rays["direction"][k + j * m: k + (j + 1) * m] = np.array(a, dtype=np.float32)
    return rays


def camera_ray_batch(g, orc, scene, campos, lookat, fov, w, h, seed):
    """Primary rays (jittered) + secondary rays from their hit points + shadow rays: the SURVEY 8d mix."""
    rng = np.random.RandomState(seed)
    cam = orc.camera(campos, lookat, fov, w, h)
    pos, U, V, Wv = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    ys, xs = np.mgrid[0:h, 0:w]
    sx = (xs + rng.rand(h, w)) / w * 2 - 1
    sy = (ys + rng.rand(h, w)) / h * 2 - 1
    d = sx[..., None] * U + sy[..., None] * V + Wv
    d = (d / np.linalg.norm(d, axis=2, keepdims=True)).reshape(-1, 3)
    prim = np.zeros(w * h, dtype=g.RAY_DTYPE)
    prim["origin"] = pos
    prim["direction"] = d.astype(np.float32)
    prim["t_max"] = np.float32(g.RAY_MAX)
    hits = orc.trace_batch(scene.desc, prim)
    ok = hits["leaf_tri"] >= 0
    hp = prim["origin"][ok] + hits["t"][ok, None] * prim["direction"][ok]
    d2 = rng.randn(hp.shape[0], 3)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    sec = np.zeros(hp.shape[0], dtype=g.RAY_DTYPE)
    sec["origin"] = (hp + 1e-4 * d2).astype(np.float32)
    sec["direction"] = d2.astype(np.float32)
    sec["t_max"] = np.float32(g.RAY_MAX)
    sh = sec.copy()
    sh["flags"] = 1
    sh["t_max"] = (rng.rand(hp.shape[0]) * 3).astype(np.float32)
    return np.concatenate([prim, sec, sh])


def assert_hits_equal(a, b, what=""):
    """Bit-exact: hit flag / triangle id, distance and barycentrics."""
    bad = np.nonzero(a["leaf_tri"] != b["leaf_tri"])[0]
    assert bad.size == 0, "%s: %d rays differ in hit flag / triangle id (first: %s)" % (what, bad.size, bad[:5])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32)), what + ": distances differ"
    assert np.array_equal(a["bary_y"].view(np.uint32), b["bary_y"].view(np.uint32)), what + ": barycentric y differs"
    assert np.array_equal(a["bary_z"].view(np.uint32), b["bary_z"].view(np.uint32)), what + ": barycentric z differs"


def _log_parity(kind, value, **extra):
    """GPUGI_B200_PARITY_LOG=<file>: every measured parity statistic is appended as one JSON line (what the tolerances are set from)."""
    path = os.environ.get("GPUGI_B200_PARITY_LOG")
    if path:
        import json
        with open(path, "a") as f:
            f.write(json.dumps(dict(test=os.environ.get("PYTEST_CURRENT_TEST", ""), kind=kind, value=value, **extra)) + "\n")


@pytest.fixture(scope="session", autouse=True)
def _parity_log_rel_mse():
    """With GPUGI_B200_PARITY_LOG set, rel_mse values the tests compute are logged too."""
    if not os.environ.get("GPUGI_B200_PARITY_LOG"):
        yield
        return
    from gpugi_b200 import parallel
    plain = parallel.rel_mse

    def logged(*a, **k):
        v = plain(*a, **k)
        _log_parity("rel_mse", float(v))
        return v
    parallel.rel_mse = logged
    yield
    parallel.rel_mse = plain


def pixel_mismatch_fraction(img, ref, rel_tol=1e-2):
    rel = np.abs(img[..., :3] - ref[..., :3]) / (np.abs(ref[..., :3]) + 1e-3)
    value = float((rel.max(axis=2) > rel_tol).mean())
    _log_parity("pixel_mismatch_fraction", value, rel_tol=rel_tol)
    return value


def robust_block_compare(a, b, block=8, clip=20.0):
    """Two independent noisy estimates of the same image (BPT: the light-cache slot order, hence every connection, is scheduling
    dependent).  Fireflies make plain (rel)MSE a coin toss at a few iterations, so: clip both at `clip` x the mean, average over
    block x block pixels, and return (relative difference of the clipped means, median and 90th percentile of the per-block relative
    differences)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)[..., :3]
    b = np.asarray(b, dtype=np.float64)[..., :3]
    cap = clip * max(b.mean(), 1e-12)
    a, b = np.minimum(a, cap), np.minimum(b, cap)
    h, w = (a.shape[0] // block) * block, (a.shape[1] // block) * block
    ab = a[:h, :w].reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3, 4))
    bb = b[:h, :w].reshape(h // block, block, w // block, block, 3).mean(axis=(1, 3, 4))
    rel = np.abs(ab - bb) / (bb + 0.05 * bb.mean())
    return abs(a.mean() - b.mean()) / b.mean(), float(np.median(rel)), float(np.percentile(rel, 90))
