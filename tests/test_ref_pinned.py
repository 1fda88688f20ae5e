"""SURVEY 8 row a-15 and the epsilon intersection routines, pinned against the REFERENCE'S OWN object code.

oracle/_ref/libref_bim.so = dependencies/bim/src/*.cpp + dependencies/epsilon/src/*.cpp compiled where they lie (oracle/Makefile;
oracle/ref_shims/bim_prelude.hpp lists the MSVC-isms resolved from outside).  Golden files written by it (tests/golden/make_golden_bim.py):
  ref_box1_t2.json/.bim   a scene file pair stored by BinaryModel::storeEnvironmentFile / storeBinaryHeader / storeChunk
  ref_bim_arrays.npz      hierarchy / parents / leaves / AABoxes / tree levels of buildHierarchy(SAH) and (KD_TREE) + computeBVHAABoxes
  ref_ei_kat.npz          ei::intersects(Ray, Box) and (Ray, Triangle) on seeded random cases
The golden tests run everywhere (CPU suite, GPU box); the `live` tests additionally push fresh inputs through the library where it exists.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
live = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_bim.so not built here (needs /root/reference)")


def _same_hierarchy(scene, ref, prefix):
    n = scene.nodes()
    lt = scene.leaf_triangles()
    assert len(n) == len(ref[prefix + "hierarchy"])
    assert np.array_equal(n["first_child"], ref[prefix + "hierarchy"]["first_child"])
    assert np.array_equal(n["escape"], ref[prefix + "hierarchy"]["escape"])
    # AABoxes bit for bit
    assert np.array_equal(n["min"].view(np.uint32), ref[prefix + "aaboxes"][:, :3].view(np.uint32))
    assert np.array_equal(n["max"].view(np.uint32), ref[prefix + "aaboxes"][:, 3:].view(np.uint32))
    assert np.array_equal(scene.parents(), ref[prefix + "parents"])
    assert np.array_equal(lt["v"], ref[prefix + "leaves"][:, :3]) and np.array_equal(lt["material"], ref[prefix + "leaves"][:, 3])
    assert scene.info.num_tree_levels == int(ref[prefix + "tree_levels"])


def test_host_reads_reference_written_bim(g):
    """host/bim.cpp (Scene::Scene -> BinaryModel::load + makeChunkResident) on a file pair written by the reference's writer."""
    z = np.load(os.path.join(GOLDEN, "ref_bim_arrays.npz"))
    sc = g.Scene.load(os.path.join(GOLDEN, "ref_box1_t2.json"))
    assert np.array_equal(sc.positions().view(np.uint32), z["box1_in_positions"].view(np.uint32))
    assert np.array_equal(sc.triangles(), z["box1_in_triangles"])
    assert np.array_equal(sc.triangle_materials(), z["box1_in_triangle_materials"])
    # vertex info = (atan2(n.y, n.x), n.z, u, v) of the stored normals / texcoords (scene.cpp:81-83)
    vi = sc.vertex_info()
    nrm, uv = z["box1_in_normals"], z["box1_in_texcoords"]
    assert np.array_equal(vi[:, 1].view(np.uint32), nrm[:, 2].view(np.uint32)) and np.array_equal(vi[:, 2:].view(np.uint32), uv.view(np.uint32))
    assert np.allclose(vi[:, 0], np.arctan2(nrm[:, 1], nrm[:, 0]), atol=1e-6)
    _same_hierarchy(sc, z, "box1_sah_")
    assert sc.info.num_materials == 9
    lo, hi = z["box1_sah_bounding_box"][:3], z["box1_sah_bounding_box"][3:]
    assert np.array_equal(np.array(sc.info.bbox_min, dtype=np.float32), lo) and np.array_equal(np.array(sc.info.bbox_max, dtype=np.float32), hi)


def test_reference_json_is_what_the_host_expects():
    """The environment file as the reference's nlohmann writer emits it: materials keyed by name, "scene" = bim file, "accelerator"."""
    env = json.load(open(os.path.join(GOLDEN, "ref_box1_t2.json")))
    assert env["scene"] == "ref_box1_t2.bim" and env["accelerator"] == "aabox"
    assert set(env["materials"]) == {"DefaultMaterial", "Glass", "Floor", "RightWall", "Wall", "LeftWall", "White", "LightR", "LightL"}
    assert env["materials"]["LightR"]["emissivity"] == pytest.approx([803.9, 192.2, 19.6])


@pytest.mark.parametrize("method", ["sah", "kd"])
@pytest.mark.parametrize("mesh", ["box1", "soup"])
def test_host_builders_equal_reference_builders_golden(g, scenes, mesh, method):
    """Chunk::buildHierarchy(SAH / KD_TREE) + remapNodePointers + computeBVHAABoxes of host/bim.cpp vs the reference's, on the committed arrays."""
    z = np.load(os.path.join(GOLDEN, "ref_bim_arrays.npz"))
    inp = [z["%s_in_%s" % (mesh, k)] for k in ("positions", "normals", "texcoords", "triangles", "triangle_materials")]
    mats = scenes.box1_materials() if mesh == "box1" else {"a": scenes.legacy_material((0.7, 0.7, 0.7)), "b": scenes.legacy_material((0.3, 0.5, 0.7))}
    sc = g.Scene.from_mesh(*inp, mats, "", method)
    _same_hierarchy(sc, z, "%s_%s_" % (mesh, method))


def test_oracle_box_test_vs_epsilon_golden(orc):
    """IntersectBox (intersectiontests.glsl:9-20) as restated in the oracle vs ei::intersects(Ray, Box) -- whose commented-out body
    (3dintersection.cpp:343-348) is the GLSL text.  Hit flags must agree on every case that is not within rounding of a box edge;
    distances to float rounding (the GLSL multiplies by a reciprocal where epsilon divides)."""
    z = np.load(os.path.join(GOLDEN, "ref_ei_kat.npz"))
    out = z["box_out"]
    n = len(out)
    disagree = 0
    for i in range(n):
        hit, first, last = orc.intersect_box(z["box_o"][i], z["box_d"][i], z["box_lo"][i], z["box_hi"][i])
        if hit != bool(out["hit"][i]):
            disagree += 1
            assert abs(first - last) <= 1e-5 * max(1.0, abs(last)), "case %d: a clear hit / miss differs from epsilon" % i
            continue
        if hit:
            assert abs(first - out["t0"][i]) <= 4e-7 * max(1.0, abs(out["t0"][i])) + 1e-7
            assert abs(last - out["t1"][i]) <= 4e-7 * max(1.0, abs(out["t1"][i])) + 1e-7
    assert disagree <= n // 500
    assert out["hit"].sum() > 200  # the batch exercises both outcomes
    # epsilon's (distance, distanceExit) overload lacks the final `tmax >= 0` test: boxes entirely behind the origin along z are "hit" with a
    # negative exit.  The shader (and the oracle) follow the plain overload.
    behind = (out["hit_distance_overload"] == 1) & (out["hit"] == 0)
    assert np.all(out["t1"][behind] < 0.0)


def test_oracle_triangle_test_vs_epsilon_golden(orc):
    """IntersectTriangle (intersectiontests.glsl:63-79) vs ei::intersects(Ray, Triangle, distance, barycentric)
    (3dintersection.cpp: same edge / normal algebra, division instead of reciprocal multiply, no 1e-4 distance epsilon)."""
    z = np.load(os.path.join(GOLDEN, "ref_ei_kat.npz"))
    out = z["tri_out"]
    n = len(out)
    disagree = 0
    for i in range(n):
        hit, t, bary, _ = orc.intersect_triangle(z["tri_o"][i], z["tri_d"][i], *z["tri_p"][i])
        if hit != bool(out["hit"][i]):
            disagree += 1
            near_edge = np.abs(out["bary"][i]).min() < 1e-4 or (out["hit"][i] and out["t"][i] <= 2e-4)
            assert near_edge, "case %d: a clear hit / miss differs from epsilon" % i
            continue
        if hit:
            assert abs(t - out["t"][i]) <= 4e-6 * max(1.0, abs(out["t"][i]))
            assert np.abs(bary - out["bary"][i]).max() <= 2e-5
    assert disagree <= n // 500
    assert out["hit"].sum() > 300


# ------------------------------------------------------------------------------------------------- live: fresh inputs through the reference's code
def _soup(n, seed):
    rng = np.random.RandomState(seed)
    c = rng.uniform(-10, 10, size=(n, 1, 3))
    p = (c + rng.uniform(-1, 1, size=(n, 3, 3))).astype(np.float32).reshape(-1, 3)
    nrm = np.tile(np.array([0, 1, 0], dtype=np.float32), (3 * n, 1))
    uv = rng.rand(3 * n, 2).astype(np.float32)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mat = (np.arange(n) % 2).astype(np.uint32)
    return p, nrm, uv, tri, mat


@live
@pytest.mark.parametrize("method", ["sah", "kd"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 64, 3000])
def test_host_builders_equal_reference_builders_live(g, scenes, n, method):
    mesh = _soup(n, 100 + n)
    ref = pyref.RefModel.build(*mesh, ["a", "b"], method).arrays()
    mats = {"a": scenes.legacy_material((0.7, 0.7, 0.7)), "b": scenes.legacy_material((0.3, 0.5, 0.7))}
    sc = g.Scene.from_mesh(*mesh, mats, "", method)
    _same_hierarchy(sc, {"x_" + k: v for k, v in ref.items()}, "x_")


@live
def test_sponza_shaped_scene_builder_equals_reference(g, scenes, sponza_small):
    """The bench scene's generator at reduced size (textured, 30 k triangles) through both SAH builders."""
    sc = sponza_small
    pos, tri, mat = sc.positions(), sc.triangles(), sc.triangle_materials()
    ref = pyref.RefModel.build(pos, None, None, tri, mat, ["m%d" % i for i in range(sc.info.num_materials)], "sah").arrays()
    _same_hierarchy(sc, {"x_" + k: v for k, v in ref.items()}, "x_")


@live
def test_reference_reads_host_written_bim(g, scenes, tmp_path):
    """BinaryModel::load of the reference on a file pair written by host/bim.cpp: every array it would upload is identical."""
    sc = scenes.make_box1(tess=2)
    sc.store(str(tmp_path / "host.json"), "host.bim")
    got = pyref.RefModel.load(str(tmp_path / "host.json"))
    z = np.load(os.path.join(GOLDEN, "ref_bim_arrays.npz"))
    a = got.arrays()
    for k in ("hierarchy", "parents", "leaves", "aaboxes", "bounding_box"):
        assert np.array_equal(a[k], z["box1_sah_" + k]), k
    for k in ("positions", "normals", "texcoords", "triangles", "triangle_materials"):
        assert np.array_equal(a[k], z["box1_in_" + k]), k
    assert got.material_names() == scenes.BOX1_MATERIAL_NAMES
    assert got.counts()["accelerator"] == 0x01000000 and int(a["tree_levels"]) == int(z["box1_sah_tree_levels"])


@live
def test_round_trip_through_both_writers_and_readers(g, scenes, tmp_path):
    """mesh -> reference build + store -> host load -> host store -> reference load: nothing changes on the way."""
    mesh = _soup(200, 9)
    env_in = tmp_path / "in.json"
    mats = {"a": scenes.legacy_material((0.7, 0.7, 0.7)), "b": scenes.legacy_material((0.3, 0.5, 0.7), (0.04, 0.04, 0.04), 0.3)}
    json.dump({"scene": "", "accelerator": "aabox", "materials": mats, "lights": {}, "cameras": {}}, open(env_in, "w"))
    ref = pyref.RefModel.build(*mesh, ["a", "b"], "sah", env_json=str(env_in))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.store("r.json", "r.bim")
    finally:
        os.chdir(cwd)
    sc = g.Scene.load(str(tmp_path / "r.json"))
    _same_hierarchy(sc, {"x_" + k: v for k, v in ref.arrays().items()}, "x_")
    sc.store(str(tmp_path / "h.json"), "h.bim")
    back = pyref.RefModel.load(str(tmp_path / "h.json")).arrays()
    for k, v in ref.arrays().items():
        assert np.array_equal(back[k], v), k


@live
def test_oracle_vs_epsilon_live(orc):
    rng = np.random.RandomState(2024)
    n = 3000
    o = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    d = rng.randn(n, 3)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    lo = rng.uniform(-3, 2, size=(n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0.05, 3, size=(n, 3))).astype(np.float32)
    clear_disagreements = 0
    for i in range(n):
        eh, t0, t1, _ = pyref.ei_ray_box(o[i], d[i], lo[i], hi[i])
        hit, first, last = orc.intersect_box(o[i], d[i], lo[i], hi[i])
        if hit != eh and abs(first - last) > 1e-5 * max(1.0, abs(last)):
            clear_disagreements += 1
    assert clear_disagreements == 0


def test_camera_params_vs_epsilon_golden(g, orc):
    """Camera::ComputeCameraParams (camera.cpp:76-85): host Camera and the oracle's camera vs the formula evaluated with epsilon's own
    normalize / cross / operator*= (ref_ei_camera_params), bit for bit on 512 seeded cameras."""
    z = np.load(os.path.join(GOLDEN, "ref_ei_kat.npz"))
    for row, want in zip(z["cam_in"], z["cam_out"]):
        w, h = (1920, 1080) if abs(row[7] - 16.0 / 9.0) < 1e-6 else ((1024, 768) if abs(row[7] - 4.0 / 3.0) < 1e-6 else (512, 512))
        got = np.frombuffer(g.camera_params(row[:3], row[3:6], float(row[6]), w, h), dtype=np.float32)
        assert np.array_equal(got[3:12].view(np.uint32), want.view(np.uint32))
        assert np.array_equal(orc.camera(row[:3], row[3:6], float(row[6]), w, h)[3:12].view(np.uint32), want.view(np.uint32))


def test_light_sample_point_vs_epsilon_golden(g, orc, scenes):
    """Position / normal of an area-light sample (lightsampler.cpp:55-69) vs epsilon's vector arithmetic: a scene whose only emitter is
    ONE triangle makes every sample fall on it, and the XorshiftF stream (pinned by tests/golden/ref_rng.npz) gives the barycentrics."""
    import ctypes as C
    from gpugi_b200 import _capi as capi
    z = np.load(os.path.join(GOLDEN, "ref_ei_kat.npz"))
    for k in range(0, 64):
        tri = z["ls_tri"][k]
        pos = np.concatenate([tri, tri + np.float32(20.0)]).astype(np.float32)
        nrm = np.tile(np.array([0, 1, 0], dtype=np.float32), (6, 1))
        uv = np.zeros((6, 2), dtype=np.float32)
        mats = {"light": scenes.legacy_material((0.5, 0.5, 0.5), emissivity=(3.0, 2.0, 1.0)), "dark": scenes.legacy_material((0.5, 0.5, 0.5))}
        sc = g.Scene.from_mesh(pos, nrm, uv, np.array([[0, 1, 2], [3, 4, 5]], dtype=np.uint32), np.array([0, 1], dtype=np.uint32), mats, "")
        ls = g.LightSampler(sc)
        ls.seed(1000 + k)
        s = ls.generate(4)
        state = C.c_uint64(1000 + k)
        for i in range(4):
            capi.host.gih_xorshift_f(C.byref(state))  # triangle pick
            xi0 = np.float32(capi.host.gih_xorshift_f(C.byref(state)))
            xi1 = np.float32(capi.host.gih_xorshift_f(C.byref(state)))
            x0 = np.sqrt(xi0)
            a, b, c = np.float32(1.0) - x0, x0 * (np.float32(1.0) - xi1), x0 * xi1
            if pyref.available():
                p, n = pyref.ei_light_sample_point(tri[0], tri[1], tri[2], a, b, c, 1.0e-4)
                assert np.array_equal(s["position"][i].view(np.uint32), p.view(np.uint32))
                assert s["normal_theta_cos"][i] == n[2] and abs(s["normal_phi"][i] - np.arctan2(n[1], n[0])) <= 2.4e-7
    # the committed vectors: same formula, barycentrics from the file
    if pyref.available():
        for k in range(64, 128):
            xi = z["ls_xi"][k]
            x0 = np.sqrt(xi[0])
            p, n = pyref.ei_light_sample_point(*z["ls_tri"][k], np.float32(1.0) - x0, x0 * (np.float32(1.0) - xi[1]), x0 * xi[1], 1.0e-4)
            assert np.array_equal(np.concatenate([p, n]).view(np.uint32), z["ls_out"][k].view(np.uint32))
