"""SURVEY 8 row f-2, bvhmake's legacy pipeline: the `.rawscene` container (bvhmake/filedef.hpp:11-89) and the kdtree / sweep / lds build methods over
8-triangle leaves (host/bvhmake.cpp), pinned against the REFERENCE'S OWN build methods (oracle/_ref/libref_bvhmake.so = buildmethods/kdtree.cpp,
sweep.cpp + fitmethods/aaboxfit.cpp compiled where they lie; golden arrays in tests/golden/ref_bvhmake_arrays.npz)."""
import os
import struct

import numpy as np
import pytest

from oracle import pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["m%d" % i for i in range(9)]


def _children(nodes):
    """{left, right} of the file's preorder nodes: firstChild, and the node behind firstChild's subtree (= its escape)."""
    out = np.zeros((len(nodes), 2), np.uint32)
    for i in range(len(nodes)):
        fc = nodes["first_child"][i]
        out[i] = (fc, 0) if fc & 0x80000000 else (fc, nodes["escape"][fc])
    return out


def _leaf_array(raw):
    l = raw.leaves()
    return np.concatenate([l["triangles"]["v"], l["triangles"]["material"][..., None]], axis=2)


def _check_same(raw, children, boxes, leaves):
    assert np.array_equal(_children(raw.nodes()), children)
    assert np.array_equal(raw.boxes().view(np.uint32), boxes.view(np.uint32))
    assert np.array_equal(_leaf_array(raw), leaves)


@pytest.mark.parametrize("method", ["kdtree", "sweep"])
@pytest.mark.parametrize("mesh", ["box1", "soup"])
def test_builders_equal_reference_golden(g, mesh, method):
    z = np.load(os.path.join(GOLDEN, "ref_bvhmake_arrays.npz"))
    raw = g.RawScene.build(z[mesh + "_in_positions"], None, None, z[mesh + "_in_triangles"], z[mesh + "_in_materials"], NAMES, method)
    _check_same(raw, z["%s_%s_children" % (mesh, method)], z["%s_%s_boxes" % (mesh, method)], z["%s_%s_leaves" % (mesh, method)])


@pytest.mark.skipif(not pyref.bvhmake_available(), reason="oracle/_ref/libref_bvhmake.so not built here (needs /root/reference)")
@pytest.mark.parametrize("method", ["kdtree", "sweep"])
@pytest.mark.parametrize("n", [64, 100, 333, 2500])
def test_builders_equal_reference_live(g, n, method, capfd):
    # (the reference sizes its node pool as 4 T / 8 and parks temporaries in the pool's last three entries, bvhbuilder.cpp:530, sweep.cpp:219-222:
    # below a few dozen triangles it writes past its own buffers, so the comparison starts at 64)
    rng = np.random.RandomState(40 + n)
    c = rng.uniform(-10, 10, size=(n, 1, 3))
    pos = (c + rng.uniform(-1, 1, size=(n, 3, 3))).astype(np.float32).reshape(-1, 3)
    if n >= 100:
        pos[: 3 * 20] = pos[3 * 20: 3 * 40]   # 20 duplicated triangles: equal centres (the median's tie handling, equal Hilbert keys)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mat = (np.arange(n) % 3).astype(np.uint32)
    children, boxes, leaves = pyref.bvhmake_build(pos, tri, mat, method)
    raw = g.RawScene.build(pos, None, None, tri, mat, NAMES, method)
    _check_same(raw, children, boxes, leaves)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 17])
@pytest.mark.parametrize("method", ["kdtree", "sweep", "lds"])
def test_small_inputs(g, n, method):
    """Fewer triangles than the reference's pools can take: one leaf up to 8 triangles, a valid tree above."""
    rng = np.random.RandomState(n)
    pos = rng.uniform(-1, 1, size=(3 * n, 3)).astype(np.float32)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    raw = g.RawScene.build(pos, None, None, tri, None, ["a"], method)
    leaves = _leaf_array(raw)
    assert (leaves[..., 0] != leaves[..., 1]).sum() == n
    assert len(raw.nodes()) == 2 * len(leaves) - 1 and (len(leaves) == 1) == (n <= 8)


@pytest.mark.parametrize("method", ["kdtree", "sweep", "lds"])
def test_hierarchy_invariants_and_file_round_trip(g, scenes, tmp_path, method):
    """Every build method: preorder layout with parent / escape links, boxes that contain their triangles and children, every triangle in exactly
    one leaf; the file written by StoreRawScene is the documented sequence of named arrays and loads back unchanged."""
    pos, nrm, uv, tri, mat = scenes.box1_mesh(3)
    raw = g.RawScene.build(pos, nrm, uv, tri, mat, scenes.BOX1_MATERIAL_NAMES, method)
    nodes, boxes, leaves = raw.nodes(), raw.boxes(), _leaf_array(raw)
    n = len(nodes)
    assert nodes["parent"][0] == 0 and nodes["escape"][0] == 0
    seen = []
    for i in range(n):
        fc = int(nodes["first_child"][i])
        if fc & 0x80000000:
            leaf = leaves[fc & 0x7FFFFFFF]
            valid = leaf[leaf[:, 0] != leaf[:, 1]]
            assert 1 <= len(valid) <= 8
            p = pos[valid[:, :3].reshape(-1)]
            assert np.all(p.min(0) >= boxes[i, :3]) and np.all(p.max(0) <= boxes[i, 3:])
            seen += [tuple(t) for t in valid]
        else:
            l, r = fc, int(nodes["escape"][fc])
            assert l == i + 1 and nodes["parent"][l] == i and nodes["parent"][r] == i and nodes["escape"][r] == nodes["escape"][i]
            for c in (l, r):
                assert np.all(boxes[c, :3] >= boxes[i, :3]) and np.all(boxes[c, 3:] <= boxes[i, 3:])
    want = sorted(tuple(int(x) for x in list(t) + [m]) for t, m in zip(tri, mat))
    assert sorted(tuple(int(x) for x in t) for t in seen) == want
    path = str(tmp_path / "box1.rawscene")
    raw.store(path)
    blob = open(path, "rb").read()
    off, names = 0, []
    while off < len(blob):
        name, count, size = struct.unpack_from("<32sII", blob, off)
        names.append((name.split(b"\0")[0].decode(), count, size))
        off += 40 + count * size
    assert off == len(blob)
    assert names == [("materialref", 9, 32), ("vertices", len(pos), 32), ("bounding_aabox", n, 24), ("hierarchy", n, 12), ("triangles", len(leaves), 128)]
    back = g.RawScene.load(path)
    assert back.material_names() == scenes.BOX1_MATERIAL_NAMES
    assert np.array_equal(back.nodes(), nodes) and np.array_equal(back.boxes(), boxes) and np.array_equal(_leaf_array(back), leaves)
    v = back.vertices()
    assert np.array_equal(v["position"], pos) and np.array_equal(v["normal"], nrm) and np.array_equal(v["texcoord"], uv)
    # an unknown array (bvhmake also writes approx_sggx) is skipped, a truncated file is refused
    with open(path, "ab") as f:
        f.write(struct.pack("<32sII", b"approx_sggx", n, 12) + bytes(12 * n))
    assert np.array_equal(g.RawScene.load(path).nodes(), nodes)
    open(str(tmp_path / "cut.rawscene"), "wb").write(blob[: len(blob) - 50])
    with pytest.raises(g.GpugiError):
        g.RawScene.load(str(tmp_path / "cut.rawscene"))


@pytest.mark.parametrize("method", ["kdtree", "sweep", "lds"])
def test_raw_scene_becomes_a_scene_the_oracle_can_trace(g, orc, scenes, method):
    """ToModel: 8-triangle leaves expanded into 2-triangle leaves under the file's inner nodes -- a hierarchy in the chunk's layout (preorder, escape
    links, leaf pairs).  Closest hits on it equal those on the bim SAH hierarchy of the same mesh (distance bit for bit; ids differ by layout)."""
    from tests.conftest import make_ray_batch
    pos, nrm, uv, tri, mat = scenes.box1_mesh(3)
    raw = g.RawScene.build(pos, nrm, uv, tri, mat, scenes.BOX1_MATERIAL_NAMES, method)
    sc = raw.to_scene(scenes.box1_materials())
    ref = scenes.make_box1(tess=3)
    assert sc.info.num_triangles == ref.info.num_triangles and sc.info.num_materials == 9
    lt = sc.leaf_triangles()
    assert len(lt) % 2 == 0 and (lt["v"][:, 0] != lt["v"][:, 1]).sum() == len(tri)
    nodes = sc.nodes()
    leaf_ids = nodes["first_child"][(nodes["first_child"] & 0x80000000) != 0] & 0x7FFFFFFF
    assert np.array_equal(leaf_ids, np.arange(len(leaf_ids)))          # leaves numbered in preorder
    rays = make_ray_batch(g, ref, 3000, 5, any_fraction=0.25)
    a, b = orc.trace_batch(sc.desc, rays), orc.trace_batch(ref.desc, rays)
    closest = rays["flags"] == 0
    assert np.array_equal(a["t"][closest].view(np.uint32), b["t"][closest].view(np.uint32))
    assert np.array_equal(a["leaf_tri"][~closest], b["leaf_tri"][~closest])
    assert np.array_equal(a["leaf_tri"] >= 0, b["leaf_tri"] >= 0)
    # the same triangle was hit: compare vertex triples
    ha = lt["v"][a["leaf_tri"][closest & (a["leaf_tri"] >= 0)]]
    hb = ref.leaf_triangles()["v"][b["leaf_tri"][closest & (b["leaf_tri"] >= 0)]]
    same = (ha == hb).all(axis=1)
    assert same.mean() > 0.99   # (exact ties between coplanar neighbours may resolve to the other triangle: the visit order differs)


@pytest.mark.gpu
def test_raw_scene_on_the_device(g, orc, scenes):
    """The expanded hierarchy through gi_upload_scene (incl. the upload-time re-clustering) and both walks: bit-exact against the oracle."""
    from tests.conftest import make_ray_batch
    pos, nrm, uv, tri, mat = scenes.box1_mesh(4)
    for method in ("kdtree", "sweep", "lds"):
        sc = g.RawScene.build(pos, nrm, uv, tri, mat, scenes.BOX1_MATERIAL_NAMES, method).to_scene(scenes.box1_materials())
        rays = make_ray_batch(g, sc, 100000, 9, any_fraction=0.25)
        want = orc.trace_batch(sc.desc, rays)
        dev = g.Device(0)
        dev.upload_scene(sc)
        for mode in (g.GI_TRAVERSAL_FAST, g.GI_TRAVERSAL_REFERENCE):
            dev.set_option(g.GI_OPT_TRAVERSAL, mode)
            got = dev.trace(rays)
            assert np.array_equal(got["leaf_tri"], want["leaf_tri"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
        dev.close()
