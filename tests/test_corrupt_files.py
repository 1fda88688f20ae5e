"""Corrupt input files are refused, not followed: `.bim` (host/bim.cpp), `.rawscene` (host/bvhmake.cpp), the JSON environment file (host/json.hpp).
The reference trusts its files (scene.cpp:15-59, bim_io.cpp:78-166, bvhmake main.cpp:127-143; `Scene::SanityCheck`, scene.hpp:191, only runs in
debug builds); a renderer service that loads user-supplied scenes must not.  The same mutators ran 20 000 - 60 000 times per format against the host layer
built with -fsanitize=address,undefined while this was written; the cases below are the classes that had crashed it, plus a random sample."""
import os
import random
import struct
import zlib

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHUNK, AABOX, HIERARCHY, PARENTS, LEAVES, POSITION = 0x3, 0x01000000, 0x08000000, 0x08000001, 0x08000002, 0x1


def _sections(buf):
    out, pos = [], 0
    while pos + 24 <= len(buf):
        typ, _, size, usize = struct.unpack_from("<IIQQ", buf, pos)
        out.append([typ, bytes(buf[pos + 24: pos + 24 + size]), usize])
        pos += 24 + size
    return out


def _pack(sections):
    return b"".join(struct.pack("<IIQQ", t, 0, len(d), u) + d for t, d, u in sections)


def _chunk_sections(top):
    """sub-sections of the chunk, inflated: [[type, raw bytes], ...]"""
    data = next(d for t, d, _ in top if t == CHUNK)
    return [[t, bytearray(zlib.decompress(d) if u else d)] for t, d, u in _sections(data)]


def _write_case(tmp_path, top, subs):
    body = _pack([(t, bytes(d), 0) for t, d in subs])
    blob = _pack([(t, body if t == CHUNK else d, 0 if t == CHUNK else u) for t, d, u in top])
    (tmp_path / "ref_box1_t2.bim").write_bytes(blob)
    (tmp_path / "x.json").write_text(open(os.path.join(GOLDEN, "ref_box1_t2.json")).read())
    return str(tmp_path / "x.json")


@pytest.fixture(scope="module")
def ref_bim():
    return _sections(open(os.path.join(GOLDEN, "ref_box1_t2.bim"), "rb").read())


def test_bim_rewritten_uncompressed_still_loads(g, ref_bim, tmp_path):
    """the mutator's own round trip: the reference-written file with every chunk section stored raw loads like the original"""
    a = g.Scene.load(os.path.join(GOLDEN, "ref_box1_t2.json"))
    b = g.Scene.load(_write_case(tmp_path, ref_bim, _chunk_sections(ref_bim)))
    ia, ib = a.info, b.info
    for f in ("num_triangles", "num_vertices", "num_nodes", "num_leaf_triangles", "num_materials", "num_tree_levels", "num_light_triangles"):
        assert getattr(ia, f) == getattr(ib, f), f
    assert np.array_equal(a.nodes(), b.nodes())


def test_bim_with_broken_links_is_refused(g, ref_bim, tmp_path):
    def case(edit):
        subs = _chunk_sections(ref_bim)
        edit({t: d for t, d in subs}, subs)
        return _write_case(tmp_path, ref_bim, subs)

    def child_in_front(by_type, subs):      # an inner node whose first child is itself: the box pass and every walk would loop
        nodes = np.frombuffer(by_type[HIERARCHY], dtype=np.uint32).reshape(-1, 2)
        inner = int(np.nonzero((nodes[:, 0] & 0x80000000) == 0)[0][1])
        struct.pack_into("<I", by_type[HIERARCHY], inner * 8, inner)

    def escape_in_front(by_type, subs):
        struct.pack_into("<I", by_type[HIERARCHY], 5 * 8 + 4, 2)

    def leaf_out_of_range(by_type, subs):
        nodes = np.frombuffer(by_type[HIERARCHY], dtype=np.uint32).reshape(-1, 2)
        leaf = int(np.nonzero(nodes[:, 0] & 0x80000000)[0][0])
        struct.pack_into("<I", by_type[HIERARCHY], leaf * 8, 0x80000000 | 100000)

    def vertex_out_of_range(by_type, subs):
        struct.pack_into("<I", by_type[LEAVES], 0, 1 << 30)

    def parents_missing(by_type, subs):
        subs[:] = [s for s in subs if s[0] != PARENTS]

    def fewer_boxes(by_type, subs):
        del by_type[AABOX][-24:]

    for edit in (child_in_front, escape_in_front, leaf_out_of_range, vertex_out_of_range, parents_missing, fewer_boxes):
        with pytest.raises(g.GpugiError):
            g.Scene.load(case(edit))


def test_bim_without_hierarchy_ignores_the_file_s_boxes(g, ref_bim, tmp_path):
    """boxes of a tree that is not in the file are void: the hierarchy is rebuilt and the boxes with it (they had been read past their end)"""
    subs = [s for s in _chunk_sections(ref_bim) if s[0] not in (HIERARCHY, PARENTS, LEAVES)]
    sc = g.Scene.load(_write_case(tmp_path, ref_bim, subs))
    ref = g.Scene.load(os.path.join(GOLDEN, "ref_box1_t2.json"))
    assert sc.info.num_triangles == ref.info.num_triangles and sc.info.num_nodes > 0


def test_bim_sections_cannot_claim_more_than_the_file_holds(g, ref_bim, tmp_path):
    subs = _chunk_sections(ref_bim)
    body = bytearray(_pack([(t, bytes(d), 0) for t, d in subs]))
    off = 0
    for t, d in subs:                                                # the vertex positions: a section of a terabyte
        if t == POSITION:
            break
        off += 24 + len(d)
    struct.pack_into("<Q", body, off + 8, 1 << 40)
    blob = _pack([(t, bytes(body) if t == CHUNK else d, 0 if t == CHUNK else u) for t, d, u in ref_bim])
    (tmp_path / "ref_box1_t2.bim").write_bytes(blob)
    (tmp_path / "x.json").write_text(open(os.path.join(GOLDEN, "ref_box1_t2.json")).read())
    with pytest.raises(g.GpugiError):
        g.Scene.load(str(tmp_path / "x.json"))
    # a compressed section that claims to inflate to 2^40 bytes
    data = next(d for t, d, _ in ref_bim if t == CHUNK)
    inner = _sections(data)
    k = next(i for i, (t, d, u) in enumerate(inner) if t == POSITION and u)
    inner[k][2] = 1 << 40
    blob = _pack([(t, _pack(inner) if t == CHUNK else d, u) for t, d, u in ref_bim])
    (tmp_path / "ref_box1_t2.bim").write_bytes(blob)
    with pytest.raises(g.GpugiError):
        g.Scene.load(str(tmp_path / "x.json"))


def test_random_bim_mutations_load_or_are_refused(g, ref_bim, tmp_path):
    rng = random.Random(11)
    picks = [0, 1, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF]
    loaded = refused = 0
    for _ in range(300):
        subs = _chunk_sections(ref_bim)
        for s in subs:
            if rng.random() < 0.12 and len(s[1]) >= 4:
                for _ in range(rng.randint(1, 3)):
                    mode, p = rng.randint(0, 3), rng.randrange(0, max(len(s[1]) // 4, 1)) * 4
                    if len(s[1]) < 4:
                        break
                    if mode == 0:
                        s[1][p:p + 4] = struct.pack("<I", rng.choice(picks + [rng.randrange(0, 4000), 0x80000000 | rng.randrange(0, 4000)]))
                    elif mode == 1:
                        s[1][rng.randrange(len(s[1]))] = rng.randrange(256)
                    elif mode == 2:
                        del s[1][rng.randrange(0, len(s[1]) + 1) // 4 * 4:]
                    else:
                        s[1][p:p + 4] = struct.pack("<f", rng.choice([float("nan"), float("inf"), -float("inf"), 1e38, 0.0]))
            if rng.random() < 0.03:
                s[0] = rng.choice([0x1, 0x2, 0x20, 0x10000, 0x20000, AABOX, 0x02000000, HIERARCHY, PARENTS, LEAVES, 0x6])
        subs = [s for s in subs if rng.random() >= 0.03]
        try:
            sc = g.Scene.load(_write_case(tmp_path, ref_bim, subs), bvh_type=rng.randint(0, 1))
            assert sc.info.num_nodes > 0
            loaded += 1
        except g.GpugiError:
            refused += 1
    assert loaded > 30 and refused > 30, (loaded, refused)


def test_rawscene_with_broken_hierarchy_is_refused(g, scenes, tmp_path):
    pos, nrm, uv, tri, mat = _grid_mesh()
    raw = g.RawScene.build(pos, nrm, uv, tri, mat % 2, ["A", "B"], method="kdtree")
    path = str(tmp_path / "m.rawscene")
    raw.store(path)
    blob = bytearray(open(path, "rb").read())
    arrays, off = {}, 0
    while off < len(blob):
        name, count, size = struct.unpack_from("<32sII", blob, off)
        arrays[name.split(b"\0")[0].decode()] = (off + 40, count, size)
        off += 40 + count * size
    h0, n, _ = arrays["hierarchy"]
    nodes = np.frombuffer(bytes(blob[h0:h0 + 12 * n]), dtype=np.uint32).reshape(-1, 3)      # parent, firstChild, escape
    inner = int(np.nonzero((nodes[:, 1] & 0x80000000) == 0)[0][-1])

    def variant(offset, value):
        b = bytearray(blob)
        struct.pack_into("<I", b, offset, value)
        open(path, "wb").write(bytes(b))
        return path

    for offset, value in ((h0 + 12 * inner + 4, 0),                 # first child = the root: the expansion walked in circles until the stack ran out
                          (h0 + 12 * inner + 4, inner),             # first child = the node itself
                          (h0 + 12 * (inner + 1) + 8, 0),           # the first child's escape (= its sibling) points at the root
                          (arrays["vertices"][0] - 40 + 32, 1 << 30)):   # an array header that claims 2^30 elements
        with pytest.raises(g.GpugiError):
            g.RawScene.load(variant(offset, value))
    open(path, "wb").write(bytes(blob))
    assert g.RawScene.load(path).nodes().shape[0] == n


def _grid_mesh(n=6):
    ys, xs = np.mgrid[0:n + 1, 0:n + 1]
    pos = np.stack([xs.ravel(), ys.ravel(), (xs * ys).ravel() % 3], axis=1).astype(np.float32)
    nrm = np.tile(np.float32([0, 0, 1]), (pos.shape[0], 1))
    uv = np.stack([xs.ravel() / n, ys.ravel() / n], axis=1).astype(np.float32)
    a = (ys[:-1, :-1] * (n + 1) + xs[:-1, :-1]).ravel()
    tri = np.concatenate([np.stack([a, a + 1, a + n + 1], axis=1), np.stack([a + 1, a + n + 2, a + n + 1], axis=1)]).astype(np.uint32)
    return pos, nrm, uv, tri, np.arange(tri.shape[0], dtype=np.uint32)


def test_deeply_nested_json_is_refused(g, tmp_path):
    for text in ('{"materials": ' + "[" * 200000 + "}", '{"a":' * 200000):
        p = tmp_path / "deep.json"
        p.write_text(text)
        with pytest.raises(g.GpugiError) as e:
            g.Scene.load(str(p))
        assert "nested" in str(e.value)


def test_upload_refuses_cyclic_links(g, emul, box1):
    """gi_upload_scene's validation (csrc/gi_layout.h buildHostLayout, the same code in the host-built emulator): links that form a cycle are
    refused -- also one that only a MISSED inner node would take, which the all-boxes-hit walk of the layout code never follows -- because the
    stackless walk would never leave it on the device.  In-range, acyclic links keep loading."""
    import ctypes as C
    rays = np.zeros(1, dtype=g.RAY_DTYPE)
    rays["direction"][0] = (0, 0, -1)
    rays["origin"][0] = (0, 1.5, 2.0)
    rays["t_max"] = g.RAY_MAX
    d = box1.desc
    nodes = box1.nodes()
    inner = np.nonzero((nodes["first_child"] & 0x80000000) == 0)[0]
    emul.trace(d, rays, 0)                                   # the untouched scene is accepted

    def with_nodes(n):
        d2 = box1.desc
        d2.nodes = C.cast(n.ctypes.data, type(d2.nodes))
        return d2

    deep = int(inner[len(inner) // 2])
    anc = int(inner[1])
    assert anc < deep and nodes["escape"][deep] != anc
    cyc = nodes.copy()
    cyc["escape"][deep] = anc                                # a miss at `deep` jumps back to an earlier inner node that leads to `deep` again
    with pytest.raises(RuntimeError):
        emul.trace(with_nodes(cyc), rays, 0)
    cyc = nodes.copy()
    cyc["first_child"][deep] = deep                          # self loop
    with pytest.raises(RuntimeError):
        emul.trace(with_nodes(cyc), rays, 0)
    ok = nodes.copy()
    ok["escape"][deep] = 0                                   # ends the walk early: wrong picture, but every ray terminates
    emul.trace(with_nodes(ok), rays, 0)
