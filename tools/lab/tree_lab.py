#!/usr/bin/env python
"""CPU lab for the upload topology (GI_OPT_UPLOAD_TOPOLOGY): builds trees over the REFERENCE'S leaves of the bench scene in numpy, runs
the product's lane walk on them through the host emulator (tests/emul, em_set_tree) and prints box / triangle tests per ray, tree depth
and SAH cost -- to compare builder variants without a GPU.  Test / analysis tooling only; nothing in the product imports it.
Usage: python tools/lab/tree_lab.py [--rays 30000] [--variants ref,ploc4,ploc16,...]"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def expand21(v):
    v = v.astype(np.uint64) & np.uint64(0x1fffff)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton_order(bmin, bmax, rlo, rhi):
    c = 0.5 * bmin + 0.5 * bmax
    ext = np.maximum(rhi - rlo, 1e-30)
    q = (np.clip((c - rlo) / ext, 0, 1) * 2097151.0).astype(np.uint64)
    codes = (expand21(q[:, 0]) << np.uint64(2)) | (expand21(q[:, 1]) << np.uint64(1)) | expand21(q[:, 2])
    return np.argsort(codes, kind="stable")


def area(mn, mx):
    d = mx - mn
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]


def ploc(bmin, bmax, order, radius, radius_schedule=None):
    """PLOC as in gi_lbvh.cuh (nearest within +-radius by union area, ties to the smaller position, mutual pairs merge, in-order compaction)."""
    nl = len(order)
    N = 2 * nl - 1
    mn = np.zeros((N, 3), np.float32); mx = np.zeros((N, 3), np.float32)
    mn[:nl] = bmin[order]; mx[:nl] = bmax[order]
    left = np.zeros(N, np.uint32); right = np.zeros(N, np.uint32)
    clusters = np.arange(nl, dtype=np.int64)
    nxt = nl
    p = 0
    while len(clusters) > 1:
        n = len(clusters)
        r = radius if radius_schedule is None else radius_schedule(p, n)
        cmn, cmx = mn[clusters], mx[clusters]
        best = np.full(n, np.inf, np.float32); bj = np.full(n, -1, np.int64)
        for o in list(range(-r, 0)) + list(range(1, r + 1)):
            j = np.arange(n) + o
            ok = (j >= 0) & (j < n)
            jj = np.clip(j, 0, n - 1)
            a = area(np.minimum(cmn, cmn[jj]), np.maximum(cmx, cmx[jj])).astype(np.float32)
            a[~ok] = np.inf
            better = a < best
            best[better] = a[better]; bj[better] = j[better]
        i = np.arange(n)
        mutual = (bj >= 0) & (bj[np.clip(bj, 0, n - 1)] == i)
        lead = mutual & (i < bj)
        drop = mutual & (i > bj)
        ids = nxt + np.arange(lead.sum())
        l = clusters[lead]; rr = clusters[bj[lead]]
        mn[ids] = np.minimum(mn[l], mn[rr]); mx[ids] = np.maximum(mx[l], mx[rr])
        left[ids] = l; right[ids] = rr
        out = clusters.copy(); out[lead] = ids
        clusters = out[~drop]
        nxt += len(ids)
        p += 1
    return left, right, int(clusters[0]), mn, mx, p


def sah_sweep(bmin, bmax):
    """Top-down full-sweep SAH over the leaves, all three axes (the quality yardstick).  Returns left, right, leaf_of, root, mn, mx."""
    nl = len(bmin)
    N = 2 * nl - 1
    mn = np.zeros((N, 3), np.float32); mx = np.zeros((N, 3), np.float32)
    mn[:nl] = bmin; mx[:nl] = bmax
    left = np.zeros(N, np.uint32); right = np.zeros(N, np.uint32)
    cen = 0.5 * bmin + 0.5 * bmax
    nxt = [nl]

    def alloc():
        i = nxt[0]; nxt[0] += 1
        return i
    root = alloc()
    stack = [(np.arange(nl), root)]
    while stack:
        ids, node = stack.pop()
        n = len(ids)
        mn[node] = bmin[ids].min(axis=0); mx[node] = bmax[ids].max(axis=0)
        best = (np.inf, None, None)
        for ax in range(3):
            o = ids[np.argsort(cen[ids, ax], kind="stable")]
            lmn = np.minimum.accumulate(bmin[o], axis=0); lmx = np.maximum.accumulate(bmax[o], axis=0)
            rmn = np.minimum.accumulate(bmin[o][::-1], axis=0)[::-1]; rmx = np.maximum.accumulate(bmax[o][::-1], axis=0)[::-1]
            k = np.arange(1, n)
            cost = area(lmn[:-1], lmx[:-1]) * k + area(rmn[1:], rmx[1:]) * (n - k)
            i = int(np.argmin(cost))
            if cost[i] < best[0]:
                best = (float(cost[i]), o, i + 1)
        _, o, split = best
        for part, side in ((o[:split], left), (o[split:], right)):
            if len(part) == 1:
                side[node] = part[0]
            else:
                c = alloc(); side[node] = c
                stack.append((part, c))
    return left, right, np.arange(nl, dtype=np.uint32), root, mn, mx


def tree_stats(left, right, root, mn, mx, nl):
    depth = np.zeros(len(left), np.int32)
    stack = [(root, 1)]
    sah = 0.0
    maxd = 0; sumd = 0
    ra = float(area(mn[root], mx[root]))
    while stack:
        x, dd = stack.pop()
        if x < nl:
            maxd = max(maxd, dd); sumd += dd
            continue
        sah += float(area(mn[x], mx[x])) / ra
        stack.append((int(left[x]), dd + 1)); stack.append((int(right[x]), dd + 1))
    return {"sah_inner": sah, "max_depth": maxd, "mean_leaf_depth": sumd / nl}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=30000)
    ap.add_argument("--variants", default="ref,ploc1,ploc2,ploc4,ploc8,ploc16,ploc32")
    ap.add_argument("--triangles", type=int, default=262144)
    args = ap.parse_args()
    import gpugi_b200 as g
    from gpugi_b200 import scenes
    from oracle import pyoracle as orc
    from tests.emul import pyemul as emul
    from tests.conftest import camera_ray_batch

    tmp = tempfile.mkdtemp(prefix="gpugi_lab_")
    sc = scenes.make_sponza_like(tmp, args.triangles, 32)
    nodes = sc.nodes()
    is_leaf = (nodes["first_child"] & 0x80000000) != 0
    leaf_idx = (nodes["first_child"][is_leaf] & 0x7fffffff).astype(np.int64)
    nl = len(leaf_idx)
    bmin = np.zeros((nl, 3), np.float32); bmax = np.zeros((nl, 3), np.float32)
    bmin[leaf_idx] = nodes["min"][is_leaf]; bmax[leaf_idx] = nodes["max"][is_leaf]
    rlo, rhi = nodes["min"][0], nodes["max"][0]
    order = morton_order(bmin, bmax, rlo, rhi)
    C = scenes.SPONZA_CAMERA
    rays = camera_ray_batch(g, orc, sc, C["pos"], C["lookat"], C["fov"], 160, 90, 4)
    rng = np.random.RandomState(1)
    rays = rays[rng.permutation(len(rays))[: args.rays]]
    print("leaves", nl, "rays", len(rays), "any-hit share %.2f" % (rays["flags"] & 1).mean())
    ref_hits = None
    for v in args.variants.split(","):
        t0 = time.time()
        if v == "ref":
            emul.set_topology(0)
            stats = {}
        else:
            if v.startswith("ploc"):
                left, right, root, mn, mx, passes = ploc(bmin, bmax, order, int(v[4:]))
            elif v.startswith("sched"):
                # radius grows as the cluster list shrinks: small windows at the bottom, wide at the top
                lo_r, hi_r = [int(x) for x in v[5:].split("_")]
                left, right, root, mn, mx, passes = ploc(bmin, bmax, order, lo_r, lambda p, n: hi_r if n < nl // 64 else lo_r)
            elif v == "sah":
                left, right, order_v, root, mn, mx = sah_sweep(bmin, bmax)
                passes = 0
            else:
                raise SystemExit("unknown variant " + v)
            stats = tree_stats(left, right, root, mn, mx, nl)
            stats["passes"] = passes
            emul.set_tree(left, right, order_v if v == "sah" else order, root)
        tb = time.time() - t0
        hits, cnt = emul.trace(sc.desc, rays, 2)
        assert v == "ref" or emul.topology_applied()
        if ref_hits is None:
            ref_hits = hits.copy()
        assert np.array_equal(hits["leaf_tri"], ref_hits["leaf_tri"]) and np.array_equal(hits["t"].view(np.uint32), ref_hits["t"].view(np.uint32))
        print("%-10s box tests/ray %7.2f  tri tests/ray %6.2f  fallback %5d  build %.1fs  %s" % (
            v, cnt["node_tests"] / len(rays), cnt["tri_tests"] / len(rays), cnt["fallback_rays"], tb,
            "  ".join("%s=%.4g" % kv for kv in stats.items())), flush=True)
    emul.set_topology(0)


if __name__ == "__main__":
    main()
