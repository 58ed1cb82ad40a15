"""Lockstep model (tools/bvh_cost.py) of vote-free leaf batching: every lane parks up to K leaves and all lanes
flush every M node steps (a uniform loop counter instead of warp votes). Best case -14 % per later Cornell wave,
+16..+28 % on coherent waves; given that the model was optimistic for every variant that was built, not built
(profiles/r02_experiments.md).

    python tools/model_periodic_flush.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from bvh_cost import *  # noqa: E402,F401,F403

def walk_periodic(o, d, layouts, tris, K, M, book=3):
    """per-lane FIFO of K parked leaves, flushed every M node steps (uniform loop counter, no votes)"""
    offs, Bs, Ss, Fs, Cs = [0], [], [], [], []
    for b, s, f, c in layouts:
        Bs.append(b); Ss.append(np.where(s >= 0, s + offs[-1], END)); Fs.append(f); Cs.append(c)
        offs.append(offs[-1] + len(b))
    bounds, skip, lfirst, lcnt = np.concatenate(Bs), np.concatenate(Ss), np.concatenate(Fs), np.concatenate(Cs)
    octant = (d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4
    R = len(o); Rp = (R + 31)//32*32
    node = np.full(Rp, END, np.int64); node[:R] = np.array(offs[:8])[octant]
    oo = np.zeros((Rp,3),np.float32); oo[:R]=o; dd=np.ones((Rp,3),np.float32); dd[:R]=d
    best = np.full(Rp, np.inf, np.float32); btri = np.full(Rp, -1, np.int64)
    pend = np.full((Rp, K), -1, np.int64); npend = np.zeros(Rp, np.int64); stalled = np.zeros(Rp, bool)
    W = Rp // 32
    iters = np.zeros(W, np.int64); rounds = np.zeros(W, np.int64); lanes=0; leaf_lanes=0
    nodes_per_ray = np.zeros(Rp, np.int64)
    with np.errstate(all="ignore"):
        inv = (1.0 / dd).astype(np.float32)
    it = 0
    while True:
        alive_lane = (node >= 0) | (npend > 0)
        if not alive_lane.any(): break
        it += 1
        adv = (node >= 0) & ~stalled
        idx = np.nonzero(adv)[0]
        warp_alive = alive_lane.reshape(W,32).any(1)
        iters[warp_alive] += 1   # a warp pays the node-step slot while any of its lanes is still in the loop
        if len(idx):
            nd = node[idx]; b = bounds[nd]
            with np.errstate(all="ignore"):
                tx0, tx1 = (b[:, 0] - oo[idx, 0]) * inv[idx, 0], (b[:, 1] - oo[idx, 0]) * inv[idx, 0]
                ty0, ty1 = (b[:, 2] - oo[idx, 1]) * inv[idx, 1], (b[:, 3] - oo[idx, 1]) * inv[idx, 1]
                tz0, tz1 = (b[:, 4] - oo[idx, 2]) * inv[idx, 2], (b[:, 5] - oo[idx, 2]) * inv[idx, 2]
                t0 = np.fmax(np.fmax(np.fmin(tx0, tx1), np.fmin(ty0, ty1)), np.fmax(np.fmin(tz0, tz1), 0))
                t1 = np.fmin(np.fmin(np.fmax(tx0, tx1), np.fmax(ty0, ty1)), np.fmin(np.fmax(tz0, tz1), best[idx]))
            hit = t1 >= t0
            nodes_per_ray[idx] += 1; lanes += len(idx)
            isleaf = hit & (lfirst[nd] >= 0)
            room = npend[idx] < K
            park = isleaf & room; stall = isleaf & ~room
            pi = idx[park]
            pend[pi, npend[pi]] = nd[park]; npend[pi] += 1
            stalled[idx[stall]] = True
            newnode = np.where(hit & (lfirst[nd] < 0), nd + 1, skip[nd])
            node[idx] = np.where(stall, nd, newnode)
        if it % M == 0:
            cnt = np.where(pend >= 0, lcnt[np.maximum(pend, 0)], 0).sum(1)
            rw = cnt.reshape(W, 32).max(1)
            rounds += rw
            leaf_lanes += int(cnt.sum())
            li = np.nonzero(npend > 0)[0]
            for k in range(K):
                has = li[npend[li] > k]
                if not len(has): break
                leafs = pend[has, k]; maxc = lcnt[leafs]
                for j in range(int(maxc.max())):
                    sel = has[j < maxc]; lf = pend[sel, k]
                    tri_idx = lfirst[lf] + j
                    ok, t, nn = tri_hit(oo[sel], dd[sel], tris, tri_idx, best[sel])
                    best[sel[ok]] = t[ok]; btri[sel[ok]] = tri_idx[ok]
            pend[li] = -1; npend[li] = 0; stalled[:] = False
    return btri[:R], dict(iters=iters, rounds=rounds, lanes=lanes/max(iters.sum(),1), leaf_lanes=leaf_lanes/max(rounds.sum(),1), npr=nodes_per_ray[:R].mean())

def run(wl, W=480, H=272, waves=4):
    sname, pose, fov = WORKLOADS[wl]
    scene = rv.builtin_scene() if sname == "builtin" else rv.cornell_scene()
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = np.ascontiguousarray(scene.triangles[perm])
    ftb = front_to_back_layouts(nodes, tris)
    rng = np.random.default_rng(7)
    o, d = camera_rays(W, H, pose, fov, rng)
    for wave in range(waves):
        if wave > 0:
            p = np.argsort((d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4, kind="stable")
            o, d = o[p], d[p]
        best, btri, bn, st = walk(o, d, ftb, tris)
        nc = (NODE_PRIMARY if wave == 0 else NODE_BOUNCE) + LOOP
        base = nc * st["iters"].sum() + LEAF * st["leaf"].sum()
        line = f"{wl} wave {wave}: base {base/1e6:.2f}M (n/r {st['nodes_per_ray'].mean():.1f})"
        for K, M in ((2, 2), (4, 4), (4, 6), (4, 8), (6, 12)):
            t2, s2 = walk_periodic(o, d, ftb, tris, K, M)
            assert np.array_equal(t2, btri), (K, M, (t2 != btri).sum())
            c = (nc + 3) * s2["iters"].sum() + (LEAF + 6) * s2["rounds"].sum() + 4 * s2["iters"].sum() / M
            line += f" | K{K} M{M}: {100*c/base-100:+.0f}% n/r {s2['npr']:.1f} lanes {s2['lanes']:.1f}/{s2['leaf_lanes']:.1f}"
        print(line)
        o, d = lambert_bounce(o, d, best, bn, rng)
run("Cornell box (C3)")
run("built-in, default pose", waves=2)
