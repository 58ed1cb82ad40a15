"""Lockstep model of warp-level leaf-test batching (the "N1" experiments of round 2; results in
profiles/r02_experiments.md): every lane parks up to K leaves while it walks on, the warp flushes
when T (ray, triangle) pairs are parked — either redistributed over all 32 lanes with shuffles
("coop": 40 instructions of overhead per round assumed) or tested by the lanes that own them.
Built on tools/bvh_cost.py. The model predicted a gain for the Cornell box's bounce waves; the
kernel that was built from it (per-lane flush, K = 4, T = 24) measured 30 % SLOWER on the B200 —
the per-iteration warp votes cost more than the divergence they remove.

    python tools/leaf_batching_model.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from bvh_cost import *  # noqa: E402,F401,F403
COOP_OVERHEAD = 40  # gather of (ray, triangle) pairs by shuffle + segmented min back to the owning lane

def walk_coop(o, d, layouts, tris, K, T):
    """per-lane FIFO of up to K postponed leaves; the warp flushes all pending (ray, triangle) pairs
    cooperatively (32 pairs per round) when >= T pairs are pending or nobody can advance"""
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
    iters = np.zeros(W, np.int64); rounds = np.zeros(W, np.int64); pairs_total = 0
    nodes_per_ray = np.zeros(Rp, np.int64)
    with np.errstate(all="ignore"):
        inv = (1.0 / dd).astype(np.float32)
    while True:
        adv = (node >= 0) & ~stalled
        if not adv.any() and not (npend > 0).any(): break
        idx = np.nonzero(adv)[0]
        if len(idx):
            nd = node[idx]; b = bounds[nd]
            with np.errstate(all="ignore"):
                tx0, tx1 = (b[:, 0] - oo[idx, 0]) * inv[idx, 0], (b[:, 1] - oo[idx, 0]) * inv[idx, 0]
                ty0, ty1 = (b[:, 2] - oo[idx, 1]) * inv[idx, 1], (b[:, 3] - oo[idx, 1]) * inv[idx, 1]
                tz0, tz1 = (b[:, 4] - oo[idx, 2]) * inv[idx, 2], (b[:, 5] - oo[idx, 2]) * inv[idx, 2]
                t0 = np.fmax(np.fmax(np.fmin(tx0, tx1), np.fmin(ty0, ty1)), np.fmax(np.fmin(tz0, tz1), 0))
                t1 = np.fmin(np.fmin(np.fmax(tx0, tx1), np.fmax(ty0, ty1)), np.fmin(np.fmax(tz0, tz1), best[idx]))
            hit = t1 >= t0
            nodes_per_ray[idx] += 1
            iters[np.unique(idx // 32)] += 1
            isleaf = hit & (lfirst[nd] >= 0)
            room = npend[idx] < K
            park = isleaf & room
            stall = isleaf & ~room
            pi = idx[park]
            pend[pi, npend[pi]] = nd[park]; npend[pi] += 1
            stalled[idx[stall]] = True
            newnode = np.where(hit & (lfirst[nd] < 0), nd + 1, skip[nd])
            node[idx] = np.where(stall, nd, newnode)
        # pairs pending per warp (leaf triangle counts)
        cnt = np.where(pend >= 0, lcnt[np.maximum(pend, 0)], 0).sum(1)
        pairs_w = cnt.reshape(W, 32).sum(1)
        adv_w = ((node >= 0) & ~stalled).reshape(W, 32).any(1)
        flush_w = (pairs_w >= T) | (~adv_w & (pairs_w > 0))
        if flush_w.any():
            rounds[flush_w] += (pairs_w[flush_w] + 31) // 32
            pairs_total += int(pairs_w[flush_w].sum())
            fl = np.repeat(flush_w, 32) & (npend > 0)
            li = np.nonzero(fl)[0]
            # test in FIFO order per lane (visit order), exact semantics
            for k in range(K):
                has = li[npend[li] > k]
                if not len(has): break
                leafs = pend[has, k]
                maxc = lcnt[leafs]
                for j in range(int(maxc.max())):
                    sel = has[j < maxc]; lf = pend[sel, k]
                    tri_idx = lfirst[lf] + j
                    ok, t, nn = tri_hit(oo[sel], dd[sel], tris, tri_idx, best[sel])
                    best[sel[ok]] = t[ok]; btri[sel[ok]] = tri_idx[ok]
            pend[li] = -1; npend[li] = 0
            stalled[np.repeat(flush_w, 32)] = False
    return btri[:R], dict(iters=iters, rounds=rounds, pairs=pairs_total, nodes_per_ray=nodes_per_ray[:R])

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
        line = f"{wl} wave {wave}: base {base/1e6:.2f}M (box {nc*st['iters'].sum()/1e6:.2f} + leaf {LEAF*st['leaf'].sum()/1e6:.2f}; nodes/ray {st['nodes_per_ray'].mean():.1f})"
        for K, T in ((1, 8), (2, 16), (4, 32), (4, 16), (8, 32)):
            t2, s2 = walk_coop(o, d, ftb, tris, K, T)
            assert np.array_equal(t2, btri)
            c = (nc + 4) * s2["iters"].sum() + (LEAF + COOP_OVERHEAD) * s2["rounds"].sum()
            line += f" | K{K} T{T}: {c/1e6:.2f}M ({100*c/base-100:+.0f}%) n/r {s2['nodes_per_ray'].mean():.1f} pairs/round {s2['pairs']/max(s2['rounds'].sum(),1):.1f}"
        print(line)
        o, d = lambert_bounce(o, d, best, bn, rng)
run("Cornell box (C3)")
run("built-in, default pose", waves=2)
