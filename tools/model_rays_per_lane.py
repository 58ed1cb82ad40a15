"""Lockstep model (tools/bvh_cost.py) of a bounce wave in which every lane walks K rays back to back in
one loop, the warp reconverging only after 32 x K rays: lane utilisation of the box tests rises from 14 to 24 of 32
in the Cornell box, but the leaf-test block then runs in nearly every iteration — total cost -8 % at best, +20 % on
coherent waves. Not built (profiles/r02_experiments.md).

    python tools/model_rays_per_lane.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from bvh_cost import *  # noqa: E402,F401,F403
SWITCH = 14   # divergent block when a lane moves on to its next ray (store hit, load o/d/inv, octant base)

def walk_kray(o, d, layouts, tris, K):
    """each lane walks K rays back to back in ONE loop (no reconvergence between rays); warp = 32 lanes x K rays"""
    offs, Bs, Ss, Fs, Cs = [0], [], [], [], []
    for b, s, f, c in layouts:
        Bs.append(b); Ss.append(np.where(s >= 0, s + offs[-1], END)); Fs.append(f); Cs.append(c)
        offs.append(offs[-1] + len(b))
    bounds, skip, lfirst, lcnt = np.concatenate(Bs), np.concatenate(Ss), np.concatenate(Fs), np.concatenate(Cs)
    R = len(o); per = 32 * K
    Rp = (R + per - 1)//per*per
    oo = np.zeros((Rp,3),np.float32); oo[:R]=o; dd=np.ones((Rp,3),np.float32); dd[:R]=d
    valid = np.zeros(Rp,bool); valid[:R]=True
    octant = (dd[:, 0] < 0) * 1 + (dd[:, 1] < 0) * 2 + (dd[:, 2] < 0) * 4
    start = np.array(offs[:8])[octant]
    W = Rp // per
    # lane l of warp w handles rays w*per + j*32 + l, j = 0..K-1
    lane_ray = (np.arange(W)[:,None,None]*per + np.arange(K)[None,None,:]*32 + np.arange(32)[None,:,None])  # [W,32,K]
    j = np.zeros((W,32), np.int64)
    cur = lane_ray[:,:,0].copy()
    node = np.where(valid[cur], start[cur], END)
    best = np.full((W,32), np.inf, np.float32); btri_all = np.full(Rp, -1, np.int64)
    btri = np.full((W,32), -1, np.int64)
    iters = np.zeros(W, np.int64); leafs = np.zeros(W, np.int64); switches = np.zeros(W, np.int64)
    lanes_sum = 0
    with np.errstate(all="ignore"):
        inv = (1.0 / dd).astype(np.float32)
    done = np.zeros((W,32), bool)
    while not done.all():
        # lanes whose ray ended switch to the next one (divergent block)
        ended = (node < 0) & ~done
        if ended.any():
            switches[ended.any(1)] += 1
            wi, li = np.nonzero(ended)
            btri_all[cur[wi,li]] = btri[wi,li]
            j[wi,li] += 1
            fin = j[wi,li] >= K
            done[wi[fin], li[fin]] = True
            wi2, li2 = wi[~fin], li[~fin]
            cur[wi2,li2] = lane_ray[wi2,li2,j[wi2,li2]]
            nv = valid[cur[wi2,li2]]
            node[wi2,li2] = np.where(nv, start[cur[wi2,li2]], END)
            best[wi2,li2] = np.inf; btri[wi2,li2] = -1
        act = (node >= 0)
        if not act.any():
            continue
        wi, li = np.nonzero(act)
        r = cur[wi,li]; nd = node[wi,li]; b = bounds[nd]
        with np.errstate(all="ignore"):
            tx0, tx1 = (b[:, 0] - oo[r, 0]) * inv[r, 0], (b[:, 1] - oo[r, 0]) * inv[r, 0]
            ty0, ty1 = (b[:, 2] - oo[r, 1]) * inv[r, 1], (b[:, 3] - oo[r, 1]) * inv[r, 1]
            tz0, tz1 = (b[:, 4] - oo[r, 2]) * inv[r, 2], (b[:, 5] - oo[r, 2]) * inv[r, 2]
            t0 = np.fmax(np.fmax(np.fmin(tx0, tx1), np.fmin(ty0, ty1)), np.fmax(np.fmin(tz0, tz1), 0))
            t1 = np.fmin(np.fmin(np.fmax(tx0, tx1), np.fmax(ty0, ty1)), np.fmin(np.fmax(tz0, tz1), best[wi,li]))
        hit = t1 >= t0
        iters[np.unique(wi)] += 1
        lanes_sum += len(wi)
        leaf = hit & (lfirst[nd] >= 0)
        if leaf.any():
            lw, ll, lr, ln = wi[leaf], li[leaf], r[leaf], nd[leaf]
            maxc = lcnt[ln]
            for k in range(int(maxc.max())):
                m = k < maxc
                tri_idx = lfirst[ln[m]] + k
                ok, t, nn = tri_hit(oo[lr[m]], dd[lr[m]], tris, tri_idx, best[lw[m], ll[m]])
                bw, bl = lw[m][ok], ll[m][ok]
                best[bw, bl] = t[ok]; btri[bw, bl] = tri_idx[ok]
                leafs[np.unique(lw[m])] += 1
        node[wi,li] = np.where(hit & (lfirst[nd] < 0), nd + 1, skip[nd])
    return btri_all[:R], dict(iters=iters, leafs=leafs, switches=switches, lanes=lanes_sum / max(iters.sum(),1))

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
        line = f"{wl} wave {wave}: base {base/1e6:.2f}M lanes {st['node_lanes'].sum()/st['iters'].sum():.1f}"
        for K in (2, 4, 8, 16):
            t2, s2 = walk_kray(o, d, ftb, tris, K)
            assert np.array_equal(t2, btri), (K, (t2 != btri).sum())
            c = nc * s2["iters"].sum() + LEAF * s2["leafs"].sum() + SWITCH * s2["switches"].sum()
            line += f" | K{K}: {c/1e6:.2f}M ({100*c/base-100:+.0f}%) lanes {s2['lanes']:.1f}"
        print(line)
        o, d = lambert_bounce(o, d, best, bn, rng)
run("Cornell box (C3)")
run("built-in, default pose", waves=2)
