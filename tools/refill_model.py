"""Lane-refill model for the bounce waves (CPU, numpy) — companion of tools/bvh_cost.py.

Today a warp loads 32 queued rays, walks them in lockstep and shades them together: lanes whose
walk ends early idle until the longest one is done, and a leaf test runs for whichever lanes stand
on a leaf in that iteration. This script simulates the alternative — persistent lanes: a lane whose
traversal ends parks until enough lanes are parked (or nobody walks), then the parked lanes are
shaded and refilled from the warp's share of the queue together; lanes that reach a leaf park the
same way and are tested together — and prints the warp-instruction cost of both for the bounce
waves of the bench workloads, over a few thresholds. Same walk, same per-lane visiting order.

    python tools/refill_model.py            # markdown table -> stdout
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import bvh_cost as bc  # noqa: E402
import rvpt_b200 as rv  # noqa: E402

NODE, LOOP, LEAF, VOTES, SHADE = 22, 6, 45, 8, 260  # instruction weights; SHADE = load + shade + push per batch
EMPTY, WALK, ATLEAF, DONE = 0, 1, 2, 3
THRESHOLDS = ((1, 8), (4, 16), (4, 24), (8, 24), (16, 24), (32, 32))


def simulate(o, d, layouts, tris, rays_per_warp, t_leaf, t_done):
    """persistent-lane simulation; returns (cost, box iterations, leaf executions, refills)"""
    offs, Bs, Ss, Fs, Cs = [0], [], [], [], []
    for b, s, f, c in layouts:
        Bs.append(b); Ss.append(np.where(s >= 0, s + offs[-1], -1)); Fs.append(f); Cs.append(c)
        offs.append(offs[-1] + len(b))
    bounds, skip, lfirst, lcnt = np.concatenate(Bs), np.concatenate(Ss), np.concatenate(Fs), np.concatenate(Cs)
    octant = (d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4
    start = np.array(offs[:8])[octant]
    R = len(o)
    W = (R + rays_per_warp - 1) // rays_per_warp
    nxt = np.arange(W) * rays_per_warp                   # next unfetched ray of each warp's share
    end = np.minimum(nxt + rays_per_warp, R)
    ray = np.full((W, 32), -1, np.int64)
    state = np.full((W, 32), EMPTY, np.int8)
    node = np.full((W, 32), -1, np.int64)
    tri_k = np.zeros((W, 32), np.int64)
    best = np.full(R, np.inf, np.float32)
    with np.errstate(all="ignore"):
        inv = (1.0 / d).astype(np.float32)
    cost = np.zeros(W, np.int64)
    n_iter = n_leaf = n_refill = 0
    while True:
        walking = (state == WALK).sum(1)
        atleaf = (state == ATLEAF).sum(1)
        idle = ((state == DONE) | (state == EMPTY)).sum(1)
        has_work = nxt < end
        live = (walking + atleaf + (state == DONE).sum(1) > 0) | has_work
        if not live.any():
            break
        # 1. refill + shade: enough idle lanes (or nothing else to do) and something to shade or fetch
        do_refill = live & (idle > 0) & ((idle >= t_done) | (walking + atleaf == 0)) & \
            (has_work | ((state == DONE).sum(1) > 0))
        if do_refill.any():
            ws = np.nonzero(do_refill)[0]
            n_refill += len(ws)
            cost[ws] += SHADE
            for w in ws:                                 # fetch for idle lanes
                lanes = np.nonzero((state[w] == DONE) | (state[w] == EMPTY))[0]
                k = min(len(lanes), end[w] - nxt[w])
                state[w, lanes] = EMPTY
                if k > 0:
                    ids = np.arange(nxt[w], nxt[w] + k)
                    ray[w, lanes[:k]] = ids
                    node[w, lanes[:k]] = start[ids]
                    state[w, lanes[:k]] = WALK
                    nxt[w] += k
        # 2. leaf phase
        walking = (state == WALK).sum(1)
        atleaf = (state == ATLEAF).sum(1)
        do_leaf = (atleaf > 0) & ((atleaf >= t_leaf) | (walking == 0))
        if do_leaf.any():
            sel = do_leaf[:, None] & (state == ATLEAF)
            wi, li = np.nonzero(sel)
            r = ray[wi, li]
            nd = node[wi, li]
            k = tri_k[wi, li]
            tri_idx = lfirst[nd] + k
            ok, t, _ = bc.tri_hit(o[r], d[r], tris, tri_idx, best[r])
            best[r[ok]] = t[ok]
            more = k + 1 < lcnt[nd]
            tri_k[wi, li] = np.where(more, k + 1, 0)
            nxt_node = skip[nd]
            fin = ~more
            node[wi[fin], li[fin]] = nxt_node[fin]
            state[wi[fin], li[fin]] = np.where(nxt_node[fin] >= 0, WALK, DONE)
            cost[np.unique(wi)] += LEAF
            n_leaf += len(np.unique(wi))
        # 3. box step for walking lanes
        wi, li = np.nonzero(state == WALK)
        if len(wi):
            r = ray[wi, li]
            nd = node[wi, li]
            b = bounds[nd]
            with np.errstate(all="ignore"):
                tx0, tx1 = (b[:, 0] - o[r, 0]) * inv[r, 0], (b[:, 1] - o[r, 0]) * inv[r, 0]
                ty0, ty1 = (b[:, 2] - o[r, 1]) * inv[r, 1], (b[:, 3] - o[r, 1]) * inv[r, 1]
                tz0, tz1 = (b[:, 4] - o[r, 2]) * inv[r, 2], (b[:, 5] - o[r, 2]) * inv[r, 2]
                t0 = np.fmax(np.fmax(np.fmin(tx0, tx1), np.fmin(ty0, ty1)), np.fmax(np.fmin(tz0, tz1), 0))
                t1 = np.fmin(np.fmin(np.fmax(tx0, tx1), np.fmax(ty0, ty1)), np.fmin(np.fmax(tz0, tz1), best[r]))
            hit = t1 >= t0
            leaf = hit & (lfirst[nd] >= 0)
            nn = np.where(hit & (lfirst[nd] < 0), nd + 1, skip[nd])
            nn = np.where(leaf, nd, nn)                  # a lane at a hit leaf parks on it
            node[wi, li] = nn
            state[wi, li] = np.where(leaf, ATLEAF, np.where(nn >= 0, WALK, DONE))
            uw = np.unique(wi)
            cost[uw] += NODE + LOOP + VOTES
            n_iter += len(uw)
    return int(cost.sum()), n_iter, n_leaf, n_refill


def main():
    print("\n# Lane refill in the bounce waves: simulated cost (tools/refill_model.py, CPU)\n")
    print("`today` = 32-ray groups walked in lockstep, leaf test nested in the box loop, shading once per group "
          f"(box {NODE}+{LOOP}, leaf {LEAF}, load+shade+push {SHADE} warp-instructions). `refill(tl, td)` = persistent lanes: "
          f"leaf tests run when >= tl lanes stand on a leaf, shading + refill when >= td lanes are idle (+{VOTES} "
          "instructions of votes per box step), every warp working through a share of 1024 queued rays. Front-to-back arrays for the Cornell box, reference order for the built-in scene, queue order unless stated, 480x272 primary rays.\n")
    rng = np.random.default_rng(7)
    for wl in ("built-in, default pose", "Cornell box (C3)"):
        sname, pose, fov = bc.WORKLOADS[wl]
        scene = rv.builtin_scene() if sname == "builtin" else rv.cornell_scene()
        nodes, perm = rv.build_bvh(scene.triangles)
        tris = np.ascontiguousarray(scene.triangles[perm])
        ftb = bc.front_to_back_layouts(nodes, tris)
        lay = ftb if sname == "cornell" else [bc.reference_layout(nodes)] * 8
        o, d = bc.camera_rays(480, 272, pose, fov, rng)
        best, _, bn, _ = bc.walk(o, d, ftb, tris)
        print(f"## {wl}\n")
        print("| wave | rays | today (M) | " + " | ".join(f"refill({a},{b}) (M)" for a, b in
                                                             THRESHOLDS) + " | today, octant-sorted (M) | refill(4,24), sorted (M) |")
        print("|---|---|---|" + "---|" * (len(THRESHOLDS) + 2))
        for wave in (1, 2):
            o, d = bc.lambert_bounce(o, d, best, bn, rng)
            if len(o) < 2000:
                break
            best, _, bn, st = bc.walk(o, d, lay, tris)
            groups = (len(o) + 31) // 32
            today = (NODE + LOOP) * st["iters"].sum() + LEAF * st["leaf"].sum() + SHADE * groups
            cells = []
            for tl, td in THRESHOLDS:
                c, _, _, _ = simulate(o, d, lay, tris, 1024, tl, td)
                cells.append(f"{c / 1e6:.2f}")
            p = np.argsort((d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4, kind="stable")
            _, _, _, st2 = bc.walk(o[p], d[p], lay, tris)
            today_sorted = (NODE + LOOP) * st2["iters"].sum() + LEAF * st2["leaf"].sum() + SHADE * groups
            c2, _, _, _ = simulate(o[p], d[p], lay, tris, 1024, 4, 24)
            print(f"| {wave} | {len(o)} | {today / 1e6:.2f} | " + " | ".join(cells) +
                  f" | {today_sorted / 1e6:.2f} | {c2 / 1e6:.2f} |")
        print()


if __name__ == "__main__":
    main()
