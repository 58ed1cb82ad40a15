"""SIMT cost model of the stackless BVH walk, on the CPU (numpy).

Steps the 32 rays of every warp in lockstep through the kernel's walk ("hit -> next record,
miss -> skip link", leaf test nested in the box loop) and counts what a warp executes: box-test
iterations, leaf-test executions and how many lanes are active in each. It reproduces the lane
utilisation ncu measures on the B200 (built-in scene bounce wave: 27.4 vs 27 lanes per box test,
Cornell box: 12.2 vs 12), so layout and scheduling ideas can be screened without a GPU:

    python tools/bvh_cost.py                 # table for the three bench workloads -> stdout (markdown)

Variants per workload: the reference's child order vs the engine's front-to-back octant arrays
(the real ones, from rvpt_b200_octant_layouts), bounce rays in queue order vs sorted by direction
octant, and the cost a perfect lane refill would reach ("ideal": lane-instructions / 32).
Arithmetic is numpy float32 in natural order, not the bit-exact contract: statistics only.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rvpt_b200 as rv  # noqa: E402
from rvpt_b200 import _lib  # noqa: E402

END = -1
# instruction weights of the kernel's blocks (profiles/r01_k_frame_hot_blocks.txt)
NODE_PRIMARY, NODE_BOUNCE, LOOP, LEAF = 16, 22, 6, 45


def reference_layout(nodes):
    """reference node array -> walk-order arrays (bounds[n,6], skip[n], leaf_first[n], leaf_count[n])"""
    b, first, cnt, inner = [], [], [], []
    todo = [0]
    while todo:
        nd = nodes[todo.pop()]
        b.append(nd["bounds"])
        if nd["primitive_count"] > 0:
            first.append(int(nd["first_child_or_primitive"]))
            cnt.append(int(nd["primitive_count"]))
            inner.append(0)
        else:
            first.append(-1)
            cnt.append(0)
            inner.append(1)
            todo.append(int(nd["first_child_or_primitive"]) + 1)
            todo.append(int(nd["first_child_or_primitive"]))
    n = len(b)
    size = [1] * n
    for k in range(n - 1, -1, -1):
        if inner[k]:
            c0 = k + 1
            size[k] = 1 + size[c0] + size[c0 + size[c0]]
    skip = np.array([k + size[k] if k + size[k] < n else END for k in range(n)], np.int64)
    return np.array(b, np.float32), skip, np.array(first), np.array(cnt)


def front_to_back_layouts(nodes, tris):
    """the engine's eight per-octant arrays (engine.cu::build_octant_layouts) in the same form"""
    lib = _lib.load()
    n = C.c_size_t(0)
    lib.rvpt_b200_octant_layouts(nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris), None, 0, C.byref(n))
    out = np.zeros(n.value * 64, np.float32)
    rc = lib.rvpt_b200_octant_layouts(nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris),
                                      out.ctypes.data, out.size, C.byref(n))
    assert rc == 0
    A = out[: n.value * 32].reshape(8, n.value, 4)
    B = out[n.value * 32:].reshape(8, n.value, 4)
    # DevTri ranges: leaves of the reference layout are stored in walk order, one record per triangle
    _, _, rfirst, rcnt = reference_layout(nodes)
    dev_first = np.cumsum(np.where(rfirst >= 0, rcnt, 0)) - np.where(rfirst >= 0, rcnt, 0)
    count_of = {int(dev_first[i]): (int(rfirst[i]), int(rcnt[i])) for i in range(len(rfirst)) if rfirst[i] >= 0}
    layouts = []
    for k in range(8):
        sx, sy, sz = k & 1, (k >> 1) & 1, (k >> 2) & 1
        pick = lambda near, far, s, lo: np.where(bool(s) == lo, far, near)  # noqa: E731
        bounds = np.stack([pick(A[k, :, 0], A[k, :, 1], sx, True), pick(A[k, :, 0], A[k, :, 1], sx, False),
                           pick(A[k, :, 2], A[k, :, 3], sy, True), pick(A[k, :, 2], A[k, :, 3], sy, False),
                           pick(B[k, :, 0], B[k, :, 1], sz, True), pick(B[k, :, 0], B[k, :, 1], sz, False)], 1)
        skip_u = B[k, :, 2].copy().view(np.uint32).astype(np.int64)
        skip = np.where(skip_u == 0xFFFFFFFF, END, skip_u)
        leaf_u = B[k, :, 3].copy().view(np.uint32).astype(np.int64)
        first = np.array([count_of[int(v)][0] if not (v & 0x80000000) else -1 for v in leaf_u])
        cnt = np.array([count_of[int(v)][1] if not (v & 0x80000000) else 0 for v in leaf_u])
        layouts.append((bounds.astype(np.float32), skip, first, cnt))
    return layouts


def tri_hit(o, d, tris, idx, best):
    v0, v1, v2 = tris["vertex0"][idx, :3], tris["vertex1"][idx, :3], tris["vertex2"][idx, :3]
    e0, e1 = v1 - v0, v2 - v0
    n = np.cross(e0, e1)
    with np.errstate(all="ignore"):
        t = np.einsum("ij,ij->i", v0 - o, n) / np.einsum("ij,ij->i", d, n)
        p0 = o + t[:, None] * d - v0
        b0, b1 = np.einsum("ij,ij->i", p0, e0), np.einsum("ij,ij->i", p0, e1)
        g11, g01, g00 = (np.einsum("ij,ij->i", e1, e1), np.einsum("ij,ij->i", e0, e1),
                         np.einsum("ij,ij->i", e0, e0))
        inv = 1.0 / (g11 * g00 - g01 * g01)
        u = inv * (g11 * b0 - g01 * b1)
        v = inv * (-g01 * b0 + g00 * b1)
    ok = (t > 0) & (t < best) & (u > 0) & (v > 0) & (u + v < 1)
    return ok, t, n


def walk(o, d, layouts, tris):
    """Lockstep walk; ray i uses layouts[octant(d_i)] (pass the same layout eight times for one
    order). Warp = 32 consecutive rays. Returns nearest t, hit triangle, its normal and counters."""
    offs, Bs, Ss, Fs, Cs = [0], [], [], [], []
    for b, s, f, c in layouts:
        Bs.append(b)
        Ss.append(np.where(s >= 0, s + offs[-1], END))
        Fs.append(f)
        Cs.append(c)
        offs.append(offs[-1] + len(b))
    bounds, skip, lfirst, lcnt = np.concatenate(Bs), np.concatenate(Ss), np.concatenate(Fs), np.concatenate(Cs)
    octant = (d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4
    R = len(o)
    node = np.array(offs[:8])[octant].astype(np.int64)
    best = np.full(R, np.inf, np.float32)
    btri = np.full(R, -1, np.int64)
    bn = np.zeros((R, 3), np.float32)
    W = (R + 31) // 32
    st = dict(iters=np.zeros(W, np.int64), node_lanes=np.zeros(W, np.int64), leaf=np.zeros(W, np.int64),
              leaf_lanes=np.zeros(W, np.int64), nodes_per_ray=np.zeros(R, np.int64),
              tris_per_ray=np.zeros(R, np.int64))
    with np.errstate(all="ignore"):
        inv = (1.0 / d).astype(np.float32)
    act = node >= 0
    while act.any():
        idx = np.nonzero(act)[0]
        nd = node[idx]
        b = bounds[nd]
        with np.errstate(all="ignore"):
            tx0, tx1 = (b[:, 0] - o[idx, 0]) * inv[idx, 0], (b[:, 1] - o[idx, 0]) * inv[idx, 0]
            ty0, ty1 = (b[:, 2] - o[idx, 1]) * inv[idx, 1], (b[:, 3] - o[idx, 1]) * inv[idx, 1]
            tz0, tz1 = (b[:, 4] - o[idx, 2]) * inv[idx, 2], (b[:, 5] - o[idx, 2]) * inv[idx, 2]
            t0 = np.fmax(np.fmax(np.fmin(tx0, tx1), np.fmin(ty0, ty1)), np.fmax(np.fmin(tz0, tz1), 0))
            t1 = np.fmin(np.fmin(np.fmax(tx0, tx1), np.fmax(ty0, ty1)), np.fmin(np.fmax(tz0, tz1), best[idx]))
        hit = t1 >= t0
        st["nodes_per_ray"][idx] += 1
        wid = idx // 32
        np.add.at(st["node_lanes"], wid, 1)
        st["iters"][np.unique(wid)] += 1
        leaf = hit & (lfirst[nd] >= 0)
        if leaf.any():
            li = idx[leaf]
            maxc = lcnt[node[li]]
            for k in range(int(maxc.max())):
                sel = li[k < maxc]
                tri_idx = lfirst[node[sel]] + k
                ok, t, nn = tri_hit(o[sel], d[sel], tris, tri_idx, best[sel])
                st["tris_per_ray"][sel] += 1
                best[sel[ok]] = t[ok]
                btri[sel[ok]] = tri_idx[ok]
                bn[sel[ok]] = nn[ok]
                w = sel // 32
                st["leaf"][np.unique(w)] += 1
                np.add.at(st["leaf_lanes"], w, 1)
        node[idx] = np.where(hit & (lfirst[nd] < 0), nd + 1, skip[nd])
        act = node >= 0
    return best, btri, bn, st


def camera_rays(W, H, pose, fov_deg, rng):
    """jittered pinhole rays in the kernel's chunk order (16x16 tiles, 8x4 pixels per warp)"""
    ys, xs = np.mgrid[0:H, 0:W]
    key = ((((ys // 16) * (W // 16) + xs // 16) * 8 + ((ys % 16) // 4) * 2 + (xs % 16) // 8) * 32
           + (ys % 4) * 8 + (xs % 8)).ravel()
    order = np.argsort(key)
    xs, ys = xs.ravel()[order], ys.ravel()[order]
    cx = (xs + rng.random(len(xs))) / W
    cy = 1.0 - (ys + rng.random(len(xs))) / H
    u, v, w = (W / H) * (2 * cx - 1), 2 * cy - 1, 1.0 / np.tan(0.5 * np.radians(fov_deg))
    d = np.stack([u, v, np.full_like(u, w)], 1).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.broadcast_to(np.array(pose, np.float32), d.shape).copy(), d


def lambert_bounce(o, d, best, bn, rng):
    hit = np.isfinite(best)
    n = bn[hit] / np.linalg.norm(bn[hit], axis=1, keepdims=True)
    dn = d[hit] / np.linalg.norm(d[hit], axis=1, keepdims=True)
    n[np.einsum("ij,ij->i", dn, n) > 0] *= -1
    pos = o[hit] + best[hit][:, None] * d[hit]
    uu, vv = rng.random(len(n)), rng.random(len(n))
    phi, c = 2 * np.pi * uu, 1 - 2 * vv
    s = np.sqrt(np.maximum(0, 1 - c * c))
    return ((pos + 0.005 * n).astype(np.float32),
            (n + np.stack([s * np.cos(phi), s * np.sin(phi), c], 1)).astype(np.float32))


WORKLOADS = {
    "built-in, default pose": ("builtin", (0.0, 0.0, 0.0), 90.0),
    "built-in, pose (0, 0.8, -2.5)": ("builtin", (0.0, 0.8, -2.5), 90.0),
    "Cornell box (C3)": ("cornell", (0.0, 1.2, -3.4), 60.0),
}


def evaluate(workload, primary_ftb, bounce_ftb, sort_bounce, W=480, H=272, waves=4):
    sname, pose, fov = WORKLOADS[workload]
    scene = rv.builtin_scene() if sname == "builtin" else rv.cornell_scene()
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = np.ascontiguousarray(scene.triangles[perm])
    ref = [reference_layout(nodes)] * 8
    ftb = front_to_back_layouts(nodes, tris)
    rng = np.random.default_rng(7)
    o, d = camera_rays(W, H, pose, fov, rng)
    rows, hits = [], []
    for wave in range(waves):
        if wave > 0 and sort_bounce:
            p = np.argsort((d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4, kind="stable")
            o, d = o[p], d[p]
        lay = ftb if (primary_ftb if wave == 0 else bounce_ftb) else ref
        best, btri, bn, st = walk(o, d, lay, tris)
        nc = (NODE_PRIMARY if wave == 0 else NODE_BOUNCE) + LOOP
        cost = nc * st["iters"].sum() + LEAF * st["leaf"].sum()
        ideal = (nc * st["node_lanes"].sum() + LEAF * st["leaf_lanes"].sum()) / 32
        rows.append(dict(wave=wave, rays=len(o), nodes=st["nodes_per_ray"].mean(), tris=st["tris_per_ray"].mean(),
                         node_lanes=st["node_lanes"].sum() / max(st["iters"].sum(), 1),
                         leaf_lanes=st["leaf_lanes"].sum() / max(st["leaf"].sum(), 1), cost=cost, ideal=ideal))
        hits.append(btri)
        if not np.isfinite(best).any():
            break
        o, d = lambert_bounce(o, d, best, bn, rng)
    return rows, hits


def main():
    print("# SIMT cost model of the BVH walk (tools/bvh_cost.py, CPU, 480x272 rays per wave 0)\n")
    print("Warp-instructions a wave's traversal costs (box test 16/22 + 6 loop, leaf test 45), lanes active per box / "
          "leaf test, and the cost at perfect lane utilisation (what lane refill could approach). `ftb` = the "
          "engine's front-to-back octant arrays, `ref` = the reference's child order; `sorted` = bounce rays "
          "grouped by direction octant before they are dealt to warps.\n")
    for wl in WORKLOADS:
        print(f"## {wl}\n")
        print("| primary / bounce order | bounce rays | wave | rays | box tests / ray | tri tests / ray | lanes box / leaf | "
              "cost (M) | ideal (M) |")
        print("|---|---|---|---|---|---|---|---|---|")
        for pf, bf, srt in ((False, False, False), (True, False, False), (True, True, False), (True, False, True),
                            (True, True, True)):
            rows, _ = evaluate(wl, pf, bf, srt)
            for r in rows:
                if r["rays"] < 1000:
                    continue
                print(f"| {'ftb' if pf else 'ref'} / {'ftb' if bf else 'ref'} | {'sorted' if srt else 'queue order'} | "
                      f"{r['wave']} | {r['rays']} | {r['nodes']:.2f} | {r['tris']:.2f} | {r['node_lanes']:.1f} / "
                      f"{r['leaf_lanes']:.1f} | {r['cost'] / 1e6:.2f} | {r['ideal'] / 1e6:.2f} |")
        print()


if __name__ == "__main__":
    main()
