"""SIMT cost model of the BVH walk on the CPU (numpy): how many node-test and leaf-test
*warp instructions-slots* a BVH costs for primary and one-bounce rays, stepping the 32 rays of a
warp in lockstep exactly like the kernel's stackless if-if loop does. Used to compare BVH builder
variants without a GPU:   python tools/bvh_cost.py [builtin|cornell] [default|pinned]
Approximate arithmetic (numpy float32, not the bit-exact contract): for statistics only."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rvpt_b200 as rv  # noqa: E402


def preorder(nodes):
    """reference node array -> (bounds[n,6], skip[n], leaf_first[n], leaf_count[n]) in walk order"""
    order, skip_stack = [], []
    out_b, out_first, out_cnt, inner = [], [], [], []
    todo = [0]
    while todo:
        i = todo.pop()
        nd = nodes[i]
        out_b.append(nd["bounds"])
        if nd["primitive_count"] > 0:
            out_first.append(int(nd["first_child_or_primitive"]))
            out_cnt.append(int(nd["primitive_count"]))
            inner.append(0)
        else:
            out_first.append(-1)
            out_cnt.append(0)
            inner.append(1)
            todo.append(int(nd["first_child_or_primitive"]) + 1)
            todo.append(int(nd["first_child_or_primitive"]))
    n = len(out_b)
    size = [1] * n
    for k in range(n - 1, -1, -1):
        if inner[k]:
            c0 = k + 1
            c1 = c0 + size[c0]
            size[k] = 1 + size[c0] + size[c1]
    skip = np.array([k + size[k] if k + size[k] < n else -1 for k in range(n)], np.int64)
    return np.array(out_b, np.float32), skip, np.array(out_first), np.array(out_cnt)


def tri_hit(o, d, tris, idx, best):
    v0, v1, v2 = tris["vertex0"][idx, :3], tris["vertex1"][idx, :3], tris["vertex2"][idx, :3]
    e0, e1 = v1 - v0, v2 - v0
    n = np.cross(e0, e1)
    with np.errstate(all="ignore"):
        t = np.einsum("ij,ij->i", v0 - o, n) / np.einsum("ij,ij->i", d, n)
        p0 = o + t[:, None] * d - v0
        b0, b1 = np.einsum("ij,ij->i", p0, e0), np.einsum("ij,ij->i", p0, e1)
        g11, g01, g00 = np.einsum("ij,ij->i", e1, e1), np.einsum("ij,ij->i", e0, e1), np.einsum("ij,ij->i", e0, e0)
        inv = 1.0 / (g11 * g00 - g01 * g01)
        u = inv * (g11 * b0 - g01 * b1)
        v = inv * (-g01 * b0 + g00 * b1)
    ok = (t > 0) & (t < best) & (u > 0) & (v > 0) & (u + v < 1)
    return ok, t, n


def walk(o, d, bounds, skip, lfirst, lcnt, tris):
    """lockstep walk of all rays; returns (best_t, best_n, per-warp node iterations, per-warp leaf
    executions, per-ray node tests, per-ray triangle tests). Warp = 32 consecutive rays."""
    R = len(o)
    pad = (-R) % 32
    node = np.zeros(R, np.int64)
    best = np.full(R, np.inf, np.float32)
    bn = np.zeros((R, 3), np.float32)
    n_nodes = np.zeros(R, np.int64)
    n_tris = np.zeros(R, np.int64)
    W = (R + pad) // 32
    warp_iters = np.zeros(W, np.int64)
    warp_leaf = np.zeros(W, np.int64)
    warp_leaf_lanes = np.zeros(W, np.int64)
    warp_node_lanes = np.zeros(W, np.int64)
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
        n_nodes[idx] += 1
        wid = idx // 32
        np.add.at(warp_node_lanes, wid, 1)
        warp_iters[np.unique(wid)] += 1
        leaf = hit & (lfirst[nd] >= 0)
        if leaf.any():
            li = idx[leaf]
            # a warp executes the leaf block once per triangle slot while any lane needs it
            maxc = lcnt[node[li]]
            for k in range(int(maxc.max())):
                sel = li[k < maxc]
                tri_idx = lfirst[node[sel]] + k
                ok, t, nn = tri_hit(o[sel], d[sel], tris, tri_idx, best[sel])
                n_tris[sel] += 1
                best[sel[ok]] = t[ok]
                bn[sel[ok]] = nn[ok]
                w = sel // 32
                warp_leaf[np.unique(w)] += 1
                np.add.at(warp_leaf_lanes, w, 1)
        nxt = np.where(hit & (lfirst[nd] < 0), nd + 1, skip[nd])
        node[idx] = nxt
        act = node >= 0
    return best, bn, warp_iters, warp_leaf, n_nodes, n_tris, warp_node_lanes, warp_leaf_lanes


def camera_rays(W, H, pose, fov_deg, rng):
    # 8x4 pixel blocks in the kernel's chunk order (tile 16x16 -> 8 chunks)
    ys, xs = np.mgrid[0:H, 0:W]
    ty, tx = ys // 16, xs // 16
    wy, wx = (ys % 16) // 4, (xs % 16) // 8
    key = (((ty * (W // 16) + tx) * 8 + wy * 2 + wx) * 32 + (ys % 4) * 8 + (xs % 8)).ravel()
    order = np.argsort(key)
    xs, ys = xs.ravel()[order], ys.ravel()[order]
    cx = (xs + rng.random(len(xs))) / W
    cy = 1.0 - (ys + rng.random(len(xs))) / H
    aspect = W / H
    u, v, w = aspect * (2 * cx - 1), 2 * cy - 1, 1.0 / np.tan(0.5 * np.radians(fov_deg))
    d = np.stack([u, v, np.full_like(u, w)], 1).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(np.array(pose, np.float32), d.shape).copy()
    return o, d


def evaluate(scene_name="builtin", pose_name="default", W=480, H=272, verbose=True, nodes_perm=None):
    scene = rv.builtin_scene() if scene_name == "builtin" else rv.cornell_scene()
    pose = {"default": (0.0, 0.0, 0.0), "pinned": (0.0, 0.8, -2.5)}[pose_name]
    fov = 90.0
    if scene_name == "cornell":
        pose, fov = (0.0, 1.2, -3.4), 60.0
    nodes, perm = nodes_perm if nodes_perm else rv.build_bvh(scene.triangles)
    tris = scene.triangles[perm]
    bounds, skip, lfirst, lcnt = preorder(nodes)
    rng = np.random.default_rng(7)
    o, d = camera_rays(W, H, pose, fov, rng)
    res = {}
    total = 0.0
    for wave in range(3):
        best, bn, wi, wl, nn, nt, wnl, wll = walk(o, d, bounds, skip, lfirst, lcnt, tris)
        node_cost = 16 if wave == 0 else 22
        cost = (node_cost + 6) * wi.sum() + 45 * wl.sum()
        res[wave] = dict(rays=len(o), node_per_ray=nn.mean(), tri_per_ray=nt.mean(), warp_iters=int(wi.sum()),
                         warp_leaf=int(wl.sum()), node_lanes=wnl.sum() / max(wi.sum(), 1),
                         leaf_lanes=wll.sum() / max(wl.sum(), 1), cost=cost)
        total += cost
        hit = np.isfinite(best)
        if not hit.any():
            break
        # Lambert bounce off the hit point (compaction keeps the order)
        n = bn[hit]
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        dn = d[hit] / np.linalg.norm(d[hit], axis=1, keepdims=True)
        flip = np.einsum("ij,ij->i", dn, n) > 0
        n[flip] *= -1
        pos = o[hit] + best[hit][:, None] * d[hit]
        uu, vv = rng.random(len(n)), rng.random(len(n))
        phi, c = 2 * np.pi * uu, 1 - 2 * vv
        s = np.sqrt(np.maximum(0, 1 - c * c))
        o = (pos + 0.005 * n).astype(np.float32)
        d = (n + np.stack([s * np.cos(phi), s * np.sin(phi), c], 1)).astype(np.float32)
    if verbose:
        print(f"{scene_name}/{pose_name}: {len(nodes)} nodes, {len(tris)} tris, depth-model cost {total/1e6:.2f} M warp-instr")
        for wv, r in res.items():
            print(f"  wave {wv}: rays {r['rays']:7d} nodes/ray {r['node_per_ray']:.2f} tris/ray {r['tri_per_ray']:.2f} "
                  f"warp node iters {r['warp_iters']:8d} (lanes {r['node_lanes']:.1f}) leaf execs {r['warp_leaf']:8d} "
                  f"(lanes {r['leaf_lanes']:.1f}) cost {r['cost']/1e6:.2f}M")
    return total, res


if __name__ == "__main__":
    sc = sys.argv[1] if len(sys.argv) > 1 else "builtin"
    po = sys.argv[2] if len(sys.argv) > 2 else "default"
    evaluate(sc, po)
