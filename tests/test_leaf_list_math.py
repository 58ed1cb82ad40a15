"""The interval arithmetic behind the batched primary wave's leaf lists (kernels.cu,
build_leaf_list), restated in numpy float32 and checked on the CPU against per-ray slab tests.

The device code lists, per 8x4 pixel block, the leaves whose boxes ANY primary ray of the block
can enter, and the block's rays then test only those boxes. The result is the tree walk's only if
the list is COMPLETE: no ray of the block may pass the slab test (intersection.glsl:327-357, the
walk's arithmetic: (bound - origin) * invdir, every operation rounded to float32) of a leaf that
the beam test dropped. This file checks exactly that property for the formulas the kernel uses —
direction bounds from the affine pre-normalisation direction (camera.glsl:41-47), |d| intervals
per axis, entry / exit bounds widened by 1e-5 of their scale — on the built-in scene and the
Cornell box, over poses inside, outside and far from the scene, for jitters that include the
corners 0 and 1 (rand() may return 1.0, util.glsl:49). It is a check of the ARGUMENT; the device
code itself is held to the oracle bit for bit by tests/test_gpu_parity.py
(test_leaf_lists_of_the_batched_primary_wave and every batched test).
"""
import numpy as np
import pytest

f32 = np.float32


def _camera_rays(cam, W, H, px, py, jx, jy):
    """camera_pinhole_ray for pixels (px, py) + jitter, float32 op by op (compute_pass.comp:153-154,
    camera.glsl:41-47; summation order of include/rvpt_math.h)."""
    M = cam[:16].astype(f32)
    aspect, hfov = f32(cam[16]), f32(cam[17])
    w = f32(1.0) / np.tan(f32(0.5) * hfov, dtype=f32)
    inv_x, inv_y = f32(1.0) / f32(W), f32(1.0) / f32(H)
    cx = (px.astype(f32) + jx) * inv_x
    cy = f32(1.0) - (py.astype(f32) + jy) * inv_y
    u = aspect * ((cx + cx) - f32(1.0))
    v = (cy + cy) - f32(1.0)
    d = []
    for i in range(3):
        acc = (M[i] * u + M[4 + i] * v) + M[8 + i] * w
        acc = acc + M[12 + i] * f32(0.0)
        d.append(acc.astype(f32))
    dot = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
    inv = f32(1.0) / np.sqrt(dot, dtype=f32)
    return [c * inv for c in d]


def _ray_passes(lo, hi, d):
    """intersect_aabb with mint = 0, maxt = INF for boxes (lo, hi) relative to the origin: (n_leaves, n_rays)."""
    t0 = np.zeros((lo.shape[0], d[0].shape[0]), f32)
    t1 = np.full_like(t0, np.inf)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for a in range(3):
            inv = (f32(1.0) / d[a])[None, :]
            ta, tb = lo[:, a:a + 1] * inv, hi[:, a:a + 1] * inv
            t0 = np.fmax(t0, np.fmin(ta, tb))
            t1 = np.fmin(t1, np.fmax(ta, tb))
    return t1 >= t0


def _beam_lists(cam, W, H, x0, y0, lo, hi):
    """build_leaf_list for blocks with corner (x0, y0): returns (usable, pass[n_blocks, n_leaves])."""
    M = cam[:16].astype(f32)
    aspect, hfov = f32(cam[16]), f32(cam[17])
    w = f32(1.0) / np.tan(f32(0.5) * hfov, dtype=f32)
    inv_x, inv_y = f32(1.0) / f32(W), f32(1.0) / f32(H)
    pad = f32(0.015625)
    cxa, cxb = (x0.astype(f32) - pad) * inv_x, (x0.astype(f32) + (f32(8.0) + pad)) * inv_x
    cya = f32(1.0) - (y0.astype(f32) - pad) * inv_y
    cyb = f32(1.0) - (y0.astype(f32) + (f32(4.0) + pad)) * inv_y
    ua, ub = aspect * ((cxa + cxa) - f32(1.0)), aspect * ((cxb + cxb) - f32(1.0))
    va, vb = (cya + cya) - f32(1.0), (cyb + cyb) - f32(1.0)
    dlo, dhi, L = [], [], np.zeros_like(ua)
    for i in range(3):
        a, b, c, e = M[i] * ua, M[i] * ub, M[4 + i] * va, M[4 + i] * vb
        wz = M[8 + i] * w
        dlo.append((np.fmin(a, b) + np.fmin(c, e)) + wz)
        dhi.append((np.fmax(a, b) + np.fmax(c, e)) + wz)
        L = L + np.fmax(np.abs(dlo[i]), np.abs(dhi[i]))
    usable = (L > 0) & (L < f32(1e30))
    widen, apart = f32(1e-5) * L, f32(1e-4) * L
    R = f32(np.max(np.abs(np.concatenate([lo[0], hi[0]]))))  # the root's record
    eps = f32(1e-5) * R
    t0 = np.zeros((x0.shape[0], lo.shape[0]), f32)
    t1 = np.full_like(t0, np.inf)
    for i in range(3):
        l, h = dlo[i] - widen, dhi[i] + widen
        neg = h < -apart
        usable &= neg | (l > apart)
        alo, ahi = np.where(neg, -h, l), np.where(neg, -l, h)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv_lo = (f32(1.0) / ahi) * f32(0.99999)
            inv_hi = (f32(1.0) / alo) * f32(1.00001)
        near = np.where(neg[:, None], -hi[None, :, i], lo[None, :, i]) - eps
        far = np.where(neg[:, None], -lo[None, :, i], hi[None, :, i]) + eps
        with np.errstate(invalid="ignore", over="ignore"):
            t0 = np.fmax(t0, np.fmin(near * inv_lo[:, None], near * inv_hi[:, None]))
            t1 = np.fmin(t1, np.fmax(far * inv_lo[:, None], far * inv_hi[:, None]))
    return usable, ~(t1 < t0)


POSES = [  # scene, translation, rotation, fov
    ("builtin", (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 90.0),
    ("builtin", (0.0, 0.8, -2.5), (0.0, 0.0, 0.0), 90.0),
    ("builtin", (0.6, 0.9, -1.2), (0.3, -0.5, 0.2), 120.0),
    ("builtin", (0.0, 1.0, -14.0), (0.0, 0.0, 0.0), 30.0),
    ("cornell", (0.0, 1.2, -3.4), (0.0, 0.0, 0.0), 60.0),
    ("cornell", (0.95, 0.05, -0.9), (0.02, -0.1, 0.0), 100.0),
    ("cornell", (0.0, 1.0, 0.0), (1.3, 0.4, 0.0), 150.0),
]


@pytest.mark.parametrize("scene_name,pose,rot,fov", POSES)
def test_beam_lists_are_complete(rv, builtin, cornell, scene_name, pose, rot, fov):
    prep = builtin if scene_name == "builtin" else cornell
    W, H = 640, 360
    cam = rv.camera_data(translation=pose, rotation=rot, aspect=W / H, fov=fov)
    o = cam[12:15].astype(f32)
    leaf = prep.nodes["primitive_count"] > 0
    b = prep.nodes["bounds"].astype(f32)
    root_lo, root_hi = b[0, 0::2] - o, b[0, 1::2] - o
    lo = np.vstack([root_lo[None], b[leaf][:, 0::2] - o])  # row 0 = the root (scale of eps), then the leaves
    hi = np.vstack([root_hi[None], b[leaf][:, 1::2] - o])
    xs, ys = np.meshgrid(np.arange(0, W, 8), np.arange(0, H, 4))
    x0, y0 = xs.ravel(), ys.ravel()
    usable, listed = _beam_lists(cam, W, H, x0, y0, lo, hi)
    assert usable.mean() > 0.9  # only blocks on a sign change of a direction component fall back
    rng = np.random.default_rng(7)
    missed = 0
    checked = 0
    for _ in range(6):
        # one random pixel of every block, jitter drawn from {0, 1, uniform}: corners and edges included
        px = x0 + rng.integers(0, 8, x0.shape)
        py = y0 + rng.integers(0, 4, y0.shape)
        jx = rng.choice([f32(0.0), f32(1.0), f32(rng.random())], x0.shape).astype(f32)
        jy = rng.choice([f32(0.0), f32(1.0), f32(rng.random())], x0.shape).astype(f32)
        d = _camera_rays(cam, W, H, px, py, jx, jy)
        passes = _ray_passes(lo, hi, d).T  # (n_blocks, n_leaves + 1)
        bad = passes & ~listed & usable[:, None]
        missed += int(bad.sum())
        checked += int((passes & usable[:, None]).sum())
    assert checked > 0 or scene_name == "builtin"
    assert missed == 0, f"{missed} (ray, leaf) pairs pass the slab test of a leaf their block's list lacks"
    # and the lists are worth having: short, mostly empty where there is sky
    n = listed[usable][:, 1:].sum(1)
    assert n.mean() < 12
