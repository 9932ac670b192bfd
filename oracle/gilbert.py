"""ORACLE (test infrastructure, not product code) -- generalized Hilbert ("Gilbert") curve.

CPU restatement of the space-filling curve the reference uses to reorder video tokens.
Follows the algorithm of /root/reference/wanx/train/special_attentions_local/utils/gilbert3d.py:6-167
(Jakub Cerveny's BSD-2 "gilbert3d") and the permutation tables built from it in
wanx_blocksparseattn.py:102-129 (W) / cogvideo_blocksparseattn.py:110-128 (C).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may import this.

The restatement is vector-based and iterative (explicit work stack, numpy 3-vectors) instead of the
reference's nine-scalar recursive generator; the visiting order it produces is identical and is
pinned by tests/test_oracle_gilbert.py against (a) the reference module when /root/reference is
present and (b) committed golden hashes in tests/golden/gilbert_hashes.json.
"""
from __future__ import annotations

import numpy as np


def _sgn(v):
    return np.sign(v).astype(np.int64)


def _half(v):
    # floor division toward -inf, like Python's // on each component (gilbert3d.py:69-71)
    return np.floor_divide(v, 2)


def gilbert3d_order(width: int, height: int, depth: int) -> np.ndarray:
    """Return int64 array [W*H*D, 3] of (x, y, z) in curve order (gilbert3d.py:6-29)."""
    W, H, D = int(width), int(height), int(depth)
    o = np.zeros(3, np.int64)
    ex = np.array([W, 0, 0], np.int64)
    ey = np.array([0, H, 0], np.int64)
    ez = np.array([0, 0, D], np.int64)
    if W >= H and W >= D:
        root = (o, ex, ey, ez)
    elif H >= W and H >= D:
        root = (o, ey, ex, ez)
    else:
        root = (o, ez, ex, ey)

    out = np.empty((W * H * D, 3), np.int64)
    n = 0
    stack = [root]
    while stack:
        p, a, b, c = stack.pop()
        w, h, d = abs(int(a.sum())), abs(int(b.sum())), abs(int(c.sum()))
        da, db, dc = _sgn(a), _sgn(b), _sgn(c)

        # straight runs (gilbert3d.py:51-67)
        if h == 1 and d == 1:
            out[n:n + w] = p + np.arange(w)[:, None] * da
            n += w
            continue
        if w == 1 and d == 1:
            out[n:n + h] = p + np.arange(h)[:, None] * db
            n += h
            continue
        if w == 1 and h == 1:
            out[n:n + d] = p + np.arange(d)[:, None] * dc
            n += d
            continue

        a2, b2, c2 = _half(a), _half(b), _half(c)
        w2, h2, d2 = abs(int(a2.sum())), abs(int(b2.sum())), abs(int(c2.sum()))
        # prefer even steps (gilbert3d.py:77-85)
        if (w2 % 2) and w > 2:
            a2 = a2 + da
        if (h2 % 2) and h > 2:
            b2 = b2 + db
        if (d2 % 2) and d > 2:
            c2 = c2 + dc

        if 2 * w > 3 * h and 2 * w > 3 * d:
            # wide case: split along a only (gilbert3d.py:88-97)
            kids = [
                (p, a2, b, c),
                (p + a2, a - a2, b, c),
            ]
        elif 3 * h > 4 * d:
            # do not split in d (gilbert3d.py:100-116)
            kids = [
                (p, b2, c, a2),
                (p + b2, a, b - b2, c),
                (p + (a - da) + (b2 - db), -b2, c, -(a - a2)),
            ]
        elif 3 * d > 4 * h:
            # do not split in h (gilbert3d.py:119-135)
            kids = [
                (p, c2, a2, b),
                (p + c2, a, b, c - c2),
                (p + (a - da) + (c2 - dc), -c2, -(a - a2), b),
            ]
        else:
            # regular case: split in all three (gilbert3d.py:138-167)
            kids = [
                (p, b2, c2, a2),
                (p + b2, c, a2, b - b2),
                (p + (b2 - db) + (c - dc), a, -b2, -(c - c2)),
                (p + (a - da) + b2 + (c - dc), -c, -(a - a2), b - b2),
                (p + (a - da) + (b2 - db), -b2, c2, -(a - a2)),
            ]
        stack.extend(reversed(kids))
    assert n == W * H * D
    return out


def gilbert_permutations(width: int, height: int, depth: int):
    """Permutation tables of GilbertRearranger (W:102-129).

    Returns (curve2raster, raster2curve), both int64 [W*H*D]:
      curve2raster[c] = raster index x + W*(y + H*z) of the c-th curve point
                        (the reference's `original_order2gilbert_order`, used by `rearrange`)
      raster2curve[r] = curve position of raster index r
                        (the reference's `gilbert_order2original_order`, used by `reversed_rearrange`)
    """
    xyz = gilbert3d_order(width, height, depth)
    curve2raster = xyz[:, 0] + width * (xyz[:, 1] + height * xyz[:, 2])
    raster2curve = np.empty_like(curve2raster)
    raster2curve[curve2raster] = np.arange(curve2raster.size, dtype=np.int64)
    return curve2raster, raster2curve
