"""Synthetic inputs shared by the oracle, the parity tests and bench.py.

TEST / BENCH INFRASTRUCTURE (lives under oracle/: never imported by `sqair_b200/`).

No MNIST offline, so sequences are procedural "stroke" sprites moving like the reference's
moving-MNIST generator: #objects ~ U{0..n_max} (data/data.py:84), non-overlapping initial
positions with retries (data/data.py:90-107), `NoisyAccelerationTrajectory(noise_std=.01,
max_speed=10, max_acc=3, bounce=True)` dynamics (data/trajectory.py:118-143,
data/create_seq_mnist.py:43-56), `np.maximum` blending at rounded positions
(data/template.py:69-104), uint8 -> float32/255 (data/data.py:199).

Also holds the numpy restatement of the counter-based noise generator (Philox4x32-10 +
Box-Muller) that `sqair_fill_noise` implements on the device, so both sides can draw the same
eps / u tensors from (seed, global row, frame, slot).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------------
# sprites and sequences
# --------------------------------------------------------------------------------------------
def make_sprites(n_sprites=10, size=28, seed=1234):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    sprites = []
    for _ in range(n_sprites):
        npts = rng.integers(4, 7)
        pts = rng.uniform(4, size - 4, (npts, 2))
        thick = rng.uniform(1.0, 1.6)
        img = np.zeros((size, size))
        for a, b in zip(pts[:-1], pts[1:]):
            d = b - a
            tt = np.clip(((yy - a[0]) * d[0] + (xx - a[1]) * d[1]) / max(d @ d, 1e-6), 0, 1)
            dist2 = (yy - (a[0] + tt * d[0])) ** 2 + (xx - (a[1] + tt * d[1])) ** 2
            img = np.maximum(img, np.exp(-dist2 / (2 * thick ** 2)))
        img = img / img.max() * 255.
        img[img < 20] = 0
        ys, xs = np.nonzero(img)
        sprites.append(img[ys.min():ys.max() + 1, xs.min():xs.max() + 1])      # tight bbox, data.py:136-138
    return sprites


def _blend(canvas, tmpl, pos):
    h, w = tmpl.shape
    H, W = canvas.shape
    y0, x0 = int(np.round(pos[0])), int(np.round(pos[1]))
    ys, xs = max(y0, 0), max(x0, 0)
    ye, xe = min(y0 + h, H), min(x0 + w, W)
    if ye <= ys or xe <= xs:
        return
    canvas[ys:ye, xs:xe] = np.maximum(canvas[ys:ye, xs:xe], tmpl[ys - y0:ye - y0, xs - x0:xe - x0])


def make_sequences(T, B, H, W, n_max, seed=1234, obj_size=28):
    """Returns (imgs float32 [T,B,H,W] in [0,1], nums int [B])."""
    rng = np.random.default_rng(seed)
    sprites = make_sprites(seed=seed)
    lo = np.array([0., 0.])
    hi = np.array([H - obj_size, W - obj_size], dtype=np.float64)
    imgs = np.zeros((T, B, H, W), dtype=np.float64)
    nums = rng.integers(0, n_max + 1, B)
    for b in range(B):
        placed = []
        for _ in range(nums[b]):
            pos = None
            for _retry in range(5):
                cand = rng.uniform(lo, hi)
                if all(np.abs(cand - q).max() >= obj_size * 0.5 for q in placed):
                    pos = cand
                    break
            if pos is None:
                pos = rng.uniform(lo, hi)
            placed.append(pos)
            tmpl = sprites[rng.integers(len(sprites))]
            vel = rng.uniform(-10, 10, 2)
            acc = rng.uniform(-3, 3, 2)
            p = pos.copy()
            for t in range(T):
                if t > 0:
                    p = p + vel
                    vel = vel + acc
                    acc = acc + rng.normal(0, .01, 2)
                    for d in range(2):
                        if p[d] < lo[d]:
                            p[d] = 2 * lo[d] - p[d]; vel[d] *= -1; acc[d] *= -1
                        elif p[d] > hi[d]:
                            p[d] = 2 * hi[d] - p[d]; vel[d] *= -1; acc[d] *= -1
                    p = np.clip(p, lo, hi)
                    vel = np.clip(vel, -10, 10)
                    acc = np.clip(acc, -3, 3)
                _blend(imgs[t, b], tmpl, p)
    imgs = imgs.astype(np.uint8).astype(np.float32) / 255.
    return imgs, nums


# --------------------------------------------------------------------------------------------
# counter-based noise: Philox4x32-10 keyed by seed, counter = (row, frame, slot, block)
# --------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counters/keys are uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def _u01(x):            # [0,1), exact in fp32
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def _box_muller(xa, xb):
    u1 = ((xa >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)   # (0,1]
    u2 = _u01(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(2.0 * np.pi) * u2).astype(np.float32)
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def philox_noise(T, rows, n, nw, seed, row_offset=0):
    """eps_where [T,rows,2n,4], eps_what [T,rows,2n,nw], u_pres [T,rows,2n]; identical to the device
    generator for global rows row_offset..row_offset+rows-1 (so shards reproduce the 1-GPU draw).
    Block j of counter word 3: j=0 -> eps_where, j=1.. -> eps_what[4(j-1):4j], j=63 -> u_pres."""
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    t = np.arange(T, dtype=np.uint32)[:, None, None, None]
    r = (np.arange(rows, dtype=np.uint32) + np.uint32(row_offset))[None, :, None, None]
    s = np.arange(2 * n, dtype=np.uint32)[None, None, :, None]
    nblk = (nw + 3) // 4
    j = np.arange(nblk + 1, dtype=np.uint32)[None, None, None, :]
    x0, x1, x2, x3 = philox4x32_10(r, t, s, j, k0, k1)
    z0, z1 = _box_muller(x0, x1)
    z2, z3 = _box_muller(x2, x3)
    z = np.stack((z0, z1, z2, z3), -1)                                   # [T,rows,2n,nblk+1,4]
    eps_where = z[..., 0, :]
    eps_what = z[..., 1:, :].reshape(T, rows, 2 * n, nblk * 4)[..., :nw]
    y0, _, _, _ = philox4x32_10(r, t, s, np.uint32(63), k0, k1)
    u = _u01(y0)[..., 0]
    return dict(eps_where=np.ascontiguousarray(eps_where), eps_what=np.ascontiguousarray(eps_what),
                u_pres=np.ascontiguousarray(u))


def numpy_noise(T, rows, n, nw, seed=7):
    """Plain numpy Generator noise (for fixtures that pre-date / do not need the device generator)."""
    rng = np.random.default_rng(seed)
    return dict(eps_where=rng.standard_normal((T, rows, 2 * n, 4)).astype(np.float32),
                eps_what=rng.standard_normal((T, rows, 2 * n, nw)).astype(np.float32),
                u_pres=rng.random((T, rows, 2 * n)).astype(np.float32))
