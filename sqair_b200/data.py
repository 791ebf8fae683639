"""Synthetic moving-sprite sequences (no MNIST offline): procedural stroke sprites, bouncing
noisy-acceleration trajectories and max-blending, following the semantics of the reference's
generator (data/create_seq_mnist.py:43-56, data/trajectory.py:118-143, data/template.py:69-104).
Host-side numpy; used by bench.py and examples to produce frames of the right shape and statistics."""
import numpy as np


def _sprites(rng, n_sprites=10, size=28):
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    out = []
    for _ in range(n_sprites):
        pts = rng.uniform(4, size - 4, (rng.integers(4, 7), 2))
        thick = rng.uniform(1.0, 1.6)
        img = np.zeros((size, size))
        for a, b in zip(pts[:-1], pts[1:]):
            d = b - a
            tt = np.clip(((yy - a[0]) * d[0] + (xx - a[1]) * d[1]) / max(float(d @ d), 1e-6), 0, 1)
            img = np.maximum(img, np.exp(-((yy - a[0] - tt * d[0]) ** 2 + (xx - a[1] - tt * d[1]) ** 2) / (2 * thick ** 2)))
        img = img / img.max() * 255.
        img[img < 20] = 0
        ys, xs = np.nonzero(img)
        out.append(img[ys.min():ys.max() + 1, xs.min():xs.max() + 1])
    return out


def moving_sprites(T, B, H, W, n_max, seed=1234, obj_size=28):
    """-> (imgs float32 [T,B,H,W] in [0,1], nums int [B])."""
    rng = np.random.default_rng(seed)
    sprites = _sprites(rng)
    lo, hi = np.zeros(2), np.array([H - obj_size, W - obj_size], dtype=np.float64)
    imgs = np.zeros((T, B, H, W))
    nums = rng.integers(0, n_max + 1, B)
    for b in range(B):
        for _ in range(nums[b]):
            p, v, a = rng.uniform(lo, hi), rng.uniform(-10, 10, 2), rng.uniform(-3, 3, 2)
            tmpl = sprites[rng.integers(len(sprites))]
            for t in range(T):
                if t:
                    p, v, a = p + v, v + a, a + rng.normal(0, .01, 2)
                    for d in range(2):
                        if p[d] < lo[d] or p[d] > hi[d]:
                            p[d] = 2 * (lo[d] if p[d] < lo[d] else hi[d]) - p[d]
                            v[d], a[d] = -v[d], -a[d]
                    p, v, a = np.clip(p, lo, hi), np.clip(v, -10, 10), np.clip(a, -3, 3)
                y0, x0 = int(np.round(p[0])), int(np.round(p[1]))
                ys, xs, ye, xe = max(y0, 0), max(x0, 0), min(y0 + tmpl.shape[0], H), min(x0 + tmpl.shape[1], W)
                if ye > ys and xe > xs:
                    imgs[t, b, ys:ye, xs:xe] = np.maximum(imgs[t, b, ys:ye, xs:xe], tmpl[ys - y0:ye - y0, xs - x0:xe - x0])
    return imgs.astype(np.uint8).astype(np.float32) / 255., nums
