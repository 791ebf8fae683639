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


def moving_sprites(T, B, H, W, n_max, seed=1234, obj_size=28, return_tracks=False):
    """-> (imgs float32 [T,B,H,W] in [0,1], nums int [B]); with `return_tracks` also coords float64 [T,B,n_max,4] =
    (y, x, height, width) of every object and frame, zeros for absent objects (create_seq_mnist.py:65-83), and labels
    uint8 [B, n_max] (sprite index)."""
    rng = np.random.default_rng(seed)
    sprites = _sprites(rng)
    lo, hi = np.zeros(2), np.array([H - obj_size, W - obj_size], dtype=np.float64)
    imgs = np.zeros((T, B, H, W))
    nums = rng.integers(0, n_max + 1, B)
    coords = np.zeros((T, B, max(n_max, 1), 4))
    labels = np.zeros((B, max(n_max, 1)), dtype=np.uint8)
    for b in range(B):
        for j in range(nums[b]):
            p, v, a = rng.uniform(lo, hi), rng.uniform(-10, 10, 2), rng.uniform(-3, 3, 2)
            labels[b, j] = rng.integers(len(sprites))
            tmpl = sprites[labels[b, j]]
            for t in range(T):
                if t:
                    p, v, a = p + v, v + a, a + rng.normal(0, .01, 2)
                    for d in range(2):
                        if p[d] < lo[d] or p[d] > hi[d]:
                            p[d] = 2 * (lo[d] if p[d] < lo[d] else hi[d]) - p[d]
                            v[d], a[d] = -v[d], -a[d]
                    p, v, a = np.clip(p, lo, hi), np.clip(v, -10, 10), np.clip(a, -3, 3)
                y0, x0 = int(np.round(p[0])), int(np.round(p[1]))
                coords[t, b, j] = (p[0], p[1], tmpl.shape[0], tmpl.shape[1])
                ys, xs, ye, xe = max(y0, 0), max(x0, 0), min(y0 + tmpl.shape[0], H), min(x0 + tmpl.shape[1], W)
                if ye > ys and xe > xs:
                    imgs[t, b, ys:ye, xs:xe] = np.maximum(imgs[t, b, ys:ye, xs:xe], tmpl[ys - y0:ye - y0, xs - x0:xe - x0])
    imgs = imgs.astype(np.uint8).astype(np.float32) / 255.
    return (imgs, nums, coords, labels) if return_tracks else (imgs, nums)


# ---------------------------------------------------------------------------------------------
# the reference's dataset files (data/create_seq_mnist.py:126-131; data/data.py:189-201)
# ---------------------------------------------------------------------------------------------
def make_dataset(n_samples, n_timesteps=10, canvas_size=(50, 50), n_objects=2, seed=1234):
    """A dataset dictionary in the layout `create_seq_mnist.py` pickles: imgs uint8 [T,N,H,W], labels uint8 [N,n_max],
    nums uint8 [1,N,n_max+1] (unary count: the first n entries are 1, data.py:170-174 transposed by fix_data),
    coords float64 [T,N,n_max,4].  Sprites stand in for the MNIST digits (not available offline)."""
    imgs, nums, coords, labels = moving_sprites(n_timesteps, n_samples, canvas_size[0], canvas_size[1], n_objects, seed=seed,
                                                return_tracks=True)
    unary = np.zeros((1, n_samples, n_objects + 1), dtype=np.uint8)
    for i, n in enumerate(nums):
        unary[0, i, :n] = 1
    return dict(imgs=np.round(imgs * 255.).astype(np.uint8), labels=labels, nums=unary, coords=coords)


def save_data(data, path):
    import pickle
    with open(path, 'wb') as f:
        pickle.dump({k: v for k, v in data.items()}, f, protocol=2)        # protocol 2: readable by the reference's Python 2


def load_data(path, data_path=None):
    """data/data.py:189-201: unpickle, imgs -> float32 in [0, 1], nums -> float32.  Files written by the reference
    (Python 2 cPickle) load with latin-1 decoding."""
    import os
    import pickle
    if data_path is not None:
        path = os.path.join(data_path, path)
    with open(path, 'rb') as f:
        try:
            data = pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            data = pickle.load(f, encoding='latin1')
    data = {(k.decode() if isinstance(k, bytes) else k): v for k, v in data.items()}
    data['imgs'] = data['imgs'].astype(np.float32) / 255.
    data['nums'] = data['nums'].astype(np.float32)
    return data


class Batcher(object):
    """data/data.py:204-240 `tensors_from_data`: every call draws one minibatch along the per-key batch axes --
    `np.random.choice(n, batch_size)` (with replacement) when shuffling, else consecutive windows cycling through the
    data.  Returns numpy arrays; `device_batch` stages them through pinned memory."""

    def __init__(self, data_dict, batch_size, axes=None, shuffle=False, seed=None):
        self.data = {k: v for k, v in data_dict.items() if isinstance(v, np.ndarray)}
        self.keys = list(self.data.keys())
        self.axes = axes if axes is not None else {k: 0 for k in self.keys}
        self.batch_size, self.shuffle = batch_size, shuffle
        self.n_entries = self.data[self.keys[0]].shape[self.axes[self.keys[0]]]
        self._rng = np.random.RandomState(seed)
        self._start = 0

    def next_indices(self):
        if self.shuffle:
            return self._rng.choice(self.n_entries, self.batch_size)
        starts = range(0, self.n_entries - self.batch_size + 1, self.batch_size)
        start = starts[self._start % len(starts)]
        self._start += 1
        return np.arange(start, start + self.batch_size)

    def __call__(self):
        idx = self.next_indices()
        return {k: self.data[k].take(idx, self.axes[k]) for k in self.keys}


# ---------------------------------------------------------------------------------------------
# device-side renderer (C ABI: sqair_render_sprites)
# ---------------------------------------------------------------------------------------------
def sprite_atlas(seed=1234, cell=28):
    """The sprite set of `moving_sprites(seed=seed)` as an atlas: uint8 [S, cell, cell] (top-left aligned) + int32 [S, 2]."""
    sprites = _sprites(np.random.default_rng(seed))
    atlas = np.zeros((len(sprites), cell, cell), dtype=np.uint8)
    hw = np.zeros((len(sprites), 2), dtype=np.int32)
    for i, s in enumerate(sprites):
        atlas[i, :s.shape[0], :s.shape[1]] = s.astype(np.uint8)
        hw[i] = s.shape
    return atlas, hw


def render_on_device(coords, labels, nums, H, W, device, atlas=None, seed=1234):
    """Frames [T,B,H,W] float32 on `device` from object tracks: coords [T,B,n,4] (y, x, h, w), labels [B,n] sprite ids,
    nums [B] object counts -- the same pixels `moving_sprites` renders on the host, without a host-to-device copy of
    the frames (only the few-KB tracks travel)."""
    import ctypes as C
    import torch
    from . import _capi
    atl, hw = atlas if atlas is not None else sprite_atlas(seed)
    T, B, n = coords.shape[:3]
    pos = np.round(coords[..., :2]).astype(np.int32)
    sp = np.where(np.arange(n)[None, :] < np.asarray(nums)[:, None], labels.astype(np.int32), -1).astype(np.int32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    atl_d, hw_d, pos_d, sp_d = d(atl), d(hw), d(pos), d(sp)
    frames = torch.empty(T, B, H, W, dtype=torch.float32, device=device)
    p = lambda t: C.c_void_p(t.data_ptr())
    _capi.check(_capi.lib().sqair_render_sprites(p(atl_d), p(hw_d), p(pos_d), p(sp_d), p(frames), T, B, n, H, W, atl.shape[0], atl.shape[1],
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return frames
