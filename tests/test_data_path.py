"""Data path: the reference's dataset file format (data/create_seq_mnist.py:65-87,126-131), loader
(data/data.py:189-240), `mnist_tools.load` and the sequence-length curriculum (data/mnist_tools.py:62-108)."""
import os
import pickle

import numpy as np
import pytest

from sqair_b200 import data, mnist_tools
from sqair_b200 import tf_flags as flags
from sqair_b200.configs import seq_mnist_data as data_config


def test_dataset_layout_matches_the_reference_files(tmp_path):
    ds = data.make_dataset(12, n_timesteps=5, canvas_size=(50, 50), n_objects=2, seed=3)
    assert ds['imgs'].shape == (5, 12, 50, 50) and ds['imgs'].dtype == np.uint8
    assert ds['nums'].shape == (1, 12, 3) and ds['coords'].shape == (5, 12, 2, 4) and ds['labels'].shape == (12, 2)
    n = ds['nums'][0].astype(int).sum(-1)                                 # unary count: first n entries are 1 (data.py:170-174)
    for i in range(12):
        assert list(ds['nums'][0, i]) == [1] * n[i] + [0] * (3 - n[i])
        assert (ds['coords'][:, i, n[i]:] == 0).all()                     # absent objects: zeros (create_seq_mnist.py:78-83)
        assert (ds['coords'][:, i, :n[i], 2:] > 0).all()                  # (height, width) of the template
        assert (ds['imgs'][:, i].reshape(5, -1).max(-1) > 0).all() == (n[i] > 0)
    path = str(tmp_path / 'seq_mnist_train.pickle')
    data.save_data(ds, path)
    raw = pickle.load(open(path, 'rb'))
    assert set(raw) == {'imgs', 'labels', 'nums', 'coords'}
    back = data.load_data(path)
    assert back['imgs'].dtype == np.float32 and back['imgs'].max() <= 1.0 and back['nums'].dtype == np.float32
    np.testing.assert_array_equal(np.round(back['imgs'] * 255).astype(np.uint8), ds['imgs'])
    # a file as the reference's Python 2 writes it (str keys pickled as bytes) loads too
    py2 = {k.encode(): v for k, v in ds.items()}
    pickle.dump(py2, open(str(tmp_path / 'py2.pickle'), 'wb'), protocol=2)
    assert set(data.load_data(str(tmp_path / 'py2.pickle'))) == {'imgs', 'labels', 'nums', 'coords'}


def test_batcher_semantics():
    d = dict(imgs=np.arange(3 * 10 * 2).reshape(3, 10, 2).astype(np.float32), labels=np.arange(10)[:, None])
    b = data.Batcher(d, 4, dict(imgs=1, labels=0), shuffle=False)
    firsts = [int(b()['labels'][0, 0]) for _ in range(5)]
    assert firsts == [0, 4, 0, 4, 0]                                       # windows 0..3, 4..7, then it cycles (data.py:216-221)
    mb = b()
    assert mb['imgs'].shape == (3, 4, 2) and np.array_equal(mb['imgs'][:, :, 0] // 2 % 10, np.tile(mb['labels'][:, 0], (3, 1)))
    s = data.Batcher(d, 64, dict(imgs=1, labels=0), shuffle=True, seed=0)
    lab = s()['labels'][:, 0]
    assert lab.shape == (64,) and len(set(lab.tolist())) <= 10 and lab.max() <= 9     # drawn with replacement (data.py:213)


def test_load_and_curriculum(tmp_path):
    for part, n, seed in (('train', 40, 1), ('validation', 16, 2)):
        data.save_data(data.make_dataset(n, n_timesteps=6, n_objects=2, seed=seed), str(tmp_path / ('seq_mnist_%s.pickle' % part)))
    F = flags.FLAGS
    F.train_path, F.valid_path = str(tmp_path / 'seq_mnist_train.pickle'), str(tmp_path / 'seq_mnist_validation.pickle')
    F.seq_len, F.stage_itr = 3, 100
    dd = data_config.load(8, seed=0)
    assert dd.train_data['imgs'].shape == (6, 40, 50, 50) and dd.axes == mnist_tools.axes
    assert dd.train_img(0).shape == (3, 8, 50, 50)                         # seq_len at step 0
    assert dd.train_img(250).shape == (5, 8, 50, 50)                       # + global_step // stage_itr
    assert dd.train_img(10 ** 6).shape == (6, 8, 50, 50)                   # capped at the data length
    mb = dd.valid_tensors.next(0)
    assert mb['nums'].shape == (3, 8, 3) and mb['coords'].shape == (3, 8, 3, 4)        # nums tiled over time, coords padded to n+1 slots
    assert mnist_tools.stage_seq_len(99, 3, 100, 10) == 3 and mnist_tools.stage_seq_len(100, 3, 100, 10) == 4
    assert mnist_tools.stage_seq_len(5, 0, 100, 10) == 10 and mnist_tools.stage_seq_len(5, 3, 0, 10) == 10
    # no curriculum: seq_len truncates the data itself (mnist_tools.py:69-73)
    F.stage_itr = 0
    dd = data_config.load(4)
    assert dd.train_data['imgs'].shape[0] == 3 and dd.valid_img().shape == (3, 4, 50, 50)
    F.seq_len = 0
