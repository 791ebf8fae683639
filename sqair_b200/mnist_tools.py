"""Loading the sequential multi-MNIST files and the sequence-length curriculum (reference: sqair/data/mnist_tools.py:33-108).
`load(batch_size, n_timesteps)` returns the same dictionary of train / validation streams; what TF expresses as graph
tensors fed by a `py_func` (data/data.py:237) are callables here: `data.train_img()` draws the next minibatch."""
import numpy as np

from . import data as _data
from . import tf_flags as flags

flags.DEFINE_integer('seq_len', 0, 'Length of loaded data sequences. If 0, it defaults to the maximum length.')
flags.DEFINE_integer('stage_itr', 0, 'If > 0 it setups a curriculum learning where `seq_len` starts as given and '
                                     'increases by one every `stage_itr` until it gets to the maximum value.')

axes = {'imgs': 1, 'labels': 0, 'nums': 1, 'coords': 1}


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def truncate(data_dict, n_timesteps):
    for k in ('imgs', 'coords', 'nums'):
        data_dict[k] = data_dict[k][:n_timesteps]
    return data_dict


def process_data(data, n_timesteps):
    """mnist_tools.py:49-59: optional truncation; coords padded with zeros up to the slot count of `nums`."""
    if n_timesteps is not None:
        truncate(data, n_timesteps)
    n_steps = data['nums'].shape[-1]
    to_pad = n_steps - data['coords'].shape[-2]
    if to_pad > 0:
        shape = list(data['coords'].shape)
        shape[-2] = to_pad
        data['coords'] = np.concatenate((data['coords'], np.zeros(shape, dtype=data['coords'].dtype)), -2)


def stage_seq_len(global_step, seq_len, stage_itr, n_timesteps):
    """mnist_tools.py:84-87: the curriculum length min(seq_len + global_step // stage_itr, T)."""
    if seq_len == 0 or stage_itr <= 0:
        return n_timesteps
    return int(min(seq_len + int(global_step) // stage_itr, n_timesteps))


class Stream(object):
    """One minibatch source (train: shuffled with replacement, validation: rolling windows).  `next(global_step)` returns
    {imgs [T',B,H,W], nums [T',B,n+1], coords [T',B,n,4], labels [B,n]} truncated to the curriculum length T'."""

    def __init__(self, data, batch_size, shuffle, seq_len, stage_itr, seed=None):
        self._batcher = _data.Batcher(data, batch_size, axes, shuffle=shuffle, seed=seed)
        self._seq_len, self._stage_itr = seq_len, stage_itr
        self.n_timesteps = data['imgs'].shape[0]
        self._tile_nums = data['imgs'].shape[0] != data['nums'].shape[0]

    def next(self, global_step=0):
        mb = self._batcher()
        if self._tile_nums:                                             # mnist_tools.py:79-81
            mb['nums'] = np.tile(mb['nums'], (self.n_timesteps, 1, 1))
        t = stage_seq_len(global_step, self._seq_len, self._stage_itr, self.n_timesteps)
        for k in ('imgs', 'nums', 'coords'):                            # index.dynamic_truncate (index.py:224-241)
            mb[k] = mb[k][:t]
        return mb


def load(batch_size, n_timesteps=None, seed=None):
    F = flags.FLAGS
    valid_data = _data.load_data(F.valid_path)
    train_data = _data.load_data(F.train_path)
    if F.stage_itr == 0 and n_timesteps is None and F.seq_len != 0:
        n_timesteps = F.seq_len
    process_data(valid_data, n_timesteps)
    process_data(train_data, n_timesteps)
    train = Stream(train_data, batch_size, True, F.seq_len, F.stage_itr, seed)
    valid = Stream(valid_data, batch_size, False, F.seq_len, F.stage_itr)
    return AttrDict(train_tensors=train, valid_tensors=valid, train_data=train_data, valid_data=valid_data, axes=axes,
                    train_img=lambda step=0: train.next(step)['imgs'], valid_img=lambda step=0: valid.next(step)['imgs'],
                    train_num=lambda step=0: train.next(step)['nums'], valid_num=lambda step=0: valid.next(step)['nums'],
                    train_coord=lambda step=0: train.next(step)['coords'], valid_coord=lambda step=0: valid.next(step)['coords'])
