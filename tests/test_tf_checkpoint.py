"""TensorFlow tensor-bundle reader / writer (sqair_b200/tf_checkpoint.py): the route by which a checkpoint of the
reference (`tf.train.Saver`, scripts/experiment.py:165-168; released model evaluated in notebooks/play.ipynb:421-480)
gets into the parameter store.  The released checkpoint is not available offline, so the format code is exercised on
bundles it writes itself, plus the published CRC32C test vectors and the structural constants of the format."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from sqair_b200 import _capi, optim, tf_checkpoint as ck
from sqair_b200.params import ParamStore

HERE = os.path.dirname(os.path.abspath(__file__))


def test_crc32c_known_answers():
    assert ck.crc32c(b'123456789') == 0xe3069283                      # the CRC-32C check value
    assert ck.crc32c(b'\x00' * 32) == 0x8a9136aa                      # RFC 3720 B.4 test vectors
    assert ck.crc32c(b'\xff' * 32) == 0x62a8ab43
    assert ck.crc32c(bytes(range(32))) == 0x46dd794e
    assert ck.mask_crc(0) == 0xa282ead8


def _reference_like_tensors(rng):
    ref = json.load(open(os.path.join(HERE, 'golden', 'ref_variables.json')))
    tensors = {}
    for name, shape in ref['variables'].items():          # notebooks/play.ipynb:239-362
        tensors[name] = rng.standard_normal(tuple(shape)).astype(np.float32)
        tensors[name + '/RMSProp'] = rng.random(tuple(shape)).astype(np.float32)
    tensors['global_step'] = np.asarray(1000000, dtype=np.int64)
    return tensors


def test_round_trip_with_the_reference_variable_names(tmp_path):
    rng = np.random.default_rng(0)
    tensors = _reference_like_tensors(rng)
    prefix = str(tmp_path / 'model.ckpt-1000000')
    ck.write_checkpoint(prefix, tensors, block_size=512)                # many data blocks, prefix-compressed keys
    raw = open(prefix + '.index', 'rb').read()
    assert raw[-8:] == bytes.fromhex('57fb808b247547db')                # table magic number, little endian
    listing = ck.list_variables(prefix)
    assert list(listing) == sorted(tensors, key=lambda n: n.encode())   # byte order, header entry not listed
    w = 'discovery/discovery_core/encoder/mlp/linear/w'
    assert listing[w]['shape'] == tensors[w].shape and listing[w]['dtype'] == 1
    back = ck.read_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape
        np.testing.assert_array_equal(back[k], v)
    assert int(back['global_step']) == 1000000
    sub = ck.read_checkpoint(prefix, names=[w])
    assert list(sub) == [w]
    with pytest.raises(KeyError):
        ck.read_checkpoint(prefix, names=['no/such/variable'])
    assert ck.find_model_files(str(tmp_path)) == {1000000: prefix}


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(1)
    prefix = str(tmp_path / 'm')
    ck.write_checkpoint(prefix, {'a/w': rng.standard_normal((7, 5)).astype(np.float32), 'a/b': np.zeros(5, np.float32)})
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[11] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    with pytest.raises(ValueError, match='checksum mismatch'):
        ck.read_checkpoint(prefix)
    assert ck.read_checkpoint(prefix, verify=False)['a/w'].shape == (7, 5)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[5] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError, match='block checksum'):
        ck.list_variables(prefix)
    idx[-1] ^= 0xff
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError, match='magic'):
        ck.list_variables(prefix)


def test_parameter_store_round_trip_with_optimizer_slots(tmp_path):
    cfg = _capi.make_cfg(1, 1, 1, 3, 50, 50)
    cpu = torch.device('cpu')
    a = ParamStore(cfg, cpu, seed=1)
    opt = optim.make_optimizer('rmsprop', 1e-5)
    s0, s1 = opt._get_slots(a)
    s0.uniform_(0.5, 2.0); s1.normal_()
    opt.global_step = 123456
    prefix = ck.save_from(a, str(tmp_path / 'model.ckpt-123456'), optimizer=opt)
    names = list(ck.list_variables(prefix))
    assert 'global_step' in names and len(names) == 3 * len(a.table) + 1
    assert sum(int(np.prod(s)) for s, _ in a.table.values()) == 2951522          # notebooks/play.ipynb:362 (n = 3)
    b = ParamStore(cfg, cpu, seed=2)
    opt2 = optim.make_optimizer('rmsprop', 1e-5)
    assert not torch.equal(a.flat, b.flat)
    unused = ck.load_into(b, prefix, optimizer=opt2)
    assert unused == []
    assert torch.equal(a.flat, b.flat) and opt2.global_step == 123456
    t0, t1 = opt2._get_slots(b)
    assert torch.equal(t0, s0) and torch.equal(t1, s1)
    # a model with another slot count does not accept the checkpoint silently
    c = ParamStore(_capi.make_cfg(1, 1, 1, 4, 50, 50), cpu)
    with pytest.raises(ValueError, match='shape'):
        ck.load_into(c, prefix)
    partial = {k: v for k, v in ck.read_checkpoint(prefix).items() if not k.startswith('decoder/')}
    ck.write_checkpoint(str(tmp_path / 'partial'), partial)
    with pytest.raises(KeyError, match='lacks'):
        ck.load_into(b, str(tmp_path / 'partial'))
    assert ck.load_into(b, str(tmp_path / 'partial'), strict=False) != []
