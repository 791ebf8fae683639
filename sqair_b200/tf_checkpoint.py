"""Reader / writer for TensorFlow "tensor bundle" checkpoints (what `tf.train.Saver` of TF >= 1.0 writes:
`<prefix>.index` + `<prefix>.data-00000-of-00001`), in pure Python -- no TensorFlow needed.

Why it is here: the reference trains with `tf.train.Saver` (sqair/scripts/experiment.py:165-168,228-234) and its released
model (`scripts/download_models.sh`, evaluated in `notebooks/play.ipynb:421-480`) is such a bundle whose variable names
are exactly the names of `sqair_param_layout` (`notebooks/play.ipynb:239-362`).  `load_into(store, prefix)` therefore
puts a reference checkpoint into a `ParamStore` by name, and `save_from(store, prefix)` writes one the reference could
restore.  The released checkpoint itself is not available offline, so the format code is tested on bundles written by
this module (round trip, prefix-compressed keys, multi-block index, CRC checks) -- `tests/test_tf_checkpoint.py`.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table = LevelDB's SSTable):
  index file  = data blocks + metaindex block + index block + 48-byte footer (two block handles, padding, magic
                0xdb4775248b80fb57).  A block = prefix-compressed entries (varint32 shared, non_shared, value_len, key
                suffix, value) + uint32 restart offsets + uint32 restart count, followed on disk by a 1-byte compression
                type (0 = none, 1 = snappy) and a masked CRC32C.  Keys are tensor names in byte order; the value of a
                key is a serialised BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}; the empty key holds
                the BundleHeaderProto {num_shards, endianness, version}.
  data shards = raw little-endian tensor bytes at the recorded offsets.
"""
import os
import re
import struct
from collections import OrderedDict

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------
# CRC32C (Castagnoli), masked the LevelDB way
# ---------------------------------------------------------------------------------------------
def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_table()
_CRC_NP = np.array(_CRC_TABLE, dtype=np.uint32)


def crc32c(data, crc=0):
    crc ^= 0xffffffff
    tab = _CRC_TABLE
    for b in bytes(data):
        crc = tab[(crc ^ b) & 0xff] ^ (crc >> 8)
    return crc ^ 0xffffffff


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + _MASK_DELTA) & 0xffffffff


# ---------------------------------------------------------------------------------------------
# varints and the three protobuf messages involved (hand-rolled: wire types 0, 2 and 5 only)
# ---------------------------------------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError('malformed varint')


def _parse_message(buf):
    """-> {field number: [values]}; varints as int, length-delimited as bytes, fixed32/64 as int."""
    fields, pos = {}, 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        fields.setdefault(num, []).append(v)
    return fields


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _encode_entry(dtype_id, shape, offset, size, crc):
    dims = b''.join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
    msg = _field(1, 0, _put_varint(dtype_id)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _put_varint(offset))           # (shard_id = 0 and zero offsets are proto3 defaults: omitted)
    msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack('<I', crc))
    return msg


def _decode_entry(buf):
    f = _parse_message(buf)
    shape = []
    for shp in f.get(2, []):
        for dim in _parse_message(shp).get(2, []):
            size = _parse_message(dim).get(1, [0])[0]
            shape.append(size - (1 << 64) if size >= (1 << 63) else size)
    return dict(dtype=f.get(1, [0])[0], shape=tuple(shape), shard=f.get(3, [0])[0], offset=f.get(4, [0])[0],
                size=f.get(5, [0])[0], crc=f.get(6, [None])[0], sliced=7 in f)


# ---------------------------------------------------------------------------------------------
# SSTable blocks
# ---------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
    contents = buf[offset:offset + size]
    ctype = buf[offset + size]
    if verify:
        stored = struct.unpack_from('<I', buf, offset + size + 1)[0]
        if mask_crc(crc32c(buf[offset:offset + size + 1])) != stored:
            raise ValueError('checkpoint index: block checksum mismatch at offset %d' % offset)
    if ctype == 1:
        raise NotImplementedError('snappy-compressed index blocks (tf.train.Saver writes them uncompressed)')
    if ctype != 0:
        raise ValueError('unknown block compression type %d' % ctype)
    return bytes(contents)


def _block_entries(block):
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b'', restart_interval

    def add(self, key, value):
        shared = 0
        if self.count and self.count % self.interval == 0:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def _handle(offset, size):
    return _put_varint(offset) + _put_varint(size)


# ---------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------
def list_variables(prefix, verify=True):
    """-> OrderedDict name -> dict(dtype, shape, shard, offset, size, crc) of every tensor in the bundle."""
    with open(prefix + '.index', 'rb') as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError('%s.index is not a TensorFlow checkpoint index (bad magic number)' % prefix)
    pos = len(buf) - 48
    _, pos = _get_varint(buf, pos)                 # metaindex handle (unused)
    _, pos = _get_varint(buf, pos)
    ioff, pos = _get_varint(buf, pos)
    isize, pos = _get_varint(buf, pos)
    entries = OrderedDict()
    header = None
    for _, hv in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p = _get_varint(hv, 0)
        bsize, _ = _get_varint(hv, p)
        for key, value in _block_entries(_read_block(buf, boff, bsize, verify)):
            if key == b'':
                header = _parse_message(value)
            else:
                entries[key.decode('utf-8')] = _decode_entry(value)
    if header is None:
        raise ValueError('checkpoint index without a bundle header')
    if header.get(2, [0])[0] != 0:
        raise NotImplementedError('big-endian tensor bundles')
    list_variables.num_shards = header.get(1, [1])[0]
    return entries


def read_checkpoint(prefix, names=None, verify=True):
    """-> OrderedDict name -> numpy array.  `names`: subset to read (default: everything).  `verify` checks the CRC32C of
    the index blocks and of every tensor read."""
    entries = list_variables(prefix, verify)
    num_shards = list_variables.num_shards
    out, shards = OrderedDict(), {}
    for name in (names if names is not None else entries):
        if name not in entries:
            raise KeyError('tensor "%s" is not in checkpoint %s' % (name, prefix))
        e = entries[name]
        if e['sliced']:
            raise NotImplementedError('partitioned variable "%s"' % name)
        if e['dtype'] not in _DTYPES:
            raise NotImplementedError('dtype %d of "%s"' % (e['dtype'], name))
        if e['shard'] not in shards:
            shards[e['shard']] = open('%s.data-%05d-of-%05d' % (prefix, e['shard'], num_shards), 'rb')
        f = shards[e['shard']]
        f.seek(e['offset'])
        raw = f.read(e['size'])
        if len(raw) != e['size']:
            raise ValueError('checkpoint data shard truncated while reading "%s"' % name)
        if verify and e['crc'] is not None and mask_crc(crc32c(raw)) != e['crc']:
            raise ValueError('checksum mismatch for tensor "%s"' % name)
        dt = np.dtype(_DTYPES[e['dtype']])
        if int(np.prod(e['shape'], dtype=np.int64)) * dt.itemsize != e['size']:
            raise ValueError('size of "%s" does not match its shape %s' % (name, (e['shape'],)))
        out[name] = np.frombuffer(raw, dtype=dt).reshape(e['shape']).copy()
    for f in shards.values():
        f.close()
    return out


def write_checkpoint(prefix, tensors, block_size=4096, restart_interval=16):
    """Writes {name: array} as a single-shard tensor bundle (the layout `tf.train.Saver` produces)."""
    names = sorted(tensors, key=lambda n: n.encode('utf-8'))
    data, entries, offset = [], [], 0
    for n in names:
        a = np.asarray(tensors[n], order='C')           # (ascontiguousarray would turn scalars into 1-vectors)
        if a.dtype not in _DTYPE_IDS:
            raise NotImplementedError('dtype %s of "%s"' % (a.dtype, n))
        raw = a.tobytes()
        entries.append((n.encode('utf-8'), _encode_entry(_DTYPE_IDS[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
        data.append(raw)
        offset += len(raw)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for raw in data:
            f.write(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))   # 1 shard, version {producer 1}
    out, index = bytearray(), _BlockBuilder(restart_interval=1)

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)
        out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\0'))))
        return off, len(block)

    bb = _BlockBuilder(restart_interval)
    for key, value in [(b'', header)] + entries:
        bb.add(key, value)
        if len(bb.buf) >= block_size:
            index.add(key, _handle(*emit(bb.finish())))
            bb = _BlockBuilder(restart_interval)
    if bb.count:
        index.add(bb.last, _handle(*emit(bb.finish())))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = _handle(*meta) + _handle(*idx)
    out.extend(footer + b'\0' * (40 - len(footer)) + struct.pack('<Q', _MAGIC))
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
    return prefix


def find_model_files(model_dir):
    """{iteration: checkpoint prefix} of the `model.ckpt-<itr>` bundles in a run directory (experiment_tools.py:135-144)."""
    found = {}
    for name in os.listdir(model_dir):
        m = re.match(r'^(model\.ckpt-(\d+))\.index$', name)
        if m:
            found[int(m.group(2))] = os.path.join(model_dir, m.group(1))
    return found


def load_into(store, prefix, optimizer=None, strict=True, verify=True):
    """Reference checkpoint -> ParamStore (by TF variable name; shapes must match `sqair_param_layout`).  With
    `optimizer` (an `optim.RMSPropOptimizer`) the `<var>/RMSProp` (mean square) and `<var>/RMSProp_1` (momentum) slots and
    `global_step` are restored as well (scripts/experiment.py:140,165-168).  Returns the names found in the checkpoint
    but not used."""
    import torch
    entries = list_variables(prefix, verify)
    missing = [n for n in store.table if n not in entries]
    if missing and strict:
        raise KeyError('checkpoint %s lacks %d variables, e.g. %s' % (prefix, len(missing), missing[:3]))
    names = [n for n in store.table if n in entries]
    values = read_checkpoint(prefix, names, verify)
    sd = store.state_dict()
    for n in names:
        if tuple(values[n].shape) != tuple(store.table[n][0]):
            raise ValueError('variable "%s": checkpoint shape %s, model shape %s' % (n, values[n].shape, store.table[n][0]))
        sd[n] = torch.from_numpy(values[n].astype(np.float32))
    store.load_state_dict(sd)
    used = set(names)
    if optimizer is not None:
        s0, s1 = optimizer._get_slots(store)
        for suffix, slot in (('/RMSProp', s0), ('/RMSProp_1', s1)):
            have = [n for n in store.table if n + suffix in entries]
            vals = read_checkpoint(prefix, [n + suffix for n in have], verify)
            for n in have:
                shape, off = store.table[n]
                cnt = int(np.prod(shape)) if len(shape) else 1
                slot[off:off + cnt].copy_(torch.from_numpy(vals[n + suffix].astype(np.float32).reshape(-1)))
                used.add(n + suffix)
        if 'global_step' in entries:
            optimizer.global_step = int(read_checkpoint(prefix, ['global_step'], verify)['global_step'])
            used.add('global_step')
    return [n for n in entries if n not in used]


def save_from(store, prefix, optimizer=None, global_step=None):
    """ParamStore (+ optimiser slots) -> a bundle with the reference's variable names, readable by `tf.train.Saver`."""
    tensors = OrderedDict((n, v.detach().cpu().numpy()) for n, v in store.state_dict().items())
    if optimizer is not None:
        s0, s1 = optimizer._get_slots(store)
        for suffix, slot in (('/RMSProp', s0), ('/RMSProp_1', s1)):
            if slot is None:
                continue
            flat = slot.detach().cpu().numpy()
            for n, (shape, off) in store.table.items():
                cnt = int(np.prod(shape)) if len(shape) else 1
                tensors[n + suffix] = flat[off:off + cnt].reshape(shape)
        global_step = optimizer.global_step if global_step is None else global_step
    if global_step is not None:
        tensors['global_step'] = np.asarray(global_step, dtype=np.int64)
    return write_checkpoint(prefix, tensors)
