"""CPU tests of the C-ABI library: it loads, exports every symbol include/sqair_b200.h declares,
its host-side tables agree with the reference's variable listing, and errors surface like the
reference's (ValueError with the same message).  No compute calls (no GPU here)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import sqair_testlib as TL
from oracle import sqair_oracle as O
from sqair_b200 import _capi

ROOT = TL.ROOT


def _lib():
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _capi.lib()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'sqair_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(sqair_[a-z_0-9]+)\s*\(', hdr))
    assert declared == set(_capi.EXPORTED), declared ^ set(_capi.EXPORTED)
    lib = C.CDLL(_capi.LIB_PATH) if os.path.exists(_capi.LIB_PATH) else _lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib().sqair_version() >= 100


def test_struct_layouts_match_header():
    # sizes implied by the header (all int32/float fields, 8-byte pointers)
    assert C.sizeof(_capi.SqairCfg) == 13 * 4 + 6 * 4 + 8 * 4
    assert C.sizeof(_capi.SqairOutputs) == 38 * 8
    assert C.sizeof(_capi.SqairSizes) == 5 * 8 + 6 * 4


def test_param_layout_matches_reference_listing():
    _lib()
    ref = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'ref_variables.json')))
    cfg = _capi.make_cfg(10, 32, 5, 3, 50, 50)
    table = _capi.param_layout(cfg)
    assert {n: list(s) for n, s, _, _ in table} == ref['variables']
    assert _capi.query_sizes(cfg).param_count == ref['total']
    # canonical order/offsets identical to the oracle's flatten_params
    ocfg = O.Cfg(T=10, B=32, K=5, n=3)
    off = 0
    for (name, shape, o, po), (oname, oshape) in zip(table, O.param_shapes(ocfg).items()):
        assert name == oname and tuple(shape) == tuple(oshape) and o == off and po % 4 == 0
        off += int(np.prod(shape))


def test_query_sizes_and_shapes():
    _lib()
    cfg = _capi.make_cfg(10, 32, 5, 4, 50, 50)
    s = _capi.query_sizes(cfg)
    assert s.rows == 160 and s.n_ctas // s.cluster_size * s.rows_per_cta >= 160 and s.smem_bytes <= 232448
    assert s.eps_what_floats == 10 * 160 * 8 * 50 and s.u_pres_floats == 10 * 160 * 8
    shapes = _capi.output_shapes(cfg)
    assert shapes['what'] == (10, 160, 4, 50) and shapes['canvas'] == (10, 160, 50, 50)
    assert shapes['disc_prob'] == (10, 160, 5) and shapes['log_weights_per_timestep'] == (10, 160)


def test_errors_mirror_reference():
    _lib()
    with pytest.raises(ValueError, match='Invalid prior type'):          # propagate.py:42-43
        _capi.make_cfg(3, 4, 1, 2, 50, 50, prior_type='bogus')
    with pytest.raises(ValueError, match='Invalid prior type'):          # sqair_modules.py:223-224
        _capi.make_cfg(3, 4, 1, 2, 50, 50, disc_prior_type='bogus')
    bad = _capi.make_cfg(3, 4, 1, 2, 50, 50)
    bad.n = 0
    with pytest.raises(ValueError, match='n_steps_per_image'):
        _capi.query_sizes(bad)
    assert b'n_steps_per_image' in _capi.lib().sqair_last_error()


def test_ops_refuse_cpu_tensors():
    """The product path has no CPU fallback: CPU tensors are rejected, not silently computed."""
    import torch
    from sqair_b200 import ops
    cfg = _capi.make_cfg(3, 4, 1, 2, 50, 50)
    with pytest.raises(ValueError, match='CUDA'):
        ops.pack_params(cfg, torch.zeros(_capi.query_sizes(cfg).param_count))
    with pytest.raises(ValueError, match='CUDA'):
        ops.stn_glimpse(torch.zeros(1, 50, 50), torch.zeros(1, 4), 20)
