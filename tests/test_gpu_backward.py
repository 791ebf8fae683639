"""GPU parity of the backward pass: sqair_forward_train + sqair_objective_grad + sqair_backward (C ABI) against torch
autograd through the oracle evaluated in float64 (`opt.compute_gradients(target)`, model.py:150-168; targets.py:46-75).

Tolerance per variable: |got - want| <= 1e-3 |want| + 2e-4 max|want| (+ 3 x the fp32 oracle's own distance from the
float64 value: the measured fp32 noise floor of sums with heavy cancellation).  Every variable must receive a gradient
(model.py:163-166)."""
import numpy as np
import pytest
import torch

import sqair_testlib as TL
from oracle import sqair_oracle as O

pytestmark = pytest.mark.gpu

CASES = {
    'c1_T3_B4_K1_n2': dict(T=3, B=4, K=1, n=2),                       # BASELINE configs[0] (K = 1: -elbo_iwae target)
    'small_c2_T4_B3_K5_n4': dict(T=4, B=3, K=5, n=4),                 # configs[1] / configs[2] (VIMCO) shape
    'n3_K2': dict(T=3, B=3, K=2, n=3),
    'rw_prior': dict(T=3, B=2, K=2, n=2, prior_type='rw'),
    'guided_geom': dict(T=3, B=2, K=2, n=2, prior_type='guided', disc_prior_type='geom'),
    'no_rec_no_mask': dict(T=3, B=2, K=2, n=2, rec_where_prior=False, masked_glimpse=False),
    'bg_std_differs': dict(T=3, B=2, K=2, n=2, bg_std=0.5),              # d std / d mask term of the likelihood (modules.py:453)
    'c4_like_64px_n6': dict(T=2, B=2, K=2, n=6, H=64, W=64),
    'one_slot': dict(T=3, B=3, K=2, n=1),
    'max_slots_n8': dict(T=2, B=2, K=2, n=8),
    'odd_pixels_45x35': dict(T=2, B=3, K=2, n=2, H=45, W=35),
}


def _check(cfg, with_floor=True, smooth=True):
    assert torch.cuda.is_available()
    if smooth:       # small cases: a noise seed without resampler coordinates on a derivative jump (TL.smooth_inputs)
        imgs, params, noise, want = TL.smooth_inputs(cfg)
    else:            # full size: ~50 such samples are unavoidable, each moves 1/50 of one row-frame's likelihood term
        imgs, params, noise = TL.make_inputs(cfg)
        want, _ = TL.oracle_gradients(cfg, imgs, params, noise, double=True)
    floor = TL.oracle_gradients(cfg, imgs, params, noise)[0] if with_floor else None
    got, outs, launches = TL.run_cuda_backward(cfg, imgs, params, noise, return_outputs=True)
    fwd, _ = TL.run_oracle(cfg, imgs, params, noise)
    bad = TL.compare_outputs(outs, fwd)                        # the stash-writing forward is still the forward
    assert not bad, '\n'.join(bad)
    bad = TL.compare_gradients(got, want, floor=floor)
    assert not bad, '\n'.join(bad)
    assert launches > 0
    return got


@pytest.mark.parametrize('name', list(CASES))
def test_backward_parity(name):
    cfg = O.Cfg(**CASES[name])
    got = _check(cfg)
    if cfg.disc_prior_type == 'cat':
        missing = [k for k, v in got.items() if not (np.abs(v).max() > 0)]
        assert not missing, 'variables without a gradient: %s' % missing


def test_full_size_c2_backward_parity():
    """BASELINE configs[1] / configs[2]: T=10, B=32, K=5, n=4, 50x50, VIMCO target."""
    _check(O.Cfg(T=10, B=32, K=5, n=4), smooth=False)


def test_c4_shaped_backward_parity():
    """BASELINE configs[3] shape (100x100 canvas, n=6 objects, B=16, K=10) with three frames: the frames are read from
    global memory (they do not fit the shared-memory staging), 60 glimpses per row in the canvas stage."""
    _check(O.Cfg(T=3, B=16, K=10, n=6, H=100, W=100), with_floor=False, smooth=False)


def test_backward_is_reproducible_and_workspace_independent():
    """Two runs with differently poisoned scratch memory agree to accumulation-order noise (atomics in the weight
    gradient reductions) -- nothing reads uninitialised workspace."""
    cfg = O.Cfg(T=3, B=3, K=2, n=3)
    imgs, params, noise = TL.make_inputs(cfg)
    a = TL.run_cuda_backward(cfg, imgs, params, noise)
    b = TL.run_cuda_backward(cfg, imgs, params, noise)
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-4, atol=1e-5 * max(np.abs(a[k]).max(), 1e-30), err_msg=k)


@pytest.mark.parametrize('name', ['small_c2_T4_B3_K5_n4', 'no_rec_no_mask', 'max_slots_n8'])
def test_program_kernel_matches_launch_per_operation(name, monkeypatch):
    """The reverse program as one persistent cluster kernel (default; dgrad products on the tensor cores with the forward's
    fp32-faithful tf32 split) against the same program issued as one launch per operation (SQAIR_BWD_LAUNCHES=1; fp32
    FFMA dgrad): the two paths share the stage code and the weight-gradient GEMMs, not the product kernels."""
    cfg = O.Cfg(**CASES[name])
    imgs, params, noise = TL.make_inputs(cfg)
    monkeypatch.delenv('SQAIR_BWD_LAUNCHES', raising=False)
    a, _, la = TL.run_cuda_backward(cfg, imgs, params, noise, return_outputs=True)
    monkeypatch.setenv('SQAIR_BWD_LAUNCHES', '1')
    b, _, lb = TL.run_cuda_backward(cfg, imgs, params, noise, return_outputs=True)
    assert la < lb / 2, (la, lb)                  # at BASELINE configs[1]: ~120 launches (program + weight-gradient GEMMs) instead of ~1 450
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=2e-4, atol=2e-5 * max(np.abs(b[k]).max(), 1e-30), err_msg=k)


@pytest.mark.parametrize('shape', ['1:1', '3:2', '7:4', '12:8'])
def test_program_kernel_launch_shapes(shape, monkeypatch):
    """Rows per cluster / cluster size of the reverse-program kernel are free parameters (SQAIR_BWD_ROWS / SQAIR_BWD_CLUSTER):
    a single block, a ragged last cluster (15 rows), operands wider than one pass of three MMA n-tiles (12 rows x 3 slots)."""
    cfg = O.Cfg(T=2, B=5, K=3, n=3)
    imgs, params, noise = TL.make_inputs(cfg)
    monkeypatch.setenv('SQAIR_BWD_LAUNCHES', '1')
    want = TL.run_cuda_backward(cfg, imgs, params, noise)
    monkeypatch.delenv('SQAIR_BWD_LAUNCHES')
    r, c = shape.split(':')
    monkeypatch.setenv('SQAIR_BWD_ROWS', r)
    monkeypatch.setenv('SQAIR_BWD_CLUSTER', c)
    got = TL.run_cuda_backward(cfg, imgs, params, noise)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=2e-4, atol=2e-5 * max(np.abs(want[k]).max(), 1e-30), err_msg=k)


def test_program_tables_are_cached_and_evicted():
    """Every (configuration, buffer set) records its own operation table; the library keeps 64 of them.  More than 64 distinct
    buffer sets (all kept alive, so the addresses differ) exercise the eviction path; results stay identical."""
    cfg = O.Cfg(T=1, B=1, K=2, n=1)
    imgs, params, noise = TL.make_inputs(cfg)
    keep, first = [], None
    for i in range(70):
        g, outs, _ = TL.run_cuda_backward(cfg, imgs, params, noise, return_outputs=True, keep_alive=keep)
        if first is None:
            first = g
        elif i % 23 == 0 or i == 69:
            for k in first:
                np.testing.assert_allclose(g[k], first[k], rtol=1e-4, atol=1e-5 * max(np.abs(first[k]).max(), 1e-30), err_msg=k)


def test_long_sequence_program_matches_launches(monkeypatch):
    """T = 100 (the roll-out length of BASELINE configs[4]): ~14 000 recorded operations in one table."""
    cfg = O.Cfg(T=100, B=2, K=2, n=2)
    imgs, params, noise = TL.make_inputs(cfg)
    monkeypatch.delenv('SQAIR_BWD_LAUNCHES', raising=False)
    a = TL.run_cuda_backward(cfg, imgs, params, noise)
    monkeypatch.setenv('SQAIR_BWD_LAUNCHES', '1')
    b = TL.run_cuda_backward(cfg, imgs, params, noise)
    for k in a:
        assert np.isfinite(a[k]).all(), k
        np.testing.assert_allclose(a[k], b[k], rtol=2e-4, atol=2e-5 * max(np.abs(b[k]).max(), 1e-30), err_msg=k)
