"""CPU check of the backward pass (no GPU): the stage functions and the driver of sqair_backward.h, compiled for the
host (tests/host_emu: sequential loops instead of kernels, stash and workspace poisoned with NaN), against torch
autograd through the oracle (model.py:150-168 `opt.compute_gradients(target)`).  Every variable of the reference must
receive its gradient (model.py:163-166)."""
import pytest

import sqair_testlib as TL
from oracle import sqair_oracle as O

CASES = [      # (config, rows per cluster, blocks per cluster) of the emulated forward that fills the stash
    (dict(T=3, B=4, K=1, n=2), 2, 1),                                   # BASELINE configs[0]; K = 1 -> -elbo_iwae target
    (dict(T=3, B=2, K=3, n=3), 3, 2),                                   # VIMCO
    (dict(T=2, B=2, K=2, n=2, prior_type='guided', disc_prior_type='geom'), 1, 4),
    (dict(T=2, B=2, K=2, n=2, prior_type='rw', rec_where_prior=False, masked_glimpse=False), 2, 2),
    (dict(T=3, B=2, K=2, n=2, rec_where_prior=False, masked_glimpse=False), 2, 1),
    (dict(T=2, B=2, K=2, n=2, bg_std=0.5), 2, 2),                      # background std differs from the object std
         # seed 7 puts a canvas row on a bilinear kink
]


@pytest.mark.parametrize('kw,R,C', CASES)
def test_emulated_backward_matches_autograd(kw, R, C):
    cfg = O.Cfg(**kw)
    imgs, params, noise, want = TL.smooth_inputs(cfg)
    floor, _ = TL.oracle_gradients(cfg, imgs, params, noise)
    got, outs = TL.run_emu_backward(cfg, imgs, params, noise, R, cluster=C)
    fwd, _ = TL.run_oracle(cfg, imgs, params, noise)
    assert not TL.compare_outputs(outs, fwd)                  # stashing must not disturb the forward results
    bad = TL.compare_gradients(got, want, floor=floor)
    assert not bad, '\n'.join(bad)
    if cfg.disc_prior_type == 'cat':
        assert all(abs(v).max() > 0 for v in got.values())     # no variable is left without a gradient
