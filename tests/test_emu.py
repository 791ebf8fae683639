"""CPU check of the kernel program itself: sqair_device.cuh compiled for the host (one sequential
"thread" per block, tests/host_emu) must reproduce the oracle.  This validates indexing, slot
bookkeeping, the layer/packing tables and every formula of the device code without a GPU; the
`-m gpu` tests then validate the real kernel."""
import numpy as np
import pytest

import sqair_testlib as TL
from oracle import sqair_oracle as O

CASES = [      # (config, rows per cluster, blocks per cluster)
    (dict(T=3, B=4, K=1, n=2), 2, 1),
    (dict(T=3, B=3, K=2, n=3), 4, 2),
    (dict(T=2, B=2, K=5, n=4), 5, 4),
    (dict(T=3, B=2, K=2, n=2, prior_type='rw'), 1, 8),
    (dict(T=3, B=2, K=2, n=2, prior_type='guided', disc_prior_type='geom'), 3, 4),
    (dict(T=2, B=2, K=1, n=2, rec_where_prior=False, masked_glimpse=False), 2, 2),
    (dict(T=2, B=2, K=2, n=6, H=64, W=64), 4, 4),
    (dict(T=3, B=3, K=2, n=1), 2, 2),          # one slot: the reference squeezes the slot axis of the per-slot log-probs
    (dict(T=2, B=2, K=1, n=8), 6, 3),          # maximum slot count, largest rows-per-cluster instantiation
    (dict(T=2, B=3, K=1, n=2, H=45, W=35), 1, 5),
]


@pytest.mark.parametrize('kw,R,C', CASES)
def test_emulated_kernel_matches_oracle(kw, R, C):
    cfg = O.Cfg(**kw)
    imgs, params, noise = TL.make_inputs(cfg)
    want, _ = TL.run_oracle(cfg, imgs, params, noise)
    got = TL.run_emu(cfg, imgs, params, noise, R, cluster=C)
    bad = TL.compare_outputs(got, want)
    assert not bad, '\n'.join(bad)


@pytest.mark.parametrize('kw,R,C', [
    (dict(T=4, B=3, K=2, n=3, sample_from_prior=True, generate_after=1), 2, 2),       # frames 2, 3 roll forward from the prior
    (dict(T=3, B=2, K=2, n=2, sample_from_prior=True), 2, 1),                         # q evaluated at prior draws, nothing replaced
    (dict(T=4, B=2, K=1, n=2, sample_from_prior=True, generate_after=2, prior_type='guided', rec_where_prior=False), 1, 4),
])
def test_emulated_generation_matches_oracle(kw, R, C):
    """sample_from_prior / generate_after (seq.py:198-203; sqair_modules.py:157-170,294-302) in the kernel program."""
    cfg = O.Cfg(**kw)
    imgs, params, noise = TL.make_inputs(cfg)
    noise = TL.with_prior_noise(cfg, noise)
    want, _ = TL.run_oracle(cfg, imgs, params, noise)
    got = TL.run_emu(cfg, imgs, params, noise, R, cluster=C)
    bad = TL.compare_outputs(got, want)
    assert not bad, '\\n'.join(bad)
    if cfg.generate_after > 0:
        assert (want['disc_pres'][cfg.generate_after + 1:] == 0).all()                # generated frames discover nothing
        plain, _ = TL.run_oracle(O.Cfg(**{k: v for k, v in kw.items() if k not in ('sample_from_prior', 'generate_after')}), imgs, params, noise)
        assert np.array_equal(plain['what'][:cfg.generate_after + 1], want['what'][:cfg.generate_after + 1])   # observed prefix unchanged
        assert not np.array_equal(plain['what'][cfg.generate_after + 1:], want['what'][cfg.generate_after + 1:])
