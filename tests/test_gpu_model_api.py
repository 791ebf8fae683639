"""GPU: the reference-facing plugin surface (config module -> Model) against the oracle."""
import numpy as np
import pytest
import torch

import sqair_testlib as TL
from oracle import sqair_oracle as O

pytestmark = pytest.mark.gpu


def test_config_load_model_matches_oracle():
    assert torch.cuda.is_available()
    from sqair_b200.common_model_flags import flags
    from sqair_b200.configs import mlp_mnist_model as config
    F = flags.FLAGS
    cfg = O.Cfg(T=4, B=6, K=5, n=3)
    F.n_steps_per_image, F.k_particles = cfg.n, cfg.K
    imgs, params, noise = TL.make_inputs(cfg)
    dev = torch.device('cuda:0')
    obs = torch.from_numpy(imgs).to(dev)
    model = config.load(obs, None, None, mean_img=imgs.mean((0, 1)))
    # same variables (by TF name) and the same draws as the oracle
    model.sequence.param_store(cfg.H, cfg.W, dev).load_state_dict(params)
    model._build(noise={k: torch.from_numpy(v).to(dev) for k, v in noise.items()})
    want, wobj = TL.run_oracle(cfg, imgs, params, noise)
    got = {k: getattr(model, k).cpu().numpy() for k in want}
    bad = TL.compare_outputs(got, want)
    assert not bad, '\n'.join(bad)
    for k in ('elbo_vae', 'elbo_iwae', 'ess', 'data_ll', 'kl', 'log_p_z', 'log_q_z_given_x', 'num_steps', 'raw_mse'):
        np.testing.assert_allclose(float(getattr(model, k)), float(wobj[k]), rtol=2e-4, atol=1e-4, err_msg=k)
    np.testing.assert_allclose(model.log_weights.cpu().numpy(), wobj['log_weights'], rtol=1e-4, atol=1e-3)
    target, gvs = model.make_target()
    np.testing.assert_allclose(float(target), float(wobj['vimco_target']), rtol=2e-4, atol=1e-4)
    assert gvs is None
    assert model.resampled_canvas.shape == (cfg.T, cfg.B, cfg.H, cfg.W)
    assert tuple(model.iw_resampling_idx.shape) == (cfg.B,)
    # a second step on new frames through the same API
    res = model.step(obs, seed=5)
    assert res['log_weights'].shape == (cfg.B, cfg.K) and torch.isfinite(res['scalars'][:5]).all()


def test_generation_through_the_plugin_surface():
    """The config's `sample_from_prior` flag and `SequentialAIR(generate_after=...)` (configs/mlp_mnist_model.py:48,144;
    seq.py:46): observed frames are inferred, later frames are rolled forward from the propagation prior."""
    from sqair_b200.common_model_flags import flags
    from sqair_b200.configs import mlp_mnist_model as config
    F = flags.FLAGS
    cfg = O.Cfg(T=6, B=4, K=2, n=3)
    F.n_steps_per_image, F.k_particles, F.sample_from_prior = cfg.n, cfg.K, True
    try:
        imgs, params, _ = TL.make_inputs(cfg)
        dev = torch.device('cuda:0')
        obs = torch.from_numpy(imgs).to(dev)
        model = config.load(obs, None, None, mean_img=imgs.mean((0, 1)))
        seq = model.sequence
        assert seq._sample_from_prior and seq._generate_after == -1
        seq._generate_after = 2                                  # what SequentialAIR(..., generate_after=2) sets (seq.py:64)
        a = {k: v.clone() for k, v in seq(obs, k_particles=cfg.K, seed=3).items()}
        b = {k: v.clone() for k, v in seq(obs, k_particles=cfg.K, seed=3).items()}
        assert all(torch.equal(a[k], b[k]) for k in a)           # counter-based draws: reproducible
        assert (a['disc_pres'][3:] == 0).all() and torch.isfinite(a['log_weights_per_timestep']).all()
        assert a['canvas'].shape == (cfg.T, cfg.B * cfg.K, cfg.H, cfg.W)
    finally:
        F.sample_from_prior = False
