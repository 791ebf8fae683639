"""CPU tests of the host-side mirror of the reference's operator surface (no GPU, no compute):
flag defaults, config wiring, architecture validation and error behaviour."""
import json
import os

import pytest
import torch

import sqair_testlib as TL
from sqair_b200 import rnn as snt
from sqair_b200.common_model_flags import flags, get_params
from sqair_b200.core import DiscoveryCore, PropagationCore
from sqair_b200.modules import (AIRDecoder, AIREncoder, Decoder, Encoder, SpatialTransformer, StepsPredictor,
                                StochasticTransformParam)
from sqair_b200.propagate import SequentialSSM, make_prior
from sqair_b200.seq import SequentialAIR
from sqair_b200.sqair_modules import Discover, Propagate
from sqair_b200 import targets, index, parallel
from sqair_b200.configs import mlp_mnist_model as config   # defines the config's flags


def build_sequence(n=3, transition='VanillaRNN', prior='rnn', disc_prior='cat', rec=True, masked=True, nh=256):
    img_size, gs, nw = [50, 50, 1], [20, 20], 50
    hid = ([nh, nh],)
    rnn_class = getattr(snt, transition)
    enc = lambda: AIREncoder(img_size, gs, nw, Encoder(hid), masked_glimpse=masked)
    dcell = DiscoveryCore(img_size, gs, nw, rnn_class(nh), lambda: Encoder(hid), enc,
                          lambda: StochasticTransformParam(hid, -3.), lambda: StepsPredictor(([nh // 2],), 1.))
    disc = Discover(n, dcell, 0.75, where_mean=[-2., -2., 0, 0], disc_prior_type=disc_prior, rec_where_prior=rec)
    pcell = PropagationCore(img_size, gs, nw, rnn_class(nh), lambda: dcell._input_encoder, lambda: dcell._glimpse_encoder,
                            lambda: StochasticTransformParam(hid, -3.), lambda: StepsPredictor(([nh // 2],), 5.), snt.GRU(nh))
    prop = Propagate(SequentialSSM(pcell), make_prior(prior, nw, snt.GRU(nh), 10.))
    dec = AIRDecoder(img_size, gs, lambda size: Decoder(hid, size, output_scale=.25), mean_img=None, output_std=.3)
    return SequentialAIR(n, gs, disc, prop, snt.GRU(nh), dec)


def test_flag_defaults_match_released_run():
    """release_models/mnist_mlp/1/flags.json (copied values): the defaults of the model flags."""
    F = flags.FLAGS
    released = dict(disc_prior_type='cat', disc_step_bias=1.0, glimpse_size=20, k_particles=5, masked_glimpse=True,
                    n_steps_per_image=3, n_units=8, n_what=50, output_scale=0.25, output_std=0.3,
                    prior_transition='GRU', prop_prior_step_bias=10.0, prop_prior_type='rnn', prop_step_bias=5.0,
                    rec_where_prior=True, sample_from_prior=False, step_success_prob=0.75, time_transition='GRU',
                    transform_var_bias=-3.0, transition='VanillaRNN')
    for k, v in released.items():
        assert getattr(F, k) == v, k
    p = get_params()
    assert p.n_hidden == 256 and p.glimpse_size == [20, 20] and p.n_hiddens == ([256, 256],) and p.steps_pred_hidden == ([128],)


def test_sequence_reads_architecture_into_kernel_config():
    seq = build_sequence(n=4, prior='guided', disc_prior='geom', rec=False, masked=False)
    cfg = seq.make_cfg(10, 32, 5, 50, 50)
    assert (cfg.T, cfg.B, cfg.K, cfg.n, cfg.H, cfg.W, cfg.G, cfg.n_what, cfg.n_hidden) == (10, 32, 5, 4, 50, 50, 20, 50, 256)
    assert cfg.prior_type == 2 and cfg.disc_prior_type == 1 and cfg.rec_where_prior == 0 and cfg.masked_glimpse == 0
    assert abs(cfg.prop_prior_step_bias - 10.) < 1e-6 and abs(cfg.output_std - .3) < 1e-6 and abs(cfg.bg_std - .3) < 1e-6


def test_unsupported_architectures_fail_loudly():
    with pytest.raises(NotImplementedError, match='VanillaRNN'):
        build_sequence(transition='GRU')
    with pytest.raises(ValueError, match='Invalid prior type'):
        build_sequence(prior='nope')
    with pytest.raises(ValueError, match='Invalid prior type'):
        build_sequence(disc_prior='nope')
    with pytest.raises(ValueError, match='Only one of'):
        StepsPredictor([128], max_rel_logit_change=.1, max_logit_change=.1)
    st = SpatialTransformer((50, 50), (20, 20))
    with pytest.raises(ValueError, match='coords or logits'):
        st(torch.zeros(1, 50, 50))
    with pytest.raises(ValueError, match='not both'):
        st(torch.zeros(1, 50, 50), coords=torch.zeros(1, 4), logits=torch.zeros(1, 4))


def test_config_load_needs_a_gpu_not_a_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(Exception):
        config.load(torch.zeros(3, 4, 50, 50), None, None)


def test_targets_match_oracle_formulas():
    from oracle import sqair_oracle as O
    lw, lp = torch.randn(6, 5) * 3, torch.randn(6, 5)
    assert torch.allclose(targets.iwae(lw), O.iwae(lw))
    assert torch.allclose(targets.vimco_control_variate(lw), O.vimco_control_variate(lw))
    assert torch.allclose(targets.vimco(lw, lp), O.vimco(lw, lp))
    pres = (torch.rand(5, 8) < .5).float()
    x = torch.randn(5, 8, 3)
    assert torch.equal(index.select_present(x, pres), O.select_present(x, pres))
    assert index.tile_input_for_iwae(torch.arange(6.).reshape(2, 3, 1), 2, with_time=True)[0, :, 0].tolist() == [0, 0, 1, 1, 2, 2]


def test_shard_ranges_cover_batch():
    for n, w in ((32, 8), (33, 8), (5, 4), (3, 8)):
        got = [parallel.shard_range(n, w, r) for r in range(w)]
        assert sum(c for _, c in got) == n
        assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(w - 1))
    assert parallel.row_offset(32, 8, 3, 5) == 60


def test_learning_rate_schedule_matches_the_training_script():
    """scripts/experiment.py:128-136: stages '4,6,10' of train_itr, the rate divided by 3 at each boundary
    (tf.train.piecewise_constant: the old value holds AT the boundary step)."""
    from sqair_b200 import optim
    lr = optim.make_schedule(1e-5, '4,6,10', 2000000)          # release_models/mnist_mlp/1/flags.json
    assert lr(0) == pytest.approx(1e-5) and lr(400000) == pytest.approx(1e-5)
    assert lr(400001) == pytest.approx(1e-5 / 3) and lr(1000000) == pytest.approx(1e-5 / 3)
    assert lr(1000001) == pytest.approx(1e-5 / 9) and lr(5000000) == pytest.approx(1e-5 / 9)
    assert optim.make_schedule(3e-4, '', 10)(7) == pytest.approx(3e-4)
    with pytest.raises(ValueError):
        optim.piecewise_constant(0, [1, 2], [1.0])
    with pytest.raises(ValueError):
        optim.make_optimizer('lbfgs', 1e-3)
    assert optim.make_optimizer('rmsprop', 1e-5).momentum == 0.9   # experiment.py:140
