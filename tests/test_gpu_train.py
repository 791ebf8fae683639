"""GPU: the training surface -- `Model.make_target(opt) -> (target, gvs)` (model.py:150-168), the optimiser update
(scripts/experiment.py:138-155, TF 1.x rules) and the data-parallel gradient combination -- against the oracle
(torch autograd on the CPU restatement + a numpy restatement of the TF update rules)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import sqair_testlib as TL
from oracle import sqair_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_model(cfg, imgs, params, noise, dev, row_offset=0):
    from sqair_b200.common_model_flags import flags
    from sqair_b200.configs import mlp_mnist_model as config
    F = flags.FLAGS
    F.n_steps_per_image, F.k_particles = cfg.n, cfg.K
    obs = torch.from_numpy(imgs).to(dev)
    model = config.load(obs, None, None, mean_img=imgs.mean((0, 1)))
    model._row_offset = row_offset
    model.sequence.param_store(cfg.H, cfg.W, dev).load_state_dict(params)
    nz = {k: torch.from_numpy(v).to(dev) for k, v in noise.items()}
    model._build(noise=nz)
    return model, obs, nz


def test_make_target_returns_a_gradient_for_every_variable():
    from sqair_b200 import optim
    dev = torch.device('cuda:0')
    cfg = O.Cfg(T=3, B=3, K=4, n=3)
    imgs, params, noise, want = TL.smooth_inputs(cfg)
    model, obs, nz = _load_model(cfg, imgs, params, noise, dev)
    opt = optim.make_optimizer('rmsprop', 1e-5)
    target, gvs = model.make_target(opt)
    _, wobj = TL.run_oracle(cfg, imgs, params, noise)
    np.testing.assert_allclose(float(target), float(wobj['vimco_target']), rtol=2e-4, atol=1e-4)
    names = [n for n, _ in gvs.names]
    assert names == list(want.keys())                                 # one pair per variable, canonical order
    got = {n: g.cpu().numpy() for n, (g, v) in zip(names, gvs)}
    for n, (g, v) in zip(names, gvs):
        assert tuple(g.shape) == tuple(v.shape) == tuple(want[n].shape), n
    floor, _ = TL.oracle_gradients(cfg, imgs, params, noise)
    bad = TL.compare_gradients(got, want, floor=floor)
    assert not bad, '\n'.join(bad)
    missing = [k for k, v in got.items() if not (np.abs(v).max() > 0)]
    assert not missing, missing                                       # model.py:163-166
    # L2 regulariser (targets.py:31-35): value += w/2 sum v^2, gradient += w v
    t2, gvs2 = model.make_target(opt, l2_reg=0.1)
    flat = O.flatten_params(params, cfg).double()
    np.testing.assert_allclose(float(t2), float(wobj['vimco_target']) + 0.05 * float((flat ** 2).sum()), rtol=2e-4)
    k = 'decoder/air_decoder/decoder/mlp/linear/w'
    i = names.index(k)
    np.testing.assert_allclose(gvs2[i][0].cpu().numpy(), got[k] + 0.1 * params[k].numpy(), rtol=1e-4, atol=1e-5)


def _tf_update(kind, w, g, s0, s1, lr, t):
    """numpy (float32) restatement of the TF 1.x kernels ApplyRMSProp / ApplyAdam / ApplyMomentum / ApplyGradientDescent."""
    f = np.float32
    if kind == 'rmsprop':
        s0 += (g * g - s0) * f(1 - 0.9)
        s1[:] = s1 * f(0.9) + (g * f(lr)) / np.sqrt(s0 + f(1e-10))
        w -= s1
    elif kind == 'adam':
        lr_t = f(lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t))
        s0 += (g - s0) * f(1 - 0.9)
        s1 += (g * g - s1) * f(1 - 0.999)
        w -= (s0 * lr_t) / (np.sqrt(s1) + f(1e-8))
    elif kind == 'momentum':
        s0[:] = s0 * f(0.9) + g
        w -= f(lr) * s0
    else:
        w -= f(lr) * g


@pytest.mark.parametrize('kind', ['rmsprop', 'adam', 'momentum', 'sgd'])
def test_optimizer_update_follows_tf_rules(kind):
    from sqair_b200 import optim
    from sqair_b200.params import ParamStore
    from sqair_b200 import _capi
    dev = torch.device('cuda:0')
    store = ParamStore(_capi.make_cfg(1, 1, 1, 2, 50, 50), dev)
    rng = np.random.default_rng(3)
    w = store.flat.cpu().numpy().copy()
    s0 = np.ones_like(w) if kind == 'rmsprop' else np.zeros_like(w)
    s1 = np.zeros_like(w)
    sched = optim.make_schedule(1e-3, '4,6,10', 20)                   # boundaries at steps 4 and 10
    opt = optim.make_optimizer(kind, sched)
    gvs = optim.GradsAndVars()
    gvs.store = store
    for t in range(1, 8):
        g = (rng.standard_normal(w.shape) * 10.0 ** rng.integers(-3, 2)).astype(np.float32)
        gvs.flat_grad = torch.from_numpy(g).to(dev)
        lr = sched(t - 1)
        assert lr == pytest.approx(1e-3 if t - 1 <= 4 else 1e-3 / 3)
        _tf_update(kind, w, g, s0, s1, lr, t)
        opt.apply_gradients(gvs)
        np.testing.assert_allclose(store.flat.cpu().numpy(), w, rtol=2e-6, atol=1e-7)
    assert opt.global_step == 7


def test_training_steps_follow_an_oracle_driven_loop():
    """Three iterations of the reference's loop (compute_gradients -> apply_gradients) through Model.train_step against
    the same loop driven by the oracle's autograd gradients and the numpy RMSProp: per-variable parameter change."""
    from sqair_b200 import optim
    dev = torch.device('cuda:0')
    cfg = O.Cfg(T=3, B=2, K=3, n=2)
    imgs, params, noise, _ = TL.smooth_inputs(cfg)
    model, obs, nz = _load_model(cfg, imgs, params, noise, dev)
    store = model.sequence.param_store(cfg.H, cfg.W, dev)
    lr = 1e-5                                      # the released run's rate (release_models/mnist_mlp/1/flags.json)
    opt = optim.make_optimizer('rmsprop', lr)
    p = {k: v.clone() for k, v in params.items()}
    flat0 = O.flatten_params(p, cfg).numpy().copy()
    w = flat0.copy()
    s0, s1 = np.ones_like(w), np.zeros_like(w)
    for it in range(3):
        gvs = model.compute_gradients(obs, noise=nz)
        opt.apply_gradients(gvs)
        g, _ = TL.oracle_gradients(cfg, imgs, O.unflatten_params(torch.from_numpy(w.copy()), cfg), noise, double=True)
        g = {k: v.astype(np.float32) for k, v in g.items()}
        gflat = O.flatten_params({k: torch.from_numpy(v) for k, v in g.items()}, cfg).numpy()
        _tf_update('rmsprop', w, gflat, s0, s1, lr, it + 1)
    got = O.unflatten_params(torch.from_numpy(store.flat.cpu().numpy() - flat0), cfg)
    want = O.unflatten_params(torch.from_numpy(w - flat0), cfg)
    bad = []
    # RMSProp turns a gradient g into lr g / sqrt(ms): saturated (~3 lr per step) where |g| is large, ~lr g where it is small.
    # The gradient parity bar allows an absolute error of 2e-4 max|g| per variable; on an unsaturated entry of a variable
    # whose largest gradient is ~300 that is ~0.06 lr per step = ~1 % of the variable's largest parameter change.  So:
    # 0.5 % of the change + 2 % of the variable's largest change + two ulps of the parameter itself.
    for k in want:
        g_, w_, p_ = got[k].numpy(), want[k].numpy(), params[k].numpy()
        tol = 5e-3 * np.abs(w_) + 2e-2 * np.abs(w_).max() + 2.4e-7 * np.abs(p_)
        if (np.abs(g_ - w_) > tol).any():
            bad.append('%s: %d entries, worst %.3e of %.3e' % (k, int((np.abs(g_ - w_) > tol).sum()), np.abs(g_ - w_).max(), np.abs(w_).max()))
    assert not bad, '\n'.join(bad)
    assert float(np.abs(w - flat0).max()) > 2e-5                       # the parameters did move (~3 lr per step at most)


def test_sharded_gradients_combine_to_the_unsharded_gradient():
    """Data parallelism: the batch splits by sequences, noise is keyed by the global row, and the global
    gradient is the sequence-weighted mean of the shard gradients -- what the all-reduce computes."""
    dev = torch.device('cuda:0')
    cfg = O.Cfg(T=3, B=4, K=3, n=2)
    imgs, params, _ = TL.make_inputs(cfg)
    from sqair_b200 import ops
    model, obs, _ = _load_model(cfg, imgs, params, TL.make_inputs(cfg)[2], dev)
    full = model.compute_gradients(obs, seed=11).flat_grad.clone()
    full_elbo = float(model.last_scalars[1])
    half = O.Cfg(T=3, B=2, K=3, n=2)
    acc, elbo = torch.zeros_like(full), 0.
    for r in range(2):
        m, o, _ = _load_model(half, imgs[:, 2 * r:2 * r + 2], params, TL.make_inputs(half)[2], dev, row_offset=r * 2 * cfg.K)
        acc += m.compute_gradients(o, seed=11).flat_grad * 0.5
        elbo += 0.5 * float(m.last_scalars[1])
    assert elbo == pytest.approx(full_elbo, rel=1e-5)
    scale = float(full.abs().max())
    np.testing.assert_allclose(acc.cpu().numpy(), full.cpu().numpy(), rtol=1e-3, atol=2e-5 * scale)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_two_ranks_reproduce_one_rank_over_nccl():
    """N ranks with an NCCL all-reduce reproduce the single-rank ELBO, gradient and parameter update."""
    n = min(torch.cuda.device_count(), 8)
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
                        '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tools', 'dp_check.py')],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'DP_CHECK_OK' in r.stdout


def test_device_renderer_matches_the_host_generator():
    """sqair_render_sprites (data path) against the numpy renderer of the same tracks: identical pixels."""
    from sqair_b200 import data
    dev = torch.device('cuda:0')
    T, B, H, W, n = 6, 9, 50, 50, 3
    imgs, nums, coords, labels = data.moving_sprites(T, B, H, W, n, seed=5, return_tracks=True)
    got = data.render_on_device(coords, labels, nums, H, W, dev, seed=5).cpu().numpy()
    np.testing.assert_array_equal(got, imgs)
    assert imgs.max() > 0.5
