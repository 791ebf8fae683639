"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance: 1e-4 relative (+1e-4 absolute near zero) on every real-valued output, bit-exact on the
integer-valued ones (presence, object ids, step counts) -- BASELINE.json north_star.
"""
import os

import numpy as np
import pytest
import torch

import sqair_testlib as TL
from oracle import sqair_oracle as O
from oracle import synthetic as S

pytestmark = pytest.mark.gpu


def _gpu():
    if not torch.cuda.is_available():
        pytest.fail('no CUDA device: the gpu-marked tests need a B200')
    from sqair_b200 import ops
    return ops, torch.device('cuda:0')


def run_cuda(cfg, imgs, params, noise, rows_per_cta=None, cluster=None):
    ops, dev = _gpu()
    ccfg = TL.capi_cfg(cfg)
    old = os.environ.pop('SQAIR_ROWS_PER_CTA', None)
    oldc = os.environ.pop('SQAIR_CLUSTER', None)
    if rows_per_cta:
        os.environ['SQAIR_ROWS_PER_CTA'] = str(rows_per_cta)
    if cluster:
        os.environ['SQAIR_CLUSTER'] = str(cluster)
    try:
        flat = O.flatten_params(params, cfg).to(dev)
        packed = ops.pack_params(ccfg, flat)
        nz = {k: torch.from_numpy(v).to(dev) for k, v in noise.items()}
        pn = {k: nz[k + '_prior'] for k in ('eps_where', 'eps_what', 'u_pres')} if cfg.sample_from_prior else None
        out = ops.forward(ccfg, packed, torch.from_numpy(imgs).to(dev), nz, prior_noise=pn, generate_after=cfg.generate_after)
        obj = ops.objective(out['log_weights_per_timestep'], out['discrete_log_prob'], cfg.B, cfg.K)
        torch.cuda.synchronize()
    finally:
        os.environ.pop('SQAIR_ROWS_PER_CTA', None)
        os.environ.pop('SQAIR_CLUSTER', None)
        if old is not None:
            os.environ['SQAIR_ROWS_PER_CTA'] = old
        if oldc is not None:
            os.environ['SQAIR_CLUSTER'] = oldc
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res['_objective'] = obj['scalars'].cpu().numpy()
    return res


CASES = {
    'c1_T3_B4_K1_n2': dict(T=3, B=4, K=1, n=2),                       # BASELINE configs[0]
    'small_c2_T4_B3_K5_n4': dict(T=4, B=3, K=5, n=4),                 # configs[1] shape, fewer sequences
    'n3_K2': dict(T=3, B=3, K=2, n=3),
    'rw_prior': dict(T=3, B=2, K=2, n=2, prior_type='rw'),
    'guided_geom': dict(T=3, B=2, K=2, n=2, prior_type='guided', disc_prior_type='geom'),
    'no_rec_no_mask': dict(T=3, B=2, K=1, n=2, rec_where_prior=False, masked_glimpse=False),
    'bg_std_differs': dict(T=3, B=2, K=2, n=2, bg_std=0.5),             # modules.py:406-426,453: std = mask fg + (1 - mask) bg
    'c4_like_64px_n6': dict(T=2, B=2, K=2, n=6, H=64, W=64),
    # edges: a single slot, the maximum slot count, one frame, one row, non-square canvases (one of them with
    # H*W not a multiple of 4: frames are then read from global memory instead of the TMA-staged copy)
    'one_slot': dict(T=3, B=3, K=2, n=1),
    'max_slots_n8': dict(T=2, B=2, K=1, n=8),
    'single_frame_single_row': dict(T=1, B=1, K=1, n=2),
    'non_square_40x60': dict(T=2, B=2, K=2, n=2, H=40, W=60),
    'odd_pixels_45x35': dict(T=2, B=3, K=1, n=2, H=45, W=35),
    'seven_particles': dict(T=2, B=2, K=7, n=3),
}


@pytest.mark.parametrize('name', list(CASES))
def test_forward_parity(name):
    cfg = O.Cfg(**CASES[name])
    imgs, params, noise = TL.make_inputs(cfg)
    want, obj = TL.run_oracle(cfg, imgs, params, noise)
    got = run_cuda(cfg, imgs, params, noise)
    bad = TL.compare_outputs(got, want) + TL.compare_objective(got['_objective'], obj, cfg)
    assert not bad, '\n'.join(bad)


@pytest.mark.parametrize('R,C', [(1, 1), (2, 1), (3, 2), (4, 2), (5, 4), (2, 4), (1, 8), (3, 8), (6, 4), (6, 1), (4, 5), (2, 6), (3, 7)])
def test_launch_shape_variants(R, C):
    """Every instantiation of the persistent kernel (rows per cluster R) and every cluster size C,
    including row counts that do not divide."""
    cfg = O.Cfg(T=3, B=3, K=3, n=2)
    imgs, params, noise = TL.make_inputs(cfg)
    want, _ = TL.run_oracle(cfg, imgs, params, noise)
    got = run_cuda(cfg, imgs, params, noise, rows_per_cta=R, cluster=C)
    bad = TL.compare_outputs(got, want)
    assert not bad, '\n'.join(bad)


def test_full_size_c2_parity():
    """BASELINE configs[1]: T=10, B=32, K=5, n=4, 50x50 (the oracle needs a few seconds)."""
    cfg = O.Cfg(T=10, B=32, K=5, n=4)
    imgs, params, noise = TL.make_inputs(cfg)
    want, obj = TL.run_oracle(cfg, imgs, params, noise)
    got = run_cuda(cfg, imgs, params, noise)
    bad = TL.compare_outputs(got, want) + TL.compare_objective(got['_objective'], obj, cfg)   # ELBO-VAE / IWAE, ESS, targets
    assert not bad, '\n'.join(bad)
    assert 0.05 < want['presence'].mean() < 0.95     # both branches exercised


def test_full_size_c4_parity():
    """BASELINE configs[3]: 100x100 canvas, n=6 objects, B=16, K=10 (glimpse-bandwidth stress)."""
    cfg = O.Cfg(T=10, B=16, K=10, n=6, H=100, W=100)
    imgs, params, noise = TL.make_inputs(cfg)
    want, obj = TL.run_oracle(cfg, imgs, params, noise)
    got = run_cuda(cfg, imgs, params, noise)
    bad = TL.compare_outputs(got, want) + TL.compare_objective(got['_objective'], obj, cfg)   # ELBO-VAE / IWAE, ESS, targets
    assert not bad, '\n'.join(bad)


def test_c5_long_rollout_parity():
    """BASELINE configs[4]: seq_len=100 inference rollout, 2 objects, B=8, K=1.  One hundred frames of recurrence:
    the integer-valued decisions must still agree exactly and the real outputs within tolerance."""
    cfg = O.Cfg(T=100, B=8, K=1, n=2)
    imgs, params, noise = TL.make_inputs(cfg)
    want, obj = TL.run_oracle(cfg, imgs, params, noise)
    got = run_cuda(cfg, imgs, params, noise)
    bad = TL.compare_outputs(got, want) + TL.compare_objective(got['_objective'], obj, cfg)   # ELBO-VAE / IWAE, ESS, targets
    assert not bad, '\n'.join(bad)


GENERATION = {
    'generate_after_1': dict(T=4, B=3, K=2, n=3, sample_from_prior=True, generate_after=1),
    'prior_draws_only': dict(T=3, B=2, K=2, n=2, sample_from_prior=True),
    'guided_no_rec': dict(T=4, B=2, K=1, n=2, sample_from_prior=True, generate_after=2, prior_type='guided', rec_where_prior=False),
    'c2_shaped_rollout': dict(T=10, B=4, K=5, n=4, sample_from_prior=True, generate_after=3),
}


@pytest.mark.parametrize('name', list(GENERATION))
def test_generation_parity(name):
    """`SequentialAIR(..., sample_from_prior=True, generate_after=g)` (seq.py:46,198-203; sqair_modules.py:157-170,
    294-302): posterior evaluated at draws from the propagation prior; frames t > g roll forward from the prior and
    discover nothing."""
    cfg = O.Cfg(**GENERATION[name])
    imgs, params, noise = TL.make_inputs(cfg)
    noise = TL.with_prior_noise(cfg, noise)
    want, obj = TL.run_oracle(cfg, imgs, params, noise)
    got = run_cuda(cfg, imgs, params, noise)
    bad = TL.compare_outputs(got, want) + TL.compare_objective(got['_objective'], obj, cfg)
    assert not bad, '\n'.join(bad)
    if cfg.generate_after > 0:
        assert (got['disc_pres'][cfg.generate_after + 1:] == 0).all()


def test_one_packed_buffer_serves_calls_with_different_rows_per_cluster():
    """The packed parameters depend on the cluster size only; the per-call layer table (shared-memory offsets, strides,
    frame staging) is NOT part of them.  One buffer, two batch sizes that the library runs with different rows per
    cluster: both must match the oracle."""
    ops, dev = _gpu()
    from sqair_b200 import _capi
    seen = {}
    for B in (1, 2, 3, 4, 6):
        s = _capi.query_sizes(TL.capi_cfg(O.Cfg(T=2, B=B, K=5, n=2)))
        seen.setdefault(s.cluster_size, {}).setdefault(s.rows_per_cta, B)
    pair = next(((c, list(r.values())) for c, r in seen.items() if len(r) >= 2), None)
    assert pair is not None, 'no two batch sizes share a cluster size with different rows per cluster: %r' % seen
    (Ba, Bb) = pair[1][:2]
    cfg_a, cfg_b = O.Cfg(T=2, B=Ba, K=5, n=2), O.Cfg(T=2, B=Bb, K=5, n=2)
    imgs, params, noise = TL.make_inputs(cfg_a if Ba > Bb else cfg_b)
    flat = O.flatten_params(params, cfg_a).to(dev)
    packed = ops.pack_params(TL.capi_cfg(cfg_a), flat)                  # packed ONCE
    for cfg in (cfg_a, cfg_b, cfg_a):
        im = np.ascontiguousarray(imgs[:, :cfg.B])
        nz = {k: np.ascontiguousarray(v[:, :cfg.B * cfg.K]) for k, v in noise.items()}
        want, _ = TL.run_oracle(cfg, im, params, nz)
        out = ops.forward(TL.capi_cfg(cfg), packed, torch.from_numpy(im).to(dev), {k: torch.from_numpy(v).to(dev) for k, v in nz.items()})
        torch.cuda.synchronize()
        bad = TL.compare_outputs({k: v.cpu().numpy() for k, v in out.items()}, want)
        assert not bad, 'B=%d: %s' % (cfg.B, '\n'.join(bad))


def test_device_noise_matches_numpy_philox():
    ops, dev = _gpu()
    cfg = O.Cfg(T=3, B=4, K=2, n=3)
    ccfg = TL.capi_cfg(cfg)
    for off in (0, 5):
        nz = ops.fill_noise(ccfg, seed=0x1234567890abcdef, row_offset=off, device=dev)
        torch.cuda.synchronize()
        ref = S.philox_noise(cfg.T, cfg.rows, cfg.n, cfg.nw, 0x1234567890abcdef, row_offset=off)
        assert np.array_equal(nz['u_pres'].cpu().numpy(), ref['u_pres'])              # uniforms are exact
        np.testing.assert_allclose(nz['eps_where'].cpu().numpy(), ref['eps_where'], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(nz['eps_what'].cpu().numpy(), ref['eps_what'], rtol=1e-4, atol=2e-5)


def test_sharded_rows_reproduce_unsharded_run():
    """Multi-GPU invariant: sequences are independent, so running two shards of the batch
    with row_offset-keyed noise must reproduce the single-call result exactly."""
    ops, dev = _gpu()
    cfg = O.Cfg(T=3, B=4, K=3, n=2)
    imgs, params, _ = TL.make_inputs(cfg)
    ccfg = TL.capi_cfg(cfg)
    flat = O.flatten_params(params, cfg).to(dev)
    packed = ops.pack_params(ccfg, flat)
    obs = torch.from_numpy(imgs).to(dev)
    full = ops.forward(ccfg, packed, obs, ops.fill_noise(ccfg, 99, 0, device=dev))
    half = O.Cfg(T=3, B=2, K=3, n=2)
    hcfg = TL.capi_cfg(half)
    packed_h = ops.pack_params(hcfg, flat)          # the packed layout follows the launch shape of the call
    parts = []
    for s in range(2):
        o = obs[:, 2 * s:2 * s + 2].contiguous()
        parts.append(ops.forward(hcfg, packed_h, o, ops.fill_noise(hcfg, 99, s * half.rows, device=dev)))
    torch.cuda.synchronize()
    for k in full:
        cat = torch.cat([p[k] for p in parts], 1)
        assert torch.equal(cat, full[k]), k


def test_stn_glimpse_op():
    ops, dev = _gpu()
    rng = np.random.default_rng(0)
    N, H, W, G = 37, 50, 50, 20
    img = rng.random((N, H, W), dtype=np.float32)
    where = (rng.standard_normal((N, 4)) * 1.5).astype(np.float32)
    got = ops.stn_glimpse(torch.from_numpy(img).to(dev), torch.from_numpy(where).to(dev), G).cpu().numpy()
    want = O.stn_forward(torch.from_numpy(img), O.to_coords(torch.from_numpy(where)), G).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # identity transform copies the image when G == H (invariant)
    ident = np.tile(np.array([[20., 20., 0., 0.]], dtype=np.float32), (N, 1))
    got = ops.stn_glimpse(torch.from_numpy(img).to(dev), torch.from_numpy(ident).to(dev), H).cpu().numpy()
    np.testing.assert_allclose(got, img, rtol=0, atol=1e-5)


def test_canvas_ll_op():
    ops, dev = _gpu()
    rng = np.random.default_rng(1)
    N, n, H, W, G = 9, 4, 50, 50, 20
    cfg = O.Cfg(T=1, B=N, K=1, n=n)
    p = O.init_params(cfg, 3, mean_img=rng.random((H, W)) * 0.2)
    what = torch.from_numpy(rng.standard_normal((N, n, cfg.nw)).astype(np.float32))
    where = torch.from_numpy((rng.standard_normal((N, n, 4))).astype(np.float32))
    pres = torch.from_numpy((rng.random((N, n, 1)) < 0.6).astype(np.float32))
    img = torch.from_numpy(rng.random((N, H, W), dtype=np.float32))
    canvas, std, glimpse = O.air_decoder(p, cfg, what, where, pres)
    want_ll = O.normal_log_prob(img, canvas, std).sum((1, 2)).numpy()
    got_canvas, got_ll = ops.canvas_ll(glimpse.contiguous().to(dev), where.to(dev), pres[..., 0].contiguous().to(dev),
                                       p['decoder/air_decoder/Variable'][..., 0].contiguous().to(dev), img.to(dev))
    np.testing.assert_allclose(got_canvas.cpu().numpy(), canvas.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got_ll.cpu().numpy(), want_ll, rtol=1e-4, atol=1e-3)


def test_objective_op():
    ops, dev = _gpu()
    from sqair_b200 import _capi
    rng = np.random.default_rng(2)
    T, B, K = 10, 32, 5
    lw_t = (rng.standard_normal((T, B * K)) * 3 + 50).astype(np.float32)
    lp_t = (rng.standard_normal((T, B * K))).astype(np.float32)
    got = ops.objective(torch.from_numpy(lw_t).to(dev), torch.from_numpy(lp_t).to(dev), B, K)
    lw = torch.from_numpy(lw_t).sum(0).reshape(B, K)
    lp = torch.from_numpy(lp_t).sum(0).reshape(B, K)
    np.testing.assert_allclose(got['log_weights'].cpu().numpy(), lw.numpy(), rtol=1e-5)
    np.testing.assert_allclose(got['elbo_iwae_per_example'].cpu().numpy(), O.iwae(lw).numpy(), rtol=1e-5)
    np.testing.assert_allclose(got['importance_weights'].cpu().numpy(), torch.softmax(lw, -1).numpy(),
                               rtol=1e-4, atol=1e-6)
    sc = got['scalars'].cpu().numpy()
    np.testing.assert_allclose(sc[_capi.OBJ_ELBO_VAE], lw.mean().item(), rtol=1e-5)
    np.testing.assert_allclose(sc[_capi.OBJ_ELBO_IWAE], O.iwae(lw).mean().item(), rtol=1e-5)
    np.testing.assert_allclose(sc[_capi.OBJ_ESS], O.ess(torch.softmax(lw, -1)).mean().item(), rtol=1e-4)
    np.testing.assert_allclose(sc[_capi.OBJ_VIMCO_TARGET], (O.vimco(lw, lp, O.iwae(lw)) / T).item(), rtol=1e-4)
    np.testing.assert_allclose(sc[_capi.OBJ_IWAE_TARGET], (-O.iwae(lw).mean() / T).item(), rtol=1e-5)


def test_objective_grad_op():
    """First stage of the backward pass: d(VIMCO target)/d(log weights, discrete log-probs) against torch autograd
    on the oracle's objective (targets.py:46-75 semantics incl. the stop-gradient on the learning signal)."""
    ops, dev = _gpu()
    rng = np.random.default_rng(3)
    T, B, K = 10, 32, 5
    lw_t = (rng.standard_normal((T, B * K)) * 3 + 50).astype(np.float32)
    lp_t = (rng.standard_normal((T, B * K))).astype(np.float32)
    d_lw, d_lp = ops.objective_grad(torch.from_numpy(lw_t).to(dev), torch.from_numpy(lp_t).to(dev), B, K)
    a = torch.from_numpy(lw_t).requires_grad_(True)
    b = torch.from_numpy(lp_t).requires_grad_(True)
    lw, lp = a.sum(0).reshape(B, K), b.sum(0).reshape(B, K)
    (O.vimco(lw, lp, O.iwae(lw)) / T).backward()
    for t in range(T):          # every frame of a row receives the row's gradient
        np.testing.assert_allclose(d_lw.cpu().numpy().reshape(-1), a.grad[t].numpy(), rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(d_lp.cpu().numpy().reshape(-1), b.grad[t].numpy(), rtol=1e-4, atol=1e-8)


def test_stn_glimpse_grad_op():
    """Glimpse-sampler backward w.r.t. the where-logits against torch autograd through the oracle's transformer
    (resampler warp gradient, modules.py:165-227).  Scales stay above the 1e-4 floor, where the oracle's clamp and
    the reference's straight-through clip have the same gradient."""
    ops, dev = _gpu()
    rng = np.random.default_rng(4)
    N, H, W, G = 53, 50, 50, 20
    img = torch.from_numpy(rng.random((N, H, W), dtype=np.float32))
    where = torch.from_numpy((rng.standard_normal((N, 4)) * 1.2).astype(np.float32)).requires_grad_(True)
    dg = torch.from_numpy(rng.standard_normal((N, G, G)).astype(np.float32))
    O.stn_forward(img, O.to_coords(where), G).backward(dg)
    got = ops.stn_glimpse_grad(img.to(dev), where.detach().to(dev), dg.to(dev)).cpu().numpy()
    want = where.grad.numpy()
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-3 * np.abs(want).max())


def test_canvas_ll_grad_op():
    """Canvas composition + pixel likelihood backward (inverse-warp gradient w.r.t. data and warp, mean-image mask)
    against torch autograd through the oracle's composition (modules.py:435-467; seq.py:272-273)."""
    ops, dev = _gpu()
    rng = np.random.default_rng(5)
    N, n, H, W, G = 11, 3, 50, 50, 20
    glimpse = torch.from_numpy((rng.random((N, n, G, G)) * 0.5).astype(np.float32)).requires_grad_(True)
    where = torch.from_numpy((rng.standard_normal((N, n, 4)) * 0.8).astype(np.float32)).requires_grad_(True)
    pres = torch.from_numpy((rng.random((N, n)) < 0.7).astype(np.float32))
    mean_img = torch.from_numpy((rng.random((H, W)) * 0.2).astype(np.float32)).requires_grad_(True)
    img = torch.from_numpy(rng.random((N, H, W), dtype=np.float32))
    d_ll = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    coords = O.to_coords(where).reshape(N * n, 4)
    pr = pres.reshape(N, n, 1, 1)
    canvas = (O.stn_inverse(glimpse.reshape(N * n, G, G), coords, H, W).reshape(N, n, H, W) * pr).sum(1)
    nz = (O.stn_inverse(torch.ones(N * n, G, G), coords, H, W).reshape(N, n, H, W) * pr).sum(1)
    mask = torch.sigmoid(-10. + nz * 20.)
    canvas = canvas + mean_img[None] * mask
    std = float(np.float32(np.sqrt(np.float32(0.3))) ** 2)
    ll = O.normal_log_prob(img, canvas, mask * std + (1. - mask) * std).sum((1, 2))
    ll.backward(d_ll)
    d_gl, d_wh, d_mi = ops.canvas_ll_grad(glimpse.detach().to(dev), where.detach().to(dev), pres.to(dev),
                                          mean_img.detach().to(dev), img.to(dev), d_ll.to(dev))
    for got, want in ((d_gl, glimpse.grad), (d_wh, where.grad), (d_mi, mean_img.grad)):
        w = want.numpy()
        np.testing.assert_allclose(got.cpu().numpy(), w, rtol=2e-3, atol=2e-3 * np.abs(w).max())


@pytest.mark.parametrize('M,K,N', [(6400, 672, 256), (1600, 264, 109), (37, 5, 3), (640, 400, 256)])
def test_wgrad_op(M, K, N):
    """Weight-gradient GEMM of the backward pass (x^T dy; tcgen05 3xTF32 path for TMA-addressable shapes, else the mma.sync
    split kernel) against float64.  Tolerance: 1e-5 relative + 8e-6 of the typical magnitude sqrt(M) * 0.1 -- the TMEM
    accumulator rounds toward zero on each of the ~100 chained MMAs of a CTA (measured worst case 5e-6 of that scale)."""
    ops, dev = _gpu()
    rng = np.random.default_rng(6)
    x = rng.standard_normal((M, K)).astype(np.float32)
    dy = (rng.standard_normal((M, N)) * 0.1).astype(np.float32)
    want = x.astype(np.float64).T @ dy.astype(np.float64)
    got = ops.wgrad(torch.from_numpy(x).to(dev), torch.from_numpy(dy).to(dev)).cpu().numpy()
    scale = np.sqrt(M) * 0.1
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=8e-6 * scale)
    base = torch.from_numpy(rng.standard_normal((K, N)).astype(np.float32)).to(dev)
    got2 = ops.wgrad(torch.from_numpy(x).to(dev), torch.from_numpy(dy).to(dev), out=base.clone()).cpu().numpy()
    np.testing.assert_allclose(got2, want + base.cpu().numpy().astype(np.float64), rtol=1e-5, atol=8e-6 * scale + 1e-6)


@pytest.mark.parametrize('M,K,N', [(6400, 672, 256), (1600, 264, 109), (37, 5, 3), (640, 400, 256)])
def test_dgrad_op(M, K, N):
    """Input-gradient GEMM of the backward pass (dy w^T, same split tensor-core arithmetic) against float64."""
    ops, dev = _gpu()
    rng = np.random.default_rng(7)
    dy = (rng.standard_normal((M, N)) * 0.1).astype(np.float32)
    w = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    want = dy.astype(np.float64) @ w.astype(np.float64).T
    got = ops.dgrad(torch.from_numpy(dy).to(dev), torch.from_numpy(w).to(dev)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6 * np.sqrt(N) * 0.1 / np.sqrt(K))
