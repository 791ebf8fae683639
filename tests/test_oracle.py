"""CPU tests pinning the oracle (no GPU): the reference's own variable listing, third-party
semantics that can be checked here, and oracle-free invariants."""
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

import sqair_testlib as TL
from oracle import sqair_oracle as O
from oracle import synthetic as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_variable_inventory_matches_reference_notebook():
    """notebooks/play.ipynb:239-362 lists every trainable variable of the released model (n=3)."""
    ref = json.load(open(os.path.join(GOLD, 'ref_variables.json')))
    cfg = O.Cfg(n=ref['n_steps_per_image'])
    mine = {k: list(v) for k, v in O.param_shapes(cfg).items()}
    assert mine == ref['variables']
    assert O.param_count(cfg) == ref['total'] == 2951522
    by_scope = {}
    for k, s in mine.items():
        by_scope[k.split('/')[0]] = by_scope.get(k.split('/')[0], 0) + int(np.prod(s))
    for scope in ('decoder', 'discovery', 'model', 'propagation'):
        assert by_scope[scope] == ref['scope_counts_printed'][scope]
    assert by_scope['sequence'] == 14848 + 256 * 256      # the notebook prints the last variable separately


def test_stn_matches_grid_sample():
    torch.manual_seed(0)
    N, H, W, G = 6, 50, 40, 20
    img = torch.rand(N, H, W)
    coords = torch.stack([torch.rand(N) * 1.2 + .05, torch.rand(N) * 1.2 + .05,
                          torch.rand(N) * 1.6 - .8, torch.rand(N) * 1.6 - .8], -1)
    sx, sy, tx, ty = coords.unbind(-1)
    theta = torch.zeros(N, 2, 3)
    theta[:, 0, 0], theta[:, 0, 2], theta[:, 1, 1], theta[:, 1, 2] = sx, tx, sy, ty
    ref = F.grid_sample(img[:, None], F.affine_grid(theta, (N, 1, G, G), align_corners=True),
                        mode='bilinear', padding_mode='zeros', align_corners=True)[:, 0]
    assert (O.stn_forward(img, coords, G) - ref).abs().max() < 1e-5
    gl = torch.rand(N, G, G)
    theta = torch.zeros(N, 2, 3)
    theta[:, 0, 0], theta[:, 0, 2], theta[:, 1, 1], theta[:, 1, 2] = 1 / sx, -tx / sx, 1 / sy, -ty / sy
    ref = F.grid_sample(gl[:, None], F.affine_grid(theta, (N, 1, H, W), align_corners=True),
                        mode='bilinear', padding_mode='zeros', align_corners=True)[:, 0]
    assert (O.stn_inverse(gl, coords, H, W) - ref).abs().max() < 1e-5


def test_stn_identity_and_round_trip():
    img = torch.rand(3, 50, 50)
    ident = torch.tensor([[1., 1., 0., 0.]]).expand(3, -1)
    assert (O.stn_forward(img, ident, 50) - img).abs().max() < 1e-5
    # inverse(forward) reproduces the crop region (interior, same resolution)
    coords = torch.tensor([[0.4, 0.4, 0.1, -0.2]]).expand(3, -1)
    smooth = F.avg_pool2d(img[:, None], 9, 1, 4)[:, 0]
    g = O.stn_forward(smooth, coords, 20)
    back = O.stn_inverse(g, coords, 50, 50)
    fwd_again = O.stn_forward(back, coords, 20)
    assert (fwd_again[:, 2:-2, 2:-2] - g[:, 2:-2, 2:-2]).abs().max() < 0.05


def test_select_present_is_stable_partition():
    rng = np.random.default_rng(0)
    x = torch.arange(5 * 8 * 3, dtype=torch.float32).reshape(5, 8, 3)
    pres = torch.from_numpy((rng.random((5, 8)) < 0.5).astype(np.float32))
    y = O.select_present(x, pres)
    for b in range(5):
        want = [x[b, k] for k in range(8) if pres[b, k] == 1] + [x[b, k] for k in range(8) if pres[b, k] == 0]
        assert torch.equal(y[b], torch.stack(want))


def test_compute_object_ids():
    last = torch.tensor([[2.], [-1.]])
    prev = torch.tensor([[[0.], [2.]], [[-1.], [-1.]]])
    pp = torch.tensor([[[1.], [0.]], [[0.], [0.]]])
    dp = torch.tensor([[[1.], [1.]], [[1.], [0.]]])
    new_last, ids = O.compute_object_ids(last, prev, pp, dp)
    assert ids[..., 0].tolist() == [[0., -1., 3., 4.], [-1., -1., 0., -1.]]
    assert new_last[:, 0].tolist() == [4., 0.]


def test_num_steps_distribution_normalised_and_matches_definition():
    p = torch.rand(7, 4)
    joint = O.bernoulli_to_modified_geometric(p)
    assert torch.allclose(joint.sum(-1), torch.ones(7), atol=1e-6)
    pd = p.double()
    for i in range(5):
        want = torch.prod(pd[:, :i], -1) * ((1 - pd[:, i]) if i < 4 else 1.)
        assert torch.allclose(joint[:, i].double(), want, atol=1e-6)


def test_vimco_control_variate_is_leave_one_out():
    lw = torch.randn(6, 5) * 3
    cv = O.vimco_control_variate(lw)
    K = 5
    for j in range(K):
        others = [i for i in range(K) if i != j]
        repl = lw.clone()
        repl[:, j] = lw[:, others].mean(-1)
        want = torch.logsumexp(repl, -1) - np.log(K)
        assert torch.allclose(cv[:, j], want, atol=1e-5)
    assert torch.isnan(O.vimco_control_variate(torch.randn(3, 1))).all()       # K = 1: targets.py:55 divides by 0


def test_fill_triangular_tf_order():
    L = O.fill_triangular(torch.arange(10.))
    assert L.tolist() == [[4, 0, 0, 0], [8, 9, 0, 0], [7, 6, 5, 0], [3, 2, 1, 0]]


def test_gru_is_sonnet_not_torch():
    """snt.GRU applies the reset gate before Uh and mixes (1-z) h + z h~."""
    cfg = O.Cfg()
    p = O.init_params(cfg, 1, jitter=0.1)
    x, h = torch.randn(3, cfg.nw + 4), torch.randn(3, cfg.nh)
    s = 'propagation/gru_1'
    z = torch.sigmoid(x @ p[s + '/wz'] + h @ p[s + '/uz'] + p[s + '/bz'])
    r = torch.sigmoid(x @ p[s + '/wr'] + h @ p[s + '/ur'] + p[s + '/br'])
    c = torch.tanh(x @ p[s + '/wh'] + (r * h) @ p[s + '/uh'] + p[s + '/bh'])
    assert torch.allclose(O.gru(p, s, x, h), (1 - z) * h + z * c)


def test_iwae_tiling_order():
    x = torch.arange(2 * 3, dtype=torch.float32).reshape(2, 3, 1)
    t = O.tile_input_for_iwae(x, 2)
    assert t[0, :, 0].tolist() == [0, 0, 1, 1, 2, 2]          # row = b*K + k (index.py:106-129)


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    out = S.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.uint32(0xffffffff)
    out = S.philox4x32_10(f, f, f, f, 0xffffffff, 0xffffffff)
    assert [int(v) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = S.philox4x32_10(np.uint32(0x243f6a88), np.uint32(0x85a308d3), np.uint32(0x13198a2e), np.uint32(0x03707344),
                          0xa4093822, 0x299f31d0)
    assert [int(v) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_oracle_golden_fixture():
    """Regression fixture of the oracle itself (tests/golden/make_oracle_fixture.py) on BASELINE configs[0]."""
    fx = np.load(os.path.join(GOLD, 'oracle_c1.npz'))
    cfg = O.Cfg(T=3, B=4, K=1, n=2)
    imgs, params, noise = TL.make_inputs(cfg)
    out, obj = TL.run_oracle(cfg, imgs, params, noise)
    for k in ('presence', 'obj_id'):
        assert np.array_equal(out[k], fx[k]), k
    for k in ('what', 'where', 'canvas', 'log_weights_per_timestep', 'data_ll_per_sample', 'kl_per_sample', 'disc_prob'):
        np.testing.assert_allclose(out[k], fx[k], rtol=2e-4, atol=2e-4, err_msg=k)
    np.testing.assert_allclose(obj['elbo_iwae'], fx['elbo_iwae'], rtol=2e-4)


def test_synthetic_sequences_shape_and_range():
    imgs, nums = S.make_sequences(5, 6, 50, 50, 3, seed=3)
    assert imgs.shape == (5, 6, 50, 50) and imgs.dtype == np.float32
    assert 0. <= imgs.min() and imgs.max() <= 1. and (nums <= 3).all()
    empty = [b for b in range(6) if nums[b] == 0]
    for b in empty:
        assert imgs[:, b].max() == 0.


def test_oracle_gru_against_torch_grucell():
    """Independent form of the Sonnet GRU: h' = (1 - z) h + z c with z = sigmoid(x Wz + h Uz + bz),
    r = sigmoid(x Wr + h Ur + br), c = tanh(x Wh + (r h) Uh + bh).  torch.nn.GRUCell computes
    h' = (1 - z') n + z' h with n = tanh(W_in x + b_in + r (W_hn h + b_hn)): the two agree when z' = 1 - z -- i.e. with the
    update-gate weights negated -- provided the candidate uses (r * h) Uh, which GRUCell does NOT (it applies r after the
    matrix product).  So the check runs GRUCell's gates and rebuilds the candidate the Sonnet way from its own r."""
    torch.manual_seed(0)
    nin, nh, B = 7, 5, 4
    scope = 'g'
    p = {scope + '/' + k: torch.randn((nin if k[0] == 'w' else nh), nh) * 0.5 for k in ('wz', 'wr', 'wh', 'uz', 'ur', 'uh')}
    p.update({scope + '/' + k: torch.randn(nh) * 0.5 for k in ('bz', 'br', 'bh')})
    x, h = torch.randn(B, nin), torch.randn(B, nh)
    got = O.gru(p, scope, x, h)
    got = got[0] if isinstance(got, tuple) else got
    cell = torch.nn.GRUCell(nin, nh)
    with torch.no_grad():     # GRUCell rows: [r | z | n]; z' = 1 - z  <=>  negate the z weights
        cell.weight_ih.copy_(torch.cat((p['g/wr'].T, -p['g/wz'].T, p['g/wh'].T)))
        cell.weight_hh.copy_(torch.cat((p['g/ur'].T, -p['g/uz'].T, p['g/uh'].T)))
        cell.bias_ih.copy_(torch.cat((p['g/br'], -p['g/bz'], p['g/bh'])))
        cell.bias_hh.zero_()
        gi, gh = x @ cell.weight_ih.T + cell.bias_ih, h @ cell.weight_hh.T
        r = torch.sigmoid(gi[:, :nh] + gh[:, :nh])
        zc = torch.sigmoid(gi[:, nh:2 * nh] + gh[:, nh:2 * nh])                  # = 1 - z
        cand = torch.tanh(gi[:, 2 * nh:] + (r * h) @ p['g/uh'])                   # Sonnet: reset applied BEFORE the product
        want = (1 - zc) * cand + zc * h
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    # and the gates themselves are GRUCell's (same r, z' = 1 - z): with Uh = identity-free check of the torch form
    lin_cell = cell(x, h)
    n_torch = torch.tanh(gi[:, 2 * nh:] + r * gh[:, 2 * nh:])
    torch.testing.assert_close(lin_cell, (1 - zc) * n_torch + zc * h, rtol=1e-5, atol=1e-6)
