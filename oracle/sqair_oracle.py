"""CPU oracle for SQAIR's per-frame Discover/Propagate hot path.

TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared with; nothing in
`sqair_b200/` imports it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import or execute anything under `oracle/`.

PARITY UNPINNED against real TF1: the reference (akosiorek/sqair @ 474f5d0) is Python-2 /
TensorFlow-1.6 / Sonnet-1.14 and cannot be imported in this environment, and it ships no tests or
golden vectors.  This is an op-for-op torch-CPU fp32 restatement of the reference graph, every
function citing the reference file:line it follows; TF/Sonnet semantics (un-vendored third-party
code: tensorflow==1.6.0, dm_sonnet==1.14) are restated from their documented behaviour
(SURVEY.md Appendix B).  What *is* pinned: the variable names / shapes / per-scope parameter
counts printed by the reference's own notebook (`notebooks/play.ipynb:239-362`,
tests/golden/ref_variables.json) and the STN forward/inverse formulas against
`torch.nn.functional.grid_sample(align_corners=True, padding_mode='zeros')`.

Randomness: TF's Philox op streams cannot be reproduced outside TF, so every random draw is an
explicit input ("identical seeds" == identical eps/u tensors):
    noise['eps_where'] [T, B', 2n, 4]   noise['eps_what'] [T, B', 2n, nw]   noise['u_pres'] [T, B', 2n]
slots 0..n-1 are the propagation draws, n..2n-1 the discovery draws.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------
# configuration (flags of sqair/common_model_flags.py:32-56 and sqair/configs/mlp_mnist_model.py:42-52)
# --------------------------------------------------------------------------------------------
@dataclass
class Cfg:
    T: int = 3
    B: int = 4
    K: int = 1                      # k_particles
    n: int = 2                      # n_steps_per_image
    H: int = 50
    W: int = 50
    G: int = 20                     # glimpse_size
    nw: int = 50                    # n_what
    nh: int = 256                   # 32 * n_units
    prior_type: str = 'rnn'         # prop_prior_type: rnn | rw | guided (propagate.py:35-45)
    disc_prior_type: str = 'cat'    # cat | geom (sqair_modules.py:205-224)
    rec_where_prior: bool = True
    masked_glimpse: bool = True
    step_success_prob: float = 0.75
    prop_prior_step_bias: float = 10.0
    output_std: float = 0.3
    bg_std: float = None            # modules.py:406-407: None -> output_std
    where_update_scale: float = 1.0
    min_std: float = 1e-2
    where_mean: tuple = (-2., -2., 0., 0.)   # only used when rec_where_prior is False
    where_std: tuple = (1., 1., 1., 1.)
    sample_from_prior: bool = False          # seq.py:46,63: evaluate q at prior samples; with generate_after also ...
    generate_after: int = -1                 # ... replace the latents by prior samples in frames t > generate_after (seq.py:198-203)

    @property
    def P(self):
        return self.H * self.W

    @property
    def g(self):
        return self.G * self.G

    @property
    def rows(self):
        return self.B * self.K


# --------------------------------------------------------------------------------------------
# parameter inventory: TF variable names / shapes (notebooks/play.ipynb:239-362)
# --------------------------------------------------------------------------------------------
_RN = 'discovery/discover/recurrent_normal_impl/'


def param_shapes(cfg: Cfg) -> "OrderedDict[str, tuple]":
    n, nw, nh, g, P = cfg.n, cfg.nw, cfg.nh, cfg.g, cfg.P
    s = nh // 2
    d = OrderedDict()

    def lin(name, i, o):
        d[name + '/w'] = (i, o)
        d[name + '/b'] = (o,)

    # decoder scope (modules.py:131-147,399)
    d['decoder/air_decoder/Variable'] = (cfg.H, cfg.W, 1)
    lin('decoder/air_decoder/decoder/mlp/linear', nw, nh)
    lin('decoder/air_decoder/decoder/mlp/linear_1', nh, nh)
    lin('decoder/air_decoder/decoder/mlp/linear_2', nh, g)
    d['decoder/air_decoder/decoder/output_scale'] = ()
    # discovery scope
    d['discovery/discover/discovery/vanilla_rnn_initial_state_0/w'] = (1, nh)       # core.py:130
    lin('discovery/discover/mlp/linear', 1, 10)                                      # sqair_modules.py:218
    lin('discovery/discover/mlp/linear_1', 10, n + 1)
    if cfg.rec_where_prior:                                                          # modules.py:559-565
        d[_RN + 'discovery/discover/recurrent_normal_impl/vanilla_rnn_initial_state_0/w'] = (1, 4)
        d[_RN + 'init_sample'] = (1, 4)
        lin(_RN + 'linear', 4, 8)
        lin(_RN + 'linear_1', 4 + nh + 1, 128)
        lin(_RN + 'vanilla_rnn/hidden_to_hidden', 128, 4)
        lin(_RN + 'vanilla_rnn/in_to_hidden', 4, 4)
    lin('discovery/discovery_core/air_encoder/gaussian_from_param_vec/linear', nh, 2 * nw)
    if cfg.masked_glimpse:                                                           # modules.py:322-324
        lin('discovery/discovery_core/air_encoder/mlp/linear', nh, 128)
        lin('discovery/discovery_core/air_encoder/mlp/linear_1', 128, g)
    lin('discovery/discovery_core/encoder/mlp/linear', P, nh)                        # image encoder
    lin('discovery/discovery_core/encoder/mlp/linear_1', nh, nh)
    lin('discovery/discovery_core/encoder_1/mlp/linear', g, nh)                      # glimpse encoder (shared)
    lin('discovery/discovery_core/encoder_1/mlp/linear_1', nh, nh)
    lin('discovery/discovery_core/steps_predictor/mlp/linear', nh + nw, s)
    lin('discovery/discovery_core/steps_predictor/mlp/linear_1', s, 1)
    lin('discovery/discovery_core/stochastic_transform_param/mlp/linear', nh, nh)
    lin('discovery/discovery_core/stochastic_transform_param/mlp/linear_1', nh, nh)
    lin('discovery/discovery_core/stochastic_transform_param/mlp/linear_2', nh, 8)
    d['discovery/discovery_core/stochastic_transform_param/scale_offset'] = ()
    lin('discovery/vanilla_rnn/hidden_to_hidden', nh, nh)
    lin('discovery/vanilla_rnn/in_to_hidden', 2 * nh + nw + 5, nh)
    # model scope (sqair_modules.py:209-214)
    d['model/sequential_air/while/sqair_timestep/discover/step_prior_bias'] = (n + 1,)
    d['model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias'] = (n + 1,)
    # propagation scope
    for scope, nin in (('propagation/gru', nh + 4 + 2 * nw), ('propagation/gru_1', nw + 4)):
        for gate in 'zrh':
            d['%s/w%s' % (scope, gate)] = (nin, nh)
            d['%s/u%s' % (scope, gate)] = (nh, nh)
            d['%s/b%s' % (scope, gate)] = (nh,)
    lin('propagation/propagate_prior/linear', nh, 2 * (4 + nw) + 1)
    d['propagation/propagation_core/affine_diag_normal/cholesky_scale'] = (10,)
    lin('propagation/propagation_core/rnn_inpt/mlp/linear', nh, 128)
    lin('propagation/propagation_core/rnn_inpt/mlp/linear_1', 128, 4)
    lin('propagation/propagation_core/steps_predictor/mlp/linear', 2 * nh + nw, s)
    lin('propagation/propagation_core/steps_predictor/mlp/linear_1', s, 1)
    lin('propagation/propagation_core/stochastic_transform_param/mlp/linear', 2 * nh + 4, nh)
    lin('propagation/propagation_core/stochastic_transform_param/mlp/linear_1', nh, nh)
    lin('propagation/propagation_core/stochastic_transform_param/mlp/linear_2', nh, 8)
    d['propagation/propagation_core/stochastic_transform_param/scale_offset'] = ()
    lin('propagation/propagation_core/what/gaussian_from_param_vec/linear', nh, 2 * nw)
    lin('propagation/propagation_core/what/linear', nh, 3 * nw)
    d['propagation/sequential_ssm/propagation/vanilla_rnn_initial_state_0/w'] = (1, nh)
    lin('propagation/vanilla_rnn/hidden_to_hidden', nh, nh)
    lin('propagation/vanilla_rnn/in_to_hidden', 3 * nw + 10 + nh, nh)
    # sequence scope
    d['sequence/sequential_air/propagation/gru_1_initial_state_0/w'] = (1, nh)       # prior h0
    d['sequence/sequential_air/propagation/gru_initial_state_0/w'] = (1, nh)         # temporal h0
    lin('sequence/sequential_air/sqair_timestep/mlp/linear', nw + 4, nh)
    lin('sequence/sequential_air/sqair_timestep/mlp/linear_1', nh, nh)
    return d


def param_count(cfg: Cfg) -> int:
    return int(sum(int(np.prod(s)) for s in param_shapes(cfg).values()))


def init_params(cfg: Cfg, seed: int = 42, mean_img=None, jitter: float = 0.0):
    """Sonnet-style initial values: w ~ TruncNormal(0, 1/sqrt(fan_in)) (+-2 sigma),
    b = 0, GRU matrices glorot-uniform, trainable initial states 0, and the constant initialisers of
    the reference (configs/mlp_mnist_model.py:42-52, core.py:345, modules.py:323-324,
    sqair_modules.py:81-83,209-214).  `jitter` > 0 perturbs biases / initial states so that parity
    tests exercise every term with non-trivial values."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        base = name.rsplit('/', 1)[-1]
        if base in ('wz', 'wr', 'wh', 'uz', 'ur', 'uh'):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-lim, lim, shape)
        elif base == 'w' and 'initial_state' not in name:
            v = rng.standard_normal(shape)
            bad = np.abs(v) > 2.0
            while bad.any():
                v[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(v) > 2.0
            v = v / math.sqrt(shape[0])
        else:
            v = np.zeros(shape)
        out[name] = v
    out['decoder/air_decoder/decoder/output_scale'] = np.asarray(0.25)
    out['discovery/discovery_core/stochastic_transform_param/scale_offset'] = np.asarray(-3.0)
    out['propagation/propagation_core/stochastic_transform_param/scale_offset'] = np.asarray(-3.0)
    out['discovery/discovery_core/steps_predictor/mlp/linear_1/b'][:] = 1.0      # disc_step_bias
    out['propagation/propagation_core/steps_predictor/mlp/linear_1/b'][:] = 5.0  # prop_step_bias
    out['propagation/propagation_core/what/linear/b'][:] = 1.0                   # remember_bias core.py:345
    if cfg.masked_glimpse:
        out['discovery/discovery_core/air_encoder/mlp/linear_1/b'][:] = 1.0      # modules.py:324
    if cfg.rec_where_prior:
        out[_RN + 'linear/b'][:] = np.asarray(list(cfg.where_mean) + list(cfg.where_std))
        out[_RN + 'init_sample'] = rng.uniform(-0.5, 0.5, (1, 4))
    out['model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias'][0] = 10.0
    out['propagation/propagation_core/affine_diag_normal/cholesky_scale'] = rng.uniform(-0.3, 0.3, (10,))
    if mean_img is not None:
        out['decoder/air_decoder/Variable'] = np.asarray(mean_img, dtype=np.float64).reshape(cfg.H, cfg.W, 1)
    if jitter > 0:
        for name, v in out.items():
            base = name.rsplit('/', 1)[-1]
            if base in ('b', 'bz', 'br', 'bh', 'step_prior_bias') or 'initial_state' in name:
                out[name] = v + jitter * rng.standard_normal(v.shape)
    return OrderedDict((k, torch.tensor(np.asarray(v), dtype=torch.float32)) for k, v in out.items())


def flatten_params(params, cfg: Cfg) -> torch.Tensor:
    """Canonical flat layout: variables in `param_shapes` order, each row-major."""
    return torch.cat([params[k].reshape(-1).float() for k in param_shapes(cfg)])


def unflatten_params(flat, cfg: Cfg):
    out, off = OrderedDict(), 0
    for k, s in param_shapes(cfg).items():
        num = int(np.prod(s))
        out[k] = flat[off:off + num].reshape(s)
        off += num
    return out


# --------------------------------------------------------------------------------------------
# L0 semantics: snt.Linear / Nonlinear / MLP / VanillaRNN / GRU (neural.py:34-116; Appendix B)
# --------------------------------------------------------------------------------------------
def linear(p, name, x):
    return x @ p[name + '/w'] + p[name + '/b']


def mlp(p, scope, x, n_layers, out_transfer=None):
    """neural.py:111-116 MLP: ELU on hidden layers, `out_transfer` (or none) on the last one."""
    for i in range(n_layers):
        name = scope + '/linear' + ('' if i == 0 else '_%d' % i)
        x = linear(p, name, x)
        if i < n_layers - 1:
            x = F.elu(x)
        elif out_transfer is not None:
            x = out_transfer(x)
    return x


def mlp_hidden(p, scope, x, n_layers):
    """MLP with no output layer: every layer has ELU (modules.py:108-112 Encoder)."""
    for i in range(n_layers):
        name = scope + '/linear' + ('' if i == 0 else '_%d' % i)
        x = F.elu(linear(p, name, x))
    return x


def vanilla_rnn(p, scope, x, h):
    """snt.VanillaRNN: tanh(in_to_hidden(x) + hidden_to_hidden(h))."""
    return torch.tanh(linear(p, scope + '/in_to_hidden', x) + linear(p, scope + '/hidden_to_hidden', h))


def gru(p, scope, x, h):
    """snt.GRU (Appendix B): reset gate applied before Uh; h' = (1-z) h + z h~."""
    z = torch.sigmoid(x @ p[scope + '/wz'] + h @ p[scope + '/uz'] + p[scope + '/bz'])
    r = torch.sigmoid(x @ p[scope + '/wr'] + h @ p[scope + '/ur'] + p[scope + '/br'])
    c = torch.tanh(x @ p[scope + '/wh'] + (r * h) @ p[scope + '/uh'] + p[scope + '/bh'])
    return (1. - z) * h + z * c


def normal_log_prob(x, loc, scale):
    """tfd.Normal.log_prob."""
    return -0.5 * ((x - loc) / scale) ** 2 - torch.log(scale) - 0.5 * LOG_2PI


def bernoulli_log_prob(x, logits):
    """tfd.Bernoulli(logits).log_prob(x) = -sigmoid_cross_entropy(labels=x, logits)."""
    return -(torch.clamp(logits, min=0.) - logits * x + torch.log1p(torch.exp(-torch.abs(logits))))


# --------------------------------------------------------------------------------------------
# spatial transformer (modules.py:150-227) -- snt.AffineGridWarper + tf.contrib.resampler
# --------------------------------------------------------------------------------------------
def clip_preserve(x, lo, hi):
    """ops.py:33-42: clipped value, identity gradient (stop_gradient(clipped - x) + x)."""
    return x + (x.clamp(lo, hi) - x).detach()


def to_coords(logits):
    """modules.py:220-227: (scale, shift) = (sigmoid(l[:2]), tanh(l[2:]))."""
    return torch.cat((torch.sigmoid(logits[..., :2]), torch.tanh(logits[..., 2:])), -1)


def _bilinear_zero_pad(img, x, y):
    """tf.contrib.resampler semantics (Appendix B): img [N,h,w]; x,y [N,M] pixel coordinates."""
    N, h, w = img.shape
    fx, fy = torch.floor(x), torch.floor(y)
    cx, cy = fx + 1, fy + 1
    dx, dy = cx - x, cy - y

    def fetch(ix, iy):
        ok = (ix >= 0) & (ix <= w - 1) & (iy >= 0) & (iy <= h - 1)
        ixc = ix.clamp(0, w - 1).long()
        iyc = iy.clamp(0, h - 1).long()
        v = img.reshape(N, h * w).gather(1, iyc * w + ixc)
        return torch.where(ok, v, torch.zeros_like(v))

    out = dx * dy * fetch(fx, fy) + (1 - dx) * (1 - dy) * fetch(cx, cy) \
        + dx * (1 - dy) * fetch(fx, cy) + (1 - dx) * dy * fetch(cx, fy)
    inside = (x > -1) & (y > -1) & (x < w) & (y < h)
    return torch.where(inside, out, torch.zeros_like(out))


def stn_forward(img, coords, G):
    """modules.py:165-172,204-218 (forward warper): img [N,H,W], coords [N,4]=(sx,sy,tx,ty) ->
    glimpse [N,G,G].  x_pix = (W-1)/2 (sx u + tx) + (W-1)/2, u in linspace(-1,1,G)."""
    N, H, W = img.shape
    sx, sy, tx, ty = coords.unbind(-1)
    sx, sy = clip_preserve(sx, 1e-4, None), clip_preserve(sy, 1e-4, None)      # modules.py:206; ops.py:33-42
    lin = torch.linspace(-1., 1., G, dtype=img.dtype)
    xs = (W - 1) / 2. * (sx[:, None] * lin[None] + tx[:, None]) + (W - 1) / 2.       # [N,G]
    ys = (H - 1) / 2. * (sy[:, None] * lin[None] + ty[:, None]) + (H - 1) / 2.
    x = xs[:, None, :].expand(N, G, G).reshape(N, G * G)
    y = ys[:, :, None].expand(N, G, G).reshape(N, G * G)
    return _bilinear_zero_pad(img, x, y).reshape(N, G, G)


def stn_inverse(glimpse, coords, H, W):
    """modules.py:167-168 (`warper.inverse()`): glimpse [N,G,G] -> [N,H,W];
    x_g = (G-1)/2 ((u - tx)/sx) + (G-1)/2 for u in linspace(-1,1,W)."""
    N, G, _ = glimpse.shape
    sx, sy, tx, ty = coords.unbind(-1)
    sx, sy = clip_preserve(sx, 1e-4, None), clip_preserve(sy, 1e-4, None)
    lx = torch.linspace(-1., 1., W, dtype=glimpse.dtype)
    ly = torch.linspace(-1., 1., H, dtype=glimpse.dtype)
    xs = (G - 1) / 2. * ((lx[None] - tx[:, None]) / sx[:, None]) + (G - 1) / 2.       # [N,W]
    ys = (G - 1) / 2. * ((ly[None] - ty[:, None]) / sy[:, None]) + (G - 1) / 2.       # [N,H]
    x = xs[:, None, :].expand(N, H, W).reshape(N, H * W)
    y = ys[:, :, None].expand(N, H, W).reshape(N, H * W)
    return _bilinear_zero_pad(glimpse, x, y).reshape(N, H, W)


# --------------------------------------------------------------------------------------------
# glimpse encoder (modules.py:306-364 AIREncoder + Encoder + GaussianFromParamVec)
# --------------------------------------------------------------------------------------------
_DC = 'discovery/discovery_core/'


def glimpse_mask(p, mask_inpt):
    """modules.py:322-324,350-356: sigmoid(MLP(128 -> g)(temporal_state))."""
    return mlp(p, _DC + 'air_encoder/mlp', mask_inpt, 2, torch.sigmoid)


def air_encoder(p, cfg, img, where, mask_inpt=None):
    """modules.py:326-364: extract glimpse at to_coords(where), optional mask, encode ->
    (loc, scale) of Normal(what)."""
    glimpse = stn_forward(img, to_coords(where), cfg.G).reshape(img.shape[0], cfg.g)
    if cfg.masked_glimpse and mask_inpt is not None:
        glimpse = glimpse * glimpse_mask(p, mask_inpt)
    hdn = mlp_hidden(p, _DC + 'encoder_1/mlp', glimpse, 2)
    params = linear(p, _DC + 'air_encoder/gaussian_from_param_vec/linear', hdn)
    loc, scale = params[:, :cfg.nw], params[:, cfg.nw:]
    return loc, F.softplus(scale) + cfg.min_std, glimpse      # modules.py:70-74 (scale_offset 0)


# --------------------------------------------------------------------------------------------
# propagation prior (propagate.py:48-158)
# --------------------------------------------------------------------------------------------
def propagate_prior(p, cfg, z_tm1, prior_state):
    what_tm1, where_tm1, pres_tm1, pres_logit_tm1 = z_tm1
    Bp, n = what_tm1.shape[:2]
    x = torch.cat((what_tm1, where_tm1), -1).reshape(Bp * n, -1)              # propagate.py:77
    h = gru(p, 'propagation/gru_1', x, prior_state.reshape(Bp * n, -1))       # BatchApply :78-80
    stats = linear(p, 'propagation/propagate_prior/linear', h).reshape(Bp, n, -1)
    logit, stats = stats[..., :1], stats[..., 1:]                             # :84
    logit = logit + cfg.prop_prior_step_bias
    logit = pres_tm1 * logit + (pres_tm1 - 1.) * 88.                          # :86
    m = 4 + cfg.nw
    locs, scales = stats[..., :m], stats[..., m:]
    where_loc, what_loc = locs[..., :4], locs[..., 4:]
    where_scale = F.softplus(scales[..., :4]) + 1e-2                          # :91
    what_scale = F.softplus(scales[..., 4:]) + 1e-2
    if cfg.prior_type == 'rw':                                                # :123-139
        where_loc, what_loc = where_tm1, what_tm1
        logit = pres_logit_tm1 + .1 * logit
    elif cfg.prior_type == 'guided':                                          # :142-158
        where_loc, what_loc = where_tm1 + .1 * where_loc, what_tm1 + .1 * what_loc
        logit = pres_logit_tm1 + .1 * logit
    elif cfg.prior_type != 'rnn':
        raise ValueError('Invalid prior type: "{}". Choose from {}.'.format(cfg.prior_type, ['rnn', 'rw', 'guided']))
    return (where_loc, where_scale, what_loc, what_scale, logit), h.reshape(Bp, n, -1)


# --------------------------------------------------------------------------------------------
# propagation core (core.py:230-359) and SSM (propagate.py:161-184)
# --------------------------------------------------------------------------------------------
_PC = 'propagation/propagation_core/'


def fill_triangular(x):
    """tfd.fill_triangular (lower) for a 10-vector -> 4x4, TF's element order (Appendix B)."""
    n = 4
    full = torch.cat((x[n:], torch.flip(x, [0]))).reshape(n, n)
    return torch.tril(full)


def affine_diag_normal_tril(p, scale):
    """modules.py:535-545: L = fill_triangular(c) * scale[..., None] + diag(scale)."""
    L0 = fill_triangular(p[_PC + 'affine_diag_normal/cholesky_scale'])
    return L0[None] * scale[..., None] + torch.diag_embed(scale)


def mvn_tril_log_prob(x, loc, L):
    """tfd.MultivariateNormalTriL.log_prob, d = 4."""
    d = (x - loc)
    ys = []
    for i in range(4):                                 # forward substitution L y = d
        acc = d[..., i]
        for j in range(i):
            acc = acc - L[..., i, j] * ys[j]
        ys.append(acc / L[..., i, i])
    y = torch.stack(ys, -1)
    logdet = torch.log(torch.abs(torch.diagonal(L, dim1=-2, dim2=-1))).sum(-1)
    return -0.5 * (y ** 2).sum(-1) - logdet - 2. * LOG_2PI


def steps_predictor(p, scope, prev_presence, feats):
    """modules.py:506-524: logit = MLP(128 -> 1)(feats); gated by previous presence."""
    logit = mlp(p, scope + 'steps_predictor/mlp', feats, 2)
    return prev_presence * logit + (prev_presence - 1.) * 88.


def propagation_core_step(p, cfg, img, z_tm1_k, temporal_state, state, eps_where, eps_what, u_pres):
    """core.py:280-359 for one object slot.  state = (what_km1, where_km1, pres_km1, h)."""
    what_tm1, where_tm1, pres_tm1, _ = z_tm1_k
    what_km1, where_km1, pres_km1, h = state
    where_bias = mlp(p, _PC + 'rnn_inpt/mlp', temporal_state, 2) * .1              # core.py:291
    loc1, _, _ = air_encoder(p, cfg, img, where_tm1 + where_bias, temporal_state)  # :292-293
    rnn_inpt = torch.cat((loc1, what_km1, where_km1, pres_km1,
                          what_tm1, where_tm1, pres_tm1, temporal_state), -1)      # :295-301
    h = vanilla_rnn(p, 'propagation/vanilla_rnn', rnn_inpt, h)                      # :302
    # where (core.py:321-333)
    t_in = torch.cat((h, where_tm1, temporal_state), -1)
    prm = mlp(p, _PC + 'stochastic_transform_param/mlp', t_in, 3)
    loc_w = prm[:, :4]
    scale_w = prm[:, 4:] + p[_PC + 'stochastic_transform_param/scale_offset']       # modules.py:96-97
    where_loc = where_tm1 + cfg.where_update_scale * loc_w
    where_scale = F.softplus(scale_w - 1.) + 1e-2
    L = affine_diag_normal_tril(p, where_scale)
    where = where_loc + (L @ eps_where[..., None])[..., 0]                          # MVN-TriL sample
    # what (core.py:335-359)
    loc2, scale2, _ = air_encoder(p, cfg, img, where, temporal_state)
    g_in = torch.cat((h, where, loc2, scale2), -1)
    new_temporal = gru(p, 'propagation/gru', g_in, temporal_state)
    tprm = linear(p, _PC + 'what/gaussian_from_param_vec/linear', new_temporal)
    loc_t, scale_t = tprm[:, :cfg.nw], F.softplus(tprm[:, cfg.nw:]) + cfg.min_std
    gates = torch.sigmoid(linear(p, _PC + 'what/linear', new_temporal)) * .9999
    fgate, igate, tgate = gates[:, :cfg.nw], gates[:, cfg.nw:2 * cfg.nw], gates[:, 2 * cfg.nw:]
    what_loc = fgate * what_tm1 + (1. - igate) * loc2 + (1. - tgate) * loc_t
    what_scale = (1. - igate) * scale2 + (1. - tgate) * scale_t
    what = what_loc + what_scale * eps_what
    # presence (core.py:141-144,311-313)
    logit = steps_predictor(p, _PC, pres_tm1, torch.cat((h, temporal_state, what), -1))
    prob = torch.sigmoid(logit)
    pres = (u_pres[:, None] < prob).to(prob.dtype) * pres_tm1
    out = dict(what=what, what_loc=what_loc, what_scale=what_scale, where=where, where_loc=where_loc,
               where_scale=where_scale, presence_prob=prob, presence=pres, presence_logit=logit,
               temporal_state=new_temporal)
    return out, (what, where, pres, h)


def _stack(dicts):
    return {k: torch.stack([d[k] for d in dicts], 1) for k in dicts[0]}


def sequential_ssm(p, cfg, img, z_tm1, temporal_state, eps_where, eps_what, u_pres):
    """propagate.py:168-184: static_rnn of PropagationCore over the n slots."""
    Bp = img.shape[0]
    state = (torch.zeros(Bp, cfg.nw), torch.zeros(Bp, 4), torch.zeros(Bp, 1),         # core.py:132-139,238
             p['propagation/sequential_ssm/propagation/vanilla_rnn_initial_state_0/w'].expand(Bp, -1))
    outs = []
    for k in range(cfg.n):
        zk = tuple(z[:, k] for z in z_tm1)
        o, state = propagation_core_step(p, cfg, img, zk, temporal_state[:, k], state,
                                         eps_where[:, k], eps_what[:, k], u_pres[:, k])
        outs.append(o)
    ho = _stack(outs)
    return ho, ho['presence'][..., 0].sum(-1)


def propagate(p, cfg, img, z_tm1, temporal_state, prior_state, eps_where, eps_what, u_pres, do_generate=False,
              prior_noise=None):
    """sqair_modules.py:250-329.  `prior_noise` = (eps_where, eps_what, u_pres) of the prior draws, each [B', n, ...]
    (only with cfg.sample_from_prior)."""
    pres_tm1 = z_tm1[2][..., 0]
    prior_stats, prior_state = propagate_prior(p, cfg, z_tm1, prior_state)
    ho, num_steps = sequential_ssm(p, cfg, img, z_tm1, temporal_state, eps_where, eps_what, u_pres)
    pres = ho['presence'][..., 0]                                   # (:286: taken BEFORE any replacement and used for all masks)
    pw_loc, pw_scale, pa_loc, pa_scale, p_logit = prior_stats
    s_what, s_where, s_pres = ho['what'], ho['where'], pres         # :292 samples at which q is evaluated
    if cfg.sample_from_prior:                                       # :294-302
        ew, ea, up = prior_noise
        s_what = pa_loc + pa_scale * ea                             # [p.sample() for p in priors] (propagate.py:113-120)
        s_where = pw_loc + pw_scale * ew
        s_pres = (up < torch.sigmoid(p_logit[..., 0])).to(pres.dtype)
        if do_generate:
            ho = dict(ho)
            ho['what'], ho['where'], ho['presence'] = s_what, s_where, s_pres[..., None]
    # posteriors (sqair_modules.py:49-60,290,304)
    q_what = normal_log_prob(s_what, ho['what_loc'], ho['what_scale']).sum(-1)
    q_where = mvn_tril_log_prob(s_where, ho['where_loc'], affine_diag_normal_tril(p, ho['where_scale']))
    q_pres = bernoulli_log_prob(s_pres, ho['presence_logit'][..., 0])
    # priors (propagate.py:113-120; sqair_modules.py:306-307)
    p_what = normal_log_prob(ho['what'], pa_loc, pa_scale).sum(-1)
    p_where = normal_log_prob(ho['where'], pw_loc, pw_scale).sum(-1)
    p_pres = bernoulli_log_prob(pres, p_logit[..., 0])
    prop_prob = torch.exp(q_pres) * pres_tm1                                         # :308
    m = pres_tm1 * pres                                                              # :315
    q_what, q_where, p_what, p_where = q_what * m, q_where * m, p_what * m, p_where * m
    q_pres = (q_pres * pres_tm1).sum(-1)                                             # :317
    p_pres = (p_pres * pres_tm1).sum(-1)
    o = dict(ho)
    o.update(prior_stats=prior_stats, prior_state=prior_state, num_steps=num_steps,
             what_log_prob=q_what, where_log_prob=q_where, prop_log_prob=q_pres,
             what_prior_log_prob=p_what, where_prior_log_prob=p_where, prop_prior_log_prob=p_pres,
             prop_prob=prop_prob,
             q_z_given_x=(q_what + q_where).sum(-1) + q_pres,                        # :62-63,319-320
             p_z=(p_what + p_where).sum(-1) + p_pres)
    return o


# --------------------------------------------------------------------------------------------
# discovery core (core.py:147-227) and Discover (sqair_modules.py:66-229)
# --------------------------------------------------------------------------------------------
def image_encoder(p, img):
    return mlp_hidden(p, _DC + 'encoder/mlp', img.reshape(img.shape[0], -1), 2)       # core.py:165


def discovery_core_step(p, cfg, img, conditioning, state, eps_where, eps_what, u_pres):
    """core.py:192-227.  The `is_allowed` input is unpacked and never used (core.py:192)."""
    what_km1, where_km1, pres_km1, h = state
    rnn_inpt = torch.cat((image_encoder(p, img), conditioning, what_km1, where_km1, pres_km1), -1)
    h = vanilla_rnn(p, 'discovery/vanilla_rnn', rnn_inpt, h)
    prm = mlp(p, _DC + 'stochastic_transform_param/mlp', h, 3)
    where_loc = prm[:, :4]
    where_scale = F.softplus(prm[:, 4:] + p[_DC + 'stochastic_transform_param/scale_offset']) + 1e-2
    where = where_loc + where_scale * eps_where
    what_loc, what_scale, _ = air_encoder(p, cfg, img, where)                          # no mask, core.py:217
    what = what_loc + what_scale * eps_what
    logit = steps_predictor(p, _DC, pres_km1, torch.cat((h, what), -1))
    prob = torch.sigmoid(logit)
    pres = (u_pres[:, None] < prob).to(prob.dtype) * pres_km1
    out = dict(what=what, what_loc=what_loc, what_scale=what_scale, where=where, where_loc=where_loc,
               where_scale=where_scale, presence_prob=prob, presence=pres, presence_logit=logit)
    return out, (what, where, pres, h)


def bernoulli_to_modified_geometric(presence_prob):
    """prior.py:61-67, float64 island."""
    pp = presence_prob.double()
    inv = 1. - pp
    prob = torch.cumprod(pp, -1)                                                       # prior.py:34-58
    mod = torch.cat((inv[..., :1], inv[..., 1:] * prob[..., :-1], prob[..., -1:]), -1)
    mod = mod / mod.sum(-1, keepdim=True)
    return mod.to(presence_prob.dtype)          # float32 in the reference (prior.py:67)


def recurrent_normal_log_prob(p, cfg, samples, conditioning):
    """modules.py:548-630 with override_samples: the conditioned 128-d state is never updated
    (modules.py:582-593); the 4-unit VanillaRNN sees (previous sample, state)."""
    Bp, n, _ = samples.shape
    state = p[_RN + 'discovery/discover/recurrent_normal_impl/vanilla_rnn_initial_state_0/w'].expand(Bp, -1)
    state = F.elu(linear(p, _RN + 'linear_1', torch.cat((state, conditioning), -1)))
    prev = p[_RN + 'init_sample'].expand(Bp, -1)
    lps = []
    for i in range(n):
        out = vanilla_rnn(p, _RN + 'vanilla_rnn', prev, state)
        stats = linear(p, _RN + 'linear', out)
        loc, scale = stats[:, :4], F.softplus(stats[:, 4:]) + 1e-2
        lps.append(normal_log_prob(samples[:, i], loc, scale))
        prev = samples[:, i]
    return torch.stack(lps, 1)


def recurrent_normal_sample(p, cfg, conditioning, eps):
    """modules.py:621-630 `RecurrentNormal.sample`: like the log-prob loop, but each step feeds its own draw."""
    Bp, n, _ = eps.shape
    state = p[_RN + 'discovery/discover/recurrent_normal_impl/vanilla_rnn_initial_state_0/w'].expand(Bp, -1)
    state = F.elu(linear(p, _RN + 'linear_1', torch.cat((state, conditioning), -1)))
    prev = p[_RN + 'init_sample'].expand(Bp, -1)
    out = []
    for i in range(n):
        stats = linear(p, _RN + 'linear', vanilla_rnn(p, _RN + 'vanilla_rnn', prev, state))
        prev = stats[:, :4] + (F.softplus(stats[:, 4:]) + 1e-2) * eps[:, i]
        out.append(prev)
    return torch.stack(out, 1)


def discover(p, cfg, img, conditioning, time_step, prior_conditioning, eps_where, eps_what, u_pres, do_generate=False,
             prior_noise=None):
    """sqair_modules.py:94-229."""
    Bp = img.shape[0]
    n = cfg.n
    state = (torch.zeros(Bp, cfg.nw), torch.zeros(Bp, 4), torch.ones(Bp, 1),             # core.py:153
             p['discovery/discover/discovery/vanilla_rnn_initial_state_0/w'].expand(Bp, -1))
    outs = []
    for k in range(n):
        o, state = discovery_core_step(p, cfg, img, conditioning, state,
                                       eps_where[:, k], eps_what[:, k], u_pres[:, k])
        outs.append(o)
    ho = _stack(outs)
    num_steps = ho['presence'][..., 0].sum(-1)                                        # :145 (before any replacement)
    if cfg.sample_from_prior and do_generate:                                         # :157-170
        ew, ea, _ = prior_noise
        ho = dict(ho)
        ho['what'] = ea                                                               # N(0, 1) prior (:79)
        if cfg.rec_where_prior:
            ho['where'] = recurrent_normal_sample(p, cfg, torch.cat((conditioning, prior_conditioning), -1), ew)
        else:
            ho['where'] = torch.tensor(cfg.where_mean, dtype=ew.dtype) + torch.tensor(cfg.where_std, dtype=ew.dtype) * ew
        ho['presence'] = torch.zeros_like(ho['presence'])                             # pres_sample * 0. (:164): nothing is discovered
    pres = ho['presence'][..., 0]
    # posteriors (:177-179)
    q_what = normal_log_prob(ho['what'], ho['what_loc'], ho['what_scale']).sum(-1) * pres
    q_where = normal_log_prob(ho['where'], ho['where_loc'], ho['where_scale']).sum(-1) * pres
    joint = bernoulli_to_modified_geometric(ho['presence_prob'][..., 0])               # [B', n+1]
    idx = num_steps.to(torch.int64)[:, None]
    q_num = torch.log(clip_preserve(joint.gather(1, idx)[:, 0], 1e-16, 1.))            # prior.py:95-102
    # priors (:199-226)
    p_what = normal_log_prob(ho['what'], torch.zeros(()), torch.ones(())).sum(-1) * pres
    if cfg.rec_where_prior:
        where_cond = torch.cat((conditioning, prior_conditioning), -1)                 # :154
        p_where = recurrent_normal_log_prob(p, cfg, ho['where'], where_cond).sum(-1) * pres
    else:
        p_where = normal_log_prob(ho['where'], torch.tensor(cfg.where_mean, dtype=pres.dtype),
                                  torch.tensor(cfg.where_std, dtype=pres.dtype)).sum(-1) * pres
    if cfg.disc_prior_type == 'cat':
        logits = p['model/sequential_air/while/sqair_timestep/discover/step_prior_bias'] \
            + (0. if time_step == 0 else 1.) * \
            p['model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias']
        logits = logits[None] + mlp(p, 'discovery/discover/mlp', prior_conditioning, 2)   # :218
        logits = F.elu(logits)
        p_num = torch.log_softmax(logits, -1).gather(1, idx)[:, 0]
    elif cfg.disc_prior_type == 'geom':
        pr = 1. - cfg.step_success_prob                                                # tfd.Geometric(probs)
        p_num = num_steps * math.log1p(-pr) + math.log(pr)
    else:
        raise ValueError('Invalid prior type: {}'.format(cfg.disc_prior_type))
    o = dict(ho)
    o.update(num_steps=num_steps, what_log_prob=q_what, where_log_prob=q_where, num_step_log_prob=q_num,
             what_prior_log_prob=p_what, where_prior_log_prob=p_where, num_step_prior_log_prob=p_num,
             num_steps_prob=joint,
             q_z_given_x=(q_what + q_where).sum(-1) + q_num,
             p_z=(p_what + p_where).sum(-1) + p_num)
    return o


# --------------------------------------------------------------------------------------------
# slot bookkeeping (index.py:132-221; sqair_modules.py:368-385,514-582)
# --------------------------------------------------------------------------------------------
def encode_latents(p, what, where, presence):
    """sqair_modules.py:368-385 (relation_embedding=False)."""
    x = torch.cat((what, where), -1)
    f = mlp_hidden(p, 'sequence/sequential_air/sqair_timestep/mlp', x, 2) * presence
    return f.sum(-2)


def select_present(x, presence):
    """index.py:132-165: per-row stable partition, present entries first.  x [B',K,d]."""
    order = torch.argsort(1 - presence.to(torch.int64), dim=1, stable=True)
    return x.gather(1, order[..., None].expand_as(x))


def compute_object_ids(last_used_id, prev_ids, prop_pres, disc_pres):
    """index.py:198-221.  last_used_id [B',1]; prev_ids/pres [B',n,1]."""
    prop_ids = prev_ids * prop_pres - (1 - prop_pres)
    inc = torch.cumsum(disc_pres, 1)
    disc_ids = inc + last_used_id[:, None]
    last_used_id = last_used_id + inc[:, -1]
    disc_ids = disc_ids * disc_pres - (1 - disc_pres)
    return last_used_id, torch.cat((prop_ids, disc_ids), 1)


HEADS = 'what what_loc what_scale where where_loc where_scale presence_prob presence presence_logit'.split()


def choose_latents(p, cfg, prop, disc, last_used_id, prev_ids):
    """sqair_modules.py:514-582."""
    Bp, n = prev_ids.shape[:2]
    t0 = p['sequence/sequential_air/propagation/gru_initial_state_0/w'].expand(Bp, n, -1)
    p0 = p['sequence/sequential_air/propagation/gru_1_initial_state_0/w'].expand(Bp, n, -1)
    temporal = torch.cat((prop['temporal_state'], t0), 1)
    prior = torch.cat((prop['prior_state'], p0), 1)
    merged = [torch.cat((prop[k], disc[k]), 1) for k in HEADS]
    last_used_id, new_ids = compute_object_ids(last_used_id, prev_ids, prop['presence'], disc['presence'])
    pres = merged[HEADS.index('presence')][..., 0]
    parts = [select_present(v, pres)[:, :n] for v in merged + [new_ids, prior, temporal]]
    ho = dict(zip(HEADS, parts[:len(HEADS)]))
    ids, prior, temporal = parts[len(HEADS):]
    return ho, ids, prior, temporal, last_used_id


# --------------------------------------------------------------------------------------------
# decoder (modules.py:367-467) and per-frame log weights (seq.py:271-276)
# --------------------------------------------------------------------------------------------
def air_decoder(p, cfg, what, where, presence):
    Bp, n = what.shape[:2]
    g = mlp(p, 'decoder/air_decoder/decoder/mlp', what.reshape(Bp * n, -1), 3)
    g = g * p['decoder/air_decoder/decoder/output_scale']                                # modules.py:147
    coords = to_coords(where).reshape(Bp * n, 4)
    pres = presence.reshape(Bp, n, 1, 1)
    canvas = (stn_inverse(g.reshape(Bp * n, cfg.G, cfg.G), coords, cfg.H, cfg.W)
              .reshape(Bp, n, cfg.H, cfg.W) * pres).sum(1)                              # :435-445
    ones = torch.ones(Bp * n, cfg.G, cfg.G)
    nz = (stn_inverse(ones, coords, cfg.H, cfg.W).reshape(Bp, n, cfg.H, cfg.W) * pres).sum(1)
    mask = torch.sigmoid(-10. + nz * 20.)                                              # :462
    canvas = canvas + p['decoder/air_decoder/Variable'][None, :, :, 0] * mask          # :465
    std = float(np.float32(np.sqrt(np.float32(cfg.output_std))) ** 2)                   # :419-422
    bg = cfg.output_std if cfg.bg_std is None else cfg.bg_std                           # :406-407
    bg = float(np.float32(np.sqrt(np.float32(bg))) ** 2)
    out_std = mask * std + (1. - mask) * bg                                             # :453
    return canvas, out_std, g.reshape(Bp, n, cfg.G, cfg.G)


def sqair_timestep(p, cfg, img, z_tm1, temporal_state, prior_state, last_used_id, prev_ids, t, noise_t, prior_noise_t=None):
    """sqair_modules.py:446-512."""
    n = cfg.n
    ew, ea, up = noise_t
    dg = cfg.sample_from_prior and cfg.generate_after > 0 and t > cfg.generate_after   # seq.py:198-203
    pn_prop = pn_disc = None
    if cfg.sample_from_prior:
        pn_prop = tuple(x[:, :n] for x in prior_noise_t)
        pn_disc = tuple(x[:, n:] for x in prior_noise_t)
    prop = propagate(p, cfg, img, z_tm1, temporal_state, prior_state, ew[:, :n], ea[:, :n], up[:, :n], dg, pn_prop)
    conditioning = encode_latents(p, prop['what'], prop['where'], prop['presence'])     # :501
    prior_logit = prop['prior_stats'][-1][..., 0]
    expected = ((torch.sigmoid(prior_logit) - .5) / n).sum(-1, keepdim=True)             # :505-507
    disc = discover(p, cfg, img, conditioning, t, expected, ew[:, n:], ea[:, n:], up[:, n:], dg, pn_disc)
    ho, ids, prior_state, temporal_state, last_used_id = choose_latents(p, cfg, prop, disc, last_used_id, prev_ids)
    return prop, disc, ho, ids, prior_state, temporal_state, last_used_id


OUTPUT_NAMES = (HEADS + 'obj_id step_log_prob canvas glimpse '
                'disc_what_log_prob disc_where_log_prob disc_what_prior_log_prob disc_where_prior_log_prob '
                'disc_log_prob disc_prior_log_prob disc_prob '
                'prop_what_log_prob prop_where_log_prob prop_what_prior_log_prob prop_where_prior_log_prob '
                'prop_log_prob prop_prior_log_prob prop_prob discrete_log_prob '
                'num_prop_steps_per_sample num_disc_steps_per_sample num_steps_per_sample prop_pres disc_pres '
                'data_ll_per_sample kl_per_sample log_q_z_given_x_per_sample log_p_z_per_sample '
                'log_weights_per_timestep'.split())
assert len(OUTPUT_NAMES) == 38


def sequential_air(p, cfg, obs, noise):
    """seq.py:69-279.  obs [T,B',H,W] (already tiled for IWAE) -> dict of the 38 [T,B',...] outputs."""
    T, Bp = obs.shape[:2]
    n = cfg.n
    z = (torch.zeros(Bp, n, cfg.nw), torch.zeros(Bp, n, 4), torch.zeros(Bp, n, 1), torch.zeros(Bp, n, 1))
    temporal = p['sequence/sequential_air/propagation/gru_initial_state_0/w'].expand(Bp, n, -1)
    prior = p['sequence/sequential_air/propagation/gru_1_initial_state_0/w'].expand(Bp, n, -1)
    prev_ids = -torch.ones(Bp, n, 1)
    last_id = -torch.ones(Bp, 1)
    tas = {k: [] for k in OUTPUT_NAMES}
    for t in range(T):
        img = obs[t]
        noise_t = (noise['eps_where'][t], noise['eps_what'][t], noise['u_pres'][t])
        pn_t = None
        if cfg.sample_from_prior:          # the prior draws (samples of `p.sample()`): a second set of the same shapes
            pn_t = (noise['eps_where_prior'][t], noise['eps_what_prior'][t], noise['u_pres_prior'][t])
        prop, disc, ho, ids, prior, temporal, last_id = sqair_timestep(
            p, cfg, img, z, temporal, prior, last_id, prev_ids, t, noise_t, pn_t)
        z = (ho['what'], ho['where'], ho['presence'], ho['presence_logit'])
        prev_ids = ids
        canvas, std, glimpse = air_decoder(p, cfg, ho['what'], ho['where'], ho['presence'])
        data_ll = normal_log_prob(img, canvas, std).sum((1, 2))                          # seq.py:272-273
        q = disc['q_z_given_x'] + prop['q_z_given_x']
        pz = disc['p_z'] + prop['p_z']
        kl = q - pz
        vals = dict(ho)
        vals.update(
            obj_id=ids, step_log_prob=prop['prop_log_prob'] + disc['num_step_log_prob'],
            canvas=canvas, glimpse=glimpse,
            disc_what_log_prob=disc['what_log_prob'], disc_where_log_prob=disc['where_log_prob'],
            disc_what_prior_log_prob=disc['what_prior_log_prob'],
            disc_where_prior_log_prob=disc['where_prior_log_prob'],
            disc_log_prob=disc['num_step_log_prob'], disc_prior_log_prob=disc['num_step_prior_log_prob'],
            disc_prob=disc['num_steps_prob'],
            prop_what_log_prob=prop['what_log_prob'], prop_where_log_prob=prop['where_log_prob'],
            prop_what_prior_log_prob=prop['what_prior_log_prob'],
            prop_where_prior_log_prob=prop['where_prior_log_prob'],
            prop_log_prob=prop['prop_log_prob'], prop_prior_log_prob=prop['prop_prior_log_prob'],
            prop_prob=prop['prop_prob'],
            discrete_log_prob=prop['prop_log_prob'] + disc['num_step_log_prob'],
            num_prop_steps_per_sample=prop['num_steps'], num_disc_steps_per_sample=disc['num_steps'],
            num_steps_per_sample=ho['presence'][..., 0].sum(-1),
            prop_pres=prop['presence'], disc_pres=disc['presence'],
            data_ll_per_sample=data_ll, kl_per_sample=kl, log_q_z_given_x_per_sample=q,
            log_p_z_per_sample=pz, log_weights_per_timestep=data_ll - kl)
        for k in OUTPUT_NAMES:
            v = vals[k]
            if v.dim() > 1 and v.shape[-1] == 1:                                          # seq.py:254-255
                v = v[..., 0]
            tas[k].append(v)
    return {k: torch.stack(v, 0) for k, v in tas.items()}


# --------------------------------------------------------------------------------------------
# objective (model.py:79-168; targets.py:38-75; ops.py:52-59)
# --------------------------------------------------------------------------------------------
def tile_input_for_iwae(x, K):
    """index.py:106-129 with_time=True: [T,B,...] -> [T,B*K,...], row = b*K + k."""
    return x.repeat_interleave(K, dim=1)


def iwae(log_weights):
    return torch.logsumexp(log_weights, -1) - math.log(float(log_weights.shape[-1]))


def vimco_control_variate(lw):
    K = lw.shape[-1]
    s = lw.sum(-1, keepdim=True)
    abo = (s - lw) / (K - 1.)
    base = lw[..., None] + torch.diag_embed(abo - lw)
    return torch.logsumexp(base, -2) - math.log(float(K))


def vimco(log_weights, log_probs, elbo_iwae=None):
    cv = vimco_control_variate(log_weights)
    signal = (log_weights - cv).detach()
    log_probs = log_probs.reshape(log_weights.shape)
    if elbo_iwae is None:
        elbo_iwae = iwae(log_weights)
    return (-elbo_iwae[..., None] - signal * log_probs).mean()


def ess(weights):
    return weights.sum(-1) ** 2 / (weights ** 2).sum(-1)


def model_forward(p, cfg, obs, noise):
    """model.py:43-168.  obs [T,B,H,W].  Returns (outputs dict, objective dict)."""
    tiled = tile_input_for_iwae(obs, cfg.K)
    out = sequential_air(p, cfg, tiled, noise)
    T = obs.shape[0]
    lw = out['log_weights_per_timestep'].sum(0).reshape(cfg.B, cfg.K)
    obj = dict(log_weights=lw, elbo_vae=lw.mean())
    obj['elbo_iwae_per_example'] = iwae(lw)
    obj['elbo_iwae'] = obj['elbo_iwae_per_example'].mean()
    obj['normalised_elbo_vae'] = obj['elbo_vae'] / T
    obj['normalised_elbo_iwae'] = obj['elbo_iwae'] / T
    iw = torch.softmax(lw, -1)
    obj['importance_weights'] = iw
    obj['ess'] = ess(iw).mean()
    if cfg.K > 1:
        lp = out['discrete_log_prob'].sum(0)
        obj['vimco_target'] = vimco(lw, lp, obj['elbo_iwae_per_example']) / T           # model.py:150-158
    obj['iwae_target'] = -obj['elbo_iwae'] / T

    def iwm(x):                                                                         # model.py:202-205
        x = x.reshape(-1, cfg.B, cfg.K).mean(0)
        return (iw * x * cfg.K).mean()

    for name, key in (('data_ll', 'data_ll_per_sample'), ('log_p_z', 'log_p_z_per_sample'),
                      ('log_q_z_given_x', 'log_q_z_given_x_per_sample'), ('kl', 'kl_per_sample'),
                      ('num_steps', 'num_steps_per_sample'), ('num_disc_steps', 'num_disc_steps_per_sample'),
                      ('num_prop_steps', 'num_prop_steps_per_sample')):
        obj[name] = iwm(out[key])
    mse = ((tiled - out['canvas']) ** 2).mean((0, 2, 3))
    obj['mse'] = iwm(mse)
    obj['raw_mse'] = mse.mean()
    return out, obj


# --------------------------------------------------------------------------------------------
# gradients (model.py:150-168: opt.compute_gradients(target)) -- torch autograd through the restatement
# --------------------------------------------------------------------------------------------
def model_gradients(p, cfg, obs, noise, target='auto'):
    """d target / d every variable, target = VIMCO / T (model.py:152-158; NaN at K = 1, targets.py:55) or the
    `-elbo_iwae / T` branch (model.py:156).  'auto' = VIMCO when K > 1.  Returns (grads by TF name, objective dict)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out, obj = model_forward(leaves, cfg, obs, noise)
    use_vimco = cfg.K > 1 if target == 'auto' else target == 'vimco'
    tgt = obj['vimco_target'] if use_vimco else obj['iwae_target']
    gs = torch.autograd.grad(tgt, list(leaves.values()), allow_unused=True)
    grads = OrderedDict((k, (torch.zeros_like(v) if g is None else g)) for (k, v), g in zip(leaves.items(), gs))
    missing = [k for (k, _), g in zip(leaves.items(), gs) if g is None]
    return grads, {k: v.detach() for k, v in obj.items()}, missing
