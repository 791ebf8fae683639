"""Thin torch-facing wrappers over the C ABI (include/sqair_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation below is a
call into libsqair_b200.so on the caller's current stream.  There is no fallback: tensors must be
CUDA tensors and the library must be built.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from ._capi import SqairCfg, SqairOutputs, OUTPUT_NAMES, check, make_cfg, output_shapes, query_sizes, param_layout


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError('sqair_b200 expects contiguous float32 CUDA tensors (no CPU fallback exists)')


def pack_params(cfg: SqairCfg, flat: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Canonical flat parameter vector (TF variable order) -> kernel-side buffer."""
    _need_cuda(flat)
    sizes = query_sizes(cfg)
    if flat.numel() != sizes.param_count:
        raise ValueError('expected %d parameters, got %d' % (sizes.param_count, flat.numel()))
    packed = out if out is not None else torch.empty(sizes.packed_floats, dtype=torch.float32, device=flat.device)
    if packed.numel() != sizes.packed_floats:
        raise ValueError('packed buffer has the wrong size')
    check(_capi.lib().sqair_pack_params(C.byref(cfg), _ptr(flat), _ptr(packed), _stream()))
    return packed


def alloc_noise(cfg: SqairCfg, device):
    rows, n2 = cfg.B * cfg.K, 2 * cfg.n
    return dict(eps_where=torch.empty(cfg.T, rows, n2, 4, dtype=torch.float32, device=device),
                eps_what=torch.empty(cfg.T, rows, n2, cfg.n_what, dtype=torch.float32, device=device),
                u_pres=torch.empty(cfg.T, rows, n2, dtype=torch.float32, device=device))


def fill_noise(cfg: SqairCfg, seed: int, row_offset: int = 0, noise=None, device=None):
    """Counter-based draws for every noise site; `row_offset` = global index of this shard's row 0."""
    if noise is None:
        noise = alloc_noise(cfg, device or torch.device('cuda', torch.cuda.current_device()))
    _need_cuda(noise['eps_where'], noise['eps_what'], noise['u_pres'])
    check(_capi.lib().sqair_fill_noise(C.byref(cfg), C.c_uint64(seed & (2 ** 64 - 1)), int(row_offset),
                                       _ptr(noise['eps_where']), _ptr(noise['eps_what']), _ptr(noise['u_pres']),
                                       _stream()))
    return noise


def alloc_outputs(cfg: SqairCfg, device, names=None):
    shapes = output_shapes(cfg)
    return {k: torch.empty(shapes[k], dtype=torch.float32, device=device) for k in (names or OUTPUT_NAMES)}


def forward(cfg: SqairCfg, packed: torch.Tensor, obs: torch.Tensor, noise: dict, outputs: dict = None, names=None,
            stash: torch.Tensor = None, prior_noise: dict = None, generate_after: int = -1):
    """SequentialAIR over a [T,B,H,W] batch (seq.py:69-84); returns {name: [T, B*K, ...] tensor}.  With `stash` (a
    float32 buffer of `query_train_sizes(cfg).stash_floats`) the kernel also records what `backward` needs."""
    _need_cuda(packed, obs, noise['eps_where'], noise['eps_what'], noise['u_pres'])
    if tuple(obs.shape) != (cfg.T, cfg.B, cfg.H, cfg.W):
        raise ValueError('obs must be [T,B,H,W] = %s, got %s' % ((cfg.T, cfg.B, cfg.H, cfg.W), tuple(obs.shape)))
    rows, n2 = cfg.B * cfg.K, 2 * cfg.n
    if tuple(noise['eps_where'].shape) != (cfg.T, rows, n2, 4) or \
            tuple(noise['eps_what'].shape) != (cfg.T, rows, n2, cfg.n_what) or \
            tuple(noise['u_pres'].shape) != (cfg.T, rows, n2):
        raise ValueError('noise tensors have the wrong shape for this configuration')
    if outputs is None:
        outputs = alloc_outputs(cfg, obs.device, names)
    so = SqairOutputs()
    for k, v in outputs.items():
        _need_cuda(v)
        setattr(so, k, v.data_ptr())
    if prior_noise is not None:            # generation: sample_from_prior / generate_after (seq.py:198-203)
        if stash is not None:
            raise ValueError('generation is an inference mode: no stash')
        _need_cuda(prior_noise['eps_where'], prior_noise['eps_what'], prior_noise['u_pres'])
        for k in ('eps_where', 'eps_what', 'u_pres'):
            if tuple(prior_noise[k].shape) != tuple(noise[k].shape):
                raise ValueError('prior noise tensors must have the shapes of the posterior noise tensors')
        check(_capi.lib().sqair_forward_generate(C.byref(cfg), _ptr(packed), _ptr(obs), _ptr(noise['eps_where']),
                                                 _ptr(noise['eps_what']), _ptr(noise['u_pres']), _ptr(prior_noise['eps_where']),
                                                 _ptr(prior_noise['eps_what']), _ptr(prior_noise['u_pres']), int(generate_after),
                                                 C.byref(so), _stream()))
    elif stash is None:
        check(_capi.lib().sqair_forward(C.byref(cfg), _ptr(packed), _ptr(obs), _ptr(noise['eps_where']),
                                        _ptr(noise['eps_what']), _ptr(noise['u_pres']), C.byref(so), _stream()))
    else:
        _need_cuda(stash)
        if stash.numel() < _capi.query_train_sizes(cfg).stash_floats:
            raise ValueError('stash buffer too small for this configuration')
        check(_capi.lib().sqair_forward_train(C.byref(cfg), _ptr(packed), _ptr(obs), _ptr(noise['eps_where']),
                                              _ptr(noise['eps_what']), _ptr(noise['u_pres']), C.byref(so), _ptr(stash), _stream()))
    return outputs


def pack_backward(cfg: SqairCfg, flat: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Canonical flat parameters -> the parameter copy the backward GEMMs read (once per parameter update)."""
    _need_cuda(flat)
    n = _capi.query_train_sizes(cfg).backward_param_floats
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=flat.device)
    check(_capi.lib().sqair_pack_backward(C.byref(cfg), _ptr(flat), _ptr(out), _stream()))
    return out


def backward(cfg: SqairCfg, flat: torch.Tensor, bw_params: torch.Tensor, obs: torch.Tensor, noise: dict, stash: torch.Tensor,
             d_log_w: torch.Tensor, d_disc_lp: torch.Tensor = None, workspace: torch.Tensor = None, d_params: torch.Tensor = None):
    """`opt.compute_gradients(target)` (model.py:160): gradient of the target w.r.t. every variable, in the canonical
    flat layout.  d_log_w / d_disc_lp: [B, K] from `objective_grad`.  Returns (d_params, kernels enqueued)."""
    _need_cuda(flat, bw_params, obs, noise['eps_where'], noise['eps_what'], stash, d_log_w)
    if workspace is None:
        workspace = torch.empty(_capi.query_train_sizes(cfg).workspace_floats, dtype=torch.float32, device=flat.device)
    if d_params is None:
        d_params = torch.empty_like(flat)
    _need_cuda(workspace, d_params)
    if d_disc_lp is not None:
        _need_cuda(d_disc_lp)
    n = C.c_int32(0)
    check(_capi.lib().sqair_backward(C.byref(cfg), _ptr(flat), _ptr(bw_params), _ptr(obs), _ptr(noise['eps_where']),
                                     _ptr(noise['eps_what']), _ptr(stash), _ptr(d_log_w),
                                     _ptr(d_disc_lp) if d_disc_lp is not None else None, _ptr(workspace), _ptr(d_params),
                                     C.byref(n), _stream()))
    return d_params, n.value


def optimizer_update(kind, params, grad, slot0, slot1, lr, hyper_a, hyper_b, epsilon, grad_scale=1., l2_weight=0.):
    """One optimiser step on the flat canonical buffers, in place (scripts/experiment.py:138-155; TF 1.x update rules)."""
    _need_cuda(params, grad)
    if grad.numel() != params.numel():
        raise ValueError('gradient and parameter buffers differ in size')
    for s in (slot0, slot1):
        if s is not None:
            _need_cuda(s)
    check(_capi.lib().sqair_optimizer_update(int(kind), _ptr(params), _ptr(grad), _ptr(slot0) if slot0 is not None else None,
                                             _ptr(slot1) if slot1 is not None else None, params.numel(), float(lr),
                                             float(hyper_a), float(hyper_b), float(epsilon), float(grad_scale),
                                             float(l2_weight), _stream()))
    return params


def objective(log_w_t: torch.Tensor, disc_lp_t: torch.Tensor, B: int, K: int):
    """Particle reductions of Model._build / make_target (model.py:88-103,150-158)."""
    _need_cuda(log_w_t, disc_lp_t)
    T = log_w_t.shape[0]
    dev = log_w_t.device
    lw = torch.empty(B, K, dtype=torch.float32, device=dev)
    pe = torch.empty(B, dtype=torch.float32, device=dev)
    iw = torch.empty(B, K, dtype=torch.float32, device=dev)
    sc = torch.empty(_capi.OBJ_N, dtype=torch.float32, device=dev)
    check(_capi.lib().sqair_objective(_ptr(log_w_t), _ptr(disc_lp_t), T, B, K, _ptr(lw), _ptr(pe), _ptr(iw), _ptr(sc),
                                      _stream()))
    return dict(log_weights=lw, elbo_iwae_per_example=pe, importance_weights=iw, scalars=sc)


def objective_grad(log_w_t: torch.Tensor, disc_lp_t: torch.Tensor, B: int, K: int, out=None):
    """Gradients of the VIMCO target (model.py:150-158 / targets.py:62-75) w.r.t. the rows' summed log weights and
    discrete log-probs, both [B,K]; every per-frame term of a row receives the row's value."""
    _need_cuda(log_w_t, disc_lp_t)
    T = log_w_t.shape[0]
    if out is not None:
        d_lw, d_lp = out
        _need_cuda(d_lw, d_lp)
    else:
        d_lw = torch.empty(B, K, dtype=torch.float32, device=log_w_t.device)
        d_lp = torch.empty(B, K, dtype=torch.float32, device=log_w_t.device)
    check(_capi.lib().sqair_objective_grad(_ptr(log_w_t), _ptr(disc_lp_t), T, B, K, _ptr(d_lw), _ptr(d_lp), _stream()))
    return d_lw, d_lp


def stn_glimpse(img: torch.Tensor, where: torch.Tensor, G: int) -> torch.Tensor:
    """SpatialTransformer forward at where-logits (modules.py:165-172,204-227): [N,H,W],[N,4] -> [N,G,G]."""
    _need_cuda(img, where)
    N, H, W = img.shape
    out = torch.empty(N, G, G, dtype=torch.float32, device=img.device)
    check(_capi.lib().sqair_stn_glimpse(_ptr(img), _ptr(where), _ptr(out), N, H, W, G, _stream()))
    return out


def stn_glimpse_grad(img: torch.Tensor, where: torch.Tensor, d_glimpse: torch.Tensor) -> torch.Tensor:
    """Backward of `stn_glimpse` w.r.t. the where-logits: [N,H,W], [N,4], [N,G,G] -> [N,4]."""
    _need_cuda(img, where, d_glimpse)
    N, H, W = img.shape
    G = d_glimpse.shape[-1]
    out = torch.empty(N, 4, dtype=torch.float32, device=img.device)
    check(_capi.lib().sqair_stn_glimpse_grad(_ptr(img), _ptr(where), _ptr(d_glimpse), _ptr(out), N, H, W, G, _stream()))
    return out


def canvas_ll(glimpse, where, presence, mean_img, img, output_std=0.3, bg_std=None):
    """AIRDecoder canvas composition + pixel log-likelihood (modules.py:435-467; seq.py:272-273)."""
    _need_cuda(glimpse, where, presence, mean_img, img)
    N, n, G, _ = glimpse.shape
    H, W = img.shape[1:]
    canvas = torch.empty(N, H, W, dtype=torch.float32, device=img.device)
    ll = torch.empty(N, dtype=torch.float32, device=img.device)
    check(_capi.lib().sqair_canvas_ll(_ptr(glimpse), _ptr(where), _ptr(presence), _ptr(mean_img), _ptr(img),
                                      _ptr(canvas), _ptr(ll), N, n, H, W, G, float(output_std),
                                      float(output_std if bg_std is None else bg_std), _stream()))
    return canvas, ll


def canvas_ll_grad(glimpse, where, presence, mean_img, img, d_ll, output_std=0.3, bg_std=None):
    """Backward of `canvas_ll` for an upstream gradient `d_ll` [N] on the pixel log-likelihood: returns
    (d_glimpse [N,n,G,G], d_where [N,n,4], d_mean_img [H,W])."""
    _need_cuda(glimpse, where, presence, mean_img, img, d_ll)
    N, n, G, _ = glimpse.shape
    H, W = img.shape[1:]
    d_gl = torch.empty_like(glimpse)
    d_wh = torch.empty(N, n, 4, dtype=torch.float32, device=img.device)
    d_mi = torch.zeros(H, W, dtype=torch.float32, device=img.device)
    check(_capi.lib().sqair_canvas_ll_grad(_ptr(glimpse), _ptr(where), _ptr(presence), _ptr(mean_img), _ptr(img), _ptr(d_ll),
                                           _ptr(d_gl), _ptr(d_wh), _ptr(d_mi), N, n, H, W, G, float(output_std),
                                           float(output_std if bg_std is None else bg_std), _stream()))
    return d_gl, d_wh, d_mi


def wgrad(x: torch.Tensor, dy: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Weight gradient of a dense layer: x [M,K], dy [M,N] -> x^T dy [K,N] (added into `out` when given)."""
    _need_cuda(x, dy)
    M, K = x.shape
    N = dy.shape[1]
    if dy.shape[0] != M:
        raise ValueError('x and dy must have the same number of rows')
    acc = out is not None
    if out is None:
        out = torch.empty(K, N, dtype=torch.float32, device=x.device)
    else:
        _need_cuda(out)
    check(_capi.lib().sqair_wgrad(_ptr(x), _ptr(dy), _ptr(out), M, K, N, int(acc), _stream()))
    return out


def dgrad(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Input gradient of a dense layer: dy [M,N], w [K,N] (the reference's variable layout) -> dy w^T [M,K]."""
    _need_cuda(dy, w)
    M, N = dy.shape
    K = w.shape[0]
    if w.shape[1] != N:
        raise ValueError('dy and w must have the same number of columns')
    out = torch.empty(M, K, dtype=torch.float32, device=dy.device)
    check(_capi.lib().sqair_dgrad(_ptr(dy), _ptr(w), _ptr(out), M, K, N, _stream()))
    return out
