"""Shared helpers for the parity tests: oracle <-> C-ABI struct conversion, inputs, comparison."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

from oracle import sqair_oracle as O
from oracle import synthetic as S
from sqair_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# outputs that hold integer-valued decisions: must match bit-exactly
EXACT = ('presence obj_id prop_pres disc_pres num_prop_steps_per_sample num_disc_steps_per_sample '
         'num_steps_per_sample').split()
RTOL, ATOL = 1e-4, 1e-4     # north-star tolerance: 1e-4 relative fp32 (atol for values near 0)


def capi_cfg(cfg: O.Cfg) -> _capi.SqairCfg:
    return _capi.make_cfg(cfg.T, cfg.B, cfg.K, cfg.n, cfg.H, cfg.W, cfg.G, cfg.nw, cfg.nh, cfg.prior_type,
                          cfg.disc_prior_type, cfg.rec_where_prior, cfg.masked_glimpse, cfg.step_success_prob,
                          cfg.prop_prior_step_bias, cfg.output_std, cfg.bg_std, cfg.where_update_scale, cfg.min_std,
                          cfg.where_mean, cfg.where_std)


def make_inputs(cfg: O.Cfg, data_seed=1234, weight_seed=42, noise_seed=7, jitter=0.1, n_max=None, philox=True):
    imgs, nums = S.make_sequences(cfg.T, cfg.B, cfg.H, cfg.W, cfg.n if n_max is None else n_max, seed=data_seed)
    params = O.init_params(cfg, weight_seed, mean_img=imgs.mean((0, 1)), jitter=jitter)
    if philox:
        noise = S.philox_noise(cfg.T, cfg.rows, cfg.n, cfg.nw, noise_seed)
    else:
        noise = S.numpy_noise(cfg.T, cfg.rows, cfg.n, cfg.nw, noise_seed)
    return imgs, params, noise


def run_oracle(cfg, imgs, params, noise):
    with torch.no_grad():
        out, obj = O.model_forward(params, cfg, torch.from_numpy(imgs),
                                   {k: torch.from_numpy(v) for k, v in noise.items()})
    return {k: v.numpy() for k, v in out.items()}, {k: v.numpy() for k, v in obj.items()}


# Outputs whose value is a SUM of H*W per-pixel log-densities (|summand| up to ~1e2) that largely cancel: fp32
# accumulation-order noise is ~1e-7 * sum|summands|, so their absolute tolerance scales with the pixel count.
# Measured noise floor (the fp32 oracle against the SAME oracle in float64, tools/oracle_noise_floor.py ->
# profiles/r02_oracle_fp32_vs_fp64_{c2,c4}.txt): data_ll / log weights 6.4e-3 at 50x50 (2.6e-6 H W), canvas max 1.6e-4
# with 3e-6 of the pixels beyond 1e-4 + 1e-4 |v|, every other output <= 2e-4 absolute / 1e-3 relative.  Two fp32
# implementations differ from each other by up to twice that; the tolerances below are ~4x the floor.
PIXEL_SUMS = ('data_ll_per_sample', 'log_weights_per_timestep')
PIXEL_SUM_ATOL = 1e-5          # x H*W


def compare_outputs(got: dict, want: dict, rtol=RTOL, atol=ATOL, names=None):
    """Returns a list of human-readable mismatches (empty == parity).

    * integer-valued outputs (EXACT): bit-exact;
    * canvas: the inverse transformer amplifies fp32-level (1e-6) differences of `where` by (G-1)/(2 sx) ~ 1e2 at
      glimpse edges, so >= 99.995% of the pixels must meet rtol/atol 1e-4 and every pixel 1e-3 (values in [0, 1]);
    * pixel sums: atol = 1e-5 * H*W;
    * everything else: |got - want| <= atol + rtol * |want|."""
    bad = []
    for k in (names or _capi.OUTPUT_NAMES):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape:
            bad.append('%s: shape %s vs %s' % (k, a.shape, b.shape))
            continue
        if k in EXACT:
            nbad = int((a != b).sum())
            if nbad:
                bad.append('%s: %d/%d entries differ (must be exact)' % (k, nbad, a.size))
            continue
        at = atol
        if k in PIXEL_SUMS:
            at = PIXEL_SUM_ATOL * want['canvas'].shape[-1] * want['canvas'].shape[-2]
        err = np.abs(a - b) - (at + rtol * np.abs(b))
        finite = np.isfinite(a).all()
        nviol = int((err > 0).sum())
        if k == 'canvas' and finite:
            if nviol <= 5e-5 * a.size and np.abs(a - b).max() <= 1e-3:
                continue
        if not finite or nviol:
            e2 = np.where(np.isfinite(err), err, np.inf)
            i = np.unravel_index(np.argmax(e2), err.shape)
            bad.append('%s: %d bad; worst at %s got %r want %r (|d|=%.3g)'
                       % (k, nviol, i, a[i], b[i], abs(a[i] - b[i])))
    return bad


# ---------------------------------------------------------------------------------------------
# host emulator of the kernel program (tests/host_emu) -- CPU-side logic check, test-only
# ---------------------------------------------------------------------------------------------
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, 'tests', 'host_emu')
        so = os.path.join(d, 'libsqair_emu.so')
        srcs = [os.path.join(d, 'emu.cpp'), os.path.join(ROOT, 'sqair_b200', 'csrc', 'sqair_device.cuh'),
                os.path.join(ROOT, 'sqair_b200', 'csrc', 'sqair_core.h'), os.path.join(ROOT, 'include', 'sqair_b200.h'),
                os.path.join(ROOT, 'sqair_b200', 'csrc', 'sqair_backward.h')]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-pthread', '-Wno-unknown-pragmas',
                                   '-o', so, srcs[0]])
        _emu = C.CDLL(so)
        _emu.emu_forward.argtypes = [C.POINTER(_capi.SqairCfg)] + [C.c_void_p] * 5 + \
            [C.POINTER(_capi.SqairOutputs), C.c_int, C.c_int]
        _emu.emu_forward.restype = C.c_int
        _emu.emu_smem_floats.argtypes = [C.POINTER(_capi.SqairCfg), C.c_int, C.c_int]
        _emu.emu_smem_floats.restype = C.c_int
    return _emu


def with_prior_noise(cfg, noise, seed=8):
    """Adds the second noise set of the generation mode (draws from the priors) to a noise dict."""
    pn = S.philox_noise(cfg.T, cfg.rows, cfg.n, cfg.nw, seed)
    out = dict(noise)
    out.update({k + '_prior': v for k, v in pn.items()})
    return out


def run_emu(cfg: O.Cfg, imgs, params, noise, R, cluster=1):
    ccfg = capi_cfg(cfg)
    flat = np.ascontiguousarray(O.flatten_params(params, cfg).numpy())
    shapes = _capi.output_shapes(ccfg)
    outs = {k: np.full(s, np.nan, dtype=np.float32) for k, s in shapes.items()}
    so = _capi.SqairOutputs()
    for k in _capi.OUTPUT_NAMES:
        setattr(so, k, outs[k].ctypes.data)
    imgs = np.ascontiguousarray(imgs, dtype=np.float32)
    nz = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in noise.items()}
    if cfg.sample_from_prior:
        lib = emu_lib()
        lib.emu_forward_generate.argtypes = [C.POINTER(_capi.SqairCfg)] + [C.c_void_p] * 8 + [C.c_int, C.POINTER(_capi.SqairOutputs),
                                                                                             C.c_int, C.c_int]
        lib.emu_forward_generate.restype = C.c_int
        rc = lib.emu_forward_generate(C.byref(ccfg), flat.ctypes.data, imgs.ctypes.data, nz['eps_where'].ctypes.data,
                                      nz['eps_what'].ctypes.data, nz['u_pres'].ctypes.data, nz['eps_where_prior'].ctypes.data,
                                      nz['eps_what_prior'].ctypes.data, nz['u_pres_prior'].ctypes.data, cfg.generate_after,
                                      C.byref(so), R, cluster)
    else:
        rc = emu_lib().emu_forward(C.byref(ccfg), flat.ctypes.data, imgs.ctypes.data, nz['eps_where'].ctypes.data,
                                   nz['eps_what'].ctypes.data, nz['u_pres'].ctypes.data, C.byref(so), R, cluster)
    assert rc == 0, rc
    return outs


# ---------------------------------------------------------------------------------------------
# gradients: oracle (torch autograd) and the emulated backward pass
# ---------------------------------------------------------------------------------------------
last_kink_distance = np.inf        # closest approach of a resampler coordinate to an integer in the last float64 run
KINK_EPS = 3e-5                    # fp32 rounding of a pixel coordinate (magnitude <= 100) is ~1e-5


def smooth_inputs(cfg, **kw):
    """make_inputs + float64 oracle gradients for the first noise seed whose resampler coordinates stay KINK_EPS away
    from every integer (where the reference gradient itself is one-sided and fp32 rounding picks the side).
    Returns (imgs, params, noise, want)."""
    for seed in range(7, 27):
        imgs, params, noise = make_inputs(cfg, noise_seed=seed, **kw)
        want, _ = oracle_gradients(cfg, imgs, params, noise, double=True)
        if last_kink_distance >= KINK_EPS:
            return imgs, params, noise, want
    raise AssertionError('no kink-free noise seed found')


def oracle_gradients(cfg, imgs, params, noise, target='auto', double=False):
    """d target / d every variable by torch autograd through the oracle -> ({name: array}, objective dict).
    double=True evaluates the restatement in float64 (the trustworthy value of sums with heavy cancellation).
    Every variable must receive a gradient (model.py:163-166); with the `geom` count prior the variables of the `cat`
    prior are not part of the reference graph at all and come back as zeros."""
    obs, nz = torch.from_numpy(imgs), {k: torch.from_numpy(v) for k, v in noise.items()}
    if double:
        global last_kink_distance
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        resample, dist = O._bilinear_zero_pad, [np.inf]

        def watched(img, x, y):
            # Bilinear interpolation is continuous but its position derivative jumps at integer coordinates: a sample
            # within fp32 rounding of one gets either one-sided derivative, both valid.  Record how close the run comes.
            with torch.no_grad():
                for c, size in ((x, img.shape[2]), (y, img.shape[1])):
                    c = c[(c > -1) & (c < size)]
                    if c.numel():
                        dist[0] = min(dist[0], float((c - torch.round(c)).abs().min()))
            return resample(img, x, y)

        O._bilinear_zero_pad = watched
        try:
            g, obj, missing = O.model_gradients({k: v.double() for k, v in params.items()}, cfg, obs.double(),
                                                {k: v.double() for k, v in nz.items()}, target)
        finally:
            O._bilinear_zero_pad = resample
            torch.set_default_dtype(old)
        last_kink_distance = dist[0]
    else:
        g, obj, missing = O.model_gradients(params, cfg, obs, nz, target)
    if cfg.disc_prior_type == 'geom':
        missing = [k for k in missing if 'discover/mlp/' not in k and 'step_prior' not in k]
    assert not missing, missing
    return {k: v.numpy() for k, v in g.items()}, {k: v.numpy() for k, v in obj.items()}


def objective_grads(log_w_t, disc_lp_t, B, K, vimco=True):
    """torch restatement of the first backward stage (targets.py:46-75): gradients of the target w.r.t. the rows' summed
    log weights / discrete log-probs, each [B*K]."""
    T = log_w_t.shape[0]
    a = torch.as_tensor(log_w_t).sum(0).reshape(B, K).clone().requires_grad_(True)
    b = torch.as_tensor(disc_lp_t).sum(0).reshape(B, K).clone().requires_grad_(True)
    tgt = (O.vimco(a, b, O.iwae(a)) if vimco else -O.iwae(a).mean()) / T
    ga, gb = torch.autograd.grad(tgt, [a, b], allow_unused=True)
    gb = torch.zeros_like(a) if gb is None else gb
    return ga.reshape(-1).numpy().copy(), gb.reshape(-1).numpy().copy()


def run_emu_backward(cfg: O.Cfg, imgs, params, noise, R, cluster=1, vimco=None):
    """Emulated forward (with stash) + backward -> ({name: gradient}, outputs)."""
    lib = emu_lib()
    lib.emu_forward_backward.argtypes = [C.POINTER(_capi.SqairCfg)] + [C.c_void_p] * 5 + \
        [C.POINTER(_capi.SqairOutputs), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_forward_backward.restype = C.c_int
    vimco = cfg.K > 1 if vimco is None else vimco
    # upstream gradients from the oracle's own log weights (they match the kernel's to 1e-4; parity-tested separately)
    want, _ = run_oracle(cfg, imgs, params, noise)
    gw, gp = objective_grads(want['log_weights_per_timestep'], want['discrete_log_prob'], cfg.B, cfg.K, vimco)
    ccfg = capi_cfg(cfg)
    flat = np.ascontiguousarray(O.flatten_params(params, cfg).numpy())
    shapes = _capi.output_shapes(ccfg)
    outs = {k: np.full(s, np.nan, dtype=np.float32) for k, s in shapes.items()}
    so = _capi.SqairOutputs()
    for k in _capi.OUTPUT_NAMES:
        setattr(so, k, outs[k].ctypes.data)
    imgs = np.ascontiguousarray(imgs, dtype=np.float32)
    nz = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in noise.items()}
    gw, gp = np.ascontiguousarray(gw, dtype=np.float32), np.ascontiguousarray(gp, dtype=np.float32)
    dflat = np.full_like(flat, np.nan)
    rc = lib.emu_forward_backward(C.byref(ccfg), flat.ctypes.data, imgs.ctypes.data, nz['eps_where'].ctypes.data,
                                  nz['eps_what'].ctypes.data, nz['u_pres'].ctypes.data, C.byref(so), R, cluster,
                                  gw.ctypes.data, gp.ctypes.data, dflat.ctypes.data)
    assert rc == 0, rc
    grads = {k: v.numpy() for k, v in O.unflatten_params(torch.from_numpy(dflat), cfg).items()}
    return grads, outs


def compare_gradients(got: dict, want: dict, rtol=1e-3, atol_rel=2e-4, floor: dict = None):
    """Per variable: |got - want| <= rtol * |want| + atol_rel * max|want_variable| (model.py:163-166 demands a gradient
    for every variable: none may be missing).  `floor` (optional, {name: array}): an fp32 evaluation of the same
    gradient by the oracle; 3 x its own distance from `want` (float64) is added to the tolerance -- the measured fp32
    noise floor of sums with heavy cancellation (e.g. the scalar scale offsets).  Returns mismatch descriptions."""
    bad = []
    for k, w in want.items():
        g = got.get(k)
        if g is None:
            bad.append('%s: missing' % k)
            continue
        g, w = np.asarray(g, dtype=np.float64), np.asarray(w, dtype=np.float64)
        if g.shape != w.shape:
            bad.append('%s: shape %s vs %s' % (k, g.shape, w.shape))
            continue
        scale = np.abs(w).max()
        tol = rtol * np.abs(w) + atol_rel * scale + 1e-12
        if floor is not None:
            tol = tol + 3.0 * np.abs(np.asarray(floor[k], dtype=np.float64) - w).max()
        err = np.abs(g - w) - tol
        if not np.isfinite(g).all() or (err > 0).any():
            i = np.unravel_index(np.argmax(np.where(np.isfinite(err), err, np.inf)), err.shape) if err.ndim else ()
            bad.append('%s: %d/%d bad, max|want| %.3e, worst at %s got %.6e want %.6e' %
                       (k, int((err > 0).sum()) + int((~np.isfinite(g)).sum()), g.size, scale, i, g[i], w[i]))
    return bad


def run_cuda_backward(cfg: O.Cfg, imgs, params, noise, vimco=None, return_outputs=False, keep_alive=None):
    """The product path: sqair_forward_train -> sqair_objective_grad -> sqair_backward, through the C ABI.  `keep_alive`: a
    list that receives the device buffers of the call (so that a later call cannot get the same addresses)."""
    from sqair_b200 import ops
    dev = torch.device('cuda:0')
    ccfg = capi_cfg(cfg)
    vimco = cfg.K > 1 if vimco is None else vimco
    flat = O.flatten_params(params, cfg).to(dev)
    packed = ops.pack_params(ccfg, flat)
    bw = ops.pack_backward(ccfg, flat)
    ts = _capi.query_train_sizes(ccfg)
    stash = torch.full((ts.stash_floats,), float('nan'), dtype=torch.float32, device=dev)      # poison: every read must have been written
    nz = {k: torch.from_numpy(v).to(dev) for k, v in noise.items()}
    obs = torch.from_numpy(imgs).to(dev)
    out = ops.forward(ccfg, packed, obs, nz, stash=stash)
    d_lw, d_lp = ops.objective_grad(out['log_weights_per_timestep'], out['discrete_log_prob'], cfg.B, cfg.K)
    ws = torch.full((ts.workspace_floats,), float('nan'), dtype=torch.float32, device=dev)
    d_params, launches = ops.backward(ccfg, flat, bw, obs, nz, stash, d_lw, d_lp if vimco else None, workspace=ws)
    torch.cuda.synchronize()
    if keep_alive is not None:
        keep_alive.append((flat, packed, bw, stash, nz, obs, out, d_lw, d_lp, ws, d_params))
    grads = {k: v.numpy() for k, v in O.unflatten_params(d_params.cpu(), cfg).items()}
    if return_outputs:
        return grads, {k: v.cpu().numpy() for k, v in out.items()}, launches
    return grads


def compare_objective(got_scalars, want_obj, cfg, vimco=True):
    """Objective scalars of Model._build / make_target (model.py:88-103,150-158): [elbo_vae, elbo_iwae, ess, vimco_target,
    iwae_target] from `sqair_objective` against the oracle.  The ELBOs are means of sums over T frames of pixel sums
    (atol 1e-5 * H*W per frame, as for `log_weights_per_timestep`) -> rtol 1e-4 + that floor; the targets are the same
    quantities / T; ESS depends on differences of log weights through a softmax -> rtol 1e-3."""
    g = np.asarray(got_scalars, dtype=np.float64)
    floor = PIXEL_SUM_ATOL * cfg.H * cfg.W * cfg.T
    bad = []
    checks = [('elbo_vae', _capi.OBJ_ELBO_VAE, 1e-4, floor), ('elbo_iwae', _capi.OBJ_ELBO_IWAE, 1e-4, floor),
              ('ess', _capi.OBJ_ESS, 1e-3, 1e-3), ('iwae_target', _capi.OBJ_IWAE_TARGET, 1e-4, floor / cfg.T)]
    if vimco and cfg.K > 1:
        checks.append(('vimco_target', _capi.OBJ_VIMCO_TARGET, 2e-4, 2 * floor / cfg.T))
    for name, idx, rtol, atol in checks:
        w = float(want_obj[name])
        if not abs(g[idx] - w) <= atol + rtol * abs(w):
            bad.append('%s: got %.7g want %.7g' % (name, g[idx], w))
    return bad
