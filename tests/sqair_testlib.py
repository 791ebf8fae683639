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
                          cfg.prop_prior_step_bias, cfg.output_std, None, cfg.where_update_scale, cfg.min_std,
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
PIXEL_SUMS = ('data_ll_per_sample', 'log_weights_per_timestep')


def compare_outputs(got: dict, want: dict, rtol=RTOL, atol=ATOL, names=None):
    """Returns a list of human-readable mismatches (empty == parity).

    * integer-valued outputs (EXACT): bit-exact;
    * canvas: the inverse transformer amplifies fp32-level (1e-6) differences of `where` by (G-1)/(2 sx) ~ 1e2 at
      glimpse edges, so >= 99.99% of the pixels must meet rtol/atol 1e-4 and every pixel 2e-3 (values in [0, 1]);
    * pixel sums: atol = 2e-5 * H*W;
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
            at = 2e-5 * want['canvas'].shape[-1] * want['canvas'].shape[-2]
        err = np.abs(a - b) - (at + rtol * np.abs(b))
        finite = np.isfinite(a).all()
        nviol = int((err > 0).sum())
        if k == 'canvas' and finite:
            if nviol <= 1e-4 * a.size and np.abs(a - b).max() <= 2e-3:
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
                os.path.join(ROOT, 'sqair_b200', 'csrc', 'sqair_core.h'), os.path.join(ROOT, 'include', 'sqair_b200.h')]
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
    rc = emu_lib().emu_forward(C.byref(ccfg), flat.ctypes.data, imgs.ctypes.data, nz['eps_where'].ctypes.data,
                               nz['eps_what'].ctypes.data, nz['u_pres'].ctypes.data, C.byref(so), R, cluster)
    assert rc == 0, rc
    return outs
