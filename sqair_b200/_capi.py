"""ctypes binding of the C ABI declared in include/sqair_b200.h.

The CUDA library is the only compute path of this package: if `libsqair_b200.so` is missing or
fails to load, importing the binding raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SQAIR_LIB') or os.path.join(_HERE, 'csrc', 'libsqair_b200.so')   # SQAIR_LIB: instrumented build (tuning only)

OUTPUT_NAMES = (
    'what what_loc what_scale where where_loc where_scale presence_prob presence presence_logit '
    'obj_id step_log_prob canvas glimpse '
    'disc_what_log_prob disc_where_log_prob disc_what_prior_log_prob disc_where_prior_log_prob '
    'disc_log_prob disc_prior_log_prob disc_prob '
    'prop_what_log_prob prop_where_log_prob prop_what_prior_log_prob prop_where_prior_log_prob '
    'prop_log_prob prop_prior_log_prob prop_prob discrete_log_prob '
    'num_prop_steps_per_sample num_disc_steps_per_sample num_steps_per_sample prop_pres disc_pres '
    'data_ll_per_sample kl_per_sample log_q_z_given_x_per_sample log_p_z_per_sample '
    'log_weights_per_timestep').split()
assert len(OUTPUT_NAMES) == 38

PRIOR_TYPES = {'rnn': 0, 'rw': 1, 'guided': 2}
DISC_PRIOR_TYPES = {'cat': 0, 'geom': 1}
OPT_RMSPROP, OPT_ADAM, OPT_MOMENTUM, OPT_SGD = 0, 1, 2, 3
OBJ_ELBO_VAE, OBJ_ELBO_IWAE, OBJ_ESS, OBJ_VIMCO_TARGET, OBJ_IWAE_TARGET, OBJ_N = 0, 1, 2, 3, 4, 8


class SqairCfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                'T B K n H W G n_what n_hidden prior_type disc_prior_type rec_where_prior masked_glimpse'.split()] + \
               [(k, C.c_float) for k in
                'step_success_prob prop_prior_step_bias output_std bg_std where_update_scale min_std'.split()] + \
               [('where_mean', C.c_float * 4), ('where_std', C.c_float * 4)]


class SqairSizes(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                'param_count packed_floats eps_where_floats eps_what_floats u_pres_floats'.split()] + \
               [(k, C.c_int32) for k in 'rows rows_per_cta cluster_size n_ctas smem_bytes n_layers'.split()]


class SqairTrainSizes(C.Structure):
    _fields_ = [(k, C.c_int64) for k in 'stash_floats workspace_floats backward_param_floats'.split()]


class SqairParamDesc(C.Structure):
    _fields_ = [('name', C.c_char * 160), ('ndim', C.c_int32), ('shape', C.c_int32 * 3),
                ('offset', C.c_int64), ('packed_offset', C.c_int64)]


class SqairOutputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OUTPUT_NAMES]


def make_cfg(T, B, K, n, H, W, G=20, n_what=50, n_hidden=256, prior_type='rnn', disc_prior_type='cat',
             rec_where_prior=True, masked_glimpse=True, step_success_prob=0.75, prop_prior_step_bias=10.,
             output_std=0.3, bg_std=None, where_update_scale=1.0, min_std=1e-2,
             where_mean=(-2., -2., 0., 0.), where_std=(1., 1., 1., 1.)) -> SqairCfg:
    if prior_type not in PRIOR_TYPES:                                  # propagate.py:42-43
        raise ValueError('Invalid prior type: "{}". Choose from {}.'.format(prior_type, list(PRIOR_TYPES)))
    if disc_prior_type not in DISC_PRIOR_TYPES:                        # sqair_modules.py:223-224
        raise ValueError('Invalid prior type: {}'.format(disc_prior_type))
    c = SqairCfg()
    c.T, c.B, c.K, c.n, c.H, c.W, c.G = T, B, K, n, H, W, G
    c.n_what, c.n_hidden = n_what, n_hidden
    c.prior_type, c.disc_prior_type = PRIOR_TYPES[prior_type], DISC_PRIOR_TYPES[disc_prior_type]
    c.rec_where_prior, c.masked_glimpse = int(bool(rec_where_prior)), int(bool(masked_glimpse))
    c.step_success_prob, c.prop_prior_step_bias = step_success_prob, prop_prior_step_bias
    c.output_std = output_std
    c.bg_std = output_std if bg_std is None else bg_std               # modules.py:406-407
    c.where_update_scale, c.min_std = where_update_scale, min_std
    c.where_mean[:] = list(where_mean)
    c.where_std[:] = list(where_std)
    return c


def output_shapes(cfg: SqairCfg):
    """Shapes of the 38 outputs for a call with `cfg` (seq.py:121-177), leading [T, rows].

    The reference squeezes every per-frame output whose last axis has length 1 before writing it to its
    TensorArray (seq.py:253-255).  presence / ids etc. are [rows, n, 1] there and come out as [rows, n]; the
    per-slot log-probabilities are [rows, n] and therefore lose their slot axis when n_steps_per_image == 1.
    The memory layout is unchanged, only the reported shape follows the quirk."""
    T, rows, n = cfg.T, cfg.B * cfg.K, cfg.n
    per_slot = {'what': (n, cfg.n_what), 'what_loc': (n, cfg.n_what), 'what_scale': (n, cfg.n_what),
                'where': (n, 4), 'where_loc': (n, 4), 'where_scale': (n, 4),
                'canvas': (cfg.H, cfg.W), 'glimpse': (n, cfg.G, cfg.G), 'disc_prob': (n + 1,)}
    slot_vecs = 'presence_prob presence presence_logit obj_id prop_pres disc_pres'.split()
    slot_scalars = ('disc_what_log_prob disc_where_log_prob disc_what_prior_log_prob disc_where_prior_log_prob '
                    'prop_what_log_prob prop_where_log_prob prop_what_prior_log_prob prop_where_prior_log_prob '
                    'prop_prob').split()
    shapes = {}
    for k in OUTPUT_NAMES:
        if k in per_slot:
            shapes[k] = (T, rows) + per_slot[k]
        elif k in slot_vecs:
            shapes[k] = (T, rows, n)
        elif k in slot_scalars:
            shapes[k] = (T, rows, n) if n > 1 else (T, rows)
        else:
            shapes[k] = (T, rows)
    return shapes


def bind(lib):
    """Declares argtypes/restypes of every exported symbol on a loaded CDLL."""
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    lib.sqair_last_error.restype = C.c_char_p
    lib.sqair_last_error.argtypes = []
    lib.sqair_version.restype = C.c_int
    lib.sqair_version.argtypes = []
    lib.sqair_query_sizes.argtypes = [C.POINTER(SqairCfg), C.POINTER(SqairSizes)]
    lib.sqair_param_layout.argtypes = [C.POINTER(SqairCfg), C.POINTER(SqairParamDesc), C.POINTER(i32)]
    lib.sqair_pack_params.argtypes = [C.POINTER(SqairCfg), vp, vp, vp]
    lib.sqair_fill_noise.argtypes = [C.POINTER(SqairCfg), C.c_uint64, i32, vp, vp, vp, vp]
    lib.sqair_forward.argtypes = [C.POINTER(SqairCfg), vp, vp, vp, vp, vp, C.POINTER(SqairOutputs), vp]
    lib.sqair_forward_generate.argtypes = [C.POINTER(SqairCfg), vp, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(SqairOutputs), vp]
    lib.sqair_forward_train.argtypes = [C.POINTER(SqairCfg), vp, vp, vp, vp, vp, C.POINTER(SqairOutputs), vp, vp]
    lib.sqair_query_train_sizes.argtypes = [C.POINTER(SqairCfg), C.POINTER(SqairTrainSizes)]
    lib.sqair_pack_backward.argtypes = [C.POINTER(SqairCfg), vp, vp, vp]
    lib.sqair_backward.argtypes = [C.POINTER(SqairCfg), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i32), vp]
    lib.sqair_optimizer_update.argtypes = [i32, vp, vp, vp, vp, C.c_int64, f32, f32, f32, f32, f32, f32, vp]
    lib.sqair_render_sprites.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.sqair_objective.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.sqair_objective_grad.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.sqair_stn_glimpse.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.sqair_wgrad.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.sqair_dgrad.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.sqair_stn_glimpse_grad.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.sqair_canvas_ll.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp]
    lib.sqair_canvas_ll_grad.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp]
    for name in ('sqair_query_sizes sqair_param_layout sqair_pack_params sqair_fill_noise sqair_forward sqair_forward_train sqair_forward_generate '
                 'sqair_query_train_sizes sqair_pack_backward sqair_backward sqair_optimizer_update sqair_render_sprites '
                 'sqair_objective sqair_objective_grad sqair_wgrad sqair_dgrad sqair_stn_glimpse sqair_stn_glimpse_grad sqair_canvas_ll sqair_canvas_ll_grad').split():
        getattr(lib, name).restype = C.c_int
    return lib


EXPORTED = ('sqair_last_error sqair_version sqair_query_sizes sqair_param_layout sqair_pack_params sqair_forward_train sqair_forward_generate '
            'sqair_query_train_sizes sqair_pack_backward sqair_backward sqair_optimizer_update sqair_render_sprites '
            'sqair_fill_noise sqair_forward sqair_objective sqair_objective_grad sqair_stn_glimpse sqair_stn_glimpse_grad sqair_canvas_ll sqair_canvas_ll_grad sqair_wgrad sqair_dgrad').split()

_lib = None


def lib():
    """The CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('sqair_b200: %s is missing -- build it with `python -c "import __graft_entry__ as g; '
                               'g.build()"` or `make -C sqair_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


class SqairError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib().sqair_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        raise SqairError('sqair_b200 error %d: %s' % (rc, msg))


def param_layout(cfg: SqairCfg):
    """[(name, shape, offset, packed_offset)] of the reference's variables in canonical order."""
    n = C.c_int32(0)
    check(lib().sqair_param_layout(C.byref(cfg), None, C.byref(n)))
    arr = (SqairParamDesc * n.value)()
    check(lib().sqair_param_layout(C.byref(cfg), arr, C.byref(n)))
    return [(d.name.decode(), tuple(d.shape[:d.ndim]), d.offset, d.packed_offset) for d in arr]


def query_train_sizes(cfg: SqairCfg) -> SqairTrainSizes:
    s = SqairTrainSizes()
    check(lib().sqair_query_train_sizes(C.byref(cfg), C.byref(s)))
    return s


def query_sizes(cfg: SqairCfg) -> SqairSizes:
    s = SqairSizes()
    check(lib().sqair_query_sizes(C.byref(cfg), C.byref(s)))
    return s
