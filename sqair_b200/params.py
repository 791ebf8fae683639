"""Parameter store: the reference's trainable variables by TF name (notebooks/play.ipynb:239-362)
in one flat fp32 device buffer (canonical order = `sqair_param_layout`) plus its packed kernel-side copy."""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _capi, ops


def table(cfg):
    """OrderedDict name -> (shape, offset) in canonical order."""
    return OrderedDict((n, (s, o)) for n, s, o, _ in _capi.param_layout(cfg))


def initial_values(cfg, seed=42, spec=None):
    """Sonnet-style initialisation: w ~ TruncNormal(0, 1/sqrt(fan_in)) (+-2 sigma), b = 0,
    GRU matrices glorot-uniform, trainable initial states 0, plus the reference's constant initialisers taken
    from `spec` (steps biases, scale offsets, output_scale, gate / mask biases, where-prior readout bias,
    step_prior_timestep_bias = [10, 0, ...], mean image)."""
    spec = spec or {}
    rng = np.random.default_rng(seed)
    tab = table(cfg)
    total = sum(int(np.prod(s)) for s, _ in tab.values())
    flat = np.zeros(total, dtype=np.float32)

    def view(name):
        s, o = tab[name]
        return flat[o:o + int(np.prod(s))].reshape(s) if len(s) else flat[o:o + 1]

    for name, (shape, off) in tab.items():
        base = name.rsplit('/', 1)[-1]
        if base in ('wz', 'wr', 'wh', 'uz', 'ur', 'uh'):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            view(name)[...] = rng.uniform(-lim, lim, shape)
        elif base == 'w' and 'initial_state' not in name:
            v = rng.standard_normal(shape)
            bad = np.abs(v) > 2.
            while bad.any():
                v[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(v) > 2.
            view(name)[...] = v / math.sqrt(shape[0])
    DC, PC = 'discovery/discovery_core/', 'propagation/propagation_core/'
    RN = 'discovery/discover/recurrent_normal_impl/'
    view('decoder/air_decoder/decoder/output_scale')[...] = spec.get('output_scale', .25)
    view(DC + 'stochastic_transform_param/scale_offset')[...] = spec.get('disc_scale_offset', -3.)
    view(PC + 'stochastic_transform_param/scale_offset')[...] = spec.get('prop_scale_offset', -3.)
    view(DC + 'steps_predictor/mlp/linear_1/b')[...] = spec.get('disc_step_bias', 1.)
    view(PC + 'steps_predictor/mlp/linear_1/b')[...] = spec.get('prop_step_bias', 5.)
    view(PC + 'what/linear/b')[...] = 1.                                            # remember_bias, core.py:345
    if cfg.masked_glimpse:
        view(DC + 'air_encoder/mlp/linear_1/b')[...] = 1.                            # modules.py:324
    if cfg.rec_where_prior:
        view(RN + 'linear/b')[...] = np.asarray(list(spec.get('where_mean', (-2., -2., 0., 0.))) +
                                                list(spec.get('where_std', (1., 1., 1., 1.))))
        lim = math.sqrt(6.0 / 5.0)
        view(RN + 'init_sample')[...] = rng.uniform(-lim, lim, (1, 4))               # tf.get_variable default: glorot uniform
    view('model/sequential_air/while/sqair_timestep/discover/step_prior_timestep_bias')[0] = 10.
    lim = math.sqrt(6.0 / 10.0)
    view(PC + 'affine_diag_normal/cholesky_scale')[...] = rng.uniform(-lim, lim, (10,))
    if spec.get('mean_img') is not None:
        view('decoder/air_decoder/Variable')[...] = np.asarray(spec['mean_img'], dtype=np.float32).reshape(cfg.H, cfg.W, 1)
    return flat


class ParamStore(object):
    def __init__(self, cfg, device, seed=42, spec=None):
        self.cfg, self.device = cfg, device
        self.table = table(cfg)
        spec = dict(spec or {})
        spec.setdefault('where_mean', tuple(cfg.where_mean))
        spec.setdefault('where_std', tuple(cfg.where_std))
        self.flat = torch.from_numpy(initial_values(cfg, seed, spec)).to(device)
        self._version, self._packed, self._bw = 0, {}, {}

    def variables(self):
        """name -> VIEW into the flat buffer (in-place writes must be followed by `mark_dirty()`)."""
        return OrderedDict((n, self.flat[o:o + int(np.prod(s))].reshape(s)) for n, (s, o) in self.table.items())

    def state_dict(self):
        """name -> copy of the variable (checkpoint semantics: later updates do not alias into the returned tensors)."""
        return OrderedDict((n, v.clone()) for n, v in self.variables().items())

    def mark_dirty(self):
        """The flat buffer was modified in place (optimiser step, checkpoint load): kernel-side copies are stale."""
        self._version += 1

    def load_state_dict(self, sd):
        for n, (s, o) in self.table.items():
            self.flat[o:o + int(np.prod(s))].copy_(torch.as_tensor(sd[n], dtype=torch.float32).reshape(-1))
        self.mark_dirty()

    def load_flat(self, flat):
        self.flat.copy_(torch.as_tensor(flat, dtype=torch.float32).reshape(-1))
        self.mark_dirty()

    def packed(self, cfg=None):
        """Kernel-side copy for a call with configuration `cfg` (the packed layout depends on the launch shape the
        library picks for the call: the cluster size); re-packed only after the parameters changed."""
        cfg = cfg or self.cfg
        s = _capi.query_sizes(cfg)
        key = (s.cluster_size, s.packed_floats, cfg.n)
        hit = self._packed.get(key)
        if hit is None or hit[0] != self._version:
            buf = hit[1] if hit is not None else None          # re-pack into the same buffer (training: every step)
            hit = (self._version, ops.pack_params(cfg, self.flat, out=buf))
            self._packed[key] = hit
        return hit[1]

    def backward_params(self, cfg=None):
        """Parameter copy read by the backward GEMMs (`sqair_pack_backward`), cached like `packed`."""
        cfg = cfg or self.cfg
        key = (_capi.query_train_sizes(cfg).backward_param_floats, cfg.n)
        hit = self._bw.get(key)
        if hit is None or hit[0] != self._version:
            hit = (self._version, ops.pack_backward(cfg, self.flat, out=hit[1] if hit is not None else None))
            self._bw[key] = hit
        return hit[1]
