"""Building-block specifications with the constructor signatures of sqair/modules.py.

Each class records the hyper-parameters the reference's graph builder would have used; the
arithmetic (dense layers, glimpse sampler, distributions) is executed by the fused CUDA kernel.
`SpatialTransformer` is also callable on device tensors through the stand-alone sampler kernel."""
import math

import numpy as np
import torch

from . import ops
from .neural import MLP, _flatten


class GaussianFromParamVec(object):
    def __init__(self, n_params, scale_offset=0., min_std=1e-2, *args, **kwargs):   # modules.py:47
        self.n_params, self.scale_offset, self.min_std = n_params, scale_offset, min_std


class StochasticTransformParam(object):
    def __init__(self, n_hidden, scale_offset=-2.):                                 # modules.py:84
        self.n_hidden, self.scale_offset = _flatten(n_hidden), scale_offset


class Encoder(object):
    def __init__(self, n_hidden):                                                   # modules.py:104
        self.n_hidden = _flatten(n_hidden)


class Decoder(object):
    def __init__(self, n_hidden, output_size, output_scale=.25):                    # modules.py:135
        self.n_hidden, self.output_size, self.output_scale = _flatten(n_hidden), tuple(output_size), output_scale


class SpatialTransformer(object):
    """modules.py:150-227.  Forward transformer callable: `st(img, logits=...)` or `st(img, coords=...)`."""

    def __init__(self, img_size, crop_size, inverse=False):
        self.img_size, self.crop_size, self.inverse = tuple(img_size), tuple(crop_size), inverse

    def __call__(self, img, sentinel=None, coords=None, logits=None):
        if sentinel is not None:
            raise ValueError('Either coords or logits must be given by kwargs!')
        if coords is not None and logits is not None:
            raise ValueError('Please give eithe coords or logits, not both!')
        if coords is None and logits is None:
            raise ValueError('Please give coords or logits!')
        if self.inverse:
            raise NotImplementedError('the inverse transformer runs fused with the likelihood: ops.canvas_ll')
        if logits is None:
            logits = self.to_logits(coords, eps=0.)
        if img.dim() == 4:
            img = img[..., 0]
        return ops.stn_glimpse(img.contiguous(), logits.contiguous(), self.crop_size[0])

    @staticmethod
    def to_coords(logits):
        return torch.cat((torch.sigmoid(logits[..., :2]), torch.tanh(logits[..., 2:])), -1)

    @staticmethod
    def to_logits(coords, eps=1e-4):
        scale, shift = coords[..., :2], coords[..., 2:]
        scale = scale.clamp(eps, 1. - eps)
        shift = shift.clamp(eps - 1., 1. - eps)
        return torch.cat((torch.log(scale / (1. - scale)), 0.5 * (torch.log1p(shift) - torch.log1p(-shift))), -1)

    @staticmethod
    def stn_to_pixel_coord(scale, translation, length):
        return 0.5 * (length - 1.) * (translation - scale + 1.), (length + 1.) * scale


class AIREncoder(object):
    def __init__(self, img_size, glimpse_size, n_what, glimpse_encoder, scale_offset=0.,
                 masked_glimpse=False, debug=False):                                # modules.py:309
        self.img_size, self.glimpse_size, self.n_what = tuple(img_size), tuple(glimpse_size), n_what
        self.glimpse_encoder, self.scale_offset, self.masked_glimpse = glimpse_encoder, scale_offset, masked_glimpse
        if scale_offset != 0.:
            raise NotImplementedError('AIREncoder scale_offset != 0 is not wired into the fused kernel')


class AIRDecoder(object):
    def __init__(self, img_size, glimpse_size, glimpse_decoder, batch_dims=2, mean_img=None, output_std=0.3,
                 learn_std=False, bg_std=None, learn_bg_std=False, min_std=0., bg_bigger_than_fg_std=False):
        self.img_size, self.glimpse_size = tuple(img_size), tuple(glimpse_size)     # modules.py:371
        self.glimpse_decoder = glimpse_decoder(glimpse_size)
        self.mean_img = mean_img
        self.output_std = output_std
        self.bg_std = output_std if bg_std is None else bg_std                      # modules.py:406-407
        if learn_std or learn_bg_std or min_std != 0. or bg_bigger_than_fg_std:
            raise NotImplementedError('learned / lower-bounded output std is not wired into the fused kernel')


class StepsPredictor(object):
    def __init__(self, n_hidden, steps_bias=0., max_rel_logit_change=np.inf, max_logit_change=np.inf, **kwargs):
        if max_logit_change != np.inf and max_rel_logit_change != np.inf:           # modules.py:489-490
            raise ValueError('Only one of max_logit_change and max_rel_logit_change can be used!')
        if max_logit_change != np.inf or max_rel_logit_change != np.inf:
            raise NotImplementedError('logit-change clamps (unused by the MNIST config) are not in the fused kernel')
        self.n_hidden, self.steps_bias = _flatten(n_hidden), steps_bias


class AffineDiagNormal(object):
    """modules.py:527-545 (marker: the propagation `where` posterior is MVN-TriL inside the kernel)."""


class RecurrentNormal(object):
    def __init__(self, n_dim, n_hidden, conditional=False, output_initializers=None):   # modules.py:614
        self.n_dim, self.n_hidden, self.conditional, self.output_initializers = n_dim, n_hidden, conditional, output_initializers


class ConditionedNormalAdaptor(object):
    def __init__(self, loc, scale):                                                 # modules.py:633
        self.loc, self.scale = tuple(float(x) for x in loc), tuple(float(x) for x in scale)
