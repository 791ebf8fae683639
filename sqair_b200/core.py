"""RNN cores for discovery and propagation: constructor surface of sqair/core.py:51,241.
The encoder / estimator arguments are zero-argument factories, as in the reference; propagation is
expected to reuse discovery's glimpse encoder (configs/mlp_mnist_model.py:112-113)."""
import numpy as np

from .modules import AffineDiagNormal, SpatialTransformer


class BaseSQAIRCore(object):
    _n_transform_param = 4
    _init_presence_value = None
    _output_names = None

    def __init__(self, img_size, crop_size, n_what, transition, input_encoder, glimpse_encoder,
                 transform_estimator, steps_predictor, where_loc_bias=None, debug=False):
        self._img_size, self._crop_size, self._n_what = tuple(img_size), tuple(crop_size), n_what
        self._n_pix = int(np.prod(self._img_size))
        self._cell = transition
        self._n_hidden = int(self._cell.output_size[0])
        if where_loc_bias is not None:
            raise NotImplementedError('where_loc_bias is not wired into the fused kernel')
        self._debug = debug
        self._spatial_transformer = SpatialTransformer(img_size, crop_size)
        self._transform_estimator = transform_estimator()
        self._input_encoder = input_encoder()
        self._glimpse_encoder = glimpse_encoder()
        self._steps_predictor = steps_predictor()

    @property
    def n_what(self):
        return self._n_what

    @property
    def output_names(self):
        return self._output_names


class DiscoveryCore(BaseSQAIRCore):
    _output_names = 'what what_loc what_scale where where_loc where_scale presence_prob presence presence_logit'.split()
    _init_presence_value = 1.


class PropagationCore(BaseSQAIRCore):
    _output_names = ('what what_sample what_loc what_scale where where_sample where_loc where_scale presence_prob'
                     ' presence presence_logit temporal_state').split()
    _init_presence_value = 0.

    def __init__(self, img_size, crop_size, n_what, transition, input_encoder, glimpse_encoder, transform_estimator,
                 steps_predictor, temporal_cell, where_update_scale=1.0, debug=False):
        super(PropagationCore, self).__init__(img_size, crop_size, n_what, transition, input_encoder, glimpse_encoder,
                                              transform_estimator, steps_predictor, debug=debug)
        self._temporal_cell = temporal_cell
        self._where_update_scale = where_update_scale
        self._where_distrib = AffineDiagNormal()
