"""MLP-SQAIR for (multi-)MNIST: the drop-in counterpart of the reference's config module
(sqair/configs/mlp_mnist_model.py:42-150) -- same flags, same `load(img, coords, num, mean_img=None,
debug=False) -> Model` protocol, same wiring of cores / priors / decoder -- against this package."""
from functools import partial

import numpy as np

from sqair_b200 import rnn as snt
from sqair_b200.common_model_flags import flags, get_params
from sqair_b200.core import DiscoveryCore, PropagationCore
from sqair_b200.model import Model
from sqair_b200.modules import Encoder, StochasticTransformParam, StepsPredictor, Decoder, AIRDecoder, AIREncoder
from sqair_b200.propagate import make_prior, SequentialSSM
from sqair_b200.seq import SequentialAIR
from sqair_b200.sqair_modules import Propagate, Discover

flags.DEFINE_string('disc_prior_type', 'cat', 'Prior for the number of discovery steps: {geom, cat}.')
flags.DEFINE_float('step_success_prob', 0.75, 'Step success prob of the geometric discovery prior.')
flags.DEFINE_float('disc_step_bias', 1., 'Added to the logit of discovering a new object.')
flags.DEFINE_float('prop_step_bias', 5., 'Added to the logit of propagating an existing object.')
flags.DEFINE_boolean('sample_from_prior', False, 'Samples from the prior instead of q if True.')
flags.DEFINE_boolean('rec_where_prior', True, 'Uses a recurrent prior for where in discovery.')
flags.DEFINE_boolean('debug', False, 'Adds argument checks to distributions.')


def maybe_getattr(obj, name):
    return getattr(obj, name, None) if name is not None else None


def parse_string_flag(flag, dtype=np.float32, sep=',', num_elements=-1):
    try:
        values = [dtype(f.strip()) for f in str(flag).split(sep)]
    except ValueError:
        values = [np.float32(flag)]
    if len(values) == 1 and num_elements > 1:
        values *= num_elements
    elif num_elements != -1 and len(values) != num_elements:
        raise ValueError('Incorrect number of elements in flag "{}"'.format(flag))
    return values


def load(img, coords, num, mean_img=None, debug=False):
    F = flags.FLAGS
    if img.dim() == 4:
        img = img[..., None]
        if mean_img is not None:
            mean_img = np.asarray(mean_img)[..., np.newaxis]
    params = get_params()
    shape = list(img.shape)
    img_size = shape[2:]

    rnn_class = maybe_getattr(snt, F.transition)
    time_rnn_class = maybe_getattr(snt, F.time_transition)
    input_encoder = partial(Encoder, params.n_hiddens)

    def glimpse_encoder():
        return AIREncoder(img_size, params.glimpse_size, F.n_what, Encoder(params.n_hiddens),
                          masked_glimpse=F.masked_glimpse, debug=F.debug)

    transform_estimator = partial(StochasticTransformParam, params.n_hiddens, F.transform_var_bias)
    steps_predictor = partial(StepsPredictor, params.steps_pred_hidden, F.disc_step_bias)

    # discovery
    discover_cell = DiscoveryCore(img_size, params.glimpse_size, F.n_what, rnn_class(params.n_hidden),
                                  input_encoder, glimpse_encoder, transform_estimator, steps_predictor, debug=debug)
    discover = Discover(F.n_steps_per_image, discover_cell, step_success_prob=F.step_success_prob,
                        where_mean=parse_string_flag(F.scale_prior, float, num_elements=2) + [0, 0],
                        disc_prior_type=F.disc_prior_type, rec_where_prior=F.rec_where_prior)

    # propagation: its own RNN cells, every other estimator shared with discovery or freshly built
    input_encoder = lambda: discover_cell._input_encoder
    glimpse_encoder = lambda: discover_cell._glimpse_encoder
    transform_estimator = partial(StochasticTransformParam, params.n_hiddens, F.transform_var_bias)
    steps_predictor = partial(StepsPredictor, params.steps_pred_hidden, F.prop_step_bias)
    propagation_cell = PropagationCore(img_size, params.glimpse_size, F.n_what, rnn_class(params.n_hidden),
                                       input_encoder, glimpse_encoder, transform_estimator, steps_predictor,
                                       time_rnn_class(params.n_hidden), debug=debug)
    prior_rnn = maybe_getattr(snt, F.prior_transition)(params.n_hidden)
    propagation_prior = make_prior(F.prop_prior_type, F.n_what, prior_rnn, F.prop_prior_step_bias)
    propagate = Propagate(SequentialSSM(propagation_cell), propagation_prior)

    # decoder
    glimpse_decoder = partial(Decoder, params.n_hiddens, output_scale=F.output_scale)
    decoder = AIRDecoder(img_size, params.glimpse_size, glimpse_decoder, batch_dims=2, mean_img=mean_img,
                         output_std=F.output_std)

    # sequence + model
    time_cell = maybe_getattr(snt, F.time_transition)
    if time_cell is not None:
        time_cell = time_cell(params.n_hidden)
    sequence_apdr = SequentialAIR(F.n_steps_per_image, params.glimpse_size, discover, propagate, time_cell, decoder,
                                  sample_from_prior=F.sample_from_prior)
    return Model(img, coords, sequence_apdr, F.k_particles, num, debug)
