"""Model hyper-parameter flags shared by model configs (mirrors sqair/common_model_flags.py:32-71:
same names, defaults and `get_params()` contract)."""
from types import SimpleNamespace

from . import tf_flags as flags

flags.DEFINE_float('transform_var_bias', -3., 'Bias added to the variance logit of Gaussian `where` distributions.')
flags.DEFINE_float('output_scale', .25, 'Scales the output mean of the glimpse decoder.')
flags.DEFINE_string('scale_prior', '-2', 'One float or comma-separated floats: mean of the Gaussian prior for scale logit.')
flags.DEFINE_integer('glimpse_size', 20, 'Glimpse size.')
flags.DEFINE_float('prop_prior_step_bias', 10., 'Bias of the propagation prior presence logit.')
flags.DEFINE_string('prop_prior_type', 'rnn', 'Propagation prior: rnn | rw | guided.')
flags.DEFINE_boolean('masked_glimpse', True, 'Masks glimpses in propagation if True.')
flags.DEFINE_integer('k_particles', 5, 'Number of particles of the IWAE bound.')
flags.DEFINE_integer('n_steps_per_image', 3, 'Number of inference steps (object slots) per frame.')
flags.DEFINE_string('transition', 'VanillaRNN', 'RNN core of the discovery and propagation cores.')
flags.DEFINE_string('time_transition', 'GRU', 'RNN core of the temporal rnn in the propagation core.')
flags.DEFINE_string('prior_transition', 'GRU', 'RNN core of the propagation prior.')
flags.DEFINE_float('output_std', .3, 'Standard deviation of Gaussian p(x|z).')
flags.DEFINE_integer('n_units', 8, 'Width in units of 32 neurons; 8 means 256.')
flags.DEFINE_integer('n_what', 50, 'Dimensionality of `what` variables.')


def get_params():
    F = flags.FLAGS
    params = SimpleNamespace(glimpse_size=[F.glimpse_size] * 2, n_hidden=32 * F.n_units, n_layers=2)
    # the reference wraps both in 1-tuples by accident (common_model_flags.py:68-69); consumers flatten
    params.n_hiddens = ([params.n_hidden] * params.n_layers,)
    params.steps_pred_hidden = ([params.n_hidden // 2],)
    return params
