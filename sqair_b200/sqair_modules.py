"""Propagate / Discover / SQAIRTimestep: surface of sqair/sqair_modules.py:66-582."""
from .modules import ConditionedNormalAdaptor, RecurrentNormal


class Discover(object):
    def __init__(self, n_steps, cell, step_success_prob, where_mean=(-2., -2., 0., 0.), where_std=(1., 1., 1., 1.),
                 disc_prior_type='geom', rec_where_prior=False):                    # sqair_modules.py:69
        if disc_prior_type not in ('geom', 'cat'):
            raise ValueError('Invalid prior type: {}'.format(disc_prior_type))      # sqair_modules.py:223-224
        self._n_steps, self._cell = n_steps, cell
        self._init_disc_step_success_prob = step_success_prob
        self._disc_prior_type = disc_prior_type
        self._where_mean, self._where_std = tuple(where_mean), tuple(where_std)
        self._rec_where_prior = bool(rec_where_prior)
        if rec_where_prior:
            init = {'b': list(where_mean) + list(where_std)}
            self._where_prior = RecurrentNormal(4, 128, conditional=True, output_initializers=init)
        else:
            self._where_prior = ConditionedNormalAdaptor(where_mean, where_std)

    @property
    def n_what(self):
        return self._cell.n_what


class Propagate(object):
    def __init__(self, ssm, prior):                                                 # sqair_modules.py:235
        self._ssm, self._prior = ssm, prior
        self._where_posterior = self._ssm._cell._where_distrib


class SQAIRTimestep(object):
    def __init__(self, n_steps, discover, propagate, time_cell, relation_embedding=False):   # sqair_modules.py:421
        if relation_embedding:
            raise NotImplementedError('relation_embedding=True is not in the fused kernel (default False)')
        self._n_steps, self._discover, self._propagate, self._time_cell = n_steps, discover, propagate, time_cell

    @property
    def n_what(self):
        return self._discover.n_what
