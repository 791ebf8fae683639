"""State-space model and priors for propagation: surface of sqair/propagate.py:35-184."""


class PropagatePrior(object):
    name = 'rnn'

    def __init__(self, n_what, cell, prop_logit_bias, where_loc_bias=None):       # propagate.py:54
        if where_loc_bias is not None:
            raise NotImplementedError('where_loc_bias is not wired into the fused kernel')
        self._n_what, self._cell, self._prop_logit_bias = n_what, cell, prop_logit_bias


class RandomWalkPropagatePrior(PropagatePrior):
    name = 'rw'


class GuidedWalkPropagatePrior(PropagatePrior):
    name = 'guided'


def make_prior(name, *args, **kwargs):
    prior_map = {'rnn': PropagatePrior, 'rw': RandomWalkPropagatePrior, 'guided': GuidedWalkPropagatePrior}
    if name not in prior_map:                                                       # propagate.py:42-43
        raise ValueError('Invalid prior type: "{}". Choose from {}.'.format(name, list(prior_map.keys())))
    return prior_map[name](*args, **kwargs)


class SequentialSSM(object):
    def __init__(self, cell):                                                       # propagate.py:164
        self._cell = cell
