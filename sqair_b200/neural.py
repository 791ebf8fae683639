"""Specification objects for the feed-forward nets of the reference (sqair/neural.py:34-116).
They record widths / transfer functions; the layers execute inside the fused kernel, which
implements ELU hidden layers and {none, sigmoid} output transfers only."""


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for y in x:
            out.extend(_flatten(y))
        return out
    return [x]


class Nonlinear(object):
    def __init__(self, n_output, transfer='elu', initializers=None):
        self.n_output, self.transfer, self.initializers = int(n_output), transfer, initializers


class MLP(object):
    """MLP(n_hiddens, hidden_transfer=elu, n_out=None, transfer=None, ...) (neural.py:50-116)."""

    def __init__(self, n_hiddens, hidden_transfer='elu', n_out=None, transfer=None, initializers=None,
                 output_initializers=None, name=None):
        self.n_hiddens = [int(h) for h in _flatten(n_hiddens)]       # nest.flatten, neural.py:66
        self.hidden_transfer, self.n_out, self.transfer = hidden_transfer, n_out, transfer
        self.initializers = initializers
        self.output_initializers = initializers if output_initializers is None else output_initializers

    @property
    def output_size(self):
        return self.n_out if self.n_out is not None else self.n_hiddens[-1]
