"""Stand-ins for the Sonnet RNN cores the reference instantiates by name
(`maybe_getattr(snt, F.transition)`, configs/mlp_mnist_model.py:86-87): specification objects only;
the recurrences themselves run inside the fused CUDA kernel (SURVEY Appendix B semantics)."""


class _Core(object):
    kind = None

    def __init__(self, hidden_size):
        self.hidden_size = int(hidden_size)

    @property
    def output_size(self):
        return (self.hidden_size,)

    @property
    def state_size(self):
        return (self.hidden_size,)


class VanillaRNN(_Core):
    """out = tanh(in_to_hidden(x) + hidden_to_hidden(h))."""
    kind = 'VanillaRNN'


class GRU(_Core):
    """Sonnet GRU: reset gate applied before Uh; h' = (1 - z) h + z h~."""
    kind = 'GRU'
