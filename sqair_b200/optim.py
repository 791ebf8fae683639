"""Optimisers and the learning-rate schedule of the reference's training script (sqair/scripts/experiment.py:126-155),
TensorFlow 1.x update rules, as ONE fused kernel over the flat canonical parameter buffer (`sqair_optimizer_update`).

The call protocol mirrors the reference: `target, gvs = model.make_target(opt, ...)` asks the optimiser for the
gradients (`opt.compute_gradients(target)`; here the target object carries the backward pass of the CUDA library instead
of a TF graph) and `opt.apply_gradients(gvs, global_step)` applies them."""
import numpy as np
import torch

from . import _capi, ops


def piecewise_constant(step, boundaries, values):
    """tf.train.piecewise_constant: values[0] for step <= boundaries[0], values[i] for boundaries[i-1] < step <= boundaries[i]."""
    if len(values) != len(boundaries) + 1:
        raise ValueError('The length of boundaries should be 1 less than the length of values')
    for b, v in zip(boundaries, values):
        if step <= b:
            return float(v)
    return float(values[-1])


def make_schedule(learning_rate, schedule, train_itr):
    """scripts/experiment.py:128-136: `schedule` = comma-separated relative stage lengths (default '4,6,10'); the rate is
    divided by 3 at the cumulative boundaries scaled to `train_itr`.  Returns lr(global_step)."""
    if not schedule:
        return lambda step: float(learning_rate)
    stages = np.cumsum([float(f) for f in str(schedule).split(',')])
    stages = stages * train_itr / stages[-1]
    bounds = [int(b) for b in np.round(stages).astype(np.int32)]
    lrs = [float(v) for v in learning_rate * (1. / 3) ** np.arange(len(bounds))]
    return lambda step: piecewise_constant(int(step), bounds[:-1], lrs)


class GradsAndVars(list):
    """[(gradient, variable)] in `sqair_param_layout` order (views into two flat buffers, kept as attributes)."""
    flat_grad = None
    store = None
    names = ()


class Optimizer(object):
    kind, n_slots, slot0_init = _capi.OPT_SGD, 0, 0.

    def __init__(self, learning_rate):
        self._lr = learning_rate                      # float or callable(global_step)
        self._slots = {}
        self.global_step = 0

    def learning_rate(self, step=None):
        step = self.global_step if step is None else step
        return float(self._lr(step)) if callable(self._lr) else float(self._lr)

    def compute_gradients(self, target):
        """`target`: what Model.make_target builds (value + the backward pass that differentiates it)."""
        return target.compute_gradients()

    def _hyper(self, lr):
        return lr, 0., 0., 0.

    def _get_slots(self, store):
        key = id(store)
        if key not in self._slots:
            s = [torch.full_like(store.flat, self.slot0_init if i == 0 else 0.) for i in range(self.n_slots)]
            self._slots[key] = (s + [None, None])[:2]
        return self._slots[key]

    def apply_gradients(self, gvs, global_step=None, grad_scale=1., l2_weight=0.):
        """In-place update of the parameter store behind `gvs`; the kernel-side copies are re-packed lazily."""
        if not isinstance(gvs, GradsAndVars):
            raise TypeError('apply_gradients expects the gradient list returned by Model.make_target')
        store = gvs.store
        if global_step is not None:
            self.global_step = int(global_step)
        lr, a, b, eps = self._hyper(self.learning_rate(self.global_step))
        s0, s1 = self._get_slots(store)
        ops.optimizer_update(self.kind, store.flat, gvs.flat_grad, s0, s1, lr, a, b, eps, grad_scale, l2_weight)
        store.mark_dirty()
        self.global_step += 1
        return self.global_step

    def state_dict(self, store):
        s0, s1 = self._get_slots(store)
        return dict(global_step=self.global_step, slot0=s0, slot1=s1)


class GradientDescentOptimizer(Optimizer):
    pass


class MomentumOptimizer(Optimizer):
    kind, n_slots = _capi.OPT_MOMENTUM, 1

    def __init__(self, learning_rate, momentum=.9):
        super(MomentumOptimizer, self).__init__(learning_rate)
        self.momentum = momentum

    def _hyper(self, lr):
        return lr, self.momentum, 0., 0.


class RMSPropOptimizer(Optimizer):
    """tf.train.RMSPropOptimizer(lr, decay=0.9, momentum=0.0, epsilon=1e-10); the reference passes momentum=.9
    (scripts/experiment.py:140).  The mean-square slot starts at one, the momentum slot at zero (TF slot initialisers)."""
    kind, n_slots, slot0_init = _capi.OPT_RMSPROP, 2, 1.

    def __init__(self, learning_rate, decay=.9, momentum=0., epsilon=1e-10):
        super(RMSPropOptimizer, self).__init__(learning_rate)
        self.decay, self.momentum, self.epsilon = decay, momentum, epsilon

    def _hyper(self, lr):
        return lr, self.decay, self.momentum, self.epsilon


class AdamOptimizer(Optimizer):
    kind, n_slots = _capi.OPT_ADAM, 2

    def __init__(self, learning_rate, beta1=.9, beta2=.999, epsilon=1e-8):
        super(AdamOptimizer, self).__init__(learning_rate)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon

    def _hyper(self, lr):
        t = self.global_step + 1
        return lr * np.sqrt(1. - self.beta2 ** t) / (1. - self.beta1 ** t), self.beta1, self.beta2, self.epsilon


def make_optimizer(name, learning_rate):
    """scripts/experiment.py:138-146."""
    name = name.lower()
    if name == 'rmsprop':
        return RMSPropOptimizer(learning_rate, momentum=.9)
    if name == 'adam':
        return AdamOptimizer(learning_rate)
    if name == 'sgd':
        return GradientDescentOptimizer(learning_rate)
    if name == 'momentum':
        return MomentumOptimizer(learning_rate, momentum=.9)
    raise ValueError('unknown optimiser "%s"' % name)
