"""Optimisation targets and control variates: the pure functions of sqair/targets.py:31-89 on
`[..., K]` device tensors (torch element-wise plumbing for API parity; `Model` itself uses the fused
particle-objective kernel `sqair_objective`)."""
import math

import torch


def l2_reg(weight, params=None):
    if weight == 0.:
        return 0.
    return weight * sum(0.5 * (p ** 2).sum() for p in (params or []))


def iwae(log_weights):
    k = log_weights.shape[-1]
    return torch.logsumexp(log_weights, -1) - math.log(float(k))


def vimco_control_variate(target_per_particle):
    k = int(target_per_particle.shape[-1])
    summed = target_per_particle.sum(-1, keepdim=True)
    all_but_one = (summed - target_per_particle) / (k - 1.)            # NaN at K = 1, as in the reference
    baseline = target_per_particle[..., None] + torch.diag_embed(all_but_one - target_per_particle)
    return torch.logsumexp(baseline, -2) - math.log(float(k))


def vimco(log_weights, log_probs, elbo_iwae=None):
    signal = (log_weights - vimco_control_variate(log_weights)).detach()
    log_probs = log_probs.reshape(log_weights.shape)
    if elbo_iwae is None:
        elbo_iwae = iwae(log_weights)
    return (-elbo_iwae[..., None] - signal * log_probs).mean()


def reinforce(log_weights, log_probs, elbo_iwae=None):
    signal = log_weights.detach()
    log_probs = log_probs.reshape(log_weights.shape)
    if elbo_iwae is None:
        elbo_iwae = iwae(log_weights)
    return (-elbo_iwae[..., None] - signal * log_probs).mean()
