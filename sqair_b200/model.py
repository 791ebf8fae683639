"""Model: IWAE tiling, ELBO / IWAE / VIMCO objective and logging quantities around a sequence model
(reference: sqair/model.py:30-214).  The reference builds the graph once and re-runs it per
`sess.run`; here construction evaluates the model once on `obs` (filling the same attributes) and
`step(obs, seed)` re-evaluates it on a new batch -- the call bench.py times."""
import torch

from . import _capi, ops


class Target(object):
    """What `Model.make_target` hands to `opt.compute_gradients`: the value of the training target and the backward pass
    that differentiates it.  float(target) / target.value give the scalar."""

    def __init__(self, model, value, l2_reg):
        self.model, self.value, self.l2_reg = model, value, l2_reg

    def compute_gradients(self):
        return self.model.compute_gradients(l2_reg=self.l2_reg)

    def __float__(self):
        return float(self.value)


class Model(object):
    VI_TARGETS = 'iwae reinforce'.split()
    TARGETS = VI_TARGETS

    def __init__(self, obs, coords, seqence_model, k_particles, presence=None, is_training=None, debug=False,
                 seed=0, row_offset=0, device=None):
        """obs: [T,B,H,W(,1)] image sequences; coords: ground-truth boxes (unused by the path); seqence_model:
        callable like `SequentialAIR`; k_particles: IWAE particles; presence: [T,B,k] ground truth, evaluation only."""
        if device is None:
            device = obs.device if obs.is_cuda else torch.device('cuda', torch.cuda.current_device())
        self.device = device
        self.obs = obs
        self.coords = coords
        self.sequence = seqence_model
        self.k_particles = int(k_particles)
        self.gt_presence = presence
        self.debug = debug
        self.n_timesteps, self.batch_size = int(obs.shape[0]), int(obs.shape[1])
        self.img_size = tuple(obs.shape[2:])
        self.tiled_batch_size = self.batch_size * self.k_particles
        self._is_training = is_training
        self._seed, self._row_offset = seed, row_offset
        self.last_scalars = None
        self._build()

    # ------------------------------------------------------------------------------------------
    def _run(self, obs, seed, noise=None, kernel_events=None):
        obs = obs.to(self.device, non_blocking=True)
        if obs.dim() == 5:
            obs = obs[..., 0]
        out = self.sequence(obs, k_particles=self.k_particles, seed=seed, row_offset=self._row_offset, noise=noise,
                            kernel_events=kernel_events)
        obj = ops.objective(out['log_weights_per_timestep'], out['discrete_log_prob'], int(obs.shape[1]), self.k_particles)
        return obs, out, obj

    def step(self, obs, seed=0, noise=None, kernel_events=None):
        """One pass of the hot path on a batch: noise draw, SequentialAIR, particle objective.
        Returns device tensors: scalars [elbo_vae, elbo_iwae, ess, vimco_target, iwae_target, ...], log_weights [B,K]."""
        _, out, obj = self._run(obs, seed, noise, kernel_events)
        self.last_scalars = obj['scalars']
        self.outputs = out
        return dict(scalars=obj['scalars'], log_weights=obj['log_weights'], outputs=out)

    def _build(self, noise=None):
        self._noise = noise                       # explicit draws (parity tests); None = counter-based from the seed
        obs, out, obj = self._run(self.obs, self._seed, noise)
        self.outputs = out
        self.__dict__.update(out)                                                  # model.py:86
        B, K, T = self.batch_size, self.k_particles, self.n_timesteps
        sc = obj['scalars']
        self.last_scalars = sc
        self.log_weights = obj['log_weights']                                       # model.py:88-89
        self.elbo_vae = sc[_capi.OBJ_ELBO_VAE]
        self.elbo_iwae_per_example = obj['elbo_iwae_per_example']
        self.elbo_iwae = sc[_capi.OBJ_ELBO_IWAE]
        self.normalised_elbo_vae = self.elbo_vae / float(T)
        self.normalised_elbo_iwae = self.elbo_iwae / float(T)
        self.importance_weights = obj['importance_weights']                         # model.py:100
        self.ess = sc[_capi.OBJ_ESS]
        self.iw_resampling_idx = torch.multinomial(self.importance_weights, 1)[:, 0]    # model.py:102-103
        self._vimco_target, self._iwae_target = sc[_capi.OBJ_VIMCO_TARGET], sc[_capi.OBJ_IWAE_TARGET]
        for key, name in (('data_ll_per_sample', 'data_ll'), ('log_p_z_per_sample', 'log_p_z'),
                          ('log_q_z_given_x_per_sample', 'log_q_z_given_x'), ('kl_per_sample', 'kl')):
            self._log_resampled(out[key], name)
        tiled = obs.repeat_interleave(K, dim=1)                                      # logging only
        self.mse_per_sample = ((tiled - self.canvas) ** 2).mean((0, 2, 3))          # model.py:111-116
        self._log_resampled(self.mse_per_sample, 'mse')
        self.raw_mse = self.mse_per_sample.mean()
        self._log_resampled(self.num_steps_per_sample, 'num_steps')
        if self.gt_presence is not None:                                            # model.py:121-131
            self.gt_num_steps = self.gt_presence.to(self.device).float().sum(-1)
            nsp = self.num_steps_per_sample.reshape(-1, B, K)
            self.num_step_accuracy_per_example = (self.gt_num_steps[..., None] == nsp).float()
            self.raw_num_step_accuracy = self.num_step_accuracy_per_example.mean()
            self.num_step_accuracy = self._imp_weighted_mean(self.num_step_accuracy_per_example)
        for name in 'obj_id canvas glimpse presence_prob presence presence_logit where'.split():
            setattr(self, 'resampled_' + name, self.resample(getattr(self, name), axis=1))
        self._log_resampled(self.num_disc_steps_per_sample, 'num_disc_steps')
        self._log_resampled(self.num_prop_steps_per_sample, 'num_prop_steps')

    # ------------------------------------------------------------------------------------------
    def make_target(self, opt=None, n_train_itr=None, l2_reg=0.):
        """model.py:150-168: target = VIMCO(log_weights, sum_t discrete_log_prob, elbo_iwae_per_example) / T + L2 and
        `gvs = opt.compute_gradients(target)` -- one (gradient, variable) pair for EVERY trainable variable (asserted, as
        model.py:162-166 does).  The gradient comes from the backward pass of the CUDA library on the batch and draws the
        model was built on.  Without an optimiser only the value is returned (gvs = None)."""
        store = self.sequence.param_store(self.img_size[0], self.img_size[1], self.device)
        # discrete_log_prob always exists (model.py:152-154); at K = 1 the VIMCO baseline divides by K - 1 = 0
        # (targets.py:55) and the usable target is the -elbo_iwae branch (model.py:156)
        value = self._vimco_target if self.k_particles > 1 else self._iwae_target
        if l2_reg != 0.:
            value = value + l2_reg * 0.5 * (store.flat ** 2).sum()                  # targets.py:31-35
        if opt is None:
            return value, None
        target = Target(self, value, l2_reg)
        gvs = opt.compute_gradients(target)
        assert len(gvs) == len(store.table)
        for g, (name, v) in zip(gvs, gvs.names):
            assert g[0] is not None, 'Gradient for variable {} is None'.format(name)
        return target, gvs

    def compute_gradients(self, obs=None, seed=None, noise=None, l2_reg=0.):
        """Forward (with stash) + objective + backward on `obs` (default: the batch the model was built on).
        Returns GradsAndVars; `.flat_grad` is the flat gradient of THIS process's batch-mean target."""
        from .optim import GradsAndVars
        if obs is None:                           # the batch (and draws) the model was built on
            obs, noise = self.obs, (self._noise if noise is None else noise)
        obs = obs.to(self.device, non_blocking=True)
        seed = self._seed if seed is None else seed
        if int(obs.shape[1]) != self.batch_size or int(obs.shape[0]) != self.n_timesteps:
            raise ValueError('batch of shape %s does not match the model (T=%d, B=%d)' %
                             (tuple(obs.shape), self.n_timesteps, self.batch_size))
        out, obj, flat_grad = self.sequence.forward_backward(obs, k_particles=self.k_particles, noise=noise, seed=seed,
                                                             row_offset=self._row_offset)
        store = self.sequence.param_store(self.img_size[0], self.img_size[1], self.device)
        if l2_reg != 0.:
            flat_grad.add_(store.flat, alpha=float(l2_reg))
        self.outputs, self.last_scalars = out, obj['scalars']
        variables = store.variables()
        gvs = GradsAndVars()
        for name, (shape, off) in store.table.items():
            n = 1
            for d in shape:
                n *= d
            gvs.append((flat_grad[off:off + n].reshape(shape), variables[name]))
        gvs.flat_grad, gvs.store, gvs.names = flat_grad, store, list(variables.items())
        gvs.objective = obj
        return gvs

    def train_step(self, obs, opt, seed=0, global_step=None, l2_reg=0., group=None, n_global=None):
        """One iteration of the reference's training loop (scripts/experiment.py:150-155,217-218): gradients of the
        target on `obs`, data-parallel mean over the ranks of `group` (ONE all-reduce of the flat gradient, NCCL), the
        optimiser update and the re-pack of the kernel-side parameter copies.  Returns the objective scalars
        (device tensor: elbo_vae, elbo_iwae, ess, vimco_target, iwae_target)."""
        from . import parallel
        gvs = self.compute_gradients(obs, seed, l2_reg=l2_reg)
        n_global = self.batch_size * parallel.world_size(group) if n_global is None else n_global
        parallel.allreduce_flat_gradient(gvs.flat_grad, self.batch_size, n_global, group)
        opt.apply_gradients(gvs, global_step)
        return gvs.objective['scalars']

    def resample(self, *args, **kwargs):
        axis = kwargs.pop('axis', -1)
        res = [self._resample(a, axis) if self.k_particles > 1 else a for a in args]
        return res[0] if len(res) == 1 else res

    def _resample(self, arg, axis=-1):
        idx = self.iw_resampling_idx + torch.arange(self.batch_size, device=self.device) * self.k_particles
        return arg.index_select(axis if axis >= 0 else arg.dim() + axis, idx)

    def _log_resampled(self, tensor, name):
        setattr(self, 'resampled_' + name, self._resample(tensor))
        setattr(self, name, self._imp_weighted_mean(tensor))

    def _imp_weighted_mean(self, tensor):                                            # model.py:202-205
        tensor = tensor.reshape(-1, self.batch_size, self.k_particles).mean(0)
        return (self.importance_weights * tensor * self.k_particles).mean()

    def img_summaries(self):
        return (self.resampled_canvas[0].clamp(0., 1.) * 255).round().to(torch.uint8), self.obs[0]

    # ---- bench helpers -------------------------------------------------------------------------
    @property
    def cfg(self):
        return self.sequence.make_cfg(self.n_timesteps, self.batch_size, self.k_particles, self.img_size[0], self.img_size[1])

    def synthetic_obs_host(self):
        return self._host_obs


def load_synthetic_model(device, rank=0, T=10, B=32, K=5, n=4, H=50, W=50, seed=1234, row_offset=0):
    """Builds the reference's MNIST model (configs/mlp_mnist_model.py wiring, default flags) on synthetic
    moving-sprite frames of the requested shape; used by bench.py and examples."""
    import numpy as np
    from . import data
    from .configs import mlp_mnist_model as config
    from .common_model_flags import flags
    F = flags.FLAGS
    F.n_steps_per_image, F.k_particles = n, K
    imgs, nums = data.moving_sprites(T, B, H, W, n, seed=seed + rank)
    host = torch.from_numpy(imgs).pin_memory()
    mean_img = imgs.mean((0, 1))
    model = config.load(host.to(device), None, None, mean_img=mean_img)
    model._row_offset = row_offset              # global index of this shard's first row: keys the counter-based draws
    model._host_obs = host
    return model
