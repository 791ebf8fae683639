"""SequentialAIR: the operator this package accelerates (reference: sqair/seq.py:41-279).

The reference unrolls SQAIRTimestep in a `tf.while_loop` writing 38 TensorArrays; here the same
recursion is ONE persistent CUDA kernel launch (`sqair_forward`) per call.  The constructor keeps the
reference signature and validates that the wired modules form the architecture the fused kernel
implements (the reference's MNIST config and its flag-level variants)."""
import torch

from . import _capi, ops
from .params import ParamStore
from .sqair_modules import SQAIRTimestep


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class SequentialAIR(object):
    def __init__(self, max_steps, glimpse_size, discover, propagate, time_cell, decoder,
                 sample_from_prior=False, generate_after=-1):
        # seq.py:63-64,198-203: with sample_from_prior the posterior is evaluated at draws from the propagation prior and, in
        # frames t > generate_after (if > 0), those draws replace the latents (conditional generation)
        self._sample_from_prior, self._generate_after = bool(sample_from_prior), int(generate_after)
        self._max_steps, self._glimpse_size, self._decoder = max_steps, tuple(glimpse_size), decoder
        self._sqair = SQAIRTimestep(self._max_steps, discover, propagate, time_cell)
        self._discover, self._propagate = discover, propagate
        self._spec = self._read_architecture()
        self._stores = {}
        self._train_bufs = {}
        self._infer_bufs = {}

    # ---- architecture -> kernel configuration ----------------------------------------------------
    def _read_architecture(self):
        d, p = self._discover._cell, self._propagate._ssm._cell
        nh = d._n_hidden
        def need(cond, what):
            if not cond:
                raise NotImplementedError('fused SQAIR kernel: unsupported architecture: ' + what)
        need(d._cell.kind == 'VanillaRNN' and p._cell.kind == 'VanillaRNN', 'transition must be VanillaRNN')
        need(p._temporal_cell.kind == 'GRU' and self._propagate._prior._cell.kind == 'GRU',
             'time_transition / prior_transition must be GRU')
        need(p._n_hidden == nh and p._temporal_cell.hidden_size == nh and self._propagate._prior._cell.hidden_size == nh,
             'all recurrent cores must share one width')
        need(p._glimpse_encoder is d._glimpse_encoder, 'propagation must share discovery\'s glimpse encoder')
        enc = d._glimpse_encoder
        need(enc.glimpse_encoder.n_hidden == [nh, nh] and d._input_encoder.n_hidden == [nh, nh], 'encoders must be [nh, nh]')
        for te in (d._transform_estimator, p._transform_estimator):
            need(te.n_hidden == [nh, nh], 'transform estimators must be [nh, nh]')
        for sp in (d._steps_predictor, p._steps_predictor):
            need(sp.n_hidden == [nh // 2], 'steps predictors must be [nh/2]')
        dec = self._decoder.glimpse_decoder
        need(dec.n_hidden == [nh, nh] and tuple(dec.output_size) == self._glimpse_size, 'decoder must be [nh, nh] -> glimpse')
        need(self._glimpse_size[0] == self._glimpse_size[1], 'square glimpses only')
        return dict(n_hidden=nh, n_what=d.n_what, masked_glimpse=enc.masked_glimpse,
                    prior_type=self._propagate._prior.name, prop_prior_step_bias=self._propagate._prior._prop_logit_bias,
                    disc_prior_type=self._discover._disc_prior_type, rec_where_prior=self._discover._rec_where_prior,
                    step_success_prob=self._discover._init_disc_step_success_prob,
                    where_mean=self._discover._where_mean, where_std=self._discover._where_std,
                    where_update_scale=p._where_update_scale, output_std=self._decoder.output_std,
                    bg_std=self._decoder.bg_std,
                    init=dict(output_scale=dec.output_scale, disc_scale_offset=d._transform_estimator.scale_offset,
                              prop_scale_offset=p._transform_estimator.scale_offset,
                              disc_step_bias=d._steps_predictor.steps_bias, prop_step_bias=p._steps_predictor.steps_bias,
                              where_mean=self._discover._where_mean, where_std=self._discover._where_std,
                              mean_img=self._decoder.mean_img))

    def make_cfg(self, T, B, K, H, W):
        s = self._spec
        return _capi.make_cfg(T, B, K, self._max_steps, H, W, self._glimpse_size[0], s['n_what'], s['n_hidden'],
                              s['prior_type'], s['disc_prior_type'], s['rec_where_prior'], s['masked_glimpse'],
                              s['step_success_prob'], s['prop_prior_step_bias'], s['output_std'], s['bg_std'],
                              s['where_update_scale'], 1e-2, s['where_mean'], s['where_std'])

    def param_store(self, H, W, device, seed=42):
        """Variables depend on the canvas size only; one store per (H, W, device)."""
        key = (H, W, str(device))
        if key not in self._stores:
            self._stores[key] = ParamStore(self.make_cfg(1, 1, 1, H, W), device, seed, self._spec['init'])
        return self._stores[key]

    # ---- the operator ------------------------------------------------------------------------------
    def __call__(self, obs, coords=None, sample_from_prior=False, k_particles=1, noise=None, seed=0,
                 row_offset=0, outputs=None, kernel_events=None):
        """obs: [T,B,H,W] or [T,B,H,W,1] float32 CUDA tensor.  With k_particles = K > 1 the K particles of a
        sequence share its frames (virtual `tile_input_for_iwae`); outputs are [T, B*K, ...] as in the reference.
        `noise` (eps_where / eps_what / u_pres tensors) fixes every random draw; otherwise they are generated
        on the device from `seed` with rows keyed by `row_offset + row`.  (The `sample_from_prior` argument of the
        reference's `_build` is never read there, seq.py:69-84; generation is switched on in the constructor.)"""
        if obs.dim() == 5:
            if obs.shape[-1] != 1:
                raise NotImplementedError('multi-channel frames')
            obs = obs[..., 0]
        obs = obs.contiguous()
        T, B, H, W = obs.shape
        cfg = self.make_cfg(T, B, k_particles, H, W)
        store = self.param_store(H, W, obs.device)
        # the 38 output tensors and the noise tensors of a call shape are allocated once and REUSED: what a call returns
        # is overwritten by the next call with the same shape (pass `outputs=` to keep results)
        key = (T, B, k_particles, H, W, str(obs.device))
        buf = self._infer_bufs.get(key)
        if buf is None:
            buf = self._infer_bufs[key] = dict(noise=ops.alloc_noise(cfg, obs.device), outputs=ops.alloc_outputs(cfg, obs.device))
        if noise is None:
            noise = ops.fill_noise(cfg, seed, row_offset, noise=buf['noise'])
        if outputs is None:
            outputs = buf['outputs']
        prior_noise = None
        if self._sample_from_prior:                       # second noise set: the draws from the priors
            if all(k + '_prior' in noise for k in ('eps_where', 'eps_what', 'u_pres')):
                prior_noise = {k: noise[k + '_prior'] for k in ('eps_where', 'eps_what', 'u_pres')}
            else:
                if 'prior_noise' not in buf:
                    buf['prior_noise'] = ops.alloc_noise(cfg, obs.device)
                prior_noise = ops.fill_noise(cfg, (seed * 0x9E3779B1 + 0x7F4A7C15) & (2 ** 63 - 1), row_offset, noise=buf['prior_noise'])
        if kernel_events is not None:
            kernel_events[0].record()
        out = ops.forward(cfg, store.packed(cfg), obs, noise, outputs, prior_noise=prior_noise, generate_after=self._generate_after)
        if kernel_events is not None:
            kernel_events[1].record()
        return AttrDict(out)

    # ---- the operator with its gradient (training) -------------------------------------------------
    def _train_buffers(self, cfg, device):
        """Stash / workspace / gradient / noise / output buffers of one call shape, allocated once."""
        key = (cfg.T, cfg.B, cfg.K, cfg.H, cfg.W, str(device))
        if key not in self._train_bufs:
            ts = _capi.query_train_sizes(cfg)
            store = self.param_store(cfg.H, cfg.W, device)
            self._train_bufs[key] = AttrDict(
                stash=torch.empty(ts.stash_floats, dtype=torch.float32, device=device),
                workspace=torch.empty(ts.workspace_floats, dtype=torch.float32, device=device),
                d_params=torch.empty_like(store.flat), noise=ops.alloc_noise(cfg, device),
                outputs=ops.alloc_outputs(cfg, device))
        return self._train_bufs[key]

    def forward_backward(self, obs, k_particles=1, noise=None, seed=0, row_offset=0, vimco=None, use_graph=True):
        """Forward pass that keeps what the adjoint needs, particle objective, and the gradient of the training target
        (VIMCO / T when K > 1, else -elbo_iwae / T: model.py:150-158) w.r.t. every variable -- the work of
        `opt.compute_gradients(target)` (model.py:160).  Returns (outputs, objective dict, flat gradient in
        `sqair_param_layout` order).  The returned tensors live in per-shape buffers that the next call overwrites.
        The backward pass (one persistent reverse-program kernel + ~120 weight-gradient launches on four streams) is
        replayed from a CUDA graph after the first call of a shape (all its operands live in the per-shape buffers, so
        the graph is static; the first, eager call also uploads the program table the graph then refers to)."""
        if obs.dim() == 5:
            if obs.shape[-1] != 1:
                raise NotImplementedError('multi-channel frames')
            obs = obs[..., 0]
        T, B, H, W = obs.shape
        cfg = self.make_cfg(T, B, k_particles, H, W)
        store = self.param_store(H, W, obs.device)
        buf = self._train_buffers(cfg, obs.device)
        vimco = k_particles > 1 if vimco is None else vimco
        if 'obs' not in buf:
            buf.obs = torch.empty(T, B, H, W, dtype=torch.float32, device=obs.device)
            buf.d_lw = torch.empty(B, k_particles, dtype=torch.float32, device=obs.device)
            buf.d_lp = torch.empty(B, k_particles, dtype=torch.float32, device=obs.device)
            buf.graphs = {}
        buf.obs.copy_(obs, non_blocking=True)              # (also the host-to-device copy when `obs` is pinned host memory)
        if noise is None:
            ops.fill_noise(cfg, seed, row_offset, noise=buf.noise)
        else:
            for k in buf.noise:
                buf.noise[k].copy_(noise[k], non_blocking=True)
        out = ops.forward(cfg, store.packed(cfg), buf.obs, buf.noise, buf.outputs, stash=buf.stash)
        lw, lp = out['log_weights_per_timestep'], out['discrete_log_prob']
        obj = ops.objective(lw, lp, B, k_particles)
        ops.objective_grad(lw, lp, B, k_particles, out=(buf.d_lw, buf.d_lp))
        bwp = store.backward_params(cfg)

        def run_backward():
            ops.backward(cfg, store.flat, bwp, buf.obs, buf.noise, buf.stash, buf.d_lw, buf.d_lp if vimco else None,
                         workspace=buf.workspace, d_params=buf.d_params)

        key = (bool(vimco), bwp.data_ptr(), store.flat.data_ptr())
        state = buf.graphs.get(key)
        if not use_graph or torch.cuda.is_current_stream_capturing():
            run_backward()
        elif state is None:
            run_backward()                                  # first call of this shape: eager (also warms every cache)
            buf.graphs[key] = 'warm'
        elif state == 'warm':
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            side = torch.cuda.Stream(device=obs.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    run_backward()
            torch.cuda.current_stream().wait_stream(side)
            buf.graphs[key] = g
            g.replay()
        else:
            state.replay()
        return AttrDict(out), obj, buf.d_params
