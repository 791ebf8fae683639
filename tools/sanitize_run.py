import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
import sqair_testlib as TL
from oracle import sqair_oracle as O
from sqair_b200 import ops
dev = torch.device('cuda:0')
for kw in (dict(T=2, B=4, K=1, n=2), dict(T=2, B=3, K=5, n=4), dict(T=1, B=2, K=2, n=2, H=45, W=35)):
    cfg = O.Cfg(**kw)
    imgs, params, noise = TL.make_inputs(cfg)
    ccfg = TL.capi_cfg(cfg)
    packed = ops.pack_params(ccfg, O.flatten_params(params, cfg).to(dev))
    out = ops.forward(ccfg, packed, torch.from_numpy(imgs).to(dev), {k: torch.from_numpy(v).to(dev) for k, v in noise.items()})
    obj = ops.objective(out['log_weights_per_timestep'], out['discrete_log_prob'], cfg.B, cfg.K)
    torch.cuda.synchronize()
    print(kw, float(obj['scalars'][1]))
# backward pass (stash-writing forward, objective gradient, reverse program incl. the tcgen05 weight-gradient GEMMs),
# optimiser update and the stand-alone weight-gradient entry point
for kw in (dict(T=2, B=4, K=1, n=2), dict(T=2, B=3, K=5, n=4), dict(T=2, B=2, K=2, n=2, rec_where_prior=False, masked_glimpse=False)):
    cfg = O.Cfg(**kw)
    imgs, params, noise = TL.make_inputs(cfg)
    g = TL.run_cuda_backward(cfg, imgs, params, noise)
    print('backward', kw, float(sum(np.abs(v).sum() for v in g.values())))
x = torch.randn(512, 160, device=dev); dy = torch.randn(512, 96, device=dev)
w = ops.wgrad(x, dy)
torch.cuda.synchronize()
print('wgrad (tcgen05 path) max err', float((w - x.t() @ dy).abs().max()))
p = torch.randn(1000, device=dev); gr = torch.randn(1000, device=dev); s0 = torch.ones(1000, device=dev); s1 = torch.zeros(1000, device=dev)
ops.optimizer_update(0, p, gr, s0, s1, 1e-3, .9, .9, 1e-10)
torch.cuda.synchronize()
print('optimizer ok')
