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
