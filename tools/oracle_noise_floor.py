"""fp32 noise floor of the reference arithmetic itself: the oracle (torch-CPU restatement) evaluated in float32 against
the same restatement in float64 on identical inputs, per output.  The GPU parity tolerances (tests/sqair_testlib.py) are
read against these numbers: an fp32 implementation cannot agree with another fp32 implementation better than each agrees
with the float64 value.  Presence decisions (u < sigmoid(logit)) can flip between precisions; frames after a flip are
excluded per row (reported as `rows kept`).  CPU only.  Usage: oracle_noise_floor.py [c2|c4|small]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch

import sqair_testlib as TL
from oracle import sqair_oracle as O

which = sys.argv[1] if len(sys.argv) > 1 else 'c2'
cfg = dict(c2=O.Cfg(T=10, B=32, K=5, n=4), c4=O.Cfg(T=10, B=16, K=10, n=6, H=100, W=100), small=O.Cfg(T=4, B=3, K=5, n=4))[which]
imgs, params, noise = TL.make_inputs(cfg)
w32, o32 = TL.run_oracle(cfg, imgs, params, noise)
torch.set_default_dtype(torch.float64)
with torch.no_grad():
    out, obj = O.model_forward({k: v.double() for k, v in params.items()}, cfg, torch.from_numpy(imgs).double(),
                               {k: torch.from_numpy(v).double() for k, v in noise.items()})
torch.set_default_dtype(torch.float32)
w64 = {k: v.numpy() for k, v in out.items()}
o64 = {k: v.numpy() for k, v in obj.items()}
T, rows = cfg.T, cfg.B * cfg.K
# rows whose integer decisions agree in every frame
same = np.ones(rows, bool)
for k in TL.EXACT:
    same &= (w32[k].reshape(T, rows, -1) == w64[k].reshape(T, rows, -1)).all((0, 2))
print('%s: T=%d B=%d K=%d n=%d %dx%d; rows kept %d of %d (identical presence / ids in both precisions)'
      % (which, cfg.T, cfg.B, cfg.K, cfg.n, cfg.H, cfg.W, int(same.sum()), rows))
print('%-34s %12s %12s %12s   %s' % ('output', 'max|d|', 'max|d|/|v|', 'max|v|', 'test tolerance'))
px = cfg.H * cfg.W
for k in w32:
    if k in TL.EXACT:
        continue
    a = w32[k].reshape(T, rows, -1)[:, same].astype(np.float64)
    b = w64[k].reshape(T, rows, -1)[:, same]
    d = np.abs(a - b)
    rel = (d / np.maximum(np.abs(b), 1e-30))[np.abs(b) > 1e-3]
    tol = 'atol %.3g + 1e-4 rel' % (TL.PIXEL_SUM_ATOL * px) if k in TL.PIXEL_SUMS else ('99.995% at 1e-4, all 1e-3' if k == 'canvas' else '1e-4 + 1e-4 rel')
    print('%-34s %12.3e %12.3e %12.3e   %s' % (k, d.max(), rel.max() if rel.size else 0.0, np.abs(b).max(), tol))
if k:
    c = np.abs(w32['canvas'].reshape(T, rows, -1)[:, same].astype(np.float64) - w64['canvas'].reshape(T, rows, -1)[:, same])
    print('canvas: fraction of pixels with |d| > 1e-4 + 1e-4|v|: %.3e; > 1e-5: %.3e' % (float((c > 1e-4 + 1e-4 * np.abs(w64['canvas'].reshape(T, rows, -1)[:, same])).mean()), float((c > 1e-5).mean())))
for k in ('elbo_vae', 'elbo_iwae', 'ess', 'vimco_target', 'iwae_target'):
    if k in o64:
        print('objective %-14s fp32 %.7g  fp64 %.7g  |d| %.3e  rel %.3e%s' % (k, float(o32[k]), float(o64[k]), abs(float(o32[k]) - float(o64[k])), abs(float(o32[k]) - float(o64[k])) / abs(float(o64[k])),
              '' if same.all() else '   (includes rows whose decisions differ)'))
