"""Writes tests/golden/oracle_c1.npz: outputs of the CPU oracle on BASELINE configs[0]
(T=3, B=4, K=1, n=2, 50x50; data seed 1234, weight seed 42 with 0.1 jitter, Philox noise seed 7).
Regenerate with `python tests/golden/make_oracle_fixture.py` from the repo root."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import sqair_testlib as TL          # noqa: E402
from oracle import sqair_oracle as O   # noqa: E402

cfg = O.Cfg(T=3, B=4, K=1, n=2)
imgs, params, noise = TL.make_inputs(cfg)
out, obj = TL.run_oracle(cfg, imgs, params, noise)
keep = 'presence obj_id what where canvas log_weights_per_timestep data_ll_per_sample kl_per_sample disc_prob'.split()
np.savez_compressed(os.path.join(HERE, 'oracle_c1.npz'), elbo_iwae=obj['elbo_iwae'], **{k: out[k] for k in keep})
print('wrote', os.path.join(HERE, 'oracle_c1.npz'))
