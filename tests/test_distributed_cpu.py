"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding of sequences,
row offsets for the counter-based noise, and the single all-reduce that combines per-rank
objective means.  The sharded oracle must reproduce the unsharded oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sqair_testlib as TL
    from oracle import sqair_oracle as O
    from oracle import synthetic as S
    from sqair_b200 import parallel
    torch.set_num_threads(2)
    full = O.Cfg(T=2, B=5, K=2, n=2)
    imgs, params, _ = TL.make_inputs(full)
    start, count = parallel.shard_range(full.B, world, rank)
    cfg = O.Cfg(T=2, B=count, K=2, n=2)
    noise = S.philox_noise(cfg.T, cfg.rows, cfg.n, cfg.nw, 11, row_offset=parallel.row_offset(full.B, world, rank, full.K))
    out, obj = TL.run_oracle(cfg, imgs[:, start:start + count], params, noise)
    local = torch.tensor([float(obj['elbo_vae']), float(obj['elbo_iwae'])])
    combined = parallel.combine_batch_means(local, count)
    # flat-gradient all-reduce: per-rank gradients of shard means -> gradient of the global batch mean
    g = torch.full((7,), float(rank + 1))
    parallel.allreduce_flat_gradient(g, count, full.B)
    assert torch.allclose(g, torch.full((7,), (1. * 3 + 2. * 2) / 5.)), g
    q.put((rank, start, count, out['log_weights_per_timestep'], combined.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_rank():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import sqair_testlib as TL
    from oracle import sqair_oracle as O
    from oracle import synthetic as S
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full = O.Cfg(T=2, B=5, K=2, n=2)
    imgs, params, _ = TL.make_inputs(full)
    noise = S.philox_noise(full.T, full.rows, full.n, full.nw, 11)
    out, obj = TL.run_oracle(full, imgs, params, noise)
    assert [r[1:3] for r in res] == [(0, 3), (3, 2)]
    lw = np.concatenate([r[3] for r in res], 1)
    np.testing.assert_allclose(lw, out['log_weights_per_timestep'], rtol=1e-5, atol=1e-4)
    for r in res:
        np.testing.assert_allclose(r[4], [obj['elbo_vae'], obj['elbo_iwae']], rtol=1e-5)
