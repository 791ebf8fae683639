"""Launched under torchrun on N GPUs: a data-parallel training step (global batch split by sequences, NCCL all-reduce of
the flat gradient) must reproduce the single-rank step on the whole batch -- ELBO, gradient and updated parameters.
Noise is keyed by the global row, so the draws are identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from sqair_b200 import optim, parallel, data
from sqair_b200.common_model_flags import flags
from sqair_b200.configs import mlp_mnist_model as config

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
T, B, K, n, H, W = 4, 2 * world, 5, 4, 50, 50
F = flags.FLAGS
F.n_steps_per_image, F.k_particles = n, K
imgs, _ = data.moving_sprites(T, B, H, W, n, seed=99)
mean_img = imgs.mean((0, 1))


def run(obs_np, row_offset, group_on):
    obs = torch.from_numpy(obs_np).to(dev)
    model = config.load(obs, None, None, mean_img=mean_img)
    model._row_offset = row_offset
    store = model.sequence.param_store(H, W, dev)
    if group_on:
        parallel.broadcast_parameters(store)
    opt = optim.make_optimizer('rmsprop', 1e-3)
    gvs = model.compute_gradients(obs, seed=21)
    n_local = obs.shape[1]
    if group_on:
        parallel.allreduce_flat_gradient(gvs.flat_grad, n_local, B)
    grad = gvs.flat_grad.clone()
    before = store.flat.clone()
    opt.apply_gradients(gvs)
    means = gvs.objective['scalars'][:2].clone()
    if group_on:
        means = parallel.combine_batch_means(means, n_local)
    return grad, store.flat - before, means


start, count = parallel.shard_range(B, world, rank)
g_dp, d_dp, m_dp = run(imgs[:, start:start + count], start * K, True)
ok = True
if rank == 0:
    g_1, d_1, m_1 = run(imgs, 0, False)
    scale = float(g_1.abs().max())
    e_g = float((g_dp - g_1).abs().max()) / scale
    e_d = float((d_dp - d_1).abs().max()) / float(d_1.abs().max())
    e_m = float(((m_dp - m_1).abs() / m_1.abs()).max())
    print('ranks %d  global batch %d: max gradient error %.2e of max |g|, update error %.2e, ELBO (vae, iwae) rel error %.2e'
          % (world, B, e_g, e_d, e_m))
    ok = e_g < 2e-4 and e_d < 2e-3 and e_m < 1e-5
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
if rank == 0 and float(flag) == 1.0:
    print('DP_CHECK_OK')
dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
