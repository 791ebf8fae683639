"""Times the training-step pieces on one GPU: forward (with stash), objective gradient, backward.  GPU box only.
Usage: python tools/train_step_time.py [T B K n [H W]]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sqair_b200 import ops, _capi
from oracle import sqair_oracle as O          # inputs only (weights / frames); nothing here is timed on the CPU
from oracle import synthetic as S

a = [int(x) for x in sys.argv[1:]]
T, B, K, n = (a + [10, 32, 5, 4][len(a):])[:4]
H, W = (a[4], a[5]) if len(a) >= 6 else (50, 50)
cfg = O.Cfg(T=T, B=B, K=K, n=n, H=H, W=W)
dev = torch.device('cuda:0')
imgs, _ = S.make_sequences(T, B, H, W, n, seed=1234)
params = O.init_params(cfg, 42, mean_img=imgs.mean((0, 1)), jitter=0.1)
ccfg = _capi.make_cfg(T, B, K, n, H, W)
flat = O.flatten_params(params, cfg).to(dev)
obs = torch.from_numpy(imgs).to(dev)
packed = ops.pack_params(ccfg, flat)
bw = ops.pack_backward(ccfg, flat)
ts = _capi.query_train_sizes(ccfg)
print('stash %.1f MB, workspace %.1f MB, backward params %.1f MB' % (ts.stash_floats * 4e-6, ts.workspace_floats * 4e-6, ts.backward_param_floats * 4e-6))
stash = torch.empty(ts.stash_floats, dtype=torch.float32, device=dev)
wsb = torch.empty(ts.workspace_floats, dtype=torch.float32, device=dev)
noise = ops.fill_noise(ccfg, 7, device=dev)
outs = ops.alloc_outputs(ccfg, dev)
dp = torch.empty_like(flat)


def step():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    ops.fill_noise(ccfg, 7, noise=noise)
    ev[1].record()
    ops.forward(ccfg, packed, obs, noise, outputs=outs, stash=stash)
    ev[2].record()
    d_lw, d_lp = ops.objective_grad(outs['log_weights_per_timestep'], outs['discrete_log_prob'], B, K)
    ev[3].record()
    _, launches = ops.backward(ccfg, flat, bw, obs, noise, stash, d_lw, d_lp if K > 1 else None, workspace=wsb, d_params=dp)
    ev[4].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], launches


for _ in range(3):
    step()
acc = [0.0] * 4
N = 10
t0 = time.time()
for _ in range(N):
    t, launches = step()
    acc = [x + y for x, y in zip(acc, t)]
wall = (time.time() - t0) / N * 1e3
acc = [x / N for x in acc]
print('noise %.3f ms, forward+stash %.3f ms, objective grad %.3f ms, backward %.3f ms (%d launches); step %.3f ms device, %.3f ms wall -> %.0f frames/s'
      % (acc[0], acc[1], acc[2], acc[3], launches, sum(acc), wall, T * B / (max(sum(acc), wall) * 1e-3)))
ops.forward(ccfg, packed, obs, noise, outputs=outs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    ops.forward(ccfg, packed, obs, noise, outputs=outs)
e1.record(); torch.cuda.synchronize()
print('inference forward %.3f ms' % (e0.elapsed_time(e1) / N))
print('grad norm %.4e, finite %s' % (float(dp.norm()), bool(torch.isfinite(dp).all())))
# ---- where does the backward time go: host enqueue vs device (CUDA graph replay removes the host from the picture)
d_lw, d_lp = ops.objective_grad(outs['log_weights_per_timestep'], outs['discrete_log_prob'], B, K)
torch.cuda.synchronize()
t0 = time.perf_counter()
ops.backward(ccfg, flat, bw, obs, noise, stash, d_lw, d_lp if K > 1 else None, workspace=wsb, d_params=dp)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('backward: host enqueue %.3f ms, then %.3f ms until the device is done' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
ref = dp.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    ops.backward(ccfg, flat, bw, obs, noise, stash, d_lw, d_lp if K > 1 else None, workspace=wsb, d_params=dp)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        ops.backward(ccfg, flat, bw, obs, noise, stash, d_lw, d_lp if K > 1 else None, workspace=wsb, d_params=dp)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    g.replay()
e1.record(); torch.cuda.synchronize()
print('backward as a CUDA graph: %.3f ms per replay; max |diff| vs eager %.3e (atomics reorder)' % (e0.elapsed_time(e1) / N, float((dp - ref).abs().max())))
