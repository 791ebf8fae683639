"""Times the persistent sequence kernel for each rows-per-block choice (tuning aid, GPU box only)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sqair_b200 import ops, _capi

def main():
    T, B, K, n, H = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (10, 32, 5, 4, 50))]
    dev = torch.device('cuda:0')
    cfg = _capi.make_cfg(T, B, K, n, H, H)
    sizes = _capi.query_sizes(cfg)
    g = torch.Generator(device='cpu').manual_seed(0)
    flat = (torch.randn(sizes.param_count, generator=g) * 0.05).to(dev)
    packed = ops.pack_params(cfg, flat)
    obs = torch.rand(T, B, H, H, generator=g).to(dev)
    noise = ops.fill_noise(cfg, 7, 0, device=dev)
    outs = ops.alloc_outputs(cfg, dev)
    res = {}
    shapes = [tuple(int(v) for v in x.split(':')) for x in os.environ.get('SWEEP', '4:3,5:4,6:5,6:4,4:2,3:2,2:1,6:2,5:3,4:4' if len(sys.argv) <= 5 else '0:0').split(',')]
    for R, C in shapes:
        os.environ.pop('SQAIR_ROWS_PER_CTA', None); os.environ.pop('SQAIR_CLUSTER', None)
        if R:
            os.environ['SQAIR_ROWS_PER_CTA'] = str(R)
            os.environ['SQAIR_CLUSTER'] = str(C)
        try:
            s = _capi.query_sizes(cfg)
            packed = ops.pack_params(cfg, flat)
        except Exception as e:
            print('R=%d C=%d: %s' % (R, C, e)); continue
        for _ in range(3):
            ops.forward(cfg, packed, obs, noise, outs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.forward(cfg, packed, obs, noise, outs)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res['%d:%d' % (R, C)] = ms
        chk = ' '.join('%s %.6f' % (k, outs[k].double().sum().item()) for k in ('log_weights_per_timestep', 'canvas', 'presence', 'what'))
        print('R=%d C=%d ctas=%d smem=%d B: %.3f ms/step  %.0f frames/s | %s' % (s.rows_per_cta, s.cluster_size, s.n_ctas, s.smem_bytes, ms, B * T / ms * 1e3, chk), flush=True)
    os.environ.pop('SQAIR_ROWS_PER_CTA', None); os.environ.pop('SQAIR_CLUSTER', None)
    print(json.dumps(res))

if __name__ == '__main__':
    main()
