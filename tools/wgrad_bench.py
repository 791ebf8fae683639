import sys, torch
sys.path.insert(0, '.')
from sqair_b200 import ops
dev = torch.device('cuda:0')
for M, K, N in [(6400, 672, 256), (6400, 632, 512), (6400, 264, 256), (1600, 2504, 256)]:
    x = torch.randn(M, K, device=dev); dy = torch.randn(M, N, device=dev)
    for _ in range(3): ops.wgrad(x, dy)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.wgrad(x, dy)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print('wgrad M=%d K=%d N=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent)' % (M, K, N, ms, 2.0 * M * K * N / ms / 1e9))
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3): x.t() @ dy
    torch.cuda.synchronize(); a.record()
    for _ in range(20): x.t() @ dy
    b.record(); torch.cuda.synchronize()
    print('   cuBLAS fp32 (torch): %.3f ms' % (a.elapsed_time(b) / 20))
