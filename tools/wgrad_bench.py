"""Throughput of the weight-gradient GEMM (C ABI: sqair_wgrad) at the layer shapes of BASELINE configs[1]: back-to-back
eager calls (host enqueue included), the same calls replayed from a CUDA graph (device time only), and torch's cuBLAS
fp32 product for comparison.  "fp32-equivalent" = 2 M K N / time for a result of fp32 accuracy (3xTF32 inside).
GPU box only.  SQAIR_NO_TC=1 selects the mma.sync kernel."""
import sys, torch
sys.path.insert(0, '.')
from sqair_b200 import ops
dev = torch.device('cuda:0')
REP = 20


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


for M, K, N in [(6400, 672, 256), (6400, 632, 512), (6400, 264, 256), (6400, 256, 256), (1600, 2504, 256)]:
    x = torch.randn(M, K, device=dev); dy = torch.randn(M, N, device=dev)
    out = torch.empty(K, N, device=dev)
    for _ in range(3): ops.wgrad(x, dy)
    eager = timed(lambda: [ops.wgrad(x, dy) for _ in range(REP)]) / REP
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out.zero_(); ops.wgrad(x, dy, out=out); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REP):
                out.zero_(); ops.wgrad(x, dy, out=out)         # what the eager call does: clear, then accumulate
    g.replay()
    graph = timed(g.replay) / REP
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3): x.t() @ dy
    blas = timed(lambda: [x.t() @ dy for _ in range(REP)]) / REP
    f = 2.0 * M * K * N / 1e9
    print('wgrad M=%d K=%d N=%d: eager %.3f ms %.1f TFLOP/s | CUDA graph %.3f ms %.1f TFLOP/s (fp32-equivalent) | cuBLAS fp32 %.3f ms %.1f TFLOP/s'
          % (M, K, N, eager, f / eager, graph, f / graph, blas, f / blas))
