#!/bin/bash
# 8-GPU evidence (short): NCCL invariance check at 8 ranks, bench.py at 4 and 8 ranks
TAG=${1:-sc}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 tools/dp_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tee gpurun_out/${TAG}_dp_check_n8.txt | tail -3
for N in 4 8; do
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
done
python - <<PY
import json
for n in (4, 8):
    try:
        d = json.loads([l for l in open('gpurun_out/${TAG}_bench_n%d.json' % n) if l.startswith('{')][-1])
        print(n, 'infer %.0f f/s (%.2f ms) e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), '| train %.0f f/s %.2f ms ar %.3f ms' % (d['train']['value'], d['train']['ms_per_step'], d['train']['allreduce_ms']), '| strong %.0f f/s %.2f ms ar %.3f ms, %d seq/GPU' % (d['train_strong']['value'], d['train_strong']['ms_per_step'], d['train_strong']['allreduce_ms'], d['train_strong']['sequences_per_gpu']))
    except Exception as e:
        print(n, 'failed', e)
PY
