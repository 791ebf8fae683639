#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): training tests incl. the NCCL invariance check, bench at 1 and N ranks.
TAG=${1:-mg}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/${TAG}_pytest_train.log 2>&1; echo "train tests rc=$?"
tail -15 gpurun_out/${TAG}_pytest_train.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tee gpurun_out/${TAG}_dp_check_n$N.txt | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.loads(open('gpurun_out/${TAG}_bench_n%d.json' % n).read().strip().splitlines()[-1])
        print(n, 'infer %.0f f/s (%.2f ms)' % (d['value'], d['ms_per_step']), '| train', {k: d['train'][k] for k in ('value', 'ms_per_step', 'allreduce_ms')}, '| strong', {k: d['train_strong'][k] for k in ('value', 'ms_per_step', 'allreduce_ms', 'sequences_per_gpu')})
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 gpurun_out/${TAG}_bench_n$N.err
