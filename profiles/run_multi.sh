#!/usr/bin/env bash
# usage: bash profiles/run_multi.sh <N> <tag> [bench flags]  -- multi-GPU correctness test (N >= 2) + torchrun bench on N GPUs
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench rc $?"; grep -v "^$" gpurun_out/${tag}_bench_n$N.err | tail -4
python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench_n$N.json'))
print('value', d['value'], 'per gpu', d['value'] / d['n_gpus'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism'])
for k, v in (d.get('configs') or {}).items():
    print(k, v if not isinstance(v, dict) else (round(v['value'],1), round(v['ms_per_step'],2)))
PY
