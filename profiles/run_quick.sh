#!/usr/bin/env bash
# usage: bash profiles/run_quick.sh <tag> [pytest -k expr]   -- kernel tests (subset), bench with breakdown, GEMM shape table
tag=$1; kexpr=${2:-"flash or gemm"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "$kexpr" > gpurun_out/${tag}_ktests.log 2>&1 || { tail -30 gpurun_out/${tag}_ktests.log; exit 1; }
tail -2 gpurun_out/${tag}_ktests.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/${tag}_mtests.log 2>&1 || { tail -30 gpurun_out/${tag}_mtests.log; exit 1; }
tail -2 gpurun_out/${tag}_mtests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_${tag}.json > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}.json')); print(d['value'], d['e2e']['value'], d['clocks']); print(d['breakdown_ms_per_step'])"; tail -3 gpurun_out/bench_${tag}.err
if [ -z "$NO_SHAPES" ]; then timeout 600 python profiles/gemm_shapes.py 32 > gpurun_out/gemm_shapes_${tag}.jsonl 2> gpurun_out/gemm_shapes_${tag}.err; python -c "
import json
for l in open('gpurun_out/gemm_shapes_${tag}.jsonl'):
    r=json.loads(l); print('%-18s pair %.4f ms %6.0f TF | 1cta %.4f ms %6.0f TF | cublas %.4f ms %6.0f TF | hbm %.4f' % (r['name'], r.get('wm_bn512_ms',0), r.get('wm_bn512_tflops',0), r.get('wm_bn256_ms',0), r.get('wm_bn256_tflops',0), r['cublas_ms'], r['cublas_tflops'], r['min_hbm_ms']))"; fi
