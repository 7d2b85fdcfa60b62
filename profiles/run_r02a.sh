#!/usr/bin/env bash
# round 2, first GPU call: the whole -m gpu suite (incl. the new alias / benchmarked-shape tests), then the default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python -c "import os; print('cores', os.cpu_count())" >> gpurun_out/r02a_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02a_tests.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02a_tests.log
tail -25 gpurun_out/r02a_tests.log
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc $?"
tail -5 gpurun_out/r02a_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02a_bench.json'))
print('value', d['value'], 'e2e', d['e2e'], 'nms', d['nms'])
print(d['breakdown_ms_per_step'])
for k, v in (d.get('configs') or {}).items():
    print(k, v if not isinstance(v, dict) else (v['value'], v['ms_per_step'], v['breakdown_ms_per_step']))
print('dense', d['dense_herd'] and (d['dense_herd']['value'], d['dense_herd']['nms'], d['dense_herd'].get('nms_10k')))
print(d['cpu_baseline'])
PY
