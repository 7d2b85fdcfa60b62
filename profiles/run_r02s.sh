#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02s_tests.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02s_tests.log
grep -E "^\[vit|^\[stage|logits max-abs|FAILED|passed|failed" gpurun_out/r02s_tests.log | tail -20
python profiles/attn_small_time.py | tail -3
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
echo "bench rc $?"; tail -3 gpurun_out/r02s_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02s_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'eager', d['eager']['value'])
print(d['breakdown_ms_per_step'])
PY
