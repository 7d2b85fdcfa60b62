#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02o_tests.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02o_tests.log
tail -6 gpurun_out/r02o_tests.log
timeout 900 python bench.py --no-extras --breakdown gpurun_out/r02o_breakdown.json > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02o_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'eager', d['eager']['value'])
print(d['breakdown_ms_per_step'])
PY
