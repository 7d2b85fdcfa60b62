#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02v_tests.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02v_tests.log
grep -E "FAILED|passed|failed" gpurun_out/r02v_tests.log | tail -5
timeout 1200 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err
echo "bench rc $?"; tail -3 gpurun_out/r02v_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02v_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'eager', d['eager']['value'], 'roof', d['roofline']['frac'])
print(d['breakdown_ms_per_step'])
for k, v in (d.get('configs') or {}).items():
    print(k, v if not isinstance(v, dict) else (round(v['value'],1), round(v['ms_per_step'],2), v['breakdown_ms_per_step']))
dh = d['dense_herd']
print('dense', dh and (round(dh['value'],1), dh['nms'], dh.get('nms_10k')))
print(d['cpu_baseline'])
PY
