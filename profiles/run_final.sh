#!/usr/bin/env bash
# final state of the round: smoke(), the whole -m gpu suite, the default bench (all configs), the reference arm
tag=${1:-r02final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "pytest rc $?" >> gpurun_out/${tag}_tests.log
grep -E "FAILED|passed|failed" gpurun_out/${tag}_tests.log | tail -5
timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc $?"; tail -2 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
echo "reference rc $?"; cut -c1-300 gpurun_out/${tag}_bench_reference.json
python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'eager', d['eager']['value'], 'roof', d['roofline']['frac'], d['clocks'])
print(d['breakdown_ms_per_step'])
for k, v in (d.get('configs') or {}).items():
    print(k, v if not isinstance(v, dict) else (round(v['value'],1), round(v['ms_per_step'],2), v['breakdown_ms_per_step']))
dh = d['dense_herd']
print('dense', dh and (round(dh['value'],1), dh['nms'], dh.get('nms_10k')))
print(d['cpu_baseline'])
PY
