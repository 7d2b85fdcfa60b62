#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused_layernorm or nms or nan or attn_small" -p no:cacheprovider 2>&1 | tail -8
python profiles/nms_time.py 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_model_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
echo "bench rc $?"; tail -3 gpurun_out/r02p_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02p_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'eager', d['eager']['value'])
print(d['breakdown_ms_per_step'])
PY
