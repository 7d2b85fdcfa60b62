#!/usr/bin/env bash
# BASELINE.json configs beyond the bench headline: ViT-L, 900 queries (dense herd), plus the headline with the pipelined e2e feed
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg_vitb.json 2> gpurun_out/bench_cfg_vitb.err; tail -2 gpurun_out/bench_cfg_vitb.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --model vit_l --batch 32 > gpurun_out/bench_cfg_vitl.json 2> gpurun_out/bench_cfg_vitl.err; tail -2 gpurun_out/bench_cfg_vitl.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --queries 900 --batch 32 > gpurun_out/bench_cfg_q900.json 2> gpurun_out/bench_cfg_q900.err; tail -2 gpurun_out/bench_cfg_q900.err
python - <<'PY'
import json
for n in ("vitb","vitl","q900"):
    try:
        d=json.load(open(f"gpurun_out/bench_cfg_{n}.json")); print(n, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d.get("model_tflops"), d["breakdown_ms_per_step"])
    except Exception as e: print(n, "failed", e)
PY
