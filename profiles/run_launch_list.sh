#!/usr/bin/env bash
# ncu launch list (device time of every launch; cold-cache, serialised: compare SHARES with bench.py's event breakdown)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline"
$CMD > gpurun_out/launches_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|flash|window|layernorm|attn_small|transpose|patchify|hfc_finalize|add_cast|postprocess|nms" -c 4000 --csv --log-file gpurun_out/launches_r01z.csv $CMD > gpurun_out/launches_ncu.log 2>&1
tail -2 gpurun_out/launches_ncu.log | cut -c1-300; wc -l gpurun_out/launches_r01z.csv
