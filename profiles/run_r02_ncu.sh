#!/usr/bin/env bash
# round-2 ncu evidence: (1) launch list of one eager bench step, (2) --set full of the flash v7 kernel and of the qkv GEMM
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline --no-extras --no-graph"
$CMD > gpurun_out/r02_launches_plain.log 2>&1 || { tail -5 gpurun_out/r02_launches_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|flash|window|layernorm|attn_small|attn_mma|transpose|patchify|hfc_finalize|add_cast|postprocess|nms" -c 4000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_launches_ncu.log 2>&1
tail -2 gpurun_out/r02_launches_ncu.log | cut -c1-200; wc -l gpurun_out/r02_launches.csv
python profiles/launch_list_summary.py gpurun_out/r02_launches.csv gpurun_out/r02_ncu_launch_list
CMD8="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-extras --no-graph"
$CMD8 > gpurun_out/r02_ncu_plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flash7_kernel -s 1 -c 1 -o gpurun_out/r02_prof_flash7 $CMD8 > gpurun_out/r02_ncu_flash7.log 2>&1
tail -2 gpurun_out/r02_ncu_flash7.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/r02_prof_flash7.ncu-rep 25 > gpurun_out/r02_ncu_flash7_summary.txt 2>&1; head -5 gpurun_out/r02_ncu_flash7_summary.txt | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:attn_mma_kernel -s 2 -c 1 -o gpurun_out/r02_prof_attn_mma $CMD8 > gpurun_out/r02_ncu_attn_mma.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_prof_attn_mma.ncu-rep 12 > gpurun_out/r02_ncu_attn_mma_summary.txt 2>&1; head -3 gpurun_out/r02_ncu_attn_mma_summary.txt | cut -c1-400
