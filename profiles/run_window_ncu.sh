#!/usr/bin/env bash
# ncu --set full of the window-attention kernel in isolation (profiles/window_time.py, batch 32): bash profiles/run_window_ncu.sh <tag>
tag=${1:-r02x}
mkdir -p gpurun_out
timeout 200 python profiles/window_time.py 32 > gpurun_out/${tag}_window_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_window_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window2_kernel -s 2 -c 1 -o gpurun_out/${tag}_prof_window2 python profiles/window_time.py 32 > gpurun_out/${tag}_ncu_window2.log 2>&1
tail -2 gpurun_out/${tag}_ncu_window2.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/${tag}_prof_window2.ncu-rep 30 > gpurun_out/${tag}_ncu_window2_summary.txt 2>&1; head -40 gpurun_out/${tag}_ncu_window2_summary.txt | cut -c1-300
