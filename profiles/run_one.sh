#!/usr/bin/env bash
# generic single-command runner: bash profiles/run_one.sh <tag> <command...>   (stdout+stderr tee'd to gpurun_out/<tag>.txt)
tag=$1; shift
mkdir -p gpurun_out
"$@" 2>&1 | tee gpurun_out/${tag}.txt | tail -40
