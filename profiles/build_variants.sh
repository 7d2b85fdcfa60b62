#!/usr/bin/env bash
# build A/B variants of the library: bash profiles/build_variants.sh name1 "flags1" name2 "flags2" ...  -> libwm_b200_<name>.so
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  WM_BUILD_DIR=build_$1 WM_LIB_NAME=libwm_b200_$1.so WM_NVCC_EXTRA="$2" bash wildlifemapper_b200/csrc/build.sh | tail -1 &
  shift 2
done
wait
