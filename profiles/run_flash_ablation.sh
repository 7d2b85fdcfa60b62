#!/usr/bin/env bash
# Where does the time of the global-attention kernel go?  Builds variants of libwm_b200.so with parts of flash4_kernel
# removed (results are WRONG on purpose) -- run here, in the build container:   bash profiles/run_flash_ablation.sh build
# and times them next to the product build inside ONE gpurun call:               gpurun -- bash profiles/run_flash_ablation.sh
set -e
cd "$(dirname "$0")/.."
if [ "$1" = "build" ]; then
  WM_BUILD_DIR=build_noex2 WM_LIB_NAME=libwm_b200_noex2.so WM_NVCC_EXTRA="-DWM_F4_NO_EX2" bash wildlifemapper_b200/csrc/build.sh | tail -1
  WM_BUILD_DIR=build_onemma WM_LIB_NAME=libwm_b200_onemma.so WM_NVCC_EXTRA="-DWM_F4_ONE_MMA" bash wildlifemapper_b200/csrc/build.sh | tail -1
  WM_BUILD_DIR=build_both WM_LIB_NAME=libwm_b200_both.so WM_NVCC_EXTRA="-DWM_F4_ONE_MMA -DWM_F4_NO_EX2" bash wildlifemapper_b200/csrc/build.sh | tail -1
  WM_BUILD_DIR=build_skel WM_LIB_NAME=libwm_b200_skel.so WM_NVCC_EXTRA="-DWM_F4_SKELETON" bash wildlifemapper_b200/csrc/build.sh | tail -1
  WM_BUILD_DIR=build_skel1 WM_LIB_NAME=libwm_b200_skel1.so WM_NVCC_EXTRA="-DWM_F4_SKELETON -DWM_F4_ONE_MMA" bash wildlifemapper_b200/csrc/build.sh | tail -1
  exit 0
fi
for l in libwm_b200.so libwm_b200_noex2.so libwm_b200_onemma.so libwm_b200_both.so libwm_b200_skel.so libwm_b200_skel1.so libwm_b200.so; do
  WM_LIB_NAME=$l timeout 100 python profiles/flash_time.py 2>&1 | tail -1
done
