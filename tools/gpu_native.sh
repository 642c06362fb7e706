#!/bin/bash
# Fast GPU iteration without Python (a call costs ~30 s of box time instead of minutes): parity of the hot path against
# the oracle's stored outputs + stage timings, and the knn2 A/B check. Build first, HERE:
#   make -C orb_slam3_fast_b200/csrc -s -j8 && make -C tools/ubench -s
#   gpurun --timeout 120 -- 'bash tools/gpu_native.sh <tag> [pairs] [steps]'
TAG=${1:-n}
mkdir -p gpurun_out
timeout 60 tools/ubench/hotpath_check ${2:-1024} ${3:-10} 2>&1 | tee gpurun_out/hotpath_$TAG.log
if [ -n "$KNN" ]; then timeout 60 tools/ubench/knn2_abi_check 2>&1 | tee gpurun_out/knn2_abi_$TAG.log; fi
