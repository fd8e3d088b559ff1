#!/bin/bash
# round 2, call E: ncu --set full of representative kernels of the committed default path and of the tile kernels (one launch each, C180 L72)
mkdir -p gpurun_out
T=r02f
cap() {  # name regex env...
  local nm=$1 K=$2; shift 2
  env FV3LM_NO_GRAPH=1 "$@" timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 3 -c 1 \
      -o gpurun_out/${T}_ncu_${nm} python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_${nm}.log 2>&1
  ls -la gpurun_out/${T}_ncu_${nm}.ncu-rep 2>/dev/null | awk '{print $5, $9}'; tail -1 gpurun_out/${T}_ncu_${nm}.log | cut -c1-200
}
cap TL_ppm0   "KernTL<fv3lm::S_ppm<.int.0>"
cap AD_ppm0   "KernAD<fv3lm::S_ppm<.int.0>"
cap NL_ppm0   "KernNL<fv3lm::S_ppm<.int.0>"
cap AD_inner0 "KernAD<fv3lm::S_inner<.int.0>"
cap TL_inner0 "KernTL<fv3lm::S_inner<.int.0>"
cap AD_a2b_q1 "KernAD<fv3lm::S_a2b_q1<.int.0>"

cap ColAD_remap "KernColAD<fv3lm::S_remap"
cap AD_tpuv0 "KernAD<fv3lm::S_tpuv<.int.0>"
