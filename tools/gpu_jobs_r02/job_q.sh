#!/bin/bash
# round 2, call Q (1 GPU): packed thread mapping on small sub-domains + trimmed tile grids: whole gpu suite, C64 / C180 kernel-only A/B
mkdir -p gpurun_out
T=r02q
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/${T}_pytest_gpu.txt 2>&1; grep -E "passed|failed" gpurun_out/${T}_pytest_gpu.txt | tail -2
run() {  # name, env...
  local nm=$1; shift
  env "$@" python bench.py --kernel-only $XARGS --steps 3 --warmup 3 > gpurun_out/${T}_ko_${nm}.json 2> gpurun_out/${T}_ko_${nm}.err
  echo "$nm rc=$?"; cat gpurun_out/${T}_ko_${nm}.json; tail -c 200 gpurun_out/${T}_ko_${nm}.err
}
XARGS="--res 64"; run c64_packed FV3LM_PACKED=1; run c64_2d FV3LM_PACKED=0; run c64_auto; XARGS=""
run c180 FV3LM_DEBUG_SEG=1
run c180_budget FV3LM_AD_STORE_BUDGET=1.66e11 FV3LM_DEBUG_SEG=1
