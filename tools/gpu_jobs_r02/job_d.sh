#!/bin/bash
# round 2, call D (re-entry): gpu suite (no -x) with the parity-error table, A/B of the tile-kernel switches (kernel-only), per-op tables
mkdir -p gpurun_out
T=r02d
( time FV3LM_PARITY_OUT=gpurun_out/${T}_parity_errors_gpu.json python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.txt
run() {  # name, env...
  local nm=$1; shift
  env "$@" python bench.py --kernel-only --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_${nm}.txt > gpurun_out/${T}_ko_${nm}.json 2> gpurun_out/${T}_ko_${nm}.err
  echo "$nm rc=$?"; cat gpurun_out/${T}_ko_${nm}.json; tail -c 300 gpurun_out/${T}_ko_${nm}.err
}
run base FV3LM_FUSED_TP=0
run tp2 FV3LM_FUSED_TP=2
run a2b FV3LM_FUSED_A2B=1
run chain FV3LM_FUSED_CHAIN=1
run all FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1
run all_kpt2 FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1 FV3LM_KPT=2
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv
