#!/bin/bash
# round 2, call C: levels per thread (FV3LM_KPT) A/B; L72 parity tests + table of achieved errors; ncu --set full of representative kernels
mkdir -p gpurun_out
T=r02c
FV3LM_PARITY_OUT=gpurun_out/${T}_parity_errors_gpu.json python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for KP in 1 2 4 8; do
  FV3LM_KPT=$KP python bench.py --kernel-only --steps 3 --warmup 2 --profile-out gpurun_out/${T}_profile_kpt${KP}.txt > gpurun_out/${T}_ko_kpt${KP}.json 2> gpurun_out/${T}_ko_kpt${KP}.err
  echo "kpt $KP rc=$?"; cat gpurun_out/${T}_ko_kpt${KP}.json; tail -c 300 gpurun_out/${T}_ko_kpt${KP}.err
done
for K in "KernTL<fv3lm::S_ppm<0>" "KernAD<fv3lm::S_ppm<0>" "KernNL<fv3lm::S_inner<0>" "KernTL<fv3lm::S_dupd" "KernColTL<fv3lm::S_tri" "KernColTL<fv3lm::S_edge_profile" "KernAD<fv3lm::S_dwind2" "KernAD<fv3lm::S_del_flux<0>"; do
  nm=$(echo "$K" | sed 's/fv3lm:://g' | tr -cd 'A-Za-z0-9_')
  FV3LM_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 3 -c 1 \
      -o gpurun_out/${T}_ncu_${nm} python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_${nm}.log 2>&1
  ls -la gpurun_out/${T}_ncu_${nm}.ncu-rep 2>/dev/null | awk '{print $5, $9}'; tail -2 gpurun_out/${T}_ncu_${nm}.log
done
