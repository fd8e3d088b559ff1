#!/bin/bash
# round 2, call B: generic launcher without per-thread integer divisions; quick gpu test subset; kernel-only bench + per-op table;
# ncu --set full of a few representative kernels (one launch each) to see what bounds them now
mkdir -p gpurun_out
T=r02b
python -m pytest tests/test_tp_core.py tests/test_step_api.py tests/test_zz_fused_tp.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --kernel-only --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_base.txt > gpurun_out/${T}_ko_base.json 2> gpurun_out/${T}_ko_base.err
echo "base rc=$?"; cat gpurun_out/${T}_ko_base.json; tail -c 300 gpurun_out/${T}_ko_base.err
FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1 python bench.py --kernel-only --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_all.txt > gpurun_out/${T}_ko_all.json 2> gpurun_out/${T}_ko_all.err
echo "all rc=$?"; cat gpurun_out/${T}_ko_all.json
for K in "KernTL.*S_ppm.Li0" "KernAD.*S_ppm.Li0" "KernNL.*S_inner.Li0" "KernTL.*S_dupd" "KernColTL.*S_tri" "KernColAD.*S_remap" "KernColTL.*S_edge_profile" "KernAD.*S_dwind2" "KernAD.*S_del_flux.Li0"; do
  nm=$(echo "$K" | tr -cd 'A-Za-z0-9_')
  FV3LM_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 3 -c 1 \
      -o gpurun_out/${T}_ncu_${nm} python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_${nm}.log 2>&1
  ls -la gpurun_out/${T}_ncu_${nm}.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
