#!/bin/bash
# round 2, call G: lean tile kernels (fv_tp_2d forward / reverse, a2b_ord4): gpu tests, A/B at the bench workload, ncu summaries
mkdir -p gpurun_out
T=r02g
python -m pytest tests/test_zz_fused_tp.py tests/test_tp_core.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() {  # name, env...
  local nm=$1; shift
  env "$@" python bench.py --kernel-only --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_${nm}.txt > gpurun_out/${T}_ko_${nm}.json 2> gpurun_out/${T}_ko_${nm}.err
  echo "$nm rc=$?"; cat gpurun_out/${T}_ko_${nm}.json; tail -c 300 gpurun_out/${T}_ko_${nm}.err
}
run base FV3LM_FUSED_TP=0
run tp2 FV3LM_FUSED_TP=2
run a2b FV3LM_FUSED_A2B=1
run tp2_a2b FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1
tools/ncu_capture.sh $T TL_TpB "kern_tile<fv3lm::ftp::KernTpB<fv3lm::Dual" FV3LM_FUSED_TP=2
tools/ncu_capture.sh $T TL_TpA "kern_tile<fv3lm::ftp::KernTpA<fv3lm::Dual" FV3LM_FUSED_TP=2
tools/ncu_capture.sh $T TpRevB "kern_tile<fv3lm::ftp::KernTpRev<.int.0" FV3LM_FUSED_TP=2
tools/ncu_capture.sh $T A2bRev "kern_tile<fv3lm::ftp::KernA2bRev" FV3LM_FUSED_A2B=1
tools/ncu_capture.sh $T TL_A2b "kern_tile<fv3lm::ftp::KernA2b<fv3lm::Dual" FV3LM_FUSED_A2B=1
tools/ncu_capture.sh $T TL_ppm0 "KernTL<fv3lm::S_ppm<.int.0>"
tools/ncu_capture.sh $T AD_inner0 "KernAD<fv3lm::S_inner<.int.0>"
tools/ncu_capture.sh $T TL_dupd "KernTL<fv3lm::S_dupd"
tools/ncu_capture.sh $T ColAD_remap "KernColAD<fv3lm::S_remap"
du -sh gpurun_out
