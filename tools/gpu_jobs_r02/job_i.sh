#!/bin/bash
# round 2, call I: compiled a2b_ord4 (default) + KPT=4 default: gpu tests, kernel-only bench A/B, ncu of the new kernels
mkdir -p gpurun_out
T=r02i
python -m pytest tests/test_zz_fused_tp.py tests/test_d_sw.py tests/test_nh.py tests/test_step_api.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() {  # name, env...
  local nm=$1; shift
  env "$@" python bench.py --kernel-only $XARGS --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_${nm}.txt > gpurun_out/${T}_ko_${nm}.json 2> gpurun_out/${T}_ko_${nm}.err
  echo "$nm rc=$?"; cat gpurun_out/${T}_ko_${nm}.json; tail -c 300 gpurun_out/${T}_ko_${nm}.err
}
run default FV3LM_DEBUG_SEG=1
run a2b0 FV3LM_FUSED_A2B=0
run kpt8 FV3LM_KPT=8
run budget FV3LM_AD_STORE_BUDGET=1.62e11 FV3LM_DEBUG_SEG=1
XARGS="--res 64"; run c64; XARGS=""
tools/ncu_capture.sh $T TL_a2bReg "kern3d<fv3lm::a2bc::KReg<.bool.1"
tools/ncu_capture.sh $T AD_a2bRegT "kern3d<fv3lm::a2bc::KRegT"
tools/ncu_capture.sh $T AD_a2bRowsT "kern3d<fv3lm::a2bc::KRowsT"
du -sh gpurun_out
