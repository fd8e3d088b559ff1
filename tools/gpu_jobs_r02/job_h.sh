#!/bin/bash
# round 2, call H: FUSED_TP=2 default + in-place edge_profile: fused gpu tests, kernel-only bench (KPT A/B), per-instruction ncu counts
mkdir -p gpurun_out
T=r02h
python -m pytest tests/test_zz_fused_tp.py tests/test_nh.py tests/test_step_api.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() {  # name, env...
  local nm=$1; shift
  env "$@" python bench.py --kernel-only $XARGS --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_${nm}.txt > gpurun_out/${T}_ko_${nm}.json 2> gpurun_out/${T}_ko_${nm}.err
  echo "$nm rc=$?"; cat gpurun_out/${T}_ko_${nm}.json; tail -c 300 gpurun_out/${T}_ko_${nm}.err
}
run default FV3LM_KPT=1
run kpt2 FV3LM_KPT=2
run kpt4 FV3LM_KPT=4
XARGS="--res 64"; run c64 FV3LM_KPT=1; XARGS=""
tools/ncu_capture.sh $T TL_TpB "kern_tile<fv3lm::ftp::KernTpB<fv3lm::Dual"
tools/ncu_capture.sh $T TpRevA "kern_tile<fv3lm::ftp::KernTpRev<.int.1"
tools/ncu_capture.sh $T TL_ppm0 "KernTL<fv3lm::S_tpuv<.int.0>"
tools/ncu_capture.sh $T ColTL_edge "KernColTL<fv3lm::S_edge_profile"
tools/ncu_capture.sh $T AD_delflux "KernAD<fv3lm::S_del_flux<.int.0>"
du -sh gpurun_out
