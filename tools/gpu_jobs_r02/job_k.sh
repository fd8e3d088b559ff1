#!/bin/bash
# round 2, call K: L2 prefetch in the marching and reverse tile kernels, unrolled forward column loops: gpu tests, kernel-only bench A/B, ncu
mkdir -p gpurun_out
T=r02k
python -m pytest tests/test_zz_fused_tp.py tests/test_tp_core.py tests/test_step_api.py tests/test_tracer_2d.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() {  # name, env...
  local nm=$1; shift
  env "$@" python bench.py --kernel-only $XARGS --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_${nm}.txt > gpurun_out/${T}_ko_${nm}.json 2> gpurun_out/${T}_ko_${nm}.err
  echo "$nm rc=$?"; cat gpurun_out/${T}_ko_${nm}.json; tail -c 300 gpurun_out/${T}_ko_${nm}.err
}
run default FV3LM_TP_MARCH=1
run march0 FV3LM_TP_MARCH=0
XARGS="--res 64"; run c64; XARGS=""
tools/ncu_capture.sh $T TL_MarchB "kern_march<fv3lm::ftp::MarchB<fv3lm::Dual"
tools/ncu_capture.sh $T TpRevB "kern_tile<fv3lm::ftp::KernTpRev<.int.0"
du -sh gpurun_out
