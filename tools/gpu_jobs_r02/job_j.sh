#!/bin/bash
# round 2, call J: row-marching forward kernels of fv_tp_2d (default): gpu tests, kernel-only bench A/B, ncu
mkdir -p gpurun_out
T=r02j
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
tools/ncu_capture.sh $T TL_MarchA "kern_march<fv3lm::ftp::MarchA<fv3lm::Dual"
tools/ncu_capture.sh $T NL_MarchB "kern_march<fv3lm::ftp::MarchB<double"
du -sh gpurun_out
