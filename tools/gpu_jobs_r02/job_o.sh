#!/bin/bash
# round 2, call O: reverse tile kernel back to sequential accumulates (with the cube-edge weight tables): gpu tests with tightened tolerances,
# kernel-only bench + per-op table, ncu of the two reverse kernels
mkdir -p gpurun_out
T=r02o
python -m pytest tests/test_zz_fused_tp.py tests/test_step_api.py tests/test_fv_dynamics.py tests/test_nh.py tests/test_dyn_core.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --kernel-only --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_default.txt > gpurun_out/${T}_ko_default.json 2> gpurun_out/${T}_ko_default.err
cat gpurun_out/${T}_ko_default.json; grep tp_fused gpurun_out/${T}_profile_default.txt
tools/ncu_capture.sh $T TpRevB "kern_tile<fv3lm::ftp::KernTpRev<.int.0"
tools/ncu_capture.sh $T TpRevA "kern_tile<fv3lm::ftp::KernTpRev<.int.1"
