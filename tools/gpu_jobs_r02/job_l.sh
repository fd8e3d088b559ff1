#!/bin/bash
# round 2, call L: unrolled phase loops in the fv_tp_2d tile kernels: kernel-only bench + per-op table
mkdir -p gpurun_out
T=r02l
python -m pytest tests/test_zz_fused_tp.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
python bench.py --kernel-only --steps 3 --warmup 3 --profile-out gpurun_out/${T}_profile_default.txt > gpurun_out/${T}_ko_default.json 2> gpurun_out/${T}_ko_default.err
cat gpurun_out/${T}_ko_default.json; grep tp_fused gpurun_out/${T}_profile_default.txt
