#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --steps 1 --warmup 0 --kernel-only"
FV3LM_NO_GRAPH=1 $CMD > gpurun_out/plain180.log 2>&1 || exit 1
FV3LM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'KernColAD.*S_riem' -c 2 -o gpurun_out/r01i_KernColAD_S_riem_c180 $CMD > gpurun_out/ncu_riem180.log 2>&1
tail -2 gpurun_out/ncu_riem180.log
