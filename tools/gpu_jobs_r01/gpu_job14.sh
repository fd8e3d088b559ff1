#!/bin/bash
# CUDA-graph replay: GPU tests, single-GPU bench with and without graphs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- graphs"; python bench.py --steps 3 --warmup 3 --kernel-only 2>&1 | tail -1
echo "--- no graphs"; FV3LM_NO_GRAPH=1 python bench.py --steps 3 --warmup 3 --kernel-only 2>&1 | tail -1
echo "--- C48 graphs"; python bench.py --res 48 --steps 3 --warmup 3 --kernel-only 2>&1 | tail -1
echo "--- C48 no graphs"; FV3LM_NO_GRAPH=1 python bench.py --res 48 --steps 3 --warmup 3 --kernel-only 2>&1 | tail -1
