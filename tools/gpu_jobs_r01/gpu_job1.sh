#!/bin/bash
# round-1 GPU job: parity tests, bench, ncu launch list, ncu full capture of the top kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01_pytest_gpu.txt
cat gpurun_out/r01_pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r01b_bench_c180_hydro.json 2> gpurun_out/r01b_bench_c180_hydro.err
tail -c 600 gpurun_out/r01b_bench_c180_hydro.json
python bench.py --res 48 --steps 1 --warmup 1 --kernel-only > gpurun_out/plain48.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r01_launches_c48.csv \
    python bench.py --res 48 --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu48.log 2>&1
python bench.py --steps 1 --warmup 0 --kernel-only > gpurun_out/plain180.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'KernAD.*S_gradp|KernTL.*S_ppm|KernAD.*S_ppm' -c 6 \
    -o gpurun_out/r01_top_kernels python bench.py --steps 1 --warmup 0 --kernel-only > gpurun_out/ncu180.log 2>&1
ls -la gpurun_out
