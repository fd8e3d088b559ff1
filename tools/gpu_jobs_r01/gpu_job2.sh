#!/bin/bash
# GPU job 2: parity tests after the decomposition refactor, hydrostatic + non-hydrostatic bench, ncu full on top kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01_pytest_gpu2.txt
cat gpurun_out/r01_pytest_gpu2.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r01c_bench_c180_hydro.json 2> gpurun_out/r01c_bench_c180_hydro.err
tail -c 300 gpurun_out/r01c_bench_c180_hydro.err
python -c "import json;d=json.load(open('gpurun_out/r01c_bench_c180_hydro.json'));print('HYDRO tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb'])"
python bench.py --nonhydro --steps 3 --warmup 3 --no-cpu > gpurun_out/r01c_bench_c180_nh.json 2> gpurun_out/r01c_bench_c180_nh.err
tail -c 300 gpurun_out/r01c_bench_c180_nh.err
python -c "import json;d=json.load(open('gpurun_out/r01c_bench_c180_nh.json'));print('NH tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb']);print(json.dumps(d['top_kernels'])[:1500])"
python bench.py --res 96 --steps 1 --warmup 0 --kernel-only > gpurun_out/plain96.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'KernAD<fv3lm::S_gradp>|KernTL<fv3lm::S_ppm<0>>|KernAD<fv3lm::S_ppm<0>>|KernColAD<fv3lm::S_remap>' -c 4 \
    -o gpurun_out/r01_top_kernels python bench.py --res 96 --steps 1 --warmup 0 --kernel-only > gpurun_out/ncu96.log 2>&1
tail -5 gpurun_out/ncu96.log
ls -la gpurun_out
