#!/bin/bash
# GPU job 3: parity after the adjoint rework, NH + hydro bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r01_pytest_gpu3.txt
cat gpurun_out/r01_pytest_gpu3.txt
python bench.py --nonhydro --steps 3 --warmup 3 --no-cpu > gpurun_out/r01d_bench_c180_nh.json 2> gpurun_out/r01d_bench_c180_nh.err
tail -c 300 gpurun_out/r01d_bench_c180_nh.err
python -c "import json;d=json.load(open('gpurun_out/r01d_bench_c180_nh.json'));print('NH tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb']);print(json.dumps(d['top_kernels'])[:1800])"
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r01d_bench_c180_hydro.json 2> gpurun_out/r01d_bench_c180_hydro.err
python -c "import json;d=json.load(open('gpurun_out/r01d_bench_c180_hydro.json'));print('HYDRO tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb']);print(json.dumps(d['top_kernels'])[:1800])"
