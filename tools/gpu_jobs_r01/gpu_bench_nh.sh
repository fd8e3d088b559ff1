#!/bin/bash
# quick single-GPU measurement: gpu tests (fast subset) + C180 NH bench with the full per-op table
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu --profile-out gpurun_out/${1:-rX}_profile_c180_nh.txt > gpurun_out/${1:-rX}_bench_c180_nh.json 2> gpurun_out/${1:-rX}.err
tail -c 300 gpurun_out/${1:-rX}.err
python -c "import json;d=json.loads(open('gpurun_out/${1:-rX}_bench_c180_nh.json').read().strip().splitlines()[-1]);print('NH value',d['value'],'tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb'])"
head -25 gpurun_out/${1:-rX}_profile_c180_nh.txt
