#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu --profile-out gpurun_out/r01o_profile_c180_nh.txt > gpurun_out/r01o_bench_c180_nh.json 2> gpurun_out/r01o.err; tail -c 300 gpurun_out/r01o.err
python -c "import json;d=json.loads(open('gpurun_out/r01o_bench_c180_nh.json').read().strip().splitlines()[-1]);print('DECOMPOSED value',d['value'],'tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'pool',d['pool_peak_gb'])"
grep -E "rs_|riem" gpurun_out/r01o_profile_c180_nh.txt | awk '{s+=$3; printf "%s %s; ", $1, $3} END {print "\nriem total ms", s}'
FV3LM_RIEM_MONO=1 python bench.py --steps 3 --warmup 3 --kernel-only 2>&1 | tail -1
