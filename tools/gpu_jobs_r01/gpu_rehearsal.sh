#!/bin/bash
# round-end rehearsal: smoke(), default bench (N=1, with cpu_baseline), reference arm
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6
( time python bench.py > gpurun_out/r01q_bench_default.json 2> gpurun_out/r01q_bench_default.err ) 2>&1 | tail -3
tail -c 400 gpurun_out/r01q_bench_default.err
python -c "import json;d=json.loads(open('gpurun_out/r01q_bench_default.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','dtype','gpu_launches')});print(d['roofline']);print(d['cpu_baseline']);print(d['e2e']);print(d['clocks'])"
( time python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r01q_bench_reference.json 2> gpurun_out/r01q_ref.err ) 2>&1 | tail -3
cut -c1-700 gpurun_out/r01q_bench_reference.json
