#!/bin/bash
# NCCL exchanges inside captured CUDA graphs: 2 GPUs, sub-domain layout, graphs vs eager must agree; sharded bench
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}" 2>&1 | grep -E "kernel_only|RuntimeError|Error" | head -3; }
echo "--- N=2 C96 layout 2x2 graphs"; run 29801 --res 96 --layout 2 2 --steps 3 --warmup 3 --kernel-only
echo "--- N=2 C96 layout 2x2 eager"; FV3LM_NO_GRAPH=1 run 29802 --res 96 --layout 2 2 --steps 3 --warmup 3 --kernel-only
echo "--- N=2 C180 graphs full bench"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29803 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r01h_bench_c180_nh_2gpu.json 2> gpurun_out/r01h_2gpu.err; tail -c 300 gpurun_out/r01h_2gpu.err
python -c "import json;d=json.loads(open('gpurun_out/r01h_bench_c180_nh_2gpu.json').read().strip().splitlines()[-1]);print('NH 2GPU value',d['value'],'tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb'])"
