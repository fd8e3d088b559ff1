#!/bin/bash
# GPU job 4 (2 GPUs): NCCL multi-rank parity + sharded bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py --nonhydro 2>&1 | tail -5 | tee gpurun_out/r01_multigpu_check_2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_check.py 2>&1 | tail -3 | tee -a gpurun_out/r01_multigpu_check_2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --nonhydro --steps 3 --warmup 3 > gpurun_out/r01d_bench_c180_nh_2gpu.json 2> gpurun_out/r01d_bench_c180_nh_2gpu.err
tail -c 600 gpurun_out/r01d_bench_c180_nh_2gpu.err
python -c "import json;d=json.load(open('gpurun_out/r01d_bench_c180_nh_2gpu.json'));print('NH 2GPU value',d['value'],'tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],d['config']['multi_gpu'])"
