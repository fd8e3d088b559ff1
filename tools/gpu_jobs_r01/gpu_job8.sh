#!/bin/bash
# 8-GPU box: NCCL parity at 4 and 8 ranks (layouts 1x2 and 2x2), sharded C180 NH bench at N = 4 and 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/multigpu_check.py --nonhydro 2>&1 | grep multigpu_check | tee -a gpurun_out/r01n_multigpu_check_48.txt
done
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r01n_bench_c180_nh_${n}gpu.json 2> gpurun_out/r01n_bench_c180_nh_${n}gpu.err
  tail -c 400 gpurun_out/r01n_bench_c180_nh_${n}gpu.err
  python -c "import json;d=json.loads(open('gpurun_out/r01n_bench_c180_nh_${n}gpu.json').read().strip().splitlines()[-1]);print('NH ${n}GPU value',d['value'],'tl_ms',d['tl_ms'],'ad_ms',d['ad_ms'],'launches',d['gpu_launches'],'e2e',d['e2e']['value'],'pool',d['pool_peak_gb'],d['config']['multi_gpu'])"
done
