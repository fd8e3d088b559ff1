#!/bin/bash
# round 2, closing multi-GPU measurements: usage  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_jobs_r02/job_final_ngpu.sh N'
N=$1
mkdir -p gpurun_out
T=r02r
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/multigpu_check.py "${@:2}"; }
( run 29511 --nonhydro ) 2>&1 | grep -v "^W\|Warning" | tee gpurun_out/${T}_multigpu_check_${N}gpu.txt | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/${T}_bench_c180_nh_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; cut -c1-700 gpurun_out/${T}_bench_c180_nh_${N}gpu.json; tail -c 300 gpurun_out/${T}_bench_${N}gpu.err
