#!/bin/bash
# Round 2, first 2-GPU call: NCCL transport of the pieces added at the end of round 1 (their exchanges were only run over gloo).
# usage: gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_jobs_r02/job2_two_gpus.sh'
mkdir -p gpurun_out
T=r02b
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/multigpu_check.py "${@:2}"; }
( run 29511 --nonhydro; run 29512 --module tracer_2d; run 29513 --module c2l_ord4 ) 2>&1 | grep -v "^W\|Warning" | tee gpurun_out/${T}_multigpu_check_2gpu.txt | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 \
    > gpurun_out/${T}_bench_c180_nh_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; tail -c 400 gpurun_out/${T}_bench_c180_nh_2gpu.json
