#!/bin/bash
# round 2: 8 GPUs, layout 1x4 (180 x 45 sub-domains: full-width rows for the 32 x 16 tile kernels) against the default 2x2 (90 x 90)
mkdir -p gpurun_out
T=r02p
for L in "1 4" "4 1"; do
  nm=$(echo $L | tr ' ' 'x')
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 --layout $L --no-cpu \
    2> gpurun_out/${T}_bench_8gpu_${nm}.err | grep '^{' > gpurun_out/${T}_bench_c180_nh_8gpu_layout${nm}.json
  cut -c1-260 gpurun_out/${T}_bench_c180_nh_8gpu_layout${nm}.json; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_c180_nh_8gpu_layout${nm}.json')); print('layout ${nm}', d['value'], d['tl_ms'], d['ad_ms'])"
done
