#!/bin/bash
# sweep of the L2 residency budget of the column kernels (C180 NH, per-op table restricted to column stages)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "nh or step_api" 2>&1 | tail -2
for mb in 100000 400 192 96 48; do
  FV3LM_COL_L2_MB=$mb python bench.py --steps 2 --warmup 2 --no-cpu --profile-out gpurun_out/col_$mb.txt > gpurun_out/col_$mb.json 2>/dev/null
  python -c "import json;d=json.loads(open('gpurun_out/col_$mb.json').read().strip().splitlines()[-1]);print('L2_MB=$mb tl_ms %.1f ad_ms %.1f'%(d['tl_ms'],d['ad_ms']))"
  grep -E "riem_solver|map1_q2|map_scalar|edge_profile_crx|pk3|geopk|map1_ppm_u" gpurun_out/col_$mb.txt | awk '{printf "   %s %s ms;", $1, $3} END {print ""}'
done
