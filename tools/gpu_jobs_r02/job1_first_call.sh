#!/bin/bash
# Round 2, first GPU call (prepared at the end of round 1, when the GPU minutes were spent):
#   1. the whole -m gpu suite (includes the q_split = 0 and turbulence tests added last);
#   2. default bench line;
#   3. side measurements that have never been timed: tracer sub-cycling (q_split = 0, 3 issued sub-steps) and the
#      linearised turbulence solves (24 array passes of algorithmic traffic);
#   4. ncu --set full of the turbulence solve kernel (only after its command exited 0 without ncu).
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_jobs_r02/job1_first_call.sh'
mkdir -p gpurun_out
T=r02a
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
python bench.py > gpurun_out/${T}_bench_default_1gpu.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
python bench.py --kernel-only --turbulence --steps 5 --warmup 3 > gpurun_out/${T}_side_turbulence.json 2> gpurun_out/${T}_side_turb.err
rc=$?; cat gpurun_out/${T}_side_turbulence.json
python bench.py --kernel-only --q-split-dynamic 3 --steps 3 --warmup 3 > gpurun_out/${T}_side_qsplit0.json 2> gpurun_out/${T}_side_q.err
cat gpurun_out/${T}_side_qsplit0.json
# 5. fv_tp_2d as shared-memory-tile kernels (csrc/fused_tp.h; built at the end of round 1, CPU-emulation parity only):
#    level 0 = stage chain (the default bench line above), 1 = fused forward sweeps, 2 = fused forward + reverse kernels.
#    Same workload, kernel-only (data resident); then the complete bench line + per-op table with every tile kernel on
#    (FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1; the line carries the dot-product test at the bench workload).
for L in 1 2; do
  FV3LM_FUSED_TP=$L python bench.py --kernel-only --steps 5 --warmup 3 > gpurun_out/${T}_fused_tp_level${L}.json 2> gpurun_out/${T}_fused_tp_level${L}.err
  echo "fused level $L rc=$?"; cat gpurun_out/${T}_fused_tp_level${L}.json
done
FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1 python bench.py --no-cpu --profile-out gpurun_out/${T}_profile_fused_all.txt > gpurun_out/${T}_bench_fused_all.json 2> gpurun_out/${T}_bench_fused_all.err
echo "fused all rc=$?"; cat gpurun_out/${T}_bench_fused_all.json
# 5b. the hand-derived adjoint of nh_p_grad (default path since the end of round 1; the committed capture describes its predecessor)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:KernAD.*S_gradp -c 2 \
    -o gpurun_out/${T}_KernAD_S_gradp_c180 python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_gradp.log 2>&1
ls -la gpurun_out/${T}_KernAD_S_gradp_c180.ncu-rep
# 6. ncu of the tile kernels (only after the same command exited 0 above): launch list of one TL+AD pair and --set full of the
#    reverse kernel and kernel B (one launch each is enough: every launch of a kind does the same work)
if [ -s gpurun_out/${T}_bench_fused_all.json ]; then
  FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
      --log-file gpurun_out/${T}_ncu_launches_fused.csv python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_launches_fused.log 2>&1
  gzip -f gpurun_out/${T}_ncu_launches_fused.csv
  FV3LM_FUSED_TP=2 FV3LM_FUSED_A2B=1 FV3LM_FUSED_CHAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:kern_tile|kern_chain' -s 40 -c 16 \
      -o gpurun_out/${T}_kern_tile_c180 python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_tile.log 2>&1
  ls -la gpurun_out/${T}_kern_tile_c180.ncu-rep
fi
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:KTurbSolve -c 2 \
      -o gpurun_out/${T}_KTurbSolve_c180 python bench.py --kernel-only --turbulence --steps 1 --warmup 1 > gpurun_out/${T}_ncu_turb.log 2>&1
  ls -la gpurun_out/${T}_KTurbSolve_c180.ncu-rep
fi
