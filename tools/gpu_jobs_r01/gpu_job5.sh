#!/bin/bash
# GPU job 5: ncu --set full on representative kernels (C96, one TL+AD pair), full per-op profile table at C180 NH
mkdir -p gpurun_out
python bench.py --nonhydro --steps 2 --warmup 3 --no-cpu --profile-out gpurun_out/r01e_profile_c180_nh.txt > gpurun_out/r01e_bench_c180_nh.json 2> gpurun_out/r01e.err
CMD="python bench.py --nonhydro --res 96 --steps 1 --warmup 0 --kernel-only"
$CMD > gpurun_out/plain96.log 2>&1 || exit 1
for pat in 'KernNL.*S_ppm' 'KernTL.*S_ppm' 'KernAD.*S_ppm' 'KernTL.*S_dupd' 'KernAD.*S_gradp' 'KernColAD.*S_riem' 'KernNL.*S_dke'; do
  name=$(echo $pat | tr -d '.*')
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -c 2 -o gpurun_out/r01e_$name $CMD > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
ls -la gpurun_out
