#!/bin/bash
# round-1 closing job: GPU parity subset, the default bench line, ncu --set full of the dominant kernel at the bench workload,
# the two-sided (reference default split configuration) side measurement, ncu launch list of one TL+AD pair
mkdir -p gpurun_out
T=r01t
timeout 150 python -m pytest tests -m gpu -x -q -k "d_sw or step_api or tp_core" 2>&1 | tail -3 > gpurun_out/${T}_pytest_gpu_subset.txt; cat gpurun_out/${T}_pytest_gpu_subset.txt
timeout 300 python bench.py --profile-out gpurun_out/${T}_profile_c180_nh_perop.txt > gpurun_out/${T}_bench_default_1gpu.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_default_1gpu.json
CMD="python bench.py --steps 1 --warmup 0 --kernel-only"
FV3LM_NO_GRAPH=1 timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'KernAD.*S_gradp' -c 2 \
    -o gpurun_out/${T}_KernAD_S_gradp_c180 $CMD > gpurun_out/${T}_ncu_gradp.log 2>&1; tail -2 gpurun_out/${T}_ncu_gradp.log
timeout 200 python bench.py --two-sided --steps 2 --warmup 1 --kernel-only > gpurun_out/${T}_two_sided_c180.json 2> gpurun_out/${T}_two_sided.err; cat gpurun_out/${T}_two_sided_c180.json; tail -2 gpurun_out/${T}_two_sided.err
FV3LM_NO_GRAPH=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 60000 --csv \
    --log-file gpurun_out/${T}_launches_c180.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1; tail -1 gpurun_out/${T}_ncu_launches.log
python - <<'PY'
import csv, collections, sys
rows = collections.OrderedDict()
try:
    with open("gpurun_out/r01t_launches_c180.csv") as fh:
        rd = csv.reader(l for l in fh if not l.startswith("=="))
        hdr = next(rd)
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        for r in rd:
            if len(r) <= iv: continue
            k = r[ik][:110]; v = float(r[iv].replace(",", ""))
            a = rows.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in rows.values())
    with open("gpurun_out/r01t_launches_c180_by_kernel.txt", "w") as out:
        out.write("# ncu --metrics gpu__time_duration.sum, one TL+AD pair at C180 L72 NH (eager, no graph): kernel, launches, total ns, share\n")
        for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            out.write("%-110s %6d %14.0f %6.2f%%\n" % (k, v[0], v[1], 100.0 * v[1] / tot))
    print("launch list:", sum(v[0] for v in rows.values()), "launches,", len(rows), "kernels")
except Exception as e:
    print("launch summary failed:", e)
PY
gzip -f gpurun_out/${T}_launches_c180.csv 2>/dev/null
ls -la gpurun_out | grep ${T}
