#!/bin/bash
# round 2, closing measurements on one GPU: whole gpu suite (no -x) with the parity-error table, the default bench line, the reference arm,
# the ncu launch list of one TL+AD pair and `ncu --set full` summaries of the dominant kernels
mkdir -p gpurun_out
T=r02m
( time FV3LM_PARITY_OUT=gpurun_out/${T}_parity_errors_gpu.json python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/${T}_profile_c180_nh_perop.txt > gpurun_out/${T}_bench_default_1gpu.json 2> gpurun_out/${T}_bench_default_1gpu.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_default_1gpu.json; tail -c 300 gpurun_out/${T}_bench_default_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
echo "reference rc=$?"; cut -c1-600 gpurun_out/${T}_bench_reference_arm.json
FV3LM_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 9000 --csv --log-file gpurun_out/${T}_ncu_launches_c180_pair.csv \
    python bench.py --kernel-only --steps 1 --warmup 0 > gpurun_out/${T}_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections, gzip, shutil
rows = list(csv.reader(open('gpurun_out/r02m_ncu_launches_c180_pair.csv', errors='replace')))
hdr = next((r for r in rows if 'Kernel Name' in r), None)
if hdr:
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.Counter(); cnt = collections.Counter()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) > iv:
            try: v = float(r[iv].replace(',', ''))
            except ValueError: continue
            agg[r[ik]] += v; cnt[r[ik]] += 1
    tot = sum(agg.values())
    with open('gpurun_out/r02m_ncu_launches_c180_pair_by_kernel.txt', 'w') as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none: %d launches of one TL+AD pair (kernel-only bench, --steps 1), total %.1f ms; kernel, launches, ms, share\n" % (sum(cnt.values()), tot / 1e6))
        for k, v in agg.most_common(60):
            fh.write("%-110s %6d %10.3f %6.2f%%\n" % (k[:110], cnt[k], v / 1e6, 100.0 * v / tot))
    print(open('gpurun_out/r02m_ncu_launches_c180_pair_by_kernel.txt').read()[:1500])
with open('gpurun_out/r02m_ncu_launches_c180_pair.csv', 'rb') as f, gzip.open('gpurun_out/r02m_ncu_launches_c180_pair.csv.gz', 'wb') as g:
    shutil.copyfileobj(f, g)
PY
rm -f gpurun_out/${T}_ncu_launches_c180_pair.csv
tools/ncu_capture.sh $T TpRevB "kern_tile<fv3lm::ftp::KernTpRev<.int.0"
tools/ncu_capture.sh $T TpRevA "kern_tile<fv3lm::ftp::KernTpRev<.int.1"
tools/ncu_capture.sh $T TL_TpB "kern_tile<fv3lm::ftp::KernTpB<fv3lm::Dual"
tools/ncu_capture.sh $T NL_TpB "kern_tile<fv3lm::ftp::KernTpB<double"
tools/ncu_capture.sh $T TL_TpA "kern_tile<fv3lm::ftp::KernTpA<fv3lm::Dual"
tools/ncu_capture.sh $T ColTL_tri "KernColTL<fv3lm::S_tri"
tools/ncu_capture.sh $T TL_rs_pe "KernTL<fv3lm::S_rs_pe>"
du -sh gpurun_out
