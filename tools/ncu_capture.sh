#!/bin/bash
# tools/ncu_capture.sh TAG NAME KERNEL_REGEX [ENV=VAL ...] : one `ncu --set full` capture of the first matching launch (after 3 skipped) of the
# kernel-only bench at the bench workload; writes gpurun_out/${TAG}_ncu_${NAME}.txt (metric summary + dynamic opcode mix) and drops the
# .ncu-rep (reports with imported source are 10-25 MB each; gpurun only brings back 64 MiB)
TAG=$1; NAME=$2; K=$3; shift 3
mkdir -p gpurun_out
REP=/tmp/${TAG}_${NAME}
env FV3LM_NO_GRAPH=1 "$@" timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 3 -c 1 \
    -o $REP python bench.py --kernel-only --steps 1 --warmup 0 > /tmp/${TAG}_${NAME}.log 2>&1
if [ -f $REP.ncu-rep ]; then
  python tools/ncu_summary.py --opcodes $REP.ncu-rep > gpurun_out/${TAG}_ncu_${NAME}.txt 2>&1
  python tools/ncu_summary.py --sass-counts $REP.ncu-rep | gzip > gpurun_out/${TAG}_ncu_${NAME}_sass.csv.gz
  rm -f $REP.ncu-rep
  grep -E "^==|duration|dram__bytes|dram_throughput|sm__throughput|registers|opcode" gpurun_out/${TAG}_ncu_${NAME}.txt | cut -c1-260
else
  echo "no report for $NAME"; tail -2 /tmp/${TAG}_${NAME}.log | cut -c1-200
fi
