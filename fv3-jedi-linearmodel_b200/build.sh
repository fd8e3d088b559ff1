#!/bin/bash
# Builds (in-tree):
#   libfv3lm_b200.so     product library, nvcc, sm_100a
#   libfv3lm_hostemu.so  TEST-ONLY host emulation of the same stage functors (g++)
set -e
cd "$(dirname "$0")/csrc"
SRCS="engine.cu mosaic.cu modules.cu csw.cu dsw.cu a2b.cu dyn.cu fvdyn.cu nh.cu capi.cu step_api.cu"
OUT=..
if [ "$1" != "emu" ]; then
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
       --expt-relaxed-constexpr -o $OUT/libfv3lm_b200.so $SRCS
fi
if [ "$1" != "cuda" ]; then
  OBJS=""
  for f in $SRCS; do
    g++ -std=c++17 -O2 -fPIC -DFV3LM_HOST_EMU -x c++ -c $f -o /tmp/fv3lm_emu_${f%.cu}.o &
    OBJS="$OBJS /tmp/fv3lm_emu_${f%.cu}.o"
  done
  wait
  g++ -shared -o $OUT/libfv3lm_hostemu.so $OBJS
fi
echo built
