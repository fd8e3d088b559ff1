#!/bin/bash
# Builds (in-tree):
#   libfv3lm_b200.so     product library, nvcc, sm_100a
#   libfv3lm_hostemu.so  TEST-ONLY host emulation of the same stage functors (g++)
# usage: build.sh [cuda|emu]   (default: both).  Objects are compiled in parallel.
set -e
cd "$(dirname "$0")/csrc"
SRCS="engine.cu comm.cu mosaic.cu modules.cu tp_fwd.cu tp_rev.cu tp_march.cu csw.cu dsw.cu a2b.cu dyn.cu fvdyn.cu nh.cu capi.cu step_api.cu turb.cu"
OUT=..
OBJ=/tmp/fv3lm_obj_$(id -u)
DEFS=""
if [ -n "$FV3LM_TILE_TY" ]; then DEFS="-DFV3LM_TILE_TY=$FV3LM_TILE_TY"; fi
if [ -n "$FV3LM_TILE_MINBLOCKS" ]; then DEFS="$DEFS -DFV3LM_TILE_MINBLOCKS=$FV3LM_TILE_MINBLOCKS"; fi   # __launch_bounds__ of the tile kernels (A/B runs)    # tile height of the shared-memory-tile kernels (csrc/fused_tp.h)
mkdir -p $OBJ
pids=""
if [ "$1" != "emu" ]; then
  for f in $SRCS; do
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
         --expt-relaxed-constexpr -diag-suppress 177 $DEFS -c $f -o $OBJ/cuda_${f%.cu}.o &
    pids="$pids $!"
  done
fi
if [ "$1" != "cuda" ]; then
  for f in $SRCS; do
    g++ -std=c++17 -O2 -fPIC -DFV3LM_HOST_EMU $DEFS -x c++ -c $f -o $OBJ/emu_${f%.cu}.o &
    pids="$pids $!"
  done
fi
for p in $pids; do wait $p; done
if [ "$1" != "emu" ]; then
  OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/cuda_${f%.cu}.o"; done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libfv3lm_b200.so $OBJS
fi
if [ "$1" != "cuda" ]; then
  OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/emu_${f%.cu}.o"; done
  g++ -shared -o $OUT/libfv3lm_hostemu.so $OBJS
fi
echo built
