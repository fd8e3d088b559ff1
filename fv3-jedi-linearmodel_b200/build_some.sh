#!/bin/bash
# incremental rebuild of selected translation units: build_some.sh cuda|emu|both file1.cu file2.cu ...   (objects of the others must exist from build.sh)
set -e
cd "$(dirname "$0")/csrc"
MODE=$1; shift
SRCS="engine.cu comm.cu mosaic.cu modules.cu tp_fwd.cu tp_rev.cu tp_march.cu csw.cu dsw.cu a2b.cu dyn.cu fvdyn.cu nh.cu capi.cu step_api.cu turb.cu"
OBJ=/tmp/fv3lm_obj_$(id -u)
DEFS=""
if [ -n "$FV3LM_TILE_TY" ]; then DEFS="-DFV3LM_TILE_TY=$FV3LM_TILE_TY"; fi
if [ -n "$FV3LM_TILE_MINBLOCKS" ]; then DEFS="$DEFS -DFV3LM_TILE_MINBLOCKS=$FV3LM_TILE_MINBLOCKS"; fi   # __launch_bounds__ of the tile kernels (A/B runs)
pids=""
for f in "$@"; do
  if [ "$MODE" != "emu" ]; then
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177 $DEFS -c $f -o $OBJ/cuda_${f%.cu}.o &
    pids="$pids $!"
  fi
  if [ "$MODE" != "cuda" ]; then
    g++ -std=c++17 -O2 -fPIC -DFV3LM_HOST_EMU $DEFS -x c++ -c $f -o $OBJ/emu_${f%.cu}.o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
if [ "$MODE" != "emu" ]; then
  OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/cuda_${f%.cu}.o"; done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libfv3lm_b200.so $OBJS
fi
if [ "$MODE" != "cuda" ]; then
  OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/emu_${f%.cu}.o"; done
  g++ -shared -o ../libfv3lm_hostemu.so $OBJS
fi
echo built
