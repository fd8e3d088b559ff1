// Build-mode switch.  The product library (libfv3lm_b200.so) is compiled by nvcc for
// sm_100a and runs every kernel on the GPU.  A second, TEST-ONLY build of the same
// stage functors (FV3LM_HOST_EMU, plain g++) lets the CPU test-suite check stage
// arithmetic without a GPU; it is a different .so that the product never loads.
#pragma once
#include <stddef.h>
#ifdef FV3LM_HOST_EMU
#define HD inline
#define DEV inline
#define GLOBAL
#define LDG(p) (*(p))
#else
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define GLOBAL __global__
// read-only data (metrics, stage inputs): ld.global.nc lets the compiler keep / reorder loads across the stores of the stage
#define LDG(p) __ldg(p)
#endif

// blockIdx.z = tile * nk + level -> (tile, level).  On the device the quotient comes from a float multiply ((z + 0.5) / nk is at least
// 0.5 / nk away from an integer, z < 2^16, nk <= 128: the rounding error cannot change the truncation), a handful of instructions
// instead of the ~20 of an integer division by a run-time divisor.
DEV void split_z(int z, int nk, int& tile, int& kk) {
#ifdef FV3LM_HOST_EMU
  tile = z / nk;
#else
  tile = (int)(((float)z + 0.5f) * __frcp_rn((float)nk));
#endif
  kk = z - tile * nk;
}
