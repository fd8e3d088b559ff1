// shared between capi.cu and step_api.cu
#pragma once
#include "../../include/fv3lm_b200.h"
#include "engine.h"
#include "mosaic.h"
#include "modules.h"
#include "turb.h"
#include <map>
#include <memory>

constexpr int NFIELD = 10;

// CUDA-graph replay of a step program: the op list is static, so after one eager (pool-warming) run the whole
// TL or AD sweep -- stage kernels, memsets, NCCL exchanges -- is captured once and replayed with one launch.
struct StepGraph {
  int state = 0;            // 0: never run, 1: pool warmed, 2: captured
  void* exec = nullptr;     // cudaGraphExec_t
  long long launches = 0;   // kernels per replay
  long long exchanges = 0; double bytes_sent = 0.0;
};

// cubed_to_latlon after step_nl (traj%ua, traj%va): a two-op program on the u / v output arrays of the step program
struct C2lRunner {
  fv3lm::Program P;
  fv3lm::ModuleIO io;
  std::map<std::string, int> id;
  double* ua = nullptr; double* va = nullptr;     // compact results of the last step_nl
};

struct StepRunner {
  StepGraph graph[3];
  fv3lm::Program P;
  fv3lm::ModuleIO io;
  int nf = 8;                                     // prognostic fields: 8 hydrostatic, 10 with w, delz
  std::map<std::string, int> in_id, out_id;
  double* pert[NFIELD];                           // device-resident perturbation / adjoint state (compact)
  double* phis = nullptr;
  std::vector<std::vector<double*>> slots;        // device-resident trajectory window (compact)
  std::vector<fv3lm::TurbLtraj> turb;             // local trajectory of the turbulence scheme, per window slot
  C2lRunner* c2l = nullptr;                       // set by fv3lm_set_c2l
};

struct fv3lm_handle {
  fv3lm_config cfg;
  fv3lm::Device dv;
  fv3lm::Mosaic mo;
  fv3lm::Decomp dc;
  fv3lm::Comm comm;
  std::vector<double> ak, bk;
  std::map<std::string, double*> metric_dev;
  std::map<std::string, std::vector<double>> geo_host;   // grid / agrid lon-lat kept to derive a2b_cw
  std::string err;
  StepRunner* step = nullptr;
  int device = -1;          // CUDA device bound at create; re-asserted on entry to the ABI (a host thread may have switched devices)
};
inline void fv3lm_bind_device(const fv3lm_handle* h) {
#ifndef FV3LM_HOST_EMU
  if (h && h->device >= 0) cudaSetDevice(h->device);
#else
  (void)h;
#endif
}

extern thread_local std::string fv3lm_g_err;

// (entry points that take a handle call fv3lm_bind_device first: see FV3LM_TRY_H)
#define FV3LM_TRY try {
#define FV3LM_TRY_H(h) try { fv3lm_bind_device(h);
#define FV3LM_CATCH(h)                                            \
  }                                                               \
  catch (const std::exception& e) {                               \
    fv3lm_g_err = e.what();                                       \
    if (h) (h)->err = e.what();                                   \
    return 1;                                                     \
  }                                                               \
  return 0;
