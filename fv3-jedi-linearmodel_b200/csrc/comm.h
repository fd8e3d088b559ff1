// Inter-rank transport of the halo exchanges.  One process per GPU; the product build moves
// the packed halo strips with NCCL point-to-point groups on the library's stream (NVLink 5 /
// NVSwitch); the TEST-ONLY host emulation build calls back into the test harness (gloo).
// Replaces FMS mpp_update_domains / mpp_get_boundary and their adjoints
// (model_tlmadm/fv_mp_tlm.F90:420-852, fv_mp_adm.F90:488-725, 2581-2851).
#pragma once
#include <stddef.h>
#include <vector>

namespace fv3lm {

typedef void (*ExchangeCallback)(void* user, int npeers, const int* peers, double* const* sbuf, const size_t* scount,
                                 double* const* rbuf, const size_t* rcount);

struct Comm {
  int rank = 0, nranks = 1;
  void* nccl = nullptr;            // ncclComm_t
  ExchangeCallback cb = nullptr;   // host-emulation transport
  void* cb_user = nullptr;
  long long n_exchanges = 0;
  double bytes_sent = 0.0;
  // every rank calls this with its peer lists (matching counts on both sides); stream ordered
  void exchange(int npeers, const int* peers, double* const* sbuf, const size_t* scount, double* const* rbuf, const size_t* rcount);
  double min_over_ranks(double v);   // tiny collective built on exchange(): decisions that shape the sweep must agree on all ranks
  void init_nccl(const void* unique_id_128);
  void destroy();
  static void nccl_unique_id(void* out128);
};

}  // namespace fv3lm
