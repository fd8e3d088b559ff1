#include "comm.h"
#include "engine.h"
#include <dlfcn.h>
#include <vector>

namespace fv3lm {

#ifndef FV3LM_HOST_EMU
// NCCL is bound at run time (dlopen) so that the single-GPU library has no hard dependency on it
// and so that a host process which already loaded a libnccl.so.2 (e.g. PyTorch) shares it.
namespace {
struct ncclUniqueIdT { char internal[128]; };
typedef int (*fn_getid)(ncclUniqueIdT*);
typedef int (*fn_init)(void**, int, ncclUniqueIdT, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_void)();
typedef const char* (*fn_err)(int);
struct Nccl {
  void* lib = nullptr;
  fn_getid getid; fn_init init; fn_destroy destroy; fn_sendrecv send, recv; fn_void gstart, gend; fn_err errstr;
  void load() {
    if (lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) throw std::runtime_error("fv3lm: cannot dlopen libnccl.so.2 (multi-GPU runs need NCCL)");
    getid = (fn_getid)dlsym(lib, "ncclGetUniqueId"); init = (fn_init)dlsym(lib, "ncclCommInitRank");
    destroy = (fn_destroy)dlsym(lib, "ncclCommDestroy"); send = (fn_sendrecv)dlsym(lib, "ncclSend"); recv = (fn_sendrecv)dlsym(lib, "ncclRecv");
    gstart = (fn_void)dlsym(lib, "ncclGroupStart"); gend = (fn_void)dlsym(lib, "ncclGroupEnd"); errstr = (fn_err)dlsym(lib, "ncclGetErrorString");
    if (!getid || !init || !send || !recv || !gstart || !gend) throw std::runtime_error("fv3lm: libnccl lacks the point-to-point API");
  }
  void ck(int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string("fv3lm NCCL error in ") + what + ": " + (errstr ? errstr(rc) : "?")); }
} g_nccl;
constexpr int kNcclDouble = 8;   // ncclFloat64
}  // namespace

void Comm::nccl_unique_id(void* out128) { g_nccl.load(); g_nccl.ck(g_nccl.getid((ncclUniqueIdT*)out128), "ncclGetUniqueId"); }
void Comm::init_nccl(const void* id) {
  g_nccl.load();
  ncclUniqueIdT u; memcpy(&u, id, sizeof(u));
  g_nccl.ck(g_nccl.init(&nccl, nranks, u, rank), "ncclCommInitRank");
}
void Comm::destroy() { if (nccl && g_nccl.destroy) g_nccl.destroy(nccl); nccl = nullptr; }
void Comm::exchange(int npeers, const int* peers, double* const* sbuf, const size_t* scount, double* const* rbuf, const size_t* rcount) {
  if (npeers == 0) return;
  if (!nccl) throw std::runtime_error("fv3lm: halo exchange between ranks before fv3lm_comm_init_nccl");
  static const int sync_mode = getenv("FV3LM_SYNC_EXCHANGE") ? atoi(getenv("FV3LM_SYNC_EXCHANGE")) : 0;   // debugging aid
  if (sync_mode & 1) { if (cudaStreamSynchronize(dev::stream()) != cudaSuccess) throw std::runtime_error(std::string("fv3lm: CUDA error before exchange: ") + cudaGetErrorString(cudaGetLastError())); }
  g_nccl.ck(g_nccl.gstart(), "ncclGroupStart");
  for (int p = 0; p < npeers; p++) {
    if (scount[p]) { g_nccl.ck(g_nccl.send(sbuf[p], scount[p], kNcclDouble, peers[p], nccl, dev::stream()), "ncclSend"); bytes_sent += 8.0 * scount[p]; }
    if (rcount[p]) g_nccl.ck(g_nccl.recv(rbuf[p], rcount[p], kNcclDouble, peers[p], nccl, dev::stream()), "ncclRecv");
  }
  g_nccl.ck(g_nccl.gend(), "ncclGroupEnd");
  if (sync_mode & 2) { if (cudaStreamSynchronize(dev::stream()) != cudaSuccess) throw std::runtime_error(std::string("fv3lm: CUDA error after exchange: ") + cudaGetErrorString(cudaGetLastError())); }
  n_exchanges++;
}
#else
void Comm::nccl_unique_id(void*) { throw std::runtime_error("fv3lm: the host emulation build has no NCCL"); }
void Comm::init_nccl(const void*) { throw std::runtime_error("fv3lm: the host emulation build has no NCCL"); }
void Comm::destroy() {}
void Comm::exchange(int npeers, const int* peers, double* const* sbuf, const size_t* scount, double* const* rbuf, const size_t* rcount) {
  if (npeers == 0) return;
  if (!cb) throw std::runtime_error("fv3lm: halo exchange between ranks without a transport callback");
  cb(cb_user, npeers, peers, sbuf, scount, rbuf, rcount);
  for (int p = 0; p < npeers; p++) bytes_sent += 8.0 * scount[p];
  n_exchanges++;
}
#endif

double Comm::min_over_ranks(double v) {
  if (nranks <= 1) return v;
  const int np = nranks - 1;
  double* buf = (double*)dev::alloc(sizeof(double) * (size_t)(nranks + 1));
  dev::h2d(buf, &v, sizeof(double));
  std::vector<int> peers; std::vector<double*> sb, rb; std::vector<size_t> sc, rc;
  for (int r = 0, k = 0; r < nranks; r++) {
    if (r == rank) continue;
    peers.push_back(r); sb.push_back(buf); rb.push_back(buf + 1 + k); sc.push_back(1); rc.push_back(1); k++;
  }
  exchange(np, peers.data(), sb.data(), sc.data(), rb.data(), rc.data());
  std::vector<double> h(nranks);
  dev::d2h(h.data(), buf, sizeof(double) * (size_t)nranks);     // (synchronises the stream)
  dev::free_(buf);
  double m = h[0];
  for (int k = 1; k < nranks; k++) m = h[k] < m ? h[k] : m;
  return m;
}

}  // namespace fv3lm
