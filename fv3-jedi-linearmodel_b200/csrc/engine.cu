// Device layer, buffer pool and the program executor (NL / TL / AD sweeps).
#include "engine.h"
#include "comm.h"
#include <algorithm>

namespace fv3lm {
namespace dev {
long long launches = 0;
bool profiling = false;
std::map<std::string, ProfRow> prof;

#ifndef FV3LM_HOST_EMU
static cudaStream_t g_stream = nullptr;
cudaStream_t stream() {
  if (!g_stream) {
    if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess)
      throw std::runtime_error("fv3lm: cannot create a CUDA stream (no usable GPU; there is no CPU fallback)");
  }
  return g_stream;
}
static void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("fv3lm CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
void* alloc(size_t bytes) { void* p = nullptr; ck(cudaMalloc(&p, bytes), "cudaMalloc"); return p; }
void free_(void* p) { if (p) cudaFree(p); }
void h2d(void* d, const void* h, size_t b) { ck(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, stream()), "h2d"); }
void d2h(void* h, const void* d, size_t b) { ck(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, stream()), "d2h"); ck(cudaStreamSynchronize(stream()), "d2h sync"); }
void d2d(void* d, const void* s, size_t b) { ck(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, stream()), "d2d"); }
void zero(void* d, size_t b) { ck(cudaMemsetAsync(d, 0, b, stream()), "memset"); }
void sync() { ck(cudaStreamSynchronize(stream()), "sync"); }
void check(const char* what) { ck(cudaGetLastError(), what); }
double free_bytes() { size_t f = 0, t = 0; ck(cudaMemGetInfo(&f, &t), "cudaMemGetInfo"); return (double)f; }
#else
void* alloc(size_t bytes) { void* p = malloc(bytes); if (!p) throw std::runtime_error("malloc"); return p; }
void free_(void* p) { free(p); }
void h2d(void* d, const void* h, size_t b) { memcpy(d, h, b); }
void d2h(void* h, const void* d, size_t b) { memcpy(h, d, b); }
void d2d(void* d, const void* s, size_t b) { memcpy(d, s, b); }
void zero(void* d, size_t b) { memset(d, 0, b); }
void sync() {}
void check(const char*) {}
double free_bytes() { const char* e = getenv("FV3LM_EMU_FREE_BYTES"); return e ? atof(e) : 0.0; }
#endif
}  // namespace dev

// ---------------------------------------------------------------------------------
double* Pool::get(size_t n, bool* fresh) {
  size_t bytes = n * sizeof(double);
  auto& fl = free_[bytes];
  void* p;
  if (fresh) *fresh = false;
  if (!fl.empty()) {
    p = fl.back(); fl.pop_back();
#ifdef FV3LM_HOST_EMU
    // TEST-ONLY build: recycled buffers are poisoned so that any stage consuming a cell its producer
    // never wrote shows up as NaN in the parity tests (the product build hands them out as they are)
    if (fresh) { double* q = (double*)p; for (size_t k = 0; k < n; k++) q[k] = NAN; }
#endif
  } else {
    p = dev::alloc(bytes); bytes_total += bytes;
    // fresh device memory may hold NaN bit patterns: zero it once; recycled buffers keep stale (finite) data
    if (fresh) { dev::zero(p, bytes); *fresh = true; }
  }
  live_[p] = bytes;
  bytes_live += bytes;
  bytes_peak = std::max(bytes_peak, bytes_live);
  return (double*)p;
}
void Pool::put(double* p) {
  if (!p) return;
  auto it = live_.find(p);
  if (it == live_.end()) return;
  free_[it->second].push_back(p);
  bytes_live -= it->second;
  live_.erase(it);
}
void Pool::trim() {
  for (auto& kv : free_) { for (void* p : kv.second) { dev::free_(p); bytes_total -= kv.first; } kv.second.clear(); }
}
Pool::~Pool() {
  trim();
  for (auto& kv : live_) dev::free_(kv.first);
}

// ---------------------------------------------------------------------------------
size_t Program::val_doubles(int id) const {
  const Geom& g = dv->g;
  return (size_t)g.ntile * vals[id].nk * g.slab;
}

void Program::analyse() {
  // activity: propagate from active externals through the op list
  for (auto& v : vals) if (!v.external) v.active = false;
  for (auto& op : ops) {
    if (skipped(op)) continue;
    bool any = false;
    for (int i : op.in) any = any || vals[i].active;
    if (!op.inplace) for (int o : op.out) vals[o].active = any;
  }
  for (auto& v : vals) { v.first_def = -1; v.last_use = -1; }
  for (int n = 0; n < (int)ops.size(); n++) {
    if (skipped(ops[n])) continue;
    for (int i : ops[n].in) vals[i].last_use = n;
    for (int o : ops[n].out) { if (vals[o].first_def < 0) vals[o].first_def = n; vals[o].last_use = std::max(vals[o].last_use, n); }
  }
  // a detached view never carries a perturbation and keeps its target alive (it borrows the storage): the ops reading a view
  // count as uses of the target, which is released with them
  for (auto& v : vals) if (v.alias >= 0) v.active = false;
  for (int n = 0; n < (int)ops.size(); n++) {
    ops[n].hold.clear();
    if (skipped(ops[n])) continue;
    for (int i : ops[n].in)
      if (vals[i].alias >= 0) { ops[n].hold.push_back(vals[i].alias); vals[vals[i].alias].last_use = std::max(vals[vals[i].alias].last_use, n); }
  }
}

void Program::ensure_traj(int id) {
  Value& v = vals[id];
  if (v.alias >= 0) { v.traj = vals[v.alias].traj; return; }
  if (!v.traj) {
    if (v.external) throw std::runtime_error("external value without storage: " + v.name);
    // intermediates are only defined on the range their producer writes and consumers only read
    // that range, so recycled buffers are NOT cleared (a full-array memset per op otherwise)
    bool fresh;
    v.traj = dv->pool.get(val_doubles(id), &fresh);
  }
}
// zero_it: adjoint accumulators must start from zero; tangents (TL) are fully defined by their producer
void Program::ensure_pert(int id, bool zero_it) {
  Value& v = vals[id];
  if (!v.active) return;
  if (v.pert) return;
  if (v.external) throw std::runtime_error("external active value without pert storage: " + v.name);
  bool fresh;
  v.pert = dv->pool.get(val_doubles(id), &fresh);
  if (zero_it && !fresh) dev::zero(v.pert, val_doubles(id) * sizeof(double));
#ifdef FV3LM_HOST_EMU
  if (zero_it) dev::zero(v.pert, val_doubles(id) * sizeof(double));
#endif
}
void Program::release(int id) {
  Value& v = vals[id];
  if (v.external) return;
  if (v.alias >= 0) { v.traj = nullptr; return; }   // borrowed storage
  if (v.traj) { dv->pool.put(v.traj); v.traj = nullptr; }
  if (v.pert) { dv->pool.put(v.pert); v.pert = nullptr; }
}

void Program::check_status_flags() {
  for (const StatusFlag& f : status_flags) {
    double v = 0.0;
    dev::d2h(&v, f.flag, sizeof(double));
    if (v != 0.0) throw std::runtime_error("fv3lm: " + f.what + " " + std::to_string((long long)v));
  }
}

void Program::run_op(Op& op, int mode) {
  if (op.tl_only && mode == MODE_NL) return;          // perturbation-scheme chain: nothing of it is needed by a trajectory sweep
  for (int i : op.in) if (vals[i].alias >= 0) vals[i].traj = vals[vals[i].alias].traj;
  static const bool trace = getenv("FV3LM_TRACE") != nullptr;      // debugging aid: the op sequence on stderr
  if (trace) fprintf(stderr, "fv3lm op %-24s mode %d\n", op.name.c_str(), mode);
#ifdef FV3LM_HOST_EMU
  static const bool nancheck = getenv("FV3LM_NANCHECK") != nullptr;
  if (nancheck && mode == MODE_AD) {
    op.run(*this, op, mode);
    for (int i : op.in) {
      Value& v = vals[i];
      if (!v.pert) continue;
      size_t n = val_doubles(i), bad = 0, first = 0;
      for (size_t q = 0; q < n; q++) if (v.pert[q] != v.pert[q]) { if (!bad) first = q; bad++; }
      if (bad) {
        const Geom& g = dv->g;
        size_t pos = first % g.slab;
        fprintf(stderr, "NANCHECK ad op %s: adjoint of input %s has %zu NaN, first at tile/k %zu jj=%zu ii=%zu\n", op.name.c_str(), v.name.c_str(), bad,
                first / g.slab, pos / g.pitch, pos % g.pitch);
        static int cnt = 0; if (++cnt > 6) exit(3);
      }
    }
    return;
  }
#endif
#ifndef FV3LM_HOST_EMU
  static const bool synccheck = getenv("FV3LM_SYNC_CHECK") != nullptr;   // debugging aid: localise a faulting kernel
  if (synccheck) {
    op.run(*this, op, mode);
    cudaError_t e = cudaStreamSynchronize(dev::stream());
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error("fv3lm: CUDA error after op '" + op.name + "' (mode " + std::to_string(mode) + "): " + cudaGetErrorString(e));
    return;
  }
#endif
  if (!dev::profiling) { op.run(*this, op, mode); return; }
#ifndef FV3LM_HOST_EMU
  static cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  cudaEventRecord(e0, dev::stream());
  op.run(*this, op, mode);
  cudaEventRecord(e1, dev::stream());
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
#else
  op.run(*this, op, mode);
  float ms = 0;
#endif
  // algorithmic bytes: distinct arrays read + written, halo excluded (SURVEY 8(d) counting rule)
  const Geom& g = dv->g;
  auto cells = [&](int id) { return (double)g.ntile * vals[id].nk * g.ie * g.je * 8.0; };
  std::vector<int> ins = op.in; std::sort(ins.begin(), ins.end()); ins.erase(std::unique(ins.begin(), ins.end()), ins.end());
  double b = 0.0;
  if (op.inplace) { b = 0.0; }
  else if (mode == MODE_AD) { for (int i : ins) b += cells(i) * (vals[i].active ? 3.0 : 1.0); for (int o : op.out) if (vals[o].active) b += cells(o); }
  else { const double f = (mode == MODE_TL) ? 2.0 : 1.0; for (int i : ins) b += cells(i) * (vals[i].active ? f : 1.0); for (int o : op.out) b += cells(o) * (vals[o].active ? f : 1.0); }
  const char* tag = mode == MODE_TL ? "tl:" : mode == MODE_AD ? "ad:" : "nl:";
  dev::ProfRow& r = dev::prof[std::string(tag) + op.name];
  r.n++; r.ms += ms; r.alg_bytes += b;
}

// Fraction of the free device memory the adjoint may plan with when it decides how many forward segments to keep.  The plan also sets aside
// the largest segment for the segment being reversed (see below); at C180 on one 183 GB B200 that keeps 7 of 9 acoustic segments.
static const double kAdStoreFraction = 0.95;

// The adjoint can skip the per-segment recomputation (one whole nonlinear sweep) when the complete forward sweep
// fits in device memory -- the case once the cube is sharded over several GPUs.
bool Program::ad_fits_store_all() {
  if (ad_store_all_cached >= 0) return ad_store_all_cached != 0;     // decided once: the sweep structure must not change between runs
  double need = 0.0;
  for (int id = 0; id < (int)vals.size(); id++) {
    if (vals[id].external || vals[id].alias >= 0) continue;
    need += (double)val_doubles(id) * 8.0 * (vals[id].active ? 2.0 : 1.0);
  }
  double budget = dv->ad_store_budget;
  if (const char* e = getenv("FV3LM_AD_STORE_BUDGET")) budget = atof(e);   // tests force either path
  if (budget < 0.0) budget = kAdStoreFraction * (dev::free_bytes() + (double)(dv->pool.bytes_total - dv->pool.bytes_live));
  if (dv->comm) budget = dv->comm->min_over_ranks(budget);     // every rank must take the same path (same exchange sequence)
  ad_store_all_cached = (need <= budget) ? 1 : 0;
  return ad_store_all_cached != 0;
}

// An exception in the middle of a sweep must not leave intermediates bound: a stale non-null `pert` would make the next adjoint
// run skip the zeroing of that accumulator (ensure_pert returns early) and add onto old adjoints.
void Program::run(Mode mode) {
  try { run_sweeps(mode); }
  catch (...) {
    for (int id = 0; id < (int)vals.size(); id++) if (!vals[id].external) release(id);
    throw;
  }
}

void Program::run_sweeps(Mode mode) {
  sweep_kind = (mode == MODE_NL || mode == MODE_TL) ? VAR_FWD : VAR_AD;
  analyse();
  const int nop = (int)ops.size();
  if (mode == MODE_NL || mode == MODE_TL) {
    for (int n = 0; n < nop; n++) {
      Op& op = ops[n];
      if (skipped(op)) continue;
      if (!(op.tl_only && mode == MODE_NL)) {     // (a skipped op still ends the life of its inputs)
        for (int o : op.out) { ensure_traj(o); if (mode == MODE_TL) ensure_pert(o, false); }
        run_op(op, mode);
      }
      // free values whose last use was this op
      for (int i : op.in) if (vals[i].last_use == n) release(i);
      for (int i : op.hold) if (vals[i].last_use == n) release(i);
      for (int o : op.out) if (vals[o].last_use == n) release(o);
    }
  } else if (!seg_start.empty() && !ad_fits_store_all()) {
    // ---- segmented adjoint: checkpoint at segment boundaries, recompute one segment at a time
    // (replaces the reference's global tape, utils/tapenade/adStack.c: only the values that
    // cross a segment boundary stay in HBM, the rest is recomputed in the reverse sweep)
    std::vector<int> seg_of(nop, 0);
    {
      int s = 0; size_t nb = 0;
      for (int n = 0; n < nop; n++) { while (nb < seg_start.size() && seg_start[nb] <= n) { if (seg_start[nb] > 0) s++; nb++; } seg_of[n] = s; }
    }
    const int nseg = seg_of[nop - 1] + 1;
    auto def_seg = [&](int id) { return vals[id].first_def < 0 ? -1 : seg_of[vals[id].first_def]; };
    auto last_seg = [&](int id) { return vals[id].last_use < 0 ? -1 : seg_of[vals[id].last_use]; };
    // How many trailing segments can be kept whole during pass 1 (no recomputation for them)?  Decided once from the
    // device memory that is free beyond what one segment's recomputation + reversal needs.
    if (ad_keep_from < 0) {
      std::vector<double> seg_bytes(nseg, 0.0);
      for (int id = 0; id < (int)vals.size(); id++)
        if (!vals[id].external && vals[id].alias < 0 && vals[id].first_def >= 0) seg_bytes[seg_of[vals[id].first_def]] += (double)val_doubles(id) * 8.0;
      double largest = 0.0; for (double b : seg_bytes) largest = std::max(largest, b);
      double budget = dv->ad_store_budget;
      if (const char* e = getenv("FV3LM_AD_STORE_BUDGET")) budget = atof(e);
      if (budget < 0.0) budget = kAdStoreFraction * (dev::free_bytes() + (double)(dv->pool.bytes_total - dv->pool.bytes_live));
      if (dv->comm) budget = dv->comm->min_over_ranks(budget);   // same decision on every rank
      // reserve for the segment being reversed.  Its adjoints are allocated and released op by op and a recomputed segment reuses what the
      // already reversed ones gave back: measured 6.5 - 7.7 GB beyond the kept segments at C180 (segments of 20.5 GB; pool peak 113 / 132 GB with
      // 5 / 6 acoustic segments kept, profiles/r02i, r02v), so one whole segment is a 2.7x margin (it used to be two)
      budget -= 1.0 * largest;
      ad_keep_from = nseg - 1;                      // the last segment is always kept
      const double budget0 = budget;
      while (ad_keep_from > 0 && seg_bytes[ad_keep_from - 1] <= budget) { budget -= seg_bytes[ad_keep_from - 1]; ad_keep_from--; }
      if (getenv("FV3LM_DEBUG_SEG")) {
        fprintf(stderr, "fv3lm %s: %d segments, budget %.2f GB (after 2 x largest %.2f GB), kept from %d; bytes (GB):", name.c_str(), nseg, budget0 / 1e9, largest / 1e9, ad_keep_from);
        for (double b : seg_bytes) fprintf(stderr, " %.2f", b / 1e9);
        fprintf(stderr, "\n");
      }
    }
    const int keep_from = ad_keep_from;
    // pass 1: plain forward; segment-local values are freed at their last use, boundary-crossing ones stay
    for (int n = 0; n < nop; n++) {
      Op& op = ops[n];
      if (skipped(op)) continue;
      if (!(op.tl_only && seg_of[n] < keep_from)) {        // (else: recomputed with its segment)
        for (int o : op.out) ensure_traj(o);
        // kept segments (at least the last, reversed first) stay whole; their patch ops save what they overwrite
        run_op(op, seg_of[n] >= keep_from ? MODE_ADFWD : MODE_NL);
      }
      if (seg_of[n] >= keep_from) continue;
      for (int i : op.in) if (vals[i].last_use == n && def_seg(i) == seg_of[n]) release(i);
      for (int i : op.hold) if (vals[i].last_use == n && def_seg(i) == seg_of[n]) release(i);
      for (int o : op.out) if (vals[o].last_use == n && def_seg(o) == seg_of[n]) release(o);
    }
    for (int s = nseg - 1; s >= 0; s--) {
      int n0 = 0, n1 = nop - 1;
      while (seg_of[n0] != s) n0++;
      while (seg_of[n1] != s) n1--;
      // recompute the segment keeping every value (patch ops save what they overwrite)
      for (int n = n0; n <= n1 && s < keep_from; n++) {
        Op& op = ops[n];
        if (skipped(op)) continue;
        for (int o : op.out) ensure_traj(o);
        run_op(op, MODE_ADFWD);
      }
      for (int n = n1; n >= n0; n--) {
        Op& op = ops[n];
        if (skipped(op)) continue;
        bool any_out = false;
        for (int o : op.out) if (vals[o].active && vals[o].pert) any_out = true;
        if (any_out || op.inplace) {
          for (int o : op.out) ensure_pert(o, true);
          for (int i : op.in) ensure_pert(i, true);
          run_op(op, MODE_AD);
        }
        if (!op.inplace)
          for (int o : op.out) if (vals[o].first_def == n) release(o);
      }
      (void)last_seg;
    }
    for (int id = 0; id < (int)vals.size(); id++)
      if (!vals[id].external && vals[id].first_def < 0) release(id);
  } else {
    // forward sweep, keep everything ("device checkpoint arena")
    for (int n = 0; n < nop; n++) {
      Op& op = ops[n];
      if (skipped(op)) continue;
      for (int o : op.out) ensure_traj(o);
      run_op(op, MODE_ADFWD);
    }
    // reverse sweep
    for (int n = nop - 1; n >= 0; n--) {
      Op& op = ops[n];
      if (skipped(op)) continue;
      bool any_out = false;
      for (int o : op.out) if (vals[o].active && vals[o].pert) any_out = true;
      if (any_out || op.inplace) {
        for (int o : op.out) ensure_pert(o, true);
        for (int i : op.in) ensure_pert(i, true);
        run_op(op, MODE_AD);
      }
      if (!op.inplace)
        for (int o : op.out) if (vals[o].first_def == n) release(o);
    }
    // inputs that nobody produced (non-external temporaries) are released
    for (int id = 0; id < (int)vals.size(); id++)
      if (!vals[id].external && vals[id].first_def < 0) release(id);
  }
  dev::check(name.c_str());
}

}  // namespace fv3lm
