// fv3lm engine: device-resident fields, stencil "stages", and static programs of
// stages that can be executed in three modes:
//   NL  nonlinear trajectory only
//   TL  trajectory + tangent (dual numbers, value and perturbation in one sweep, the
//       same structure as the reference's Tapenade tangent mode)
//   AD  forward NL sweep that keeps every intermediate in HBM (the "device checkpoint
//       arena" replacing utils/tapenade/adStack.c), then the reverse sweep in
//       GATHER form: each thread owns one input cell and sums the contributions of
//       every output cell that read it; no atomics, fixed summation order.
//
// A Stage is a struct with
//     static constexpr int NI, NO;            number of input / output fields
//     struct P {...};                          POD parameters
//     static constexpr Tap taps[NT];           union of (input, di, dj, dk) it may read
//     template<class X> static void eval(X&, const P&);
// eval is written once against a context X (X::T = double or Dual).
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <stdexcept>
#include <type_traits>
#include <algorithm>
#include "platform.h"
#include "dual.h"

namespace fv3lm {

// ---------------------------------------------------------------------------------
// geometry of the device arrays
// ---------------------------------------------------------------------------------
// A device holds `ntile` SUB-DOMAINS (whole cube tiles for layout 1x1, else the layout(lx,ly)
// pieces of tools/fv_mp_nlm_mod.F90:411-441), all of the same size nxl x nyl.  Stage code
// sees two index frames: the LOCAL frame (compute domain is..ie = 1..nxl, js..je = 1..nyl;
// loop ranges / in_rect) and the TILE-GLOBAL Fortran index x.i, x.j = local + i0/j0
// (cube-edge special cases keyed on i == 1, i == npx, ... exactly like the reference).
constexpr int MAXSUB = 24;
struct Geom {
  int N, npx, npy, ng;  // cells per tile edge; npx = npy = N + 1; halo width (3)
  int is, ie, js, je;   // compute domain in the local frame: 1..nxl, 1..nyl
  int NX, NY, pitch;    // local array extents nxl + 2 ng + 1, nyl + 2 ng + 1 and the row pitch (doubles)
  int ntile;            // sub-domains resident on this device
  int K;                // npz
  int slab;             // pitch * NY
  short i0[MAXSUB], j0[MAXSUB];   // tile-global index = local index + i0 / j0
  short tile_of[MAXSUB];          // cube tile (0..5) of each resident sub-domain
};

// 2-D metric arrays, each [ntile][NY][pitch]; sin_sg/cos_sg 1..4 (+5 for sin) split out.
// Definitions: SURVEY appendix D  (model/fv_grid_utils_nlm.F90, tools/fv_grid_tools_nlm.F90)
#define FV3LM_METRIC_LIST(X)                                                              \
  X(area) X(rarea) X(area_c) X(rarea_c) X(dx) X(dy) X(rdx) X(rdy) X(dxa) X(dya) X(rdxa)   \
  X(rdya) X(dxc) X(dyc) X(rdxc) X(rdyc) X(cosa) X(sina) X(rsina) X(cosa_u) X(sina_u)      \
  X(rsin_u) X(cosa_v) X(sina_v) X(rsin_v) X(cosa_s) X(rsin2) X(sin_sg1) X(sin_sg2)        \
  X(sin_sg3) X(sin_sg4) X(cos_sg1) X(cos_sg2) X(cos_sg3) X(cos_sg4) X(divg_u) X(divg_v)   \
  X(del6_u) X(del6_v) X(f0) X(fC) X(agrid_lon) X(agrid_lat) X(grid_lon) X(grid_lat)       \
  X(edge_w) X(edge_e) X(edge_s) X(edge_n) X(edge_vect_w) X(edge_vect_e) X(edge_vect_s)    \
  X(edge_vect_n) X(a2b_cw) X(ppmw_x0) X(ppmw_x1) X(ppmw_x2) X(ppmw_x3) X(ppmw_y0) X(ppmw_y1)  \
  X(ppmw_y2) X(ppmw_y3)
// (the edge_* arrays are 1-D, stored in row 0 of a slab so they share the layout; a2b_cw holds the 4 x 3
// extrap_corner weights of a2b_ord4, derived by the library from grid / agrid once they are all uploaded; ppmw_x0..3 / ppmw_y0..3 the four
// weights of the PPM edge value on the cube edges (faces 1 and npx / npy: two one-sided extrapolations, tp_core_tlm.F90:2405-2417), derived
// from dxa / dya when those are uploaded, so that the hot kernels load four numbers there instead of dividing)

struct Metrics {
#define X(n) const double* n;
  FV3LM_METRIC_LIST(X)
#undef X
  double da_min, da_min_c;
};

struct Tap { int f, di, dj, dk; };
constexpr int KLAST = 9999;   // dk value meaning "the last level of that field" (absolute)

template <int N> struct FArr {
  double* p[N > 0 ? N : 1];
  int nk[N > 0 ? N : 1];
};

// ---------------------------------------------------------------------------------
// contexts
// ---------------------------------------------------------------------------------
struct CtxBase {
  Geom g; Metrics m;
  int ii, jj, kk, tile;   // array column/row, level, resident sub-domain
  int i, j;               // Fortran tile-global indices
  int il, jl;             // local-frame indices (compute domain 1..nxl, 1..nyl)
  int pos;                // jj * pitch + ii : all offsets are 32-bit (a rank's array has < 2^31 elements)
  // i0 / j0: origin of the sub-domain, read by the caller from the kernel parameter (constant bank);
  // indexing the per-thread copy of Geom dynamically would force it into local memory
  DEV void setpos(int ii_, int jj_, int kk_, int tile_, int i0_, int j0_) {
    ii = ii_; jj = jj_; kk = kk_; tile = tile_; il = ii_ - (g.ng - 1); jl = jj_ - (g.ng - 1);
    i = il + i0_; j = jl + j0_;
    pos = jj_ * g.pitch + ii_;
  }
  // metric at relative offset
  // TEST-ONLY build: every stencil access is bounds-checked against the sub-domain array (an out-of-slab tap is
  // silent on the GPU as long as it stays inside the allocation)
  DEV void bounds(int di, int dj, const char* what) const {
#ifdef FV3LM_HOST_EMU
    if (ii + di < 0 || ii + di >= g.NX || jj + dj < 0 || jj + dj >= g.NY)
      throw std::runtime_error(std::string("fv3lm emu: out-of-slab ") + what + " access at ii=" + std::to_string(ii) + " jj=" + std::to_string(jj) +
                               " di=" + std::to_string(di) + " dj=" + std::to_string(dj));
#else
    (void)di; (void)dj; (void)what;
#endif
  }
  DEV double M(const double* a, int di = 0, int dj = 0) const {
    bounds(di, dj, "metric");
    return LDG(a + (tile * g.slab + pos + dj * g.pitch + di));
  }
  // metric at absolute Fortran index
  DEV double Mabs(const double* a, int ai, int aj) const {
    return LDG(a + (tile * g.slab + (aj - (j - jl) + g.ng - 1) * g.pitch + (ai - (i - il) + g.ng - 1)));
  }
  // 1-D edge array (stored whole-tile in the head of the slab), tile-global Fortran index
  DEV double M1(const double* a, int ai) const { return LDG(a + (tile * g.slab + (ai + g.ng - 1))); }
  // loop ranges are given in the local frame
  DEV bool in_rect(int i0, int i1, int j0, int j1) const { return il >= i0 && il <= i1 && jl >= j0 && jl <= j1; }
  // tile-global rectangle (whole-tile operators such as a2b_ord4)
  DEV bool in_tile(int i0, int i1, int j0, int j1) const { return i >= i0 && i <= i1 && j >= j0 && j <= j1; }
  DEV int off(int nkf, int di, int dj, int dk) const {
    int k;
    if (dk == 0) k = kk < nkf ? kk : nkf - 1;          // (2-D fields read inside a 3-D launch)
    else if (dk == KLAST) k = nkf - 1;
    else { k = kk + dk; k = k > nkf - 1 ? nkf - 1 : k; k = k < 0 ? 0 : k; }
    bounds(di, dj, "field");
    return (tile * nkf + k) * g.slab + pos + dj * g.pitch + di;
  }
};

template <class S> struct CtxNL : CtxBase {
  using T = double;
  static constexpr int mode = 0;
  FArr<S::NI> in_; FArr<S::NO> out_;
  DEV T in(int f, int di = 0, int dj = 0, int dk = 0) const { return LDG(in_.p[f] + off(in_.nk[f], di, dj, dk)); }
  DEV void out(int o, T v) const { out_.p[o][off(out_.nk[o], 0, 0, 0)] = v; }
};

template <class S> struct CtxTL : CtxBase {
  using T = Dual;
  static constexpr int mode = 1;
  FArr<S::NI> in_, ind_; FArr<S::NO> out_, outd_;
  DEV T in(int f, int di = 0, int dj = 0, int dk = 0) const {
    const int o = off(in_.nk[f], di, dj, dk);
    return Dual(LDG(in_.p[f] + o), ind_.p[f] ? LDG(ind_.p[f] + o) : 0.0);
  }
  DEV void out(int o, T v) const {
    const int q = off(out_.nk[o], 0, 0, 0);
    out_.p[o][q] = v.v;
    if (outd_.p[o]) outd_.p[o][q] = v.d;
  }
};

template <class S> struct CtxAD : CtxBase {
  using T = Dual;
  static constexpr int mode = 2;
  FArr<S::NI> in_; FArr<S::NO> outad_;
  int sf, sdi, sdj, sdk;  // the seeded tap
  double acc;
  DEV T in(int f, int di = 0, int dj = 0, int dk = 0) const {
    return Dual(LDG(in_.p[f] + off(in_.nk[f], di, dj, dk)), (f == sf && di == sdi && dj == sdj && dk == sdk) ? 1.0 : 0.0);
  }
  DEV void out(int o, T v) {
    if (outad_.p[o]) {
      // an output cell nobody consumed has a zero adjoint; its value may be computed from cells that were
      // never written (generous loop ranges), so it must not enter the sum as 0 * garbage
      const double a = LDG(outad_.p[o] + off(outad_.nk[o], 0, 0, 0));
      if (a != 0.0) acc += v.d * a;
    }
  }
};

// ---------------------------------------------------------------------------------
// kernel functors (called for every (ii, jj, z = tile * nk + k) of the padded arrays)
// ---------------------------------------------------------------------------------
template <class S> struct KernNL {
  typename S::P p; Geom g; Metrics m; FArr<S::NI> in; FArr<S::NO> out; int nk;
  // (level and sub-domain come from the launcher: an integer division per thread costs ~20 instructions, as much as a simple stage)
  DEV void operator()(int ii, int jj, int kk, int tile) const {
    CtxNL<S> x; x.g = g; x.m = m; x.in_ = in; x.out_ = out;
    x.setpos(ii, jj, kk, tile, g.i0[tile], g.j0[tile]);
    S::eval(x, p);
  }
  // levels k0 .. k1-1 of one (i, j): the context is built once and the loop is not unrolled, so that the level-invariant work of
  // the stage (index set-up, rectangle / cube-edge tests, the loads of the 2-D metrics) is hoisted out of it
  DEV void run(int ii, int jj, int k0, int k1, int tile) const {
    CtxNL<S> x; x.g = g; x.m = m; x.in_ = in; x.out_ = out;
    x.setpos(ii, jj, k0, tile, g.i0[tile], g.j0[tile]);
#pragma unroll 1
    for (int k = k0; k < k1; k++) { x.kk = k; S::eval(x, p); }
  }
};
template <class S> struct KernTL {
  typename S::P p; Geom g; Metrics m; FArr<S::NI> in, ind; FArr<S::NO> out, outd; int nk;
  DEV void operator()(int ii, int jj, int kk, int tile) const {
    CtxTL<S> x; x.g = g; x.m = m; x.in_ = in; x.ind_ = ind; x.out_ = out; x.outd_ = outd;
    x.setpos(ii, jj, kk, tile, g.i0[tile], g.j0[tile]);
    S::eval(x, p);
  }
  DEV void run(int ii, int jj, int k0, int k1, int tile) const {
    CtxTL<S> x; x.g = g; x.m = m; x.in_ = in; x.ind_ = ind; x.out_ = out; x.outd_ = outd;
    x.setpos(ii, jj, k0, tile, g.i0[tile], g.j0[tile]);
#pragma unroll 1
    for (int k = k0; k < k1; k++) { x.kk = k; S::eval(x, p); }
  }
};

// ---- taps grouped by stencil offset --------------------------------------------------------------
// The reverse sweep needs d(out)/d(tap) for every tap.  Taps of DIFFERENT input fields at the SAME offset
// (di, dj, dk) are read by the same output cell, so one evaluation of the stage there with a multi-seed
// dual number (DualN<M>, M = fields at that offset) yields all M derivatives at once: the number of stage
// evaluations per thread drops from NT (taps) to the number of distinct offsets.
template <class S> struct TapInfo {
  static constexpr bool same(int a, int b) {
    return S::taps[a].di == S::taps[b].di && S::taps[a].dj == S::taps[b].dj && S::taps[a].dk == S::taps[b].dk;
  }
  static constexpr bool leader(int n) { for (int m = 0; m < n; m++) if (same(m, n)) return false; return true; }
  static constexpr int count(int n) { int c = 0; for (int m = 0; m < S::NT; m++) if (same(m, n)) c++; return c; }
  // input field of the m-th tap that shares the offset of tap n
  static constexpr int field(int n, int m) {
    int c = 0;
    for (int q = 0; q < S::NT; q++) if (same(q, n)) { if (c == m) return S::taps[q].f; c++; }
    return -1;
  }
};

template <class S, int n> struct CtxADN : CtxBase {
  static constexpr int NM = TapInfo<S>::count(n);
  using T = DualN<NM>;
  static constexpr int mode = 2;
  FArr<S::NI> in_; FArr<S::NO> outad_;
  double acc[NM];
  template <int m> DEV static void seed(T& r, int f) {
    if constexpr (m < NM) {
      constexpr int fm = TapInfo<S>::field(n, m);
      r.d[m] = (f == fm) ? 1.0 : 0.0;
      seed<m + 1>(r, f);
    }
  }
  DEV T in(int f, int di = 0, int dj = 0, int dk = 0) const {
    constexpr Tap t = S::taps[n];
    T r(LDG(in_.p[f] + off(in_.nk[f], di, dj, dk)));
    if (di == t.di && dj == t.dj && dk == t.dk) seed<0>(r, f);
    return r;
  }
  DEV void out(int o, const T& v) {
    if (outad_.p[o]) {
      // an output cell nobody consumed has a zero adjoint; its value may be computed from cells that were
      // never written (generous loop ranges), so it must not enter the sum as 0 * garbage
      const double a = LDG(outad_.p[o] + off(outad_.nk[o], 0, 0, 0));
      if (a != 0.0) {
#pragma unroll
        for (int m = 0; m < NM; m++) acc[m] += v.d[m] * a;
      }
    }
  }
};

template <class S, int n, int m> struct AdScatter {
  DEV static void run(const double* src, double* acc) {
    if constexpr (m < TapInfo<S>::count(n)) {
      constexpr int fm = TapInfo<S>::field(n, m);
      acc[fm] += src[m];
      AdScatter<S, n, m + 1>::run(src, acc);
    }
  }
};

template <class S, int n> struct AdTaps {
  template <class K> DEV static void eval_at(const K& kn, int oi, int oj, int ok, int tile, double* acc) {
    CtxADN<S, n> x; x.g = kn.g; x.m = kn.m; x.in_ = kn.in; x.outad_ = kn.outad;
#pragma unroll
    for (int m = 0; m < CtxADN<S, n>::NM; m++) x.acc[m] = 0.0;
    x.setpos(oi, oj, ok, tile, kn.g.i0[tile], kn.g.j0[tile]);
    S::eval(x, kn.p);
    AdScatter<S, n, 0>::run(x.acc, acc);
  }
  template <class K> DEV static void run(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    if constexpr (n < S::NT) {
      if constexpr (TapInfo<S>::leader(n)) {
        constexpr Tap t = S::taps[n];
        int oi = ii - t.di, oj = jj - t.dj;
        if (oi >= 0 && oi < kn.g.NX && oj >= 0 && oj < kn.g.NY) {
          if (t.dk == KLAST) {
            // every output level reads the last level of this input: its owner sums over them
            if (kk == kn.in.nk[t.f] - 1)
              for (int ok = 0; ok < kn.nk_fwd; ok++) eval_at(kn, oi, oj, ok, tile, acc);
          } else {
            int ok = kk - t.dk;
            if (ok >= 0 && ok < kn.nk_fwd) eval_at(kn, oi, oj, ok, tile, acc);
          }
        }
      }
      AdTaps<S, n + 1>::run(kn, ii, jj, kk, tile, acc);
    }
  }
};

// a stage may replace the generic seeded evaluations by a hand-derived gather adjoint:
//   static constexpr bool custom_ad = true;  template<class K> static void adjoint(kn, ii, jj, kk, tile, acc)
template <class S, class = void> struct HasCustomAd { static constexpr bool value = false; };
template <class S> struct HasCustomAd<S, std::enable_if_t<S::custom_ad>> { static constexpr bool value = true; };

template <class S> struct KernAD {
  typename S::P p; Geom g; Metrics m; FArr<S::NI> in, inad; FArr<S::NO> outad; int nk, nk_fwd;
  DEV void operator()(int ii, int jj, int kk, int tile) const {
    double acc[S::NI], old[S::NI];
    // the accumulators' current values are requested BEFORE the evaluations: issued as `p[o] += acc` at the end, every field's load would wait
    // behind the previous field's store (the compiler cannot prove the adjoint arrays distinct) -- one exposed DRAM latency per input field
#pragma unroll
    for (int f = 0; f < S::NI; f++) {
      acc[f] = 0.0; old[f] = 0.0;
      if (inad.p[f] && kk < inad.nk[f]) old[f] = inad.p[f][(tile * inad.nk[f] + kk) * g.slab + jj * g.pitch + ii];
    }
    if constexpr (HasCustomAd<S>::value) S::adjoint(*this, ii, jj, kk, tile, acc);
    else AdTaps<S, 0>::run(*this, ii, jj, kk, tile, acc);
    // the same value may be bound to two inputs (del_fx reads d2 twice): their contributions go to one accumulator
    unsigned dup = 0u;
#pragma unroll
    for (int f = S::NI - 1; f >= 1; f--) {
#pragma unroll
      for (int h = 0; h < f; h++)
        if (!(dup & (1u << f)) && inad.p[h] == inad.p[f] && inad.p[f]) { acc[h] += acc[f]; dup |= 1u << f; }
    }
#pragma unroll
    for (int f = 0; f < S::NI; f++) {
      if (!(dup & (1u << f)) && inad.p[f] && kk < inad.nk[f]) {
        const int o = (tile * inad.nk[f] + kk) * g.slab + jj * g.pitch + ii;
        inad.p[f][o] = old[f] + acc[f];
      }
    }
  }
  DEV void run(int ii, int jj, int k0, int k1, int tile) const {
#pragma unroll 1
    for (int k = k0; k < k1; k++) (*this)(ii, jj, k, tile);
  }
};

// ---------------------------------------------------------------------------------
// device layer
// ---------------------------------------------------------------------------------
namespace dev {
void* alloc(size_t bytes);
void free_(void* p);
void h2d(void* d, const void* h, size_t bytes);
void d2h(void* h, const void* d, size_t bytes);
void d2d(void* d, const void* s, size_t bytes);
void zero(void* d, size_t bytes);
void sync();
void check(const char* what);
double free_bytes();         // free device memory
extern long long launches;   // number of kernels launched (bench.py gpu_launches)
// optional per-op profile (CUDA events around every op; serialises, so only used in a
// dedicated profiling pass, never inside a timed region)
struct ProfRow { long long n = 0; double ms = 0.0, alg_bytes = 0.0; };
extern bool profiling;
extern std::map<std::string, ProfRow> prof;
#ifndef FV3LM_HOST_EMU
cudaStream_t stream();
#endif
}  // namespace dev

#ifndef FV3LM_HOST_EMU
template <class F> GLOBAL void kern3d(F f, int nx, int ny) {
  int ii = blockIdx.x * blockDim.x + threadIdx.x;
  int jj = blockIdx.y * blockDim.y + threadIdx.y;
  if (ii < nx && jj < ny) f(ii, jj, (int)blockIdx.z);
}
template <class F> void launch3d(const F& f, int nx, int ny, int nz) {
  if (nz <= 0) return;
  dim3 b(32, 8, 1), gr((nx + 31) / 32, (ny + 7) / 8, nz);
  kern3d<F><<<gr, b, 0, dev::stream()>>>(f, nx, ny);
  dev::launches++;
}
// stencil stages: every thread handles KPT consecutive levels of one (i, j): the 2-D metric loads, the
// cube-edge branch conditions and the index set-up are level-invariant and get shared across them
// (an unrolled variant was tried in round 1: ptxas did not share work across the unrolled levels; the loop in F::run is not unrolled
// and the context is built once, so loop-invariant code motion does the sharing.)  FV3LM_KPT overrides the default.
inline int stage_kpt() { static const int k = getenv("FV3LM_KPT") ? std::max(1, atoi(getenv("FV3LM_KPT"))) : 4; return k; }
// blockIdx.z = tile * nkc + level chunk.  The quotient comes from a float multiply: (z + 0.5) / nkc is at least 0.5 / nkc away from
// an integer and z < 2^16, so the rounding error of the product (< 1e-5) cannot change the truncation -- exact, 3 instructions
// instead of the ~20 of an integer division by a run-time divisor.
DEV int fast_div_small(int z, float inv) { return (int)(((float)z + 0.5f) * inv); }
// ncell > 0: the cells of a level are packed onto consecutive threads (a warp may straddle two rows): on small sub-domains a 32 x 8 block
// grid wastes lanes (97-cell rows of C180 on 8 GPUs fill 3 1/32 blocks: 76 % lane efficiency), the packed form none.  The row comes from
// the same exact float quotient as the level (cells per level < 2^16).  One kernel serves both mappings (a second instantiation of every
// stage kernel doubled the build time).
template <class F> GLOBAL void kern_stage(const __grid_constant__ F f, int nx, int ny, int nk, int nkc, float inv_nkc, int kpt, int ncell, float inv_nx) {
  int ii, jj;
  if (ncell > 0) {
    const int id = blockIdx.x * 256 + threadIdx.y * 32 + threadIdx.x;
    if (id >= ncell) return;
    jj = fast_div_small(id, inv_nx); ii = id - jj * nx;
  } else {
    ii = blockIdx.x * blockDim.x + threadIdx.x;
    jj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ii >= nx || jj >= ny) return;
  }
  const int tile = fast_div_small(blockIdx.z, inv_nkc), k0 = (blockIdx.z - tile * nkc) * kpt;
  const int k1 = k0 + kpt < nk ? k0 + kpt : nk;
  if (kpt == 1) f(ii, jj, k0, tile);
  else f.run(ii, jj, k0, k1, tile);
}
template <class F> void launch_stage(const F& f, int nx, int ny, int ntile, int nk) {
  if (ntile * nk <= 0) return;
  const int kpt = stage_kpt(), nkc = (nk + kpt - 1) / kpt;
  const int padded = ((nx + 31) / 32 * 32) * ((ny + 7) / 8 * 8), ncell = nx * ny;
  static const int packed_mode = getenv("FV3LM_PACKED") ? atoi(getenv("FV3LM_PACKED")) : -1;     // -1: automatic, 0 / 1: forced (A/B runs)
  const bool packed = ncell < 65536 && (packed_mode == 1 || (packed_mode == -1 && ncell * 100 < padded * 92));
  dim3 b(32, 8, 1), gr((nx + 31) / 32, (ny + 7) / 8, ntile * nkc);
  if (packed) gr = dim3((ncell + 255) / 256, 1, ntile * nkc);
  kern_stage<F><<<gr, b, 0, dev::stream()>>>(f, nx, ny, nk, nkc, 1.0f / (float)nkc, kpt, packed ? ncell : 0, 1.0f / (float)nx);
  dev::launches++;
}
#else
inline int stage_kpt() { static const int k = getenv("FV3LM_KPT") ? std::max(1, atoi(getenv("FV3LM_KPT"))) : 4; return k; }
template <class F> void launch_stage(const F& f, int nx, int ny, int ntile, int nk) {
  const int kpt = stage_kpt();
  for (int t = 0; t < ntile; t++)
    for (int k0 = 0; k0 < nk; k0 += kpt)
      for (int jj = 0; jj < ny; jj++)
        for (int ii = 0; ii < nx; ii++) {
          if (kpt == 1) f(ii, jj, k0, t);
          else f.run(ii, jj, k0, std::min(k0 + kpt, nk), t);
        }
  dev::launches++;
}
template <class F> void launch3d(const F& f, int nx, int ny, int nz) {
  for (int z = 0; z < nz; z++)
    for (int jj = 0; jj < ny; jj++)
      for (int ii = 0; ii < nx; ii++) f(ii, jj, z);
  dev::launches++;
}
#endif

// ---------------------------------------------------------------------------------
// programs
// ---------------------------------------------------------------------------------
enum Mode { MODE_NL = 0, MODE_TL = 1, MODE_AD = 2, MODE_ADFWD = 3 };
enum Variant { VAR_ALL = 0, VAR_FWD = 1, VAR_AD = 2 };

struct Pool {
  std::map<size_t, std::vector<void*>> free_;
  std::map<void*, size_t> live_;
  size_t bytes_total = 0, bytes_peak = 0, bytes_live = 0;
  double* get(size_t n_doubles, bool* fresh = nullptr);   // fresh != null: zero new memory, report whether it was new
  void put(double* p);
  void trim();
  ~Pool();
};

struct Value {
  std::string name;
  int nk = 1;
  bool active = false;     // carries a perturbation / adjoint
  bool external = false;   // owned by the caller of the program (never pooled)
  double* traj = nullptr;
  double* pert = nullptr;  // TL tangent, or AD adjoint
  int first_def = -1, last_use = -1;
  // >= 0: a DETACHED view of that value -- same trajectory storage, never active.  The trajectory-scheme chain of a split
  // transport (model_tlmadm/sw_core_tlm.F90:1664-1682: linear scheme for the perturbation, the nonlinear model's scheme
  // for the trajectory) reads its inputs through such views, so no perturbation / adjoint flows through it.
  int alias = -1;
};

struct Device;  // per-handle state (geometry, metrics, pool)

struct Op {
  std::string name;
  std::vector<int> in, out;
  bool inplace = false;    // patch op: out == in, mutates cells that were dead
  std::vector<int> hold;   // targets of the detached views among `in`: kept alive up to this op (filled by Program::analyse)
  bool tl_only = false;    // perturbation-scheme chain of a split transport: its trajectory values are only needed by TL / AD sweeps
  // alternative implementations of the same piece of a program: 0 = always part of it, VAR_FWD = only in NL / TL sweeps (fused
  // shared-memory-tile kernels that keep their intermediates on chip), VAR_AD = only in adjoint runs (the stage-by-stage chain
  // whose intermediates the reverse sweep needs in HBM).  Program::analyse / run ignore the ops of the other kind.
  int variant = 0;
  int nk_launch = 1;
  std::function<void(struct Program&, Op&, int /*Mode or 3 = AD reverse*/)> run;
};

struct Program {
  Device* dv = nullptr;
  std::vector<Value> vals;
  std::vector<Op> ops;
  std::string name;
  std::vector<int> seg_start;        // op indices where a recompute segment of the adjoint begins
  void mark_segment() { seg_start.push_back((int)ops.size()); }
  // device-side status flags raised by ops (value != 0 = error; checked by the step API after a sweep) and small device tables
  // that live as long as the program
  struct StatusFlag { const double* flag; std::string what; };
  std::vector<StatusFlag> status_flags;
  std::vector<std::shared_ptr<void>> keep_alive;
  void check_status_flags();
  bool tl_only = false;              // builders set this around the ops of a perturbation-scheme chain (copied into Op::tl_only)
  std::map<int, int> detached_of;
  int detached(int id) {             // detached view of a value (cached); negative ids (absent optional inputs) pass through
    if (id < 0) return id;
    if (vals[id].alias >= 0) return id;
    auto it = detached_of.find(id);
    if (it != detached_of.end()) return it->second;
    Value v; v.name = vals[id].name + "~"; v.nk = vals[id].nk; v.alias = id;
    vals.push_back(v);
    return detached_of[id] = (int)vals.size() - 1;
  }

  int val(const std::string& nm, int nk, bool external = false) {
    Value v; v.name = nm; v.nk = nk; v.external = external;
    vals.push_back(v);
    return (int)vals.size() - 1;
  }
  size_t val_doubles(int id) const;
  int sweep_kind = VAR_FWD;          // which alternative ops (Op::variant) the current run uses: set by run(), or by the caller before analyse()
  int variant = 0;                   // builders set this around alternative ops (copied into Op::variant)
  bool skipped(const Op& op) const { return op.variant != VAR_ALL && op.variant != sweep_kind; }
  void analyse();                    // activity + liveness
  bool ad_fits_store_all();
  int ad_store_all_cached = -1;
  int ad_keep_from = -1;             // first segment kept whole by the adjoint's forward pass (decided once)
  void run(Mode mode);               // NL, TL, or AD (forward store-all + reverse); releases every intermediate if a sweep throws
  void run_sweeps(Mode mode);
  void run_op(Op& op, int mode);     // one op, optionally profiled
  void ensure_traj(int id);
  void ensure_pert(int id, bool zero_it);
  void release(int id);
  template <class S>
  void add(const char* nm, const typename S::P& prm, std::vector<int> ins, std::vector<int> outs, int nk_launch);
};

// ---------------------------------------------------------------------------------
// column stages: one thread owns a whole (i, j) column and marches in k (vertical
// solvers, prefix sums).  NL/TL share one templated eval; the adjoint is hand-written
// (eval_ad) as a thread-local reverse sweep -- columns never interact, so no reduction.
//   struct S { NI, NO; struct P; template<class X> eval(X&, P&); template<class X> eval_ad(X&, P&); }
// ---------------------------------------------------------------------------------
template <class S> struct ColNL : CtxBase {
  using T = double;
  FArr<S::NI> in_; FArr<S::NO> out_;
  DEV int o2(int nkf, int k, int di, int dj) const { bounds(di, dj, "column field"); return (tile * nkf + (nkf == 1 ? 0 : k)) * g.slab + pos + dj * g.pitch + di; }
  // inputs are never outputs of the same op (SSA values): the read-only path lets the compiler start the loads of the next levels
  // before the stores of this one (a column sweep is otherwise one exposed DRAM latency per level)
  DEV T in(int f, int k, int di = 0, int dj = 0) const { return LDG(in_.p[f] + o2(in_.nk[f], k, di, dj)); }
  DEV void out(int o, int k, T v) const { out_.p[o][o2(out_.nk[o], k, 0, 0)] = v; }
  DEV T rd(int o, int k) const { return out_.p[o][o2(out_.nk[o], k, 0, 0)]; }   // read back own output
};
template <class S> struct ColTL : CtxBase {
  using T = Dual;
  FArr<S::NI> in_, ind_; FArr<S::NO> out_, outd_;
  DEV int o2(int nkf, int k, int di, int dj) const { bounds(di, dj, "column field"); return (tile * nkf + (nkf == 1 ? 0 : k)) * g.slab + pos + dj * g.pitch + di; }
  DEV T in(int f, int k, int di = 0, int dj = 0) const {
    const int o = o2(in_.nk[f], k, di, dj);
    return Dual(LDG(in_.p[f] + o), ind_.p[f] ? LDG(ind_.p[f] + o) : 0.0);
  }
  DEV void out(int o, int k, T v) const {
    const int q = o2(out_.nk[o], k, 0, 0);
    out_.p[o][q] = v.v;
    if (outd_.p[o]) outd_.p[o][q] = v.d;
  }
  DEV T rd(int o, int k) const { const int q = o2(out_.nk[o], k, 0, 0); return Dual(out_.p[o][q], outd_.p[o] ? outd_.p[o][q] : 0.0); }
};
template <class S> struct ColAD : CtxBase {
  FArr<S::NI> in_, inad_; FArr<S::NO> out_, outad_;
  DEV int o2(int nkf, int k) const { return (tile * nkf + (nkf == 1 ? 0 : k)) * g.slab + pos; }
  DEV double in(int f, int k) const { return LDG(in_.p[f] + o2(in_.nk[f], k)); }
  DEV double outv(int o, int k) const { return LDG(out_.p[o] + o2(out_.nk[o], k)); }          // stored forward value
  DEV double oad(int o, int k) const { return outad_.p[o] ? outad_.p[o][o2(outad_.nk[o], k)] : 0.0; }
  DEV void oad_add(int o, int k, double v) const { if (outad_.p[o]) outad_.p[o][o2(outad_.nk[o], k)] += v; }  // workspace use
  DEV void add(int f, int k, double v) const { if (inad_.p[f]) inad_.p[f][o2(inad_.nk[f], k)] += v; }
  // split form of add(): request the accumulators of a level (or of a few levels) first, store afterwards -- `add, add, add` makes every
  // load wait behind the previous store (the compiler cannot prove the adjoint arrays distinct): one exposed DRAM latency per call
  DEV double iad(int f, int k) const { return inad_.p[f] ? inad_.p[f][o2(inad_.nk[f], k)] : 0.0; }
  DEV void iad_set(int f, int k, double v) const { if (inad_.p[f]) inad_.p[f][o2(inad_.nk[f], k)] = v; }
  DEV bool same_ad(int f, int h) const { return inad_.p[f] == inad_.p[h]; }
  DEV bool active(int f) const { return inad_.p[f] != nullptr; }
};
template <class S> struct KernColNL {
  typename S::P p; Geom g; Metrics m; FArr<S::NI> in; FArr<S::NO> out;
  DEV void operator()(int ii, int jj, int z) const {
    ColNL<S> x; x.g = g; x.m = m; x.in_ = in; x.out_ = out; x.setpos(ii, jj, 0, z, g.i0[z], g.j0[z]);
    S::eval(x, p);
  }
};
template <class S> struct KernColTL {
  typename S::P p; Geom g; Metrics m; FArr<S::NI> in, ind; FArr<S::NO> out, outd;
  DEV void operator()(int ii, int jj, int z) const {
    ColTL<S> x; x.g = g; x.m = m; x.in_ = in; x.ind_ = ind; x.out_ = out; x.outd_ = outd; x.setpos(ii, jj, 0, z, g.i0[z], g.j0[z]);
    S::eval(x, p);
  }
};
template <class S> struct KernColAD {
  typename S::P p; Geom g; Metrics m; FArr<S::NI> in, inad; FArr<S::NO> out, outad;
  DEV void operator()(int ii, int jj, int z) const {
    ColAD<S> x; x.g = g; x.m = m; x.in_ = in; x.inad_ = inad; x.out_ = out; x.outad_ = outad; x.setpos(ii, jj, 0, z, g.i0[z], g.j0[z]);
    // a column whose outputs nobody consumed (all adjoints zero) contributes nothing; skipping it also keeps
    // 0 * (values derived from never-written cells) out of the sums
    bool any = false;
#pragma unroll
    for (int o = 0; o < S::NO; o++)
      if (outad.p[o]) for (int k = 0; k < outad.nk[o] && !any; k++) any = x.oad(o, k) != 0.0;
    if (!any) return;
    S::eval_ad(x, p);
  }
};

// splice of a split transport: trajectory from the trajectory-scheme chain (input 1, detached), perturbation / adjoint from the
// perturbation-scheme chain (input 0).  Whole array.   in: a b ; out: (traj b, pert a)
HD double with_val(double, double v) { return v; }
HD Dual with_val(Dual a, double v) { a.v = v; return a; }
template <int M> HD DualN<M> with_val(DualN<M> a, double v) { a.v = v; return a; }
struct S_splice {
  static constexpr int NI = 2, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    if constexpr (X::mode == 0) x.out(0, x.in(1));      // the perturbation chain is skipped in pure trajectory sweeps: input 0 has no storage
    else x.out(0, with_val(x.in(0), val(x.in(1))));
  }
};
// the same with per-level switches: a side that is switched off on a level contributes an exact zero there (its chain wrote
// nothing on that level).  Vorticity damping fluxes, model_tlmadm/sw_core_tlm.F90:2436-2451, 2506-2530.
struct LevMask { unsigned char v[128]; };
struct S_splice_lev {
  static constexpr int NI = 2, NO = 1;
  struct P { LevMask on_a, on_b; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const bool a = p.on_a.v[x.kk] != 0, b = p.on_b.v[x.kk] != 0;
    if constexpr (X::mode == 0) x.out(0, b ? x.in(1) : T(0.0));
    else x.out(0, with_val(a ? x.in(0) : T(0.0), b ? val(x.in(1)) : 0.0));
  }
};

struct Device {
  Geom g;
  Metrics m;
  double ad_store_budget = -1.0;   // bytes the adjoint may use to keep the whole forward sweep (< 0: ask the device)
  struct Comm* comm = nullptr;     // for the collective (min over ranks) version of that budget
  std::vector<double*> metric_bufs;
  Pool pool;
  std::shared_ptr<void> a2b_tables;   // a2b_ord4 compiled into a constant stencil + sparse edge rows (a2b.cu), built on first use
  std::string err;
};

template <class S>
void Program::add(const char* nm, const typename S::P& prm, std::vector<int> ins, std::vector<int> outs, int nk_launch) {
  if ((int)ins.size() != S::NI || (int)outs.size() != S::NO) throw std::runtime_error(std::string("arity mismatch in ") + nm);
  Op op; op.name = nm; op.in = ins; op.out = outs; op.nk_launch = nk_launch; op.tl_only = tl_only; op.variant = variant;
  typename S::P p = prm;
  op.run = [p](Program& P, Op& o, int mode) {
    const Geom& g = P.dv->g;
    if (mode == MODE_NL || mode == MODE_TL || mode == MODE_ADFWD) {
      FArr<S::NI> in, ind; FArr<S::NO> out, outd;
      for (int f = 0; f < S::NI; f++) { Value& v = P.vals[o.in[f]]; in.p[f] = v.traj; in.nk[f] = v.nk; ind.p[f] = v.active ? v.pert : nullptr; ind.nk[f] = v.nk; }
      for (int f = 0; f < S::NO; f++) { Value& v = P.vals[o.out[f]]; out.p[f] = v.traj; out.nk[f] = v.nk; outd.p[f] = v.active ? v.pert : nullptr; outd.nk[f] = v.nk; }
      bool any_active = false;
      for (int f = 0; f < S::NI; f++) any_active = any_active || ind.p[f];
      if (mode != MODE_TL || !any_active) {      // nothing to propagate (e.g. the trajectory-scheme chain of a split transport): values only
        KernNL<S> k{p, g, P.dv->m, in, out, o.nk_launch};
        launch_stage(k, g.NX, g.NY, g.ntile, o.nk_launch);
      } else {
        KernTL<S> k{p, g, P.dv->m, in, ind, out, outd, o.nk_launch};
        launch_stage(k, g.NX, g.NY, g.ntile, o.nk_launch);
      }
    } else {  // AD reverse
      FArr<S::NI> in, inad; FArr<S::NO> outad;
      int nk = o.nk_launch; bool any = false;
      for (int f = 0; f < S::NI; f++) { Value& v = P.vals[o.in[f]]; in.p[f] = v.traj; in.nk[f] = v.nk; inad.p[f] = v.active ? v.pert : nullptr; inad.nk[f] = v.nk; if (v.active) { any = true; if (v.nk > nk) nk = v.nk; } }
      for (int f = 0; f < S::NO; f++) { Value& v = P.vals[o.out[f]]; outad.p[f] = v.active ? v.pert : nullptr; outad.nk[f] = v.nk; }
      if (!any) return;
      KernAD<S> k{p, g, P.dv->m, in, inad, outad, nk, o.nk_launch};
      launch_stage(k, g.NX, g.NY, g.ntile, nk);
    }
  };
  ops.push_back(op);
}

template <class S>
void add_col(Program& P, const char* nm, const typename S::P& prm, std::vector<int> ins, std::vector<int> outs) {
  if ((int)ins.size() != S::NI || (int)outs.size() != S::NO) throw std::runtime_error(std::string("arity mismatch in ") + nm);
  Op op; op.name = nm; op.in = ins; op.out = outs; op.nk_launch = 1; op.tl_only = P.tl_only; op.variant = P.variant;
  typename S::P p = prm;
  op.run = [p](Program& P, Op& o, int mode) {
    const Geom& g = P.dv->g;
    FArr<S::NI> in, ind; FArr<S::NO> out, outd;
    for (int f = 0; f < S::NI; f++) { Value& v = P.vals[o.in[f]]; in.p[f] = v.traj; in.nk[f] = v.nk; ind.p[f] = v.active ? v.pert : nullptr; ind.nk[f] = v.nk; }
    for (int f = 0; f < S::NO; f++) { Value& v = P.vals[o.out[f]]; out.p[f] = v.traj; out.nk[f] = v.nk; outd.p[f] = v.active ? v.pert : nullptr; outd.nk[f] = v.nk; }
    if (mode == MODE_TL) launch3d(KernColTL<S>{p, g, P.dv->m, in, ind, out, outd}, g.NX, g.NY, g.ntile);
    else if (mode == MODE_AD) launch3d(KernColAD<S>{p, g, P.dv->m, in, ind, out, outd}, g.NX, g.NY, g.ntile);
    else launch3d(KernColNL<S>{p, g, P.dv->m, in, out}, g.NX, g.NY, g.ntile);
  };
  P.ops.push_back(op);
}


}  // namespace fv3lm
