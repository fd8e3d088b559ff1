// fv_tp_2d as two shared-memory-tile kernels (forward sweeps only: NL and TL).
//
// The stage-by-stage chain of build_fv_tp_2d (modules.cu) sends six intermediates through HBM per transport
// (fy2, q_i, fx_ou, fx2, q_j, fy_ou: 30 array passes per call, model_tlmadm/tp_core_tlm.F90:2123-2324).  Here a block owns a
// TX x TY tile of one level of one sub-domain and walks the same chain in shared memory:
//   kernel A (after copy_corners_y):  q tile -> inner y-flux fy2 -> q_i -> outer x-flux fx_ou     writes fy2, fx_ou   (7 passes)
//   kernel B (after copy_corners_x):  q tile -> inner x-flux fx2 -> q_j -> outer y-flux, then
//                                     fx = 0.5 (fx_ou + fx2) mx,  fy = 0.5 (fy_ou + fy2) my        writes fx, fy       (11 passes)
// The arithmetic is the stage functors' own (tp::ppm_flux on a context whose input 0 is the tile, the S_inner / S_favg
// expressions verbatim), so the results are those of the unfused chain.  A kernel is a sequence of PHASES with a block
// barrier between them: `phase(ph, tid, ...)` is called by every thread of the block on the device (__syncthreads() between
// phases) and in a loop over the block's threads per phase in the host-emulation build, which stays the parity gate.
//
// The adjoint keeps the stage-by-stage chain (its reverse sweep needs the intermediates in HBM): the two kernels are
// VAR_FWD ops, the chain ops VAR_AD (engine.h, Op::variant).
#pragma once
#include "stages_tp.h"

namespace fv3lm {
namespace ftp {

constexpr int TX = 32, TY = 8, NTHR = TX * TY;
constexpr int QW = TX + 6, QH = TY + 6;      // tile of the transported field with the PPM footprint (-3 .. +2) on both axes

struct Fld { const double* v; const double* d; int nk; };
struct OFld { double* v; double* d; int nk; };

template <class TT> struct Num;
template <> struct Num<double> {
  DEV static double ld(const Fld& f, int o) { return LDG(f.v + o); }
  DEV static double lds(const double* sv, const double*, int o) { return sv[o]; }
  DEV static void sts(double* sv, double*, int o, double a) { sv[o] = a; }
  DEV static void st(const OFld& f, int o, double a) { f.v[o] = a; }
};
template <> struct Num<Dual> {
  DEV static Dual ld(const Fld& f, int o) { return Dual(LDG(f.v + o), f.d ? LDG(f.d + o) : 0.0); }
  DEV static Dual lds(const double* sv, const double* sd, int o) { return Dual(sv[o], sd[o]); }
  DEV static void sts(double* sv, double* sd, int o, Dual a) { sv[o] = a.v; sd[o] = a.d; }
  DEV static void st(const OFld& f, int o, Dual a) { f.v[o] = a.v; if (f.d) f.d[o] = a.d; }
};
template <bool TLM, int N> struct SBuf { double v[N]; double d[TLM ? N : 1]; };

// the context tp::ppm_flux sees: the transported field (its only input) is a shared-memory tile, metrics stay global
template <class TT> struct TileCtx : CtxBase {
  using T = TT;
  static constexpr int mode = std::is_same<TT, double>::value ? 0 : 1;
  const double* sv; const double* sd; int sw, sp;     // tile (value, perturbation), its row width, position of the current cell
  DEV T in(int, int di = 0, int dj = 0, int = 0) const { return Num<TT>::lds(sv, sd, sp + dj * sw + di); }
};

// ---- kernel A: inner y sweep, outer x sweep ------------------------------------------------------------
template <class TT, bool FULL> struct KernTpA {
  static constexpr bool TLM = !std::is_same<TT, double>::value;
  static constexpr int NPH = 4;
  Geom g; Metrics m; LevOrd ord; int nk;
  Fld q, cry, yfx, ray, crx; OFld fy2, fxo;
  struct Smem { SBuf<TLM, QW * QH> q; SBuf<TLM, QW*(TY + 1)> fy; SBuf<TLM, QW * TY> qi; };
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& s) const {
    using N = Num<TT>;
    const int tile = z / nk, kk = z % nk, ii0 = bx * TX, jj0 = by * TY;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    const int is = g.is, ie = g.ie, js = g.js, je = g.je, isd = is - g.ng, ied = ie + g.ng;
    TileCtx<TT> x; x.g = g; x.m = m;
    if (ph == 0) {                 // q with its footprint
      for (int c = tid; c < QW * QH; c += NTHR) {
        const int ii = ii0 - 3 + c % QW, jj = jj0 - 3 + c / QW;
        TT a = TT(0.0);
        if (ii >= 0 && ii < g.NX && jj >= 0 && jj < g.NY) { x.setpos(ii, jj, kk, tile, i0, j0); a = N::ld(q, x.off(q.nk, 0, 0, 0)); }
        N::sts(s.q.v, s.q.d, c, a);
      }
    } else if (ph == 1) {          // fy2 = yppm(q, cry) on (isd:ied, js:je+1)
      for (int c = tid; c < QW * (TY + 1); c += NTHR) {
        const int ii = ii0 - 3 + c % QW, jj = jj0 + c / QW;
        if (ii < 0 || ii >= g.NX || jj >= g.NY) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!x.in_rect(isd, ied, js, je + 1)) continue;
        x.sv = s.q.v; x.sd = s.q.d; x.sw = QW; x.sp = (jj - jj0 + 3) * QW + (ii - ii0 + 3);
        const TT f = tp::ppm_flux<1, FULL>(x, 0, N::ld(cry, x.off(cry.nk, 0, 0, 0)), ord.v[kk]);
        N::sts(s.fy.v, s.fy.d, c, f);
        if (ii >= ii0 && ii < ii0 + TX && jj < jj0 + TY) N::st(fy2, x.off(fy2.nk, 0, 0, 0), f);
      }
    } else if (ph == 2) {          // q_i = (q area + yfx fy2 (j) - yfx fy2 (j+1)) / ra_y on (isd:ied, js:je)
      for (int c = tid; c < QW * TY; c += NTHR) {
        const int ii = ii0 - 3 + c % QW, jj = jj0 + c / QW;
        if (ii < 0 || ii >= g.NX || jj >= g.NY) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!x.in_rect(isd, ied, js, je)) continue;
        const TT f0 = N::ld(yfx, x.off(yfx.nk, 0, 0, 0)) * N::lds(s.fy.v, s.fy.d, c);
        const TT f1 = N::ld(yfx, x.off(yfx.nk, 0, 1, 0)) * N::lds(s.fy.v, s.fy.d, c + QW);
        const TT qq = N::lds(s.q.v, s.q.d, (jj - jj0 + 3) * QW + (ii - ii0 + 3));
        N::sts(s.qi.v, s.qi.d, c, (qq * x.M(x.m.area) + f0 - f1) / N::ld(ray, x.off(ray.nk, 0, 0, 0)));
      }
    } else {                       // fx_ou = xppm(q_i, crx) on (is:ie+1, js:je)
      const int ii = ii0 + tid % TX, jj = jj0 + tid / TX;
      if (ii >= g.NX || jj >= g.NY) return;
      x.setpos(ii, jj, kk, tile, i0, j0);
      if (!x.in_rect(is, ie + 1, js, je)) return;
      x.sv = s.qi.v; x.sd = s.qi.d; x.sw = QW; x.sp = (jj - jj0) * QW + (ii - ii0 + 3);
      N::st(fxo, x.off(fxo.nk, 0, 0, 0), tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, x.off(crx.nk, 0, 0, 0)), ord.v[kk]));
    }
  }
};

// ---- kernel B: inner x sweep, outer y sweep, flux averages -------------------------------------------------
template <class TT, bool FULL> struct KernTpB {
  static constexpr bool TLM = !std::is_same<TT, double>::value;
  static constexpr int NPH = 4;
  static constexpr int FW = TX + 1;
  Geom g; Metrics m; LevOrd ord; int nk;
  Fld q, crx, xfx, rax, cry, fy2, fxo, mx, my; OFld fx, fy;
  struct Smem { SBuf<TLM, QW * QH> q; SBuf<TLM, FW * QH> fx; SBuf<TLM, TX * QH> qj; };
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& s) const {
    using N = Num<TT>;
    const int tile = z / nk, kk = z % nk, ii0 = bx * TX, jj0 = by * TY;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    const int is = g.is, ie = g.ie, js = g.js, je = g.je, jsd = js - g.ng, jed = je + g.ng;
    TileCtx<TT> x; x.g = g; x.m = m;
    if (ph == 0) {
      for (int c = tid; c < QW * QH; c += NTHR) {
        const int ii = ii0 - 3 + c % QW, jj = jj0 - 3 + c / QW;
        TT a = TT(0.0);
        if (ii >= 0 && ii < g.NX && jj >= 0 && jj < g.NY) { x.setpos(ii, jj, kk, tile, i0, j0); a = N::ld(q, x.off(q.nk, 0, 0, 0)); }
        N::sts(s.q.v, s.q.d, c, a);
      }
    } else if (ph == 1) {          // fx2 = xppm(q, crx) on (is:ie+1, jsd:jed)
      for (int c = tid; c < FW * QH; c += NTHR) {
        const int ii = ii0 + c % FW, jj = jj0 - 3 + c / FW;
        if (ii >= g.NX || jj < 0 || jj >= g.NY) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!x.in_rect(is, ie + 1, jsd, jed)) continue;
        x.sv = s.q.v; x.sd = s.q.d; x.sw = QW; x.sp = (jj - jj0 + 3) * QW + (ii - ii0 + 3);
        N::sts(s.fx.v, s.fx.d, c, tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, x.off(crx.nk, 0, 0, 0)), ord.v[kk]));
      }
    } else if (ph == 2) {          // q_j = (q area + xfx fx2 (i) - xfx fx2 (i+1)) / ra_x on (is:ie, jsd:jed)
      for (int c = tid; c < TX * QH; c += NTHR) {
        const int ii = ii0 + c % TX, jj = jj0 - 3 + c / TX;
        if (ii >= g.NX || jj < 0 || jj >= g.NY) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!x.in_rect(is, ie, jsd, jed)) continue;
        const int cf = (jj - jj0 + 3) * FW + (ii - ii0);
        const TT f0 = N::ld(xfx, x.off(xfx.nk, 0, 0, 0)) * N::lds(s.fx.v, s.fx.d, cf);
        const TT f1 = N::ld(xfx, x.off(xfx.nk, 1, 0, 0)) * N::lds(s.fx.v, s.fx.d, cf + 1);
        const TT qq = N::lds(s.q.v, s.q.d, (jj - jj0 + 3) * QW + (ii - ii0 + 3));
        N::sts(s.qj.v, s.qj.d, c, (qq * x.M(x.m.area) + f0 - f1) / N::ld(rax, x.off(rax.nk, 0, 0, 0)));
      }
    } else {                       // fy_ou = yppm(q_j, cry) on (is:ie, js:je+1); the two averages (tp_core_tlm.F90:2268-2313)
      const int ii = ii0 + tid % TX, jj = jj0 + tid / TX;
      if (ii >= g.NX || jj >= g.NY) return;
      x.setpos(ii, jj, kk, tile, i0, j0);
      if (x.in_rect(is, ie, js, je + 1)) {
        x.sv = s.qj.v; x.sd = s.qj.d; x.sw = TX; x.sp = (jj - jj0 + 3) * TX + (ii - ii0);
        const TT fyo = tp::ppm_flux<1, FULL>(x, 0, N::ld(cry, x.off(cry.nk, 0, 0, 0)), ord.v[kk]);
        N::st(fy, x.off(fy.nk, 0, 0, 0), 0.5 * (fyo + N::ld(fy2, x.off(fy2.nk, 0, 0, 0))) * N::ld(my, x.off(my.nk, 0, 0, 0)));
      }
      if (x.in_rect(is, ie + 1, js, je)) {
        const TT fx2 = N::lds(s.fx.v, s.fx.d, (jj - jj0 + 3) * FW + (ii - ii0));
        N::st(fx, x.off(fx.nk, 0, 0, 0), 0.5 * (N::ld(fxo, x.off(fxo.nk, 0, 0, 0)) + fx2) * N::ld(mx, x.off(mx.nk, 0, 0, 0)));
      }
    }
  }
};

#ifndef FV3LM_HOST_EMU
template <class K> GLOBAL void __launch_bounds__(NTHR) kern_tile(const __grid_constant__ K k) {
  __shared__ typename K::Smem s;
  const int tid = threadIdx.y * TX + threadIdx.x;
#pragma unroll
  for (int ph = 0; ph < K::NPH; ph++) {
    k.phase(ph, tid, blockIdx.x, blockIdx.y, blockIdx.z, s);
    if (ph + 1 < K::NPH) __syncthreads();
  }
}
template <class K> void launch_tile(const K& k, int nx, int ny, int nz) {
  if (nz <= 0) return;
  dim3 b(TX, TY, 1), gr((nx + TX - 1) / TX, (ny + TY - 1) / TY, nz);
  kern_tile<K><<<gr, b, 0, dev::stream()>>>(k);
  dev::launches++;
}
#else
template <class K> void launch_tile(const K& k, int nx, int ny, int nz) {
  std::unique_ptr<typename K::Smem> s(new typename K::Smem);
  for (int z = 0; z < nz; z++)
    for (int by = 0; by < (ny + TY - 1) / TY; by++)
      for (int bx = 0; bx < (nx + TX - 1) / TX; bx++) {
        // a fresh block finds arbitrary shared memory: poison it so that a read of a cell no phase wrote shows up
        memset(s.get(), 0xff, sizeof(typename K::Smem));
        for (int ph = 0; ph < K::NPH; ph++)
          for (int tid = 0; tid < NTHR; tid++) k.phase(ph, tid, bx, by, z, *s);
      }
  dev::launches++;
}
#endif

inline Fld fld(const Value& v, bool tl) { return Fld{v.traj, (tl && v.active) ? v.pert : nullptr, v.nk}; }
inline OFld ofld(const Value& v, bool tl) { return OFld{v.traj, (tl && v.active) ? v.pert : nullptr, v.nk}; }

template <template <class, bool> class K, class Fill>
void run_fused(Program& P, Op& o, int mode, bool full, const Fill& fill) {
  const Geom& g = P.dv->g;
  bool tl = false;
  if (mode == MODE_TL) for (int i : o.in) tl = tl || (P.vals[i].active && P.vals[i].pert);
  auto go = [&](auto kern) {
    kern.g = g; kern.m = P.dv->m; kern.nk = o.nk_launch;
    fill(kern, tl);
    launch_tile(kern, g.NX, g.NY, g.ntile * o.nk_launch);
  };
  if (tl) { if (full) go(K<Dual, true>{}); else go(K<Dual, false>{}); }
  else { if (full) go(K<double, true>{}); else go(K<double, false>{}); }
}

// appends the two fused ops (VAR_FWD) to a program; the caller has already added the copy_corners patches around them
inline void add_fused_a(Program& P, const std::string& nm, int q, int cry, int yfx, int ra_y, int crx, int fy2, int fxo,
                        const LevOrd& hord, bool full, int nk) {
  Op op; op.name = nm; op.in = {q, cry, yfx, ra_y, crx}; op.out = {fy2, fxo}; op.nk_launch = nk; op.tl_only = P.tl_only; op.variant = VAR_FWD;
  op.run = [hord, full](Program& P, Op& o, int mode) {
    if (mode != MODE_NL && mode != MODE_TL) throw std::runtime_error("fused fv_tp_2d kernels run in forward sweeps only");
    run_fused<KernTpA>(P, o, mode, full, [&](auto& k, bool tl) {
      k.ord = hord;
      k.q = fld(P.vals[o.in[0]], tl); k.cry = fld(P.vals[o.in[1]], tl); k.yfx = fld(P.vals[o.in[2]], tl);
      k.ray = fld(P.vals[o.in[3]], tl); k.crx = fld(P.vals[o.in[4]], tl);
      k.fy2 = ofld(P.vals[o.out[0]], tl); k.fxo = ofld(P.vals[o.out[1]], tl);
    });
  };
  P.ops.push_back(op);
}
inline void add_fused_b(Program& P, const std::string& nm, int q, int crx, int xfx, int ra_x, int cry, int fy2, int fxo, int mx, int my,
                        int fx, int fy, const LevOrd& hord, bool full, int nk) {
  Op op; op.name = nm; op.in = {q, crx, xfx, ra_x, cry, fy2, fxo, mx, my}; op.out = {fx, fy}; op.nk_launch = nk; op.tl_only = P.tl_only; op.variant = VAR_FWD;
  op.run = [hord, full](Program& P, Op& o, int mode) {
    if (mode != MODE_NL && mode != MODE_TL) throw std::runtime_error("fused fv_tp_2d kernels run in forward sweeps only");
    run_fused<KernTpB>(P, o, mode, full, [&](auto& k, bool tl) {
      k.ord = hord;
      k.q = fld(P.vals[o.in[0]], tl); k.crx = fld(P.vals[o.in[1]], tl); k.xfx = fld(P.vals[o.in[2]], tl); k.rax = fld(P.vals[o.in[3]], tl);
      k.cry = fld(P.vals[o.in[4]], tl); k.fy2 = fld(P.vals[o.in[5]], tl); k.fxo = fld(P.vals[o.in[6]], tl);
      k.mx = fld(P.vals[o.in[7]], tl); k.my = fld(P.vals[o.in[8]], tl);
      k.fx = ofld(P.vals[o.out[0]], tl); k.fy = ofld(P.vals[o.out[1]], tl);
    });
  };
  P.ops.push_back(op);
}

}  // namespace ftp
}  // namespace fv3lm
