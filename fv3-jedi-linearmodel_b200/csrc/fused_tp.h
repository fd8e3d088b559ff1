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

// a block of NTHR threads owns TX x TY cells (every phase is a loop over its cell set, so the tile height can be tuned
// independently of the block size: a taller tile has less footprint overhead, a flatter one more blocks per SM)
#ifndef FV3LM_TILE_TY
#define FV3LM_TILE_TY 16     // build.sh passes -DFV3LM_TILE_TY=$FV3LM_TILE_TY for A/B runs (8: 2.1 field cells loaded and 1.8 inner fluxes per owned cell, 16: 1.6 and 1.4)
#endif
constexpr int TX = 32, TY = FV3LM_TILE_TY, NTHR = 256;
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
    int tile, kk; split_z(z, nk, tile, kk);
    const int ii0 = bx * TX, jj0 = by * TY;
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
      for (int c = tid; c < TX * TY; c += NTHR) {
        const int ii = ii0 + c % TX, jj = jj0 + c / TX;
        if (ii >= g.NX || jj >= g.NY) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!x.in_rect(is, ie + 1, js, je)) continue;
        x.sv = s.qi.v; x.sd = s.qi.d; x.sw = QW; x.sp = (jj - jj0) * QW + (ii - ii0 + 3);
        N::st(fxo, x.off(fxo.nk, 0, 0, 0), tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, x.off(crx.nk, 0, 0, 0)), ord.v[kk]));
      }
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
    int tile, kk; split_z(z, nk, tile, kk);
    const int ii0 = bx * TX, jj0 = by * TY;
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
      auto own = [&](int ii, int jj) {
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
      };
      for (int c = tid; c < TX * TY; c += NTHR) own(ii0 + c % TX, jj0 + c / TX);
    }
  }
};

#ifndef FV3LM_HOST_EMU
template <class K> GLOBAL void __launch_bounds__(NTHR) kern_tile(const __grid_constant__ K k) {
  __shared__ typename K::Smem s;
  const int tid = threadIdx.x;
#pragma unroll
  for (int ph = 0; ph < K::NPH; ph++) {
    k.phase(ph, tid, blockIdx.x, blockIdx.y, blockIdx.z, s);
    if (ph + 1 < K::NPH) __syncthreads();
  }
}
template <class K> void launch_tile(const K& k, int nx, int ny, int nz) {
  if (nz <= 0) return;
  dim3 b(NTHR, 1, 1), gr((nx + TX - 1) / TX, (ny + TY - 1) / TY, nz);
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


// =====================================================================================================================
// Reverse sweep of one half of fv_tp_2d (inner sweep along axis DIN, outer sweep along the other axis) as ONE tile kernel in
// gather form.  AVG = false is the reverse of kernel A (DIN = 1: y inside, x outside; incoming adjoints: fx_ou and fy2), AVG = true
// the reverse of kernel B (DIN = 0; incoming adjoints fx, fy; the two flux averages are part of it).  Only the op's INPUT values
// are read: the intermediates (inner flux Fi, updated field qm, outer flux Fo) are recomputed in the tile, so the forward sweep of
// the adjoint does not have to keep them in HBM.  With I the inner and O the outer axis, for the cells of the block's tile:
//   Fo_ad(c)  = adjoint of the outer flux at c                                        tile + (-3..+2 along I, -2..+2 along O)
//   qm_ad(c)  = sum_s  dFo(c + s eO) / dqm(c) * Fo_ad(c + s eO),  s = -2 .. 3           tile + (-3..+2 along I)
//   Fi_ad(c)  = Fi_ext_ad(c) + fi(c) * (qm_ad(c) / ra(c) - qm_ad(c - eI) / ra(c - eI))  tile + (-2..+2 along I)
//   q_ad(c)  += qm_ad(c) area / ra(c) + sum_s dFi(c + s eI) / dq(c) * Fi_ad(c + s eI)
// and the adjoints of the Courant numbers, flux areas and ra at the cell itself.  Every thread owns the cells it adds to: no
// atomics, fixed summation order.  The linear orders only (1, 2, 333): nothing is ever differentiated through the others.
// Hand-derived from tp_core_tlm.F90:2123-2324, 2328-2660 (the coefficient formulas are those of S_ppm::adjoint).
// =====================================================================================================================
struct Box {
  int x0, y0, w, h;
  DEV int idx(int ii, int jj) const { return (jj - y0) * w + (ii - x0); }
  DEV bool has(int ii, int jj) const { return ii >= x0 && ii < x0 + w && jj >= y0 && jj < y0 + h; }
  DEV int n() const { return w * h; }
};
// I-range [ia, ib], O-range [oa, ob] (array coordinates along the inner / outer axis) -> box in (ii, jj)
template <int DIN> DEV Box box_of(int ia, int ib, int oa, int ob) {
  return DIN == 0 ? Box{ia, oa, ib - ia + 1, ob - oa + 1} : Box{oa, ia, ob - oa + 1, ib - ia + 1};
}

// d flux(face) / d q(face + d) of the linear PPM fluxes; x is positioned at the face
template <int DIR, class X> DEV double ppm_coef(const X& x, double c, int ord, int d) {
  if (ord == 1) return (c > 0.0) ? (d == -1 ? 1.0 : 0.0) : (d == 0 ? 1.0 : 0.0);
  if (ord == ORD333) {
    const double c2 = c * c / 6.0;
    if (c > 0.0) return d == 0 ? 2.0 / 6.0 - 0.5 * c + c2 : d == -1 ? 5.0 / 6.0 + 0.5 * c - 2.0 * c2 : d == -2 ? -1.0 / 6.0 + c2 : 0.0;
    return d == -1 ? 2.0 / 6.0 + 0.5 * c + c2 : d == 0 ? 5.0 / 6.0 - 0.5 * c - 2.0 * c2 : d == 1 ? -1.0 / 6.0 + c2 : 0.0;
  }
  using S = S_ppm<DIR>;
  {
    // regular faces: the three edge values involved use the uniform weights (p2, p1, p1, p2) -- no metric loads, no edge tests
    const int ia = DIR == 0 ? x.i : x.j, np = DIR == 0 ? x.g.npx : x.g.npy;
    if (ia >= 4 && ia <= np - 3) {
      auto W = [](int n) { return (n == 0 || n == 3) ? tp::p2 : ((n == 1 || n == 2) ? tp::p1 : 0.0); };
      if (c > 0.0) return (d == -1 ? 1.0 + (1.0 - c) * (2.0 * c - 1.0) : 0.0) + (1.0 - c) * (1.0 - c) * W(d + 2) - (1.0 - c) * c * W(d + 3);
      return (d == 0 ? 1.0 - (1.0 + c) * (1.0 + 2.0 * c) : 0.0) + (1.0 + c) * (1.0 + c) * W(d + 2) + (1.0 + c) * c * W(d + 1);
    }
  }
  if (c > 0.0) {
    const double dqt = 1.0 + (1.0 - c) * (2.0 * c - 1.0), dal0 = (1.0 - c) * (1.0 - c), dalm = -(1.0 - c) * c;
    return (d == -1 ? dqt : 0.0) + dal0 * S::al_w(x, 0, d + 2) + dalm * S::al_w(x, -1, d + 3);
  }
  const double dqt = 1.0 - (1.0 + c) * (1.0 + 2.0 * c), dal0 = (1.0 + c) * (1.0 + c), dalp = (1.0 + c) * c;
  return (d == 0 ? dqt : 0.0) + dal0 * S::al_w(x, 0, d + 2) + dalp * S::al_w(x, 1, d + 1);
}
// d flux(face) / d c(face); x (positioned at the face) reads the transported field from its tile
template <int DIR, class X> DEV double ppm_dc(const X& x, double c, int ord) {
  if (ord == 1) return 0.0;
  if (ord == ORD333) {
    const double qm1 = tp::Q<DIR>(x, 0, -1), q0 = tp::Q<DIR>(x, 0, 0);
    const double curv = c > 0.0 ? q0 - 2.0 * qm1 + tp::Q<DIR>(x, 0, -2) : tp::Q<DIR>(x, 0, 1) - 2.0 * q0 + qm1;
    return -0.5 * (q0 - qm1) + c / 3.0 * curv;
  }
  const double al0 = tp::edge_al<DIR>(x, 0, 0);
  if (c > 0.0) {
    const double qt = tp::Q<DIR>(x, 0, -1), b = tp::edge_al<DIR>(x, 0, -1) + al0 - (qt + qt);
    return -(al0 - qt - c * b) - (1.0 - c) * b;
  }
  const double qt = tp::Q<DIR>(x, 0, 0), b = al0 + tp::edge_al<DIR>(x, 0, 1) - (qt + qt);
  return (al0 - qt + c * b) + (1.0 + c) * b;
}

template <int DIN, bool AVG> struct KernTpRev {
  static constexpr int DOUT = 1 - DIN;
  static constexpr int NPH = 5;
  static constexpr int TA = DIN == 0 ? TX : TY, TB = DIN == 0 ? TY : TX;
  Geom g; Metrics m; LevOrd ord; int nk;
  // values: transported field, inner Courant number / flux area / ra, outer Courant number; AVG: the other two fluxes and the multipliers
  Fld q, ci, fi, ra, co, fin2, fout2, mI, mO;
  // incoming adjoints (.v = adjoint array): !AVG: aI = fy2_ad (inner flux), aO = fx_ou_ad (outer flux); AVG: aI = fx_ad, aO = fy_ad
  Fld aI, aO;
  // accumulated adjoints (.v = adjoint array, null = inactive)
  OFld q_ad, ci_ad, fi_ad, ra_ad, co_ad, fin2_ad, fout2_ad, mI_ad, mO_ad;
  struct Smem {
    double q[QW * QH], Fi[(TA + 1) * (TB + 6)], qm[TA * (TB + 6)];
    double Fo_ad[(TA + 6) * (TB + 5)], qm_ad[(TA + 6) * TB], Fi_ad[(TA + 5) * TB];
  };
  DEV static int eIx() { return DIN == 0; }
  DEV static int eIy() { return DIN == 1; }
  DEV static int eOx() { return DIN == 1; }
  DEV static int eOy() { return DIN == 0; }
  DEV bool rect_i(const CtxBase& x) const {   // inner flux: faces along I, the full halo along O
    return DIN == 0 ? x.in_rect(g.is, g.ie + 1, g.js - g.ng, g.je + g.ng) : x.in_rect(g.is - g.ng, g.ie + g.ng, g.js, g.je + 1);
  }
  DEV bool rect_m(const CtxBase& x) const {
    return DIN == 0 ? x.in_rect(g.is, g.ie, g.js - g.ng, g.je + g.ng) : x.in_rect(g.is - g.ng, g.ie + g.ng, g.js, g.je);
  }
  DEV bool rect_o(const CtxBase& x) const {   // outer flux: faces along O, compute domain along I
    return DIN == 0 ? x.in_rect(g.is, g.ie, g.js, g.je + 1) : x.in_rect(g.is, g.ie + 1, g.js, g.je);
  }
  DEV bool rect_fi(const CtxBase& x) const {  // AVG: where the averaged flux along I (fx) lives
    return x.in_rect(g.is, g.ie + 1, g.js, g.je);
  }
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& s) const {
    int tile, kk; split_z(z, nk, tile, kk);
    const int ii0 = bx * TX, jj0 = by * TY;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    const int a0 = DIN == 0 ? ii0 : jj0, b0 = DIN == 0 ? jj0 : ii0;
    const int od = ord.v[kk];
    const Box bq = box_of<DIN>(a0 - 3, a0 + TA + 2, b0 - 3, b0 + TB + 2);
    const Box bFi = box_of<DIN>(a0, a0 + TA, b0 - 3, b0 + TB + 2);
    const Box bqm = box_of<DIN>(a0, a0 + TA - 1, b0 - 3, b0 + TB + 2);
    const Box bFoad = box_of<DIN>(a0 - 3, a0 + TA + 2, b0 - 2, b0 + TB + 2);
    const Box bqmad = box_of<DIN>(a0 - 3, a0 + TA + 2, b0, b0 + TB - 1);
    const Box bFiad = box_of<DIN>(a0 - 2, a0 + TA + 2, b0, b0 + TB - 1);
    TileCtx<double> x; x.g = g; x.m = m; x.sd = nullptr;
    auto inside = [&](int ii, int jj) { return ii >= 0 && ii < g.NX && jj >= 0 && jj < g.NY; };
    auto gl = [&](const Fld& f, int di = 0, int dj = 0) { return LDG(f.v + x.off(f.nk, di, dj, 0)); };
    if (ph == 0) {
      // the transported field with its footprint; the adjoint of the outer flux where it exists, zero elsewhere
      for (int c = tid; c < bq.n(); c += NTHR) {
        const int ii = bq.x0 + c % bq.w, jj = bq.y0 + c / bq.w;
        double a = 0.0;
        if (inside(ii, jj)) { x.setpos(ii, jj, kk, tile, i0, j0); a = gl(q); }
        s.q[c] = a;
      }
      for (int c = tid; c < bFoad.n(); c += NTHR) {
        const int ii = bFoad.x0 + c % bFoad.w, jj = bFoad.y0 + c / bFoad.w;
        double a = 0.0;
        if (inside(ii, jj)) {
          x.setpos(ii, jj, kk, tile, i0, j0);
          if (rect_o(x)) { a = gl(aO); if (AVG && a != 0.0) a *= 0.5 * gl(mO); }
        }
        s.Fo_ad[c] = a;
      }
    } else if (ph == 1) {          // inner flux values
      for (int c = tid; c < bFi.n(); c += NTHR) {
        const int ii = bFi.x0 + c % bFi.w, jj = bFi.y0 + c / bFi.w;
        if (!inside(ii, jj)) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!rect_i(x)) continue;
        x.sv = s.q; x.sw = bq.w; x.sp = bq.idx(ii, jj);
        s.Fi[c] = tp::ppm_flux<DIN, false>(x, 0, gl(ci), od);
      }
    } else if (ph == 2) {          // qm values (S_inner); qm_ad by gathering the outer faces that read the cell
      for (int c = tid; c < bqm.n(); c += NTHR) {
        const int ii = bqm.x0 + c % bqm.w, jj = bqm.y0 + c / bqm.w;
        if (!inside(ii, jj)) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (!rect_m(x)) continue;
        const double f0 = gl(fi) * s.Fi[bFi.idx(ii, jj)], f1 = gl(fi, eIx(), eIy()) * s.Fi[bFi.idx(ii + eIx(), jj + eIy())];
        s.qm[c] = (s.q[bq.idx(ii, jj)] * x.M(x.m.area) + f0 - f1) / gl(ra);
      }
      for (int c = tid; c < bqmad.n(); c += NTHR) {
        const int ii = bqmad.x0 + c % bqmad.w, jj = bqmad.y0 + c / bqmad.w;
        double sum = 0.0;
        bool in_m = false;
        if (inside(ii, jj)) { x.setpos(ii, jj, kk, tile, i0, j0); in_m = rect_m(x); }
        if (in_m) {
          for (int sft = -2; sft <= 3; sft++) {       // face f = cell + sft along O reads the cell at offset d = -sft
            const int fi_ = ii + sft * eOx(), fj_ = jj + sft * eOy();
            if (!inside(fi_, fj_) || !bFoad.has(fi_, fj_)) continue;
            const double a = s.Fo_ad[bFoad.idx(fi_, fj_)];
            if (a == 0.0) continue;
            x.setpos(fi_, fj_, kk, tile, i0, j0);
            sum += ppm_coef<DOUT>(x, gl(co), od, -sft) * a;
          }
        }
        s.qm_ad[c] = sum;
      }
    } else if (ph == 3) {          // adjoint of the inner flux
      for (int c = tid; c < bFiad.n(); c += NTHR) {
        const int ii = bFiad.x0 + c % bFiad.w, jj = bFiad.y0 + c / bFiad.w;
        double a = 0.0;
        if (inside(ii, jj)) {
          x.setpos(ii, jj, kk, tile, i0, j0);
          if (rect_i(x)) {
            if (AVG) { if (rect_fi(x)) { a = gl(aI); if (a != 0.0) a *= 0.5 * gl(mI); } }
            else a = gl(aI);
            double t = 0.0;
            const double qa0 = rect_m(x) ? s.qm_ad[bqmad.idx(ii, jj)] : 0.0;
            if (qa0 != 0.0) t += qa0 / gl(ra);
            const int pi = ii - eIx(), pj = jj - eIy();
            if (inside(pi, pj) && bqmad.has(pi, pj)) {
              const double qa1 = s.qm_ad[bqmad.idx(pi, pj)];      // (zero outside rect_m)
              if (qa1 != 0.0) t -= qa1 / gl(ra, -eIx(), -eIy());
            }
            if (t != 0.0) a += gl(fi) * t;
          }
        }
        s.Fi_ad[c] = a;
      }
    } else if (ph == 4) {          // everything the block's own cells accumulate
      auto own = [&](int ii, int jj) {
      if (!inside(ii, jj)) return;
      x.setpos(ii, jj, kk, tile, i0, j0);
      const bool in_i = rect_i(x), in_m = rect_m(x), in_o = rect_o(x);
      double aq = 0.0;
      if (in_m) {
        const double qa = s.qm_ad[bqmad.idx(ii, jj)];
        if (qa != 0.0) {
          const double r = gl(ra);
          aq += qa * x.M(x.m.area) / r;
          if (ra_ad.v) ra_ad.v[x.off(ra_ad.nk, 0, 0, 0)] += -s.qm[bqm.idx(ii, jj)] / r * qa;
        }
      }
      if (in_i) {
        const double fa = s.Fi_ad[bFiad.idx(ii, jj)];
        if (fi_ad.v) {          // flux area: Fi(c) * (qm_ad(c) / ra(c) - qm_ad(c - eI) / ra(c - eI))
          double t = 0.0;
          const double qa0 = in_m ? s.qm_ad[bqmad.idx(ii, jj)] : 0.0;
          if (qa0 != 0.0) t += qa0 / gl(ra);
          const int pi = ii - eIx(), pj = jj - eIy();
          if (inside(pi, pj)) { const double qa1 = s.qm_ad[bqmad.idx(pi, pj)]; if (qa1 != 0.0) t -= qa1 / gl(ra, -eIx(), -eIy()); }
          if (t != 0.0) fi_ad.v[x.off(fi_ad.nk, 0, 0, 0)] += s.Fi[bFi.idx(ii, jj)] * t;
        }
        if (ci_ad.v && fa != 0.0) {
          x.sv = s.q; x.sw = bq.w; x.sp = bq.idx(ii, jj);
          ci_ad.v[x.off(ci_ad.nk, 0, 0, 0)] += ppm_dc<DIN>(x, gl(ci), od) * fa;
        }
      }
      if (in_o) {
        const double oa = s.Fo_ad[bFoad.idx(ii, jj)];
        if (co_ad.v && oa != 0.0) {
          x.sv = s.qm; x.sw = bqm.w; x.sp = bqm.idx(ii, jj);
          co_ad.v[x.off(co_ad.nk, 0, 0, 0)] += ppm_dc<DOUT>(x, gl(co), od) * oa;
        }
      }
      if (AVG) {
        if (in_o) {             // fy = 0.5 (Fo + fy2) mO
          const double a = gl(aO);
          if (a != 0.0) {
            if (fin2_ad.v) fin2_ad.v[x.off(fin2_ad.nk, 0, 0, 0)] += 0.5 * gl(mO) * a;
            if (mO_ad.v) {
              x.sv = s.qm; x.sw = bqm.w; x.sp = bqm.idx(ii, jj);
              const double Fo = tp::ppm_flux<DOUT, false>(x, 0, gl(co), od);
              mO_ad.v[x.off(mO_ad.nk, 0, 0, 0)] += 0.5 * (Fo + gl(fin2)) * a;
            }
          }
        }
        if (rect_fi(x)) {       // fx = 0.5 (fx_ou + Fi) mI
          const double a = gl(aI);
          if (a != 0.0) {
            if (fout2_ad.v) fout2_ad.v[x.off(fout2_ad.nk, 0, 0, 0)] += 0.5 * gl(mI) * a;
            if (mI_ad.v) mI_ad.v[x.off(mI_ad.nk, 0, 0, 0)] += 0.5 * (gl(fout2) + s.Fi[bFi.idx(ii, jj)]) * a;
          }
        }
      }
      // the inner faces that read this cell
      if (q_ad.v) {
        for (int sft = -2; sft <= 3; sft++) {
          const int fi_ = ii + sft * eIx(), fj_ = jj + sft * eIy();
          if (!inside(fi_, fj_)) continue;
          const double a = s.Fi_ad[bFiad.idx(fi_, fj_)];      // (zero outside rect_i)
          if (a == 0.0) continue;
          x.setpos(fi_, fj_, kk, tile, i0, j0);
          aq += ppm_coef<DIN>(x, gl(ci), od, -sft) * a;
        }
        x.setpos(ii, jj, kk, tile, i0, j0);
        if (aq != 0.0) q_ad.v[x.off(q_ad.nk, 0, 0, 0)] += aq;
      }
      };
      for (int c = tid; c < TX * TY; c += NTHR) own(ii0 + c % TX, jj0 + c / TX);
    }
  }
};


// =====================================================================================================================
// Generic pieces for fusing a chain of stencil STAGES (engine.h) in a tile: a context whose inputs / outputs are shared-memory
// tiles or global fields, so that a stage's own eval() runs unchanged, and the gather-form reverse of a LINEAR stage on tiles.
// =====================================================================================================================
struct TRef { const double* v; const double* d; Box b; };    // input tile (v == nullptr: the input is the global field gi[f])
struct TOut { double* v; double* d; Box b; };                // output tile (v == nullptr: not staged)
template <class TT, int NI, int NO> struct TCtx : CtxBase {
  using T = TT;
  static constexpr int mode = std::is_same<TT, double>::value ? 0 : 1;
  TRef ti[NI]; Fld gi[NI]; TOut to[NO]; OFld go[NO];
  DEV T in(int f, int di = 0, int dj = 0, int dk = 0) const {
    if (ti[f].v) {
#ifdef FV3LM_HOST_EMU
      if (!ti[f].b.has(ii + di, jj + dj)) throw std::runtime_error("fv3lm emu: tile read outside its box");
#endif
      return Num<TT>::lds(ti[f].v, ti[f].d, ti[f].b.idx(ii + di, jj + dj));
    }
    return Num<TT>::ld(gi[f], off(gi[f].nk, di, dj, dk));
  }
  DEV void out(int o, T v) const {
    if (to[o].v) Num<TT>::sts(to[o].v, to[o].d, to[o].b.idx(ii, jj), v);
    if (go[o].v) Num<TT>::st(go[o], off(go[o].nk, 0, 0, 0), v);
  }
};

// one seeded evaluation of a LINEAR stage S at an output cell: the inputs are zeros, the tap (F, di, dj) carries the seed, the
// output adjoints come from tiles.  (The coefficients of a linear stage do not depend on its inputs.)
template <class S> struct TCtxLinAD : CtxBase {
  using T = DualN<1>;
  static constexpr int mode = 2;
  const double* oad[S::NO]; Box ob[S::NO];
  int sf, sdi, sdj;
  double acc;
  DEV T in(int f, int di = 0, int dj = 0, int = 0) const {
    T r(0.0);
    if (f == sf && di == sdi && dj == sdj) r.d[0] = 1.0;
    return r;
  }
  DEV void out(int o, const T& v) {
    if (oad[o] && ob[o].has(ii, jj)) {
      const double a = oad[o][ob[o].idx(ii, jj)];
      if (a != 0.0) acc += v.d[0] * a;
    }
  }
};
// adjoint of input F at the cell (ii, jj): sum over the taps of F of d out(cell - tap) / d in * out_ad(cell - tap)
template <class S, int F, int n = 0> struct TileLinGather {
  template <class C> DEV static void run(C& x, const typename S::P& p, int ii, int jj, int kk, int tile, int i0, int j0) {
    if constexpr (n < S::NT) {
      constexpr Tap t = S::taps[n];
      if constexpr (t.f == F) {
        const int oi = ii - t.di, oj = jj - t.dj;
        if (oi >= 0 && oi < x.g.NX && oj >= 0 && oj < x.g.NY) {
          x.sf = F; x.sdi = t.di; x.sdj = t.dj;
          x.setpos(oi, oj, kk, tile, i0, j0);
          S::eval(x, p);
        }
      }
      TileLinGather<S, F, n + 1>::run(x, p, ii, jj, kk, tile, i0, j0);
    }
  }
};

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

inline Fld adj_in(const Value& v) { return Fld{v.active ? v.pert : nullptr, nullptr, v.nk}; }
inline OFld adj_out(const Value& v) { return OFld{v.active ? v.pert : nullptr, nullptr, v.nk}; }
inline Fld val_in(const Value& v) { return Fld{v.traj, nullptr, v.nk}; }

// appends the two fused ops to a program; the caller has already added the copy_corners patches around them.
// with_ad = false: forward sweeps only (VAR_FWD, the stage chain stays for adjoint runs); true: the ops also carry the reverse
// tile kernel (KernTpRev), the stage chain is not built at all and its six intermediates never reach HBM in any sweep.
inline void add_fused_a(Program& P, const std::string& nm, int q, int cry, int yfx, int ra_y, int crx, int fy2, int fxo,
                        const LevOrd& hord, bool full, int nk, bool with_ad) {
  Op op; op.name = nm; op.in = {q, cry, yfx, ra_y, crx}; op.out = {fy2, fxo}; op.nk_launch = nk; op.tl_only = P.tl_only;
  op.variant = with_ad ? P.variant : (int)VAR_FWD;
  for (int id : op.in) if (P.vals[id].nk != nk) throw std::runtime_error("fused fv_tp_2d: every field must have the launch's number of levels");
  op.run = [hord, full, with_ad](Program& P, Op& o, int mode) {
    if (mode == MODE_AD) {
      if (!with_ad || full) throw std::runtime_error("fused fv_tp_2d: no reverse kernel for this op");
      const Geom& g = P.dv->g;
      KernTpRev<1, false> k{};
      k.g = g; k.m = P.dv->m; k.ord = hord; k.nk = o.nk_launch;
      const Value &vq = P.vals[o.in[0]], &vci = P.vals[o.in[1]], &vfi = P.vals[o.in[2]], &vra = P.vals[o.in[3]], &vco = P.vals[o.in[4]];
      k.q = val_in(vq); k.ci = val_in(vci); k.fi = val_in(vfi); k.ra = val_in(vra); k.co = val_in(vco);
      k.aI = adj_in(P.vals[o.out[0]]); k.aO = adj_in(P.vals[o.out[1]]);
      if (!k.aI.v || !k.aO.v) throw std::runtime_error("fused fv_tp_2d: output adjoints missing");
      k.q_ad = adj_out(vq); k.ci_ad = adj_out(vci); k.fi_ad = adj_out(vfi); k.ra_ad = adj_out(vra); k.co_ad = adj_out(vco);
      launch_tile(k, g.NX, g.NY, g.ntile * o.nk_launch);
      return;
    }
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
                        int fx, int fy, const LevOrd& hord, bool full, int nk, bool with_ad) {
  Op op; op.name = nm; op.in = {q, crx, xfx, ra_x, cry, fy2, fxo, mx, my}; op.out = {fx, fy}; op.nk_launch = nk; op.tl_only = P.tl_only;
  op.variant = with_ad ? P.variant : (int)VAR_FWD;
  for (int id : op.in) if (P.vals[id].nk != nk) throw std::runtime_error("fused fv_tp_2d: every field must have the launch's number of levels");
  op.run = [hord, full, with_ad](Program& P, Op& o, int mode) {
    if (mode == MODE_AD) {
      if (!with_ad || full) throw std::runtime_error("fused fv_tp_2d: no reverse kernel for this op");
      const Geom& g = P.dv->g;
      KernTpRev<0, true> k{};
      k.g = g; k.m = P.dv->m; k.ord = hord; k.nk = o.nk_launch;
      const Value &vq = P.vals[o.in[0]], &vci = P.vals[o.in[1]], &vfi = P.vals[o.in[2]], &vra = P.vals[o.in[3]], &vco = P.vals[o.in[4]];
      const Value &vfy2 = P.vals[o.in[5]], &vfxo = P.vals[o.in[6]], &vmx = P.vals[o.in[7]], &vmy = P.vals[o.in[8]];
      k.q = val_in(vq); k.ci = val_in(vci); k.fi = val_in(vfi); k.ra = val_in(vra); k.co = val_in(vco);
      k.fin2 = val_in(vfy2); k.fout2 = val_in(vfxo); k.mI = val_in(vmx); k.mO = val_in(vmy);
      k.aI = adj_in(P.vals[o.out[0]]); k.aO = adj_in(P.vals[o.out[1]]);
      if (!k.aI.v || !k.aO.v) throw std::runtime_error("fused fv_tp_2d: output adjoints missing");
      k.q_ad = adj_out(vq); k.ci_ad = adj_out(vci); k.fi_ad = adj_out(vfi); k.ra_ad = adj_out(vra); k.co_ad = adj_out(vco);
      k.fin2_ad = adj_out(vfy2); k.fout2_ad = adj_out(vfxo); k.mI_ad = adj_out(vmx); k.mO_ad = adj_out(vmy);
      launch_tile(k, g.NX, g.NY, g.ntile * o.nk_launch);
      return;
    }
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
