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
// L2 prefetch of what a later phase (tile kernels) or the next row (marching kernels) reads: these kernels start every phase with global
// loads followed by a block barrier and were found latency bound (ncu r02g / r02j: long-scoreboard stalls 14 per issue at < 50 % occupancy)
DEV void pf_l2(const double* p) {
#ifndef FV3LM_HOST_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
DEV void pf_fld(const Fld& f, int o) { pf_l2(f.v + o); if (f.d) pf_l2(f.d + o); }

template <bool TLM, int N> struct SBuf { double v[N]; double d[TLM ? N : 1]; };

// the context tp::ppm_flux sees: the transported field (its only input) is a shared-memory tile, metrics stay global
template <class TT> struct TileCtx : CtxBase {
  using T = TT;
  static constexpr int mode = std::is_same<TT, double>::value ? 0 : 1;
  const double* sv; const double* sd; int sw, sp;     // tile (value, perturbation), its row width, position of the current cell
  DEV T in(int, int di = 0, int dj = 0, int = 0) const { return Num<TT>::lds(sv, sd, sp + dj * sw + di); }
};

// ---- lean tile machinery ---------------------------------------------------------------------------------------------
// ncu of the first version of these kernels (profiles/r02e_*): 850 instructions per owned cell in the TL kernel B, 22 % of them fp64 --
// the rest was per-cell index set-up (CtxBase::setpos / off with level clamps, rectangle tests, box look-ups).  Here everything that is
// uniform over a block (level offset, rectangles clipped to the tile, index origin) is computed once per phase, a cell costs one
// multiply-add for its offset, and the flux code sees a context that holds only what it reads.
struct Rect {
  int x0, x1, y0, y1;                       // inclusive, array coordinates
  DEV bool has(int ii, int jj) const { return ii >= x0 && ii <= x1 && jj >= y0 && jj <= y1; }
};
DEV int imax(int a, int b) { return a > b ? a : b; }
DEV int imin(int a, int b) { return a < b ? a : b; }
struct Blk {                                // block-uniform geometry
  int tile, kk, ii0, jj0;                   // sub-domain, level, array coordinates of the block's first cell
  int ci, cj;                               // tile-global Fortran index = array index + ci / cj
  int pitch, base, mb, NX, NY;              // row pitch, offset of (tile, level) in a 3-D field, of the tile in a 2-D metric
  int xs, xe, ys, ye, ng;                   // compute domain is..ie, js..je in array coordinates; halo width
  // the rectangle (il0..il1, jl0..jl1) of the local frame, shifted by (a, b, c, d), in array coordinates
  DEV Rect rect(int a, int b, int c, int d) const { return Rect{xs + a, xe + b, ys + c, ye + d}; }
  DEV Rect clip(Rect r, int x0, int x1, int y0, int y1) const {
    return Rect{imax(imax(r.x0, x0), 0), imin(imin(r.x1, x1), NX - 1), imax(imax(r.y0, y0), 0), imin(imin(r.y1, y1), NY - 1)};
  }
  DEV int off(int ii, int jj) const { return base + jj * pitch + ii; }
  DEV bool inside(int ii, int jj) const { return (unsigned)ii < (unsigned)NX && (unsigned)jj < (unsigned)NY; }
};
DEV Blk make_blk(const Geom& g, int nk, int bx, int by, int z) {
  Blk b; split_z(z, nk, b.tile, b.kk);
  b.ii0 = bx * TX; b.jj0 = by * TY;
  const int lo = g.ng - 1;
  b.ci = g.i0[b.tile] - lo; b.cj = g.j0[b.tile] - lo;
  b.pitch = g.pitch; b.base = (b.tile * nk + b.kk) * g.slab; b.mb = b.tile * g.slab; b.NX = g.NX; b.NY = g.NY;
  b.xs = g.is + lo; b.xe = g.ie + lo; b.ys = g.js + lo; b.ye = g.je + lo; b.ng = g.ng;
  return b;
}
DEV Rect isect(Rect a, Rect b) { return Rect{imax(a.x0, b.x0), imin(a.x1, b.x1), imax(a.y0, b.y0), imin(a.y1, b.y1)}; }
DEV bool rempty(Rect r) { return r.x1 < r.x0 || r.y1 < r.y0; }
// "lean rectangle + generic complement": a phase evaluates its regular cells (a block-uniform rectangle: away from the cube edges and
// the borders of the compute domain) with straight-line code, and the few cells of its box outside that rectangle -- up to four thin
// strips, empty for most blocks -- with the stage's own generic eval(), densely mapped onto the block's threads.
template <class F> DEV void for_strip(int tid, Rect s, F f) {
  if (rempty(s)) return;
  const int w = s.x1 - s.x0 + 1, n = w * (s.y1 - s.y0 + 1);
  for (int c = tid; c < n; c += NTHR) { const int r = c / w; f(s.x0 + c - r * w, s.y0 + r); }
}
template <class F> DEV void for_complement(int tid, Rect box, Rect rr, F f) {
  const Rect in = isect(box, rr);
  if (rempty(in)) { for_strip(tid, box, f); return; }
  for_strip(tid, Rect{box.x0, box.x1, box.y0, in.y0 - 1}, f);
  for_strip(tid, Rect{box.x0, box.x1, in.y1 + 1, box.y1}, f);
  for_strip(tid, Rect{box.x0, in.x0 - 1, in.y0, in.y1}, f);
  for_strip(tid, Rect{in.x1 + 1, box.x1, in.y0, in.y1}, f);
}
// what tp::ppm_flux / ppm_coef / ppm_dc / S_ppm::al_w read: the tile, the Fortran index of the cell, npx / npy, dxa / dya
template <class TT> struct LTile {
  using T = TT;
  static constexpr int mode = std::is_same<TT, double>::value ? 0 : 1;
  struct { int npx, npy; } g;
  struct MM { const double *dxa, *dya, *ppmw_x0, *ppmw_x1, *ppmw_x2, *ppmw_x3, *ppmw_y0, *ppmw_y1, *ppmw_y2, *ppmw_y3; } m;
  const double* sv; const double* sd; int sw, sp;
  int i, j, mpos, pitch;
  DEV void init(const Geom& g_, const Metrics& m_) {
    g.npx = g_.npx; g.npy = g_.npy; pitch = g_.pitch; sd = nullptr;
    m = MM{m_.dxa, m_.dya, m_.ppmw_x0, m_.ppmw_x1, m_.ppmw_x2, m_.ppmw_x3, m_.ppmw_y0, m_.ppmw_y1, m_.ppmw_y2, m_.ppmw_y3};
  }
  DEV void at(const Blk& b, int ii, int jj) { i = ii + b.ci; j = jj + b.cj; mpos = b.mb + jj * b.pitch + ii; }
  DEV T in(int, int di = 0, int dj = 0, int = 0) const { return Num<TT>::lds(sv, sd, sp + dj * sw + di); }
  DEV double M(const double* a, int di = 0, int dj = 0) const { return LDG(a + (mpos + dj * pitch + di)); }
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
    const Blk b = make_blk(g, nk, bx, by, z);
    const int ii0 = b.ii0, jj0 = b.jj0, od = ord.v[b.kk];
    if (ph == 0) {                 // q with its footprint
      for (int c = tid; c < QW * QH; c += NTHR) {
        const int ii = ii0 - 3 + c % QW, jj = jj0 - 3 + c / QW;
        TT a = TT(0.0);
        if (b.inside(ii, jj)) a = N::ld(q, b.off(ii, jj));
        N::sts(s.q.v, s.q.d, c, a);
      }
    } else if (ph == 1) {          // fy2 = yppm(q, cry) on (isd:ied, js:je+1)
      const Rect r = b.clip(b.rect(-b.ng, b.ng, 0, 1), ii0 - 3, ii0 + TX + 2, jj0, jj0 + TY);
      LTile<TT> x; x.init(g, m); x.sv = s.q.v; x.sd = s.q.d; x.sw = QW;
      for (int c = tid; c < QW * (TY + 1); c += NTHR) {
        const int cx = c % QW, cy = c / QW, ii = ii0 - 3 + cx, jj = jj0 + cy;
        if (!r.has(ii, jj)) continue;
        const int o = b.off(ii, jj);
        x.at(b, ii, jj); x.sp = (cy + 3) * QW + cx;
        const TT f = tp::ppm_flux<1, FULL>(x, 0, N::ld(cry, o), od);
        N::sts(s.fy.v, s.fy.d, c, f);
        if (cx >= 3 && cx < TX + 3 && cy < TY) N::st(fy2, o, f);
      }
    } else if (ph == 2) {          // q_i = (q area + yfx fy2 (j) - yfx fy2 (j+1)) / ra_y on (isd:ied, js:je)
      const Rect r = b.clip(b.rect(-b.ng, b.ng, 0, 0), ii0 - 3, ii0 + TX + 2, jj0, jj0 + TY - 1);
      for (int c = tid; c < QW * TY; c += NTHR) {
        const int cx = c % QW, cy = c / QW, ii = ii0 - 3 + cx, jj = jj0 + cy;
        if (!r.has(ii, jj)) continue;
        const int o = b.off(ii, jj);
        const TT f0 = N::ld(yfx, o) * N::lds(s.fy.v, s.fy.d, c);
        const TT f1 = N::ld(yfx, o + b.pitch) * N::lds(s.fy.v, s.fy.d, c + QW);
        const TT qq = N::lds(s.q.v, s.q.d, (cy + 3) * QW + cx);
        N::sts(s.qi.v, s.qi.d, c, (qq * LDG(m.area + (b.mb + jj * b.pitch + ii)) + f0 - f1) / N::ld(ray, o));
      }
    } else {                       // fx_ou = xppm(q_i, crx) on (is:ie+1, js:je)
      const Rect r = b.clip(b.rect(0, 1, 0, 0), ii0, ii0 + TX - 1, jj0, jj0 + TY - 1);
      LTile<TT> x; x.init(g, m); x.sv = s.qi.v; x.sd = s.qi.d; x.sw = QW;
      for (int c = tid; c < TX * TY; c += NTHR) {
        const int cx = c % TX, cy = c / TX, ii = ii0 + cx, jj = jj0 + cy;
        if (!r.has(ii, jj)) continue;
        const int o = b.off(ii, jj);
        x.at(b, ii, jj); x.sp = cy * QW + cx + 3;
        N::st(fxo, o, tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, o), od));
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
    const Blk b = make_blk(g, nk, bx, by, z);
    const int ii0 = b.ii0, jj0 = b.jj0, od = ord.v[b.kk];
    if (ph == 0) {
      for (int c = tid; c < QW * QH; c += NTHR) {
        const int ii = ii0 - 3 + c % QW, jj = jj0 - 3 + c / QW;
        TT a = TT(0.0);
        if (b.inside(ii, jj)) a = N::ld(q, b.off(ii, jj));
        N::sts(s.q.v, s.q.d, c, a);
      }
    } else if (ph == 1) {          // fx2 = xppm(q, crx) on (is:ie+1, jsd:jed)
      const Rect r = b.clip(b.rect(0, 1, -b.ng, b.ng), ii0, ii0 + TX, jj0 - 3, jj0 + TY + 2);
      LTile<TT> x; x.init(g, m); x.sv = s.q.v; x.sd = s.q.d; x.sw = QW;
      for (int c = tid; c < FW * QH; c += NTHR) {
        const int cx = c % FW, cy = c / FW, ii = ii0 + cx, jj = jj0 - 3 + cy;
        if (!r.has(ii, jj)) continue;
        x.at(b, ii, jj); x.sp = cy * QW + cx + 3;
        N::sts(s.fx.v, s.fx.d, c, tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, b.off(ii, jj)), od));
      }
    } else if (ph == 2) {          // q_j = (q area + xfx fx2 (i) - xfx fx2 (i+1)) / ra_x on (is:ie, jsd:jed)
      const Rect r = b.clip(b.rect(0, 0, -b.ng, b.ng), ii0, ii0 + TX - 1, jj0 - 3, jj0 + TY + 2);
      for (int c = tid; c < TX * QH; c += NTHR) {
        const int cx = c % TX, cy = c / TX, ii = ii0 + cx, jj = jj0 - 3 + cy;
        if (!r.has(ii, jj)) continue;
        const int o = b.off(ii, jj), cf = cy * FW + cx;
        const TT f0 = N::ld(xfx, o) * N::lds(s.fx.v, s.fx.d, cf);
        const TT f1 = N::ld(xfx, o + 1) * N::lds(s.fx.v, s.fx.d, cf + 1);
        const TT qq = N::lds(s.q.v, s.q.d, cy * QW + cx + 3);
        N::sts(s.qj.v, s.qj.d, c, (qq * LDG(m.area + (b.mb + jj * b.pitch + ii)) + f0 - f1) / N::ld(rax, o));
      }
    } else {                       // fy_ou = yppm(q_j, cry) on (is:ie, js:je+1); the two averages (tp_core_tlm.F90:2268-2313)
      const Rect ry = b.clip(b.rect(0, 0, 0, 1), ii0, ii0 + TX - 1, jj0, jj0 + TY - 1);
      const Rect rx = b.clip(b.rect(0, 1, 0, 0), ii0, ii0 + TX - 1, jj0, jj0 + TY - 1);
      LTile<TT> x; x.init(g, m); x.sv = s.qj.v; x.sd = s.qj.d; x.sw = TX;
      for (int c = tid; c < TX * TY; c += NTHR) {
        const int cx = c % TX, cy = c / TX, ii = ii0 + cx, jj = jj0 + cy;
        const int o = b.off(ii, jj);
        if (ry.has(ii, jj)) {
          x.at(b, ii, jj); x.sp = (cy + 3) * TX + cx;
          const TT fyo = tp::ppm_flux<1, FULL>(x, 0, N::ld(cry, o), od);
          N::st(fy, o, 0.5 * (fyo + N::ld(fy2, o)) * N::ld(my, o));
        }
        if (rx.has(ii, jj)) {
          const TT fx2 = N::lds(s.fx.v, s.fx.d, (cy + 3) * FW + cx);
          N::st(fx, o, 0.5 * (N::ld(fxo, o) + fx2) * N::ld(mx, o));
        }
      }
    }
  }
};

#ifndef FV3LM_HOST_EMU
#ifndef FV3LM_TILE_MINBLOCKS
#define FV3LM_TILE_MINBLOCKS 5     // 48 registers: five resident blocks instead of four (reverse kernels 1.79 -> 1.58 ms, profiles/r02t_*)
#endif
template <class K> GLOBAL void __launch_bounds__(NTHR, FV3LM_TILE_MINBLOCKS) kern_tile(const __grid_constant__ K k) {
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
  {
    // regular faces: straight-line code, no edge tests (the same fast path as tp::ppm_flux)
    const int ia = DIR == 0 ? x.i : x.j, np = DIR == 0 ? x.g.npx : x.g.npy;
    if (ia >= 4 && ia <= np - 3) {
      const double qm2 = tp::Q<DIR>(x, 0, -2), qm1 = tp::Q<DIR>(x, 0, -1), q0 = tp::Q<DIR>(x, 0, 0), q1 = tp::Q<DIR>(x, 0, 1);
      const double al0 = tp::p1 * (qm1 + q0) + tp::p2 * (qm2 + q1);
      if (c > 0.0) {
        const double b = tp::p1 * (qm2 + qm1) + tp::p2 * (tp::Q<DIR>(x, 0, -3) + q0) + al0 - (qm1 + qm1);
        return -(al0 - qm1 - c * b) - (1.0 - c) * b;
      }
      const double b = al0 + (tp::p1 * (q0 + q1) + tp::p2 * (qm1 + tp::Q<DIR>(x, 0, 2))) - (q0 + q0);
      return (al0 - q0 + c * b) + (1.0 + c) * b;
    }
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
  static constexpr int NPH = 6;
  static constexpr int TA = DIN == 0 ? TX : TY, TB = DIN == 0 ? TY : TX;
  Geom g; Metrics m; LevOrd ord; int nk;
  // values: transported field, inner Courant number / flux area / ra, outer Courant number; AVG: the other two fluxes and the multipliers
  Fld q, ci, fi, ra, co, fin2, fout2, mI, mO;
  // incoming adjoints (.v = adjoint array): !AVG: aI = fy2_ad (inner flux), aO = fx_ou_ad (outer flux); AVG: aI = fx_ad, aO = fy_ad
  Fld aI, aO;
  // accumulated adjoints (.v = adjoint array, null = inactive)
  OFld q_ad, ci_ad, fi_ad, ra_ad, co_ad, fin2_ad, fout2_ad, mI_ad, mO_ad;
  struct GenSmem {
    double q[QW * QH], Fi[(TA + 1) * (TB + 6)], qm[TA * (TB + 6)];
    double Fo_ad[(TA + 6) * (TB + 5)], qm_ad[(TA + 6) * TB], Fi_ad[(TA + 5) * TB];
  };
  // ---- lean path (unlimited PPM, ord = 2): boxes with compile-time extents, relative to the block origin.  A box spans
  // [a0 + LI, a0 + TA - 1 + HI] along the inner axis and [b0 + LO, b0 + TB - 1 + HO] along the outer one.
  template <int LI, int HI, int LO, int HO> struct CB {
    static constexpr int NI_ = TA + HI - LI, NO_ = TB + HO - LO;
    static constexpr int W = DIN == 0 ? NI_ : NO_, H = DIN == 0 ? NO_ : NI_;
    static constexpr int X0 = DIN == 0 ? LI : LO, Y0 = DIN == 0 ? LO : LI;     // relative to (ii0, jj0)
    static constexpr int N = W * H;
    static constexpr int SI = DIN == 0 ? 1 : W, SO = DIN == 0 ? W : 1;          // index steps along the inner / outer axis
    DEV static int idx(int rx, int ry) { return (ry - Y0) * W + (rx - X0); }
  };
  using BQ = CB<-3, 3, -3, 3>;      // transported field
  using BFI = CB<0, 1, -3, 3>;      // inner flux (values)
  using BQM = CB<0, 0, -3, 3>;      // updated field (values)
  using BT = CB<-3, 3, 0, 0>;       // t = qm_ad / ra
  using BAO = CB<-3, 3, -2, 3>;     // adjoint of the outer flux and its Courant number
  using BALO = CB<-3, 3, -1, 2>;    // adjoint of the outer sweep's edge values
  using BFIAD = CB<-2, 3, 0, 0>;    // adjoint of the inner flux and its Courant number
  using BALI = CB<-1, 2, 0, 0>;     // adjoint of the inner sweep's edge values
  static constexpr int NU = (2 * BAO::N + BALO::N) > (2 * BFIAD::N + BALI::N) ? (2 * BAO::N + BALO::N) : (2 * BFIAD::N + BALI::N);
  struct LeanSmem { double q[BQ::N], Fi[BFI::N], qm[BQM::N], t[BT::N], u[NU]; };   // u: outer-sweep tiles, then the inner-sweep ones
  union Smem { GenSmem gen; LeanSmem lean; };
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
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& su) const {
    int tile_, kk_; split_z(z, nk, tile_, kk_);
    if (ord.v[kk_] == 2) lean_phase(ph, tid, bx, by, z, su.lean);
    else if (ph < 5) gen_phase(ph, tid, bx, by, z, su.gen);
  }
  // d F / d qt, d F / d al0, d F / d (alm | alp) of the unlimited PPM flux, S_ppm::adjoint's formulas
  DEV static void flux_partials(double c, double& dqt, double& dal0, double& dal2) {
    if (c > 0.0) { dqt = 1.0 + (1.0 - c) * (2.0 * c - 1.0); dal0 = (1.0 - c) * (1.0 - c); dal2 = -(1.0 - c) * c; }
    else { dqt = 1.0 - (1.0 + c) * (1.0 + 2.0 * c); dal0 = (1.0 + c) * (1.0 + c); dal2 = (1.0 + c) * c; }
  }
  // adjoint of the edge value AL(e) from the flux adjoints A and Courant numbers C of the faces e-1, e, e+1 (tile index n, step st):
  // AL(e) is al0 of face e, alm of face e+1 (c > 0) and alp of face e-1 (c <= 0)
  DEV static double edge_adjoint(const double* A, const double* C, int n, int st) {
    double dqt, d0, d2, r = 0.0;
    const double a0 = A[n], ap = A[n + st], am = A[n - st];
    if (a0 != 0.0) { flux_partials(C[n], dqt, d0, d2); r += a0 * d0; }
    if (ap != 0.0) { const double c = C[n + st]; if (c > 0.0) { flux_partials(c, dqt, d0, d2); r += ap * d2; } }
    if (am != 0.0) { const double c = C[n - st]; if (!(c > 0.0)) { flux_partials(c, dqt, d0, d2); r += am * d2; } }
    return r;
  }
  // adjoint of the cell value q(p) of a 1-D sweep along D: the faces p (c <= 0) and p+1 (c > 0) that use it as qt, and the four
  // edge values AL(p-1 .. p+2) that it enters.  A, C: face tiles at index n (step st); AL: edge tile at index ne (step se).
  template <int D, class X> DEV static double cell_adjoint(X& x, const Blk& b, int ii, int jj, const double* A, const double* C, int n, int st,
                                                       const double* AL, int ne, int se) {
    double dqt, d0, d2, r = 0.0;
    const double a0 = A[n], ap = A[n + st];
    if (a0 != 0.0) { const double c = C[n]; if (!(c > 0.0)) { flux_partials(c, dqt, d0, d2); r += a0 * dqt; } }
    if (ap != 0.0) { const double c = C[n + st]; if (c > 0.0) { flux_partials(c, dqt, d0, d2); r += ap * dqt; } }
    const int pos = (D == 0 ? ii + b.ci : jj + b.cj), np = D == 0 ? x.g.npx : x.g.npy;
    if (pos >= 4 && pos <= np - 4) {
      r += tp::p2 * (AL[ne - se] + AL[ne + 2 * se]) + tp::p1 * (AL[ne] + AL[ne + se]);
    } else {
#pragma unroll
      for (int e = -1; e <= 2; e++) {
        const double al = AL[ne + e * se];
        if (al != 0.0) { x.at(b, ii + (D == 0 ? e : 0), jj + (D == 1 ? e : 0)); r += S_ppm<D>::al_w(x, 0, 2 - e) * al; }
      }
    }
    return r;
  }
  DEV void lean_phase(int ph, int tid, int bx, int by, int z, LeanSmem& s) const {
    const Blk b = make_blk(g, nk, bx, by, z);
    const int ii0 = b.ii0, jj0 = b.jj0;
    constexpr int eIx = DIN == 0, eIy = DIN == 1;
    const int sI = DIN == 0 ? 1 : b.pitch;                       // global index step along the inner axis
    const Rect r_i = DIN == 0 ? b.rect(0, 1, -b.ng, b.ng) : b.rect(-b.ng, b.ng, 0, 1);
    const Rect r_m = DIN == 0 ? b.rect(0, 0, -b.ng, b.ng) : b.rect(-b.ng, b.ng, 0, 0);
    const Rect r_o = DIN == 0 ? b.rect(0, 0, 0, 1) : b.rect(0, 1, 0, 0);
    const Rect r_fi = b.rect(0, 1, 0, 0);
    double* AO = s.u; double* CO = s.u + BAO::N; double* ALO = s.u + 2 * BAO::N;
    double* FA = s.u; double* CI = s.u + BFIAD::N; double* ALI = s.u + 2 * BFIAD::N;
    LTile<double> x; x.init(g, m);
    if (ph == 0) {
      for (int c = tid; c < BQ::N; c += NTHR) {
        const int ii = ii0 + BQ::X0 + c % BQ::W, jj = jj0 + BQ::Y0 + c / BQ::W;
        s.q[c] = b.inside(ii, jj) ? LDG(q.v + b.off(ii, jj)) : 0.0;
      }
      for (int c = tid; c < BAO::N; c += NTHR) {
        const int ii = ii0 + BAO::X0 + c % BAO::W, jj = jj0 + BAO::Y0 + c / BAO::W;
        double a = 0.0, cc = 0.0;
        if (r_o.has(ii, jj)) {
          const int o = b.off(ii, jj);
          a = LDG(aO.v + o); cc = LDG(co.v + o);
          if (AVG && a != 0.0) a *= 0.5 * LDG(mO.v + o);
        }
        AO[c] = a; CO[c] = cc;
      }
    } else if (ph == 1) {
      x.sv = s.q; x.sw = BQ::W;
      for (int c = tid; c < BFI::N; c += NTHR) {
        const int rx = BFI::X0 + c % BFI::W, ry = BFI::Y0 + c / BFI::W, ii = ii0 + rx, jj = jj0 + ry;
        if (!r_i.has(ii, jj)) continue;
        x.at(b, ii, jj); x.sp = BQ::idx(rx, ry);
        s.Fi[c] = tp::ppm_flux<DIN, false>(x, 0, LDG(ci.v + b.off(ii, jj)), 2);
      }
      for (int c = tid; c < BALO::N; c += NTHR) {
        const int rx = BALO::X0 + c % BALO::W, ry = BALO::Y0 + c / BALO::W;
        ALO[c] = edge_adjoint(AO, CO, BAO::idx(rx, ry), BAO::SO);
      }
    } else if (ph == 2) {
      for (int c = tid; c < BQM::N; c += NTHR) {
        const int rx = BQM::X0 + c % BQM::W, ry = BQM::Y0 + c / BQM::W, ii = ii0 + rx, jj = jj0 + ry;
        if (!r_m.has(ii, jj)) continue;
        const int o = b.off(ii, jj), nf = BFI::idx(rx, ry);
        const double f0 = LDG(fi.v + o) * s.Fi[nf], f1 = LDG(fi.v + o + sI) * s.Fi[nf + BFI::SI];
        s.qm[c] = (s.q[BQ::idx(rx, ry)] * LDG(m.area + (b.mb + jj * b.pitch + ii)) + f0 - f1) / LDG(ra.v + o);
      }
      for (int c = tid; c < BT::N; c += NTHR) {
        const int rx = BT::X0 + c % BT::W, ry = BT::Y0 + c / BT::W, ii = ii0 + rx, jj = jj0 + ry;
        double t = 0.0;
        if (r_m.has(ii, jj)) {
          const double qa = cell_adjoint<DOUT>(x, b, ii, jj, AO, CO, BAO::idx(rx, ry), BAO::SO, ALO, BALO::idx(rx, ry), BALO::SO);
          if (qa != 0.0) t = qa / LDG(ra.v + b.off(ii, jj));
        }
        s.t[c] = t;
      }
    } else if (ph == 3) {
      for (int c = tid; c < BFIAD::N; c += NTHR) {
        const int rx = BFIAD::X0 + c % BFIAD::W, ry = BFIAD::Y0 + c / BFIAD::W, ii = ii0 + rx, jj = jj0 + ry;
        double a = 0.0, cc = 0.0;
        if (r_i.has(ii, jj)) {
          const int o = b.off(ii, jj);
          cc = LDG(ci.v + o);
          if (AVG) { if (r_fi.has(ii, jj)) { a = LDG(aI.v + o); if (a != 0.0) a *= 0.5 * LDG(mI.v + o); } }
          else a = LDG(aI.v + o);
          const int nt = BT::idx(rx, ry);
          const double dt = s.t[nt] - s.t[nt - BT::SI];
          if (dt != 0.0) a += LDG(fi.v + o) * dt;
        }
        FA[c] = a; CI[c] = cc;
      }
    } else if (ph == 4) {
      for (int c = tid; c < BALI::N; c += NTHR) {
        const int rx = BALI::X0 + c % BALI::W, ry = BALI::Y0 + c / BALI::W;
        ALI[c] = edge_adjoint(FA, CI, BFIAD::idx(rx, ry), BFIAD::SI);
      }
    } else {
      for (int c = tid; c < TX * TY; c += NTHR) {
        const int rx = c % TX, ry = c / TX, ii = ii0 + rx, jj = jj0 + ry;
        if (!b.inside(ii, jj)) continue;
        const int o = b.off(ii, jj);
        const bool in_i = r_i.has(ii, jj), in_m = r_m.has(ii, jj), in_o = r_o.has(ii, jj);
        const int nt = BT::idx(rx, ry);
        // (requesting all nine accumulators ahead of the stores was tried: 64 -> 112 registers, 1.9 -> 2.8 ms per launch, profiles/r02m_ncu_TpRevB.txt)
        double aq = 0.0;
        if (in_m) {
          const double tt = s.t[nt];
          if (tt != 0.0) {
            aq += tt * LDG(m.area + (b.mb + jj * b.pitch + ii));
            if (ra_ad.v) ra_ad.v[o] += -s.qm[BQM::idx(rx, ry)] * tt;
          }
        }
        if (in_i) {
          if (fi_ad.v) {
            const double dt = s.t[nt] - s.t[nt - BT::SI];
            if (dt != 0.0) fi_ad.v[o] += s.Fi[BFI::idx(rx, ry)] * dt;
          }
          const double fa = FA[BFIAD::idx(rx, ry)];
          if (ci_ad.v && fa != 0.0) {
            x.sv = s.q; x.sw = BQ::W; x.sp = BQ::idx(rx, ry); x.at(b, ii, jj);
            ci_ad.v[o] += ppm_dc<DIN>(x, LDG(ci.v + o), 2) * fa;
          }
        }
        if (in_o) {
          double oa = LDG(aO.v + o);
          if (AVG && oa != 0.0) oa *= 0.5 * LDG(mO.v + o);
          if (co_ad.v && oa != 0.0) {
            x.sv = s.qm; x.sw = BQM::W; x.sp = BQM::idx(rx, ry); x.at(b, ii, jj);
            co_ad.v[o] += ppm_dc<DOUT>(x, LDG(co.v + o), 2) * oa;
          }
        }
        if (AVG) {
          if (in_o) {             // fy = 0.5 (Fo + fy2) mO
            const double a = LDG(aO.v + o);
            if (a != 0.0) {
              if (fin2_ad.v) fin2_ad.v[o] += 0.5 * LDG(mO.v + o) * a;
              if (mO_ad.v) {
                x.sv = s.qm; x.sw = BQM::W; x.sp = BQM::idx(rx, ry); x.at(b, ii, jj);
                const double Fo = tp::ppm_flux<DOUT, false>(x, 0, LDG(co.v + o), 2);
                mO_ad.v[o] += 0.5 * (Fo + LDG(fin2.v + o)) * a;
              }
            }
          }
          if (r_fi.has(ii, jj)) {       // fx = 0.5 (fx_ou + Fi) mI
            const double a = LDG(aI.v + o);
            if (a != 0.0) {
              if (fout2_ad.v) fout2_ad.v[o] += 0.5 * LDG(mI.v + o) * a;
              if (mI_ad.v) mI_ad.v[o] += 0.5 * (LDG(fout2.v + o) + s.Fi[BFI::idx(rx, ry)]) * a;
            }
          }
        }
        if (q_ad.v) {
          aq += cell_adjoint<DIN>(x, b, ii, jj, FA, CI, BFIAD::idx(rx, ry), BFIAD::SI, ALI, BALI::idx(rx, ry), BALI::SI);
          if (aq != 0.0) q_ad.v[o] += aq;
        }
      }
    }
  }
  DEV void gen_phase(int ph, int tid, int bx, int by, int z, GenSmem& s) const {
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

inline Fld adj_in(const Value& v) { return Fld{v.active ? v.pert : nullptr, nullptr, v.nk}; }
inline OFld adj_out(const Value& v) { return OFld{v.active ? v.pert : nullptr, nullptr, v.nk}; }
inline Fld val_in(const Value& v) { return Fld{v.traj, nullptr, v.nk}; }

}  // namespace ftp
}  // namespace fv3lm
