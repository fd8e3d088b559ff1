// fv_tp_2d forward sweeps (NL, TL) as ROW-MARCHING kernels.
//
// The shared-memory-tile kernels of fused_tp.h spend ~700 instructions per owned cell (ncu, profiles/r02h_ncu_TL_TpB*): every stencil
// value goes through shared memory, every phase recomputes cell indices, and a 32 x 16 tile evaluates 1.4 inner fluxes per owned cell.
// Here a block spans the WHOLE row of a sub-domain (one thread per column i) and marches over MARCH_RY + halo rows of one level:
//   * stencils along j live in REGISTERS: a thread keeps a sliding window of the six rows the PPM flux reads (no shared memory, no
//     index arithmetic: the window shifts by one row per step);
//   * stencils along i go through ONE shared-memory row: write the row, barrier, read the five neighbours;
//   * every global access of a step is the row pointer plus a fixed offset; there is no halo recomputation along i at all and
//     6 / MARCH_RY extra rows along j.
//   kernel A (after copy_corners_y): q -> fy2 = yppm(q, cry) [registers] -> q_i -> fx_ou = xppm(q_i, crx) [row exchange]
//   kernel B (after copy_corners_x): q -> fx2 = xppm(q, crx) [row exchange] -> q_j -> fy_ou = yppm(q_j, cry) [registers] -> fx, fy
// The flux arithmetic is tp::ppm_flux itself on a context whose input is the register window, so the cube-edge cases and every order
// (linear and, for the trajectory side, monotone) are the stage code's own (model_tlmadm/tp_core_tlm.F90:2123-2324, 2328-2660).
// Host emulation: the per-thread state is an explicit struct, a row step is a sequence of phases with a block barrier between them.
#pragma once
#include "fused_tp.h"

namespace fv3lm {
namespace ftp {

#ifndef FV3LM_MARCH_RY
#define FV3LM_MARCH_RY 24
#endif
constexpr int MARCH_RY = FV3LM_MARCH_RY;       // rows a block owns
constexpr int MARCH_MAXT = 384;                // a block has NX rounded up to a warp threads (C360 on one GPU: NX = 367)

// what tp::ppm_flux reads, served from a register window: w[n] = cell (face - 3 + n) along DIR
template <class TT, int DIR> struct WinCtx {
  using T = TT;
  static constexpr int mode = std::is_same<TT, double>::value ? 0 : 1;
  struct { int npx, npy; } g;
  typename LTile<TT>::MM m;
  TT w[6];
  int i, j, mpos, pitch;
  DEV T in(int, int di = 0, int dj = 0, int = 0) const { return w[(DIR == 0 ? di : dj) + 3]; }
  DEV double M(const double* a, int di = 0, int dj = 0) const { return LDG(a + (mpos + dj * pitch + di)); }
};

struct MarchBlk {
  int tile, kk, jj0, ci, cj, pitch, base, mb, NX, NY, xs, xe, ys, ye, ng;
};
DEV MarchBlk make_march_blk(const Geom& g, int nk, int by, int z) {
  MarchBlk b; split_z(z, nk, b.tile, b.kk);
  b.jj0 = by * MARCH_RY;
  const int lo = g.ng - 1;
  b.ci = g.i0[b.tile] - lo; b.cj = g.j0[b.tile] - lo;
  b.pitch = g.pitch; b.base = (b.tile * nk + b.kk) * g.slab; b.mb = b.tile * g.slab; b.NX = g.NX; b.NY = g.NY;
  b.xs = g.is + lo; b.xe = g.ie + lo; b.ys = g.js + lo; b.ye = g.je + lo; b.ng = g.ng;
  return b;
}
template <class TT> struct RowBuf;     // one row of values (and tangents) in shared memory
template <> struct RowBuf<double> {
  double v[MARCH_MAXT];
  DEV void put(int n, double a) { v[n] = a; }
  DEV double get(int n) const { return v[n]; }
};
template <> struct RowBuf<Dual> {
  double v[MARCH_MAXT], d[MARCH_MAXT];
  DEV void put(int n, Dual a) { v[n] = a.v; d[n] = a.d; }
  DEV Dual get(int n) const { return Dual(v[n], d[n]); }
};

// ---- kernel A -------------------------------------------------------------------------------------------------------
template <class TT, bool FULL> struct MarchA {
  static constexpr int NPH = 2;
  Geom g; Metrics m; LevOrd ord; int nk;
  Fld q, cry, yfx, ray, crx; OFld fy2, fxo;
  struct State { TT w[6]; TT fprev; };          // q rows r-3 .. r+2 of this column; inner flux of the previous face
  struct Smem { RowBuf<TT> row[2]; };
  DEV int r_begin(int by) const { return by * MARCH_RY; }
  DEV int r_end(int by) const { return imin(by * MARCH_RY + MARCH_RY, g.NY - 1); }
  DEV void init(int tid, int by, int z, State& st) const {
    using N = Num<TT>;
    const MarchBlk b = make_march_blk(g, nk, by, z);
#pragma unroll
    for (int n = 0; n < 6; n++) st.w[n] = TT(0.0);
    st.fprev = TT(0.0);
    if (tid >= b.NX) return;
#pragma unroll
    for (int n = 1; n < 6; n++) {               // rows jj0-3 .. jj0+1 wait in w[1..5]; the first step shifts them down and loads row jj0+2
      const int jj = b.jj0 - 4 + n;
      if (jj >= 0 && jj < b.NY) st.w[n] = N::ld(q, b.base + jj * b.pitch + tid);
    }
  }
  DEV void step(int ph, int r, int tid, int by, int z, State& st, Smem& s) const {
    using N = Num<TT>;
    const MarchBlk b = make_march_blk(g, nk, by, z);
    const int ii = tid, od = ord.v[b.kk];
    const bool col = ii < b.NX;
    const int c = r - 1;                         // the cell row completed by this face
    const bool row_c = r > b.jj0 && c >= b.ys && c <= b.ye;
    if (ph == 0) {
      if (!col) return;
#pragma unroll
      for (int n = 0; n < 5; n++) st.w[n] = st.w[n + 1];
      st.w[5] = (r + 2 < b.NY) ? N::ld(q, b.base + (r + 2) * b.pitch + ii) : TT(0.0);
      const int o = b.base + r * b.pitch + ii;
      if (r + 3 < b.NY) { pf_fld(q, o + 3 * b.pitch); pf_fld(cry, o + b.pitch); pf_fld(yfx, o + b.pitch); pf_fld(ray, o); pf_fld(crx, o); }
      TT f = TT(0.0);
      const bool col_y = ii <= b.xe + b.ng;      // (isd:ied) = array columns 0 .. xe+ng
      if (col_y && r >= b.ys && r <= b.ye + 1) {
        WinCtx<TT, 1> x; x.g.npx = g.npx; x.g.npy = g.npy; x.m = typename LTile<TT>::MM{m.dxa, m.dya, m.ppmw_x0, m.ppmw_x1, m.ppmw_x2, m.ppmw_x3, m.ppmw_y0, m.ppmw_y1, m.ppmw_y2, m.ppmw_y3}; x.pitch = b.pitch;
#pragma unroll
        for (int n = 0; n < 6; n++) x.w[n] = st.w[n];
        x.i = ii + b.ci; x.j = r + b.cj; x.mpos = b.mb + r * b.pitch + ii;
        f = tp::ppm_flux<1, FULL>(x, 0, N::ld(cry, o), od);
        if (r < b.jj0 + MARCH_RY) N::st(fy2, o, f);
      }
      TT qi = TT(0.0);
      if (col_y && row_c) {
        const int oc = o - b.pitch;
        qi = (st.w[2] * LDG(m.area + (b.mb + c * b.pitch + ii)) + N::ld(yfx, oc) * st.fprev - N::ld(yfx, o) * f) / N::ld(ray, oc);
      }
      st.fprev = f;
      if (r > b.jj0) s.row[r & 1].put(ii, qi);
    } else {
      if (!col || !row_c || ii < b.xs || ii > b.xe + 1) return;
      WinCtx<TT, 0> x; x.g.npx = g.npx; x.g.npy = g.npy; x.m = typename LTile<TT>::MM{m.dxa, m.dya, m.ppmw_x0, m.ppmw_x1, m.ppmw_x2, m.ppmw_x3, m.ppmw_y0, m.ppmw_y1, m.ppmw_y2, m.ppmw_y3}; x.pitch = b.pitch;
      const RowBuf<TT>& rb = s.row[r & 1];
#pragma unroll
      for (int n = 0; n < 6; n++) x.w[n] = rb.get(ii - 3 + n);
      x.i = ii + b.ci; x.j = c + b.cj; x.mpos = b.mb + c * b.pitch + ii;
      const int oc = b.base + c * b.pitch + ii;
      N::st(fxo, oc, tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, oc), od));
    }
  }
};

// ---- kernel B -------------------------------------------------------------------------------------------------------
template <class TT, bool FULL> struct MarchB {
  static constexpr int NPH = 3;
  Geom g; Metrics m; LevOrd ord; int nk;
  Fld q, crx, xfx, rax, cry, fy2, fxo, mx, my; OFld fx, fy;
  struct State { TT w[6]; TT qv; TT fx2; };     // q_j rows r-5 .. r of this column; q and the inner flux of the current row
  struct Smem { RowBuf<TT> qrow, frow; };
  DEV int r_begin(int by) const { return imax(by * MARCH_RY - 3, 0); }
  DEV int r_end(int by) const { return imin(by * MARCH_RY + MARCH_RY + 1, g.NY - 1); }
  DEV void init(int, int, int, State& st) const {
#pragma unroll
    for (int n = 0; n < 6; n++) st.w[n] = TT(0.0);
    st.qv = TT(0.0); st.fx2 = TT(0.0);
  }
  DEV void step(int ph, int r, int tid, int by, int z, State& st, Smem& s) const {
    using N = Num<TT>;
    const MarchBlk b = make_march_blk(g, nk, by, z);
    const int ii = tid, od = ord.v[b.kk];
    if (ii >= b.NX) return;
    const int o = b.base + r * b.pitch + ii;
    const bool row_in = r >= b.ys - b.ng && r <= b.ye + b.ng;       // (jsd:jed)
    if (ph == 0) {
      st.qv = N::ld(q, o);
      s.qrow.put(ii, st.qv);
      if (r + 1 < b.NY && r >= 1) {
        const int o1 = o + b.pitch;
        pf_fld(q, o1); pf_fld(crx, o1); pf_fld(xfx, o1); pf_fld(rax, o1); pf_fld(fxo, o1); pf_fld(mx, o1);
        pf_fld(cry, o1 - 2 * b.pitch); pf_fld(fy2, o1 - 2 * b.pitch); pf_fld(my, o1 - 2 * b.pitch);
      }
    } else if (ph == 1) {
      TT fxx = TT(0.0);
      st.fx2 = TT(0.0);
      if (row_in && ii >= b.xs && ii <= b.xe + 1) {
        WinCtx<TT, 0> x; x.g.npx = g.npx; x.g.npy = g.npy; x.m = typename LTile<TT>::MM{m.dxa, m.dya, m.ppmw_x0, m.ppmw_x1, m.ppmw_x2, m.ppmw_x3, m.ppmw_y0, m.ppmw_y1, m.ppmw_y2, m.ppmw_y3}; x.pitch = b.pitch;
#pragma unroll
        for (int n = 0; n < 6; n++) x.w[n] = s.qrow.get(ii - 3 + n);
        x.i = ii + b.ci; x.j = r + b.cj; x.mpos = b.mb + r * b.pitch + ii;
        st.fx2 = tp::ppm_flux<0, FULL>(x, 0, N::ld(crx, o), od);
        fxx = N::ld(xfx, o) * st.fx2;
      }
      s.frow.put(ii, fxx);
    } else {
      TT qj = TT(0.0);
      if (row_in && ii >= b.xs && ii <= b.xe)
        qj = (st.qv * LDG(m.area + (b.mb + r * b.pitch + ii)) + s.frow.get(ii) - s.frow.get(ii + 1)) / N::ld(rax, o);
#pragma unroll
      for (int n = 0; n < 5; n++) st.w[n] = st.w[n + 1];
      st.w[5] = qj;
      const int f = r - 2;                       // the outer face whose six rows are now complete
      if (f >= b.jj0 && f < b.jj0 + MARCH_RY && f >= b.ys && f <= b.ye + 1 && ii >= b.xs && ii <= b.xe) {
        WinCtx<TT, 1> x; x.g.npx = g.npx; x.g.npy = g.npy; x.m = typename LTile<TT>::MM{m.dxa, m.dya, m.ppmw_x0, m.ppmw_x1, m.ppmw_x2, m.ppmw_x3, m.ppmw_y0, m.ppmw_y1, m.ppmw_y2, m.ppmw_y3}; x.pitch = b.pitch;
#pragma unroll
        for (int n = 0; n < 6; n++) x.w[n] = st.w[n];
        x.i = ii + b.ci; x.j = f + b.cj; x.mpos = b.mb + f * b.pitch + ii;
        const int of = b.base + f * b.pitch + ii;
        const TT fyo = tp::ppm_flux<1, FULL>(x, 0, N::ld(cry, of), od);
        N::st(fy, of, 0.5 * (fyo + N::ld(fy2, of)) * N::ld(my, of));
      }
      if (r >= b.jj0 && r < b.jj0 + MARCH_RY && r >= b.ys && r <= b.ye && ii >= b.xs && ii <= b.xe + 1)
        N::st(fx, o, 0.5 * (N::ld(fxo, o) + st.fx2) * N::ld(mx, o));
    }
  }
};

#ifndef FV3LM_HOST_EMU
template <class K> GLOBAL void __launch_bounds__(MARCH_MAXT) kern_march(const __grid_constant__ K k) {
  __shared__ typename K::Smem s;
  typename K::State st;
  const int tid = threadIdx.x, by = blockIdx.x, z = blockIdx.y;
  k.init(tid, by, z, st);
  const int r1 = k.r_end(by);
  for (int r = k.r_begin(by); r <= r1; r++) {
#pragma unroll
    for (int ph = 0; ph < K::NPH; ph++) {
      k.step(ph, r, tid, by, z, st, s);
      if (ph + 1 < K::NPH) __syncthreads();
    }
  }
}
template <class K> void launch_march(const K& k, int nx, int ny, int nz) {
  if (nz <= 0) return;
  if (nx > MARCH_MAXT) throw std::runtime_error("fv_tp_2d marching kernels: sub-domain rows longer than 384 cells");
  dim3 b((nx + 31) / 32 * 32, 1, 1), gr((ny + MARCH_RY - 1) / MARCH_RY, nz, 1);
  kern_march<K><<<gr, b, 0, dev::stream()>>>(k);
  dev::launches++;
}
#else
template <class K> void launch_march(const K& k, int nx, int ny, int nz) {
  if (nx > MARCH_MAXT) throw std::runtime_error("fv_tp_2d marching kernels: sub-domain rows longer than 384 cells");
  const int nt = (nx + 31) / 32 * 32;
  std::unique_ptr<typename K::Smem> s(new typename K::Smem);
  std::vector<typename K::State> st(nt);
  for (int z = 0; z < nz; z++)
    for (int by = 0; by < (ny + MARCH_RY - 1) / MARCH_RY; by++) {
      memset(s.get(), 0xff, sizeof(typename K::Smem));              // a fresh block finds arbitrary shared memory
      for (int tid = 0; tid < nt; tid++) k.init(tid, by, z, st[tid]);
      for (int r = k.r_begin(by); r <= k.r_end(by); r++)
        for (int ph = 0; ph < K::NPH; ph++)
          for (int tid = 0; tid < nt; tid++) k.step(ph, r, tid, by, z, st[tid], *s);
    }
  dev::launches++;
}
#endif

inline bool march_enabled(int nx) {
  const char* e = getenv("FV3LM_TP_MARCH");          // read per call (a host-side branch per op): tests switch it
  const int on = e ? atoi(e) : 0;   // opt-in: measured slower than the tile kernels (profiles/r02k_*)
  return on != 0 && nx <= MARCH_MAXT;
}

template <template <class, bool> class K, class Fill>
void run_march(Program& P, Op& o, int mode, bool full, const Fill& fill) {
  const Geom& g = P.dv->g;
  bool tl = false;
  if (mode == MODE_TL) for (int i : o.in) tl = tl || (P.vals[i].active && P.vals[i].pert);
  auto go = [&](auto kern) {
    kern.g = g; kern.m = P.dv->m; kern.nk = o.nk_launch;
    fill(kern, tl);
    launch_march(kern, g.NX, g.NY, g.ntile * o.nk_launch);
  };
  if (tl) { if (full) go(K<Dual, true>{}); else go(K<Dual, false>{}); }
  else { if (full) go(K<double, true>{}); else go(K<double, false>{}); }
}
// forward sweeps of the two fused fv_tp_2d ops (declared in fused_tp.h, called from add_fused_a / add_fused_b)
inline bool march_fwd_a(Program& P, Op& o, int mode, bool full, const LevOrd& hord) {
  if (!march_enabled(P.dv->g.NX)) return false;
  run_march<MarchA>(P, o, mode, full, [&](auto& k, bool tl) {
    k.ord = hord;
    k.q = fld(P.vals[o.in[0]], tl); k.cry = fld(P.vals[o.in[1]], tl); k.yfx = fld(P.vals[o.in[2]], tl);
    k.ray = fld(P.vals[o.in[3]], tl); k.crx = fld(P.vals[o.in[4]], tl);
    k.fy2 = ofld(P.vals[o.out[0]], tl); k.fxo = ofld(P.vals[o.out[1]], tl);
  });
  return true;
}
inline bool march_fwd_b(Program& P, Op& o, int mode, bool full, const LevOrd& hord) {
  if (!march_enabled(P.dv->g.NX)) return false;
  run_march<MarchB>(P, o, mode, full, [&](auto& k, bool tl) {
    k.ord = hord;
    k.q = fld(P.vals[o.in[0]], tl); k.crx = fld(P.vals[o.in[1]], tl); k.xfx = fld(P.vals[o.in[2]], tl); k.rax = fld(P.vals[o.in[3]], tl);
    k.cry = fld(P.vals[o.in[4]], tl); k.fy2 = fld(P.vals[o.in[5]], tl); k.fxo = fld(P.vals[o.in[6]], tl);
    k.mx = fld(P.vals[o.in[7]], tl); k.my = fld(P.vals[o.in[8]], tl);
    k.fx = ofld(P.vals[o.out[0]], tl); k.fy = ofld(P.vals[o.out[1]], tl);
  });
  return true;
}

}  // namespace ftp
}  // namespace fv3lm
