// fv_dynamics: the model-step driver (model/fv_dynamics_nlm.F90:70-760; TL fv_dynamics_tlm.F90:87,
// AD fv_dynamics_adm.F90:110/874) = theta_v conversion, k_split x (dyn_core, tracer_2d,
// Lagrangian_to_Eulerian), and the fv3jedi_lm dynamics-component wrapper
// (src/dynamics/fv3jedi_lm_dynamics_mod.F90:268-689).
#include "fvdyn.h"
#include "nh.h"
#include "modules.h"
#include "comm.h"
#include "stages_c2l.h"
#include <memory>

namespace fv3lm {

// ---------------------------------------------------------------------------------
// tracer_2d (model/fv_tracer2d_nlm.F90:275-516, q_split = 1)
// ---------------------------------------------------------------------------------
// xfx / yfx from the accumulated Courant numbers (:329-349).  in: cx cy ; out: xfx yfx
struct S_trc_fx {
  static constexpr int NI = 2, NO = 2;
  struct P { int dummy; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int isd = g.is - g.ng, ied = g.ie + g.ng, jsd = g.js - g.ng, jed = g.je + g.ng;
    if (x.in_rect(g.is, g.ie + 1, jsd, jed)) {
      T c = x.in(0);
      x.out(0, val(c) > 0.0 ? c * x.M(x.m.dxa, -1, 0) * x.M(x.m.dy) * x.M(x.m.sin_sg3, -1, 0) : c * x.M(x.m.dxa) * x.M(x.m.dy) * x.M(x.m.sin_sg1));
    }
    if (x.in_rect(isd, ied, g.js, g.je + 1)) {
      T c = x.in(1);
      x.out(1, val(c) > 0.0 ? c * x.M(x.m.dya, 0, -1) * x.M(x.m.dx) * x.M(x.m.sin_sg4, 0, -1) : c * x.M(x.m.dya) * x.M(x.m.dx) * x.M(x.m.sin_sg2));
    }
  }
};
// dp2 = dp1 + div(mfx, mfy) * rarea (:441-445).  in: dp1 mfx mfy ; out: dp2
struct S_trc_dp2 {
  static constexpr int NI = 3, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 5;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {2, 0, 0, 0}, {2, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    x.out(0, x.in(0) + ((x.in(1) - x.in(1, 1, 0)) + (x.in(2) - x.in(2, 0, 1))) * x.M(x.m.rarea));
  }
};
// q = (q*dp1 + div(fx, fy)*rarea) / dp2 (:470-476).  in: q dp1 fx fy dp2 ; out: q_new (halo copied)
struct S_trc_upd {
  static constexpr int NI = 5, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 7;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 1, 0, 0}, {3, 0, 0, 0}, {3, 0, 1, 0}, {4, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    x.out(0, (x.in(0) * x.in(1) + ((x.in(2) - x.in(2, 1, 0)) + (x.in(3) - x.in(3, 0, 1))) * x.M(x.m.rarea)) / x.in(4));
  }
};

// ---------------------------------------------------------------------------------
// q_split = 0: the reference chooses the number of tracer sub-steps from the accumulated Courant numbers at run time
// (fv_tracer2d_nlm.F90:351-420: cmax(k) -> mp_reduce_max -> nsplt = int(1 + max_k cmax), ksplt(k) = int(1 + cmax(k)), fluxes
// scaled by 1/ksplt(k)).  Here the program stays static (q_split_max sub-steps are always issued, so it replays as one CUDA
// graph and needs no host synchronisation): a small device table holds frac(k), ksplt(k) and nsplt, and the stages of a sub-step
// `it` pass a level through unchanged when it > ksplt(k).  nsplt > q_split_max raises a flag that fv3lm_step_* turns into an error.
// The table is trajectory-only (the TL/AD scale the flux increments by the same 1/ksplt(k), fv_tracer2d_tlm.F90:240-242).
// ---------------------------------------------------------------------------------
struct TrcTable {      // device doubles: [0,K) frac ; [K,2K) ksplt ; [2K] nsplt ; [2K+1] overflow ; [2K+2, 3K+2) cmax ; then (nranks-1) K received
  double* p = nullptr; int K = 0;
  ~TrcTable() { dev::free_(p); }
};
struct KTrcCmax {      // per-level maximum of the Courant numbers over the cells this rank owns (:353-371)
  Geom g; Metrics m; const double* cx; const double* cy; double* cmax; int K;
  DEV void operator()(int ii, int jj, int z) const {
    const int il = ii - (g.ng - 1), jl = jj - (g.ng - 1);
    if (il < g.is || il > g.ie || jl < g.js || jl > g.je) return;
    const int k = z % K, tile = z / K;
    const int o = (tile * K + k) * g.slab + jj * g.pitch + ii;
    double v = fmax(fabs(cx[o]), fabs(cy[o]));
    if (!(k + 1 < K / 6)) v += 1.0 - 1.0 / sqrt(m.rsin2[tile * g.slab + jj * g.pitch + ii]);     // + 1 - sin_sg(5)
#ifdef FV3LM_HOST_EMU
    if (v > cmax[k]) cmax[k] = v;
#else
    atomicMax((unsigned long long*)(cmax + k), (unsigned long long)__double_as_longlong(v));      // v >= 0: the bit patterns order like the values
#endif
  }
};
struct KTrcKsplt {     // one thread: global maximum, nsplt, ksplt(k), frac(k) (:374-420)
  double* t; int K, nother, nmax;
  DEV void operator()(int ii, int, int) const {
    if (ii != 0) return;
    double* cmax = t + 2 * K + 2;
    double cg = 0.0;
    for (int k = 0; k < K; k++) {
      double c = cmax[k];
      for (int r = 0; r < nother; r++) c = fmax(c, cmax[(r + 1) * K + k]);
      cmax[k] = c; cg = fmax(cg, c);
    }
    const int nsplt = (int)(1.0 + cg);
    for (int k = 0; k < K; k++) {
      const int ks = nsplt != 1 ? (int)(1.0 + cmax[k]) : 1;
      t[K + k] = (double)ks; t[k] = 1.0 / (double)ks;
    }
    t[2 * K] = (double)nsplt;
    t[2 * K + 1] = nsplt > nmax ? (double)nsplt : 0.0;
  }
};
// a * frac(k) over the whole array.   in: a ; out: a frac
struct S_lev_scale {
  static constexpr int NI = 1, NO = 1;
  struct P { const double* frac; };
  static constexpr int NT = 1;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) { x.out(0, x.in(0) * LDG(p.frac + x.kk)); }
};
// sub-step `it` of the tracer update: levels with it > ksplt(k) pass through.  in: q dp1 fx fy dp2 ; out: q_new
struct S_trc_upd_dyn {
  static constexpr int NI = 5, NO = 1;
  struct P { const double* tab; int K, it; };
  static constexpr int NT = 7;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 1, 0, 0}, {3, 0, 0, 0}, {3, 0, 1, 0}, {4, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    if ((double)p.it > LDG(p.tab + p.K + x.kk)) { x.out(0, x.in(0)); return; }
    x.out(0, (x.in(0) * x.in(1) + ((x.in(2) - x.in(2, 1, 0)) + (x.in(3) - x.in(3, 0, 1))) * x.M(x.m.rarea)) / x.in(4));
  }
};
// dp1 for the next sub-step (:488-494): dp2 where this level took sub-step `it` and it is not the last one.  in: dp1 dp2 ; out: dp1'
struct S_trc_dp1_dyn {
  static constexpr int NI = 2, NO = 1;
  struct P { const double* tab; int K, it; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    const bool take = (double)p.it <= LDG(p.tab + p.K + x.kk) && (double)p.it != LDG(p.tab + 2 * p.K);
    x.out(0, take ? x.in(1) : x.in(0));
  }
};

static std::shared_ptr<TrcTable> add_trc_table(Program& P, int cx, int cy, int nmax) {
  const Geom& g = P.dv->g;
  const int K = g.K;
  Comm* comm = P.dv->comm;
  const int nother = comm && comm->nranks > 1 ? comm->nranks - 1 : 0;
  auto tab = std::make_shared<TrcTable>();
  tab->K = K;
  const size_t n = (size_t)(2 * K + 2) + (size_t)(1 + nother) * K;
  tab->p = (double*)dev::alloc(n * sizeof(double));
  dev::zero(tab->p, n * sizeof(double));
  Op op; op.name = "trc_cmax"; op.in = {cx, cy}; op.nk_launch = 1; op.tl_only = false;
  op.run = [tab, comm, nother, nmax, n](Program& P, Op& o, int mode) {
    if (mode == MODE_AD) return;                 // trajectory-only: the table of the forward sweep is still valid
    const Geom& g = P.dv->g;
    const int K = tab->K;
    double* cmax = tab->p + 2 * K + 2;
    dev::zero(cmax, (size_t)(1 + nother) * K * sizeof(double));
    launch3d(KTrcCmax{g, P.dv->m, P.vals[o.in[0]].traj, P.vals[o.in[1]].traj, cmax, K}, g.NX, g.NY, g.ntile * K);
    if (nother) {                                // mp_reduce_max: every rank sends its K maxima to every other rank
      std::vector<int> peers; std::vector<double*> sb, rb; std::vector<size_t> sc, rc;
      for (int r = 0, k = 0; r < comm->nranks; r++) {
        if (r == comm->rank) continue;
        peers.push_back(r); sb.push_back(cmax); rb.push_back(cmax + (size_t)(1 + k) * K); sc.push_back(K); rc.push_back(K); k++;
      }
      comm->exchange(nother, peers.data(), sb.data(), sc.data(), rb.data(), rc.data());
    }
    launch3d(KTrcKsplt{tab->p, K, nother, nmax}, 1, 1, 1);
    (void)n;
  };
  P.ops.push_back(op);
  P.status_flags.push_back({tab->p + 2 * K + 1, "tracer_2d (q_split = 0) needs more sub-steps than fv3lm_config.q_split_max: nsplt ="});
  P.keep_alive.push_back(tab);
  return tab;
}

std::vector<int> build_tracer_2d(Program& P, Mosaic& mo, std::vector<int> q, int dp1, int mfx, int mfy, int cx, int cy, int hord_tr, const std::string& tag,
                                 int hord_tr_pert, int q_split, int q_split_max) {
  const int K = P.dv->g.K;
  auto nm = [&](const std::string& s) { return tag + "." + s; };
  if (q_split != 0 && q_split != 1) throw std::runtime_error("tracer_2d: a fixed q_split > 1 is not built (1 = one sub-step, 0 = from the Courant numbers)");
  if (q_split == 0) {
    const int nmax = q_split_max > 0 ? q_split_max : 3;
    int xfx0 = P.val(nm("xfx0"), K), yfx0 = P.val(nm("yfx0"), K);
    P.add<S_trc_fx>("trc_fx", {0}, {cx, cy}, {xfx0, yfx0}, K);
    auto tab = add_trc_table(P, cx, cy, nmax);
    auto scaled = [&](int a, const char* n) { int o = P.val(nm(n), K); P.add<S_lev_scale>("trc_scale", {tab->p}, {a}, {o}, K); return o; };
    const int cxs = scaled(cx, "cx_s"), cys = scaled(cy, "cy_s"), xfx = scaled(xfx0, "xfx"), yfx = scaled(yfx0, "yfx"), mfxs = scaled(mfx, "mfx_s"), mfys = scaled(mfy, "mfy_s");
    int ra_x = P.val(nm("ra_x"), K), ra_y = P.val(nm("ra_y"), K);
    P.add<S_ra>("trc_ra", {0}, {xfx, yfx}, {ra_x, ra_y}, K);
    LevOrd ho; for (int k = 0; k < 128; k++) ho.v[k] = (signed char)enc_hord(hord_tr);
    LevOrd hp = ho; if (hord_tr_pert) for (int k = 0; k < 128; k++) hp.v[k] = (signed char)enc_hord(hord_tr_pert);
    const bool split = hord_tr_pert != 0 && hord_tr_pert != hord_tr;
    auto D = [&](int id) { return P.detached(id); };
    for (int it = 1; it <= nmax; it++) {
      const std::string ti = "it" + std::to_string(it) + ".";
      int dp2 = P.val(nm(ti + "dp2"), K);
      P.add<S_trc_dp2>("trc_dp2", {0}, {dp1, mfxs, mfys}, {dp2}, K);
      std::vector<int> out;
      for (size_t n = 0; n < q.size(); n++) {
        const std::string tq = ti + "tp_q" + std::to_string(n);
        if (it > 1) add_patch(P, "halo_q", &mo.h_center, {q[n]});       // (:403-407: the update of the previous sub-step completes here)
        TpOut f;
        if (!split) f = build_fv_tp_2d(P, mo, q[n], cxs, cys, xfx, yfx, ra_x, ra_y, mfxs, mfys, ho, K, nm(tq));
        else {
          P.tl_only = true;
          TpOut a = build_fv_tp_2d(P, mo, q[n], cxs, cys, xfx, yfx, ra_x, ra_y, mfxs, mfys, hp, K, nm(tq + "_p"));
          P.tl_only = false;
          TpOut b = build_fv_tp_2d(P, mo, D(q[n]), D(cxs), D(cys), D(xfx), D(yfx), D(ra_x), D(ra_y), D(mfxs), D(mfys), ho, K, nm(tq + "_t"));
          f.fx = P.val(nm(tq + ".fx"), K); f.fy = P.val(nm(tq + ".fy"), K);
          P.add<S_splice>("splice", {0}, {a.fx, b.fx}, {f.fx}, K);
          P.add<S_splice>("splice", {0}, {a.fy, b.fy}, {f.fy}, K);
        }
        int qn = P.val(nm(ti + "q" + std::to_string(n)), K);
        P.add<S_trc_upd_dyn>("trc_upd", {tab->p, K, it}, {q[n], dp1, f.fx, f.fy, dp2}, {qn}, K);
        out.push_back(qn);
      }
      q = out;
      if (it < nmax) { int d1 = P.val(nm(ti + "dp1"), K); P.add<S_trc_dp1_dyn>("trc_dp1", {tab->p, K, it}, {dp1, dp2}, {d1}, K); dp1 = d1; }
    }
    return q;
  }
  int xfx = P.val(nm("xfx"), K), yfx = P.val(nm("yfx"), K), ra_x = P.val(nm("ra_x"), K), ra_y = P.val(nm("ra_y"), K), dp2 = P.val(nm("dp2"), K);
  P.add<S_trc_fx>("trc_fx", {0}, {cx, cy}, {xfx, yfx}, K);
  P.add<S_ra>("trc_ra", {0}, {xfx, yfx}, {ra_x, ra_y}, K);
  P.add<S_trc_dp2>("trc_dp2", {0}, {dp1, mfx, mfy}, {dp2}, K);
  LevOrd ho; for (int k = 0; k < 128; k++) ho.v[k] = (signed char)enc_hord(hord_tr);
  std::vector<int> out;
  for (size_t n = 0; n < q.size(); n++) {
    TpOut f;
    if (hord_tr_pert == 0 || hord_tr_pert == hord_tr) f = build_fv_tp_2d(P, mo, q[n], cx, cy, xfx, yfx, ra_x, ra_y, mfx, mfy, ho, K, nm("tp_q" + std::to_string(n)));
    else {   // model_tlmadm/fv_tracer2d_tlm.F90:1060-1090: perturbation with hord_tr_pert, trajectory with hord_tr
      LevOrd hp; for (int k = 0; k < 128; k++) hp.v[k] = (signed char)enc_hord(hord_tr_pert);
      auto D = [&](int id) { return P.detached(id); };
      P.tl_only = true;
      TpOut a = build_fv_tp_2d(P, mo, q[n], cx, cy, xfx, yfx, ra_x, ra_y, mfx, mfy, hp, K, nm("tp_q" + std::to_string(n) + "_p"));
      P.tl_only = false;
      TpOut b = build_fv_tp_2d(P, mo, D(q[n]), D(cx), D(cy), D(xfx), D(yfx), D(ra_x), D(ra_y), D(mfx), D(mfy), ho, K, nm("tp_q" + std::to_string(n) + "_t"));
      f.fx = P.val(nm("fx_q" + std::to_string(n)), K); f.fy = P.val(nm("fy_q" + std::to_string(n)), K);
      P.add<S_splice>("splice", {0}, {a.fx, b.fx}, {f.fx}, K);
      P.add<S_splice>("splice", {0}, {a.fy, b.fy}, {f.fy}, K);
    }
    int qn = P.val(nm("q" + std::to_string(n)), K);
    P.add<S_trc_upd>("trc_upd", {0}, {q[n], dp1, f.fx, f.fy, dp2}, {qn}, K);
    out.push_back(qn);
  }
  return out;
}

// ---------------------------------------------------------------------------------
// Lagrangian_to_Eulerian (model/fv_mapz_nlm.F90:60-958), remap_option = 0, |kord| = 17
// ---------------------------------------------------------------------------------
// new Eulerian interface pressures (:313-317, :339-350).  in: pe peln pk ; out: pe2 pn2 pk2   (K+1 levels)
struct S_rm_pe2 {
  static constexpr int NI = 3, NO = 3;
  struct P { LevD ak, bk; double ptop, akap; int K; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, KLAST}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    const int k = x.kk;
    if (k == 0 || k == p.K) {
      x.out(0, k == 0 ? T(p.ptop) : x.in(0, 0, 0, KLAST));
      x.out(1, x.in(1)); x.out(2, x.in(2));
    } else {
      T pe2 = p.ak.v[k] + p.bk.v[k] * x.in(0, 0, 0, KLAST);
      T pn = m_log(pe2);
      x.out(0, pe2); x.out(1, pn); x.out(2, m_exp(p.akap * pn));
    }
  }
};
// theta_v -> T_v, new delp, new pkz (:242-246, :318-328, :468-472)   in: pt pk peln pe2 pk2 pn2 ; out: tv dp2 pkz
struct S_rm_tv {
  static constexpr int NI = 6, NO = 3;
  struct P { double akap; };
  static constexpr int NT = 11;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, 0, 0, 1}, {2, 0, 0, 0}, {2, 0, 0, 1},
                                   {3, 0, 0, 0}, {3, 0, 0, 1}, {4, 0, 0, 0}, {4, 0, 0, 1}, {5, 0, 0, 0}, {5, 0, 0, 1}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    x.out(0, x.in(0) * (x.in(1, 0, 0, 1) - x.in(1)) / (p.akap * (x.in(2, 0, 0, 1) - x.in(2))));
    x.out(1, x.in(3, 0, 0, 1) - x.in(3));
    x.out(2, (x.in(4, 0, 0, 1) - x.in(4)) / (p.akap * (x.in(5, 0, 0, 1) - x.in(5))));
  }
};
// interface pressures at the D-grid wind points (:544-596).  DIR 0: u rows (j-1, j) ; DIR 1: v columns (i-1, i)
// in: pe ; out: pe0 pe3   (K+1 levels)
template <int DIR> struct S_rm_pew {
  static constexpr int NI = 1, NO = 2;
  struct P { LevD ak, bk; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, DIR == 1 ? -1 : 0, DIR == 0 ? -1 : 0, 0}, {0, 0, 0, KLAST}, {0, DIR == 1 ? -1 : 0, DIR == 0 ? -1 : 0, KLAST}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (DIR == 0) { if (!x.in_rect(g.is, g.ie, g.js, g.je + 1)) return; }
    else { if (!x.in_rect(g.is, g.ie + 1, g.js, g.je)) return; }
    constexpr int di = DIR == 1 ? -1 : 0, dj = DIR == 0 ? -1 : 0;
    const int k = x.kk;
    x.out(0, k == 0 ? x.in(0) : 0.5 * (x.in(0, di, dj, 0) + x.in(0)));
    if (DIR == 1 && k == 0) x.out(1, T(p.ak.v[0]));
    else x.out(1, p.ak.v[k] + (0.5 * p.bk.v[k]) * (x.in(0, di, dj, KLAST) + x.in(0, 0, 0, KLAST)));
  }
};
// back to the model's thermodynamic variable (:883-887 last step, :923-931 otherwise).  in: tn q1 pkz ; out: pt
struct S_rm_pt {
  static constexpr int NI = 3, NO = 1;
  struct P { double zvir; int last_step; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    if (p.last_step) x.out(0, x.in(0) / (1.0 + p.zvir * x.in(1)));
    else x.out(0, x.in(0) / x.in(2));
  }
};
// out = in on a rectangle (insert a remapped field into a full-size array value)
struct S_copy {
  static constexpr int NI = 1, NO = 1;
  struct P { int i0, i1, j0, j1; };
  static constexpr int NT = 1;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    x.out(0, x.in(0));
  }
};


// non-hydrostatic thermodynamics of the remap (model/fv_mapz_nlm.F90:247-279):
//   T_v = theta_v * exp(k1k log(rrg delp/delz theta_v)),  delz -> -delz/delp,  dp2 = new delp
// in: pt delp delz pe2 ; out: tv dp2 dzr
struct S_rm_tv_nh {
  static constexpr int NI = 4, NO = 3;
  struct P { double k1k, rrg; };
  static constexpr int NT = 5;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0}, {3, 0, 0, 1}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    T pt = x.in(0), dp = x.in(1), dz = x.in(2);
    x.out(0, pt * m_exp(p.k1k * m_log(p.rrg * dp / dz * pt)));
    x.out(1, x.in(3, 0, 0, 1) - x.in(3));
    x.out(2, -dz / dp);
  }
};
// after the remap (:455-466, :497-505): delz = -dzr*dp2 ; pkz = exp(akap log(rrg dp2/delz T_v))
// in: dzr_n dp2 tn ; out: delz pkz
struct S_rm_post_nh {
  static constexpr int NI = 3, NO = 2;
  struct P { double akap, rrg; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    T dp2 = x.in(1);
    T dz = -x.in(0) * dp2;
    x.out(0, dz);
    x.out(1, m_exp(p.akap * m_log(p.rrg * dp2 / dz * x.in(2))));
  }
};
// non-hydrostatic pkz at the start of the step (model/fv_dynamics_nlm.F90:345-356):
//   pkz = exp(kappa log(rdg delp T (1 + zvir q) / delz)).   in: delp t q1 delz ; out: pkz
struct S_pkz_nh {
  static constexpr int NI = 4, NO = 1;
  struct P { double akap, rdg, zvir; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    x.out(0, m_exp(p.akap * m_log(p.rdg * x.in(0) * x.in(1) * (1.0 + p.zvir * x.in(2)) / x.in(3))));
  }
};

static LevD lev_of(const std::vector<double>& a) {
  LevD o; for (int k = 0; k < 96; k++) o.v[k] = k < (int)a.size() ? a[k] : 0.0; return o;
}

RemapOut build_remap(Program& P, Mosaic& mo, const DynConfig& c, const std::vector<double>& ak, const std::vector<double>& bk,
                     int pe, int pk, int peln, int pt, std::vector<int> q, int u, int v, bool last_step, const std::string& tag,
                     int delp, int w, int delz, int ws) {
  (void)mo;
  const bool nh = !c.hydrostatic;
  const Geom& g = P.dv->g;
  const int K = g.K, is = g.is, ie = g.ie, js = g.js, je = g.je;
  auto nm = [&](const std::string& s) { return tag + "." + s; };
  LevD AK = lev_of(ak), BK = lev_of(bk);
  RemapOut o;
  int pe2 = P.val(nm("pe2"), K + 1), pn2 = P.val(nm("pn2"), K + 1), pk2 = P.val(nm("pk2"), K + 1);
  P.add<S_rm_pe2>("rm_pe2", {AK, BK, c.ptop, c.akap, K}, {pe, peln, pk}, {pe2, pn2, pk2}, K + 1);
  int tv = P.val(nm("tv"), K), dp2 = P.val(nm("dp2"), K), pkz = P.val(nm("pkz"), K);
  int dzr = -1;
  const double rrg = -c.rdgas / c.grav;
  if (nh) { dzr = P.val(nm("dzr"), K); P.add<S_rm_tv_nh>("rm_tv_nh", {c.akap / (1.0 - c.akap), rrg}, {pt, delp, delz, pe2}, {tv, dp2, dzr}, K); }
  else P.add<S_rm_tv>("rm_tv", {c.akap}, {pt, pk, peln, pe2, pk2, pn2}, {tv, dp2, pkz}, K);
  int dummy2d = P.val(nm("qs0"), 1);
  int tn = P.val(nm("tn"), K);
  // one remap call site.  Two-sided mode with a monotone trajectory kord (split_kord, model_tlmadm/fv_mapz_tlm.F90:494-506, 611-655):
  // the increment is remapped with the linear |kord| = 17 scheme about the same inputs, the trajectory with the nonlinear model's kord
  auto remap_site = [&](const char* nm_, int iv, int use_dp2, int i0, int i1, int j0, int j1, int kord_t, bool cs, double qmin,
                        std::vector<int> ins, int out) {
    const int ak = kord_t < 0 ? -kord_t : kord_t;
    if (!c.pert.on || ak > 16) { add_col<S_remap>(P, nm_, {K, iv, use_dp2, i0, i1, j0, j1}, ins, {out}); return; }
    int a = P.val(P.vals[out].name + ".p", K), b = P.val(P.vals[out].name + ".t", K);
    P.tl_only = true;
    add_col<S_remap>(P, nm_, {K, iv, use_dp2, i0, i1, j0, j1}, ins, {a});
    P.tl_only = false;
    std::vector<int> dins; for (int i : ins) dins.push_back(P.detached(i));
    add_col<S_remap_nl>(P, (std::string(nm_) + "_traj").c_str(), {K, iv, use_dp2, i0, i1, j0, j1, ak, cs ? 1 : 0, qmin}, dins, {b});
    P.add<S_splice>("splice", {0}, {a, b}, {out}, K);
  };
  constexpr double t_min = 184.0;    // model/fv_mapz_nlm.F90:40
  remap_site("map_scalar_T", 1, 0, is, ie, js, je, c.kord_tm, false, t_min, {tv, peln, pn2, dummy2d, dp2}, tn);
  o.w = -1; o.delz = -1;
  if (nh) {
    o.w = P.val(nm("w"), K);
    remap_site("map1_ppm_w", -2, 0, is, ie, js, je, c.kord_wz, true, 0.0, {w, pe, pe2, ws, dp2}, o.w);
    int dzn = P.val(nm("dzr_n"), K);
    remap_site("map1_ppm_delz", 1, 0, is, ie, js, je, c.kord_tm, true, 0.0, {dzr, pe, pe2, dummy2d, dp2}, dzn);
    o.delz = P.val(nm("delz"), K);
    P.add<S_rm_post_nh>("rm_post_nh", {c.akap, rrg}, {dzn, dp2, tn}, {o.delz, pkz}, K);
  }
  for (size_t n = 0; n < q.size(); n++) {
    int qn = P.val(nm("q" + std::to_string(n)), K);
    remap_site("map1_q2", 0, 1, is, ie, js, je, c.kord_tr, false, 0.0, {q[n], pe, pe2, dummy2d, dp2}, qn);
    o.q.push_back(qn);
  }
  int pe0u = P.val(nm("pe0u"), K + 1), pe3u = P.val(nm("pe3u"), K + 1), pe0v = P.val(nm("pe0v"), K + 1), pe3v = P.val(nm("pe3v"), K + 1);
  P.add<S_rm_pew<0>>("rm_pe_u", {AK, BK}, {pe}, {pe0u, pe3u}, K + 1);
  P.add<S_rm_pew<1>>("rm_pe_v", {AK, BK}, {pe}, {pe0v, pe3v}, K + 1);
  o.u = P.val(nm("u"), K); o.v = P.val(nm("v"), K);
  remap_site("map1_ppm_u", -1, 0, is, ie, js, je + 1, c.kord_mt, true, 0.0, {u, pe0u, pe3u, dummy2d, dp2}, o.u);
  remap_site("map1_ppm_v", -1, 0, is, ie + 1, js, je, c.kord_mt, true, 0.0, {v, pe0v, pe3v, dummy2d, dp2}, o.v);
  o.pt = P.val(nm("pt"), K);
  P.add<S_rm_pt>("rm_pt", {c.zvir, last_step ? 1 : 0}, {tn, o.q.empty() ? tn : o.q[0], pkz}, {o.pt}, K);
  o.delp = dp2; o.pkz = pkz; o.pe = pe2; o.pk = pk2; o.peln = pn2;
  return o;
}

// ---------------------------------------------------------------------------------
// fv_dynamics
// ---------------------------------------------------------------------------------
// pt = T (1 + zvir q) / pkz  (:396-403).   in: t q1 pkz ; out: theta_v
struct S_thv {
  static constexpr int NI = 3, NO = 1;
  struct P { double zvir; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    x.out(0, x.in(0) * (1.0 + p.zvir * x.in(1)) / x.in(2));
  }
};

FvOut build_fv_dynamics(Program& P, Mosaic& mo, const DynConfig& c, const std::vector<double>& ak, const std::vector<double>& bk, FvState s) {
  const Geom& g = P.dv->g;
  const int K = g.K;
  FvOut o;
  // hydrostatic: pkz from compute_fv3_pressures (fv_pressure.F90:22-69) = geopk restricted to the compute domain
  int pk0 = P.val("fv.pk0", K + 1), gz0 = P.val("fv.gz0", K + 1), pe0 = P.val("fv.pe0", K + 1), pl0 = P.val("fv.peln0", K + 1), pkz0 = P.val("fv.pkz0", K);
  add_col<S_geopk>(P, "compute_fv3_pressures", {c.ptop, c.akap, c.cp_air, 0, 0, K}, {s.delp, s.pt, s.phis}, {pk0, gz0, pe0, pl0, pkz0});
  if (!c.hydrostatic) {
    pkz0 = P.val("fv.pkz_nh", K);
    P.add<S_pkz_nh>("pkz_nh", {c.akap, -c.rdgas / c.grav, c.zvir}, {s.delp, s.pt, s.q.empty() ? s.pt : s.q[0], s.delz}, {pkz0}, K);
  }
  int pt = P.val("fv.thv", K);
  P.add<S_thv>("thv", {c.zvir}, {s.pt, s.q.empty() ? s.pt : s.q[0], pkz0}, {pt}, K);
  int u = s.u, v = s.v, delp = s.delp, w = s.w, delz = s.delz;
  std::vector<int> q = s.q;
  DynConfig cd = c; cd.bdt = c.bdt / c.k_split;
  for (int n_map = 1; n_map <= c.k_split; n_map++) {
    const std::string tg = "m" + std::to_string(n_map);
    add_patch(P, "halo_delp0", &mo.h_center, {delp});
    add_patch(P, "halo_pt0", &mo.h_center, {pt});
    add_patch(P, "halo_uv0", &mo.h_dgrid, {u, v});
    const int dp1 = delp;
    DynState ds; ds.u = u; ds.v = v; ds.w = w; ds.delz = delz; ds.pt = pt; ds.delp = delp; ds.phis = s.phis;
    DynOut d = c.hydrostatic ? build_dyn_core(P, mo, cd, ds, tg) : build_dyn_core_nh(P, mo, cd, ak, bk, ds, tg);
    P.mark_segment();
    for (int& x : q) add_patch(P, "halo_q", &mo.h_center, {x});
    q = build_tracer_2d(P, mo, q, dp1, d.mfx, d.mfy, d.cx, d.cy, c.hord_tr, tg + ".trc", c.pert.on ? c.pert.hord_tr : 0, c.q_split, c.q_split_max);
    P.mark_segment();
    RemapOut r = build_remap(P, mo, c, ak, bk, d.pe, d.pk, d.peln, d.pt, q, d.u, d.v, n_map == c.k_split, tg + ".rm",
                             d.delp, c.hydrostatic ? -1 : d.w, c.hydrostatic ? -1 : d.delz, c.hydrostatic ? -1 : d.ws);
    u = r.u; v = r.v; pt = r.pt; delp = r.delp; q = r.q;
    if (!c.hydrostatic) { w = r.w; delz = r.delz; }
  }
  o.u = u; o.v = v; o.pt = pt; o.delp = delp; o.q = q; o.w = w; o.delz = delz;
  return o;
}

// ---------------------------------------------------------------------------------
// module wrappers
// ---------------------------------------------------------------------------------
void mod_remap(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm); c.hydrostatic = true;
  int pe = io.in(P, "pe", K + 1), pk = io.in(P, "pk", K + 1), peln = io.in(P, "peln", K + 1), pt = io.in(P, "pt", K);
  int q0 = io.in(P, "q0", K), u = io.in(P, "u", K), v = io.in(P, "v", K);
  RemapOut r = build_remap(P, mo, c, *prm.ak, *prm.bk, pe, pk, peln, pt, {q0}, u, v, prm.geti("last_step", 1) != 0, "rm", -1, -1, -1, -1);
  io.out(P, "pt_n", r.pt); io.out(P, "q0_n", r.q[0]); io.out(P, "u_n", r.u); io.out(P, "v_n", r.v); io.out(P, "delp_n", r.delp);
  io.out(P, "pkz_n", r.pkz); io.out(P, "pe_n", r.pe);
}

void mod_c2l_ord4(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams&) {
  const int K = P.dv->g.K;
  int u = io.in(P, "u", K), v = io.in(P, "v", K);
  int a11 = io.in(P, "a11", 1), a12 = io.in(P, "a12", 1), a21 = io.in(P, "a21", 1), a22 = io.in(P, "a22", 1);
  add_patch(P, "halo_dgrid", &mo.h_dgrid, {u, v});            // cubed_to_latlon(..., mode = 1): mpp_update_domains(u, v, DGRID_NE)
  int ua = P.val("ua", K), va = P.val("va", K);
  P.add<S_c2l>("c2l_ord4", {0}, {u, v, a11, a12, a21, a22}, {ua, va}, K);
  io.out(P, "ua", ua); io.out(P, "va", va);
}

void mod_tracer_2d(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  // tracer_2d alone (fv_tracer2d_nlm.F90:275-516) on two tracers: the sub-step logic of q_split = 0 needs Courant numbers a
  // model step at test sizes never reaches
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm);
  int q0 = io.in(P, "q0", K), q1 = io.in(P, "q1", K), dp1 = io.in(P, "dp1", K);
  int mfx = io.in(P, "mfx", K), mfy = io.in(P, "mfy", K), cx = io.in(P, "cx", K), cy = io.in(P, "cy", K);
  std::vector<int> q = {q0, q1};
  for (int& x : q) add_patch(P, "halo_q", &mo.h_center, {x});
  q = build_tracer_2d(P, mo, q, dp1, mfx, mfy, cx, cy, c.hord_tr, "trc", c.pert.on ? c.pert.hord_tr : 0, c.q_split, c.q_split_max);
  io.out(P, "q0_n", q[0]); io.out(P, "q1_n", q[1]);
}

void mod_step(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  // one fv3jedi_lm dynamics step on halo'd arrays whose compute domain holds the API state
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm);
  FvState s;
  s.u = io.in(P, "u", K); s.v = io.in(P, "v", K); s.pt = io.in(P, "t", K); s.delp = io.in(P, "delp", K);
  const char* qn[4] = {"qv", "ql", "qi", "o3"};
  for (int n = 0; n < c.nq; n++) s.q.push_back(io.in(P, qn[n], K));
  s.w = io.in(P, "w", K); s.delz = c.hydrostatic ? -1 : io.in(P, "delz", K); s.phis = io.in(P, "phis", 1);
  // traj_to_fv3 / pert_to_fv3: shared edge rows of the D-grid winds, halo of phis (fv3jedi_lm_dynamics_mod.F90:782-800)
  add_patch(P, "get_boundary_in", &mo.gb_dgrid, {s.u, s.v});
  add_patch(P, "halo_phis", &mo.h_center, {s.phis});
  FvOut o = build_fv_dynamics(P, mo, c, *prm.ak, *prm.bk, s);
  io.out(P, "u_n", o.u); io.out(P, "v_n", o.v); io.out(P, "t_n", o.pt); io.out(P, "delp_n", o.delp);
  for (int n = 0; n < c.nq; n++) io.out(P, std::string(qn[n]) + "_n", o.q[n]);
  if (!c.hydrostatic) { io.out(P, "w_n", o.w); io.out(P, "delz_n", o.delz); }
}

}  // namespace fv3lm
