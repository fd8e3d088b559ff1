#include "mosaic.h"
#include <algorithm>
#include <cmath>
#include <memory>
#include <map>
#include <tuple>

namespace fv3lm {

// ---------------------------------------------------------------------------------
// connectivity.  Contact table tools/fv_mp_nlm_mod.F90:524-573: odd tiles (1,3,5) have
// an aligned east (t+1 west) and south (t-1 north) neighbour and rotated north (t+2
// west) and west (t-2 north) neighbours; even tiles the mirror image.  A point of tile
// A's extended index space is mapped into the neighbour's frame in continuous
// coordinates (cell (1,1) spans [0,1]^2), which handles every staggering at once.
// rot = +1: A's e_x -> -B's e_y, e_y -> +B's e_x ; rot = -1 the opposite.
// ---------------------------------------------------------------------------------
static void to_neighbour(int t, double x, double y, int N, int& tb, double& xb, double& yb, int& rot) {
  bool odd = (t % 2 == 0);  // 0-based 0,2,4 are tiles 1,3,5
  rot = 0;
  if (odd) {
    if (x > N) { tb = (t + 1) % 6; xb = x - N; yb = y; }
    else if (y < 0) { tb = (t + 5) % 6; xb = x; yb = y + N; }
    else if (y > N) { tb = (t + 2) % 6; xb = y - N; yb = N - x; rot = +1; }
    else { tb = (t + 4) % 6; xb = N - y; yb = x + N; rot = -1; }
  } else {
    if (y > N) { tb = (t + 1) % 6; xb = x; yb = y - N; }
    else if (x < 0) { tb = (t + 5) % 6; xb = x + N; yb = y; }
    else if (x > N) { tb = (t + 2) % 6; xb = N - y; yb = x - N; rot = -1; }
    else { tb = (t + 4) % 6; xb = y + N; yb = N - x; rot = +1; }
  }
}

static void stag_off(int st, double& ox, double& oy) {
  switch (st) {
    case ST_CENTER: ox = 0.5; oy = 0.5; break;
    case ST_CORNER: ox = 0.0; oy = 0.0; break;
    case ST_YSTAG: ox = 0.5; oy = 0.0; break;
    default: ox = 0.0; oy = 0.5; break;
  }
}

// halo map for a pair (comp0 staggering, comp1 staggering); scalar fields use npair = 1.
static void build_halo(PatchMap& pm, const Geom& g, int npair, const int* stags, bool vector_sign) {
  const int N = g.N, ng = g.ng, o = ng - 1;
  for (int comp = 0; comp < npair; comp++) {
    double ox, oy; stag_off(stags[comp], ox, oy);
    int nxi = N + (ox == 0.0 ? 1 : 0), nyj = N + (oy == 0.0 ? 1 : 0);
    for (int t = 0; t < 6; t++)
      for (int j = 1 - ng; j <= nyj + ng; j++)
        for (int i = 1 - ng; i <= nxi + ng; i++) {
          double x = i - 1 + ox, y = j - 1 + oy;
          bool outx = (x < 0) || (x > N), outy = (y < 0) || (y > N);
          if (outx == outy) continue;  // interior / shared edge, or corner ghost block
          int tb, rot; double xb, yb;
          to_neighbour(t, x, y, N, tb, xb, yb, rot);
          int sc = comp; double sg = 1.0;
          if (npair == 2 && rot != 0) {
            sc = 1 - comp;
            if (vector_sign) sg = (rot == +1) ? (comp == 0 ? -1.0 : 1.0) : (comp == 0 ? 1.0 : -1.0);
          }
          double sox, soy;
          if (npair == 2) stag_off(stags[sc], sox, soy);
          else { sox = (rot != 0) ? oy : ox; soy = (rot != 0) ? ox : oy; }
          int si = (int)std::lround(xb - sox + 1), sj = (int)std::lround(yb - soy + 1);
          PatchEntry e;
          e.dtile = t; e.dpos = (j + o) * g.pitch + (i + o); e.dcomp = comp;
          e.stile = tb; e.spos = (sj + o) * g.pitch + (si + o); e.scomp = sc; e.sign = sg;
          pm.host.push_back(e);
        }
  }
}

// in-tile corner fills: list of (dst i,j,comp) <- sign * (src i,j,comp), replicated per tile
struct CF { int di, dj, dc, si, sj, sc; double sg; };
static void build_local(PatchMap& pm, const Geom& g, const std::vector<CF>& cf) {
  const int o = g.ng - 1;
  for (int t = 0; t < 6; t++)
    for (const CF& c : cf) {
      PatchEntry e;
      e.dtile = t; e.dpos = (c.dj + o) * g.pitch + (c.di + o); e.dcomp = c.dc;
      e.stile = t; e.spos = (c.sj + o) * g.pitch + (c.si + o); e.scomp = c.sc; e.sign = c.sg;
      pm.host.push_back(e);
    }
  pm.restore = true;
}

static int* up_i(const std::vector<int>& v) {
  int* d = (int*)dev::alloc(std::max<size_t>(1, v.size()) * sizeof(int));
  if (!v.empty()) dev::h2d(d, v.data(), v.size() * sizeof(int));
  return d;
}

void PatchMap::upload() {
  n = (int)host.size();
  std::vector<int> dt(n), dp(n), dc(n), st(n), sp(n), sc(n); std::vector<double> sg(n);
  for (int k = 0; k < n; k++) { dt[k] = host[k].dtile; dp[k] = host[k].dpos; dc[k] = host[k].dcomp; st[k] = host[k].stile; sp[k] = host[k].spos; sc[k] = host[k].scomp; sg[k] = host[k].sign; }
  d_dtile = up_i(dt); d_dpos = up_i(dp); d_dcomp = up_i(dc); d_stile = up_i(st); d_spos = up_i(sp); d_scomp = up_i(sc);
  d_sign = (double*)dev::alloc(std::max(1, n) * sizeof(double));
  if (n) dev::h2d(d_sign, sg.data(), n * sizeof(double));
  // transpose: group forward entries by source cell, forward order inside a row
  std::map<std::tuple<int, int, int>, std::vector<int>> rows;
  for (int k = 0; k < n; k++) rows[std::make_tuple(host[k].scomp, host[k].stile, host[k].spos)].push_back(k);
  nsrc = (int)rows.size();
  std::vector<int> rt, rp, rc, rr, re;
  rr.push_back(0);
  for (auto& kv : rows) {
    rc.push_back(std::get<0>(kv.first)); rt.push_back(std::get<1>(kv.first)); rp.push_back(std::get<2>(kv.first));
    for (int k : kv.second) re.push_back(k);
    rr.push_back((int)re.size());
  }
  a_stile = up_i(rt); a_spos = up_i(rp); a_scomp = up_i(rc); a_row = up_i(rr); a_ent = up_i(re);
  dev::sync();
}
void PatchMap::destroy() {
  for (int* p : {d_dtile, d_dpos, d_dcomp, d_stile, d_spos, d_scomp, a_stile, a_spos, a_scomp, a_row, a_ent}) dev::free_(p);
  dev::free_(d_sign);
  d_dtile = nullptr;
}

void Mosaic::build(const Geom& g_) {
  g = g_;
  const int npx = g.npx, npy = g.npy, ng = g.ng;
  { int s[1] = {ST_CENTER}; build_halo(h_center, g, 1, s, false); h_center.name = "halo_center"; }
  { int s[1] = {ST_CORNER}; build_halo(h_corner, g, 1, s, false); h_corner.name = "halo_corner"; }
  { int s[2] = {ST_YSTAG, ST_XSTAG}; build_halo(h_dgrid, g, 2, s, true); h_dgrid.name = "halo_dgrid"; }
  { int s[2] = {ST_XSTAG, ST_YSTAG}; build_halo(h_cgrid, g, 2, s, true); h_cgrid.name = "halo_cgrid"; }
  // mpp_get_boundary for the D grid (dyn_core_nlm.F90:943-955, fv3jedi_lm_dynamics_mod.F90:386-399): the
  // shared north row of u and east column of v are taken from the tile that owns them
  // as its south / west edge.  fields {u, v}
  {
    const int N = g.N, o = g.ng - 1;
    int stags[2] = {ST_YSTAG, ST_XSTAG};
    for (int t = 0; t < 6; t++)
      for (int e = 0; e < 2; e++)
        for (int n = 1; n <= N; n++) {
          int i = (e == 0) ? n : N + 1, j = (e == 0) ? N + 1 : n;   // e = 0: u(i, N+1) ; e = 1: v(N+1, j)
          double ox, oy; stag_off(stags[e], ox, oy);
          double x = i - 1 + ox, y = j - 1 + oy;
          int tb, rot; double xb, yb;
          to_neighbour(t, e == 0 ? x : x + 0.25, e == 0 ? y + 0.25 : y, N, tb, xb, yb, rot);
          if (e == 0) { if (rot == 0) yb -= 0.25; else xb -= 0.25; } else { if (rot == 0) xb -= 0.25; else yb -= 0.25; }
          int sc = e; double sg = 1.0;
          if (rot != 0) { sc = 1 - e; sg = (rot == +1) ? (e == 0 ? -1.0 : 1.0) : (e == 0 ? 1.0 : -1.0); }
          double sox, soy; stag_off(stags[sc], sox, soy);
          int si = (int)std::lround(xb - sox + 1), sj = (int)std::lround(yb - soy + 1);
          PatchEntry pe;
          pe.dtile = t; pe.dpos = (j + o) * g.pitch + (i + o); pe.dcomp = e;
          pe.stile = tb; pe.spos = (sj + o) * g.pitch + (si + o); pe.scomp = sc; pe.sign = sg;
          gb_dgrid.host.push_back(pe);
        }
    gb_dgrid.name = "get_boundary_dgrid";
  }
  // copy_corners, model/tp_core_nlm.F90:214-289
  for (int dir = 1; dir <= 2; dir++) {
    std::vector<CF> cf;
    for (int j = 1 - ng; j <= 0; j++) for (int i = 1 - ng; i <= 0; i++)
      cf.push_back(dir == 1 ? CF{i, j, 0, j, 1 - i, 0, 1.0} : CF{i, j, 0, 1 - j, i, 0, 1.0});
    for (int j = 1 - ng; j <= 0; j++) for (int i = npx; i <= npx + ng - 1; i++)
      cf.push_back(dir == 1 ? CF{i, j, 0, npy - j, i - npx + 1, 0, 1.0} : CF{i, j, 0, npy + j - 1, npx - i, 0, 1.0});
    for (int j = npy; j <= npy + ng - 1; j++) for (int i = npx; i <= npx + ng - 1; i++)
      cf.push_back(dir == 1 ? CF{i, j, 0, j, 2 * npx - 1 - i, 0, 1.0} : CF{i, j, 0, 2 * npy - 1 - j, i, 0, 1.0});
    for (int j = npy; j <= npy + ng - 1; j++) for (int i = 1 - ng; i <= 0; i++)
      cf.push_back(dir == 1 ? CF{i, j, 0, npy - j, i - 1 + npx, 0, 1.0} : CF{i, j, 0, j + 1 - npx, npy - i, 0, 1.0});
    build_local(dir == 1 ? cc1 : cc2, g, cf);
  }
  cc1.name = "copy_corners_x"; cc2.name = "copy_corners_y";
  // fill_4corners, model/sw_core_nlm.F90:3102-3295
  {
    std::vector<CF> a = {{-1, 0, 0, 0, 2, 0, 1.0}, {0, 0, 0, 0, 1, 0, 1.0}, {npx + 1, 0, 0, npx, 2, 0, 1.0}, {npx, 0, 0, npx, 1, 0, 1.0},
                         {npx, npy, 0, npx, npy - 1, 0, 1.0}, {npx + 1, npy, 0, npx, npy - 2, 0, 1.0}, {0, npy, 0, 0, npy - 1, 0, 1.0}, {-1, npy, 0, 0, npy - 2, 0, 1.0}};
    std::vector<CF> b = {{0, 0, 0, 1, 0, 0, 1.0}, {0, -1, 0, 2, 0, 0, 1.0}, {npx, 0, 0, npx - 1, 0, 0, 1.0}, {npx, -1, 0, npx - 2, 0, 0, 1.0},
                         {npx, npy, 0, npx - 1, npy, 0, 1.0}, {npx, npy + 1, 0, npx - 2, npy, 0, 1.0}, {0, npy, 0, 1, npy, 0, 1.0}, {0, npy + 1, 0, 2, npy, 0, 1.0}};
    build_local(f4c1, g, a); build_local(f4c2, g, b);
    f4c1.name = "fill_4corners_x"; f4c2.name = "fill_4corners_y";
  }
  // fill_corners BGRID, tools/fv_mp_nlm_mod.F90:1046-1083
  {
    std::vector<CF> a, b;
    for (int j = 1; j <= ng; j++) for (int i = 1; i <= ng; i++) {
      a.push_back({1 - i, 1 - j, 0, 1 - j, i + 1, 0, 1.0}); a.push_back({1 - i, npy + j, 0, 1 - j, npy - i, 0, 1.0});
      a.push_back({npx + i, 1 - j, 0, npx + j, i + 1, 0, 1.0}); a.push_back({npx + i, npy + j, 0, npx + j, npy - i, 0, 1.0});
      b.push_back({1 - j, 1 - i, 0, i + 1, 1 - j, 0, 1.0}); b.push_back({1 - j, npy + i, 0, i + 1, npy + j, 0, 1.0});
      b.push_back({npx + j, 1 - i, 0, npx - i, 1 - j, 0, 1.0}); b.push_back({npx + j, npy + i, 0, npx - i, npy + j, 0, 1.0});
    }
    build_local(fcb_x, g, a); build_local(fcb_y, g, b);
    fcb_x.name = "fill_corners_bgrid_x"; fcb_y.name = "fill_corners_bgrid_y";
  }
  // fill_corners(x, y, VECTOR, DGRID), tools/fv_mp_nlm_mod.F90:1271-1303, mySign = -1
  {
    std::vector<CF> a; const double s = -1.0;
    for (int j = 1; j <= ng; j++) for (int i = 1; i <= ng; i++) {
      a.push_back({1 - i, 1 - j, 0, 1 - j, i, 1, s}); a.push_back({1 - i, npy + j, 0, 1 - j, npy - i, 1, 1.0});
      a.push_back({npx - 1 + i, 1 - j, 0, npx + j, i, 1, 1.0}); a.push_back({npx - 1 + i, npy + j, 0, npx + j, npy - i, 1, s});
      a.push_back({1 - i, 1 - j, 1, j, 1 - i, 0, s}); a.push_back({1 - i, npy - 1 + j, 1, j, npy + i, 0, 1.0});
      a.push_back({npx + i, 1 - j, 1, npx - j, 1 - i, 0, 1.0}); a.push_back({npx + i, npy - 1 + j, 1, npx - j, npy + i, 0, s});
    }
    build_local(fc_dgrid_vec, g, a); fc_dgrid_vec.name = "fill_corners_dgrid_vec";
  }
  // d2a2c_vect: corner values of the A-grid winds are taken from the other component of the
  // neighbouring face (dst comp 0 <- src comp 1), model/sw_core_nlm.F90:2884-2925, :2986-3030
  {
    const int je = g.je, ie = g.ie;
    std::vector<CF> ut, ua, vt, va;
    for (int i = -2; i <= 0; i++) { ut.push_back({i, 0, 0, 0, 1 - i, 1, -1.0}); ut.push_back({i, npy, 0, 0, je + i, 1, 1.0}); }
    for (int i = 0; i <= 2; i++) { ut.push_back({npx + i, 0, 0, npx, i + 1, 1, 1.0}); ut.push_back({npx + i, npy, 0, npx, je - i, 1, -1.0}); }
    ua = {{-1, 0, 0, 0, 2, 1, -1.0}, {0, 0, 0, 0, 1, 1, -1.0}, {npx, 0, 0, npx, 1, 1, 1.0}, {npx + 1, 0, 0, npx, 2, 1, 1.0},
          {npx, npy, 0, npx, npy - 1, 1, -1.0}, {npx + 1, npy, 0, npx, npy - 2, 1, -1.0}, {-1, npy, 0, 0, npy - 2, 1, 1.0}, {0, npy, 0, 0, npy - 1, 1, 1.0}};
    for (int j = -2; j <= 0; j++) { vt.push_back({0, j, 0, 1 - j, 0, 1, -1.0}); vt.push_back({npx, j, 0, ie + j, 0, 1, 1.0}); }
    for (int j = 0; j <= 2; j++) { vt.push_back({0, npy + j, 0, j + 1, npy, 1, 1.0}); vt.push_back({npx, npy + j, 0, ie - j, npy, 1, -1.0}); }
    va = {{0, -1, 0, 2, 0, 1, -1.0}, {0, 0, 0, 1, 0, 1, -1.0}, {npx, 0, 0, npx - 1, 0, 1, 1.0}, {npx, -1, 0, npx - 2, 0, 1, 1.0},
          {npx, npy, 0, npx - 1, npy, 1, -1.0}, {npx, npy + 1, 0, npx - 2, npy, 1, -1.0}, {0, npy, 0, 1, npy, 1, 1.0}, {0, npy + 1, 0, 2, npy, 1, 1.0}};
    build_local(c_utmp, g, ut); build_local(c_ua, g, ua); build_local(c_vtmp, g, vt); build_local(c_va, g, va);
    c_utmp.name = "d2a2c_utmp_corners"; c_ua.name = "d2a2c_ua_corners"; c_vtmp.name = "d2a2c_vtmp_corners"; c_va.name = "d2a2c_va_corners";
  }
  for (PatchMap* p : all()) p->upload();
}
std::vector<PatchMap*> Mosaic::all() {
  return {&h_center, &h_corner, &h_dgrid, &h_cgrid, &gb_dgrid, &cc1, &cc2, &f4c1, &f4c2, &fcb_x, &fcb_y, &fc_dgrid_vec, &c_utmp, &c_ua, &c_vtmp, &c_va};
}
void Mosaic::destroy() {
  for (PatchMap* p : all()) p->destroy();
}

// ---------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------
struct PatchArgs {
  const int *dtile, *dpos, *dcomp, *stile, *spos, *scomp; const double* sign;
  const int *a_stile, *a_spos, *a_scomp, *a_row, *a_ent;
  double* F[4]; int nkf[4]; int slab; int n;
};
struct KPatchFwd {
  PatchArgs a; double* save;
  DEV void operator()(int n, int k, int) const {
    int dc = a.dcomp[n], sc = a.scomp[n];
    if (k >= a.nkf[dc] || !a.F[dc]) return;
    size_t d = ((size_t)a.dtile[n] * a.nkf[dc] + k) * a.slab + a.dpos[n];
    size_t s = ((size_t)a.stile[n] * a.nkf[sc] + k) * a.slab + a.spos[n];
    if (save) save[(size_t)k * a.n + n] = a.F[dc][d];
    a.F[dc][d] = a.F[sc] ? a.sign[n] * a.F[sc][s] : 0.0;
  }
};
struct KPatchAdjGather {   // one thread per (source cell, level)
  PatchArgs a;
  DEV void operator()(int r, int k, int) const {
    int sc = a.a_scomp[r];
    if (k >= a.nkf[sc] || !a.F[sc]) return;
    double sum = 0.0;
    for (int e = a.a_row[r]; e < a.a_row[r + 1]; e++) {
      int n = a.a_ent[e]; int dc = a.dcomp[n];
      if (!a.F[dc]) continue;
      sum += a.sign[n] * a.F[dc][((size_t)a.dtile[n] * a.nkf[dc] + k) * a.slab + a.dpos[n]];
    }
    a.F[sc][((size_t)a.a_stile[r] * a.nkf[sc] + k) * a.slab + a.a_spos[r]] += sum;
  }
};
struct KPatchAdjZero {     // zero the adjoint of the overwritten cells, restore trajectory
  PatchArgs a; double* T[4]; const double* save;
  DEV void operator()(int n, int k, int) const {
    int dc = a.dcomp[n];
    if (k >= a.nkf[dc]) return;
    size_t d = ((size_t)a.dtile[n] * a.nkf[dc] + k) * a.slab + a.dpos[n];
    if (a.F[dc]) a.F[dc][d] = 0.0;
    if (save && T[dc]) T[dc][d] = save[(size_t)k * a.n + n];
  }
};

void add_patch(Program& P, const char* nm, PatchMap* map, std::vector<int> fields) {
  Op op; op.name = nm; op.in = fields; op.out = fields; op.inplace = true; op.nk_launch = 1;
  auto scratch = std::make_shared<double*>(nullptr);
  op.run = [map, scratch](Program& P, Op& o, int mode) {
    const Geom& g = P.dv->g;
    PatchArgs a;
    a.dtile = map->d_dtile; a.dpos = map->d_dpos; a.dcomp = map->d_dcomp; a.stile = map->d_stile; a.spos = map->d_spos; a.scomp = map->d_scomp;
    a.sign = map->d_sign; a.a_stile = map->a_stile; a.a_spos = map->a_spos; a.a_scomp = map->a_scomp; a.a_row = map->a_row; a.a_ent = map->a_ent;
    a.slab = g.slab; a.n = map->n;
    int nkmax = 1;
    for (int c = 0; c < 4; c++) { a.F[c] = nullptr; a.nkf[c] = 1; }
    for (size_t c = 0; c < o.in.size(); c++) { a.nkf[c] = P.vals[o.in[c]].nk; nkmax = std::max(nkmax, a.nkf[c]); }
    if (mode == MODE_NL || mode == MODE_TL || mode == MODE_ADFWD) {
      for (size_t c = 0; c < o.in.size(); c++) a.F[c] = P.vals[o.in[c]].traj;
      double* save = nullptr;
      if (mode == MODE_ADFWD && map->restore) { save = P.dv->pool.get((size_t)map->n * nkmax); *scratch = save; }
      launch3d(KPatchFwd{a, save}, map->n, nkmax, 1);
      if (mode == MODE_TL) {
        bool any = false;
        for (size_t c = 0; c < o.in.size(); c++) { Value& v = P.vals[o.in[c]]; a.F[c] = v.active ? v.pert : nullptr; any = any || a.F[c]; }
        // inactive sources contribute zero: only run when every field of the patch is active or none
        if (any) launch3d(KPatchFwd{a, nullptr}, map->n, nkmax, 1);
      }
    } else {  // reverse
      bool any = false;
      for (size_t c = 0; c < o.in.size(); c++) { Value& v = P.vals[o.in[c]]; a.F[c] = v.active ? v.pert : nullptr; any = any || a.F[c]; }
      if (any) launch3d(KPatchAdjGather{a}, map->nsrc, nkmax, 1);
      KPatchAdjZero z; z.a = a; z.save = *scratch;
      for (int c = 0; c < 4; c++) z.T[c] = nullptr;
      for (size_t c = 0; c < o.in.size(); c++) z.T[c] = P.vals[o.in[c]].traj;
      if (any || z.save) launch3d(z, map->n, nkmax, 1);
      if (*scratch) { P.dv->pool.put(*scratch); *scratch = nullptr; }
    }
  };
  P.ops.push_back(op);
}

}  // namespace fv3lm
