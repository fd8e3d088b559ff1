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

// One entry of the GLOBAL patch list: destination / source sub-domain ids are global, positions
// are relative to the (identically shaped) sub-domain arrays.
struct GEntry { int dsub, dpos, dcomp, ssub, spos, scomp; double sign; };

static int lpos(const Geom& g, const Decomp& dc, int d, int i, int j) {   // tile-global (i, j) -> position in sub-domain d's slab
  const int o = g.ng - 1;
  const int ii = i - dc.i0_of(d) + o, jj = j - dc.j0_of(d) + o;
  if (ii < 0 || ii >= g.NX || jj < 0 || jj >= g.NY) return -1;
  return jj * g.pitch + ii;
}

// halo map for a pair (comp0 staggering, comp1 staggering); scalar fields use npair = 1.
// Every cell of a sub-domain's extended index range that lies outside its own closed region is
// filled from its owner: another sub-domain of the same tile, or -- across a cube edge -- the
// neighbouring tile with the rotation / sign rules of the contact table.  The ng x ng ghost
// blocks at the 8 cube vertices have no owner and are left to the corner fills.
static void build_halo(std::vector<GEntry>& out, const Geom& g, const Decomp& dc, int npair, const int* stags, bool vector_sign) {
  const int N = g.N, ng = g.ng;
  for (int comp = 0; comp < npair; comp++) {
    double ox, oy; stag_off(stags[comp], ox, oy);
    for (int d = 0; d < dc.nsub_total; d++) {
      const int t = dc.tile_of(d), i0 = dc.i0_of(d), j0 = dc.j0_of(d);
      const int nxi = dc.nxl + (ox == 0.0 ? 1 : 0), nyj = dc.nyl + (oy == 0.0 ? 1 : 0);
      for (int jl = 1 - ng; jl <= nyj + ng; jl++)
        for (int il = 1 - ng; il <= nxi + ng; il++) {
          const int i = il + i0, j = jl + j0;
          double x = i - 1 + ox, y = j - 1 + oy;
          if (x >= i0 && x <= i0 + dc.nxl && y >= j0 && y <= j0 + dc.nyl) continue;   // own (closed) region incl. shared edges
          bool outx = (x < 0) || (x > N), outy = (y < 0) || (y > N);
          if (outx && outy) continue;                                                // cube-vertex ghost block
          int tb = t, rot = 0, si = i, sj = j, sc = comp; double sg = 1.0;
          if (outx || outy) {
            double xb, yb;
            to_neighbour(t, x, y, N, tb, xb, yb, rot);
            if (npair == 2 && rot != 0) {
              sc = 1 - comp;
              if (vector_sign) sg = (rot == +1) ? (comp == 0 ? -1.0 : 1.0) : (comp == 0 ? 1.0 : -1.0);
            }
            double sox, soy;
            if (npair == 2) stag_off(stags[sc], sox, soy);
            else { sox = (rot != 0) ? oy : ox; soy = (rot != 0) ? ox : oy; }
            si = (int)std::lround(xb - sox + 1); sj = (int)std::lround(yb - soy + 1);
          }
          const int s = dc.sub_of(tb, si, sj);
          GEntry e;
          e.dsub = d; e.dpos = lpos(g, dc, d, i, j); e.dcomp = comp;
          e.ssub = s; e.spos = lpos(g, dc, s, si, sj); e.scomp = sc; e.sign = sg;
          if (e.dpos < 0 || e.spos < 0) throw std::runtime_error("fv3lm mosaic: halo entry outside the sub-domain arrays");
          out.push_back(e);
        }
    }
  }
}

// in-tile corner fills: list of (dst i,j,comp) <- sign * (src i,j,comp) in tile-global indices, applied
// by the sub-domain that holds the tile corner (the only one whose arrays contain the ghost block)
struct CF { int di, dj, dc, si, sj, sc; double sg; };
static void build_local(std::vector<GEntry>& out, const Geom& g, const Decomp& dc, const std::vector<CF>& cf) {
  for (int d = 0; d < dc.nsub_total; d++)
    for (const CF& c : cf) {
      GEntry e;
      e.dsub = d; e.dpos = lpos(g, dc, d, c.di, c.dj); e.dcomp = c.dc;
      e.ssub = d; e.spos = lpos(g, dc, d, c.si, c.sj); e.scomp = c.sc; e.sign = c.sg;
      if (e.dpos < 0) continue;
      if (e.spos < 0) throw std::runtime_error("fv3lm mosaic: corner-fill source outside the sub-domain arrays");
      out.push_back(e);
    }
}

// split the global list into the rank-local part and the per-peer send / receive lists
static void finish(PatchMap& pm, const std::vector<GEntry>& all, const Decomp& dc, bool restore) {
  pm.restore = restore;
  std::map<int, PeerList> peers;
  for (const GEntry& e : all) {
    const int od = dc.owner(e.dsub), os = dc.owner(e.ssub);
    if (od != dc.rank && os != dc.rank) continue;
    PatchEntry pe;
    pe.dtile = dc.local_index(e.dsub); pe.dpos = e.dpos; pe.dcomp = e.dcomp;
    pe.stile = dc.local_index(e.ssub); pe.spos = e.spos; pe.scomp = e.scomp; pe.sign = e.sign;
    if (od == dc.rank && os == dc.rank) pm.host.push_back(pe);
    else if (od == dc.rank) { peers[os].peer = os; peers[os].recv.push_back(pe); }
    else { peers[od].peer = od; peers[od].send.push_back(pe); }
  }
  for (auto& kv : peers) pm.peers.push_back(kv.second);
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
  for (PeerList& pl : peers) {
    pl.n_send = (int)pl.send.size(); pl.n_recv = (int)pl.recv.size();
    std::vector<int> a(pl.n_send), b(pl.n_send), c(pl.n_send);
    for (int k = 0; k < pl.n_send; k++) { a[k] = pl.send[k].stile; b[k] = pl.send[k].spos; c[k] = pl.send[k].scomp; }
    pl.s_tile = up_i(a); pl.s_pos = up_i(b); pl.s_comp = up_i(c);
    std::vector<int> d(pl.n_recv), e(pl.n_recv), f(pl.n_recv); std::vector<double> sgn(pl.n_recv);
    for (int k = 0; k < pl.n_recv; k++) { d[k] = pl.recv[k].dtile; e[k] = pl.recv[k].dpos; f[k] = pl.recv[k].dcomp; sgn[k] = pl.recv[k].sign; }
    pl.r_tile = up_i(d); pl.r_pos = up_i(e); pl.r_comp = up_i(f);
    pl.r_sign = (double*)dev::alloc(std::max(1, pl.n_recv) * sizeof(double));
    if (pl.n_recv) dev::h2d(pl.r_sign, sgn.data(), pl.n_recv * sizeof(double));
    // adjoint: distinct source cells of the send list, each summing its entries in list order
    std::map<std::tuple<int, int, int>, std::vector<int>> srows;
    for (int k = 0; k < pl.n_send; k++) srows[std::make_tuple(pl.send[k].scomp, pl.send[k].stile, pl.send[k].spos)].push_back(k);
    pl.n_srow = (int)srows.size();
    std::vector<int> qt, qp, qc, qr, qe; qr.push_back(0);
    for (auto& kv : srows) {
      qc.push_back(std::get<0>(kv.first)); qt.push_back(std::get<1>(kv.first)); qp.push_back(std::get<2>(kv.first));
      for (int k : kv.second) qe.push_back(k);
      qr.push_back((int)qe.size());
    }
    pl.a_tile = up_i(qt); pl.a_pos = up_i(qp); pl.a_comp = up_i(qc); pl.a_row = up_i(qr); pl.a_ent = up_i(qe);
  }
  dev::sync();
}
void PatchMap::destroy() {
  for (int* p : {d_dtile, d_dpos, d_dcomp, d_stile, d_spos, d_scomp, a_stile, a_spos, a_scomp, a_row, a_ent}) dev::free_(p);
  dev::free_(d_sign);
  d_dtile = nullptr;
  for (PeerList& pl : peers) {
    for (int* p : {pl.s_tile, pl.s_pos, pl.s_comp, pl.r_tile, pl.r_pos, pl.r_comp, pl.a_tile, pl.a_pos, pl.a_comp, pl.a_row, pl.a_ent}) dev::free_(p);
    dev::free_(pl.r_sign);
  }
  peers.clear();
}

void Mosaic::build(const Geom& g_, const Decomp& dc_, Comm* comm_) {
  g = g_; dc = dc_; comm = comm_;
  const int npx = g.npx, npy = g.npy, ng = g.ng;
  { int s[1] = {ST_CENTER}; { std::vector<GEntry> all; build_halo(all, g, dc, 1, s, false); finish(h_center, all, dc, false); } h_center.name = "halo_center"; }
  { int s[1] = {ST_CORNER}; { std::vector<GEntry> all; build_halo(all, g, dc, 1, s, false); finish(h_corner, all, dc, false); } h_corner.name = "halo_corner"; }
  { int s[2] = {ST_YSTAG, ST_XSTAG}; { std::vector<GEntry> all; build_halo(all, g, dc, 2, s, true); finish(h_dgrid, all, dc, false); } h_dgrid.name = "halo_dgrid"; }
  { int s[2] = {ST_XSTAG, ST_YSTAG}; { std::vector<GEntry> all; build_halo(all, g, dc, 2, s, true); finish(h_cgrid, all, dc, false); } h_cgrid.name = "halo_cgrid"; }
  // mpp_get_boundary for the D grid (dyn_core_nlm.F90:943-955, fv3jedi_lm_dynamics_mod.F90:386-399): the
  // shared north row of u and east column of v are taken from the tile that owns them
  // as its south / west edge.  fields {u, v}
  {
    const int N = g.N;
    int stags[2] = {ST_YSTAG, ST_XSTAG};
    std::vector<GEntry> all;
    for (int d = 0; d < dc.nsub_total; d++)
      for (int e = 0; e < 2; e++) {
        const int t = dc.tile_of(d), i0 = dc.i0_of(d), j0 = dc.j0_of(d);
        const int nn = (e == 0) ? dc.nxl : dc.nyl;
        for (int n = 1; n <= nn; n++) {
          int i = (e == 0) ? i0 + n : i0 + dc.nxl + 1, j = (e == 0) ? j0 + dc.nyl + 1 : j0 + n;   // e = 0: u(i, je+1) ; e = 1: v(ie+1, j)
          int tb = t, si = i, sj = j, sc = e; double sg = 1.0;
          if ((e == 0 && j > N) || (e == 1 && i > N)) {
            double ox, oy; stag_off(stags[e], ox, oy);
            double x = i - 1 + ox, y = j - 1 + oy;
            int rot; double xb, yb;
            to_neighbour(t, e == 0 ? x : x + 0.25, e == 0 ? y + 0.25 : y, N, tb, xb, yb, rot);
            if (e == 0) { if (rot == 0) yb -= 0.25; else xb -= 0.25; } else { if (rot == 0) xb -= 0.25; else yb -= 0.25; }
            if (rot != 0) { sc = 1 - e; sg = (rot == +1) ? (e == 0 ? -1.0 : 1.0) : (e == 0 ? 1.0 : -1.0); }
            double sox, soy; stag_off(stags[sc], sox, soy);
            si = (int)std::lround(xb - sox + 1); sj = (int)std::lround(yb - soy + 1);
          }
          const int s2 = dc.sub_of(tb, si, sj);
          GEntry ge;
          ge.dsub = d; ge.dpos = lpos(g, dc, d, i, j); ge.dcomp = e;
          ge.ssub = s2; ge.spos = lpos(g, dc, s2, si, sj); ge.scomp = sc; ge.sign = sg;
          if (ge.dpos < 0 || ge.spos < 0) throw std::runtime_error("fv3lm mosaic: get_boundary entry outside the sub-domain arrays");
          all.push_back(ge);
        }
      }
    finish(gb_dgrid, all, dc, false);
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
    { std::vector<GEntry> all; build_local(all, g, dc, cf); finish(dir == 1 ? cc1 : cc2, all, dc, true); }
  }
  cc1.name = "copy_corners_x"; cc2.name = "copy_corners_y";
  // fill_4corners, model/sw_core_nlm.F90:3102-3295
  {
    std::vector<CF> a = {{-1, 0, 0, 0, 2, 0, 1.0}, {0, 0, 0, 0, 1, 0, 1.0}, {npx + 1, 0, 0, npx, 2, 0, 1.0}, {npx, 0, 0, npx, 1, 0, 1.0},
                         {npx, npy, 0, npx, npy - 1, 0, 1.0}, {npx + 1, npy, 0, npx, npy - 2, 0, 1.0}, {0, npy, 0, 0, npy - 1, 0, 1.0}, {-1, npy, 0, 0, npy - 2, 0, 1.0}};
    std::vector<CF> b = {{0, 0, 0, 1, 0, 0, 1.0}, {0, -1, 0, 2, 0, 0, 1.0}, {npx, 0, 0, npx - 1, 0, 0, 1.0}, {npx, -1, 0, npx - 2, 0, 0, 1.0},
                         {npx, npy, 0, npx - 1, npy, 0, 1.0}, {npx, npy + 1, 0, npx - 2, npy, 0, 1.0}, {0, npy, 0, 1, npy, 0, 1.0}, {0, npy + 1, 0, 2, npy, 0, 1.0}};
    { std::vector<GEntry> all; build_local(all, g, dc, a); finish(f4c1, all, dc, true); } { std::vector<GEntry> all; build_local(all, g, dc, b); finish(f4c2, all, dc, true); }
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
    { std::vector<GEntry> all; build_local(all, g, dc, a); finish(fcb_x, all, dc, true); } { std::vector<GEntry> all; build_local(all, g, dc, b); finish(fcb_y, all, dc, true); }
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
    { std::vector<GEntry> all; build_local(all, g, dc, a); finish(fc_dgrid_vec, all, dc, true); } fc_dgrid_vec.name = "fill_corners_dgrid_vec";
  }
  // d2a2c_vect: corner values of the A-grid winds are taken from the other component of the
  // neighbouring face (dst comp 0 <- src comp 1), model/sw_core_nlm.F90:2884-2925, :2986-3030
  {
    const int je = g.N, ie = g.N;   // tile-global
    std::vector<CF> ut, ua, vt, va;
    for (int i = -2; i <= 0; i++) { ut.push_back({i, 0, 0, 0, 1 - i, 1, -1.0}); ut.push_back({i, npy, 0, 0, je + i, 1, 1.0}); }
    for (int i = 0; i <= 2; i++) { ut.push_back({npx + i, 0, 0, npx, i + 1, 1, 1.0}); ut.push_back({npx + i, npy, 0, npx, je - i, 1, -1.0}); }
    ua = {{-1, 0, 0, 0, 2, 1, -1.0}, {0, 0, 0, 0, 1, 1, -1.0}, {npx, 0, 0, npx, 1, 1, 1.0}, {npx + 1, 0, 0, npx, 2, 1, 1.0},
          {npx, npy, 0, npx, npy - 1, 1, -1.0}, {npx + 1, npy, 0, npx, npy - 2, 1, -1.0}, {-1, npy, 0, 0, npy - 2, 1, 1.0}, {0, npy, 0, 0, npy - 1, 1, 1.0}};
    for (int j = -2; j <= 0; j++) { vt.push_back({0, j, 0, 1 - j, 0, 1, -1.0}); vt.push_back({npx, j, 0, ie + j, 0, 1, 1.0}); }
    for (int j = 0; j <= 2; j++) { vt.push_back({0, npy + j, 0, j + 1, npy, 1, 1.0}); vt.push_back({npx, npy + j, 0, ie - j, npy, 1, -1.0}); }
    va = {{0, -1, 0, 2, 0, 1, -1.0}, {0, 0, 0, 1, 0, 1, -1.0}, {npx, 0, 0, npx - 1, 0, 1, 1.0}, {npx, -1, 0, npx - 2, 0, 1, 1.0},
          {npx, npy, 0, npx - 1, npy, 1, -1.0}, {npx, npy + 1, 0, npx - 2, npy, 1, -1.0}, {0, npy, 0, 1, npy, 1, 1.0}, {0, npy + 1, 0, 2, npy, 1, 1.0}};
    { std::vector<GEntry> all; build_local(all, g, dc, ut); finish(c_utmp, all, dc, true); } { std::vector<GEntry> all; build_local(all, g, dc, ua); finish(c_ua, all, dc, true); } { std::vector<GEntry> all; build_local(all, g, dc, vt); finish(c_vtmp, all, dc, true); } { std::vector<GEntry> all; build_local(all, g, dc, va); finish(c_va, all, dc, true); }
    c_utmp.name = "d2a2c_utmp_corners"; c_ua.name = "d2a2c_ua_corners"; c_vtmp.name = "d2a2c_vtmp_corners"; c_va.name = "d2a2c_va_corners";
  }
  for (PatchMap* p : all()) { p->comm = comm; p->upload(); }
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

// ---- remote part: pack / unpack kernels (one thread per (entry, level)); message layout
// [set][level][entry], set 0 = trajectory (or adjoint), set 1 = perturbation in TL mode
struct KPack {       // send list: source cells -> buffer
  const int *tile, *pos, *comp; double* F[4]; int nkf[4]; int slab, n, nk; double* buf;
  DEV void operator()(int e, int k, int) const {
    int c = comp[e];
    double v = 0.0;
    if (F[c] && k < nkf[c]) v = F[c][((size_t)tile[e] * nkf[c] + k) * slab + pos[e]];
    buf[(size_t)k * n + e] = v;
  }
};
struct KUnpack {     // buffer -> destination cells (sign applied here)
  const int *tile, *pos, *comp; const double* sign; double* F[4]; int nkf[4]; int slab, n, nk; const double* buf;
  DEV void operator()(int e, int k, int) const {
    int c = comp[e];
    if (!F[c] || k >= nkf[c]) return;
    F[c][((size_t)tile[e] * nkf[c] + k) * slab + pos[e]] = sign[e] * buf[(size_t)k * n + e];
  }
};
struct KPackAdj {    // recv list: sign * adjoint of the destination cells -> buffer, then zero them
  const int *tile, *pos, *comp; const double* sign; double* F[4]; int nkf[4]; int slab, n, nk; double* buf;
  DEV void operator()(int e, int k, int) const {
    int c = comp[e];
    double v = 0.0;
    if (F[c] && k < nkf[c]) {
      size_t o = ((size_t)tile[e] * nkf[c] + k) * slab + pos[e];
      v = sign[e] * F[c][o];
      F[c][o] = 0.0;
    }
    buf[(size_t)k * n + e] = v;
  }
};
struct KUnpackAdjAdd {   // one thread per (distinct source cell, level): fixed-order sum of its entries
  const int *tile, *pos, *comp, *row, *ent; double* F[4]; int nkf[4]; int slab, n, nk; const double* buf;
  DEV void operator()(int r, int k, int) const {
    int c = comp[r];
    if (!F[c] || k >= nkf[c]) return;
    double sum = 0.0;
    for (int q = row[r]; q < row[r + 1]; q++) sum += buf[(size_t)k * n + ent[q]];
    F[c][((size_t)tile[r] * nkf[c] + k) * slab + pos[r]] += sum;
  }
};

// forward exchange of `nset` field sets (F0 = trajectory, F1 = perturbation or null)
static void remote_forward(Program& P, PatchMap* map, Comm* comm, double* const F0[4], double* const F1[4], const int nkf[4], int nkmax) {
  const Geom& g = P.dv->g;
  const int np = (int)map->peers.size();
  if (np == 0) return;
  if (!comm) throw std::runtime_error("fv3lm: patch with remote entries but no communicator");
  const int nset = F1 ? 2 : 1;
  std::vector<int> peers(np); std::vector<double*> sb(np), rb(np); std::vector<size_t> sc(np), rc(np);
  for (int p = 0; p < np; p++) {
    PeerList& pl = map->peers[p];
    peers[p] = pl.peer;
    sc[p] = (size_t)nset * nkmax * pl.n_send; rc[p] = (size_t)nset * nkmax * pl.n_recv;
    sb[p] = sc[p] ? P.dv->pool.get(sc[p]) : nullptr; rb[p] = rc[p] ? P.dv->pool.get(rc[p]) : nullptr;
    for (int s = 0; s < nset && pl.n_send; s++) {
      KPack k; k.tile = pl.s_tile; k.pos = pl.s_pos; k.comp = pl.s_comp; k.slab = g.slab; k.n = pl.n_send; k.nk = nkmax;
      for (int c = 0; c < 4; c++) { k.F[c] = s == 0 ? F0[c] : F1[c]; k.nkf[c] = nkf[c]; }
      k.buf = sb[p] + (size_t)s * nkmax * pl.n_send;
      launch3d(k, pl.n_send, nkmax, 1);
    }
  }
  comm->exchange(np, peers.data(), sb.data(), sc.data(), rb.data(), rc.data());
  for (int p = 0; p < np; p++) {
    PeerList& pl = map->peers[p];
    for (int s = 0; s < nset && pl.n_recv; s++) {
      KUnpack k; k.tile = pl.r_tile; k.pos = pl.r_pos; k.comp = pl.r_comp; k.sign = pl.r_sign; k.slab = g.slab; k.n = pl.n_recv; k.nk = nkmax;
      for (int c = 0; c < 4; c++) { k.F[c] = s == 0 ? F0[c] : F1[c]; k.nkf[c] = nkf[c]; }
      k.buf = rb[p] + (size_t)s * nkmax * pl.n_recv;
      launch3d(k, pl.n_recv, nkmax, 1);
    }
    if (sb[p]) P.dv->pool.put(sb[p]);
    if (rb[p]) P.dv->pool.put(rb[p]);
  }
}

// adjoint exchange: halo -> owner accumulate, then the halo adjoint is zero (mpp_update_domains_ad + zero_domain)
static void remote_adjoint(Program& P, PatchMap* map, Comm* comm, double* const Fad[4], const int nkf[4], int nkmax) {
  const Geom& g = P.dv->g;
  const int np = (int)map->peers.size();
  if (np == 0) return;
  if (!comm) throw std::runtime_error("fv3lm: patch with remote entries but no communicator");
  std::vector<int> peers(np); std::vector<double*> sb(np), rb(np); std::vector<size_t> sc(np), rc(np);
  for (int p = 0; p < np; p++) {
    PeerList& pl = map->peers[p];
    peers[p] = pl.peer;
    // roles are reversed: the destination side sends, the source side receives
    sc[p] = (size_t)nkmax * pl.n_recv; rc[p] = (size_t)nkmax * pl.n_send;
    sb[p] = sc[p] ? P.dv->pool.get(sc[p]) : nullptr; rb[p] = rc[p] ? P.dv->pool.get(rc[p]) : nullptr;
    if (pl.n_recv) {
      KPackAdj k; k.tile = pl.r_tile; k.pos = pl.r_pos; k.comp = pl.r_comp; k.sign = pl.r_sign; k.slab = g.slab; k.n = pl.n_recv; k.nk = nkmax;
      for (int c = 0; c < 4; c++) { k.F[c] = Fad[c]; k.nkf[c] = nkf[c]; }
      k.buf = sb[p];
      launch3d(k, pl.n_recv, nkmax, 1);
    }
  }
  comm->exchange(np, peers.data(), sb.data(), sc.data(), rb.data(), rc.data());
  for (int p = 0; p < np; p++) {
    PeerList& pl = map->peers[p];
    if (pl.n_send) {
      KUnpackAdjAdd k; k.tile = pl.a_tile; k.pos = pl.a_pos; k.comp = pl.a_comp; k.row = pl.a_row; k.ent = pl.a_ent; k.slab = g.slab; k.n = pl.n_send; k.nk = nkmax;
      for (int c = 0; c < 4; c++) { k.F[c] = Fad[c]; k.nkf[c] = nkf[c]; }
      k.buf = rb[p];
      launch3d(k, pl.n_srow, nkmax, 1);
    }
    if (sb[p]) P.dv->pool.put(sb[p]);
    if (rb[p]) P.dv->pool.put(rb[p]);
  }
}

void add_patch(Program& P, const char* nm, PatchMap* map, std::vector<int> fields) {
  Comm* comm = map->comm;
  Op op; op.name = nm; op.in = fields; op.out = fields; op.inplace = true; op.nk_launch = 1; op.tl_only = P.tl_only;
  auto scratch = std::make_shared<double*>(nullptr);
  op.run = [map, scratch, comm](Program& P, Op& o, int mode) {
    const Geom& g = P.dv->g;
    PatchArgs a;
    a.dtile = map->d_dtile; a.dpos = map->d_dpos; a.dcomp = map->d_dcomp; a.stile = map->d_stile; a.spos = map->d_spos; a.scomp = map->d_scomp;
    a.sign = map->d_sign; a.a_stile = map->a_stile; a.a_spos = map->a_spos; a.a_scomp = map->a_scomp; a.a_row = map->a_row; a.a_ent = map->a_ent;
    a.slab = g.slab; a.n = map->n;
    int nkmax = 1;
    for (int c = 0; c < 4; c++) { a.F[c] = nullptr; a.nkf[c] = 1; }
    for (size_t c = 0; c < o.in.size(); c++) { a.nkf[c] = P.vals[o.in[c]].nk; nkmax = std::max(nkmax, a.nkf[c]); }
    double* Ft[4] = {nullptr, nullptr, nullptr, nullptr}; double* Fp[4] = {nullptr, nullptr, nullptr, nullptr};
    bool any = false;
    for (size_t c = 0; c < o.in.size(); c++) { Value& v = P.vals[o.in[c]]; Ft[c] = v.traj; Fp[c] = v.active ? v.pert : nullptr; any = any || Fp[c]; }
    if (mode == MODE_NL || mode == MODE_TL || mode == MODE_ADFWD) {
      for (int c = 0; c < 4; c++) a.F[c] = Ft[c];
      double* save = nullptr;
      if (mode == MODE_ADFWD && map->restore && map->n) { save = P.dv->pool.get((size_t)map->n * nkmax); *scratch = save; }
      if (map->n) launch3d(KPatchFwd{a, save}, map->n, nkmax, 1);
      if (mode == MODE_TL && any) {
        // inactive sources contribute zero: only run when every field of the patch is active or none
        for (int c = 0; c < 4; c++) a.F[c] = Fp[c];
        if (map->n) launch3d(KPatchFwd{a, nullptr}, map->n, nkmax, 1);
      }
      remote_forward(P, map, comm, Ft, (mode == MODE_TL && any) ? Fp : nullptr, a.nkf, nkmax);
    } else {  // reverse
      for (int c = 0; c < 4; c++) a.F[c] = Fp[c];
      if (any && map->n) launch3d(KPatchAdjGather{a}, map->nsrc, nkmax, 1);
      KPatchAdjZero z; z.a = a; z.save = *scratch;
      for (int c = 0; c < 4; c++) z.T[c] = Ft[c];
      if ((any || z.save) && map->n) launch3d(z, map->n, nkmax, 1);
      if (any) remote_adjoint(P, map, comm, Fp, a.nkf, nkmax);
      if (*scratch) { P.dv->pool.put(*scratch); *scratch = nullptr; }
    }
  };
  P.ops.push_back(op);
}

}  // namespace fv3lm
