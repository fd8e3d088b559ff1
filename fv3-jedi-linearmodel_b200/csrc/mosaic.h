// Cubed-sphere mosaic: halo-exchange and corner ghost-cell fills as precomputed index
// maps ("patch" operators).  Replaces FMS mpp_update_domains / mpp_get_boundary and
// their adjoints (tools/fv_mp_nlm_mod.F90:285-591 contact table, :966-1471 corner
// fills; model_tlmadm/fv_mp_adm.F90:488-725 adjoint halo = accumulate then zero).
//
// forward :  F[dcomp][dst] = sign * F[scomp][src]            (dst cells were dead)
// adjoint :  ad[scomp][src] += sum_dst sign * ad[dcomp][dst]  (CSR gather, fixed order)
//            ad[dcomp][dst]  = 0
#pragma once
#include <vector>
#include "engine.h"
#include "decomp.h"
#include "comm.h"

namespace fv3lm {

// dtile / stile: resident sub-domain index (after Mosaic::build filtered the global list)
struct PatchEntry { int dtile, dpos, dcomp, stile, spos, scomp; double sign; };

// entries of a patch whose source lives on another rank: what this rank sends (its source cells)
// and receives (its destination cells), both in the global entry order so that the two sides agree
struct PeerList {
  int peer = -1;
  int n_send = 0, n_recv = 0, n_srow = 0;
  std::vector<PatchEntry> send, recv;
  int *s_tile = nullptr, *s_pos = nullptr, *s_comp = nullptr;                      // send list (source cells)
  int *r_tile = nullptr, *r_pos = nullptr, *r_comp = nullptr; double* r_sign = nullptr;   // recv list (destination cells)
  int *a_tile = nullptr, *a_pos = nullptr, *a_comp = nullptr, *a_row = nullptr, *a_ent = nullptr;   // CSR: distinct source cells -> send entries
};

struct PatchMap {
  std::string name;
  int n = 0;           // forward entries
  int nsrc = 0;        // distinct sources (adjoint rows)
  bool restore = false;  // corner-type patch: save & restore overwritten trajectory cells
  // device arrays
  int *d_dtile = nullptr, *d_dpos = nullptr, *d_dcomp = nullptr, *d_stile = nullptr, *d_spos = nullptr, *d_scomp = nullptr;
  double* d_sign = nullptr;
  int *a_stile = nullptr, *a_spos = nullptr, *a_scomp = nullptr, *a_row = nullptr;  // CSR rows -> forward entry ids
  int* a_ent = nullptr;
  std::vector<PatchEntry> host;
  std::vector<PeerList> peers;   // remote part (empty on a single rank)
  Comm* comm = nullptr;
  void upload();
  void destroy();
};

// staggering codes (x = i - 1 + ox, y = j - 1 + oy)
enum Stag { ST_CENTER = 0, ST_CORNER = 1, ST_YSTAG = 2 /* D-grid u, C-grid vc */, ST_XSTAG = 3 /* D-grid v, C-grid uc */ };

struct Mosaic {
  Geom g;
  Decomp dc;
  Comm* comm = nullptr;
  // halo exchanges
  PatchMap h_center, h_corner, h_dgrid, h_cgrid;
  PatchMap gb_dgrid;        // mpp_get_boundary(u, v, DGRID_NE): north row of u, east column of v from the owner tile
  // corner fills
  PatchMap cc1, cc2;        // copy_corners(dir)        model/tp_core_nlm.F90:214
  PatchMap f4c1, f4c2;      // fill_4corners(dir)       model/sw_core_nlm.F90:3102
  PatchMap fcb_x, fcb_y;    // fill_corners BGRID X/Y   tools/fv_mp_nlm_mod.F90:1046
  PatchMap fc_dgrid_vec;    // fill_corners(vc,uc,VECTOR,DGRID) :1271
  // d2a2c_vect corner exchanges between the x and y components (model/sw_core_nlm.F90:2884-2925, :2986-3030)
  PatchMap c_utmp, c_ua, c_vtmp, c_va;   // fields {dst, src}
  std::vector<PatchMap*> all();
  void build(const Geom& g, const Decomp& dc, Comm* comm);
  void destroy();
};

void add_patch(Program& P, const char* nm, PatchMap* map, std::vector<int> fields);

}  // namespace fv3lm
