// Domain decomposition of the six cube tiles into layout(lx, ly) sub-domains and their
// assignment to ranks (one rank = one GPU).  Mirrors the reference's only parallelism,
// the 2-D horizontal decomposition of tools/fv_mp_nlm_mod.F90:411-441 (layout, npes multiple
// of 6) -- except that a rank may own several sub-domains, batched in one launch.
#pragma once
#include <stdexcept>
#include <string>

namespace fv3lm {

struct Decomp {
  int N = 0, lx = 1, ly = 1, nranks = 1, rank = 0;
  int nxl = 0, nyl = 0;      // cells per sub-domain edge
  int nsub_total = 6;        // 6 * lx * ly, global ids d = tile * lx * ly + sy * lx + sx
  int per_rank = 6;

  void init(int N_, int rank_, int nranks_, int lx_, int ly_) {
    N = N_; rank = rank_; nranks = nranks_ < 1 ? 1 : nranks_;
    if (lx_ <= 0 || ly_ <= 0) {
      // smallest layout whose sub-domain count divides evenly over the ranks
      static const int cand[][2] = {{1, 1}, {1, 2}, {2, 2}, {2, 3}, {3, 3}, {2, 4}, {4, 4}};
      bool ok = false;
      for (auto& c : cand) if ((6 * c[0] * c[1]) % nranks == 0 && N % c[0] == 0 && N % c[1] == 0) { lx_ = c[0]; ly_ = c[1]; ok = true; break; }
      if (!ok) throw std::runtime_error("fv3lm: no layout splits the cube evenly over " + std::to_string(nranks) + " ranks");
    }
    lx = lx_; ly = ly_;
    if (N % lx || N % ly) throw std::runtime_error("fv3lm: layout does not divide the tile");
    nxl = N / lx; nyl = N / ly;
    if (nxl < 4 || nyl < 4) throw std::runtime_error("fv3lm: sub-domains must be at least 4 cells wide (halo 3)");
    nsub_total = 6 * lx * ly;
    if (nsub_total % nranks) throw std::runtime_error("fv3lm: 6*layout is not a multiple of the number of ranks");
    per_rank = nsub_total / nranks;
    if (rank < 0 || rank >= nranks) throw std::runtime_error("fv3lm: bad rank");
  }
  int tile_of(int d) const { return d / (lx * ly); }
  int sx_of(int d) const { return d % lx; }
  int sy_of(int d) const { return (d % (lx * ly)) / lx; }
  int i0_of(int d) const { return sx_of(d) * nxl; }
  int j0_of(int d) const { return sy_of(d) * nyl; }
  int owner(int d) const { return d / per_rank; }
  int local_index(int d) const { return d % per_rank; }
  int global_id(int r, int l) const { return r * per_rank + l; }
  // sub-domain of tile t that owns tile-global point index (i, j) of a field whose index range is
  // 1..N (+1 for staggered points: the extra row/column belongs to the last sub-domain)
  int sub_of(int t, int i, int j) const {
    int sx = (i - 1) / nxl; if (sx > lx - 1) sx = lx - 1; if (sx < 0) sx = 0;
    int sy = (j - 1) / nyl; if (sy > ly - 1) sy = ly - 1; if (sy < 0) sy = 0;
    return t * lx * ly + sy * lx + sx;
  }
};

}  // namespace fv3lm
