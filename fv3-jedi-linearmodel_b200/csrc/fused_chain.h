// Automatic tile fusion of a patch-free CHAIN of stencil stages for forward sweeps (NL, TL).
//
//   add_chain<S1, S2, ...>(P, name, {p1, p2, ...}, {ins of S1, ins of S2, ...}, {outs of S1, ...}, live_out, nk)
//
// builds ONE op (VAR_FWD) that evaluates the stages' own eval() one after the other inside a block: a value produced by one stage
// and read by a later one lives in a shared-memory tile whose box is derived from the consumers' taps (reverse data-flow over the
// chain: box(v) = union over consumers c of region(c) (+) taps_c(v), plus the block's own cells if v leaves the chain); values that
// enter the chain are read from global memory, values in `live_out` are also written to it.  One phase per stage, a block barrier
// between phases.  The stage-by-stage ops of the same chain stay in the program as VAR_AD ops (the adjoint needs the
// intermediates in HBM), so the caller brackets them with P.variant = VAR_AD and then calls add_chain.
// Restrictions: taps of chain-internal values must have dk = 0; every value has the launch's number of levels or is 2-D input.
#pragma once
#include "fused_tp.h"

namespace fv3lm {
namespace ftp {

constexpr int CH_MAXV = 24, CH_MAXI = 12, CH_MAXO = 4;
struct ChainVal {
  Fld gin; OFld gout;          // global storage: gin.v for values entering the chain, gout.v for values leaving it
  int soff;                    // tile offset in shared memory (doubles), -1 = no tile
  int x0, y0, w, h;            // tile box relative to the block origin
};
template <int NS> struct ChainParams {
  ChainVal v[CH_MAXV];
  int in_slot[NS][CH_MAXI], out_slot[NS][CH_MAXO];
  int rx0[NS], ry0[NS], rw[NS], rh[NS];      // region each stage is evaluated on, relative to the block origin
  int tile_doubles;                          // shared memory per number kind (value; TL: + as much for the perturbation)
};

template <int n, class... S> struct NthStage;
template <class S0, class... S> struct NthStage<0, S0, S...> { using type = S0; };
template <int n, class S0, class... S> struct NthStage<n, S0, S...> { using type = typename NthStage<n - 1, S...>::type; };

// the stages' parameter structs, one after the other (a plain aggregate: kernel parameters must be trivially copyable)
template <class... S> struct PPack;
template <> struct PPack<> {};
template <class S0, class... S> struct PPack<S0, S...> { typename S0::P head; PPack<S...> tail; };
template <int n, class S0, class... S> HD const auto& pget(const PPack<S0, S...>& p) {
  if constexpr (n == 0) return p.head; else return pget<n - 1, S...>(p.tail);
}
inline PPack<> ppack() { return {}; }
template <class S0, class... S, class P0, class... Ps> PPack<S0, S...> ppack_of(const P0& p0, const Ps&... ps) {
  PPack<S0, S...> r; r.head = p0;
  if constexpr (sizeof...(S) > 0) r.tail = ppack_of<S...>(ps...);
  return r;
}

template <class TT, class... S> struct KernChain {
  static constexpr int NS = sizeof...(S);
  static constexpr int NPH = NS;
  static constexpr bool TLM = !std::is_same<TT, double>::value;
  Geom g; Metrics m; int nk;
  ChainParams<NS> cp;
  PPack<S...> prm;
  struct Smem { double* base; };            // dynamic shared memory (device) / heap (emulation)
  template <int s> DEV void stage(int tid, int ii0, int jj0, int kk, int tile, double* sm) const {
    using St = typename NthStage<s, S...>::type;
    TCtx<TT, St::NI, St::NO> x; x.g = g; x.m = m;
#pragma unroll
    for (int f = 0; f < St::NI; f++) {
      const ChainVal& v = cp.v[cp.in_slot[s][f]];
      x.gi[f] = v.gin;
      if (v.soff >= 0) x.ti[f] = TRef{sm + v.soff, sm + cp.tile_doubles + v.soff, Box{ii0 + v.x0, jj0 + v.y0, v.w, v.h}};
      else x.ti[f] = TRef{nullptr, nullptr, Box{0, 0, 0, 0}};
    }
#pragma unroll
    for (int o = 0; o < St::NO; o++) {
      const ChainVal& v = cp.v[cp.out_slot[s][o]];
      x.go[o] = v.gout;
      if (v.soff >= 0) x.to[o] = TOut{sm + v.soff, sm + cp.tile_doubles + v.soff, Box{ii0 + v.x0, jj0 + v.y0, v.w, v.h}};
      else x.to[o] = TOut{nullptr, nullptr, Box{0, 0, 0, 0}};
    }
    const int n = cp.rw[s] * cp.rh[s], i0 = g.i0[tile], j0 = g.j0[tile];
    for (int c = tid; c < n; c += NTHR) {
      const int ii = ii0 + cp.rx0[s] + c % cp.rw[s], jj = jj0 + cp.ry0[s] + c / cp.rw[s];
      if (ii < 0 || ii >= g.NX || jj < 0 || jj >= g.NY) continue;
      x.setpos(ii, jj, kk, tile, i0, j0);
      // a cell of the region that lies outside the block's own cells must not reach global memory (its owner writes it)
      const bool own = ii >= ii0 && ii < ii0 + TX && jj >= jj0 && jj < jj0 + TY;
      if (!own) {
#pragma unroll
        for (int o = 0; o < St::NO; o++) x.go[o].v = nullptr;
      }
      St::eval(x, pget<s, S...>(prm));
      if (!own) {
#pragma unroll
        for (int o = 0; o < St::NO; o++) x.go[o] = cp.v[cp.out_slot[s][o]].gout;
      }
    }
  }
  template <int s> DEV void dispatch(int ph, int tid, int ii0, int jj0, int kk, int tile, double* sm) const {
    if constexpr (s < NS) {
      if (ph == s) stage<s>(tid, ii0, jj0, kk, tile, sm);
      else dispatch<s + 1>(ph, tid, ii0, jj0, kk, tile, sm);
    }
  }
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& s) const {
    int tile, kk; split_z(z, nk, tile, kk);
    dispatch<0>(ph, tid, bx * TX, by * TY, kk, tile, s.base);
  }
};

#ifndef FV3LM_HOST_EMU
template <class K> GLOBAL void __launch_bounds__(NTHR) kern_chain(const __grid_constant__ K k) {
  extern __shared__ double chain_smem[];
  typename K::Smem s{chain_smem};
  const int tid = threadIdx.x;
#pragma unroll
  for (int ph = 0; ph < K::NPH; ph++) {
    k.phase(ph, tid, blockIdx.x, blockIdx.y, blockIdx.z, s);
    if (ph + 1 < K::NPH) __syncthreads();
  }
}
// a B200 SM has 227 KB of shared memory a block may opt in to (the default limit without the attribute is 48 KB)
constexpr size_t CHAIN_SMEM_MAX = 227 * 1024;
template <class K> void launch_chain(const K& k, int nx, int ny, int nz, size_t smem_bytes) {
  if (nz <= 0) return;
  if (smem_bytes > CHAIN_SMEM_MAX) throw std::runtime_error("fused chain: tiles exceed 227 KB of shared memory");
  static size_t opted = 0;     // per kernel instantiation
  if (smem_bytes > 48 * 1024 && smem_bytes > opted) {
    if (cudaFuncSetAttribute(kern_chain<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM_MAX) != cudaSuccess)
      throw std::runtime_error("fused chain: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
    opted = CHAIN_SMEM_MAX;
  }
  dim3 b(NTHR, 1, 1), gr((nx + TX - 1) / TX, (ny + TY - 1) / TY, nz);
  kern_chain<K><<<gr, b, smem_bytes, dev::stream()>>>(k);
  dev::launches++;
}
#else
constexpr size_t CHAIN_SMEM_MAX = 227 * 1024;
template <class K> void launch_chain(const K& k, int nx, int ny, int nz, size_t smem_bytes) {
  // the same limit as the device launcher, so that the CPU suite fails where the GPU would
  if (smem_bytes > CHAIN_SMEM_MAX) throw std::runtime_error("fused chain: tiles exceed 227 KB of shared memory");
  std::vector<double> buf(smem_bytes / sizeof(double) + 1);
  typename K::Smem s{buf.data()};
  for (int z = 0; z < nz; z++)
    for (int by = 0; by < (ny + TY - 1) / TY; by++)
      for (int bx = 0; bx < (nx + TX - 1) / TX; bx++) {
        memset(buf.data(), 0xff, buf.size() * sizeof(double));       // a new block finds arbitrary shared memory
        for (int ph = 0; ph < K::NPH; ph++)
          for (int tid = 0; tid < NTHR; tid++) k.phase(ph, tid, bx, by, z, s);
      }
  dev::launches++;
}
#endif

// ---- host side: regions and tiles from the taps -------------------------------------------------------------
struct HBox {
  int x0 = 0, y0 = 0, x1 = -1, y1 = -1;      // inclusive; empty when x1 < x0
  bool empty() const { return x1 < x0; }
  void add(int a0, int b0, int a1, int b1) {
    if (empty()) { x0 = a0; y0 = b0; x1 = a1; y1 = b1; }
    else { x0 = std::min(x0, a0); y0 = std::min(y0, b0); x1 = std::max(x1, a1); y1 = std::max(y1, b1); }
  }
};
template <class St> void chain_taps(int f, int& dx0, int& dx1, int& dy0, int& dy1, bool& has_dk) {
  dx0 = dy0 = 1 << 20; dx1 = dy1 = -(1 << 20); has_dk = false;
  for (int n = 0; n < St::NT; n++)
    if (St::taps[n].f == f) {
      dx0 = std::min(dx0, St::taps[n].di); dx1 = std::max(dx1, St::taps[n].di);
      dy0 = std::min(dy0, St::taps[n].dj); dy1 = std::max(dy1, St::taps[n].dj);
      if (St::taps[n].dk != 0) has_dk = true;
    }
}

template <class... S> struct ChainBuilder {
  static constexpr int NS = sizeof...(S);
  std::vector<std::vector<int>> ins, outs;     // program value ids per stage
  std::vector<int> ids;                        // distinct values of the chain -> slot
  std::vector<HBox> need;                      // per slot: where consumers inside the chain read it
  std::vector<HBox> region;                    // per stage
  int slot(int id) { for (size_t n = 0; n < ids.size(); n++) if (ids[n] == id) return (int)n; ids.push_back(id); need.emplace_back(); return (int)ids.size() - 1; }
  bool produced(int id, int before) const { for (int s = 0; s < before; s++) for (int o : outs[s]) if (o == id) return true; return false; }
  template <int s> void back(const std::vector<int>& live_out) {
    if constexpr (s >= 0) {
      using St = typename NthStage<s, S...>::type;
      HBox r;
      for (int o : outs[s]) {
        const int sl = slot(o);
        if (!need[sl].empty()) r.add(need[sl].x0, need[sl].y0, need[sl].x1, need[sl].y1);
        for (int l : live_out) if (l == o) r.add(0, 0, TX - 1, TY - 1);
      }
      if (r.empty()) throw std::runtime_error("fused chain: a stage whose outputs nobody reads");
      region[s] = r;
      for (int f = 0; f < St::NI; f++) {
        const int id = ins[s][f];
        if (!produced(id, s)) { slot(id); continue; }
        int dx0, dx1, dy0, dy1; bool dk;
        chain_taps<St>(f, dx0, dx1, dy0, dy1, dk);
        if (dk) throw std::runtime_error("fused chain: a chain-internal value is read at another level");
        if (dx1 < dx0) continue;
        need[slot(id)].add(r.x0 + dx0, r.y0 + dy0, r.x1 + dx1, r.y1 + dy1);
      }
      back<s - 1>(live_out);
    }
  }
};

template <class... S>
void add_chain(Program& P, const std::string& name, PPack<S...> prm, std::vector<std::vector<int>> ins,
               std::vector<std::vector<int>> outs, std::vector<int> live_out, int nk) {
  constexpr int NS = sizeof...(S);
  if ((int)ins.size() != NS || (int)outs.size() != NS) throw std::runtime_error("fused chain: arity");
  ChainBuilder<S...> cb; cb.ins = ins; cb.outs = outs; cb.region.resize(NS);
  for (int s = 0; s < NS; s++) { for (int o : outs[s]) cb.slot(o); for (int i : ins[s]) cb.slot(i); }
  cb.template back<NS - 1>(live_out);
  if ((int)cb.ids.size() > CH_MAXV) throw std::runtime_error("fused chain: too many values");
  // op inputs = values entering the chain, outputs = live_out
  Op op; op.name = name; op.nk_launch = nk; op.tl_only = P.tl_only; op.variant = VAR_FWD;
  for (size_t n = 0; n < cb.ids.size(); n++) {
    bool prod = false;
    for (int s = 0; s < NS; s++) for (int o : outs[s]) if (o == cb.ids[n]) prod = true;
    if (!prod) op.in.push_back(cb.ids[n]);
  }
  op.out = live_out;
  auto shared = std::make_shared<ChainBuilder<S...>>(cb);
  op.run = [shared, prm, live_out](Program& P, Op& o, int mode) {
    if (mode != MODE_NL && mode != MODE_TL) throw std::runtime_error("fused chain ops run in forward sweeps only");
    const ChainBuilder<S...>& cb = *shared;
    const Geom& g = P.dv->g;
    bool tl = false;
    if (mode == MODE_TL) for (int i : o.in) tl = tl || (P.vals[i].active && P.vals[i].pert);
    ChainParams<NS> cp{};
    for (size_t n = 0; n < cb.ids.size(); n++) {
      const Value& v = P.vals[cb.ids[n]];
      ChainVal& cv = cp.v[n];
      bool prod = false, live = false;
      for (int s = 0; s < NS; s++) for (int oo : cb.outs[s]) if (oo == cb.ids[n]) prod = true;
      for (int l : live_out) if (l == cb.ids[n]) live = true;
      cv.gin = prod ? Fld{nullptr, nullptr, v.nk} : fld(v, tl);
      cv.gout = (prod && live) ? ofld(v, tl) : OFld{nullptr, nullptr, v.nk};
      cv.soff = -1; cv.x0 = cv.y0 = cv.w = cv.h = 0;
    }
    int off = 0;
    for (int s = 0; s < NS; s++) {
      for (size_t f = 0; f < cb.ins[s].size(); f++) for (size_t n = 0; n < cb.ids.size(); n++) if (cb.ids[n] == cb.ins[s][f]) cp.in_slot[s][f] = (int)n;
      for (size_t f = 0; f < cb.outs[s].size(); f++) for (size_t n = 0; n < cb.ids.size(); n++) if (cb.ids[n] == cb.outs[s][f]) cp.out_slot[s][f] = (int)n;
      cp.rx0[s] = cb.region[s].x0; cp.ry0[s] = cb.region[s].y0;
      cp.rw[s] = cb.region[s].x1 - cb.region[s].x0 + 1; cp.rh[s] = cb.region[s].y1 - cb.region[s].y0 + 1;
      // a value that a later stage of the chain reads gets a tile as large as its producer's region (the union of what the
      // consumers of ALL outputs of that stage need), so that out() never stores outside a box
      for (size_t f = 0; f < cb.outs[s].size(); f++) {
        ChainVal& cv = cp.v[cp.out_slot[s][f]];
        if (cb.need[cp.out_slot[s][f]].empty()) continue;
        cv.x0 = cp.rx0[s]; cv.y0 = cp.ry0[s]; cv.w = cp.rw[s]; cv.h = cp.rh[s];
        cv.soff = off; off += cv.w * cv.h;
      }
    }
    cp.tile_doubles = off;
    const size_t bytes = (size_t)off * sizeof(double) * (tl ? 2 : 1) + sizeof(double);
    auto go = [&](auto kern) {
      kern.g = g; kern.m = P.dv->m; kern.nk = o.nk_launch; kern.cp = cp; kern.prm = prm;
      launch_chain(kern, g.NX, g.NY, g.ntile * o.nk_launch, bytes);
    };
    if (tl) go(KernChain<Dual, S...>{}); else go(KernChain<double, S...>{});
  };
  P.ops.push_back(op);
}

}  // namespace ftp
}  // namespace fv3lm
