"""Python cubed-sphere index maps (synthetic-input generation and the oracle's halo exchange;
the product's halo maps are built independently in csrc/mosaic.cu).

Cubed-sphere mosaic connectivity, halo-exchange index maps and corner ghost-cell
fills, restated from the reference as pure index maps.

parity unpinned: the reference ships no golden vectors; the
vector rotation/sign rules live in un-vendored FMS (mpp_domains_mod, no version
pin, CMakeLists.txt:70-74) and are re-derived here from the contact table
 tools/fv_mp_nlm_mod.F90:524-573 and checked by geometric invariants in
tests/test_invariants.py.

Array convention used by the whole oracle: a per-tile 2-D slab is stored as
[..., NY, NX] with NX = NY = N + 2*ng + 1 (ng = 3, tools/fv_mp_nlm_mod.F90:63),
Fortran tile-global index (i, j) lives at [..., j + ng - 1, i + ng - 1].
"""
import numpy as np

NG = 3

# staggering of a field: continuous coordinate of index (i,j) is
#   x = i - 1 + ox ,  y = j - 1 + oy       (cell (1,1) spans [0,1]x[0,1])
STAG = {
    "center": (0.5, 0.5),   # A-grid scalars
    "corner": (0.0, 0.0),   # B-grid / cell corners
    "ystag": (0.5, 0.0),    # D-grid u, C-grid vc   (x centre, y corner)
    "xstag": (0.0, 0.5),    # D-grid v, C-grid uc   (x corner, y centre)
}


def off(ng=NG):
    return ng - 1


def R(a, b, ng=NG):
    """python slice for the inclusive Fortran index range a..b"""
    return slice(a + ng - 1, b + ng)


def neighbour(t, side):
    """Contact table tools/fv_mp_nlm_mod.F90:524-573, 0-based
    tile t.  Returns (tile_b, kind) with kind in {'aligned', 'rotp', 'rotm'}:
    the map from A's extended continuous coords (x,y) to B's coords for that side."""
    odd = (t % 2 == 0)  # tiles 1,3,5 are 0-based 0,2,4
    if odd:
        return {"E": ((t + 1) % 6, "E_al"), "N": ((t + 2) % 6, "N_rot"),
                "W": ((t - 2) % 6, "W_rot"), "S": ((t - 1) % 6, "S_al")}[side]
    return {"N": ((t + 1) % 6, "N_al"), "E": ((t + 2) % 6, "E_rot"),
            "S": ((t - 2) % 6, "S_rot"), "W": ((t - 1) % 6, "W_al")}[side]


def to_neighbour(kind, x, y, N):
    """(x,y) in A's frame -> (xb, yb, rot) in B's frame. rot = 0: axes aligned;
    rot = +1: A's e_x -> -B's e_y, A's e_y -> +B's e_x   (uA = -vB, vA = +uB);
    rot = -1: A's e_x -> +B's e_y, A's e_y -> -B's e_x   (uA = +vB, vA = -uB)."""
    if kind == "E_al":
        return x - N, y, 0
    if kind == "W_al":
        return x + N, y, 0
    if kind == "N_al":
        return x, y - N, 0
    if kind == "S_al":
        return x, y + N, 0
    if kind == "N_rot":   # odd tile north -> (t+2) west
        return y - N, N - x, +1
    if kind == "S_rot":   # even tile south -> (t-2) east
        return y + N, N - x, +1
    if kind == "W_rot":   # odd tile west -> (t-2) north
        return N - y, x + N, -1
    if kind == "E_rot":   # even tile east -> (t+2) south
        return N - y, x - N, -1
    raise ValueError(kind)


def _side_of(x, y, N):
    outx = (x < 0) or (x > N)
    outy = (y < 0) or (y > N)
    if outx and outy:
        return None  # corner ghost region: never exchanged
    if x > N:
        return "E"
    if x < 0:
        return "W"
    if y > N:
        return "N"
    if y < 0:
        return "S"
    return "in"


class HaloMap:
    """dst[tile, j, i] = sign * src_field[src_tile, sj, si]; comp says which member
    of a (x-comp, y-comp) pair is read (0 or 1).  Flat int arrays."""

    def __init__(self, dt, dj, di, st, sc, sj, si, sg):
        self.dt, self.dj, self.di = dt, dj, di
        self.st, self.sc, self.sj, self.si, self.sg = st, sc, sj, si, sg


def build_halo_map(N, stag, ng=NG, pair_stag=None, comp=0, vector_sign=True, halo=None):
    """Halo map for one field of staggering `stag`.  For a vector/scalar pair pass
    pair_stag = (stag_x_component, stag_y_component) and comp = which one `stag` is.
    vector_sign=False reproduces mpp SCALAR_PAIR (swap, no sign flip)."""
    ox, oy = STAG[stag]
    h = ng if halo is None else halo
    o = ng - 1
    nxi = N + (1 if ox == 0.0 else 0)   # number of owned indices in x
    nyj = N + (1 if oy == 0.0 else 0)
    rec = []
    for t in range(6):
        for j in range(1 - h, nyj + h + 1):
            for i in range(1 - h, nxi + h + 1):
                x = i - 1 + ox
                y = j - 1 + oy
                side = _side_of(x, y, N)
                if side is None or side == "in":
                    continue
                tb, kind = neighbour(t, side)
                xb, yb, rot = to_neighbour(kind, x, y, N)
                sgn = 1.0
                scomp = comp
                if pair_stag is not None and rot != 0:
                    scomp = 1 - comp
                    if vector_sign:
                        # rot=+1: uA=-vB, vA=+uB ; rot=-1: uA=+vB, vA=-uB
                        if rot == +1:
                            sgn = -1.0 if comp == 0 else 1.0
                        else:
                            sgn = 1.0 if comp == 0 else -1.0
                if pair_stag is not None:
                    sox, soy = STAG[pair_stag[scomp]]
                else:
                    sox, soy = ox, oy
                    if rot != 0:
                        sox, soy = oy, ox  # centre/corner are symmetric anyway
                si = xb - sox + 1
                sj = yb - soy + 1
                assert abs(si - round(si)) < 1e-9 and abs(sj - round(sj)) < 1e-9, (stag, t, i, j)
                si, sj = int(round(si)), int(round(sj))
                rec.append((t, j + o, i + o, tb, scomp, sj + o, si + o, sgn))
    a = np.array(rec, dtype=np.float64)
    return HaloMap(a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2].astype(np.int64),
                   a[:, 3].astype(np.int64), a[:, 4].astype(np.int64), a[:, 5].astype(np.int64),
                   a[:, 6].astype(np.int64), a[:, 7].copy())


class Halo:
    """FMS mpp_update_domains replacement for whole-tile decomposition (6 tiles).
    All methods are functional (return new arrays) and work on numpy arrays or
    torch tensors shaped [6, ..., NY, NX]."""

    def __init__(self, N, ng=NG):
        self.N, self.ng = N, ng
        self.m_center = build_halo_map(N, "center", ng)
        self.m_corner = build_halo_map(N, "corner", ng)
        # D-grid (u: ystag = x component, v: xstag = y component)
        ps = ("ystag", "xstag")
        self.m_du = build_halo_map(N, "ystag", ng, ps, 0)
        self.m_dv = build_halo_map(N, "xstag", ng, ps, 1)
        self.m_du_sp = build_halo_map(N, "ystag", ng, ps, 0, vector_sign=False)
        self.m_dv_sp = build_halo_map(N, "xstag", ng, ps, 1, vector_sign=False)
        # C-grid (uc: xstag = x component, vc: ystag = y component)
        pc = ("xstag", "ystag")
        self.m_cu = build_halo_map(N, "xstag", ng, pc, 0)
        self.m_cv = build_halo_map(N, "ystag", ng, pc, 1)
        self.m_cu_sp = build_halo_map(N, "xstag", ng, pc, 0, vector_sign=False)
        self.m_cv_sp = build_halo_map(N, "ystag", ng, pc, 1, vector_sign=False)

    @staticmethod
    def _apply(m, dst, srcs):
        """dst, srcs[c]: [6, K..., NY, NX] -> new dst with halo filled."""
        is_np = isinstance(dst, np.ndarray)
        out = dst.copy() if is_np else dst.clone()
        for c in (0, 1):
            sel = (m.sc == c)
            if not sel.any():
                continue
            src = srcs[c]
            if is_np:
                sg = m.sg[sel]
                vals = src[m.st[sel], ..., m.sj[sel], m.si[sel]]      # [n, K...]
                sg = sg.reshape((-1,) + (1,) * (vals.ndim - 1))
                out[m.dt[sel], ..., m.dj[sel], m.di[sel]] = sg * vals
            else:
                import torch
                st = torch.as_tensor(m.st[sel]); sj = torch.as_tensor(m.sj[sel]); si = torch.as_tensor(m.si[sel])
                dt = torch.as_tensor(m.dt[sel]); dj = torch.as_tensor(m.dj[sel]); di = torch.as_tensor(m.di[sel])
                sg = torch.as_tensor(m.sg[sel], dtype=src.dtype)
                nd = src.dim()
                if nd == 3:
                    out[dt, dj, di] = sg * src[st, sj, si]
                else:
                    # [6, K, NY, NX] -> index with K kept
                    vals = src[st, :, sj, si]                              # [n, K]
                    out[dt, :, dj, di] = sg[:, None] * vals
        return out

    def scalar(self, q):
        """mpp_update_domains(q, domain)  (CENTER)"""
        return self._apply(self.m_center, q, (q, q))

    def corner(self, q):
        """mpp_update_domains(q, domain, position=CORNER)"""
        return self._apply(self.m_corner, q, (q, q))

    def dgrid(self, u, v, scalar_pair=False):
        """mpp_update_domains(u, v, domain, gridtype=DGRID_NE)"""
        mu, mv = (self.m_du_sp, self.m_dv_sp) if scalar_pair else (self.m_du, self.m_dv)
        return self._apply(mu, u, (u, v)), self._apply(mv, v, (u, v))

    def cgrid(self, uc, vc, scalar_pair=False):
        """mpp_update_domains(uc, vc, domain, gridtype=CGRID_NE)"""
        mu, mv = (self.m_cu_sp, self.m_cv_sp) if scalar_pair else (self.m_cu, self.m_cv)
        return self._apply(mu, uc, (uc, vc)), self._apply(mv, vc, (uc, vc))


# ----------------------------------------------------------------------------------
# corner ghost fills -- all are "dst cells <- src cells" within one tile slab; every
# tile of a whole-tile decomposition has all four cube corners.
# ----------------------------------------------------------------------------------

def _assign(q, pairs, sign=1.0, src=None):
    """pairs: list of ((i,j) dst, (i,j) src) in Fortran indices. functional."""
    o = NG - 1
    is_np = isinstance(q, np.ndarray)
    out = q.copy() if is_np else q.clone()
    s = q if src is None else src
    dj = [p[0][1] + o for p in pairs]; di = [p[0][0] + o for p in pairs]
    sj = [p[1][1] + o for p in pairs]; si = [p[1][0] + o for p in pairs]
    if isinstance(sign, (list, tuple)):
        if is_np:
            sg = np.array(sign)
        else:
            import torch
            sg = torch.tensor(sign, dtype=q.dtype)
        out[..., dj, di] = sg * s[..., sj, si]
    else:
        out[..., dj, di] = sign * s[..., sj, si]
    return out


def copy_corners_pairs(npx, npy, dir, ng=NG):
    """model/tp_core_nlm.F90:214-289"""
    p = []
    for j in range(1 - ng, 1):
        for i in range(1 - ng, 1):
            p.append(((i, j), (j, 1 - i)) if dir == 1 else ((i, j), (1 - j, i)))          # SW
    for j in range(1 - ng, 1):
        for i in range(npx, npx + ng):
            p.append(((i, j), (npy - j, i - npx + 1)) if dir == 1 else ((i, j), (npy + j - 1, npx - i)))  # SE
    for j in range(npy, npy + ng):
        for i in range(npx, npx + ng):
            p.append(((i, j), (j, 2 * npx - 1 - i)) if dir == 1 else ((i, j), (2 * npy - 1 - j, i)))      # NE
    for j in range(npy, npy + ng):
        for i in range(1 - ng, 1):
            p.append(((i, j), (npy - j, i - 1 + npx)) if dir == 1 else ((i, j), (j + 1 - npx, npy - i)))  # NW
    return p


def copy_corners(q, npx, npy, dir):
    return _assign(q, copy_corners_pairs(npx, npy, dir))


def fill_4corners_pairs(npx, npy, dir):
    """model/sw_core_nlm.F90:3102-3295 (fill_4corners / fill2_4corners / fill3_4corners)"""
    if dir == 1:
        return [((-1, 0), (0, 2)), ((0, 0), (0, 1)),                               # SW
                ((npx + 1, 0), (npx, 2)), ((npx, 0), (npx, 1)),                    # SE
                ((npx, npy), (npx, npy - 1)), ((npx + 1, npy), (npx, npy - 2)),    # NE
                ((0, npy), (0, npy - 1)), ((-1, npy), (0, npy - 2))]               # NW
    return [((0, 0), (1, 0)), ((0, -1), (2, 0)),                                   # SW
            ((npx, 0), (npx - 1, 0)), ((npx, -1), (npx - 2, 0)),                   # SE
            ((npx, npy), (npx - 1, npy)), ((npx, npy + 1), (npx - 2, npy)),        # NE
            ((0, npy), (1, npy)), ((0, npy + 1), (2, npy))]                        # NW


def fill_4corners(q, npx, npy, dir):
    return _assign(q, fill_4corners_pairs(npx, npy, dir))


def fill_corners_bgrid_pairs(npx, npy, fill, ng=NG):
    """tools/fv_mp_nlm_mod.F90:1046-1083 (BGRID), fill in {'x','y'}"""
    p = []
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            if fill == "x":
                p += [((1 - i, 1 - j), (1 - j, i + 1)), ((1 - i, npy + j), (1 - j, npy - i)),
                      ((npx + i, 1 - j), (npx + j, i + 1)), ((npx + i, npy + j), (npx + j, npy - i))]
            else:
                p += [((1 - j, 1 - i), (i + 1, 1 - j)), ((1 - j, npy + i), (i + 1, npy + j)),
                      ((npx + j, 1 - i), (npx - i, 1 - j)), ((npx + j, npy + i), (npx - i, npy + j))]
    return p


def fill_corners_bgrid(q, npx, npy, fill):
    return _assign(q, fill_corners_bgrid_pairs(npx, npy, fill))


def fill_corners_agrid_scalar(q, npx, npy, fill, ng=NG):
    """tools/fv_mp_nlm_mod.F90:1085-1116 (AGRID scalar)"""
    p = []
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            if fill == "x":
                p += [((1 - i, 1 - j), (1 - j, i)), ((1 - i, npy - 1 + j), (1 - j, npy - i)),
                      ((npx - 1 + i, 1 - j), (npx - 1 + j, i)), ((npx - 1 + i, npy - 1 + j), (npx - 1 + j, npy - i))]
            else:
                p += [((1 - j, 1 - i), (i, 1 - j)), ((1 - j, npy - 1 + i), (i, npy - 1 + j)),
                      ((npx - 1 + j, 1 - i), (npx - i, 1 - j)), ((npx - 1 + j, npy - 1 + i), (npx - i, npy - 1 + j))]
    return _assign(q, p)


def fill_corners_dgrid(x, y, npx, npy, sign, ng=NG):
    """tools/fv_mp_nlm_mod.F90:1271-1303.  sign = -1 for VECTOR=.true., +1 otherwise."""
    px, sx, py, sy = [], [], [], []
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            px += [((1 - i, 1 - j), (1 - j, i)), ((1 - i, npy + j), (1 - j, npy - i)),
                   ((npx - 1 + i, 1 - j), (npx + j, i)), ((npx - 1 + i, npy + j), (npx + j, npy - i))]
            sx += [sign, 1.0, 1.0, sign]
            py += [((1 - i, 1 - j), (j, 1 - i)), ((1 - i, npy - 1 + j), (j, npy + i)),
                   ((npx + i, 1 - j), (npx - j, 1 - i)), ((npx + i, npy - 1 + j), (npx - j, npy + i))]
            sy += [sign, 1.0, 1.0, sign]
    xn = _assign(x, px, sx, src=y)
    yn = _assign(y, py, sy, src=x)   # reads the ORIGINAL x (sources are non-corner cells)
    return xn, yn


def fill_corners_cgrid(x, y, npx, npy, sign, ng=NG):
    """tools/fv_mp_nlm_mod.F90:1383-1407"""
    px, sx, py, sy = [], [], [], []
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            px += [((1 - i, 1 - j), (j, 1 - i)), ((1 - i, npy - 1 + j), (j, npy + i)),
                   ((npx + i, 1 - j), (npx - j, 1 - i)), ((npx + i, npy - 1 + j), (npx - j, npy + i))]
            sx += [1.0, sign, sign, 1.0]
            py += [((1 - i, 1 - j), (1 - j, i)), ((1 - i, npy + j), (1 - j, npy - i)),
                   ((npx - 1 + i, 1 - j), (npx + j, i)), ((npx - 1 + i, npy + j), (npx + j, npy - i))]
            sy += [1.0, sign, sign, 1.0]
    xn = _assign(x, px, sx, src=y)
    yn = _assign(y, py, sy, src=x)
    return xn, yn


def fill_corners_agrid_pair(x, y, npx, npy, sign, ng=NG):
    """tools/fv_mp_nlm_mod.F90:1440-1470"""
    px, sx, py, sy = [], [], [], []
    for j in range(1, ng + 1):
        for i in range(1, ng + 1):
            px += [((1 - i, 1 - j), (1 - j, i)), ((1 - i, npy - 1 + j), (1 - j, npy - i)),
                   ((npx - 1 + i, 1 - j), (npx - 1 + j, i)), ((npx - 1 + i, npy - 1 + j), (npx - 1 + j, npy - i))]
            sx += [sign, 1.0, 1.0, sign]
            py += [((1 - j, 1 - i), (i, 1 - j)), ((1 - j, npy - 1 + i), (i, npy - 1 + j)),
                   ((npx - 1 + j, 1 - i), (npx - i, 1 - j)), ((npx - 1 + j, npy - 1 + i), (npx - i, npy - 1 + j))]
            sy += [sign, 1.0, 1.0, sign]
    xn = _assign(x, px, sx, src=y)
    yn = _assign(y, py, sy, src=x)
    return xn, yn


def fill_ghost(q, npx, npy, value, ng=NG):
    """model/fv_grid_utils_nlm.F90:3037-3074: set the four corner ghost blocks of an
    A-grid array (indices 1-ng..npx-1+ng) to `value`."""
    out = q.copy() if isinstance(q, np.ndarray) else q.clone()
    lo = R(1 - ng, 0)
    hix = R(npx, npx - 1 + ng)
    hiy = R(npy, npy - 1 + ng)
    out[..., lo, lo] = value
    out[..., lo, hix] = value
    out[..., hiy, hix] = value
    out[..., hiy, lo] = value
    return out
