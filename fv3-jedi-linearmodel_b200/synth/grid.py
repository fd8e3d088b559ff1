"""Synthetic grid metrics (input generation for tests and bench.py -- in a real coupling the
host model passes its own gridstruct through fv3lm_set_metric).  Gnomonic equal-edge
cubed-sphere grid and the FV3 metric terms, restated in numpy from

  model/fv_grid_utils_nlm.F90  gnomonic_ed :1240-1336, symm_ed :1514, grid_utils_init :78-760,
                               edge_factors :1105, efactor_a2c_v :926, get_center_vect :1722
  tools/fv_grid_tools_nlm.F90  init_grid :417-940, mirror_grid :2181-2314, grid_area :1932-2143

Deviation (documented in DESIGN.md): the reference sorts the four corner points of a
cell before summing (sorted_inta/sorted_intb) to make tiles bitwise mirror-symmetric;
we sum in natural order (differences are O(1e-16) relative).

The reference holds no golden grid data; tests/test_invariants.py
checks invariants (sum(area)=4 pi R^2, edge continuity, sin^2+cos^2=1, symmetry).

Every array is [6, NY, NX] (or [6, NY, NX, c]); Fortran index (i,j) -> [j+2, i+2].
"""
import numpy as np
from . import cubed_sphere as cs
from .cubed_sphere import R, NG

RADIUS = 6371.0e3               # utils/fv3jedi_lm_const_mod.F90
OMEGA = 2.0 * np.pi / 86164.0
BIG = 1.0e8                     # model/fv_grid_utils_nlm.F90:49-50
TINY = 1.0e-8


# ------------------------------------------------------------------ small helpers
def ll2xyz(p):
    lon, lat = p[..., 0], p[..., 1]
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], -1)


def xyz2ll(e):
    e = e / np.linalg.norm(e, axis=-1, keepdims=True)
    lon = np.where((np.abs(e[..., 0]) + np.abs(e[..., 1])) < 1e-10, 0.0, np.arctan2(e[..., 1], e[..., 0]))
    lon = np.where(lon < 0.0, lon + 2.0 * np.pi, lon)
    lat = np.arcsin(np.clip(e[..., 2], -1.0, 1.0))
    return np.stack([lon, lat], -1)


def gc_dist(q1, q2, radius=1.0):
    """great_circle_dist  model/fv_grid_utils_nlm.F90:1967"""
    beta = 2.0 * np.arcsin(np.sqrt(np.sin((q1[..., 1] - q2[..., 1]) / 2.0) ** 2 +
                                   np.cos(q1[..., 1]) * np.cos(q2[..., 1]) *
                                   np.sin((q1[..., 0] - q2[..., 0]) / 2.0) ** 2))
    return radius * beta


def mid3(e1, e2):
    e = e1 + e2
    return e / np.linalg.norm(e, axis=-1, keepdims=True)


def mid_ll(p1, p2):
    return xyz2ll(mid3(ll2xyz(p1), ll2xyz(p2)))


def cross(a, b):
    return np.cross(a, b)


def unit(a):
    return a / np.linalg.norm(a, axis=-1, keepdims=True)


def spherical_angle(p1, p2, p3):
    """model/fv_grid_utils_nlm.F90:2765"""
    p = cross(p1, p2)
    q = cross(p1, p3)
    ddd = (p * p).sum(-1) * (q * q).sum(-1)
    c = (p * q).sum(-1) / np.sqrt(np.where(ddd > 0, ddd, 1.0))
    ang = np.arccos(np.clip(c, -1.0, 1.0))
    return np.where(ddd <= 0.0, 0.0, ang)


def cos_angle(p1, p2, p3):
    """model/fv_grid_utils_nlm.F90:2825"""
    p = cross(p1, p2)
    q = cross(p1, p3)
    ddd = np.sqrt((p * p).sum(-1) * (q * q).sum(-1))
    return np.where(ddd > 0.0, (p * q).sum(-1) / np.where(ddd > 0, ddd, 1.0), 1.0)


def get_area(p1, p4, p2, p3, radius=RADIUS):
    """model/fv_grid_utils_nlm.F90:2676 (argument order of the reference kept)"""
    e1, e2, e3, e4 = ll2xyz(p1), ll2xyz(p2), ll2xyz(p3), ll2xyz(p4)
    a1 = spherical_angle(e1, e2, e4)
    a2 = spherical_angle(e2, e3, e1)
    a3 = spherical_angle(e3, e4, e2)
    a4 = spherical_angle(e4, e3, e1)
    return (a1 + a2 + a3 + a4 - 2.0 * np.pi) * radius ** 2


def get_area_tri(p1, p2, p3, radius=RADIUS):
    e1, e2, e3 = ll2xyz(p1), ll2xyz(p2), ll2xyz(p3)
    # get_angle(ndims=2,p1,p2,p3) = spherical_angle(e(p2), e(p1), e(p3))
    a = spherical_angle(e2, e1, e3)
    b = spherical_angle(e3, e2, e1)
    c = spherical_angle(e1, e3, e2)
    return (a + b + c - np.pi) * radius ** 2


# ------------------------------------------------------------------ tile-1 grid
def _mirror_latlon(lon1, lat1, lon2, lat2, lon0, lat0):
    p0 = ll2xyz(np.array([lon0, lat0])); p1 = ll2xyz(np.array([lon1, lat1])); p2 = ll2xyz(np.array([lon2, lat2]))
    nb = unit(cross(p1, p2))
    pp = p0 - 2.0 * (p0 * nb).sum() * nb
    return xyz2ll(pp)


def gnomonic_ed(im):
    """model/fv_grid_utils_nlm.F90:1240-1336; returns lamda, theta [im+1(i), im+1(j)] indexed [i-1, j-1]"""
    rsq3 = 1.0 / np.sqrt(3.0)
    alpha = np.arcsin(rsq3)
    dely = 2.0 * alpha / im
    lam = np.zeros((im + 1, im + 1)); the = np.zeros((im + 1, im + 1))
    for j in range(im + 1):
        lam[0, j] = 0.75 * np.pi; lam[im, j] = 1.25 * np.pi
        the[0, j] = -alpha + dely * j; the[im, j] = the[0, j]
    for i in range(1, im):
        ll = _mirror_latlon(lam[0, 0], the[0, 0], lam[im, im], the[im, im], lam[0, i], the[0, i])
        lam[i, 0], the[i, 0] = ll[0], ll[1]
        lam[i, im] = lam[i, 0]; the[i, im] = -the[i, 0]
    pp = np.zeros((3, im + 1, im + 1))
    for (i, j) in ((0, 0), (im, 0), (0, im), (im, im)):
        pp[:, i, j] = ll2xyz(np.array([lam[i, j], the[i, j]]))
    for j in range(1, im):
        e = ll2xyz(np.array([lam[0, j], the[0, j]]))
        pp[:, 0, j] = e
        pp[1, 0, j] = -e[1] * rsq3 / e[0]; pp[2, 0, j] = -e[2] * rsq3 / e[0]
    for i in range(1, im):
        e = ll2xyz(np.array([lam[i, 0], the[i, 0]]))
        pp[:, i, 0] = e
        pp[1, i, 0] = -e[1] * rsq3 / e[0]; pp[2, i, 0] = -e[2] * rsq3 / e[0]
    pp[0, :, :] = -rsq3
    for j in range(1, im + 1):
        for i in range(1, im + 1):
            pp[1, i, j] = pp[1, i, 0]
            pp[2, i, j] = pp[2, 0, j]
    ll = xyz2ll(np.moveaxis(pp, 0, -1))
    return ll[..., 0].copy(), ll[..., 1].copy()


def symm_ed(im, lam, the):
    """model/fv_grid_utils_nlm.F90:1514-1553 (arrays indexed [i-1, j-1])"""
    for j in range(1, im + 1):
        for i in range(1, im):
            lam[i, j] = lam[i, 0]
    for j in range(im + 1):
        for i in range(im // 2):
            ip = im - i
            avg = 0.5 * (lam[i, j] - lam[ip, j])
            lam[i, j] = avg + np.pi; lam[ip, j] = np.pi - avg
            avg = 0.5 * (the[i, j] + the[ip, j])
            the[i, j] = avg; the[ip, j] = avg
    for j in range(im // 2):
        jp = im - j
        for i in range(1, im):
            avg = 0.5 * (lam[i, j] + lam[i, jp])
            lam[i, j] = avg; lam[i, jp] = avg
            avg = 0.5 * (the[i, j] - the[i, jp])
            the[i, j] = avg; the[i, jp] = -avg
    return lam, the


def _rot3d(axis, lon, lat, ang_deg):
    """tools/fv_grid_tools_nlm.F90:1830 with degrees+convert: spherical in/out"""
    a = np.deg2rad(ang_deg)
    c, s = np.cos(a), np.sin(a)
    # spherical_to_cartesian without -DRIGHT_HAND (the reference build does not define it,
    # cmake/fv3jedilm_compiler_flags.cmake:8): z = -r sin(lat)
    x1 = np.cos(lat) * np.cos(lon); y1 = np.cos(lat) * np.sin(lon); z1 = -np.sin(lat)
    if axis == 1:
        x2, y2, z2 = x1, c * y1 + s * z1, -s * y1 + c * z1
    elif axis == 2:
        x2, y2, z2 = c * x1 - s * z1, y1, s * x1 + c * z1
    else:
        x2, y2, z2 = c * x1 + s * y1, -s * x1 + c * y1, z1
    # cartesian_to_spherical (tools/fv_grid_tools_nlm.F90:1793): lon in [-pi,pi], lat = acos(z/r) - pi/2
    r = np.sqrt(x2 * x2 + y2 * y2 + z2 * z2)
    lon2 = np.where((np.abs(x2) + np.abs(y2)) < 1e-10, 0.0, np.arctan2(y2, x2))
    lat2 = np.arccos(np.clip(z2 / r, -1.0, 1.0)) - np.pi / 2
    return lon2, lat2


def mirror_grid(lam1, the1, npx):
    """tools/fv_grid_tools_nlm.F90:2181-2314.  in: tile-1 lon/lat [i-1,j-1]; out g[6, j-1, i-1, 2]"""
    npy = npx
    g1 = np.stack([lam1, the1], -1)    # [i, j, 2]
    half = int(np.ceil(npx / 2.0))
    for j in range(half):
        for i in range(half):
            ii, jj = npx - 1 - i, npy - 1 - j
            for c in (0, 1):
                v = 0.25 * (abs(g1[i, j, c]) + abs(g1[ii, j, c]) + abs(g1[i, jj, c]) + abs(g1[ii, jj, c]))
                for (a, b) in ((i, j), (ii, j), (i, jj), (ii, jj)):
                    g1[a, b, c] = np.copysign(v, g1[a, b, c])
            if npx % 2 != 0 and i == (npx - 1) // 2:
                g1[i, j, 0] = 0.0; g1[i, jj, 0] = 0.0
    lon, lat = g1[..., 0], g1[..., 1]
    out = np.zeros((6, npx, npy, 2))
    out[0] = g1
    mid = (npx - 1) // 2          # 0-based index of the centre line when npx odd
    I, J = np.meshgrid(np.arange(npx), np.arange(npy), indexing="ij")
    # tile 2
    l, t = _rot3d(3, lon, lat, -90.0); out[1, ..., 0], out[1, ..., 1] = l, t
    # tile 3
    l, t = _rot3d(3, lon, lat, -90.0); l, t = _rot3d(1, l, t, 90.0)
    if npx % 2 != 0:
        l = np.where((I == mid) & (I == J), 0.0, l); t = np.where((I == mid) & (I == J), np.pi / 2, t)
        l = np.where((J == mid) & (I < mid), 0.0, l)
        l = np.where((J == mid) & (I > mid), np.pi, l)
    out[2, ..., 0], out[2, ..., 1] = l, t
    # tile 4
    l, t = _rot3d(3, lon, lat, -180.0); l, t = _rot3d(1, l, t, 90.0)
    if npx % 2 != 0:
        l = np.where(J == mid, np.pi, l)
    out[3, ..., 0], out[3, ..., 1] = l, t
    # tile 5
    l, t = _rot3d(3, lon, lat, 90.0); l, t = _rot3d(2, l, t, 90.0)
    out[4, ..., 0], out[4, ..., 1] = l, t
    # tile 6
    l, t = _rot3d(2, lon, lat, 90.0); l, t = _rot3d(3, l, t, 0.0)
    if npx % 2 != 0:
        l = np.where((I == mid) & (I == J), 0.0, l); t = np.where((I == mid) & (I == J), -np.pi / 2, t)
        l = np.where((I == mid) & (J > mid), 0.0, l)
        l = np.where((I == mid) & (J < mid), np.pi, l)
    out[5, ..., 0], out[5, ..., 1] = l, t
    return np.transpose(out, (0, 2, 1, 3)).copy()      # -> [6, j, i, 2]


def global_grid(N, shift_fac=18.0):
    """init_grid, tools/fv_grid_tools_nlm.F90:571-622: corner lon/lat of the 6 tiles [6, npy, npx, 2]"""
    npx = N + 1
    lam, the = gnomonic_ed(N)
    lam, the = symm_ed(N, lam, the)
    lam = lam - np.pi
    g = mirror_grid(lam, the, npx)
    g[..., 0] -= np.pi / shift_fac
    g[..., 0] = np.where(g[..., 0] < 0.0, g[..., 0] + 2.0 * np.pi, g[..., 0])
    g[0] = np.where(np.abs(g[0]) < 1e-10, 0.0, g[0])
    # shared edges copied for bitwise consistency (:600-622); g[t, j-1, i-1]
    e = npx - 1
    g[1, :, 0] = g[0, :, e]
    g[2, :, 0] = g[0, e, ::-1]
    g[4, e, :] = g[0, ::-1, 0]
    g[5, e, :] = g[0, 0, :]
    g[2, 0, :] = g[1, e, :]
    g[3, 0, :] = g[1, ::-1, e]
    g[5, :, e] = g[1, 0, ::-1]
    g[3, :, 0] = g[2, :, e]
    g[4, :, 0] = g[2, e, ::-1]
    g[2, :, e] = g[3, :, 0]
    g[4, 0, :] = g[3, e, :]
    g[5, 0, :] = g[3, ::-1, e]
    g[5, :, 0] = g[4, :, e]
    return g


# ------------------------------------------------------------------ metrics
def build_metrics(N, radius=RADIUS, omega=OMEGA, shift_fac=18.0):
    """Returns a dict of numpy arrays [6, NY, NX(,c)] and 1-D edge factor arrays, plus scalars."""
    ng = NG
    npx = npy = N + 1
    NXP = N + 2 * ng + 1
    halo = cs.Halo(N, ng)
    o = ng - 1
    M = {}

    def Z(*extra):
        return np.zeros((6, NXP, NXP) + extra)

    # ---- corner lon/lat incl. halo (init_grid :644-650)
    gg = global_grid(N, shift_fac)
    grid = Z(2)
    grid[:, R(1, npy), R(1, npx), :] = gg
    for c in (0, 1):
        q = halo.corner(np.ascontiguousarray(grid[..., c]))
        grid[..., c] = cs.fill_corners_bgrid(q, npx, npy, "x")
    grid3 = ll2xyz(grid)

    # ---- dx, dy on the extended grid, then the D-grid corner fill (:653-686)
    dx = Z(); dy = Z()
    dx[:, :, :-1] = gc_dist(grid[:, :, 1:], grid[:, :, :-1], radius)        # dx(i,j): (i,j)->(i+1,j)
    dy[:, :-1, :] = gc_dist(grid[:, 1:, :], grid[:, :-1, :], radius)        # dy(i,j): (i,j)->(i,j+1)
    dx, dy = cs.fill_corners_dgrid(dx, dy, npx, npy, 1.0)

    # ---- agrid (:690-713)
    agrid = Z(2)
    ec = grid3[:, :-1, :-1] + grid3[:, :-1, 1:] + grid3[:, 1:, :-1] + grid3[:, 1:, 1:]
    agrid[:, :-1, :-1, :] = xyz2ll(ec)
    agrid[..., 0] = cs.fill_corners_agrid_scalar(np.ascontiguousarray(agrid[..., 0]), npx, npy, "x")
    agrid[..., 1] = cs.fill_corners_agrid_scalar(np.ascontiguousarray(agrid[..., 1]), npx, npy, "y")

    # ---- dxa, dya (:715-727)
    dxa = Z(); dya = Z()
    p1 = mid_ll(grid[:, :-1, :-1], grid[:, 1:, :-1]); p2 = mid_ll(grid[:, :-1, 1:], grid[:, 1:, 1:])
    dxa[:, :-1, :-1] = gc_dist(p2, p1, radius)
    p1 = mid_ll(grid[:, :-1, :-1], grid[:, :-1, 1:]); p2 = mid_ll(grid[:, 1:, :-1], grid[:, 1:, 1:])
    dya[:, :-1, :-1] = gc_dist(p2, p1, radius)
    dxa, dya = cs.fill_corners_agrid_pair(dxa, dya, npx, npy, 1.0)

    # ---- dxc, dyc (:734-752, edge overrides :766-826, exchange+corner fill :865-867)
    dxc = Z(); dyc = Z()
    isd, ied = 1 - ng, N + ng
    dxc[:, R(isd, ied), R(isd + 1, ied)] = gc_dist(agrid[:, R(isd, ied), R(isd + 1, ied)],
                                                   agrid[:, R(isd, ied), R(isd, ied - 1)], radius)
    dxc[:, R(isd, ied), isd + o] = dxc[:, R(isd, ied), isd + 1 + o]
    dxc[:, R(isd, ied), ied + 1 + o] = dxc[:, R(isd, ied), ied + o]
    dyc[:, R(isd + 1, ied), R(isd, ied)] = gc_dist(agrid[:, R(isd + 1, ied), R(isd, ied)],
                                                   agrid[:, R(isd, ied - 1), R(isd, ied)], radius)
    dyc[:, isd + o, R(isd, ied)] = dyc[:, isd + 1 + o, R(isd, ied)]
    dyc[:, ied + 1 + o, R(isd, ied)] = dyc[:, ied + o, R(isd, ied)]
    js = R(1, N)
    # west / east
    pm = mid_ll(grid[:, R(1, N), 1 + o], grid[:, R(2, N + 1), 1 + o])
    dxc[:, js, 1 + o] = 2.0 * gc_dist(pm, agrid[:, js, 1 + o], radius)
    pm = mid_ll(grid[:, R(1, N), npx + o], grid[:, R(2, N + 1), npx + o])
    dxc[:, js, npx + o] = 2.0 * gc_dist(agrid[:, js, npx - 1 + o], pm, radius)
    # south / north
    pm = mid_ll(grid[:, 1 + o, R(1, N)], grid[:, 1 + o, R(2, N + 1)])
    dyc[:, 1 + o, js] = 2.0 * gc_dist(pm, agrid[:, 1 + o, js], radius)
    pm = mid_ll(grid[:, npy + o, R(1, N)], grid[:, npy + o, R(2, N + 1)])
    dyc[:, npy + o, js] = 2.0 * gc_dist(agrid[:, npy - 1 + o, js], pm, radius)
    dxc, dyc = halo.cgrid(dxc, dyc, scalar_pair=True)
    dxc, dyc = cs.fill_corners_cgrid(dxc, dyc, npx, npy, 1.0)

    # ---- area (grid_area :1987-2008), exchange, ghost corners (:869, :898)
    area = Z()
    A = R(1, N); B = R(2, N + 1)
    area[:, A, A] = get_area(grid[:, A, A], grid[:, B, A], grid[:, A, B], grid[:, B, B], radius)
    area = halo.scalar(area)
    area = cs.fill_ghost(area, npx, npy, -BIG)

    # ---- area_c (grid_area :2051-2142; edge/corner overrides init_grid :766-860)
    area_c = Z()
    C = R(1, N + 1); Cm = R(0, N)
    area_c[:, C, C] = get_area(agrid[:, Cm, Cm], agrid[:, C, Cm], agrid[:, Cm, C], agrid[:, C, C], radius)
    jj = R(1, N + 1); jm = R(0, N); jp = R(2, N + 2)
    # west edge i=1
    pa = mid_ll(grid[:, jm, 1 + o], grid[:, jj, 1 + o]); pd = mid_ll(grid[:, jj, 1 + o], grid[:, jp, 1 + o])
    area_c[:, jj, 1 + o] = 2.0 * get_area(pa, pd, agrid[:, jm, 1 + o], agrid[:, jj, 1 + o], radius)
    # east edge i=npx
    pb = mid_ll(grid[:, jm, npx + o], grid[:, jj, npx + o]); pc = mid_ll(grid[:, jj, npx + o], grid[:, jp, npx + o])
    area_c[:, jj, npx + o] = 2.0 * get_area(agrid[:, jm, npx - 1 + o], agrid[:, jj, npx - 1 + o], pb, pc, radius)
    # south edge j=1
    pa = mid_ll(grid[:, 1 + o, jm], grid[:, 1 + o, jj]); pb = mid_ll(grid[:, 1 + o, jj], grid[:, 1 + o, jp])
    area_c[:, 1 + o, jj] = 2.0 * get_area(pa, agrid[:, 1 + o, jm], pb, agrid[:, 1 + o, jj], radius)
    # north edge j=npy
    pc = mid_ll(grid[:, npy + o, jj], grid[:, npy + o, jp]); pd = mid_ll(grid[:, npy + o, jm], grid[:, npy + o, jj])
    area_c[:, npy + o, jj] = 2.0 * get_area(agrid[:, npy - 1 + o, jm], pd, agrid[:, npy - 1 + o, jj], pc, radius)
    # four cube corners: 3 x third-cell
    def G(i, j): return grid[:, j + o, i + o]
    def Ag(i, j): return agrid[:, j + o, i + o]
    i, j = 1, 1
    area_c[:, j + o, i + o] = 3.0 * get_area(G(i, j), mid_ll(G(i, j), G(i, j + 1)), mid_ll(G(i, j), G(i + 1, j)), Ag(i, j), radius)
    i, j = npx, 1
    area_c[:, j + o, i + o] = 3.0 * get_area(mid_ll(G(i - 1, j), G(i, j)), Ag(i - 1, j), G(i, j), mid_ll(G(i, j), G(i, j + 1)), radius)
    i, j = npx, npy
    area_c[:, j + o, i + o] = 3.0 * get_area(Ag(i - 1, j - 1), mid_ll(G(i - 1, j), G(i, j)), mid_ll(G(i, j - 1), G(i, j)), G(i, j), radius)
    i, j = 1, npy
    area_c[:, j + o, i + o] = 3.0 * get_area(mid_ll(G(i, j - 1), G(i, j)), G(i, j), Ag(i, j - 1), mid_ll(G(i, j), G(i + 1, j)), radius)
    area_c = halo.corner(area_c)
    area_c = cs.fill_corners_bgrid(area_c, npx, npy, "x")

    def recip(a):
        return np.where(a != 0.0, 1.0 / np.where(a != 0.0, a, 1.0), 0.0)

    M.update(grid=grid, agrid=agrid, dx=dx, dy=dy, dxa=dxa, dya=dya, dxc=dxc, dyc=dyc, area=area, area_c=area_c,
             rdx=recip(dx), rdy=recip(dy), rdxa=recip(dxa), rdya=recip(dya), rdxc=recip(dxc), rdyc=recip(dyc),
             rarea=recip(area), rarea_c=recip(area_c))

    # ---- grid_utils_init: cos_sg / sin_sg (model/fv_grid_utils_nlm.F90:318-392)
    cos_sg = np.zeros((6, NXP, NXP, 10)); sin_sg = np.zeros((6, NXP, NXP, 10))
    g00 = grid3[:, :-1, :-1]; g10 = grid3[:, :-1, 1:]; g01 = grid3[:, 1:, :-1]; g11 = grid3[:, 1:, 1:]
    p3 = ll2xyz(agrid[:, :-1, :-1])
    S = (slice(None), slice(0, NXP - 1), slice(0, NXP - 1))
    cos_sg[S + (6,)] = cos_angle(g00, g10, g01)
    cos_sg[S + (7,)] = -cos_angle(g10, g00, g11)
    cos_sg[S + (8,)] = cos_angle(g11, g10, g01)
    cos_sg[S + (9,)] = -cos_angle(g01, g00, g11)
    cos_sg[S + (1,)] = cos_angle(mid3(g00, g01), p3, g01)
    cos_sg[S + (2,)] = cos_angle(mid3(g00, g10), g10, p3)
    cos_sg[S + (3,)] = cos_angle(mid3(g10, g11), p3, g10)
    cos_sg[S + (4,)] = cos_angle(mid3(g01, g11), g01, p3)
    # ec1, ec2 (get_center_vect :1722) -> cos_sg(5)
    pc = unit(g00 + g10 + g01 + g11)
    ec1 = unit(cross(pc, cross(mid3(g10, g11), mid3(g00, g01))))
    ec2 = unit(cross(pc, cross(mid3(g01, g11), mid3(g00, g10))))
    cos_sg[S + (5,)] = (ec1 * ec2).sum(-1)
    sin_sg = np.minimum(1.0, np.sqrt(np.maximum(0.0, 1.0 - cos_sg ** 2)))
    # first corner patch (sin only, :362-392) -- it precedes cosa_u/sina_u etc.
    sin0 = sin_sg.copy(); cos0 = cos_sg.copy()
    def SG0(i, j, k): return sin_sg[:, j + (ng - 1), i + (ng - 1), k]
    for i in (-2, -1, 0):
        sin0[:, i + o, 0 + o, 3] = SG0(i, 1, 2); sin0[:, 0 + o, i + o, 4] = SG0(1, i, 1)      # sw
        sin0[:, npy + o, i + o, 2] = SG0(1, npx + i, 1)                                        # nw (reference quirk: npx+i)
        sin0[:, i + o, npx + o, 1] = SG0(npx - i, 1, 2)                                        # se
    for i in range(npy, npy + 3):
        sin0[:, i + o, 0 + o, 3] = SG0(npy - i, npy - 1, 4)                                    # nw
        sin0[:, 0 + o, i + o, 4] = SG0(npx - 1, npx - i, 3)                                    # se
        sin0[:, i + o, npx + o, 1] = SG0(i, npy - 1, 4); sin0[:, npy + o, i + o, 2] = SG0(npx - 1, i, 3)   # ne
    # ghost corners set to tiny/big, then the corner-adjacent patches (:571-627)
    for k in range(1, 10):
        sin_sg[..., k] = cs.fill_ghost(np.ascontiguousarray(sin_sg[..., k]), npx, npy, TINY)
        cos_sg[..., k] = cs.fill_ghost(np.ascontiguousarray(cos_sg[..., k]), npx, npy, BIG)
    src_s, src_c = sin_sg.copy(), cos_sg.copy()
    def SG(a, i, j, k): return a[:, j + o, i + o, k]
    for a, s in ((sin_sg, src_s), (cos_sg, src_c)):
        for i in (0, -1, -2):                       # sw
            a[:, i + o, 0 + o, 3] = SG(s, i, 1, 2)
            a[:, 0 + o, i + o, 4] = SG(s, 1, i, 1)
        for i in range(npy, npy + 3):               # nw
            a[:, i + o, 0 + o, 3] = SG(s, npy - i, npy - 1, 4)
        for i in (0, -1, -2):
            a[:, npy + o, i + o, 2] = SG(s, 1, npy - i, 1)
        for j in (0, -1, -2):                       # se
            a[:, j + o, npx + o, 1] = SG(s, npx - j, 1, 2)
        for i in range(npx, npx + 3):
            a[:, 0 + o, i + o, 4] = SG(s, npx - 1, npx - i, 3)
        for i in (0, 1, 2):                         # ne
            a[:, npy + i + o, npx + o, 1] = SG(s, npx + i, npy - 1, 4)
            a[:, npy + o, npx + i + o, 2] = SG(s, npx - 1, npy + i, 3)
    M["cos_sg"] = cos_sg; M["sin_sg"] = sin_sg

    # cosa/sina etc. are built from sin0/cos0 = the sg arrays BEFORE the ghost fill
    # (reference order: :487-557 precede :571).
    cosa = np.full((6, NXP, NXP), BIG); sina = np.full((6, NXP, NXP), BIG)
    cosa_u = np.full((6, NXP, NXP), BIG); sina_u = np.full((6, NXP, NXP), BIG); rsin_u = np.full((6, NXP, NXP), BIG)
    cosa_v = np.full((6, NXP, NXP), BIG); sina_v = np.full((6, NXP, NXP), BIG); rsin_v = np.full((6, NXP, NXP), BIG)
    rsina = np.full((6, NXP, NXP), BIG)
    Cc = R(1, N + 1); Cm1 = R(0, N)
    cosa[:, Cc, Cc] = 0.5 * (cos0[:, Cm1, Cm1, 8] + cos0[:, Cc, Cc, 6])
    sina[:, Cc, Cc] = 0.5 * (sin0[:, Cm1, Cm1, 8] + sin0[:, Cc, Cc, 6])
    D = R(isd, ied); Dp = R(isd + 1, ied); Dm = R(isd, ied - 1)
    cosa_u[:, D, Dp] = 0.5 * (cos0[:, D, Dm, 3] + cos0[:, D, Dp, 1])
    sina_u[:, D, Dp] = 0.5 * (sin0[:, D, Dm, 3] + sin0[:, D, Dp, 1])
    rsin_u[:, D, Dp] = 1.0 / np.maximum(TINY, sina_u[:, D, Dp] ** 2)
    cosa_v[:, Dp, D] = 0.5 * (cos0[:, Dm, D, 4] + cos0[:, Dp, D, 2])
    sina_v[:, Dp, D] = 0.5 * (sin0[:, Dm, D, 4] + sin0[:, Dp, D, 2])
    rsin_v[:, Dp, D] = 1.0 / np.maximum(TINY, sina_v[:, Dp, D] ** 2)
    cosa_s = np.full((6, NXP, NXP), BIG); rsin2 = np.full((6, NXP, NXP), BIG)
    cosa_s[:, D, D] = cos0[:, D, D, 5]
    rsin2[:, D, D] = 1.0 / np.maximum(TINY, sin0[:, D, D, 5] ** 2)
    cosa_s = cs.fill_ghost(cosa_s, npx, npy, BIG)
    rs = 1.0 / np.maximum(TINY, sina[:, Cc, Cc] ** 2)
    II, JJ = np.meshgrid(np.arange(1, N + 2), np.arange(1, N + 2))    # [j, i]
    edge = (II == 1) | (II == npx) | (JJ == 1) | (JJ == npy)
    rsina[:, Cc, Cc] = np.where(edge[None], BIG, rs)
    for i in (1, npx):
        s = sina_u[:, D, i + o]
        rsin_u[:, D, i + o] = 1.0 / (np.sign(s) * np.maximum(TINY, np.abs(s)))
    for j in (1, npy):
        s = sina_v[:, j + o, D]
        rsin_v[:, j + o, D] = 1.0 / (np.sign(s) * np.maximum(TINY, np.abs(s)))
    M.update(cosa=cosa, sina=sina, rsina=rsina, cosa_u=cosa_u, sina_u=sina_u, rsin_u=rsin_u,
             cosa_v=cosa_v, sina_v=sina_v, rsin_v=rsin_v, cosa_s=cosa_s, rsin2=rsin2)

    # ---- divg_u/v, del6_u/v (:700-726) then SCALAR_PAIR CGRID exchange (:743-746)
    divg_u = Z(); del6_u = Z(); divg_v = Z(); del6_v = Z()
    E = R(isd, ied + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        divg_u[:, E, D] = sina_v[:, E, D] * dyc[:, E, D] / dx[:, E, D]
        del6_u[:, E, D] = sina_v[:, E, D] * dx[:, E, D] / dyc[:, E, D]
        for j in (1, npy):
            f = 0.5 * (sin_sg[:, j + o, D, 2] + sin_sg[:, j - 1 + o, D, 4])
            divg_u[:, j + o, D] = f * dyc[:, j + o, D] / dx[:, j + o, D]
            del6_u[:, j + o, D] = f * dx[:, j + o, D] / dyc[:, j + o, D]
        divg_v[:, D, E] = sina_u[:, D, E] * dxc[:, D, E] / dy[:, D, E]
        del6_v[:, D, E] = sina_u[:, D, E] * dy[:, D, E] / dxc[:, D, E]
        for i in (1, npx):
            f = 0.5 * (sin_sg[:, D, i + o, 1] + sin_sg[:, D, i - 1 + o, 3])
            divg_v[:, D, i + o] = f * dxc[:, D, i + o] / dy[:, D, i + o]
            del6_v[:, D, i + o] = f * dy[:, D, i + o] / dxc[:, D, i + o]
    for a in (divg_u, del6_u, divg_v, del6_v):
        a[~np.isfinite(a)] = 0.0
    divg_v, divg_u = halo.cgrid(divg_v, divg_u, scalar_pair=True)
    del6_v, del6_u = halo.cgrid(del6_v, del6_u, scalar_pair=True)
    M.update(divg_u=divg_u, divg_v=divg_v, del6_u=del6_u, del6_v=del6_v)

    # ---- edge factors (edge_factors :1105, efactor_a2c_v :926); stored [6, NXP], index i -> i+o
    edge_w = np.full((6, NXP), BIG); edge_e = np.full((6, NXP), BIG)
    edge_s = np.full((6, NXP), BIG); edge_n = np.full((6, NXP), BIG)
    for (i, arr) in ((1, edge_w), (npx, edge_e)):
        py = mid_ll(agrid[:, R(1, N), i - 1 + o], agrid[:, R(1, N), i + o])        # j = 1..N
        gpt = grid[:, R(2, N), i + o]                                              # j = 2..N
        d1 = gc_dist(py[:, :-1], gpt); d2 = gc_dist(py[:, 1:], gpt)
        arr[:, R(2, N)] = d2 / (d1 + d2)
    for (j, arr) in ((1, edge_s), (npy, edge_n)):
        px = mid_ll(agrid[:, j - 1 + o, R(1, N)], agrid[:, j + o, R(1, N)])
        gpt = grid[:, j + o, R(2, N)]
        d1 = gc_dist(px[:, :-1], gpt); d2 = gc_dist(px[:, 1:], gpt)
        arr[:, R(2, N)] = d2 / (d1 + d2)
    M.update(edge_w=edge_w, edge_e=edge_e, edge_s=edge_s, edge_n=edge_n)

    im2 = (npx - 1) // 2
    ev_w = np.full((6, NXP), BIG); ev_e = np.full((6, NXP), BIG); ev_s = np.full((6, NXP), BIG); ev_n = np.full((6, NXP), BIG)
    for (i, arr) in ((1, ev_w), (npx, ev_e)):
        jr = R(-1, N + 2)                                                            # j = js-2..je+2
        py = mid_ll(agrid[:, jr, i - 1 + o], agrid[:, jr, i + o])
        p2 = mid_ll(grid[:, jr, i + o], grid[:, R(0, N + 3), i + o])
        for j in range(0, N + 2):                                                    # js-1..je+1
            a = j - (-1)                                                             # position in jr arrays
            if j <= im2:
                d1 = gc_dist(py[:, a], p2[:, a]); d2 = gc_dist(py[:, a + 1], p2[:, a])
            else:
                d2 = gc_dist(py[:, a - 1], p2[:, a]); d1 = gc_dist(py[:, a], p2[:, a])
            arr[:, j + o] = d1 / (d1 + d2)
        arr[:, 0 + o] = arr[:, 1 + o]
        arr[:, npy + o] = arr[:, N + o]
    for (j, arr) in ((1, ev_s), (npy, ev_n)):
        ir = R(-1, N + 2)
        px = mid_ll(agrid[:, j - 1 + o, ir], agrid[:, j + o, ir])
        p1 = mid_ll(grid[:, j + o, ir], grid[:, j + o, R(0, N + 3)])
        for i in range(0, N + 2):
            a = i + 1
            if i <= im2:
                d1 = gc_dist(px[:, a], p1[:, a]); d2 = gc_dist(px[:, a + 1], p1[:, a])
            else:
                d2 = gc_dist(px[:, a - 1], p1[:, a]); d1 = gc_dist(px[:, a], p1[:, a])
            arr[:, i + o] = d1 / (d1 + d2)
        arr[:, 0 + o] = arr[:, 1 + o]
        arr[:, npx + o] = arr[:, N + o]
    M.update(edge_vect_w=ev_w, edge_vect_e=ev_e, edge_vect_s=ev_s, edge_vect_n=ev_n)

    # ---- init_cubed_to_latlon (model/fv_grid_utils_nlm.F90:2248-2310): local cell-centre winds -> lon / lat components
    lon_c, lat_c = agrid[:, :-1, :-1, 0], agrid[:, :-1, :-1, 1]
    vlon = np.stack([-np.sin(lon_c), np.cos(lon_c), np.zeros_like(lon_c)], axis=-1)                       # unit_vect_latlon :2213
    vlat = np.stack([-np.sin(lat_c) * np.cos(lon_c), -np.sin(lat_c) * np.sin(lon_c), np.cos(lat_c)], axis=-1)
    z11 = (ec1 * vlon).sum(-1); z12 = (ec1 * vlat).sum(-1); z21 = (ec2 * vlon).sum(-1); z22 = (ec2 * vlat).sum(-1)
    s5 = sin_sg[:, :-1, :-1, 5]
    for nm, z, sgn in (("a11", z22, 0.5), ("a12", z12, -0.5), ("a21", z21, -0.5), ("a22", z11, 0.5)):
        a = np.zeros((6, NXP, NXP))
        a[:, :-1, :-1] = np.where(s5 > 1e-10, sgn * z / np.maximum(s5, 1e-10), 0.0)       # (ghost corners hold sin = tiny: unused)
        M[nm] = a

    # ---- Coriolis (fv3jedi_lm_dynamics_mod.F90:126-139, f_coriolis_angle = 0)
    M["fC"] = 2.0 * omega * np.sin(grid[..., 1])
    M["f0"] = 2.0 * omega * np.sin(agrid[..., 1])

    inner = area[:, R(1, N), R(1, N)]
    M["da_min"] = float(inner.min()); M["da_max"] = float(inner.max())
    ac = area_c[:, R(1, N), R(1, N)]                     # global_mx_c(area_c(is:ie,js:je)) :735
    M["da_min_c"] = float(ac.min()); M["da_max_c"] = float(ac.max())
    M["N"] = N; M["npx"] = npx; M["npy"] = npy; M["ng"] = ng
    return M
