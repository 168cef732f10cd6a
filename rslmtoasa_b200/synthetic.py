"""Synthetic clusters + Hamiltonian block sets in the reference's own data conventions.

Shared by the oracle, the CPU baseline, the tests and bench.py.  Nothing here is on the
product's compute path: it only manufactures *inputs* shaped like the ones `lattice`/`hamiltonian` hand to
`recursion` in the reference:

* ``nn(kk, ncols)`` int32, Fortran order: ``nn[i,0]`` = number of slots incl. the on-site slot (reference
  ``nn(i,1)``), ``nn[i,m]`` (m>=1) = 1-based neighbour site of slot m+1 or 0 if the neighbour is outside the
  cluster (reference `lattice.f90:1856-1860`, `2889-2892`).
* ``iz(kk)`` int32 1-based atom type (`lattice.f90`, member ``iz``).
* ``ee/eeo(18,18,nslot,ntype)``, ``hall/hallo(18,18,nslot,nmax)``, ``lsham/enim(18,18,ntype)`` complex128 Fortran
  order (`hamiltonian.f90:294-301`), with ``nslot = max(nn[:,0]) + 1`` exactly like the reference allocation.
"""
from __future__ import annotations

import dataclasses
import numpy as np

NB = 18  # spd x spin block size (2*(lmax+1)**2, lmax = 2)

BCC_DISP = np.array(
    [[0, 0, 0]]
    + [[sx, sy, sz] for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)]          # 8 NN, units of a/2
    + [[2, 0, 0], [-2, 0, 0], [0, 2, 0], [0, -2, 0], [0, 0, 2], [0, 0, -2]],        # 6 NNN
    dtype=np.int64)
FCC_DISP = np.array(
    [[0, 0, 0]]
    + [[sx, sy, 0] for sx in (1, -1) for sy in (1, -1)]
    + [[sx, 0, sz] for sx in (1, -1) for sz in (1, -1)]
    + [[0, sy, sz] for sy in (1, -1) for sz in (1, -1)]                              # 12 NN, units of a/2
    + [[2, 0, 0], [-2, 0, 0], [0, 2, 0], [0, -2, 0], [0, 0, 2], [0, 0, -2]],        # 6 NNN
    dtype=np.int64)


def _opposite_slots(disp: np.ndarray) -> np.ndarray:
    """slot index of -R for every slot R (0-based, slot 0 = on-site maps to itself)."""
    opp = np.empty(len(disp), dtype=np.int64)
    for m, d in enumerate(disp):
        opp[m] = int(np.where((disp == -d).all(axis=1))[0][0])
    return opp


@dataclasses.dataclass
class Lattice:
    """The members of the reference `lattice` type that the recursion reads (`lattice.f90:144-309`)."""
    kk: int
    nn: np.ndarray          # (kk, ncols) int32, Fortran order
    iz: np.ndarray          # (kk,) int32, 1-based types
    ntype: int
    nmax: int               # sites 1..nmax use the site-indexed `hall`
    irec: np.ndarray        # (nrec,) int32 1-based recursion sites
    cr: np.ndarray | None = None   # (3, kk) integer coordinates in units of a/2 (kept only for small clusters)
    disp: np.ndarray | None = None

    @property
    def nslot(self) -> int:
        return int(self.nn[:, 0].max()) + 1

    @property
    def ncols(self) -> int:
        return int(self.nn.shape[1])


@dataclasses.dataclass
class Hamiltonian:
    """The members of the reference `hamiltonian` type on the hot path (`hamiltonian.f90:43-113`)."""
    ee: np.ndarray
    lsham: np.ndarray
    hall: np.ndarray | None = None
    eeo: np.ndarray | None = None
    hallo: np.ndarray | None = None
    enim: np.ndarray | None = None
    hoh: bool = False
    v_a: np.ndarray | None = None
    v_b: np.ndarray | None = None


def _sphere_points(kind: str, r2: float) -> np.ndarray:
    """All lattice points with |r|^2 <= r2 (alat^2), as integer coords in units of a/2, centre first."""
    n = int(np.ceil(np.sqrt(r2))) + 1
    g = np.arange(-2 * n, 2 * n + 1, dtype=np.int64)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    if kind == "bcc":      # all even (corner) or all odd (centre)
        par = pts & 1
        keep = (par[:, 0] == par[:, 1]) & (par[:, 1] == par[:, 2])
    elif kind == "fcc":    # x+y+z even
        keep = (pts.sum(axis=1) & 1) == 0
    else:
        raise ValueError(kind)
    pts = pts[keep]
    d2 = (pts ** 2).sum(axis=1)
    pts = pts[d2 <= 4.0 * r2 + 1e-9]
    d2 = (pts ** 2).sum(axis=1)
    order = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0], d2))
    return pts[order]


def _nn_from_points(pts: np.ndarray, disp: np.ndarray) -> np.ndarray:
    kk = len(pts)
    off = int(np.abs(pts).max()) + 3
    span = 2 * off + 1
    key = lambda p: ((p[:, 0] + off) * span + (p[:, 1] + off)) * span + (p[:, 2] + off)
    keys = key(pts)
    order = np.argsort(keys)
    skeys = keys[order]
    nn = np.zeros((kk, len(disp)), dtype=np.int32, order="F")
    nn[:, 0] = len(disp)
    for m in range(1, len(disp)):
        q = key(pts + disp[m])
        pos = np.searchsorted(skeys, q)
        pos = np.clip(pos, 0, kk - 1)
        hit = skeys[pos] == q
        nn[:, m] = np.where(hit, order[pos] + 1, 0).astype(np.int32)
    return nn


def sphere_cluster(kind: str = "bcc", r2: float = 8.0, ntype: int = 1, nmax: int = 0,
                   type_rule: str = "single") -> Lattice:
    """Open-boundary spherical cluster cut (the reference's `bravais`/`cut`), site 1 = centre.

    type_rule: "single" (all sites type 1), "b2" (bcc sublattices = types 1/2; ntype>=2), "layer"
    (type = 1 + min(|z| layer, ntype-1), a slab-like layer typing).  Sites 1..nmax form the site-indexed
    ("impurity", `hall`) region; with ntype==3 and rule "b2" the centre site gets its own type 3.
    """
    disp = BCC_DISP if kind == "bcc" else FCC_DISP
    pts = _sphere_points(kind, r2)
    if len(pts) % 2 == 1 and len(pts) > 1:   # the reference makes kk even (`lattice.f90:1091`)
        pts = pts[:-1]
    nn = _nn_from_points(pts, disp)
    kk = len(pts)
    iz = np.ones(kk, dtype=np.int32)
    if type_rule == "b2":
        iz = (1 + (pts[:, 0] & 1)).astype(np.int32)
        if ntype >= 3:
            iz[0] = 3
    elif type_rule == "layer":
        iz = (1 + np.minimum(np.abs(pts[:, 2]) // (1 if kind == "bcc" else 1), ntype - 1)).astype(np.int32)
    return Lattice(kk=kk, nn=nn, iz=iz, ntype=int(max(ntype, iz.max())), nmax=nmax,
                   irec=np.array([1], dtype=np.int32), cr=pts.T.copy(), disp=disp)


def periodic_bcc_positions(nx: int, ny: int, nz: int) -> np.ndarray:
    """(3, kk) coordinates in units of a/2 of periodic_bcc's sites (the reference's lattice%cr up to the factor alat/2)."""
    idx = np.arange(nx * ny * nz, dtype=np.int64)
    cell = np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.float64) * 2.0
    cr = np.empty((3, 2 * nx * ny * nz), dtype=np.float64, order="F")
    cr[:, 0::2] = cell
    cr[:, 1::2] = cell + 1.0
    return cr


def periodic_bcc(nx: int, ny: int, nz: int, ntype: int = 1) -> Lattice:
    """bcc with periodic boundaries, nx*ny*nz cubic cells x 2 atoms, generated analytically (O(kk)).

    Site numbering: ((z*ny + y)*nx + x)*2 + basis + 1 (x fastest), the lattice-loop order of the reference's
    PBC builder; every slot is populated (no zeros).
    """
    kk = 2 * nx * ny * nz
    idx = np.arange(nx * ny * nz, dtype=np.int64)
    x = idx % nx
    y = (idx // nx) % ny
    z = idx // (nx * ny)
    nn = np.zeros((kk, len(BCC_DISP)), dtype=np.int32, order="F")
    nn[:, 0] = len(BCC_DISP)

    def site(xc, yc, zc, b):
        return (((zc % nz) * ny + (yc % ny)) * nx + (xc % nx)) * 2 + b + 1

    for m in range(1, len(BCC_DISP)):
        dx, dy, dz = (int(v) for v in BCC_DISP[m])
        for b in (0, 1):
            # positions in units of a/2: corner (2x,2y,2z), centre (2x+1, 2y+1, 2z+1)
            px, py, pz = 2 * x + b + dx, 2 * y + b + dy, 2 * z + b + dz
            nb = px & 1
            nn[b::2, m] = site((px - nb) // 2, (py - nb) // 2, (pz - nb) // 2, nb).astype(np.int32)
    iz = np.ones(kk, dtype=np.int32)
    if ntype == 2:
        iz[1::2] = 2
    return Lattice(kk=kk, nn=nn, iz=iz, ntype=ntype, nmax=0, irec=np.array([1], dtype=np.int32),
                   cr=None, disp=BCC_DISP)


def _rand_block(rng, sigma):
    return (rng.normal(0.0, sigma, (NB, NB)) + 1j * rng.normal(0.0, sigma, (NB, NB))) / np.sqrt(2.0)


def make_hamiltonian(lat: Lattice, seed: int = 20260101, sigma: float = 0.05, hoh: bool = False,
                     spin_orbit: bool = True, velocity: bool = False) -> Hamiltonian:
    """Random Hermitian-consistent block set: H_slot(m) of type t  <->  H_slot(-m)^H of the neighbour's type.

    Hermitian consistency needs the neighbour's type to be a function of (type, slot); that holds for the
    "single"/"b2" rules and for PBC lattices.  For other typings the blocks are still deterministic but the
    assembled operator is only approximately Hermitian (the recursion formulas never assume hermiticity).
    Spectrum: on-site diag in [-0.4, 0.4] Ry, hoppings Gaussian sigma -> ||H|| < ~1.2 Ry for sigma = 0.05.
    """
    rng = np.random.default_rng(seed)
    disp = lat.disp
    nslot_used = len(disp)
    nslot = lat.nslot
    ntype = lat.ntype
    opp = _opposite_slots(disp)
    ee = np.zeros((NB, NB, nslot, ntype), dtype=np.complex128, order="F")
    # type of the neighbour in slot m for a site of type t (taken from the first site of that type that has it)
    nbr_type = np.zeros((ntype, nslot_used), dtype=np.int64)
    for t in range(ntype):
        sites = np.where(lat.iz == t + 1)[0]
        for m in range(1, nslot_used):
            nb = lat.nn[sites, m]
            nb = nb[nb > 0]
            nbr_type[t, m] = lat.iz[nb[0] - 1] - 1 if len(nb) else t
    for t in range(ntype):
        on = _rand_block(rng, sigma)
        on = 0.5 * (on + on.conj().T)
        on[np.diag_indices(NB)] = rng.uniform(-0.4, 0.4, NB)
        ee[:, :, 0, t] = on
    for t in range(ntype):
        for m in range(1, nslot_used):
            t2, m2 = int(nbr_type[t, m]), int(opp[m])
            if (t2, m2) < (t, m):
                ee[:, :, m, t] = ee[:, :, m2, t2].conj().T
            else:
                ee[:, :, m, t] = _rand_block(rng, sigma)
    if not spin_orbit:   # collinear: spin-block-diagonal (structural zeros the reference still multiplies)
        ee[:9, 9:] = 0.0
        ee[9:, :9] = 0.0
    lsham = np.zeros((NB, NB, ntype), dtype=np.complex128, order="F")
    for t in range(ntype):
        ls = _rand_block(rng, 0.01) if spin_orbit else np.zeros((NB, NB), dtype=np.complex128)
        lsham[:, :, t] = 0.5 * (ls + ls.conj().T)
    ham = Hamiltonian(ee=ee, lsham=lsham, hoh=hoh)
    if lat.nmax > 0:
        hall = np.zeros((NB, NB, nslot, lat.nmax), dtype=np.complex128, order="F")
        for i in range(lat.nmax):
            hall[:, :, :, i] = ee[:, :, :, lat.iz[i] - 1]
            d = _rand_block(rng, 0.02)
            hall[:, :, 0, i] += 0.5 * (d + d.conj().T)
        for i in range(lat.nmax):
            for m in range(1, nslot_used):
                j = int(lat.nn[i, m]) - 1
                if 0 <= j < lat.nmax and i < j:
                    d = _rand_block(rng, 0.01)
                    hall[:, :, m, i] += d
                    hall[:, :, int(opp[m]), j] += d.conj().T
        ham.hall = hall
    if hoh:
        # eeo = ee * obarm(type of neighbour) (`hamiltonian.f90:1597-1606`), enim per type; obarm small & Hermitian
        obarm = np.zeros((NB, NB, ntype), dtype=np.complex128)
        for t in range(ntype):
            o = _rand_block(rng, 0.02)
            obarm[:, :, t] = 0.5 * (o + o.conj().T) + np.diag(rng.uniform(-0.1, 0.1, NB))
        enim = np.zeros((NB, NB, ntype), dtype=np.complex128, order="F")
        for t in range(ntype):
            enim[:, :, t] = np.diag(rng.uniform(-0.05, 0.05, NB))
        eeo = np.zeros_like(ee)
        for t in range(ntype):
            for m in range(nslot_used):
                t2 = t if m == 0 else int(nbr_type[t, m])
                eeo[:, :, m, t] = ee[:, :, m, t] @ obarm[:, :, t2]
        ham.eeo, ham.enim = eeo, enim
        if lat.nmax > 0:
            hallo = np.zeros_like(ham.hall)
            for i in range(lat.nmax):
                for m in range(nslot_used):
                    j = i if m == 0 else int(lat.nn[i, m]) - 1
                    if j >= 0:
                        hallo[:, :, m, i] = ham.hall[:, :, m, i] @ obarm[:, :, lat.iz[j] - 1]
            ham.hallo = hallo
    if velocity:
        # v_m = -i (d . R_m) H_m (`hamiltonian.f90:1348`); direction x for v_a, y for v_b
        v_a = np.zeros_like(ee)
        v_b = np.zeros_like(ee)
        for m in range(nslot_used):
            v_a[:, :, m, :] = -1j * 0.5 * disp[m, 0] * ee[:, :, m, :]
            v_b[:, :, m, :] = -1j * 0.5 * disp[m, 1] * ee[:, :, m, :]
        ham.v_a, ham.v_b = v_a, v_b
    return ham


def random_phases(kk: int, nvec: int, seed: int = 20260104) -> np.ndarray:
    """u ~ U(0,1) per (site, vector); the KPM start block is exp(2 pi i u_k) I / sqrt(kk) (`recursion.f90:1135-1142`).

    The reference calls `random_seed()` without arguments (non-repeatable); the ABI therefore takes the phases
    from the host.  Counter-based Philox so that any rank can generate its own shard."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    return np.asfortranarray(rng.random((kk, nvec)))


def partition(rank: int, nprocs: int, n: int) -> tuple[int, int]:
    """`get_mpi_variables` (`mpi.f90:32-58`): 1-based inclusive (start_atom, end_atom) of this rank."""
    per = n // nprocs
    rem = n % nprocs
    if rank < rem:
        per += 1
        start = rank * per + 1
    else:
        start = rank * per + rem + 1
    return start, start + per - 1
