"""Host-side mirror of the reference's `type recursion` (source/recursion.f90:41-116) on top of the C ABI.

Same procedure names, the same result members with the same shapes and index conventions, and the same error
behaviour (a fatal condition raises instead of `g_logger%fatal`).  Everything numerical happens inside
librsrec.so on the GPU; this module only marshals the reference's arrays.

    rec = Recursion(hamiltonian, lattice, control, energy)      # recursion(hamiltonian_obj, energy_obj), 132-143
    rec.recur_b()            -> rec.a_b, rec.b2_b (18,18,lld,nrec_local), rec.a, rec.b2      (1807-1866)
    rec.recur_b_ij()         -> rec.a_b, rec.b2_b (18,18,lld,4*njij_local)                   (1655-1737)
    rec.recur()              -> rec.a, rec.b2 (lld,18,nrec_local,1)                          (3485-3532)
    rec.chebyshev_recur()    -> rec.mu_n (18,18,2*lld+2,nrec_local)                          (3057-3130)
    rec.chebyshev_recur_ij() -> rec.mu_n (18,18,2*lld+2,4*njij_local)                        (2376-2487)
    rec.compute_moments_stochastic() -> rec.mu_nm_stochastic (18,18,M,M,ntype|nvec)          (979-1234)
    rec.zsqr()               -> rec.b2_b <- sqrt                                             (1980-2023)
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import numpy as np

from . import _lib
from .synthetic import partition

NB = 18
ONE_OVER_SQRT_TWO = 1.0 / np.sqrt(2.0)


@dataclasses.dataclass
class Control:
    """The `&control` entries the recursion reads (reference control.f90:356-384)."""
    lld: int = 16
    recur: str = "block"            # lanczos | block | chebyshev
    cond_ll: int = 200
    cond_calctype: str = "per_type"  # per_type | random_vec
    random_vec_num: int = 1


@dataclasses.dataclass
class Energy:
    """`&energy energy_min, energy_max` (reference energy.f90); defines the Chebyshev scale a and shift b."""
    energy_min: float = -1.5
    energy_max: float = 1.5
    channels_ldos: int = 2500
    fermi: float = 0.0
    ene: object = None
    edel: float = 0.0
    nv1: int = 0
    ik1: int = 0
    fix_fermi: bool = False

    def e_mesh(self):
        """energy%e_mesh (energy.f90:175-208): channels_ldos+10 points from energy_min with a step that hits fermi;
        nv1 = ik1 = the odd channel count."""
        if self.channels_ldos % 2 == 0:
            self.nv1 = self.channels_ldos + 1
        else:
            self.nv1 = self.channels_ldos
            self.channels_ldos -= 1
        self.ik1 = self.nv1
        edel = (self.energy_max - self.energy_min) / self.channels_ldos
        r = (self.fermi - self.energy_min) / edel
        edel = (self.fermi - self.energy_min) / (np.sign(r) * np.floor(abs(r) + 0.5))  # nint
        self.edel = float(edel)
        self.ene = self.energy_min + edel * np.arange(self.channels_ldos + 10, dtype=np.float64)
        return self.ene

    def scale_shift(self):
        # recursion.f90:3078-3079
        return (self.energy_max - self.energy_min) / (2 - 0.3), (self.energy_max + self.energy_min) / 2


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fc(a):
    return None if a is None else np.asfortranarray(a, dtype=np.complex128)


def orbital_tail(mu_n_orb, nstart, en, path=None):
    """Tail of chebyshev_orbital_mod (recursion.f90:3009-3049), host side like in the reference: the moments are divided by the
    number of start sites (the reference loops over all kk sites and divides by kk), weighted with the Jackson kernel
    (math.f90:1641-1655), doubled for n >= 2 and summed to g0(E) = sum_i mu_i Im(-i e^{-i (i-1) acos w(E)}) / sqrt(a^2 - (E-b)^2)
    on the mesh en%ene; lzi(E) = Re tr g0(E); lz(E) = its T = 0 Fermi-weighted Simpson integral up to E (simpson_f,
    math.f90:1600-1632, fermi = .true.).  Returns (rows, lz, lzi): rows = the three columns of the reference's `fort.50`
    (E - E_F, -lz/pi, -lzi/pi) and writes them in its format ('(3es16.6)') when `path` is given.  The reference reads one
    element past the mesh in simpson_f when channels_ldos is even (nv1 + 10 > size(ene)); that term is taken as zero."""
    mu = np.array(mu_n_orb, dtype=np.complex128)
    lld = mu.shape[2]
    if en.ene is None:
        en.e_mesh()
    ene = np.asarray(en.ene, dtype=np.float64)
    nv = len(ene)
    a, b = en.scale_shift()
    n = float(lld)
    theta = np.pi * np.arange(lld) / (n + 1.0)
    kernel = ((n - np.arange(lld) + 1.0) * np.cos(theta) + np.sin(theta) / np.tan(np.pi / (n + 1.0))) / (n + 1.0)
    mu = mu / float(nstart)
    mu = mu * kernel[None, None, :]
    mu[:, :, 1:] *= 2.0
    w = (ene - b) / a
    ang = np.arange(lld)[:, None] * np.arccos(w)[None, :]                      # (i-1) acos(wscale(ie))
    fac = np.imag(-1j * np.exp(-1j * ang))                                     # aimag(exp_factor)
    tr = np.einsum("lli->i", mu)                                               # trace of every moment
    lzi = np.real(tr @ fac) / np.sqrt(a * a - (ene - b) ** 2)                  # rtrace(g0(:,:,ie))
    # simpson_f(lz, ene, ene(ie), nv1, lzi, fermi=.true., T=0): panels I = 2, 4, ..., nv1+9 (1-based), weights 1,4,1
    npts = en.nv1
    wgt = np.zeros(nv + 1)
    for i1 in range(2, npts + 10, 2):
        for off, c in ((-1, 1.0), (0, 4.0), (1, 1.0)):
            k = i1 + off - 1
            if k < nv:
                wgt[k] += c
    wgt = wgt[:nv]
    h = ene[1] - ene[0]
    lz = np.zeros(nv)
    with np.errstate(over="ignore"):
        for ie in range(nv):
            f = 1.0 / (np.exp((ene - ene[ie]) / 1.0e-15) + 1.0)                # fermifun (math.f90:994-1000), kBT = kB*0 + 1e-15
            lz[ie] = h * np.sum(wgt * lzi * f) / 3.0
    out = np.stack([ene - en.fermi, -lz / np.pi, -lzi / np.pi], axis=1)
    if path is not None:
        with open(path, "w") as fh:
            for row in out:
                fh.write("".join("%16.6E" % v for v in row) + "\n")
    return out, lz, lzi



class Recursion:
    def __init__(self, hamiltonian, lattice, control: Control | None = None, energy: Energy | None = None,
                 device: int = 0, rank: int = 0, numprocs: int = 1, ijpair=None, atlist=None, phases=None):
        self.hamiltonian, self.lattice = hamiltonian, lattice
        self.control = control or Control()
        self.en = energy or Energy()
        self.rank, self.numprocs = rank, numprocs
        self.ijpair = None if ijpair is None else np.asarray(ijpair, dtype=np.int32).reshape(-1, 2)
        self.atlist = atlist
        self.phases = phases          # (kk, random_vec_num): the host supplies the random numbers
        self.a = self.b2 = self.a_b = self.b2_b = self.mu_n = self.mu_nm_stochastic = None
        L = _lib.load()
        self._L = L
        self._h = C.c_void_p()
        _lib.check(L.rsrec_create(C.byref(self._h), device, lattice.kk, lattice.ncols, lattice.nslot, lattice.ntype,
                                  lattice.nmax))
        self.upload()

    # -- data export of the lattice / hamiltonian builders ------------------------------------------------
    def upload(self):
        """Re-export nn/iz and the block sets (what `self%run_recursion` does every SCF iteration, self.f90:777-797)."""
        L, lat, ham = self._L, self.lattice, self.hamiltonian
        nn = np.asfortranarray(lat.nn, dtype=np.int32)
        iz = np.ascontiguousarray(lat.iz, dtype=np.int32)
        _lib.check(L.rsrec_set_lattice(self._h, _p(nn), _p(iz)))
        if getattr(lat, "cr", None) is not None:      # lattice%cr: work ordering only (L2 locality)
            cr = np.asfortranarray(lat.cr, dtype=np.float64)
            _lib.check(L.rsrec_set_positions(self._h, _p(cr)))
        arrs = [_fc(getattr(ham, k, None)) for k in ("ee", "eeo", "hall", "hallo", "lsham", "enim")]
        _lib.check(L.rsrec_set_hamiltonian(self._h, *[_p(x) for x in arrs], int(bool(ham.hoh))))
        if getattr(ham, "v_a", None) is not None:
            for slot in ("a", "b"):
                v, vo = _fc(getattr(ham, "v_" + slot)), _fc(getattr(ham, "vo_" + slot, None))
                _lib.check(L.rsrec_set_operator(self._h, ord(slot), _p(v), _p(vo)))

    def build_hamiltonian(self, hhh, jt, it, pot, mom, lsham, hoh=False, download=True):
        """Device-side build_bulkham / build_locham (hamiltonian.f90:1553-1667): assembles the block sets on the GPU
        from structure-constant blocks and potential parameters (see include/rsrec.h).  `pot` is a dict with the
        (9,ntype) arrays wx0 wx1 cx0 cx1 cex0 cex1 obx0 obx1 and the (9,2,ntype) arrays cx, cex.  With download=True the
        reference's arrays come back as a dict (ee, eeo, hall, hallo, enim, obarm)."""
        lat = self.lattice
        nt, ns, nl = lat.ntype, lat.nslot, lat.nmax
        packed = np.zeros((9, 12, nt), np.complex128, order="F")
        for k, name in enumerate(("wx0", "wx1", "cx0", "cx1", "cex0", "cex1", "obx0", "obx1")):
            packed[:, k, :] = pot[name]
        packed[:, 8, :], packed[:, 9, :] = pot["cx"][:, 0, :], pot["cx"][:, 1, :]
        packed[:, 10, :], packed[:, 11, :] = pot["cex"][:, 0, :], pot["cex"][:, 1, :]
        hhh = np.asfortranarray(hhh, dtype=np.float64)
        jt = np.asfortranarray(jt, dtype=np.int32)
        it = np.ascontiguousarray(it, dtype=np.int32)
        mom = np.asfortranarray(mom, dtype=np.float64)
        ls = _fc(lsham)
        out = {}
        if download:
            out = {"ee": np.zeros((NB, NB, ns, nt), np.complex128, order="F"),
                   "eeo": np.zeros((NB, NB, ns, nt), np.complex128, order="F") if hoh else None,
                   "hall": np.zeros((NB, NB, ns, nl), np.complex128, order="F") if nl else None,
                   "hallo": np.zeros((NB, NB, ns, nl), np.complex128, order="F") if (nl and hoh) else None,
                   "enim": np.zeros((NB, NB, nt), np.complex128, order="F"),
                   "obarm": np.zeros((NB, NB, nt), np.complex128, order="F")}
        _lib.check(self._L.rsrec_build_hamiltonian(self._h, _p(hhh), _p(jt), _p(it), _p(packed), _p(mom), _p(ls), int(hoh),
                                                   *[_p(out.get(k)) for k in ("ee", "eeo", "hall", "hallo", "enim", "obarm")]))
        return out

    def rotate_to_local_axis(self, m_loc):
        """hamiltonian%rotate_to_local_axis (hamiltonian.f90:2442-2463) on the device-resident sets."""
        m = np.ascontiguousarray(m_loc, dtype=np.float64)
        _lib.check(self._L.rsrec_rotate_to_local_axis(self._h, _p(m)))

    def rotate_from_local_axis(self):
        _lib.check(self._L.rsrec_rotate_from_local_axis(self._h))

    def recur_b_local_axis(self, mom):
        """recur_b with hamiltonian%local_axis (recursion.f90:1826-1832): mom (3, nrec) = the moments of the recursion atoms."""
        s, e = self._local_units(len(self.lattice.irec))
        sites = np.ascontiguousarray(self.lattice.irec[s - 1:e], dtype=np.int32)
        mm = np.asfortranarray(np.asarray(mom, dtype=np.float64)[:, s - 1:e])
        lld, n = self.control.lld, len(sites)
        self.a_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        self.b2_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        _lib.check(self._L.rsrec_lanczos_block_local_axis(self._h, n, _p(sites), _p(mm), lld, _p(self.a_b), _p(self.b2_b)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.rsrec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers --------------------------------------------------------------------------------------------
    def _local_units(self, n):
        s, e = partition(self.rank, self.numprocs, n)   # get_mpi_variables, mpi.f90:32-58
        return s, e

    def set_kernel_family(self, family: int):
        _lib.check(self._L.rsrec_set_kernel_family(self._h, family))

    def set_fusion(self, lanczos: int = -1, cheb: int = -1):
        """Gram products inside the SpMV kernel: lanczos (default on), cheb (default off); -1 = unchanged."""
        _lib.check(self._L.rsrec_set_fusion(self._h, lanczos, cheb))

    @property
    def launch_count(self) -> int:
        return int(self._L.rsrec_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.rsrec_stream(self._h) or 0)

    def _pair_units(self):
        """Unit list of recur_b_ij / chebyshev_recur_ij: slot ij_loc*4-4+reci; i==j keeps only reci=1 with signs 1,1."""
        s, e = self._local_units(len(self.ijpair))
        nloc = max(e - s + 1, 0)
        slots, si, sj, asg, bsg = [], [], [], [], []
        signs = [(1, 1), (1, -1), (1, 1j), (1, -1j)]
        for ij in range(s, e + 1):
            i, j = (int(v) for v in self.ijpair[ij - 1])
            for reci in range(1, 5):
                if i == j and reci > 1:
                    continue
                a_, b_ = ((1.0, 1.0) if i == j else
                          (signs[reci - 1][0] * ONE_OVER_SQRT_TWO, signs[reci - 1][1] * ONE_OVER_SQRT_TWO))
                slots.append((ij - s) * 4 + reci - 1)
                si.append(i); sj.append(j); asg.append(a_); bsg.append(b_)
        return nloc, slots, (np.array(si, np.int32), np.array(sj, np.int32), np.array(asg, np.complex128),
                             np.array(bsg, np.complex128))

    # -- block Lanczos --------------------------------------------------------------------------------------
    def _lanczos_block(self, si, sj, asg, bsg):
        lld, n = self.control.lld, len(si)
        a_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        b2_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        _lib.check(self._L.rsrec_lanczos_block(self._h, n, _p(si), _p(sj), _p(asg), _p(bsg), lld, _p(a_b), _p(b2_b)))
        return a_b, b2_b

    def recur_b(self):
        s, e = self._local_units(len(self.lattice.irec))
        sites = np.ascontiguousarray(self.lattice.irec[s - 1:e], dtype=np.int32)
        self.a_b, self.b2_b = self._lanczos_block(sites, None, None, None)
        lld = self.control.lld
        self.a = np.zeros((lld, NB, len(sites), 3), order="F")
        self.b2 = np.zeros((lld, NB, len(sites), 3), order="F")
        d = np.arange(NB)
        self.a[:, :, :, 0] = np.real(self.a_b[d, d]).transpose(1, 0, 2)    # a(ll,l,i,1) = real(atemp_b(l,l,ll))
        self.b2[:, :, :, 0] = np.real(self.b2_b[d, d]).transpose(1, 0, 2)

    def recur_b_ij(self):
        nloc, slots, (si, sj, asg, bsg) = self._pair_units()
        a_b, b2_b = self._lanczos_block(si, sj, asg, bsg)
        lld = self.control.lld
        self.a_b = np.zeros((NB, NB, lld, 4 * nloc), np.complex128, order="F")
        self.b2_b = np.zeros((NB, NB, lld, 4 * nloc), np.complex128, order="F")
        self.a_b[..., slots] = a_b
        self.b2_b[..., slots] = b2_b

    def zsqr(self):
        b = np.asfortranarray(self.b2_b)
        _lib.check(self._L.rsrec_zsqr(self._h, _p(b), b.shape[2], b.shape[3]))
        self.b2_b = b

    # -- scalar Lanczos -------------------------------------------------------------------------------------
    def recur(self):
        s, e = self._local_units(len(self.lattice.irec))
        sites = np.ascontiguousarray(self.lattice.irec[s - 1:e], dtype=np.int32)
        lld = self.control.lld
        a = np.zeros((lld, NB, len(sites)), order="F")
        b2 = np.zeros((lld, NB, len(sites)), order="F")
        _lib.check(self._L.rsrec_lanczos_scalar(self._h, len(sites), _p(sites), lld, _p(a), _p(b2)))
        self.a = np.zeros((lld, NB, len(sites), 3), order="F")
        self.b2 = np.zeros((lld, NB, len(sites), 3), order="F")
        self.a[..., 0], self.b2[..., 0] = a, b2

    # -- Chebyshev ------------------------------------------------------------------------------------------
    def _cheb(self, si, sj, asg, bsg):
        lld, n = self.control.lld, len(si)
        a, b = self.en.scale_shift()
        mu = np.zeros((NB, NB, 2 * lld + 2, n), np.complex128, order="F")
        _lib.check(self._L.rsrec_cheb_moments(self._h, n, _p(si), _p(sj), _p(asg), _p(bsg), lld, a, b, _p(mu)))
        return mu

    def chebyshev_recur(self):
        s, e = self._local_units(len(self.lattice.irec))
        sites = np.ascontiguousarray(self.lattice.irec[s - 1:e], dtype=np.int32)
        self.mu_n = self._cheb(sites, None, None, None)

    def chebyshev_recur_ij(self):
        nloc, slots, (si, sj, asg, bsg) = self._pair_units()
        mu = self._cheb(si, sj, asg, bsg)
        self.mu_n = np.zeros((NB, NB, 2 * self.control.lld + 2, 4 * nloc), np.complex128, order="F")
        self.mu_n[..., slots] = mu

    def chebyshev_recur_random(self, phases=None):
        """KPM moments from random-phase start blocks (the start vector of recursion.f90:1131-1143), sharded over
        ranks with the reference's block rule; the caller all-reduces `mu_n.sum(-1)`."""
        ph = np.asfortranarray(self.phases if phases is None else phases, dtype=np.float64)
        s, e = self._local_units(ph.shape[1])
        loc = np.asfortranarray(ph[:, s - 1:e])
        lld = self.control.lld
        a, b = self.en.scale_shift()
        mu = np.zeros((NB, NB, 2 * lld + 2, loc.shape[1]), np.complex128, order="F")
        if loc.shape[1]:
            _lib.check(self._L.rsrec_cheb_moments_random(self._h, loc.shape[1], _p(loc), lld, a, b, _p(mu)))
        self.mu_n = mu

    # -- Kubo-Bastin ----------------------------------------------------------------------------------------
    def compute_moments_stochastic(self):
        M = self.control.cond_ll
        a, b = self.en.scale_shift()
        if self.control.cond_calctype == "per_type":
            sites = np.ascontiguousarray(self.atlist, dtype=np.int32)
            mu = np.zeros((NB, NB, M, M, len(sites)), np.complex128, order="F")
            _lib.check(self._L.rsrec_kubo_moments(self._h, len(sites), 0, _p(sites), None, M, a, b, _p(mu)))
        else:
            ph = np.asfortranarray(self.phases, dtype=np.float64)
            mu = np.zeros((NB, NB, M, M, ph.shape[1]), np.complex128, order="F")
            _lib.check(self._L.rsrec_kubo_moments(self._h, ph.shape[1], 1, None, _p(ph), M, a, b, _p(mu)))
        self.mu_nm_stochastic = mu

    # -- reachability map and the experimental orbital-moment KPM -----------------------------------------------
    def create_ll_map(self, site: int):
        """create_ll_map (3277-3303) for the start mask izeroll(site,1) = 1: -> self.izeroll (kk+1, lld+1) int32."""
        m = np.zeros((self.lattice.kk + 1, self.control.lld + 1), np.int32, order="F")
        _lib.check(self._L.rsrec_create_ll_map(self._h, int(site), self.control.lld, _p(m)))
        self.izeroll = m
        return m

    def chebyshev_orbital_mod(self, start_sites, cr, alat):
        """moment part of chebyshev_orbital_mod (2901-3008) for the given start sites: mu_n_orb (18,18,lld), un-normalised."""
        s = np.ascontiguousarray(start_sites, dtype=np.int32)
        crf = np.asfortranarray(cr, dtype=np.float64)
        a, b = self.en.scale_shift()
        mu = np.zeros((NB, NB, self.control.lld), np.complex128, order="F")
        _lib.check(self._L.rsrec_orbital_moments(self._h, len(s), _p(s), _p(crf), float(alat), self.control.lld, a, b, _p(mu)))
        self.mu_n_orb = mu
        return mu

    def chebyshev_orbital_tail(self, mu_n_orb=None, nstart=None, path=None):
        """tail of chebyshev_orbital_mod on the moments of the last `chebyshev_orbital_mod` call (see `orbital_tail`)"""
        out, self.lz_orb, self.lzi_orb = orbital_tail(self.mu_n_orb if mu_n_orb is None else mu_n_orb,
                                                      self.lattice.kk if nstart is None else nstart, self.en, path)
        return out

    # -- single operator applications -------------------------------------------------------------------------
    def ham_vec_matmul(self, psi_in, a, b):
        pin = _fc(psi_in)
        out = np.zeros_like(pin, order="F")
        _lib.check(self._L.rsrec_ham_vec_matmul(self._h, _p(pin), _p(out), a, b))
        return out

    def velo_vec_matmul(self, slot, psi_in):
        pin = _fc(psi_in)
        out = np.zeros_like(pin, order="F")
        _lib.check(self._L.rsrec_velo_vec_matmul(self._h, ord(slot), _p(pin), _p(out)))
        return out

    # -- device-resident stepping (bench) -----------------------------------------------------------------------
    def cheb_begin_random(self, phases, lld):
        ph = np.asfortranarray(phases, dtype=np.float64)
        a, b = self.en.scale_shift()
        self._sess = (ph.shape[1], lld)
        _lib.check(self._L.rsrec_cheb_begin_random(self._h, ph.shape[1], _p(ph), lld, a, b))

    def cheb_begin_sites(self, sites, lld):
        si = np.ascontiguousarray(sites, dtype=np.int32)
        a, b = self.en.scale_shift()
        self._sess = (len(si), lld)
        _lib.check(self._L.rsrec_cheb_begin_sites(self._h, len(si), _p(si), None, None, None, lld, a, b))

    def cheb_run_steps(self, n):
        _lib.check(self._L.rsrec_cheb_run_steps(self._h, n))

    def cheb_end(self):
        n, lld = self._sess
        mu = np.zeros((NB, NB, 2 * lld + 2, n), np.complex128, order="F")
        _lib.check(self._L.rsrec_cheb_end(self._h, _p(mu)))
        return mu

    @property
    def h2d_bytes(self) -> int:
        return int(self._L.rsrec_h2d_bytes(self._h))

    @property
    def d2h_bytes(self) -> int:
        return int(self._L.rsrec_d2h_bytes(self._h))

    def profile(self, enable: bool):
        _lib.check(self._L.rsrec_profile(self._h, int(enable)))

    def profile_read(self):
        """-> (summed gather-SpMV kernel milliseconds, launches) since profiling was enabled."""
        ms, n = C.c_double(0.0), C.c_int(0)
        _lib.check(self._L.rsrec_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def synchronize(self):
        _lib.check(self._L.rsrec_synchronize(self._h))

    # -- per-phase device timing under the reference's g_timer labels (recursion.f90:1902-1970, 3104-3127) ------
    def phase_timing(self, enable: bool = True):
        _lib.check(self._L.rsrec_phase_timing(self._h, int(enable)))

    def phase_read(self) -> dict:
        """-> {g_timer label: (milliseconds on the device, calls)} since the last read."""
        n = self._L.rsrec_phase_count()
        ms = np.zeros(n, np.float64)
        calls = np.zeros(n, np.int64)
        _lib.check(self._L.rsrec_phase_read(self._h, _p(ms), _p(calls)))
        return {self._L.rsrec_phase_label(k).decode(): (float(ms[k]), int(calls[k])) for k in range(n) if calls[k]}

    def host_phase_read(self) -> dict:
        """host wall-clock seconds per stage of the sharded calls since the last read."""
        sec = np.zeros(5, np.float64)
        _lib.check(self._L.rsrec_host_phase_read(self._h, _p(sec)))
        return dict(zip(("tables", "plan", "recursion", "exchange", "download"), (float(x) for x in sec)))

    # -- the exchange step of the unit-sharded path: NCCL inside the library -------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0 creates the NCCL id; the host broadcasts these 128 bytes (MPI_Bcast in the Fortran host)."""
        buf = (C.c_ubyte * 128)()
        _lib.check(_lib.load().rsrec_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _lib.check(self._L.rsrec_comm_init(self._h, nranks, rank, buf))
        self.rank, self.numprocs = rank, nranks

    def comm_init_torch(self):
        """bench/tests convenience: the 128-byte id travels over an existing torch.distributed group (any backend)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(world, rank, box[0])

    def comm_destroy(self):
        _lib.check(self._L.rsrec_comm_destroy(self._h))

    def comm_info(self):
        n, r, v = C.c_int(0), C.c_int(0), C.c_int(0)
        _lib.check(self._L.rsrec_comm_info(self._h, C.byref(n), C.byref(r), C.byref(v)))
        return n.value, r.value, v.value

    def allreduce(self, arr: np.ndarray) -> np.ndarray:
        """MPI_ALLREDUCE(MPI_IN_PLACE, arr, SUM) of a host array (float64 / complex128 / int32), in place."""
        kind = {np.dtype(np.float64): 0, np.dtype(np.complex128): 1, np.dtype(np.int32): 2}[arr.dtype]
        assert arr.flags.f_contiguous or arr.flags.c_contiguous
        _lib.check(self._L.rsrec_allreduce(self._h, _p(arr), arr.size, kind))
        return arr

    def allgather_units(self, local: np.ndarray, n_units: int) -> np.ndarray:
        """per-unit results (last axis = local unit) of all ranks in global unit order (recursion.f90:1788-1799)."""
        loc = np.asfortranarray(local)
        full = np.zeros(loc.shape[:-1] + (n_units,), dtype=loc.dtype, order="F")
        per = int(np.prod(loc.shape[:-1])) * (2 if loc.dtype == np.complex128 else 1)
        _lib.check(self._L.rsrec_allgather_units(self._h, _p(loc) if loc.size else None, _p(full), per, n_units))
        return full

    def recur_b_sharded(self):
        """recur_b over ALL lattice.irec: every rank runs its block-rule shard, a_b/b2_b (18,18,lld,nrec) are gathered on
        the device over NCCL and returned on every rank."""
        sites = np.ascontiguousarray(self.lattice.irec, dtype=np.int32)
        lld, n = self.control.lld, len(sites)
        self.a_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        self.b2_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        _lib.check(self._L.rsrec_lanczos_block_sharded(self._h, n, _p(sites), None, None, None, lld, _p(self.a_b), _p(self.b2_b)))

    def recur_b_ij_sharded(self):
        """recur_b_ij over ALL pairs of the job (units = 4 start vectors per pair, 1 for i == j): each rank runs its block-rule
        shard of the unit list, the coefficients are gathered on the device -> a_b, b2_b (18,18,lld,4*njij) on every rank."""
        rank, nprocs = self.rank, self.numprocs
        self.rank, self.numprocs = 0, 1                      # the full unit list; the library shards it
        try:
            nloc, slots, (si, sj, asg, bsg) = self._pair_units()
        finally:
            self.rank, self.numprocs = rank, nprocs
        lld, n = self.control.lld, len(si)
        a_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        b2_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        _lib.check(self._L.rsrec_lanczos_block_sharded(self._h, n, _p(si), _p(sj), _p(asg), _p(bsg), lld, _p(a_b), _p(b2_b)))
        self.a_b = np.zeros((NB, NB, lld, 4 * nloc), np.complex128, order="F")
        self.b2_b = np.zeros((NB, NB, lld, 4 * nloc), np.complex128, order="F")
        self.a_b[..., slots] = a_b
        self.b2_b[..., slots] = b2_b

    def chebyshev_recur_random_sum(self, phases=None, sharded: bool = True):
        """stochastic-trace moments: sum over ALL random vectors of the job (this rank runs its shard of the columns of
        `phases`; the sum over local vectors and over ranks happens on the device) -> mu_sum (18,18,2*lld+2)."""
        ph = np.asfortranarray(self.phases if phases is None else phases, dtype=np.float64)
        if sharded:
            s, e = self._local_units(ph.shape[1])
            loc = np.asfortranarray(ph[:, s - 1:e])
        else:                       # `phases` already is this rank's shard (e.g. a pinned staging buffer)
            loc = ph
        lld = self.control.lld
        a, b = self.en.scale_shift()
        mu = np.zeros((NB, NB, 2 * lld + 2), np.complex128, order="F")
        _lib.check(self._L.rsrec_cheb_moments_random_sum(self._h, loc.shape[1], _p(loc) if loc.shape[1] else None, lld, a, b, _p(mu)))
        self.mu_sum = mu
        return mu
