"""Host-side mirrors of the reference types that consume the recursion results:
`green` (green.f90), `dos` (density_of_states.f90) and `conductivity` (conductivity.f90), on top of the C ABI.

Same procedure names and result members as the reference; everything numerical runs in librsrec.so on the GPU.

    g = Green(recursion)                 # green(dos_obj) -> recursion, energy, control
    g.block_green()   -> g.g0 (18,18,nv,nrec_local)     green.f90:588-621   (needs recursion.zsqr() first, like run_dos)
    g.chebyshev_green() -> g.g0, recursion.mu_ng        green.f90:1030-1108
    g.sgreen()        -> g.g0                           green.f90:628-705
    g.bgreen(ia, ie_start, ie_len, a_inf, b_inf, eta)   green.f90:1191-1339
    Dos(recursion).density(ia, mdir) -> tdens (18,nv)   density_of_states.f90:248-372
    Conductivity(recursion).calculate_conductivity_tensor() -> integrand (18,nv), integrand_at   conductivity.f90:228-306
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib

NB = 18


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a, dt):
    return np.asfortranarray(a, dtype=dt)


class _Consumer:
    def __init__(self, recursion):
        self.recursion = recursion
        self.en = recursion.en
        self.control = recursion.control
        self._L, self._h = recursion._L, recursion._h
        if self.en.ene is None:
            self.en.e_mesh()

    @property
    def ene(self):
        return np.ascontiguousarray(self.en.ene, dtype=np.float64)


class Green(_Consumer):
    def __init__(self, recursion, sym_term: bool = False):
        super().__init__(recursion)
        self.sym_term = sym_term          # control%sym_term (green.f90:1240)
        self.g0 = None

    def get_terminf(self):
        """recursion%get_terminf (recursion.f90:2092-2138) on the recursion's a_b / b2_b (= B after zsqr)."""
        a_b, b_b = _f(self.recursion.a_b, np.complex128), _f(self.recursion.b2_b, np.complex128)
        ll, na = a_b.shape[2], a_b.shape[3]
        a_inf = np.zeros((NB, NB, na), order="F"); b_inf = np.zeros((NB, NB, na), order="F")
        a0 = np.zeros(na); b0 = np.zeros(na)
        _lib.check(self._L.rsrec_get_terminf(self._h, _p(a_b), _p(b_b), na, ll, _p(a_inf), _p(b_inf), _p(a0), _p(b0)))
        return a_inf, b_inf, a0, b0

    def bgreen(self, ia, ie_start, ie_len, a_inf, b_inf, eta=0.0):
        """one unit `ia` (1-based local index); returns g_out (18,18,nv)."""
        a_b = _f(self.recursion.a_b[..., ia - 1], np.complex128)
        b_b = _f(self.recursion.b2_b[..., ia - 1], np.complex128)
        ene = self.ene
        g = np.zeros((NB, NB, len(ene)), np.complex128, order="F")
        ai, bi = _f(a_inf, np.float64), _f(b_inf, np.float64)
        eta = complex(eta)
        _lib.check(self._L.rsrec_bgreen(self._h, _p(a_b), _p(b_b), a_b.shape[2], _p(ene), len(ene), ie_start, ie_len,
                                        _p(ai), _p(bi), eta.real, eta.imag, int(self.sym_term), _p(g)))
        return g

    def block_green(self):
        a_b, b_b = _f(self.recursion.a_b, np.complex128), _f(self.recursion.b2_b, np.complex128)
        ll, na = a_b.shape[2], a_b.shape[3]
        ene = self.ene
        self.g0 = np.zeros((NB, NB, len(ene), na), np.complex128, order="F")
        _lib.check(self._L.rsrec_block_green(self._h, _p(a_b), _p(b_b), na, ll, _p(ene), len(ene), int(self.sym_term),
                                             _p(self.g0)))
        return self.g0

    def chebyshev_green(self):
        mu = _f(self.recursion.mu_n, np.complex128)
        nk, na = mu.shape[2], mu.shape[3]
        ene = self.ene
        self.recursion.mu_ng = np.zeros_like(mu, order="F")
        self.g0 = np.zeros((NB, NB, len(ene), na), np.complex128, order="F")
        _lib.check(self._L.rsrec_chebyshev_green(self._h, _p(mu), na, (nk - 2) // 2, _p(ene), len(ene),
                                                 self.en.energy_min, self.en.energy_max, _p(self.recursion.mu_ng),
                                                 _p(self.g0)))
        return self.g0

    # -- fused: recursion + Green function without a host round trip of the coefficients ------------------------
    def recur_b_green(self, download_g0: bool = True):
        """run_recursion + run_dos of the block path (self.f90:799-856) in one call: fills recursion.a_b, b2_b (= B^2,
        as recur_b leaves it), recursion.a, b2 and self.g0.  download_g0 = False leaves g0 on the device only (for
        `Bands`)."""
        rec = self.recursion
        s, e = rec._local_units(len(rec.lattice.irec))
        sites = np.ascontiguousarray(rec.lattice.irec[s - 1:e], dtype=np.int32)
        lld, n = self.control.lld, len(sites)
        ene = self.ene
        rec.a_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        rec.b2_b = np.zeros((NB, NB, lld, n), np.complex128, order="F")
        self.g0 = np.zeros((NB, NB, len(ene), n), np.complex128, order="F") if download_g0 else None
        _lib.check(self._L.rsrec_recur_b_green(self._h, n, _p(sites), lld, _p(ene), len(ene), int(self.sym_term),
                                               _p(rec.a_b), _p(rec.b2_b), _p(self.g0)))
        d = np.arange(NB)
        rec.a = np.zeros((lld, NB, n, 3), order="F"); rec.b2 = np.zeros((lld, NB, n, 3), order="F")
        rec.a[:, :, :, 0] = np.real(rec.a_b[d, d]).transpose(1, 0, 2)
        rec.b2[:, :, :, 0] = np.real(rec.b2_b[d, d]).transpose(1, 0, 2)
        return self.g0

    def chebyshev_recur_green(self, keep_moments: bool = True, download_g0: bool = True):
        """chebyshev_recur + chebyshev_green in one call."""
        rec = self.recursion
        s, e = rec._local_units(len(rec.lattice.irec))
        sites = np.ascontiguousarray(rec.lattice.irec[s - 1:e], dtype=np.int32)
        lld, n = self.control.lld, len(sites)
        ene = self.ene
        if keep_moments:
            rec.mu_n = np.zeros((NB, NB, 2 * lld + 2, n), np.complex128, order="F")
            rec.mu_ng = np.zeros((NB, NB, 2 * lld + 2, n), np.complex128, order="F")
        self.g0 = np.zeros((NB, NB, len(ene), n), np.complex128, order="F") if download_g0 else None
        _lib.check(self._L.rsrec_cheb_recur_green(self._h, n, _p(sites), lld, self.en.energy_min, self.en.energy_max,
                                                  _p(ene), len(ene), _p(rec.mu_n) if keep_moments else None,
                                                  _p(rec.mu_ng) if keep_moments else None, _p(self.g0)))
        return self.g0

    # -- exchange path: green%calculate_intersite_gf (green.f90:425-469) ------------------------------------------------
    def calculate_intersite_gf(self, fused: bool = True):
        """gij, gji (18,18,nv,njij_local) and their spin components Ginmag, Gix, Giy, Giz, Gjnmag, Gjx, Gjy, Gjz
        (9,9,nv,njij_local) for this rank's pairs of `recursion.ijpair`, `control.recur` = 'block' | 'chebyshev'.
        fused: the pair recursion (recur_b_ij / chebyshev_recur_ij) and the Green functions run in one call with g0 left on
        the device; otherwise recursion.a_b / b2_b (after zsqr) or mu_n with four slots per pair are used, like the
        reference does after run_recursion."""
        rec = self.recursion
        nloc, slots, (si, sj, asg, bsg) = rec._pair_units()
        s0, _ = rec._local_units(len(rec.ijpair))
        pairs = np.asarray(rec.ijpair[s0 - 1:s0 - 1 + nloc], dtype=np.int32).reshape(-1, 2)
        pi, pj = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
        ene = self.ene
        nv, lld = len(ene), self.control.lld
        block = getattr(self.control, "recur", "block") == "block"
        if fused:
            n = len(si)
            if block:
                _lib.check(self._L.rsrec_recur_b_ij_green(self._h, n, _p(si), _p(sj), _p(asg), _p(bsg), lld, _p(ene), nv,
                                                          int(self.sym_term), None, None, None))
            else:
                _lib.check(self._L.rsrec_cheb_recur_ij_green(self._h, n, _p(si), _p(sj), _p(asg), _p(bsg), lld,
                                                             self.en.energy_min, self.en.energy_max, _p(ene), nv, None, None, None))
        elif block:
            self.block_green()
        else:
            self.chebyshev_green()
        self.gij = np.zeros((NB, NB, nv, nloc), np.complex128, order="F")
        self.gji = np.zeros((NB, NB, nv, nloc), np.complex128, order="F")
        gs = np.zeros((9, 9, nv, nloc, 8), np.complex128, order="F")
        _lib.check(self._L.rsrec_intersite_gf(self._h, nloc, _p(pi), _p(pj), int(fused), _p(self.gij), _p(self.gji), _p(gs)))
        (self.ginmag, self.gix, self.giy, self.giz, self.gjnmag, self.gjx, self.gjy, self.gjz) = (gs[..., c] for c in range(8))
        return self.gij, self.gji

    def sgreen(self, dw_l, cshi, nmdir: int = 1):
        """dw_l, cshi (18,na): potential parameters sqrt(Delta) and the band-centre shift of each atom
        (density_of_states.f90:300-304)."""
        a, b2 = _f(self.recursion.a, np.float64), _f(self.recursion.b2, np.float64)
        lld, na = a.shape[0], a.shape[2]
        ene = self.ene
        dw, cs = _f(dw_l, np.float64), _f(cshi, np.float64)
        self.g0 = np.zeros((NB, NB, len(ene), na), np.complex128, order="F")
        _lib.check(self._L.rsrec_sgreen(self._h, _p(a), _p(b2), lld, na, nmdir, _p(ene), len(ene), _p(dw), _p(cs),
                                        _p(self.g0)))
        return self.g0


class Dos(_Consumer):
    def density_all(self, dw_l, cshi, nmdir: int = 1):
        """tdens (18,nv,na,nmdir) for every atom and direction in one launch set."""
        a, b2 = _f(self.recursion.a, np.float64), _f(self.recursion.b2, np.float64)
        lld, na = a.shape[0], a.shape[2]
        ene = self.ene
        dw, cs = _f(dw_l, np.float64), _f(cshi, np.float64)
        td = np.zeros((NB, len(ene), na, nmdir), order="F")
        _lib.check(self._L.rsrec_density(self._h, _p(a), _p(b2), lld, na, nmdir, _p(ene), len(ene), _p(dw), _p(cs),
                                         _p(td)))
        return td

    def density(self, ia, mdir, dw_l, cshi):
        """dos%density(tdens, ia, mdir): one atom (1-based local index), one direction; dw_l, cshi (18)."""
        a = _f(self.recursion.a[:, :, ia - 1:ia, mdir - 1:mdir], np.float64)
        b2 = _f(self.recursion.b2[:, :, ia - 1:ia, mdir - 1:mdir], np.float64)
        ene = self.ene
        dw, cs = _f(np.reshape(dw_l, (NB, 1)), np.float64), _f(np.reshape(cshi, (NB, 1)), np.float64)
        td = np.zeros((NB, len(ene)), order="F")
        _lib.check(self._L.rsrec_density(self._h, _p(a), _p(b2), a.shape[0], 1, 1, _p(ene), len(ene), _p(dw), _p(cs),
                                         _p(td)))
        return td

    def bpopt(self, a, rb):
        """batched bpopt: a, rb (ll,nchains) -> ainf, rbinf, ifail (nchains)."""
        a, rb = _f(a, np.float64), _f(rb, np.float64)
        if a.ndim == 1:
            a, rb = a.reshape(-1, 1, order="F"), rb.reshape(-1, 1, order="F")
        n = a.shape[1]
        ainf, rbinf, ifail = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
        _lib.check(self._L.rsrec_bpopt(self._h, n, a.shape[0], _p(a), _p(rb), _p(ainf), _p(rbinf), _p(ifail)))
        return ainf, rbinf, ifail


class Conductivity(_Consumer):
    def calculate_conductivity_tensor(self):
        """Gamma_nm contracted with the diagonal of mu_nm_stochastic: the energy integrand of
        conductivity.f90:267-290 (the Simpson integration and file output stay with the host program)."""
        mu = _f(self.recursion.mu_nm_stochastic, np.complex128)
        M, nloop = mu.shape[2], mu.shape[4]
        ene = self.ene
        per_type = self.control.cond_calctype == "per_type"
        self.integrand = np.zeros((NB, len(ene)), np.complex128, order="F")
        self.integrand_at = np.zeros((NB, len(ene), nloop), np.complex128, order="F")
        _lib.check(self._L.rsrec_conductivity_integrand(self._h, _p(mu), M, nloop, _p(ene), len(ene), self.en.energy_min,
                                                        self.en.energy_max, int(per_type), _p(self.integrand),
                                                        _p(self.integrand_at)))
        return self.integrand, self.integrand_at

    def integrate_conductivity(self):
        """tail of calculate_conductivity_tensor (conductivity.f90:300-372): sigma(E_F) for every mesh energy from
        self.integrand / self.integrand_at -> self.sigma (2,19,nv,1+nat): (re|im, total|orbital, energy, summed|per type),
        the numbers the reference writes to cond_total*.out and <symbol>_cond*.out."""
        ene = self.ene
        a, b = self.en.scale_shift()
        ws = (ene - b) / a
        per_type = self.control.cond_calctype == "per_type"
        nat = self.integrand_at.shape[2] if per_type else 0
        loop_over = self.integrand_at.shape[2]
        integ = _f(self.integrand, np.complex128)
        iat = _f(self.integrand_at, np.complex128) if nat else None
        self.sigma = np.zeros((2, 19, len(ene), 1 + nat), order="F")
        _lib.check(self._L.rsrec_conductivity_cumulative(self._h, _p(integ), _p(iat), len(ene), self.en.nv1, nat,
                                                         float(ws[1] - ws[0]), loop_over, _p(self.sigma)))
        return self.sigma

    def compute_conductivity(self, keep_moments: bool = False):
        """compute_moments_stochastic + calculate_conductivity_tensor's integrand in one call; only the diagonals of
        mu_nm_stochastic the integrand consumes are kept (on the device) unless keep_moments is set."""
        rec = self.recursion
        M = self.control.cond_ll
        ene = self.ene
        per_type = self.control.cond_calctype == "per_type"
        if per_type:
            sites = np.ascontiguousarray(rec.atlist, dtype=np.int32)
            nstart, ph = len(sites), None
        else:
            # random vectors are the units of this path: each rank runs its block-rule shard (mpi.f90:32-58) and the library
            # all-reduces the integrand over the communicator attached with Recursion.comm_init
            ph = _f(rec.phases, np.float64)
            s, e = rec._local_units(ph.shape[1])
            ph = np.asfortranarray(ph[:, s - 1:e])
            nstart, sites = ph.shape[1], None
        mu = np.zeros((NB, NB, M, M, nstart), np.complex128, order="F") if keep_moments else None
        self.integrand = np.zeros((NB, len(ene)), np.complex128, order="F")
        self.integrand_at = np.zeros((NB, len(ene), max(nstart, 1)), np.complex128, order="F")
        _lib.check(self._L.rsrec_kubo_conductivity(self._h, nstart, 0 if per_type else 1, _p(sites), _p(ph) if nstart else None, M,
                                                   self.en.energy_min, self.en.energy_max, _p(ene), len(ene), _p(mu),
                                                   _p(self.integrand), _p(self.integrand_at)))
        if keep_moments:
            rec.mu_nm_stochastic = mu
        return self.integrand, self.integrand_at
