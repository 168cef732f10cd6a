"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- never imported by the product package).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY PINNED by the reference's bccFe golden fixtures (oracle/ref_bccfe.py, tests/test_reference_golden.py) and by
oracle/dense_check*.py for what those do not reach -- see rsrec_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librsrec_oracle.so")
_lib = None

c_i32p = C.POINTER(C.c_int32)
c_dp = C.POINTER(C.c_double)
c_vp = C.c_void_p


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("rsrec_oracle.c", "rsrec_oracle_post.c", "rsrec_oracle_lattice.c", "rsrec_oracle_bands.c", "rsrec_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librsrec_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()   # no-op unless a source is newer than the .so
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_create.restype = c_vp
        _lib.orc_create.argtypes = [C.c_int] * 5 + [c_vp, c_vp]
        _lib.orc_destroy.argtypes = [c_vp]
        _lib.orc_set_hamiltonian.argtypes = [c_vp] * 7 + [C.c_int]
        _lib.orc_set_operator.argtypes = [c_vp, C.c_int, c_vp, c_vp]
        _lib.orc_set_use_mask.argtypes = [c_vp, C.c_int]
        _lib.orc_set_threads.argtypes = [C.c_int]
        _lib.orc_get_max_threads.restype = C.c_int
        _lib.orc_lanczos_block.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp]
        _lib.orc_lanczos_scalar.argtypes = [c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp]
        _lib.orc_cheb_moments.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_double, C.c_double, c_vp]
        _lib.orc_cheb_moments_random.argtypes = [c_vp, C.c_int, c_vp, C.c_int, C.c_double, C.c_double, c_vp]
        _lib.orc_cheb_time_steps.argtypes = [c_vp, c_vp, C.c_int, C.c_double, C.c_double, c_vp]
        _lib.orc_kubo_moments.argtypes = [c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, C.c_double, C.c_double, c_vp]
        _lib.orc_kubo_moments_cols.argtypes = [c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, C.c_double, C.c_double, c_vp, C.c_int, c_vp]
        _lib.orc_zsqr.argtypes = [c_vp, C.c_int, C.c_int]
        _lib.orc_ham_vec_matmul.argtypes = [c_vp, c_vp, c_vp, C.c_double, C.c_double, c_vp]
        _lib.orc_velo_vec_matmul.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp]
        _lib.orc_heev18.argtypes = [c_vp, c_vp]
        _lib.orc_last_irnum.argtypes = [c_vp]
        _lib.orc_create_ll_map.argtypes = [c_vp, C.c_int, c_vp]
        _lib.orc_orbital_moments.argtypes = [c_vp, C.c_int, c_vp, c_vp, C.c_double, C.c_int, C.c_double, C.c_double, c_vp]
        _lib.orc_last_irnum.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _f(a, dtype):
    return None if a is None else np.asfortranarray(a, dtype=dtype)


class Oracle:
    """Holds one (lattice, hamiltonian) pair; methods mirror the reference's `recursion` procedures."""

    def __init__(self, lat, ham, use_mask: bool = True, threads: int | None = None):
        L = lib()
        self.lat, self.ham = lat, ham
        self.nn = _f(lat.nn, np.int32)
        self.iz = _f(lat.iz, np.int32)
        self._keep = [_f(getattr(ham, k), np.complex128) for k in ("ee", "eeo", "hall", "hallo", "lsham", "enim")]
        self.h = L.orc_create(lat.kk, lat.ncols, lat.nslot, lat.ntype, lat.nmax, _p(self.nn), _p(self.iz))
        L.orc_set_hamiltonian(self.h, *[_p(a) for a in self._keep], int(ham.hoh))
        self._ops = [_f(getattr(ham, k, None), np.complex128) for k in ("v_a", "v_b")]
        self._vops = [_f(getattr(ham, k, None), np.complex128) for k in ("vo_a", "vo_b")]
        if self._ops[0] is not None:
            L.orc_set_operator(self.h, ord("a"), _p(self._ops[0]), _p(self._vops[0]))
            L.orc_set_operator(self.h, ord("b"), _p(self._ops[1]), _p(self._vops[1]))
        L.orc_set_use_mask(self.h, int(use_mask))
        if threads:
            L.orc_set_threads(threads)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    @staticmethod
    def _units(site_i, site_j, asign, bsign):
        si = np.ascontiguousarray(site_i, dtype=np.int32)
        n = len(si)
        sj = np.zeros(n, np.int32) if site_j is None else np.ascontiguousarray(site_j, dtype=np.int32)
        a = np.ones(n, np.complex128) if asign is None else np.ascontiguousarray(asign, dtype=np.complex128)
        b = np.ones(n, np.complex128) if bsign is None else np.ascontiguousarray(bsign, dtype=np.complex128)
        return n, si, sj, a, b

    def lanczos_block(self, site_i, lld, site_j=None, asign=None, bsign=None):
        n, si, sj, a, b = self._units(site_i, site_j, asign, bsign)
        a_b = np.zeros((18, 18, lld, n), np.complex128, order="F")
        b2_b = np.zeros((18, 18, lld, n), np.complex128, order="F")
        lib().orc_lanczos_block(self.h, n, _p(si), _p(sj), _p(a), _p(b), lld, _p(a_b), _p(b2_b))
        return a_b, b2_b

    def lanczos_scalar(self, sites, lld):
        s = np.ascontiguousarray(sites, dtype=np.int32)
        a = np.zeros((lld, 18, len(s)), np.float64, order="F")
        b2 = np.zeros((lld, 18, len(s)), np.float64, order="F")
        lib().orc_lanczos_scalar(self.h, len(s), _p(s), lld, _p(a), _p(b2))
        return a, b2

    def cheb_moments(self, site_i, lld, a, b, site_j=None, asign=None, bsign=None):
        n, si, sj, as_, bs_ = self._units(site_i, site_j, asign, bsign)
        mu = np.zeros((18, 18, 2 * lld + 2, n), np.complex128, order="F")
        rc = lib().orc_cheb_moments(self.h, n, _p(si), _p(sj), _p(as_), _p(bs_), lld, a, b, _p(mu))
        return mu, rc

    def cheb_moments_random(self, phases, lld, a, b):
        ph = np.asfortranarray(phases, dtype=np.float64)
        nvec = ph.shape[1]
        mu = np.zeros((18, 18, 2 * lld + 2, nvec), np.complex128, order="F")
        rc = lib().orc_cheb_moments_random(self.h, nvec, _p(ph), lld, a, b, _p(mu))
        return mu, rc

    def cheb_time_steps(self, phases, nsteps, a, b):
        """seconds for `nsteps` chebyshev_recur_ll steps of one random vector (CPU baseline)."""
        ph = np.ascontiguousarray(phases, dtype=np.float64)
        sec = C.c_double(0.0)
        rc = lib().orc_cheb_time_steps(self.h, _p(ph), nsteps, a, b, C.byref(sec))
        assert rc == 0
        return sec.value

    def kubo_moments(self, cond_ll, a, b, start_sites=None, phases=None):
        if start_sites is not None:
            s = np.ascontiguousarray(start_sites, dtype=np.int32)
            n, kind, ph = len(s), 0, None
        else:
            ph = np.asfortranarray(phases, dtype=np.float64)
            n, kind, s = ph.shape[1], 1, None
        mu = np.zeros((18, 18, cond_ll, cond_ll, n), np.complex128, order="F")
        rc = lib().orc_kubo_moments(self.h, n, kind, _p(s), _p(ph), cond_ll, a, b, _p(mu))
        assert rc == 0
        return mu

    def kubo_moments_cols(self, cond_ll, a, b, msel, start_sites=None, phases=None):
        """compute_moments_stochastic restricted to the left indices `msel` (1-based): mu(18,18,cond_ll,len(msel),nstart).
        Same chains as kubo_moments; for full-size lattices where all cond_ll^2 contractions take minutes on a CPU."""
        if start_sites is not None:
            s = np.ascontiguousarray(start_sites, dtype=np.int32)
            n, kind, ph = len(s), 0, None
        else:
            ph = np.asfortranarray(phases, dtype=np.float64)
            n, kind, s = ph.shape[1], 1, None
        ms = np.ascontiguousarray(msel, dtype=np.int32)
        mu = np.zeros((18, 18, cond_ll, len(ms), n), np.complex128, order="F")
        rc = lib().orc_kubo_moments_cols(self.h, n, kind, _p(s), _p(ph), cond_ll, a, b, _p(ms), len(ms), _p(mu))
        assert rc == 0
        return mu

    def create_ll_map(self, site, lld):
        """izeroll (kk+1, lld+1) with column 1 = the start mask of chebyshev_recur (recursion.f90:3086-3091)."""
        m = np.zeros((self.lat.kk + 1, lld + 1), np.int32, order="F")
        m[site, 0] = 1
        lib().orc_create_ll_map(self.h, lld, _p(m))
        return m

    def orbital_moments(self, start_sites, cr, alat, lld, a, b):
        s = np.ascontiguousarray(start_sites, dtype=np.int32)
        crf = np.asfortranarray(cr, dtype=np.float64)
        mu = np.zeros((18, 18, lld), np.complex128, order="F")
        rc = lib().orc_orbital_moments(self.h, len(s), _p(s), _p(crf), float(alat), lld, a, b, _p(mu))
        assert rc == 0
        return mu

    def zsqr(self, b2_b):
        out = np.array(b2_b, dtype=np.complex128, order="F", copy=True)
        lib().orc_zsqr(_p(out), out.shape[2], out.shape[3])
        return out

    def ham_vec_matmul(self, psi_in, a, b, izero):
        pin = np.asfortranarray(psi_in, dtype=np.complex128)
        out = np.zeros_like(pin, order="F")
        iz = np.ascontiguousarray(izero, dtype=np.int32).copy()
        lib().orc_ham_vec_matmul(self.h, _p(pin), _p(out), a, b, _p(iz))
        return out, iz

    def velo_vec_matmul(self, slot, psi_in, izero):
        pin = np.asfortranarray(psi_in, dtype=np.complex128)
        out = np.zeros_like(pin, order="F")
        iz = np.ascontiguousarray(izero, dtype=np.int32).copy()
        lib().orc_velo_vec_matmul(self.h, ord(slot), _p(pin), _p(out), _p(iz))
        return out, iz

    def last_irnum(self):
        return lib().orc_last_irnum(self.h)


def heev18(m):
    u = np.array(m, dtype=np.complex128, order="F", copy=True)
    ev = np.zeros(18)
    lib().orc_heev18(_p(u), _p(ev))
    return ev, u


def cheb_scale(emin: float, emax: float):
    """a, b of `recursion.f90:3078-3079`."""
    return (emax - emin) / (2 - 0.3), (emax + emin) / 2


# ---- consumers either side of the hot path (rsrec_oracle_post.c) -----------------------------------
def _post():
    L = lib()
    if not getattr(L, "_post_ready", False):
        L.orc_emami.argtypes = [C.c_int, c_vp, c_vp, c_vp, c_vp]
        L.orc_bpopt.argtypes = [C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.orc_get_terminf.argtypes = [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp]
        L.orc_bgreen.argtypes = [c_vp, c_vp, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp, C.c_double, C.c_double,
                                 C.c_int, c_vp]
        L.orc_block_green.argtypes = [c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, c_vp]
        L.orc_jackson_kernel.argtypes = [C.c_int, c_vp]
        L.orc_lorentz_kernel.argtypes = [C.c_int, C.c_double, c_vp]
        L.orc_chebyshev_green.argtypes = [c_vp, C.c_int, C.c_int, c_vp, C.c_int, C.c_double, C.c_double, c_vp, c_vp]
        L.orc_bprldos.argtypes = [C.c_double, c_vp, c_vp, C.c_int, c_vp]
        L.orc_bprldos.restype = C.c_double
        L.orc_density.argtypes = [c_vp, c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp, c_vp]
        L.orc_sgreen.argtypes = [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp, c_vp]
        L.orc_gamma_nm.argtypes = [c_vp, C.c_int, C.c_int, C.c_double, C.c_double, c_vp]
        L.orc_conductivity_integrand.argtypes = [c_vp, C.c_int, C.c_int, c_vp, C.c_int, C.c_double, C.c_double, C.c_int,
                                                 c_vp, c_vp]
        L._post_ready = True
    return L


def e_mesh(energy_min, energy_max, channels_ldos, fermi):
    """energy%e_mesh (energy.f90:175-208) -> ene(channels_ldos+10) (channels_ldos made odd-compatible like the reference)."""
    if channels_ldos % 2 != 0:
        channels_ldos -= 1
    edel = (energy_max - energy_min) / channels_ldos
    edel = (fermi - energy_min) / np.rint((fermi - energy_min) / edel)
    return energy_min + edel * np.arange(channels_ldos + 10, dtype=np.float64)


def emami(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    emax, emin = C.c_double(0), C.c_double(0)
    _post().orc_emami(len(a), _p(a), _p(b), C.byref(emax), C.byref(emin))
    return emax.value, emin.value


def bpopt(a, rb):
    a = np.ascontiguousarray(a, np.float64); rb = np.ascontiguousarray(rb, np.float64)
    ainf, rbinf, ifail = C.c_double(0), C.c_double(0), C.c_int(0)
    _post().orc_bpopt(len(a), _p(a), _p(rb), C.byref(ainf), C.byref(rbinf), C.byref(ifail))
    return ainf.value, rbinf.value, ifail.value


def get_terminf(a_b, b_b):
    a_b = _f(a_b, np.complex128); b_b = _f(b_b, np.complex128)
    ll, na = a_b.shape[2], a_b.shape[3]
    a_inf = np.zeros((18, 18, na), order="F"); b_inf = np.zeros((18, 18, na), order="F")
    a0 = np.zeros(na); b0 = np.zeros(na)
    _post().orc_get_terminf(_p(a_b), _p(b_b), na, ll, _p(a_inf), _p(b_inf), _p(a0), _p(b0))
    return a_inf, b_inf, a0, b0


def bgreen(a_b, b_b, ene, a_inf, b_inf, eta=0.0, sym_term=False, ie_start=1, ie_len=None):
    """one unit: a_b, b_b (18,18,ll); returns g_out (18,18,nv)."""
    a_b = _f(a_b, np.complex128); b_b = _f(b_b, np.complex128)
    ene = np.ascontiguousarray(ene, np.float64)
    nv = len(ene)
    g = np.zeros((18, 18, nv), np.complex128, order="F")
    ai = _f(a_inf, np.float64); bi = _f(b_inf, np.float64)
    _post().orc_bgreen(_p(a_b), _p(b_b), a_b.shape[2], _p(ene), nv, ie_start, nv if ie_len is None else ie_len, _p(ai),
                       _p(bi), complex(eta).real, complex(eta).imag, int(sym_term), _p(g))
    return g


def block_green(a_b, b_b, ene, sym_term=False):
    a_b = _f(a_b, np.complex128); b_b = _f(b_b, np.complex128)
    ene = np.ascontiguousarray(ene, np.float64)
    ll, na, nv = a_b.shape[2], a_b.shape[3], len(ene)
    g0 = np.zeros((18, 18, nv, na), np.complex128, order="F")
    _post().orc_block_green(_p(a_b), _p(b_b), na, ll, _p(ene), nv, int(sym_term), _p(g0))
    return g0


def jackson_kernel(n):
    k = np.zeros(n)
    _post().orc_jackson_kernel(n, _p(k))
    return k


def lorentz_kernel(n, lam):
    k = np.zeros(n)
    _post().orc_lorentz_kernel(n, lam, _p(k))
    return k


def chebyshev_green(mu_n, ene, energy_min, energy_max):
    mu_n = _f(mu_n, np.complex128)
    ene = np.ascontiguousarray(ene, np.float64)
    nk, na, nv = mu_n.shape[2], mu_n.shape[3], len(ene)
    mu_ng = np.zeros_like(mu_n, order="F")
    g0 = np.zeros((18, 18, nv, na), np.complex128, order="F")
    _post().orc_chebyshev_green(_p(mu_n), na, (nk - 2) // 2, _p(ene), nv, energy_min, energy_max, _p(mu_ng), _p(g0))
    return mu_ng, g0


def bprldos(e, a, b2, edges):
    a = np.ascontiguousarray(a, np.float64); b2 = np.ascontiguousarray(b2, np.float64)
    ed = np.ascontiguousarray(edges, np.float64)
    return _post().orc_bprldos(e, _p(a), _p(b2), len(a), _p(ed))


def density(a, b2, ene, dw_l, cshi):
    """a, b2: (lld,18) of one (atom, direction); returns tdens (18,nv)."""
    a = _f(a, np.float64); b2 = _f(b2, np.float64)
    ene = np.ascontiguousarray(ene, np.float64)
    dw = np.ascontiguousarray(dw_l, np.float64); cs = np.ascontiguousarray(cshi, np.float64)
    td = np.zeros((18, len(ene)), order="F")
    _post().orc_density(_p(a), _p(b2), a.shape[0], _p(ene), len(ene), _p(dw), _p(cs), _p(td))
    return td


def sgreen(a, b2, nmdir, ene, dw_l, cshi):
    """a, b2: (lld,18,na,3) like recursion%a; dw_l, cshi: (18,na); returns g0 (18,18,nv,na)."""
    a = _f(a, np.float64); b2 = _f(b2, np.float64)
    ene = np.ascontiguousarray(ene, np.float64)
    dw = _f(dw_l, np.float64); cs = _f(cshi, np.float64)
    lld, na, nv = a.shape[0], a.shape[2], len(ene)
    g0 = np.zeros((18, 18, nv, na), np.complex128, order="F")
    _post().orc_sgreen(_p(a), _p(b2), lld, na, nmdir, _p(ene), nv, _p(dw), _p(cs), _p(g0))
    return g0


def gamma_nm(ene, M, energy_min, energy_max):
    ene = np.ascontiguousarray(ene, np.float64)
    g = np.zeros((len(ene), M, M), np.complex128, order="F")
    _post().orc_gamma_nm(_p(ene), len(ene), M, energy_min, energy_max, _p(g))
    return g


def conductivity_integrand(mu_nm, ene, energy_min, energy_max, per_type):
    mu = _f(mu_nm, np.complex128)
    ene = np.ascontiguousarray(ene, np.float64)
    M, nloop, nv = mu.shape[2], mu.shape[4], len(ene)
    integ = np.zeros((18, nv), np.complex128, order="F")
    integ_at = np.zeros((18, nv, nloop), np.complex128, order="F")
    _post().orc_conductivity_integrand(_p(mu), M, nloop, _p(ene), nv, energy_min, energy_max, int(per_type), _p(integ),
                                       _p(integ_at))
    return integ, integ_at


# ---- neighbour-table construction (rsrec_oracle_lattice.c) ----------------------------------------
def build_nn(crd, no, iu, ct, pbc=None, nrep=(1, 1, 1), a=None, alat=1.0, ncols=None):
    """lattice%nncal + lattice%remd.  crd (3,kk) = cr*alat; no (kk) bravais type; iu (ntot) representative sites;
    pbc: None (open cluster) or 3 flags with nrep = (n1,n2,n3) and a (3,3) = lattice%a.  -> nn (kk, nm+1), nm."""
    L = lib()
    L.orc_build_nn.argtypes = [C.c_int, c_vp, c_vp, C.c_int, c_vp, C.c_double, C.c_int, c_vp, c_vp, c_vp, C.c_double,
                               C.c_int, c_vp, c_vp]
    crd = np.asfortranarray(crd, dtype=np.float64)
    kk = crd.shape[1]
    no = np.ascontiguousarray(no, dtype=np.int32); iu = np.ascontiguousarray(iu, dtype=np.int32)
    b = np.ascontiguousarray(pbc if pbc is not None else (0, 0, 0), dtype=np.int32)
    nr = np.ascontiguousarray(nrep, dtype=np.int32)
    av = np.asfortranarray(a if a is not None else np.eye(3), dtype=np.float64)
    nm = C.c_int(0)
    cols = ncols or 1
    while True:
        nn = np.zeros((kk, cols), np.int32, order="F")
        rc = L.orc_build_nn(kk, _p(crd), _p(no), len(iu), _p(iu), float(ct), int(pbc is not None), _p(b), _p(nr), _p(av),
                            float(alat), cols, _p(nn), C.byref(nm))
        if rc == -1 and ncols is None:
            cols = nm.value + 1
            continue
        return nn, nm.value, rc


# ---- `type bands` (bands.f90; rsrec_oracle_bands.c) -------------------------------------------------------------------
def _bands():
    L = lib()
    if not getattr(L, "_bands_ready", False):
        d, i = C.c_double, C.c_int
        L.orc_bands_dos.argtypes = [c_vp, i, i, c_vp, c_vp, c_vp]
        L.orc_bands_fermi.argtypes = [c_vp, i, d, d, d, i, C.POINTER(d), C.POINTER(i), C.POINTER(d), C.POINTER(i)]
        L.orc_simpson_m.argtypes = [d, d, i, c_vp, d, i, c_vp]
        L.orc_simpson_m.restype = d
        L.orc_bands_magnetic_moments.argtypes = [c_vp, i, i, c_vp, d, d, i, d, c_vp, c_vp]
        L.orc_bands_moments.argtypes = [c_vp, i, i, i, c_vp, c_vp, c_vp, d, d, i, d, c_vp, c_vp]
        L._bands_ready = True
    return L


def e_mesh_full(energy_min, energy_max, channels_ldos, fermi):
    """energy%e_mesh (energy.f90:175-208) -> dict(ene, edel, nv1, channels_ldos)"""
    if channels_ldos % 2 == 0:
        nv1 = channels_ldos + 1
    else:
        nv1 = channels_ldos
        channels_ldos -= 1
    edel = (energy_max - energy_min) / channels_ldos
    r = (fermi - energy_min) / edel
    edel = (fermi - energy_min) / (np.sign(r) * np.floor(abs(r) + 0.5))
    return {"ene": energy_min + edel * np.arange(channels_ldos + 10, dtype=np.float64), "edel": float(edel), "nv1": nv1,
            "channels_ldos": channels_ldos}


def l_spherical():
    """hcpx(L_x), hcpx(L_y), hcpx(L_z) as calculate_orbital_moments forms them (bands.f90:1094-1101) -> (9,9,3)"""
    from . import ham_oracle as HO
    from .ref_bccfe import _L_matrices
    return np.asfortranarray(np.stack([HO.hcpx_cart2sph(m) for m in _L_matrices()], axis=2))


def bands_dos(g0):
    L = _bands()
    g0 = _f(g0, np.complex128)
    nv, nu = g0.shape[2], g0.shape[3]
    dtot = np.zeros(nv); dosia = np.zeros((nv, nu), order="F"); dosial = np.zeros((18, nv, nu), order="F")
    L.orc_bands_dos(_p(g0), nv, nu, _p(dtot), _p(dosia), _p(dosial))
    return dtot, dosia, dosial


def bands_fermi(dtot, edel, energy_min, qqv, fermi, ik1, fix_fermi=False):
    """-> fermi, nv1, e1, ifail"""
    L = _bands()
    dtot = np.ascontiguousarray(dtot, dtype=np.float64)
    f, n, e, fl = C.c_double(fermi), C.c_int(ik1), C.c_double(0.0), C.c_int(0)
    L.orc_bands_fermi(_p(dtot), len(dtot), float(edel), float(energy_min), float(qqv), int(fix_fermi), C.byref(f), C.byref(n),
                      C.byref(e), C.byref(fl))
    return f.value, n.value, e.value, fl.value


def simpson_m(h, ef, npts, y, ea, nexp, ene):
    L = _bands()
    y = np.ascontiguousarray(y, dtype=np.float64); ene = np.ascontiguousarray(ene, dtype=np.float64)
    return L.orc_simpson_m(float(h), float(ef), int(npts), _p(y), float(ea), int(nexp), _p(ene))


def bands_magnetic_moments(g0, ene, edel, fermi, nv1, e1):
    L = _bands()
    g0 = _f(g0, np.complex128); ene = np.ascontiguousarray(ene, dtype=np.float64)
    nv, nu = g0.shape[2], g0.shape[3]
    m0 = np.zeros((3, nu), order="F"); m1 = np.zeros((3, nu), order="F")
    L.orc_bands_magnetic_moments(_p(g0), nv, nu, _p(ene), float(edel), float(fermi), int(nv1), float(e1), _p(m0), _p(m1))
    return m0, m1


def bands_moments(g0, channels_ldos, mom, ene, edel, fermi, nv1, e1):
    """-> occ (3,6,nunits) = sgef, pmef, smef; lmom (3,nunits)"""
    L = _bands()
    g0 = _f(g0, np.complex128); ene = np.ascontiguousarray(ene, dtype=np.float64)
    nv, nu = g0.shape[2], g0.shape[3]
    mom = _f(mom, np.float64); lsph = l_spherical()
    occ = np.zeros((3, 6, nu), order="F"); lmom = np.zeros((3, nu), order="F")
    L.orc_bands_moments(_p(g0), nv, int(channels_ldos), nu, _p(mom), _p(lsph), _p(ene), float(edel), float(fermi), int(nv1),
                        float(e1), _p(occ), _p(lmom))
    return occ, lmom


def intersite_gf(g0, pairs):
    """calculate_intersite_gf (green.f90:425-469): g0 (18,18,nv,4*njij) in the four-slot layout, pairs (njij,2)
    -> gij, gji (18,18,nv,njij), gspin (9,9,nv,njij,8) = Ginmag, Gix, Giy, Giz, Gjnmag, Gjx, Gjy, Gjz"""
    L = lib()
    L.orc_intersite_gf.argtypes = [c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
    g0 = _f(g0, np.complex128)
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    nv, njij = g0.shape[2], len(pairs)
    assert g0.shape[3] == 4 * njij
    pi, pj = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
    gij = np.zeros((18, 18, nv, njij), np.complex128, order="F"); gji = np.zeros_like(gij, order="F")
    gs = np.zeros((9, 9, nv, njij, 8), np.complex128, order="F")
    L.orc_intersite_gf(_p(g0), nv, njij, _p(pi), _p(pj), _p(gij), _p(gji), _p(gs))
    return gij, gji, gs


def conductivity_cumulative(integrand, integrand_at, nv1, wscale, loop_over):
    """tail of calculate_conductivity_tensor (conductivity.f90:300-372), the literal O(nv^2) simpson_f loop:
    integrand (18,nv), integrand_at (18,nv,nat) or None -> sigma (2,19,nv,1+nat)"""
    L = lib()
    L.orc_conductivity_cumulative.argtypes = [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp]
    integrand = _f(integrand, np.complex128)
    nv = integrand.shape[1]
    nat = 0 if integrand_at is None else integrand_at.shape[2]
    iat = None if integrand_at is None else _f(integrand_at, np.complex128)
    ws = np.ascontiguousarray(wscale, dtype=np.float64)
    sigma = np.zeros((2, 19, nv, 1 + nat), order="F")
    L.orc_conductivity_cumulative(_p(integrand), _p(iat) if nat else None, nv, int(nv1), nat, _p(ws), int(loop_over), _p(sigma))
    return sigma
