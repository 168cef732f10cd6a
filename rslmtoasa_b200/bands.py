"""Host-side mirror of the reference's `type bands` (bands.f90) on top of the C ABI: what the SCF loop takes from the
on-site Green function -- total DOS, Fermi level, magnetic / orbital moments, band moments (charges) and the band
energy.  g0 stays on the device: every Green-function call of `Green` leaves it there, and only O(nv) + O(nunits)
numbers come back.

    b = Bands(green, qqv)                # bands(green_obj); qqv = total valence (bands.f90:252)
    b.calculate_fermi()           -> b.dtot, en.fermi, b.nv1, b.e1          bands.f90:227-347
    b.calculate_magnetic_moments()-> b.mx/my/mz (mom0), b.mom1, b.mtot, b.mom    791-855
    b.calculate_moments()         -> b.ql (3,0:2,2,n), b.gravity_center, b.lmom  409-524 (+ 1075-1156)
    b.calculate_band_energy()     -> b.eband                                 354-359

Multi-rank: dtot is all-reduced (bands.f90:276) and the per-unit results all-gathered, through `parallel.py`.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib, parallel
from .green import _Consumer, _p


class Bands(_Consumer):
    def __init__(self, green, qqv: float, nsp: int = 2, vmad=None, device=None):
        super().__init__(green.recursion)
        self.green = green
        self.qqv = float(qqv)
        self.nsp = nsp
        self.vmad = vmad                 # potential%vmad per unit (gravity_center is measured from it), default 0
        self.device = device             # torch device for the collectives (None = host tensors / gloo)
        self.nv1, self.e1 = self.en.nv1 or self.en.ik1, 0.0
        self.dtot = self.dosia = self.dosial = None
        self.eband = None
        self.mom = None                  # (3, n_local) unit vectors potential%mom

    # -- helpers ------------------------------------------------------------------------------------------------
    def _shape(self):
        nu, nv = C.c_int(0), C.c_int(0)
        _lib.check(self._L.rsrec_bands_g0_shape(self._h, C.byref(nu), C.byref(nv)))
        return nu.value, nv.value

    def _n_global(self):
        return len(self.recursion.lattice.irec)

    def set_g0(self, g0):
        """use a g0 (18,18,nv,n_local) computed elsewhere (the staged reference flow keeps green%g0 on the host)"""
        g0 = np.asfortranarray(g0, dtype=np.complex128)
        _lib.check(self._L.rsrec_bands_set_g0(self._h, _p(g0), g0.shape[3], g0.shape[2]))

    # -- bands.f90:227-347 --------------------------------------------------------------------------------------
    def calculate_fermi(self, ldos: bool = False):
        nu, nv = self._shape()
        dtot = np.zeros(nv)
        if ldos:
            self.dosia = np.zeros((nv, nu), order="F"); self.dosial = np.zeros((18, nv, nu), order="F")
        _lib.check(self._L.rsrec_bands_dos(self._h, _p(dtot), _p(self.dosia) if ldos else None, _p(self.dosial) if ldos else None))
        self.dtot = parallel.allreduce_sum(dtot, self.device)
        fermi, nv1, e1, ifail = C.c_double(self.en.fermi), C.c_int(self.en.ik1), C.c_double(0.0), C.c_int(0)
        _lib.check(self._L.rsrec_bands_fermi(self._h, _p(self.dtot), nv, self.en.edel, self.en.energy_min, self.qqv,
                                             int(self.en.fix_fermi), C.byref(fermi), C.byref(nv1), C.byref(e1), C.byref(ifail)))
        self.ifail = ifail.value
        if self.ifail == 0:
            self.en.fermi, self.nv1, self.e1 = fermi.value, nv1.value, e1.value
        return self.en.fermi

    def calculate_band_energy(self):
        eb = C.c_double(0.0)
        ene = self.ene
        _lib.check(self._L.rsrec_bands_band_energy(self._h, _p(self.dtot), len(self.dtot), _p(ene), self.en.edel, self.en.fermi,
                                                   self.nv1, self.e1, C.byref(eb)))
        self.eband = eb.value
        return self.eband

    # -- bands.f90:791-855 --------------------------------------------------------------------------------------
    def calculate_magnetic_moments(self):
        nu, _ = self._shape()
        m0 = np.zeros((3, nu), order="F"); m1 = np.zeros((3, nu), order="F")
        ene = self.ene
        _lib.check(self._L.rsrec_bands_magnetic_moments(self._h, _p(ene), self.en.edel, self.en.fermi, self.nv1, self.e1, _p(m0), _p(m1)))
        self.mx, self.my, self.mz = m0
        self.mom0, self.mom1 = m0, m1
        self.mtot = np.sqrt(m0[0] ** 2 + m0[1] ** 2 + m0[2] ** 2) + 1.0e-15
        self.mom = np.asfortranarray(m0 / self.mtot)
        if self.nsp < 3:
            self.mom[:] = np.array([0.0, 0.0, 1.0])[:, None]
        return self.mom0

    # -- bands.f90:409-524 --------------------------------------------------------------------------------------
    def calculate_moments(self):
        nu, _ = self._shape()
        if self.mom is None:
            self.mom = np.asfortranarray(np.tile(np.array([0.0, 0.0, 1.0])[:, None], (1, nu)))
        occ = np.zeros((3, 6, nu), order="F"); lmom = np.zeros((3, nu), order="F")
        ene = self.ene
        mom = np.asfortranarray(self.mom, dtype=np.float64)
        _lib.check(self._L.rsrec_bands_moments(self._h, self.en.channels_ldos, _p(ene), self.en.edel, self.en.fermi, self.nv1,
                                               self.e1, _p(mom), _p(occ), _p(lmom)))
        self.occ, self.lmom = occ, lmom
        sgef, pmef, smef = occ[0], occ[1], occ[2]                       # (6, nu)
        vmad = np.zeros(nu) if self.vmad is None else np.asarray(self.vmad, dtype=np.float64)
        cg = pmef / sgef
        self.gravity_center = np.asfortranarray((cg - vmad[None, :]).reshape(2, 3, nu).transpose(1, 0, 2))   # (l, spin, unit)
        ql = np.zeros((3, 3, 2, nu), order="F")                          # ql(1:3, 0:2, 1:2)
        ql[0] = sgef.reshape(2, 3, nu).transpose(1, 0, 2)
        ql[2] = (smef - 2.0 * cg * pmef + cg ** 2 * sgef).reshape(2, 3, nu).transpose(1, 0, 2)
        self.ql = ql
        return ql

    # -- all ranks' units in global order (the MPI_ALLREDUCE of the flattened potentials, bands.f90:502-512) ---------
    def gather(self, arr):
        return parallel.allgather_units(np.asfortranarray(arr), self._n_global(), self.device)
