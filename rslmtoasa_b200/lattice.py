"""Host-side mirror of the neighbour-table part of the reference's `type lattice` (lattice.f90:1835-1870):
`nncal` + `remd` run on the GPU through the C ABI (`rsrec_build_nn`)."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_nn(crd, no, iu, ct, pbc=None, nrep=(1, 1, 1), a=None, alat=1.0, device: int = 0):
    """-> nn (kk, nm+1) int32 (Fortran order), nm.  Arguments as lattice%nncal / lattice%remd hold them:
    crd (3,kk) = cr*alat, no (kk) = num, iu (ntot) representatives, ct = ct(1); pbc = (b1,b2,b3) or None."""
    L = _lib.load()
    crd = np.asfortranarray(crd, dtype=np.float64)
    kk = crd.shape[1]
    no = np.ascontiguousarray(no, dtype=np.int32)
    iu = np.ascontiguousarray(iu, dtype=np.int32)
    b = None if pbc is None else np.ascontiguousarray(pbc, dtype=np.int32)
    nr = np.ascontiguousarray(nrep, dtype=np.int32)
    av = None if a is None else np.asfortranarray(a, dtype=np.float64)
    nm = C.c_int(0)
    rc = L.rsrec_build_nn(device, kk, _p(crd), _p(no), len(iu), _p(iu), float(ct), _p(b), _p(nr), _p(av), float(alat), 0,
                          None, C.byref(nm))
    if rc != 0 and nm.value == 0:
        _lib.check(rc)
    nn = np.zeros((kk, nm.value + 1), np.int32, order="F")
    _lib.check(L.rsrec_build_nn(device, kk, _p(crd), _p(no), len(iu), _p(iu), float(ct), _p(b), _p(nr), _p(av), float(alat),
                                nn.shape[1], _p(nn), C.byref(nm)))
    return nn, nm.value
