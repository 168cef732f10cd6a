"""CPU tests of the neighbour-table oracle (nncal + remd, lattice.f90:3035-3123 / 2823-2907): pinned by an independent
numpy restatement (brute-force distance matrix, argsort-free construction) and by structural invariants."""
import numpy as np
import pytest

from rslmtoasa_b200 import synthetic as S

ALAT = 5.42


def numpy_nn(crd, no, iu, ct, cell=None):
    """neighbours of i = ascending j with |min-image(r_j - r_i)|^2 < ct^2; slots ordered like the representative's."""
    kk = crd.shape[1]
    d = crd[:, None, :] - crd[:, :, None]                       # d[:, i, j] = r_j - r_i
    if cell is not None:
        frac = np.linalg.solve(cell, d.reshape(3, -1)).reshape(3, kk, kk)
        d = (cell @ (frac - np.rint(frac)).reshape(3, -1)).reshape(3, kk, kk)
    r2 = (d ** 2).sum(0)
    adj = (r2 < ct * ct) & ~np.eye(kk, dtype=bool)
    lists = [np.nonzero(adj[i])[0] for i in range(kk)]
    nm = max(len(l) for l in lists) + 1
    sign = 1.0 if cell is not None else -1.0                     # open clusters store r_i - r_j, periodic ones r_j - r_i
    nn = np.zeros((kk, nm + 1), np.int32)
    for i in range(kk):
        rep = iu[no[i] - 1] - 1
        vecs = sign * d[:, rep, lists[rep]]
        nn[i, 0] = len(lists[rep]) + 1
        for j in lists[i]:
            k = np.nonzero(((sign * d[:, i, j][:, None] - vecs) ** 2).sum(0) < 1e-4)[0]
            assert len(k) >= 1
            nn[i, 1 + k[0]] = j + 1
    return nn, nm


def _bcc_sphere():
    lat = S.sphere_cluster("bcc", 6.0)
    return 0.5 * lat.cr.astype(np.float64) * ALAT, lat


def _pbc_bcc(nx, ny, nz):
    pts = [(x + s, y + s, z + s) for z in range(nz) for y in range(ny) for x in range(nx) for s in (0.0, 0.5)]
    crd = np.array(pts, dtype=np.float64).T * ALAT
    return crd, np.eye(3), (nx, ny, nz)


def test_open_bcc_cluster_vs_numpy(oracle_mod):
    crd, lat = _bcc_sphere()
    no, iu = np.ones(lat.kk, np.int32), [1]
    nn, nm, rc = oracle_mod.build_nn(crd, no, iu, 1.1 * ALAT)
    ref, rnm = numpy_nn(crd, no, iu, 1.1 * ALAT)
    assert rc == 0 and nm == rnm == 15 and np.array_equal(nn, ref)
    assert (nn[:, 0] == 15).all() and nn[0, 1:15].tolist() == sorted(nn[0, 1:15].tolist())   # centre: ascending (nncal order)
    # same neighbour sets as the synthetic generator used everywhere else (slot order differs by construction)
    for i in range(lat.kk):
        mine = set(nn[i, 1:][nn[i, 1:] > 0])
        theirs = lat.nn[i, 1:lat.nn[i, 0]]
        assert mine == set(theirs[theirs > 0])


def test_two_bravais_types_and_fcc(oracle_mod):
    lat = S.sphere_cluster("fcc", 5.0)
    crd = 0.5 * lat.cr.astype(np.float64) * ALAT
    no = np.ones(lat.kk, np.int32)
    nn, nm, rc = oracle_mod.build_nn(crd, no, [1], 1.05 * ALAT)
    ref, _ = numpy_nn(crd, no, [1], 1.05 * ALAT)
    assert rc == 0 and nm == 19 and np.array_equal(nn, ref)
    crd, lat = _bcc_sphere()                                    # B2: corner and body-centre sublattices as two types
    no = (1 + (lat.cr[0] % 2)).astype(np.int32)
    iu = [1, int(np.nonzero(no == 2)[0][0]) + 1]
    nn, nm, rc = oracle_mod.build_nn(crd, no, iu, 1.1 * ALAT)
    ref, _ = numpy_nn(crd, no, iu, 1.1 * ALAT)
    assert rc == 0 and np.array_equal(nn, ref)


def test_periodic_bcc_vs_numpy_and_reciprocity(oracle_mod):
    crd, a, nrep = _pbc_bcc(4, 3, 3)
    kk = crd.shape[1]
    no = np.ones(kk, np.int32)
    nn, nm, rc = oracle_mod.build_nn(crd, no, [1], 1.1 * ALAT, pbc=(1, 1, 1), nrep=nrep, a=a, alat=ALAT)
    cell = np.diag(np.array(nrep, float) * ALAT)
    ref, _ = numpy_nn(crd, no, [1], 1.1 * ALAT, cell)
    assert rc == 0 and nm == 15 and np.array_equal(nn, ref)
    assert (nn[:, 1:15] > 0).all()                               # no missing neighbours under PBC
    for i in range(kk):                                          # i in nn(j) <=> j in nn(i)
        for j in nn[i, 1:15]:
            assert i + 1 in nn[j - 1, 1:15]


def test_errors(oracle_mod):
    crd, lat = _bcc_sphere()
    no = np.ones(lat.kk, np.int32)
    nn, nm, rc = oracle_mod.build_nn(crd, no, [1], 1.1 * ALAT, ncols=4)
    assert rc == -1 and nm == 15                                 # caller's table too narrow: nm reported
    nn, nm, rc = oracle_mod.build_nn(crd, no, [lat.kk], 1.1 * ALAT)
    assert rc == -2                                              # an edge site as representative: "VECTOR NOT FOUND"
