"""Parity at the FULL sizes of BASELINE.json configs 2, 3 and 4, CUDA path through the C ABI against the
CPU oracle on the same seeded inputs, with the north-star tolerances (a_n/b_n 1e-10, mu 1e-9 relative).  The oracle needs
a few seconds per case on the GPU box's host cores; for config 4 (cond_ll^2 contractions over 8000 sites take minutes on a
CPU) it restates every chain and contracts a subset of the left indices, all other entries are covered by properties."""
import numpy as np
import pytest

from rslmtoasa_b200 import synthetic as S
from tests.cases import relerr, EMIN, EMAX

pytestmark = pytest.mark.gpu

TOL_AB = 1e-10
TOL_MU = 1e-9


def _rec(lat, ham, **kw):
    from rslmtoasa_b200 import Recursion, Control, Energy
    ctl = Control(**{k: v for k, v in kw.items() if k in ("lld", "cond_ll", "cond_calctype")})
    extra = {k: v for k, v in kw.items() if k in ("atlist", "phases")}
    return Recursion(ham, lat, ctl, Energy(EMIN, EMAX), **extra)


def test_config2_surface_full_size(oracle_mod):
    """config 2: fcc sphere r^2 = 100 -> 16756 sites, 7 layer types, 19 slots, 6 recursion sites in one batch, lld = 21."""
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    assert lat.kk == 16756 and lat.ncols == 19
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260102)
    rec = _rec(lat, ham, lld=21)
    rec.recur_b()
    a_b, b2_b = oracle_mod.Oracle(lat, ham).lanczos_block(lat.irec, 21)
    assert relerr(rec.a_b, a_b) < TOL_AB
    assert relerr(rec.b2_b, b2_b) < TOL_AB
    rec.close()


@pytest.mark.parametrize("hoh", [False, True])
def test_config1_collinear_full_size(oracle_mod, hoh, monkeypatch):
    """config 1 (bcc sphere r^2 = 80 -> 5984 sites, lld = 21) with collinear blocks -- what every bccFe regression case of the
    reference runs: the spin-resolved SpMV with 8 consumer warps (k_apply_dmma_sd8, Lanczos A-products inside) against the
    oracle, against its 4-warp form (RSREC_SD_WARPS=4) and, for the Chebyshev moments, the same"""
    lat = S.sphere_cluster("bcc", 80.0)
    assert lat.kk == 5984
    ham = S.make_hamiltonian(lat, seed=20260101, spin_orbit=False, hoh=hoh)
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block(lat.irec, 21)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = orc.cheb_moments(lat.irec, 21, a, b)
    got = {}
    for warps in ("8", "4"):
        monkeypatch.setenv("RSREC_SD_WARPS", warps)
        rec = _rec(lat, ham, lld=21)
        rec.recur_b()
        rec.chebyshev_recur()
        assert rec._L.rsrec_spin_diag_launch_count(rec._h) > 0
        assert relerr(rec.a_b, a_b) < TOL_AB and relerr(rec.b2_b, b2_b) < TOL_AB
        assert relerr(rec.mu_n, mu) < TOL_MU
        got[warps] = (rec.a_b.copy(), rec.b2_b.copy(), rec.mu_n.copy())
        rec.close()
    for x, y in zip(got["8"], got["4"]):
        assert relerr(x, y) < 1e-12


@pytest.mark.parametrize("hoh", [False, True])
def test_config3_impurity_full_size(oracle_mod, hoh):
    """config 3: B2 sphere r^2 = 60 -> 3838 sites, 3 types, 15 site-indexed (hall) sites, lld = 21."""
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    assert lat.kk == 3838 and lat.nmax == 15
    ham = S.make_hamiltonian(lat, seed=20260103, hoh=hoh)
    rec = _rec(lat, ham, lld=21)
    rec.recur_b()
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block(lat.irec, 21)
    assert relerr(rec.a_b, a_b) < TOL_AB
    assert relerr(rec.b2_b, b2_b) < TOL_AB
    if not hoh:   # the Chebyshev path of the same case
        rec.control.lld = 30
        rec.chebyshev_recur()
        a, b = oracle_mod.cheb_scale(EMIN, EMAX)
        mu, _ = orc.cheb_moments(lat.irec, 30, a, b)
        assert relerr(rec.mu_n, mu) < TOL_MU
    rec.close()


def test_config4_conductivity_random_full_size(oracle_mod):
    """config 4: bcc PBC 8000 sites, cond_ll = 50, R = 8 random vectors (compute_moments_stochastic, random_vec)."""
    lat = S.periodic_bcc(10, 20, 20)
    assert lat.kk == 8000
    ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    M, R = 50, 8
    ph = S.random_phases(lat.kk, R, seed=20260104)
    rec = _rec(lat, ham, lld=21, cond_ll=M, cond_calctype="random_vec", phases=ph)
    rec.compute_moments_stochastic()
    mu = rec.mu_nm_stochastic
    assert mu.shape == (18, 18, M, M, R) and np.isfinite(mu).all()
    # oracle: both chains of vectors 1 and R in full, contracted against four left indices (all right indices n)
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    msel = np.array([1, 2, M // 2, M], dtype=np.int32)
    for v in (0, R - 1):
        ref = orc.kubo_moments_cols(M, a, b, msel, phases=ph[:, v:v + 1])[..., 0]
        scale = np.abs(ref).max()
        assert np.abs(mu[:, :, :, msel - 1, v] - ref).max() / scale < TOL_MU
    # every vector: the fused conductivity path contracts only the diagonals mu(l,l,n,m) with a different kernel
    # (k_kubo_diag); its integrand must equal the integrand of the full moments
    from rslmtoasa_b200 import Conductivity
    rec.en.channels_ldos = 400
    con = Conductivity(rec)
    i_full, _ = con.calculate_conductivity_tensor()
    i_full = i_full.copy()
    i_fused, _ = con.compute_conductivity()
    ok = np.isfinite(i_full)
    assert ok.sum() > 0.9 * ok.size
    assert np.abs(i_fused[ok] - i_full[ok]).max() / np.abs(i_full[ok]).max() < 1e-10
    rec.close()


def _bcc_shell_lattice(r2, nshells):
    """bcc sphere with the first `nshells` neighbour shells in the table (8, 6, 12, 24, 8, 6, 24, 24 ... neighbours)."""
    pts = S._sphere_points("bcc", r2)
    if len(pts) % 2 == 1:
        pts = pts[:-1]
    g = np.arange(-6, 7)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    d = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    par = d & 1
    d = d[(par[:, 0] == par[:, 1]) & (par[:, 1] == par[:, 2])]
    n2 = (d ** 2).sum(axis=1)
    shells = np.unique(n2)[1:nshells + 1]
    keep = d[np.isin(n2, shells)]
    keep = keep[np.lexsort((keep[:, 2], keep[:, 1], keep[:, 0], (keep ** 2).sum(axis=1)))]
    disp = np.concatenate([np.zeros((1, 3), np.int64), keep]).astype(np.int64)
    nn = S._nn_from_points(pts, disp)
    return S.Lattice(kk=len(pts), nn=nn, iz=np.ones(len(pts), np.int32), ntype=1, nmax=0, irec=np.array([1], np.int32),
                     cr=pts.T.copy(), disp=disp)


@pytest.mark.parametrize("hoh", [False, True])
def test_long_neighbour_lists_run_the_whole_call_in_one_family(oracle_mod, hoh):
    """ncols beyond the tensor pipeline's stage list (DM_MAXST = 48): the whole call must run in the SIMT family -- mixing
    the families inside one recursion gave A = 0 and truncated moments (round-1 advisor finding)."""
    lat = _bcc_shell_lattice(5.0, 5)     # 1 + 8 + 6 + 12 + 24 + 8 = 59 slots
    assert lat.ncols == 59
    ham = S.make_hamiltonian(lat, seed=11, sigma=0.02, hoh=hoh)
    rec = _rec(lat, ham, lld=6)
    rec.recur_b()
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block(lat.irec, 6)
    assert np.abs(a_b[:, :, 0, 0]).max() > 1e-3          # A_1 is the on-site block, certainly not zero
    assert relerr(rec.a_b, a_b) < TOL_AB
    assert relerr(rec.b2_b, b2_b) < TOL_AB
    rec.chebyshev_recur()
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = orc.cheb_moments(lat.irec, 6, a, b)
    assert relerr(rec.mu_n, mu) < TOL_MU
    rec.close()


def test_many_units_are_split_into_batches_that_fit(oracle_mod, monkeypatch):
    """unit batches are sized from the real work-vector count (psi, pmn, hpsi [+ hoh scratch]); a forced small batch must give
    the same coefficients as one batch."""
    from tests.cases import case
    lat, ham = case("bulk_hoh")
    lat.irec = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    rec = _rec(lat, ham, lld=6)
    rec.recur_b()
    one = rec.a_b.copy()
    monkeypatch.setenv("RSREC_UNIT_BATCH", "2")
    rec.recur_b()
    assert relerr(rec.a_b, one) < 1e-13
    rec.close()


@pytest.mark.parametrize("hoh", [False, True])
def test_large_site_indexed_region(oracle_mod, hoh):
    """a local (hall) region of hundreds of sites: site-indexed sites are packed two per tile (one per half tile, each half
    with the Hamiltonian class of its site) -- odd and even counts, region larger than the bulk remainder of a class"""
    for nmax in (301, 64):
        lat = S.sphere_cluster("bcc", 24.0, ntype=3, nmax=nmax, type_rule="b2")
        assert lat.kk > 2 * nmax
        lat.irec = np.array([1, nmax, nmax + 1], dtype=np.int32)     # starts inside, at the edge of, and outside the region
        ham = S.make_hamiltonian(lat, seed=20260103 + nmax, hoh=hoh)
        rec = _rec(lat, ham, lld=7)
        rec.recur_b()
        orc = oracle_mod.Oracle(lat, ham)
        a_b, b2_b = orc.lanczos_block(lat.irec, 7)
        assert relerr(rec.a_b, a_b) < TOL_AB and relerr(rec.b2_b, b2_b) < TOL_AB
        rec.chebyshev_recur()
        a, b = oracle_mod.cheb_scale(EMIN, EMAX)
        mu, _ = orc.cheb_moments(lat.irec, 7, a, b)
        assert relerr(rec.mu_n, mu) < TOL_MU
        rec.close()
