"""Parity of the device-side consumers of the recursion results against the CPU oracle, through
the C ABI: terminator (bit-exact: it is branchy integer-like work on doubles), block / Chebyshev Green functions,
scalar continued-fraction DOS, Kubo-Bastin integrand (floating point, tolerances below, relative to the largest entry).
"""
import numpy as np
import pytest

from tests.cases import case, relerr, EMIN, EMAX

pytestmark = pytest.mark.gpu

TOL_G = 1e-10      # Green functions: 20 levels of 18x18 inversions; measured ~1e-13
TOL_SUM = 1e-12    # plain weighted sums (chebyshev_green, conductivity integrand)


def _rec(lat, ham, **kw):
    from rslmtoasa_b200 import Recursion, Control, Energy
    ctl = Control(**{k: v for k, v in kw.items() if k in ("lld", "cond_ll", "cond_calctype", "random_vec_num")})
    en = Energy(EMIN, EMAX, channels_ldos=kw.get("channels", 160), fermi=kw.get("fermi", 0.05))
    extra = {k: v for k, v in kw.items() if k in ("ijpair", "atlist", "phases")}
    return Recursion(ham, lat, ctl, en, **extra)


@pytest.fixture(scope="module")
def block_rec():
    """recur_b + zsqr on the GPU for 4 units of the layered fcc case, lld = 9."""
    lat, ham = case("surface")
    rec = _rec(lat, ham, lld=9)
    rec.recur_b()
    rec.zsqr()
    return rec


def test_e_mesh_is_the_reference_rule(oracle_mod):
    from rslmtoasa_b200 import Energy
    en = Energy(-1.2, 1.0, channels_ldos=161, fermi=0.05)
    assert np.array_equal(en.e_mesh(), oracle_mod.e_mesh(-1.2, 1.0, 161, 0.05)) and en.channels_ldos == 160


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_bpopt_bit_exact(oracle_mod, block_rec, kernel, monkeypatch):
    """both terminator kernels: one warp per chain (31 speculative bisection nodes per round) and one thread per chain"""
    from rslmtoasa_b200 import Dos
    monkeypatch.setenv("RSREC_BPOPT_KERNEL", kernel)
    rng = np.random.default_rng(11)
    ll, n = 14, 200
    a = 0.1 + 0.3 * rng.normal(size=(ll, n))
    rb = 0.3 + 0.1 * rng.normal(size=(ll, n))          # some negative / tiny entries on purpose
    a[:, 0], rb[:, 0] = 0.3, 0.25                        # constant chain
    a[:, 1], rb[:, 1] = 0.0, 0.0                         # dead chain (off-diagonal block entries look like this)
    ainf, rbinf, ifail = Dos(block_rec).bpopt(a, rb)
    for c in range(n):
        oa, ob, of = oracle_mod.bpopt(a[:, c], rb[:, c])
        assert (ainf[c] == oa or (np.isnan(ainf[c]) and np.isnan(oa))) and (rbinf[c] == ob or np.isnan(ob)), c
        assert ifail[c] == of


def test_get_terminf_bit_exact(oracle_mod, block_rec):
    from rslmtoasa_b200 import Green
    a_inf, b_inf, a0, b0 = Green(block_rec).get_terminf()
    oa, ob, oa0, ob0 = oracle_mod.get_terminf(block_rec.a_b, block_rec.b2_b)
    assert np.array_equal(a_inf, oa) and np.array_equal(b_inf, ob)
    assert np.array_equal(a0, oa0) and np.array_equal(b0, ob0)


@pytest.mark.parametrize("sym_term", [False, True])
@pytest.mark.parametrize("eta", [0.0, 0.02j])
def test_bgreen(oracle_mod, block_rec, sym_term, eta):
    from rslmtoasa_b200 import Green
    g = Green(block_rec, sym_term=sym_term)
    a_inf, b_inf, _, _ = g.get_terminf()
    ene = g.ene
    for ia in (1, 3):
        got = g.bgreen(ia, 1, len(ene), a_inf[..., ia - 1], b_inf[..., ia - 1], eta)
        ref = oracle_mod.bgreen(block_rec.a_b[..., ia - 1], block_rec.b2_b[..., ia - 1], ene, a_inf[..., ia - 1],
                                b_inf[..., ia - 1], eta, sym_term)
        assert relerr(got, ref) < TOL_G


def test_bgreen_channel_window(oracle_mod, block_rec):
    from rslmtoasa_b200 import Green
    g = Green(block_rec)
    a_inf, b_inf, _, _ = g.get_terminf()
    full = g.bgreen(2, 1, len(g.ene), a_inf[..., 1], b_inf[..., 1])
    one = g.bgreen(2, 40, 3, a_inf[..., 1], b_inf[..., 1])
    assert np.array_equal(one[:, :, 39:42], full[:, :, 39:42])
    assert not one[:, :, :39].any() and not one[:, :, 42:].any()
    empty = g.bgreen(2, 5, 0, a_inf[..., 1], b_inf[..., 1])
    assert not empty.any()


@pytest.mark.parametrize("sym_term", [False, True])
def test_block_green(oracle_mod, block_rec, sym_term):
    from rslmtoasa_b200 import Green
    g = Green(block_rec, sym_term=sym_term)
    g0 = g.block_green()
    ref = oracle_mod.block_green(block_rec.a_b, block_rec.b2_b, g.ene, sym_term)
    assert g0.shape == ref.shape == (18, 18, 170, 4)
    assert relerr(g0, ref) < TOL_G
    d = np.arange(18)
    assert (-g0[d, d].imag.sum(0)).min() > -1e-8          # Herglotz: non-negative LDOS


def test_block_green_odd_sizes(oracle_mod):
    """mesh lengths that do not fill the last CTA, a single unit, the shortest chain the terminator accepts"""
    from rslmtoasa_b200 import Green
    lat, ham = case("tiny")
    for lld, channels in ((2, 2), (3, 14), (5, 36)):
        rec = _rec(lat, ham, lld=lld, channels=channels, fermi=-0.3)
        rec.recur_b()
        rec.zsqr()
        g = Green(rec)
        g0 = g.block_green()
        ref = oracle_mod.block_green(rec.a_b, rec.b2_b, g.ene)
        both_nan = np.isnan(g0) & np.isnan(ref)
        assert relerr(np.where(both_nan, 0, g0), np.where(both_nan, 0, ref)) < TOL_G


@pytest.mark.parametrize("name,lld", [("bulk", 12), ("pbc", 30)])
def test_chebyshev_green(oracle_mod, name, lld):
    from rslmtoasa_b200 import Green
    lat, ham = case(name)
    rec = _rec(lat, ham, lld=lld, channels=211)
    rec.chebyshev_recur()
    g = Green(rec)
    g0 = g.chebyshev_green()
    mu_ng, ref = oracle_mod.chebyshev_green(rec.mu_n, g.ene, EMIN, EMAX)
    assert relerr(g0, ref) < TOL_SUM
    assert relerr(rec.mu_ng, mu_ng) < 1e-15


def _scalar_rec():
    from rslmtoasa_b200 import synthetic as S
    lat, _ = case("bulk")
    lat.irec = np.array([1, 2, 5], dtype=np.int32)
    ham1 = S.make_hamiltonian(lat, seed=20260101, spin_orbit=False)
    rec = _rec(lat, ham1, lld=10)
    rec.recur()
    return rec


def test_density_and_sgreen(oracle_mod):
    from rslmtoasa_b200 import Green, Dos
    rec = _scalar_rec()
    rng = np.random.default_rng(8)
    dw = 1.0 + 0.05 * rng.normal(size=(18, 3))
    cs = 0.02 * rng.normal(size=(18, 3))
    dos = Dos(rec)
    ene = dos.ene
    for ia in (1, 3):
        td = dos.density(ia, 1, dw[:, ia - 1], cs[:, ia - 1])
        ref = oracle_mod.density(rec.a[:, :, ia - 1, 0], rec.b2[:, :, ia - 1, 0], ene, dw[:, ia - 1], cs[:, ia - 1])
        assert relerr(td, ref) < 1e-13
    g0 = Green(rec).sgreen(dw, cs, 1)
    assert relerr(g0, oracle_mod.sgreen(rec.a, rec.b2, 1, ene, dw, cs)) < 1e-13
    # three quantisation directions (non-collinear scalar path, green.f90:688-699)
    rec.a[..., 1], rec.b2[..., 1] = rec.a[..., 0] * 1.01, rec.b2[..., 0]
    rec.a[..., 2], rec.b2[..., 2] = rec.a[..., 0] * 0.99, rec.b2[..., 0]
    g3 = Green(rec).sgreen(dw, cs, 3)
    assert relerr(g3, oracle_mod.sgreen(rec.a, rec.b2, 3, ene, dw, cs)) < 1e-13
    td_all = dos.density_all(dw, cs, 3)
    assert td_all.shape == (18, len(ene), 3, 3)
    assert relerr(td_all[:, :, 1, 2], oracle_mod.density(rec.a[:, :, 1, 2], rec.b2[:, :, 1, 2], ene, dw[:, 1], cs[:, 1])) < 1e-13


@pytest.mark.parametrize("calctype,M", [("per_type", 7), ("random_vec", 24)])
def test_conductivity_integrand(oracle_mod, calctype, M):
    from rslmtoasa_b200 import Conductivity
    lat, ham = case("pbc")
    rng = np.random.default_rng(5)
    kw = dict(cond_ll=M, cond_calctype=calctype, channels=60, fermi=0.0)
    if calctype == "per_type":
        rec = _rec(lat, ham, atlist=np.array([1, 2], np.int32), **kw)
    else:
        rec = _rec(lat, ham, phases=rng.random((lat.kk, 3)), random_vec_num=3, **kw)
    rec.compute_moments_stochastic()
    cond = Conductivity(rec)
    integ, integ_at = cond.calculate_conductivity_tensor()
    ri, rat = oracle_mod.conductivity_integrand(rec.mu_nm_stochastic, cond.ene, EMIN, EMAX, calctype == "per_type")
    # the 10 extra mesh points of e_mesh lie beyond |w| = 1 on this coarse mesh: acos -> NaN on both sides
    nan = np.isnan(ri)
    assert np.array_equal(np.isnan(integ), nan) and 0 < nan.sum() < nan.size / 2
    assert relerr(np.nan_to_num(integ), np.nan_to_num(ri)) < TOL_SUM
    assert relerr(np.nan_to_num(integ_at), np.nan_to_num(rat)) < TOL_SUM
    if calctype == "random_vec":
        assert not np.nan_to_num(integ_at).any()


def test_fused_recur_b_green_is_the_staged_pipeline(oracle_mod):
    """recur_b -> zsqr -> block_green through host arrays == the fused device-resident call, bit for bit"""
    from rslmtoasa_b200 import Green
    for name in ("surface", "impurity_hoh"):
        lat, ham = case(name)
        rec = _rec(lat, ham, lld=8)
        rec.recur_b()
        a_b, b2_b = rec.a_b.copy(), rec.b2_b.copy()
        rec.zsqr()
        staged = Green(rec).block_green().copy()
        rec2 = _rec(lat, ham, lld=8)
        g0 = Green(rec2).recur_b_green()
        assert np.array_equal(rec2.a_b, a_b) and np.array_equal(rec2.b2_b, b2_b)
        assert np.array_equal(g0, staged)
        orc = oracle_mod.Oracle(lat, ham)
        oa, ob2 = orc.lanczos_block(lat.irec, 8)
        assert relerr(g0, oracle_mod.block_green(oa, orc.zsqr(ob2), Green(rec2).ene)) < 1e-8   # end to end vs the oracle


def test_fused_cheb_recur_green(oracle_mod):
    from rslmtoasa_b200 import Green
    lat, ham = case("surface")
    rec = _rec(lat, ham, lld=10, channels=100)
    rec.chebyshev_recur()
    staged = Green(rec).chebyshev_green().copy()
    rec2 = _rec(lat, ham, lld=10, channels=100)
    g = Green(rec2)
    g0 = g.chebyshev_recur_green()
    assert np.array_equal(rec2.mu_n, rec.mu_n) and np.array_equal(rec2.mu_ng, rec.mu_ng)
    assert np.array_equal(g0, staged, equal_nan=True) and not np.isnan(g0[:, :, :100]).any()   # mesh tail has |w| > 1
    assert np.array_equal(g.chebyshev_recur_green(keep_moments=False), staged, equal_nan=True)


@pytest.mark.parametrize("calctype", ["per_type", "random_vec"])
def test_fused_kubo_conductivity(oracle_mod, calctype):
    from rslmtoasa_b200 import Conductivity
    lat, ham = case("pbc")
    kw = dict(cond_ll=9, cond_calctype=calctype, channels=60, fermi=0.0)
    if calctype == "per_type":
        rec = _rec(lat, ham, atlist=np.array([2, 1], np.int32), **kw)
    else:
        rec = _rec(lat, ham, phases=np.random.default_rng(6).random((lat.kk, 2)), random_vec_num=2, **kw)
    rec.compute_moments_stochastic()
    staged, staged_at = (x.copy() for x in Conductivity(rec).calculate_conductivity_tensor())
    c = Conductivity(rec)
    integ, integ_at = c.compute_conductivity(keep_moments=True)
    assert np.array_equal(integ, staged, equal_nan=True) and np.array_equal(integ_at, staged_at, equal_nan=True)
    assert np.array_equal(c.recursion.mu_nm_stochastic, rec.mu_nm_stochastic)
    # without the moments only the 18 consumed diagonals are contracted (k_kubo_diag): same numbers to rounding
    integ2, at2 = c.compute_conductivity(keep_moments=False)
    assert np.array_equal(np.isnan(integ2), np.isnan(staged))
    assert relerr(np.nan_to_num(integ2), np.nan_to_num(staged)) < 1e-12
    assert relerr(np.nan_to_num(at2), np.nan_to_num(staged_at)) < 1e-12 or not np.nan_to_num(staged_at).any()
    ri, _ = oracle_mod.conductivity_integrand(rec.mu_nm_stochastic, c.ene, EMIN, EMAX, calctype == "per_type")
    assert relerr(np.nan_to_num(integ), np.nan_to_num(ri)) < TOL_SUM


def test_post_argument_errors(block_rec):
    from rslmtoasa_b200 import Green, RsrecError
    g = Green(block_rec)
    a_inf, b_inf, _, _ = g.get_terminf()
    with pytest.raises(RsrecError):
        g.bgreen(1, 0, 3, a_inf[..., 0], b_inf[..., 0])          # channels are 1-based
    with pytest.raises(RsrecError):
        g.bgreen(1, 169, 3, a_inf[..., 0], b_inf[..., 0])        # window runs past the mesh


def test_gpu_against_committed_golden_vectors():
    """the CUDA path against tests/golden/*.npz: hot-path coefficients recomputed on the GPU, then every consumer fed
    with the COMMITTED coefficients so that each stage is compared on identical inputs"""
    import os
    from rslmtoasa_b200 import Green, Dos, Conductivity
    here = os.path.join(os.path.dirname(__file__), "golden")
    g, p = np.load(os.path.join(here, "oracle_golden.npz")), np.load(os.path.join(here, "post_golden.npz"))
    lat, ham = case("impurity_hoh")
    rec = _rec(lat, ham, lld=6)
    rec.en.ene = p["ene"]
    rec.recur_b()
    assert relerr(rec.a_b, g["imp_a_b"]) < 1e-10 and relerr(rec.b2_b, g["imp_b2_b"]) < 1e-10
    rec.chebyshev_recur()
    assert relerr(rec.mu_n, g["imp_mu"]) < 1e-9
    rec.a_b, rec.b2_b = g["imp_a_b"], g["imp_b2_b"].copy()
    rec.zsqr()
    assert relerr(rec.b2_b, p["b_b"]) < 1e-12
    rec.b2_b = p["b_b"]
    gr = Green(rec)
    a_inf, b_inf, a0, b0 = gr.get_terminf()
    assert np.array_equal(a_inf, p["a_inf"]) and np.array_equal(b_inf, p["b_inf"])
    assert np.array_equal(a0, p["a_inf0"]) and np.array_equal(b0, p["b_inf0"])
    assert relerr(gr.block_green(), p["g0_block"]) < TOL_G
    gr.sym_term = True
    assert relerr(gr.block_green(), p["g0_block_sym"]) < TOL_G
    rec.mu_n = g["imp_mu"]
    assert relerr(gr.chebyshev_green(), p["g0_cheb"]) < TOL_SUM and relerr(rec.mu_ng, p["mu_ng"]) < 1e-15
    rec.a = np.zeros((8, 18, 1, 3), order="F"); rec.b2 = np.ones((8, 18, 1, 3), order="F")
    rec.a[..., 0], rec.b2[..., 0] = g["bulk_sa"], g["bulk_sb"]
    assert relerr(Dos(rec).density(1, 1, p["dw"][:, 0], p["cs"][:, 0]), p["tdens"]) < 1e-13
    rec.mu_nm_stochastic = g["pbc_kubo"]
    rec.control.cond_calctype = "per_type"
    integ, integ_at = Conductivity(rec).calculate_conductivity_tensor()
    assert relerr(np.nan_to_num(integ), np.nan_to_num(p["integrand"])) < TOL_SUM
    assert relerr(np.nan_to_num(integ_at), np.nan_to_num(p["integrand_at"])) < TOL_SUM


def test_block_green_full_mesh_size(oracle_mod):
    """the reference's mesh size (channels_ldos + 10 = 2510 energies), lld = 21, 3 units of the layered fcc case"""
    from rslmtoasa_b200 import Green
    lat, ham = case("surface")
    lat.irec = np.array([1, 2, 9], dtype=np.int32)
    rec = _rec(lat, ham, lld=21, channels=2500, fermi=0.0)
    g = Green(rec)
    g0 = g.recur_b_green()
    assert g0.shape == (18, 18, 2510, 3)
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block(lat.irec, 21)
    ref = oracle_mod.block_green(a_b, orc.zsqr(b2_b), g.ene)
    # lld = 21 on a 250-site cluster runs into an exhausted Krylov space: compare where the chain is still stable
    assert relerr(rec.a_b[:, :, :8], a_b[:, :, :8]) < 1e-8
    staged = oracle_mod.block_green(rec.a_b, orc.zsqr(rec.b2_b), g.ene)
    ok = np.isfinite(staged) & np.isfinite(g0)
    assert ok.mean() > 0.99 and relerr(g0[ok], staged[ok]) < 1e-8
    d = np.arange(18)
    ldos = -g0[d, d].imag.sum(0) / np.pi
    assert np.nanmin(ldos) > -1e-6


@pytest.mark.parametrize("seed", range(4))
def test_randomised_consumers(oracle_mod, block_rec, seed):
    """fuzz: random Hermitian A_l / positive definite B_l chains, random moments, random scalar chains and mesh sizes"""
    from rslmtoasa_b200 import Green, Dos, Conductivity, Energy
    rng = np.random.default_rng(300 + seed)
    ll, na = int(rng.integers(2, 12)), int(rng.integers(1, 4))
    channels = int(rng.integers(2, 90)) * 2

    def herm(scale):
        m = rng.normal(size=(18, 18)) + 1j * rng.normal(size=(18, 18))
        return scale * (m + m.conj().T) / 2

    a_b = np.zeros((18, 18, ll, na), complex, order="F"); b_b = np.zeros_like(a_b)
    for u in range(na):
        for l in range(ll):
            a_b[:, :, l, u] = herm(0.1) + np.diag(rng.uniform(-0.3, 0.3, 18))
            h = herm(0.05)
            b_b[:, :, l, u] = h @ h + 0.3 * np.eye(18)                 # Hermitian positive definite, like zsqr's output
    rec = block_rec
    saved = (rec.a_b, rec.b2_b, rec.en)
    try:
        rec.en = Energy(-1.5, 1.5, channels_ldos=channels, fermi=float(rng.uniform(-0.5, 0.5)))
        rec.en.e_mesh()
        rec.a_b, rec.b2_b = a_b, b_b
        for sym in (False, True):
            g = Green(rec, sym_term=sym)
            assert relerr(g.block_green(), oracle_mod.block_green(a_b, b_b, g.ene, sym)) < TOL_G
        a_inf, b_inf, a0, b0 = g.get_terminf()
        oa, ob, oa0, ob0 = oracle_mod.get_terminf(a_b, b_b)
        assert np.array_equal(a_inf, oa, equal_nan=True) and np.array_equal(b_inf, ob, equal_nan=True)
        lld = int(rng.integers(0, 9))
        mu = np.asfortranarray(rng.normal(size=(18, 18, 2 * lld + 2, na)) + 1j * rng.normal(size=(18, 18, 2 * lld + 2, na)))
        rec.mu_n = mu
        g0 = Green(rec).chebyshev_green()
        mu_ng, ref = oracle_mod.chebyshev_green(mu, g.ene, -1.5, 1.5)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(g0), ok) and relerr(g0[ok], ref[ok]) < TOL_SUM
        lls = int(rng.integers(2, 15))
        a = np.asfortranarray(rng.uniform(-0.3, 0.3, size=(lls, 18, na, 3)))
        b2 = np.asfortranarray(rng.uniform(0.02, 0.2, size=(lls, 18, na, 3)))
        rec.a, rec.b2 = a, b2
        dw = 1.0 + 0.05 * rng.normal(size=(18, na)); cs = 0.02 * rng.normal(size=(18, na))
        for nmdir in (1, 3):
            got = Green(rec).sgreen(dw, cs, nmdir)
            assert relerr(got, oracle_mod.sgreen(a, b2, nmdir, g.ene, dw, cs)) < 1e-12
    finally:
        rec.a_b, rec.b2_b, rec.en = saved


@pytest.mark.parametrize("M,nvec", [(1, 1), (5, 2), (16, 1), (17, 1), (35, 2)])
def test_kubo_diagonal_contraction_sizes(oracle_mod, M, nvec):
    """the diagonal-only contraction against the oracle's full moments: cond_ll below, at and above the 16-vector
    blocks of the kernel, several start vectors, hoh and plain"""
    from rslmtoasa_b200 import Conductivity, synthetic as S
    for name in ("pbc", "pbc_hoh"):
        lat, ham = case(name)
        ph = S.random_phases(lat.kk, nvec, seed=77 + M)
        rec = _rec(lat, ham, cond_ll=M, cond_calctype="random_vec", channels=40, fermi=0.0, phases=ph, random_vec_num=nvec)
        c = Conductivity(rec)
        integ, _ = c.compute_conductivity(keep_moments=False)
        a, b = oracle_mod.cheb_scale(EMIN, EMAX)
        mu = oracle_mod.Oracle(lat, ham).kubo_moments(M, a, b, phases=ph)
        ri, _ = oracle_mod.conductivity_integrand(mu, c.ene, EMIN, EMAX, False)
        assert np.array_equal(np.isnan(integ), np.isnan(ri))
        assert relerr(np.nan_to_num(integ), np.nan_to_num(ri)) < 1e-10


# ---- exchange path: calculate_intersite_gf (green.f90:425-469) -----------------------------------------------------------
PAIRS = np.array([[1, 2], [3, 3], [2, 9], [5, 1]], np.int32)


def _oracle_intersite(oracle_mod, lat, ham, rec, recur, lld, ene):
    orc = oracle_mod.Oracle(lat, ham)
    _, slots, (si, sj, asg, bsg) = rec._pair_units()
    n4 = 4 * len(PAIRS)
    if recur == "block":
        a_c, b_c = orc.lanczos_block(si, lld, site_j=sj, asign=asg, bsign=bsg)
        g_c = oracle_mod.block_green(a_c, orc.zsqr(b_c), ene)
    else:
        a, b = oracle_mod.cheb_scale(EMIN, EMAX)
        mu_c, _ = orc.cheb_moments(si, lld, a, b, site_j=sj, asign=asg, bsign=bsg)
        g_c = oracle_mod.chebyshev_green(mu_c, ene, EMIN, EMAX)[1]
    g0 = np.zeros((18, 18, len(ene), n4), complex, order="F")
    g0[..., slots] = g_c
    return oracle_mod.intersite_gf(g0, PAIRS)


@pytest.mark.parametrize("recur", ["block", "chebyshev"])
@pytest.mark.parametrize("fused", [True, False])
def test_calculate_intersite_gf(oracle_mod, recur, fused):
    from rslmtoasa_b200 import Green
    lat, ham = case("bulk_hoh" if recur == "block" else "bulk")
    lld = 8
    rec = _rec(lat, ham, lld=lld, ijpair=PAIRS, channels=60)
    rec.control.recur = recur
    gr = Green(rec)
    if not fused:                                  # the reference's staged flow: recursion results on the host, four slots per pair
        if recur == "block":
            rec.recur_b_ij(); rec.zsqr()
        else:
            rec.chebyshev_recur_ij()
    gij, gji = gr.calculate_intersite_gf(fused=fused)
    oij, oji, ogs = _oracle_intersite(oracle_mod, lat, ham, rec, recur, lld, gr.ene)
    ok = np.isfinite(oij)                          # chebyshev: the mesh tail beyond |w| = 1 is NaN on both sides
    assert ok.mean() > 0.7 and np.array_equal(np.isfinite(gij), ok)
    tol = TOL_G if recur == "block" else TOL_SUM
    assert relerr(gij[ok], oij[ok]) < tol and relerr(gji[ok], oji[ok]) < tol
    mine = np.stack([gr.ginmag, gr.gix, gr.giy, gr.giz, gr.gjnmag, gr.gjx, gr.gjy, gr.gjz], axis=-1)
    oks = np.isfinite(ogs)
    assert relerr(mine[oks], ogs[oks]) < tol
    # i == j pair: gij = gji = the on-site Green function of that site
    assert np.array_equal(gij[..., 1][ok[..., 1]], gji[..., 1][ok[..., 1]])


def test_intersite_gf_needs_matching_resident_g0(block_rec):
    from rslmtoasa_b200 import Green, RsrecError
    import ctypes as C
    gr = Green(block_rec)
    gr.block_green()                               # 4 on-site units, not pair units
    z = np.zeros((18, 18, len(gr.ene), 3), complex, order="F")
    pi = np.array([1, 2, 3], np.int32); pj = np.array([2, 3, 4], np.int32)
    rc = block_rec._L.rsrec_intersite_gf(block_rec._h, 3, pi.ctypes.data_as(C.c_void_p), pj.ctypes.data_as(C.c_void_p), 1,
                                         z.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), None)
    assert rc == -1


def test_conductivity_cumulative_bit_exact(oracle_mod, block_rec):
    """tail of calculate_conductivity_tensor: the running-sum kernel == the literal O(nv^2) simpson_f loop, bit for bit"""
    from rslmtoasa_b200 import Conductivity
    cond = Conductivity(block_rec)
    en = block_rec.en
    nv = len(cond.ene)
    rng = np.random.default_rng(17)
    cond.integrand = np.asfortranarray(rng.standard_normal((18, nv)) + 1j * rng.standard_normal((18, nv)))
    cond.integrand_at = np.asfortranarray(rng.standard_normal((18, nv, 3)) + 1j * rng.standard_normal((18, nv, 3)))
    a, b = en.scale_shift()
    ws = (cond.ene - b) / a
    for calctype in ("per_type", "random_vec"):
        block_rec.control.cond_calctype = calctype
        sig = cond.integrate_conductivity()
        ref = oracle_mod.conductivity_cumulative(cond.integrand, cond.integrand_at if calctype == "per_type" else None, en.nv1, ws, 3)
        assert sig.shape == ref.shape and np.array_equal(sig, ref)
    block_rec.control.cond_calctype = "per_type"
    # a non-finite sample poisons every integral of its series in the literal loop (NaN * 0): same here
    cond.integrand[4, 10] = np.nan
    sig = cond.integrate_conductivity()
    assert np.isnan(sig[0, 0, :, 0]).all() and np.isnan(sig[0, 5, :, 0]).all()          # real total, real orbital l2 = 5
    assert np.isfinite(sig[1, :, :, 0]).all() and np.isfinite(sig[:, 3, :, 0]).all()      # imaginary parts, other orbitals
