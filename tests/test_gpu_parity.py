"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: a_n/b_n 1e-10 relative (over the stable depth), mu_n 1e-9 relative.
"""
import numpy as np
import pytest

from tests.cases import case, relerr, EMIN, EMAX

pytestmark = pytest.mark.gpu

TOL_AB = 1e-10
TOL_MU = 1e-9


def _rec(lat, ham, **kw):
    from rslmtoasa_b200 import Recursion, Control, Energy
    ctl = Control(**{k: v for k, v in kw.items() if k in ("lld", "cond_ll", "cond_calctype", "random_vec_num")})
    extra = {k: v for k, v in kw.items() if k in ("ijpair", "atlist", "phases", "rank", "numprocs")}
    return Recursion(ham, lat, ctl, Energy(EMIN, EMAX), **extra)


@pytest.mark.parametrize("family", [0, 1])
@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "surface", "impurity", "impurity_hoh", "pbc", "tiny"])
def test_recur_b(oracle_mod, name, family):
    lat, ham = case(name)
    lld = 8
    rec = _rec(lat, ham, lld=lld)
    rec.set_kernel_family(family)
    rec.recur_b()
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block(lat.irec, lld)
    assert rec.a_b.shape == a_b.shape
    assert relerr(rec.a_b, a_b) < TOL_AB
    assert relerr(rec.b2_b, b2_b) < TOL_AB
    d = np.arange(18)
    assert np.array_equal(rec.a[:, :, :, 0], np.real(rec.a_b[d, d]).transpose(1, 0, 2))


@pytest.mark.parametrize("family", [0, 1])
@pytest.mark.parametrize("name", ["bulk", "impurity_hoh"])
def test_recur_b_ij(oracle_mod, name, family):
    lat, ham = case(name)
    lld = 6
    pairs = np.array([[1, 2], [3, 3], [2, 7]], dtype=np.int32)
    rec = _rec(lat, ham, lld=lld, ijpair=pairs)
    rec.set_kernel_family(family)
    rec.recur_b_ij()
    orc = oracle_mod.Oracle(lat, ham)
    s = 1 / np.sqrt(2)
    signs = [(s, s), (s, -s), (s, 1j * s), (s, -1j * s)]
    for ij, (i, j) in enumerate(pairs):
        for reci in range(4):
            slot = ij * 4 + reci
            if i == j and reci > 0:
                assert not rec.a_b[..., slot].any()
                continue
            a_, b_ = (1.0, 1.0) if i == j else signs[reci]
            a_b, b2_b = orc.lanczos_block([i], lld, site_j=[j], asign=[a_], bsign=[b_])
            assert relerr(rec.a_b[..., slot], a_b[..., 0]) < TOL_AB
            assert relerr(rec.b2_b[..., slot], b2_b[..., 0]) < TOL_AB


@pytest.mark.parametrize("name", ["bulk", "impurity"])
def test_recur_scalar(oracle_mod, name):
    lat, ham = case(name)
    lld = 9
    rec = _rec(lat, ham, lld=lld)
    rec.recur()
    a, b2 = oracle_mod.Oracle(lat, ham).lanczos_scalar(lat.irec, lld)
    assert relerr(rec.a[..., 0], a) < TOL_AB
    assert relerr(rec.b2[..., 0], b2) < TOL_AB


def test_zsqr(oracle_mod):
    lat, ham = case("bulk")
    rec = _rec(lat, ham, lld=6)
    rec.recur_b()
    ref = oracle_mod.Oracle(lat, ham).zsqr(rec.b2_b)
    b2 = rec.b2_b.copy()
    rec.zsqr()
    assert relerr(rec.b2_b, ref) < 1e-10
    for ll in range(6):
        assert relerr(rec.b2_b[:, :, ll, 0] @ rec.b2_b[:, :, ll, 0], b2[:, :, ll, 0]) < 1e-11


@pytest.mark.parametrize("family", [0, 1])
@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "surface", "impurity", "impurity_hoh", "pbc", "tiny"])
def test_chebyshev_recur(oracle_mod, name, family):
    lat, ham = case(name)
    lld = 12
    rec = _rec(lat, ham, lld=lld)
    rec.set_kernel_family(family)
    rec.chebyshev_recur()
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, rc = oracle_mod.Oracle(lat, ham).cheb_moments(lat.irec, lld, a, b)
    assert rc == 0
    assert rec.mu_n.shape == mu.shape
    assert relerr(rec.mu_n, mu) < TOL_MU


@pytest.mark.parametrize("family", [0, 1])
def test_chebyshev_recur_ij(oracle_mod, family):
    lat, ham = case("bulk")
    lld = 7
    pairs = np.array([[1, 4], [5, 5]], dtype=np.int32)
    rec = _rec(lat, ham, lld=lld, ijpair=pairs)
    rec.set_kernel_family(family)
    rec.chebyshev_recur_ij()
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    orc = oracle_mod.Oracle(lat, ham)
    s = 1 / np.sqrt(2)
    mu, _ = orc.cheb_moments([1, 1, 1, 1, 5], lld, a, b, site_j=[4, 4, 4, 4, 5],
                             asign=[s, s, s, s, 1.0], bsign=[s, -s, 1j * s, -1j * s, 1.0])
    assert relerr(rec.mu_n[..., [0, 1, 2, 3, 4]], mu) < TOL_MU
    assert not rec.mu_n[..., 5:].any()


@pytest.mark.parametrize("family", [0, 1])
@pytest.mark.parametrize("name", ["pbc", "bulk_hoh"])
def test_chebyshev_random(oracle_mod, name, family):
    from rslmtoasa_b200 import synthetic as S
    lat, ham = case(name)
    lld = 10
    ph = S.random_phases(lat.kk, 3)
    rec = _rec(lat, ham, lld=lld, phases=ph)
    rec.set_kernel_family(family)
    rec.chebyshev_recur_random()
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, rc = oracle_mod.Oracle(lat, ham).cheb_moments_random(ph, lld, a, b)
    assert relerr(rec.mu_n, mu) < TOL_MU


def test_chebyshev_diverged(oracle_mod):
    """energy window far too small -> the reference calls fatal; we return RSREC_EDIVERGED."""
    from rslmtoasa_b200 import Recursion, Control, Energy, RsrecError
    lat, ham = case("bulk")
    rec = Recursion(ham, lat, Control(lld=40), Energy(-0.05, 0.05))
    with pytest.raises(RsrecError) as ei:
        rec.chebyshev_recur()
    assert ei.value.code == -2
    a, b = oracle_mod.cheb_scale(-0.05, 0.05)
    _, rc = oracle_mod.Oracle(lat, ham).cheb_moments(lat.irec, 40, a, b)
    assert rc == -2


@pytest.mark.parametrize("name", ["pbc", "impurity", "pbc_hoh"])
def test_ham_and_velo_vec_matmul(oracle_mod, name):
    lat, ham = case(name)
    if getattr(ham, "v_a", None) is None:
        from rslmtoasa_b200 import synthetic as S
        h2 = S.make_hamiltonian(lat, seed=20260103, velocity=True)
        ham.v_a, ham.v_b = h2.v_a, h2.v_b
    rng = np.random.default_rng(5)
    psi = np.asfortranarray(rng.normal(size=(18, 18, lat.kk)) + 1j * rng.normal(size=(18, 18, lat.kk)))
    rec = _rec(lat, ham)
    orc = oracle_mod.Oracle(lat, ham)
    ones = np.ones(lat.kk + 1, np.int32); ones[0] = 0
    a, b = 1.7, -0.2
    ref, _ = orc.ham_vec_matmul(psi, a, b, ones)
    assert relerr(rec.ham_vec_matmul(psi, a, b), ref) < 1e-12
    for slot in "ab":
        ref, _ = orc.velo_vec_matmul(slot, psi, ones)
        assert relerr(rec.velo_vec_matmul(slot, psi), ref) < 1e-12


@pytest.mark.parametrize("kind,M", [("per_type", 6), ("random_vec", 6), ("per_type", 1), ("random_vec", 9), ("random_vec", 4)])
@pytest.mark.parametrize("name", ["pbc", "pbc_hoh"])
def test_compute_moments_stochastic(oracle_mod, name, kind, M):
    """M = 1, 4, 6, 9 cover full, partial and single left blocks / right batches of the GEMM contraction."""
    from rslmtoasa_b200 import synthetic as S
    lat, ham = case(name)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    orc = oracle_mod.Oracle(lat, ham)
    if kind == "per_type":
        rec = _rec(lat, ham, cond_ll=M, cond_calctype=kind, atlist=[1])
        ref = orc.kubo_moments(M, a, b, start_sites=[1])
    else:
        ph = S.random_phases(lat.kk, 2)
        rec = _rec(lat, ham, cond_ll=M, cond_calctype=kind, phases=ph)
        ref = orc.kubo_moments(M, a, b, phases=ph)
    rec.compute_moments_stochastic()
    assert rec.mu_nm_stochastic.shape == ref.shape
    assert relerr(rec.mu_nm_stochastic, ref) < TOL_MU


def test_rank_partition_matches_reference_rule():
    """units are sharded with get_mpi_variables' rule; concatenating the shards reproduces the 1-rank result."""
    lat, ham = case("surface")
    full = _rec(lat, ham, lld=5)
    full.chebyshev_recur()
    parts = []
    for r in range(3):
        rec = _rec(lat, ham, lld=5, rank=r, numprocs=3)
        rec.chebyshev_recur()
        parts.append(rec.mu_n)
    assert np.array_equal(np.concatenate(parts, axis=-1), full.mu_n)


def test_device_resident_stepping_matches_one_shot():
    from rslmtoasa_b200 import synthetic as S
    lat, ham = case("pbc")
    ph = S.random_phases(lat.kk, 2)
    rec = _rec(lat, ham, lld=9, phases=ph)
    rec.chebyshev_recur_random()
    rec.cheb_begin_random(ph, 9)
    rec.cheb_run_steps(4)
    rec.cheb_run_steps(5)
    mu = rec.cheb_end()
    assert np.array_equal(mu, rec.mu_n)
    assert rec.launch_count > 0


# ---- BASELINE.json full sizes: size-independent properties (the oracle cannot run these in seconds) ----------
def test_full_size_1M_sites_properties():
    """config 5 (1M-site bcc, KPM random vector): mu_0 = kk/sqrtf(kk)^2 I exactly-ish, moments of a Hermitian H are
    Hermitian, the stochastic trace of T_1 is small, and stepping in two chunks equals one shot (determinism)."""
    from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
    lat = S.periodic_bcc(100, 100, 50)
    assert lat.kk == 1_000_000
    ham = S.make_hamiltonian(lat, seed=20260105)
    ph = S.random_phases(lat.kk, 1, seed=20260105)
    rec = Recursion(ham, lat, Control(lld=4), Energy(EMIN, EMAX), phases=ph)
    rec.chebyshev_recur_random()
    mu = rec.mu_n[..., 0]
    nrm = float(np.sqrt(np.float32(lat.kk)))
    assert relerr(mu[:, :, 0], np.eye(18) * lat.kk / nrm ** 2) < 1e-12
    for k in range(mu.shape[2]):
        assert np.abs(mu[:, :, k] - mu[:, :, k].conj().T).max() < 1e-10
    assert np.abs(mu).max() < 1.5                     # |T_n| <= 1 inside the window
    rec.cheb_begin_random(ph, 4)
    rec.cheb_run_steps(1)
    rec.cheb_run_steps(3)
    assert np.array_equal(rec.cheb_end()[..., 0], mu)  # bit-reproducible run to run


def test_kernel_families_agree_at_config1_size():
    """config 1 (bulk bcc Fe, kk = 5984): tensor-pipe and SIMT families against each other at full size, plus
    the invariants of the block recursion (B^2 Hermitian positive definite, A Hermitian, mu_0 = I)."""
    from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
    lat = S.sphere_cluster("bcc", 80.0)
    ham = S.make_hamiltonian(lat, seed=20260101)
    out = []
    for fam in (0, 1):
        rec = Recursion(ham, lat, Control(lld=21), Energy(EMIN, EMAX))
        rec.set_kernel_family(fam)
        rec.recur_b()
        rec.chebyshev_recur()
        out.append((rec.a_b.copy(), rec.b2_b.copy(), rec.mu_n.copy()))
    assert relerr(out[1][0], out[0][0]) < 1e-11 and relerr(out[1][1], out[0][1]) < 1e-11
    assert relerr(out[1][2], out[0][2]) < 1e-12
    a_b, b2_b, mu = out[1]
    assert relerr(mu[:, :, 0, 0], np.eye(18)) < 1e-14
    for ll in range(21):
        assert relerr(b2_b[:, :, ll, 0], b2_b[:, :, ll, 0].conj().T) < 1e-12
        assert np.linalg.eigvalsh(b2_b[:, :, ll, 0]).min() > 0
        assert np.abs(a_b[:, :, ll, 0] - a_b[:, :, ll, 0].conj().T).max() < 1e-11


def test_linearity_of_ham_vec_matmul_at_config4_size():
    """config 4 lattice (kk = 8000 PBC): H~(x + c y) = H~ x + c H~ y, and <x|H y> = <H x|y> for the Hermitian H"""
    from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
    lat = S.periodic_bcc(10, 20, 20)
    ham = S.make_hamiltonian(lat, seed=20260104)
    rec = Recursion(ham, lat, Control(), Energy(EMIN, EMAX))
    rng = np.random.default_rng(11)
    x = np.asfortranarray(rng.normal(size=(18, 18, lat.kk)) + 1j * rng.normal(size=(18, 18, lat.kk)))
    y = np.asfortranarray(rng.normal(size=(18, 18, lat.kk)) + 1j * rng.normal(size=(18, 18, lat.kk)))
    c = 0.3 - 0.7j
    hx, hy, hxy = rec.ham_vec_matmul(x, 1.0, 0.0), rec.ham_vec_matmul(y, 1.0, 0.0), rec.ham_vec_matmul(x + c * y, 1.0, 0.0)
    assert relerr(hxy, hx + c * hy) < 1e-13
    assert abs(np.vdot(x, hy) - np.vdot(hx, y)) / abs(np.vdot(x, hy)) < 1e-12


@pytest.mark.parametrize("family", [0, 1])
def test_results_are_bitwise_reproducible(family):
    """every reduction runs in a fixed order (no atomics on the data path): same handle or a fresh one, same bits"""
    lat, ham = case("impurity_hoh")
    outs = []
    for fresh in range(2):
        rec = _rec(lat, ham, lld=7)
        rec.set_kernel_family(family)
        for rep in range(2):
            rec.recur_b()
            rec.chebyshev_recur()
            outs.append((rec.a_b.copy(), rec.b2_b.copy(), rec.mu_n.copy()))
    for o in outs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(o, outs[0]))


@pytest.mark.parametrize("name", ["bulk", "tiny", "pbc", "impurity"])
def test_create_ll_map_bit_exact(oracle_mod, name):
    lat, ham = case(name)
    rec = _rec(lat, ham, lld=6)
    orc = oracle_mod.Oracle(lat, ham)
    for site in (1, lat.kk):
        assert np.array_equal(rec.create_ll_map(site), orc.create_ll_map(site, 6))


@pytest.mark.parametrize("family", [0, 1])
@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "impurity"])
def test_chebyshev_orbital_mod(oracle_mod, name, family):
    lat, ham = case(name)
    rec = _rec(lat, ham, lld=7)
    rec.set_kernel_family(family)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    cr = 0.5 * lat.cr.astype(np.float64)
    starts = [1, 4, lat.kk]
    mu = rec.chebyshev_orbital_mod(starts, cr, 5.42)
    ref = oracle_mod.Oracle(lat, ham).orbital_moments(starts, cr, 5.42, 7, a, b)
    assert relerr(mu, ref) < TOL_MU
    assert np.array_equal(mu, rec.chebyshev_orbital_mod(starts, cr, 5.42))       # reproducible
    # the tail of the routine (Jackson weights, Chebyshev sum, trace, Simpson integral -> the rows of fort.50) on both moment sets
    from oracle import dense_check_post as D
    rec.en.channels_ldos, rec.en.fermi = 400, 0.0
    rec.en.e_mesh()
    rows = rec.chebyshev_orbital_tail(nstart=len(starts))
    want = D.orbital_tail(ref, len(starts), rec.en.ene, rec.en.fermi, rec.en.energy_min, rec.en.energy_max, rec.en.nv1)
    assert np.abs(rows - want).max() <= 1e-9 * np.abs(want).max()


def test_fresh_handles_give_identical_kubo_moments():
    """every handle uploads and packs the block sets again: the set-up path must be stream-ordered with the kernels
    (a synchronous cudaMemcpy from pageable memory returns before the DMA lands), so 12 fresh handles agree bit for bit"""
    from rslmtoasa_b200 import synthetic as S
    lat, ham = case("pbc_hoh")
    ph = S.random_phases(lat.kk, 2)
    first = None
    for it in range(12):
        rec = _rec(lat, ham, cond_ll=6, cond_calctype="random_vec", phases=ph)
        if it % 3 == 1:
            rec.recur_b()
        rec.compute_moments_stochastic()
        if first is None:
            first = rec.mu_nm_stochastic.copy()
        assert np.array_equal(rec.mu_nm_stochastic, first)
        rec.close()


def test_edge_sizes_and_empty_inputs(oracle_mod):
    """lld = 1 (no step at all), lld = 2, zero units, a single-site 'cluster', units on the last site"""
    from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
    lat, ham = case("tiny")
    orc = oracle_mod.Oracle(lat, ham)
    for lld in (1, 2):
        rec = _rec(lat, ham, lld=lld)
        rec.recur_b()
        a_b, b2_b = orc.lanczos_block(lat.irec, lld)
        assert rec.a_b.shape == (18, 18, lld, len(lat.irec))
        assert relerr(rec.b2_b, b2_b) < TOL_AB and np.abs(rec.a_b - a_b).max() < 1e-12
    rec = _rec(lat, ham, lld=0)
    rec.chebyshev_recur()                                    # lld = 0: only mu(1), mu(2)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = orc.cheb_moments(lat.irec, 0, a, b)
    assert rec.mu_n.shape == (18, 18, 2, len(lat.irec)) and relerr(rec.mu_n, mu) < TOL_MU
    # no units on this rank (more ranks than units): empty results, no launch
    rec = _rec(lat, ham, lld=4, rank=3, numprocs=4)
    n0 = rec.launch_count
    rec.recur_b(); rec.chebyshev_recur()
    assert rec.a_b.shape[-1] == 0 and rec.mu_n.shape[-1] == 0 and rec.launch_count == n0
    # the last site of the cluster as the recursion site (edge of the site range, mostly missing neighbours)
    lat.irec = np.array([lat.kk], dtype=np.int32)
    rec = _rec(lat, ham, lld=5)
    rec.recur_b()
    a_b, b2_b = orc.lanczos_block(lat.irec, 5)
    assert relerr(rec.a_b, a_b) < TOL_AB and relerr(rec.b2_b, b2_b) < TOL_AB
    # one-site lattice: H reduces to the on-site block; A_1 = H_on + lsham, B_2^2 = 0
    one = S.sphere_cluster("bcc", 0.1)
    assert one.kk == 1
    h1 = S.make_hamiltonian(one, seed=3)
    rec = _rec(one, h1, lld=2)
    rec.recur_b()
    assert relerr(rec.a_b[:, :, 0, 0], h1.ee[:, :, 0, 0] + h1.lsham[:, :, 0]) < 1e-14
    assert np.abs(rec.b2_b[:, :, 1, 0]).max() < 1e-25


def test_positions_only_reorder_the_work():
    """rsrec_set_positions (lattice%cr) sorts the tiles along a Morton curve for L2 locality: the Chebyshev moments are
    bit-identical (site-ordered reductions); the Lanczos A = sum psi^H H psi is reduced per CTA of the SpMV kernel in tile
    order, so its rounding follows the tile order (1e-13)"""
    from rslmtoasa_b200 import synthetic as S
    lat = S.periodic_bcc(6, 5, 4)
    ham = S.make_hamiltonian(lat, seed=20260104)
    ph = S.random_phases(lat.kk, 2)
    outs = []
    for with_pos in (False, True):
        lat.cr = S.periodic_bcc_positions(6, 5, 4) if with_pos else None
        rec = _rec(lat, ham, lld=6, phases=ph)
        rec.chebyshev_recur_random()
        mu = rec.mu_n.copy()
        rec.recur_b()
        outs.append((mu, rec.a_b.copy(), rec.b2_b.copy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert relerr(outs[1][1], outs[0][1]) < 1e-13 and relerr(outs[1][2], outs[0][2]) < 1e-13


@pytest.mark.parametrize("seed", range(8))
def test_randomised_lattices_and_holes(oracle_mod, seed):
    """fuzz: random cluster radius / typing / local region / hoh, neighbour entries zeroed at random inside the slot
    lists (what chbar_nc does when hmfind fails, hamiltonian.f90:2350-2352), random recursion sites and pair units"""
    from rslmtoasa_b200 import synthetic as S
    rng = np.random.default_rng(1000 + seed)
    kind = ["bcc", "fcc"][seed % 2]
    ntype = int(rng.integers(1, 4))
    nmax = int(rng.integers(0, 6))
    lat = S.sphere_cluster(kind, float(rng.uniform(2.0, 7.0)), ntype=ntype, nmax=nmax,
                           type_rule="layer" if ntype > 1 else "single")
    ham = S.make_hamiltonian(lat, seed=seed, hoh=bool(seed & 2))
    nn = lat.nn.copy(order="F")
    holes = rng.random(nn.shape) < 0.08
    holes[:, 0] = False
    nn[holes] = 0
    lat.nn = nn
    nrec = int(rng.integers(1, 4))
    lat.irec = rng.choice(np.arange(1, lat.kk + 1), size=min(nrec, lat.kk), replace=False).astype(np.int32)
    lld = int(rng.integers(3, 8))
    family = seed % 2
    rec = _rec(lat, ham, lld=lld)
    rec.set_kernel_family(family)
    orc = oracle_mod.Oracle(lat, ham)
    rec.recur_b()
    a_b, b2_b = orc.lanczos_block(lat.irec, lld)
    # holes make H non-Hermitian: B^2 can leave the PSD cone late in the chain on both sides alike; compare the stable part
    ok = np.isfinite(a_b).all(axis=(0, 1, 3)) & np.isfinite(rec.a_b).all(axis=(0, 1, 3))
    n_ok = int(np.argmin(ok)) if not ok.all() else lld
    assert n_ok >= 2
    assert relerr(rec.a_b[:, :, :n_ok], a_b[:, :, :n_ok]) < 1e-8 and relerr(rec.b2_b[:, :, :n_ok], b2_b[:, :, :n_ok]) < 1e-8
    rec.chebyshev_recur()
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = orc.cheb_moments(lat.irec, lld, a, b)
    assert relerr(rec.mu_n, mu) < TOL_MU
    if lat.kk >= 3:
        i, j = (int(v) for v in rng.choice(np.arange(1, lat.kk + 1), size=2, replace=False))
        rec2 = _rec(lat, ham, lld=lld, ijpair=np.array([[i, j]], dtype=np.int32))
        rec2.set_kernel_family(family)
        rec2.chebyshev_recur_ij()
        s = 1 / np.sqrt(2)
        mu2, _ = orc.cheb_moments([i] * 4, lld, a, b, site_j=[j] * 4, asign=[s] * 4, bsign=[s, -s, 1j * s, -1j * s])
        assert relerr(rec2.mu_n, mu2) < TOL_MU


@pytest.mark.parametrize("seed", range(6))
def test_randomised_kubo_and_scalar(oracle_mod, seed):
    """fuzz: Kubo-Bastin moments (per_type and random vectors, odd cond_ll, local region present so that the velocity
    operators skip sites 1..nmax like the reference) and the scalar recursion on random open clusters with holes"""
    from rslmtoasa_b200 import synthetic as S
    rng = np.random.default_rng(2000 + seed)
    kind = ["bcc", "fcc"][seed % 2]
    ntype, nmax = int(rng.integers(1, 3)), int(rng.integers(0, 4))
    lat = S.sphere_cluster(kind, float(rng.uniform(2.0, 5.0)), ntype=ntype, nmax=nmax,
                           type_rule="layer" if ntype > 1 else "single")
    hoh = bool(seed & 1) and nmax == 0                       # velo_hoh uses the type-indexed sets only
    ham = S.make_hamiltonian(lat, seed=seed, hoh=hoh, velocity=True)
    if hoh:
        ham.vo_a = np.asfortranarray(0.3j * ham.eeo); ham.vo_b = np.asfortranarray(-0.2 * ham.eeo)
    nn = lat.nn.copy(order="F")
    holes = rng.random(nn.shape) < 0.05
    holes[:, 0] = False
    nn[holes] = 0
    lat.nn = nn
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    M = int(rng.integers(1, 8))
    orc = oracle_mod.Oracle(lat, ham)
    family = (seed // 2) % 2
    if seed % 3 == 0:
        sites = rng.choice(np.arange(nmax + 1, lat.kk + 1), size=2, replace=False).astype(np.int32)
        rec = _rec(lat, ham, cond_ll=M, cond_calctype="per_type", atlist=sites)
        ref = orc.kubo_moments(M, a, b, start_sites=sites)
    else:
        ph = np.asfortranarray(rng.random((lat.kk, 2)))
        rec = _rec(lat, ham, cond_ll=M, cond_calctype="random_vec", phases=ph)
        ref = orc.kubo_moments(M, a, b, phases=ph)
    rec.set_kernel_family(family)
    rec.compute_moments_stochastic()
    assert relerr(rec.mu_nm_stochastic, ref) < TOL_MU
    # scalar recursion on the spin-diagonal part
    ham1 = S.make_hamiltonian(lat, seed=seed, spin_orbit=False)
    lat.irec = rng.choice(np.arange(1, lat.kk + 1), size=2, replace=False).astype(np.int32)
    lld = int(rng.integers(2, 7))
    rec = _rec(lat, ham1, lld=lld)
    rec.set_kernel_family(family)
    rec.recur()
    sa, sb = oracle_mod.Oracle(lat, ham1).lanczos_scalar(lat.irec, lld)
    assert relerr(rec.a[..., 0], sa) < 1e-8 and relerr(rec.b2[..., 0], sb) < 1e-8


def test_unit_batching_when_memory_is_short(oracle_mod, monkeypatch):
    """units that do not fit in device memory together are processed in batches (RSREC_UNIT_BATCH forces it here):
    every driver gives the unbatched result (the reduction grids differ, so to rounding)"""
    lat, ham = case("surface")                       # 4 recursion sites
    pairs = np.array([[1, 2], [5, 5], [2, 9]], dtype=np.int32)
    ref = _rec(lat, ham, lld=6, ijpair=pairs)
    ref.recur_b(); a0, b0 = ref.a_b.copy(), ref.b2_b.copy()
    ref.chebyshev_recur(); m0 = ref.mu_n.copy()
    ref.recur_b_ij(); aij0 = ref.a_b.copy()
    for nb in ("1", "3"):
        monkeypatch.setenv("RSREC_UNIT_BATCH", nb)
        rec = _rec(lat, ham, lld=6, ijpair=pairs)
        rec.recur_b()
        assert relerr(rec.a_b, a0) < 1e-12 and relerr(rec.b2_b, b0) < 1e-12
        rec.chebyshev_recur()
        assert relerr(rec.mu_n, m0) < 1e-12
        rec.recur_b_ij()
        assert relerr(rec.a_b, aij0) < 1e-12
        from rslmtoasa_b200 import Green
        g = Green(rec)
        rec2 = _rec(lat, ham, lld=6)
        g2 = Green(rec2)
        monkeypatch.delenv("RSREC_UNIT_BATCH")
        want = g2.recur_b_green().copy()
        monkeypatch.setenv("RSREC_UNIT_BATCH", nb)
        rec3 = _rec(lat, ham, lld=6)
        got = Green(rec3).recur_b_green()
        ok = np.isfinite(want)
        assert relerr(got[ok], want[ok]) < 1e-9
    monkeypatch.delenv("RSREC_UNIT_BATCH")


# ---- spin-diagonal (collinear) hopping blocks: k_apply_dmma_sd ------------------------------------------------------------
def _collinear(ham):
    """zero the blocks between the two spins on every hopping slot (m >= 2); the on-site slot keeps its l.s-like coupling"""
    import copy
    h = copy.deepcopy(ham)
    for name in ("ee", "eeo", "hall", "hallo", "v_a", "v_b", "vo_a", "vo_b"):
        a = getattr(h, name, None)
        if a is None or a.size == 0:
            continue
        a = np.array(a, order="F")
        a[:9, 9:, 1:, ...] = 0.0
        a[9:, :9, 1:, ...] = 0.0
        setattr(h, name, a)
    return h


@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "impurity", "surface", "pbc_hoh"])
def test_spin_diagonal_hoppings_take_the_two_block_kernel(oracle_mod, name, monkeypatch):
    """collinear Hamiltonians (hopping blocks spin-diagonal, on-site block full) run the SpMV as two 18x18 products per
    slot; results must agree with the oracle and, to rounding, with the full-block kernel (RSREC_NO_SPIN_DIAG=1)"""
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case(name)
    ham = _collinear(ham)
    lld = 7
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    a_o, b_o = orc.lanczos_block(lat.irec, lld)
    mu_o, _ = orc.cheb_moments(lat.irec, lld, a, b)
    res = {}
    for mode in ("sd", "full"):
        if mode == "full":
            monkeypatch.setenv("RSREC_NO_SPIN_DIAG", "1")
        rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX))
        n0 = rec.launch_count
        rec.recur_b()
        rec.chebyshev_recur()
        res[mode] = (rec.a_b.copy(), rec.b2_b.copy(), rec.mu_n.copy())
        nsd = rec._L.rsrec_spin_diag_launch_count(rec._h)
        assert (nsd > 0) == (mode == "sd"), (mode, nsd)
        assert relerr(rec.a_b, a_o) < 1e-10 and relerr(rec.b2_b, b_o) < 1e-10 and relerr(rec.mu_n, mu_o) < 1e-9
        rec.close()
    for x, y in zip(res["sd"], res["full"]):
        assert relerr(x, y) < 1e-12


@pytest.mark.parametrize("name", ["pbc", "pbc_hoh"])
def test_spin_diagonal_velocity_sets_in_kubo_moments(oracle_mod, name):
    """Kubo-Bastin moments with collinear blocks: the velocity products v*psi (and vo*, hoh) take the two-block kernel too"""
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case(name)
    ham = _collinear(ham)
    M = 6
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu_o = oracle_mod.Oracle(lat, ham).kubo_moments(M, a, b, start_sites=[1, 5])
    rec = Recursion(ham, lat, Control(cond_ll=M, cond_calctype="per_type"), Energy(EMIN, EMAX), atlist=[1, 5])
    rec.compute_moments_stochastic()
    assert rec._L.rsrec_spin_diag_launch_count(rec._h) > 0
    assert relerr(rec.mu_nm_stochastic, mu_o) < 1e-9


@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "surface", "impurity_hoh", "pbc"])
def test_gram_products_inside_the_spmv_kernel(oracle_mod, name):
    """the fused forms (SpMV + three-term update + the 18x18 reductions in one kernel) against the oracle and against the
    separate-kernel forms: Lanczos fused is the default, Chebyshev fused is opt-in (rsrec_set_fusion)"""
    from rslmtoasa_b200 import synthetic as S
    lat, ham = case(name)
    lld = 7
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    a_b, b2_b = orc.lanczos_block(lat.irec, lld)
    mu, _ = orc.cheb_moments(lat.irec, lld, a, b)
    res = {}
    for fused in (0, 1):
        rec = _rec(lat, ham, lld=lld)
        rec.set_fusion(lanczos=fused, cheb=fused)
        l0 = rec.launch_count
        rec.recur_b()
        nl = rec.launch_count - l0
        ph = S.random_phases(lat.kk, 2)
        res[fused] = (rec.a_b.copy(), rec.b2_b.copy(), nl)
        assert relerr(rec.a_b, a_b) < TOL_AB and relerr(rec.b2_b, b2_b) < TOL_AB
        rec.chebyshev_recur()
        assert relerr(rec.mu_n, mu) < TOL_MU
        rec.chebyshev_recur_random(ph)
        ref, _ = orc.cheb_moments_random(ph, lld, a, b)
        assert relerr(rec.mu_n, ref) < TOL_MU
        rec.close()
    assert res[1][2] < res[0][2]                      # one launch less per step
    assert relerr(res[1][0], res[0][0]) < 1e-12 and relerr(res[1][1], res[0][1]) < 1e-12



# ---- pipelined block Lanczos step (k_rotortho_dmma + the square root on a second stream) ------------------------------------------
@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "impurity", "impurity_hoh", "surface", "pbc"])
def test_pipelined_lanczos_step_against_the_oracle_and_the_five_launch_step(oracle_mod, name, monkeypatch):
    """H applied to the unnormalised residual while B = (B^2)^1/2 is formed on a second stream, rotation + three-term update +
    orthogonalisation in one pass (A_n = B^-1 (R^H H R) B^-1): same algebra as crecal_b (recursion.f90:1873-1973), different
    association -- a_b / b2_b must agree with the oracle to 1e-10 and with the five-launch step to rounding, for site and pair
    starts, with hoh, a site-indexed region, several units in a batch, and for the scalar recursion."""
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case(name)
    lld = 9
    orc = oracle_mod.Oracle(lat, ham)
    a_o, b_o = orc.lanczos_block(lat.irec, lld)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RSREC_LZ_PIPELINE", mode)
        rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX))
        n0 = rec.launch_count
        rec.recur_b()
        launches = rec.launch_count - n0
        assert relerr(rec.a_b, a_o) < 1e-10 and relerr(rec.b2_b, b_o) < 1e-10, mode
        out = [rec.a_b.copy(), rec.b2_b.copy()]
        rec.recur_b()                        # two streams, fixed-order reductions: run-to-run bitwise reproducible
        assert np.array_equal(rec.a_b, out[0]) and np.array_equal(rec.b2_b, out[1]), mode
        rec.ijpair = np.array([[1, 2], [3, 3]], dtype=np.int32)
        rec.recur_b_ij()
        out += [rec.a_b.copy(), rec.b2_b.copy()]
        res[mode] = (out, launches)
        rec.close()
    assert res["1"][1] < res["0"][1]          # four launches per step (one of them off the critical path) instead of five
    for x, y in zip(res["1"][0], res["0"][0]):
        assert relerr(x, y) < 1e-11


def test_pipelined_scalar_lanczos(oracle_mod, monkeypatch):
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case("impurity")
    lld = 9
    a_o, b_o = oracle_mod.Oracle(lat, ham).lanczos_scalar(lat.irec, lld)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RSREC_LZ_PIPELINE", mode)
        rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX))
        rec.recur()
        assert relerr(rec.a[..., 0], a_o) < 1e-10 and relerr(rec.b2[..., 0], b_o) < 1e-10, mode
        res[mode] = (rec.a.copy(), rec.b2.copy())
        rec.close()
    for x, y in zip(res["1"], res["0"]):
        assert relerr(x, y) < 1e-11


@pytest.mark.parametrize("lld", [1, 2, 3])
def test_pipelined_lanczos_shortest_recursions(oracle_mod, lld, monkeypatch):
    """lld = 1 (no step at all), 2 (one square root, no merged pass) and 3 (one merged pass) through the pipelined driver"""
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case("bulk")
    a_o, b_o = oracle_mod.Oracle(lat, ham).lanczos_block(lat.irec, lld)
    monkeypatch.setenv("RSREC_LZ_PIPELINE", "1")
    rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX))
    rec.recur_b()
    assert relerr(rec.b2_b, b_o) < 1e-10
    if lld > 1:
        assert relerr(rec.a_b, a_o) < 1e-10
    else:
        assert not rec.a_b.any() and not np.abs(a_o).any()
    rec.close()
