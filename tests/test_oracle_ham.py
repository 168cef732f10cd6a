"""CPU tests of the Hamiltonian-assembly oracle (oracle/ham_oracle.py): pinned by the independent Pauli-algebra statement
and by invariants (unitarity of the orbital transformation, collinear structure, Hermiticity, rotation invariance)."""
import numpy as np
import pytest

from oracle import ham_oracle as HO


def random_inputs(seed=3, ntype=2, nslot=5, nloc=2, collinear=False):
    rng = np.random.default_rng(seed)
    ncls = ntype + nloc
    pot = {k: (rng.normal(size=(9, ntype)) * (0.3 if k.startswith("w") else 0.1)).astype(complex) for k in HO.POT_KEYS}
    pot["cx"] = rng.normal(size=(9, 2, ntype)).astype(complex) * 0.2
    pot["cex"] = rng.normal(size=(9, 2, ntype)).astype(complex) * 0.1
    mom = rng.normal(size=(3, ntype))
    mom /= np.linalg.norm(mom, axis=0)
    if collinear:
        mom[:] = np.array([0.0, 0.0, 1.0])[:, None]
    hhh = rng.normal(size=(9, 9, nslot, ncls))
    it = np.array([t + 1 for t in range(ntype)] + list(rng.integers(1, ntype + 1, nloc)), np.int32)
    jt = rng.integers(1, ntype + 1, size=(nslot, ncls)).astype(np.int32)
    jt[0] = it                       # slot 1 is the atom itself
    if nslot > 3:
        jt[3, 1] = 0                 # a missing neighbour
    return hhh, jt, it, pot, mom


def test_hcpx_is_a_unitary_similarity():
    assert np.allclose(HO.VC, HO.V.conj().T) and np.allclose(HO.V @ HO.VC, np.eye(9))
    rng = np.random.default_rng(0)
    h = rng.normal(size=(9, 9)); h = h + h.T
    assert np.allclose(np.linalg.eigvalsh(HO.hcpx_cart2sph(h.astype(complex))), np.linalg.eigvalsh(h))


@pytest.mark.parametrize("hoh", [False, True])
def test_blocks_vs_pauli_statement(hoh):
    hhh, jt, it, pot, mom = random_inputs()
    blk, blko, obarm, enim = HO.build_blocks(hhh, jt, it, pot, mom, hoh)
    for c in range(jt.shape[1]):
        for m in range(jt.shape[0]):
            if jt[m, c] == 0:
                assert not blk[:, :, m, c].any()
                continue
            ref = HO.pauli_block(int(it[c]), int(jt[m, c]), m == 0, hhh[:, :, m, c], pot, mom, hoh)
            assert np.abs(blk[:, :, m, c] - ref).max() < 1e-13
            if hoh:
                assert np.allclose(blko[:, :, m, c], ref @ obarm[:, :, jt[m, c] - 1])
    assert blko.any() == hoh


def test_collinear_structure_and_obarm_enim():
    hhh, jt, it, pot, mom = random_inputs(collinear=True)
    blk, _, obarm, enim = HO.build_blocks(hhh, jt, it, pot, mom, True)
    assert np.abs(blk[:9, 9:]).max() < 1e-15 and np.abs(blk[9:, :9]).max() < 1e-15      # no spin mixing
    for t in range(mom.shape[1]):
        d = np.arange(9)
        # diagonal orbital matrices stay block diagonal in l under the cart->sph transformation; spin up/down = o0 +- o1
        up = HO.hcpx_cart2sph(np.diag(pot["obx0"][:, t] + pot["obx1"][:, t]))
        assert np.allclose(obarm[:9, :9, t], up) and np.abs(obarm[:9, 9:, t]).max() < 1e-15
        eu = pot["cx"][:, 0, t] - pot["cex"][:, 0, t]
        assert np.allclose(enim[:9, :9, t], HO.hcpx_cart2sph(np.diag(eu)))


def test_hermiticity_of_reciprocal_pairs():
    """S(R) real with S(-R) = S(R)^T and real potential parameters: block_ji(-R) = block_ij(R)^H"""
    hhh, jt, it, pot, mom = random_inputs(ntype=2, nslot=2, nloc=0)
    h01 = HO.ham0m_nc(1, 2, False, hhh[:, :, 1, 0], pot, mom, False)
    h10 = HO.ham0m_nc(2, 1, False, hhh[:, :, 1, 0].T.copy(), pot, mom, False)
    b01 = HO.spin_block(np.stack([HO.hcpx_cart2sph(h01[:, :, k]) for k in range(4)], axis=2))
    b10 = HO.spin_block(np.stack([HO.hcpx_cart2sph(h10[:, :, k]) for k in range(4)], axis=2))
    assert np.abs(b10 - b01.conj().T).max() < 1e-13


def _spherical_L():
    """L_z, L_x, L_y in the basis (00)(1-1)(10)(11)(2-2)..(22) from the ladder-operator definitions (independent of the
    reference's tables)"""
    lz = np.zeros((9, 9)); lp = np.zeros((9, 9))
    for l in range(3):
        s = l * l + l
        for m in range(-l, l + 1):
            lz[s + m, s + m] = m
            if m < l:
                lp[s + m + 1, s + m] = np.sqrt(l * (l + 1) - m * (m + 1))
    lm = lp.T
    return 0.5 * (lp + lm), -0.5j * (lp - lm), lz


def test_local_axis_rotation_is_the_rotation_that_takes_m_to_z():
    rng = np.random.default_rng(5)
    assert np.allclose(HO.local_axis_rmat(np.array([0.0, 0.0, 1.0])), np.eye(18))
    lx, ly, lz = _spherical_L()
    for _ in range(4):
        m = rng.normal(size=3); m /= np.linalg.norm(m)
        r = HO.local_axis_rmat(m)
        assert np.allclose(r.conj().T @ r, np.eye(18), atol=1e-13)                  # unitary
        # spin: R^H (m.sigma) R = sigma_z ; orbital: R^H (m.L) R = L_z  (the moment direction becomes the z axis)
        ms = np.kron(sum(m[k] * HO.SIG[k] for k in range(3)), np.eye(9))
        assert np.allclose(r.conj().T @ ms @ r, np.kron(HO.SIG[2], np.eye(9)), atol=1e-12)
        ml = np.kron(np.eye(2), m[0] * lx + m[1] * ly + m[2] * lz)
        assert np.allclose(r.conj().T @ ml @ r, np.kron(np.eye(2), lz), atol=1e-12)


def test_rotated_blocks_of_a_ferromagnet_are_spin_diagonal():
    """all moments along m: after rotate_to_local_axis(m) the exchange part is sigma_z, so no spin mixing remains in
    blocks built from rotation-invariant orbital parts (s-like structure constants)"""
    hhh, jt, it, pot, mom = random_inputs(ntype=1, nslot=3, nloc=0)
    hhh[:] = 0.0
    hhh[0, 0] = 0.3                                    # s-s hopping only: orbital part invariant under rotation
    for k in ("cx0", "cx1", "cex0", "cex1"):
        pot[k][1:] = 0.0
    blk, _, _, _ = HO.build_blocks(hhh, jt, it, pot, mom, False)
    rot = HO.rotmag_loc(blk, mom[:, 0])
    assert np.abs(blk[0, 9]).max() > 1e-3                                           # spin mixing in the global frame
    assert np.abs(rot[:9, 9:]).max() < 1e-13 and np.abs(rot[9:, :9]).max() < 1e-13  # gone in the local frame
