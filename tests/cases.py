"""Shared seeded test cases (small enough for the oracle to finish in seconds)."""
import numpy as np

from rslmtoasa_b200 import synthetic as S

EMIN, EMAX = -2.0, 2.0   # spectrum of the synthetic H is within about +-1.6 Ry


def case(name):
    """-> (lattice, hamiltonian) for a named configuration (scaled-down BASELINE.json configs)."""
    if name == "bulk":            # config 1 shape: bcc sphere, 1 type, open boundary
        lat = S.sphere_cluster("bcc", 6.0)
        ham = S.make_hamiltonian(lat, seed=20260101)
    elif name == "bulk_hoh":
        lat = S.sphere_cluster("bcc", 6.0)
        ham = S.make_hamiltonian(lat, seed=20260101, hoh=True)
    elif name == "surface":       # config 2 shape: fcc, layer-typed, several units
        lat = S.sphere_cluster("fcc", 5.0, ntype=4, type_rule="layer")
        lat.irec = np.array([1, 2, 5, 9], dtype=np.int32)
        ham = S.make_hamiltonian(lat, seed=20260102)
    elif name == "impurity":      # config 3 shape: B2, 3 types, site-indexed local region
        lat = S.sphere_cluster("bcc", 6.0, ntype=3, nmax=9, type_rule="b2")
        ham = S.make_hamiltonian(lat, seed=20260103)
    elif name == "impurity_hoh":
        lat = S.sphere_cluster("bcc", 6.0, ntype=3, nmax=9, type_rule="b2")
        ham = S.make_hamiltonian(lat, seed=20260103, hoh=True)
    elif name == "pbc":           # config 4/5 shape: periodic bcc, no missing neighbours
        lat = S.periodic_bcc(4, 4, 3)
        ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    elif name == "pbc_hoh":
        lat = S.periodic_bcc(4, 3, 3)
        ham = S.make_hamiltonian(lat, seed=20260104, velocity=True, hoh=True)
        ham.vo_a = _vo(ham.eeo, 0.3j)
        ham.vo_b = _vo(ham.eeo, -0.2)
    elif name == "tiny":          # ragged: 2-site-deep cluster where most slots are missing
        lat = S.sphere_cluster("bcc", 0.8)
        ham = S.make_hamiltonian(lat, seed=7)
    else:
        raise KeyError(name)
    return lat, ham


def _vo(eeo, scale):
    """a deterministic 'v*o' operator set with the right shape (the reference builds vo from v and obarm)."""
    return np.asfortranarray(scale * eeo)


def relerr(x, ref):
    return float(np.abs(np.asarray(x) - np.asarray(ref)).max() / max(np.abs(np.asarray(ref)).max(), 1e-300))
