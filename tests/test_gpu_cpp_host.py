"""The C++ host-side mirror (include/rsrec.hpp) of the reference's `type recursion`, compiled with g++ against
librsrec.so and checked against the oracle through a flat binary hand-off."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.cases import case, EMIN, EMAX

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _w(f, arr, dtype):
    a = np.asarray(arr, dtype=dtype).ravel(order="F") if arr is not None else np.zeros(0, dtype)
    f.write(np.int64(a.size).tobytes())
    f.write(a.tobytes())


def _build(tmp_path):
    exe = os.path.join(str(tmp_path), "host_mirror_check")
    libdir = os.path.join(ROOT, "rslmtoasa_b200")
    subprocess.check_call([shutil.which("g++") or "/usr/bin/g++", "-O2", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp"),
                           "-L" + libdir, "-lrsrec", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_host_mirror_compiles(tmp_path):
    """CPU: the header-only mirror compiles and links against the C ABI."""
    from rslmtoasa_b200 import build as B
    B.build_library()
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_host_mirror_matches_oracle(tmp_path, oracle_mod):
    lat, ham = case("impurity")
    lld = 7
    pairs = np.array([[1, 2], [3, 3]], dtype=np.int32)
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    a_b, b2_b = orc.lanczos_block(lat.irec, lld)
    mu, _ = orc.cheb_moments(lat.irec, lld, a, b)
    s = 1 / np.sqrt(2)
    aij = np.zeros((18, 18, lld, 8), complex, order="F")
    r, _ = orc.lanczos_block([1, 1, 1, 1, 3], lld, site_j=[2, 2, 2, 2, 3], asign=[s, s, s, s, 1], bsign=[s, -s, 1j * s, -1j * s, 1])
    aij[..., [0, 1, 2, 3, 4]] = r
    path = os.path.join(str(tmp_path), "case.bin")
    with open(path, "wb") as f:
        _w(f, [lat.kk, lat.ncols, lat.ntype, lat.nmax, lld], np.int32)
        _w(f, lat.nn, np.int32); _w(f, lat.iz, np.int32); _w(f, lat.irec, np.int32); _w(f, pairs, np.int32)
        _w(f, ham.ee, np.complex128); _w(f, ham.lsham, np.complex128); _w(f, ham.hall, np.complex128)
        _w(f, [EMIN, EMAX], np.float64)
        _w(f, a_b, np.complex128); _w(f, b2_b, np.complex128); _w(f, mu, np.complex128); _w(f, aij, np.complex128)
        _w(f, [70, 0.05], np.float64)
        ene = oracle_mod.e_mesh(EMIN, EMAX, 70, 0.05)              # the C++ mirror builds the same mesh (energy.f90:175-208)
        _w(f, oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene), np.complex128)
        _w(f, oracle_mod.chebyshev_green(mu, ene, EMIN, EMAX)[1], np.complex128)   # NaN tail (|w| > 1) is skipped by the C++ max
        # bands on the block g0: valence, then the oracle's fermi, nv1, e1, eband, occ(3,6,n), mom0(3,n), lmom(3,n)
        g0 = oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene)
        m = oracle_mod.e_mesh_full(EMIN, EMAX, 70, 0.05)
        dtot = oracle_mod.bands_dos(g0)[0]
        qqv = 0.4 * oracle_mod.simpson_m(m["edel"], ene[m["nv1"] - 1], m["nv1"], dtot, ene[m["nv1"] - 1], 0, ene)
        ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot, m["edel"], EMIN, qqv, 0.05, m["nv1"])
        assert ifail == 0
        nu = g0.shape[3]
        mom = np.tile(np.array([0.0, 0.0, 1.0])[:, None], (1, nu))
        occ, lmom = oracle_mod.bands_moments(g0, m["channels_ldos"], mom, ene, m["edel"], ef, nv1, e1)
        m0, _ = oracle_mod.bands_magnetic_moments(g0, ene, m["edel"], ef, nv1, e1)
        _w(f, [qqv, ef, nv1, e1, oracle_mod.simpson_m(m["edel"], ef, nv1, dtot, e1, 1, ene)], np.float64)
        _w(f, occ, np.float64); _w(f, m0, np.float64); _w(f, lmom, np.float64)
    out = subprocess.run([_build(tmp_path), path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = {}
    for line in out.stdout.splitlines():
        t = line.split()
        if t[0] == "sharded":
            vals["sharded.a_b"], vals["sharded.phases"], vals["sharded.label"], vals["sharded.allreduce"] = float(t[2]), int(t[4]), t[6], float(t[8])
        elif t[0] in ("recur_b", "chebyshev_recur", "recur_b_ij", "block_green", "recur_b_green", "chebyshev_green", "bands"):
            for k, v in zip(t[1::2], t[2::2]):
                vals[t[0] + "." + k] = float(v)
    assert vals["recur_b.a_b"] < 1e-10 and vals["recur_b.b2_b"] < 1e-10
    assert vals["chebyshev_recur.mu_n"] < 1e-9
    assert vals["recur_b_ij.a_b"] < 1e-10
    assert vals["sharded.a_b"] < 1e-10 and vals["sharded.phases"] == 4 and vals["sharded.label"] == "H|PSI_n>" and vals["sharded.allreduce"] == 0.0
    assert vals["block_green.g0"] < 1e-8 and vals["recur_b_green.g0"] == 0.0
    assert vals["chebyshev_green.g0"] < 1e-9
    assert vals["bands.fermi"] < 1e-10 and vals["bands.nv1"] == 0.0 and vals["bands.eband"] < 1e-10
    assert vals["bands.occ"] < 1e-10 and vals["bands.mom0"] < 1e-10 and vals["bands.lmom"] < 1e-10
    assert "fatal -2" in out.stdout and "did not converge" in out.stdout
