"""The exchange step inside the library (NCCL on device-resident results) and the per-phase timers, through
the C ABI.  The single-rank communicator runs on any GPU box; the 2-rank test needs 2 GPUs (one process per GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

from tests.cases import case, relerr, EMIN, EMAX

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rec(lat, ham, device=0, **kw):
    from rslmtoasa_b200 import Recursion, Control, Energy
    ctl = Control(**{k: v for k, v in kw.items() if k in ("lld", "cond_ll", "cond_calctype")})
    extra = {k: v for k, v in kw.items() if k in ("atlist", "phases")}
    return Recursion(ham, lat, ctl, Energy(EMIN, EMAX, channels_ldos=300), device=device, **extra)


def test_phase_timers_carry_the_reference_labels():
    """g_timer labels of crecal_b / chebyshev_recur (recursion.f90:1902-1970, 3104-3127) with device times."""
    lat, ham = case("bulk")
    rec = _rec(lat, ham, lld=7)
    rec.phase_timing(True)
    rec.recur_b()
    ph = rec.phase_read()
    assert set(ph) == {"H|PSI_n>", "H|Psi_n-A_n|Psi_n-B_n|Psi_n-1", "B_n+1", "<PSI|B_n+1|PSI>"}
    assert all(ms > 0.0 for ms, _ in ph.values())
    assert all(ph[k][1] == 6 for k in ("H|PSI_n>", "H|Psi_n-A_n|Psi_n-B_n|Psi_n-1", "B_n+1"))
    # the pipelined step (small lattices, one unit) merges the rotation into the orthogonalisation pass and skips the last one,
    # whose result nothing reads; the five-launch step rotates after every level like the reference
    assert ph["<PSI|B_n+1|PSI>"][1] in (5, 6)
    rec.chebyshev_recur()
    ph = rec.phase_read()
    assert ph["<PSI_0|PSI_0>"][1] == 1 and ph["<PSI_0|PSI_1>"][1] == 1 and ph["<PSI_0|PSI_n>"][1] == 7
    rec.phase_timing(False)
    rec.recur_b()
    assert rec.phase_read() == {}
    rec.close()


def test_single_rank_communicator(oracle_mod):
    from rslmtoasa_b200 import Recursion, synthetic as S
    lat, ham = case("pbc")
    lat.irec = np.array([1, 4, 9, 12, 30], dtype=np.int32)
    rec = _rec(lat, ham, lld=5)
    rec.comm_init(1, 0, Recursion.comm_unique_id())
    n, r, ver = rec.comm_info()
    assert (n, r) == (1, 0) and ver >= 22000
    x = np.arange(7, dtype=np.float64)
    assert np.array_equal(rec.allreduce(x.copy()), x)
    rec.recur_b()
    a_loc = rec.a_b.copy()
    rec.recur_b_sharded()
    assert np.array_equal(rec.a_b, a_loc)
    rec.ijpair = np.array([[1, 2], [3, 3]], dtype=np.int32)
    rec.recur_b_ij()
    a_ij = rec.a_b.copy()
    rec.recur_b_ij_sharded()
    assert np.array_equal(rec.a_b, a_ij)
    assert np.array_equal(rec.allgather_units(a_loc, 5), a_loc)
    ph = S.random_phases(lat.kk, 3)
    rec.chebyshev_recur_random(ph)
    per_vec = rec.mu_n.copy()
    mu_sum = rec.chebyshev_recur_random_sum(ph)
    assert relerr(mu_sum, per_vec.sum(axis=-1)) < 1e-14
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    ref, _ = oracle_mod.Oracle(lat, ham).cheb_moments_random(ph, 5, a, b)
    assert relerr(mu_sum, ref.sum(axis=-1)) < 1e-9
    rec.comm_destroy()
    rec.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, id_bytes, out_dir):
    sys.path.insert(0, ROOT)
    from rslmtoasa_b200 import Conductivity, synthetic as S
    from rslmtoasa_b200.green import Green
    from rslmtoasa_b200.bands import Bands
    lat, ham = case("pbc")
    lat.irec = np.array([1, 4, 9, 12, 30], dtype=np.int32)
    ph = S.random_phases(lat.kk, 5)
    rec = _rec(lat, ham, device=rank, lld=5, cond_ll=6, cond_calctype="random_vec", phases=ph)
    rec.ijpair = np.array([[1, 2], [3, 3], [2, 7]], dtype=np.int32)
    rec.comm_init(world, rank, id_bytes)
    rec.recur_b_ij_sharded()                                # pair units sharded, gathered on the device
    a_ij = rec.a_b.copy()
    rec.recur_b_sharded()                                   # device all-gather of a_b / b2_b
    a_dev = rec.a_b.copy()
    mu_sum = rec.chebyshev_recur_random_sum(ph)             # device sum + all-reduce of the moments
    rec.recur_b()                                           # this rank's shard only
    gathered = rec.allgather_units(rec.a_b, len(lat.irec))  # host-array all-gather
    x = np.full(4, float(rank + 1))
    rec.allreduce(x)
    con = Conductivity(rec)
    integ, _ = con.compute_conductivity()                   # vectors sharded, integrand all-reduced in the library
    g = Green(rec)
    g.recur_b_green(download_g0=False)                      # this rank's units, g0 resident
    dtot = np.zeros(len(g.ene))
    from rslmtoasa_b200 import _lib
    _lib.check(rec._L.rsrec_bands_dos(rec._h, dtot.ctypes.data, None, None))   # dtot all-reduced on the device (bands.f90:276)
    np.savez(os.path.join(out_dir, f"c{rank}.npz"), a_dev=a_dev, a_ij=a_ij, a_sh=gathered, mu_sum=mu_sum, x=x,
             integ=integ, dtot=dtot)
    rec.comm_destroy()
    rec.close()


def test_two_rank_exchange_inside_the_library(tmp_path, oracle_mod):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one process per GPU)")
    import torch.multiprocessing as mp
    from rslmtoasa_b200 import Recursion, Conductivity, synthetic as S
    from rslmtoasa_b200.green import Green
    world = 2
    mp.spawn(_worker, args=(world, Recursion.comm_unique_id(), str(tmp_path)), nprocs=world, join=True)
    # single-rank results of the same job
    lat, ham = case("pbc")
    lat.irec = np.array([1, 4, 9, 12, 30], dtype=np.int32)
    ph = S.random_phases(lat.kk, 5)
    rec = _rec(lat, ham, lld=5, cond_ll=6, cond_calctype="random_vec", phases=ph)
    rec.ijpair = np.array([[1, 2], [3, 3], [2, 7]], dtype=np.int32)
    rec.recur_b_ij()
    a_ij = rec.a_b.copy()
    rec.recur_b()
    mu_sum = rec.chebyshev_recur_random_sum(ph)
    integ, _ = Conductivity(rec).compute_conductivity()
    g = Green(rec)
    g.recur_b_green(download_g0=False)
    from rslmtoasa_b200 import _lib
    dtot = np.zeros(len(g.ene))
    _lib.check(rec._L.rsrec_bands_dos(rec._h, dtot.ctypes.data, None, None))
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"c{r}.npz"))
        assert relerr(d["a_sh"], rec.a_b) < 1e-13
        assert relerr(d["a_dev"], rec.a_b) < 1e-13
        assert relerr(d["a_ij"], a_ij) < 1e-13
        assert relerr(d["mu_sum"], mu_sum) < 1e-13
        assert np.array_equal(d["x"], np.full(4, 3.0))
        ok = np.isfinite(integ)
        assert relerr(d["integ"][ok], integ[ok]) < 1e-11
        assert np.allclose(d["dtot"], dtot, rtol=1e-12, atol=1e-13)
    rec.close()
