"""world_size-2 (and 3) gloo tests of the N>1 host logic: unit sharding with the reference's block rule and the
one collective per recursion call.  The per-rank compute is played by the CPU oracle here (no GPU in this
container); on the GPU box bench.py runs the same parallel.py functions over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from rslmtoasa_b200 import synthetic as S, parallel as P
    from tests.cases import case, EMIN, EMAX
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lat, ham = case("pbc")
    orc = O.Oracle(lat, ham, threads=1)
    a, b = O.cheb_scale(EMIN, EMAX)
    # (1) random KPM vectors sharded over ranks, moments all-reduced (stochastic trace)
    nvec = 5
    ph = S.random_phases(lat.kk, nvec)
    lo, hi = P.shard_range(nvec, rank, world)
    mu_loc, _ = orc.cheb_moments_random(ph[:, lo:hi], 5, a, b) if hi > lo else (np.zeros((18, 18, 12, 0), complex), 0)
    mu_sum = P.allreduce_sum(np.asfortranarray(mu_loc.sum(axis=-1)))
    # (2) recursion sites sharded over ranks, site-resolved coefficients gathered
    sites = np.array([1, 4, 9, 12, 30], dtype=np.int32)
    lo, hi = P.shard_range(len(sites), rank, world)
    a_loc, b_loc = orc.lanczos_block(sites[lo:hi], 5)
    a_all = P.allgather_units(a_loc, len(sites))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), mu_sum=mu_sum, a_all=a_all)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_units_and_collectives(tmp_path, oracle_mod, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from rslmtoasa_b200 import synthetic as S
    from tests.cases import case, EMIN, EMAX, relerr
    lat, ham = case("pbc")
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = orc.cheb_moments_random(S.random_phases(lat.kk, 5), 5, a, b)
    a_b, _ = orc.lanczos_block([1, 4, 9, 12, 30], 5)
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert relerr(g["mu_sum"], mu.sum(axis=-1)) < 1e-13        # every rank holds the full sum
        assert relerr(g["a_all"], a_b) < 1e-13                      # gather reproduces the 1-rank result


def _bands_worker(rank, world, port, out_dir):
    """the N>1 flow of `Bands` (rslmtoasa_b200/bands.py): per-rank dtot -> all-reduce -> Fermi scan on every rank ->
    per-unit moments -> all-gather; the per-rank compute is played by the oracle"""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from rslmtoasa_b200 import parallel as P
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g0, m, qqv = _bands_case(O)
    nu = g0.shape[3]
    lo, hi = P.shard_range(nu, rank, world)
    loc = np.asfortranarray(g0[..., lo:hi])
    dtot = O.bands_dos(loc)[0] if hi > lo else np.zeros(g0.shape[2])
    dtot = P.allreduce_sum(dtot)
    ef, nv1, e1, ifail = O.bands_fermi(dtot, m["edel"], -1.2, qqv, 0.05, m["nv1"])
    mom = np.asfortranarray(np.tile(np.array([0.0, 0.0, 1.0])[:, None], (1, hi - lo)))
    occ = O.bands_moments(loc, m["channels_ldos"], mom, m["ene"], m["edel"], ef, nv1, e1)[0] if hi > lo else np.zeros((3, 6, 0), order="F")
    occ_all = P.allgather_units(occ, nu)
    np.savez(os.path.join(out_dir, f"b{rank}.npz"), dtot=dtot, ef=ef, nv1=nv1, occ=occ_all)
    dist.destroy_process_group()


def _bands_case(O):
    m = O.e_mesh_full(-1.2, 0.9, 120, 0.05)
    nv, nu = len(m["ene"]), 5
    rng = np.random.default_rng(3)
    g0 = rng.standard_normal((18, 18, nv, nu)) + 1j * rng.standard_normal((18, 18, nv, nu))
    g0[np.arange(18), np.arange(18)] -= 3j
    return np.asfortranarray(g0), m, 20.0


@pytest.mark.parametrize("world", [2, 3])
def test_bands_flow_over_ranks(tmp_path, oracle_mod, world):
    mp.spawn(_bands_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, m, qqv = _bands_case(oracle_mod)
    dtot = oracle_mod.bands_dos(g0)[0]
    ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot, m["edel"], -1.2, qqv, 0.05, m["nv1"])
    assert ifail == 0
    mom = np.tile(np.array([0.0, 0.0, 1.0])[:, None], (1, g0.shape[3]))
    occ = oracle_mod.bands_moments(g0, m["channels_ldos"], mom, m["ene"], m["edel"], ef, nv1, e1)[0]
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), f"b{r}.npz"))
        assert np.allclose(g["dtot"], dtot, rtol=1e-13, atol=1e-13)      # sum order differs across ranks: rounding only
        assert int(g["nv1"]) == nv1 and abs(float(g["ef"]) - ef) < 1e-12
        assert np.allclose(g["occ"], occ, rtol=1e-10, atol=1e-12)
