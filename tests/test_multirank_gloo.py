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
