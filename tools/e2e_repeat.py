import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import bench as B
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
lat = S.periodic_bcc(100, 100, 50)
ham = S.make_hamiltonian(lat, seed=20260105)
a_sc = None
rec = Recursion(ham, lat, Control(lld=249), Energy(B.EMIN, B.EMAX))
ph = S.random_phases(lat.kk, 1, seed=20260105)
for i in range(4):
    t0 = time.perf_counter(); rec.upload(); t1 = time.perf_counter()
    mu = rec.chebyshev_recur_random_sum(ph, sharded=False); t2 = time.perf_counter()
    print("e2e pass %d: upload %.3f s  recursion %.3f s  total %.3f s -> %.2f steps/s" % (i, t1 - t0, t2 - t1, t2 - t0, 249 / (t2 - t0)), flush=True)
