import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
lat = S.sphere_cluster("bcc", 80.0); ham = S.make_hamiltonian(lat, seed=20260101)
rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0))
rec.recur()
best = 1e9
for _ in range(7):
    t0 = time.perf_counter(); rec.recur(); best = min(best, time.perf_counter() - t0)
print(f"scalar recur lld=21 config 1: {best*1e3:.3f} ms, sd launches {rec._L.rsrec_spin_diag_launch_count(rec._h)}, checksum {np.abs(rec.a).sum():.12e}")
