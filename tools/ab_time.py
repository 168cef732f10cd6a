#!/usr/bin/env python
"""best-of-N wall time of one driver on a named config: python tools/ab_time.py <bulk|surface|impurity> <what> [lld] [hoh] [reps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S

cfg, what = sys.argv[1], sys.argv[2]
lld = int(sys.argv[3]) if len(sys.argv) > 3 else 21
hoh = len(sys.argv) > 4 and sys.argv[4] == "hoh"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
if cfg == "bulk":
    lat = S.sphere_cluster("bcc", 80.0); ham = S.make_hamiltonian(lat, seed=20260101, hoh=hoh)
elif cfg == "surface":
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer"); lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260102, hoh=hoh)
else:
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2"); ham = S.make_hamiltonian(lat, seed=20260103, hoh=hoh)
rec = Recursion(ham, lat, Control(lld=lld), Energy(-2.0, 2.0))
fn = getattr(rec, what)
fn()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
print(cfg, what, "lld", lld, "hoh", hoh, "block_plan", "off" if os.environ.get("RSREC_NO_BLOCK_PLAN") else "on",
      "best ms %.3f median ms %.3f" % (1e3 * min(ts), 1e3 * float(np.median(ts))))
