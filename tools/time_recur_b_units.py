#!/usr/bin/env python
"""recur_b on the config-1 lattice (5984 sites) and the config-3 lattice with 1, 2, 4, 8 recursion sites in one batch, pipelined
step forced on / off (RSREC_LZ_PIPELINE): where the pipelined form stops paying.  Best of N host-timed calls."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
    out = {}
    for name, lat in (("config1", S.sphere_cluster("bcc", 80.0)), ("config3", S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2"))):
        ham = S.make_hamiltonian(lat, seed=20260101)
        for nu in (1, 2, 4, 8):
            lat.irec = np.arange(1, nu + 1, dtype=np.int32)
            rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0))
            for _ in range(3): rec.recur_b()
            ts = []
            for _ in range(12):
                t0 = time.perf_counter(); rec.recur_b(); ts.append(time.perf_counter() - t0)
            out["%s x%d" % (name, nu)] = round(1e3 * min(ts), 3)
            rec.close()
    print(json.dumps(out))
else:
    for v in ("1", "0"):
        env = dict(os.environ, RSREC_LZ_PIPELINE=v)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("RSREC_LZ_PIPELINE=" + v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
