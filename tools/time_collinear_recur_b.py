import os,sys,time
sys.path.insert(0,".")
import numpy as np
from rslmtoasa_b200 import Recursion, Control, Energy, Green, synthetic as S
lat = S.sphere_cluster("bcc", 80.0)
ham = S.make_hamiltonian(lat, seed=20260101, spin_orbit=False)
rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0, channels_ldos=2500, fermi=0.0))
rec.recur_b()
ts=[]
for _ in range(15):
    t0=time.perf_counter(); rec.recur_b(); ts.append(time.perf_counter()-t0)
g = Green(rec); g.recur_b_green(download_g0=False); tg=[]
for _ in range(15):
    t0=time.perf_counter(); g.recur_b_green(download_g0=False); tg.append(time.perf_counter()-t0)
print("SD_WARPS", os.environ.get("RSREC_SD_WARPS","8"), "recur_b best ms %.3f median %.3f"%(1e3*min(ts),1e3*float(np.median(ts))), "recur_b_green best %.3f"%(1e3*min(tg)), "sd launches", int(rec._L.rsrec_spin_diag_launch_count(rec._h)))
