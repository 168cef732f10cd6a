import os, sys
sys.path.insert(0, ".")
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
lat = S.sphere_cluster("bcc", 80.0)
ham = S.make_hamiltonian(lat, seed=20260101)
rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0))
rec.recur_b()
print("done", rec.launch_count)
