"""Small fixed run of the generic kernel (NSTR=32, 65 layers) for ncu."""
import sys; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
w = workloads.mls_shortwave(nstr=32, nlyr=65, wlinf=0.25, wlsup=4.0, wlinc=0.02, cloud_tau=10.0)
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for k in ("dtauc", "ssalb", "pmom"):
    w[k] = np.tile(w[k], (rep,) + (1,) * (w[k].ndim - 1))
w["bins"] = np.tile(w["bins"], rep)
s = sb.Solver(0)
for _ in range(2):
    out = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=32, temper=w["temper"])
print("bins", w["dtauc"].shape[0], "bad", int((out["status"] != 0).sum()))
