"""Small fixed radiance run (config C3 shape, NSTR=8, 10 zenith angles) of the generic kernel for ncu."""
import sys; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
w = workloads.mls_shortwave(nstr=8, wlinf=4.0, wlsup=20.0, wlinc=0.05)
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for k in ("dtauc", "ssalb", "pmom"):
    w[k] = np.tile(w[k], (rep,) + (1,) * (w[k].ndim - 1))
w["bins"] = np.tile(w["bins"], rep)
umu = np.sort(np.cos(np.deg2rad(np.linspace(5.0, 85.0, 10))))
s = sb.Solver(0)
for _ in range(2):
    out = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, temper=w["temper"],
                         umu=umu, phi=np.array([0.0]), uu_levels=[0])
print("bins", w["dtauc"].shape[0], "bad", int((out["status"] != 0).sum()))
