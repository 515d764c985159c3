"""NSTR=20 radiance batch (SBDART's default stream count for radiance output) on the radiance register kernel, for ncu."""
import sys; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
umu = np.array([-1.0, -0.8, -0.5, -0.2, -0.05, 0.05, 0.3, 0.6, 0.9, 1.0])
phi = np.array([0.0, 60.0, 180.0])
w = workloads.retrieval_batch(2048, nstr=20, nlyr=33, ncols=8, seed=20)
w["bins"]["phi0"] = 30.0
s = sb.Solver(0)
for _ in range(3):
    o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=20, umu=umu, phi=phi, uu_levels=[0], uu_packed=True)
print("bins", len(w["bins"]), "bad", int((o["status"] != 0).sum()))
