"""Probe: near-vertical user angle, register vs general kernel vs the CPU checker."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
from oracle import oracle
umu = np.array([float(sys.argv[1]) if len(sys.argv) > 1 else -0.99999, 0.5]); phi = np.array([25.0])
w = workloads.retrieval_batch(6, nstr=8, nlyr=7, ncols=2, seed=5)
w["bins"]["phi0"] = 0.0
res = {}
for mode in ("register", "general"):
    if mode == "general": os.environ["SBD_FORCE_GENERIC"] = "1"
    s = sb.Solver(0)
    res[mode] = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=umu, phi=phi)["uu"]
    s.close()
b = w["bins"]
for i in range(len(b)):
    r = oracle.disort(w["dtauc"][i], w["ssalb"][i], w["pmom"][i], nstr=8, umu=umu, phi=phi, fbeam=b["fbeam"][i],
                      umu0=b["umu0"][i], phi0=0.0, fisot=b["fisot"][i], albedo=b["albedo"][i], onlyfl=False)
    sc = np.abs(r["uu"]).max()
    print(i, "umu0", b["umu0"][i], "reg-chk", np.abs(res["register"][i] - r["uu"]).max() / sc, "gen-chk", np.abs(res["general"][i] - r["uu"]).max() / sc,
          "reg-gen", np.abs(res["register"][i] - res["general"][i]).max() / sc)
