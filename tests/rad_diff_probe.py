"""Where do the register and the general kernel disagree on radiances?  Worst bins against the CPU checker."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
from oracle import oracle

umu = np.array([-1.0, -0.8, -0.5, -0.2, -0.05, 0.05, 0.3, 0.6, 0.9, 1.0])
phi = np.array([0.0, 60.0, 180.0])
nstr = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
w = workloads.retrieval_batch(B, nstr=nstr, nlyr=33, ncols=8, seed=nstr)
w["bins"]["phi0"] = 30.0
res = {}
for mode in ("register", "generic"):
    if mode == "generic": os.environ["SBD_FORCE_GENERIC"] = "1"
    else: os.environ.pop("SBD_FORCE_GENERIC", None)
    s = sb.Solver(0)
    res[mode] = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr, umu=umu, phi=phi)
    res[mode + "2"] = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr, umu=umu, phi=phi)
a, g = res["register"]["uu"], res["generic"]["uu"]
print("repeatable: register", np.array_equal(a, res["register2"]["uu"]), "generic", np.array_equal(g, res["generic2"]["uu"]))
sc = np.abs(g).reshape(B, -1).max(1)
d = np.abs(a - g).reshape(B, -1).max(1) / np.maximum(sc, 1e-300)
order = np.argsort(-d)
print("bins with rel diff > 1e-6:", int((d > 1e-6).sum()), "of", B)
b = w["bins"]
for i in order[:6]:
    r = oracle.disort(w["dtauc"][i], w["ssalb"][i], w["pmom"][i], nstr=nstr, temper=w["temper"][b["col"][i]] if w.get("temper") is not None else None,
                      umu=umu, phi=phi, fbeam=b["fbeam"][i], umu0=b["umu0"][i], phi0=b["phi0"][i], fisot=b["fisot"][i],
                      albedo=b["albedo"][i], btemp=b["btemp"][i], ttemp=b["ttemp"][i], temis=b["temis"][i],
                      wvnmlo=b["wvnmlo"][i], wvnmhi=b["wvnmhi"][i], plank=bool(b["plank"][i]), onlyfl=False)
    so = np.abs(r["uu"]).max()
    print(i, "diff", d[i], "reg-vs-chk", np.abs(a[i] - r["uu"]).max() / so, "gen-vs-chk", np.abs(g[i] - r["uu"]).max() / so,
          "umu0", b["umu0"][i], "tau", w["dtauc"][i].sum(), "fbeam", b["fbeam"][i], "plank", b["plank"][i], "status", res["register"]["status"][i])
