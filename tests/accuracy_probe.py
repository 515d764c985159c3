"""Error of the CUDA path against the CPU oracle on the C2 bench spectrum, relative to
each bin's largest flux (what the 5-digit SBDART records can resolve is ~1e-5)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import sbdart_b200 as sb
from oracle import oracle

if len(sys.argv) > 2 and sys.argv[2] == "smoke":
    from sbdart_b200 import workloads
    w = workloads.mls_shortwave(nstr=16, wlinc=0.25)        # the workload of __graft_entry__.smoke()
else:
    w = bench.build_workload(1)
idx = np.arange(0, w["dtauc"].shape[0], int(sys.argv[1]) if len(sys.argv) > 1 else 7)
b = w["bins"][idx]
s = sb.Solver(0)
got = s.disort_batch(w["dtauc"][idx], w["ssalb"][idx], w["pmom"][idx], b, nstr=16, temper=w["temper"])
ref = oracle.disort_flux_batch(
    w["dtauc"][idx], w["ssalb"][idx], w["pmom"][idx], nstr=16, fbeam=b["fbeam"], umu0=b["umu0"],
    albedo=b["albedo"], plank=b["plank"], wvnmlo=b["wvnmlo"], wvnmhi=b["wvnmhi"], btemp=b["btemp"],
    ttemp=b["ttemp"], temis=b["temis"], fisot=b["fisot"], temper=w["temper"], col=b["col"],
    nthreads=os.cpu_count() or 1)
assert (got["status"] == ref["status"]).all()
scale = np.max([np.abs(ref[k]).max(axis=1) for k in ("rfldir", "rfldn", "flup")], axis=0)[:, None]
for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg"):
    e = np.abs(got[k] - ref[k])
    big = np.abs(ref[k]) > 1e-6 * scale
    print(f"{k:7s} max|err|/scale = {np.max(e / scale):.2e}   max rel err where |ref| > 1e-6 scale = "
          f"{np.max(e[big] / np.abs(ref[k][big])) if big.any() else 0:.2e}   worst bin {idx[np.argmax((e / scale).max(axis=1))]}")
