"""Radiance bins/s by NSTR: register kernel (adding form) against the general kernel (SBD_FORCE_GENERIC=1)."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import sbdart_b200 as sb
from sbdart_b200 import workloads

umu = np.array([-1.0, -0.8, -0.5, -0.2, -0.05, 0.05, 0.3, 0.6, 0.9, 1.0])
phi = np.array([0.0, 60.0, 180.0])
top_only = len(sys.argv) > 1 and sys.argv[1] == 'top'      # top level only: what SBDART's iout=5/20 consume
for nstr in (16, 20, 24, 32):
    B = 4096 if nstr <= 20 else 2048
    w = workloads.retrieval_batch(B, nstr=nstr, nlyr=33, ncols=8, seed=nstr)
    w["bins"]["phi0"] = 30.0
    for mode in ("register", "generic"):
        if mode == "generic": os.environ["SBD_FORCE_GENERIC"] = "1"
        else: os.environ.pop("SBD_FORCE_GENERIC", None)
        s = sb.Solver(0)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr, umu=umu, phi=phi,
                                 uu_levels=[0] if top_only else None, uu_packed=top_only)
            torch.cuda.synchronize(); best = min(best, time.time() - t0)
        if mode == "register": ref = got
        else:
            sc = np.abs(got["uu"]).max()
            print("   max |register - generic| / max", float(np.abs(ref["uu"] - got["uu"]).max() / sc), "status", int((ref["status"] != got["status"]).sum()))
        print(f"nstr {nstr} {mode:8s} {B / best:10.0f} bins/s (host-buffer call, best of 3)")
