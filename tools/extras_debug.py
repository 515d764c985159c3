"""Worst GPU-vs-checker mismatches of selected namelist runs (debug aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbdart_b200 as sb                      # noqa: E402
from sbdart_b200.frontend import Sbdart      # noqa: E402
from solvers import make_solve_cuda, solve_oracle  # noqa: E402

RUNS = [
    "&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=8, wlsup=12, wlinc=20, iout=10 /",
    "&INPUT idatm=2, nstr=8, tcloud=2, zcloud=2, wlinf=.5, wlsup=.7, wlinc=.05, sza=30, iout=21,"
    " uzen=100,140,145,150,155,160, phi=0,10,180, corint=t /",
]
s = sb.Solver(0)
for nl in RUNS:
    run = Sbdart(nl)
    b = run.batch(run.bins())
    b.pop("uu_levels", None)
    g, c = make_solve_cuda(s)(b), solve_oracle(b)
    print(nl[:90])
    keys = ["rfldir", "rfldn", "flup"] + (["uu"] if "uu" in c else [])
    for k in keys:
        ax = tuple(range(1, c[k].ndim))
        scale = np.abs(c[k]).max(axis=ax, keepdims=True)
        d = np.abs(g[k] - c[k])
        rel = d / np.maximum(np.abs(c[k]), 1e-300)
        viol = d - 1e-5 * np.abs(c[k])
        i = np.unravel_index(np.argmax(viol / scale), d.shape)
        print(f"  {k}: worst (abs err - 1e-5|ref|)/binscale = {(viol / scale)[i]:.3e} at {i}: gpu {g[k][i]:.9e} cpu {c[k][i]:.9e}"
              f" binscale {scale[i[0]].ravel()[0]:.3e}")
    if "uu" in c:
        bad = np.argwhere(np.abs(g["uu"] - c["uu"]) > 1e-5 * np.abs(c["uu"]) + 1e-9 * np.abs(c["uu"]).max(axis=(1, 2, 3), keepdims=True))
        print("  uu violations:", len(bad), bad[:12].tolist(), "umu", b["umu"], "phi", b["phi"])
        for i in bad[:6]:
            i = tuple(i)
            print("   ", i, g["uu"][i], c["uu"][i])
s.close()
