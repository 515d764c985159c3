"""Radiance debug: fast-kernel radiance path vs generic kernel vs CPU checker."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
from solvers import make_solve_cuda, solve_oracle
from test_gpu_extras import CORINT_RUNS

s = sb.Solver(0)
for nl in CORINT_RUNS[:3]:
    for corint in (False, True):
        run = Sbdart(nl); b = run.batch(run.bins()); b.pop("uu_levels", None)
        b["corint"] = corint
        cpu = solve_oracle(b)
        fast = make_solve_cuda(s)(b)
        os.environ["SBD_FORCE_GENERIC"] = "1"
        gen = make_solve_cuda(s)(b)
        del os.environ["SBD_FORCE_GENERIC"]
        scale = np.abs(cpu["uu"]).max(axis=(1, 2, 3), keepdims=True)
        for name, x, y in (("fast-cpu", fast, cpu), ("gen-cpu", gen, cpu)):
            e = np.abs(x["uu"] - y["uu"]) / scale
            i = np.unravel_index(np.argmax(e), e.shape)
            print(f"corint={corint} {name}: max err/scale {e.max():.2e} at (bin,phi,lev,umu)={i} umu={b['umu'][i[3]]:.3f} "
                  f"got {x['uu'][i]:.6e} ref {y['uu'][i]:.6e} ncutinfo status {x['status'][i[0]]}")
