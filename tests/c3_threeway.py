"""Three-way flux comparison on the C3 thermal bins: generic kernel (radiance call), fast
kernel (flux call), CPU checker.  Tells which of them is the outlier."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
from solvers import make_solve_cuda, solve_oracle
from test_gpu_full_parity import C3_NIGHT, C3_DAY

s = sb.Solver(0)
for nl in (C3_NIGHT, C3_DAY):
    run = Sbdart(nl); b = run.batch(run.bins())
    gen = make_solve_cuda(s)(b)
    bf = {k: v for k, v in b.items() if k not in ("umu", "phi", "uu_levels", "corint")}
    fast = make_solve_cuda(s)(bf)
    ref = solve_oracle(bf)
    os.environ["SBD_FORCE_GENERIC"] = "1"
    genf = make_solve_cuda(s)(bf)
    del os.environ["SBD_FORCE_GENERIC"]
    scale = np.max([np.abs(ref[k]).max(axis=1) for k in ("rfldir", "rfldn", "flup")], axis=0)[:, None]
    for k in ("rfldn", "flup"):
        for name, a, c in (("generic(rad)-oracle", gen, ref), ("fast-oracle", fast, ref), ("generic(rad)-fast", gen, fast),
                           ("generic(flux)-fast", genf, fast)):
            e = np.abs(a[k] - c[k]) / scale
            i = np.unravel_index(np.argmax(e), e.shape)
            print(f"{k:6s} {name:20s} max err/scale {e.max():.2e} at bin {i[0]} level {i[1]} wl {run.last if False else ''}")
