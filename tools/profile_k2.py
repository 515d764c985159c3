"""Whole-spectrum columns call (K2 + solve) on a reduced batch, for ncu: tools/profile_k2.py [columns]."""
import sys; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
from sbdart_b200.frontend.device import ColumnRunner
from bench import C2_NAMELIST
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 16
run = Sbdart(C2_NAMELIST)
s = sb.Solver(0)
cr = ColumnRunner(run, s, ncol, levels=[run.ntop - 1, run.nbot - 1])
for _ in range(3):
    cr.step()
print("columns", ncol, "bins", cr.nbins.value)
