"""Small fixed run for ncu: a few launches of the hot kernel on a reduced batch."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import sbdart_b200 as sb
from bench import build_workload
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = build_workload(rep)
s = sb.Solver(0)
for _ in range(3):
    out = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
print("bins", w["dtauc"].shape[0], "bad", int((out["status"] != 0).sum()))
