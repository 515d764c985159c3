"""Debug: split of the fast kernel's time over its phases (needs a library built with
-DSBD_PHASE_TIMING; see the SBD_TICK macro in sbd_fast.cu)."""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from bench import build_workload
w = build_workload(8)
s = sb.Solver(0)
L = sb.lib()
t = (C.c_ulonglong * 8)()
s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
L.sbd_debug_phase_ticks(t, 1)
s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
L.sbd_debug_phase_ticks(t, 1)
tot = sum(t[:4])
for name, v in zip(("prologue", "phase 1", "phase 2", "phase 3"), t):
    print(f"{name:9s} {100.0 * v / tot:5.1f} %   {v / (w['dtauc'].shape[0] / 8.0):10.0f} ticks per CTA round-bin")
print("inside phase 3 (warp 0 of each CTA): fetch+wait %.1f %%, back substitution %.1f %%, fluxes %.1f %% of phase 3" %
      tuple(100.0 * t[i] / max(t[3], 1) for i in (4, 5, 6)))
