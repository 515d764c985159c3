"""Debug: split of the radiance register kernel's time over its phases (library built with
-DSBD_PHASE_TIMING: tools/build_variant.py radtiming sbd_fast.cu -DSBD_PHASE_TIMING, then
SBD_LIB_PATH=tools/experiments/libsbd_radtiming.so SBD_SKIP_BUILD_ID_CHECK=1)."""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
nstr = int(sys.argv[1]) if len(sys.argv) > 1 else 16
top_only = len(sys.argv) > 2 and sys.argv[2] == 'top'      # what SBDART's iout=5/20 consume
umu = np.array([-1.0, -0.8, -0.5, -0.2, -0.05, 0.05, 0.3, 0.6, 0.9, 1.0])
phi = np.array([0.0, 60.0, 180.0])
w = workloads.retrieval_batch(4096, nstr=nstr, nlyr=33, ncols=8, seed=nstr)
w["bins"]["phi0"] = 30.0
s = sb.Solver(0)
L = sb.lib()
t = (C.c_ulonglong * 8)()
for rep in range(2):
    s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr, umu=umu, phi=phi,
                   uu_levels=[0] if top_only else None, uu_packed=top_only)
    L.sbd_debug_phase_ticks(t, 1)
tot = sum(t[:4])
for name, v in zip(("prologue", "phase 1", "phase 2", "phase 3"), t):
    print(f"{name:9s} {100.0 * v / tot:5.1f} %")
print("inside phase 3 (warp 0 of each CTA): fetch+wait %.1f %%, layer solution %.1f %%, fluxes %.1f %%, user angles %.1f %% of phase 3" %
      tuple(100.0 * t[i] / max(t[3], 1) for i in (4, 5, 6, 7)))
