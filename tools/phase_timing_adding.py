"""Debug: split of the adding kernel's time over its phases (library built with
-DSBD_PHASE_TIMING: tools/build_variant.py addtiming sbd_adding.cu -DSBD_PHASE_TIMING,
then SBD_LIB_PATH=tools/experiments/libsbd_addtiming.so SBD_SKIP_BUILD_ID_CHECK=1)."""
import ctypes as C, sys
sys.path.insert(0, '.')
import sbdart_b200 as sb
from bench import build_workload
w = build_workload(8)
s = sb.Solver(0)
L = sb.lib()
t = (C.c_ulonglong * 8)()
s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
L.sbd_debug_add_ticks(t, 1)
s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
L.sbd_debug_add_ticks(t, 1)
tot = sum(t[:4])
for name, v in zip(("prologue", "phase 1", "phase 2", "phase 3"), t):
    print(f"{name:9s} {100.0 * v / tot:5.1f} %   {v / (w['dtauc'].shape[0] / 8.0):10.0f} ticks per CTA round-bin")
print("Jacobi sweeps per layer round: %.2f (%d rounds)" % (t[6] / max(t[7], 1), t[7]))
