"""Device-timed rate of the C4-shaped set (NSTR=32, 65 layers) -- kernel tuning runs."""
import sys; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
from sbdart_b200.timing import BatchTimer
nstr = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 16
w = workloads.mls_shortwave(nstr=nstr, nlyr=65 if nstr >= 32 else 33, wlinf=0.25, wlsup=4.0, wlinc=0.02, cloud_tau=10.0)
for k in ("dtauc", "ssalb", "pmom"):
    w[k] = np.tile(w[k], (rep,) + (1,) * (w[k].ndim - 1))
w["bins"] = np.tile(w["bins"], rep)
s = sb.Solver(0)
t = BatchTimer(s, w)
ms = t.device_ms(steps=3, warmup=1)
B = len(w["bins"])
print(f"nstr {nstr} bins {B} device {ms:.2f} ms  {B / ms * 1e3:.0f} bins/s  bad {int((t.results(True)['status'] != 0).sum())}")
