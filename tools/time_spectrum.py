"""Wall time of one whole SBDART run (config C2 namelist) on the GPU path: optical
properties produced by the K2 kernel, solved by the batched kernel, records formatted."""
import sys, time; sys.path.insert(0, '.')
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
import bench
s = sb.Solver(0)
t0 = time.perf_counter(); run = Sbdart(bench.C2_NAMELIST); t1 = time.perf_counter()
out = run.run_device(s); t2 = time.perf_counter()
ts = []
for _ in range(5):
    run = Sbdart(bench.C2_NAMELIST)
    ta = time.perf_counter(); out = run.run_device(s); ts.append(time.perf_counter() - ta)
print(f"front-end setup {1e3*(t1-t0):.1f} ms; first run_device {1e3*(t2-t1):.1f} ms; steady run_device {1e3*min(ts):.1f} ms "
      f"({len(out.splitlines())} output lines)")
from sbdart_b200.frontend import device
import time as _t
run = Sbdart(bench.C2_NAMELIST)
ta = _t.perf_counter(); rows, res = device.run_spectrum(run, s); tb = _t.perf_counter()
print(f"run_spectrum (K2 + solve + copies) {1e3*(tb-ta):.1f} ms for {len(rows)} bins -> {len(rows)/(tb-ta):.0f} bins/s")
