"""One whole SBDART run of each named namelist on the device path (K2 producer kernel,
solve kernel, CORINT kernel); run under `ncu --metrics gpu__time_duration.sum` for the
launch list in profiles/."""
import sys; sys.path.insert(0, '.')
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
import bench
RUNS = {
    "C2": bench.C2_NAMELIST,
    "C4": "&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=.25, wlsup=100, wlinc=20, iout=10 /",
    "C3": "&INPUT idatm=2, wlinf=4, wlsup=80, wlinc=20, nstr=8, iout=20, uzen=5,25,45,65,85, sza=30 /",
    "corint": "&INPUT idatm=2, nstr=16, iaer=1, vis=10, wlinf=.45, wlsup=.65, wlinc=.01, sza=55, iout=20,"
              " uzen=0,30,60,80,100,127,170, phi=0,45,90,180, corint=t /",
}
s = sb.Solver(0)
for name in sys.argv[1:] or RUNS:
    out = Sbdart(RUNS[name]).run_device(s)
    print(name, len(out.splitlines()), "lines;", out.splitlines()[-1][:80] if name != "C2" else "")
