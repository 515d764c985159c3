import sys, time
sys.path.insert(0, ".")
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
s = sb.Solver(0)
for nstr in (20, 16, 8):
    nml = f"&INPUT\n idatm=2, wlinf=.4, wlsup=.9, wlinc=.0025, iout=5, sza=40, nstr={nstr}, uzen=0,20,40,60,80,100,120,140,160,180, phi=0,45,90,135,180, tcloud=5, zcloud=2\n /"
    for rep in range(2):
        r = Sbdart(nml)
        t0 = time.time(); out = r.run_device(s); t1 = time.time()
    rows = len(r.bins()) if False else 0
    l0 = s.kernel_launches
    print("nstr", nstr, "whole run_device", round(t1 - t0, 3), "s", len(out.splitlines()), "lines")
