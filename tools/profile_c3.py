"""C3 namelist (sunlit: 8 azimuth modes, 10 zenith x 19 azimuth angles) on the radiance register kernel, for ncu."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import sbdart_b200 as sb
from bench_configs import namelist_workload, tile
sza = sys.argv[1] if len(sys.argv) > 1 else "30"
nl = f"&INPUT idatm=2, wlinf=4, wlsup=80, wlinc=20, nstr=8, iout=20, sza={sza}, uzen=0,20,40,60,80,100,120,140,160,180 /"
w = tile(namelist_workload(nl), 16)
s = sb.Solver(0)
for _ in range(3):
    o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=w["nstr"], temper=w["temper"], umu=w["umu"], phi=w["phi"],
                       uu_levels=w.get("uu_levels"), uu_packed=True)
print("bins", len(w["bins"]), "bad", int((o["status"] != 0).sum()))
