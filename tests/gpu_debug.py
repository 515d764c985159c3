import sys, time; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
from oracle import oracle
sys.path.insert(0, 'tests')
from test_gpu_parity import oracle_flux, KEYS

s = sb.Solver(0)
w = workloads.mls_shortwave(nstr=16, wlinc=0.05)
got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
ref = oracle_flux(w)
for k in KEYS:
    err = np.abs(got[k]-ref[k])/np.maximum(np.abs(ref[k]),1e-300)
    i = np.unravel_index(np.nanargmax(err), err.shape)
    print(k, 'max rel', err[i], 'at', i, 'got', got[k][i], 'ref', ref[k][i], 'bin scale', np.abs(ref['flup'][i[0]]).max(), 'wl', w['wl'][i[0]], 'plank', w['bins']['plank'][i[0]])
# timing
for nb in (2253, 2253*16):
    w = workloads.mls_shortwave(nstr=16, replicate=nb//2253 if nb>2253 else 1)
    B = w['dtauc'].shape[0]
    t0=time.time(); got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"]); t1=time.time()
    t0=time.time(); got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"]); t1=time.time()
    print('B',B,'e2e s',t1-t0,'bins/s',B/(t1-t0), 'bad', (got['status']!=0).sum())
