import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
from sbdart_b200.frontend.device import run_spectrum
nl = sys.argv[1] if len(sys.argv)>1 else "&INPUT tcloud=0, zcloud=8, nre=10, idatm=4, sza=95, wlinf=4, wlsup=20, wlinc=-.01, iout=1 /"
s = sb.Solver(0)
run = Sbdart(nl); rows0 = run.bins(); ref = run.batch(rows0)
rows, res, dev = run_spectrum(Sbdart(nl), s, want_inputs=True)
print('bins', len(rows), len(rows0))
for k in ("dtauc","ssalb","pmom"):
    d = np.abs(dev[k]-ref[k]); r = d/np.maximum(np.abs(ref[k]),1e-300)
    i = np.unravel_index(np.argmax(d), d.shape)
    print(k, 'max abs', d.max(), 'at', i, dev[k][i], ref[k][i], 'bins with rel>1e-9:', np.unique(np.nonzero((r>1e-9)&(d>1e-14))[0])[:20])
for f in ("fbeam","umu0","albedo","wvnmlo","wvnmhi"):
    print(f, np.abs(dev['bins'][f]-ref['bins'][f]).max())
print([ (r['il'],r['kd'],r['nk']) for r in rows[:6]], [ (r['il'],r['kd'],r['nk']) for r in rows0[:6]])
print('wt', [r['wt'] for r in rows[:6]], [r['wt'] for r in rows0[:6]])
b=0
print('dtauc dev', dev['dtauc'][b][:8]); print('dtauc ref', ref['dtauc'][b][:8])
