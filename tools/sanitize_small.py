"""Tiny mixed run for compute-sanitizer: flux NSTR 4..32 (adding kernel; SBD_FORCE_ELIM=1: the
elimination kernels) incl. thermal and truncated bins, a bin with a negative optical depth (handed
to the elimination kernel), USRTAU, radiances (register kernel NSTR 8/20/32 and the general kernel), BRDF surfaces (fluxes and radiances)."""
import sys; sys.path.insert(0, '.')
import numpy as np
import sbdart_b200 as sb
from sbdart_b200 import workloads
s = sb.Solver(0)
for nstr in (4, 8, 16):
    w = workloads.retrieval_batch(40, nstr=nstr, nlyr=33, ncols=4, seed=nstr)
    o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr)
    print(nstr, "retrieval bad", int((o["status"] != 0).sum()))
w = workloads.mls_shortwave(nstr=16, wlinc=0.25)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"])
print("thermal bad", int((o["status"] != 0).sum()))
tot = w["dtauc"].sum(axis=1)
utau = np.stack([0 * tot, 0.3 * tot, tot], axis=1)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16, temper=w["temper"], utau=utau)
print("usrtau bad", int((o["status"] != 0).sum()))
# NSTR 20 / 24 / 32: the CTA-per-bin register kernel (beam, and beam + Planck on a 65-layer grid)
for nstr in (20, 24, 32):
    w = workloads.retrieval_batch(12, nstr=nstr, nlyr=33, ncols=3, seed=nstr)
    o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr)
    print(nstr, "wide retrieval bad", int((o["status"] != 0).sum()))
w = workloads.mls_shortwave(nstr=32, nlyr=65, wlinf=1.8, wlsup=2.2, wlinc=0.1, cloud_tau=10.0)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=32, temper=w["temper"])
print("wide thermal bad", int((o["status"] != 0).sum()))
# radiances: register kernel (NSTR 8, 20, 32; azimuth modes, packed levels) and general kernel
w = workloads.retrieval_batch(6, nstr=8, nlyr=6, ncols=2, seed=3)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=np.array([-0.5, 0.5]), phi=np.array([0.0, 90.0]),
                   uu_levels=[0, 6], uu_packed=True)
print("packed radiance bad", int((o["status"] != 0).sum()))
w = workloads.retrieval_batch(4, nstr=20, nlyr=6, ncols=2, seed=5)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=20, umu=np.array([-0.5, 0.5]), phi=np.array([0.0, 90.0]))
print("20 radiance bad", int((o["status"] != 0).sum()))
import os
os.environ["SBD_FORCE_GENERIC"] = "1"
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=20, umu=np.array([-0.5, 0.5]), phi=np.array([0.0, 90.0]))
del os.environ["SBD_FORCE_GENERIC"]
print("generic radiance bad", int((o["status"] != 0).sum()))
w = workloads.retrieval_batch(3, nstr=32, nlyr=5, ncols=2, seed=6)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=32, umu=np.array([-0.5, 0.5]), phi=np.array([0.0]))
print("32 radiance bad", int((o["status"] != 0).sum()))
w = workloads.retrieval_batch(6, nstr=8, nlyr=6, ncols=2, seed=3)
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=np.array([-0.5, 0.5]), phi=np.array([0.0, 90.0]))
print("radiance bad", int((o["status"] != 0).sum()))
w["dtauc"][2, 3] = -0.01       # radiance bin handed to the general kernel
o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=np.array([-0.5, 0.5]), phi=np.array([0.0, 90.0]))
print("radiance with a handed-over bin bad", int((o["status"] != 0).sum()))
# a negative optical depth: TAUC is not monotone, the adding kernel hands the bin over
for nstr in (16, 32):
    w = workloads.retrieval_batch(8, nstr=nstr, nlyr=12, ncols=2, seed=11)
    w["dtauc"][3, 5] = -0.01
    o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr)
    print(nstr, "handed-over bin bad", int((o["status"] != 0).sum()))
# BRDF surfaces: fluxes (adding kernel) and radiances (register kernel)
from sbdart_b200.frontend import brdf
model = brdf.SurfaceModel(9, [0.2, 0.1, 0.05, 1.5, 2.0])
umu = np.array([-0.5, 0.3, 0.8])
for nstr, rad in ((8, False), (20, False), (8, True), (20, True)):
    mu, _ = sb.quadrature(nstr // 2)
    w = workloads.retrieval_batch(4, nstr=nstr, nlyr=6, ncols=1, seed=7)
    umu0 = float(w["bins"]["umu0"][0])
    tab = brdf.surface_tables(model, None, mu, umu0, True, nstr if rad else 1, umu=umu if rad else None)
    s.set_surfaces(nstr, tab["bdr"][None], tab["bem"][None], tab["rmu"][None] if rad else None, tab["emu"][None] if rad else None)
    b = w["bins"].copy(); b["albedo"] = sb.surface_albedo(0)
    o = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], b, nstr=nstr, umu=umu if rad else None,
                       phi=np.array([0.0, 90.0]) if rad else None)
    print(nstr, "brdf", "radiance" if rad else "flux", "bad", int((o["status"] != 0).sum()))
    s.set_surfaces()
s.close()
