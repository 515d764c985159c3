"""External-file modes of the front end: usrcld.dat (usrcloud, taucloud.f:142-274),
aerosol.dat (aeread, tauaero.f:1526-1714) and the k-distribution files CKATM / CKTAU
(gasinit / readk, taugas.f:7297-7389, :7695-7792).  The reference ships no fixture for any of
them: each is checked against the same physics entered through the NAMELIST, run through the
CPU checker."""
import os

import numpy as np
import pytest

from sbdart_b200.frontend import Sbdart, extras
from solvers import solve_oracle


@pytest.fixture
def workdir(tmp_path):
    old = os.getcwd()
    os.chdir(tmp_path)
    yield tmp_path
    os.chdir(old)


def test_usrcld_matches_the_namelist_cloud(workdir):
    """A water path in one layer of usrcld.dat = the same lwp / nre / zcloud given in INPUT."""
    base = "&INPUT\n idatm=4, isat=0, wlinf=.5, wlsup=.7, wlinc=.05, iout=1, sza=30,\n {}\n /"
    ref_run = Sbdart(base.format("zcloud=3, lwp=40, nre=8"))
    ref = ref_run.run(solve_oracle)
    nz = ref_run.nz
    lay = ref_run.clouds.lcld[0]                   # 1 = top layer
    lines = ["/"] * (nz - lay) + ["40. 8. 0. -1. 1."]       # bottom-up; a slash keeps the defaults
    (workdir / "usrcld.dat").write_text("\n".join(lines) + "\n")
    got = Sbdart(base.format("nre=0")).run(solve_oracle)
    assert got == ref
    # cloud fraction: optical depth x cldfrac**1.5 (taucloud.f:261)
    lines[-1] = "40. 8. 0. -1. 0.5"
    (workdir / "usrcld.dat").write_text("\n".join(lines) + "\n")
    half = Sbdart(base.format("nre=0")).run(solve_oracle)
    eq = Sbdart(base.format(f"zcloud=3, lwp={40 * 0.5 ** 1.5!r}, nre=8")).run(solve_oracle)
    from sbchk_cases import compare_records
    compare_records(half, eq, rel=1e-4)       # (the upward flux over the black surface is round-off)
    (workdir / "usrcld.dat").write_text("0. 8. 5. 30. 1.\n")
    with pytest.raises(NotImplementedError, match="rhoice"):
        Sbdart(base.format("nre=0")).run(solve_oracle)


def test_aerosol_dat_interpolation_and_run(workdir):
    """aeread: log-wavelength interpolation (geometric for the optical depth), spectrally
    uniform single record, getmom for nmom = 1."""
    nz = 33
    (workdir / "aerosol.dat").write_text(
        "2 1\n0.4\n0.02 0.9 0.6\n0.20 0.95 0.7\n0.8\n0.01 0.8 0.5\n0.05 0.85 0.65\n")
    af = extras.AerosolFile(nz, 3)
    pm = np.zeros((nz, 7))
    dt, w = af(0.4 * 2 ** 0.5, 6, pm)                     # half way in log wavelength
    assert np.allclose(dt[-2:], [np.sqrt(0.02 * 0.01), np.sqrt(0.20 * 0.05)]) and (dt[:-2] == 0).all()
    assert np.allclose(w[-2:], [0.85, 0.9])
    g = 0.5 * (0.7 + 0.65)
    assert np.allclose(pm[-1, 1:], g ** np.arange(1, 7) * dt[-1] * w[-1])
    dt2, _ = af(5.0, 6, np.zeros((nz, 7)))                # beyond the file: the last record
    assert np.allclose(dt2[-2:], [0.01, 0.05])
    # a whole run: more extinction -> less direct beam at the surface, same top-of-atmosphere input
    base = "&INPUT\n idatm=4, isat=0, wlinf=.5, wlsup=.6, wlinc=.05, iout=10, sza=30,\n {}\n /"
    clear = np.array(Sbdart(base.format("iaer=0")).run(solve_oracle).split(), float)
    hazy = np.array(Sbdart(base.format("iaer=-1")).run(solve_oracle).split(), float)
    assert hazy[3] == clear[3] and hazy[8] < clear[8] and hazy[4] > clear[4]
    # single record = spectrally uniform
    (workdir / "aerosol.dat").write_text("1 1\n0.55\n0.3 0.9 0.7\n")
    af = extras.AerosolFile(nz, 3)
    for wl in (0.3, 0.6, 2.0):
        dt, w = af(wl, 6, np.zeros((nz, 7)))
        assert dt[-1] == 0.3 and w[-1] == 0.9


def test_cktau_reproduces_the_band_model_run(workdir):
    """kdist = -1: CKATM / CKTAU written from a kdist = 1 run (its atmosphere, spectral limits,
    k-weights and gas optical depths) must give back that run's records."""
    nml = "&INPUT\n idatm=4, isat=0, wlinf=.9, wlsup=1.0, wlinc=.02, iout=1, sza=30, kdist=1,\n nf=0\n /"
    src = Sbdart(nml)
    ref = src.run(solve_oracle)
    rows = src.bins()
    z, pr, t = src.z, src.pr, src.t
    with open("CKATM", "w") as f:
        f.write(f"{len(z)} {float(src.wh[0])!r}\n")
        for arr in (z, pr, t):
            f.write(" ".join(repr(float(v)) for v in arr) + "\n")
    recs = []
    by_wl = {}
    for r in rows:
        by_wl.setdefault(r["il"], []).append(r)
    from sbdart_b200.frontend import rayleigh
    for il in sorted(by_wl, reverse=True):            # CKTAU runs from high to low wavenumber
        rr = by_wl[il]
        dtaur = rayleigh(rr[0]["wl"], src.z, src.pr, src.t)
        # gas optical depth of every k-term = total - Rayleigh (no clouds, no aerosols in this run)
        recs.append(dict(iv=il, ib=1, nb=1, vnu0=10000. / rr[0]["wl"], vnu1=rr[0]["wvnmlo"], vnu2=rr[0]["wvnmhi"],
                         etf=0.0, ewc=1.0, gw=[r["wt"] for r in rr],
                         dtk=np.stack([np.maximum(r["dtau"] - dtaur, 0.0) for r in rr], axis=1)))
    extras.write_cktau("CKTAU", recs)
    # (the REAL*4 wavenumbers of the file must fall inside the requested range: widen it a little)
    got = Sbdart(nml.replace("kdist=1", "kdist=-1").replace("wlinf=.9", "wlinf=.895").replace("wlsup=1.0", "wlsup=1.005")
                 ).run(solve_oracle)
    a, b = got.splitlines(), ref.splitlines()
    assert a[:3] == b[:3] and len(a) == len(b)
    # CKTAU holds REAL*4 values and lists the spectrum from the blue end
    va = np.array([l.split() for l in a[3:]], float)
    vb = np.array([l.split() for l in b[3:]], float)[::-1]
    assert np.allclose(va, vb, rtol=2e-4, atol=1e-12)
