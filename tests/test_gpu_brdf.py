"""BRDF surfaces (LAMBER = .FALSE.; SURFAC disort.f:3765-3907 with the BDREF models of
spectra.f:249-1357) on the GPU against the CPU oracle: the adding kernel (fluxes) and the
general kernel (radiances, every azimuth mode), and the Fortran entry with a BDREF callback."""
import ctypes as C

import numpy as np
import pytest

import sbdart_b200 as sb
from oracle import oracle
from sbdart_b200.frontend import brdf
from sbdart_b200.frontend import tables

pytestmark = pytest.mark.gpu

MODELS = [(8, [0.6, 0.3, 0.4, 0.1]), (9, [0.2, 0.1, 0.05, 1.5, 2.0]), (7, [0.5, 6.0, 34.3])]


def _atm(nstr, L, rng):
    dt = 10 ** rng.uniform(-2.5, 0.3, L)
    ss = 1 - 10 ** rng.uniform(-3, -0.3, L)
    g = rng.uniform(0.1, 0.85, L)
    pm = g[:, None] ** np.arange(nstr + 3)[None, :]
    return dt, ss, pm


def _set_oracle(model, state):
    oracle.set_bdref(model.ibdrf, model.params(), *((state["nr"], state["ni"], state["rsw"]) if state else (0, 0, 0)))


@pytest.mark.parametrize("isalb,sc", MODELS)
@pytest.mark.parametrize("nstr", [8, 16, 20])
def test_flux_with_brdf_surface(isalb, sc, nstr):
    """NSTR 8/16: adding kernel; NSTR 20: general kernel."""
    rng = np.random.default_rng(isalb * 100 + nstr)
    L = 9
    model = brdf.SurfaceModel(isalb, sc)
    state = brdf.ocean_state(tables(), model, 0.55) if model.spectral else None
    _set_oracle(model, state)
    mu, _ = sb.quadrature(nstr // 2)
    s = sb.Solver(0)
    for umu0, plank in ((0.7, False), (0.35, True)):
        dt, ss, pm = _atm(nstr, L, rng)
        temper = np.linspace(240, 295, L + 1)
        kw = dict(fbeam=1.0, umu0=umu0, fisot=0.0)
        if plank:
            kw.update(wvnmlo=18100.0, wvnmhi=18260.0, btemp=300.0, ttemp=0.0, temis=0.0)
        ref = oracle.disort(dt, ss, pm, nstr=nstr, lamber=False, plank=plank, temper=temper, **kw)
        assert ref["status"] == 0
        tab = brdf.surface_tables(model, state, mu, umu0, True, 1)
        s.set_surfaces(nstr, tab["bdr"][None], tab["bem"][None])
        bins = sb.make_bins(1, albedo=sb.surface_albedo(0), plank=int(plank), **kw)
        got = s.disort_batch(dt[None], ss[None], pm[None], bins, nstr=nstr, temper=temper[None])
        assert got["status"][0] == 0
        sc_ = max(np.abs(ref[k]).max() for k in ("rfldir", "rfldn", "flup"))
        for k in ("rfldir", "rfldn", "flup", "uavg"):
            assert np.abs(got[k][0] - ref[k]).max() <= 1e-8 * sc_, (k, umu0)
    s.set_surfaces()
    s.close()


@pytest.mark.parametrize("isalb,sc", MODELS)
@pytest.mark.parametrize("nstr,variant", [(8, "beam"), (20, "beam"), (8, "thermal"), (16, "thick")])
def test_radiances_with_brdf_surface(isalb, sc, nstr, variant):
    """beam: sunlit; thermal: surface emission EMU B(Ts) and an emitting top boundary besides the beam;
    thick: an absorbing layer deep enough for the layer truncation (the surface is then not seen)."""
    rng = np.random.default_rng(isalb * 10 + nstr)
    L = 6
    model = brdf.SurfaceModel(isalb, sc)
    state = brdf.ocean_state(tables(), model, 0.55) if model.spectral else None
    _set_oracle(model, state)
    mu, _ = sb.quadrature(nstr // 2)
    umu = np.array([-0.9, -0.4, 0.15, 0.5, 0.95])
    phi = np.array([0.0, 45.0, 120.0, 180.0])
    dt, ss, pm = _atm(nstr, L, rng)
    umu0 = 0.6
    kw = dict(fbeam=1.0, umu0=umu0, phi0=20.0)
    okw = dict(kw)
    temper = np.linspace(230.0, 290.0, L + 1)
    if variant == "thermal":
        kw.update(wvnmlo=2400.0, wvnmhi=2600.0, btemp=305.0, ttemp=250.0, temis=0.4, fisot=0.01)
        okw = dict(kw, plank=True, temper=temper)
    if variant == "thick":
        dt[3] = 40.0; ss[3] = 0.3
    ref = oracle.disort(dt, ss, pm, nstr=nstr, lamber=False, onlyfl=False, umu=umu, phi=phi, **okw)
    assert ref["status"] == 0
    tab = brdf.surface_tables(model, state, mu, umu0, True, nstr, umu=umu)
    s = sb.Solver(0)
    s.set_surfaces(nstr, tab["bdr"][None], tab["bem"][None], tab["rmu"][None], tab["emu"][None])
    bins = sb.make_bins(1, albedo=sb.surface_albedo(0), plank=int(variant == "thermal"), **kw)
    got = s.disort_batch(dt[None], ss[None], pm[None], bins, nstr=nstr, umu=umu, phi=phi,
                         temper=temper[None] if variant == "thermal" else None)
    s.set_surfaces()
    s.close()
    assert got["status"][0] == 0
    assert np.abs(got["flup"][0] - ref["flup"]).max() <= 1e-8 * np.abs(ref["flup"]).max()
    assert np.abs(got["uu"][0] - ref["uu"]).max() <= 1e-7 * np.abs(ref["uu"]).max()


def test_fortran_entry_calls_the_hosts_bdref():
    """disort_ with LAMBER = .FALSE.: the library runs SURFAC with the host program's BDREF
    (here a ctypes callback: the Ross-Li model of spectra.f:350)."""
    model = brdf.SurfaceModel(9, [0.2, 0.1, 0.05, 1.5, 2.0])
    _set_oracle(model, None)
    dp = C.POINTER(C.c_double)
    proto = C.CFUNCTYPE(C.c_double, dp, dp, dp, dp, dp)

    def bd(wlo, whi, mur, mui, phir):
        return float(brdf.bdref(model, None, mur[0], mui[0], phir[0]))

    cb = proto(bd)
    L_ = sb.lib()
    L_.sbd_set_bdref_callback(C.cast(cb, C.c_void_p))
    nstr, L = 8, 4
    rng = np.random.default_rng(3)
    dt, ss, pm = _atm(nstr, L, rng)
    umu = np.array([-0.7, 0.3, 0.8])
    phi = np.array([0.0, 90.0])
    out = sb.disort(L, dt, ss, nstr + 2, pm.T, np.zeros(L + 1), 0.0, 0.0, False, 0, None, nstr, True, 3, umu, 2, phi,
                    0, 1.0, 0.6, 10.0, 0.0, False, 0.0, 0.0, 0.0, 0.0, False, False, 0.0)
    L_.sbd_set_bdref_callback(None)
    ref = oracle.disort(dt, ss, pm, nstr=nstr, lamber=False, onlyfl=False, umu=umu, phi=phi, fbeam=1.0, umu0=0.6, phi0=10.0)
    assert out["status"] == 0
    assert np.abs(out["flup"] - ref["flup"]).max() <= 1e-8 * np.abs(ref["flup"]).max()
    uu = np.transpose(out["uu"], (2, 1, 0))[:, :, :3]      # UU(iu, lu, j) -> [j][lu][iu]
    assert np.abs(uu - ref["uu"]).max() <= 1e-7 * np.abs(ref["uu"]).max()


@pytest.mark.parametrize("nml", [
    "&INPUT\n idatm=4, isat=0, wlinf=.4, wlsup=.7, wlinc=.02, iout=1, isalb=7, sc=0.5,6.0,34.3, sza=40, nstr=8\n /",
    "&INPUT\n idatm=4, isat=0, wlinf=.5, wlsup=.6, wlinc=.05, iout=20, isalb=7, sc=0.5,6.0,34.3, sza=40, nstr=8,"
    " uzen=0,30,60, phi=0,90,180\n /",
    "&INPUT\n idatm=4, isat=0, wlinf=.5, wlsup=.6, wlinc=.05, iout=21, isalb=8, sc=0.6,0.3,0.4,0.1, sza=55,"
    " uzen=100,130,160,180, phi=0,60,120,180\n /"])
def test_whole_runs_with_brdf_surfaces(nml):
    """isalb = 7 / 8 whole runs (ocean: one surface table per wavelength; iout = 21: SBDART's
    default NSTR = 20 on the general kernel) through Sbdart.run_device: records identical to the
    CPU checker's up to the last printed digit."""
    from sbchk_cases import compare_records
    from sbdart_b200.frontend import Sbdart
    from solvers import solve_oracle
    ref = Sbdart(nml).run(solve_oracle)
    s = sb.Solver(0)
    got = Sbdart(nml).run_device(s)
    s.close()
    nval, nexact, worst = compare_records(got, ref, rel=1.01e-4)      # one unit of the last digit
    assert nexact >= 0.98 * nval, (nval, nexact, worst)


@pytest.mark.parametrize("nstr", [8, 20])
def test_brdf_flux_bin_with_a_negative_optical_depth(nstr):
    """A BRDF flux bin whose TAUC is not monotone: the adding kernel hands it over, and because the
    elimination kernels are Lambertian it is the general kernel that solves it."""
    isalb, sc = MODELS[0]
    rng = np.random.default_rng(1000 + nstr)
    L = 9
    model = brdf.SurfaceModel(isalb, sc)
    state = brdf.ocean_state(tables(), model, 0.55) if model.spectral else None
    _set_oracle(model, state)
    mu, _ = sb.quadrature(nstr // 2)
    dt, ss, pm = _atm(nstr, L, rng)
    dt[4] = -0.2 * dt[3]
    kw = dict(fbeam=1.0, umu0=0.6, fisot=0.0)
    ref = oracle.disort(dt, ss, pm, nstr=nstr, lamber=False, **kw)
    assert ref["status"] == 0
    tab = brdf.surface_tables(model, state, mu, 0.6, True, 1)
    s = sb.Solver(0)
    s.set_surfaces(nstr, tab["bdr"][None], tab["bem"][None])
    bins = sb.make_bins(1, albedo=sb.surface_albedo(0), **kw)
    got = s.disort_batch(dt[None], ss[None], pm[None], bins, nstr=nstr)
    s.set_surfaces()
    s.close()
    assert got["status"][0] == 0
    sc_ = max(np.abs(ref[k]).max() for k in ("rfldir", "rfldn", "flup"))
    for k in ("rfldir", "rfldn", "flup", "uavg"):
        assert np.abs(got[k][0] - ref[k]).max() <= 1e-8 * sc_, k
