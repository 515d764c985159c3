"""GPU parity for the front-end options of frontend/extras.py: aerosols in the
producer kernel (K2) against the host front end, and whole runs (K2 + solve, and
host optical properties + GPU solve) against the CPU checker.

Tolerances: the K2 inputs are compared at 1e-9 relative (same formulas, different
pow/exp implementations); fluxes and radiances at the 1e-5 of BASELINE.json's north
star, through the printed records (5 digits) and directly on the arrays."""
import numpy as np
import pytest

import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
from sbchk_cases import compare_records
from solvers import make_solve_cuda, solve_oracle

pytestmark = pytest.mark.gpu

AEROSOL_RUNS = [
    # C4 slice: rural aerosol by visibility, stratus cloud, 65-level grid
    "&INPUT idatm=2, nstr=8, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=.3, wlsup=3.0, wlinc=.05, iout=1 /",
    # oceanic aerosol by optical depth, thermal range in wavenumber steps
    "&INPUT idatm=1, nstr=4, iaer=3, tbaer=0.25, rhaer=0.85, wlinf=4, wlsup=40, wlinc=20, sza=95, iout=1 /",
    # two stratospheric layers on top of an urban boundary layer, wavelengths beyond both table ends
    "&INPUT idatm=3, nstr=4, iaer=2, vis=10, jaer=2,4, zaer=18,25, taerst=0.1,0.02, wlinf=.21, wlsup=.8,"
    " wlinc=.01, sza=50, iout=1 /",
    # stratospheric aerosol only (iaer=0)
    "&INPUT idatm=4, nstr=4, jaer=3, zaer=22, taerst=0.2, wlinf=.4, wlsup=2.5, wlinc=.1, iout=1 /",
    # user model with an Angstrom continuation; Rayleigh-like phase function; no-scattering option
    "&INPUT idatm=2, nstr=4, iaer=5, wlbaer=.4,.55,1., qbaer=1.5,1.,.4, wbaer=.9,.92,.95, gbaer=.7,.68,.6,"
    " abaer=1.3, tbaer=.5, wlinf=.3, wlsup=2., wlinc=.05, iout=1 /",
    "&INPUT idatm=2, nstr=4, iaer=4, tbaer=.2, imoma=2, wlinf=.3, wlsup=1., wlinc=.05, iout=1 /",
    "&INPUT idatm=2, nstr=4, iaer=1, vis=15, nosct=3, wlinf=.3, wlsup=1., wlinc=.05, iout=1 /",
]


@pytest.mark.parametrize("nl", AEROSOL_RUNS)
def test_producer_kernel_aerosols_match_host_front_end(nl):
    from sbdart_b200.frontend.device import run_spectrum
    s = sb.Solver(0)
    run = Sbdart(nl)
    assert run.aerosols.active
    ref = run.batch(run.bins())
    rows, res, dev = run_spectrum(Sbdart(nl), s, want_inputs=True)
    assert len(rows) == len(ref["bins"])
    for k in ("dtauc", "ssalb", "pmom"):
        np.testing.assert_allclose(dev[k], ref[k], rtol=1e-9, atol=1e-13 * np.abs(ref[k]).max())
    # and the aerosol setting does not leak into the next run on the same handle
    clean = "&INPUT idatm=2, nstr=4, wlinf=.3, wlsup=1., wlinc=.05, iout=1 /"
    ref0 = Sbdart(clean)
    ref0 = ref0.batch(ref0.bins())
    _, _, dev0 = run_spectrum(Sbdart(clean), s, want_inputs=True)
    np.testing.assert_allclose(dev0["dtauc"], ref0["dtauc"], rtol=1e-9, atol=1e-13 * ref0["dtauc"].max())
    s.close()


@pytest.mark.parametrize("nl", [
    # BASELINE.json config 4 on a slice of its spectrum: NSTR=32, 65 layers, cloud + aerosol
    "&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=.4, wlsup=.7, wlinc=.02, iout=1 /",
    "&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=8, wlsup=12, wlinc=20, iout=10 /",
    # NSTR=16 with aerosol (fast kernel), sensor filter, date/place geometry, in-cloud humidity
    "&INPUT idatm=2, nstr=16, iaer=3, tbaer=.3, isat=4, wlinc=.005, iday=172, time=18, alat=34.4, alon=-119.8,"
    " tcloud=5, zcloud=1, rhcld=1., iout=10 /",
])
def test_whole_runs_match_the_cpu_checker(nl):
    s = sb.Solver(0)
    want = Sbdart(nl).run(solve_oracle)
    got_dev = Sbdart(nl).run_device(s)
    got_host = Sbdart(nl).run(make_solve_cuda(s))
    for got in (got_dev, got_host):
        nval, nexact, worst = compare_records(got, want, rel=1.2e-4)
        assert nval >= 9 and worst <= 1.2e-4
    # the arrays themselves at the north-star tolerance
    run = Sbdart(nl)
    b = run.batch(run.bins())
    r_gpu, r_cpu = make_solve_cuda(s)(b), solve_oracle(b)
    assert (r_gpu["status"] == r_cpu["status"]).all()
    for k in ("rfldir", "rfldn", "flup"):
        scale = np.abs(r_cpu["flup"]).max(axis=1, keepdims=True) + np.abs(r_cpu["rfldn"]).max(axis=1, keepdims=True) \
            + np.abs(r_cpu["rfldir"]).max(axis=1, keepdims=True)
        # floor: 2e-8 of the bin's flux scale (NSTR=32 with 65 layers leaves ~7e-9 of round-off on
        # fluxes that are 1e-7 of the bin maximum; the NSTR <= 16 cases sit below 1e-10)
        assert (np.abs(r_gpu[k] - r_cpu[k]) <= 1e-5 * np.abs(r_cpu[k]) + 2e-8 * scale).all(), k
    assert s.kernel_launches >= 3
    s.close()


def test_radiance_run_with_aerosol_and_filter():
    """iout=20 radiances (generic kernel, all azimuth modes) through K2 with aerosols."""
    nl = ("&INPUT idatm=2, nstr=8, iaer=1, vis=23, isat=-3, wlinf=.65, wlsup=.05, wlinc=.01, sza=30, iout=20,"
          " uzen=0,20,40,60,80, phi=0,90,180 /")
    s = sb.Solver(0)
    want = Sbdart(nl).run(solve_oracle)
    got = Sbdart(nl).run_device(s)
    nval, nexact, worst = compare_records(got, want, rel=1.2e-4)
    assert nval > 20 and worst <= 1.2e-4
    s.close()


def test_table_phase_function_takes_the_host_property_path():
    """imoma=4 (haze-L table moments) is not evaluated by K2: optical properties on the
    host, the solve still on the GPU."""
    nl = "&INPUT idatm=2, nstr=8, iaer=1, vis=23, imoma=4, wlinf=.5, wlsup=.6, wlinc=.02, iout=10 /"
    s = sb.Solver(0)
    n0 = s.kernel_launches
    got = Sbdart(nl).run_device(s)
    assert s.kernel_launches > n0
    nval, nexact, worst = compare_records(got, Sbdart(nl).run(solve_oracle), rel=1.2e-4)
    assert worst <= 1.2e-4
    s.close()


# ------------------------------------------------------------------ CORINT (INTCOR on the GPU)
CORINT_RUNS = [
    # aureole of a thin water cloud: view angles around the solar direction (IMS active), three azimuths
    "&INPUT idatm=2, nstr=8, tcloud=2, zcloud=2, wlinf=.5, wlsup=.7, wlinc=.05, sza=30, iout=21,"
    " uzen=100,140,145,151,155,160, phi=0,10,180, corint=t /",
    # rural aerosol, NSTR=16, upward and downward views, top of the atmosphere
    "&INPUT idatm=2, nstr=16, iaer=1, vis=10, wlinf=.45, wlsup=.65, wlinc=.1, sza=55, iout=20,"
    " uzen=0,30,60,80,100,127,170, phi=0,45,90,180, corint=t /",
    # deep cloud (layer truncation below the cloud in the visible), ice cloud on top
    "&INPUT idatm=1, nstr=8, tcloud=60,3, zcloud=1,9, nre=8,-30, wlinf=1.5, wlsup=1.7, wlinc=.1, sza=20, iout=23,"
    " uzen=10,150,165, phi=0,90, corint=t /",
    # SBDART's default stream count for radiance output (NSTR=20: register kernel, 10-lane layer groups)
    "&INPUT idatm=2, nstr=20, tcloud=1, zcloud=3, wlinf=.55, wlsup=.65, wlinc=.1, sza=40, iout=20,"
    " uzen=0,45,95,150, phi=0,60,180, corint=t /",
]


@pytest.mark.parametrize("nl", CORINT_RUNS)
def test_corint_matches_the_cpu_checker(nl):
    """Nakajima-Tanaka corrections (sbd_intcor.cu) against INTCOR of the checker, which the
    DISORT self test pins (UU = 47.865571, tests/test_oracle_golden.py).

    The view angles avoid uzen = 180 - sza: there SINSCA's prefactor 1/(1 + umu/umu0) is ~1e11
    (SBDART's pi = 3.1415926536 makes umu + umu0 ~ 5e-12, above DISORT's DITHER of 2e-14 that
    selects the closed form) and multiplies differences of exponentials that agree to 11
    digits, so any two implementations -- the reference included -- differ by ~2e-5 there."""
    s = sb.Solver(0)
    run = Sbdart(nl)
    b = run.batch(run.bins())
    assert b["corint"] and b["pmom"].shape[2] == 300
    b.pop("uu_levels")
    r_cpu = solve_oracle(b)
    r_gpu = make_solve_cuda(s)(b)                       # every level
    assert (r_gpu["status"] == 0).all() and (r_cpu["status"] == 0).all()
    scale = np.abs(r_cpu["uu"]).max(axis=(1, 2, 3), keepdims=True)
    assert (np.abs(r_gpu["uu"] - r_cpu["uu"]) <= 1e-5 * np.abs(r_cpu["uu"]) + 1e-9 * scale).all()
    # the correction is not a no-op: it changes the radiances by more than a per cent somewhere
    plain = dict(b, corint=False)
    r_plain = make_solve_cuda(s)(plain)
    assert np.abs(r_gpu["uu"] - r_plain["uu"]).max() > 1e-2 * scale.max()
    np.testing.assert_array_equal(r_gpu["flup"], r_plain["flup"])       # fluxes untouched
    # records through the selected-level path
    got = Sbdart(nl).run_device(s)
    nval, nexact, worst = compare_records(got, Sbdart(nl).run(solve_oracle), rel=1.2e-4)
    assert worst <= 1.2e-4
    s.close()


def test_corint_through_the_fortran_entry():
    """disort_ with CORINT=.TRUE.: the self-test atmosphere of disort.f:6393-6430 cut into two
    layers so that its output level (tau = 0.5) is a layer boundary."""
    from oracle import oracle
    pm = np.array([1.0, 0.8042, 0.646094, 0.481851, 0.359056])
    kw = dict(fbeam=3.14159265, umu0=0.866, phi0=0.0, fisot=1.0, albedo=0.7, btemp=300.0, ttemp=100.0,
              temis=0.8, wvnmlo=0.0, wvnmhi=50000.0)
    want = oracle.disort([0.5, 0.5], [0.9, 0.9], np.stack([pm, pm]), nstr=4, temper=[210., 205., 200.],
                         umu=[-0.5, 0.5], phi=[0.0, 90.0], plank=True, onlyfl=False, corint=True, **kw)
    got = sb.disort(2, [0.5, 0.5], [0.9, 0.9], 4, np.stack([pm, pm]).T, [210., 205., 200.], kw["wvnmlo"],
                    kw["wvnmhi"], False, 0, None, 4, True, 2, [-0.5, 0.5], 2, [0.0, 90.0], 0, kw["fbeam"],
                    kw["umu0"], kw["phi0"], kw["fisot"], True, kw["albedo"], kw["btemp"], kw["ttemp"],
                    kw["temis"], True, False, accur=0.0, corint=True)
    assert got["status"] == 0 and want["status"] == 0
    uu = np.transpose(got["uu"][:2, :3, :2], (2, 1, 0))          # UU(iu, lu, j) -> [j][lu][iu]
    np.testing.assert_allclose(uu, want["uu"], rtol=1e-6, atol=1e-9 * np.abs(want["uu"]).max())
    np.testing.assert_allclose(got["flup"][:3], want["flup"], rtol=1e-8)


@pytest.mark.parametrize("nl", [
    # nstr left to its default: 20 streams for radiance output (drt.f:241-247)
    "&INPUT idatm=2, tcloud=3, zcloud=2, wlinf=.5, wlsup=.8, wlinc=.05, sza=35, iout=5, uzen=0,30,60,85,100,150, phi=0,90,180 /",
    "&INPUT idatm=4, wlinf=9, wlsup=12, wlinc=.5, sza=95, iout=6, uzen=5,50,95,140,175, phi=0 /",
    "&INPUT idatm=2, nstr=24, iaer=1, vis=15, wlinf=.45, wlsup=.55, wlinc=.05, sza=60, iout=21, uzen=20,70,110,160, phi=0,45 /",
])
def test_whole_radiance_runs_at_the_default_stream_count(nl):
    """iout = 5 / 6 / 21 radiance records through the device path (K2 + radiance register kernel at
    NSTR 20 / 24) against the front end driven by the CPU checker."""
    s = sb.Solver(0)
    run = Sbdart(nl)
    assert run.p["nstr"] in (20, 24)
    want = Sbdart(nl).run(solve_oracle)
    got = Sbdart(nl).run_device(s)
    nval, nexact, worst = compare_records(got, want, rel=1.2e-4)
    assert nval > 20 and worst <= 1.2e-4
    s.close()
