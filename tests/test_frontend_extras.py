"""Front-end options no stored output of the reference exercises (aerosols, zgrid,
in-cloud humidity, zensun, sensor filters; sbdart_b200/frontend/extras.py).

There is no golden for them, so the checks are the properties the
reference's own code guarantees: normalisation identities of denprfl (tauaero.f:1425-1443),
the grid end points of zgrid (atms.f:566-585), conservation of the water column in
saturate (atms.f:196-213), Beer's law through the full solve, and the byte format of
the records.  The solve is done by the CPU checker (tests/solvers.py).
"""
import math

import numpy as np
import pytest

from sbdart_b200.frontend import Sbdart, atms, extras, f32, setfilt
from solvers import solve_oracle


def _tau_aer(run, wl, nmom=6):
    pm = np.zeros((run.nz, nmom + 1))
    dtaua, waer = run.aerosols(wl, nmom, pm)
    return dtaua, waer, pm


# ------------------------------------------------------------------ zgrid
def test_zgrid_end_points_and_spacing():
    z, p, t, wh, wo = atms(2)
    zz, pp, tt, hh, oo = extras.zgrid(z, p, t, wh, wo, 1.0, 30.0, 65)
    assert len(zz) == 65 and zz[0] == 0.0 and abs(zz[-1] - z[-1]) < 1e-9
    assert (np.diff(zz) > 0).all()
    assert abs((zz[-1] - zz[-2]) - 30.0) < 1e-6          # zgrid2 = thickness of the top layer
    assert abs((zz[1] - zz[0]) - 1.0) < 1e-3             # zgrid1 = resolution at the bottom
    # the first points coincide with original levels 0, 1, 2 km: the fields are reproduced there
    for k in range(3):
        assert abs(pp[k] - p[k]) <= 1e-12 * p[k] and abs(tt[k] - t[k]) <= 1e-12 * t[k]
    assert (np.diff(pp) < 0).all() and pp.min() > 0 and hh.min() >= 0 and oo.min() >= 0


def test_zgrid_run_has_65_layers_and_close_fluxes():
    a = Sbdart("&INPUT idatm=2, wlinf=0.55, wlsup=0.55, iout=10 /")
    b = Sbdart("&INPUT idatm=2, wlinf=0.55, wlsup=0.55, iout=10, ngrid=65 /")
    assert a.nz == 33 and b.nz == 65 and len(b.temper) == 66
    fa = [float(x) for x in a.run(solve_oracle).split()[3:]]
    fb = [float(x) for x in b.run(solve_oracle).split()[3:]]
    # same atmosphere on a finer grid: top-of-atmosphere and surface fluxes agree to ~0.1 %
    for x, y in zip(fa, fb):
        if abs(x) > 1e-6:
            assert abs(x - y) <= 3e-3 * abs(x)


# ------------------------------------------------------------------ aerosols
@pytest.mark.parametrize("iaer", [1, 2, 3, 4])
def test_tbaer_is_the_optical_depth_at_055(iaer):
    run = Sbdart(f"&INPUT idatm=2, iaer={iaer}, tbaer=0.37, wlinf=0.55, wlsup=0.55, iout=10 /")
    dtaua, waer, pm = _tau_aer(run, f32(0.55))
    assert abs(dtaua.sum() - 0.37) < 1e-12                # tauaero.f:1431-1433
    assert (waer > 0).all() and (waer <= 1).all()
    # Henyey-Greenstein moments weighted with the scattering depth (tauaero.f:1296-1301)
    g = pm[:, 1] / (dtaua * waer)
    assert np.allclose(pm[:, 2], g ** 2 * dtaua * waer, rtol=1e-12)
    assert 0.5 < g[0] < 0.85


def test_visibility_normalisation():
    run = Sbdart("&INPUT idatm=2, iaer=1, vis=23, wlinf=0.55, wlsup=0.55, iout=10 /")
    aer = run.aerosols
    dtaua, _, _ = _tau_aer(run, f32(0.55))
    # sigma = 3.912 / (ext55 vis n(0)), tau = sigma sum n dz (tauaero.f:1434-1440), n from the
    # 23 km profile whatever vis is (inverted test in aerzstd, tauaero.f:129-134)
    nd = aer._aervint()
    assert aer.aeroden(0.0) == pytest.approx(2828.0)
    assert dtaua.sum() == pytest.approx(f32(3.912) / 23. * nd.sum() / 2828.0, rel=1e-12)
    assert nd[0] == pytest.approx(aer.aeroden(100.) * 5.)
    # a thicker haze scales every layer by the same factor
    run5 = Sbdart("&INPUT idatm=2, iaer=1, vis=5, wlinf=0.55, wlsup=0.55, iout=10 /")
    d5, _, _ = _tau_aer(run5, f32(0.55))
    assert np.allclose(d5, dtaua * 23. / 5., rtol=1e-12)


def test_aerosol_spectral_dependence_and_humidity():
    dry = Sbdart("&INPUT idatm=2, iaer=1, vis=23, rhaer=0.0, wlinf=0.55, wlsup=0.55, iout=10 /").aerosols
    wet = Sbdart("&INPUT idatm=2, iaer=1, vis=23, rhaer=0.95, wlinf=0.55, wlsup=0.55, iout=10 /").aerosols
    for a in (dry, wet):
        e = [a.aerbwi(w)[0] for w in (0.3, 0.55, 1.0, 2.0, 10.0)]
        assert e[0] > e[1] > e[2] > e[3] > e[4] > 0          # rural aerosol: extinction falls with wavelength
        assert a.aerbwi(f32(0.55))[0] == pytest.approx(1.0, abs=2e-4)   # tables are normalised at 0.55 um
    assert wet.aerbwi(0.55)[1] > dry.aerbwi(0.55)[1]         # wet particles absorb less
    # the table end points continue as a power law with exponent abaer = 0 for the standard models
    assert dry.aerbwi(0.1)[0] == pytest.approx(dry.aerext[0])
    assert dry.aerbwi(500.)[0] == pytest.approx(dry.aerext[-1])


def test_stratospheric_layer():
    run = Sbdart("&INPUT idatm=2, jaer=1, zaer=20, taerst=0.05, wlinf=0.55, wlsup=0.55, iout=10 /")
    dtaua, waer, pm = _tau_aer(run, f32(0.55))
    nl = run.aerosols.laer[0]
    assert run.z[run.nz - nl] <= 20.0 + 1e-3 < run.z[run.nz - nl + 1]
    assert np.count_nonzero(dtaua) == 1
    assert dtaua[nl - 1] == pytest.approx(0.05, rel=1e-6)   # background model: extinction 1 at 0.55 um
    assert 0.99 < waer[nl - 1] <= 1.0
    # on top of a boundary-layer aerosol the single-scattering albedo is the depth-weighted mean
    both = Sbdart("&INPUT idatm=2, iaer=2, tbaer=0.3, jaer=3, zaer=20, taerst=0.05, wlinf=0.55, wlsup=0.55, iout=10 /")
    d2, w2, _ = _tau_aer(both, f32(0.55))
    bl = Sbdart("&INPUT idatm=2, iaer=2, tbaer=0.3, wlinf=0.55, wlsup=0.55, iout=10 /")
    d1, w1, _ = _tau_aer(bl, f32(0.55))
    qa, wa, _ = both.aerosols.aestrat(3, f32(0.55))
    j = both.aerosols.laer[0] - 1
    assert d2[j] == pytest.approx(d1[j] + 0.05 * qa, rel=1e-12)
    assert w2[j] == pytest.approx((w1[j] * d1[j] + wa * 0.05 * qa) / d2[j], rel=1e-12)
    assert np.allclose(np.delete(d2, j), np.delete(d1, j))


def test_user_aerosol_model():
    run = Sbdart("&INPUT idatm=2, iaer=5, wlbaer=0.4,0.55,1.0, qbaer=1.5,1.0,0.4, wbaer=0.9,0.92,0.95,"
                 " gbaer=0.7,0.68,0.6, wlinf=0.55, wlsup=0.55, iout=10 /")
    dtaua, waer, _ = _tau_aer(run, f32(0.55))
    assert dtaua.sum() == pytest.approx(1.0, rel=1e-6)      # tbaer defaults to q(0.55) (tauaero.f:1413)
    assert waer[0] == pytest.approx(0.92, rel=1e-6)
    e, w, g = run.aerosols.aerbwi(0.7416198487)             # geometric mean of 0.55 and 1.0
    assert e == pytest.approx(math.sqrt(1.0 * 0.4), rel=1e-6) and g == pytest.approx(0.64, rel=1e-6)
    with pytest.raises(ValueError):
        Sbdart("&INPUT iaer=5, wlbaer=0.4,0.55, qbaer=1.5, wbaer=0.9,0.92, gbaer=.7,.7 /")
    with pytest.raises(ValueError):
        Sbdart("&INPUT iaer=1 /")                           # must specify either tbaer or vis


def test_direct_beam_obeys_beers_law_through_the_solve():
    """Monochromatic run with and without aerosol: the surface direct beam differs by
    exp(-tau_aer / mu0) exactly; the diffuse field grows."""
    base = "&INPUT idatm=2, wlinf=0.7, wlsup=0.7, sza=60, kdist=0, iout=10"
    a = Sbdart(base + " /")
    b = Sbdart(base + ", iaer=1, tbaer=0.4 /")
    fa = [float(x) for x in a.run(solve_oracle).split()]
    fb = [float(x) for x in b.run(solve_oracle).split()]
    tau = _tau_aer(b, f32(0.7))[0].sum()
    assert fb[8] / fa[8] == pytest.approx(math.exp(-tau / 0.5), rel=2e-4)   # botdir, 5 printed digits
    assert fb[4] > fa[4]                                                     # more light scattered back to space
    assert (fb[6] - fb[8]) > (fa[6] - fa[8])                                 # more diffuse light at the surface


# ------------------------------------------------------------------ humidity in clouds
def test_saturate_conserves_the_water_column():
    def column(z, wh):
        tot = 0.0
        for i in range(len(z) - 1):
            d1, d2, dz = wh[i], wh[i + 1], z[i + 1] - z[i]
            tot += .5 * dz * (d1 + d2) if abs(d1 - d2) <= 1e-3 * d1 else dz * (d1 - d2) / math.log(d1 / d2)
        return tot
    clear = Sbdart("&INPUT idatm=2, tcloud=10, zcloud=2 /")
    cloudy = Sbdart("&INPUT idatm=2, tcloud=10, zcloud=2, rhcld=1.0 /")
    # the layer-mean column the routine starts from (atms.f:85-96) equals the log-mean column it ends with
    z = clear.z
    edges = np.concatenate([[z[0]], .5 * (z[1:] + z[:-1]), [z[-1]]])
    assert column(z, cloudy.wh) == pytest.approx(float((np.diff(edges) * clear.wh).sum()), rel=1e-12)
    # saturated at the cloud level relative to the levels around it
    rh = [extras.relhum(cloudy.t[i], cloudy.wh[i]) for i in range(6)]
    assert rh[2] > rh[0] and rh[2] > rh[4] and rh[2] == pytest.approx(rh[3], rel=1e-9)
    forced = Sbdart("&INPUT idatm=2, tcloud=10, zcloud=2, rhcld=0.9, krhclr=1 /")
    assert extras.relhum(forced.t[2], forced.wh[2]) == pytest.approx(0.9, rel=1e-12)
    assert forced.wh[10] == clear.wh[10]


# ------------------------------------------------------------------ solar geometry
def test_zensun():
    # equinox (day 80), local noon on the Greenwich meridian at the equator: sun near the zenith
    zen, azm, solfac = extras.zensun(80, 12.0, 0.0, 0.0)
    assert zen < 2.5
    # summer solstice at 23.44 N: overhead sun; winter: 46.9 degrees from the zenith
    assert extras.zensun(172, 12.0, 23.44, 0.0)[0] < 1.0
    assert extras.zensun(355, 12.0, 23.44, 0.0)[0] == pytest.approx(46.88, abs=0.3)
    # earth-sun distance: perihelion early January, aphelion early July
    assert extras.zensun(2, 12., 0., 0.)[2] == pytest.approx(1.0 / (1 - 0.01671) ** 2, rel=1e-6)
    assert extras.zensun(185, 12., 0., 0.)[2] == pytest.approx(1.0 / (1 + 0.01671) ** 2, rel=1e-4)
    run = Sbdart("&INPUT iday=172, time=12, alat=23.44, alon=0, wlinf=.55, wlsup=.55 /")
    assert run.sza < 1.0 and run.p["solfac"] < 1.0 and 0.0 <= run.phi0 < 360.0


# ------------------------------------------------------------------ sensor filters
def test_filters():
    wl1, wl2, nwl, wlinc, filt = setfilt(1, 0.0, 0.0, 0.005, want_filter=True)      # METEOSAT
    assert (wl1, wl2) == (f32(.355), f32(1.11)) and nwl == 152
    assert filt(wl1) == pytest.approx(0.005) and filt(0.72) > 0.95 and filt(5.0) == filt(wl2)
    assert setfilt(1, 0.0, 0.0, 0.005)[:3] == (wl1, wl2, nwl)
    wl1, wl2, nwl, wlinc, tri = setfilt(-3, 0.6, 0.1, 0.01, want_filter=True)
    assert (wl1, wl2, nwl) == (0.5, 0.7, 21)
    assert tri(0.6) == 1.0 and tri(0.55) == pytest.approx(0.5) and tri(0.5) == 0.0
    wl1, wl2, nwl, wlinc, gau = setfilt(-4, 0.6, 0.05, 0.01, want_filter=True)
    assert gau(0.6) == pytest.approx(1.0, abs=1e-4) and gau(wl1) == pytest.approx(math.exp(-4 * math.pi), rel=1e-6)
    with pytest.raises(ValueError):
        setfilt(-3, 0.6, 0.0, 0.01, want_filter=True)
    flat = setfilt(0, 0.3, 0.4, 0.01, want_filter=True)[4]
    assert flat(0.35) == 1.0


def test_filtered_run_is_the_weighted_mean():
    """iout=10 with a triangular filter = sum of the unfiltered spectral fluxes times filter x dwl."""
    run = Sbdart("&INPUT idatm=2, isat=-3, wlinf=0.6, wlsup=0.05, wlinc=0.01, kdist=0, iout=10 /")
    out = [float(x) for x in run.run(solve_oracle).split()]
    rows, res = run.last["rows"], run.last["result"]
    assert len(rows) == 11 and rows[0]["ff"] == 0.0 and rows[5]["ff"] == 1.0
    top = sum((res["rfldn"][i][run.ntop - 1] + res["rfldir"][i][run.ntop - 1]) * r["wt"] * r["ff"]
              for i, r in enumerate(rows))
    assert out[3] == pytest.approx(top, rel=1e-4)
    assert out[2] == pytest.approx(sum(r["dwl"] * r["ff"] for r in rows), abs=5e-5)


def test_modtran_solar_table():
    from sbdart_b200.frontend import Sun
    s3, s2 = Sun(3), Sun(2)
    assert (np.diff(s3.wl) > 0).all() and len(s3.wl) == 2494
    wl = np.linspace(0.3, 3.0, 200)
    # two solar spectra: the same sun (W/m2/um) to a few per cent when integrated
    i3 = np.trapezoid([s3(w) for w in wl], wl)
    i2 = np.trapezoid([s2(w) for w in wl], wl)
    assert i3 == pytest.approx(i2, rel=0.03) and 1200 < i3 < 1400


# ------------------------------------------------------------------ the C4 namelist
def test_c4_namelist_builds_bins():
    """BASELINE.json config 4: NSTR=32, stratus cloud + rural aerosol, 65 layers (a slice of the spectrum)."""
    run = Sbdart("&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=.5, wlsup=.6,"
                 " wlinc=.02, iout=10 /")
    rows = run.bins()
    b = run.batch(rows)
    assert b["dtauc"].shape[1] == 65 and b["pmom"].shape[2] == 35 and b["nstr"] == 32
    assert (b["ssalb"] >= 0).all() and (b["ssalb"] <= 1).all() and (b["dtauc"] >= 0).all()
    assert np.allclose(b["pmom"][:, :, 0], 1.0)
    lc = run.clouds.lcld[0] - 1
    assert b["dtauc"][0, lc] > 9.0                            # the cloud layer
    assert b["dtauc"][0].sum() - b["dtauc"][0, lc] > 0.3      # aerosol + Rayleigh + gas
    res = solve_oracle(b)
    assert (res["status"] == 0).all()
    txt = run.records(rows, res)
    assert len(txt.split()) == 9


def test_iout22_holds_the_iout20_and_iout21_radiances():
    """iout=22 prints fluxes and radiances at every level (drt.f:1153-1163); its first and last
    level blocks are the iout=20 (top) and iout=21 (bottom) radiance tables."""
    base = "&INPUT idatm=2, nstr=8, wlinf=.55, wlsup=.56, wlinc=.01, sza=30, uzen=10,100,170, phi=0,90, iout={} /"
    t22, t20, t21 = (Sbdart(base.format(i)).run(solve_oracle).split() for i in (22, 20, 21))
    assert t22[:3] == ["2", "3", "33"]
    nz = 33
    uurl = t22[4 + 2 + 3 + 4 * nz:]
    assert len(uurl) == 2 * 3 * nz
    assert uurl[:6] == t20[-6:] and uurl[-6:] == t21[-6:]
    z = [float(x) for x in t22[9:9 + nz]]
    assert z[0] == 100.0 and z[-1] == 0.0
    fxdn = [float(x) for x in t22[9 + nz:9 + 2 * nz]]
    assert fxdn[0] == pytest.approx(float(t20[3]), rel=1e-4) and fxdn[-1] == pytest.approx(float(t20[6]), rel=1e-4)


def test_corint_restores_the_aureole():
    """CORINT=.TRUE. (299 moments, INTCOR): the forward peak the delta-M truncation removes comes
    back near the solar direction; far from it and in the fluxes nothing changes much."""
    base = ("&INPUT idatm=2, nstr=8, tcloud=2, zcloud=2, wlinf=.55, wlsup=.55, sza=30, iout=21,"
            " uzen=100,151,175, phi=0,180, corint={} /")
    off, on = (Sbdart(base.format(c)).run(solve_oracle).split() for c in ("f", "t"))
    assert off[:9] == on[:9]                                   # the flux record
    r_off, r_on = [float(x) for x in off[-6:]], [float(x) for x in on[-6:]]
    assert r_on[2] > 2.5 * r_off[2]                            # uzen=151, phi=0: one degree from the sun
    assert abs(r_on[3] / r_off[3] - 1) < 0.15                  # same zenith angle, opposite azimuth


# ------------------------------------------------------------------ user data files
def test_user_files_reproduce_the_built_in_tables(tmp_path, monkeypatch):
    """atms.dat / albedo.dat / solar.dat / filter.dat written from the built-in tables give the
    same run as the built-in options (useratm atms.f:468, rdspec spectra.f:4382)."""
    from sbdart_b200.frontend import Albedo, Sun
    monkeypatch.chdir(tmp_path)
    z, p, t, wh, wo = atms(2)
    rows = [f"{len(z)}"] + [f"{z[i]:.6f} {p[i]:.8e} {t[i]:.6f} {wh[i]:.8e} {wo[i]:.8e} trailing text ignored"
                            for i in range(len(z) - 1, -1, -1)]
    (tmp_path / "atms.dat").write_text("\n".join(rows) + "\n")
    alb = Albedo(5, 0.0, [1, 0, 0, 0, 0])
    (tmp_path / "albedo.dat").write_text("".join(f"{w:.9f}, {a:.9e}\n" for w, a in zip(alb.wl, alb.alb)))
    sun = Sun(1)
    (tmp_path / "solar.dat").write_text("".join(f"{w:.9f} {v:.9e}\n" for w, v in zip(sun.wl, sun.s)))
    (tmp_path / "filter.dat").write_text("0.50 0.0\n0.60 1.0\n\n0.70 0.0\n")
    want = Sbdart("&INPUT idatm=2, isalb=5, nf=1, isat=-3, wlinf=.6, wlsup=.1, wlinc=.02, sza=40, kdist=0, iout=10 /")
    got = Sbdart("&INPUT idatm=0, isalb=-1, nf=-1, isat=-1, wlinc=.02, sza=40, kdist=0, iout=10 /")
    assert got.nz == want.nz and np.allclose(got.z, want.z) and np.allclose(got.pr, want.pr, rtol=1e-8)
    assert (got.wl1, got.wl2, got.nwl) == (0.5, 0.7, want.nwl)
    a = [float(x) for x in want.run(solve_oracle).split()]
    b = [float(x) for x in got.run(solve_oracle).split()]
    assert b == pytest.approx(a, rel=2e-4)
    # amix blends a user atmosphere into a standard one on the same grid (atms.f:451-463)
    mixed = Sbdart("&INPUT idatm=4, amix=0.5 /")
    z4, p4, t4, _, _ = atms(4)
    assert np.allclose(mixed.t, 0.5 * (t + t4))
    (tmp_path / "atms.dat").write_text("2\n10. 200. 220. 1. 1.\n0. 1000. 290. 5. 1.\n")
    with pytest.raises(ValueError):
        Sbdart("&INPUT idatm=4, amix=0.5 /")


def test_chkin_rejects_what_the_reference_rejects():
    """Range checks of drt.f:568-735 with the reference's wording."""
    for nl, word in (("&INPUT iout=3 /", "iout"), ("&INPUT wlinf=0.1, wlsup=0.5 /", "wlinf"),
                     ("&INPUT wlinf=0.6, wlsup=0.5 /", "wlsup"), ("&INPUT isat=40 /", "isat"),
                     ("&INPUT nf=7 /", "nf"), ("&INPUT tcloud=5, lwp=10, zcloud=1 /", "TCLOUD or LWP"),
                     ("&INPUT zpres=2, pbar=900 /", "zpres or pbar"), ("&INPUT nre=1, tcloud=5, zcloud=1 /", "nre"),
                     ("&INPUT isalb=11 /", "isalb"), ("&INPUT jaer=5, zaer=20, taerst=.1 /", "jaer")):
        with pytest.raises(ValueError, match=word):
            Sbdart(nl)
    ok = Sbdart("&INPUT vis=23 /")                       # a warning, not an error (errmsg 16)
    assert ok.warnings == [(16, "CHKIN--IAER=0, though VIS or TBAER set")]


def test_print_and_stop_modes():
    """idatm < 0 / ngrid < 0 print the (regridded) atmosphere (prnatm, drt.f:803-809), iday < 0 the
    solar geometry (drt.f:285-299); the reference then executes STOP."""
    from sbdart_b200.frontend import SbdartStop
    with pytest.raises(SbdartStop) as e:
        Sbdart("&INPUT idatm=-2 /")
    lines = e.value.text.splitlines()
    assert lines[0] == "          33" and len(lines) == 34
    assert lines[1] == "      0.000  1.013E+03  2.940E+02  1.400E+01  6.000E-05"
    with pytest.raises(SbdartStop) as e:
        Sbdart("&INPUT idatm=2, ngrid=-20 /")
    assert e.value.text.splitlines()[0].strip() == "20" and len(e.value.text.splitlines()) == 21
    with pytest.raises(SbdartStop) as e:
        Sbdart("&INPUT iday=-172, time=18, alat=34.4, alon=-119.8 /")
    assert e.value.text.splitlines()[1].split()[:4] == ["172", "18.000", "34.400", "-119.800"]


def test_brdf_surface_runs_and_lambertian_limit():
    """isalb = 7, 8, 9 (drt.f:469-477, LAMBER = .FALSE.).  A Ross-Li surface with only the
    isotropic kernel is a Lambertian surface of that albedo: the records must be identical to the
    isalb = 0 run (SURFAC's quadrature of a constant, disort.f:3765-3829)."""
    from solvers import solve_oracle
    base = "&INPUT\n idatm=4, isat=0, wlinf=.5, wlsup=.6, wlinc=.05, iout=10, sza=40, {}\n /"
    lam = Sbdart(base.format("isalb=0, albcon=0.2")).run(solve_oracle)
    iso = Sbdart(base.format("isalb=9, sc=0.2,0,0,1,1")).run(solve_oracle)
    assert lam == iso
    hap = Sbdart(base.format("isalb=8, sc=0.6,0.3,0.4,0.1")).run(solve_oracle)
    sea = Sbdart(base.format("isalb=7, sc=0.5,6.0,34.3")).run(solve_oracle)
    f = lambda rec: np.array(rec.split()[3:], float)      # noqa: E731
    # same incident flux, different upward fluxes; the ocean is the darkest of the three
    assert f(hap)[0] == f(lam)[0] == f(sea)[0]
    assert f(sea)[1] < f(lam)[1] and f(hap)[1] != f(lam)[1]
    # radiances: one surface per wavelength for the ocean model, every azimuth mode
    rad = Sbdart("&INPUT\n idatm=4, isat=0, wlinf=.5, wlsup=.6, wlinc=.05, iout=20, isalb=7, sc=0.5,6.0,34.3,"
                 " sza=40, nstr=8, uzen=0,30,60, phi=0,90,180\n /").run(solve_oracle)
    rows = [np.array(l.split(), float) for l in rad.splitlines()[4:7]]
    assert rows[0][0] == rows[0][1] == rows[0][2]          # nadir view: no azimuth dependence
    assert rows[1][0] != rows[1][2]                        # sun glint side vs. the far side


def test_spowder_granular_surface_layer():
    """spowder (drt.f:337-349, taugas.f:7592-7595): a sub-surface layer between -1 and 0 km, no gas
    and no Rayleigh scattering in it, temper(nz+1) left unset as in the reference.  The manual's
    example (rtdoc.txt:1418-1424: a semi-infinite layer of 100 um ice grains under a thin water
    cloud) must behave like the reference's own snow albedo table (isalb = 1) to a few percent."""
    from solvers import solve_oracle
    base = "&INPUT\n sza=30, idatm=4, wlinf=.4, wlsup=.8, wlinc=.1, iout=1,\n {}\n /"
    pw = Sbdart(base.format("spowder=t, tcloud=10000,10, zcloud=-1,2, nre=-100,10"))
    assert pw.nz == 34 and pw.z[0] == -1.0 and pw.temper[-1] == 0.0 and pw.clouds.lcld[0] == 34
    rows = pw.bins()
    assert all(5000.0 < r["dtau"][-1] < 20000.0 for r in rows)   # only the grains in the bottom layer
    got = np.array([l.split() for l in pw.run(solve_oracle).splitlines()[3:]], float)
    snow = Sbdart(base.format("tcloud=0,10, zcloud=0,2, nre=8,10, isalb=1")).run(solve_oracle)
    ref = np.array([l.split() for l in snow.splitlines()[3:]], float)
    assert np.allclose(got[:, 2], ref[:, 2])                     # same solar input
    assert np.allclose(got[:, 6] / got[:, 5], ref[:, 6] / ref[:, 5], rtol=0.03)      # surface albedo
