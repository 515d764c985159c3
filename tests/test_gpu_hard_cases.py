"""Hard cases of the flux solve on the GPU (adding kernel, NSTR 4..32) against the CPU checker:
conservative scattering over a perfectly reflecting surface, optically huge and empty layers,
single-layer atmospheres, grazing sun, strongly forward-peaked phase functions, hot thin layers."""
import numpy as np
import pytest

import sbdart_b200 as sb
from oracle import oracle

pytestmark = pytest.mark.gpu


def _hg(g, nmom):
    return np.asarray(g)[:, None] ** np.arange(nmom + 1)[None, :]


def _cases(nstr):
    nm = nstr + 2
    c = []
    # (name, dtauc, ssalb, g, kwargs)
    c.append(("thick conservative cloud over albedo 1", [0.1, 300.0, 0.2], [1.0, 1.0, 1.0], [0.0, 0.85, 0.3],
              dict(fbeam=1.0, umu0=0.5, albedo=1.0)))
    c.append(("thin conservative over albedo 1", [1e-4, 1e-3, 1e-5], [1.0, 1.0, 1.0], [0.1, 0.2, 0.3],
              dict(fbeam=1.0, umu0=0.9, albedo=1.0)))
    c.append(("all layers empty", [0.0, 0.0, 0.0, 0.0], [0.5, 0.9, 1.0, 0.0], [0.5, 0.5, 0.5, 0.5],
              dict(fbeam=1.0, umu0=0.7, albedo=0.3)))
    c.append(("single layer", [0.7], [0.95], [0.75], dict(fbeam=1.0, umu0=0.3, albedo=0.1)))
    c.append(("huge absorbing depth (truncation)", [5.0, 2000.0, 1.0, 3.0], [0.5, 0.4, 0.9, 0.2], [0.7, 0.8, 0.1, 0.5],
              dict(fbeam=1.0, umu0=0.6, albedo=0.5)))
    c.append(("huge depth with thermal emission (no truncation)", [5.0, 2000.0, 1.0], [0.5, 0.4, 0.9], [0.7, 0.8, 0.1],
              dict(fbeam=0.0, umu0=1.0, albedo=0.2, plank=True)))
    c.append(("grazing sun", [0.05, 0.5, 2.0], [0.99, 0.9, 0.8], [0.6, 0.7, 0.8], dict(fbeam=1.0, umu0=0.011, albedo=0.4)))
    c.append(("forward peak", [0.3, 4.0, 0.3], [0.999, 0.99999, 0.9], [0.93, 0.97, 0.9], dict(fbeam=1.0, umu0=0.8, albedo=0.0)))
    c.append(("isotropic illumination only", [0.5, 0.5], [0.9, 0.3], [0.2, 0.6], dict(fbeam=0.0, umu0=1.0, albedo=0.7, fisot=1.0)))
    c.append(("pure absorption", [0.3, 1.0, 2.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], dict(fbeam=1.0, umu0=0.5, albedo=0.3, plank=True)))
    c.append(("mixed thin / thick / conservative", [1e-7, 30.0, 1e-3, 8.0, 1e-9, 0.4], [1.0, 0.999999, 0.2, 1.0, 0.7, 0.95],
              [0.0, 0.86, 0.4, 0.8, 0.1, 0.7], dict(fbeam=1.0, umu0=0.45, albedo=0.9, plank=True)))
    return [(n, np.array(d, float), np.array(s, float), _hg(g, nm), kw) for n, d, s, g, kw in c]


@pytest.mark.parametrize("nstr", [4, 8, 16, 20, 32])
def test_hard_cases_match_the_checker(nstr):
    s = sb.Solver(0)
    worst = 0.0
    for name, dt, ss, pm, kw in _cases(nstr):
        L = len(dt)
        kw = dict(kw)
        plank = kw.pop("plank", False)
        temper = np.linspace(210.0, 300.0, L + 1)
        okw = dict(kw)
        if plank:
            okw.update(plank=True, temper=temper, wvnmlo=600.0, wvnmhi=700.0, btemp=310.0, ttemp=100.0, temis=1.0)
        ref = oracle.disort(dt, ss, pm, nstr=nstr, **okw)
        if "mixed thin" in name:
            # hot, optically negligible layers: DISORT's thermal particular solution cancels
            # 1e9-sized terms there (tests/test_adding_math_cpu.py), so the reference itself is
            # only good to ~1e-3 in this case; the numpy statement of the adding form is the checker
            import ctypes as C
            import adding_model as am
            mu, w = sb.quadrature(nstr // 2)
            plk = lambda t: oracle.lib().sbdo_plkavg(600.0, 700.0, float(t), C.byref(C.c_int(0)))   # noqa: E731
            mod = am.fluxes(dt, ss, pm, nstr, mu, w, fbeam=kw["fbeam"], umu0=kw["umu0"], albedo=kw["albedo"],
                            pk=np.array([plk(t) for t in temper]), tplank=plk(100.0), bplank=plk(310.0))
            assert max(np.abs(mod[k] - ref[k]).max() for k in ("flup", "rfldn")) < 2e-3 * np.abs(ref["flup"]).max()
            ref = dict(mod, status=0)
        bins = sb.make_bins(1, plank=int(plank), **{k: v for k, v in okw.items() if k not in ("plank", "temper")})
        got = s.disort_batch(dt[None], ss[None], pm[None], bins, nstr=nstr, temper=temper[None])
        assert got["status"][0] == ref["status"], (name, got["status"][0], ref["status"])
        if ref["status"] != 0:
            continue
        scale = max(np.abs(ref[k]).max() for k in ("rfldir", "rfldn", "flup"))
        for k in ("rfldir", "rfldn", "flup", "uavg"):
            err = np.abs(got[k][0] - ref[k]).max() / scale
            worst = max(worst, err)
            # conservative / albedo-1 problems are ill conditioned in both formulations
            tol = 1e-6 if "albedo 1" in name else 1e-7
            assert err <= tol, (name, k, err)
    s.close()
    print("worst", worst)


@pytest.mark.parametrize("nstr", [8, 16, 20])
def test_hard_cases_radiances(nstr):
    """The same atmospheres with user angles (radiance register kernel: adding sweeps, layer solutions
    recovered from the interface intensities -- the recovery multiplies by k, which is ~1e-7 for the
    dithered conservative layers)."""
    umu = np.array([-1.0, -0.5, -0.1, 0.1, 0.6, 1.0])
    phi = np.array([0.0, 90.0])
    s = sb.Solver(0)
    worst = 0.0
    for name, dt, ss, pm, kw in _cases(nstr):
        if "mixed thin" in name:
            continue          # the checker itself is only good to 1e-3 there (see above)
        L = len(dt)
        kw = dict(kw)
        plank = kw.pop("plank", False)
        temper = np.linspace(210.0, 300.0, L + 1)
        okw = dict(kw)
        if plank:
            okw.update(plank=True, temper=temper, wvnmlo=600.0, wvnmhi=700.0, btemp=310.0, ttemp=100.0, temis=1.0)
        ref = oracle.disort(dt, ss, pm, nstr=nstr, umu=umu, phi=phi, phi0=30.0, onlyfl=False, **okw)
        bins = sb.make_bins(1, plank=int(plank), phi0=30.0, **{k: v for k, v in okw.items() if k not in ("plank", "temper")})
        got = s.disort_batch(dt[None], ss[None], pm[None], bins, nstr=nstr, temper=temper[None], umu=umu, phi=phi)
        assert got["status"][0] == ref["status"], (name, got["status"][0], ref["status"])
        if ref["status"] != 0:
            continue
        scale = max(np.abs(ref["uu"]).max(), np.abs(ref["flup"]).max() / np.pi, np.abs(ref["rfldn"]).max() / np.pi)
        err = np.abs(got["uu"][0] - ref["uu"]).max() / scale
        worst = max(worst, err)
        tol = 1e-5 if "albedo 1" in name else 1e-6
        assert err <= tol, (name, err)
    s.close()
    print("worst", worst)


@pytest.mark.parametrize("umu", [[1.0], [-1.0], [-1.0, 1.0], [0.3], [-0.99999, 0.5]])
def test_special_user_angle_sets(umu):
    """DISORT sums the m = 0 mode only when every user direction is vertical or the sun is overhead
    (disort.f:577-586); one and two user angles are the special cases of that rule."""
    from sbdart_b200 import workloads
    umu = np.array(umu)
    phi = np.array([25.0])
    w = workloads.retrieval_batch(6, nstr=8, nlyr=7, ncols=2, seed=5)
    w["bins"]["phi0"] = 0.0
    w["bins"]["umu0"][3] = 1.0 - 1e-6          # sun overhead: one mode
    s = sb.Solver(0)
    got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=umu, phi=phi)
    s.close()
    b = w["bins"]
    for i in range(len(b)):
        r = oracle.disort(w["dtauc"][i], w["ssalb"][i], w["pmom"][i], nstr=8, umu=umu, phi=phi, fbeam=b["fbeam"][i],
                          umu0=b["umu0"][i], phi0=0.0, fisot=b["fisot"][i], albedo=b["albedo"][i], onlyfl=False)
        assert got["status"][i] == r["status"]
        if r["status"] == 0:
            scale = max(np.abs(r["uu"]).max(), np.abs(r["flup"]).max() / np.pi)
            assert np.abs(got["uu"][i] - r["uu"]).max() <= 1e-7 * scale, (i, umu)


@pytest.mark.parametrize("kernel", ["register", "general"])
def test_viewing_against_the_beam(monkeypatch, kernel):
    """|1 + umu/umu0| < 1e-4 but not zero: the reference's L'Hospital form of the beam term takes its
    exponential at the user level for every layer of the path (disort.f:4560-4566), which is not what an
    exact integration gives at that distance from the limit (relative difference ~ dtau x 1e-4 / umu) --
    parity means the reference's form."""
    from sbdart_b200 import workloads
    if kernel == "general":
        monkeypatch.setenv("SBD_FORCE_GENERIC", "1")
    w = workloads.retrieval_batch(6, nstr=8, nlyr=9, ncols=2, seed=9)
    w["bins"]["umu0"] = 0.6
    w["bins"]["phi0"] = 0.0
    w["dtauc"] *= 3.0
    umu = np.array([-0.6 * (1.0 + 8.0e-5), -0.6 * (1.0 - 5.0e-5), 0.4])        # increasing, as CHEKIN demands
    phi = np.array([180.0, 0.0])
    s = sb.Solver(0)
    got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=umu, phi=phi)
    s.close()
    b = w["bins"]
    for i in range(len(b)):
        r = oracle.disort(w["dtauc"][i], w["ssalb"][i], w["pmom"][i], nstr=8, umu=umu, phi=phi, fbeam=b["fbeam"][i],
                          umu0=0.6, phi0=0.0, fisot=b["fisot"][i], albedo=b["albedo"][i], onlyfl=False)
        assert got["status"][i] == r["status"]
        if r["status"] == 0:
            scale = max(np.abs(r["uu"]).max(), np.abs(r["flup"]).max() / np.pi)
            assert np.abs(got["uu"][i] - r["uu"]).max() <= 1e-7 * scale, i


def test_radiances_with_negligible_layers_and_many_angles():
    """Layers thinner than 1e-6 in scaled depth (the reference drops their source integral for the level
    they touch, disort.f:4635-4641) between ordinary ones; 36 user angles x 5 azimuths (more angles
    than lanes)."""
    nstr = 8
    dt = np.array([1e-8, 0.5, 1e-7, 0.3, 1e-9, 0.7, 2e-7])
    ss = np.array([0.9, 0.95, 0.5, 0.8, 1.0, 0.6, 0.99])
    pm = _hg([0.1, 0.8, 0.3, 0.7, 0.2, 0.6, 0.4], nstr + 2)
    umu = np.concatenate([-np.linspace(1.0, 0.02, 18), np.linspace(0.02, 1.0, 18)])
    phi = np.array([0.0, 45.0, 90.0, 135.0, 180.0])
    kw = dict(fbeam=1.0, umu0=0.55, albedo=0.35, fisot=0.02)
    ref = oracle.disort(dt, ss, pm, nstr=nstr, umu=umu, phi=phi, phi0=15.0, onlyfl=False, **kw)
    assert ref["status"] == 0
    for env in (None, "SBD_FORCE_GENERIC"):
        import os
        if env:
            os.environ[env] = "1"
        try:
            s = sb.Solver(0)
            got = s.disort_batch(dt[None], ss[None], pm[None], sb.make_bins(1, phi0=15.0, **kw), nstr=nstr, umu=umu, phi=phi)
            s.close()
        finally:
            if env:
                del os.environ[env]
        assert got["status"][0] == 0
        scale = np.abs(ref["uu"]).max()
        assert np.abs(got["uu"][0] - ref["uu"]).max() <= 1e-7 * scale, env


def test_more_user_angles_than_the_register_kernel_holds():
    """200 user angles x 3 azimuths at NSTR 32: the register kernel's work area does not fit the SM's
    shared memory, the general kernel takes the run (no error)."""
    nstr = 32
    rng = np.random.default_rng(2)
    L = 4
    dt = rng.uniform(0.05, 1.0, L); ss = rng.uniform(0.5, 1.0, L)
    pm = _hg(rng.uniform(0.0, 0.8, L), nstr + 2)
    umu = np.concatenate([-np.linspace(1.0, 0.01, 100), np.linspace(0.01, 1.0, 100)])
    phi = np.array([0.0, 90.0, 180.0])
    kw = dict(fbeam=1.0, umu0=0.5, albedo=0.2)
    ref = oracle.disort(dt, ss, pm, nstr=nstr, umu=umu, phi=phi, phi0=0.0, onlyfl=False, **kw)
    s = sb.Solver(0)
    got = s.disort_batch(dt[None], ss[None], pm[None], sb.make_bins(1, **kw), nstr=nstr, umu=umu, phi=phi)
    s.close()
    assert got["status"][0] == 0 and ref["status"] == 0
    assert np.abs(got["uu"][0] - ref["uu"]).max() <= 1e-7 * np.abs(ref["uu"]).max()


@pytest.mark.parametrize("nstr,rad", [(16, False), (32, False), (8, True), (20, True)])
def test_maximum_layer_count(nstr, rad):
    """SBD_MAX_NLYR = 128 layers: the CTA shapes shrink to what the shared memory holds."""
    from sbdart_b200 import workloads
    w = workloads.retrieval_batch(6, nstr=nstr, nlyr=128, ncols=2, seed=3)
    w["dtauc"] *= 0.3
    umu = np.array([-0.7, -0.2, 0.3, 0.9]) if rad else None
    phi = np.array([0.0, 120.0]) if rad else None
    s = sb.Solver(0)
    got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr, umu=umu, phi=phi)
    s.close()
    b = w["bins"]
    for i in range(len(b)):
        r = oracle.disort(w["dtauc"][i], w["ssalb"][i], w["pmom"][i], nstr=nstr, umu=umu, phi=phi, fbeam=b["fbeam"][i],
                          umu0=b["umu0"][i], phi0=b["phi0"][i], fisot=b["fisot"][i], albedo=b["albedo"][i], onlyfl=not rad)
        assert got["status"][i] == r["status"]
        if r["status"] != 0:
            continue
        scale = max(np.abs(r[k]).max() for k in ("rfldir", "rfldn", "flup"))
        for k in ("rfldir", "rfldn", "flup"):
            assert np.abs(got[k][i] - r[k]).max() <= 1e-7 * scale, (i, k)
        if rad:
            assert np.abs(got["uu"][i] - r["uu"]).max() <= 1e-7 * max(np.abs(r["uu"]).max(), scale / np.pi), i


@pytest.mark.parametrize("nstr,rad", [(16, False), (12, False), (32, False), (8, True), (20, True)])
def test_rejected_inputs_are_reported_per_bin(nstr, rad):
    """CHEKIN's input checks (disort.f:4920-5155) end the reference program; here the offending bin
    reports a status and the other bins of the batch are solved as usual."""
    from sbdart_b200 import workloads
    w = workloads.retrieval_batch(12, nstr=nstr, nlyr=6, ncols=2, seed=13)
    good = {k: w[k].copy() for k in ("dtauc", "ssalb", "pmom")}
    good_bins = w["bins"].copy()
    b = w["bins"]
    w["ssalb"][1, 2] = 1.0 + 1e-9
    w["ssalb"][2, 0] = -0.1
    w["dtauc"][3, 4] = np.nan
    w["pmom"][4, 1, 3] = 1.5
    b["fbeam"][5] = -1.0
    b["umu0"][6] = 1.5
    b["albedo"][7] = 1.2
    b["fisot"][8] = -0.5
    b["umu0"][9] = 0.0
    umu = np.array([-0.5, 0.4]) if rad else None
    phi = np.array([0.0]) if rad else None
    s = sb.Solver(0)
    got = s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], b, nstr=nstr, umu=umu, phi=phi)
    ref = s.disort_batch(good["dtauc"], good["ssalb"], good["pmom"], good_bins, nstr=nstr, umu=umu, phi=phi)
    s.close()
    bad = np.arange(1, 10)
    assert (got["status"][bad] == sb.BIN_BAD_INPUT).all(), got["status"]
    ok = np.array([0, 10, 11])
    assert (got["status"][ok] == 0).all()
    for k in ("rfldir", "rfldn", "flup") + (("uu",) if rad else ()):
        assert np.array_equal(got[k][ok], ref[k][ok]), k      # placement independent, bit for bit


def test_argument_errors():
    from sbdart_b200 import workloads
    w = workloads.retrieval_batch(2, nstr=8, nlyr=4, ncols=1, seed=1)
    s = sb.Solver(0)
    with pytest.raises(sb.SbdError):
        s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=7)           # odd
    with pytest.raises(sb.SbdError):
        s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=42)          # > SBD_MAX_NSTR
    with pytest.raises(sb.SbdError):
        s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=np.array([0.5, -0.5]), phi=np.array([0.0]))   # not increasing
    with pytest.raises(sb.SbdError):
        s.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=8, umu=np.array([0.0, 0.5]), phi=np.array([0.0]))    # horizontal
    s.close()
