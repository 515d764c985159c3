"""The flux algorithm of the `adding` kernel (tests/adding_model.py, numpy) against the CPU
oracle: layer reflection / transmission operators + interaction principle give the same
fluxes as DISORT's boundary-value system (disort.f:2702-3616, :1780-2006)."""
import ctypes as C

import numpy as np
import pytest

import adding_model as am
from oracle import oracle


def _quad(n):
    mu, w = np.zeros(n), np.zeros(n)
    oracle.lib().sbdo_qgausn(n, oracle._dp(mu), oracle._dp(w))
    return mu, w


def _plk(lo, hi, t):
    ie = C.c_int(0)
    return oracle.lib().sbdo_plkavg(lo, hi, t, C.byref(ie))


def _both(dt, ss, pm, nstr, temper=None, **kw):
    okw = dict(kw)
    pk, tpl, bpl = None, 0.0, 0.0
    if temper is not None:
        okw.update(plank=True, temper=temper, wvnmlo=800.0, wvnmhi=900.0, btemp=305.0, ttemp=250.0, temis=0.5)
        pk = np.array([_plk(800.0, 900.0, t) for t in temper])
        tpl, bpl = 0.5 * _plk(800.0, 900.0, 250.0), _plk(800.0, 900.0, 305.0)
    ref = oracle.disort(dt, ss, pm, nstr=nstr, **okw)
    mu, w = _quad(nstr // 2)
    out = am.fluxes(dt, ss, pm, nstr, mu, w, pk=pk, tplank=tpl, bplank=bpl, **kw)
    return ref, out


def _worst(ref, out):
    sc = max(np.abs(ref[k]).max() for k in ("rfldir", "rfldn", "flup"))
    worst = 0.0
    for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg"):
        s = max(sc, np.abs(ref[k]).max()) if k == "dfdt" else sc
        worst = max(worst, np.abs(out[k] - ref[k]).max() / s)
    return worst


@pytest.mark.parametrize("nstr", [4, 8, 16])
def test_adding_fluxes_match_oracle(nstr):
    rng = np.random.default_rng(100 + nstr)
    for case in range(24):
        L = int(rng.integers(1, 34))
        plank = case % 2 == 1
        # (the reference's thermal particular solution is itself noisy for conservative or very
        # thin layers with a temperature step, see the next test: keep those to the beam cases)
        dt = 10 ** rng.uniform(-3 if plank else -6, 1.0, L)
        ss = 1 - 10 ** rng.uniform(-4 if plank else -7, 0, L)
        if case % 6 == 0:
            ss[:] = 1.0                                       # conservative (dithered, disort.f:486)
        if case % 8 == 2:
            ss[:] = 0.3; dt[:] = 2.0                          # absorbing: NCUT truncation
        if case % 6 == 4:
            dt[L // 2] = 0.0                                  # an empty layer
        g = rng.uniform(0, 0.9, L)
        pm = g[:, None] ** np.arange(nstr + 3)[None, :]
        kw = dict(fbeam=0.0 if case % 4 == 3 else 1.0, umu0=float(rng.uniform(0.15, 0.98)),
                  albedo=float(rng.uniform(0, 1)), fisot=0.1 if case % 3 == 0 else 0.0)
        temper = np.linspace(220, 300, L + 1) + rng.uniform(-5, 5, L + 1) if plank else None
        ref, out = _both(dt, ss, pm, nstr, temper, **kw)
        assert ref["status"] == 0
        assert _worst(ref, out) < 2e-8, (case, L, _worst(ref, out))


def test_thin_layer_with_temperature_step():
    """A layer of optical depth 1e-9 cannot change the fluxes below it by more than ~1e-8.  The
    adding form honours that; DISORT's own particular solution (Z0 = B -+ (dB/dtau) q, 1e9-sized
    terms that cancel) moves by 4e-6 -- the reference's round-off, not a parity target."""
    nstr = 16
    g = np.array([0.1, 0.8, 0.5, 0.85])
    pm = g[:, None] ** np.arange(nstr + 3)[None, :]
    kw = dict(fbeam=0.0, umu0=0.6, albedo=0.3, fisot=0.0)
    dt4, t4 = np.array([1e-9, 2.0, 0.1, 5.0]), np.array([220.0, 240, 260, 280, 300])
    ss = np.full(4, 1.0 - 1e-6)
    ref4, out4 = _both(dt4, ss, pm, nstr, t4, **kw)
    ref3, out3 = _both(dt4[1:], ss[1:], pm[1:], nstr, t4[1:], **kw)
    assert abs(out3["flup"][0] / ref3["flup"][0] - 1) < 1e-10
    assert abs(out4["flup"][1] / out3["flup"][0] - 1) < 1e-8
    assert abs(ref4["flup"][1] / ref3["flup"][0] - 1) > 1e-7        # documents the reference's noise
