"""Bidirectional surface reflectance (isalb = 7, 8, 9): the BDREF models of spectra.f:249-1357
and DISORT's SURFAC quadrature (disort.f:3639-3912) that turns them into the per-mode tables
the solver consumes.

    bdref(model, wl, mur, mui, phir)    ocean (6S sun glint + foam + Morel case-I water),
                                        Hapke, Ross-thick / Li-sparse
    surface_tables(...)                 BDR[m][iq][0..n], BEM[iq], RMU[m][iu][0..n], EMU[iu]

Default-REAL literals of the reference are rounded to float32 like everywhere in the front end.
Reference quirks kept: `suralb` takes the salinity from sc(3) (the manual says sc(4),
spectra.f:139-144) and `seabdrf` hands the chlorophyll concentration to `indwat` as the salinity
(spectra.f:455).
"""
from __future__ import annotations

import numpy as np

PI = 3.1415926536          # params.f:29 (a double-precision literal there)


def f32(x):
    return float(np.float32(x))


class SurfaceModel:
    """suralb (spectra.f:61-176) for isalb = 7, 8, 9: ibdrf and the model parameters."""

    def __init__(self, isalb: int, sc):
        sc = [float(v) for v in sc] + [0.0] * 5
        self.ibdrf = {7: 1, 8: 2, 9: 3}[abs(int(isalb))]
        if self.ibdrf == 1:
            self.chlor, self.wndspd, self.salin = sc[0], sc[1], sc[2]
        elif self.ibdrf == 2:
            self.hssa, self.hasym, self.hotspt, self.hotwdth = sc[0:4]
        else:
            self.rliso, self.rlvol, self.rlgeo, self.rlhot, self.rlwdth = sc[0:5]

    @property
    def spectral(self):
        return self.ibdrf == 1          # only the ocean model depends on the wavelength

    def params(self):
        if self.ibdrf == 1:
            return [self.chlor, self.wndspd, self.salin]
        if self.ibdrf == 2:
            return [self.hssa, self.hasym, self.hotspt, self.hotwdth]
        return [self.rliso, self.rlvol, self.rlgeo, self.rlhot, self.rlwdth]


def _locate(xx, x):
    """Numerical-Recipes LOCATE as used by the reference: j with xx[j] <= x < xx[j+1] (1-based j,
    0 below the table, n at / above its end), for an ascending table."""
    n = len(xx)
    jl, ju = 0, n + 1
    while ju - jl > 1:
        jm = (ju + jl) // 2
        if x > xx[jm - 1]:
            jl = jm
        else:
            ju = jm
    return jl


def indwat(tables, wl, xsal):
    """Refractive index of sea water (spectra.f:594-1222)."""
    wltab, mr, mi = (tables["spectra/indwat/" + k] for k in ("wltab", "mrtab", "mitab"))
    i = _locate(wltab, wl)
    i = min(max(i, 1), len(wltab) - 1)
    wt = (wl - wltab[i - 1]) / (wltab[i] - wltab[i - 1])
    wt = max(0.0, min(1.0, wt))
    nr = mr[i - 1] * (mr[i] / mr[i - 1]) ** wt
    ni = mi[i - 1] * (mi[i] / mi[i - 1]) ** wt
    nr = nr + f32(0.006) * (xsal / f32(34.3))
    ni = ni + f32(0.000) * (xsal / f32(34.3))
    return nr, ni


def morcasiwat(tables, wl, c):
    """Sub-surface reflectance of case-I water, Morel 1988 (spectra.f:467-592)."""
    if wl < f32(0.400) or wl > f32(0.700):
        return 0.0
    t = {k: tables["spectra/morcasiwat/" + k] for k in ("tkw", "txc", "te", "tbw")}
    iwl = int(np.rint((wl - f32(0.400)) / f32(0.005)))          # 0-based
    kw, xc, e, bw = (t[k][iwl] for k in ("tkw", "txc", "te", "tbw"))
    if abs(c) < f32(0.0001):
        bb, kd = f32(0.5) * bw, kw
    else:
        b = f32(0.30) * c ** f32(0.62)
        bbt = f32(0.002) + f32(0.02) * (f32(0.5) - f32(0.25) * np.log10(c)) * f32(0.550) / wl
        bb = f32(0.5) * bw + bbt * b
        kd = kw + xc * c ** e
    u1 = f32(0.75)
    r1 = f32(0.33) * bb / u1 / kd
    while True:
        u2 = f32(0.90) * (1.0 - r1) / (1.0 + f32(2.25) * r1)
        rsw = f32(0.33) * bb / u2 / kd
        if abs((rsw - r1) / rsw) < f32(0.0001):
            return float(rsw)
        r1 = rsw


def fresnel(nr, ni, coschi, sinchi):
    """spectra.f:1320-1352"""
    a1 = np.abs(nr * nr - ni * ni - sinchi * sinchi)
    a2 = np.sqrt((nr * nr - ni * ni - sinchi * sinchi) ** 2 + 4 * nr * nr * ni * ni)
    u = np.sqrt(0.5 * (a1 + a2))
    v = np.sqrt(0.5 * (-a1 + a2))
    rr2 = ((coschi - u) ** 2 + v * v) / ((coschi + u) ** 2 + v * v)
    b1 = (nr * nr - ni * ni) * coschi
    b2 = 2 * nr * ni * coschi
    rl2 = ((b1 - u) ** 2 + (b2 + v) ** 2) / ((b1 + u) ** 2 + (b2 - v) ** 2)
    return (rr2 + rl2) / 2.0


def sunglint(wndspd, nr, ni, csin, cvin, phi):
    """Cox-Munk sun glint averaged over the wind direction (spectra.f:1224-1316)."""
    cs = np.maximum(csin, f32(0.05))
    cv = np.maximum(cvin, f32(0.05))
    ss = np.sqrt(1.0 - cs ** 2)
    sv = np.sqrt(1.0 - cv ** 2)
    zx = -sv * np.sin(PI - phi) / (cs + cv)
    zy = (ss + sv * np.cos(PI - phi)) / (cs + cv)
    tilt = np.arctan(np.sqrt(zx * zx + zy * zy))
    sigmac = f32(0.003) + f32(0.00192) * wndspd
    sigmau = f32(0.00316) * wndspd
    c40, c22, c04 = f32(0.40), f32(0.12), f32(0.23)
    r2 = zx ** 2 + zy ** 2
    axe2 = f32(.5) * r2 / sigmac
    axn2 = f32(.5) * r2 / sigmau
    q4 = 3 * zx ** 4 + 6 * zx ** 2 * zy ** 2 + 3 * zy ** 4
    axe4 = q4 / (8 * sigmac ** 2)
    axn4 = q4 / (8 * sigmau ** 2)
    axe2xn2 = (zx ** 4 + 10 * zx ** 2 * zy ** 2 + zy ** 4) / (8 * sigmau * sigmac)
    coef = 1.0
    coef = coef + c40 / 24. * (axe4 - 6 * axe2 + 3)
    coef = coef + c04 / 24. * (axn4 - 6 * axn2 + 3)
    coef = coef + c22 / 4. * (axe2xn2 - axn2 - axe2 + 1)
    coef = coef / (2. * PI * np.sqrt(sigmau) * np.sqrt(sigmac))
    proba = coef * np.exp(-(axe2 + axn2) / 2.)
    cos2chi = cv * cs + sv * ss * np.cos(PI - phi)
    cos2chi = np.where(cos2chi > 1.0, 0.99999999999, cos2chi)
    cos2chi = np.where(cos2chi < -1.0, -0.99999999999, cos2chi)
    coschi = np.sqrt(0.5 * (1 + cos2chi))
    sinchi = np.sqrt(0.5 * (1 - cos2chi))
    r1 = fresnel(nr, ni, coschi, sinchi)
    return PI * r1 * proba / (4. * cs * cv * np.cos(tilt) ** 4)


def ocean_state(tables, model: SurfaceModel, wl):
    """The wavelength-dependent part of seabdrf (spectra.f:453-461): nr, ni, rsw, wndwt, rfoam."""
    nr, ni = indwat(tables, wl, model.chlor)
    rsw = morcasiwat(tables, wl, model.chlor)
    if model.chlor == 0.0:
        rsw = 0.0
    wndwt = f32(2.951e-6) * model.wndspd ** f32(3.52)
    rfoam = wndwt * f32(0.22)
    return dict(nr=float(nr), ni=float(ni), rsw=float(rsw), wndwt=float(wndwt), rfoam=float(rfoam))


def bdref(model: SurfaceModel, state, mur, mui, phir):
    """BDREF (spectra.f:249-296): mur / mui = cosines of reflection / incidence (positive),
    phir = azimuth difference in radians; numpy broadcasting over the three."""
    mur, mui, phir = np.broadcast_arrays(np.asarray(mur, float), np.asarray(mui, float), np.asarray(phir, float))
    if model.ibdrf == 1:                       # seabdrf(wl, mus = mui, muv = mur, phir)
        rgl = sunglint(model.wndspd, state["nr"], state["ni"], mui, mur, phir)
        return state["rfoam"] + (1. - state["wndwt"]) * rgl + (1. - state["rfoam"]) * state["rsw"]
    if model.ibdrf == 2:                       # hapkbdrf(ui = mui, ur = mur)
        ui, ur = mui, mur
        coss = ui * ur + np.sqrt(1. - ur ** 2) * np.sqrt(1. - ui ** 2) * np.cos(PI - phir)
        s = np.arccos(coss)
        g = model.hasym
        pfun = (1. - g ** 2) / (1 + g ** 2 + 2 * g * coss) ** 1.5
        pfun0 = (1. - g ** 2) / (1 + g) ** 3
        b0 = model.hotspt / (model.hssa * pfun0)
        bfun = b0 / (1. + np.tan(s / 2) / model.hotwdth)
        rt = np.sqrt(1. - model.hssa)
        hfunr = (1. + 2 * ur) / (1. + 2. * ur * rt)
        hfuni = (1. + 2 * ui) / (1. + 2. * ui * rt)
        bd = (1. + bfun) * pfun + hfunr * hfuni - 1.
        return f32(.25) * model.hssa * bd / (ur + ui)
    # rtlsbdrf(mui, mur, phir): Ross-thick, Li-sparse (spectra.f:350-419)
    ui = np.maximum(mui, f32(.01))
    ur = np.maximum(mur, f32(.01))
    cosra = np.cos(PI - phir)
    coss = np.clip(ui * ur + np.sqrt(1. - ur ** 2) * np.sqrt(1. - ui ** 2) * cosra, -1.0, 1.0)
    s = np.arccos(coss)
    f1 = (PI / 2 - s) * coss + np.sin(s)
    f1 = f1 / (ui + ur) - PI / 4.
    vza, sza = np.arccos(ur), np.arccos(ui)
    tanvzap = model.rlwdth * np.tan(vza)
    tanszap = model.rlwdth * np.tan(sza)
    if model.rlwdth == 1.0:
        vzap, szap = vza, sza
    else:
        vzap, szap = np.arctan(tanvzap), np.arctan(tanszap)
    cossp = np.clip(np.cos(szap) * np.cos(vzap) + np.sin(szap) * np.sin(vzap) * cosra, -1.0, 1.0)
    dd = tanszap ** 2 + tanvzap ** 2 - 2 * tanszap * tanvzap * cosra
    secsum = 1. / np.cos(szap) + 1. / np.cos(vzap)
    cost = model.rlhot * np.sqrt(dd + (tanszap * tanvzap * np.sin(PI - phir)) ** 2)
    cost = np.clip(cost / secsum, -1.0, 1.0)
    t = np.arccos(cost)
    f2 = (t - np.sin(t) * cost) * secsum / PI
    f2 = f2 - 1. / np.cos(vzap) + f32(.5) * (1. + cossp) / (np.cos(szap) * np.cos(vzap))
    return model.rliso + model.rlvol * f1 + model.rlgeo * f2


def _gauss01(m):
    """Gauss-Legendre nodes / weights on (0, 1) (QGAUSN, disort.f:5984)."""
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


NMUG = 50


def surface_tables(model: SurfaceModel, state, cmu, umu0, fbeam_on, nmodes, umu=None, pi=3.1415927410125732):
    """SURFAC for a non-Lambertian surface (disort.f:3765-3907).

    cmu: the n positive quadrature cosines; umu: user cosines (radiance runs) or None.
    Returns dict(bdr [M][n][n+1], bem [n], rmu [M][NU][n+1], emu [NU]); column 0 of bdr / rmu is
    the direct-beam direction umu0.  `pi` is DISORT's single-precision constant (disort.f:441).
    """
    cmu = np.asarray(cmu, float)
    n = len(cmu)
    g, w = _gauss01(NMUG // 2)
    gmu = np.concatenate([g, -g])
    gwt = np.concatenate([w, w])
    phi = pi * gmu

    def fourier(mur, mui):
        """[M][len(mur)][len(mui)]: 0.5 (2 - delta_m0) sum_k gwt_k bdref cos(m pi gmu_k)"""
        vals = bdref(model, state, mur[:, None, None], mui[None, :, None], phi[None, None, :])
        out = np.empty((nmodes, len(mur), len(mui)))
        for m in range(nmodes):
            out[m] = (0.5 * (2.0 - (1.0 if m == 0 else 0.0))) * (vals * (gwt * np.cos(m * pi * gmu))).sum(axis=2)
        return out

    def emissivity(mur):
        """1 - directional reflectivity (disort.f:3806-3829)"""
        vals = bdref(model, state, mur[:, None, None], g[None, :, None], phi[None, None, :])   # [r][k][jg]
        inner = (vals * (w * g)[None, :, None]).sum(axis=1)
        return 1.0 - (inner * gwt[None, :]).sum(axis=1)

    inc = np.concatenate([[umu0], cmu])
    bdr = fourier(cmu, inc)
    if not fbeam_on:
        bdr[:, :, 0] = 0.0
    out = dict(bdr=bdr, bem=emissivity(cmu))
    if umu is not None:
        umu = np.asarray(umu, float)
        up = umu > 0.0
        rmu = np.zeros((nmodes, len(umu), n + 1))
        emu = np.zeros(len(umu))
        if up.any():
            rmu[:, up, :] = fourier(umu[up], inc)
            if not fbeam_on:
                rmu[:, :, 0] = 0.0
            emu[up] = emissivity(umu[up])
        out.update(rmu=rmu, emu=emu)
    return out


class SurfaceSpec:
    """The BRDF surfaces of one run: one per wavelength for the ocean model, one in all
    otherwise.  `tables(nstr)` returns SURFAC's arrays stacked over the surfaces, in the layout
    of sbd_set_surfaces (include/sbdart_b200.h)."""

    def __init__(self, model: SurfaceModel, states, umu0, fbeam_on, umu=None):
        self.model, self.states = model, list(states)
        self.umu0, self.fbeam_on, self.umu = float(umu0), bool(fbeam_on), umu
        self._cache = {}

    def __len__(self):
        return len(self.states)

    def tables(self, nstr, quadrature=None):
        if nstr not in self._cache:
            cmu = (quadrature or _gauss01)(nstr // 2)[0]
            nmodes = nstr if self.umu is not None else 1
            tabs = [surface_tables(self.model, st, cmu, self.umu0, self.fbeam_on, nmodes, umu=self.umu)
                    for st in self.states]
            self._cache[nstr] = {k: np.stack([t[k] for t in tabs]) for k in tabs[0]}
        return self._cache[nstr]
