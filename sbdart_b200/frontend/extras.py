"""Front-end pieces beyond the five TestRuns cases: the variable
vertical grid, boundary-layer and stratospheric aerosols, in-cloud humidity,
solar geometry from date/time/place and the sensor filter functions.

  zgrid                    atms.f:505-617
  satcloud, saturate       atms.f:11-221
  relhum                   tauaero.f:1499-1524
  aeroblk / tauaero        tauaero.f:12-1497  (aerzstd, stdaer, usraer, aerbwi,
                           aestrat, phaerw, aeroden, aervint, denprfl)
  zensun                   spectra.f:4440-4556
  setfilt / filter         spectra.f:3240-3411 (isat -4 .. 29)

None of the reference's stored outputs exercises these options, so
they are checked by properties in tests/test_frontend_extras.py, not by a
golden.  Default-REAL literals and REAL() conversions of the reference are
rounded through float32 where they appear (f32 / np.float32 below).
"""
from __future__ import annotations

import math

import numpy as np

from . import MXLY, PI_KR, TZERO, ZIP, T, f32, getmom, levrng, locate, numset, zlayer

NAERZ, NAERB, NAERW = 5, 150, 47
WL55 = f32(0.55)


# ------------------------------------------------------------------ user data files
def _records(path, nitems):
    """List-directed READ of `nitems` values per record: the first nitems numbers of every
    non-blank line (a Fortran `read(13,*) a, b` ignores the rest of the line)."""
    out = []
    with open(path) as f:
        for line in f:
            tok = [t for t in line.replace(",", " ").split() if t]
            if not tok:
                continue
            if len(tok) < nitems:
                raise ValueError(f"{path}: expected {nitems} values per line, got {line.strip()!r}")
            out.append([float(t.lower().replace("d", "e")) for t in tok[:nitems]])
    return out


def useratm(path="atms.dat"):
    """User atmosphere (useratm, atms.f:468-503): nz, then nz lines z p t wh wo from the top
    down (or bottom up: the arrays are reversed when needed).  Returns bottom-up arrays."""
    with open(path) as f:
        first = f.readline().replace(",", " ").split()
    nz = int(float(first[0]))
    if nz > MXLY:
        raise ValueError(f"error in USERATM: {nz} layers specified in ATMS.DAT, but current limit is {MXLY}")
    recs = []
    with open(path) as f:
        f.readline()
        for line in f:
            tok = [t for t in line.replace(",", " ").split() if t]
            if tok:
                recs.append([float(t.lower().replace("d", "e")) for t in tok[:5]])
            if len(recs) == nz:
                break
    if len(recs) < nz:
        raise ValueError("atms.dat: fewer levels than announced")
    a = np.array(recs)[::-1]                    # read(13,*) z(i), ... for i = nz, 1, -1
    if a[0, 0] > a[-1, 0]:
        a = a[::-1]
    return tuple(np.ascontiguousarray(a[:, k]) for k in range(5))


def rdspec(path, nmax=5000):
    """Two-column spectral file (rdspec, spectra.f:4382-4437): wavelength, value."""
    recs = _records(path, 2)
    if len(recs) > nmax:
        raise ValueError(f"Error in rdspec -- file {path} should not specify more than {nmax} values")
    a = np.array(recs)
    a = a[a[:, 0] != 0.]
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])


class ListReader:
    """List-directed sequential READs of a formatted file (`read(u,*) a, b, ...`): every READ
    starts on a new line and takes as many lines as its item list needs; what is left of
    the last line is skipped.  read() returns None at end of file (the END= branch)."""

    def __init__(self, path):
        with open(path) as f:
            self.lines = [[t for t in ln.replace(",", " ").split() if t] for ln in f]
        self.pos = 0

    def rewind(self):
        self.pos = 0

    def read(self, nitems):
        vals = []
        while len(vals) < nitems:
            if self.pos >= len(self.lines):
                return None
            tok = self.lines[self.pos]
            self.pos += 1
            if "/" in tok:                       # a slash ends the record: the rest keeps its value
                vals += tok[:tok.index("/")]
                vals += [None] * (nitems - len(vals))
                break
            vals += tok
        return [None if v is None else float(v.lower().replace("d", "e")) for v in vals[:nitems]]


class AerosolFile:
    """aeread (tauaero.f:1526-1714): boundary-layer aerosols from aerosol.dat (iaer = -1).

    File: `nn nmom`, then per wavelength `wl` and nn records `taer waer pm(1..nmom)` for the
    levels ns = nz-nn+1 .. nz (top-down layer index, nz = surface).  The reference keeps two
    wavelength slots (ind, 3-ind) and reads forward as the requested wavelength grows; that
    state machine is reproduced as it stands, including its treatment of a first wavelength
    at or below the file's first one (the record is then taken as spectrally uniform between
    wl/2 and 2 wl, and later records are still read)."""

    def __init__(self, nz, imoma, path="aerosol.dat"):
        self.nz, self.imoma = nz, imoma
        self.f = ListReader(path)
        self.wlbaer = [0.0, 0.0]
        self.wl0 = 0.0
        self.ind, self.more = 1, 1
        self.taer = np.zeros((nz + 1, 3))
        self.waer = np.zeros((nz + 1, 3))
        self.gaer = None

    def _layer_block(self, slot, zero_first):
        for i in range(self.ns, self.nz + 1):
            if zero_first:
                self.taer[i, slot] = 0.; self.waer[i, slot] = 0.; self.gaer[:, i, slot] = 0.
            rec = self.f.read(2 + self.nmom)
            if rec is None:
                raise ValueError(f"not enough aerosol records\n need {self.nz - self.ns + 1} records")
            self.taer[i, slot], self.waer[i, slot] = rec[0], rec[1]
            self.gaer[:, i, slot] = rec[2:]

    def __call__(self, wl, nmom_out, pmom):
        """Adds the aerosol moments x scattering depth to pmom[nz][nmom_out+1]; returns dtau, wbaer."""
        nz = self.nz
        if self.wlbaer[0] == 0.:
            hdr = self.f.read(2)
            nn, self.nmom = int(hdr[0]), int(hdr[1])
            self.ns = nz - nn + 1
            if self.ns <= 0:
                raise ValueError(f"nz  nn {nz} {nn}\n too many layers specified in aerosol.dat")
            self.gaer = np.zeros((self.nmom, nz + 1, 3))
            first = self.f.read(1)
            if first is None:
                raise ValueError("no data found in aerosol.dat")
            self.wlbaer[0] = first[0]
            self.wl0 = self.wlbaer[0]
            self._layer_block(1, False)
            self.ind = 1
        elif wl < min(self.wlbaer) and wl > self.wl0:
            self.f.rewind()
            self.f.read(2)
            self.wlbaer[0] = self.f.read(1)[0]
            self._layer_block(1, True)
            self.wlbaer[1] = 0.
            self.ind, self.more = 1, 1
        if self.more == 1:
            hit_eof = False
            while wl > max(self.wlbaer):
                self.more = 0
                rec = self.f.read(1)
                if rec is None:
                    hit_eof = True
                    break
                self.ind = 3 - self.ind
                self._layer_block(self.ind, True)
                self.wlbaer[self.ind - 1] = rec[0]
            if not hit_eof:
                self.more = 1
        ns, ind, oth = self.ns, self.ind, 3 - self.ind
        if self.wlbaer[1] == 0.:
            self.wlbaer = [.5 * wl, 2 * wl]
            self.taer[ns:, 2] = self.taer[ns:, 1]
            self.waer[ns:, 2] = self.waer[ns:, 1]
            self.gaer[:, ns:, 2] = self.gaer[:, ns:, 1]
            self.wl0 = 0.
        wt = math.log(wl / self.wlbaer[ind - 1]) / math.log(self.wlbaer[oth - 1] / self.wlbaer[ind - 1])
        wt = max(0.0, min(wt, 1.0))
        dtau, wbaer = np.zeros(nz), np.zeros(nz)
        from . import getmom, MAXMOM
        for i in range(ns, nz + 1):
            ta, tb = self.taer[i, ind], self.taer[i, oth]
            if min(ta, tb) > 0.:
                dtau[i - 1] = ta * (tb / ta) ** wt
            else:
                dtau[i - 1] = ta * (1. - wt) + tb * wt
            wbaer[i - 1] = self.waer[i, ind] * (1. - wt) + self.waer[i, oth] * wt
            if self.nmom == 1:
                gg = self.gaer[0, i, ind] * (1. - wt) + self.gaer[0, i, oth] * wt
                pm = getmom(self.imoma, gg, MAXMOM)[1:]
            else:
                pm = self.gaer[:, i, ind] * (1. - wt) + self.gaer[:, i, oth] * wt
            m = min(len(pm), nmom_out)
            pmom[i - 1, 1:m + 1] += pm[:m] * dtau[i - 1] * wbaer[i - 1]
        return dtau, wbaer


def usrcloud_table(nz, path="usrcld.dat"):
    """The file part of usrcloud (taucloud.f:203-211): records lwp, reff, fwp, reice, cldfrac from
    the lowest layer upwards; missing records and items behind a slash keep the defaults
    0, 8, 0, -1, 1.  Returned arrays are indexed like the reference's (1 = top layer)."""
    tab = np.tile(np.array([0., 8., 0., -1., 1.]), (nz, 1))
    f = ListReader(path)
    for i in range(nz, 0, -1):
        rec = f.read(5)
        if rec is None:
            break
        for k, v in enumerate(rec):
            if v is not None:
                tab[i - 1, k] = v
    return tab


# ------------------------------------------------------------------ k-distribution files
def ckatm(path="CKATM"):
    """gasinit, first part (taugas.f:7309-7330): `nz h2oden`, then z(1:nz), p(1:nz), t(1:nz)
    (list-directed).  Returns bottom-up z, p, t and the surface water-vapour density."""
    f = ListReader(path)
    hdr = f.read(2)
    nz, h2oden = int(hdr[0]), hdr[1]
    if nz > MXLY:
        raise ValueError(f"gasinit --- nz gt mxly {nz} {MXLY}")
    z, p, t = (np.array(f.read(nz)) for _ in range(3))
    if z[0] > z[-1]:
        z, p, t = z[::-1].copy(), p[::-1].copy(), t[::-1].copy()
    return z, p, t, h2oden


def cktau_records(nz, path="CKTAU"):
    """The records of the unformatted sequential file CKTAU (readk, taugas.f:7723-7725):
    iv, ib, nb, nk (integers), vnu0, vnu1, vnu2, etf, ewc, gw(1:nk), dtk(1:nz,1:nk) (REAL*4),
    each between two 4-byte record lengths."""
    recs = []
    with open(path, "rb") as f:
        raw = f.read()
    pos = 0
    while pos + 4 <= len(raw):
        n = int(np.frombuffer(raw, "<i4", 1, pos)[0])
        body = raw[pos + 4: pos + 4 + n]
        pos += 8 + n
        iv, ib, nb, nk = (int(v) for v in np.frombuffer(body, "<i4", 4, 0))
        fl = np.frombuffer(body, "<f4", 5 + nk + nz * nk, 16).astype(float)
        recs.append(dict(iv=iv, ib=ib, nb=nb, nk=nk, vnu0=fl[0], vnu1=fl[1], vnu2=fl[2], etf=fl[3], ewc=fl[4],
                         gw=fl[5:5 + nk].copy(), dtk=fl[5 + nk:].reshape(nk, nz).T.copy()))
    return recs


def write_cktau(path, recs):
    """Writes records in the layout cktau_records reads (test fixtures, converters)."""
    with open(path, "wb") as f:
        for r in recs:
            nk = len(r["gw"])
            body = np.array([r["iv"], r["ib"], r["nb"], nk], "<i4").tobytes()
            body += np.array([r["vnu0"], r["vnu1"], r["vnu2"], r["etf"], r["ewc"]], "<f4").tobytes()
            body += np.asarray(r["gw"], "<f4").tobytes()
            body += np.asarray(r["dtk"], "<f4").T.copy().tobytes()       # dtk(1:nz, 1:nk), column-major
            n = np.array([len(body)], "<i4").tobytes()
            f.write(n + body + n)


# ------------------------------------------------------------------ vertical grid
def zgrid(z, p, t, wh, wo, zgrid1, zgrid2, ngrid):
    """Regrid the model atmosphere to |ngrid| levels (atms.f:505-617).  Arrays are
    bottom-up as in the reference; returns new (z, p, t, wh, wo)."""
    nz = len(z)
    ng = min(MXLY, abs(int(ngrid)))
    tol = f32(0.99)
    a = zgrid1 * float(np.float32(ng - 1))
    ztop = z[nz - 1]
    if a >= ztop:
        a, beta = ztop, 0.0
    else:
        a = min(a, tol * (ztop - zgrid2))
        toprat = float(np.float32(ng - 2) / np.float32(ng - 1))
        beta = math.log(((ztop - zgrid2) / (toprat * a) - 1.) / (ztop / a - 1.)) / math.log(toprat)
    zz, pp, tt, hh, oo = (np.zeros(ng) for _ in range(5))
    j = 2
    for i in range(1, ng + 1):
        x = float(np.float32(i - 1) / np.float32(ng - 1))
        zz[i - 1] = a * x * (1. + (ztop / a - 1.) * x ** beta)
        jj = j
        while jj <= nz and not zz[i - 1] <= z[jj - 1]:
            jj += 1
        j = min(jj, nz)
        fz = (zz[i - 1] - z[j - 2]) / (z[j - 1] - z[j - 2])
        fz = min(max(fz, 0.0), 1.0)
        pp[i - 1] = p[j - 2] * (p[j - 1] / p[j - 2]) ** fz
        tt[i - 1] = t[j - 2] * (1. - fz) + t[j - 1] * fz
        for src, dst in ((wh, hh), (wo, oo)):
            if min(src[j - 1], src[j - 2]) > 0.:
                dst[i - 1] = src[j - 2] * (src[j - 1] / src[j - 2]) ** fz
            else:
                dst[i - 1] = src[j - 2] * (1. - fz) + src[j - 1] * fz
    return zz, pp, tt, hh, oo


# ------------------------------------------------------------------ humidity
def satden(a):
    """Saturation water-vapour density (g/m3) at a = tzero/T (atms.f:61-65)."""
    return a * math.exp(18.916758 - a * (14.845878 + a * 2.4918766))


def relhum(t, h2o):
    """tauaero.f:1499-1524."""
    return h2o / satden(TZERO / t)


def satcloud(lcld, t, rhcld, wh):
    """Set the humidity inside the cloud layers (krhclr=1; atms.f:11-59).  In place."""
    nz, n = len(t), len(lcld)
    for i in range(1, n + 1):
        lbot, ltop = levrng(lcld, i)
        if lbot != 0:
            for j in range(max(ltop - 1, 1), lbot + 1):
                jj = nz - j + 1
                wh[jj - 1] = rhcld * satden(TZERO / t[jj - 1])


def saturate(lcld, z, t, rhcld, wh):
    """In-cloud humidity with the column water vapour conserved (atms.f:70-221).  In place."""
    nz, n = len(z), len(lcld)

    def column(select):
        tot, zbot = 0.0, z[0]
        for i in range(1, select[1] + 1):
            if i == 1:
                ztop = .5 * (z[1] + zbot)
            elif i == nz:
                ztop = z[nz - 1]
            else:
                ztop = .5 * (z[i] + z[i - 1])
            tot += select[0](i, f32(.1) * (ztop - zbot))
            zbot = ztop
        return tot

    wvp = column((lambda i, dz: dz * wh[i - 1], nz))
    if wvp == 0.:
        raise ValueError("Error in saturate ---  original column water vapor is zero -- can not modify")
    for i in range(1, n + 1):
        lbot, ltop = levrng(lcld, i)
        if lbot != 0:
            for j in range(max(ltop - 1, 1), lbot + 1):
                jj = nz - j + 1
                wh[jj - 1] = -rhcld * satden(TZERO / t[jj - 1])
    wvpclr = column((lambda i, dz: dz * wh[i - 1] if wh[i - 1] > 0. else 0.0, nz - 1))
    wvpcld = column((lambda i, dz: -dz * wh[i - 1] if wh[i - 1] < 0. else 0.0, nz - 1))
    if wvpcld == 0:
        raise ValueError("Error in saturate --- water vapor density in cloud = 0 ?")
    if wvpclr == 0:
        cldfac, clrfac = wvp / wvpcld, f32(1.e-30)
    else:
        clrfac, cldfac = (wvp - wvpcld) / wvpclr, 1.
        if clrfac < 0:
            clrfac, cldfac = f32(1.e-30), wvp / wvpcld
    for i in range(nz):
        wh[i] = -cldfac * wh[i] if wh[i] < 0 else clrfac * wh[i]
    wvpn = 0.
    for i in range(nz - 1):
        dz, d1, d2 = z[i + 1] - z[i], wh[i], wh[i + 1]
        if abs(d1 - d2) <= f32(.001) * d1 or min(d1, d2) == 0.:
            du = .5 * dz * (d1 + d2)
        else:
            du = dz * (d1 - d2) / math.log(d1 / d2)
        wvpn += f32(.1) * du
    wh *= wvp / wvpn


# ------------------------------------------------------------------ solar geometry
def zensun(iday, time, alat, alon):
    """Solar zenith, azimuth (degrees) and flux multiplier (spectra.f:4440-4556)."""
    nday, eqt, dec = T("spectra/zensun/nday"), T("spectra/zensun/eqt"), T("spectra/zensun/dec")
    degpday = float(np.float32(360.) / np.float32(365.242))
    eccen, dayph = f32(0.01671), 2.0
    dtor = PI_KR / 180.
    dd = float((iday - 1) % 365 + 1)
    i = 0
    while i < 74 and not nday[i] > dd:
        i += 1
    i = min(i, 73)
    frac = (dd - nday[i - 1]) / (nday[i] - nday[i - 1])
    eqtime = eqt[i - 1] * (1. - frac) + frac * eqt[i]
    decang = dec[i - 1] * (1. - frac) + frac * dec[i]
    sunlat = decang
    sunlon = -15. * (time - 12. + eqtime / 60.)
    t0, t1 = (90. - alat) * dtor, (90. - sunlat) * dtor
    p0, p1 = alon * dtor, sunlon * dtor
    zz = math.cos(t0) * math.cos(t1) + math.sin(t0) * math.sin(t1) * math.cos(p1 - p0)
    xx = math.sin(t1) * math.sin(p1 - p0)
    yy = math.sin(t0) * math.cos(t1) - math.cos(t0) * math.sin(t1) * math.cos(p1 - p0)
    azimuth = math.atan2(xx, yy) / dtor
    zenith = math.acos(zz) / dtor
    rsun = 1. - eccen * math.cos(degpday * (dd - dayph) * dtor)
    return zenith, azimuth, 1. / rsun ** 2


# ------------------------------------------------------------------ sensor filters
_FILTERS = {1: "meteo", 2: "goese", 3: "goesw", 4: "avhr81", 5: "avhr82", 6: "avhr91", 7: "avhr92",
            8: "avhr101", 9: "avhr102", 10: "avhr111", 11: "avhr112", 12: "gtr1", 13: "gtr2",
            14: "nm410", 15: "nm936", 16: "mfrsr1", 17: "mfrsr2", 18: "mfrsr3", 19: "mfrsr4",
            20: "mfrsr5", 21: "mfrsr6", 22: "avhr83", 23: "avhr84", 24: "avhr85", 25: "setlow",
            26: "airs1", 27: "airs2", 28: "airs3", 29: "airs4"}


class Filter:
    """filter(w) (spectra.f:3390-3411) over the table setfilt prepared."""

    def __init__(self, wlfilt=None, filt=None):
        self.wl, self.f = wlfilt, filt

    def __call__(self, w):
        if self.wl is None:
            return 1.0
        i = locate(self.wl, w)
        wt = (w - self.wl[i - 1]) / (self.wl[i] - self.wl[i - 1])
        wt = max(0.0, min(1.0, wt))
        return self.f[i - 1] * (1. - wt) + self.f[i] * wt


def filter_table(isat, wlinf, wlsup):
    """The isat != 0/-2 cases of setfilt (spectra.f:3253-3330): wlmin, wlmax, Filter."""
    if wlsup == 0. and isat < -2:
        raise ValueError(f"Error -- WLSUP must be non-zero when ISAT={isat}")
    if isat == -4:                       # centre, equivalent width, gaussian
        sqpi = math.sqrt(PI_KR)
        xc, nnf = 2, 1000
        xlim = xc * sqpi
        wlmin, wlmax = wlinf - xc * wlsup, wlinf + xc * wlsup
        i = np.arange(nnf)
        xx = -xlim + i * (2 * xlim) / (nnf - 1)
        return wlmin, wlmax, Filter(wlmin + (wlmax - wlmin) * i / (nnf - 1), np.exp(-xx ** 2))
    if isat == -3:                       # centre, equivalent width, triangular
        wlmin, wlmax = wlinf - wlsup, wlinf + wlsup
        return wlmin, wlmax, Filter(np.array([wlmin, wlinf, wlmax]), np.array([0., 1., 0.]))
    if isat in _FILTERS:
        nm = _FILTERS[isat]
        sr = T(f"spectra/{nm}/sr").astype(float)
        wlmin, wlmax = float(T(f"spectra/{nm}/wmn")), float(T(f"spectra/{nm}/wmx"))
        # (wlmax-wlmin)*real(i-1)/(nnf-1) is evaluated left to right in double
        return wlmin, wlmax, Filter(wlmin + (wlmax - wlmin) * np.arange(len(sr)) / (len(sr) - 1), sr)
    if isat == -1:                       # filter.dat
        wlf, filt = rdspec("filter.dat", 1000)
        return wlf[0], wlf[-1], Filter(wlf, filt)
    raise ValueError(f"isat={isat} out of range")


# ------------------------------------------------------------------ aerosols
AERO_DEFAULTS = dict(
    zaer=[0.0] * NAERZ, taerst=[0.0] * NAERZ, jaer=[0] * NAERZ, zbaer=[ZIP] * MXLY,
    dbaer=[ZIP] * MXLY, vis=ZIP, tbaer=ZIP, abaer=0.0, wlbaer=[ZIP] * NAERB, qbaer=[ZIP] * NAERB,
    wbaer=[ZIP] * NAERB, gbaer=[ZIP] * NAERB, pmaer=[ZIP] * (NAERB * 299), rhaer=ZIP, imoma=3)


def _full(v, n, fill):
    a = np.full(n, fill, dtype=float)
    v = np.atleast_1d(np.asarray(v, dtype=float))
    a[: len(v)] = v
    return a


class Aerosols:
    """module aeroblk + tauaero (tauaero.f:12-1497).

    Everything the reference does on the first call (denprfl: vertical profile, spectral
    model, normalisation to vis / tbaer; zlayer of the stratospheric layers) is done in
    the constructor; __call__ is the per-wavelength part (tauaero.f:1223-1331)."""

    def __init__(self, p, z, rhaer):
        self.iaer = int(p["iaer"])
        self.imoma = int(p["imoma"])
        self.nosct = int(p["nosct"])
        self.abaer = float(p["abaer"])
        self.vis, self.tbaer = float(p["vis"]), float(p["tbaer"])
        self.jaer = _full(p["jaer"], NAERZ, 0).astype(int)
        self.zaer = _full(p["zaer"], NAERZ, 0.0)
        self.taerst = _full(p["taerst"], NAERZ, 0.0)
        self.zbaer = _full(p["zbaer"], MXLY, ZIP)
        self.dbaer = _full(p["dbaer"], MXLY, ZIP)
        self.wlbaer = _full(p["wlbaer"], NAERB, ZIP)
        self.qbaer = _full(p["qbaer"], NAERB, ZIP)
        self.wbaer = _full(p["wbaer"], NAERB, ZIP)
        self.gbaer = _full(p["gbaer"], NAERB, ZIP)
        self.pmaer = np.atleast_1d(np.asarray(p["pmaer"], dtype=float))
        self.awl = T("tauaero/aeroblk/awl")
        self.aerstr = T("tauaero/aestrat/aerstr")
        self.z = np.asarray(z, dtype=float)
        self.nz = len(z)
        self.npmaer = 0
        self.nwlbaer = NAERW
        self.afile = AerosolFile(self.nz, self.imoma) if self.iaer == -1 else None
        if self.iaer < -1 or self.iaer > 5:
            raise ValueError("iaer out of range [-1,5]")
        if self.jaer.min() < 0 or self.jaer.max() > 4:
            raise ValueError("jaer out of range [0,4]")
        self.active = self.iaer != 0 or bool((self.jaer != 0).any())
        self.laer = [0] * NAERZ
        if (self.jaer != 0).any():
            n = numset(0.0, self.taerst)
            self.laer = zlayer(self.z, self.zaer[:n]) + [0] * (NAERZ - n)
        self.dtsv = np.zeros(self.nz)
        if self.iaer > 0:
            self._denprfl(rhaer)

    # ---- vertical profile (aerzstd, aeroden, aervint)
    def _aerzstd(self):
        alt, a05, a23 = (T(f"tauaero/aerzstd/{k}") for k in ("alt", "aden05", "aden23"))
        if self.vis <= 0.:
            den = float(np.float32(1.) / np.float32(23.) - np.float32(1.) / np.float32(5.))
            wtv = (1. / self.vis - f32(0.2)) / den if self.vis != 0. else math.inf
        else:
            wtv = 1.                       # the reference's test is inverted: 23 km profile
        wtv = max(0.0, min(1.0, wtv))
        self.nzbaer = len(alt)
        self.zbaer[: self.nzbaer] = alt
        self.dbaer[: self.nzbaer] = a05 * (1. - wtv) + a23 * wtv

    def aeroden(self, zz):
        """tauaero.f:1134-1172."""
        z = max(0.0, min(100.0, zz))
        zb, db, n = self.zbaer, self.dbaer, self.nzbaer
        if z > zb[n - 1]:
            return 0.0
        k = locate(zb[:n], z)
        f = (z - zb[k - 1]) / (zb[k] - zb[k - 1])
        if min(db[k - 1], db[k]) <= 0.:
            return max(db[k - 1] * (1. - f) + db[k] * f, 0.0)
        return db[k - 1] * (db[k] / db[k - 1]) ** f

    def _aervint(self):
        """n(i) dz per DISORT layer, top-down (tauaero.f:1449-1497)."""
        z, nz = self.z, self.nz
        vint = np.zeros(nz)
        bup = z[0] < z[nz - 1]
        vint[0] = self.aeroden(100.) * 5.
        zu = max(z[0], z[nz - 1])
        for i in range(2, nz + 1):
            ii = nz - i + 1 if bup else i
            zd = z[ii - 1]
            vint[i - 1] = (zu - zd) * self.aeroden(zd)
            zu = zd
        return vint

    # ---- spectral models (stdaer, usraer)
    def _stdaer(self, humid):
        rhzone = [0., f32(.7), f32(.8), f32(.99)]
        tiny = f32(.00000001)
        self.abaer = 0.
        rhum = max(0.0, min(1.0, humid))
        j = 1 if rhum < rhzone[1] else (2 if rhum < rhzone[2] else 3)
        wt = (rhum - rhzone[j - 1]) / (rhzone[j] - rhzone[j - 1])
        pre = {1: "rur", 2: "urb", 3: "ocn", 4: "tro"}[abs(self.iaer)]
        tab = lambda s: T(f"tauaero/stdaer/{pre}{s}").reshape(4, NAERW)  # noqa: E731  [humidity][wl]
        self.nwlbaer = NAERW
        self.wlb = self.awl.copy()
        out = []
        for s in ("e", "a", "g"):
            t = tab(s)
            v1, v2 = np.maximum(t[j - 1], tiny), np.maximum(t[j], tiny)
            out.append(v1 * (v2 / v1) ** wt)
        self.aerext, self.aerabs, self.aerasm = out

    def _usraer(self):
        """User-defined spectral model, iaer=5 (tauaero.f:405-594).  Returns q55."""
        wlb, qb, wb, gb = self.wlbaer, self.qbaer, self.wbaer, self.gbaer
        nwl, nq, nw, ng = (numset(ZIP, a) for a in (wlb, qb, wb, gb))
        npm = numset(ZIP, self.pmaer)
        err = []
        if nwl == 0:
            qb[0], nq = 1., 1
            if nw != 1:
                err.append("specify one value of wbaer when wlbaer not set")
        elif nwl == 1:
            if nq > 1:
                err.append("number of elements must match: wlbaer, qbaer")
            elif nq == 0:
                qb[0] = 1.
            if nw != 1:
                err.append("number of elements must match: wlbaer, wbaer")
        else:
            if nwl != nq:
                err.append("number of elements must match: wlbaer, qbaer")
            if nwl != nw:
                err.append("number of elements must match: wlbaer, wbaer")
            if ng == 0:
                if npm >= 1:
                    if npm % nwl != 0:
                        err.append("incorrect number of phase function moments")
                    else:
                        npm //= nwl
                else:
                    err.append("must specify either gbaer or pmaer")
            elif ng != nw:
                err.append("number of elements must match: wlbaer, gbaer")
        if ((ng == 0) == (npm == 0)) and self.imoma == 3:
            err.append("must specify either gbaer or pmaer, not both")
        if err:
            raise ValueError("Error in user specified aerosols (iaer=5): " +
                             "; ".join("Error in USRAER -- " + e for e in err))
        pm = self.pmaer.copy()
        if nw == 1:
            if nwl == 0:
                wlb[0] = WL55
            wlb[1] = 2 * wlb[0]
            nwl = 2
            qb[1] = qb[0] * (wlb[0] / wlb[1]) ** self.abaer
            wb[1], gb[1] = wb[0], gb[0]
            if npm > 0:
                pm = np.repeat(pm[:npm], 2)
        self.nwlbaer = nwl
        self.wlb = wlb[:nwl].copy()
        self.aerext = qb[:nwl].copy()
        self.aerabs = (1. - wb[:nwl]) * self.aerext
        self.aerasm = gb[:nwl].copy()
        self.npmaer = npm
        self.pm = pm[: npm * nwl].reshape(npm, nwl) if npm > 0 else None
        j = locate(self.wlb, WL55)
        f = math.log(WL55 / self.wlb[j - 1]) / math.log(self.wlb[j] / self.wlb[j - 1])
        if WL55 < self.wlb[0]:
            q55 = qb[0] * (self.wlb[0] / WL55) ** self.abaer
        elif WL55 > self.wlb[nwl - 1]:
            q55 = qb[nwl - 1] * (self.wlb[nwl - 1] / WL55) ** self.abaer
        else:
            q55 = qb[j - 1] * (qb[j] / qb[j - 1]) ** f
        if npm > 0:
            self.imoma = 0
        return q55

    def aerbwi(self, wl):
        """Boundary-layer extinction, SSA, asymmetry at wl (tauaero.f:177-250)."""
        if self.iaer == 0:
            return 0., 0., 0.
        wlb, ext, ab, asm, n = self.wlb, self.aerext, self.aerabs, self.aerasm, self.nwlbaer
        wa = 0.
        if wl <= wlb[0]:
            extinc = ext[0] * (wlb[0] / wl) ** self.abaer
            wa, ga = 1. - ab[0] / ext[0], asm[0]
        elif wl >= wlb[n - 1]:
            extinc = ext[n - 1] * (wlb[n - 1] / wl) ** self.abaer
            wa, ga = 1. - ab[n - 1] / ext[n - 1], asm[n - 1]
        else:
            k = locate(wlb[:n], wl)
            wt = math.log(wl / wlb[k - 1]) / math.log(wlb[k] / wlb[k - 1])
            extinc = ext[k - 1] * (ext[k] / ext[k - 1]) ** wt
            if ab[k - 1] > 0. and ab[k] > 0.:
                absorp = ab[k - 1] * (ab[k] / ab[k - 1]) ** wt
            else:
                absorp = ab[k - 1] * (1. - wt) + ab[k] * wt
            if extinc > 0.:
                wa = max(0.0, min(1. - absorp / extinc, 1.0))
            ga = (1. - wt) * asm[k - 1] + wt * asm[k]
        return extinc, wa, ga

    def aestrat(self, ja, wl):
        """Stratospheric models 1-4 (tauaero.f:253-403)."""
        awl, a = self.awl, self.aerstr[:, :, ja - 1]
        k = locate(awl, wl)
        wa = 0.
        if wl <= awl[0]:
            qa = a[0, 0] * (awl[0] / wl) ** self.abaer
            wa, ga = 1. - a[0, 1] / a[0, 0], a[k - 1, 2]
        elif wl >= awl[NAERW - 1]:
            qa = a[NAERW - 1, 0] * (awl[0] / wl) ** self.abaer     # awl(1) as in the reference
            wa, ga = 1. - a[NAERW - 1, 1] / a[NAERW - 1, 0], a[NAERW - 1, 2]
        else:
            wt = math.log(wl / awl[k - 1]) / math.log(awl[k] / awl[k - 1])
            qa = a[k - 1, 0] * (a[k, 0] / a[k - 1, 0]) ** wt
            absorp = a[k - 1, 1] * (a[k, 1] / a[k - 1, 1]) ** wt
            if qa > 0.:
                wa = max(0.0, min(1. - absorp / qa, 1.0))
            ga = (1. - wt) * a[k - 1, 2] + wt * a[k, 2]
        return qa, wa, ga

    def _phaerw(self, w):
        """User phase-function moments interpolated linearly, no extrapolation (tauaero.f:144-174)."""
        k = locate(self.wlb, w)
        wt = (w - self.wlb[k - 1]) / (self.wlb[k] - self.wlb[k - 1])
        wt = max(0.0, min(wt, 1.0))
        return self.pm[:, k - 1] * (1. - wt) + self.pm[:, k] * wt

    def _denprfl(self, rhaer):
        """sigma * n * dz per layer (tauaero.f:1361-1446)."""
        visfac = f32(3.912)
        nzb, ndb = numset(ZIP, self.zbaer), numset(ZIP, self.dbaer)
        if ndb > 0:
            if nzb == 1:
                raise ValueError("Error -- only one value of zbaer set")
            if nzb != ndb and nzb > 1:
                raise ValueError("Error -- number of elements must match: zbaer, dbaer")
            if nzb == 0:
                nzb = ndb
                self.zbaer[:nzb] = self.z[:nzb]
            self.nzbaer = nzb
        else:
            self._aerzstd()
        if self.iaer == 5:
            ext55 = self._usraer()
            if self.vis == ZIP and self.tbaer == ZIP:
                self.tbaer = ext55
        else:
            self._stdaer(rhaer)
            ext55 = self.aerbwi(WL55)[0]
            if self.vis == ZIP and self.tbaer == ZIP:
                raise ValueError("must specify either tbaer or vis")
        dtsv = self._aervint()
        sigma = 0.
        if ext55 > 0.:
            if self.tbaer >= 0:
                vint = dtsv.sum()
                if vint != 0:
                    sigma = self.tbaer / (ext55 * vint)
            else:
                sigma = visfac / (ext55 * self.vis * self.aeroden(0.0))
        self.dtsv = sigma * dtsv

    # ---- per wavelength (tauaero.f:1223-1331)
    def __call__(self, wl, nmom, pmom):
        """Adds (moments x scattering depth) to pmom[nz][nmom+1] in place; returns dtaua, waer."""
        nz = self.nz
        dtauab, waer = np.zeros(nz), np.zeros(nz)
        if self.iaer > 0:
            extinc, wa, ga = self.aerbwi(wl)
            if self.nosct == 1:
                extinc *= 1. - wa
            if self.nosct == 3:
                extinc *= 1. - wa * ga
            if self.nosct != 0:
                wa = ga = 0.
            if self.imoma > 0:
                pm = getmom(self.imoma, ga, nmom)
                namom = nmom
            else:
                namom = self.npmaer
                pm = np.concatenate([[1.0], self._phaerw(wl)])
            dtauab = extinc * self.dtsv
            waer[:] = wa
            m = min(namom, nmom)
            pmom[:, 1:m + 1] += (pm[None, 1:m + 1] * dtauab[:, None]) * waer[:, None]
        elif self.iaer == -1:                 # aerosol.dat (tauaero.f:1250-1254)
            dtauab, waer = self.afile(wl, nmom, pmom)
        dtaua = dtauab.copy()
        for i in range(NAERZ):
            if self.jaer[i] != 0 and self.taerst[i] > 0.:
                nl = self.laer[i]
                if nl <= 0:
                    raise ValueError("stratospheric aerosol layer outside the grid")
                extinc, wa, ga = self.aestrat(int(self.jaer[i]), wl)
                dt = self.taerst[i] * extinc
                pm = getmom(3, ga, nmom)
                pmom[nl - 1, 1:] += pm[1:] * dt * wa
                waer[nl - 1] = (waer[nl - 1] * dtaua[nl - 1] + wa * dt) / (dtaua[nl - 1] + dt)
                dtaua[nl - 1] += dt
        return dtaua, waer
