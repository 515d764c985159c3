"""SBDART front end: NAMELIST INPUT -> per-bin DISORT inputs -> IOUT records.

Host-side mirror of the reference driver (drt.f) and of the optical-property
producers it calls once per wavelength (8f-1/2):
  atms / modatm            atms.f:223-466        standard atmospheres
  absint, volmix, modmix   taugas.f:1924, 7178   absorber path integrals
  taugas + continua        taugas.f:2236-6821    LOWTRAN7 band model
  kdistr, gasset, taucor   taugas.f:1802, 7392   3-term k-distribution
  depthscl                 taugas.f:7512         per-k total depth / SSA
  rayleigh, solirr, salbedo, setfilt, filter     spectra.f
  taucloud, cloudpar       taucloud.f:10, :344   Mie-table clouds
  wllimits, zlayer, nearest, normom, stdout0/1/2 drt.f
The physical tables are data extracted from the reference's DATA statements
by tools/extract_tables.py (tables.npz).  The radiative-transfer solve itself
is NOT here: `run()` takes a `solve(batch)` callable -- the CUDA batch solver
(Solver.disort_batch); the golden tests also pass their CPU checker.

Aerosols, zgrid, in-cloud humidity, zensun and the sensor filters live in extras.py.
User files read from the working directory like the reference: atms.dat, albedo.dat,
filter.dat, solar.dat, aerosol.dat, usrcld.dat, CKATM / CKTAU (kdist=-1).
"""
from __future__ import annotations

import math
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_T = None

# params.f:17-30 (default-REAL literals widened to double)
f32 = lambda x: float(np.float32(x))  # noqa: E731
PZERO, TZERO, RE = f32(1013.25), f32(273.15), f32(6371.2)
PMO, GRAV, ALOSCH = f32(2.6568e-23), f32(9.80665), f32(2.6868e19)
ZIP = -1.0
PI_KR = 3.1415926536
MXLY, NSTRMS, NCLDZ, MXQ, MAXMOM = 65, 40, 5, 63, 299


def tables():
    global _T
    if _T is None:
        z = np.load(os.path.join(_HERE, "tables.npz"))
        _T = {k: z[k] for k in z.files}
    return _T


def T(key):
    return tables()[key]


# ------------------------------------------------------------------ NAMELIST
# names of the two NAMELIST groups (drt.f:200-215); a name outside its group aborts the read
INPUT_NAMES = frozenset("""idatm amix isat wlinf wlsup wlinc sza csza solfac nf iday time alat alon zpres pbar
    sclh2o uw uo3 o3trp ztrp xrsc xn2 xo2 xco2 xch4 xn2o xco xno2 xso2 xnh3 xno xhno3 xo4 isalb albcon sc
    zcloud tcloud lwp nre rhcld krhclr jaer zaer taerst iaer vis rhaer tbaer wlbaer qbaer abaer wbaer gbaer
    pmaer zbaer dbaer nothrm nosct kdist zgrid1 zgrid2 ngrid idb zout iout prnt temis nstr nzen uzen vzen
    nphi phi saza imomc imoma ttemp btemp corint spowder""".split())
DINPUT_NAMES = frozenset("ibcnd phi0 prnt ipth fisot temis nstr nzen uzen vzen nphi phi ttemp btemp".split())


def parse_namelist(text, group="input", names=None):
    """Fortran NAMELIST reader for the subset SBDART's INPUT files use (drt.f:200-231):
    scalars, lists, repeat counts n*v, logicals t/f, case-insensitive names, array
    elements `name(k) = v1, v2, ...` (values fill elements k, k+1, ...), `/`, `&end` or
    `$end` terminators with or without leading blanks, `!` comments.  Returns a list of
    (name, first_element_or_None, values); None when the group is absent.  A name that
    is not in `names` raises ValueError, as the Fortran read aborts on it."""
    m = re.search(r"[&$]\s*" + group + r"\b(.*?)(?:/|&end|\$end)", text, re.I | re.S)
    if not m:
        return None
    body = re.sub(r"!.*", "", m.group(1))
    items = []
    lhs = r"([A-Za-z_]\w*)\s*(?:\(\s*(\d+)\s*\))?\s*="
    ahead = r"[A-Za-z_]\w*\s*(?:\(\s*\d+\s*\))?\s*="
    for name, sub, val in re.findall(lhs + r"\s*(.*?)(?=\s*" + ahead + r"|\Z)", body, re.S):
        name = name.lower()
        if names is not None and name not in names:
            raise ValueError(f"namelist ${group.upper()}: unknown variable '{name}'")
        toks = [t for t in re.split(r"[\s,]+", val.strip()) if t]
        vals = []
        for t in toks:
            rep = 1
            if "*" in t:
                r, t = t.split("*", 1)
                rep = int(r)
            tl = t.lower().strip(".")
            if tl in ("t", "true"):
                v = True
            elif tl in ("f", "false"):
                v = False
            else:
                try:
                    v = float(t.lower().replace("d", "e"))
                except ValueError:
                    raise ValueError(f"namelist ${group.upper()}: bad value '{t}' for '{name}'") from None
            vals += [v] * rep
        if not vals:
            raise ValueError(f"namelist ${group.upper()}: no value for '{name}'")
        items.append((name, int(sub) if sub else None, vals))
    return items


def apply_namelist(p, items):
    """Assign parsed NAMELIST items onto the parameter dict: whole-variable assignments
    replace the leading elements of an array default (the rest keeps its default, as in
    Fortran), `name(k) = ...` assigns elements k, k+1, ... (1-based)."""
    for name, sub, vals in items:
        cur = p.get(name)
        if isinstance(cur, (list, np.ndarray)) or sub is not None:
            base = list(cur) if isinstance(cur, (list, np.ndarray)) else ([] if cur is None else [cur])
            k0 = (sub or 1) - 1
            if len(base) < k0 + len(vals):
                if name in DEFAULTS and isinstance(DEFAULTS[name], list) and name != "pmaer":
                    raise ValueError(f"namelist: subscript of '{name}' out of range")
                base += [ZIP] * (k0 + len(vals) - len(base))
            base[k0:k0 + len(vals)] = vals
            p[name] = base
        else:
            p[name] = vals[0] if len(vals) == 1 else vals


DEFAULTS = dict(
    idatm=4, amix=ZIP, isat=0, wlinf=0.55, wlsup=0.55, wlinc=0.0, sza=0.0, csza=ZIP, solfac=1.0,
    nf=2, iday=0, time=16.0, alat=-64.767, alon=-64.067, zpres=ZIP, pbar=ZIP, sclh2o=ZIP, uw=ZIP,
    uo3=ZIP, o3trp=ZIP, ztrp=0.0, xrsc=1.0, xn2=ZIP, xo2=ZIP, xco2=ZIP, xch4=ZIP, xn2o=ZIP,
    xco=ZIP, xno2=ZIP, xso2=ZIP, xnh3=ZIP, xno=ZIP, xhno3=ZIP, xo4=1.0, isalb=0, albcon=0.0,
    zcloud=[0.0] * NCLDZ, tcloud=[0.0] * NCLDZ, lwp=[0.0] * NCLDZ, nre=[8.0] * NCLDZ, rhcld=ZIP,
    krhclr=0, iaer=0, nothrm=-1, nosct=0, kdist=3, zgrid1=1.0, zgrid2=30.0, ngrid=0,
    zout=[0.0, 100.0], iout=10, temis=0.0, nstr=0, nzen=0, uzen=[ZIP] * NSTRMS,
    vzen=[90.0] * NSTRMS, nphi=0, phi=[ZIP] * NSTRMS, saza=180.0, imomc=3, ttemp=ZIP, btemp=ZIP,
    corint=False, ibcnd=0, fisot=0.0,
    # module aeroblk (tauaero.f:27-47)
    zaer=[0.0] * 5, taerst=[0.0] * 5, jaer=[0] * 5, zbaer=[ZIP] * MXLY, dbaer=[ZIP] * MXLY, vis=ZIP,
    tbaer=ZIP, abaer=0.0, wlbaer=[ZIP] * 150, qbaer=[ZIP] * 150, wbaer=[ZIP] * 150, gbaer=[ZIP] * 150,
    pmaer=[ZIP], rhaer=ZIP, imoma=3,
)


def _arr(v, n, fill):
    a = np.full(n, fill, dtype=float)
    v = np.atleast_1d(np.asarray(v, dtype=float))
    a[: len(v)] = v
    return a


def numset(zipv, zz):
    idx = np.nonzero(np.asarray(zz) != zipv)[0]
    return int(idx[-1]) + 1 if len(idx) else 0


def locate(xx, x):
    """drt.f:1181-1235 (1-based result)."""
    n = len(xx)
    if x == xx[0]:
        return 1
    if x == xx[-1]:
        return n - 1
    jl, ju = 1, n
    asc = xx[-1] > xx[0]
    while ju - jl > 1:
        jm = (ju + jl) // 2
        if asc == (x > xx[jm - 1]):
            jl = jm
        else:
            ju = jm
    return jl


def interp_table(wlt, tab, wl):
    """locate + clipped linear weight (spectra.f:53-56, :1409-1412, :3406-3409)."""
    j = locate(wlt, wl)
    wt = (wl - wlt[j - 1]) / (wlt[j] - wlt[j - 1])
    wt = max(0.0, min(1.0, wt))
    return tab[j - 1] * (1.0 - wt) + tab[j] * wt


# ------------------------------------------------------------------ atmosphere
_ATM = {1: ("tropic", 1), 2: ("midsum", 2), 3: ("midwin", 3), 4: ("subsum", 4), 5: ("subwin", 5),
        6: ("us62", 6)}


def atms(idatm):
    name, k = _ATM[abs(idatm)]
    g = lambda v: T(f"atms/{name}/{v}{k}").astype(float).copy()  # noqa: E731
    return g("z"), g("p"), g("t"), g("wh"), g("wo")


def modatm(z, p, wh, wo, sclh2o, uw, uo3, o3trp, ztrp, pbar):
    """atms.f:223-420."""
    nz = len(z)
    if uw >= 0.0:
        if sclh2o > 0.0:
            wh[:] = uw / sclh2o * np.exp(-z / sclh2o)
        else:
            tot = 0.0
            for i in range(nz - 2, -1, -1):
                dz = z[i + 1] - z[i]
                d1, d2 = wh[i], wh[i + 1]
                if abs(d1 - d2) <= f32(.001) * d1 or min(d1, d2) == 0.0:
                    du = .5 * dz * (d1 + d2)
                else:
                    du = dz * (d1 - d2) / math.log(d1 / d2)
                tot += du
            wh *= uw / (f32(0.1) * tot)
    if uo3 >= 0.0 or o3trp >= 0.0:
        ofac = .5 / (30 * PMO * ALOSCH)
        tropo3 = tropo3p = strto3 = 0.0
        for i in range(nz - 2, -1, -1):
            if z[i] >= ztrp:
                strto3 += ofac * (z[i + 1] - z[i]) * (wo[i] + wo[i + 1])
            elif tropo3p == 0:
                tropo3p = ofac * (z[i + 1] - z[i]) * wo[i + 1]
                tropo3 = ofac * (z[i + 1] - z[i]) * wo[i]
            else:
                tropo3 += ofac * (z[i + 1] - z[i]) * (wo[i] + wo[i + 1])
        factrp = facstr = 1.0
        if uo3 >= 0.0:
            facstr = uo3 / strto3
        if o3trp >= 0:
            factrp = max(o3trp - tropo3p * facstr, 0.0) / tropo3
        wo[:] = np.where(z < ztrp, factrp * wo, facstr * wo)
    if pbar >= 0.0:
        p *= pbar / p[0]


class TraceGases:
    """trcblk + volmix + modmix (taugas.f:6823-7294)."""

    NAMES = ("n2", "o2", "co2", "ch4", "n2o", "co", "no2", "so2", "nh3", "no", "hno3")

    def __init__(self, x):
        self.alt = T("taugas/trcblk/alt")
        self.tab = {n: T(f"taugas/trcblk/{n}") for n in self.NAMES}
        self.u = {n: 1.0 for n in self.NAMES}
        self.modify = 0
        vals = []
        for n in self.NAMES:
            xv = x.get("x" + n, ZIP)
            vals.append(xv)
            if xv >= 0.0:
                self.u[n] = xv / self.tab[n][0]
        if max(vals) > f32(-0.99):
            self.modify = 1

    def volmix(self, zz):
        z = max(0.0, min(zz, 100.0))
        if z < 25.:
            k = int(z) + 1
        elif z < 50.:
            k = int(26 + (z - 25.) / 5.)
        elif z < 70.:
            k = 31
        else:
            k = 32
        f = (z - self.alt[k - 1]) / (self.alt[k] - self.alt[k - 1])
        out = {}
        for n in self.NAMES:
            v = self.tab[n][k - 1] * (1. - f) + self.tab[n][k] * f
            if self.modify:
                v *= self.u[n]
            out[n] = v
        return out


_PSS_TSS = {  # dd(k) = con * pss**a * tss**b  (taugas.f:2039-2120), k 1-based
    17: (0.9810, 0.3324), 18: (1.1406, -2.6343), 19: (0.9834, -2.5294), 20: (1.0443, -2.4359),
    21: (0.9681, -1.9537), 22: (0.9555, -1.5378), 23: (0.9362, -1.6338), 24: (0.9233, -0.9398),
    25: (0.8658, -0.1034), 26: (0.8874, -0.2576), 27: (0.7982, 0.0588), 28: (0.8088, 0.2816),
    29: (0.6642, 0.2764), 30: (0.6656, 0.5061),
    31: (0.4200, 1.3909), 32: (0.4221, 0.7678), 33: (0.3739, 0.1225), 34: (0.1770, 0.9827),
    35: (0.3921, 0.1942),
    36: (0.6705, -2.2560), 37: (0.7038, -5.0768), 38: (0.7258, -1.6740), 39: (0.6982, -1.8107),
    40: (0.8867, -0.5327), 41: (0.7883, -1.3244), 42: (0.6899, -0.8152), 43: (0.6035, 0.6026),
    44: (0.7589, 0.6911), 45: (0.9267, 0.1716), 46: (0.7139, -0.4185),
    47: (0.3783, 0.9399), 48: (0.7203, -0.1836), 49: (0.7764, 1.1931),
    50: (1.1879, 2.9738), 51: (0.9353, 0.1936), 52: (0.8023, -0.9111), 53: (0.6968, 0.3377),
    54: (0.5265, -0.4702), 55: (0.3956, -0.0545), 56: (0.2943, 1.2316), 57: (0.2135, 0.0733),
}
_GROUP = {  # species index -> which concentration it scales
    **{k: "h2o" for k in range(17, 31)}, **{k: "o3" for k in range(31, 36)},
    **{k: "co2" for k in range(36, 44)}, 44: "co", 45: "co", 46: "ch4", 47: "n2o", 48: "n2o",
    49: "n2o", 50: "o2", 51: "o2", 52: "nh3", 53: "nh3", 54: "no", 55: "no2", 56: "so2", 57: "so2",
}


def absint(z, p, t, wh, wo, trace):
    """Absorber amounts uu(63, nz) above each level (taugas.f:1924-2233)."""
    nz = len(z)
    dd = np.zeros((MXQ + 1, nz))          # 1-based species index
    xlosch = ALOSCH * f32(1.e5)
    conjoe = f32(0.1) / ALOSCH
    con = f32(3.3429e21)
    rhzero = TZERO / 296.0
    vf = None
    for i in range(nz):
        vf = trace.volmix(z[i])
        tt, pp = t[i], p[i]
        pss, tss = pp / PZERO, TZERO / tt
        f1 = (pp / PZERO) / (tt / TZERO)
        f2 = (pp / PZERO) * math.sqrt(TZERO / tt)
        wair = ALOSCH * f1
        rhoair = f1
        rhoh2o = con * wh[i] / xlosch
        rhofrn = rhoair - rhoh2o
        wo2d = conjoe * wair * vf["o2"] * pss
        vfo3 = wo[i] / (3 * PMO * wair)
        dd[1, i] = wo2d * tt
        dd[2, i] = wo2d * (tt - 220.) ** 2
        dd[3, i] = f1 ** 2
        dd[4, i] = f32(1.e-6) * vf["n2"] * f1 * f2
        dd[5, i] = xlosch * rhoh2o ** 2 / rhzero
        dd[6, i] = f1
        dd[8, i] = conjoe * wair * vfo3
        dd[10, i] = xlosch * rhoh2o * rhofrn / rhzero
        dd[11, i] = f1 * vf["hno3"] * (f32(1.e-6) * f32(1.e5))
        dd[63, i] = wo2d
        conc = {"h2o": wh[i] * f32(.1), "o3": conjoe * wair * vfo3}
        for g in ("co2", "co", "ch4", "n2o", "o2", "nh3", "no", "no2", "so2"):
            conc[g] = conjoe * wair * vf[g]
        for k, (a, b) in _PSS_TSS.items():
            dd[k, i] = conc[_GROUP[k]] * pss ** f32(a) * tss ** f32(b)
        dd[58, i] = (1. + f32(.83) * f1) * conc["o2"]
    uu = np.zeros((MXQ + 1, nz + 1))
    scfac = math.exp(-1.)
    du5 = du8 = 0.0
    for i in range(nz - 1, -1, -1):
        if i == nz - 1:
            ztop, ptop, ttop = 2 * z[i] - z[i - 1], p[i] ** 2 / p[i - 1], t[i]
        else:
            ztop, ptop, ttop = z[i + 1], p[i + 1], t[i + 1]
        dz = ztop - z[i]
        if p[i] == ptop:
            tbar = .5 * (ttop + t[i])
        else:
            dp = (p[i] - ptop) / math.log(p[i] / ptop)
            drho = (p[i] / t[i] - ptop / ttop) / math.log(p[i] * ttop / (ptop * t[i]))
            tbar = dp / drho
        for k in range(1, MXQ + 1):
            den1 = dd[k, i]
            if i == nz - 1:
                uutop, den2 = 0.0, dd[k, i] * scfac
            else:
                uutop, den2 = uu[k, i + 1], dd[k, i + 1]
            ozn = k == 8 or 31 <= k <= 35 or 59 <= k <= 60
            denmin, denave = min(den1, den2), .5 * (den1 + den2)
            if denmin > 0. and denmin < f32(0.999) * denave and not ozn:
                du = dz * (den1 - den2) / math.log(den1 / den2)
            else:
                du = dz * denave
            if k == 5:
                du5 = du
            if k == 8:
                du8 = du
            if k == 9:
                tfac = max(0.0, min(1.0, (296. - tbar) / (296. - 260.)))
                uu[9, i] = uutop + du5 * tfac
            elif k == 59:
                uu[59, i] = uutop + f32(.269) * du8 * (tbar - f32(273.15))
            elif k == 60:
                uu[60, i] = uutop + f32(.269) * du8 * (tbar - f32(273.15)) ** 2
            else:
                uu[k, i] = uutop + du
    return uu           # uu[k, i], k 1-based species, i 0-based level; uu[:, nz] = 0


# ------------------------------------------------------------------ gas continua
def _sint(tab, v):
    """taugas.f:3854-3871 (v1=-20, dv=10, npts=2003)."""
    i = int((v + 20.) / 10. + f32(1.00001))
    if i >= 2003:
        return 0.0
    c = tab[i - 1]
    if int(v) % 10 > 0:
        c = (tab[i - 1] + tab[i]) / 2.
    return c


def _c4dta(v):
    if v < 2080. or v > 2740.:
        return 0.0
    return T("taugas/c4dta/c4")[(int(v) - 2080) // 5]


def _hno3(v):
    if 850. <= v <= 920.:
        return T("taugas/hno3/h1")[int((v - 845.) / 5.) - 1]
    if 1275. <= v <= 1350.:
        return T("taugas/hno3/h2")[int((v - 1270.) / 5.) - 1]
    if 1675. <= v <= 1735.:
        return T("taugas/hno3/h3")[int((v - 1670.) / 5.) - 1]
    return 0.0


def _hertda(v):
    if v <= 36000.:
        return 0.0
    corr = 0.0
    if v <= 40000.:
        corr = ((40000. - v) / 4000.) * f32(7.917e-27)
    rlosch = f32(2.6868e24) * f32(1.0e-5)
    yr = v / 48811.0
    return (f32(6.884e-24) * yr * math.exp(f32(-69.738) * math.log(yr) ** 2) - corr) * rlosch


def _o2cont(v):
    if v < 1395 or v > 1760:
        return 0.0, 0.0, 0.0
    i = int((v - 1395.0) / 5.0 + f32(1.00001))
    if i < 1 or i > 74:
        c = a = b = 0.0
    else:
        c, a, b = (T("taugas/o2cont/o2s0")[i - 1], T("taugas/o2cont/o2a")[i - 1],
                   T("taugas/o2cont/o2b")[i - 1])
    return c / f32(0.20946), a, a ** 2 / 2. + b


def _o4cont(wl, xo4):
    wnm = 1000. * wl
    inm = int(wnm)
    f = wnm - inm
    inm = inm - 335 + 1
    if 1 <= inm <= 1015:
        fraco2, fracn2, effn2 = f32(.209), f32(.781), f32(.2)
        factor = fraco2 ** 2
        if wl > f32(1.2):
            factor = fraco2 * (fraco2 + effn2 * fracn2)
        sig = T("taugas/o4cont/sig")
        return xo4 * factor * (sig[inm - 1] * (1. - f) + sig[inm] * f)
    return 0.0


def _o3hht(v):
    i = int((v - 27370.) / 5. + f32(1.00001))
    if i < 1 or i > 2687:
        return 0.0, 0.0, 0.0
    return (T("taugas/o3hht/s0")[i - 1], T("taugas/o3hht/s1")[i - 1], T("taugas/o3hht/s2")[i - 1])


def _o3uv(v):
    s = T("taugas/o3uv/s")
    i = int((v - 40800.) / 100. + f32(1.00001))
    if i < 1 or i > 133:
        return 0.0
    vr = i * 100. + 40800.
    if (v - f32(.1)) <= vr <= (v + f32(.1)):
        return s[i - 1]
    if i == 133:
        i = 132
    am = (s[i] - s[i - 1]) / 100.
    return am * v + (s[i - 1] - am * vr)


def _c8dta(v):
    if v < 13000. or v > 50000.:
        return 0.0
    iv = int(v)
    if 24200 < iv < 27500:
        return 0.0
    xi = (v - 13000.0) / 200.0 + 1.
    if iv >= 27500:
        xi = (v - 27500.0) / 500. + 57.
    n = int(xi + f32(1.001))
    xd = xi - float(n)
    c8 = T("taugas/c8dta/c8")
    return c8[n - 1] + xd * (c8[n - 1] - c8[n - 2])


def _cxdta(v, iwl, iwh, cp):
    """Band coefficient lookup (taugas.f:6416-6456), stateless form."""
    iv = int(v)
    nb = int(np.nonzero(iwl == -999)[0][0])
    ic = 0
    for b in range(nb):
        if iwl[b] <= iv <= iwh[b]:
            return cp[ic + (iv - iwl[b]) // 5]
        ic += (iwh[b] - iwl[b]) // 5 + 1
    return -20.0


_MOLS = ("h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "no", "so2", "no2", "nh3")   # imol 1..11

# abcdta band windows (taugas.f:6547-6731): imol -> list of (iw, [(lo, hi), ...])
_BANDS = {
    1: [(17, [(0, 345)]), (18, [(350, 1000)]), (19, [(1005, 1640)]), (20, [(1645, 2530)]),
        (21, [(2535, 3420)]), (22, [(3425, 4310)]), (23, [(4315, 6150)]), (24, [(6155, 8000)]),
        (25, [(8005, 9615)]), (26, [(9620, 11540)]), (27, [(11545, 13070)]),
        (28, [(13075, 14860)]), (29, [(14865, 16045)]), (30, [(16340, 17860)])],
    3: [(31, [(0, 200)]), (32, [(515, 1275)]), (33, [(1630, 2295)]), (34, [(2670, 2845)]),
        (35, [(2850, 3260)])],
    2: [(36, [(425, 835)]), (37, [(840, 1440)]), (38, [(1805, 2855)]), (39, [(3070, 3755)]),
        (40, [(3760, 4065)]), (41, [(4530, 5380)]), (42, [(5905, 7025)]),
        (43, [(7395, 7785), (8030, 8335), (9340, 9670)])],
    5: [(44, [(0, 175)]), (45, [(1940, 2285), (4040, 4370)])],
    6: [(46, [(1065, 1775), (2345, 3230), (4110, 4690), (5865, 6135)])],
    4: [(47, [(0, 120)]),
        (48, [(490, 775), (865, 995), (1065, 1385), (1545, 2040), (2090, 2655)]),
        (49, [(2705, 2865), (3245, 3925), (4260, 4470), (4540, 4785), (4910, 5165)])],
    7: [(50, [(0, 265)]),
        (51, [(7650, 8080), (9235, 9490), (12850, 13220), (14300, 14600), (15695, 15955),
              (49600, 52710)])],
    11: [(52, [(0, 385)]), (53, [(390, 2150)])],
    8: [(54, [(1700, 2005)])],
    10: [(55, [(580, 925), (1515, 1695), (2800, 2970)])],
    9: [(56, [(0, 185)]), (57, [(400, 650), (950, 1460), (2415, 2580)])],
}
_BASE = {1: 16, 3: 30, 2: 35, 5: 43, 6: 45, 4: 46, 7: 49, 11: 51, 8: 53, 10: 54, 9: 55}


def abcdta(iv):
    """Band-model parameters for wavenumber iv: ibnd, bms, bma, bmb, bmc per molecule."""
    ibnd = np.full(12, -1, dtype=int)
    bms, bma, bmb, bmc = (np.zeros(12) for _ in range(4))
    for imol, bands in _BANDS.items():
        iw = -1
        for w, rngs in bands:
            if any(lo <= iv <= hi for lo, hi in rngs):
                iw = w
        ibnd[imol] = iw
        if iw > 0:
            ib = iw - _BASE[imol] - 1
            m = _MOLS[imol - 1]
            bms[imol] = T(f"taugas/abcdta/a{m}")[ib]
            if imol == 7 and 49600 <= iv <= 52710:
                bms[imol] = f32(.4704)
            bma[imol] = T(f"taugas/abcdta/aa{m}")[ib]
            bmb[imol] = T(f"taugas/abcdta/bb{m}")[ib]
            bmc[imol] = T(f"taugas/abcdta/cc{m}")[ib]
    return ibnd, bms, bma, bmb, bmc


def _schrun(v):
    i = int((v - 49600.) / 5. + f32(1.0001))
    if 1 <= i <= 423:
        return T("taugas/schrun/shn")[i - 1]
    return -20.0


def raysig(v):
    return v ** 4 / (f32(9.38076e+18) + f32(-1.08426e+09) * v ** 2)


class GasState:
    """Per-wavelength module state of gasblk (cps, ibnd, bm*)."""


def taugas(wl, uu, amu0, z, xo4):
    """Continuum and band-model optical depth increments, top-down
    (taugas.f:2236-2534).  Returns dtauc[nz], dtaul[nz] and the band state."""
    nz = len(z)
    iv = 5 * (int(10000.0 / wl) // 5)
    v = 10000. / wl
    s0 = _sint(T("taugas/slf296/s"), v)
    s1 = _sint(T("taugas/slf260/s"), v)
    fh2o = _sint(T("taugas/frn296/f"), v)
    t0, t1 = 296., 260.
    if s0 > 0.:
        alpha2 = 200. ** 2
        xh2o = 1. - f32(0.2333) * (alpha2 / ((v - 1050.) ** 2 + alpha2))
        s0 *= xh2o
        s1 *= xh2o
    if (v / f32(0.6952)) / t1 <= 87.:
        xd = math.exp(-v / (t0 * f32(0.6952)))
        radfn0 = v * (1. - xd) / (1. + xd)
        xd = math.exp(-v / (t1 * f32(0.6952)))
        radfn1 = v * (1. - xd) / (1. + xd)
    else:
        radfn0 = radfn1 = v
    wfac = f32(1.e-20)
    ya = math.exp(-math.log(f32(1.025) * f32(3.159e-8)) + f32(2.75e-4) * v)
    yb = math.exp(-math.log(f32(8.97e-6)) + f32(1.300e-3) * v)
    fdg = 1. / (ya + yb)
    abn2 = _c4dta(v)
    abno3 = _hno3(v)
    abo2 = _hertda(v)
    sigo20, sigo2a, sigo2b = _o2cont(v)
    sigo4 = _o4cont(wl, xo4)
    doz1 = doz2 = doz3 = 0.
    if v > 40800:
        doz1 = f32(.269) * _o3uv(v)
    elif v > 24370:
        c0, ct1, ct2 = _o3hht(v)
        doz1, doz2, doz3 = f32(.269) * c0, c0 * ct1, c0 * ct2
    elif 13000. <= v <= 24200:
        doz1 = _c8dta(v)
    cps = np.full(12, -20.0)
    for imol, m in enumerate(_MOLS, start=1):
        cps[imol] = _cxdta(v, T(f"taugas/gasblk/iwl{m}"), T(f"taugas/gasblk/iwh{m}"),
                           T(f"taugas/gasblk/cp{m}"))
    ibnd, bms, bma, bmb, bmc = abcdta(iv)
    if v > 49600:
        cps[7] = _schrun(v)

    # slant-path weighted absorber amounts, accumulated from the top (taugas.f:2425-2437)
    def amuz(zzz):
        return np.sqrt(1. - (1. - amu0 ** 2) * (RE / (RE + zzz)) ** 2)

    duu = uu[:, :nz] - uu[:, 1:nz + 1]                 # uu[:, nz] = 0
    zbar = np.empty(nz)
    zbar[nz - 1] = z[nz - 1]
    zbar[:nz - 1] = 0.5 * (z[:nz - 1] + z[1:])
    inc = duu / amuz(zbar)[None, :]
    w = np.cumsum(inc[:, ::-1], axis=1)                # w[k, im-1], im = 1..nz top-down
    tcunif = sigo4 * w[3] + abn2 * w[4] + sigo20 * (w[63] + sigo2a * (w[1] - 220 * w[63]) +
                                                    sigo2b * w[2]) + abo2 * w[58]
    tch2o = (s0 * radfn0 * (wfac * w[5]) + ((s1 * radfn1) - (s0 * radfn0)) * (wfac * w[9]) +
             (fh2o + fdg) * radfn0 * (wfac * w[10]))
    tco3 = doz1 * w[8] + doz2 * w[59] + doz3 * w[60]
    tctrc = abno3 * w[11]
    tauc = tcunif + tch2o + tco3 + tctrc
    taul = np.zeros(nz)
    for k in range(1, 12):
        ib = ibnd[k]
        if ib > 0:
            cp = cps[k]
            if cp > -20.:
                wk = w[ib]
                ok = wk > 1.e-20
                awl = bms[k] * (cp + np.log10(np.where(ok, wk, 1.0)))
                awl = np.minimum(awl, 20.)
                taul += np.where(ok, 10. ** awl, 0.0)
    dtauc = np.diff(tauc, prepend=0.0)
    dtaul = np.diff(taul, prepend=0.0)
    st = GasState()
    st.cps, st.ibnd, st.bms, st.bma, st.bmb, st.bmc = cps, ibnd, bms, bma, bmb, bmc
    return dtauc, dtaul, st


def kdistr(uu, st, nz):
    """Three-term k-distribution of the band-model transmission (taugas.f:1802-1920)."""
    fac = T("taugas/kdistr/fac")
    cp1s = 10. ** st.cps
    dtauk = np.zeros((nz, 3))
    twgp = np.zeros((nz, 3))
    for k in range(3):
        for mol in range(1, 12):
            ib = st.ibnd[mol]
            if ib < 0:
                continue
            gk = fac[k] * st.bmc[mol]
            dp = (st.bma[mol], st.bmb[mol], 1. - st.bma[mol] - st.bmb[mol])[k]
            duu = (uu[ib, :nz] - uu[ib, 1:nz + 1])[::-1]            # n = 1..nz top-down
            wpth = duu * gk
            dtauk[:, k] += wpth * cp1s[mol]
            twgp[:, k] += wpth * cp1s[mol] * dp
    wtk = np.where(dtauk != 0, twgp / np.where(dtauk != 0, dtauk, 1.0), 1. / 3.)
    wtk = wtk / wtk.sum(axis=1, keepdims=True)
    tk = np.cumsum(dtauk, axis=0)
    return dtauk, tk, wtk


def taucor(gwk, tau, amu, utau):
    cf = 1.
    if utau > 12.0:
        return cf
    for _ in range(20):
        e = gwk * np.exp(-cf * tau / amu)
        ff = e.sum()
        f = math.log(ff) + utau
        if abs(f) < f32(0.000001):
            return cf
        fp = -(e * tau).sum() / (ff * amu)
        cf += -f / fp
    raise RuntimeError("TAUCOR: iteration did not converge")


def gasset(kdist, wl, uu, amu0, z, xo4):
    """taugas.f:7392-7510.  Returns nk, gwk[3], dtauk[nz][6], dtaugc[nz]."""
    nz = len(z)
    dtcv, dtlv, st = taugas(wl, uu, 1.0, z, xo4)
    if amu0 > 0.:
        dtcs, dtls, _ = taugas(wl, uu, amu0, z, xo4)
    else:
        dtcs, dtls = dtcv, dtlv
    dtaugc = dtcv.copy()
    gwk = np.zeros(3)
    dtk = wtk = None
    if kdist == 0 or dtlv.sum() < f32(.01):
        nk = 1
        gwk[0] = 1.
    else:
        dtk, tk, wtk = kdistr(uu, st, nz)
        if tk[nz - 1].max() < f32(0.01):
            nk = 1
            gwk[0] = 1.
        else:
            nk = 3
            gwk = (dtlv[:, None] * wtk).sum(axis=0)
            wnorm = gwk.sum()
            gwk = np.array([1., 0., 0.]) if wnorm == 0 else gwk / wnorm
    dtauk = np.zeros((nz, 6))
    if kdist == 0 or nk == 1:
        dtauk[:, 0] = dtlv
        dtauk[:, 3] = amu0 * dtls
    else:
        dtauk[:, 0:3] = dtk
        dtauk[:, 3:6] = dtk
        if kdist >= 2 and amu0 > 0.:
            tauls = 0.
            tglc = np.zeros(3)
            for j in range(nz):
                tauls += dtls[j]
                tglc = dtk[j] + tglc
                cf = taucor(gwk, tglc, amu0, tauls)
                dtauk[j, 3:6] = tglc * (cf - 1.0) + dtk[j]
                tglc = cf * tglc
    if amu0 <= 0.:
        dtauk[:, 3] = dtlv
    return nk, gwk, dtauk, dtaugc


def _rolloff(wl, tsc):
    ramp = (f32(4.1) - wl) / (f32(4.1) - f32(3.9))
    ramp = max(min(1.0, ramp), 0.0)
    return ramp * math.exp(1. - max(tsc, 1.0))


def depthscl(kdist, kd, nk, wl, dtaur, dtaua, waer, dtauc, wcld, gwk, dtauk, dtaugc, spowder=False):
    """taugas.f:7512-7621 (kd 0-based).  Returns dtau, wreal, wt."""
    nz = len(dtaur)
    wt = gwk[kd]
    dtaug = np.zeros(nz)
    if kdist == -1:                       # optical depths from CKTAU
        dtaug = dtauk[:, kd].copy()
    elif kdist == 0 or nk == 1:
        wt = 1.
        tsc = tglv = tgls = 0.
        for i in range(nz):
            tglv += dtauk[i, 0]
            tgls += dtauk[i, 3]
            tsc += dtaur[i] + dtauc[i] + dtaua[i]
            afac = 1.
            if tglv > f32(.001):
                afac = tgls / tglv
            ramp = _rolloff(wl, tsc)
            afac = afac * ramp + 1. - ramp
            dtaug[i] = dtaugc[i] + dtauk[i, 0] * afac
    elif kdist == 1:
        dtaug = dtaugc + dtauk[:, kd]
    elif kdist == 2:
        dtaug = dtaugc + dtauk[:, kd + 3]
    else:
        tsc = 0.
        for i in range(nz):
            tsc += dtaur[i] + dtauc[i] + dtaua[i]
            ramp = _rolloff(wl, tsc)
            dtaug[i] = dtaugc[i] + dtauk[i, kd] * (1. - ramp) + dtauk[i, kd + 3] * ramp
    if spowder:                           # no gas, no Rayleigh scattering in the sub-surface layer
        dtaur[nz - 1] = 0.                # (in place, like the reference: later k-terms see it too)
        dtaug[nz - 1] = 0.
    dtau = dtaug + dtauc + dtaua + dtaur
    tiny = np.finfo(float).tiny
    sca = dtauc * wcld + dtaua * waer + dtaur
    wreal = np.where(dtau > tiny, sca / np.where(dtau > tiny, dtau, 1.0), 0.0)
    return dtau, wreal, wt


def rayleigh(wl, z, p, t):
    """spectra.f:206-247 (top-down)."""
    nz = len(z)
    sig = raysig(10000. / wl)
    d = np.zeros(nz)
    d[0] = sig * (p[nz - 1] / PZERO) / (t[nz - 1] / TZERO) * 5.
    for i in range(2, nz + 1):
        im = nz - i + 1
        rhom = (p[im - 1] / PZERO) / (t[im - 1] / TZERO)
        rhop = (p[im] / PZERO) / (t[im] / TZERO)
        dz = z[im] - z[im - 1]
        if rhom == rhop:
            d[i - 1] = .5 * sig * dz * (rhom + rhop)
        else:
            d[i - 1] = sig * dz * (rhop - rhom) / math.log(rhop / rhom)
    return d


# ------------------------------------------------------------------ spectra
def _f32ramp(n):
    i = np.arange(n)
    return (i.astype(np.float32) / np.float32(n - 1)).astype(float)


class Sun:
    """solirr (spectra.f:1367-1415) with the nf=1/2/3 tables."""

    def __init__(self, nf):
        self.nf = nf
        if nf == 2:     # sunlow, spectra.f:2394-2410
            s2a, s2b = T("spectra/sunlow/sun2a"), T("spectra/sunlow/sun2b")
            n1, n2 = 2910, 1440
            r1 = ((n1 - 1 - np.arange(n1)).astype(np.float32) / np.float32(n1 - 1)).astype(float)
            r2 = ((n2 - 1 - np.arange(n2)).astype(np.float32) / np.float32(n2 - 1)).astype(float)
            wn1 = 28400. + (57490. - 28400.) * r1
            wn2 = 0. + (28780. - 0.) * r2
            self.wl = np.concatenate([10000. / wn1, 10000. / np.maximum(wn2, 1.0)])
            self.s = np.concatenate([s2b[::-1], s2a[::-1]])
        elif nf == 1:   # sun1s: 0.25..4.0 um, 751 points
            self.s = T("spectra/sun1s/sun1")
            self.wl = f32(.25) + (4.0 - f32(.25)) * _f32ramp(751)
        elif nf == 3:   # sunmod, spectra.f:2884-2893: 100-49960 cm-1 every 20 cm-1
            s3 = T("spectra/sunmod/sun3")
            n = 2494
            r = ((n - 1 - np.arange(n)).astype(np.float32) / np.float32(n - 1)).astype(float)
            self.wl = 10000. / (100. + (49960. - 100.) * r)
            self.s = s3[::-1]
        elif nf == -1:  # solar.dat (wavelength um, W/m2/um)
            from .extras import rdspec
            self.wl, self.s = rdspec("solar.dat")
        elif nf != 0:
            raise ValueError(f"ERROR in solirr --- illegal value of nf {nf}")

    def __call__(self, wl):
        if self.nf == 0:
            return 1.0
        return interp_table(self.wl, self.s, wl)


class Albedo:
    """suralb + salbedo (spectra.f:28-176)."""

    NAMES = {1: "snow", 2: "clearw", 3: "lakew", 4: "seaw", 5: "sand", 6: "vegeta"}

    def __init__(self, isalb, albcon, sc):
        if isalb == 0:
            self.wl = np.array([0.0, np.finfo(float).max])
            self.alb = np.array([albcon, albcon])
        elif isalb in self.NAMES:
            self.alb = T(f"spectra/{self.NAMES[isalb]}/albx")
            self.wl = f32(.25) + (4.0 - f32(.25)) * _f32ramp(751)
        elif isalb == 10:
            self.wl = f32(.25) + (4.0 - f32(.25)) * _f32ramp(751)
            self.alb = sum(T(f"spectra/{n}/albx") * sc[i]
                           for i, n in enumerate(("snow", "seaw", "sand", "vegeta")))
        elif isalb == -1:   # albedo.dat
            from .extras import rdspec
            self.wl, self.alb = rdspec("albedo.dat")
        else:
            raise NotImplementedError(f"isalb={isalb}")

    def __call__(self, wl, warn=None):
        # spectra.f:44-56: outside the table a warning (errmsg 18) and the end value
        if warn is not None:
            if wl < self.wl[0]:
                warn(18, "SALBEDO--spectral range error, wlinf lt " + f"{self.wl[0]:9.3f}")
            if wl > self.wl[-1]:
                warn(18, "SALBEDO--spectral range error, wlsup gt " + f"{self.wl[-1]:9.3f}")
        return interp_table(self.wl, self.alb, wl)


def setfilt(isat, wlinf, wlsup, wlinc, want_filter=False):
    """spectra.f:3240-3387.  Returns wlmin, wlmax, nwl, wlinc [, filter function]."""
    if want_filter:
        from .extras import Filter, filter_table
        if isat in (0, -2):
            return setfilt(isat, wlinf, wlsup, wlinc) + (Filter(),)
        wlmin, wlmax, filt = filter_table(isat, wlinf, wlsup)
        return _setfilt_grid(wlmin, wlmax, wlinc) + (filt,)
    if isat == 0:
        wlmin, wlmax = wlinf, wlsup
        if wlinf == wlsup:
            return wlmin, wlmax, 1, f32(.001)
    elif isat == -2:
        wlmin, wlmax = wlinf - .5 * wlsup, wlinf + .5 * wlsup
        if wlsup == 0.:
            return wlmin, wlmax, 1, f32(.001)
    else:
        from .extras import filter_table
        wlmin, wlmax, _ = filter_table(isat, wlinf, wlsup)
    return _setfilt_grid(wlmin, wlmax, wlinc)


def _setfilt_grid(wlmin, wlmax, wlinc):
    """Number of wavelength steps (spectra.f:3358-3385)."""
    if wlmin < f32(0.199):
        raise ValueError("Error in SETFILT -- illegal wavelength limits")
    if wlinc > 1.:
        nwl = int(((10000. / wlmin) - (10000. / wlmax)) / wlinc + 1.)
    elif wlinc < 0.:
        nwl = int(1 + math.log(wlmax / wlmin) / abs(wlinc))
    else:
        if wlinc == 0:
            wlinc = (wlmax - wlmin) / max(10, 1 + int((wlmax - wlmin) / f32(0.005)))
        nwl = int(round((wlmax - wlmin) / wlinc)) + 1       # nint
    if wlmin != wlmax and nwl == 1:
        nwl = 2
    return wlmin, wlmax, nwl, wlinc


def wllimits(il, nwl, wlinc, wl1, wl2):
    """drt.f:1657-1740 for step il (0-based).  Returns wl, wvnmhi, wvnmlo."""
    wi = float(il)
    if wlinc > 1:
        f = lambda x: wl1 * wl2 / ((1. - x / (nwl - 1)) * wl2 + x / (nwl - 1) * wl1)  # noqa: E731
        wl, ww1, ww2 = f(wi), f(wi - .5), f(wi + .5)
    elif wlinc < 0.:
        wr = wl2 / wl1
        f = lambda x: wl1 * wr ** (x / (nwl - 1))  # noqa: E731
        wl, ww1, ww2 = f(wi), f(wi - .5), f(wi + .5)
    else:
        wl = wl1 + wi * wlinc
        ww1, ww2 = wl - .5 * wlinc, wl + .5 * wlinc
    if il == 0 and il != nwl - 1:
        ww1 = wl
    if il == nwl - 1 and il != 0:
        ww2 = wl
    if ww1 == wl and ww2 == wl:
        ww1, ww2 = wl - f32(.0005), wl + f32(.0005)
    return wl, 10000. / ww1, 10000. / ww2


# ------------------------------------------------------------------ clouds
def zlayer(z, zz):
    """drt.f:1268-1327: layer index (top-down, 1-based) of each altitude."""
    nz = len(z)
    lz = []
    for kk, zk in enumerate(zz):
        isgn = 1
        if kk == 0:
            zcmpr = zk + f32(.001)
        else:
            if zk < 0.:
                isgn = -1
            zcmpr = abs(zk + f32(.001))
        j = None
        for jj in range(nz, 0, -1):
            if z[jj - 1] <= zcmpr:
                j = jj
                break
        if j is None:
            lz.append(0)
        else:
            lz.append(isgn * (nz - j + 1))
    return lz


def levrng(lz, n):
    nzz = len(lz)
    if n > nzz or lz[n - 1] <= 0:
        return 0, 0
    lbot = ltop = lz[n - 1]
    if n != nzz and lz[n] < 0:
        ltop = -lz[n]
    return lbot, ltop


def cloudpar(wl, re):
    """Bilinear Mie lookup in (log wl, log2 re) (taucloud.f:6726-6762)."""
    wmin, wmax = math.log(f32(0.29)), math.log(f32(333.33))
    wstep = (wmax - wmin) / (400 - 1)
    eps = f32(.000001)
    fw = 1 + (math.log(wl) - wmin) / wstep
    fw = min(max(fw, 1.0), 400.0 - eps)
    iw = int(fw)
    fw -= iw
    fr = 1. + (math.log(abs(re)) / math.log(2.) - 1.) * 2
    fr = min(max(fr, 1.0), 13.0 - eps)
    ir = int(fr)
    fr -= ir
    sfx = "i" if re < 0. else ""
    out = []
    for nm in ("qq", "ww", "gg"):
        a = T(f"taucloud/cloudpar/{nm}{sfx}")
        out.append(a[iw - 1, ir - 1] * (1. - fw) * (1. - fr) + a[iw, ir - 1] * fw * (1. - fr) +
                   a[iw - 1, ir] * (1. - fw) * fr + a[iw, ir] * fw * fr)
    return tuple(out)


def getmom(iphas, gg, nmom):
    """disutil.f:2104-2209."""
    pm = np.zeros(nmom + 1)
    pm[0] = 1.0
    k = np.arange(1, nmom + 1)
    if iphas == 2:
        pm[2] = f32(0.1)
    elif iphas == 3:
        pm[1:] = gg ** k
    elif iphas == 4:
        h = T("disutil/getmom/hazelm")
        m = min(82, nmom)
        pm[1:m + 1] = h[:m] / (2 * k[:m] + 1)
    elif iphas == 5:
        c = T("disutil/getmom/cldmom")
        m = min(298, nmom)
        pm[1:m + 1] = c[:m] / (2 * k[:m] + 1)
    return pm


class Clouds:
    """taucloud (taucloud.f:10-140) incl. the first-call q550 normalisation."""

    def __init__(self, z, zcloud, tcloud, lwp, nre, imomc):
        self.nz = len(z)
        self.tcloud, self.lwp, self.nre, self.imomc = tcloud, lwp, nre, imomc
        self.mcldz = max(numset(0.0, tcloud), numset(0.0, lwp))
        self.lcld = zlayer(z, zcloud[: self.mcldz]) + [0] * (NCLDZ - self.mcldz)
        self.q550 = {}

    def __call__(self, wl, nmom):
        nz = self.nz
        taucld, wcld, qcld = np.zeros(nz), np.zeros(nz), np.zeros(nz)
        icnt = np.zeros(nz, dtype=int)
        pmom = np.zeros((nz, nmom + 1))
        tcl, lw, nre = self.tcloud, self.lwp, self.nre
        for i in range(1, NCLDZ + 1):
            lbot, ltop = levrng(self.lcld, i)
            if lbot == 0:
                continue
            if tcl[i - 1] == 0. and lw[i - 1] == 0.:
                continue
            for j in range(ltop, lbot + 1):
                if ltop == lbot:
                    reff, tcld, lwpth = nre[i - 1], tcl[i - 1], lw[i - 1]
                else:
                    wt = float(np.float32(j - ltop) / np.float32(lbot - ltop))
                    reff = nre[i] * (nre[i - 1] / nre[i]) ** wt
                    if tcl[i] == 0.:
                        tcld = tcl[i - 1] / (lbot - ltop + 1)
                    else:
                        tcld = 2 * tcl[i - 1] / ((lbot - ltop + 1) * (1. + tcl[i]))
                        tcld = tcld + (lbot - j) * tcld * (tcl[i] - 1.) / (lbot - ltop)
                    if lw[i] == 0.:
                        lwpth = lw[i - 1] / (lbot - ltop + 1)
                    else:
                        lwpth = 2 * lw[i - 1] / ((lbot - ltop + 1) * (1. + lw[i]))
                        lwpth = lwpth + (lbot - j) * lwpth * (lw[i] - 1.) / (lbot - ltop)
                qc, wc, gc = cloudpar(wl, reff)
                pm = getmom(self.imomc, gc, nmom)
                pmom[j - 1, 1:] += pm[1:]
                qcld[j - 1] += qc
                wcld[j - 1] += wc
                icnt[j - 1] += 1
                if tcl[i - 1] != 0.:
                    if j not in self.q550:
                        self.q550[j] = cloudpar(f32(0.55), reff)[0]
                    taucld[j - 1] += tcld * qc / self.q550[j]
                elif lwpth != 0.:
                    if reff < 0.:
                        taucld[j - 1] += f32(-.75) * qc * lwpth / reff / f32(.917)
                    else:
                        taucld[j - 1] += f32(.75) * qc * lwpth / reff
        for j in range(nz):
            if icnt[j]:
                wcld[j] /= icnt[j]
                pmom[j, 1:] = taucld[j] * wcld[j] * pmom[j, 1:] / icnt[j]
        return taucld, wcld, pmom


# ------------------------------------------------------------------ Fortran formats
def _es(x, w, d):
    """Fortran ESw.d of a value already rounded through REAL(4) when the reference does so."""
    if not np.isfinite(x):
        # gfortran's edit of non-finite values: right-justified NaN / Infinity (Inf when the
        # field is narrower than 8), so a solver failure is visible in the records
        if np.isnan(x):
            t = "NaN"
        else:
            t = ("-" if x < 0 else "") + ("Infinity" if w >= 8 + (x < 0) else "Inf")
        return t.rjust(w) if len(t) <= w else "*" * w
    s = f"{x:.{d}E}"
    mant, ex = s.split("E")
    e = int(ex)
    if abs(e) > 99:
        s = f"{mant}{'+' if e >= 0 else '-'}{abs(e):03d}"
    else:
        s = f"{mant}E{'+' if e >= 0 else '-'}{abs(e):02d}"
    return s.rjust(w)


def _f(x, w, d):
    s = f"{x:.{d}f}"
    if s.startswith("0.") and len(s) > w:
        s = s[1:]
    if s.startswith("-0.") and len(s) > w:
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def _r4(x):
    return float(np.float32(x))


# ------------------------------------------------------------------ the driver
class SbdartStop(Exception):
    """The reference printed a diagnostic table and executed STOP (idatm < 0, ngrid < 0,
    iday < 0): `text` is what it wrote to stdout."""

    def __init__(self, text):
        super().__init__(text)
        self.text = text


class SbdartFatal(Exception):
    """errmsg(0, ...) of the reference (disutil.f:278-325): the run wrote SBDART_WARNING.00 and
    stopped.  `text` is the message."""

    def __init__(self, text):
        super().__init__(text)
        self.text = text


def warning_file_text(num, message, input_text):
    """Contents of SBDART_WARNING.nn (errmsg, disutil.f:297-321): the message, a rule of 70
    '#', then a copy of the INPUT file with trailing blanks removed."""
    head = ("ERROR  >>>>>>" if num == 0 else "WARNING >>>>>") + " " + message
    body = "".join(ln.rstrip() + "\n" for ln in input_text.splitlines())
    return head + "\n\n" + "#" * 70 + "\n\n" + body


def write_warning_files(warnings, input_text, directory="."):
    """One file per message number, as the reference leaves them in the working directory
    (a number is reported once per run: msgset, disutil.f:295-299)."""
    paths = []
    for num, message in warnings:
        path = os.path.join(directory, f"SBDART_WARNING.{num:02d}")
        with open(path, "w") as fh:
            fh.write(warning_file_text(num, message, input_text))
        paths.append(path)
    return paths


class Sbdart:
    """One SBDART run (program sbdart, drt.f:90-563)."""

    def _warn(self, num, message):
        """errmsg(num > 0): remembered once per number (disutil.f:295-299)."""
        if all(n != num for n, _ in self.warnings):
            self.warnings.append((num, message))

    def __init__(self, namelist_text=None, **overrides):
        p = dict(DEFAULTS)
        if namelist_text is not None:
            nl = parse_namelist(namelist_text, "input", INPUT_NAMES)
            if nl is None:
                raise ValueError("error: namelist block $INPUT not found")
            apply_namelist(p, nl)
            dn = parse_namelist(namelist_text, "dinput", DINPUT_NAMES)
            if dn:
                apply_namelist(p, dn)
        p.update(overrides)
        for k in ("zcloud", "tcloud", "lwp"):
            p[k] = _arr(p[k], NCLDZ, 0.0)
        p["nre"] = _arr(p["nre"], NCLDZ, 8.0)
        p["uzen"] = _arr(p["uzen"], NSTRMS, ZIP)
        p["vzen"] = _arr(p["vzen"], NSTRMS, 90.0)
        p["phi"] = _arr(p["phi"], NSTRMS, ZIP)
        p["zout"] = _arr(p["zout"], 2, 0.0)
        for k in ("idatm", "isat", "nf", "isalb", "iout", "nstr", "nzen", "nphi", "kdist", "nothrm",
                  "imomc", "iaer", "ngrid", "nosct", "iday"):
            p[k] = int(p[k])
        self.p = p
        self._setup()

    # ---- setup part of the main program (drt.f:233-421)
    def _setup(self):
        p = self.p
        from . import extras
        iout = p["iout"]
        self.radcalc = iout in (5, 6, 20, 21, 22, 23)
        self.onlyfl = not self.radcalc
        if p["nstr"] == 0:
            p["nstr"] = min(20, NSTRMS) if self.radcalc else 4
        if p["isalb"] not in (7, 8, 9, -7, -8, -9):
            sc = p.get("sc", None)
            self.sc = [1., 0., 0., 0., 0.] if sc is None else list(np.atleast_1d(sc)) + [0.] * 5
        if self.radcalc:
            self._vuangles()
        self._chkin()
        sza = p["sza"]
        dtor = PI_KR / 180.
        if p["iday"] != 0:                                  # drt.f:276-277
            sza, p["saza"], p["solfac"] = extras.zensun(abs(p["iday"]), p["time"], p["alat"], p["alon"])
        elif p["csza"] != ZIP:
            sza = math.acos(p["csza"]) / dtor
        if abs(sza - 90) < f32(.01):
            sza = 95.
        self.sza = sza
        self.phi0 = math.fmod(p["saza"] - 180.0 + 360.0, 360.0)
        if p["iday"] < 0:                                   # drt.f:285-299: print the geometry and stop
            out = ["  day     time      lat      lon      sza      azm   solfac",
                   f"{abs(p['iday']):5d}" + "".join(f"{v:9.3f}" for v in (p["time"], p["alat"], p["alon"], sza,
                                                                          p["saza"], p["solfac"]))]
            if self.radcalc:
                out.append("      phi   rel_az")
                for ph in self.phi:
                    rel = ph + self.phi0 - (360. if self.phi0 > 180. and self.phi[-1] + self.phi0 > 360 else 0.)
                    out.append(f"{ph:9.3f}{rel:9.3f}")
            raise SbdartStop("\n".join(out) + "\n")
        self.wl1, self.wl2, self.nwl, self.wlinc, self.filter = setfilt(
            p["isat"], p["wlinf"], p["wlsup"], p["wlinc"], want_filter=True)
        kdist = 0 if iout == 2 else p["kdist"]
        self.kdist = kdist
        if kdist == -1:
            self._setup_ck()
            return
        if p["idatm"] == 0:                                 # atms.f:443
            z, pr, t, wh, wo = extras.useratm()
        else:
            z, pr, t, wh, wo = atms(p["idatm"])
        if p["amix"] > -1.:                                 # atms.f:451-463
            zz, pp, tt, hh, oo = extras.useratm()
            if len(zz) != len(z) or (np.abs(zz - z) > 0.01).any():
                raise ValueError("atms -- vertical grids do not match")
            am = p["amix"]
            pr, t = pr * (1. - am) + pp * am, t * (1. - am) + tt * am
            wh, wo = wh * (1. - am) + hh * am, wo * (1. - am) + oo * am
        if p["ngrid"] != 0:                                 # drt.f:307
            z, pr, t, wh, wo = extras.zgrid(z, pr, t, wh, wo, p["zgrid1"], p["zgrid2"], p["ngrid"])
        if p["zpres"] != ZIP:
            j = locate(z, p["zpres"])
            fj = (p["zpres"] - z[j - 1]) / (z[j] - z[j - 1])
            p["pbar"] = pr[j - 1] * (pr[j] / pr[j - 1]) ** fj
        modatm(z, pr, wh, wo, p["sclh2o"], p["uw"], p["uo3"], p["o3trp"], p["ztrp"], p["pbar"])
        self.trace = TraceGases(p)
        rhaer = p["rhaer"]
        if rhaer < 0.:                                      # drt.f:319
            rhaer = extras.relhum(t[0], wh[0])
        self.rhaer = rhaer
        self._setup_levels(z, pr, t, wh, wo)

    def _setup_ck(self):
        """kdist = -1 (drt.f:321-323, gasinit taugas.f:7297-7389): the atmosphere comes from the
        file CKATM, the gas optical depths of every spectral interval from CKTAU."""
        from . import extras
        p = self.p
        z, pr, t, h2oden = extras.ckatm()
        rhaer = p["rhaer"]
        if rhaer < 0.:
            rhaer = extras.relhum(t[0], h2oden)
        self.rhaer = rhaer
        self.cktau = extras.cktau_records(len(z))
        if self.wl1 == self.wl2:
            self.vnulo, self.vnuhi = 0., float(np.finfo(np.float32).max)
        else:
            self.vnuhi, self.vnulo = 10000. / self.wl1, 10000. / self.wl2
        nvnu, ended = 0, False
        for r in self.cktau:
            if r["vnu0"] > self.vnuhi:
                continue
            if r["vnu0"] < self.vnulo:
                ended = True
                break
            if r["ib"] == 1:
                nvnu += 1
        if ended and nvnu == 0:
            raise ValueError(f"Error --- gasinit\n no frequency samples within {self.wl1} {self.wl2}")
        self.nwl = nvnu
        self.trace = None
        self._setup_levels(z, pr, t, np.zeros(len(z)), np.zeros(len(z)))

    def _setup_levels(self, z, pr, t, wh, wo):
        p = self.p
        from . import extras
        dtor = PI_KR / 180.
        sza = self.sza
        self.z, self.pr, self.t, self.wh, self.wo = z, pr, t, wh, wo
        nz = len(z)
        self.nz = nz
        self.temper = np.concatenate([[t[nz - 1]], t[::-1]])          # drt.f:330-333
        self.btemp = self.temper[nz] if p["btemp"] < 0. else p["btemp"]
        self.ttemp = self.temper[0] if p["ttemp"] < 0. else p["ttemp"]
        self.spowder = bool(p.get("spowder"))
        if self.spowder:
            # a sub-surface layer between -1 and 0 km (drt.f:337-349).  The reference extends z, p, t,
            # wh, wo AFTER it has filled temper(0:nz): DISORT then reads temper(nz+1), an element of
            # the static array that was never set (0 K), while t(1) = btemp only enters the gas and
            # Rayleigh amounts.  Kept as built.
            if nz >= MXLY:
                raise ValueError("Error --- nz < mxly is required with spowder option")
            z = np.concatenate([[-1.0], z]); pr = np.concatenate([[f32(1.1) * pr[0]], pr])
            t = np.concatenate([[self.btemp], t]); wh = np.concatenate([[0.0], wh]); wo = np.concatenate([[0.0], wo])
            self.z, self.pr, self.t, self.wh, self.wo = z, pr, t, wh, wo
            nz += 1
            self.nz = nz
            self.temper = np.concatenate([self.temper, [0.0]])
        self.nstrsv = p["nstr"]
        self.clouds = Clouds(z, p["zcloud"], p["tcloud"], p["lwp"], p["nre"], p["imomc"])
        if p["rhcld"] >= 0 and self.kdist >= 0:             # drt.f:357-364
            if int(p["krhclr"]) == 1:
                extras.satcloud(self.clouds.lcld, t, p["rhcld"], wh)
            else:
                extras.saturate(self.clouds.lcld, z, t, p["rhcld"], wh)
        if p["ngrid"] < 0 or p["idatm"] < 0:                # prnatm (drt.f:803-809)
            raise SbdartStop(f"{nz:12d}\n" + "".join(
                _f(z[i], 11, 3) + "".join(_es(v[i], 11, 3) for v in (pr, t, wh, wo)) + "\n" for i in range(nz)))
        self.uu = absint(z, pr, t, wh, wo, self.trace) if self.kdist >= 0 else None
        zout = np.abs(p["zout"]) if p["zout"].min() < 0 else p["zout"]
        nbot = self._nearest(z, zout[0])
        ntop = self._nearest(z, zout[1])
        self.nbot = nz - nbot + 2
        self.ntop = nz - ntop + 2
        if self.ntop == 2:
            self.ntop = 1
        if self.radcalc:
            numu = self.nzen
            umu = np.zeros(numu)
            for j in range(1, numu + 1):
                u = min(1.0, max(math.cos(self.uzen[numu - j] * dtor), -1.0))
                if u == 0.:
                    u = f32(-.0001) if j == numu else f32(.0001)
                umu[j - 1] = u
            self.umu = umu
            if self.nphi == 0:
                self.nphi = 1
                self.phi = np.array([0.0])
        self.surface = None
        if p["isalb"] in (7, 8, 9):
            # BRDF surface: LAMBER = .FALSE., rsfc unused (drt.f:469-477)
            from . import brdf
            self.surface = brdf.SurfaceModel(p["isalb"], list(np.atleast_1d(p.get("sc") if p.get("sc") is not None else [0.] * 5)))
            self.albedo = lambda wl, warn=None: 0.0
        else:
            self.albedo = Albedo(p["isalb"], p["albcon"], self.sc)
        self.amu0 = math.cos(sza * dtor)
        self.sun = Sun(p["nf"])
        self.aerosols = extras.Aerosols(p, z, self.rhaer)

    def _chkin(self):
        """Range checks of the main program (chkin / ck, drt.f:568-735): the reference prints
        the offending parameters and stops; here the same text is the message of a ValueError."""
        p = self.p
        msgs = []

        def ck(name, rng, shown):
            if not msgs:
                msgs.append("CHKIN --- Errors detected in INPUT")
            msgs.append(f"\n     Input parameter {name} not within {rng}")
            msgs.append(f" {name}= {shown}")

        self.warnings = []
        if p["iaer"] == 0 and (p["vis"] != ZIP or p["tbaer"] != ZIP):
            self._warn(16, "CHKIN--IAER=0, though VIS or TBAER set")                   # drt.f:583
        if p["corint"] and self.onlyfl:
            self._warn(17, "CHKIN--CORINT=t, but flux output selected")                # drt.f:586
        if not -6 <= p["idatm"] <= 6:
            ck("idatm", "[-6,6]", p["idatm"])
        if p["wlinf"] < f32(0.199):
            ck("wlinf", "[0.2,-]", p["wlinf"])
        if p["isat"] <= -2:
            if p["wlsup"] < 0. or p["wlsup"] >= p["wlinf"]:
                ck("wlsup", "[0,wlinf]", p["wlsup"])
        elif p["wlsup"] < p["wlinf"] or p["wlsup"] > 100.:
            ck("wlsup", "[wlinf,100]", p["wlsup"])
        if not -4 <= p["isat"] <= 29:
            ck("isat", "[-4,29]", p["isat"])
        if p["solfac"] < 0.:
            ck("solfac", "[0,inf]", p["solfac"])
        if p["zcloud"].min() < -100. or p["zcloud"].max() > 100:
            ck("zcloud", "[-100,100]", p["zcloud"])
        if (np.abs(p["nre"]).min() < 2 or np.abs(p["nre"]).max() > 128) and p["nre"][0] != 0.:
            ck("nre", "[2,128]", p["nre"])
        if ((p["tcloud"] == 0) & (p["zcloud"] < 0)).any():
            msgs += ["CHKIN --- Error detected in input", "TCLOUD(k)=0 when ZCLOUD(k)<0"]
        if p["lwp"].min() < 0.:
            ck("lwp", "[0,inf]", p["lwp"])
        if np.abs(_arr(p["zaer"], 5, 0.0)).max() > 100.:
            ck("zaer", "[-100,100]", p["zaer"])
        if _arr(p["taerst"], 5, 0.0).min() < 0.:
            ck("taerst", "[0,inf]", p["taerst"])
        ja = _arr(p["jaer"], 5, 0.0)
        if ja.min() < 0 or ja.max() > 4:
            ck("jaer", "[0,4]", p["jaer"])
        if not -2 <= p["nf"] <= 3:
            ck("nf", "[-2,3]", p["nf"])
        if not -1 <= p["iaer"] <= 5:
            ck("iaer", "[-1,5]", p["iaer"])
        if p["isalb"] not in (-7, -8, -9, -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
            ck("isalb", "[-7,-8,-9,-1,0,1,2,3,4,5,6,7,8,9,10]", p["isalb"])
        if p["isalb"] == 0 and p["albcon"] < 0.:
            ck("albcon", "[0,inf]", p["albcon"])
        if p["zout"].min() < 0. or p["zout"].max() > 100:
            ck("zout", "[0,100]", p["zout"])
        if p["iout"] not in (1, 2, 5, 6, 7, 10, 11, 20, 21, 22, 23):
            ck("iout", "[1,2,5,6,7,10,11,20,21,22,23]", p["iout"])
        if not 0 <= p["nphi"] <= NSTRMS:
            ck("nphi", "[0,nstrms]", p["nphi"])
        if p["iout"] in (20, 21, 22, 23) and self.nzen == 0:
            msgs.append(f" iout ={p['iout']:2d} implies radiance calculation, but nzen=0 produces no radiance output")
        if p["zpres"] != ZIP and p["pbar"] != ZIP:
            msgs.append(" set zpres or pbar but not both")
        if numset(0.0, p["tcloud"]) != 0 and numset(0.0, p["lwp"]) != 0:
            msgs.append(" set TCLOUD or LWP, but not both")
        if msgs:
            raise ValueError("\n".join(msgs))
        # options of the reference this front end does not implement: fail loudly instead of
        # running something else (module docstring, "not covered")
        todo = []
        if p["kdist"] < -1:
            todo.append(f"kdist={p['kdist']}")
        if p["isalb"] in (-7, -8, -9):
            todo.append(f"isalb={p['isalb']} (dref, not in the reference source either)")
        if int(p.get("ibcnd", 0)) != 0:
            todo.append("ibcnd=1 (ALBTRN, disort.f:6718)")
        todo += self._unsupported_surface()
        if todo:
            raise NotImplementedError("SBDART option not supported by this front end: " + "; ".join(todo))

    def _usrcloud(self, wl, nmom):
        """usrcloud (taucloud.f:142-274): cloud layers from usrcld.dat (nre(1) = 0, no tcloud / lwp).
        Liquid water only: with a frozen water path the reference divides by its INTEGER
        parameter rhoice = .917 -> 0 (taucloud.f:193, :244) and the optical depth is infinite."""
        from . import extras
        if self.p["imomc"] < 0:
            raise ValueError("imomc < 0 not allowed with usrcld.dat option")
        if getattr(self, "_usrcld", None) is None:
            self._usrcld = extras.usrcloud_table(self.nz)
        tab = self._usrcld
        if (tab[:, 2] != 0.).any():
            raise NotImplementedError("usrcld.dat with a frozen water path: the reference's integer rhoice = 0 "
                                      "makes the ice optical depth infinite (taucloud.f:193, :244)")
        nz = self.nz
        taucld, wcld = np.zeros(nz), np.zeros(nz)
        pmom = np.zeros((nz, nmom + 1))
        for i in range(nz):
            lwp, reff, cldfrac = tab[i, 0], tab[i, 1], tab[i, 4]
            if lwp > 0.:
                qw, ww, gw = cloudpar(wl, reff)
                tauw = f32(.75) * qw * lwp / reff
            else:
                qw = ww = gw = tauw = 0.
            taucld[i] = tauw
            if taucld[i] != 0.:
                wcld[i] = (tauw * ww) / taucld[i]
                gcld = (tauw * gw) / taucld[i]
                pmom[i] = getmom(self.p["imomc"], gcld, nmom)
            taucld[i] = taucld[i] * cldfrac ** 1.5
            pmom[i, 1:] = taucld[i] * wcld[i] * pmom[i, 1:]
        return taucld, wcld, pmom

    def _unsupported_surface(self):
        if self.p["isalb"] in (7, 8, 9) and self.radcalc and self.p["corint"]:
            return ["corint with a BRDF surface"]
        return []

    @staticmethod
    def _nearest(xx, x):
        k = locate(xx, x)
        return k + 1 if abs(x - xx[k]) < abs(x - xx[k - 1]) else k

    def _vuangles(self):
        p = self.p
        phi, uzen, vzen = p["phi"].copy(), p["uzen"].copy(), p["vzen"]
        nphi, nzen, iout = p["nphi"], p["nzen"], p["iout"]
        if nphi > 0:
            p1, p2 = min(phi[0], phi[1]), max(phi[0], phi[1])
            phi[:nphi] = [p1 + i * (p2 - p1) / float(np.float32(nphi - 1)) for i in range(nphi)]
        else:
            nphi = numset(ZIP, phi)
            if nphi == 0:
                nphi = 19
                phi[:nphi] = [0. + i * 180. / float(np.float32(nphi - 1)) for i in range(nphi)]
        nvzen = numset(90.0, vzen)
        uzen[:nvzen] = 180. - vzen[:nvzen]
        if nzen > 0:
            z1, z2 = min(uzen[0], uzen[1]), max(uzen[0], uzen[1])
            ii = 0
            for i in range(nzen):
                x = z1 + i * (z2 - z1) / float(np.float32(nzen - 1))
                if abs(x - 90.) > f32(.05):
                    uzen[ii] = x
                    ii += 1
            nzen = ii
        else:
            nzen = numset(ZIP, uzen)
            if nzen == 0:
                nzen, z1, z2 = {5: (18, 0., 85.), 20: (18, 0., 85.), 6: (18, 95., 180.),
                                21: (18, 95., 180.)}.get(iout, (36, 0., 180.))
                uzen[:nzen] = [z1 + (z2 - z1) * i / float(np.float32(nzen - 1)) for i in range(nzen)]
                if p["nstr"] == 4:
                    p["nstr"] = min(2 * (max(nphi, nzen) // 2), NSTRMS)
        self.nphi, self.nzen = nphi, nzen
        self.phi, self.uzen = phi[:nphi].copy(), uzen[:nzen].copy()

    # ---- wavelength loop, optical properties only (drt.f:425-533)
    def bins(self):
        """All (wavelength, k-term) bins of the run as stacked arrays."""
        p, nz = self.p, self.nz
        nstr = p["nstr"]
        nmom = min(nstr + 2, NSTRMS)
        if self.radcalc and p["corint"]:
            nmom = MAXMOM                 # the full phase function for INTCOR (drt.f:490-491)
        rows = []
        if self.kdist == -1:
            # readk (taugas.f:7695-7760): the spectral intervals of CKTAU inside the range
            spectrum = []
            for r in self.cktau:
                if r["vnu0"] < self.vnulo:
                    break
                if r["vnu0"] <= self.vnuhi:
                    if min(r["vnu1"], r["vnu2"]) <= 0.:
                        raise ValueError(f"readk --- wvnmlo,wvnmhi:  {r['vnu1']} {r['vnu2']}")
                    if r["dtk"].min() < 0.:
                        raise ValueError("readk --- negative dtauk")
                    spectrum.append(r)
        else:
            spectrum = range(self.nwl)
        for il, rec in enumerate(spectrum):
            amu0 = self.amu0
            ib = nb = 1
            ewcoef = 1.
            if self.kdist == -1:
                wl, wvnmlo, wvnmhi = 10000. / rec["vnu0"], rec["vnu1"], rec["vnu2"]
                ib, nb, nk, ewcoef = rec["ib"], rec["nb"], rec["nk"], rec["ewc"]
                gwk, dtauk, dtaugc = rec["gw"], rec["dtk"], None
            else:
                wl, wvnmhi, wvnmlo = wllimits(il, self.nwl, self.wlinc, self.wl1, self.wl2)
                nk, gwk, dtauk, dtaugc = gasset(self.kdist, wl, self.uu, amu0, self.z, p["xo4"])
            dwl = 10000. / wvnmlo - 10000. / wvnmhi
            etirr = rec["etf"] if (self.kdist == -1 and p["nf"] == -2) else self.sun(wl) * dwl
            flxin = etirr * p["solfac"]
            if p["nf"] == 0:
                flxin = dwl
            if self.sza >= 90.:
                flxin = 0.
                amu0 = 1.
                self.amu0 = 1.          # the reference overwrites amu0 for good (drt.f:456-459)
            ff = self.filter(wl) * ewcoef
            plank = (wl > 2.) if p["nothrm"] < 0 else (p["nothrm"] == 0)
            rsfc = max(0.0, min(self.albedo(wl, self._warn), 1.0))
            pmom = np.zeros((nz, nmom + 1))
            dtauc, wcld = np.zeros(nz), np.zeros(nz)
            if self.clouds.mcldz > 0:
                dtauc, wcld, pmom = self.clouds(wl, nmom)
            elif p["nre"][0] == 0.:
                dtauc, wcld, pmom = self._usrcloud(wl, nmom)
            dtaua, waer = np.zeros(nz), np.zeros(nz)
            if self.aerosols.active:
                dtaua, waer = self.aerosols(wl, nmom, pmom)
            dtaur = rayleigh(wl, self.z, self.pr, self.t)
            if p["xrsc"] != 1.0:
                dtaur = p["xrsc"] * dtaur
            # normom (drt.f:1366-1397)
            pmom[:, 2] += f32(.1) * dtaur
            dtsct = dtauc * wcld + dtaua * waer + dtaur
            nzr = dtsct != 0.
            pmom[nzr] = pmom[nzr] / dtsct[nzr, None]
            pmom[:, 0] = 1.
            for kd in range(nk):
                dtau, wreal, wt = depthscl(self.kdist, kd, nk, wl, dtaur, dtaua, waer, dtauc, wcld,
                                           gwk, dtauk, dtaugc, self.spowder)
                rows.append(dict(il=il, kd=kd, nk=nk, ib=ib, nb=nb, wl=wl, dwl=dwl, wt=wt, ff=ff, dtau=dtau,
                                 ssalb=wreal, pmom=pmom, flxin=flxin, amu0=amu0, rsfc=rsfc,
                                 surf=(il if (self.surface is not None and self.surface.spectral) else 0),
                                 plank=plank, wvnmlo=wvnmlo, wvnmhi=wvnmhi))
        return rows

    def batch(self, rows):
        """Stack bins into the arrays of the batched C ABI (include/sbdart_b200.h)."""
        from .. import make_bins
        B = len(rows)
        g = lambda k: np.array([r[k] for r in rows])  # noqa: E731
        bins = make_bins(B, fbeam=g("flxin"), umu0=g("amu0"), phi0=self.phi0, fisot=self.p["fisot"],
                         albedo=g("rsfc"), btemp=self.btemp, ttemp=self.ttemp, temis=self.p["temis"],
                         wvnmlo=g("wvnmlo"), wvnmhi=g("wvnmhi"), accur=0.0,
                         plank=g("plank").astype(np.int32), col=0)
        d = dict(dtauc=np.stack([r["dtau"] for r in rows]), ssalb=np.stack([r["ssalb"] for r in rows]),
                 pmom=np.stack([r["pmom"] for r in rows]), bins=bins, temper=self.temper[None, :],
                 nstr=self.p["nstr"], group=g("il"))
        self._chekin_warnings(d)
        if self.surface is not None:
            # LAMBER = .FALSE. (drt.f:469-470): SURFAC's tables per surface; one surface per
            # wavelength for the ocean model (wl = 20000 / (wvnmlo + wvnmhi), spectra.f:283)
            from . import brdf
            ids = sorted(set(int(r["surf"]) for r in rows))
            first = {int(r["surf"]): r for r in reversed(rows)}
            states = [brdf.ocean_state(tables(), self.surface, 20000. / (first[i]["wvnmhi"] + first[i]["wvnmlo"]))
                      if self.surface.spectral else None for i in ids]
            remap = {i: k for k, i in enumerate(ids)}
            d["surf"] = np.array([remap[int(r["surf"])] for r in rows], dtype=np.int32)
            d["surface"] = brdf.SurfaceSpec(self.surface, states, rows[0]["amu0"], bool((g("flxin") > 0).any()),
                                            umu=self.umu if self.radcalc else None)
        if self.radcalc:
            d["umu"], d["phi"] = self.umu, self.phi
            d["corint"] = bool(self.p["corint"])
            # output levels whose intensities the records consume (drt.f:1008-1016,
            # :1143-1151): iout 5/20 the top level, 6/21 the bottom, 23 both
            iout = self.p["iout"]
            lv = {5: [self.ntop - 1], 20: [self.ntop - 1], 6: [self.nbot - 1], 21: [self.nbot - 1],
                  23: [self.ntop - 1, self.nbot - 1]}.get(iout)
            if lv is not None:
                d["uu_levels"] = sorted(set(lv))
        return d

    def _chekin_warnings(self, d):
        """The non-fatal messages of CHEKIN (disort.f:4938-4941, :5158-5172)."""
        corint = bool(self.p["corint"]) and self.radcalc
        nmom = d["pmom"].shape[2] - 1
        if corint and nmom > 10 and (d["pmom"][:, :, nmom] > f32(1.e-3)).any():
            self._warn(5, "CHEKIN-- phase function not sufficiently resolved for use with corint=.true.")
        if d["bins"]["plank"].any() and (np.abs(np.diff(self.temper)) > 10.0).any():
            self._warn(6, "CHEKIN--vertical temperature step may be too large for good accuracy")
        if self.radcalc and not corint:
            sun = d["bins"]["fbeam"] > 0.0
            if sun.any() and (d["ssalb"][sun] > 0.0).any():      # YESSCT > 0, DELTAM
                self._warn(7, "CHEKIN--intensity correction is off; intensities may be less accurate")

    _FATAL = {-1: "DISORT--input and/or dimension errors", -2: "ASYMTX--convergence problems",
              -3: "SOLVE0--boundary system singular"}

    def _check_status(self, status):
        """Per-bin failures are fatal in the reference (errmsg(0, ...) then STOP)."""
        status = np.asarray(status)
        if (status == 1).any():
            self._warn(1, "SETDIS--beam angle=computational angle; change NSTR")
        bad = status[status < 0]
        if len(bad):
            raise SbdartFatal(self._FATAL.get(int(bad[0]), f"DISORT--bin status {int(bad[0])}"))

    # ---- accumulation and records (stdout0/1/2, drt.f:892-1165)
    def run(self, solve):
        """solve(batch) -> dict(rfldir, rfldn, flup [B][nz+1][, uu [B][nphi][nz+1][numu]], status)."""
        rows = self.bins()
        res = None
        if rows:
            # ff == 0 bins are skipped by the reference (drt.f:539); none for isat <= 0
            b = self.batch(rows)
            res = solve(b)
            bad = np.asarray(res["status"]) != 0
            if bad.any():
                res = self._retry(b, res, solve)
        return self.records(rows, res)

    def run_sharded(self, solve, dist=None):
        """One run over several GPUs: every rank builds the bin list, solves its
        contiguous block (sharding.bin_partition never splits the k-terms of a wavelength), one
        all-gather returns the per-bin outputs in loop order and every rank does the ordered
        accumulation of drt.f:977-982 itself, so the records are byte-identical for any number of
        ranks.  `dist` is torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
        import torch
        from ..sharding import bin_partition, gather_outputs
        rows = self.bins()
        if not rows:
            return self.records(rows, None)
        b = self.batch(rows)
        world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        parts = bin_partition(len(rows), world, b["group"])
        lo, hi = parts[rank]
        NT = self.nz + 1
        keys = ["rfldir", "rfldn", "flup"] + (["uu"] if self.radcalc else [])
        empty = {k: np.zeros((0, NT)) for k in keys[:3]}
        lv = b.get("uu_levels")             # levels the records consume: only those are gathered
        if self.radcalc:
            empty["uu"] = np.zeros((0, len(self.phi), NT if lv is None else len(lv), len(self.umu)))
        res, failure = empty, None
        if hi > lo:
            sub = dict(b)
            for k in ("dtauc", "ssalb", "pmom", "bins", "group"):
                sub[k] = b[k][lo:hi]
            try:
                res = solve(sub)
                if (np.asarray(res["status"]) != 0).any():
                    res = self._retry(sub, res, solve)
                if lv is not None and "uu_levels" not in res:
                    res = dict(res, uu=np.ascontiguousarray(res["uu"][:, :, lv, :]))
            except Exception as e:          # noqa: BLE001 -- reported on every rank below
                if world == 1:
                    raise
                res, failure = empty, e
        if lv is not None:
            res = dict(res, uu_levels=lv)
        if world == 1:
            return self.records(rows, res)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        # a rank whose solve failed must not leave the others waiting in the all-gather:
        # exchange an ok flag first and raise on every rank
        ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if failure is not None:
                raise failure
            raise RuntimeError("run_sharded: the solve failed on another rank")
        full = {k: gather_outputs(torch.from_numpy(np.ascontiguousarray(res[k], dtype=np.float64)).to(dev),
                                  parts, dist).cpu().numpy() for k in keys}
        if lv is not None:
            full["uu_levels"] = lv
        return self.records(rows, full)

    def run_device(self, solver):
        """Whole-spectrum GPU path: the optical properties of every bin are produced
        by the K2 kernel and never leave the device (frontend/device.py)."""
        from .device import device_aerosols_supported, run_spectrum
        def host_solve(b):
            bins = b["bins"]
            if "surface" in b:          # BRDF surfaces: tables to the device, bins select them
                from .. import quadrature, surface_albedo
                tab = b["surface"].tables(b["nstr"], quadrature)
                solver.set_surfaces(b["nstr"], tab["bdr"], tab["bem"], tab.get("rmu"), tab.get("emu"))
                bins = bins.copy()
                bins["albedo"] = surface_albedo(b["surf"])
            try:
                return solver.disort_batch(
                    b["dtauc"], b["ssalb"], b["pmom"], bins, nstr=b["nstr"], temper=b["temper"],
                    umu=b.get("umu"), phi=b.get("phi"), uu_levels=b.get("uu_levels"),
                    corint=b.get("corint", False), uu_packed=True)
            finally:
                if "surface" in b:
                    solver.set_surfaces()
        if (not device_aerosols_supported(self.aerosols) or self.p["imomc"] not in (2, 3) or
                (self.radcalc and self.p["corint"]) or self.surface is not None or self.kdist < 0 or self.spowder or
                (self.clouds.mcldz == 0 and self.p["nre"][0] == 0.)):
            # table phase functions (getmom 4/5, pmaer) and the 299-moment CORINT runs:
            # optical properties on the host, solve on the GPU
            return self.run(host_solve)
        rows, res = run_spectrum(self, solver)
        # the CHEKIN / SALBEDO warnings the host-side batch would have raised
        nothrm = self.p["nothrm"]
        if (nothrm == 0 or (nothrm < 0 and self.wl2 > 2.0)) and (np.abs(np.diff(self.temper)) > 10.0).any():
            self._warn(6, "CHEKIN--vertical temperature step may be too large for good accuracy")
        if self.radcalc and not self.p["corint"] and self.sza < 90. and self.p["xrsc"] > 0.:
            self._warn(7, "CHEKIN--intensity correction is off; intensities may be less accurate")
        for wl in (self.wl1, self.wl2):
            self.albedo(wl, self._warn)
        if (res["status"] != 0).any():
            # beam / quadrature clash (drt.f:536-554): fall back to the host-side batch
            # for the retry logic (rare: one NSTR-dependent angle)
            return self.run(host_solve)
        return self.records(rows, res)

    def records(self, rows, res):
        """stdout0/1/2 (drt.f:892-1165): ordered accumulation and IOUT records."""
        # a solver that returns the selected levels only (packed layout of the C ABI) says which
        lv = None if res is None else res.get("uu_levels")
        uslot = (lambda j: j) if lv is None else (lambda j: list(lv).index(j))
        out = []
        iout, nz = self.p["iout"], self.nz
        if iout in (1, 5, 6):
            out.append("")
            out.append('"tbf')
            out.append(f"{self.nwl:15d}")
        elif iout == 7:
            out.append("")
            out.append('"fzw')
            out.append(f"{nz:15d}")
        ntop, nbot = self.ntop - 1, self.nbot - 1          # 0-based level indices
        topdn = topup = topdir = botdn = botup = botdir = 0.0
        weq = wfull = phidw = 0.0
        fxdn, fxup, fxdir = np.zeros(nz), np.zeros(nz), np.zeros(nz)
        uurs = np.zeros((self.nzen, self.nphi)) if self.radcalc else None
        uurl = np.zeros((self.nphi, self.nzen, nz)) if iout == 22 else None
        for ib, r in enumerate(rows):
            rfldir, rfldn, flup = res["rfldir"][ib], res["rfldn"][ib], res["flup"][ib]
            dwt = r["wt"] * r["ff"]
            kd, nk = r["kd"] + 1, r["nk"]
            # sub-bands of the CKTAU files: a wavelength starts with (kd = 1, ib = nb) and is
            # finished with (kd = nk, ib = 1) (drt.f:967, :989); ib = nb = 1 otherwise
            first = kd == 1 and r.get("ib", 1) == r.get("nb", 1)
            last = kd == nk and r.get("ib", 1) == 1
            if iout in (1, 5, 6):
                if first:
                    topdn = topup = topdir = botdn = botup = botdir = 0.0
                    weq = wfull = 0.0
                topdn += (rfldn[ntop] + rfldir[ntop]) * dwt
                topup += flup[ntop] * dwt
                topdir += rfldir[ntop] * dwt
                botdn += (rfldn[nbot] + rfldir[nbot]) * dwt
                botup += flup[nbot] * dwt
                botdir += rfldir[nbot] * dwt
                if kd == nk:
                    weq += r["dwl"] * r["ff"]
                    wfull += r["dwl"]
                if last:
                    if weq == 0.:
                        weq = f32(1.e-30)
                    out.append(_f(r["wl"], 12, 8) + _f(weq / wfull, 9, 5) + "".join(
                        _es(_r4(x / weq), 12, 4) for x in (topdn, topup, topdir, botdn, botup, botdir)))
                if iout in (5, 6):
                    j = ntop if iout == 5 else nbot
                    if first:
                        uurs[:] = 0.
                    uurs += res["uu"][ib][:, uslot(j), :].T * dwt
                    if last:
                        out.append(f"{self.nphi:4d}{self.nzen:4d}")
                        out += self._rows([_r4(x) for x in self.phi], 10)
                        out += self._rows([_r4(x) for x in self.uzen], 10)
                        for i in range(self.nzen - 1, -1, -1):
                            out += self._rows([_r4(x / weq) for x in uurs[i]], 10)
            if iout in (10, 11, 20, 21, 22, 23) and kd == nk:
                phidw += r["dwl"] * r["ff"]
            if iout in (7, 11, 22):
                if iout == 7 and first:
                    fxdn[:] = fxup[:] = fxdir[:] = 0.
                fxdn += (rfldn[1:nz + 1] + rfldir[1:nz + 1]) * dwt
                fxup += flup[1:nz + 1] * dwt
                fxdir += rfldir[1:nz + 1] * dwt
            if iout == 7 and last:
                out += ["", "", _f(r["wl"], 12, 8)]
                for arr in (self.z[::-1], [_r4(x) for x in fxdir], [_r4(x) for x in fxdn - fxdir],
                            [_r4(x) for x in fxdn], [_r4(x) for x in fxup]):
                    out.append("")
                    out += self._rows(arr, 10, w=11, d=3)
            if iout in (10, 20, 21, 23):
                topdn += (rfldn[ntop] + rfldir[ntop]) * dwt
                topup += flup[ntop] * dwt
                topdir += rfldir[ntop] * dwt
                botdn += (rfldn[nbot] + rfldir[nbot]) * dwt
                botup += flup[nbot] * dwt
                botdir += rfldir[nbot] * dwt
            if iout in (20, 21):
                j = ntop if iout == 20 else nbot
                uurs += res["uu"][ib][:, uslot(j), :].T * dwt
            if iout == 22:          # radiance at every level (drt.f:1065-1073)
                uurl += np.transpose(res["uu"][ib][:, 1:nz + 1, :], (0, 2, 1)) * dwt
            if iout == 23:
                for i in range(self.nzen):
                    j = ntop if self.uzen[self.nzen - 1 - i] < 90. else nbot
                    uurs[i] += res["uu"][ib][:, uslot(j), i] * dwt
        # stdout2
        p = self.p
        if iout == 11:
            out.append(f"{nz:4d}" + _es(phidw, 15, 7))
            fntm = zm = pm = 0.
            for i in range(nz):
                zz, pp = self.z[nz - 1 - i], self.pr[nz - 1 - i]
                fnt = fxdn[i] - fxup[i]
                if i > 0:
                    dfdz = (fntm - fnt) / (zm - zz)
                    heat = f32(.01) * 3600 * 24 * GRAV * (fntm - fnt) / (1004. * (pp - pm))
                else:
                    dfdz = heat = 0.
                fntm, zm, pm = fnt, zz, pp
                out += self._rows([zz, pp, _r4(fxdn[i]), _r4(fxup[i]), _r4(fxdir[i]), _r4(dfdz),
                                   _r4(heat)], 10)
        if iout in (10, 20, 21, 23):
            out.append(_f(self.wl1, 11, 4) + _f(self.wl2, 11, 4) + _f(phidw, 11, 4) + "".join(
                _es(_r4(x), 12, 4) for x in (topdn, topup, topdir, botdn, botup, botdir)))
        if iout in (20, 21, 23):
            out.append(f"{self.nphi:4d}{self.nzen:4d}")
            out += self._rows(self.phi, 10)
            out += self._rows(self.uzen, 10)
            for i in range(self.nzen - 1, -1, -1):
                out += self._rows([_r4(x) for x in uurs[i]], 20)
        if iout == 22:              # drt.f:1153-1163
            out.append(f"{self.nphi:4d}{self.nzen:4d}{nz:4d}" + _es(phidw, 12, 4))
            out += self._rows(self.phi, 10)
            out += self._rows(self.uzen, 10)
            out += self._rows(self.z[::-1], 10)
            for arr in (fxdn, fxup, fxdir):
                out += self._rows([_r4(x) for x in arr], 10)
            out += self._rows([_r4(uurl[i, j, k]) for k in range(nz) for j in range(self.nzen - 1, -1, -1)
                               for i in range(self.nphi)], 10)
        self.last = dict(rows=rows, result=res if rows else None)
        return "\n".join(out) + "\n"

    def _retry(self, b, res, solve):
        """NSTR dithering: bins that report the beam/quadrature clash are re-solved
        with NSTR-2, then NSTR+2 (drt.f:536-554)."""
        res = {k: (np.array(v, copy=True) if k != "uu_levels" else v) for k, v in res.items()}
        self._check_status(res["status"])
        for j in (1, 2):
            idx = np.nonzero(res["status"] == 1)[0]
            if len(idx) == 0:
                break
            nstr = self.nstrsv + j * (3 * j - 5)
            if nstr < 4 or nstr > NSTRMS:
                continue
            sub = dict(b)
            for k in ("dtauc", "ssalb", "pmom", "bins", "group") + (("surf",) if "surf" in b else ()):
                sub[k] = b[k][idx]
            sub["nstr"] = nstr
            r2 = solve(sub)
            self._check_status(r2["status"])
            for k in res:
                if k != "uu_levels":
                    res[k][idx] = r2[k]
        if (res["status"] != 0).any():
            raise RuntimeError("Error --- NSTR dithering procedure failed")      # drt.f:550-553
        return res

    @staticmethod
    def _rows(vals, per, w=12, d=4):
        vals = list(vals)
        return ["".join(_es(v, w, d) for v in vals[i:i + per]) for i in range(0, len(vals), per)]
