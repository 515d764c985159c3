"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.  The product package (sbdart_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SbdoInput(C.Structure):
    """Mirror of sbdo_input (oracle/disort_oracle.h)."""

    _fields_ = [(k, C.c_int) for k in (
        "nlyr", "nstr", "nmom", "usrtau", "ntau", "usrang", "numu", "nphi",
        "plank", "onlyfl", "corint", "lamber")] + [(k, C.c_double) for k in (
            "fbeam", "umu0", "phi0", "fisot", "albedo", "btemp", "ttemp",
            "temis", "wvnmlo", "wvnmhi", "accur")]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsbd_oracle.so")
    src = os.path.join(_HERE, "disort_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libsbd_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def build_ref() -> bool:
    """Try to build the untouched reference into oracle/_ref/sbdart (needs gfortran and
    /root/reference; neither exists on the GPU box).  Returns True when the binary exists."""
    subprocess.run(["make", "-C", _HERE, "_ref"], check=False, stdout=subprocess.DEVNULL)
    return os.path.exists(os.path.join(_HERE, "_ref", "sbdart"))


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        _LIB.sbdo_disort.restype = C.c_int
        _LIB.sbdo_disort.argtypes = [C.POINTER(SbdoInput)] + [dp] * 14 + [ip]
        _LIB.sbdo_qgausn.argtypes = [C.c_int, dp, dp]
        _LIB.sbdo_plkavg.restype = C.c_double
        _LIB.sbdo_plkavg.argtypes = [C.c_double, C.c_double, C.c_double, ip]
        _LIB.sbdo_asymtx.restype = C.c_int
        _LIB.sbdo_asymtx.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_int, dp]
        _LIB.sbdo_set_bdref.restype = None
        _LIB.sbdo_set_bdref.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_double]
        _LIB.sbdo_bdref_eval.restype = C.c_double
        _LIB.sbdo_bdref_eval.argtypes = [C.c_double] * 3
        _LIB.sbdo_disort_flux_batch.restype = C.c_int
        _LIB.sbdo_disort_flux_batch.argtypes = (
            [C.c_int] * 4 + [dp] * 6 + [ip] + [dp] * 7 + [ip] + [dp] * 5 + [ip, C.c_int])
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def disort(dtauc, ssalb, pmom, *, nstr, temper=None, utau=None, umu=None,
           phi=None, fbeam=0.0, umu0=1.0, phi0=0.0, fisot=0.0, albedo=0.0,
           btemp=0.0, ttemp=0.0, temis=0.0, wvnmlo=0.0, wvnmhi=0.0,
           plank=False, onlyfl=True, corint=False, accur=0.0, lamber=True):
    """One DISORT call through the oracle.

    pmom is [nlyr][nmom+1].  Returns a dict with rfldir, rfldn, flup, dfdt,
    uavg [NT], uu [nphi][NT][numu] (radiance runs), u0u, status, warn.
    """
    dtauc = _f64(dtauc)
    ssalb = _f64(ssalb)
    pmom = _f64(pmom)
    nlyr = dtauc.shape[0]
    nmom = pmom.shape[1] - 1
    utau = _f64(utau)
    umu = _f64(umu)
    phi = _f64(phi)
    temper = _f64(temper)
    inp = SbdoInput()
    inp.nlyr, inp.nstr, inp.nmom = nlyr, nstr, nmom
    inp.usrtau = int(utau is not None)
    inp.ntau = 0 if utau is None else utau.shape[0]
    inp.usrang = int(umu is not None)
    inp.numu = 0 if umu is None else umu.shape[0]
    inp.nphi = 0 if phi is None else phi.shape[0]
    inp.plank, inp.onlyfl, inp.corint, inp.lamber = int(plank), int(onlyfl), int(corint), int(lamber)
    inp.fbeam, inp.umu0, inp.phi0, inp.fisot, inp.albedo = fbeam, umu0, phi0, fisot, albedo
    inp.btemp, inp.ttemp, inp.temis = btemp, ttemp, temis
    inp.wvnmlo, inp.wvnmhi, inp.accur = wvnmlo, wvnmhi, accur
    nt = inp.ntau if inp.usrtau else nlyr + 1
    nu = inp.numu if (inp.usrang and not onlyfl) else nstr
    out = {k: np.zeros(nt) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
    uu = np.zeros((max(inp.nphi, 1), nt, nu))
    u0u = np.zeros((nt, nu))
    warn = C.c_int(0)
    st = lib().sbdo_disort(C.byref(inp), _dp(dtauc), _dp(ssalb), _dp(pmom), _dp(temper),
                           _dp(utau), _dp(umu), _dp(phi), _dp(out["rfldir"]),
                           _dp(out["rfldn"]), _dp(out["flup"]), _dp(out["dfdt"]),
                           _dp(out["uavg"]), _dp(uu), _dp(u0u), C.byref(warn))
    out.update(uu=uu, u0u=u0u, status=st, warn=warn.value)
    return out


def set_bdref(ibdrf, sc, nr=0.0, ni=0.0, rsw=0.0):
    """Surface model of the following lamber=False calls (suralb, spectra.f:139-162)."""
    p = np.zeros(5)
    p[:len(sc)] = sc
    lib().sbdo_set_bdref(int(ibdrf), _dp(p), float(nr), float(ni), float(rsw))


def bdref(mur, mui, phir):
    return lib().sbdo_bdref_eval(float(mur), float(mui), float(phir))


def disort_flux_batch(dtauc, ssalb, pmom, *, nstr, fbeam, umu0, albedo,
                      plank=None, wvnmlo=None, wvnmhi=None, btemp=None,
                      ttemp=None, temis=None, fisot=None, temper=None,
                      col=None, nthreads=1):
    """Flux-only batch [B][L] through the oracle (OpenMP over bins)."""
    dtauc, ssalb, pmom = _f64(dtauc), _f64(ssalb), _f64(pmom)
    nb, nlyr = dtauc.shape
    nmom = pmom.shape[2] - 1
    z = np.zeros(nb)

    def per(a):
        return z if a is None else _f64(np.broadcast_to(a, (nb,)))

    fb, mu0, alb = per(fbeam), per(umu0), per(albedo)
    pl = np.zeros(nb, np.int32) if plank is None else np.ascontiguousarray(
        np.broadcast_to(plank, (nb,)), np.int32)
    cl = np.zeros(nb, np.int32) if col is None else np.ascontiguousarray(col, np.int32)
    tp = _f64(np.zeros((1, nlyr + 1)) if temper is None else np.atleast_2d(temper))
    out = {k: np.zeros((nb, nlyr + 1)) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
    status = np.zeros(nb, np.int32)
    lib().sbdo_disort_flux_batch(
        nb, nlyr, nstr, nmom, _dp(dtauc), _dp(ssalb), _dp(pmom), _dp(fb), _dp(mu0),
        _dp(alb), _ip(pl), _dp(per(wvnmlo)), _dp(per(wvnmhi)), _dp(per(btemp)),
        _dp(per(ttemp)), _dp(per(temis)), _dp(per(fisot)), _dp(tp), _ip(cl),
        _dp(out["rfldir"]), _dp(out["rfldn"]), _dp(out["flup"]), _dp(out["dfdt"]),
        _dp(out["uavg"]), _ip(status), int(nthreads))
    out["status"] = status
    return out
