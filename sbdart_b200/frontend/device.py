"""Whole-spectrum GPU path: pack the tables for the producer kernel (K2,
csrc/sbd_optics.cu), hand the per-run setup to sbd_spectrum_run and feed the
per-bin fluxes back into the record formatter of frontend.Sbdart.

Table order = enum OpticsTable in csrc/sbd_optics.cuh.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import BIN_DTYPE, SbdError, lib
from . import (NCLDZ, T, _BANDS, _MOLS, Sbdart, cloudpar, f32, levrng)


class OpticsParams(C.Structure):
    """struct sbd_optics_params (include/sbdart_b200.h)."""

    _fields_ = [(k, C.c_int32) for k in (
        "nz", "nwl", "kdist", "nstr", "nf", "nothrm", "imomc", "ncloud", "nalb", "nsun", "night",
        "pad_")] + [(k, C.c_double) for k in (
            "wl1", "wl2", "wlinc", "amu0", "xo4", "xrsc", "solfac", "phi0", "fisot", "temis",
            "btemp", "ttemp")]


CLOUD_DTYPE = np.dtype([("layer", "<i4"), ("use_tau", "<i4"), ("reff", "<f8"), ("tcld", "<f8"),
                        ("lwpth", "<f8"), ("q550", "<f8")])


class AerosolParams(C.Structure):
    """struct sbd_aerosol_params (include/sbdart_b200.h)."""

    _fields_ = [(k, C.c_int32) for k in ("nwlbaer", "imoma", "nosct", "nstrat", "nz", "pad_")] + [
        ("abaer", C.c_double)]


STRAT_DTYPE = np.dtype([("layer", "<f8"), ("taerst", "<f8"), ("ext", "<f8", 47), ("absb", "<f8", 47),
                        ("asym", "<f8", 47)])


def device_aerosols_supported(aer):
    """The producer kernel evaluates getmom(2|3) only; table phase functions (imoma 4, 5)
    and user moments (pmaer) stay on the host path."""
    if aer.iaer == -1:          # aerosol.dat is read on the host
        return False
    return (not aer.active) or aer.iaer <= 0 or aer.imoma in (2, 3)


def set_aerosols(L, solver, aer):
    """Hand the wavelength-independent aerosol setup (extras.Aerosols) to the handle."""
    if not aer.active:
        rc = L.sbd_spectrum_set_aerosols(solver._h, None, None, None, None, None, None, None, None)
        if rc:
            raise SbdError(rc, "sbd_spectrum_set_aerosols")
        return
    A = AerosolParams()
    A.nz, A.nosct, A.abaer, A.imoma = aer.nz, aer.nosct, aer.abaer, aer.imoma
    cd = lambda v: np.ascontiguousarray(v, dtype=float)  # noqa: E731
    wlb = ext = ab = asm = dtsv = None
    if aer.iaer > 0:
        n = aer.nwlbaer
        A.nwlbaer = n
        wlb, ext, ab, asm, dtsv = cd(aer.wlb[:n]), cd(aer.aerext[:n]), cd(aer.aerabs[:n]), cd(aer.aerasm[:n]), cd(aer.dtsv)
    ent = [(aer.laer[i], aer.taerst[i], aer.aerstr[:, 0, aer.jaer[i] - 1], aer.aerstr[:, 1, aer.jaer[i] - 1],
            aer.aerstr[:, 2, aer.jaer[i] - 1]) for i in range(5) if aer.jaer[i] != 0 and aer.taerst[i] > 0.]
    strat = np.array(ent, dtype=STRAT_DTYPE) if ent else None
    A.nstrat = len(ent)
    awl = cd(aer.awl)
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    rc = L.sbd_spectrum_set_aerosols(solver._h, C.byref(A), ptr(wlb), ptr(ext), ptr(ab), ptr(asm), ptr(dtsv),
                                     ptr(awl), ptr(strat))
    if rc:
        raise SbdError(rc, "sbd_spectrum_set_aerosols")


class InputsOut(C.Structure):
    _fields_ = [("dtauc", C.c_void_p), ("ssalb", C.c_void_p), ("pmom", C.c_void_p),
                ("bins", C.c_void_p)]


def pack_tables():
    """Flat float64 bundle + (offsets, lengths) in enum OpticsTable order."""
    parts = [
        T("taugas/slf296/s"), T("taugas/slf260/s"), T("taugas/frn296/f"),
        T("taugas/c4dta/c4"), T("taugas/hno3/h1"), T("taugas/hno3/h2"), T("taugas/hno3/h3"),
        T("taugas/o2cont/o2s0"), T("taugas/o2cont/o2a"), T("taugas/o2cont/o2b"),
        T("taugas/o4cont/sig"),
        T("taugas/o3hht/s0"), T("taugas/o3hht/s1"), T("taugas/o3hht/s2"), T("taugas/o3uv/s"),
        T("taugas/c8dta/c8"), T("taugas/schrun/shn"), T("taugas/kdistr/fac"),
    ]
    for nm in ("qq", "ww", "gg", "qqi", "wwi", "ggi"):
        parts.append(np.ascontiguousarray(T(f"taucloud/cloudpar/{nm}").T).ravel())   # [re][wl]
    for pre in ("taugas/gasblk/cp", "taugas/gasblk/iwl", "taugas/gasblk/iwh", "taugas/abcdta/a",
                "taugas/abcdta/aa", "taugas/abcdta/bb", "taugas/abcdta/cc"):
        for m in _MOLS:
            parts.append(np.atleast_1d(T(pre + m)).astype(float))
    quads = []
    for imol in sorted(_BANDS):
        for iw, rngs in _BANDS[imol]:
            for lo, hi in rngs:
                quads += [imol, iw, lo, hi]
    parts.append(np.array(quads, dtype=float))
    parts = [np.asarray(p, dtype=np.float64).ravel() for p in parts]
    lens = np.array([len(p) for p in parts], dtype=np.int32)
    # every table starts on an even index (16-byte alignment is not required, but harmless)
    offs = np.zeros(len(parts), dtype=np.int32)
    flat = []
    pos = 0
    for i, p in enumerate(parts):
        offs[i] = pos
        flat.append(p)
        pos += len(p)
        if pos & 1:
            flat.append(np.zeros(1))
            pos += 1
    return np.concatenate(flat), np.concatenate([offs, lens]).astype(np.int32)


def cloud_entries(clouds):
    """Resolve taucloud's (cloud layer i, DISORT layer j) loop (taucloud.f:61-124);
    everything here is independent of wavelength."""
    out = []
    tcl, lw, nre = clouds.tcloud, clouds.lwp, clouds.nre
    q550 = {}
    for i in range(1, NCLDZ + 1):
        lbot, ltop = levrng(clouds.lcld, i)
        if lbot == 0 or (tcl[i - 1] == 0. and lw[i - 1] == 0.):
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
            use_tau = int(tcl[i - 1] != 0.)
            if use_tau and j not in q550:
                q550[j] = cloudpar(f32(0.55), reff)[0]     # first-call cache of the reference
            out.append((j, use_tau, reff, tcld, lwpth, q550.get(j, 1.0)))
    return np.array(out, dtype=CLOUD_DTYPE) if out else np.zeros(0, dtype=CLOUD_DTYPE)


_UPLOADED = set()


def _bind(L):
    if getattr(L, "_spectrum_bound", False):
        return
    L.sbd_optics_upload_tables.restype = C.c_int
    L.sbd_optics_upload_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]
    L.sbd_spectrum_run.restype = C.c_int
    L.sbd_spectrum_run.argtypes = ([C.c_void_p, C.POINTER(OpticsParams)] + [C.c_void_p] * 9 +
                                   [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p] + [C.c_void_p] * 10 +
                                   [C.POINTER(InputsOut)])
    L.sbd_spectrum_run_columns.restype = C.c_int
    L.sbd_spectrum_run_columns.argtypes = ([C.c_void_p, C.POINTER(OpticsParams), C.c_int32] + [C.c_void_p] * 9 +
                                           [C.c_void_p] * 4 + [C.POINTER(C.c_int32)] + [C.c_void_p] * 4)
    L.sbd_spectrum_set_aerosols.restype = C.c_int
    L.sbd_spectrum_set_aerosols.argtypes = [C.c_void_p, C.POINTER(AerosolParams)] + [C.c_void_p] * 7
    L._spectrum_bound = True


def _ensure_tables(L, solver):
    if not getattr(solver, "_optics_tables_uploaded", False):
        tab, index = pack_tables()
        rc = L.sbd_optics_upload_tables(solver._h, tab.ctypes.data, len(tab), index.ctypes.data,
                                        len(index) // 2)
        if rc:
            raise SbdError(rc, "sbd_optics_upload_tables")
        solver._optics_tables_uploaded = True


def optics_setup(run: Sbdart):
    """struct sbd_optics_params and the wavelength-independent tables of a run."""
    p, nz, nwl = run.p, run.nz, run.nwl
    P = OpticsParams()
    P.nz, P.nwl, P.kdist, P.nstr, P.nf, P.nothrm, P.imomc = nz, nwl, run.kdist, p["nstr"], p["nf"], p["nothrm"], p["imomc"]
    ce = cloud_entries(run.clouds) if run.clouds.mcldz > 0 else np.zeros(0, dtype=CLOUD_DTYPE)
    wlalb = np.ascontiguousarray(run.albedo.wl, dtype=float)
    alb = np.ascontiguousarray(run.albedo.alb, dtype=float)
    wlsun = np.ascontiguousarray(run.sun.wl, dtype=float) if p["nf"] != 0 else np.zeros(2)
    sun = np.ascontiguousarray(run.sun.s, dtype=float) if p["nf"] != 0 else np.zeros(2)
    P.ncloud, P.nalb, P.nsun, P.night = len(ce), len(alb), len(sun), int(run.sza >= 90.)
    P.wl1, P.wl2, P.wlinc, P.amu0 = run.wl1, run.wl2, run.wlinc, run.amu0
    P.xo4, P.xrsc, P.solfac, P.phi0 = p["xo4"], p["xrsc"], p["solfac"], run.phi0
    P.fisot, P.temis, P.btemp, P.ttemp = p["fisot"], p["temis"], run.btemp, run.ttemp
    return P, ce, wlalb, alb, wlsun, sun


class ColumnRunner:
    """Whole-spectrum runs of `ncol` atmospheric columns in one call (sbd_spectrum_run_columns):
    the columns share the run's spectral grid, clouds, surface and aerosols; z / pr / t / uu may
    differ per column (default: the run's own atmosphere replicated).  Host buffers in, host
    buffers out; with `levels` the flux arrays hold those output levels only ([bin][nsel]).
    `alloc(shape, dtype)` lets the caller supply pinned memory."""

    def __init__(self, run: Sbdart, solver, ncol, levels=None, columns=None, alloc=None):
        L = lib()
        _bind(L)
        _ensure_tables(L, solver)
        self.L, self.run, self.solver, self.ncol = L, run, solver, int(ncol)
        self.P, self.ce, self.wlalb, self.alb, self.wlsun, self.sun = optics_setup(run)
        nz, nwl = run.nz, run.nwl
        alloc = alloc or (lambda shape, dtype: np.zeros(shape, dtype))
        tile = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=float), (self.ncol,) + np.shape(a)))  # noqa: E731
        if columns is None:
            self.z, self.pr, self.t, self.uu = tile(run.z), tile(run.pr), tile(run.t), tile(run.uu)
        else:
            self.z, self.pr, self.t, self.uu = (np.ascontiguousarray(columns[k], dtype=float) for k in ("z", "pr", "t", "uu"))
            self.P.btemp = self.P.ttemp = -1.0      # each column's own boundary temperatures
        assert self.z.shape == (self.ncol, nz) and self.uu.shape == (self.ncol, 64, nz + 1), (self.uu.shape, nz)
        self.levels = None if levels is None else np.ascontiguousarray(sorted(set(levels)), dtype=np.int32)
        nlev = nz + 1 if self.levels is None else len(self.levels)
        nitem = self.ncol * nwl
        self.nslot = 3 * nitem
        self.nk = alloc((nitem,), np.int32)
        self.wl, self.dwl = alloc((nwl,), np.float64), alloc((nwl,), np.float64)
        self.wt = alloc((self.nslot,), np.float64)
        self.rfldir, self.rfldn, self.flup = (alloc((self.nslot, nlev), np.float64) for _ in range(3))
        self.status = alloc((self.nslot,), np.int32)
        self.nbins = C.c_int32(0)

    def step(self):
        L, ptr = self.L, (lambda a: a.ctypes.data)
        if self.levels is not None:
            L.sbd_set_flux_levels(self.solver._h, self.levels.ctypes.data, len(self.levels))
        set_aerosols(L, self.solver, self.run.aerosols)
        try:
            rc = L.sbd_spectrum_run_columns(
                self.solver._h, C.byref(self.P), self.ncol, ptr(self.z), ptr(self.pr), ptr(self.t), ptr(self.uu),
                ptr(self.ce) if len(self.ce) else None, ptr(self.wlalb), ptr(self.alb), ptr(self.wlsun),
                ptr(self.sun), ptr(self.nk), ptr(self.wl), ptr(self.dwl), ptr(self.wt), C.byref(self.nbins),
                ptr(self.rfldir), ptr(self.rfldn), ptr(self.flup), ptr(self.status))
        finally:
            if self.levels is not None:
                L.sbd_set_flux_levels(self.solver._h, None, 0)
            if self.run.aerosols.active:
                L.sbd_spectrum_set_aerosols(self.solver._h, None, None, None, None, None, None, None, None)
        if rc:
            raise SbdError(rc, "sbd_spectrum_run_columns")
        return self.nbins.value

    def transfer_bytes(self):
        h2d, d2h = C.c_int64(0), C.c_int64(0)
        self.L.sbd_last_transfer_bytes(self.solver._h, C.byref(h2d), C.byref(d2h))
        return h2d.value, d2h.value


def run_spectrum(run: Sbdart, solver, want_inputs=False):
    """All bins of `run` produced and solved on the GPU.  Returns (rows, result[, inputs])
    shaped like Sbdart.bins() / a solve() result, in loop order."""
    L = lib()
    _bind(L)
    _ensure_tables(L, solver)
    p, nz, nwl = run.p, run.nz, run.nwl
    P, ce, wlalb, alb, wlsun, sun = optics_setup(run)
    nmom = min(p["nstr"] + 2, 40)
    nslot, NT = 3 * nwl, nz + 1
    nk = np.zeros(nwl, np.int32)
    wl, dwl, wt = np.zeros(nwl), np.zeros(nwl), np.zeros(nslot)
    nbins = C.c_int32(0)
    rfldir, rfldn, flup = (np.zeros((nslot, NT)) for _ in range(3))
    status = np.zeros(nslot, np.int32)
    numu = nphi = 0
    umu = phi = uu = None
    if run.radcalc:
        umu, phi = np.ascontiguousarray(run.umu), np.ascontiguousarray(run.phi)
        numu, nphi = len(umu), len(phi)
        uu = np.zeros((nslot, nphi, NT, numu))
    io = None
    inputs = None
    if want_inputs:
        inputs = dict(dtauc=np.zeros((nslot, nz)), ssalb=np.zeros((nslot, nz)),
                      pmom=np.zeros((nslot, nz, nmom + 1)), bins=np.zeros(nslot, dtype=BIN_DTYPE))
        io = InputsOut(inputs["dtauc"].ctypes.data, inputs["ssalb"].ctypes.data,
                       inputs["pmom"].ctypes.data, inputs["bins"].ctypes.data)
    z, pr, t = (np.ascontiguousarray(a, dtype=float) for a in (run.z, run.pr, run.t))
    uua = np.ascontiguousarray(run.uu, dtype=float)
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    # intensities only at the levels the records consume (see Sbdart.batch)
    lv = None
    if run.radcalc:
        lv = {5: [run.ntop - 1], 20: [run.ntop - 1], 6: [run.nbot - 1], 21: [run.nbot - 1],
              23: [run.ntop - 1, run.nbot - 1]}.get(run.p["iout"])
    if lv is not None:
        solver.set_radiance_levels(sorted(set(lv)))
    set_aerosols(L, solver, run.aerosols)
    try:
        rc = L.sbd_spectrum_run(
            solver._h, C.byref(P), ptr(z), ptr(pr), ptr(t), ptr(uua), ptr(ce) if len(ce) else None,
            ptr(wlalb), ptr(alb), ptr(wlsun), ptr(sun), numu, ptr(umu), nphi, ptr(phi), ptr(nk), ptr(wl),
            ptr(dwl), ptr(wt), C.byref(nbins), ptr(rfldir), ptr(rfldn), ptr(flup), ptr(uu), ptr(status),
            C.byref(io) if io is not None else None)
    finally:
        if lv is not None:
            solver.set_radiance_levels(None)
        if run.aerosols.active:
            L.sbd_spectrum_set_aerosols(solver._h, None, None, None, None, None, None, None, None)
    if rc:
        raise SbdError(rc, "sbd_spectrum_run")
    B = nbins.value
    rows = []
    slots = []
    for il in range(nwl):
        for kd in range(nk[il]):
            rows.append(dict(il=il, kd=kd, nk=int(nk[il]), wl=wl[il], dwl=dwl[il], wt=wt[3 * il + kd],
                             ff=run.filter(wl[il])))
            slots.append(3 * il + kd)
    res = dict(rfldir=rfldir[:B], rfldn=rfldn[:B], flup=flup[:B], status=status[:B])
    if uu is not None:
        res["uu"] = uu[:B]
    if want_inputs:
        sl = np.array(slots)
        inputs = {k: v[sl] for k, v in inputs.items()}
        return rows, res, inputs
    return rows, res
