"""sbdart_b200 -- B200-native batched discrete-ordinate solver behind SBDART's
DISORT call (reference drt.f:541-546, disort.f:1-6).

This package is a thin host mirror of the C ABI in include/sbdart_b200.h.
All computation happens in hand-written CUDA (sbdart_b200/csrc); there is no
CPU fallback: loading fails loudly when the CUDA library is missing and every
solve raises when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# SBD_LIB_PATH: an alternative build of the same sources (tuning experiments)
LIB_PATH = os.environ.get("SBD_LIB_PATH") or os.path.join(_PKG, "libsbdart_b200.so")

SBD_SUCCESS = 0
SBD_ERR_CUDA = -100
SBD_ERR_ARG = -101
SBD_ERR_UNSUPPORTED = -102
BIN_OK = 0
BIN_ANGLE_CLASH = 1
BIN_BAD_INPUT = -1
BIN_EIG_FAIL = -2
BIN_SINGULAR = -3

EXPORTS = (
    "sbd_create", "sbd_destroy", "sbd_disort_batch", "sbd_disort_batch_device",
    "sbd_synchronize", "sbd_stream", "sbd_kernel_launches", "sbd_quadrature",
    "sbd_status_string", "sbd_abi_version", "disort_", "sbd_disort_last_status",
    "sbd_measure_fp64_peak", "sbd_optics_upload_tables", "sbd_spectrum_run",
    "sbd_set_radiance_levels", "sbd_spectrum_set_aerosols", "sbd_set_corint", "sbd_build_id", "sbd_set_radiance_layout", "sbd_set_flux_levels",
    "sbd_spectrum_run_columns", "sbd_last_transfer_bytes", "sbd_spectrum_device_fluxes",
    "sbd_set_surfaces", "sbd_set_bdref_callback",
)


class SbdDims(C.Structure):
    """struct sbd_dims (include/sbdart_b200.h)."""

    _fields_ = [(k, C.c_int32) for k in (
        "nbins", "nlyr", "nstr", "nmom", "ntau", "numu", "nphi", "ncol")]


class SbdBin(C.Structure):
    """struct sbd_bin (include/sbdart_b200.h)."""

    _fields_ = [(k, C.c_double) for k in (
        "fbeam", "umu0", "phi0", "fisot", "albedo", "btemp", "ttemp", "temis",
        "wvnmlo", "wvnmhi", "accur")] + [("plank", C.c_int32), ("col", C.c_int32)]


BIN_DTYPE = np.dtype([(k, "<f8") for k in (
    "fbeam", "umu0", "phi0", "fisot", "albedo", "btemp", "ttemp", "temis",
    "wvnmlo", "wvnmhi", "accur")] + [("plank", "<i4"), ("col", "<i4")])
assert BIN_DTYPE.itemsize == C.sizeof(SbdBin)


class SbdError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().sbd_status_string(code).decode() if _LIB is not None else str(code)
        super().__init__(f"{where}: {msg} ({code})")


_LIB = None


def lib():
    """Load the CUDA library.  No fallback: a missing build is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.sbd_build_id.restype = C.c_char_p
    if os.path.isdir(os.path.join(_PKG, "csrc")) and not os.environ.get("SBD_SKIP_BUILD_ID_CHECK"):
        from . import _build
        have, want = L.sbd_build_id().decode(), _build.source_id()
        if have != want:
            raise ImportError(
                f"{LIB_PATH} was built from other sources (build id {have}, sources {want}); "
                "rebuild with `python -c 'import __graft_entry__ as g; g.build()'`")
    dp, ip = C.c_void_p, C.c_void_p
    L.sbd_create.restype = C.c_int
    L.sbd_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.sbd_destroy.argtypes = [C.c_void_p]
    L.sbd_destroy.restype = None
    batch_args = [C.c_void_p, C.POINTER(SbdDims)] + [dp] * 3 + [C.c_void_p] + [dp] * 4 + [dp] * 6 + [ip]
    L.sbd_disort_batch.restype = C.c_int
    L.sbd_disort_batch.argtypes = batch_args
    L.sbd_disort_batch_device.restype = C.c_int
    L.sbd_disort_batch_device.argtypes = batch_args + [C.c_void_p]
    L.sbd_synchronize.restype = C.c_int
    L.sbd_synchronize.argtypes = [C.c_void_p]
    L.sbd_stream.restype = C.c_void_p
    L.sbd_stream.argtypes = [C.c_void_p]
    L.sbd_kernel_launches.restype = C.c_int64
    L.sbd_kernel_launches.argtypes = [C.c_void_p]
    L.sbd_quadrature.restype = C.c_int
    L.sbd_quadrature.argtypes = [C.c_int, dp, dp]
    L.sbd_status_string.restype = C.c_char_p
    L.sbd_status_string.argtypes = [C.c_int]
    L.sbd_abi_version.restype = C.c_int
    L.sbd_set_corint.restype = C.c_int
    L.sbd_set_corint.argtypes = [C.c_void_p, C.c_int32]
    L.sbd_set_radiance_levels.restype = C.c_int
    L.sbd_set_radiance_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.sbd_set_flux_levels.restype = C.c_int
    L.sbd_set_flux_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.sbd_last_transfer_bytes.restype = C.c_int
    L.sbd_last_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.sbd_set_radiance_layout.restype = C.c_int
    L.sbd_set_radiance_layout.argtypes = [C.c_void_p, C.c_int32]
    L.sbd_measure_fp64_peak.restype = C.c_int
    L.sbd_measure_fp64_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.sbd_set_surfaces.restype = C.c_int
    L.sbd_set_surfaces.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 4
    L.sbd_set_bdref_callback.restype = None
    L.sbd_set_bdref_callback.argtypes = [C.c_void_p]
    L.sbd_disort_last_status.restype = C.c_int
    L.disort_.restype = None
    _LIB = L
    return L


def quadrature(m: int):
    """Gauss-Legendre nodes/weights on (0,1) for NSTR = 2*m (QGAUSN, disort.f:5984)."""
    mu = np.zeros(m)
    wt = np.zeros(m)
    rc = lib().sbd_quadrature(m, mu.ctypes.data, wt.ctypes.data)
    if rc:
        raise SbdError(rc, "sbd_quadrature")
    return mu, wt


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def surface_albedo(s):
    """The `albedo` value that selects BRDF surface s (SBD_SURFACE, include/sbdart_b200.h)."""
    return -(np.asarray(s, dtype=np.float64) + 1.0)


def make_bins(nbins, *, fbeam=0.0, umu0=1.0, phi0=0.0, fisot=0.0, albedo=0.0, btemp=0.0,
              ttemp=0.0, temis=0.0, wvnmlo=0.0, wvnmhi=0.0, accur=0.0, plank=0, col=0):
    """Array of struct sbd_bin; every argument is a scalar or a length-B array."""
    b = np.zeros(nbins, dtype=BIN_DTYPE)
    for k, v in dict(fbeam=fbeam, umu0=umu0, phi0=phi0, fisot=fisot, albedo=albedo, btemp=btemp,
                     ttemp=ttemp, temis=temis, wvnmlo=wvnmlo, wvnmhi=wvnmhi, accur=accur, plank=plank,
                     col=col).items():
        b[k] = v
    return b


class Solver:
    """Owns one sbd_handle (a CUDA stream plus device scratch) on `device`."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().sbd_create(C.byref(self._h), device)
        if rc:
            self._h = None
            raise SbdError(rc, "sbd_create")
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().sbd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(lib().sbd_kernel_launches(self._h))

    @property
    def stream(self) -> int:
        return int(lib().sbd_stream(self._h) or 0)

    def measure_fp64_peak(self, reps: int = 5) -> float:
        """Sustained DFMA throughput of this device in TFLOP/s (roofline denominator)."""
        v = C.c_double(0.0)
        rc = lib().sbd_measure_fp64_peak(self._h, reps, C.byref(v))
        if rc:
            raise SbdError(rc, "sbd_measure_fp64_peak")
        return v.value

    def synchronize(self):
        rc = lib().sbd_synchronize(self._h)
        if rc:
            raise SbdError(rc, "sbd_synchronize")

    def set_surfaces(self, nstr=0, bdr=None, bem=None, rmu=None, emu=None):
        """BRDF surfaces of the following calls (LAMBER = .FALSE.): SURFAC's tables
        bdr [ns][modes][n][n+1], bem [ns][n], and for radiance runs rmu [ns][modes][numu][n+1],
        emu [ns][numu] (sbdart_b200.frontend.brdf.surface_tables builds them).  A bin selects
        surface s with albedo = surface_albedo(s).  No arguments: back to Lambertian surfaces."""
        if bdr is None:
            rc = lib().sbd_set_surfaces(self._h, 0, 0, 0, 0, None, None, None, None)
        else:
            bdr, bem = _f64(bdr), _f64(bem)
            ns, modes = bdr.shape[0], bdr.shape[1]
            numu = 0 if rmu is None else np.shape(rmu)[2]
            rmu = None if rmu is None else _f64(rmu)
            emu = None if emu is None else _f64(emu)
            rc = lib().sbd_set_surfaces(self._h, ns, nstr, modes, numu, bdr.ctypes.data, bem.ctypes.data,
                                        None if rmu is None else rmu.ctypes.data,
                                        None if emu is None else emu.ctypes.data)
        if rc:
            raise SbdError(rc, "sbd_set_surfaces")

    # -- host-buffer call: the reference-facing path (H2D + kernel + D2H) ----
    def set_radiance_levels(self, levels=None):
        """Restrict the intensities (uu) of the following calls to these output levels
        (None / empty: all levels).  SBDART prints one or two levels only."""
        lv = None if levels is None else np.ascontiguousarray(levels, dtype=np.int32)
        rc = lib().sbd_set_radiance_levels(self._h, None if lv is None or lv.size == 0 else lv.ctypes.data,
                                           0 if lv is None else int(lv.size))
        if rc:
            raise SbdError(rc, "sbd_set_radiance_levels")

    def disort_batch(self, dtauc, ssalb, pmom, bins, *, nstr, temper=None, utau=None,
                     umu=None, phi=None, out=None, uu_levels=None, corint=False, uu_packed=False):
        """Batched DISORT on host arrays.

        dtauc, ssalb [B][L]; pmom [B][L][nmom+1]; bins = make_bins(...);
        temper [ncol][L+1]; utau [B][ntau] (USRTAU) or None (layer boundaries).
        Returns dict(rfldir, rfldn, flup, dfdt, uavg [B][NT], status [B]) and, for radiance
        runs, uu [B][nphi][NT][numu]; with uu_levels and uu_packed=True uu holds the selected
        levels only, ascending ([B][nphi][nsel][numu], returned with "uu_levels").
        """
        dtauc, ssalb, pmom = _f64(dtauc), _f64(ssalb), _f64(pmom)
        B, L = dtauc.shape
        assert ssalb.shape == (B, L) and pmom.shape[:2] == (B, L)
        bins = np.ascontiguousarray(bins, dtype=BIN_DTYPE)
        assert bins.shape == (B,)
        d = SbdDims()
        d.nbins, d.nlyr, d.nstr, d.nmom = B, L, nstr, pmom.shape[2] - 1
        tp = None
        if temper is not None:
            tp = _f64(np.atleast_2d(temper))
            assert tp.shape[1] == L + 1
            d.ncol = tp.shape[0]
        ut = None
        if utau is not None:
            ut = _f64(utau)
            assert ut.shape[0] == B
            d.ntau = ut.shape[1]
        um = None if umu is None else _f64(umu)
        ph = None if phi is None else _f64(phi)
        d.numu = 0 if um is None else um.shape[0]
        d.nphi = 0 if ph is None else ph.shape[0]
        NT = d.ntau if d.ntau > 0 else L + 1
        if out is None:
            out = {k: np.empty((B, NT)) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
            out["status"] = np.empty(B, np.int32)
            if d.numu > 0 and uu_levels is not None and uu_packed:
                lv = sorted(set(int(v) for v in uu_levels))
                out["uu"] = np.empty((B, d.nphi, len(lv), d.numu))
                out["uu_levels"] = lv
            elif d.numu > 0:
                # selected levels only are copied back (sbd_set_radiance_levels): the rest stays zero
                out["uu"] = (np.zeros if uu_levels is not None else np.empty)((B, d.nphi, NT, d.numu))
        p = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        if uu_levels is not None:
            self.set_radiance_levels(uu_levels)
        if corint:          # CORINT=.TRUE.: Nakajima-Tanaka corrections after the solve
            lib().sbd_set_corint(self._h, 1)
        packed = bool(uu_packed and uu_levels is not None and d.numu > 0)
        if packed:
            lib().sbd_set_radiance_layout(self._h, 1)
        try:
            rc = lib().sbd_disort_batch(
                self._h, C.byref(d), p(dtauc), p(ssalb), p(pmom), p(bins), p(tp), p(ut), p(um),
                p(ph), p(out["rfldir"]), p(out["rfldn"]), p(out["flup"]), p(out["dfdt"]),
                p(out["uavg"]), p(out.get("uu")), p(out["status"]))
        finally:
            if uu_levels is not None:
                self.set_radiance_levels(None)
            if corint:
                lib().sbd_set_corint(self._h, 0)
            if packed:
                lib().sbd_set_radiance_layout(self._h, 0)
        if rc:
            raise SbdError(rc, "sbd_disort_batch")
        return out

    # -- device-pointer call: inputs already resident in HBM -----------------
    def disort_batch_device(self, dims: SbdDims, ptrs: dict, stream: int | None = None):
        """Enqueue a batch whose buffers are device pointers (ints), e.g.
        torch tensors' .data_ptr().  Keys: dtauc ssalb pmom bins temper utau umu
        phi rfldir rfldn flup dfdt uavg uu status.  Does not synchronise."""
        g = lambda k: ptrs.get(k) or None  # noqa: E731
        rc = lib().sbd_disort_batch_device(
            self._h, C.byref(dims), g("dtauc"), g("ssalb"), g("pmom"), g("bins"), g("temper"),
            g("utau"), g("umu"), g("phi"), g("rfldir"), g("rfldn"), g("flup"), g("dfdt"),
            g("uavg"), g("uu"), g("status"), stream)
        if rc:
            raise SbdError(rc, "sbd_disort_batch_device")


def disort(nlyr, dtauc, ssalb, nmom, pmom, temper, wvnmlo, wvnmhi, usrtau, ntau, utau, nstr,
           usrang, numu, umu, nphi, phi, ibcnd, fbeam, umu0, phi0, fisot, lamber, albedo, btemp,
           ttemp, temis, plank, onlyfl, accur=0.0, corint=False, maxcly=None, maxulv=None,
           maxumu=None, maxphi=None, maxmom=None):
    """Host mirror of SUBROUTINE DISORT (disort.f:1-6) through the gfortran-ABI
    entry `disort_`: same argument names and meaning; arrays are numpy, Fortran
    order where 2-D (pmom is (0:maxmom, maxcly)).  Returns a dict with the output
    arrays plus the mutated nstr / ntau / utau / numu / umu and `status`."""
    L = int(nlyr)
    maxcly = maxcly or L
    maxulv = maxulv or L + 1
    maxumu = maxumu or max(int(nstr), int(numu), 1)
    maxphi = maxphi or max(int(nphi), 1)
    pm = np.asfortranarray(pmom, dtype=np.float64)
    maxmom = maxmom if maxmom is not None else pm.shape[0] - 1
    assert pm.shape[0] == maxmom + 1
    ci = lambda v: C.byref(C.c_int(int(v)))  # noqa: E731
    cd = lambda v: C.byref(C.c_double(float(v)))  # noqa: E731
    dt = np.zeros(maxcly); dt[:L] = dtauc
    ss = np.zeros(maxcly); ss[:L] = ssalb
    tp = np.zeros(maxcly + 1)
    if temper is not None:
        tp[:L + 1] = temper
    ut = np.zeros(maxulv)
    if usrtau:
        ut[:ntau] = utau
    um = np.zeros(maxumu)
    if usrang:
        um[:numu] = umu
    ph = np.zeros(maxphi)
    if nphi:
        ph[:nphi] = phi
    o = {k: np.zeros(maxulv) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
    uu = np.zeros((maxumu, maxulv, maxphi), order="F")
    albmed, trnmed = np.zeros(maxumu), np.zeros(maxumu)
    c_nstr, c_ntau, c_numu = C.c_int(int(nstr)), C.c_int(int(ntau)), C.c_int(int(numu))
    prnt = (C.c_int * 7)()
    header = C.create_string_buffer(b" " * 127, 127)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().disort_(
        ci(L), P(dt), P(ss), ci(corint), ci(nmom), P(pm), P(tp), cd(wvnmlo), cd(wvnmhi),
        ci(usrtau), C.byref(c_ntau), P(ut), C.byref(c_nstr), ci(usrang), C.byref(c_numu), P(um),
        ci(nphi), P(ph), ci(ibcnd), cd(fbeam), cd(umu0), cd(phi0), cd(fisot), ci(lamber),
        cd(albedo), cd(btemp), cd(ttemp), cd(temis), ci(plank), ci(onlyfl), cd(accur), prnt,
        header, ci(maxcly), ci(maxulv), ci(maxumu), ci(maxphi), ci(maxmom), P(o["rfldir"]),
        P(o["rfldn"]), P(o["flup"]), P(o["dfdt"]), P(o["uavg"]), P(uu), P(albmed), P(trnmed),
        C.c_size_t(127))
    o.update(uu=uu, nstr=c_nstr.value, ntau=c_ntau.value, utau=ut, numu=c_numu.value, umu=um,
             dtauc=dt, ssalb=ss, pmom=pm, status=lib().sbd_disort_last_status())
    return o
