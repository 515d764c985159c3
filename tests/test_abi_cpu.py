"""CPU-side checks: the C-ABI library loads, exports every symbol declared in
include/sbdart_b200.h, and its host-only helpers agree with the oracle.  No
compute entry is exercised here (no GPU in this tier)."""
import os
import re

import numpy as np

import sbdart_b200 as sb
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sbdart_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|int64_t|char)\s*\*?\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = sb.lib()
    names = declared_symbols()
    assert set(names) == set(sb.EXPORTS), (names, sb.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.sbd_abi_version() == 1


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(sb.SbdDims) == 8 * 4
    assert C.sizeof(sb.SbdBin) == 11 * 8 + 2 * 4
    assert sb.BIN_DTYPE.itemsize == C.sizeof(sb.SbdBin)
    b = sb.make_bins(3, fbeam=[1, 2, 3], umu0=0.5, plank=[0, 1, 0])
    raw = np.frombuffer(b.tobytes(), dtype=np.uint8).reshape(3, -1)
    one = sb.SbdBin.from_buffer_copy(raw[1].tobytes())
    assert one.fbeam == 2.0 and one.umu0 == 0.5 and one.plank == 1


def test_host_quadrature_matches_oracle_qgausn():
    import ctypes as C
    dp = C.POINTER(C.c_double)
    for m in (2, 4, 5, 8, 10, 16, 20):
        mu, wt = sb.quadrature(m)
        mo, wo = np.zeros(m), np.zeros(m)
        oracle.lib().sbdo_qgausn(m, mo.ctypes.data_as(dp), wo.ctypes.data_as(dp))
        np.testing.assert_allclose(mu, mo, rtol=0, atol=4e-16)
        np.testing.assert_allclose(wt, wo, rtol=2e-13, atol=0)


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        return
    try:
        sb.Solver(0)
    except sb.SbdError as e:
        assert e.code == sb.SBD_ERR_CUDA
    else:
        raise AssertionError("Solver() must fail without a CUDA device")


def test_status_strings():
    lib = sb.lib()
    for code in (0, 1, -1, -2, -3, -100, -101, -102):
        assert lib.sbd_status_string(code)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sbdart_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), f


def test_aerosol_structs_match_header():
    """sbd_aerosol_params / sbd_strat_entry (include/sbdart_b200.h) as packed by frontend/device.py."""
    import ctypes as C
    from sbdart_b200.frontend import device
    assert C.sizeof(device.AerosolParams) == 6 * 4 + 8
    assert device.STRAT_DTYPE.itemsize == (2 + 3 * 47) * 8
    src = open(os.path.join(ROOT, "include", "sbdart_b200.h")).read()
    assert "#define SBD_NAERW 47" in src and "double ext[SBD_NAERW], absb[SBD_NAERW], asym[SBD_NAERW]" in src


def test_aerosol_setup_handed_to_the_producer_kernel():
    """Host logic of device.set_aerosols: which arrays reach sbd_spectrum_set_aerosols for a
    boundary-layer + stratospheric run, and that runs without aerosols clear the setting."""
    import ctypes as C
    from sbdart_b200.frontend import Sbdart, device

    class FakeLib:
        def __init__(self):
            self.calls = []

        def sbd_spectrum_set_aerosols(self, h, p, *arrays):
            self.calls.append((p, arrays))
            return 0

    class FakeSolver:
        _h = None

    run = Sbdart("&INPUT idatm=2, iaer=2, vis=10, jaer=2,4, zaer=18,25, taerst=0.1,0.02, wlinf=.5, wlsup=.6, wlinc=.05 /")
    L = FakeLib()
    device.set_aerosols(L, FakeSolver(), run.aerosols)
    p, arrays = L.calls[0]
    par = C.cast(p, C.POINTER(device.AerosolParams)).contents if not isinstance(p, device.AerosolParams) else p
    par = p._obj if hasattr(p, "_obj") else par
    assert par.nwlbaer == 47 and par.nstrat == 2 and par.nz == run.nz and par.imoma == 3
    assert all(a is not None for a in arrays)                 # wlb, ext, abs, asm, dtsv, awl, strat
    assert device.device_aerosols_supported(run.aerosols)
    clean = Sbdart("&INPUT idatm=2, wlinf=.5, wlsup=.6, wlinc=.05 /")
    device.set_aerosols(L, FakeSolver(), clean.aerosols)
    assert L.calls[1][0] is None
    table = Sbdart("&INPUT idatm=2, iaer=1, vis=23, imoma=4, wlinf=.5, wlsup=.6, wlinc=.05 /")
    assert not device.device_aerosols_supported(table.aerosols)
