"""NAMELIST reader and record formatting of the front end (drt.f:200-231, :991-1163):
array subscripts, terminators, unknown names, non-finite values, unsupported options."""
import numpy as np
import pytest

from sbdart_b200.frontend import INPUT_NAMES, Sbdart, _es, parse_namelist


def test_array_elements_keep_their_subscript():
    r = Sbdart("&INPUT tcloud(2)=5, zcloud(2)=3, nre(2)=12,16, iout=10 /")
    assert list(r.p["tcloud"]) == [0, 5, 0, 0, 0]
    assert list(r.p["zcloud"]) == [0, 3, 0, 0, 0]
    assert list(r.p["nre"]) == [8, 12, 16, 8, 8]
    # whole-array assignment fills the leading elements, repeat counts expand
    r = Sbdart("&INPUT tcloud=2*4.,1, zcloud=1,2,3 /")
    assert list(r.p["tcloud"]) == [4, 4, 1, 0, 0]
    with pytest.raises(ValueError, match="out of range"):
        Sbdart("&INPUT tcloud(6)=1 /")


@pytest.mark.parametrize("text", ["&INPUT iout=1/", "$INPUT iout=1 $END", "&input\n iout = 1\n/\n",
                                  " &INPUT iout=1, ! comment\n &end"])
def test_terminators(text):
    assert parse_namelist(text, "input", INPUT_NAMES) == [("iout", None, [1.0])]


def test_second_group_and_unknown_names():
    r = Sbdart("&INPUT iout=10 /\n&DINPUT fisot=2.5, btemp=280 /")
    assert r.p["fisot"] == 2.5 and r.p["btemp"] == 280
    with pytest.raises(ValueError, match="unknown variable 'bogus'"):
        Sbdart("&INPUT bogus=3 /")
    with pytest.raises(ValueError, match="unknown variable 'idatm'"):      # not a DINPUT name
        Sbdart("&INPUT iout=10 /\n&DINPUT idatm=2 /")
    with pytest.raises(ValueError, match="not found"):
        Sbdart("&OTHER iout=10 /")


def test_non_finite_values_are_printed_as_gfortran_prints_them():
    assert _es(float("nan"), 12, 4) == "         NaN"
    assert _es(float("inf"), 12, 4) == "    Infinity"
    assert _es(float("-inf"), 12, 4) == "   -Infinity"
    assert _es(float("-inf"), 8, 2) == "    -Inf"
    assert _es(0.0, 12, 4) == "  0.0000E+00" and _es(-3.599e-18, 12, 4) == " -3.5990E-18"
    assert _es(1e-100, 12, 4) == "  1.0000-100"

    def solve_nan(b):          # a solver that fails numerically must show up in the record
        n, nt = len(b["bins"]), b["dtauc"].shape[1] + 1
        o = {k: np.full((n, nt), np.nan) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
        o["status"] = np.zeros(n, np.int32)
        return o
    assert "NaN" in Sbdart("&INPUT iout=10 /").run(solve_nan)


@pytest.mark.parametrize("nl,what", [("kdist=-2", "kdist"), ("isalb=-7", "dref")])
def test_options_outside_the_front_end_raise(nl, what):
    with pytest.raises(NotImplementedError, match=what):
        Sbdart(f"&INPUT {nl} /")
    with pytest.raises(NotImplementedError, match="ibcnd"):
        Sbdart("&INPUT iout=10 /\n&DINPUT ibcnd=1 /")


def test_warning_files_follow_errmsg(tmp_path):
    """SBDART_WARNING.nn (errmsg, disutil.f:278-325): message, rule of 70 '#', copy of INPUT."""
    from sbdart_b200.frontend import write_warning_files
    from solvers import solve_oracle
    text = "&INPUT\n vis=23   \n iout=10\n/\n"
    r = Sbdart(text)
    assert r.warnings == [(16, "CHKIN--IAER=0, though VIS or TBAER set")]
    paths = write_warning_files(r.warnings, text, str(tmp_path))
    assert [p.rsplit("/", 1)[1] for p in paths] == ["SBDART_WARNING.16"]
    got = open(paths[0]).read()
    assert got == ("WARNING >>>>> CHKIN--IAER=0, though VIS or TBAER set\n\n" + "#" * 70 + "\n\n"
                   "&INPUT\n vis=23\n iout=10\n/\n")
    # CHEKIN's temperature-step warning (thermal run) and the intensity-correction one
    r = Sbdart("&INPUT wlinf=10, wlsup=10, iout=20, uzen=30, nstr=4 /")
    r.run(solve_oracle)
    nums = [n for n, _ in r.warnings]
    assert 6 in nums and 7 in nums and nums.count(6) == 1
    # albedo table left on the long-wave side: SALBEDO warning, end value used (spectra.f:44-56)
    r = Sbdart("&INPUT isalb=4, wlinf=4.5, wlsup=4.5, iout=10, sza=95 /")
    r.run(solve_oracle)
    assert any(n == 18 and "wlsup gt     4.000" in m for n, m in r.warnings)


def test_failed_bins_are_fatal_like_errmsg_0():
    from sbdart_b200.frontend import SbdartFatal, warning_file_text

    def solve_bad(b):
        n, nt = len(b["bins"]), b["dtauc"].shape[1] + 1
        o = {k: np.zeros((n, nt)) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
        o["status"] = np.full(n, -2, np.int32)
        return o
    with pytest.raises(SbdartFatal, match="ASYMTX--convergence problems"):
        Sbdart("&INPUT iout=10 /").run(solve_bad)
    assert warning_file_text(0, "x", "a  \n").startswith("ERROR  >>>>>> x\n\n###")
