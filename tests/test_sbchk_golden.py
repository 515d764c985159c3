"""End-to-end pin against the reference's own golden outputs.

TestRuns/sbchk.1-5 (copied verbatim to tests/golden/) are the regression
vectors the reference ships (TestRuns/test_runs:6).  The SBDART front end
(sbdart_b200/frontend) produces the per-bin DISORT inputs and formats the IOUT
records; the solve is done (a) by the CPU oracle -- which pins the oracle and
the front end together -- and (b) on the GPU through the C ABI.
Goldens carry 5 significant digits; see sbchk_cases.compare_records.
"""
import pytest

import sbdart_b200 as sb
from sbdart_b200.frontend import Sbdart
from sbchk_cases import case_inputs, compare_records, golden_text
from solvers import make_solve_cuda, solve_oracle


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_oracle_reproduces_sbchk(n):
    txt = "".join(Sbdart(nl).run(solve_oracle) for nl in case_inputs(n))
    nval, nexact, worst = compare_records(txt, golden_text(n))
    assert nval > 400 and worst <= 1.5e-4
    assert nexact >= 0.85 * nval        # the rest differ in the last printed digit or are round-off noise


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_cuda_reproduces_sbchk(n):
    s = sb.Solver(0)
    solve = make_solve_cuda(s)
    txt = "".join(Sbdart(nl).run(solve) for nl in case_inputs(n))
    nval, nexact, worst = compare_records(txt, golden_text(n))
    assert worst <= 1.5e-4 and nexact >= 0.85 * nval
    assert s.kernel_launches >= len(case_inputs(n))
    s.close()


@pytest.mark.gpu
def test_sbdart_executable_drop_in(tmp_path):
    """bin/sbdart: ./INPUT in, IOUT records on stdout (TestRuns/test_runs:29-37)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "INPUT").write_text(case_inputs(1)[0])
    r = subprocess.run([sys.executable, os.path.join(root, "bin", "sbdart")], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    nval, nexact, worst = compare_records(r.stdout, golden_text(1))
    assert worst <= 1.5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_whole_spectrum_gpu_path_reproduces_sbchk(n):
    """K2 (optical properties on the GPU) + K1, nothing but the setup crosses PCIe."""
    s = sb.Solver(0)
    txt = "".join(Sbdart(nl).run_device(s) for nl in case_inputs(n))
    nval, nexact, worst = compare_records(txt, golden_text(n))
    assert worst <= 1.5e-4 and nexact >= 0.85 * nval
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nl", [
    "&INPUT idatm=2, wlinf=.25, wlsup=4.0, wlinc=.005, nstr=16, iout=1 /",
    "&INPUT idatm=4, wlinf=4, wlsup=20, wlinc=-.01, sza=95, tcloud=5, zcloud=8, nre=10, iout=1 /",
    "&INPUT idatm=1, wlinf=.3, wlsup=3.0, wlinc=.05, sza=40, lwp=20,50,0, zcloud=1,4,0, nre=6,-20,8, isalb=6, kdist=2, iout=1 /",
    "&INPUT idatm=6, wlinf=5, wlsup=50, wlinc=20, sza=20, kdist=1, nothrm=0, iout=1 /",
])
def test_producer_kernel_matches_host_front_end(nl):
    """Per-bin DISORT inputs from the K2 kernel vs the host front end (which the
    sbchk goldens pin): optical depths, single-scattering albedos, moments, scalars."""
    import numpy as np
    from sbdart_b200.frontend.device import run_spectrum
    s = sb.Solver(0)
    run = Sbdart(nl)
    ref = run.batch(run.bins())
    rows, res, dev = run_spectrum(Sbdart(nl), s, want_inputs=True)
    assert len(rows) == len(ref["bins"])
    for k in ("dtauc", "ssalb", "pmom"):
        scale = np.abs(ref[k]).max(axis=tuple(range(1, ref[k].ndim)), keepdims=True)
        np.testing.assert_allclose(dev[k], ref[k], rtol=1e-9, atol=1e-13 * scale.max())
    for f in ("fbeam", "umu0", "albedo", "btemp", "ttemp", "temis", "wvnmlo", "wvnmhi", "fisot", "phi0"):
        np.testing.assert_allclose(dev["bins"][f], ref["bins"][f], rtol=1e-12, atol=0)
    assert (dev["bins"]["plank"] == ref["bins"]["plank"]).all()
    s.close()
