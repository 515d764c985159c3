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
