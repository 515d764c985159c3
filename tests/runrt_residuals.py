"""Per-column deviation of this code (front end + CPU checker) from a RunRT sweep of the
reference, and the extra vertical optical depth its direct beam implies.

    python tests/runrt_residuals.py btemp_uw_iout_1 0 [extra namelist text]

Used to characterise the sweeps that another build of sbdart wrote (see the docstring of
tests/test_runrt_golden.py).  Reads tests/golden/runrt/<name>.sbd, or RUNRT_DIR if set
(e.g. the reference's RunRT/RUNS)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import runrt_cases                                   # noqa: E402
from sbdart_b200.frontend import Sbdart             # noqa: E402
from solvers import solve_oracle                    # noqa: E402

if os.environ.get("RUNRT_DIR"):
    runrt_cases.GOLDEN = os.environ["RUNRT_DIR"]


def main():
    name, run = sys.argv[1], int(sys.argv[2])
    extra = sys.argv[3] if len(sys.argv) > 3 else ""
    inputs, outputs = runrt_cases.parse_sbd(name)
    nl = inputs[run].replace("\n/", "\n " + extra + "\n/")
    print(nl.replace("\n", " "))
    got = Sbdart(nl).run(solve_oracle)
    print("      wl    wvnm   ours/ref: topdn topup topdir botdn botup botdir | extra depth (ref - ours)")
    for x, y in zip(got.splitlines(), outputs[run].splitlines()):
        tx, ty = x.split(), y.split()
        if len(tx) != 8:
            continue
        vx, vy = np.array([float(t) for t in tx]), np.array([float(t) for t in ty])
        with np.errstate(all="ignore"):
            r = np.where(vy != 0, vx / vy, np.nan)
            dtau = np.log(vx[7] / vx[4]) - np.log(vy[7] / vy[4])
        print("%9.4f %8.2f " % (vy[0], 1e4 / vy[0]), " ".join("%7.4f" % v for v in r[2:]), " | %8.4f" % dtau)


if __name__ == "__main__":
    main()
