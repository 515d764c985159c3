"""The five TestRuns examples of the reference (TestRuns/test_runs:29-146) as
NAMELIST texts, and a tolerant comparison of IOUT records with the shipped
golden outputs (tests/golden/sbchk.N, copied verbatim from TestRuns/)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_inputs(n):
    if n == 1:
        return ["&INPUT\n idatm=4, isat=0, wlinf=.25, wlsup=1.0, wlinc=.005, iout=1,\n /"]
    if n == 2:
        return [f"&INPUT\n tcloud={tc}\n albcon={al}\n idatm=4\n isat=0\n wlinf=.55\n wlsup=.55\n isalb=0\n"
                f" iout=10\n sza=30\n /"
                for al in ("0", ".2", ".4", ".6", ".8", "1") for tc in (0, 1, 2, 4, 8, 16, 32, 64)]
    if n == 3:
        return [f"&INPUT\n tcloud={tc}\n zcloud=8\n nre=10\n idatm=4\n sza=95\n wlinf=4\n wlsup=20\n"
                f" wlinc=-.01\n iout=1\n /" for tc in (0, 1, 5)]
    if n == 4:
        return [f"&INPUT\n tcloud={tc}\n nre={nre}\n wlinf={wl}\n wlsup={wl}\n idatm=1\n isat=0\n isalb=6\n"
                f" iout=10\n sza=0\n /"
                for tc in (0, 1, 2, 4, 8, 16, 32, 64, 128) for nre in (2, 4, 8, 16, 32, 64, 128)
                for wl in (".55", "2.16")]
    if n == 5:
        return [f"&INPUT\n tcloud= {tc}\n zcloud= 1\n wlinf=.72\n wlsup=.72\n idatm=1\n isalb=4\n sza=60\n"
                f" iout=21\n nstr=20\n uzen=5,15,25,35,45,55,65,75,85,95,105,115,125,135,145,155,165,175\n"
                f" phi=0,15,30,45,60,75,90,105,120,135,150,165,180\n /" for tc in (5, 15)]
    raise ValueError(n)


def golden_text(n):
    return open(os.path.join(GOLDEN, f"sbchk.{n}")).read()


def compare_records(got, ref, rel=1.5e-4, noise=1e-6):
    """Line-by-line numeric comparison.  The goldens carry 5 significant digits
    (ES12.4), i.e. up to 5e-5 relative quantisation on each side; values below
    `noise` x the largest value of the record are round-off (e.g. upward flux
    over a black surface, -3.6E-18 in sbchk.1) and only need to be small.
    Returns (n_values, n_exact_strings, worst_relative_error)."""
    a, b = got.splitlines(), ref.splitlines()
    assert len(a) == len(b), (len(a), len(b))
    nval = nexact = 0
    worst = 0.0
    for i, (x, y) in enumerate(zip(a, b)):
        tx, ty = x.split(), y.split()
        assert len(tx) == len(ty), (i, x, y)
        try:
            vy = [float(t) for t in ty]
            vx = [float(t) for t in tx]
        except ValueError:
            assert x.strip() == y.strip(), (i, x, y)
            continue
        scale = max([abs(v) for v in vy] + [1e-300])
        for sx, sy, u, v in zip(tx, ty, vx, vy):
            nval += 1
            nexact += sx == sy
            if abs(v) < noise * scale:
                assert abs(u) < 10 * noise * scale, (i, x, y)
                continue
            err = abs(u - v) / abs(v)
            worst = max(worst, err)
            assert err <= rel, (i, sx, sy, x, y)
    return nval, nexact, worst
