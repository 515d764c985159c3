"""RunRT/RUNS/*.sbd of the reference: parameter sweeps stored TOGETHER WITH the
outputs the reference produced (everything after the `_DATA_` line).  They are
golden vectors for the front end + solver beyond TestRuns/sbchk.1-5 (SURVEY 8c,
item 3): clouds at several heights and drop sizes, solar zenith angle, surface
albedo, surface pressure, water vapour / ozone / CO2 amounts, surface temperature in
the thermal window, iout = 1, 10, 11, 21.

File format (RunRT/RunRT.py:1975-2000, RunRT/GenInput.py:107-148): lines `NAME=v1;v2;...`
are swept, the FIRST such line varying fastest; a trailing `&` ties a line to the
previous one (they vary together); `NAME=value` lines are constants; names that do not
start with a letter are plot tags, not NAMELIST variables; `#` starts a comment.
The outputs of the runs follow `_DATA_`, concatenated in loop order.
"""
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runrt")


def parse_sbd(name):
    """-> (list of NAMELIST texts in loop order, list of golden output texts)."""
    txt = open(os.path.join(GOLDEN, name + ".sbd")).read()
    head, data = txt.split("_DATA_\n", 1)
    groups, consts = [], []          # groups: list of dicts name -> values (tied lines share a group)
    for line in head.splitlines():
        p = line.split("#")[0].strip()
        if not p or "=" not in p:
            continue
        nm, val = p.split("=", 1)
        nm, val = nm.strip(), val.strip()
        tied = val.endswith("&")           # varies together with the previous swept line
        if tied:
            val = val[:-1].strip()
        if ";" in val:
            vals = [v.strip() for v in val.split(";")]
            if tied and groups:
                groups[-1][nm] = vals
            else:
                groups.append({nm: vals})
        else:
            consts.append((nm, val))
    cycles = [len(next(iter(g.values()))) for g in groups]
    niter = 1
    for c in cycles:
        niter *= c
    inputs = []
    for it in range(niter):
        k, lines = it, []
        for g, c in zip(groups, cycles):
            i = k % c
            k //= c
            lines += [f" {nm}={vals[i]}" for nm, vals in g.items() if nm[0].isalpha()]
        lines += [f" {nm}={val}" for nm, val in consts if nm[0].isalpha()]
        inputs.append("&INPUT\n" + "\n".join(lines) + "\n/")
    rows = data.split("\n")
    if rows and rows[-1] == "":
        rows.pop()
    assert len(rows) % niter == 0, (name, len(rows), niter)
    nl = len(rows) // niter
    outputs = ["\n".join(rows[i * nl:(i + 1) * nl]) + "\n" for i in range(niter)]
    return inputs, outputs
