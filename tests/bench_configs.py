"""Throughput of the other BASELINE.json configurations (not the headline bench line).

bench.py measures config C2.  This script times the same C-ABI host call
(sbd_disort_batch: pinned-free numpy host buffers, H2D + kernel + D2H inside) on
bin sets shaped like C1, C3, C4 and C5 and checks a sample of every
set against the CPU oracle.  One JSON line per configuration.

    python tests/bench_configs.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
import sbdart_b200 as sb                      # noqa: E402
from sbdart_b200 import workloads            # noqa: E402
from oracle import oracle                    # noqa: E402


def tile(w, rep):
    if rep <= 1:
        return w
    w = dict(w)
    for k in ("dtauc", "ssalb", "pmom"):
        w[k] = np.tile(w[k], (rep,) + (1,) * (w[k].ndim - 1))
    w["bins"] = np.tile(w["bins"], rep)
    return w


def check(w, got, nsample, kw):
    """Largest mismatch against the oracle on the first nsample bins, relative to the bin's
    largest flux (and, for radiance runs, largest intensity)."""
    idx = np.arange(min(nsample, len(w["bins"])))
    sub = {k: (w[k][idx] if k in ("dtauc", "ssalb", "pmom", "bins") else w[k]) for k in w}
    sub["nstr"] = w["nstr"]
    if kw.get("umu") is not None:
        sub.update(umu=kw["umu"], phi=kw["phi"])
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from solvers import solve_oracle
    ref = solve_oracle(sub, nthreads=os.cpu_count() or 1)
    ok = ref["status"] == 0
    worst = 0.0
    for k in ("rfldir", "rfldn", "flup"):
        scale = np.abs(ref["flup"][ok]).max(axis=1, keepdims=True) + np.abs(ref["rfldir"][ok]).max(axis=1, keepdims=True)
        err = np.abs(got[k][idx][ok] - ref[k][ok]) / np.maximum(scale, 1e-300)
        worst = max(worst, float(err.max()))
    worst_uu = None
    if "uu" in ref and "uu" in got:
        lv = got.get("uu_levels") or list(range(ref["uu"].shape[2]))
        r = ref["uu"][ok][:, :, lv, :]
        scale = np.abs(r).reshape(len(r), -1).max(axis=1).reshape(-1, 1, 1, 1)
        worst_uu = float((np.abs(got["uu"][idx][ok] - r) / np.maximum(scale, 1e-300)).max())
    return worst, worst_uu, int((got["status"][idx] != ref["status"]).sum())


def run(name, w, solver, reps=3, nsample=64, **kw):
    """Device-resident and end-to-end (pinned host buffers, packed intensities) rates of one
    batched call, and the mismatch of its results against the oracle on a sample."""
    from sbdart_b200.timing import BatchTimer
    t = BatchTimer(solver, w, umu=kw.get("umu"), phi=kw.get("phi"), uu_levels=kw.get("uu_levels"))
    ms_dev = t.device_ms(steps=reps, warmup=2)
    ms_e2e = t.e2e_ms(steps=reps, warmup=1)
    got = t.results(device=False)
    worst, worst_uu, stat_mismatch = check(w, got, nsample, kw)
    B = len(w["bins"])
    print(json.dumps({"config": name, "bins": B, "nstr": w["nstr"], "nlyr": int(w["dtauc"].shape[1]),
                      "device_bins_per_s": B / (ms_dev * 1e-3), "device_ms": ms_dev,
                      "e2e_bins_per_s": B / (ms_e2e * 1e-3), "e2e_ms": ms_e2e,
                      "h2d_bytes": t.h2d_bytes, "d2h_bytes": t.d2h_bytes,
                      "bad_bins": int((got["status"] != 0).sum()),
                      "max_flux_err_over_bin_scale_vs_oracle": worst,
                      "max_radiance_err_over_bin_scale_vs_oracle": worst_uu, "status_mismatch": stat_mismatch,
                      "oracle_sample": min(nsample, B), **{k: (len(v) if hasattr(v, "__len__") else v) for k, v in kw.items()}}),
          flush=True)


def namelist_workload(nl):
    """Bins of a real SBDART run (front end on the host): the BASELINE.json namelists."""
    from sbdart_b200.frontend import Sbdart
    t0 = time.perf_counter()
    r = Sbdart(nl)
    b = r.batch(r.bins())
    b["frontend_s"] = time.perf_counter() - t0
    b["run"] = r
    return b


def run_namelists(s, quick):
    """C3 and C4 exactly as BASELINE.md section 2 writes them, plus the whole-run
    time of the device front end (K2 + solve + records) for the same namelist."""
    from sbdart_b200.frontend import Sbdart
    uz = ",".join(str(x) for x in np.linspace(5.0, 85.0, 10))
    cases = [
        ("C3 namelist thermal night (M=1), iout=20, 10 zenith angles",
         f"&INPUT idatm=2, wlinf=4, wlsup=80, wlinc=20, nstr=8, iout=20, uzen={uz}, sza=95 /", 64),
        ("C3 namelist sunlit (M=8 azimuth modes), iout=20, 10 zenith angles",
         f"&INPUT idatm=2, wlinf=4, wlsup=80, wlinc=20, nstr=8, iout=20, uzen={uz}, sza=30 /", 64),
        ("C4 namelist nstr32 ngrid65 stratus + rural aerosol, 0.25-100 um at 20 cm-1",
         "&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23, wlinf=.25, wlsup=100, wlinc=20,"
         " iout=10 /", 8),
    ]
    for name, nl, rep in cases:
        w = namelist_workload(nl)
        kw = {}
        if "umu" in w:
            kw = dict(umu=w["umu"], phi=w["phi"], uu_levels=w.get("uu_levels"))
        wt = tile({k: w[k] for k in ("dtauc", "ssalb", "pmom", "bins")}, max(1, rep // (4 if quick else 1)))
        wt["nstr"], wt["temper"] = w["nstr"], w["temper"]
        run(name + f" (x{len(wt['bins']) // len(w['bins'])})", wt, s, nsample=32, **kw)
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            txt = Sbdart(nl).run_device(s)
            t.append(time.perf_counter() - t0)
        print(json.dumps({"config": name, "whole_run_device_front_end_s": min(t), "bins": len(w["bins"]),
                          "host_front_end_s": w["frontend_s"], "record": txt.splitlines()[0 if "iout=10" in nl else 0][:120]}),
              flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--namelists", action="store_true", help="only the real-namelist C3 / C4 runs")
    a = ap.parse_args()
    s = sb.Solver(0)
    if a.namelists:
        run_namelists(s, a.quick)
        s.close()
        return
    q = 8 if a.quick else 1
    # C1: NSTR=4, 33 layers, 151 wavelengths (x k-terms), replicated
    w = workloads.mls_shortwave(nstr=4, wlinf=0.25, wlsup=1.0, wlinc=0.005)
    run("C1 nstr4 L33 shortwave", tile(w, 256 // q), s)
    # C3: NSTR=8 thermal, radiances at 10 zenith angles (generic kernel), and its flux-only twin
    w = workloads.mls_shortwave(nstr=8, wlinf=4.0, wlsup=20.0, wlinc=0.05)
    umu = np.cos(np.deg2rad(np.linspace(5.0, 85.0, 10)))[::-1].copy()
    run("C3 nstr8 L33 thermal flux", tile(w, 64 // q), s)
    run("C3 nstr8 L33 thermal radiance 10 zenith x 1 azimuth, all 34 levels", tile(w, 64 // q), s,
        umu=np.sort(umu), phi=np.array([0.0]))
    run("C3 nstr8 L33 thermal radiance 10 zenith x 1 azimuth, top level only (what iout=20 prints)",
        tile(w, 64 // q), s, umu=np.sort(umu), phi=np.array([0.0]), uu_levels=[0])
    # C4: NSTR=32, 65 layers (generic kernel)
    w = workloads.mls_shortwave(nstr=32, nlyr=65, wlinf=0.25, wlsup=4.0, wlinc=0.02, cloud_tau=10.0)
    run("C4 nstr32 L65 cloud", tile(w, 16 // min(q, 4)), s, nsample=32)
    # C5: retrieval batch, NSTR=16, one GPU's share of 10^6 bins
    w = workloads.retrieval_batch(125000 // q, nstr=16, nlyr=33, ncols=125 // q or 1)
    run("C5 nstr16 L33 retrieval (1/8 of 1e6 bins)", w, s)
    s.close()


if __name__ == "__main__":
    main()
