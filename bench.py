#!/usr/bin/env python
"""bench.py -- DISORT spectral-points/sec on B200 (BASELINE.json metric).

Workload (config C2): shortwave 0.25-4.0 um at 0.005 um, NSTR=16,
33 layers (mid-latitude summer); the bin set (one bin per wavelength x k-term,
2037 bins, produced by the SBDART front end sbdart_b200/frontend) is replicated
R times so that one step is one batched launch over R complete spectra.

A step = one pass of the hot path (one batched DISORT solve of every bin).
  value : bins/s with inputs resident in HBM (device pointers, CUDA events
          on the solver's stream), max over ranks.
  e2e   : the same spectra through the whole-spectrum C-ABI call
          sbd_spectrum_run_columns with HOST buffers: the setup arrays of every column go
          in, the optical properties are produced on the device (K2), the solve follows, and
          the top / bottom fluxes the record reads come back (copies inside the timed region).
          e2e_host_buffers keeps the round-1 route (sbd_disort_batch: every bin's optical
          properties over PCIe, all levels back).
  configs : C3 / C4 / C5 of BASELINE.json as ONE job split over the ranks (strong scaling).
  --impl reference : the CPU restatement (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DISORT spectral-points/sec (NSTR=16, 33-layer MLS)"
UNIT = "spectral-points/s"


def algorithmic_flops(nstr, nlyr, ncut, plank, beam, nt):
    """SURVEY 8(d) agreed FLOP count per bin (flux path, M=1)."""
    N, n = nstr, nstr // 2
    b = 3 * n - 1
    per_layer = 3 * n * N * N + 29 * n ** 3
    per_layer = per_layer + beam * (2.0 / 3.0 * N ** 3 + 10 * N * N)
    per_layer = per_layer + plank * (2.0 / 3.0 * N ** 3 + 9 * N * N)
    return ncut * (per_layer + 4 * b * b * N + 26 * b * N) + 5 * N * N * nt


def workload_ncut(w):
    """NCUT per bin (disort.f:2557-2605): first layer where the cumulative
    absorption depth reaches 10, applied only without a thermal source."""
    ss = np.where(w["ssalb"] == 1.0, 1.0 - 100 * 2.0 ** -52, w["ssalb"])
    ab = np.cumsum((1.0 - ss) * np.maximum(w["dtauc"], 0.0), axis=1)
    L = ab.shape[1]
    before = np.concatenate([np.zeros((ab.shape[0], 1)), ab[:, :-1]], axis=1)
    ncut = (before < 10.0).sum(axis=1)
    cut = (ab[:, -1] >= 10.0) & (w["bins"]["plank"] == 0) & (L > 1)
    return np.where(cut, ncut, L)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


C2_NAMELIST = "&INPUT idatm=2, wlinf=.25, wlsup=4.0, wlinc=.005, nstr=16, iout=1 /"


def build_workload(replicate):
    """Config C2: the bins SBDART itself would hand to DISORT for the
    mid-latitude-summer atmosphere, 0.25-4.0 um at 0.005 um, NSTR=16 -- produced by
    the front end (LOWTRAN7 band model + 3-term k-distribution + Rayleigh)."""
    from sbdart_b200.frontend import Sbdart
    run = Sbdart(C2_NAMELIST)
    w = run.batch(run.bins())
    w["name"] = "sbdart_C2_mls_0.25-4.0um@0.005_nstr16_L33"
    base = w["dtauc"].shape[0]
    if replicate > 1:
        for k in ("dtauc", "ssalb", "pmom"):
            w[k] = np.tile(w[k], (replicate,) + (1,) * (w[k].ndim - 1))
        w["bins"] = np.tile(w["bins"], replicate)
    w["base_bins"] = base
    return w


def run_reference(args, rank, world):
    """CPU arm: the oracle port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import oracle
    w = build_workload(1)
    cores = os.cpu_count() or 1
    sample = min(args.ref_bins, w["dtauc"].shape[0] * 32)
    rep = -(-sample // w["dtauc"].shape[0])
    idx = np.arange(sample) % w["dtauc"].shape[0]
    b = w["bins"][idx]
    kw = dict(nstr=16, fbeam=b["fbeam"], umu0=b["umu0"], albedo=b["albedo"], plank=b["plank"],
              wvnmlo=b["wvnmlo"], wvnmhi=b["wvnmhi"], btemp=b["btemp"], ttemp=b["ttemp"],
              temis=b["temis"], fisot=b["fisot"], temper=w["temper"], col=b["col"], nthreads=cores)
    dt, ss, pm = w["dtauc"][idx], w["ssalb"][idx], w["pmom"][idx]
    for _ in range(args.warmup):
        oracle.disort_flux_batch(dt, ss, pm, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = oracle.disort_flux_batch(dt, ss, pm, **kw)
    dtm = (time.perf_counter() - t0) / args.steps
    val = sample / dtm
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dtm * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (no dataset: optical properties computed by the SBDART front end from the MLS model atmosphere)",
        "config": {"workload": w["name"], "bins_per_step": int(sample),
                   "note": "CPU restatement of disort.f (oracle/, OpenMP over bins); "
                           "no Fortran compiler in the image so the reference itself cannot be built"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} bins of the C2 set per step ({rep} spectra)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bad_bins": int((r["status"] != 0).sum()),
    }
    print(json.dumps(line), flush=True)


def config_workloads(world, rank):
    """The other BASELINE.json configurations as this rank's contiguous share of ONE job
    (strong scaling: C3 and C4 are single SBDART runs, C5 is the 10^6-bin retrieval batch)."""
    from sbdart_b200 import workloads
    from sbdart_b200.frontend import Sbdart
    from sbdart_b200.sharding import bin_partition
    uz = ",".join(str(x) for x in np.linspace(5.0, 85.0, 10))
    out = []
    for name, nl in (
            ("C3_thermal_radiances_nstr8_10zen", f"&INPUT idatm=2, wlinf=4, wlsup=80, wlinc=20, nstr=8, iout=20, uzen={uz}, sza=95 /"),
            ("C4_nstr32_65layers_cloud_aerosol", "&INPUT idatm=2, nstr=32, ngrid=65, tcloud=10, zcloud=1, iaer=1, vis=23,"
                                                 " wlinf=.25, wlsup=100, wlinc=20, iout=10 /")):
        run = Sbdart(nl)
        b = run.batch(run.bins())
        lo, hi = bin_partition(len(b["bins"]), world, b["group"])[rank]
        w = {k: b[k][lo:hi] for k in ("dtauc", "ssalb", "pmom", "bins")}
        w["temper"], w["nstr"] = b["temper"], b["nstr"]
        out.append(dict(name=name, total=len(b["bins"]), w=w, umu=b.get("umu"), phi=b.get("phi"),
                        uu_levels=b.get("uu_levels")))
    total = 1000000
    lo, hi = bin_partition(total, world)[rank]
    # every rank draws its own share (same distributions; rank-dependent seed)
    w = workloads.retrieval_batch(hi - lo, nstr=16, nlyr=33, ncols=max(1, (hi - lo) // 1000), seed=20261017 + rank)
    out.append(dict(name="C5_retrieval_1e6bins_nstr16", total=total, w=w, umu=None, phi=None, uu_levels=None))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicate", type=int, default=128, help="spectra per step per GPU (SURVEY 8d asks for a large replication of the 2037-bin spectrum)")
    ap.add_argument("--ref-bins", type=int, default=65536, help="bins per CPU reference step (~2.5 s on 16 threads)")
    ap.add_argument("--cpu-sample", type=int, default=262144, help="bins for cpu_baseline (about 10 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 strong-scaling lines")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import sbdart_b200 as sb
    from sbdart_b200.frontend import Sbdart
    from sbdart_b200.frontend.device import ColumnRunner
    from sbdart_b200.timing import BatchTimer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = build_workload(args.replicate)
    w["nstr"] = 16
    B, L = w["dtauc"].shape
    NT = L + 1
    nmom = w["pmom"].shape[2] - 1
    solver = sb.Solver(local_rank)
    ext = torch.cuda.ExternalStream(solver.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs, CUDA events on the solver stream ------------
    bt = BatchTimer(solver, w, device=local_rank)
    d_out = bt.d
    # compact top/bottom spectrum gathered at the end of a step
    d_spec = torch.empty((B, 6), dtype=torch.float64, device=dev)
    g_spec = torch.empty((world * B, 6), dtype=torch.float64, device=dev) if world > 1 else None

    def gather_device():
        if world > 1:
            with torch.cuda.stream(ext):
                torch.stack([d_out["rfldn"][:, 0], d_out["flup"][:, 0], d_out["rfldir"][:, 0],
                             d_out["rfldn"][:, -1], d_out["flup"][:, -1], d_out["rfldir"][:, -1]],
                            dim=1, out=d_spec)
                dist.all_gather_into_tensor(g_spec, d_spec)

    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = solver.kernel_launches
    ms_dev = bt.device_ms(steps=args.steps, warmup=args.warmup, after_step=gather_device)
    barrier()
    clocks = sampler.stop()
    launches = solver.kernel_launches - l0 - args.warmup * ((solver.kernel_launches - l0) // (args.steps + args.warmup))
    bad = int((d_out["status"] != 0).sum().item())

    # ---- e2e, production route: R atmospheric columns through the whole-spectrum C-ABI call.
    # Host buffers in (setup arrays of every column), host buffers out (the top / bottom fluxes
    # the iout=1 record reads); optical properties are produced on the device (K2) and never
    # cross PCIe; multi-GPU: the ranks' spectra are all-gathered on the device.
    run = Sbdart(C2_NAMELIST)
    pin = lambda shape, dtype: torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory().numpy()  # noqa: E731
    cr = ColumnRunner(run, solver, args.replicate, levels=[run.ntop - 1, run.nbot - 1], alloc=pin)
    g_e2e = [None]

    class _Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    import ctypes as C
    sb.lib().sbd_spectrum_device_fluxes.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]

    def step_e2e():
        cr.step()
        if world > 1:
            ptr, n = C.c_void_p(), C.c_int64()
            sb.lib().sbd_spectrum_device_fluxes(solver._h, C.byref(ptr), C.byref(n))
            mine = torch.as_tensor(_Dev(ptr.value, n.value), device=dev)
            if g_e2e[0] is None or g_e2e[0].numel() != world * n.value:
                g_e2e[0] = torch.empty(world * n.value, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(g_e2e[0], mine)
            torch.cuda.synchronize()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    nb_e2e = cr.nbins.value
    h2d_e2e, d2h_e2e = cr.transfer_bytes()
    # the production route must reproduce the device arm's fluxes (first spectrum, top / bottom)
    nb1 = int(w["base_bins"])
    ref_top = d_out["flup"][:nb1, 0].cpu().numpy()
    ref_bot = (d_out["rfldn"][:nb1, -1] + d_out["rfldir"][:nb1, -1]).cpu().numpy()
    got_top, got_bot = cr.flup[:nb1, 0], cr.rfldn[:nb1, 1] + cr.rfldir[:nb1, 1]
    scale = max(np.abs(ref_bot).max(), 1e-300)
    e2e_check = float(max(np.abs(got_top - ref_top).max(), np.abs(got_bot - ref_bot).max()) / scale)
    e2e_bad = int((cr.status[:nb_e2e] != 0).sum())

    # ---- e2e through the host-buffer batch call (every optical property over PCIe, all
    # levels back): the number the previous round reported, kept beside the production one
    ms_hb = bt.e2e_ms(steps=max(2, args.steps // 2), warmup=2)
    barrier()

    t_dev = torch.tensor([ms_dev, ms_e2e, ms_hb], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_hb = (float(x) for x in t_dev)

    # ---- the other configurations: this rank's share of one job (strong scaling) --------
    cfg_lines = {}
    if not args.no_configs:
        for c in config_workloads(world, rank):
            wc = c["w"]
            nb = len(wc["bins"])
            t = BatchTimer(solver, wc, umu=c["umu"], phi=c["phi"], uu_levels=c["uu_levels"], device=local_rank)
            gat = None
            if world > 1:        # one all-gather of the top / bottom fluxes per step (padded to the largest share)
                nmax = -(-c["total"] // world) + 4
                mine = torch.zeros((nmax, 6), dtype=torch.float64, device=dev)
                allr = torch.empty((world * nmax, 6), dtype=torch.float64, device=dev)

                def gat(t=t, mine=mine, allr=allr, nb=nb):
                    with torch.cuda.stream(ext):
                        torch.stack([t.d["rfldn"][:, 0], t.d["flup"][:, 0], t.d["rfldir"][:, 0], t.d["rfldn"][:, -1],
                                     t.d["flup"][:, -1], t.d["rfldir"][:, -1]], dim=1, out=mine[:nb])
                        dist.all_gather_into_tensor(allr, mine)
            barrier()
            msd = t.device_ms(steps=3, warmup=2, after_step=gat)
            barrier()
            mse = t.e2e_ms(steps=3, warmup=1)
            barrier()
            tt = torch.tensor([msd, mse], dtype=torch.float64, device=dev)
            nbad = torch.tensor([int((t.results(True)["status"] != 0).sum())], device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dist.all_reduce(nbad)
            cfg_lines[c["name"]] = {
                "bins_total": c["total"], "bins_this_rank": nb, "scaling": "strong",
                "value": c["total"] / (float(tt[0]) * 1e-3), "ms_per_step": float(tt[0]),
                "e2e_value": c["total"] / (float(tt[1]) * 1e-3), "e2e_ms_per_step": float(tt[1]),
                "unit": UNIT, "bad_bins": int(nbad[0]),
                "note": "device-timed incl. the all-gather of the top/bottom fluxes; e2e = host-buffer C-ABI call "
                        "of this rank's share (pinned buffers), max over ranks"}
            del t

    if rank == 0:
        h2d = (w["dtauc"].nbytes + w["ssalb"].nbytes + w["pmom"].nbytes + w["bins"].nbytes +
               w["temper"].nbytes)
        d2h = 5 * B * NT * 8 + 4 * B
        ncut = workload_ncut(w)
        beam = (w["bins"]["fbeam"] > 0).astype(float)
        fl = algorithmic_flops(16, L, ncut, w["bins"]["plank"].astype(float), beam, NT).sum()
        fp64_peak = solver.measure_fp64_peak(5)
        achieved = fl / (ms_dev * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = B * (8 * (2 * L + (nmom + 1) * L) + w["bins"].dtype.itemsize + 8 * 5 * NT + 4)
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
            if prof.get("bins_per_launch"):
                traffic = prof["dram_bytes_per_launch"] * (B / prof["bins_per_launch"])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * B / (ms_dev * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (no dataset: optical properties computed by the SBDART front end from the MLS model atmosphere)",
            "config": {"workload": w["name"], "bins_per_gpu_per_step": int(B),
                       "spectra_per_step": args.replicate, "bins_per_spectrum": int(w["base_bins"]),
                       "levels_out": NT,
                       "l2": f"inputs larger than L2: {h2d / 2**20:.0f} MiB read per step",
                       "parallelism": f"bins sharded over {world} GPU(s), one all-gather of the "
                                      "top/bottom spectrum per step" if world > 1 else "1 GPU"},
            "clocks": clocks,
            "e2e": {"value": world * nb_e2e / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d_e2e), "d2h_bytes_per_step": int(d2h_e2e),
                    "ms_per_step": ms_e2e, "bins_per_gpu_per_step": int(nb_e2e),
                    "route": "sbd_spectrum_run_columns: host setup arrays of every column in, top/bottom fluxes "
                             "out; optical properties produced on the device (K2), solve, level pack"
                             + (", device all-gather of the ranks' spectra" if world > 1 else ""),
                    "max_flux_diff_vs_device_arm_over_scale": e2e_check, "bad_bins": e2e_bad},
            "e2e_host_buffers": {"value": world * B / (ms_hb * 1e-3), "unit": UNIT, "ms_per_step": ms_hb,
                                 "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                                 "route": "sbd_disort_batch: every bin's dtauc/ssalb/pmom over PCIe, all levels back"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                         "traffic": traffic,
                         "peak_source": "sbd_measure_fp64_peak (live DFMA microbenchmark; "
                                        "MEASURED_PEAKS.json has no FP64 entry)",
                         "flops_per_launch": float(fl),
                         "hbm": {"achieved": alg_bytes / (ms_dev * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s",
                                 "frac": alg_bytes / (ms_dev * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "bad_bins": bad,
        }
        if cfg_lines:
            line["configs"] = cfg_lines
        if not args.no_cpu_baseline and world >= 1:
            from oracle import oracle
            cores = os.cpu_count() or 1
            ns = min(args.cpu_sample, B)
            idx = np.arange(ns)
            b = w["bins"][idx]
            t0 = time.perf_counter()
            oracle.disort_flux_batch(
                w["dtauc"][idx], w["ssalb"][idx], w["pmom"][idx], nstr=16, fbeam=b["fbeam"],
                umu0=b["umu0"], albedo=b["albedo"], plank=b["plank"], wvnmlo=b["wvnmlo"],
                wvnmhi=b["wvnmhi"], btemp=b["btemp"], ttemp=b["ttemp"], temis=b["temis"],
                fisot=b["fisot"], temper=w["temper"], col=b["col"], nthreads=cores)
            dtc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dtc, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {ns} bins of the same step, one pass"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
