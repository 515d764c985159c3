"""Timing helper shared by bench.py and tests/bench_configs.py: one batched DISORT call on a
bin set, (a) with every buffer resident in HBM (CUDA events on the solver's stream) and
(b) end to end through the host-buffer C-ABI call with pinned host buffers.

Uses torch only for device / pinned memory and CUDA events (plumbing)."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import SbdDims, SbdError, lib


class BatchTimer:
    """w: dict(dtauc [B][L], ssalb, pmom [B][L][nmom+1], bins, temper [ncol][L+1] | None, nstr);
    umu / phi: user angles of a radiance run; uu_levels: output levels whose intensities are
    wanted (packed layout: uu is [B][nphi][nsel][numu])."""

    def __init__(self, solver, w, umu=None, phi=None, uu_levels=None, corint=False, device=0):
        import torch
        self.torch, self.solver, self.dev = torch, solver, torch.device("cuda", device)
        self.B, self.L = w["dtauc"].shape
        self.NT = self.L + 1
        self.nmom = w["pmom"].shape[2] - 1
        self.nstr = int(w["nstr"])
        self.umu = None if umu is None else np.ascontiguousarray(umu, dtype=np.float64)
        self.phi = None if phi is None else np.ascontiguousarray(phi, dtype=np.float64)
        self.levels = None if uu_levels is None else np.ascontiguousarray(sorted(set(uu_levels)), dtype=np.int32)
        self.corint = bool(corint)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
        B, NT = self.B, self.NT
        self.h = dict(dtauc=pin(w["dtauc"]), ssalb=pin(w["ssalb"]), pmom=pin(w["pmom"]),
                      bins=torch.from_numpy(w["bins"].view(np.uint8).reshape(B, -1).copy()).pin_memory())
        temper = w.get("temper")
        self.ncol = 0
        if temper is not None:
            temper = np.atleast_2d(temper)
            self.ncol = temper.shape[0]
            self.h["temper"] = pin(temper)
        for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg"):
            self.h[k] = torch.empty((B, NT), dtype=torch.float64).pin_memory()
        self.h["status"] = torch.empty(B, dtype=torch.int32).pin_memory()
        self.numu = 0 if self.umu is None else len(self.umu)
        self.nphi = 0 if self.phi is None else len(self.phi)
        self.nlev = NT if self.levels is None else len(self.levels)
        if self.numu:
            self.h["uu"] = torch.empty((B, self.nphi, self.nlev, self.numu), dtype=torch.float64).pin_memory()
        self.d = {k: v.to(self.dev) for k, v in self.h.items()}
        torch.cuda.synchronize()
        d = SbdDims()
        d.nbins, d.nlyr, d.nstr, d.nmom, d.ncol = B, self.L, self.nstr, self.nmom, self.ncol
        d.numu, d.nphi = self.numu, self.nphi
        self.dims = d
        self.h2d_bytes = sum(self.h[k].numel() * self.h[k].element_size()
                             for k in ("dtauc", "ssalb", "pmom", "bins", "temper") if k in self.h)
        self.d2h_bytes = sum(self.h[k].numel() * self.h[k].element_size()
                             for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg", "status", "uu") if k in self.h)

    # -- handle state for radiance runs -----------------------------------------------
    def _enter(self):
        L = lib()
        if self.levels is not None:
            L.sbd_set_radiance_levels(self.solver._h, self.levels.ctypes.data, len(self.levels))
            L.sbd_set_radiance_layout(self.solver._h, 1)
        if self.corint:
            L.sbd_set_corint(self.solver._h, 1)

    def _exit(self):
        L = lib()
        if self.levels is not None:
            L.sbd_set_radiance_levels(self.solver._h, None, 0)
            L.sbd_set_radiance_layout(self.solver._h, 0)
        if self.corint:
            L.sbd_set_corint(self.solver._h, 0)

    def _call(self, bufs, device):
        p = lambda k: bufs[k].data_ptr() if k in bufs else None  # noqa: E731
        um = None if self.umu is None else self.umu.ctypes.data    # user angles: host arrays in both calls
        ph = None if self.phi is None else self.phi.ctypes.data
        args = [self.solver._h, C.byref(self.dims), p("dtauc"), p("ssalb"), p("pmom"), p("bins"), p("temper"),
                None, um, ph, p("rfldir"), p("rfldn"), p("flup"), p("dfdt"), p("uavg"), p("uu"), p("status")]
        rc = lib().sbd_disort_batch_device(*args, None) if device else lib().sbd_disort_batch(*args)
        if rc:
            raise SbdError(rc, "sbd_disort_batch" + ("_device" if device else ""))

    def device_ms(self, steps=5, warmup=3, after_step=None):
        """Milliseconds per call with device-resident buffers (events on the solver's stream)."""
        torch = self.torch
        ext = torch.cuda.ExternalStream(self.solver.stream, device=self.dev)
        self._enter()
        try:
            for _ in range(warmup):
                self._call(self.d, True)
                if after_step:
                    after_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(steps):
                self._call(self.d, True)
                if after_step:
                    after_step()
            e1.record(ext)
            torch.cuda.synchronize()
        finally:
            self._exit()
        return e0.elapsed_time(e1) / steps

    def e2e_ms(self, steps=5, warmup=2, after_step=None):
        """Milliseconds per host-buffer call (pinned inputs -> H2D -> kernels -> D2H)."""
        self._enter()
        try:
            for _ in range(warmup):
                self._call(self.h, False)
                if after_step:
                    after_step()
            self.torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                self._call(self.h, False)
                if after_step:
                    after_step()
            self.torch.cuda.synchronize()
        finally:
            self._exit()
        return (time.perf_counter() - t0) * 1e3 / steps

    def results(self, device=False):
        bufs = self.d if device else self.h
        out = {k: bufs[k].cpu().numpy() for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg", "status", "uu")
               if k in bufs}
        if self.levels is not None:
            out["uu_levels"] = [int(v) for v in self.levels]
        return out
