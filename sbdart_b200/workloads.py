"""Synthetic bin sets shaped like BASELINE.json's configs.

These only build INPUT arrays (optical depths, single-scattering albedos,
phase-function moments, per-bin scalars); they never compute radiative
transfer.  Shapes follow SURVEY section 8(d):
  * mls_shortwave  -- config C2: 0.25-4.0 um at 0.005 um (751 wavelengths, up
    to three k-distribution terms each), NSTR=16, 33 layers; Rayleigh + band
    absorption + an optional aerosol/cloud layer, beam source below 2 um and
    beam + Planck above (drt.f:463-467).
  * retrieval_batch -- config C5: random columns x bins, NSTR=16.
"""
from __future__ import annotations

import numpy as np

from . import make_bins, quadrature

# mid-latitude-summer-like level temperatures (K) top-down at 34 levels; only
# the Planck source uses them, so a smooth profile of the right range suffices.
_Z33 = np.array([100, 70, 50, 45, 40, 35, 30, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14,
                 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0], dtype=float)


def _mls_like_temperature(z):
    t = np.where(z <= 13, 294.0 - 6.0 * z, 216.0)
    t = np.where((z > 25) & (z <= 50), 216.0 + 2.4 * (z - 25), t)
    t = np.where(z > 50, 276.0 - 2.2 * (z - 50), t)
    return np.maximum(t, 180.0)


def hg_moments(g, nmom):
    k = np.arange(nmom + 1)
    return np.asarray(g)[..., None] ** k


def avoid_quadrature_angles(umu0, nstr, rng, tol=2e-4):
    """Re-draw beam cosines that sit on a Gauss node (disort.f:2645 would ask
    for a different NSTR; SURVEY 8d prescribes rejection for the synthetic set)."""
    mu, _ = quadrature(nstr // 2)
    umu0 = np.array(umu0, dtype=float, copy=True)
    for _ in range(100):
        bad = (np.abs(umu0[:, None] - mu[None, :]) / umu0[:, None] < tol).any(axis=1)
        if not bad.any():
            break
        umu0[bad] = rng.uniform(0.15, 0.98, size=int(bad.sum()))
    return umu0


def mls_shortwave(nstr=16, nlyr=33, wlinf=0.25, wlsup=4.0, wlinc=0.005, sza=30.0,
                  albedo=0.2, seed=20261017, replicate=1, cloud_tau=0.0):
    """Config C2-shaped bins: one per (wavelength, k-term)."""
    rng = np.random.default_rng(seed)
    nwl = int(round((wlsup - wlinf) / wlinc)) + 1
    wl = wlinf + wlinc * np.arange(nwl)
    z = _Z33 if nlyr == 33 else np.linspace(100.0, 0.0, nlyr)
    temper = np.concatenate([[0.0], _mls_like_temperature(z)])
    temper[0] = temper[1]                      # drt.f:330-333
    p = 1013.0 * np.exp(-z / 7.5)
    dp = np.diff(np.concatenate([[0.0], p]))   # layer 1 = cap above the top level
    dp[0] = p[0]
    nmom = nstr + 2
    dtau_l, ssa_l, pm_l, wl_l, k_l = [], [], [], [], []
    for iw, w in enumerate(wl):
        tau_ray = 0.008569 * w ** -4 * (1 + 0.0113 * w ** -2) * dp / 1013.0
        # band absorption: smooth pseudo-spectrum with strong bands + noise
        band = (np.exp(-((w - 0.94) / 0.02) ** 2) + 3 * np.exp(-((w - 1.38) / 0.05) ** 2) +
                6 * np.exp(-((w - 1.87) / 0.07) ** 2) + 20 * np.exp(-((w - 2.7) / 0.15) ** 2) +
                30 * np.exp(-((w - 0.27) / 0.03) ** 2) + 0.02)
        nk = 3 if band > 0.05 else 1
        gk = np.array([1.0]) if nk == 1 else np.array([0.05, 1.0, 12.0])
        aer = 0.1 * (w / 0.55) ** -1.3 * np.exp(-z / 1.5)
        aer = aer * np.concatenate([[0.0], -np.diff(z)]) / 1.5
        g_aer = 0.7
        tau_cld = np.zeros(nlyr)
        if cloud_tau > 0:
            tau_cld[nlyr - 3] = cloud_tau
        for kd in range(nk):
            tau_abs = band * gk[kd] * (dp / 1013.0) * rng.uniform(0.9, 1.1)
            sca = tau_ray + 0.95 * aer + tau_cld * 0.999999
            tot = tau_ray + aer + tau_cld + tau_abs
            pm = np.zeros((nlyr, nmom + 1))
            pm[:, 0] = 1.0
            num = (0.95 * aer)[:, None] * hg_moments(g_aer, nmom)[None, :] + \
                tau_cld[:, None] * 0.999999 * hg_moments(0.85, nmom)[None, :]
            num[:, 2] += 0.1 * tau_ray                      # drt.f:1387
            num[:, 0] += tau_ray
            pm = num / np.maximum(sca, 1e-300)[:, None]
            pm[:, 0] = 1.0
            dtau_l.append(tot)
            ssa_l.append(np.minimum(sca / tot, 1.0))
            pm_l.append(pm)
            wl_l.append(w)
            k_l.append(kd)
    dtauc = np.array(dtau_l)
    ssalb = np.array(ssa_l)
    pmom = np.array(pm_l)
    wl_b = np.array(wl_l)
    B = dtauc.shape[0]
    plank = (wl_b > 2.0).astype(np.int32)                  # drt.f:463-467
    dwn = 1e4 / np.maximum(wl_b - wlinc / 2, 1e-3) - 1e4 / (wl_b + wlinc / 2)
    wvhi = 1e4 / np.maximum(wl_b - wlinc / 2, 1e-3)
    wvlo = wvhi - dwn
    umu0 = float(np.cos(np.deg2rad(sza)))
    bins = make_bins(B, fbeam=1.0 / np.maximum(wl_b, 0.25) ** 2, umu0=umu0, albedo=albedo,
                     btemp=temper[-1], ttemp=temper[0], temis=0.0, wvnmlo=wvlo, wvnmhi=wvhi,
                     plank=plank, col=0)
    if replicate > 1:
        dtauc = np.tile(dtauc, (replicate, 1))
        ssalb = np.tile(ssalb, (replicate, 1))
        pmom = np.tile(pmom, (replicate, 1, 1))
        bins = np.tile(bins, replicate)
    return dict(dtauc=dtauc, ssalb=ssalb, pmom=pmom, bins=bins, temper=temper[None, :],
                nstr=nstr, wl=wl_b, name=f"mls_shortwave_{wlinf}-{wlsup}um@{wlinc}_nstr{nstr}_L{nlyr}")


def retrieval_batch(nbins, nstr=16, nlyr=33, ncols=None, seed=20261017):
    """Config C5 distributions: dtau = 10**U(-3,0.5), ssalb =
    1-10**U(-6,0), HG g ~ U(0,0.9); per column umu0 ~ U(0.15,0.98) (rejected
    near Gauss nodes), albedo ~ U(0,1), fbeam = 1, no thermal source."""
    rng = np.random.default_rng(seed)
    nmom = nstr + 2
    ncols = ncols or max(1, nbins // 1000)
    dtauc = 10.0 ** rng.uniform(-3, 0.5, size=(nbins, nlyr))
    ssalb = 1.0 - 10.0 ** rng.uniform(-6, 0, size=(nbins, nlyr))
    g = rng.uniform(0, 0.9, size=(nbins, nlyr))
    pmom = hg_moments(g, nmom)
    col = (np.arange(nbins) * ncols) // nbins
    umu0_c = avoid_quadrature_angles(rng.uniform(0.15, 0.98, size=ncols), nstr, rng)
    alb_c = rng.uniform(0, 1, size=ncols)
    bins = make_bins(nbins, fbeam=1.0, umu0=umu0_c[col], albedo=alb_c[col], plank=0, col=0)
    return dict(dtauc=dtauc, ssalb=ssalb, pmom=pmom, bins=bins, temper=None, nstr=nstr,
                name=f"retrieval_{nbins}bins_nstr{nstr}_L{nlyr}")
