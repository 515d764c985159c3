"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerance: north_star asks <= 1e-5 relative on fluxes; the two
implementations use different eigen-solvers and eliminations, so agreement is
expected near 1e-10.  The tests assert 1e-7 relative with an absolute floor of
1e-12 x the largest flux of the bin.
"""
import numpy as np
import pytest

import sbdart_b200 as sb
from sbdart_b200 import workloads
from oracle import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-7
KEYS = ("rfldir", "rfldn", "flup", "dfdt", "uavg")


@pytest.fixture(scope="module")
def solver():
    s = sb.Solver(0)
    yield s
    s.close()


def oracle_flux(w, nthreads=8):
    b = w["bins"]
    return oracle.disort_flux_batch(
        w["dtauc"], w["ssalb"], w["pmom"], nstr=w["nstr"], fbeam=b["fbeam"], umu0=b["umu0"],
        albedo=b["albedo"], plank=b["plank"], wvnmlo=b["wvnmlo"], wvnmhi=b["wvnmhi"],
        btemp=b["btemp"], ttemp=b["ttemp"], temis=b["temis"], fisot=b["fisot"],
        temper=w["temper"], col=b["col"], nthreads=nthreads)


def assert_close(got, ref, rtol=RTOL, atol_scale=1e-10):
    """|got-ref| <= rtol*|ref| + atol_scale*scale, scale = largest flux of the bin
    (x 4 pi for dfdt, which is a flux divergence per unit optical depth)."""
    assert (got["status"] == ref["status"]).all()
    ok = ref["status"] == 0
    scale = np.max([np.abs(ref[k][ok]).max(axis=1) for k in ("rfldir", "rfldn", "flup")], axis=0)
    for k in KEYS:
        atol = atol_scale * scale[:, None] * (4 * np.pi if k == "dfdt" else 1.0)
        err = np.abs(got[k][ok] - ref[k][ok]) - atol
        bad = err > rtol * np.abs(ref[k][ok])
        assert not bad.any(), (k, np.argwhere(bad)[:5], got[k][ok][bad][:5], ref[k][ok][bad][:5])


@pytest.mark.parametrize("nstr", [4, 8, 16])
def test_retrieval_bins_match_oracle(solver, nstr):
    w = workloads.retrieval_batch(96, nstr=nstr, nlyr=33, ncols=6, seed=nstr)
    got = solver.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=nstr)
    assert_close(got, oracle_flux(w))


def test_shortwave_with_thermal_matches_oracle(solver):
    w = workloads.mls_shortwave(nstr=16, wlinc=0.05)   # 76 wavelengths, beam + Planck mix
    got = solver.disort_batch(w["dtauc"], w["ssalb"], w["pmom"], w["bins"], nstr=16,
                              temper=w["temper"])
    ref = oracle_flux(w)
    assert (ref["status"] == 0).all()
    assert_close(got, ref)


def test_self_test_fluxes_through_c_abi(solver):
    # DISORT self test (disort.f:6393-6449), flux part, USRTAU level 0.5
    pm = np.zeros((1, 1, 5)); pm[0, 0] = [1.0, 0.8042, 0.646094, 0.481851, 0.359056]
    bins = sb.make_bins(1, fbeam=3.14159265, umu0=0.866, fisot=1.0, albedo=0.7, btemp=300.0,
                        ttemp=100.0, temis=0.8, wvnmlo=0.0, wvnmhi=50000.0, plank=1)
    got = solver.disort_batch([[1.0]], [[0.9]], pm, bins, nstr=4, temper=[[210.0, 200.0]],
                              utau=[[0.5]])
    assert got["status"][0] == 0
    assert abs(got["rfldir"][0, 0] / 1.527286 - 1) < 1e-7
    assert abs(got["rfldn"][0, 0] / 28.372225 - 1) < 1e-7
    assert abs(got["flup"][0, 0] / 152.585284 - 1) < 1e-7


@pytest.mark.parametrize("nstr", [4, 16])
def test_conservative_scattering_layers(solver, nstr):
    """SSALB == 1 exactly (dithered to 1-DITHER, disort.f:486) in Rayleigh and
    cloud layers: the near-null eigenvalue k -> 0 is the hardest case for
    every eigen-solver; fluxes must still agree."""
    rng = np.random.default_rng(5)
    B, L, nmom = 48, 33, nstr + 2
    dtauc = 10.0 ** rng.uniform(-3, 0.3, size=(B, L))
    ssalb = np.ones((B, L))
    ssalb[:, ::5] = 1.0 - 10.0 ** rng.uniform(-9, -2, size=(B, len(range(0, L, 5))))
    pmom = np.zeros((B, L, nmom + 1))
    pmom[:, :, 0] = 1.0
    pmom[:, :, 2] = 0.1                                    # Rayleigh
    cloud = workloads.hg_moments(np.full((B,), 0.85), nmom)
    pmom[:, 20, :] = cloud
    dtauc[:, 20] = rng.uniform(1, 30, size=B)
    bins = sb.make_bins(B, fbeam=1.0, umu0=rng.uniform(0.2, 0.95, B), albedo=rng.uniform(0, 1, B))
    w = dict(dtauc=dtauc, ssalb=ssalb, pmom=pmom, bins=bins, nstr=nstr, temper=None)
    got = solver.disort_batch(dtauc, ssalb, pmom, bins, nstr=nstr)
    ref = oracle_flux(w)
    assert (ref["status"] == 0).all()
    assert_close(got, ref, rtol=1e-6, atol_scale=1e-9)
