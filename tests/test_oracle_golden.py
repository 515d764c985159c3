"""Pin the CPU oracle against the reference's own known-answer test.

DISORT self test: inputs disort.f:6393-6430, expected values
disort.f:6446-6449, reference tolerance 1e-4 (disort.f:6341).  The four
printed constants carry 8-9 significant digits; the oracle must reproduce
them to 1e-7.
"""
import numpy as np

from oracle import oracle

SLFTST_PMOM = np.array([[1.0, 0.8042, 0.646094, 0.481851, 0.359056]])
SLFTST_KW = dict(nstr=4, temper=[210.0, 200.0], utau=[0.5], umu=[0.5], phi=[90.0],
                 fbeam=3.14159265, umu0=0.866, phi0=0.0, fisot=1.0, albedo=0.7,
                 btemp=300.0, ttemp=100.0, temis=0.8, wvnmlo=0.0, wvnmhi=50000.0,
                 plank=True, onlyfl=False, corint=True, accur=1e-4)
SLFTST_EXPECT = dict(uu=47.865571, rfldir=1.527286, rfldn=28.372225, flup=152.585284)


def test_disort_self_test_constants():
    r = oracle.disort([1.0], [0.9], SLFTST_PMOM, **SLFTST_KW)
    assert r["status"] == 0
    assert abs(r["uu"].ravel()[0] / SLFTST_EXPECT["uu"] - 1) < 1e-7
    for k in ("rfldir", "rfldn", "flup"):
        assert abs(r[k][0] / SLFTST_EXPECT[k] - 1) < 1e-7, k


def test_qgausn_matches_numpy_leggauss():
    import ctypes as C
    for m in (1, 2, 4, 8, 10, 16, 20):
        mu = np.zeros(m)
        wt = np.zeros(m)
        dp = C.POINTER(C.c_double)
        oracle.lib().sbdo_qgausn(m, mu.ctypes.data_as(dp), wt.ctypes.data_as(dp))
        x, w = np.polynomial.legendre.leggauss(m)
        np.testing.assert_allclose(mu, 0.5 * x + 0.5, rtol=0, atol=5e-15)
        np.testing.assert_allclose(wt, 0.5 * w, rtol=1e-13, atol=1e-15)


def test_plkavg_total_is_stefan_boltzmann_over_pi():
    # integral over all wavenumbers = sigma T^4 / pi  (disort.f:5410 header)
    t = 300.0
    v = oracle.lib().sbdo_plkavg(0.0, 1.0e6, t, None)
    sigma = float(np.float32(5.67032e-8))
    pi = float(np.float32(2.0) * np.arcsin(np.float32(1.0)))
    assert abs(v / (sigma * t ** 4 / pi) - 1) < 1e-6
    # narrow band falls back on Simpson and must agree with the series route
    a = oracle.lib().sbdo_plkavg(1000.0, 1005.0, t, None)
    b = oracle.lib().sbdo_plkavg(0.0, 1005.0, t, None) - oracle.lib().sbdo_plkavg(0.0, 1000.0, t, None)
    assert abs(a / b - 1) < 2e-5


def test_asymtx_against_numpy_eig():
    import ctypes as C
    rng = np.random.default_rng(7)
    dp = C.POINTER(C.c_double)
    for m in (2, 3, 4, 8, 16):
        # product of two SPD matrices has a real positive spectrum (the
        # structure SOLEIG feeds to ASYMTX, disort.f:3236-3252)
        a = rng.normal(size=(m, m)); a = a @ a.T + m * np.eye(m)
        b = rng.normal(size=(m, m)); b = b @ b.T + m * np.eye(m)
        mat = a @ b
        aa = np.asfortranarray(mat.copy())
        evec = np.zeros((m, m), order="F")
        ev = np.zeros(m)
        wk = np.zeros(2 * m)
        ier = oracle.lib().sbdo_asymtx(aa.ctypes.data_as(dp), evec.ctypes.data_as(dp),
                                       ev.ctypes.data_as(dp), m, m, m, wk.ctypes.data_as(dp))
        assert ier == 0
        np.testing.assert_allclose(np.sort(ev), np.sort(np.linalg.eigvals(mat).real), rtol=1e-10)
        for j in range(m):
            v = evec[:, j]
            np.testing.assert_allclose(mat @ v, ev[j] * v, rtol=0,
                                       atol=1e-9 * np.abs(ev).max() * np.abs(v).max())
