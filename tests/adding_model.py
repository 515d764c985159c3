"""numpy statement of the flux algorithm of the `adding` kernel (csrc/sbd_adding.cu).

Same discrete-ordinate equations as DISORT, solved without the
N L x N L boundary system (SETMTX / SOLVE0, disort.f:2702, :3322): every layer is reduced to its
reflection and transmission operators and the layers are combined by the interaction principle.

In the flux-weighted variables u^ = sqrt(w mu) u the sums s^ = u^+ + u^- and differences
d^ = u^+ - u^- obey  d s^/d tau = Po d^,  d d^/d tau = Pe s^  with the symmetric positive
definite operators Pe, Po of the eigenproblem (disort.f:3221-3269 in symmetric form).  With
Po = L L^T, L^T Pe L = V K^2 V^T and P = L V:

    A+ = P diag(coth(k h) / k) P^T,   A- = P diag(tanh(k h) / k) P^T,   h = dtau' / 2
    R + T = I - 2 (I + A+)^-1,        R - T = I - 2 (I + A-)^-1

evaluated in mode space, (I + P Lam P^T)^-1 = I - P (Lam^-1 + P^T P)^-1 P^T, where the diagonal terms
k tanh(kh) and k coth(kh) stay finite for k -> 0 (conservative scattering) and only grade the
diagonal for h -> 0: two SPD inversions without pivoting, absolute accuracy ~ eps cond(Po).  The interaction principle for a layer with particular solution p(tau):

    u+_top = R u-_top + T u+_bot + s_up,   s_up = p+_top - R p-_top - T p+_bot
    u-_bot = T u-_top + R u+_bot + s_dn,   s_dn = p-_bot - T p-_top - R p+_bot

Bottom-up pass ("everything below interface l reflects Rb_l and emits sb_l"), then a top-down
pass for the downward intensities, then FLUXES (disort.f:1926-2006) at the interfaces.

Test infrastructure: compared with the CPU oracle in tests/test_adding_math_cpu.py.
"""
import numpy as np

DITHER = 100.0 * 2.220446049250313e-16
PI = 3.1415927410125732          # the reference's single-precision pi (disort.f:441)


def legendre(nmax, x):
    x = np.atleast_1d(np.asarray(x, float))
    y = np.zeros((nmax, x.size))
    y[0] = 1.0
    if nmax > 1:
        y[1] = x
    for l in range(2, nmax):
        y[l] = ((2 * l - 1) * x * y[l - 1] - (l - 1) * y[l - 2]) / l
    return y


def fluxes(dtauc, ssalb, pmom, nstr, mu, wt, *, fbeam=0.0, umu0=1.0, albedo=0.0, fisot=0.0,
           pk=None, tplank=0.0, bplank=0.0):
    """pk: band-integrated Planck function at the L+1 levels (None: no thermal source);
    tplank = temis * B(ttemp), bplank = B(btemp).  mu, wt: the n Gauss nodes / weights on (0,1).
    Returns rfldir, rfldn, flup, dfdt, uavg at the L+1 levels."""
    N, n = nstr, nstr // 2
    L = len(dtauc)
    plank = pk is not None
    D = np.sqrt(wt * mu)
    csq = np.sqrt(wt / mu)
    yl = legendre(N, mu)                       # [l][i]
    y0 = legendre(N, -umu0)[:, 0]
    ss = np.where(ssalb == 1.0, 1.0 - DITHER, ssalb)
    dt = np.maximum(dtauc, 0.0)
    f = pmom[:, N]
    oprim = ss * (1 - f) / (1 - f * ss)
    dtp = (1 - f * ss) * dt
    tauc = np.concatenate([[0.0], np.cumsum(dtauc)])
    taup = np.concatenate([[0.0], np.cumsum(dtp)])
    # truncation (disort.f:2557-2605)
    ncut, abstau = L, 0.0
    for lc in range(L):
        if abstau < 10.0:
            ncut = lc + 1
        abstau += (1 - ss[lc]) * dt[lc]
    lyrcut = abstau >= 10.0 and not plank and L > 1
    if not lyrcut:
        ncut = L
    eb = np.exp(-taup / umu0) if fbeam > 0 else np.zeros(L + 1)
    edir = np.exp(-tauc / umu0) if fbeam > 0 else np.zeros(L + 1)

    Rl, Tl, sup, sdn = [], [], [], []
    for lc in range(ncut):
        pm = pmom[lc, :N].copy()
        pm[0] = 1.0
        gl = (2 * np.arange(N) + 1) * oprim[lc] * (pm - f[lc]) / (1 - f[lc])
        even = (np.arange(N) % 2) == 0
        Se = (yl[even].T * gl[even]) @ yl[even]
        So = (yl[~even].T * gl[~even]) @ yl[~even]
        Pe = np.diag(1 / mu) - np.outer(csq, csq) * Se
        Po = np.diag(1 / mu) - np.outer(csq, csq) * So
        Lc = np.linalg.cholesky(Po)
        k2, V = np.linalg.eigh(Lc.T @ Pe @ Lc)
        k = np.sqrt(np.abs(k2))
        P = Lc @ V
        h = 0.5 * dtp[lc]
        # mode space: (I + P Lam P^T)^-1 = I - P (Lam^-1 + G)^-1 P^T with G = P^T P -- the diagonal
        # terms k tanh(kh) (-> 0 for k -> 0) and k coth(kh) (-> 1/h) only grade the diagonal
        th = np.tanh(k * h)
        G = P.T @ P
        lam_p = k * th
        lam_m = np.where(th > 0, k / np.where(th > 0, th, 1.0), 1.0e300)
        Bp = np.linalg.inv(G + np.diag(lam_p))
        Bm = np.linalg.inv(G + np.diag(lam_m))
        Mp, Mm = P @ Bp @ P.T, P @ Bm @ P.T
        R = -np.eye(n) + Mp + Mm
        T = Mp - Mm
        # particular solutions, scaled (u^ = D u); directions: + up, - down
        pu_t = np.zeros(n); pd_t = np.zeros(n); pu_b = np.zeros(n); pd_b = np.zeros(n)
        # full kernel matrix over the N directions (appendix C item 3)
        mus = np.concatenate([mu, -mu])
        ws = np.concatenate([wt, wt])
        yall = legendre(N, mus)
        C = 0.5 * (yall.T * gl) @ yall * ws[None, :]
        if fbeam > 0:
            rhs = fbeam / (4 * PI) * (yall.T * gl) @ y0
            Z = np.linalg.solve(np.diag(1 + mus / umu0) - C, rhs)
            pu_t += D * Z[:n] * eb[lc]; pd_t += D * Z[n:] * eb[lc]
            pu_b += D * Z[:n] * eb[lc + 1]; pd_b += D * Z[n:] * eb[lc + 1]
        s_up = pu_t - R @ pd_t - T @ pu_b
        s_dn = pd_b - T @ pd_t - R @ pu_b
        if plank:
            # UPISOT (disort.f:4247): (I - C) 1 = (1 - w') 1 for the quadrature, so Z1 = b1 and
            # p+-(tau) = D B(tau) +- b1 q^ with q^ = Po^-1 D 1 (never singular for w' -> 1).  In
            # s_up / s_dn the b1 q^ terms combine to b1 (I + R - T) q^ = 2 b1 Mm q^, which stays
            # finite for dtau' -> 0 (b1 = dB / dtau' grows, Mm ~ h shrinks): no cancellation
            b1 = (pk[lc + 1] - pk[lc]) / dtp[lc] if dtp[lc] > 0 else 0.0
            e = 2.0 * b1 * (Mm @ np.linalg.solve(Po, D))
            ImR = np.eye(n) - R
            s_up += ImR @ (D * pk[lc]) - T @ (D * pk[lc + 1]) + e
            s_dn += ImR @ (D * pk[lc + 1]) - T @ (D * pk[lc]) - e
        Rl.append(R); Tl.append(T)
        sup.append(s_up)
        sdn.append(s_dn)

    # bottom boundary (disort.f:2919-2990, :3552-3578)
    if lyrcut:
        Rb = np.zeros((n, n)); sb = np.zeros(n)
    else:
        Rb = 2.0 * albedo * np.outer(D, D)
        sb = D * (albedo * umu0 * fbeam / PI * eb[ncut] + (1 - albedo) * bplank)
    Rbs, sbs, QT, qv = [None] * (ncut + 1), [None] * (ncut + 1), [None] * ncut, [None] * ncut
    Rbs[ncut], sbs[ncut] = Rb, sb
    for lc in range(ncut - 1, -1, -1):
        R, T = Rl[lc], Tl[lc]
        Y = np.linalg.solve(np.eye(n) - R @ Rb, np.column_stack([T, R @ sb + sdn[lc]]))
        QT[lc], qv[lc] = Y[:, :n], Y[:, n]
        W = Rb @ Y
        W[:, n] += sb
        New = np.column_stack([R, sup[lc]]) + T @ W
        Rb, sb = New[:, :n], New[:, n]
        Rbs[lc], sbs[lc] = Rb, sb
    out = {k: np.zeros(L + 1) for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg")}
    d = D * (fisot + tplank)
    for lev in range(ncut + 1):
        if lev > 0:
            d = QT[lev - 1] @ d + qv[lev - 1]
        u = Rbs[lev] @ d + sbs[lev]
        # the level belongs to the FIRST layer whose interval contains it (disort.f:2610-2625)
        lyr = next(lc for lc in range(L) if tauc[lc] <= tauc[lev] <= tauc[lc + 1])
        fldir = umu0 * fbeam * eb[lev]
        rfldir = umu0 * fbeam * edir[lev]
        flup = 2 * PI * (D @ u)
        fldn = 2 * PI * (D @ d)
        uavg = (2 * PI * (csq @ (u + d)) + fbeam * eb[lev]) / (4 * PI)
        out["rfldir"][lev] = rfldir
        out["rfldn"][lev] = fldn + fldir - rfldir
        out["flup"][lev] = flup
        out["uavg"][lev] = uavg
        out["dfdt"][lev] = (1 - ss[lyr]) * 4 * PI * (uavg - (pk[lev] if plank else 0.0))
    return out
