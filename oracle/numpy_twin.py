"""Independent NumPy restatement of one QGMAP iteration.  TEST INFRASTRUCTURE ONLY.

Written separately from oracle/qgmap_oracle.c (vectorised over the belief grid, MATLAB array
semantics via np.roll / slicing) so that the two restatements check each other: the reference ships
no golden vectors and cannot run here (PARITY UNPINNED, see oracle/qgmap_oracle.h).

Cites: gqmap_gpu_mixture.m (G:) and gqmap_gpuSuper_mix_entropy.m (S:) of the reference.
"""
import numpy as np


def gauss_hermite(n):
    """GaussHermite_2.m:21-32 with numpy's eigh in the role of MATLAB's eig."""
    i = np.arange(1, n)
    a = np.sqrt(i / 2.0)
    CM = np.diag(a, 1) + np.diag(a, -1)
    vals, vecs = np.linalg.eigh(CM)
    ind = np.argsort(vals)
    x = vals[ind]
    w = np.sqrt(np.pi) * vecs[0, ind] ** 2
    return x, w


def tables(K):
    """G:8-10.  meshgrid(X): XI[r,c]=X[c], XJ[r,c]=X[r]; flattened column-major (k = r + K*c)."""
    X, W = gauss_hermite(K)
    XI, XJ = np.meshgrid(X, X)
    WI, WJ = np.meshgrid(W, W)
    f = lambda a: a.ravel(order="F")
    return dict(XI=f(XI), XJ=f(XJ), WIWJ=f(WI * WJ), XIXJ=f(XI * XJ), A=f(XI ** 2 + XJ ** 2), D=f(XI ** 2 - XJ ** 2))


def get_vv(V):
    """G:191-208 using 2-D indexing instead of linear indices."""
    M, N = V.shape
    VV = np.zeros((M + 2, N + 2))
    VV[1:-1, 1:-1] = V
    VV[0, :] = (3.0 * VV[1, :] - 3.0 * VV[2, :]) + VV[3, :]
    VV[-1, :] = (3.0 * VV[-2, :] - 3.0 * VV[-3, :]) + VV[-4, :]
    VV[:, 0] = (3.0 * VV[:, 1] - 3.0 * VV[:, 2]) + VV[:, 3]
    VV[:, -1] = (3.0 * VV[:, -2] - 3.0 * VV[:, -3]) + VV[:, -4]
    return VV


def _cubic_w(s):
    return (((2.0 - s) * s - 1.0) * s, (3.0 * s - 5.0) * s * s + 2.0, ((4.0 - 3.0 * s) * s + 1.0) * s, (s - 1.0) * s * s)


def node_pot(I1, VV, lambdad, epsn, x1, x2, i, j):
    """G:156-179, vectorised; i,j are 1-based row/col arrays (broadcastable with x1,x2)."""
    M, N = I1.shape
    Xq = np.minimum(np.maximum(j + x1, 1.0), N)
    Yq = np.minimum(np.maximum(i + x2, 1.0), M)
    ix = np.where(Xq <= 1.0, 1.0, np.where(Xq <= N - 1, np.floor(Xq), N - 1.0))
    iy = np.where(Yq <= 1.0, 1.0, np.where(Yq <= M - 1, np.floor(Yq), M - 1.0))
    so, to = Xq - ix, Yq - iy
    ws, wt = _cubic_w(so), _cubic_w(to)
    ixi, iyi = ix.astype(np.int64), iy.astype(np.int64)
    Vq = 0.0
    for c in range(4):          # VV (0-based) row iy-1+r+... : 1-based VV(iy+r, ix+c)
        for r in range(4):
            Vq = Vq + VV[iyi - 1 + r, ixi - 1 + c] * ws[c] * wt[r]
    Vq = Vq / 4.0
    i0 = np.broadcast_to(i, Vq.shape).astype(np.int64) - 1
    j0 = np.broadcast_to(j, Vq.shape).astype(np.int64) - 1
    return -lambdad * np.sqrt(epsn + (I1[i0, j0] - Vq) ** 2)


def _spectral(tab, a, u1, u2, o1, o2, p, pot, T, kappa_sign, guard_a0):
    """Common core of node_grad_spectral (G:87-116) / edge_grad_spectral (G:118-146).
    kappa_sign = -3 for node (entropy -3T..), +1 for edge (+T..)."""
    sqrt2, const1 = np.sqrt(2.0), 1.0 + np.log(2.0 * np.pi)
    s = (np.sqrt(1 + p) + np.sqrt(1 - p)) / 2
    t = (np.sqrt(1 + p) - np.sqrt(1 - p)) / 2
    pr = 1 - p ** 2
    sqrtpr = np.sqrt(pr)
    du1 = np.zeros_like(u1); du2 = np.zeros_like(u1); do1 = np.zeros_like(u1); do2 = np.zeros_like(u1)
    dp = np.zeros_like(u1); Ei = np.zeros_like(u1)
    for k in range(tab["XI"].size):
        XI, XJ, Wk, B, A, D = (tab[n][k] for n in ("XI", "XJ", "WIWJ", "XIXJ", "A", "D"))
        zi = s * XI + t * XJ
        zj = t * XI + s * XJ
        x1 = sqrt2 * o1 * zi + u1
        x2 = sqrt2 * o2 * zj + u2
        fval = Wk * pot(x1, x2)
        dp = dp + fval * (p - p * A + 2 * B)
        du1 = du1 + fval * (zi - p * zj)
        du2 = du2 + fval * (zj - p * zi)
        do1 = do1 + fval * (A - 1 + D / sqrtpr)
        do2 = do2 + fval * (A - 1 - D / sqrtpr)
        Ei = Ei + fval
    if guard_a0:
        z = (a == 0)
        du1, du2, do1, do2, dp = (np.where(z, 0.0, v) for v in (du1, du2, do1, do2, dp))
    du1 = a * du1 * (sqrt2 / (o1 * pr)) / np.pi
    du2 = a * du2 * (sqrt2 / (o2 * pr)) / np.pi
    kT = kappa_sign * T
    da = Ei / np.pi + kT * (const1 + np.log(sqrtpr * o1 * o2))
    do1 = a * (do1 / np.pi + kT) / o1
    do2 = a * (do2 / np.pi + kT) / o2
    dp = a * (dp / np.pi - kT * p) / pr
    return da, du1, du2, do1, do2, dp, a * da


def iteration_gradients(I1, VV, st, *, K, T, lambdad, lambdas, epsn, super_=False, guard_a0=True, row_off=0):
    """One gradient pass + assembly (G:29-40 / S:28-39).  `st` holds muu,muv,sigu,sigv,pn (M,N,L), rou (M,N,L,2,2),
    alpha (L,).  Returns dict of the 14 raw arrays, the assembled dmuu.. ('G_*'), dalpha, Energy, ptdmu, ptdsigma."""
    tab = tables(K)
    muu, muv, sigu, sigv, pn, rou = (st[n] for n in ("muu", "muv", "sigu", "sigv", "pn", "rou"))
    M, N, L = muu.shape
    alpha = np.asarray(st["alpha"]).reshape(1, 1, L)
    a3 = np.broadcast_to(alpha, (M, N, L))
    ms = np.arange(1 + row_off, M + 1 + row_off).reshape(M, 1, 1)   # row_off: the arrays are rows [row_off, row_off+M) of a larger grid
    ns = np.arange(1, N + 1).reshape(1, N, 1)

    if super_:
        def npot(x1, x2):
            acc = 0.0
            for di in range(-3, 1):
                for dj in range(-3, 1):
                    acc = acc + node_pot(I1, VV, lambdad, epsn, x1, x2, 4 * ms + di, 4 * ns + dj)   # S:99-104
            return acc
    else:
        def npot(x1, x2):
            return node_pot(I1, VV, lambdad, epsn, x1, x2, ms, ns)
    dan, dmuu, dmuv, dsu, dsv, dpn, nE = _spectral(tab, a3, muu, muv, sigu, sigv, pn, npot, T, -3.0, guard_a0)

    # G:31-34: (M,N,L,e,c) inputs built with repmat / cat / circshift
    a5 = np.broadcast_to(alpha.reshape(1, 1, L, 1, 1), (M, N, L, 2, 2))
    rep = lambda x, y: np.stack([np.stack([x, x], axis=3), np.stack([y, y], axis=3)], axis=4)
    sh = lambda x, y: np.stack([np.stack([np.roll(x, -1, 0), np.roll(x, -1, 1)], axis=3),
                                np.stack([np.roll(y, -1, 0), np.roll(y, -1, 1)], axis=3)], axis=4)
    epot = lambda x1, x2: -lambdas * np.sqrt(epsn + (x1 - x2) ** 2)
    dae, dmu1, dmu2, ds1, ds2, drou, eE = _spectral(tab, a5, rep(muu, muv), sh(muu, muv), rep(sigu, sigv),
                                                    sh(sigu, sigv), rou, epot, T, +1.0, guard_a0)
    I = (slice(1, M - 1), slice(1, N - 1))
    dalpha = dan[I].sum(axis=(0, 1)) + dae[I].sum(axis=(0, 1, 3, 4))                                 # G:36
    asm = lambda node, d1, d2, c: (node + d1[:, :, :, :, c].sum(axis=3) + np.roll(d2[:, :, :, 0, c], 1, 0)
                                   + np.roll(d2[:, :, :, 1, c], 1, 1))                               # G:37-40
    G_muu, G_muv = asm(dmuu, dmu1, dmu2, 0), asm(dmuv, dmu1, dmu2, 1)
    G_su, G_sv = asm(dsu, ds1, ds2, 0), asm(dsv, ds1, ds2, 1)
    Energy = nE[I].sum() + eE[I].sum()                                                               # G:48
    return dict(dan=dan, dmuu=dmuu, dmuv=dmuv, dsigmau=dsu, dsigmav=dsv, dpn=dpn, nEnergy=nE,
                dae=dae, dmu1=dmu1, dmu2=dmu2, dsigma1=ds1, dsigma2=ds2, drou=drou, eEnergy=eE,
                G_muu=G_muu, G_muv=G_muv, G_sigu=G_su, G_sigv=G_sv, dalpha=dalpha, Energy=Energy,
                ptdmu=np.abs(G_muu[I]).mean(), ptdsigma=np.abs(G_su[I]).mean())                      # G:69-70


def apply_update(st, g, step, *, minu, maxu, minv, maxv, sigma_max, corr_tor=1 - 1e-5):
    """G:41-46 on copies; returns new dict."""
    o = {k: np.array(v, copy=True) for k, v in st.items()}
    M, N, _ = o["muu"].shape
    I = (slice(1, M - 1), slice(1, N - 1))
    cl = lambda v, lo, hi: np.minimum(np.maximum(v, lo), hi)
    o["muu"][I] = cl(st["muu"][I] + g["G_muu"][I] * step, minu, maxu)
    o["muv"][I] = cl(st["muv"][I] + g["G_muv"][I] * step, minv, maxv)
    o["sigu"][I] = cl(st["sigu"][I] + g["G_sigu"][I] * step, 0.01, sigma_max)
    o["sigv"][I] = cl(st["sigv"][I] + g["G_sigv"][I] * step, 0.01, sigma_max)
    o["rou"][I] = cl(st["rou"][I] + g["drou"][I] * step, -corr_tor, corr_tor)
    o["pn"][I] = cl(st["pn"][I] + g["dpn"][I] * step, -corr_tor, corr_tor)
    return o


def neg_mixture(x, a, u, o):
    """legacy/findMixMax.m:31-38."""
    return -np.sum(a * np.exp(-(x - u) ** 2 / (2 * o ** 2)) / (2.5066282746310002 * o))
