"""Dense fp64 dual active-set QP solver (numpy) — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

An independent check of the QP optimum the reference obtains from qpOASES (SolverMPC.cpp:702-712):

    min 1/2 x'Hx + g'x   s.t.  lbA <= A x <= ubA,   H symmetric positive definite.

It is also the executable specification of the algorithm the CUDA kernel implements
(hector_simulation_b200/csrc/hmpc_kernels.cu): Goldfarb-Idnani dual active-set iterations driven
through the explicit inverse Hessian, with the Schur complement S = A_W H^-1 A_W' of the working
set kept as a Cholesky factor.  The strictly convex QP has a unique minimiser, so any exact method
must agree with qpOASES up to its termination tolerance (1e9*EPS = 2.2e-7, Options.cpp:206).
"""
from __future__ import annotations

import numpy as np

BIG = 1e10  # bounds beyond this are treated as absent (the reference uses 5e10 as "infinity")


def solve(H, g, A, lbA, ubA, max_iter=500, tol=1e-9, triangle="upper"):
    """-> (x, info dict).  `triangle`: which triangle of a not-exactly-symmetric H defines it."""
    H = np.asarray(H, dtype=np.float64)
    n = H.shape[0]
    if triangle == "upper":
        Hs = np.triu(H) + np.triu(H, 1).T
    elif triangle == "lower":
        Hs = np.tril(H) + np.tril(H, -1).T
    else:
        Hs = 0.5 * (H + H.T)
    g = np.asarray(g, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    # one-sided constraint list  c_i' x >= d_i
    C, d = [], []
    for i in range(A.shape[0]):
        if lbA[i] > -BIG:
            C.append(A[i]); d.append(lbA[i])
        if ubA[i] < BIG:
            C.append(-A[i]); d.append(-ubA[i])
    C = np.array(C).reshape(-1, n); d = np.array(d)
    L = np.linalg.cholesky(Hs)
    Linv = np.linalg.inv(L)
    Hinv = Linv.T @ Linv
    x = -Hinv @ g
    W: list[int] = []
    lam = np.zeros(0)
    iters = 0
    scale = max(1.0, np.abs(x).max())
    status = 0
    while True:
        s = C @ x - d
        if W:
            s[W] = 0.0
        p = int(np.argmin(s))
        if s[p] >= -tol * scale:
            break
        lam_p = 0.0
        while True:
            iters += 1
            if iters > max_iter:
                status = 1
                break
            Hc = Hinv @ C[p]
            if W:
                CW = C[W]
                S = CW @ Hinv @ CW.T
                dvec = CW @ Hc
                r = np.linalg.solve(S, dvec)
                z = Hc - Hinv @ (CW.T @ r)
            else:
                r = np.zeros(0)
                z = Hc
            zn = float(C[p] @ z)
            cHc = float(C[p] @ Hc)
            t1, l = np.inf, -1
            for k in range(len(W)):
                if r[k] > 0 and lam[k] / r[k] < t1:
                    t1, l = lam[k] / r[k], k
            dependent = zn <= 1e-12 * max(cHc, 1e-300)
            t2 = np.inf if dependent else -(C[p] @ x - d[p]) / zn
            t = min(t1, t2)
            if not np.isfinite(t):
                status = 3
                break
            if not dependent:
                x = x + t * z
            lam = lam - t * r
            lam_p += t
            if t == t2:
                W.append(p)
                lam = np.append(lam, lam_p)
                break
            W.pop(l)
            lam = np.delete(lam, l)
        if status:
            break
    return x, dict(status=status, iters=iters, nactive=len(W), W=list(W), lam=lam)
