"""numpy statement of the two-stage tridiagonalisation the large-AE CUDA path implements
(saamge_b200/csrc/eigen_twostage.cuh), with the same storage conventions:

stage 1 (dense -> band, bandwidth b): for every panel j0 = 0, b, 2b, ... the block below the
  band, T[j0+b:n, j0:j0+b], is QR-factorised by Householder reflectors (reflector c has its unit
  entry in row j0+b+c; its tail overwrites T[j0+b+c+1:n, j0+c]; tau1[j0+c]); R stays in the band.
  The trailing block is updated two-sidedly, A22 <- Q^T A22 Q with Q = I - V Tf V^T:
  X = A22 V Tf, S = V^T X, Z = X - 1/2 V Tf^T S, A22 -= Z V^T + V Z^T.
stage 2 (band -> tridiagonal, bulge chasing): sweep s annihilates column s below the
  subdiagonal and chases the bulge down the band in steps of b rows; reflector (s, t) acts on
  rows i0 = s+1+t*b ... i0+L-1 and is stored in the packed lower triangle, column s:
  V2[cjm(s) + i0] = tau, V2[cjm(s) + i0 + a] = v_a (a >= 1), cjm(s) = s*n - s*(s-1)/2 - s.
back-transformation of an eigenvector y of the tridiagonal matrix: sweeps in reverse order
  (steps of one sweep commute), then the stage-1 reflectors in reverse order.

Used by tests/test_twostage_ref.py (CPU) to pin the algorithm against numpy.linalg.eigh and by
the GPU tests to compare intermediate results of the CUDA kernels."""
import numpy as np


def house(x):
    """dlarfg: returns (beta, tau, v) with v[0] = 1, (I - tau v v^T) x = beta e1."""
    alpha = x[0]
    xn2 = float(np.dot(x[1:], x[1:]))
    v = np.zeros_like(x)
    v[0] = 1.0
    if xn2 == 0.0:
        return alpha, 0.0, v
    beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return beta, tau, v


def stage1(A, b):
    """Returns (T, tau1): T holds the band (|i-j| <= b, lower part valid) and the reflectors."""
    n = A.shape[0]
    T = np.array(A, dtype=float, order="F")
    tau1 = np.zeros(n)
    j0 = 0
    while n - j0 - b >= 2:
        r = n - j0 - b
        nr = min(b, r - 1)
        P = T[j0 + b:, j0:j0 + b]  # view
        V = np.zeros((r, nr))
        for c in range(nr):
            beta, tau, v = house(P[c:, c].copy())
            tau1[j0 + c] = tau
            # apply to the remaining panel columns
            for q in range(c + 1, b):
                s = tau * np.dot(v, P[c:, q])
                P[c:, q] -= s * v
            P[c, c] = beta
            P[c + 1:, c] = v[1:]
            V[c:, c] = v
        # dlarft (forward, columnwise)
        G = V.T @ V
        Tf = np.zeros((nr, nr))
        for c in range(nr):
            Tf[c, c] = tau1[j0 + c]
            if c:
                Tf[:c, c] = -tau1[j0 + c] * (Tf[:c, :c] @ G[:c, c])
        A22 = T[j0 + b:, j0 + b:]
        Afull = np.tril(A22) + np.tril(A22, -1).T
        X = Afull @ V @ Tf
        S = V.T @ X
        Z = X - 0.5 * V @ (Tf.T @ S)
        Afull = Afull - Z @ V.T - V @ Z.T
        A22[:, :] = np.tril(Afull) + np.triu(A22, 1)  # only the lower part is maintained
        j0 += b
    return T, tau1


def extract_band(T, b):
    n = T.shape[0]
    Bd = np.zeros((2 * b, n), order="F")  # Bd[k, j] = A[j + k, j]
    for j in range(n):
        m = min(b, n - 1 - j)
        Bd[:m + 1, j] = T[j:j + m + 1, j]
    return Bd


def cjm(s, n):
    return s * n - (s * (s - 1)) // 2 - s


def stage2(Bd, b, n):
    """Bulge chasing on band storage Bd (2b x n, Bd[k, j] = A[j+k, j]).  Returns d, e, V2."""
    Bd = Bd.copy()
    V2 = np.zeros(n * (n + 1) // 2)

    def get(i, j):
        return Bd[i - j, j] if i >= j else Bd[j - i, i]

    def setv(i, j, x):
        if i >= j:
            Bd[i - j, j] = x
        else:
            Bd[j - i, i] = x

    def two_sided(i0, L, v, tau):
        D = np.array([[get(i0 + a, i0 + c) for c in range(L)] for a in range(L)])
        p = tau * (D @ v)
        K = 0.5 * tau * np.dot(p, v)
        q = p - K * v
        D = D - np.outer(v, q) - np.outer(q, v)
        for a in range(L):
            for c in range(a + 1):
                setv(i0 + a, i0 + c, D[a, c])

    if b >= 2:
        for s in range(n - 2):
            # step 0: column s
            i0 = s + 1
            L = min(b, n - i0)
            x = np.array([get(i0 + a, s) for a in range(L)])
            beta, tau, v = house(x)
            setv(i0, s, beta)
            for a in range(1, L):
                setv(i0 + a, s, 0.0)
            V2[cjm(s, n) + i0] = tau
            V2[cjm(s, n) + i0 + 1:cjm(s, n) + i0 + L] = v[1:]
            two_sided(i0, L, v, tau)
            st, Lp, vp, taup = i0, L, v, tau
            while True:
                i0 = st + b
                if i0 > n - 1:
                    break
                L = min(b, n - i0)
                O = np.array([[get(i0 + a, st + c) for c in range(Lp)] for a in range(L)])
                O = O - taup * np.outer(O @ vp, vp)  # right application of the previous reflector
                beta, tau, v = house(O[:, 0].copy())
                O[:, 1:] -= tau * np.outer(v, v @ O[:, 1:])
                O[0, 0] = beta
                O[1:, 0] = 0.0
                for a in range(L):
                    for c in range(Lp):
                        setv(i0 + a, st + c, O[a, c])
                V2[cjm(s, n) + i0] = tau
                V2[cjm(s, n) + i0 + 1:cjm(s, n) + i0 + L] = v[1:]
                two_sided(i0, L, v, tau)
                st, Lp, vp, taup = i0, L, v, tau
    d = Bd[0, :].copy()
    e = np.zeros(n)
    e[:n - 1] = Bd[1, :n - 1]
    return d, e, V2


def back_transform(y, T, tau1, V2, b):
    """z = Q1 Q2 y for one vector."""
    n = len(y)
    z = np.array(y, dtype=float)
    if b >= 2:
        for s in range(n - 3, -1, -1):
            i0 = s + 1
            while i0 <= n - 1:
                L = min(b, n - i0)
                tau = V2[cjm(s, n) + i0]
                v = np.ones(L)
                v[1:] = V2[cjm(s, n) + i0 + 1:cjm(s, n) + i0 + L]
                z[i0:i0 + L] -= tau * np.dot(v, z[i0:i0 + L]) * v
                i0 += b
    # stage-1 reflectors, last panel first, last column first
    j0s = []
    j0 = 0
    while n - j0 - b >= 2:
        j0s.append(j0)
        j0 += b
    for j0 in reversed(j0s):
        r = n - j0 - b
        nr = min(b, r - 1)
        for c in range(nr - 1, -1, -1):
            row = j0 + b + c
            v = np.ones(n - row)
            v[1:] = T[row + 1:, j0 + c]
            z[row:] -= tau1[j0 + c] * np.dot(v, z[row:]) * v
    return z


def tridiagonalise(A, b):
    T, tau1 = stage1(A, b)
    d, e, V2 = stage2(extract_band(T, b), b, A.shape[0])
    return d, e, T, tau1, V2
