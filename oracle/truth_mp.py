"""Extended-precision (mpmath, 60 digits) re-evaluation of the GP arithmetic -- TEST INFRASTRUCTURE ONLY.

Used to measure how far float64 implementations (the NumPy oracle, the CUDA path) sit from the exact
value of the same formulas on the same float64 inputs, i.e. each implementation's own rounding noise
(SURVEY.md fact 5, 8c(iii)).  Pure-Python O(n^3): keep n <= ~150.

Formulas follow BOBE/gp.py:80-96,124-178,450-466 and SURVEY.md appendix A.
"""
from __future__ import annotations

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def _kernel_entry(kind, xi, xj, ls, kv):
    q = mp.mpf(0)
    for a, b, l in zip(xi, xj, ls):
        df = (mp.mpf(float(a)) / l) - (mp.mpf(float(b)) / l)
        q += df * df
    if kind == "rbf":
        return kv * mp.exp(-q / 2), q
    r = mp.sqrt(q if q >= mp.mpf("1e-30") else mp.mpf("1e-30"))
    s5 = mp.sqrt(5)
    return kv * (1 + r * (s5 + r * mp.mpf(5) / 3)) * mp.exp(-s5 * r), q


def _cholesky(K):
    n = K.rows
    L = mp.zeros(n, n)
    for j in range(n):
        s = K[j, j] - mp.fsum(L[j, k] ** 2 for k in range(j))
        if s <= 0:
            raise ValueError("matrix not positive definite in extended precision")
        L[j, j] = mp.sqrt(s)
        for i in range(j + 1, n):
            L[i, j] = (K[i, j] - mp.fsum(L[i, k] * L[j, k] for k in range(j))) / L[j, j]
    return L


def _fwd(L, b):
    n = L.rows
    x = [mp.mpf(0)] * n
    for i in range(n):
        x[i] = (b[i] - mp.fsum(L[i, k] * x[k] for k in range(i))) / L[i, i]
    return x


def _bwd_t(L, b):
    n = L.rows
    x = [mp.mpf(0)] * n
    for i in reversed(range(n)):
        x[i] = (b[i] - mp.fsum(L[k, i] * x[k] for k in range(i + 1, n))) / L[i, i]
    return x


class TruthGP:
    """Exact-arithmetic GP on float64 inputs (X, standardised y, lengthscales, kv, noise)."""

    def __init__(self, kind, X, y_std_units, ls, kv, noise):
        self.kind = kind
        self.X = np.asarray(X, dtype=np.float64)
        self.n, self.d = self.X.shape
        self.ls = [mp.mpf(float(v)) for v in ls]
        self.kv = mp.mpf(float(kv))
        self.noise = mp.mpf(float(noise))
        self.y = [mp.mpf(float(v)) for v in np.asarray(y_std_units).ravel()]
        n = self.n
        self.K0 = mp.zeros(n, n)
        self.Q = mp.zeros(n, n)
        for i in range(n):
            for j in range(i + 1):
                v, q = _kernel_entry(kind, self.X[i], self.X[j], self.ls, self.kv)
                self.K0[i, j] = self.K0[j, i] = v
                self.Q[i, j] = self.Q[j, i] = q
        K = self.K0.copy()
        for i in range(n):
            K[i, i] += self.noise
        self.L = _cholesky(K)
        self.z = _fwd(self.L, self.y)
        self.alpha = _bwd_t(self.L, self.z)

    def mll(self):
        """log p(y) -- BOBE/gp.py:177."""
        quad = mp.fsum(a * b for a, b in zip(self.y, self.alpha))
        logdet = mp.fsum(mp.log(self.L[i, i]) for i in range(self.n))
        return float(-quad / 2 - logdet - mp.mpf(self.n) / 2 * mp.log(2 * mp.pi))

    def mll_grad(self, with_kv=True):
        """d log p / d log(l_j), d log p / d log(kv): 1/2 sum W_ik dK_ik."""
        n, d = self.n, self.d
        # K^-1 column by column
        Kinv = mp.zeros(n, n)
        for c in range(n):
            e = [mp.mpf(1) if i == c else mp.mpf(0) for i in range(n)]
            col = _bwd_t(self.L, _fwd(self.L, e))
            for i in range(n):
                Kinv[i, c] = col[i]
        s5 = mp.sqrt(5)
        g = [mp.mpf(0)] * (d + 1)
        for i in range(n):
            for k in range(n):
                W = self.alpha[i] * self.alpha[k] - Kinv[i, k]
                g[d] += W * self.K0[i, k]
                if i == k:
                    continue
                if self.kind == "rbf":
                    G = self.K0[i, k]
                else:
                    q = self.Q[i, k]
                    if q < mp.mpf("1e-30"):
                        continue
                    r = mp.sqrt(q)
                    G = self.kv * mp.mpf(5) / 3 * (1 + s5 * r) * mp.exp(-s5 * r)
                for j in range(d):
                    df = (mp.mpf(float(self.X[i, j])) - mp.mpf(float(self.X[k, j]))) / self.ls[j]
                    g[j] += W * G * df * df
        out = np.array([float(v / 2) for v in g])
        return out if with_kv else out[:d]

    def predict(self, Xq):
        """standardised mean and raw variance kk - ||L^-1 k*||^2 (no floor) -- BOBE/gp.py:455-464."""
        means, vars_ = [], []
        for x in np.atleast_2d(Xq):
            ks = [_kernel_entry(self.kind, xi, x, self.ls, self.kv)[0] for xi in self.X]
            means.append(float(mp.fsum(a * b for a, b in zip(ks, self.alpha))))
            v = _fwd(self.L, ks)
            vars_.append(float(self.kv + self.noise - mp.fsum(t * t for t in v)))
        return np.array(means), np.array(vars_)
