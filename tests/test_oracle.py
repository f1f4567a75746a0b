"""Pins the CPU oracle (oracle/gp_oracle.py) by mathematics -- the reference holds no golden vectors for this
path and cannot be run here (no JAX): closed forms, identities, gradient three ways, extended precision, and
the committed golden fixtures (regeneration check)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O
from conftest import GOLDEN_DIR, mixed_err


def toy(n=30, d=3, seed=42):
    rng = np.random.RandomState(seed)
    X = rng.uniform(0, 1, size=(n, d))
    y = -np.sum((X - 0.5) ** 2, axis=1).reshape(-1, 1)
    return X, y


def test_kernel_basic_properties():
    X, _ = toy()
    for kern in (O.rbf_kernel, O.matern_kernel):
        K = kern(X, X, np.array([0.3, 0.5, 0.7]), 2.0, 1e-6, True)
        assert np.allclose(K, K.T, atol=1e-15)
        assert np.allclose(np.diag(K), 2.0 + 1e-6, rtol=1e-12)
        K0 = kern(X, X[:5], np.array([0.3, 0.5, 0.7]), 2.0, 1e-6, False)
        assert K0.shape == (30, 5) and np.all(K0 > 0) and np.all(K0 <= 2.0 + 1e-12)


def test_kernel_closed_form_entries():
    xa, xb = np.array([[0.1, 0.2]]), np.array([[0.4, 0.6]])
    ls, kv = np.array([0.5, 0.25]), 1.7
    q = ((0.1 - 0.4) / 0.5) ** 2 + ((0.2 - 0.6) / 0.25) ** 2
    assert math.isclose(O.rbf_kernel(xa, xb, ls, kv, 0.0, False)[0, 0], kv * math.exp(-0.5 * q), rel_tol=1e-14)
    r = math.sqrt(q)
    m = kv * (1 + math.sqrt(5) * r + 5 * q / 3) * math.exp(-math.sqrt(5) * r)
    assert math.isclose(O.matern_kernel(xa, xb, ls, kv, 0.0, False)[0, 0], m, rel_tol=1e-14)


def test_mll_closed_form_n1_n2():
    # n = 1: K = kv + noise, log p = -y^2/(2K) - 1/2 log K - 1/2 log 2 pi
    k = np.array([[2.5]])
    yv = np.array([[0.7]])
    ref = -0.5 * 0.49 / 2.5 - 0.5 * math.log(2.5) - 0.5 * math.log(2 * math.pi)
    assert math.isclose(O.gp_mll(k, yv, 1), ref, rel_tol=1e-14)
    # n = 2 by the explicit 2x2 inverse / determinant
    a, b, c = 2.0, 0.6, 1.5
    K = np.array([[a, b], [b, c]])
    yv = np.array([[0.3], [-1.1]])
    det = a * c - b * b
    quad = (c * 0.09 - 2 * b * 0.3 * (-1.1) + a * 1.21) / det
    ref = -0.5 * quad - 0.5 * math.log(det) - math.log(2 * math.pi)
    assert math.isclose(O.gp_mll(K, yv, 2), ref, rel_tol=1e-13)


def test_interpolation_and_noise_level_variance():
    X, y = toy(25, 2)
    gp = O.OracleGP(X, y, noise=1e-6, lengthscales=np.array([0.3, 0.3]))
    mean = gp.predict_mean_batched(X)
    assert np.max(np.abs(mean - y.ravel())) < 1e-4
    # exact identity: K0 alpha = y - noise * alpha  (standardised units)
    ms, _ = gp.predict_batched(X)
    assert np.max(np.abs(ms - (gp.train_y.ravel() - gp.noise * gp.alphas.ravel()))) < 1e-9
    var = gp.predict_var_batched(X)
    assert np.all(var > 0) and np.all(var < 1e-3)  # reference asserts < 1e-3 (tests/test_gp.py:139)


def test_standardisation_and_zero_std():
    ys, m, s = O.standardise(np.array([[1.0], [3.0]]))
    assert m == 2.0 and s == 1.0 and np.allclose(ys.ravel(), [-1, 1])
    ys, m, s = O.standardise(np.array([[5.0], [5.0]]))
    assert s == 1.0 and np.allclose(ys, 0)


def _torch_neg_mll(gp, lp):
    """Reverse-mode through torch.linalg.cholesky: the same construction jax.value_and_grad uses."""
    X = torch.tensor(gp.train_x)
    yv = torch.tensor(gp.train_y)
    lp = torch.tensor(lp, requires_grad=True)
    hp = torch.exp(lp)
    d = gp.ndim
    ls = hp[:d]
    kv = torch.tensor(gp.kernel_variance, dtype=torch.float64) if gp.fixed_kernel_variance else hp[d]
    xs = X / ls
    dsq = ((xs[:, None, :] - xs[None, :, :]) ** 2).sum(-1)
    if gp.kernel_name == "rbf":
        K = kv * torch.exp(-0.5 * dsq)
    else:
        r = torch.sqrt(torch.where(dsq < 1e-30, torch.full_like(dsq, 1e-30), dsq))
        K = kv * (1.0 + r * (O.SQRT5 + r * 5.0 / 3.0)) * torch.exp(-O.SQRT5 * r)
    K = K + gp.noise * torch.eye(X.shape[0], dtype=torch.float64)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(yv, L)
    mll = -0.5 * (yv.T @ alpha).squeeze() - torch.log(torch.diagonal(L)).sum() - 0.5 * X.shape[0] * math.log(2 * math.pi)
    # priors
    if gp.lengthscale_prior_spec == "SAAS":
        tausq = hp[-1]
        lp_kv = -0.5 * torch.log(kv) ** 2 - math.log(math.sqrt(2 * math.pi)) - torch.log(kv)
        hc = lambda z, s: math.log(2.0) - math.log(math.pi) - math.log(s) - torch.log1p((z / s) ** 2)
        mll = mll + lp_kv + hc(tausq, 0.1) + hc(1.0 / (tausq * ls ** 2), 1.0).sum()
    elif gp.lengthscale_prior_spec == "DSLP":
        loc, sc = O.SQRT2 + 0.5 * math.log(d), O.SQRT3
        mll = mll + (-0.5 * ((torch.log(ls) - loc) / sc) ** 2 - math.log(sc * math.sqrt(2 * math.pi)) - torch.log(ls)).sum()
        mll = mll - math.log(gp.kernel_variance_bounds[1] - gp.kernel_variance_bounds[0])
    else:
        mll = mll - d * math.log(gp.lengthscale_bounds[1] - gp.lengthscale_bounds[0])
        if not gp.fixed_kernel_variance:
            mll = mll - math.log(gp.kernel_variance_bounds[1] - gp.kernel_variance_bounds[0])
    loss = -mll
    loss.backward()
    return float(loss.detach()), lp.grad.numpy()


@pytest.mark.parametrize("kernel", ["rbf", "matern"])
@pytest.mark.parametrize("prior", [None, "DSLP", "SAAS", "fixed_kv"])
def test_gradient_three_ways(kernel, prior):
    X, y = toy(40, 3, seed=1)
    kw = {}
    if prior == "fixed_kv":
        kw = dict(kernel_variance_prior="fixed", kernel_variance=1.3)
    elif prior is not None:
        kw = dict(lengthscale_prior=prior)
    gp = O.OracleGP(X, y, kernel=kernel, noise=1e-6, **kw)
    rng = np.random.default_rng(3)
    lp = rng.uniform(-1.0, 0.5, gp.num_hyperparams)
    v, g = gp.neg_mll_and_grad(lp)
    assert math.isclose(v, gp.neg_mll(lp), rel_tol=1e-12)
    vt, gt = _torch_neg_mll(gp, lp)
    assert abs(v - vt) <= 1e-10 * max(abs(vt), 40)
    assert np.max(np.abs(g - gt)) <= 1e-8 * max(np.max(np.abs(gt)), 1.0)
    h = 1e-6
    for j in range(gp.num_hyperparams):
        e = np.zeros_like(lp)
        e[j] = h
        fd = (gp.neg_mll(lp + e) - gp.neg_mll(lp - e)) / (2 * h)
        assert abs(fd - g[j]) <= 2e-5 * max(abs(g[j]), 1.0)


def test_matern_clamp_has_zero_gradient_on_duplicates():
    X, y = toy(10, 2)
    X[1] = X[0]  # exact duplicate: dsq == 0 off the diagonal too, clamp active, gradient contribution 0
    y[1] = y[0]
    gp = O.OracleGP(X, y, kernel="matern", noise=1e-4)
    lp = np.log(gp.get_hyperparams())
    v, g = gp.neg_mll_and_grad(lp)
    vt, gt = _torch_neg_mll(gp, lp)
    assert np.all(np.isfinite(g)) and np.max(np.abs(g - gt)) <= 1e-8 * max(np.max(np.abs(gt)), 1.0)


def test_non_pd_gives_nan_not_exception():
    X, y = toy(20, 2)
    X[1] = X[0]
    gp = O.OracleGP(X, y, kernel="rbf", noise=0.0, lengthscales=np.array([5.0, 5.0]))
    lp = np.log(gp.get_hyperparams())
    v, g = gp.neg_mll_and_grad(lp)
    assert np.isnan(v) and np.all(np.isnan(g))


def test_fantasy_var_equals_predict_var_of_updated_gp():
    X, y = toy(35, 3, seed=7)
    for kernel in ("rbf", "matern"):
        gp = O.OracleGP(X, y, kernel=kernel, noise=1e-6, lengthscales=np.array([0.4, 0.6, 0.5]))
        mc = np.random.default_rng(0).uniform(0, 1, (20, 3))
        xn = np.array([0.21, 0.77, 0.4])
        ktm = gp.kernel(gp.train_x, mc, gp.lengthscales, gp.kernel_variance, gp.noise, False)
        fv = gp.fantasy_var(xn, mc, ktm)
        gp2 = O.OracleGP(X, y, kernel=kernel, noise=1e-6, lengthscales=np.array([0.4, 0.6, 0.5]))
        y_std0 = gp2.y_std
        gp2.train_x = np.vstack([gp2.train_x, xn])  # value-independent: keep the standardisation fixed
        gp2.train_y = np.vstack([gp2.train_y, [[0.123]]])
        gp2.recompute_cholesky()
        raw = np.maximum(gp2._raw_var(mc), O.SAFE_NOISE_FLOOR) * y_std0 ** 2
        assert mixed_err(fv, raw, y_std0 ** 2) < 1e-9
        shared = gp.fantasy_var_shared(np.vstack([xn, mc[:3]]), mc)
        assert mixed_err(shared[0], fv, y_std0 ** 2) < 1e-9
        assert mixed_err(shared[1], gp.fantasy_var(mc[0], mc, ktm), y_std0 ** 2) < 1e-9


@pytest.mark.parametrize("kernel", ["rbf", "matern"])
@pytest.mark.parametrize("std", [False, True])
def test_wipv_gradient_matches_central_differences(kernel, std):
    """d WIPV / dx_new and d WIPStd / dx_new (wipv_values_and_grad: the derivative jax.value_and_grad takes of
    BOBE/acquisition.py:438-440,463-465 in the n <= 500 polish) against central differences of the literal
    fantasy-variance path (BOBE/gp.py:552-576), which shares no code with the analytic gradient."""
    X, y = toy(60, 3, seed=0)
    gp = O.OracleGP(X, y, kernel=kernel, noise=1e-6, lengthscales=np.array([0.15, 0.2, 0.25]), kernel_variance=1.3)
    rng = np.random.default_rng(5)
    mc, cand = rng.uniform(0, 1, (40, 3)), rng.uniform(0, 1, (4, 3))
    ktm = gp.kernel(gp.train_x, mc, gp.lengthscales, gp.kernel_variance, gp.noise, False)

    def literal(c):  # mean_j of the reference's fantasy_var, one candidate at a time
        out = []
        for x in c:
            fv = gp.fantasy_var(x, mc, ktm)
            out.append(np.mean(np.sqrt(fv) if std else fv))
        return np.array(out)

    val, g = O.wipv_values_and_grad(gp, cand, mc, std)
    assert mixed_err(val, literal(cand), 1.0) < 1e-12
    h = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd = (literal(cand + e) - literal(cand - e)) / (2 * h)
        assert np.max(np.abs(g[:, k] - fd)) < 1e-8 * np.max(np.abs(g))
    # floored terms carry no gradient: a candidate on a training point with (almost) no noise
    gp0 = O.OracleGP(X, y, kernel=kernel, noise=1e-14, lengthscales=np.array([0.15, 0.2, 0.25]))
    v0, g0 = O.wipv_values_and_grad(gp0, X[:2], mc, std)
    assert np.all(np.isfinite(v0)) and np.all(np.isfinite(g0))


def test_fast_update_cholesky_matches_full_factor():
    X, y = toy(20, 2)
    gp = O.OracleGP(X, y, noise=1e-6)
    xn = np.array([[0.33, 0.9]])
    k = gp._k12(xn).ravel()
    Lnew = O.fast_update_cholesky(gp.cholesky, k, gp.kernel_variance + gp.noise)
    Xf = np.vstack([X, xn])
    Kf = gp.kernel(Xf, Xf, gp.lengthscales, gp.kernel_variance, gp.noise, True)
    assert np.max(np.abs(Lnew - np.linalg.cholesky(Kf))) < 1e-7


def test_oracle_against_extended_precision_truth():
    from oracle.truth_mp import TruthGP
    X, y = O.synthetic_training_set(40, 3)
    for kernel, ell in (("rbf", 0.6), ("matern", 0.7)):
        gp = O.OracleGP(X, y, kernel=kernel, lengthscales=np.full(3, ell))
        T = TruthGP(gp.kernel_name, X, gp.train_y, gp.lengthscales, gp.kernel_variance, gp.noise)
        Xq = O.synthetic_queries(10, 3)
        mt, vt = T.predict(Xq)
        ms, _ = gp.predict_batched(Xq)
        assert mixed_err(ms, mt, 1.0) < 1e-9
        assert mixed_err(gp._raw_var(Xq), vt, 1.0) < 1e-9
        lp = np.log(gp.get_hyperparams())
        v, g = gp.neg_mll_and_grad(lp)
        pl, pg = gp.log_prior_and_grad(lp)
        assert abs((-v - pl) - T.mll()) < 1e-9 * 40
        gt = T.mll_grad()
        assert np.max(np.abs((-g - pg) - gt)) < 1e-7 * np.max(np.abs(gt))


def test_log_ei_against_mpmath():
    import mpmath as mp
    mp.mp.dps = 50
    us = np.array([3.0, 0.5, -0.5, -1.0, -1.5, -5.0, -30.0, -1e3, -2e6])
    got = O.log_ei_helper(us)
    for u, gv in zip(us, got):
        um = mp.mpf(float(u))
        if u > -1e4:
            ref = mp.log(mp.npdf(um) + um * mp.ncdf(um))
        else:  # asymptotic series of the Mills ratio, accurate far in the tail
            ref = mp.log(mp.npdf(um)) - 2 * mp.log(-um) + mp.log(1 - 3 / um ** 2 + 15 / um ** 4)
        tol = 1e-12 if u > -1e6 else 1e-6  # below -1e6 the reference switches to the leading term only
        assert abs(float(ref) - gv) <= tol * abs(float(ref)), (u, gv, float(ref))
    assert np.all(O.ei_values(np.array([0.1, -3.0]), np.array([[0.04], [1e-30]]), 0.0, 0.0) <= 0)


def test_priors_logprob_values():
    assert math.isclose(float(O.lognormal_logprob(2.0, 0.3, 1.7)),
                        -0.5 * ((math.log(2) - 0.3) / 1.7) ** 2 - math.log(1.7 * math.sqrt(2 * math.pi)) - math.log(2))
    assert math.isclose(float(O.halfcauchy_logprob(0.5, 0.1)), math.log(2 / (math.pi * 0.1 * (1 + 25))))
    assert math.isclose(float(O.uniform_logprob(0.3, 0.01, 5)), -math.log(4.99))


@pytest.mark.parametrize("name", ["A_banana_rbf_n100_d2", "M_matern_n300_d3", "B_rbf_n500_d2", "B_rbf_n500_d4", "B_rbf_n500_d6"])
def test_oracle_against_the_exact_values_in_the_fixtures(name):
    """The float64 oracle against the 60-digit (mpmath) values stored beside its outputs: posterior mean, posterior variance
    and the log marginal likelihood of EVERY restart row.  These gaps are the oracle's own rounding noise;
    the GPU tests hold the CUDA path to max(tolerance, 2 x this gap) against the same exact values."""
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    from oracle.gen_golden import make_case
    gp, X, y, Xq, x0, mc, cand = make_case(name)
    n = X.shape[0]
    e_mean = mixed_err(gold["mean_std"][:32], gold["truth_mean_std"], 1.0)
    e_var = mixed_err(gold["var_std"][:32], np.maximum(gold["truth_var_raw"], 1e-12), 1.0)
    gaps = []
    for r in range(x0.shape[0]):
        pl = gp.log_prior_and_grad(x0[r])[0]
        tr = float(gold["truth_mll_rows"][r])
        gaps.append(abs((-gold["neg_mll"][r] - pl) - tr) / max(abs(tr), n))
    print(f"\n[{name}] oracle vs exact: mean {e_mean:.1e} var {e_var:.1e} log-ML rows {[f'{g:.1e}' for g in gaps]}")
    assert e_mean < 1e-9 and e_var < 1e-7 and max(gaps) < 5e-9
    assert abs(float(gold["truth_mll_rows"][0]) - float(gold["truth_mll"])) <= 1e-12 * n  # row 0 = the current hyper-parameters


@pytest.mark.parametrize("name", ["A_banana_rbf_n100_d2", "M_matern_n300_d3", "B_rbf_n500_d4"])
def test_golden_fixtures_are_reproduced_by_the_oracle(name):
    from oracle.gen_golden import make_case
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    gp, X, y, Xq, x0, mc, cand = make_case(name)
    # BLAS thread counts differ between machines: the ill-conditioned cases move at the 1e-9 level
    assert mixed_err(gp.predict_mean_batched(Xq), gold["mean"], gp.y_std) < 5e-9
    assert mixed_err(gp.predict_var_batched(Xq), gold["var"], gp.y_std ** 2) < 1e-8
    v, g = gp.neg_mll_and_grad(x0[0])
    assert abs(v - gold["neg_mll"][0]) < 1e-8 * max(abs(v), X.shape[0])
    assert mixed_err(gp.fantasy_var_shared(cand, mc), gold["fantasy"], gp.y_std ** 2) < 1e-8
    if "truth_mll" in gold.files:
        # the oracle's own rounding noise at cond(K) ~ 3e9, for the record
        pl = gp.log_prior_and_grad(x0[0])[0]
        assert abs((-gold["neg_mll"][0] - pl) - float(gold["truth_mll"])) < 1e-8 * X.shape[0]
        assert mixed_err(gold["mean_std"][:32], gold["truth_mean_std"], 1.0) < 1e-9


@pytest.mark.parametrize("kernel", ["rbf", "matern"])
def test_oracle_input_gradients_match_central_differences(kernel):
    """Pins OracleGP.predict_grad_batched against central differences of the oracle's own
    predict_mean / predict_var, both flavours."""
    rng = np.random.default_rng(5)
    n, d = 60, 3
    X = rng.uniform(0, 1, (n, d))
    y = np.sin(4 * X.sum(1, keepdims=True)) + 0.3 * X[:, :1]
    gp = O.OracleGP(X, y, kernel=kernel, noise=1e-6, lengthscales=np.array([0.4, 0.7, 0.55]), kernel_variance=1.3)
    xq = rng.uniform(0.05, 0.95, (7, d))
    h = 1e-5  # |alpha| ~ 3e3 here: the difference quotient itself is only good to ~1e-7 relative
    for std in (False, True):
        mean, var, dm, dv = gp.predict_grad_batched(xq, standardised=std)
        f = (lambda z: gp.predict_batched(z)) if std else (lambda z: (gp.predict_mean_batched(z), gp.predict_var_batched(z)))
        m0, v0 = f(xq)
        assert np.allclose(mean, m0, rtol=1e-9, atol=1e-9) and np.allclose(var, np.ravel(v0), rtol=1e-8, atol=1e-12)
        for k in range(d):
            e = np.zeros(d); e[k] = h
            mp, vp = f(xq + e); mm, vm = f(xq - e)
            assert np.allclose(dm[:, k], (np.ravel(mp) - np.ravel(mm)) / (2 * h), rtol=5e-6, atol=1e-6)
            assert np.allclose(dv[:, k], (np.ravel(vp) - np.ravel(vm)) / (2 * h), rtol=5e-5, atol=1e-7)
    # at a training point the variance sits on the floor: zero gradient, like jnp.clip / jnp.where
    _, v, _, dv = gp.predict_grad_batched(X[:2], standardised=True)
    gp0 = O.OracleGP(X, y, kernel=kernel, noise=1e-14, lengthscales=np.array([0.4, 0.7, 0.55]), kernel_variance=1.3)
    _, v, _, dv = gp0.predict_grad_batched(X[:2], standardised=True)
    assert np.all(dv[v <= 1e-12] == 0.0)


def test_oracle_svm_decision_matches_sklearn():
    """Pins the restated SVM decision function (BOBE/clf.py:188-209) against scikit-learn's own decision_function --
    the third-party routine the reference extracts its parameters from (clf.py:42-47)."""
    from sklearn.svm import SVC
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (200, 4))
    labels = (np.sum((X - 0.5) ** 2, axis=1) < 0.2).astype(int)
    clf = SVC(kernel="rbf", gamma="scale", C=1e7).fit(X, labels)
    xq = rng.uniform(0, 1, (300, 4))
    dec = O.svm_decision(xq, clf.support_vectors_, clf.dual_coef_[0], float(clf.intercept_[0]), float(clf._gamma))
    ref = clf.decision_function(xq)
    assert np.allclose(dec, ref, rtol=1e-9, atol=1e-6 * np.abs(ref).max())
    assert np.array_equal(dec >= 0, clf.predict(xq) == 1)


@pytest.mark.parametrize("kernel", ["rbf", "matern"])
def test_oracle_against_sklearn_gpr(kernel):
    """External pin: scikit-learn's GaussianProcessRegressor (an independent, widely used float64 implementation of the
    same model: ARD RBF / Matern-5/2 x constant + fixed diagonal noise) must give the oracle's posterior mean, variance,
    log marginal likelihood and its gradient with respect to the log hyper-parameters.  It is not the reference, but
    it shares no code with the oracle (sklearn: cho_solve on its own kernel classes; gradient via its own einsum)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    rng = np.random.default_rng(7)
    n, d = 120, 4
    X = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * X.sum(1, keepdims=True)) + X[:, :1] ** 2
    ls, kv, noise = np.array([0.35, 0.6, 0.8, 0.5]), 1.7, 1e-6
    gp = O.OracleGP(X, y, kernel=kernel, noise=noise, lengthscales=ls, kernel_variance=kv)
    base = RBF(length_scale=ls) if kernel == "rbf" else Matern(length_scale=ls, nu=2.5)
    sk = GaussianProcessRegressor(kernel=ConstantKernel(kv) * base, alpha=noise, optimizer=None, normalize_y=False)
    sk.fit(X, gp.train_y.ravel())  # the standardised targets the GP works on (BOBE/gp.py:296-306)
    Xq = rng.uniform(0, 1, (200, d))
    m_sk, s_sk = sk.predict(Xq, return_std=True)
    m_or, v_or = gp.predict_batched(Xq)
    assert np.allclose(m_or, m_sk, rtol=1e-8, atol=1e-8)
    # sklearn's predictive variance has no noise term in k(x*, x*): add it (BOBE/gp.py:463 uses kv + noise)
    assert np.allclose(v_or.ravel(), s_sk ** 2 + noise, rtol=1e-6, atol=1e-9)
    theta = np.concatenate([[np.log(kv)], np.log(ls)])  # sklearn order: constant first, then the lengthscales
    lml, g_sk = sk.log_marginal_likelihood(theta, eval_gradient=True)
    val, grad = gp.neg_mll_and_grad(np.concatenate([np.log(ls), [np.log(kv)]]))
    lp, glp = gp.log_prior_and_grad(np.concatenate([np.log(ls), [np.log(kv)]]))
    assert abs((-val - lp) - lml) < 1e-8 * max(abs(lml), n)
    g_or = -grad - glp  # d log p(y) / d (log l_1..d, log kv)
    assert np.allclose(g_or[:d], g_sk[1:], rtol=1e-6, atol=1e-6 * np.abs(g_sk).max())
    assert np.allclose(g_or[d], g_sk[0], rtol=1e-6, atol=1e-6 * np.abs(g_sk).max())


# ---- the oracle against vectors produced by the REFERENCE'S OWN SOURCE (oracle/gen_reference_vectors.py) --------------------
def _ref_vectors():
    return np.load(os.path.join(GOLDEN_DIR, "reference_source_vectors.npz"))


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def test_oracle_kernels_match_the_reference_source():
    """dist_sq / kernel_diag / rbf_kernel / matern_kernel / gp_mll / fast_update_cholesky: BOBE/gp.py:80-197 executed from
    the reference's own file (NumPy stand-in for jax.numpy) vs the restatement -- the same float64 operations, so the
    agreement is at the last bits."""
    v = _ref_vectors()
    xa, xb, ls, kv, noise = v["k_xa"], v["k_xb"], v["k_ls"], float(v["k_kv"]), float(v["k_noise"])
    assert np.array_equal(O.dist_sq(xa, xb), v["k_dist_sq"])
    assert _rel(O.rbf_kernel(xa, xb, ls, kv, noise, False), v["k_rbf_cross"]) < 1e-15
    assert _rel(O.matern_kernel(xa, xb, ls, kv, noise, False), v["k_matern_cross"]) < 1e-15
    assert _rel(O.rbf_kernel(xa, xa, ls, kv, noise, True), v["k_rbf_square"]) < 1e-15
    assert _rel(O.matern_kernel(xa, xa, ls, kv, noise, True), v["k_matern_square"]) < 1e-15
    assert np.array_equal(O.kernel_diag(xa, kv, noise, True), v["k_diag_noise"])
    assert np.array_equal(O.kernel_diag(xa, kv, noise, False), v["k_diag_plain"])
    for kern in ("rbf", "matern"):
        got = O.gp_mll(v[f"mll_{kern}_K"], v["mll_y_std"], 60)
        assert abs(got - float(v[f"mll_{kern}_value"])) < 1e-11 * abs(float(v[f"mll_{kern}_value"]))
    newL = O.fast_update_cholesky(v["chol_L"], v["chol_k"], float(v["chol_kself"]))
    assert _rel(newL, v["chol_new_L"]) < 1e-14 and np.array_equal(np.triu(newL, 1), np.zeros_like(newL))


@pytest.mark.parametrize("p,kern", [("gp_rbf_", "rbf"), ("gp_matern_", "matern"), ("gpB_rbf_", "rbf")])
def test_oracle_gp_class_matches_the_reference_source(p, kern):
    """class GP of the reference (constructor, standardisation, Cholesky + alphas, every predict variant, neg_mll with its
    default priors, fantasy_var, WIPV / WIPStd / EI / LogEI values, update with a duplicate) vs OracleGP on the same inputs."""
    v = _ref_vectors()
    X, y, ls, Xq = v[p + "X"], v[p + "y"], v[p + "ls"], v[p + "Xq"]
    d = X.shape[1]
    gp = O.OracleGP(X, y[:, None], noise=float(v[p + "noise"]), kernel=kern, lengthscales=ls, kernel_variance=float(v[p + "kv"]))
    assert abs(gp.y_mean - float(v[p + "y_mean"])) < 1e-14 * abs(gp.y_mean) and abs(gp.y_std - float(v[p + "y_std"])) < 1e-14 * gp.y_std
    assert _rel(gp.train_y, v[p + "train_y"]) < 1e-14
    condL = float(v[p + "cond_L"])
    if p + "cholesky" in v.files:
        assert _rel(gp.cholesky, v[p + "cholesky"]) < 1e-15 * condL
    assert abs(float(np.sum(np.log(np.diag(gp.cholesky)))) - float(v[p + "logdet_half"])) < 1e-14 * condL * X.shape[0]
    assert _rel(gp.alphas, v[p + "alphas"]) < 1e-15 * condL ** 2
    scale = max(abs(float(v[p + "y_mean"])), float(v[p + "y_std"]))
    amp = float(np.abs(v[p + "alphas"]).max())  # the mean is a sum of ~n terms of size alpha: rounding ~ n eps |alpha|
    tol_mean = 200 * 2.3e-16 * amp * float(v[p + "y_std"]) + 1e-15 * scale
    assert np.max(np.abs(gp.predict_mean_batched(Xq) - v[p + "mean_batched"])) < tol_mean
    prior_var = float(v[p + "y_std"]) ** 2 * float(v[p + "kv"])
    tol_var = 1e-15 * condL ** 2 * prior_var
    assert np.max(np.abs(gp.predict_var_batched(Xq) - v[p + "var_batched"])) < tol_var
    assert abs(float(gp.predict_mean_single(Xq[3])) - float(v[p + "mean_single"])) < tol_mean
    assert abs(float(gp.predict_var_single(Xq[3])) - float(v[p + "var_single"])) < tol_var
    ms, vs = gp.predict_batched(Xq)
    assert np.max(np.abs(np.ravel(ms) - v[p + "std_mean_batched"])) < tol_mean / float(v[p + "y_std"]) + 1e-15
    assert np.max(np.abs(np.ravel(vs) - v[p + "std_var_batched"])) < tol_var / float(v[p + "y_std"]) ** 2
    assert float(v[p + "var_batched"][0]) < 1e-4 * prior_var  # (the query that is a training point)
    # log marginal likelihood with the default (Uniform) priors, and the analytic gradient against the reference's
    # neg_mll differenced centrally (jax autodiff is not available under the stand-in)
    lp = v[p + "log_params"]
    for r in range(lp.shape[0]):
        val, grad = gp.neg_mll_and_grad(lp[r])
        assert abs(val - float(v[p + "neg_mll"][r])) < 1e-10 * max(1.0, abs(val)), (r, val, float(v[p + "neg_mll"][r]))
        assert abs(float(gp.neg_mll(lp[r])) - val) < 1e-12 * max(1.0, abs(val))
        # jax.value_and_grad(neg_mll) of the reference (reverse-mode autodiff through its own statements, torch-backed
        # stand-in) vs the analytic gradient: 1e-7 (scale max|grad|), three times that at the cond(K) ~ 1e10 shape where two
        # float64 evaluations of the same gradient differ by ~1e-7 themselves
        ad = v[p + "neg_mll_ad_grad"][r]
        assert abs(val - float(v[p + "neg_mll_ad"][r])) < 3e-9 * max(abs(val), X.shape[0])
        assert np.max(np.abs(grad - ad)) < (1e-7 if condL < 1e4 else 3e-7) * max(1.0, float(np.max(np.abs(ad)))), (r,)
        if condL < 1e4:  # (central differences carry cond(K) eps / h of noise: only at the well-conditioned shapes)
            assert np.max(np.abs(grad - v[p + "neg_mll_fd_grad"][r])) < 2e-5 * max(1.0, float(np.max(np.abs(grad))))
    # fantasy variance, integrated acquisitions, EI / LogEI
    mc, cand = v[p + "mc"], v[p + "cand"]
    k_train_mc = gp.kernel(gp.train_x, mc, gp.lengthscales, gp.kernel_variance, gp.noise, False)
    fv = np.stack([gp.fantasy_var(c, mc, k_train_mc) for c in cand])
    assert np.max(np.abs(fv - v[p + "fantasy_var"])) < tol_var
    assert np.max(np.abs(fv.mean(axis=1) - v[p + "wipv"])) < tol_var
    assert np.max(np.abs(np.sqrt(fv).mean(axis=1) - v[p + "wipstd"])) < tol_var / np.sqrt(float(np.min(v[p + "fantasy_var"])))
    assert np.max(np.abs(O.wipv_values(gp, cand, mc) - v[p + "wipv"])) < tol_var
    assert np.max(np.abs(gp.fantasy_var_shared(cand, mc) - v[p + "fantasy_var"])) < 10 * tol_var
    xe, best_y, zeta = v[p + "ei_x"], float(v[p + "ei_best_y"]), float(v[p + "ei_zeta"])
    mu, var = gp.predict_batched(xe)
    ei, logei = O.ei_values(np.ravel(mu), np.ravel(var), best_y, zeta), O.logei_values(np.ravel(mu), np.ravel(var), best_y, zeta)
    assert np.max(np.abs(ei - v[p + "ei"])) < 1e-9 * max(1e-300, float(np.max(np.abs(v[p + "ei"])))) + 1e-13
    # (log EI ~ -u^2 / 2 with u^2 ~ 1 / var: it carries the RELATIVE error of a variance near the noise floor)
    assert np.max(np.abs(logei - v[p + "logei"]) / np.maximum(1.0, np.abs(v[p + "logei"]))) < max(1e-8, 1e-16 * condL ** 2)
    # update(): two new points, one duplicate (BOBE/gp.py:495-541)
    gp.update(v[p + "upd_new_x"], v[p + "upd_new_y"])
    assert gp.train_x.shape == v[p + "upd_train_x"].shape and np.array_equal(gp.train_x, v[p + "upd_train_x"])
    assert abs(gp.y_mean - float(v[p + "upd_y_mean"])) < 1e-14 * abs(gp.y_mean) and abs(gp.y_std - float(v[p + "upd_y_std"])) < 1e-14 * gp.y_std
    condU = float(v[p + "upd_cond_L"])
    assert _rel(gp.alphas, v[p + "upd_alphas"]) < 1e-15 * condU ** 2
    if p + "upd_cholesky" in v.files:
        assert _rel(gp.cholesky, v[p + "upd_cholesky"]) < 1e-15 * condU
    d_alpha = float(np.max(np.abs(gp.alphas - v[p + "upd_alphas"])))  # (the mean is k*^T alpha: it inherits alpha's rounding)
    assert np.max(np.abs(gp.predict_mean_batched(Xq[:6]) - v[p + "upd_mean_batched"])) < 10 * tol_mean + X.shape[0] * d_alpha * gp.y_std


def test_oracle_logei_helper_matches_the_reference_source():
    """BOBE/acquisition.py:21-75 over its three branches (u > -1, asymptotic, u < -1e6)."""
    v = _ref_vectors()
    u = v["logei_u"]
    got, want = O.log_ei_helper(u), v["logei_helper"]
    assert np.all(np.isfinite(want)) and np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 1e-13
    assert np.max(np.abs(O._ei_helper(u) - v["ei_helper"])) < 1e-15 * float(np.max(np.abs(v["ei_helper"])))


def test_oracle_svm_mask_matches_the_reference_source():
    """GPwithClassifier(clf_type='svm') of the reference (BOBE/clf_gp.py:16-205, BOBE/clf.py:36-83,188-214): the SVM decision
    function, the mask, and the masked mean / variance in both scalings vs the restatement, with the reference's own trained
    classifier parameters."""
    v = _ref_vectors()
    X, y, Xq = v["clf_X"], v["clf_y"], v["clf_Xq"]
    keep = y > y.max() - float(v["clf_gp_threshold"])  # BOBE/clf_gp.py:84-90
    assert np.array_equal(X[keep], v["clf_gp_train_x"])
    gp = O.OracleGP(X[keep], y[keep][:, None], noise=float(v["clf_noise"]), kernel="rbf", lengthscales=v["clf_ls"],
                    kernel_variance=float(v["clf_kv"]))
    assert abs(gp.y_mean - float(v["clf_y_mean"])) < 1e-13 * abs(gp.y_mean) and abs(gp.y_std - float(v["clf_y_std"])) < 1e-13 * gp.y_std
    params = {"support_vectors": v["clf_support_vectors"], "dual_coef": v["clf_dual_coef"], "intercept": float(v["clf_intercept"]),
              "gamma_eff": float(v["clf_gamma"])}
    dec = O.svm_decision(Xq, params["support_vectors"], params["dual_coef"], params["intercept"], params["gamma_eff"])
    assert np.max(np.abs(dec - v["clf_decision"])) < 1e-9 * max(1.0, float(np.max(np.abs(v["clf_decision"]))))
    assert np.array_equal((dec >= 0).astype(float), v["clf_mask"]) and 0 < v["clf_mask"].sum() < Xq.shape[0]
    m, var, _ = O.clf_masked_predict(gp, Xq, params, minus_inf=float(v["clf_minus_inf"]))
    assert mixed_err(m, v["clf_mean_batched"], gp.y_std) < 1e-11 and mixed_err(var, v["clf_var_batched"], gp.y_std ** 2) < 1e-11
    ms, vs, _ = O.clf_masked_predict(gp, Xq, params, minus_inf=float(v["clf_minus_inf"]), standardised=True)
    assert mixed_err(np.ravel(ms), v["clf_std_mean_batched"], 1.0) < 1e-11 and mixed_err(vs, v["clf_std_var_batched"], 1.0) < 1e-11
    # the same classifier comes out of scikit-learn when trained here on the reference's labels (BOBE/clf_gp.py:143-147)
    from sklearn.svm import SVC
    labels = np.where(y < y.max() - float(v["clf_threshold"]), 0, 1)
    clf = SVC(kernel="rbf", gamma="scale", C=1e7).fit(X, labels)
    assert np.allclose(clf.support_vectors_, v["clf_support_vectors"]) and np.allclose(clf.dual_coef_[0], v["clf_dual_coef"])


@pytest.mark.parametrize("p,kern", [("gp_rbf_", "rbf"), ("gp_matern_", "matern"), ("gpB_rbf_", "rbf")])
def test_oracle_input_gradients_match_the_reference_autodiff(p, kern):
    """d/dx of the standardised posterior mean / variance (BOBE/gp.py:476-489) and of WIPV / WIPStd
    (BOBE/acquisition.py:438-440,463-465 through GP.fantasy_var) as jax.value_and_grad yields them -- reverse mode through the
    reference's own statements (torch-backed stand-in) -- vs the analytic forms of the restatement."""
    v = _ref_vectors()
    gp = O.OracleGP(v[p + "X"], v[p + "y"][:, None], noise=float(v[p + "noise"]), kernel=kern, lengthscales=v[p + "ls"],
                    kernel_variance=float(v[p + "kv"]))
    xg, mc = v[p + "acq_grad_x"], v[p + "mc"]
    ill = float(v[p + "cond_L"]) > 1e4
    tol = 3e-6 if ill else 1e-9  # (relative to the largest gradient entry; at cond(K) ~ 1e10 float64 itself gives ~1e-6)
    m, var, dm, dv = gp.predict_grad_batched(xg, standardised=True)
    assert np.max(np.abs(np.ravel(m) - v[p + "pmean_ad_value"])) < 1e-9 * max(1.0, float(np.max(np.abs(v[p + "pmean_ad_value"]))))
    for got, key in ((dm, "pmean_ad_grad"), (dv, "pvar_ad_grad")):
        want = v[p + key]
        assert np.max(np.abs(got - want)) < tol * float(np.max(np.abs(want))), key
    for std, key in ((False, "wipv"), (True, "wipstd")):
        vals, grads = O.wipv_values_and_grad(gp, xg, mc, std=std)
        assert np.max(np.abs(vals - v[p + key + "_ad_value"]) / np.abs(v[p + key + "_ad_value"])) < (1e-7 if ill else 1e-11)
        want = v[p + key + "_ad_grad"]
        assert np.max(np.abs(grads - want)) < tol * float(np.max(np.abs(want))), key


_PRIOR_CASES = (("prior_dslp_", dict(lengthscale_prior="DSLP", kernel_variance_prior={"name": "LogNormal", "loc": 0.0, "scale": 1.0})),
                ("prior_saas_", dict(lengthscale_prior="SAAS", tausq=0.7)),
                ("prior_fixedkv_", dict(lengthscale_prior="DSLP", kernel_variance_prior="fixed")))


@pytest.mark.parametrize("tag,kw", _PRIOR_CASES)
def test_priors_match_the_reference_source_composition(tag, kw):
    """DSLP / SAAS / fixed-kernel-variance / LogNormal kernel-variance priors: the reference's own _standard_prior_logprob /
    saas_prior_logprob (BOBE/gp.py:56-78,357-366) executed with scipy.stats densities standing in for numpyro's, and the
    gradient of its full neg_mll by autodiff (torch-backed stand-in), vs the restatement AND the product's host-side priors."""
    from bobe_b200 import GP
    v = _ref_vectors()
    X, y, lp = v[tag + "X"], v[tag + "y"], v[tag + "log_params"]
    common = dict(noise=1e-6, kernel="matern", lengthscales=np.array([0.5, 0.8, 1.1]), kernel_variance=1.4)
    ref = O.OracleGP(X, y[:, None], **common, **kw)
    gp = GP(X, y[:, None], **common, **kw)  # host logic only: nothing here touches the device
    assert ref.num_hyperparams == gp.num_hyperparams == int(v[tag + "num_hyperparams"])
    assert np.allclose(ref.hyperparam_bounds, v[tag + "hyperparam_bounds"], rtol=1e-15)
    assert np.allclose(np.asarray(gp.hyperparam_bounds), v[tag + "hyperparam_bounds"], rtol=1e-15)
    for r in range(lp.shape[0]):
        want_prior = float(v[tag + "prior"][r])
        assert abs(ref.log_prior_and_grad(lp[r])[0] - want_prior) < 1e-13 * max(1.0, abs(want_prior))
        assert abs(float(np.sum(gp.prior_func(*gp._parse_hyperparams(lp[r])))) - want_prior) < 1e-13 * max(1.0, abs(want_prior))
        val, grad = ref.neg_mll_and_grad(lp[r])
        assert abs(val - float(v[tag + "neg_mll"][r])) < 1e-12 * abs(val) and abs(val - float(v[tag + "neg_mll_ad"][r])) < 1e-10 * abs(val)
        want = v[tag + "neg_mll_ad_grad"][r]
        assert np.max(np.abs(grad - want)) < 1e-9 * max(1.0, float(np.max(np.abs(want))))
        # the product's prior gradient = the reference's total gradient minus the restatement's data term
        data_grad = grad + ref.log_prior_and_grad(lp[r])[1]  # neg_mll = -(data + prior)
        prior_grad_ref = -(want) - (-data_grad)
        got = gp._prior_grad(*gp._parse_hyperparams(lp[r]))
        assert np.max(np.abs(np.asarray(got) - prior_grad_ref)) < 1e-8 * max(1.0, float(np.max(np.abs(want))))


def test_oracle_matches_the_reference_source_at_the_headline_shape():
    """BASELINE config H (n = 2000, d = 16, Matern-5/2 ARD): what the reference's own source computes for the seeded
    synthetic set -- posterior mean / variance, log-ML at three hyper-parameter rows with the autodiff gradient, fantasy
    variance -- vs the restatement."""
    v = _ref_vectors()
    n, d = int(v["gpH_n"]), int(v["gpH_d"])
    X, y = O.synthetic_training_set(n, d)
    gp = O.OracleGP(X, y, noise=1e-8, kernel="matern", lengthscales=np.ones(d), kernel_variance=1.0)
    assert abs(gp.y_mean - float(v["gpH_y_mean"])) < 1e-13 * abs(gp.y_mean) and abs(gp.y_std - float(v["gpH_y_std"])) < 1e-13 * gp.y_std
    condL = float(v["gpH_cond_L"])
    assert abs(float(np.sum(np.log(np.diag(gp.cholesky)))) - float(v["gpH_logdet_half"])) < 1e-14 * condL * n
    assert _rel(gp.alphas, v["gpH_alphas"]) < 1e-15 * condL ** 2
    Xq = O.synthetic_queries(48, d, seed=21)
    assert mixed_err(gp.predict_mean_batched(Xq), v["gpH_mean_batched"], gp.y_std) < 1e-11
    assert mixed_err(gp.predict_var_batched(Xq), v["gpH_var_batched"], gp.y_std ** 2) < 1e-11
    ms, vs = gp.predict_batched(Xq)
    assert mixed_err(np.ravel(ms), v["gpH_std_mean_batched"], 1.0) < 1e-11 and mixed_err(np.ravel(vs), v["gpH_std_var_batched"], 1.0) < 1e-11
    mc, cand = O.synthetic_queries(64, d, seed=22), O.synthetic_queries(2, d, seed=23)
    assert mixed_err(gp.fantasy_var_shared(cand, mc), v["gpH_fantasy_var"], gp.y_std ** 2) < 1e-10
    val, grad = gp.neg_mll_and_grad(v["gpH_log_params"][0])
    assert abs(val - float(v["gpH_neg_mll"][0])) < 1e-11 * max(abs(val), n)
    assert np.max(np.abs(grad - v["gpH_neg_mll_ad_grad"][0])) < 1e-8 * max(1.0, float(np.max(np.abs(grad))))


@pytest.mark.parametrize("tag", ["gpD_", "gpE_"])
def test_oracle_matches_the_reference_source_at_configs_d_and_e(tag):
    """BASELINE configs D (n = 1500, d = 27) and E (n = 4000, d = 12), RBF: the restatement vs the reference's own source."""
    v = _ref_vectors()
    n, d, ell = int(v[tag + "n"]), int(v[tag + "d"]), float(v[tag + "ell"])
    X, y = O.synthetic_training_set(n, d)
    gp = O.OracleGP(X, y, noise=1e-8, kernel="rbf", lengthscales=np.full(d, ell), kernel_variance=1.0)
    Xq = O.synthetic_queries(32, d, seed=31)
    mc, cand = O.synthetic_queries(48, d, seed=32), O.synthetic_queries(2, d, seed=33)
    assert abs(float(np.sum(np.log(np.diag(gp.cholesky)))) - float(v[tag + "logdet_half"])) < 1e-12 * n
    assert mixed_err(gp.predict_mean_batched(Xq), v[tag + "mean_batched"], gp.y_std) < 1e-11
    assert mixed_err(gp.predict_var_batched(Xq), v[tag + "var_batched"], gp.y_std ** 2) < 1e-11
    fv = gp.fantasy_var_shared(cand, mc)
    assert mixed_err(fv, v[tag + "fantasy_var"], gp.y_std ** 2) < 1e-10
    assert mixed_err(fv.mean(axis=1), v[tag + "wipv"], gp.y_std ** 2) < 1e-10


def test_oracle_matches_the_reference_source_at_config_a():
    """BASELINE config A: the 512-candidate WIPV / WIPStd sweep (BOBE/acquisition.py:385-397) of the reference's own source."""
    v = _ref_vectors()
    X, y = O.synthetic_training_set(100, 2)
    gp = O.OracleGP(X, y, noise=1e-8, kernel="rbf", lengthscales=np.full(2, 0.3), kernel_variance=1.0)
    mc = O.synthetic_queries(512, 2, seed=5)
    assert mixed_err(O.wipv_values(gp, mc, mc), v["gpA_wipv_self"], gp.y_std ** 2) < 1e-9
    assert mixed_err(O.wipv_values(gp, mc, mc, std=True), v["gpA_wipstd_self"], gp.y_std) < 1e-9
    assert mixed_err(gp.predict_mean_batched(mc[:64]), v["gpA_mean_batched"], gp.y_std) < 1e-9
