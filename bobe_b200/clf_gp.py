"""Drop-in for ``BOBE/clf_gp.py`` (SVM flavour): a GP whose predictions are masked by a feasibility classifier.

SURVEY.md 8f row 4.  The reference trains the classifier on ALL evaluated points (labels: within ``clf_threshold`` of
the best value), the GP only on the points within ``gp_threshold``, and wraps every single-point predictor in
``jnp.where(clf_probs >= probability_threshold, value, fill)`` (``BOBE/clf_gp.py:173-205``).  For the SVM classifier
(``BOBE/clf.py:36-78,188-214``) the decision function ``sum_j dual_j exp(-gamma |sv_j - x|^2) + b`` is an isotropic
RBF kernel-row product, so on the device it is ONE fused mean pass of the kernel-matrix kernel plus a select
(``bobe_svm_mask``), applied to the output of ``bobe_predict`` before anything returns to the host.

The classifier is trained on the host with scikit-learn's ``SVC`` exactly as the reference does (training is not on
the hot path).  The ``nn`` / ``ellipsoid`` classifiers of the reference need flax / optax and are out of scope:
asking for them raises ``NotImplementedError``.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import ops
from .gp import GP, safe_noise_floor as SAFE_NOISE_FLOOR, _is_t, _to_dev

log = logging.getLogger("bobe_b200.clf_gp")


def train_svm_classifier(X, Y, settings=None):
    """BOBE/clf.py:36-69 -- sklearn SVC(kernel='rbf', gamma='scale', C=1e7); returns (params, metrics)."""
    from sklearn.svm import SVC
    settings = settings or {}
    gamma, C, kernel = settings.get('gamma', 'scale'), settings.get('C', 1e7), settings.get('kernel', 'rbf')
    if kernel != 'rbf':
        raise NotImplementedError("only the RBF SVM of BOBE/clf.py:188-209 has a device decision function")
    clf = SVC(kernel=kernel, gamma=gamma, C=C)
    clf.fit(np.asarray(X), np.asarray(Y))
    params = {'support_vectors': np.array(clf.support_vectors_, dtype=np.float64),
              'dual_coef': np.array(clf.dual_coef_[0], dtype=np.float64),
              'intercept': float(clf.intercept_[0]),
              'gamma_eff': float(clf._gamma)}
    metrics = {'n_support_vectors': len(params['support_vectors']), 'gamma': f"{params['gamma_eff']:.2e}",
               'C': f"{C:.2e}", 'intercept': f"{params['intercept']:.2e}"}
    return params, metrics


class GPwithClassifier(GP):
    """BOBE/clf_gp.py:16-148 (constructor arguments and attributes) for ``clf_type='svm'``."""

    def __init__(self, train_x=None, train_y=None, clf_type='svm', clf_settings={}, clf_use_size=10, clf_update_step=1,
                 probability_threshold=0.5, minus_inf=-1e5, clf_threshold=250., gp_threshold=500., noise=1e-8,
                 kernel="rbf", optimizer="scipy", optimizer_options={}, kernel_variance_bounds=[1e-4, 1e8],
                 lengthscale_bounds=[0.01, 5.], tausq=None, tausq_bounds=[1e-4, 1e4], kernel_variance_prior=None,
                 lengthscale_prior=None, lengthscales=None, kernel_variance=1.0, param_names=None,
                 train_clf_on_init=True, device=None):
        self.train_x_clf = np.array(train_x, dtype=np.float64)
        self.train_y_clf = np.array(train_y, dtype=np.float64).reshape(-1, 1)
        self.clf_use_size, self.clf_update_step = clf_use_size, clf_update_step
        self.clf_type = clf_type.lower()
        if self.clf_type != 'svm':
            if self.clf_type in ('nn', 'ellipsoid'):
                raise NotImplementedError(f"classifier {self.clf_type!r} (flax/optax) is outside the B200 hot path; use 'svm'")
            raise ValueError(f"Unsupported classifier type: {self.clf_type}")
        self.clf_settings = clf_settings
        self.clf_params, self.clf_metrics = None, {}
        self.probability_threshold, self.minus_inf = probability_threshold, minus_inf
        self.clf_threshold, self.gp_threshold = clf_threshold, gp_threshold
        self._sv_dev = self._dual_dev = None
        if self.train_y_clf.size > 0:  # BOBE/clf_gp.py:84-90: the GP only sees points near the best value
            mask_gp = self.train_y_clf.flatten() > (self.train_y_clf.max() - self.gp_threshold)
            train_x_gp, train_y_gp = self.train_x_clf[mask_gp], self.train_y_clf[mask_gp]
        else:
            train_x_gp, train_y_gp = self.train_x_clf, self.train_y_clf
        super().__init__(train_x=train_x_gp, train_y=train_y_gp, noise=noise, kernel=kernel, optimizer=optimizer,
                         optimizer_options=optimizer_options, kernel_variance_bounds=kernel_variance_bounds,
                         lengthscale_bounds=lengthscale_bounds, lengthscales=lengthscales,
                         kernel_variance=kernel_variance,
                         lengthscale_prior=lengthscale_prior if lengthscale_prior is not None else "DSLP",
                         kernel_variance_prior=kernel_variance_prior, tausq=tausq, tausq_bounds=tausq_bounds,
                         param_names=param_names, device=device)
        self.use_clf = self.clf_data_size >= self.clf_use_size
        if self.use_clf and train_clf_on_init:
            self.train_classifier()

    @property
    def clf_data_size(self):
        return self.train_x_clf.shape[0]

    # ---- classifier ------------------------------------------------------------------------------------------
    def train_classifier(self):
        """BOBE/clf_gp.py:126-170."""
        if not self.use_clf and self.clf_data_size >= self.clf_use_size:
            self.use_clf = True
        if not self.use_clf:
            return
        labels = np.where(self.train_y_clf.flatten() < self.train_y_clf.max() - self.clf_threshold, 0, 1)
        if np.all(labels == labels[0]):
            log.debug("All labels are identical. Not using classifier for the moment")
            self.use_clf = False
            return
        self.clf_params, self.clf_metrics = train_svm_classifier(self.train_x_clf, labels, self.clf_settings)
        self._sv_dev = self._dual_dev = None

    def _clf_active(self):
        return bool(self.use_clf and self.clf_params is not None)

    def _clf_dev(self):
        if self._sv_dev is None:
            dev = self.device
            self._sv_dev = _to_dev(self.clf_params['support_vectors'], dev)
            self._dual_dev = _to_dev(self.clf_params['dual_coef'], dev)
        return self._sv_dev, self._dual_dev

    def clf_decision(self, x):
        """RBF-SVM decision values at (M, d) points (BOBE/clf.py:188-209), computed on the device."""
        sv, dual = self._clf_dev()
        xq = _to_dev(np.atleast_2d(x) if not _is_t(x) else x, self.device)
        dec = ops.svm_mask(sv, dual, self.clf_params['intercept'], self.clf_params['gamma_eff'], xq, want_decision=True)
        return dec if _is_t(x) else dec.cpu().numpy()

    # ---- prediction: the mask is an epilogue of the device call ---------------------------------------------------
    def _predict_dev(self, xq, want_mean, want_var, standardised):
        mean, var = super()._predict_dev(xq, want_mean, want_var, standardised)
        if self._clf_active():
            if not (0.0 < self.probability_threshold <= 1.0):
                raise ValueError("probability_threshold must be in (0, 1] for the 0/1 SVM probabilities")
            sv, dual = self._clf_dev()
            # fills: minus_inf for the mean, safe_noise_floor for the variance (BOBE/clf_gp.py:179,187,202-203)
            ops.svm_mask(sv, dual, self.clf_params['intercept'], self.clf_params['gamma_eff'], xq, mean, var,
                         self.minus_inf, SAFE_NOISE_FLOOR)
        return mean, var

    def predict_grad_batched(self, x, standardised=False, want_mean=True, want_var=True):
        """``jax.value_and_grad`` of the MASKED ``predict_*_single`` (BOBE/clf_gp.py:173-205): where the classifier
        excludes the point the value is the fill (``minus_inf`` / ``safe_noise_floor``) and -- ``jnp.where`` passing no
        gradient through a constant branch -- its input gradient is zero."""
        out = super().predict_grad_batched(x, standardised, want_mean, want_var)
        if not self._clf_active():
            return out
        if not (0.0 < self.probability_threshold <= 1.0):
            raise ValueError("probability_threshold must be in (0, 1] for the 0/1 SVM probabilities")
        as_np = not _is_t(x)
        dec = self.clf_decision(x if _is_t(x) else np.atleast_2d(np.asarray(x, dtype=np.float64)))
        bad = (dec < 0.0) if as_np else (dec.to(out[0].device if out[0] is not None else out[1].device) < 0.0)
        mean, var, dmean, dvar = out
        fix = (lambda a, fill: np.where(bad.reshape((-1,) + (1,) * (a.ndim - 1)), fill, a)) if as_np else \
              (lambda a, fill: torch.where(bad.reshape((-1,) + (1,) * (a.dim() - 1)), torch.as_tensor(fill, dtype=a.dtype, device=a.device), a))
        if mean is not None:
            mean, dmean = fix(mean, self.minus_inf), fix(dmean, 0.0)
        if var is not None:
            var, dvar = fix(var, SAFE_NOISE_FLOOR), fix(dvar, 0.0)
        return mean, var, dmean, dvar

    # ---- data ------------------------------------------------------------------------------------------------------
    def update(self, new_x, new_y):
        """BOBE/clf_gp.py:214-246 -- append to the classifier set, re-select the GP set, re-factorise."""
        new_x = np.atleast_2d(np.asarray(new_x, dtype=np.float64))
        new_y = np.atleast_2d(np.asarray(new_y, dtype=np.float64))
        add_x, add_y = [], []
        for i in range(new_x.shape[0]):
            if np.any(np.all(np.isclose(self.train_x_clf, new_x[i], atol=1e-6, rtol=1e-4), axis=1)):
                log.debug(f"Point {new_x[i]} already exists in the training set, not updating")
            else:
                add_x.append(new_x[i])
                add_y.append(new_y[i])
        if not add_x:
            return
        self.train_x_clf = np.concatenate([self.train_x_clf, np.atleast_2d(np.array(add_x))], axis=0)
        self.train_y_clf = np.concatenate([self.train_y_clf, np.atleast_2d(np.array(add_y)).reshape(-1, 1)], axis=0)
        mask_gp = self.train_y_clf.flatten() > (self.train_y_clf.max() - self.gp_threshold)
        self.train_x = self.train_x_clf[mask_gp]
        ty = self.train_y_clf[mask_gp].reshape(-1, 1)
        self.y_std = float(np.std(ty)) if ty.shape[0] > 1 else 1.0
        self.y_mean = float(np.mean(ty))
        self.train_y = (ty - self.y_mean) / self.y_std
        self.recompute_cholesky()  # the GP subset can shrink as well as grow: no rank-b shortcut here

    def get_random_point(self, rng=None, nstd=None):
        """BOBE/clf_gp.py:254-275."""
        rng = rng if rng is not None else np.random.default_rng()
        if not self.use_clf:
            return super().get_random_point(rng=rng, nstd=nstd)
        if nstd is not None:
            from scipy.special import erfc  # BOBE/utils/core.py:150-167 get_threshold_for_nsigma
            from scipy.stats import chi2
            threshold = 0.5 * chi2.isf(erfc(nstd / np.sqrt(2)), self.ndim)
        else:
            threshold = self.clf_threshold
        idx = np.where(self.train_y_clf.flatten() > self.train_y_clf.max() - threshold)[0]
        return self.train_x_clf[rng.choice(idx, size=1)[0]]

    # ---- state -----------------------------------------------------------------------------------------------------
    def state_dict(self):
        """BOBE/clf_gp.py:277-312 -- the GP's keys plus the classifier's."""
        state = super().state_dict()
        state.update({'train_x_clf': np.array(self.train_x_clf), 'train_y_clf': np.array(self.train_y_clf),
                      'clf_type': self.clf_type, 'clf_settings': self.clf_settings, 'clf_use_size': self.clf_use_size,
                      'clf_update_step': self.clf_update_step, 'probability_threshold': self.probability_threshold,
                      'minus_inf': self.minus_inf, 'clf_threshold': self.clf_threshold,
                      'gp_threshold': self.gp_threshold, 'use_clf': self.use_clf, 'clf_params': self.clf_params,
                      'clf_metrics': self.clf_metrics, 'gp_class': 'GPwithClassifier'})
        return state

    @classmethod
    def from_state_dict(cls, state):
        """BOBE/clf_gp.py:314-390 -- rebuilt without retraining; the stored classifier parameters are reused."""
        def _plain(v):
            return v.item() if isinstance(v, np.ndarray) and v.dtype == object and v.shape == () else v
        gp = cls(train_x=state['train_x_clf'], train_y=state['train_y_clf'], clf_type=_plain(state['clf_type']),
                 clf_settings=_plain(state['clf_settings']), clf_use_size=int(state['clf_use_size']),
                 clf_update_step=int(state['clf_update_step']),
                 probability_threshold=float(state['probability_threshold']), minus_inf=float(state['minus_inf']),
                 clf_threshold=float(state['clf_threshold']), gp_threshold=float(state['gp_threshold']),
                 noise=float(state['noise']), kernel=_plain(state['kernel_name']),
                 optimizer=_plain(state['optimizer_method']), optimizer_options=_plain(state['optimizer_options']),
                 kernel_variance_bounds=list(np.ravel(_plain(state['kernel_variance_bounds']))),
                 lengthscale_bounds=list(np.ravel(_plain(state['lengthscale_bounds']))),
                 tausq=state.get('tausq', 1.0), tausq_bounds=list(np.ravel(_plain(state.get('tausq_bounds', [1e-4, 1e4])))),
                 kernel_variance_prior=_plain(state.get('kernel_variance_prior_spec')),
                 lengthscale_prior=_plain(state.get('lengthscale_prior_spec')), lengthscales=state['lengthscales'],
                 kernel_variance=float(state['kernel_variance']), train_clf_on_init=False)
        gp.use_clf = bool(state['use_clf'])
        gp.clf_params = _plain(state['clf_params'])
        gp.clf_metrics = _plain(state['clf_metrics'])
        return gp
