"""NumPy fp64 restatement of the Gaussian-process posterior the reference queries.

TEST INFRASTRUCTURE ONLY.  The arithmetic lives in the reference's third-party
dependency scikit-learn (requirements.txt:138 pins 1.5.2; 1.9.0 is installed
here).  Call sites: sa_nsga_local.py:212-223 (SurrogateManager.predict with
return_std), sa_nsga_penalty.py:342-363 (mean only), mobo_penalty.py:252-287.
Restated from sklearn/gaussian_process/_gpr.py:446-499 (predict) and
kernels.py:1713-1743 (Matern); pinned by tests against the installed sklearn.
"""
from __future__ import annotations

import numpy as np


def matern(xa: np.ndarray, xb: np.ndarray, length_scale: float, nu: float) -> np.ndarray:
    """Matern(nu in {0.5, 1.5, 2.5}) on d = ||(x - y)/l||_2  (kernels.py:1713-1743)."""
    a = np.asarray(xa, np.float64) / length_scale
    b = np.asarray(xb, np.float64) / length_scale
    diff = a[:, None, :] - b[None, :, :]
    d = np.sqrt((diff * diff).sum(axis=2))
    if nu == 0.5:
        return np.exp(-d)
    if nu == 1.5:
        t = d * np.sqrt(3.0)
        return (1.0 + t) * np.exp(-t)
    if nu == 2.5:
        t = d * np.sqrt(5.0)
        return (1.0 + t + t * t / 3.0) * np.exp(-t)
    raise ValueError("nu must be 0.5, 1.5 or 2.5")


def posterior(x_query, x_train, alpha, chol_lower, *, amplitude: float, length_scale: float,
              nu: float, noise: float = 0.0, y_scale: float = 1.0, y_shift: float = 0.0):
    """mean = K* @ alpha; V = L^-1 K*^T; var = diag(K**) - sum(V^2) clipped at 0.

    kernel = amplitude * Matern(length_scale, nu) [+ WhiteKernel(noise)]; the white
    term only enters diag(K**) (kernels.py WhiteKernel.__call__ with Y given = 0).
    The result is mapped back with mean*y_scale + y_shift and std*y_scale, which
    covers both sklearn's normalize_y (_gpr.py:449-452, 494-497) and the reference's
    external StandardScaler un-scaling (sa_nsga_local.py:221-222).
    """
    xq = np.atleast_2d(np.asarray(x_query, np.float64))
    k_star = amplitude * matern(xq, x_train, length_scale, nu)
    mean = k_star @ np.asarray(alpha, np.float64).ravel()
    lower = np.asarray(chol_lower, np.float64)
    n = lower.shape[0]
    v = np.zeros((n, xq.shape[0]))
    rhs = k_star.T.copy()
    for i in range(n):                                   # forward substitution
        v[i] = (rhs[i] - lower[i, :i] @ v[:i]) / lower[i, i]
    var = (amplitude + noise) - (v * v).sum(axis=0)
    var = np.where(var < 0.0, 0.0, var)
    return mean * y_scale + y_shift, np.sqrt(var) * y_scale


def unpack_sklearn(gpr) -> dict:
    """Read (amplitude, length_scale, nu, noise) and (X_train_, alpha_, L_) from a fitted
    sklearn GaussianProcessRegressor whose kernel_ is Matern, C*Matern or
    C*Matern + WhiteKernel."""
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, Product, Sum, WhiteKernel

    k = gpr.kernel_
    amp, noise = 1.0, 0.0
    if isinstance(k, Sum):
        parts = [k.k1, k.k2]
        white = [p for p in parts if isinstance(p, WhiteKernel)]
        rest = [p for p in parts if not isinstance(p, WhiteKernel)]
        noise = float(white[0].noise_level)
        k = rest[0]
    if isinstance(k, Product):
        parts = [k.k1, k.k2]
        const = [p for p in parts if isinstance(p, ConstantKernel)]
        rest = [p for p in parts if not isinstance(p, ConstantKernel)]
        amp = float(const[0].constant_value)
        k = rest[0]
    if not isinstance(k, Matern):
        raise TypeError(f"unsupported kernel {gpr.kernel_}")
    return dict(amplitude=amp, length_scale=float(np.ravel(k.length_scale)[0]), nu=float(k.nu),
                noise=noise, x_train=np.asarray(gpr.X_train_, np.float64),
                alpha=np.asarray(gpr.alpha_, np.float64).ravel(), chol_lower=np.asarray(gpr.L_, np.float64),
                y_scale=float(np.ravel(getattr(gpr, "_y_train_std", 1.0))[0]),
                y_shift=float(np.ravel(getattr(gpr, "_y_train_mean", 0.0))[0]))


# ---- hyper-parameter objective (the device form is csrc/gp_lml.cu) ---------------------------------------------------
def matern_dlogl(xa: np.ndarray, length_scale: float, nu: float):
    """(k, dk/dlog l) of Matern(nu) on xa x xa (kernels.py Matern.__call__ with eval_gradient=True)."""
    a = np.asarray(xa, np.float64) / length_scale
    diff = a[:, None, :] - a[None, :, :]
    r = np.sqrt((diff * diff).sum(axis=2))
    if nu == 0.5:
        e = np.exp(-r)
        return e, r * e
    if nu == 1.5:
        s = np.sqrt(3.0) * r
        e = np.exp(-s)
        return (1.0 + s) * e, s * s * e
    if nu == 2.5:
        s = np.sqrt(5.0) * r
        e = np.exp(-s)
        return (1.0 + s + s * s / 3.0) * e, s * s / 3.0 * (1.0 + s) * e
    raise ValueError("nu must be 0.5, 1.5 or 2.5")


def blocked_sweep(k_matrix: np.ndarray, block: int = 16):
    """(K^-1, log det K) by the symmetric sweep operator in the blocked order gp_lml.cu uses: the block's columns are
    swept one pivot after the other (each pivot column kept as it was when used), the rest of the matrix then takes the
    block's rank-`block` update at once.  Returns (None, -inf) on a non-positive pivot (scikit-learn: failed Cholesky)."""
    a = np.array(k_matrix, np.float64)
    n = len(a)
    logdet = 0.0
    for k0 in range(0, n, block):
        cols = np.arange(k0, min(n, k0 + block))
        s = a[:, cols].copy()
        u = np.zeros_like(s)
        dinv = np.zeros(len(cols))
        for k, r in enumerate(cols):
            col = s[:, k].copy()
            d = col[r]
            if not d > 0.0 or not np.isfinite(d):
                return None, -np.inf
            logdet += np.log(d)
            u[:, k], dinv[k] = col, 1.0 / d
            piv = col[cols].copy()
            s -= np.outer(col, piv) / d
            s[r, :] = piv / d
            s[:, k] = col / d
            s[r, k] = -1.0 / d
        outside = np.ones(n, bool)
        outside[cols] = False
        upd = (u * dinv) @ u.T
        a[np.ix_(outside, outside)] -= upd[np.ix_(outside, outside)]
        a[:, cols] = s
        a[cols, :] = s.T
    return -a, logdet


def log_marginal_likelihood(theta, x, y, kind: int, nu: float, jitter: float = 1e-10, block: int = 16):
    """(lml, d lml / d theta) as GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)
    (sklearn/gaussian_process/_gpr.py) for kind 0 = C * Matern + WhiteKernel, theta = (log c, log l, log noise), and
    kind 1 = Matern, theta = (log l), computed the way the device kernel does (blocked sweep instead of Cholesky)."""
    theta = np.asarray(theta, np.float64)
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    n = len(x)
    amp, ell, noise = (np.exp(theta[0]), np.exp(theta[1]), np.exp(theta[2])) if kind == 0 else (1.0, np.exp(theta[0]), 0.0)
    k, dk = matern_dlogl(x, ell, nu)
    kinv, logdet = blocked_sweep(amp * k + (noise + jitter) * np.eye(n), block)
    if kinv is None:
        return -np.inf, np.zeros_like(theta)
    alpha = kinv @ y
    inner = np.outer(alpha, alpha) - kinv
    lml = -0.5 * y @ alpha - 0.5 * logdet - 0.5 * n * np.log(2.0 * np.pi)
    if kind == 0:
        grad = 0.5 * np.array([(inner * amp * k).sum(), (inner * amp * dk).sum(), noise * np.trace(inner)])
    else:
        grad = 0.5 * np.array([(inner * dk).sum()])
    return float(lml), grad
