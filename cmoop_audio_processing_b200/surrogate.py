"""Kriging / GP surrogate drop-ins with the posterior evaluated by the CUDA kernel.

Mirrors, name for name:
* ``SurrogateManager``        ablation_study/sa_nsga_local.py:169-234 (update / predict(return_std) /
                              predict_and_structure) and the mean-only form sa_nsga_penalty.py:258-363
* ``select_infill_points``    sa_nsga_penalty.py:472-518
* ``perturb_hparams`` / ``lcb_dominates`` / ``perform_local_search``   sa_nsga_local.py:351-433
* ``train_gps`` / ``predict_gps`` / ``penalized_acquisition`` / ``hparams_to_vector`` /
  ``vector_to_hparams``       mobo_penalty.py:252-338

Hyper-parameter fitting (L-BFGS-B on the log marginal likelihood with random restarts)
stays scikit-learn's arithmetic -- it is the reference's third-party dependency behind the
same API and its restarts are unseeded -- but the independent optimiser starts of the four
models run concurrently in worker processes (gp_fit.py; same objective, same RNG order,
same fitted model).  Everything that is *queried* (K*.alpha, the
triangular solve and the variance) runs on the GPU through ``cmoop_gp_predict_*``.
Because the genotype space has only 288 points, ``SurrogateManager`` evaluates the
whole space in ONE launch after every update and serves ``predict`` (including the
5 x |elite| single-point calls of the local search) from that table.
"""
from __future__ import annotations

import ctypes as C
import itertools
import random
from copy import deepcopy

import numpy as np

from . import _lib
from .nsga import EPSILON, HPARAM_SPACE

TARGET_KEYS = ("neg_acc", "size", "fpr", "cv")
NUMERICAL = ["filters", "kernel_size", "residual_blocks", "fc_layers"]
CATEGORICAL = ["use_bn", "use_dropout"]


# --------------------------------------------------------------------- device GP group
def _kernel_params(gpr) -> dict:
    """(amplitude, length_scale, nu, noise) of Matern | C*Matern | C*Matern + White."""
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, Product, Sum, WhiteKernel

    k = gpr.kernel_
    amp, noise = 1.0, 0.0
    if isinstance(k, Sum):
        a, b = k.k1, k.k2
        if isinstance(b, WhiteKernel):
            noise, k = float(b.noise_level), a
        elif isinstance(a, WhiteKernel):
            noise, k = float(a.noise_level), b
    if isinstance(k, Product):
        a, b = k.k1, k.k2
        if isinstance(a, ConstantKernel):
            amp, k = float(a.constant_value), b
        elif isinstance(b, ConstantKernel):
            amp, k = float(b.constant_value), a
    if not isinstance(k, Matern):
        raise TypeError(f"unsupported GP kernel for the CUDA posterior: {gpr.kernel_}")
    return dict(amplitude=amp, length_scale=float(np.ravel(k.length_scale)[0]), nu=float(k.nu), noise=noise)


class DeviceGPGroup:
    """A set of fitted GPs (same input dimension) resident on the GPU."""

    def __init__(self, specs: list[dict]):
        """specs: dicts with x_train [n,d], alpha [n], chol_lower [n,n] or None, amplitude,
        length_scale, nu, noise, y_scale, y_shift."""
        self._lib = _lib.load()
        _lib.bind_device()
        self.n_models = len(specs)
        self.dim = int(np.asarray(specs[0]["x_train"]).shape[1])
        self._keep = []
        arr = (_lib.GpModel * self.n_models)()
        for i, s in enumerate(specs):
            x = np.ascontiguousarray(s["x_train"], np.float64)
            a = np.ascontiguousarray(np.ravel(s["alpha"]), np.float64)
            low = None if s.get("chol_lower") is None else np.ascontiguousarray(s["chol_lower"], np.float64)
            self._keep += [x, a, low]
            arr[i].n_train, arr[i].dim = x.shape[0], x.shape[1]
            arr[i].amplitude, arr[i].length_scale = float(s["amplitude"]), float(s["length_scale"])
            arr[i].nu, arr[i].noise = float(s["nu"]), float(s.get("noise", 0.0))
            arr[i].y_scale, arr[i].y_shift = float(s.get("y_scale", 1.0)), float(s.get("y_shift", 0.0))
            arr[i].x_train = x.ctypes.data_as(_lib.c_double_p)
            arr[i].alpha = a.ctypes.data_as(_lib.c_double_p)
            arr[i].chol_lower = low.ctypes.data_as(_lib.c_double_p) if low is not None else None
        handle = C.c_void_p()
        _lib.check(self._lib.cmoop_gp_create(arr, self.n_models, C.byref(handle)), "cmoop_gp_create")
        self._handle = handle

    @classmethod
    def from_sklearn(cls, gprs, y_affine=None):
        """y_affine: optional list of (scale, shift) applied on top of sklearn's own
        normalize_y statistics (used for the reference's external StandardScaler)."""
        specs = []
        for i, g in enumerate(gprs):
            s = _kernel_params(g)
            s["x_train"], s["alpha"], s["chol_lower"] = g.X_train_, g.alpha_, g.L_
            scale = float(np.ravel(getattr(g, "_y_train_std", 1.0))[0])
            shift = float(np.ravel(getattr(g, "_y_train_mean", 0.0))[0])
            if y_affine is not None:
                o_scale, o_shift = y_affine[i]
                scale, shift = scale * o_scale, shift * o_scale + o_shift
            s["y_scale"], s["y_shift"] = scale, shift
            specs.append(s)
        return cls(specs)

    def predict(self, xq, return_std=True):
        """xq [q,dim] -> (mean [n_models,q], std [n_models,q] | None)."""
        xq = np.ascontiguousarray(np.atleast_2d(xq), np.float64)
        if xq.shape[1] != self.dim:
            raise ValueError(f"query has {xq.shape[1]} features, models were fitted on {self.dim}")
        q = xq.shape[0]
        mean = np.empty((self.n_models, q), np.float64)
        std = np.empty((self.n_models, q), np.float64) if return_std else None
        _lib.check(self._lib.cmoop_gp_predict_host(self._handle, _lib.ptr(xq), q, _lib.ptr(mean), _lib.ptr(std)),
                   "cmoop_gp_predict_host")
        return mean, std

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.cmoop_gp_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------- SurrogateManager
def _genotype_key(hp) -> tuple:
    return tuple(hp[k] for k in NUMERICAL + CATEGORICAL)


class SurrogateManager:
    """Four GPs (neg_acc, size, fpr, cv) on [filters, kernel_size, residual_blocks, fc_layers,
    onehot(use_bn), onehot(use_dropout)]; targets z-scored per model; training rows
    de-duplicated on the genotype keeping the last evaluation."""

    def __init__(self, n_restarts_optimizer: int = 10, use_table: bool = True, fit_backend: str | None = None):
        """fit_backend: None (CMOOP_GP_FIT_BACKEND, default "host": scikit-learn's objective on the host cores) or
        "device" (same optimiser and starts, objective evaluated by csrc/gp_lml.cu); see gp_fit.py."""
        import pandas as pd

        self.fit_backend = fit_backend
        self.is_fitted = False
        self.categorical_features = list(CATEGORICAL)
        self.numerical_features = list(NUMERICAL)
        self.n_restarts_optimizer = n_restarts_optimizer
        self.training_data = pd.DataFrame()
        self.models = {}
        self.scalers = {}
        self.scaler_mean = {}
        self.scaler_var = {}
        self.scaler_scale = {}
        self._categories = {}
        self._group = None
        self._use_table = use_table
        self._table = None

    # -- feature encoding (ColumnTransformer: passthrough numerics + OneHotEncoder re-fit per update)
    def _encode(self, hparams_list) -> np.ndarray:
        cols = [[float(hp[k]) for hp in hparams_list] for k in self.numerical_features]
        for k in self.categorical_features:
            for cat in self._categories[k]:
                cols.append([1.0 if hp[k] == cat else 0.0 for hp in hparams_list])   # unknown -> all zeros
        return np.ascontiguousarray(np.array(cols, dtype=np.float64).T.reshape(len(hparams_list), -1))

    def update(self, hparams_list, results_list):
        import pandas as pd
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

        new = pd.DataFrame(hparams_list)
        new["y_neg_acc"] = [r["objs"][0] for r in results_list]
        new["y_size"] = [r["objs"][1] for r in results_list]
        new["y_fpr"] = [r["objs"][2] for r in results_list]
        new["y_cv"] = [r["CV"] for r in results_list]
        genes = self.numerical_features + self.categorical_features
        self.training_data = pd.concat([self.training_data, new]).drop_duplicates(
            subset=genes, keep="last").reset_index(drop=True)
        rows = self.training_data[genes].to_dict("records")
        self._categories = {k: sorted(set(bool(r[k]) for r in rows)) for k in self.categorical_features}
        x = self._encode(rows)
        y_affine = []
        from sklearn.preprocessing import StandardScaler

        kernels, ys = [], []
        for key in TARGET_KEYS:
            y = self.training_data[f"y_{key}"].to_numpy(dtype=np.float64).reshape(-1, 1)
            scaler = StandardScaler()                       # same statistics as sa_nsga_local.py:207
            y_scaled = scaler.fit_transform(y)
            self.scalers[key] = scaler
            mean, var, scale = float(scaler.mean_[0]), float(scaler.var_[0]), float(scaler.scale_[0])
            self.scaler_mean[key], self.scaler_var[key], self.scaler_scale[key] = mean, var, scale
            kernels.append(ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1))
            ys.append(y_scaled)
            y_affine.append((scale, mean))
        # the reference fits the four regressors one after the other (sa_nsga_local.py:180-181,209); the optimiser starts
        # are independent, so they run concurrently here (fit_gprs_parallel) with unchanged arithmetic and RNG order
        gprs = fit_gprs_parallel(kernels, x, ys, n_restarts_optimizer=self.n_restarts_optimizer,
                                 backend=getattr(self, "fit_backend", None))
        for key, gpr in zip(TARGET_KEYS, gprs):
            self.models[key] = gpr
        self._install(gprs, y_affine)

    def _install(self, gprs, y_affine):
        """Upload fitted models; std follows the reference rule std*sqrt(var_) if var_>0 else 0."""
        if self._group is not None:
            self._group.close()
        self._group = DeviceGPGroup.from_sklearn(gprs, y_affine)
        self._zero_std = [j for j, k in enumerate(TARGET_KEYS) if not self.scaler_var[k] > 0]
        self._table = None
        self.is_fitted = True

    def _query(self, hparams_list):
        mean, std = self._group.predict(self._encode(hparams_list), return_std=True)
        # the kernel scaled std by scaler.scale_ (= sqrt(var_)); the reference returns zeros when var_ == 0
        for j in self._zero_std:
            std[j] = 0.0
        return mean, std

    def _build_table(self):
        keys = NUMERICAL + CATEGORICAL
        space = [dict(zip(keys, vals)) for vals in itertools.product(*[HPARAM_SPACE[k] for k in keys])]
        mean, std = self._query(space)
        self._table = {_genotype_key(hp): (mean[:, i].copy(), std[:, i].copy()) for i, hp in enumerate(space)}

    def predict(self, hparams_list, return_std=False):
        if not self.is_fitted:
            raise RuntimeError("Surrogate models must be fitted.")
        q = len(hparams_list)
        mean = np.empty((4, q))
        std = np.empty((4, q))
        served = False
        if self._use_table:
            if self._table is None:
                self._build_table()
            try:
                for i, hp in enumerate(hparams_list):
                    mean[:, i], std[:, i] = self._table[_genotype_key(hp)]
                served = True
            except KeyError:
                served = False
        if not served:
            mean, std = self._query(hparams_list)
        preds = {k: mean[j].copy() for j, k in enumerate(TARGET_KEYS)}
        if not return_std:
            return preds
        return preds, {k: std[j].copy() for j, k in enumerate(TARGET_KEYS)}

    def predict_and_structure(self, hparams_list):
        preds, _ = self.predict(hparams_list, return_std=True)
        return [{"hparams": hp, "objs": [preds["neg_acc"][i], preds["size"][i], preds["fpr"][i]],
                 "CV": max(0, preds["cv"][i])} for i, hp in enumerate(hparams_list)]

    def predict_structured_mean(self, hparams_list):
        """sa_nsga_penalty.py:342-363: ``predict`` there returns structured records directly."""
        return self.predict_and_structure(hparams_list)


class SurrogateManagerPenalty(SurrogateManager):
    """The mean-only manager of the top-level script, sa_nsga_penalty.py:258-363: same four GPs, same encoding, same
    de-duplicated training table, but ``predict(hparams_list)`` takes no ``return_std`` and returns the structured
    records ``[{'hparams','objs','CV'}]`` that script hands straight to ``select_infill_points`` (call site :563).
    Bind THIS class as ``SurrogateManager`` when rebinding sa_nsga_penalty.py;
    ``predict_and_structure`` / the dict-returning form stay reachable as ``predict_arrays``."""

    def predict_arrays(self, hparams_list, return_std=False):
        return SurrogateManager.predict(self, hparams_list, return_std=return_std)

    def predict(self, hparams_list):
        if not self.is_fitted:
            raise RuntimeError("Surrogate models must be fitted before prediction.")
        preds = SurrogateManager.predict(self, hparams_list, return_std=False)
        return [{"hparams": hp, "objs": [preds["neg_acc"][i], preds["size"][i], preds["fpr"][i]],
                 "CV": max(0, preds["cv"][i])} for i, hp in enumerate(hparams_list)]

    def predict_and_structure(self, hparams_list):
        return self.predict(hparams_list)


# --------------------------------------------------------------------- infill + local search
def select_infill_points(predicted_offspring_data, num_to_select):
    """Feasible (CV < EPSILON) first by summed min-max-normalised objectives, then infeasible by CV."""
    feasible = [(i, r) for i, r in enumerate(predicted_offspring_data) if r["CV"] < EPSILON]
    infeasible = [(i, r) for i, r in enumerate(predicted_offspring_data) if not r["CV"] < EPSILON]
    ranked: list[int] = []
    if feasible:
        objs = np.array([r["objs"] for _, r in feasible])
        lo = objs.min(axis=0)
        span = objs.max(axis=0) - lo
        span[span < EPSILON] = 1.0
        scores = ((objs - lo) / span).sum(axis=1)
        ranked += [i for i, _ in sorted(zip([i for i, _ in feasible], scores), key=lambda t: t[1])]
    if infeasible:
        ranked += [i for i, _ in sorted(infeasible, key=lambda t: t[1]["CV"])]
    chosen = ranked[:num_to_select]
    return chosen, [predicted_offspring_data[i]["hparams"] for i in chosen]


def perturb_hparams(hparams):
    """Change exactly one gene (sa_nsga_local.py:351-363); booleans flip, others move to a different value."""
    new = deepcopy(hparams)
    gene = random.choice(list(HPARAM_SPACE.keys()))
    if isinstance(HPARAM_SPACE[gene][0], bool):
        new[gene] = not new[gene]
    else:
        others = [v for v in HPARAM_SPACE[gene] if v != new[gene]]
        if others:
            new[gene] = random.choice(others)
    return new


def lcb_dominates(sol_a, sol_b):
    a, b = sol_a["lcb_objs"], sol_b["lcb_objs"]
    return all(x <= y for x, y in zip(a, b)) and any(x < y for x, y in zip(a, b))


def perform_local_search(offspring_data, surrogate_manager, k_lcb=1.0):
    """Lamarckian LCB local search (sa_nsga_local.py:370-433): LCB = mu - k*sigma on the three
    objectives, elite = LCB-non-dominated offspring, 5 sweeps of single-gene perturbations
    accepted iff the neighbour LCB-dominates the incumbent."""
    for sol in offspring_data:
        sol["lcb_objs"] = (np.array(sol["objs"]) - k_lcb * np.array(sol["stds"])).tolist()
    elite = [i for i in range(len(offspring_data))
             if not any(j != i and lcb_dominates(offspring_data[j], offspring_data[i])
                        for j in range(len(offspring_data)))]
    for _ in range(5):
        for idx in elite:
            cur = offspring_data[idx]
            cand = perturb_hparams(cur["hparams"])
            mu, sd = surrogate_manager.predict([cand], return_std=True)
            lcb = {k: mu[k][0] - k_lcb * sd[k][0] for k in mu}
            cand_sol = {"lcb_objs": [lcb["neg_acc"], lcb["size"], lcb["fpr"]]}
            if lcb_dominates(cand_sol, cur):
                cur["hparams"] = cand
                cur["lcb_objs"] = cand_sol["lcb_objs"]
                cur["objs"] = [mu["neg_acc"][0], mu["size"][0], mu["fpr"][0]]
                cur["stds"] = [sd["neg_acc"][0], sd["size"][0], sd["fpr"][0]]
    return [sol["hparams"] for sol in offspring_data]


# --------------------------------------------------------------------- MOBO helpers
class _GroupMember:
    """One fitted GP of a DeviceGPGroup; what train_gps returns in place of a sklearn GPR."""

    def __init__(self, group: DeviceGPGroup, index: int, sk_model):
        self.group, self.index, self.sk_model = group, index, sk_model

    def predict(self, x, return_std=False):
        mean, std = self.group.predict(x, return_std=return_std)
        return (mean[self.index], std[self.index]) if return_std else mean[self.index]


from .gp_fit import fit_gprs_parallel  # noqa: E402  (re-exported: the concurrent multi-start GP fit)


def train_gps(X, Y, backend=None):
    """One GaussianProcessRegressor(Matern(nu=2.5), normalize_y=True) per column of Y
    (mobo_penalty.py:252-263); fitted by scikit-learn (backend: see gp_fit.py), uploaded as one device group."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Matern

    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    fitted = fit_gprs_parallel([Matern(nu=2.5) for _ in range(Y.shape[1])], X, [Y[:, dim] for dim in range(Y.shape[1])],
                               n_restarts_optimizer=0, normalize_y=True, backend=backend)
    group = DeviceGPGroup.from_sklearn(fitted)
    return [_GroupMember(group, i, g) for i, g in enumerate(fitted)]


def predict_gps(models, X):
    """(n_samples, n_models) predicted means (mobo_penalty.py:265-273)."""
    X = np.asarray(X, np.float64)
    cache = {}
    cols = []
    for mdl in models:
        gid = id(mdl.group)
        if gid not in cache:
            cache[gid] = mdl.group.predict(X, return_std=False)[0]
        cols.append(cache[gid][mdl.index])
    return np.stack(cols, axis=1)


def penalized_acquisition(x_candidates, obj_gps, cv_gp, lam):
    """-sum_j (mu_j + lam*mu_cv) (mobo_penalty.py:275-287)."""
    obj_mu = predict_gps(obj_gps, x_candidates)
    cv_mu = predict_gps([cv_gp], x_candidates)[:, 0]
    return -np.sum(obj_mu + lam * cv_mu.reshape(-1, 1), axis=1)


_GENE_ORDER = ["filters", "kernel_size", "use_bn", "residual_blocks", "fc_layers", "use_dropout"]


def hparams_to_vector(hp):
    """Index / (n_options - 1) per gene in [0,1]^6 (mobo_penalty.py:305-318)."""
    return np.array([HPARAM_SPACE[g].index(hp[g]) / (len(HPARAM_SPACE[g]) - 1) for g in _GENE_ORDER])


def vector_to_hparams(vec):
    """Round each coordinate back to the option grid (mobo_penalty.py:320-338)."""
    return {g: HPARAM_SPACE[g][int(round(vec[i] * (len(HPARAM_SPACE[g]) - 1)))] for i, g in enumerate(_GENE_ORDER)}
