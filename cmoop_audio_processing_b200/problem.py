"""Drop-in fitness evaluation: ``evaluate_individual`` / ``compute_objectives_and_constraints``.

Mirrors the reference seam (nsga_penalty.py:368-442, sa_nsga_penalty.py:205-253, mobo_penalty.py:218-247 and the
bi-objective ablation_study/*_nsga_1.py forms): same names, same ``{'hparams','objs','CV'}`` records with Python
floats, same penalty arithmetic ``CV = sum(max(0, violation))``.  Where the reference reads module globals
(X_train, CLASSES, MIN_ACCURACY ...) a ``FitnessProblem`` is constructed once and its bound methods are installed
under the reference's names.

Underneath, the whole population is trained and scored in ONE call of ``cmoop_cnn_pop_train_eval`` (grouped CUDA
launches over all candidates) instead of the reference's serial loop; with torch.distributed initialised the
candidates are partitioned across ranks (longest-processing-time first on the analytic cost) and the objective rows
are all-gathered, so every rank returns the full, identically ordered result list.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .nsga import HPARAM_SPACE  # noqa: F401  (re-exported for drivers)

FC_UNITS = {1: [64], 2: [128, 64], 3: [256, 128, 64], 4: [512, 256, 128, 64]}


@dataclass
class TrainConfig:
    """Training recipe + per-script policy (SURVEY.md section 2.2 matrix)."""
    variant: str = "A"                   # "A": nsga_penalty.py / mobo_penalty.py ; "B": sa_nsga_*.py
    epochs: int = 300                    # nsga_penalty.py:177
    batch_size: int = 64                 # nsga_penalty.py:178
    patience: int = 5                    # nsga_penalty.py:179
    restore_best_weights: bool = False   # nsga_penalty.py:382 (False) vs sa_nsga_penalty.py:215 (True)
    acc_from: str = "history"            # "history": history['val_accuracy'][-1] ; "evaluate": model.evaluate
    y_true_mode: str = "flatten"         # "argmax_quirk": nsga_penalty.py:387 feeds all-zero labels to the FPR
    fpr_mode: str = "all"                # "filtered": sa_nsga_local.py:138-141
    learning_rate: float = 1e-3          # Keras Adam default (LEARNING_RATE at nsga_penalty.py:162 is unused)
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-7
    bn_momentum: float = 0.99
    bn_eps: float = 1e-3
    dropout_rate: float = 0.3            # nsga_penalty.py:323 (the docstring says 0.2, the code is 0.3)
    precision: str = "fp32"
    memory_budget_bytes: float = 0.0


def param_count(hp: dict, n_classes: int, variant: str = "A") -> int:
    """Keras count_params() as a closed form of the genotype (BN counts its 2 moving statistics too)."""
    f, k = int(hp["filters"]), int(hp["kernel_size"])
    bn = 4 if hp["use_bn"] else 0
    total = k * k * f + f + bn * f
    if variant == "A":
        total += k * k * f * f + f + bn * f
    for _ in range(int(hp["residual_blocks"])):
        total += f * 2 * f + 2 * f
        total += k * k * f * 2 * f + 2 * f + bn * 2 * f
        if variant == "A":
            total += k * k * 2 * f * 2 * f + 2 * f + bn * 2 * f
        f *= 2
    width = f
    for units in FC_UNITS.get(int(hp["fc_layers"]), []):
        total += width * units + units
        width = units
    return total + width * n_classes + n_classes


def compute_model_size_mb(hp: dict, n_classes: int, variant: str = "A") -> float:
    """nsga_penalty.py:337-344: count_params() * 4 / 1024**2."""
    return param_count(hp, n_classes, variant) * 4 / (1024 ** 2)


def forward_macs(hp: dict, height: int, width: int, n_classes: int, variant: str = "A") -> int:
    """Analytic forward multiply-accumulates per sample (cost model for load balancing and FLOP accounting)."""
    f, k = int(hp["filters"]), int(hp["kernel_size"])
    h, w = height, width
    macs = h * w * k * k * f
    if variant == "A":
        macs += h * w * k * k * f * f
    h, w = (h + 1) // 2, (w + 1) // 2
    for _ in range(int(hp["residual_blocks"])):
        ho, wo = (h + 1) // 2, (w + 1) // 2
        macs += ho * wo * f * 2 * f
        macs += h * w * k * k * f * 2 * f
        if variant == "A":
            macs += h * w * k * k * 2 * f * 2 * f
        f *= 2
        h, w = ho, wo
    width_ = f
    for units in FC_UNITS.get(int(hp["fc_layers"]), []):
        macs += width_ * units
        width_ = units
    return macs + width_ * n_classes


FPR_MODES = {"all": 0, "filtered": 1, "vectorised": 2}


def calculate_fpr(y_true, y_pred, num_classes, mode: str = "all", return_confusion: bool = False):
    """Macro-averaged false-positive rate of fixed label / prediction vectors -- the reference's ``calculate_fpr``:
    mode "all" nsga_penalty.py:351-364, "filtered" ablation_study/sa_nsga_local.py:138-141, "vectorised"
    ablation_study/init_sa_nsga_local.py:137-143.  Confusion matrix by integer atomics on the device, per-class rates and
    numpy-ordered mean in the library (``cmoop_fpr_from_predictions_host``): bit-identical to the reference's float."""
    lib = _lib.load()
    yt = np.ascontiguousarray(np.asarray(y_true).reshape(-1), np.int32)
    yp = np.ascontiguousarray(np.asarray(y_pred).reshape(-1), np.int32)
    if yt.shape != yp.shape:
        raise ValueError("y_true and y_pred differ in length")
    out = np.zeros(1, np.float64)
    cm = np.zeros((int(num_classes), int(num_classes)), np.int32) if return_confusion else None
    _lib.check(lib.cmoop_fpr_from_predictions_host(_lib.ptr(yt), _lib.ptr(yp), len(yt), int(num_classes),
                                                   FPR_MODES[mode], _lib.ptr(out), _lib.ptr(cm)),
               "cmoop_fpr_from_predictions_host")
    return (float(out[0]), cm) if return_confusion else float(out[0])


class _Genotype(C.Structure):
    _fields_ = [("filters", C.c_int), ("kernel_size", C.c_int), ("use_bn", C.c_int), ("residual_blocks", C.c_int),
                ("fc_layers", C.c_int), ("use_dropout", C.c_int)]


class _CnnConfig(C.Structure):
    _fields_ = [("variant", C.c_int), ("n_classes", C.c_int), ("batch_size", C.c_int), ("max_epochs", C.c_int),
                ("patience", C.c_int), ("restore_best_weights", C.c_int), ("acc_from_history", C.c_int),
                ("y_true_zero", C.c_int), ("fpr_filtered", C.c_int), ("learning_rate", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float), ("bn_momentum", C.c_float),
                ("bn_eps", C.c_float), ("dropout_rate", C.c_float), ("precision", C.c_int),
                ("memory_budget_bytes", C.c_double)]


def _genotypes(hps) -> C.Array:
    arr = (_Genotype * len(hps))()
    for i, hp in enumerate(hps):
        arr[i] = _Genotype(int(hp["filters"]), int(hp["kernel_size"]), int(bool(hp["use_bn"])),
                           int(hp["residual_blocks"]), int(hp["fc_layers"]), int(bool(hp["use_dropout"])))
    return arr


class DeviceDataset:
    """Train/validation features resident in HBM (the reference's X_train / X_validation globals)."""

    def __init__(self, x_train, y_train, x_val, y_val):
        self._lib = _lib.load()
        _lib.bind_device()
        if hasattr(x_train, "is_cuda") and x_train.is_cuda:
            self._init_from_device(x_train, y_train, x_val, y_val)
            return
        xt = np.ascontiguousarray(np.asarray(x_train, np.float32).reshape(len(x_train), *np.shape(x_train)[1:3]))
        xv = np.ascontiguousarray(np.asarray(x_val, np.float32).reshape(len(x_val), *np.shape(x_val)[1:3]))
        yt = np.ascontiguousarray(np.asarray(y_train).reshape(-1), np.int32)
        yv = np.ascontiguousarray(np.asarray(y_val).reshape(-1), np.int32)
        if xt.shape[1:] != xv.shape[1:]:
            raise ValueError("train and validation features differ in shape")
        self.n_train, self.height, self.width = xt.shape
        self.n_val = xv.shape[0]
        handle = C.c_void_p()
        _lib.check(self._lib.cmoop_cnn_dataset_create_host(_lib.ptr(xt), _lib.ptr(yt), self.n_train, _lib.ptr(xv),
                                                           _lib.ptr(yv), self.n_val, self.height, self.width,
                                                           C.byref(handle)), "cmoop_cnn_dataset_create_host")
        self._handle = handle

    def _init_from_device(self, x_train, y_train, x_val, y_val):
        """Features that already live on the GPU (CUDA torch tensors (N, H, W[, 1]) float32, e.g. straight out of
        MfccFrontEnd): device-to-device copies, the feature tensors never touch the host."""
        import torch
        if not (x_val.is_cuda and x_train.dtype == torch.float32 and x_val.dtype == torch.float32):
            raise ValueError("device features must be float32 CUDA tensors")
        xt = x_train.reshape(x_train.shape[0], x_train.shape[1], x_train.shape[2]).contiguous()
        xv = x_val.reshape(x_val.shape[0], x_val.shape[1], x_val.shape[2]).contiguous()
        if xt.shape[1:] != xv.shape[1:]:
            raise ValueError("train and validation features differ in shape")
        to_np = lambda y: np.ascontiguousarray((y.cpu().numpy() if hasattr(y, "cpu") else np.asarray(y)).reshape(-1), np.int32)
        yt, yv = to_np(y_train), to_np(y_val)
        self.n_train, self.height, self.width = (int(v) for v in xt.shape)
        self.n_val = int(xv.shape[0])
        stream = torch.cuda.current_stream(xt.device).cuda_stream
        handle = C.c_void_p()
        with torch.cuda.device(xt.device):
            _lib.check(self._lib.cmoop_cnn_dataset_create_dev(C.c_void_p(xt.data_ptr()), _lib.ptr(yt), self.n_train,
                                                              C.c_void_p(xv.data_ptr()), _lib.ptr(yv), self.n_val, self.height,
                                                              self.width, C.c_void_p(stream), C.byref(handle)),
                       "cmoop_cnn_dataset_create_dev")
        self._handle = handle

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.cmoop_cnn_dataset_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FitnessProblem:
    """The reference's module-level evaluation state as an object.

    objectives: subset/order of ("neg_acc", "size", "fpr") -- 3 for nsga_penalty.py, 2 for the ablation
    sub-problems (the dropped quantity is stored under its ``*_metric`` key and leaves the constraint sum).
    """

    def __init__(self, x_train, y_train, x_val, y_val, *, classes: int, config: TrainConfig = TrainConfig(),
                 min_accuracy: float = 0.9, max_model_size: float = 2.5, max_fpr: float = 0.1,
                 objectives=("neg_acc", "size", "fpr"), seed: int = 0, memoise: bool = False):
        self.data = x_train if isinstance(x_train, DeviceDataset) else DeviceDataset(x_train, y_train, x_val, y_val)
        self.classes = int(classes)
        self.config = config
        self.min_accuracy, self.max_model_size, self.max_fpr = min_accuracy, max_model_size, max_fpr
        self.objectives = tuple(objectives)
        self.seed = int(seed)
        self.evaluations = 0            # true evaluations so far (also the seed counter)
        # Opt-in: the genotype space has 288 points and the reference re-trains duplicates
        # (keeping the last result, sa_nsga_penalty.py:325-327); with memoise=True a genotype is trained once per problem
        # and repeated requests return the stored row.  It changes the evaluation count, hence off by default.
        self.memoise = bool(memoise)
        self._memo = {}
        self.last_details = None
        self._lib = _lib.load()

    # ---- presets named after the scripts whose globals they reproduce
    @classmethod
    def nsga_penalty(cls, *data, classes=10, **kw):
        cfg = TrainConfig(variant="A", restore_best_weights=False, acc_from="history", y_true_mode="argmax_quirk")
        return cls(*data, classes=classes, config=kw.pop("config", cfg), min_accuracy=0.9, max_model_size=2.5,
                   max_fpr=0.1, **kw)

    @classmethod
    def mobo_penalty(cls, *data, classes=10, **kw):
        cfg = TrainConfig(variant="A", restore_best_weights=True, acc_from="history")
        return cls(*data, classes=classes, config=kw.pop("config", cfg), min_accuracy=0.90, max_model_size=2.5,
                   max_fpr=0.09, **kw)

    @classmethod
    def sa_nsga_penalty(cls, *data, classes=11, **kw):
        cfg = TrainConfig(variant="B", restore_best_weights=True, acc_from="evaluate")
        return cls(*data, classes=classes, config=kw.pop("config", cfg), min_accuracy=0.75, max_model_size=2.5,
                   max_fpr=0.09, **kw)

    @classmethod
    def sa_nsga_local(cls, *data, classes=10, **kw):
        cfg = TrainConfig(variant="B", restore_best_weights=True, acc_from="evaluate", fpr_mode="filtered")
        return cls(*data, classes=classes, config=kw.pop("config", cfg), min_accuracy=0.90, max_model_size=2.5,
                   max_fpr=0.09, **kw)

    # ---- C-ABI plumbing
    def _c_config(self) -> _CnnConfig:
        c = self.config
        if c.precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        return _CnnConfig(0 if c.variant == "A" else 1, self.classes, c.batch_size, c.epochs, c.patience,
                          int(c.restore_best_weights), int(c.acc_from == "history"),
                          int(c.y_true_mode == "argmax_quirk"), int(c.fpr_mode == "filtered"), c.learning_rate,
                          c.beta1, c.beta2, c.adam_eps, c.bn_momentum, c.bn_eps, c.dropout_rate,
                          0 if c.precision == "fp32" else 1, float(c.memory_budget_bytes))

    def train_eval(self, hparams_list, seeds=None, want_history=False):
        """Array-level entry: returns out [P,6] = (acc, size_mb, fpr, epochs_run, last_val_loss, best_val_loss)
        and optionally history [P, epochs, 3]."""
        p = len(hparams_list)
        out = np.zeros((p, 6), np.float64)
        hist = np.full((p, self.config.epochs, 3), np.nan) if want_history else None
        if p == 0:
            return out, hist
        if seeds is None:
            seeds = [self.seed + self.evaluations + i for i in range(p)]
        seeds = np.ascontiguousarray(seeds, np.uint64)
        cfg = self._c_config()
        _lib.check(self._lib.cmoop_cnn_pop_train_eval(self.data._handle, _genotypes(hparams_list), _lib.ptr(seeds), p,
                                                      C.byref(cfg), _lib.ptr(out), _lib.ptr(hist)),
                   "cmoop_cnn_pop_train_eval")
        return out, hist

    def _sharded_train_eval(self, hparams_list):
        """Partition candidates over torch.distributed ranks, all-gather the [P,6] rows (NCCL / gloo)."""
        try:
            import torch
            import torch.distributed as dist
        except ImportError:  # pragma: no cover
            dist = None
        p = len(hparams_list)
        seeds = [self.seed + self.evaluations + i for i in range(p)]
        if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            out, _ = self.train_eval(hparams_list, seeds)
            return out
        from .dist import assign_lpt, candidate_cost
        world, rank = dist.get_world_size(), dist.get_rank()
        # measured per-genotype device time (dist.py), not forward MACs: MACs over-weight wide candidates ~10x
        cost = [candidate_cost(hp, self.data.height, self.data.width, self.classes, self.config.variant)
                for hp in hparams_list]
        owner = assign_lpt(cost, world)
        mine = [i for i in range(p) if owner[i] == rank]
        local, _ = self.train_eval([hparams_list[i] for i in mine], [seeds[i] for i in mine])
        full = np.zeros((p, 6), np.float64)
        for j, i in enumerate(mine):
            full[i] = local[j]
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(full).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)     # rows are disjoint: the sum is the all-gather of objective rows
        return t.cpu().numpy()

    # ---- the reference's names
    def evaluate_individual(self, hparams):
        """(accuracy, size_mb, fpr) of one genotype -- nsga_penalty.py:368-395."""
        out, _ = self.train_eval([hparams])
        self.evaluations += 1
        self.last_details = out
        return float(out[0, 0]), float(out[0, 1]), float(out[0, 2])

    def compute_objectives_and_constraints(self, population):
        """list[{'hparams','objs','CV'}] in input order -- nsga_penalty.py:418-442."""
        if getattr(self, "memoise", False):
            keys = [tuple(sorted(hp.items())) for hp in population]
            todo = [i for i, k in enumerate(keys) if k not in self._memo and k not in keys[:i]]
            if todo:
                rows = self._sharded_train_eval([population[i] for i in todo])
                self.evaluations += len(todo)
                for i, row in zip(todo, rows):
                    self._memo[keys[i]] = np.array(row, np.float64)
            out = np.stack([self._memo[k] for k in keys]) if keys else np.zeros((0, 6))
        else:
            out = self._sharded_train_eval(population)
            self.evaluations += len(population)
        self.last_details = out
        results = []
        for ind, row in zip(population, out):
            acc, size_mb, fpr = float(row[0]), float(row[1]), float(row[2])
            vals = {"neg_acc": -acc, "size": size_mb, "fpr": fpr}
            viol = {"neg_acc": max(0.0, self.min_accuracy - acc), "size": max(0.0, size_mb - self.max_model_size),
                    "fpr": max(0.0, fpr - self.max_fpr)}
            cv = 0.0
            for key in ("neg_acc", "size", "fpr"):
                if key in self.objectives:
                    cv = cv + viol[key]
            rec = {"hparams": ind, "objs": [vals[k] for k in self.objectives], "CV": cv}
            if "size" not in self.objectives:
                rec["size_metric"] = size_mb
            if "fpr" not in self.objectives:
                rec["fpr_metric"] = fpr
            if "neg_acc" not in self.objectives:
                rec["acc_metric"] = acc
            results.append(rec)
        return results

    # ---- test hooks (harness-imposed random streams shared with the CPU oracle)
    def debug_init_params(self, hparams, seed):
        n = param_count(hparams, self.classes, self.config.variant)
        out = np.zeros(n, np.float32)
        cfg = self._c_config()
        _lib.check(self._lib.cmoop_cnn_debug_init_params(_genotypes([hparams]), C.c_uint64(seed), C.byref(cfg),
                                                         _lib.ptr(out)), "cmoop_cnn_debug_init_params")
        return out

    def debug_permutation(self, seed, epoch, n=None):
        n = self.data.n_train if n is None else n
        out = np.zeros(n, np.int32)
        _lib.check(self._lib.cmoop_cnn_debug_permutation(C.c_uint64(seed), epoch, n, _lib.ptr(out)),
                   "cmoop_cnn_debug_permutation")
        return out

    def debug_train_steps(self, hparams, seed, n_steps):
        n = param_count(hparams, self.classes, self.config.variant)
        losses = np.zeros(n_steps, np.float32)
        grads = np.zeros(n, np.float32)
        params = np.zeros(n, np.float32)
        cfg = self._c_config()
        _lib.check(self._lib.cmoop_cnn_debug_train_steps(self.data._handle, _genotypes([hparams]), C.c_uint64(seed),
                                                         C.byref(cfg), n_steps, _lib.ptr(losses), _lib.ptr(grads),
                                                         _lib.ptr(params)), "cmoop_cnn_debug_train_steps")
        return losses, grads, params
