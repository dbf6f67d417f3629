import sys, time, random, numpy as np
sys.path.insert(0, '.')
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
from cmoop_audio_processing_b200.nsga import HPARAM_SPACE
rng = np.random.default_rng(0)
n_tr, n_va, H, W, C = 192, 64, 128, 313, 397
xt = rng.standard_normal((n_tr, H, W, 1)).astype(np.float32); yt = rng.integers(0, C, n_tr)
xv = rng.standard_normal((n_va, H, W, 1)).astype(np.float32); yv = rng.integers(0, C, n_va)
random.seed(1)
hps = [{k: random.choice(v) for k, v in HPARAM_SPACE.items()} for _ in range(6)]
for prec in ("bf16", "fp32"):
    prob = FitnessProblem(xt, yt, xv, yv, classes=C, config=TrainConfig(variant="B", epochs=1, patience=1, precision=prec))
    t0 = time.perf_counter()
    out, _ = prob.train_eval(hps, list(range(len(hps))))
    print(prec, f"{time.perf_counter() - t0:.2f} s", np.round(out[:, [0, 1, 2, 4]], 4).tolist(), flush=True)
    assert np.isfinite(out).all()
