"""Time cmoop_cnn_pop_train_eval on a synthetic GSC-shaped problem; prints evals/s and achieved FLOP/s."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig, forward_macs
from cmoop_audio_processing_b200.nsga import HPARAM_SPACE

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 16
variant = sys.argv[2] if len(sys.argv) > 2 else "B"
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n_train = int(sys.argv[4]) if len(sys.argv) > 4 else 3072
prec = sys.argv[5] if len(sys.argv) > 5 else "fp32"
n_val = 768
rng = np.random.default_rng(0)
xt = rng.standard_normal((n_train, 49, 40, 1)).astype(np.float32)
yt = rng.integers(0, 12, n_train)
xv = rng.standard_normal((n_val, 49, 40, 1)).astype(np.float32)
yv = rng.integers(0, 12, n_val)
random.seed(0)
hps = [{k: random.choice(v) for k, v in HPARAM_SPACE.items()} for _ in range(pop)]
prob = FitnessProblem(xt, yt, xv, yv, classes=12, config=TrainConfig(variant=variant, epochs=epochs, patience=epochs, precision=prec))
# warm-up with the WHOLE population for one short epoch so the (persistent) activation arena is allocated and the
# kernels are loaded before the timed call
warm = FitnessProblem(xt[:128], yt[:128], xv[:64], yv[:64], classes=12,
                      config=TrainConfig(variant=variant, epochs=1, patience=1, precision=prec))
warm.train_eval(hps, list(range(pop)))
t0 = time.perf_counter()
out, _ = prob.train_eval(hps, list(range(pop)))
dt = time.perf_counter() - t0
macs = [forward_macs(hp, 49, 40, 12, variant) for hp in hps]
flops = sum(o[3] * (6 * m * n_train + 2 * m * n_val) + 2 * m * n_val for o, m in zip(out, macs))
print(f"[{prec}] pop={pop} variant={variant} epochs={epochs} n_train={n_train}: {dt:.2f} s, {pop/dt:.3f} evals/s, "
      f"{flops/dt/1e12:.2f} TFLOP/s (analytic), mean fwd MACs {np.mean(macs)/1e6:.1f} M")
