"""Per-kernel-family device time of one cmoop_cnn_pop_train_eval call (CUDA-event pairs around every grouped launch,
cmoop_profile_enable): the step breakdown used to pick the next kernel to work on.

    python tools/prof_step.py [pop] [variant] [epochs] [n_train] [precision]
"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cmoop_audio_processing_b200 import _lib
from cmoop_audio_processing_b200.nsga import HPARAM_SPACE
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig, forward_macs

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variant = sys.argv[2] if len(sys.argv) > 2 else "B"
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n_train = int(sys.argv[4]) if len(sys.argv) > 4 else 3072
prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"
lib = _lib.load()
rng = np.random.default_rng(0)
xt = rng.standard_normal((n_train, 49, 40, 1)).astype(np.float32)
yt = rng.integers(0, 12, n_train)
xv = rng.standard_normal((768, 49, 40, 1)).astype(np.float32)
yv = rng.integers(0, 12, 768)
random.seed(0)
hps = [{k: random.choice(v) for k, v in HPARAM_SPACE.items()} for _ in range(pop)]
for spec in os.environ.get("PROF_FIX", "").split(","):          # e.g. PROF_FIX=filters=64,kernel_size=3: pin genes of every candidate
    if "=" in spec:
        key, val = spec.split("=")
        for hp in hps:
            hp[key] = type(HPARAM_SPACE[key][0])(int(val))
prob = FitnessProblem(xt, yt, xv, yv, classes=12, config=TrainConfig(variant=variant, epochs=epochs, patience=epochs, precision=prec))
prob.train_eval(hps, list(range(pop)))
lib.cmoop_profile_enable(1)
out, _ = prob.train_eval(hps, list(range(pop)))
lib.cmoop_profile_enable(0)
table = _lib.profile_table()
total = sum(v["ms"] for v in table.values())
macs = [forward_macs(hp, 49, 40, 12, variant) for hp in hps]
flops = sum(o[3] * (6 * m * n_train + 2 * m * 768) + 2 * m * 768 for o, m in zip(out, macs))
print(f"[{prec}] pop={pop} variant={variant} epochs={epochs} n_train={n_train}: device {lib.cmoop_cnn_last_device_ms():.1f} ms "
      f"(sum of kernels {total:.1f} ms), {flops / lib.cmoop_cnn_last_device_ms() / 1e9:.1f} TFLOP/s analytic")
for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
    tf = f"{v['flops'] / v['ms'] / 1e9:8.1f} TFLOP/s" if v["flops"] else ""
    print(f"  {k:20s} n={v['launches']:6d} ms={v['ms']:9.2f} {100 * v['ms'] / total:5.1f}%  avg_us={1e3 * v['ms'] / v['launches']:8.1f} {tf}")
