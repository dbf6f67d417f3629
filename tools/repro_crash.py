import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
n_train, group = int(sys.argv[1]), int(sys.argv[2])
seq = sys.argv[3] if len(sys.argv) > 3 else "0123"
rng = np.random.default_rng(0)
xt = rng.standard_normal((n_train, 49, 40, 1)).astype(np.float32); yt = rng.integers(0, 12, n_train)
xv = rng.standard_normal((256, 49, 40, 1)).astype(np.float32); yv = rng.integers(0, 12, 256)
prob = FitnessProblem(xt, yt, xv, yv, classes=12, config=TrainConfig(variant="B", epochs=1, patience=1, precision="bf16"))
cfgs = [dict(filters=16, kernel_size=3, use_bn=bn, residual_blocks=r, fc_layers=1, use_dropout=False) for r, bn in itertools.product((1, 2), (False, True))]
for ch in seq:
    hp = cfgs[int(ch)]
    out, _ = prob.train_eval([hp] * group, list(range(group)))
    print(ch, hp["residual_blocks"], hp["use_bn"], out[0][:3], flush=True)
print("OK", flush=True)
