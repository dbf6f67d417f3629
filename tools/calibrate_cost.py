"""Measured per-genotype cost table for the candidate sharding (dist.candidate_cost): device milliseconds of one training
epoch + validation pass per candidate, measured on groups of identical genotypes (the grouped launches scale linearly in
the number of candidates), for every (filters, kernel_size, residual_blocks, use_bn) of the reference's space and both
CNN variants at the given feature-map shape.  fc_layers / use_dropout move the cost by < 2 % and are taken at their
cheapest setting.  Writes cmoop_audio_processing_b200/cost_table.json (committed; re-run on new hardware / shapes).

    python tools/calibrate_cost.py [n_train] [group] [H] [W] [classes]
"""
import itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cmoop_audio_processing_b200 import _lib
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
group = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H = int(sys.argv[3]) if len(sys.argv) > 3 else 49
W = int(sys.argv[4]) if len(sys.argv) > 4 else 40
classes = int(sys.argv[5]) if len(sys.argv) > 5 else 12
lib = _lib.load()
rng = np.random.default_rng(0)
xt = rng.standard_normal((n_train, H, W, 1)).astype(np.float32)
yt = rng.integers(0, classes, n_train)
xv = rng.standard_normal((256, H, W, 1)).astype(np.float32)
yv = rng.integers(0, classes, 256)
table = {}
for variant in ("B", "A"):
    prob = FitnessProblem(xt, yt, xv, yv, classes=classes, config=TrainConfig(variant=variant, epochs=1, patience=1, precision="bf16"))
    for f, k, r, bn in itertools.product((16, 32, 64), (3, 5), (1, 2, 3), (False, True)):
        hp = dict(filters=f, kernel_size=k, use_bn=bn, residual_blocks=r, fc_layers=1, use_dropout=False)
        prob.train_eval([hp] * group, list(range(group)))                 # warm-up (arena, kernels)
        prob.train_eval([hp] * group, list(range(group)))
        ms = lib.cmoop_cnn_last_device_ms() / group
        table[f"{variant}:{f}:{k}:{r}:{int(bn)}"] = round(ms * 1024.0 / n_train, 4)      # per 1 024 training clips
        print(variant, hp, f"{ms:.3f} ms / candidate", flush=True)
    prob.data.close()
out = {"shape": [H, W], "classes": classes, "unit": "device ms per candidate per epoch of 1024 training clips (+ 256 validation clips)",
       "group": group, "n_train": n_train, "precision": "bf16", "table": table}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cmoop_audio_processing_b200", "cost_table.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
print("wrote", path)
