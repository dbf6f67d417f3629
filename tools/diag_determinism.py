"""Run-to-run and wave-composition determinism of the bf16 / fp32 candidate evaluation (bitwise)."""
import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_gpu_cnn as t
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
xt, yt, xv, yv = t.make_data(128, 64)
hps = [hp for _, hp in t.GENOTYPES] + [dict(filters=64, kernel_size=3, use_bn=True, residual_blocks=2, fc_layers=2, use_dropout=True)]
seeds = list(range(21, 21 + len(hps)))
ref = None
for rep, budget in enumerate((0.0, 0.0, 0.0, 1.2e8, 1.2e8, 6e7)):
    cfg = TrainConfig(variant="B", epochs=2, patience=2, precision=prec, memory_budget_bytes=budget)
    prob = FitnessProblem(xt, yt, xv, yv, classes=t.N_CLASSES, config=cfg)
    out, hist = prob.train_eval(hps, seeds, want_history=True)
    if ref is None:
        ref = out
    bad = [i for i in range(len(hps)) if not np.array_equal(out[i], ref[i])]
    print(prec, "budget", budget, "rows differing from run 0:", bad, [float(out[i, 4]) for i in bad])
for i in range(len(hps)):
    alone = FitnessProblem(xt, yt, xv, yv, classes=t.N_CLASSES, config=TrainConfig(variant="B", epochs=2, patience=2, precision=prec)).train_eval([hps[i]], [seeds[i]])[0]
    print("alone", i, np.array_equal(alone[0], ref[i]), alone[0][4], ref[i][4])
