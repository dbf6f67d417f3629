"""One SA-NSGA-II run in the shape of BASELINE configs[2] (sa_nsga_local.py flow: surrogate + Lamarckian local search,
infill 0.334) on synthetic GSC-shaped features; prints per-generation wall-clock split into true evaluations (GPU),
surrogate update (host, concurrent multi-start fit) and the rest (local search, NDS/crowding, variation), plus HV.

    python tools/bench_generation.py [pop] [generations] [n_train] [epoch_cap] [precision]
"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings
warnings.filterwarnings("ignore")
import numpy as np
from cmoop_audio_processing_b200 import drivers, synth
from cmoop_audio_processing_b200.features import MfccFrontEnd
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_train = int(sys.argv[3]) if len(sys.argv) > 3 else 12 * 256
epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 6
prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"

import torch
t0 = time.perf_counter()
wave, labels = synth.make_clips(n_train + 768, 12, seed=1234)
fe = MfccFrontEnd()
feats = fe(torch.from_numpy(wave).cuda()).cpu().numpy()
mu, sd = feats[:n_train].mean(axis=(0, 1)), feats[:n_train].std(axis=(0, 1)) + 1e-6
feats = ((feats - mu) / sd).astype(np.float32)[..., None]
print(f"features for {len(wave)} clips: {time.perf_counter() - t0:.2f} s (synthesis on the host included)", flush=True)
prob = FitnessProblem.sa_nsga_local(feats[:n_train], labels[:n_train], feats[n_train:], labels[n_train:], classes=12,
                                    config=TrainConfig(variant="B", epochs=epochs, patience=5, restore_best_weights=True,
                                                       acc_from="evaluate", fpr_mode="filtered", precision=prec))
random.seed(0)
np.random.seed(0)
ops = drivers.default_ops(prob)
upd = {"s": 0.0}
SM = ops.SurrogateManager
class TimedSM(SM):
    def update(self, *a, **k):
        t = time.perf_counter()
        r = super().update(*a, **k)
        upd["s"] += time.perf_counter() - t
        return r
ops.SurrogateManager = TimedSM
t0 = time.perf_counter()
front, history, timings = drivers.sa_nsga2(pop, gens, 0.334, ops)
total = time.perf_counter() - t0
print(f"pop {pop}, {gens} generations, {n_train} train clips, epoch cap {epochs}, {prec}: total {total:.2f} s "
      f"(initial population of {pop} true evaluations included), surrogate updates {upd['s']:.2f} s")
for t in timings:
    print(f"  gen {t['generation']}: {t['seconds']:.2f} s, {t['true_evals']} true evals in {t['eval_seconds']:.2f} s "
          f"({t['true_evals'] / t['eval_seconds']:.1f} evals/s)")
ind = drivers.front_indicators(history[-1])
print("final population:", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in ind.items()})
