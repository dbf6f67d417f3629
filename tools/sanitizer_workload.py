import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_gpu_conv_tc as t
import torch
from cmoop_audio_processing_b200.features import MfccFrontEnd, MfccConfig
rng = np.random.default_rng(0)
# patch-resident conv (forward + dgrad), stem, wgrad_tc2 on small ragged shapes
for (n, H, W, Cin, Cout, k) in [(3, 9, 8, 16, 16, 3), (5, 13, 10, 64, 128, 3), (2, 25, 20, 32, 64, 5), (9, 7, 5, 128, 256, 3)]:
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32); w = (rng.standard_normal((k, k, Cin, Cout)) * 0.05).astype(np.float32)
    b = np.zeros(Cout, np.float32); dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    t.run_conv(0, 3, x, w, b, n, H, W, Cin, Cout, k, 1, 1); t.run_conv(1, 3, dy, w, None, n, H, W, Cin, Cout, k, 1, 0)
    t.run_wgrad(3, x, dy, n, H, W, Cin, Cout, k, 1, 1); t.run_wgrad(1, x, dy, n, H, W, Cin, Cout, k, 1, 3)
for (n, H, W, Cout, k) in [(7, 49, 40, 32, 5), (1, 49, 40, 64, 3), (3, 33, 37, 16, 3)]:
    x = rng.standard_normal((n, H, W, 1)).astype(np.float32); w = rng.standard_normal((k, k, 1, Cout)).astype(np.float32)
    b = np.zeros(Cout, np.float32); dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    t.run_conv(0, 2, x, w, b, n, H, W, 1, Cout, k, 1, 1); t.run_wgrad(2, x, dy, n, H, W, 1, Cout, k, 1, 1)
# MFCC pair kernel: odd clip counts, both outputs
for clips, n_mfcc in [(1, 40), (7, 40), (33, 0), (300, 40)]:
    wave = torch.rand((clips, 16000), device="cuda") * 2 - 1
    out = MfccFrontEnd(MfccConfig(n_mfcc=n_mfcc))(wave)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
# a tiny training run through the engine (bf16 path, BN, dropout, residual blocks)
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
xt = rng.standard_normal((96, 49, 40, 1)).astype(np.float32); yt = rng.integers(0, 12, 96)
hps = [dict(filters=16, kernel_size=3, use_bn=True, residual_blocks=2, fc_layers=2, use_dropout=True),
       dict(filters=32, kernel_size=5, use_bn=False, residual_blocks=3, fc_layers=1, use_dropout=False)]
out, _ = FitnessProblem(xt, yt, xt[:40], yt[:40], classes=12, config=TrainConfig(variant="B", epochs=1, precision="bf16")).train_eval(hps, [1, 2])
assert np.isfinite(out).all()
print("sanitizer workload done")
