"""Small driver for ncu: a few launches of the MFCC kernel on device-resident clips."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmoop_audio_processing_b200.features import MfccFrontEnd, MfccConfig

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_mfcc = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wave = torch.rand((clips, 16000), device="cuda") * 2 - 1
fe = MfccFrontEnd(MfccConfig(n_mfcc=n_mfcc))
out = None
for _ in range(4):
    out = fe(wave, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fe(wave, out=out)
e1.record()
torch.cuda.synchronize()
print(f"clips={clips} n_mfcc={n_mfcc} ms/launch={e0.elapsed_time(e1)/5:.3f} checksum={float(out.sum()):.3f}")
