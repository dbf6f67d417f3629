"""Time the patch-layout weight-gradient kernel on single shapes through the debug hook (CUDA events are not exposed
there, so wall-clock of a few repeats; the kernel dominates)."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_gpu_conv_tc as t
rng = np.random.default_rng(0)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for (n, H, W, Cin, Cout, k) in [(64, 25, 20, 64, 128, 3), (64, 25, 20, 64, 64, 5), (64, 13, 10, 128, 256, 3)]:
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32)
    dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    t.run_wgrad(mode, x, dy, n, H, W, Cin, Cout, k, 1, 8)
    t0 = time.perf_counter()
    for _ in range(3):
        t.run_wgrad(mode, x, dy, n, H, W, Cin, Cout, k, 1, 8)
    print(mode, (n, H, W, Cin, Cout, k), f"{(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per call (incl. host staging)", flush=True)
