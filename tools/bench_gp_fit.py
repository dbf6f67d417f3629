"""Time the multi-start GP hyper-parameter fit of one SurrogateManager.update (4 models x 11 starts) on the host worker
pool and with the device objective (csrc/gp_lml.cu).  usage: python tools/bench_gp_fit.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

    from cmoop_audio_processing_b200 import gp_fit
    from cmoop_audio_processing_b200.gp_fit import fit_gprs_parallel

    sizes = [int(a) for a in sys.argv[1:]] or [64, 144, 288]
    space = [(f, k, r, fc, bn, 1 - bn, dr, 1 - dr) for f in (16, 32, 64, 128) for k in (3, 5) for r in (1, 2, 3)
             for fc in (1, 2, 3) for bn in (0, 1) for dr in (0, 1)]
    for n in sizes:
        rng = np.random.default_rng(n)
        x = np.asarray([space[i] for i in rng.permutation(len(space))[:n]], np.float64)
        f = x[:, 0] / 128.0
        ys = [-0.9 + 0.2 * np.exp(-f) + 0.02 * rng.standard_normal(n), 0.1 * x[:, 0] * x[:, 1] / 50.0 + 0.3 * x[:, 2],
              0.05 + 0.02 * rng.standard_normal(n) + 0.01 * x[:, 3], np.maximum(0.0, 0.3 - f + 0.05 * rng.standard_normal(n))]
        ys = [(y - y.mean()) / y.std() for y in ys]
        kernels = [ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1) for _ in ys]
        out = {"device": [], "host": []}
        for backend in ("device", "host", "device", "host", "device"):
            t0 = time.perf_counter()
            g = fit_gprs_parallel(kernels, x, ys, n_restarts_optimizer=10, random_state=1, backend=backend)
            out[backend].append(time.perf_counter() - t0)
            out[backend + "_lml"] = [m.log_marginal_likelihood_value_ for m in g]
        print(f"n={n}: device {np.round(out['device'], 3).tolist()} s ({gp_fit.LAST_DEVICE_FIT}), host pool "
              f"{np.round(out['host'], 3).tolist()} s ({os.cpu_count()} cores); best lml device "
              f"{np.round(out['device_lml'], 6).tolist()} host {np.round(out['host_lml'], 6).tolist()}", flush=True)


if __name__ == "__main__":
    main()
