#!/usr/bin/env python
"""bench.py -- candidate true-evaluations per second within a SA-NSGA-II generation (BASELINE.json metric, first clause)
on N B200s, with the MFCC front-end throughput (second clause) and the latency kernels as extra objects.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE configs[4], one generation of it): population 256, CNN variant B, 12 classes, synthetic GSC-shaped 1 s /
16 kHz clips (synth.make_clips, seed 1234) -> 49x40 MFCC features -> StandardScaler fitted on the 12x256 training clips;
12x64 validation clips; batch 64, Adam, EarlyStopping(val_loss, patience 5, restore_best_weights) with the epoch count
CAPPED at --epoch-cap (default 4; the reference's 300 is a ceiling it never reaches either, the cap is stated in `config`).

A step = ONE `compute_objectives_and_constraints(population)` call (nsga_penalty.py:418-442, the serial loop at :426) over
the whole population: the 256 candidates are partitioned over the N ranks (longest-processing-time first on a per-genotype
cost table), trained and scored on the bf16 tcgen05 path, and the objective rows are all-gathered INSIDE the timed region.
STRONG scaling: the population is fixed as N grows.  One JSON line is printed by rank 0:

  value      candidate evaluations / s over K steps: dataset resident in HBM, genotypes in, host records out
  e2e        the same metric inside a FULL SA-NSGA-II generation (drivers.sa_nsga2 = ablation_study/sa_nsga_local.py:436-554:
             tournament, variation, GP predict(return_std) + Lamarckian local search, infill selection, the true evaluations
             of the max(1, int(256*0.334)) = 85 infill points from HOST feature arrays (re-staged host->device every
             generation), surrogate update, NDS + crowding truncation of the 512 merged records, HV / IGD / Spread of the
             generation), host records to host records: true evaluations / generation wall time
  roofline   tensor: analytic FLOPs of the step (E_i*(6*MACs_i*N_train + 2*MACs_i*N_val) + 2*MACs_i*N_val per candidate)
             / device time of the step, against MEASURED_PEAKS.json bf16_tflops_sustained; `dominant_kernel` gives the same
             for the kernel family with the largest share of the step, from CUDA-event pairs around every launch
  cpu_baseline  the torch-CPU oracle (oracle/cnn_ref.py, all host threads) on a stratified sample of whole candidates
  mfcc       clips/s of the front-end kernel on 65 536 clips per GPU with its own HBM roofline and host-buffer e2e
  latency    NDS+crowding / hypervolume / GP predict: us per call device-resident and through the host ABI, next to the
             CPU port timed in the same run

--impl reference times the CPU path alone: the same candidates evaluated serially (exactly the loop of
nsga_penalty.py:426) by the torch-CPU restatement of evaluate_individual on ALL host threads (the launcher's
OMP_NUM_THREADS=1 is overridden), one or more WHOLE candidates per step drawn from cost strata of the same population.
TensorFlow / Keras are not installable here, so `kind` is "port" (DESIGN.md section 2).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import random
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate_evals_per_sec"
UNIT = "evals/s"
POP = 256
N_CLASSES = 12
N_TRAIN = 12 * 256
N_VAL = 12 * 64
INFILL = 0.334
VARIANT = "B"
H, W = 49, 40

# MFCC extra (BASELINE configs[1])
N_SAMPLES = 16000
N_FRAMES = 49
N_OUT = 40
BYTES_PER_CLIP = N_SAMPLES * 4 + N_FRAMES * N_OUT * 4      # 71 840


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:                                  # pragma: no cover
        return os.cpu_count() or 1


def use_all_host_threads() -> int:
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU legs must use every host thread at every N."""
    n = host_threads()
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = str(n)
    return n


def workload_config(args, n_gpus: int) -> dict:
    return {
        "workload": f"BASELINE configs[4] (one generation of it): SA-NSGA-II population {args.pop}, CNN variant B, "
                    f"{N_CLASSES} classes, synthetic GSC-shaped 1 s 16 kHz clips -> 49x40 MFCC, {N_TRAIN} train / {N_VAL} val "
                    f"clips, batch 64, Adam, EarlyStopping(patience 5, restore best) with the epoch count capped at "
                    f"{args.epoch_cap}; step = one compute_objectives_and_constraints(population) call",
        "population": args.pop, "classes": N_CLASSES, "n_train": N_TRAIN, "n_val": N_VAL, "epoch_cap": args.epoch_cap,
        "patience": 5, "batch_size": 64, "variant": VARIANT, "infill_percent": INFILL,
        "sharding": f"{args.pop} candidates partitioned over {n_gpus} GPU(s) (LPT on the genotype cost table), one "
                    "all-gather of the objective rows per call inside the timed region",
        "cache": "per-step working set (activations of >= 32 candidates per GPU, tens of GB) is far larger than L2 (126 MB); "
                 "no flush needed",
    }


def measured_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_sustained": float(d["bf16_tflops_sustained"]),
                "bf16_burst": float(d["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_sustained": 1400.0, "bf16_burst": 1650.0,
                "source": "fallback (B200_PROFILING.md)"}


def make_population(pop: int, seed: int = 0):
    from cmoop_audio_processing_b200.nsga import HPARAM_SPACE
    rng = random.Random(seed)
    return [{k: rng.choice(v) for k, v in HPARAM_SPACE.items()} for _ in range(pop)]


def analytic_flops(hps, epochs_run, n_final: int = 1) -> float:
    """E*(6*M*N_train + 2*M*N_val) + 2*M*N_val*n_final per candidate (n_final = the scoring passes
    actually executed: this engine takes accuracy and the confusion matrix from ONE pass)."""
    from cmoop_audio_processing_b200.problem import forward_macs
    total = 0.0
    for hp, e in zip(hps, epochs_run):
        m = forward_macs(hp, H, W, N_CLASSES, VARIANT)
        total += float(e) * (6.0 * m * N_TRAIN + 2.0 * m * N_VAL) + 2.0 * m * N_VAL * n_final
    return total


def stratified_sample(hps, m: int):
    """m candidates at the mid-quantiles of the population's analytic cost: sample i is the median of the i-th of m
    equal-count cost strata, so m / sum(t_i) is the plain stratified estimate of the population's evaluations / s."""
    from cmoop_audio_processing_b200.problem import forward_macs
    order = sorted(range(len(hps)), key=lambda i: (forward_macs(hps[i], H, W, N_CLASSES, VARIANT), i))
    picks = [order[min(len(order) - 1, int((i + 0.5) * len(order) / m))] for i in range(m)]
    return picks


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed regions run."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int, period_s: float = 0.05):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.ok and self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU side (oracle port)
def cpu_features():
    """The same synthetic clips through the plain-C MFCC oracle + StandardScaler fitted on the training split."""
    import numpy as np
    from cmoop_audio_processing_b200 import synth
    from oracle import build_c
    wave, labels = synth.make_clips(N_TRAIN + N_VAL, N_CLASSES, seed=1234)
    feats = build_c.mfcc(wave)
    mean = feats[:N_TRAIN].reshape(-1, N_OUT).mean(axis=0)
    scale = feats[:N_TRAIN].reshape(-1, N_OUT).std(axis=0)
    scale[scale == 0.0] = 1.0
    feats = ((feats - mean) / scale).astype(np.float32)[..., None]
    return (feats[:N_TRAIN], labels[:N_TRAIN].astype(np.int64), feats[N_TRAIN:], labels[N_TRAIN:].astype(np.int64))


def cpu_evaluate_candidate(hp, data, seed: int, epoch_cap: int) -> tuple[float, dict]:
    """One WHOLE candidate through the torch-CPU restatement of evaluate_individual (sa_nsga_local policy): Glorot-uniform
    initialisation, per-epoch shuffles, Adam, per-epoch validation, early stopping, restore, evaluate + FPR.  Seconds."""
    import numpy as np
    from oracle import cnn_ref
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape in cnn_ref.param_shapes(hp, N_CLASSES, VARIANT):
        if name.endswith(".w"):
            fan_in = int(np.prod(shape[:-1]))
            fan_out = int(np.prod(shape[:-2])) * shape[-1] if len(shape) == 4 else shape[-1]
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            params[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        elif name.endswith((".gamma", ".var")):
            params[name] = np.ones(shape, np.float32)
        else:
            params[name] = np.zeros(shape, np.float32)
    perms = [rng.permutation(N_TRAIN) for _ in range(epoch_cap)]
    t0 = time.perf_counter()
    res = cnn_ref.evaluate_individual(hp, data, params, perms, n_classes=N_CLASSES, variant=VARIANT, seed=seed,
                                      epochs=epoch_cap, patience=5, restore_best_weights=True, acc_from="evaluate",
                                      fpr_mode="filtered")
    return time.perf_counter() - t0, res


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()                           # before torch / the OpenMP oracle are loaded
    import torch
    torch.set_num_threads(cores)
    if torch.get_num_threads() < cores:
        raise SystemExit(f"reference arm: torch uses {torch.get_num_threads()} threads, host has {cores}")
    from cmoop_audio_processing_b200.problem import forward_macs
    data = cpu_features()
    hps = make_population(args.pop)
    per_step = max(1, -(-args.ref_min_candidates // max(1, args.steps)))   # >= 8 whole candidates in the timed region
    strata = stratified_sample(hps, args.steps * per_step)
    cheap = stratified_sample(hps, 8)[:2]                    # warm-up candidates: the two cheapest strata
    for w in range(args.warmup):
        cpu_evaluate_candidate(hps[cheap[w % len(cheap)]], data, 1000 + w, args.epoch_cap)
    times, macs, epochs = [], [], []
    t_all = time.perf_counter()
    for s in range(args.steps):
        for j in range(per_step):
            i = strata[s * per_step + j]
            dt, res = cpu_evaluate_candidate(hps[i], data, i, args.epoch_cap)
            times.append(dt)
            macs.append(forward_macs(hps[i], H, W, N_CLASSES, VARIANT))
            epochs.append(res["epochs_run"])
    total = time.perf_counter() - t_all
    n = len(times)
    value = n / total
    flops = analytic_flops([hps[i] for i in strata], epochs, n_final=2)
    sample = (f"{n} WHOLE candidates ({per_step} per step) at the mid-quantiles of {n} equal-count cost strata of the same "
              f"{args.pop}-genotype population (fwd MACs/sample {min(macs) / 1e6:.1f}-{max(macs) / 1e6:.1f} M), each trained "
              f"{min(epochs)}-{max(epochs)} epochs (cap {args.epoch_cap}) and scored, evaluated serially as "
              f"nsga_penalty.py:426 does, {total:.1f} s; torch-CPU fp32 restatement of evaluate_individual "
              f"(oracle/cnn_ref.py; TensorFlow/Keras not installable), {cores} threads; warm-up steps use the two "
              "cheapest strata")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "seconds_per_candidate": [round(t, 3) for t in times],
                         "cpu_tflops_analytic": flops / total / 1e12},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args, hps) -> dict:
    """Our arm, N = 1, rank 0: the same port on a bounded stratified sample of whole candidates (about 10-30 s)."""
    cores = host_threads()
    import torch
    torch.set_num_threads(cores)
    data = cpu_features()
    picks = stratified_sample(hps, args.cpu_candidates)
    cpu_evaluate_candidate(hps[stratified_sample(hps, 8)[0]], data, 999, 1)          # thread pool / oneDNN warm-up
    times = []
    t0 = time.perf_counter()
    for i in picks:
        dt, _ = cpu_evaluate_candidate(hps[i], data, i, args.epoch_cap)
        times.append(dt)
    total = time.perf_counter() - t0
    return {"value": len(picks) / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(picks)} whole candidates at the mid-quantiles of {len(picks)} cost strata of the population, epoch "
                      f"cap {args.epoch_cap}, serial, {total:.1f} s; torch-CPU fp32 restatement (oracle/cnn_ref.py)",
            "seconds_per_candidate": [round(t, 3) for t in times]}


# ------------------------------------------------------------------------------------------------ GPU side
def device_features(dev):
    """Synthetic clips -> MFCC -> StandardScaler(fit on train) entirely on the device; returns CUDA tensors + labels."""
    import numpy as np
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccFrontEnd, prepare_dataset_device
    wave, labels = synth.make_clips(N_TRAIN + N_VAL, N_CLASSES, seed=1234)
    w = torch.from_numpy(wave).to(dev)
    fe = MfccFrontEnd()
    x_train, x_val = prepare_dataset_device(fe, [w[:N_TRAIN], w[N_TRAIN:]], policy="fit_train")
    torch.cuda.synchronize()
    fe.close()
    return x_train, labels[:N_TRAIN].astype(np.int32), x_val, labels[N_TRAIN:].astype(np.int32)


def train_config(args):
    from cmoop_audio_processing_b200.problem import TrainConfig
    return TrainConfig(variant=VARIANT, epochs=args.epoch_cap, patience=5, restore_best_weights=True, acc_from="evaluate",
                       fpr_mode="filtered", precision="bf16")


def birdclef_extra(args, rank, world, dev, barrier) -> dict:
    """BASELINE configs[3]: BirdCLEF-shaped candidates (sa_nsga_penalty.py:42-63,102,141: 128 x 313 log-mel maps of 5 s
    32 kHz clips, 397 classes), population 64 sharded over the ranks like the headline; one epoch over a bounded synthetic
    split (the feature maps are 20x the GSC ones).  step = one compute_objectives_and_constraints call."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from cmoop_audio_processing_b200 import _lib
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig, forward_macs
    lib = _lib.load()
    bh, bw, bc, pop, n_tr, n_va = 128, 313, 397, args.birdclef_pop, args.birdclef_train, args.birdclef_val
    gen = torch.Generator(device=dev).manual_seed(7)
    x_tr = torch.randn((n_tr, bh, bw, 1), generator=gen, device=dev)
    x_va = torch.randn((n_va, bh, bw, 1), generator=gen, device=dev)
    rng = np.random.default_rng(7)
    y_tr, y_va = rng.integers(0, bc, n_tr).astype(np.int32), rng.integers(0, bc, n_va).astype(np.int32)
    hps = make_population(pop, seed=3)
    cfg = TrainConfig(variant=VARIANT, epochs=1, patience=5, restore_best_weights=True, acc_from="evaluate",
                      fpr_mode="filtered", precision="bf16")
    prob = FitnessProblem.sa_nsga_local(x_tr, y_tr, x_va, y_va, classes=bc, config=cfg)
    prob.compute_objectives_and_constraints(hps)            # warm-up: arena + kernels
    barrier()
    t0 = time.perf_counter()
    steps = 2
    dev_ms = 0.0
    for _ in range(steps):
        recs = prob.compute_objectives_and_constraints(hps)
        dev_ms += float(lib.cmoop_cnn_last_device_ms())
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, dev_ms * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, busy = float(t[0]), float(t[1])
    assert len(recs) == pop and all(np.isfinite(r["objs"]).all() for r in recs)
    flops = sum(6.0 * forward_macs(hp, bh, bw, bc, VARIANT) * n_tr + 4.0 * forward_macs(hp, bh, bw, bc, VARIANT) * n_va
                for hp in hps)
    prob.data.close()
    return {"metric": METRIC, "unit": UNIT, "value": pop * steps / dt, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "population": pop, "map": [bh, bw], "classes": bc, "n_train": n_tr, "n_val": n_va, "epochs": 1,
            "tflops_analytic_per_gpu": flops * steps / busy / 1e12 / world, "scaling": "strong",
            "workload": "BASELINE configs[3]: BirdCLEF-shaped 128x313 maps, 397 classes, variant B, bf16, one epoch, "
                        f"{pop} candidates LPT-sharded over {world} GPU(s), all-gather of the objective rows inside the timed region"}


def mfcc_extra(args, rank, world, dev, barrier) -> dict:
    """BASELINE metric, second clause: MFCC clips/s vs the HBM roofline (configs[1], 65 536 clips per GPU, weak)."""
    import torch
    import torch.distributed as dist
    from cmoop_audio_processing_b200 import _lib
    from cmoop_audio_processing_b200.features import MfccFrontEnd
    lib = _lib.load()
    clips, steps = args.mfcc_clips, args.mfcc_steps
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    wave = torch.rand((clips, N_SAMPLES), generator=gen, device=dev, dtype=torch.float32).mul_(2).sub_(1)
    out = torch.empty((clips, N_FRAMES, N_OUT), device=dev, dtype=torch.float32)
    fe = MfccFrontEnd()
    for _ in range(3):
        fe(wave, out=out)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s, e in ev:
        s.record()
        fe(wave, out=out)                                   # launched on torch's current stream (passed to the C ABI)
        e.record()
    t1.record()
    barrier()
    total_ms = t0.elapsed_time(t1)
    kernel_ms = [s.elapsed_time(e) for s, e in ev]
    # host-buffer C-ABI call on 16-bit PCM (what a wav file holds) and on fp32, pinned host memory, H2D + D2H inside
    e2e_steps = 3
    h_out = torch.empty((clips, N_FRAMES, N_OUT), dtype=torch.float32, pin_memory=True)
    h_pcm = torch.empty((clips, N_SAMPLES), dtype=torch.int16, pin_memory=True)
    h_pcm.copy_((wave * 32767.0).round().to(torch.int16))
    h_wave = torch.empty((clips, N_SAMPLES), dtype=torch.float32, pin_memory=True)
    h_wave.copy_(wave)
    torch.cuda.synchronize()
    res = {}
    for key, src in (("e2e_fp32", h_wave.numpy()), ("e2e_int16", h_pcm.numpy())):
        fe(src, out=h_out.numpy())
        barrier()
        t = time.perf_counter()
        for _ in range(e2e_steps):
            fe(src, out=h_out.numpy())
        torch.cuda.synchronize()
        res[key] = time.perf_counter() - t
    stats = torch.tensor([total_ms, res["e2e_fp32"], res["e2e_int16"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, e32, e16 = (float(v) for v in stats)
    fe.close()
    del wave, out, h_out, h_pcm, h_wave
    torch.cuda.empty_cache()
    peaks = measured_peaks()
    kavg = sum(kernel_ms) / len(kernel_ms)
    achieved = BYTES_PER_CLIP * clips / (kavg * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline.json")) as fh:
            traffic = json.load(fh)["mfcc_kernel"]["dram_bytes_per_clip"] * clips
    except Exception:
        pass
    return {
        "metric": "mfcc_clips_per_sec", "unit": "clips/s", "scaling": "weak", "clips_per_gpu": clips, "steps": steps,
        "value": world * clips * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "dtype": "f32",
        "roofline": {"bound": "hbm", "kernel": "mfcc_pair_kernel<1>", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                     "algorithmic_bytes_per_launch": BYTES_PER_CLIP * clips, "kernel_ms_avg": kavg},
        "e2e": {"value": world * clips * e2e_steps / e32, "unit": "clips/s", "h2d_bytes_per_step": clips * N_SAMPLES * 4,
                "d2h_bytes_per_step": clips * N_FRAMES * N_OUT * 4, "api": "cmoop_mfcc_fwd_host, pinned fp32 host buffers"},
        "e2e_int16": {"value": world * clips * e2e_steps / e16, "unit": "clips/s", "h2d_bytes_per_step": clips * N_SAMPLES * 2,
                      "d2h_bytes_per_step": clips * N_FRAMES * N_OUT * 4,
                      "api": "cmoop_mfcc_fwd_host_i16 (16-bit PCM host buffers, widened on the device)"},
    }


def latency_extra(dev) -> dict:
    """us per call of the latency-bound kernels (device-resident, CUDA events on the launching stream; and through the
    host ABI), next to the CPU port of the reference function (oracle/) timed in the same run."""
    import ctypes as C

    import numpy as np
    import torch
    from cmoop_audio_processing_b200 import _lib, nsga, quality
    from cmoop_audio_processing_b200.surrogate import DeviceGPGroup
    from oracle import gp_ref, hv_ref, nsga_ref
    lib = _lib.load()
    rng = np.random.default_rng(3)
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = {}

    def dev_us(fn, reps=50):
        for _ in range(5):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / reps

    def host_us(fn, reps=20):
        fn()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t) * 1e6 / reps

    for n in (128, 512):                                   # parents / merged population of pop 64 and 256, M = 3
        objs = rng.random((n, 3))
        cv = rng.random(n) * (rng.random(n) < 0.5)
        recs = [{"hparams": {}, "objs": objs[i].tolist(), "CV": float(cv[i])} for i in range(n)]
        d_objs, d_cv = torch.from_numpy(objs).to(dev), torch.from_numpy(cv).to(dev)
        d_rank = torch.empty(n, dtype=torch.int32, device=dev)
        d_order = torch.empty(n, dtype=torch.int32, device=dev)
        d_foff = torch.empty(n + 1, dtype=torch.int32, device=dev)
        d_nf = torch.empty(1, dtype=torch.int32, device=dev)
        d_crowd = torch.empty(n, dtype=torch.float64, device=dev)

        def launch():
            _lib.check(lib.cmoop_nds_crowding_dev(C.c_void_p(d_objs.data_ptr()), C.c_void_p(d_cv.data_ptr()), n, 3, 1, 7.5,
                                                  1e-6, 0, C.c_void_p(d_rank.data_ptr()), C.c_void_p(d_order.data_ptr()),
                                                  C.c_void_p(d_foff.data_ptr()), C.c_void_p(d_nf.data_ptr()),
                                                  C.c_void_p(d_crowd.data_ptr()), None, 0, C.c_void_p(stream)),
                       "cmoop_nds_crowding_dev")
        t = time.perf_counter()
        fronts = nsga_ref.fast_non_dominated_sort(recs, 7.5)
        for f in fronts:
            nsga_ref.crowding_distance(f, recs)
        cpu = (time.perf_counter() - t) * 1e6
        assert nsga.fast_non_dominated_sort(recs, 7.5) == fronts
        out[f"nds_crowding_n{n}_m3"] = {"device_us": dev_us(launch), "host_abi_us": host_us(
            lambda: nsga.nds_crowding_arrays(objs, cv, 7.5)), "cpu_port_us": cpu,
            "cpu_port": "oracle/nsga_ref.py fast_non_dominated_sort + crowding_distance (sa_nsga_local.py:247-277)"}
    pts = rng.random((256, 3))
    ref = np.array([1.1, 1.1, 1.1])
    d_pts, d_ref = torch.from_numpy(pts).to(dev), torch.from_numpy(ref).to(dev)
    d_hv = torch.empty(1, dtype=torch.float64, device=dev)
    ws_bytes = int(lib.cmoop_hypervolume_workspace_bytes(256))
    d_ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    t = time.perf_counter()
    hv_cpu = hv_ref.hypervolume(pts, ref)
    cpu = (time.perf_counter() - t) * 1e6
    assert quality.hypervolume(pts, ref) == hv_cpu
    out["hypervolume_n256_m3"] = {
        "device_us": dev_us(lambda: _lib.check(lib.cmoop_hypervolume_dev(
            C.c_void_p(d_pts.data_ptr()), 256, 3, C.c_void_p(d_ref.data_ptr()), C.c_void_p(d_hv.data_ptr()),
            C.c_void_p(d_ws.data_ptr()), ws_bytes, C.c_void_p(stream)), "cmoop_hypervolume_dev")),
        "host_abi_us": host_us(lambda: quality.hypervolume(pts, ref)), "cpu_port_us": cpu,
        "cpu_port": "oracle/hv_ref.py (pygmo hv3d restated; compare.ipynb l.230-231)"}
    n = 288                                                # the whole genotype space as training set and as queries
    x = rng.random((n, 8)) * 4
    specs = []
    for j in range(4):
        k = 1.3 * gp_ref.matern(x, x, 1.7, 1.5) + 0.1 * np.eye(n)
        low = np.linalg.cholesky(k)
        y = rng.standard_normal(n)
        specs.append(dict(x_train=x, alpha=np.linalg.solve(k, y), chol_lower=low, amplitude=1.3, length_scale=1.7, nu=1.5,
                          noise=0.1, y_scale=1.0 + j, y_shift=0.1 * j))
    grp = DeviceGPGroup(specs)
    d_x = torch.from_numpy(x).to(dev)
    d_mu = torch.empty((4, n), dtype=torch.float64, device=dev)
    d_sd = torch.empty((4, n), dtype=torch.float64, device=dev)
    t = time.perf_counter()
    for s in specs:
        gp_ref.posterior(x, s["x_train"], s["alpha"], s["chol_lower"], amplitude=1.3, length_scale=1.7, nu=1.5, noise=0.1,
                         y_scale=s["y_scale"], y_shift=s["y_shift"])
    cpu = (time.perf_counter() - t) * 1e6
    out["gp_predict_q288_n288_x4"] = {
        "device_us": dev_us(lambda: _lib.check(lib.cmoop_gp_predict_dev(
            grp._handle, C.c_void_p(d_x.data_ptr()), n, C.c_void_p(d_mu.data_ptr()), C.c_void_p(d_sd.data_ptr()),
            C.c_void_p(stream)), "cmoop_gp_predict_dev")),
        "host_abi_us": host_us(lambda: grp.predict(x)), "cpu_port_us": cpu,
        "cpu_port": "oracle/gp_ref.py posterior (NumPy/LAPACK restatement of sklearn _gpr.py:446-499), 4 models"}
    grp.close()
    return out


def run_ours(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        use_all_host_threads()                              # the cpu_baseline leg and the host GP fit use every thread
    import warnings

    import numpy as np
    import torch
    import torch.distributed as dist
    warnings.filterwarnings("ignore")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from cmoop_audio_processing_b200 import _lib, drivers, quality
    from cmoop_audio_processing_b200.problem import DeviceDataset, FitnessProblem
    lib = _lib.load()
    _lib.check(lib.cmoop_set_device(local_rank), "cmoop_set_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    sampler = ClockSampler(local_rank)
    hps = make_population(args.pop)
    x_train, y_train, x_val, y_val = device_features(dev)
    cfg = train_config(args)
    prob = FitnessProblem.sa_nsga_local(x_train, y_train, x_val, y_val, classes=N_CLASSES, config=cfg)

    # ---- value: K calls of compute_objectives_and_constraints over the whole population (sharded + all-gather inside)
    for _ in range(args.warmup):
        prob.compute_objectives_and_constraints(hps)
    barrier()
    launches0 = lib.cmoop_launch_count()
    sampler.start()
    dev_ms, epochs_run, busy = 0.0, None, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        recs = prob.compute_objectives_and_constraints(hps)
        dev_ms += float(lib.cmoop_cnn_last_device_ms())
        epochs_run = prob.last_details[:, 3].copy()
    barrier()
    total_s = time.perf_counter() - t0
    sampler.stop()
    launches = int(lib.cmoop_launch_count() - launches0)
    assert len(recs) == args.pop and all(np.isfinite(r["objs"]).all() for r in recs)
    flops_step = analytic_flops(hps, epochs_run)
    local_busy_s = dev_ms * 1e-3
    total_s, dev_s_max = reduce_max([total_s, local_busy_s])
    busy_all = [local_busy_s]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local_busy_s)
        busy_all = [float(b) for b in gathered]

    # ---- kernel-family breakdown of one more (untimed) step: CUDA-event pairs around every launch of the engine
    lib.cmoop_profile_enable(1)
    prob.compute_objectives_and_constraints(hps)
    lib.cmoop_profile_enable(0)
    table = _lib.profile_table()
    prof_ms = sum(v["ms"] for v in table.values())
    breakdown = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / prof_ms, 4),
                     **({"tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2)} if v["flops"] > 0 else {})}
                 for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- e2e: full SA-NSGA-II generations from host records / host feature arrays to host records
    h_xt = torch.empty(x_train.shape, dtype=torch.float32, pin_memory=True).copy_(x_train)
    h_xv = torch.empty(x_val.shape, dtype=torch.float32, pin_memory=True).copy_(x_val)
    torch.cuda.synchronize()
    np_xt, np_xv = h_xt.numpy(), h_xv.numpy()
    # timed generations: a fixed count (not tied to --steps): one generation in ~8 hits an optimiser start of the surrogate
    # fit that L-BFGS-B walks for seconds, so a short run is dominated by whether it caught one
    gens_w, gens_k = min(args.warmup, 2), max(1, args.e2e_generations)
    e2e_prob = FitnessProblem.sa_nsga_local(prob.data, None, None, None, classes=N_CLASSES, config=cfg, seed=10_000)
    ops = drivers.default_ops(e2e_prob)
    ops.SurrogateManager = lambda: __import__("cmoop_audio_processing_b200.surrogate", fromlist=["x"]).SurrogateManager(
        fit_backend=args.gp_fit_backend)
    staged = {"data": None}

    def evaluate_from_host(population):
        # the generation's inputs are HOST arrays: features are re-staged host->device for every evaluation batch
        if staged["data"] is not None:
            staged["data"].close()
        staged["data"] = DeviceDataset(np_xt, y_train, np_xv, y_val)
        e2e_prob.data = staged["data"]
        return e2e_prob.compute_objectives_and_constraints(population)
    ops.compute_objectives_and_constraints = evaluate_from_host
    archive = {"front": None}

    def on_generation(gen, pop_data):
        # HV (ref = max + 1e-3, compare.ipynb l.215-220) / IGD / Spread of the generation against the running union front
        feas = np.array([r["objs"] for r in pop_data if r["CV"] == 0], dtype=np.float64).reshape(-1, 3)
        if len(feas) == 0:
            return {"hv": 0.0, "n": 0}
        front = feas[quality.nondominated_mask(feas)]
        union = front if archive["front"] is None else np.vstack([archive["front"], front])
        archive["front"] = union[quality.nondominated_mask(union)]
        out = {"hv": quality.hypervolume(front, quality.reference_point(feas)), "n": int(len(front))}
        out.update(quality.front_metrics(front, archive["front"]))
        return out

    random.seed(0)
    np.random.seed(0)
    copies = []

    def mark_copies(gen, pop_data):
        out = on_generation(gen, pop_data)
        copies.append(_lib.copy_bytes())
        return out
    barrier()
    sampler.start()
    _front, _history, timings = drivers.sa_nsga2(args.pop, gens_w + gens_k, INFILL, ops, local_search=True,
                                                 on_generation=mark_copies)
    barrier()
    sampler.stop()
    timed = timings[gens_w:]
    gen_s = sum(t["seconds"] for t in timed)
    gen_evals = sum(t["true_evals"] for t in timed)
    (gen_s,) = reduce_max([gen_s])
    c_prev = copies[gens_w - 1] if gens_w > 0 else None
    c_last = copies[-1]
    if c_prev is None:                                      # no warm-up generation: count from the first timed one on
        c_prev, n_counted = copies[0], max(1, gens_k - 1)
    else:
        n_counted = gens_k
    h2d_step = (c_last[0] - c_prev[0]) / n_counted
    d2h_step = (c_last[1] - c_prev[1]) / n_counted

    peaks = measured_peaks()
    line = None
    if rank == 0:
        achieved = flops_step * args.steps / dev_s_max / 1e12 / world      # per GPU: the peak is one GPU's
        contraction = {k: v for k, v in table.items() if v["flops"] > 0}
        dom = max(contraction, key=lambda k: contraction[k]["ms"]) if contraction else None
        line = {
            "metric": METRIC, "value": args.pop * args.steps / total_s, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": gen_evals / gen_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "steps": gens_k, "true_evals_per_generation": gen_evals // gens_k,
                    "generations_per_sec": gens_k / gen_s, "seconds_per_generation": [round(t["seconds"], 3) for t in timed],
                    "eval_seconds": [round(t["eval_seconds"], 3) for t in timed],
                    "surrogate_update_seconds": [round(t["update_seconds"], 3) for t in timed],
                    "indicators_last": timed[-1]["indicators"], "gp_fit_backend": args.gp_fit_backend,
                    "api": "drivers.sa_nsga2 (ablation_study/sa_nsga_local.py:436-554) over FitnessProblem / SurrogateManager / "
                           "perform_local_search / select_infill_points / fast_non_dominated_sort / crowding_distance / quality.*; "
                           "host feature arrays re-staged host->device every generation (cmoop_cnn_dataset_create_host)"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["bf16_sustained"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"], "traffic": None,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained (the step is seconds long)",
                         "algorithmic_flops_per_step": flops_step, "device_seconds_per_step": dev_s_max / args.steps,
                         "achieved_all_gpus": achieved * world,
                         "note": "achieved = analytic training+scoring FLOPs of the whole step / device time of the step "
                                 "(max over ranks) / n_gpus, i.e. per GPU against one GPU's peak; dominant_kernel = the contraction kernel family "
                                 "with the largest share, its own 2*M*K*N flops / its own event-timed duration",
                         "dominant_kernel": None if dom is None else {
                             "name": dom, "launches": table[dom]["launches"], "ms": table[dom]["ms"],
                             "achieved": table[dom]["flops"] / (table[dom]["ms"] * 1e-3) / 1e12,
                             "frac": table[dom]["flops"] / (table[dom]["ms"] * 1e-3) / 1e12 / peaks["bf16_sustained"],
                             "share_of_step": table[dom]["ms"] / prof_ms}},
            "kernel_breakdown": breakdown,
            "rank_busy_seconds_per_step": [round(b / args.steps, 4) for b in busy_all],
            "epochs_run": {"min": int(epochs_run.min()), "max": int(epochs_run.max()), "mean": float(epochs_run.mean())},
            "clocks": sampler.summary(),
        }
    if staged["data"] is not None:
        staged["data"].close()
    del h_xt, h_xv
    # ---- extras
    mf = None
    if not args.no_mfcc:
        try:
            mf = mfcc_extra(args, rank, world, dev, barrier)
        except Exception as exc:                            # the headline line must still be printed
            mf = {"error": repr(exc)}
    bc = None
    if not args.no_birdclef:
        try:
            bc = birdclef_extra(args, rank, world, dev, barrier)
        except Exception as exc:
            bc = {"error": repr(exc)}
    if rank == 0:
        if bc is not None:
            line["birdclef"] = bc
        if mf is not None:
            line["mfcc"] = mf
        if world == 1 and not args.no_latency:
            try:
                line["latency"] = latency_extra(dev)
            except Exception as exc:
                line["latency"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args, hps)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--pop", type=int, default=POP, help="population (BASELINE configs[4]: 256)")
    ap.add_argument("--epoch-cap", type=int, default=4, help="epoch ceiling of every candidate (reference: 300, ES patience 5)")
    ap.add_argument("--e2e-generations", type=int, default=8, help="timed SA-NSGA-II generations of the e2e leg")
    ap.add_argument("--gp-fit-backend", choices=["device", "host"], default="device")
    ap.add_argument("--cpu-candidates", type=int, default=4, help="whole candidates of the N=1 cpu_baseline sample")
    ap.add_argument("--mfcc-clips", type=int, default=65536, help="clips per GPU of the MFCC extra (BASELINE configs[1])")
    ap.add_argument("--mfcc-steps", type=int, default=20)
    ap.add_argument("--ref-min-candidates", type=int, default=8, help="--impl reference: whole candidates timed at least")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mfcc", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-birdclef", action="store_true")
    ap.add_argument("--birdclef-pop", type=int, default=64, help="population of the BirdCLEF extra (BASELINE configs[3]: 64)")
    ap.add_argument("--birdclef-train", type=int, default=512)
    ap.add_argument("--birdclef-val", type=int, default=128)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                            # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
