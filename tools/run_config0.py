"""BASELINE configs[0]: nsga_penalty.py KWS CNN search, population 8 x 2 generations, synthetic GSC-shaped 1 s 16 kHz clips,
12 classes -- on the GPU through the drop-in, with every true evaluation repeated by the torch-CPU oracle (see
tests/test_gpu_config0.py for the method).

    python tools/run_config0.py [per_class_train=256] [per_class_val=64] [epoch_cap=10] [out.json]
"""
import json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np


def run(per_class_train=256, per_class_val=64, epoch_cap=10, pop=8, gens=2, oracle_dtype="float64", log=None):
    import torch
    from cmoop_audio_processing_b200 import drivers, nsga, synth
    from cmoop_audio_processing_b200.features import MfccFrontEnd, prepare_dataset_device
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    from oracle import cnn_ref, nsga_ref
    from test_gpu_cnn import unflatten
    n_tr, n_va, classes = 12 * per_class_train, 12 * per_class_val, 12
    wave, labels = synth.make_clips(n_tr + n_va, classes, seed=1234)
    w = torch.from_numpy(wave).cuda()
    fe = MfccFrontEnd()
    xt_d, xv_d = prepare_dataset_device(fe, [w[:n_tr], w[n_tr:]], policy="fit_each")      # nsga_penalty.py:111,124,137
    torch.cuda.synchronize()
    xt, xv = xt_d.cpu().numpy()[..., None], xv_d.cpu().numpy()[..., None]
    yt, yv = labels[:n_tr].astype(np.int64), labels[n_tr:].astype(np.int64)
    cfg = TrainConfig(variant="A", epochs=epoch_cap, patience=5, restore_best_weights=False, acc_from="history",
                      y_true_mode="argmax_quirk", precision="fp32")
    prob = FitnessProblem.nsga_penalty(xt, yt, xv, yv, classes=classes, config=cfg)
    evaluated = []
    inner = prob.compute_objectives_and_constraints

    def recording(population):
        first = prob.seed + prob.evaluations
        recs = inner(population)
        for i, (hp, rec) in enumerate(zip(population, recs)):
            evaluated.append((dict(hp), first + i, rec, prob.last_details[i].copy()))
        return recs
    ops = drivers.default_ops(prob, surrogate=False, script="nsga_penalty")
    ops.compute_objectives_and_constraints = recording
    random.seed(0)
    t0 = time.perf_counter()
    pareto, history, timings = drivers.nsga2(pop, gens, ops)
    gpu_s = time.perf_counter() - t0
    # ---- every true evaluation again on the CPU oracle (same init, shuffles, dropout stream)
    torch.set_num_threads(os.cpu_count() or 1)
    dt = getattr(torch, oracle_dtype)
    d_acc, d_fpr, d_loss, epochs_equal, size_exact = [], [], [], True, True
    oracle_recs = []
    t0 = time.perf_counter()
    for hp, seed, rec, det in evaluated:
        init = prob.debug_init_params(hp, seed)
        perms = [prob.debug_permutation(seed, e) for e in range(epoch_cap)]
        ref = cnn_ref.evaluate_individual(hp, (xt, yt, xv, yv), unflatten(init, hp, "A"), perms, n_classes=classes, variant="A",
                                          seed=seed, epochs=epoch_cap, patience=5, restore_best_weights=False,
                                          acc_from="history", y_true_mode="argmax_quirk", dtype=dt)
        d_acc.append(abs(-rec["objs"][0] - ref["acc"]))
        d_fpr.append(abs(rec["objs"][2] - ref["fpr"]))
        d_loss.append(abs(det[4] - ref["history"]["val_loss"][-1]))
        epochs_equal = epochs_equal and int(det[3]) == ref["epochs_run"]
        size_exact = size_exact and rec["objs"][1] == ref["size_mb"]
        cv = nsga_ref.constraint_violation(ref["acc"], ref["size_mb"], ref["fpr"], 0.9, 2.5, 0.1)
        oracle_recs.append({"hparams": hp, "objs": [-ref["acc"], ref["size_mb"], ref["fpr"]], "CV": cv})
        if log:
            log(f"  {hp} seed {seed}: gpu acc {-rec['objs'][0]:.4f} fpr {rec['objs'][2]:.5f} | oracle acc {ref['acc']:.4f} fpr {ref['fpr']:.5f}")
    cpu_s = time.perf_counter() - t0
    # ---- ranks: CUDA NDS / crowding vs the oracle's on the SAME records (bit-exact), per generation and on all evaluations
    ranks_ok, crowd_ok = True, True
    gpu_recs = [rec for _, _, rec, _ in evaluated]
    for recs, lam in [(gpu_recs, 1.0), (gpu_recs, 50.0)] + [(h, nsga.get_lambda(g, gens)) for g, h in enumerate(history)]:
        f_gpu, f_ref = nsga.fast_non_dominated_sort(recs, lam), nsga_ref.fast_non_dominated_sort(recs, lam)
        ranks_ok = ranks_ok and f_gpu == f_ref
        for fr in f_ref:
            a = nsga.crowding_distance(fr, recs, crowd_mode=nsga.CROWD_RANGE_LT)
            b = nsga_ref.crowding_distance(fr, recs, skip_on_equal=False)
            crowd_ok = crowd_ok and all(a[i] == b[i] for i in fr)
    # ranks of the oracle-evaluated records vs the GPU-evaluated ones, wherever dominance is robust to the tolerance
    from study_bf16 import robust_dominance_agreement
    pen = lambda recs, lam: np.array([[f + lam * r["CV"] for f in r["objs"]] for r in recs])     # noqa: E731
    robust, broken = robust_dominance_agreement(pen(oracle_recs, 1.0), pen(gpu_recs, 1.0), (0.30, 0.0, 0.02))
    return {"config": f"configs[0]: pop {pop} x {gens} generations, variant A (nsga_penalty.py policy), {n_tr} train / {n_va} val "
                      f"clips, epoch cap {epoch_cap}, random.seed(0), fp32 exact path vs {oracle_dtype} oracle",
            "evaluations": len(evaluated), "gpu_seconds_whole_search": gpu_s, "cpu_oracle_seconds_for_the_same_evaluations": cpu_s,
            "max_abs_d_acc": float(max(d_acc)), "max_abs_d_fpr": float(max(d_fpr)), "max_abs_d_val_loss": float(max(d_loss)),
            "median_abs_d_acc": float(np.median(d_acc)), "frac_d_acc_within_0.05": float(np.mean(np.array(d_acc) <= 0.05)),
            "frac_d_fpr_within_0.01": float(np.mean(np.array(d_fpr) <= 0.01)), "abs_d_acc": [round(float(v), 4) for v in d_acc], "epochs_equal": bool(epochs_equal), "size_exact": bool(size_exact),
            "ranks_bit_exact": bool(ranks_ok), "crowding_bit_exact": bool(crowd_ok), "robust_pairs": int(robust),
            "robust_rank_disagreements": int(broken), "pareto_size": len(pareto),
            "final_population": [{"hparams": r["hparams"], "objs": r["objs"], "CV": r["CV"]} for r in history[-1]]}


if __name__ == "__main__":
    a = sys.argv[1:]
    rep = run(int(a[0]) if len(a) > 0 else 256, int(a[1]) if len(a) > 1 else 64, int(a[2]) if len(a) > 2 else 10, log=print)
    print(json.dumps(rep, indent=1, default=str))
    if len(a) > 3:
        with open(a[3], "w") as fh:
            json.dump(rep, fh, indent=1, default=str)
