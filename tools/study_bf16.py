"""bf16 (tcgen05, bf16-stored activations) vs fp32 (exact SIMT path) vs the fp64 oracle at BASELINE configs[0] size:
12 x 256 training / 12 x 64 validation synthetic GSC-shaped clips -> MFCC -> StandardScaler(fit on train); 16 genotypes
(six with filters = 64) x 3 seeds, epoch cap 6, EarlyStopping patience 5, restore best, variant B (sa_nsga_local policy:
accuracy from evaluate, filtered FPR).  Prints / stores the distribution of |d accuracy| and |d FPR| between the two
precisions, the agreement of the non-dominated fronts, and -- for two small genotypes -- the fp64 torch-CPU oracle trained
on the same initial parameters, shuffles and dropout masks.

    python tools/study_bf16.py [out.json] [epoch_cap] [oracle_genotypes] [snr_db]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

GENOTYPES = [
    dict(filters=16, kernel_size=3, use_bn=False, residual_blocks=1, fc_layers=1, use_dropout=False),
    dict(filters=16, kernel_size=5, use_bn=True, residual_blocks=2, fc_layers=2, use_dropout=True),
    dict(filters=16, kernel_size=3, use_bn=True, residual_blocks=3, fc_layers=3, use_dropout=True),
    dict(filters=16, kernel_size=5, use_bn=False, residual_blocks=3, fc_layers=4, use_dropout=False),
    dict(filters=32, kernel_size=5, use_bn=False, residual_blocks=1, fc_layers=4, use_dropout=False),
    dict(filters=32, kernel_size=3, use_bn=True, residual_blocks=3, fc_layers=1, use_dropout=False),
    dict(filters=32, kernel_size=3, use_bn=False, residual_blocks=2, fc_layers=2, use_dropout=True),
    dict(filters=32, kernel_size=5, use_bn=True, residual_blocks=2, fc_layers=3, use_dropout=True),
    dict(filters=32, kernel_size=3, use_bn=True, residual_blocks=1, fc_layers=2, use_dropout=False),
    dict(filters=32, kernel_size=5, use_bn=False, residual_blocks=3, fc_layers=1, use_dropout=True),
    dict(filters=64, kernel_size=3, use_bn=False, residual_blocks=1, fc_layers=1, use_dropout=False),
    dict(filters=64, kernel_size=3, use_bn=True, residual_blocks=2, fc_layers=2, use_dropout=True),
    dict(filters=64, kernel_size=5, use_bn=True, residual_blocks=1, fc_layers=3, use_dropout=False),
    dict(filters=64, kernel_size=3, use_bn=True, residual_blocks=3, fc_layers=4, use_dropout=True),
    dict(filters=64, kernel_size=5, use_bn=False, residual_blocks=2, fc_layers=1, use_dropout=True),
    dict(filters=64, kernel_size=5, use_bn=True, residual_blocks=3, fc_layers=2, use_dropout=False),
]
SEEDS = (101, 202, 303)
N_CLASSES = 12


def config0_data(snr_db=10.0):
    """BASELINE configs[0] data: synthetic GSC-shaped clips through the CUDA front-end.  At the
    survey's 10 dB SNR every genotype reaches 100 % validation accuracy within a few epochs in either precision (the study
    is then vacuous), so the study also runs at a much lower SNR, where accuracies spread over 0.2 .. 0.9."""
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccFrontEnd, prepare_dataset_device
    n_tr, n_va = 12 * 256, 12 * 64
    wave, labels = synth.make_clips(n_tr + n_va, N_CLASSES, seed=1234, snr_db=snr_db)
    w = torch.from_numpy(wave).cuda()
    fe = MfccFrontEnd()
    xt, xv = prepare_dataset_device(fe, [w[:n_tr], w[n_tr:]], policy="fit_train")
    torch.cuda.synchronize()
    return (xt.cpu().numpy()[..., None], labels[:n_tr].astype(np.int64), xv.cpu().numpy()[..., None], labels[n_tr:].astype(np.int64))


def robust_dominance_agreement(a, b, tol):
    """a, b: [n, 3] objective rows (-acc, size, fpr) of the same candidates under two precisions.  A pair (i, j) is
    ROBUST in `a` when i dominates j with every deciding gap larger than tol; returns (#robust pairs, #of them on which `b`
    disagrees, i.e. i does not dominate j in b)."""
    n, robust, broken = len(a), 0, 0
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            d = a[j] - a[i]                                          # >= 0 where i is at least as good
            if (d >= 0).all() and (d > 0).any() and all(dk == 0 or dk > tk for dk, tk in zip(d, tol)):
                robust += 1
                e = b[j] - b[i]
                broken += not ((e >= 0).all() and (e > 0).any())
    return robust, broken


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n_oracle = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    snr_db = float(sys.argv[4]) if len(sys.argv) > 4 else 10.0
    from cmoop_audio_processing_b200 import nsga
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    xt, yt, xv, yv = config0_data(snr_db)
    rows = {}
    t_run = {}
    for prec in ("bf16", "fp32"):
        cfg = TrainConfig(variant="B", epochs=epochs, patience=5, restore_best_weights=True, acc_from="evaluate",
                          fpr_mode="filtered", precision=prec)
        prob = FitnessProblem.sa_nsga_local(xt, yt, xv, yv, classes=N_CLASSES, config=cfg)
        hps = [hp for _ in SEEDS for hp in GENOTYPES]
        seeds = [s + 7 * i for s in SEEDS for i in range(len(GENOTYPES))]
        t0 = time.perf_counter()
        rows[prec], _ = prob.train_eval(hps, seeds)
        t_run[prec] = time.perf_counter() - t0
        prob.data.close()
    acc = {p: rows[p][:, 0].reshape(len(SEEDS), -1) for p in rows}
    fpr = {p: rows[p][:, 2].reshape(len(SEEDS), -1) for p in rows}
    d_acc, d_fpr = np.abs(acc["bf16"] - acc["fp32"]).ravel(), np.abs(fpr["bf16"] - fpr["fp32"]).ravel()
    pct = lambda v: {q: float(np.percentile(v, q)) for q in (50, 90, 95, 100)}      # noqa: E731
    # seed-to-seed spread of the SAME precision: the noise floor the precision gap should be compared with
    spread_acc = float(np.abs(acc["fp32"] - acc["fp32"].mean(axis=0)).max())
    fronts = []
    for s in range(len(SEEDS)):
        recs = {p: [{"hparams": hp, "objs": [-acc[p][s, i], rows[p][s * len(GENOTYPES) + i, 1], fpr[p][s, i]], "CV": 0.0}
                    for i, hp in enumerate(GENOTYPES)] for p in rows}
        f = {p: nsga.fast_non_dominated_sort(recs[p], 1.0) for p in rows}
        a = np.array([r["objs"] for r in recs["fp32"]])
        b = np.array([r["objs"] for r in recs["bf16"]])
        fronts.append({"seed": SEEDS[s], "front0_fp32": sorted(f["fp32"][0]), "front0_bf16": sorted(f["bf16"][0]),
                       "robust_pairs_and_disagreements": robust_dominance_agreement(a, b, (0.03, 0.0, 0.01))})
    report = {"config": f"configs[0] size: 3072 train / 768 val clips at {snr_db} dB SNR, 16 genotypes x 3 seeds, epoch cap {epochs}, variant B",
              "seconds": t_run, "abs_d_accuracy_percentiles": pct(d_acc), "abs_d_fpr_percentiles": pct(d_fpr),
              "max_abs_d_best_val_loss": float(np.abs(rows["bf16"][:, 5] - rows["fp32"][:, 5]).max()),
              "fp32_seed_to_seed_max_abs_accuracy_spread": spread_acc,
              "accuracy_fp32": acc["fp32"].round(4).tolist(), "accuracy_bf16": acc["bf16"].round(4).tolist(),
              "fpr_fp32": fpr["fp32"].round(5).tolist(), "fpr_bf16": fpr["bf16"].round(5).tolist(), "fronts": fronts}
    if n_oracle > 0:
        import torch
        from oracle import cnn_ref
        from test_gpu_cnn import unflatten
        torch.set_num_threads(os.cpu_count() or 1)
        oracle = []
        cfg = TrainConfig(variant="B", epochs=2, patience=5, restore_best_weights=True, acc_from="evaluate", fpr_mode="filtered")
        for hp in GENOTYPES[:n_oracle]:
            entry = {"hparams": hp}
            for prec in ("fp32", "bf16"):
                cfg.precision = prec
                prob = FitnessProblem.sa_nsga_local(xt, yt, xv, yv, classes=N_CLASSES, config=cfg)
                out, _ = prob.train_eval([hp], [SEEDS[0]])
                entry[prec] = {"acc": float(out[0, 0]), "fpr": float(out[0, 2]), "val_loss": float(out[0, 5])}
                if prec == "fp32":
                    init = prob.debug_init_params(hp, SEEDS[0])
                    perms = [prob.debug_permutation(SEEDS[0], e) for e in range(2)]
                prob.data.close()
            ref = cnn_ref.evaluate_individual(hp, (xt, yt, xv, yv), unflatten(init, hp, "B"), perms, n_classes=N_CLASSES, variant="B",
                                              seed=SEEDS[0], epochs=2, patience=5, restore_best_weights=True, acc_from="evaluate",
                                              fpr_mode="filtered", dtype=torch.float64)
            entry["fp64_oracle"] = {"acc": ref["acc"], "fpr": ref["fpr"], "val_loss": min(ref["history"]["val_loss"])}
            oracle.append(entry)
        report["oracle_2_epochs"] = oracle
    print(json.dumps(report, indent=1))
    if out_path:
        with open(out_path, "w") as fh:
            json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
