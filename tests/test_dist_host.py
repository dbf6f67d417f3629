"""N>1 host logic on CPU: LPT candidate assignment and the gloo all-gather of objective rows (world_size 2)."""
import os

import numpy as np
import pytest

from cmoop_audio_processing_b200.dist import assign_lpt


def test_assign_lpt_balances_and_is_deterministic():
    costs = [100, 1, 1, 1, 50, 49, 2, 98]
    owner = assign_lpt(costs, 2)
    loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 2
    assert owner == assign_lpt(list(costs), 2)
    assert assign_lpt([], 4) == []
    assert sorted(set(assign_lpt([5] * 8, 8))) == list(range(8))


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig, compute_model_size_mb
    prob = object.__new__(FitnessProblem)
    prob.data = SimpleNamespace(height=49, width=40)
    prob.classes, prob.config, prob.seed, prob.evaluations = 12, TrainConfig(variant="B"), 5, 0
    prob.min_accuracy, prob.max_model_size, prob.max_fpr = 0.9, 2.5, 0.09
    prob.objectives = ("neg_acc", "size", "fpr")
    calls = []

    def fake_train_eval(hps, seeds=None, want_history=False):
        calls.append(list(seeds))
        out = np.zeros((len(hps), 6))
        for i, (hp, s) in enumerate(zip(hps, seeds)):
            out[i] = [0.5 + 0.001 * hp["filters"] + 1e-4 * s, compute_model_size_mb(hp, 12, "B"), 0.01 * hp["fc_layers"],
                      3, 1.0, 0.9]
        return out, None

    prob.train_eval = fake_train_eval
    pop = [dict(filters=f, kernel_size=k, use_bn=True, residual_blocks=r, fc_layers=fc, use_dropout=False)
           for f in (16, 64) for k in (3, 5) for r in (1, 3) for fc in (1, 4)][:13]
    recs = prob.compute_objectives_and_constraints(pop)
    np.save(os.path.join(tmp, f"r{rank}.npy"), np.array([r["objs"] + [r["CV"]] for r in recs]))
    np.save(os.path.join(tmp, f"n{rank}.npy"), np.array([len(c) for c in calls]))
    dist.destroy_process_group()


def test_sharded_evaluation_all_gathers_rows_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(a, b)                       # every rank holds the full, identically ordered result
    assert a.shape == (13, 4) and np.all(a[:, 1] > 0)
    n0, n1 = int(np.load(tmp_path / "n0.npy").sum()), int(np.load(tmp_path / "n1.npy").sum())
    assert n0 + n1 == 13 and n0 > 0 and n1 > 0               # disjoint shards cover the population
    # seeds follow the global candidate index, so results do not depend on the world size
    expect = [0.5 + 0.001 * (16 if i < 8 else 64) + 1e-4 * (5 + i) for i in range(13)]
    np.testing.assert_allclose(-a[:, 0], expect, rtol=0, atol=1e-12)


def _gp_worker(rank, world, port, tmp):
    import warnings

    import torch.distributed as dist
    warnings.filterwarnings("ignore")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CMOOP_GP_FIT_WORKERS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
    from cmoop_audio_processing_b200.gp_fit import fit_gprs_parallel
    rng = np.random.default_rng(0)
    x = rng.random((30, 4)) * 4
    ys = [np.sin(x[:, 0]) + 0.05 * rng.standard_normal(30), np.cos(x[:, 1])]
    state = np.random.RandomState(9)
    kernel = lambda: ConstantKernel(1.0) * Matern(length_scale=1.0, nu=1.5) + WhiteKernel(noise_level=0.1)   # noqa: E731
    gprs = fit_gprs_parallel([kernel(), kernel()], x, ys, n_restarts_optimizer=2, random_state=state)
    np.save(os.path.join(tmp, f"theta{rank}.npy"), np.array([g.kernel_.theta for g in gprs]))
    np.save(os.path.join(tmp, f"next{rank}.npy"), np.array([state.uniform()]))
    dist.destroy_process_group()


def test_gp_fit_rank0_broadcast_gloo(tmp_path):
    """The surrogate is replicated: every rank draws the same optimiser starts (streams stay in step) but rank 0 alone
    optimises and broadcasts, so all ranks end with identical hyper-parameters."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    np.testing.assert_array_equal(np.load(tmp_path / "theta0.npy"), np.load(tmp_path / "theta1.npy"))
    np.testing.assert_array_equal(np.load(tmp_path / "next0.npy"), np.load(tmp_path / "next1.npy"))


def _gp_error_worker(rank, world, port, tmp):
    import warnings

    import torch.distributed as dist
    warnings.filterwarnings("ignore")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CMOOP_GP_FIT_WORKERS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sklearn.gaussian_process.kernels import Matern
    from cmoop_audio_processing_b200 import gp_fit
    rng = np.random.default_rng(0)
    x, y = rng.random((12, 3)), rng.random(12)
    outcome = []
    # (a) an error that only rank 0 can hit (the optimiser itself runs there alone)
    real = gp_fit._optimise_start

    def broken(payload):
        raise FloatingPointError("objective blew up on the optimising rank")
    gp_fit._optimise_start = broken
    try:
        gp_fit.fit_gprs_parallel([Matern(nu=2.5)], x, [y], n_restarts_optimizer=0)
        outcome.append("no error")
    except FloatingPointError as exc:
        outcome.append(f"FloatingPointError:{exc}")
    except RuntimeError as exc:
        outcome.append(f"RuntimeError:{exc}")
    gp_fit._optimise_start = real
    # (b) backend validation happens on every rank before the rank branch
    try:
        gp_fit.fit_gprs_parallel([Matern(nu=2.5)], x, [y], backend="tpu")
        outcome.append("no error")
    except ValueError:
        outcome.append("ValueError")
    # (c) the group is still usable afterwards
    g = gp_fit.fit_gprs_parallel([Matern(nu=2.5)], x, [y], n_restarts_optimizer=0)
    outcome.append(repr(g[0].kernel_.theta.tolist()))
    with open(os.path.join(tmp, f"out{rank}.txt"), "w") as fh:
        fh.write("\n".join(outcome))
    dist.destroy_process_group()


def test_gp_fit_error_reaches_every_rank_gloo(tmp_path):
    """if rank 0 raises while optimising, the other ranks must not sit in the broadcast until the
    collective times out -- the failure is exchanged and re-raised everywhere."""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_gp_error_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = (tmp_path / "out0.txt").read_text().split("\n")
    r1 = (tmp_path / "out1.txt").read_text().split("\n")
    assert r0[0].startswith("FloatingPointError:")
    assert r1[0].startswith("RuntimeError:") and "rank 0" in r1[0] and "FloatingPointError" in r1[0]
    assert r0[1] == r1[1] == "ValueError"
    assert r0[2] == r1[2]
