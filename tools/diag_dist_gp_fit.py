"""torchrun check: under torch.distributed rank 0 runs the (device-backed) GP fit and broadcasts the optima; every rank must
end up with the same surrogate.  torchrun --nproc-per-node 2 tools/diag_dist_gp_fit.py [host|device]"""
import os
import random
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from cmoop_audio_processing_b200 import gp_fit  # noqa: E402
from cmoop_audio_processing_b200.surrogate import SurrogateManager  # noqa: E402


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else "device"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    rank = dist.get_rank()
    rng = random.Random(0)
    np.random.seed(0)
    hps = [{"filters": rng.choice([16, 32, 64, 128]), "kernel_size": rng.choice([3, 5]), "residual_blocks": rng.choice([1, 2, 3]),
            "fc_layers": rng.choice([1, 2, 3]), "use_bn": rng.choice([True, False]), "use_dropout": rng.choice([True, False])}
           for _ in range(120)]
    recs = [{"hparams": hp, "objs": [-0.5 - 0.3 * rng.random(), hp["filters"] / 50.0, 0.2 * rng.random()], "CV": rng.random()}
            for hp in hps]
    sm = SurrogateManager(fit_backend=backend)
    sm.update(hps, recs)
    preds, stds = sm.predict(hps[:40], return_std=True)
    mine = torch.tensor(np.concatenate([preds[k] for k in sorted(preds)] + [stds[k] for k in sorted(stds)]), device="cuda")
    everyone = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(everyone, mine)
    same = all(torch.equal(everyone[0], e) for e in everyone)
    print(f"rank {rank}: backend {backend}, device fit stats {gp_fit.LAST_DEVICE_FIT if rank == 0 else '(rank 0 only)'}, "
          f"identical surrogate on every rank: {same}", flush=True)
    dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
