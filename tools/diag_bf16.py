import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_gpu_cnn import make_data, unflatten, GENOTYPES, N_CLASSES
from oracle import cnn_ref
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
gi = int(sys.argv[1]) if len(sys.argv) > 1 else 1
variant, hp = GENOTYPES[gi]
xt, yt, xv, yv = make_data()
seed = 4321
for prec in ("fp32", "bf16"):
    prob = FitnessProblem(xt, yt, xv, yv, classes=N_CLASSES, config=TrainConfig(variant=variant, epochs=2, precision=prec))
    init = prob.debug_init_params(hp, seed)
    perm = prob.debug_permutation(seed, 0)
    losses, grads, _ = prob.debug_train_steps(hp, seed, 1)
    model = cnn_ref.RefModel(hp, N_CLASSES, variant, unflatten(init, hp, variant), dtype=torch.float64, bf16_convs=(prec == "bf16"))
    idx = perm[:64]
    p = model.forward(torch.from_numpy(xt[idx]), training=True, drop_ctx=(seed & 0xFFFFFFFF, 0))
    loss = cnn_ref.keras_sparse_ce(p, torch.from_numpy(yt[idx]).long()).mean(); loss.backward()
    print(prec, "loss gpu", losses[0], "oracle", float(loss.detach()))
    off = 0
    for name, shape in cnn_ref.param_shapes(hp, N_CLASSES, variant):
        n = int(np.prod(shape)); g = model.p[name].grad
        ref = np.zeros(n) if g is None else g.numpy().ravel()
        got = grads[off:off + n]; off += n
        if np.linalg.norm(ref) > 0:
            print(f"  {name:14s} relL2 {np.linalg.norm(got-ref)/np.linalg.norm(ref):.2e}  max {np.abs(got-ref).max():.2e} / {np.abs(ref).max():.2e}")
