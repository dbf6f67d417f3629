import sys, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_gpu_cnn as t
from oracle import cnn_ref
from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
for variant, hp in t.GENOTYPES:
    xt, yt, xv, yv = t.make_data()
    prob = FitnessProblem(xt, yt, xv, yv, classes=t.N_CLASSES, config=TrainConfig(variant=variant, epochs=2))
    seed = 1234
    init = prob.debug_init_params(hp, seed)
    perm = prob.debug_permutation(seed, 0)
    losses, grads, params = prob.debug_train_steps(hp, seed, 5)
    out = []
    for dt in (torch.float32, torch.float64):
        model = cnn_ref.RefModel(hp, t.N_CLASSES, variant, t.unflatten(init, hp, variant), dtype=dt)
        ref, _ = cnn_ref.train_steps(model, xt, yt, perm, 5, seed=seed & 0xFFFFFFFF)
        out.append(np.array(ref))
    print(variant, hp['filters'], hp['kernel_size'], hp['use_bn'], 'ours-vs-ref64', np.abs(losses - out[1]) / out[1], 'ref32-vs-ref64', np.abs(out[0] - out[1]) / out[1])
