"""GPU: the tcgen05 implicit-GEMM convolution (forward + data gradient) against torch on bf16-rounded operands
(products are then exact in fp32, so the tolerance only covers accumulation order) and against the fp32 SIMT kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_conv(mode, use_tc, inp, w, bias, n, H, W, Cin, Cout, k, stride, relu):
    from cmoop_audio_processing_b200 import _lib
    lib = _lib.load()
    Ho, Wo = (H, W) if stride == 1 else ((H + 1) // 2, (W + 1) // 2)
    out = np.zeros((n, Ho, Wo, Cout) if mode == 0 else (n, H, W, Cin), np.float32)
    _lib.check(lib.cmoop_cnn_debug_conv(mode, use_tc, _lib.ptr(inp), _lib.ptr(w), _lib.ptr(bias), n, H, W, Cin, Cout, k,
                                        stride, relu, _lib.ptr(out)), "cmoop_cnn_debug_conv")
    return out


def bf16_round(a):
    import torch
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


CASES = [  # n, H, W, Cin, Cout, k, stride
    (3, 9, 8, 16, 16, 3, 1),
    (64, 49, 40, 16, 16, 3, 1),      # stem2, F=16: K=144 -> one padded K block of 3
    (8, 25, 20, 32, 64, 5, 1),       # K=800 -> 13 K blocks (ring wraps 4x), N tile 64
    (5, 13, 10, 64, 128, 3, 1),
    (4, 7, 5, 128, 256, 3, 1),       # two N tiles
    (7, 25, 20, 32, 64, 1, 2),       # residual skip projection 1x1 / stride 2 (odd height)
    (2, 49, 40, 64, 64, 5, 1),
]


@pytest.mark.parametrize("n,H,W,Cin,Cout,k,stride", CASES)
def test_forward_and_dgrad(n, H, W, Cin, Cout, k, stride):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 1000 + Cin + Cout + k)
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    pad = (k - 1) // 2 if stride == 1 else 0
    xt = torch.from_numpy(bf16_round(x)).permute(0, 3, 1, 2).double()
    wt = torch.from_numpy(bf16_round(w)).permute(3, 2, 0, 1).double()
    ref = F.conv2d(xt, wt, torch.from_numpy(b).double(), stride=stride, padding=pad).permute(0, 2, 3, 1).numpy()
    for relu in (0, 1):
        want = np.maximum(ref, 0) if relu else ref
        got = run_conv(0, 1, x, w, b, n, H, W, Cin, Cout, k, stride, relu)
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-4)
    simt = run_conv(0, 0, x, w, b, n, H, W, Cin, Cout, k, stride, 0)
    full = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).permute(3, 2, 0, 1).double(),
                    torch.from_numpy(b).double(), stride=stride, padding=pad).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(simt, full, rtol=1e-4, atol=1e-4)                     # fp32 SIMT path is the exact one
    assert np.abs(got if not relu else run_conv(0, 1, x, w, b, n, H, W, Cin, Cout, k, stride, 0) - full).max() < 0.05
    # data gradient
    Ho, Wo = ref.shape[1:3]
    dy = rng.standard_normal((n, Ho, Wo, Cout)).astype(np.float32)
    dyt = torch.from_numpy(bf16_round(dy)).permute(0, 3, 1, 2).double()
    xin = torch.zeros((n, Cin, H, W), dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, wt, None, stride=stride, padding=pad).backward(dyt)
    want_dx = xin.grad.permute(0, 2, 3, 1).numpy()
    got_dx = run_conv(1, 1, dy, w, None, n, H, W, Cin, Cout, k, stride, 0)
    np.testing.assert_allclose(got_dx, want_dx, rtol=2e-4, atol=2e-4)
    xin2 = torch.zeros((n, Cin, H, W), dtype=torch.float64, requires_grad=True)
    F.conv2d(xin2, torch.from_numpy(w).permute(3, 2, 0, 1).double(), None, stride=stride, padding=pad).backward(
        torch.from_numpy(dy).permute(0, 3, 1, 2).double())
    np.testing.assert_allclose(run_conv(1, 0, dy, w, None, n, H, W, Cin, Cout, k, stride, 0),
                               xin2.grad.permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-4)


def run_wgrad(use_tc, x, dy, n, H, W, Cin, Cout, k, stride, splits):
    from cmoop_audio_processing_b200 import _lib
    lib = _lib.load()
    out = np.zeros((k * k * Cin + 1, Cout), np.float32)
    _lib.check(lib.cmoop_cnn_debug_wgrad(use_tc, _lib.ptr(x), _lib.ptr(dy), n, H, W, Cin, Cout, k, stride, splits,
                                         _lib.ptr(out)), "cmoop_cnn_debug_wgrad")
    return out


@pytest.mark.parametrize("n,H,W,Cin,Cout,k,stride", CASES)
@pytest.mark.parametrize("splits", [1, 5])
def test_weight_gradient(n, H, W, Cin, Cout, k, stride, splits):
    """dW (HWIO rows) and db (last row) on tcgen05 with MN-major operands vs torch on bf16-rounded operands; the fp32
    SIMT kernel vs torch on the unrounded operands."""
    import torch
    rng = np.random.default_rng(7 + n + Cin + Cout + k + splits)
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32)
    Ho, Wo = (H, W) if stride == 1 else ((H + 1) // 2, (W + 1) // 2)
    dy = rng.standard_normal((n, Ho, Wo, Cout)).astype(np.float32)
    pad = (k - 1) // 2 if stride == 1 else 0

    def ref(xa, ya):
        xt = torch.from_numpy(xa).permute(0, 3, 1, 2).double()
        yt = torch.from_numpy(ya).permute(0, 3, 1, 2).double()
        gw = torch.nn.grad.conv2d_weight(xt, (Cout, Cin, k, k), yt, stride=stride, padding=pad)      # OIHW
        return np.concatenate([gw.permute(2, 3, 1, 0).reshape(k * k * Cin, Cout).numpy(), yt.sum(dim=(0, 2, 3)).numpy()[None]])

    want = ref(bf16_round(x), bf16_round(dy))
    got = run_wgrad(1, x, dy, n, H, W, Cin, Cout, k, stride, splits)
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-4 * scale)
    want32 = ref(x, dy)
    got32 = run_wgrad(0, x, dy, n, H, W, Cin, Cout, k, stride, splits)
    np.testing.assert_allclose(got32, want32, rtol=1e-3, atol=2e-4 * np.abs(want32).max())


STEM_CASES = [  # n, H, W, Cout, k   (Cin = 1; the dedicated kernels of csrc/cnn/stem.cu, use_tc = 2)
    (64, 49, 40, 16, 3),       # KWS stem, every genotype corner
    (64, 49, 40, 64, 5),
    (7, 49, 40, 32, 5),        # ragged last batch: blocks past M, a block straddling two samples
    (1, 49, 40, 64, 3),
    (3, 128, 313, 32, 3),      # BirdCLEF-shaped map: a block covers ~3 rows
    (5, 33, 37, 64, 5),        # odd sizes, 32 channel groups per pixel (k = 5: 2 channels per thread)
    (2, 40, 40, 128, 3),       # 32 channel groups per pixel (k = 3: 4 channels per thread)
]


@pytest.mark.parametrize("n,H,W,Cout,k", STEM_CASES)
def test_stem_kernels(n, H, W, Cout, k):
    """Cin = 1 forward (+bias, ReLU) and weight gradient against torch fp64 and against the generic fp32 SIMT kernels."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(n + H + W + Cout + k)
    x = rng.standard_normal((n, H, W, 1)).astype(np.float32)
    w = (rng.standard_normal((k, k, 1, Cout)) / k).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    pad = (k - 1) // 2
    full = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).permute(3, 2, 0, 1).double(),
                    torch.from_numpy(b).double(), padding=pad).permute(0, 2, 3, 1).numpy()
    for relu in (0, 1):
        got = run_conv(0, 2, x, w, b, n, H, W, 1, Cout, k, 1, relu)
        np.testing.assert_allclose(got, np.maximum(full, 0) if relu else full, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(run_conv(0, 2, x, w, b, n, H, W, 1, Cout, k, 1, 0), run_conv(0, 0, x, w, b, n, H, W, 1, Cout, k, 1, 0),
                               rtol=1e-5, atol=1e-5)
    dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    yt = torch.from_numpy(dy).permute(0, 3, 1, 2).double()
    gw = torch.nn.grad.conv2d_weight(xt, (Cout, 1, k, k), yt, padding=pad)
    want = np.concatenate([gw.permute(2, 3, 1, 0).reshape(k * k, Cout).numpy(), yt.sum(dim=(0, 2, 3)).numpy()[None]])
    got = run_wgrad(2, x, dy, n, H, W, 1, Cout, k, 1, 1)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5 * np.abs(want).max())


STEM_TC_CASES = [c for c in STEM_CASES if c[3] <= 64] + [
    (64, 49, 40, 32, 3),
    (64, 49, 40, 16, 5),
    (13, 49, 40, 64, 5),       # 13 * 1960 pixels: last block and last BN tile partial
    (2, 128, 313, 64, 5),      # odd width: zero pad column in the weight gradient's pixel pairs
]


@pytest.mark.parametrize("n,H,W,Cout,k", STEM_TC_CASES)
def test_stem_tensor_core_kernels(n, H, W, Cout, k):
    """stem_tc.cu (use_tc = 4, precision bf16): mma.sync on bf16 hi + lo operands.  The arithmetic is fp32-grade, the only
    bf16 rounding is the STORED output (forward) / the stored output gradient (weight gradient), so the forward must be the
    fp64 result rounded to bf16 (one-ulp slack where the fp32-grade sum sits on a rounding boundary) and the weight gradient
    the fp64 one of the bf16-rounded dy and the UNROUNDED x.  The hook also verifies the fused BN partial sums."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(n + H + W + Cout + k)
    x = rng.standard_normal((n, H, W, 1)).astype(np.float32)
    w = (rng.standard_normal((k, k, 1, Cout)) / k).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    pad = (k - 1) // 2
    full = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).permute(3, 2, 0, 1).double(),
                    torch.from_numpy(b).double(), padding=pad).permute(0, 2, 3, 1).numpy()
    for relu in (0, 1):
        want = np.maximum(full, 0) if relu else full
        got = run_conv(0, 4, x, w, b, n, H, W, 1, Cout, k, 1, relu)
        assert np.array_equal(got, bf16_round(got))                                  # stored values are bf16
        np.testing.assert_allclose(got, want, rtol=2.0 ** -8, atol=2e-4)             # half a bf16 ulp + the hi/lo split error (2^-16 of the term magnitudes)
        assert (got == bf16_round(want.astype(np.float32))).mean() > 0.98
    dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    yt = torch.from_numpy(bf16_round(dy)).permute(0, 3, 1, 2).double()
    gw = torch.nn.grad.conv2d_weight(xt, (Cout, 1, k, k), yt, padding=pad)
    want = np.concatenate([gw.permute(2, 3, 1, 0).reshape(k * k, Cout).numpy(), yt.sum(dim=(0, 2, 3)).numpy()[None]])
    got = run_wgrad(4, x, dy, n, H, W, 1, Cout, k, 1, 1)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=5e-5 * np.abs(want).max())


SKIP_CASES = [  # n, H, W, Cin, Cout   (1x1 / stride 2, skip_tc.cu, use_tc = 5)
    (64, 25, 20, 16, 32),      # smallest genotype: 16-channel K tail only, 32-channel slice
    (64, 25, 20, 64, 128),
    (7, 13, 10, 32, 64),       # odd height, ragged last batch
    (5, 13, 10, 128, 256),
    (3, 7, 5, 256, 512),       # deepest block: K = 256 forward, K = 512 data gradient
    (9, 25, 20, 32, 64),
]


@pytest.mark.parametrize("n,H,W,Cin,Cout", SKIP_CASES)
def test_skip_projection_kernel(n, H, W, Cin, Cout):
    """skip_tc.cu: forward (+bias, ReLU) and strided-scatter data gradient vs torch fp64 on bf16-rounded operands, and
    bit-for-bit agreement in kind with the tcgen05 kernel it replaces (same operands, fp32 accumulation)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 1000 + Cin + Cout)
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((1, 1, Cin, Cout)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    xt = torch.from_numpy(bf16_round(x)).permute(0, 3, 1, 2).double()
    wt = torch.from_numpy(bf16_round(w)).permute(3, 2, 0, 1).double()
    ref = F.conv2d(xt, wt, torch.from_numpy(b).double(), stride=2).permute(0, 2, 3, 1).numpy()
    for relu in (0, 1):
        got = run_conv(0, 5, x, w, b, n, H, W, Cin, Cout, 1, 2, relu)
        np.testing.assert_allclose(got, np.maximum(ref, 0) if relu else ref, rtol=2e-4, atol=2e-4)
    Ho, Wo = ref.shape[1:3]
    dy = rng.standard_normal((n, Ho, Wo, Cout)).astype(np.float32)
    dyt = torch.from_numpy(bf16_round(dy)).permute(0, 3, 1, 2).double()
    xin = torch.zeros((n, Cin, H, W), dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, wt, None, stride=2).backward(dyt)
    got_dx = run_conv(1, 5, dy, w, None, n, H, W, Cin, Cout, 1, 2, 0)
    np.testing.assert_allclose(got_dx, xin.grad.permute(0, 2, 3, 1).numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(got_dx, run_conv(1, 1, dy, w, None, n, H, W, Cin, Cout, 1, 2, 0), rtol=1e-5, atol=1e-5)


PATCH_CASES = [c for c in CASES if c[6] == 1] + [
    (64, 25, 20, 16, 32, 3, 1),      # first residual conv of the smallest genotype: one sub-slab, nine taps
    (64, 13, 10, 128, 256, 5, 1),    # deep block: 8 sub-slabs x 25 taps, two N tiles
    (9, 7, 5, 256, 512, 3, 1),       # last block: 16 sub-slabs, four N tiles, rows mostly padding positions
    (3, 49, 40, 32, 32, 5, 1),       # variant-A stem2 shape with the widest patch (Wp = 44)
]


@pytest.mark.parametrize("n,H,W,Cin,Cout,k,stride", PATCH_CASES)
def test_patch_resident_kernel(n, H, W, Cin, Cout, k, stride):
    """conv_tc2.cu (use_tc = 3): forward (+bias, ReLU) and data gradient vs torch fp64 on bf16-rounded operands."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 1000 + Cin + Cout + k)
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    pad = (k - 1) // 2
    xt = torch.from_numpy(bf16_round(x)).permute(0, 3, 1, 2).double()
    wt = torch.from_numpy(bf16_round(w)).permute(3, 2, 0, 1).double()
    ref = F.conv2d(xt, wt, torch.from_numpy(b).double(), padding=pad).permute(0, 2, 3, 1).numpy()
    for relu in (0, 1):
        got = run_conv(0, 3, x, w, b, n, H, W, Cin, Cout, k, 1, relu)
        np.testing.assert_allclose(got, np.maximum(ref, 0) if relu else ref, rtol=2e-4, atol=2e-4)
    dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    dyt = torch.from_numpy(bf16_round(dy)).permute(0, 3, 1, 2).double()
    xin = torch.zeros((n, Cin, H, W), dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, wt, None, padding=pad).backward(dyt)
    got_dx = run_conv(1, 3, dy, w, None, n, H, W, Cin, Cout, k, 1, 0)
    np.testing.assert_allclose(got_dx, xin.grad.permute(0, 2, 3, 1).numpy(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("n,H,W,Cin,Cout,k,stride", [c for c in PATCH_CASES if c[5] in (3, 5)])
def test_patch_weight_gradient(n, H, W, Cin, Cout, k, stride):
    """wgrad_tc2.cu (use_tc = 3): dW and db from the resident patch (two taps per UMMA, bias gradient as the tap after
    the last) vs torch fp64 on bf16-rounded operands."""
    import torch
    rng = np.random.default_rng(11 + n + Cin + Cout + k)
    x = rng.standard_normal((n, H, W, Cin)).astype(np.float32)
    dy = rng.standard_normal((n, H, W, Cout)).astype(np.float32)
    pad = (k - 1) // 2
    xt = torch.from_numpy(bf16_round(x)).permute(0, 3, 1, 2).double()
    yt = torch.from_numpy(bf16_round(dy)).permute(0, 3, 1, 2).double()
    gw = torch.nn.grad.conv2d_weight(xt, (Cout, Cin, k, k), yt, padding=pad)
    want = np.concatenate([gw.permute(2, 3, 1, 0).reshape(k * k * Cin, Cout).numpy(), yt.sum(dim=(0, 2, 3)).numpy()[None]])
    got = run_wgrad(3, x, dy, n, H, W, Cin, Cout, k, 1, 1)
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-4 * np.abs(want).max())
