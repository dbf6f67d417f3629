"""GPU parity: hypervolume / GD / IGD / Spread / C-metric and the MFCC front-end."""
import math

import numpy as np
import pytest

from oracle import hv_ref, mfcc_ref, nsga_ref

pytestmark = pytest.mark.gpu


def test_hypervolume_bit_exact_vs_oracle():
    from cmoop_audio_processing_b200 import quality
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 64, 256, 1000):
        pts = rng.random((n, 3))
        if n > 4:
            pts[1] = pts[0]
            pts[2, 2] = pts[3, 2]
        ref = quality.reference_point(pts)
        assert quality.hypervolume(pts, ref) == hv_ref.hypervolume(pts, ref)
        assert quality.hypervolume(pts[:, :2], ref[:2]) == hv_ref.hypervolume(pts[:, :2], ref[:2])
    assert quality.hypervolume(np.zeros((0, 3)), [1, 1, 1]) == 0.0
    assert quality.hypervolume([[0.5, 0.5, 0.5], [2, 0, 0]], [1, 1, 1]) == 0.125
    kws = np.stack([-(0.6 + 0.4 * rng.random(200)), 3 * rng.random(200), 0.1 * rng.random(200)], axis=1)
    front = kws[nsga_ref.nondominated_mask(kws)]
    ref = hv_ref.reference_point(kws)
    assert quality.hypervolume(front, ref) == hv_ref.hypervolume(front, ref)
    assert quality.hypervolume(kws, ref) == pytest.approx(hv_ref.hypervolume(front, ref), rel=1e-12)  # dominated points add nothing


def test_front_metrics_vs_reference_golden(golden):
    from cmoop_audio_processing_b200 import quality
    for case in golden("quality")["cases"]:
        fronts = [np.array(f) for f in case["fronts"]]
        allp = np.vstack(fronts)
        mask = quality.nondominated_mask(allp)
        assert mask.tolist() == case["true_mask"]
        true = allp[mask]
        for i, f in enumerate(fronts):
            assert quality.generational_distance(f, true) == pytest.approx(case["gd"][i], rel=1e-12, abs=1e-15)
            assert quality.inverted_gd(f, true) == pytest.approx(case["igd"][i], rel=1e-12, abs=1e-15)
            s = quality.spread_metric(f, true)
            if math.isnan(case["spread"][i]):
                assert math.isnan(s)
            else:
                assert s == pytest.approx(case["spread"][i], rel=1e-12)
            for j, b in enumerate(fronts):
                assert quality.coverage_metric(f, b) == pytest.approx(case["coverage"][i][j])


def _check_features(got, want):
    """north_star: within 1e-4 relative of the fp64 oracle.  Feature values are dB-scaled (|x| up to ~100), so
    the element-wise bound is |d| <= 1e-4 * max(|want|, 1); the clip-level relative error is far smaller."""
    got = np.asarray(got, np.float64)
    err = np.abs(got - want)
    bound = 1e-4 * np.maximum(np.abs(want), 1.0)
    worst = float((err / bound).max())
    assert worst <= 1.0, f"max error is {worst:.3f}x the 1e-4 bound"
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5


@pytest.mark.parametrize("n_mfcc", [40, 0, 13])
def test_mfcc_vs_fp64_oracle(n_mfcc):
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccConfig, MfccFrontEnd
    wave, _ = synth.make_clips(257, 12, seed=1234)                 # ragged vs the 8-warp CTA and the grid
    wave[0] = 0.0                                                  # silence -> log floor
    wave[1] = synth.uniform_clips(1, seed=3)[0]
    fe = MfccFrontEnd(MfccConfig(n_mfcc=n_mfcc))
    want = mfcc_ref.mfcc(wave, mfcc_ref.MfccSpec(n_mfcc=n_mfcc))
    got_host = fe(wave)                                            # host-buffer C-ABI path
    assert got_host.shape == want.shape == (257, 49, fe.n_out)
    _check_features(got_host, want)
    dev = fe(torch.from_numpy(wave).cuda())                        # device-pointer path, torch stream
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dev.cpu().numpy(), got_host)


def test_mfcc_shapes_edges_and_standardise():
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccConfig, MfccFrontEnd
    fe = MfccFrontEnd()
    assert fe.n_frames(16000) == 49 and fe.n_frames(639) == 0 and fe.n_frames(640) == 1
    assert fe(np.zeros((0, 16000), np.float32)).shape == (0, 49, 40)
    assert fe(np.zeros((3, 100), np.float32)).shape == (3, 0, 40)
    # odd hop / odd length exercise the scalar-load path
    wave, _ = synth.make_clips(5, 12, seed=7)
    odd = np.ascontiguousarray(wave[:, :15999])
    fe2 = MfccFrontEnd(MfccConfig(hop=321, n_mfcc=0))
    want = mfcc_ref.log_mel(odd, mfcc_ref.MfccSpec(hop=321, n_mfcc=0))
    _check_features(fe2(odd), want)
    _check_features(fe2(torch.from_numpy(odd).cuda()).cpu().numpy(), want)
    # 25 ms / 10 ms framing (frame_length 400 -> structural-zero pruning path NZ=8)
    fe3 = MfccFrontEnd(MfccConfig(frame_length=400, hop=160, n_mfcc=13))
    want3 = mfcc_ref.mfcc(wave, mfcc_ref.MfccSpec(frame_length=400, hop=160, n_mfcc=13))
    got3 = fe3(wave)
    assert got3.shape == (5, 98, 13)
    _check_features(got3, want3)
    # fused StandardScaler epilogue (prepare_dataset, nsga_penalty.py:102-114)
    base = mfcc_ref.mfcc(wave)
    scaled, mean, scale = mfcc_ref.standardise(base)
    fe.set_standardise(mean, scale)
    got = fe(wave).astype(np.float64)
    assert np.abs(got - scaled).max() < 2e-4
    fe.set_standardise(None, None)
    _check_features(fe(wave), base)
    # other clip lengths on the frame-pair kernel: even (24) and odd (37) frame counts, a lone last frame per clip
    for n_samples in (8000, 12160):
        short = np.ascontiguousarray(wave[:, :n_samples])
        got_s = fe(short)
        assert got_s.shape == (5, 1 + (n_samples - 640) // 320, 40)
        _check_features(got_s, mfcc_ref.mfcc(short))
    # a device pointer that is 8- but not 16-byte aligned cannot be a bulk-TMA source: the one-frame-per-warp kernel
    # (float2 loads) serves it and must meet the same tolerance
    flat = torch.zeros(5 * 16000 + 2, device="cuda")
    view = flat[2:].view(5, 16000)
    view.copy_(torch.from_numpy(wave))
    assert view.data_ptr() % 16 == 8
    _check_features(fe(view).cpu().numpy(), base)
    with pytest.raises(RuntimeError):
        MfccFrontEnd(MfccConfig(n_fft=1000))                       # not a power of two
    with pytest.raises(RuntimeError):
        MfccFrontEnd(MfccConfig(n_fft=2048, frame_length=640, center=True))   # centring needs frame_length == n_fft


def test_mfcc_full_size_properties():
    """BASELINE config 2 shape (65 536 clips is 4.2 GB; 8 192 clips keeps the test short): linearity in
    amplitude (scaling the waveform by g shifts log-mel by 20 log10 g) and clip independence."""
    import torch
    from cmoop_audio_processing_b200.features import MfccConfig, MfccFrontEnd
    g = torch.Generator(device="cuda").manual_seed(2)
    wave = torch.rand((8192, 16000), generator=g, device="cuda") * 2 - 1
    fe = MfccFrontEnd(MfccConfig(n_mfcc=0))
    a = fe(wave)
    b = fe(wave * 0.5)
    torch.cuda.synchronize()
    assert torch.allclose(a - b, torch.full_like(a, 20 * math.log10(2.0)), atol=2e-3)
    perm = torch.randperm(8192, device="cuda", generator=g)
    c = fe(wave[perm].contiguous())
    assert torch.equal(c, a[perm])
    sub = wave[:64].cpu().numpy()
    _check_features(a[:64].cpu().numpy(), mfcc_ref.log_mel(sub, mfcc_ref.MfccSpec(n_mfcc=0)))


def test_generic_front_end_birdclef_shape_and_other_fft_sizes():
    """Generic CTA-per-frame kernel: BirdCLEF-shaped 32 kHz / n_fft 2048 / hop 512 / centred / 128 mel (313 frames for
    5 s, BASELINE configs[3]), plus a 512-point and a zero-padded 2048-point configuration."""
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccConfig, MfccFrontEnd
    wave, _ = synth.make_clips(3, 12, sample_rate=32000, seconds=5.0, seed=11)
    assert wave.shape == (3, 160000)
    cfg = dict(sample_rate=32000, frame_length=2048, hop=512, n_fft=2048, n_mels=128, n_mfcc=0, f_max=16000.0, center=True)
    fe = MfccFrontEnd(MfccConfig(**cfg))
    assert fe.n_frames(160000) == 313 and fe.n_out == 128
    want = mfcc_ref.log_mel(wave, mfcc_ref.MfccSpec(**cfg))
    got = fe(wave)
    assert got.shape == (3, 313, 128)
    _check_features(got, want)
    np.testing.assert_array_equal(fe(torch.from_numpy(wave).cuda()).cpu().numpy(), got)
    small, _ = synth.make_clips(4, 12, seed=3)
    for cfg in (dict(frame_length=400, hop=160, n_fft=512, n_mels=40, n_mfcc=13),
                dict(frame_length=640, hop=320, n_fft=2048, n_mels=64, n_mfcc=20),
                dict(frame_length=1024, hop=256, n_fft=1024, n_mels=80, n_mfcc=0, center=True),
                dict(frame_length=401, hop=160, n_fft=1024, n_mels=40, n_mfcc=0)):
        fe = MfccFrontEnd(MfccConfig(**cfg))
        want = mfcc_ref.mfcc(small, mfcc_ref.MfccSpec(**cfg))
        got = fe(small)
        assert got.shape == want.shape
        _check_features(got, want)


def test_mfcc_int16_pcm_input():
    """16-bit PCM entry points (cmoop_mfcc_fwd_{host,dev}_i16): features of sample / 32768, the widening is exact, so
    the result must equal the fp32 path on the widened waveform bit for bit and meet the oracle tolerance."""
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccFrontEnd
    wave, _ = synth.make_clips(37, 12, seed=5)
    pcm = np.clip(np.round(wave * 32767.0), -32768, 32767).astype(np.int16)
    widened = pcm.astype(np.float32) / 32768.0
    fe = MfccFrontEnd()
    got_host = fe(pcm)
    np.testing.assert_array_equal(got_host, fe(widened))
    _check_features(got_host, mfcc_ref.mfcc(widened))
    got_dev = fe(torch.from_numpy(pcm).cuda())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got_dev.cpu().numpy(), got_host)
    odd = np.ascontiguousarray(pcm[:3, :15999])                    # unaligned rows exercise the scalar tail of the widening
    np.testing.assert_array_equal(fe(odd), fe(odd.astype(np.float32) / 32768.0))


def test_mfcc_int16_two_streams_do_not_share_scratch():
    """the device int16 entry point widened PCM into process-global scratch; two calls on different streams
    (or two handles) raced on it.  The widened chunk is now a stream-ordered allocation of the caller's stream."""
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccFrontEnd
    wave_a, _ = synth.make_clips(600, 12, seed=11)
    wave_b, _ = synth.make_clips(600, 12, seed=12)
    to_pcm = lambda w: torch.from_numpy(np.clip(np.round(w * 32767.0), -32768, 32767).astype(np.int16)).cuda()   # noqa: E731
    pcm_a, pcm_b = to_pcm(wave_a), to_pcm(wave_b)
    fe_a, fe_b = MfccFrontEnd(), MfccFrontEnd()
    want_a, want_b = fe_a(pcm_a).clone(), fe_b(pcm_b).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        with torch.cuda.stream(s1):
            got_a = fe_a(pcm_a)
        with torch.cuda.stream(s2):
            got_b = fe_b(pcm_b)
        torch.cuda.synchronize()
        assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b)
