"""NumPy lane-level model of the frame-pair MFCC kernel's FFT (csrc/mfcc.cu, mfcc_pair_kernel).

Differences from fft_lane_model.py (the v3 kernel): lane = k1o + 16*h, the W32 combine partner is
lane ^ 16, the W512 twiddles are built as a power tree from W512^lane, and the real-FFT unpack is done
with shuffles (partner lane ((16-k1o)&15) + 16*(1-h), register 15-Q; lanes with k1o == 0 use their own
register Q+1 and treat Q == 15 as the self-paired bins 0 / 256 / 512).  Everything is kept at twice the
true amplitude (the kernel folds the 1/2 into the window).  Checked against numpy.fft.rfft.
"""
import numpy as np

from fft_lane_model import dft16


def pair_rfft_power(frame640, window):
    x = frame640 * window * 0.5                       # 1/2 folded into the window
    z = np.zeros(512, complex)
    z[:320] = x[0::2] + 1j * x[1::2]
    lanes = np.arange(32)
    reg = np.stack([z[32 * a + lanes] for a in range(16)], axis=1)
    y = dft16(reg)
    # W512^(lane*k1) as a power tree of depth <= 4
    w = [None] * 16
    w[1] = np.exp(-2j * np.pi * lanes / 512)
    w[2] = w[1] * w[1]; w[3] = w[2] * w[1]; w[4] = w[2] * w[2]
    w[5] = w[4] * w[1]; w[6] = w[4] * w[2]; w[7] = w[4] * w[3]; w[8] = w[4] * w[4]
    for j in range(1, 8):
        w[8 + j] = w[8] * w[j]
    for k1 in range(1, 16):
        y[:, k1] = y[:, k1] * w[k1]
    tile = np.zeros(16 * 33, complex)
    for l in lanes:
        for k1 in range(16):
            tile[k1 * 33 + l] = y[l, k1]
    k1o, h = lanes & 15, lanes >> 4
    reg2 = np.stack([tile[k1o * 33 + h + 2 * m] for m in range(16)], axis=1)
    s = dft16(reg2)
    q = np.arange(16)
    s = np.where((h == 1)[:, None], s * np.exp(-2j * np.pi * q / 32), s)
    r = s[lanes ^ 16]
    zq = np.where((h == 0)[:, None], s + r, r - s)        # lane holds Z[k1o + 16q + 256h]
    # shuffle unpack
    src = ((16 - k1o) & 15) + 16 * (1 - h)
    kb = np.where(k1o == 0, 16, k1o) + 256 * h
    power = np.full(544, np.nan)
    for Q in range(16):
        own = np.where(k1o == 0, zq[lanes, (Q + 1) & 15], zq[lanes, Q])
        recv = zq[src, 15 - Q]
        k = kb + 16 * Q
        if Q == 15:
            recv = np.where(k1o == 0, own, recv)
            k = np.where(k1o == 0, 256 * h, k)
        ar, ai, br, bi = own.real, own.imag, recv.real, -recv.imag
        er, ei, orr, oi = ar + br, ai + bi, ai - bi, br - ar
        th = 2 * np.pi * k / 1024
        c, sn = np.cos(th), np.sin(th)
        xr = er + c * orr + sn * oi
        xi = ei + c * oi - sn * orr
        power[k] = xr * xr + xi * xi
        if Q == 15:
            power[512] = (er[0] - orr[0]) ** 2
    return power[:513]


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    f = rng.standard_normal(640)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(640) / 640)
    p = pair_rfft_power(f, win)
    ref = np.abs(np.fft.rfft(f * win, n=1024)) ** 2
    assert not np.isnan(p).any()
    print("max rel err", np.max(np.abs(p - ref) / ref.max()))
    assert np.allclose(p, ref, rtol=1e-10, atol=1e-10 * ref.max())
    print("ok")
