"""NumPy lane-level model of the warp FFT used by csrc/mfcc.cu (development aid).

Mirrors the kernel's data movement exactly: 32 lanes x 16 complex registers,
radix-16 over stride-32 samples, W512 twiddle, shared-memory transpose, radix-16 over
the even/odd halves, W32 combine with the neighbour lane, natural-order scatter and
the real-FFT unpack into 513 power bins.  Checked against numpy.fft.rfft.
"""
import numpy as np


def dft16(v):
    """DIF radix-2 x4 on axis -1 (16 points), result in natural order."""
    v = v.copy()
    n = 16
    half = 8
    while half >= 1:
        for blk in range(0, n, 2 * half):
            for j in range(half):
                u = v[..., blk + j].copy()
                t = v[..., blk + j + half].copy()
                v[..., blk + j] = u + t
                v[..., blk + j + half] = (u - t) * np.exp(-2j * np.pi * j / (2 * half))
        half //= 2
    out = np.empty_like(v)
    for i in range(n):
        r = int(f"{i:04b}"[::-1], 2)
        out[..., r] = v[..., i]
    return out


def warp_rfft_power(frame640, window):
    L = len(frame640)
    x = frame640 * window
    z = np.zeros(512, complex)
    z[: L // 2] = x[0::2] + 1j * x[1::2]
    lanes = np.arange(32)
    # step 0/1: lane l holds z[32a + l]
    reg = np.stack([z[32 * a + lanes] for a in range(16)], axis=1)       # [32][16]
    y = dft16(reg)                                                          # Y[l][k1]
    # step 2
    k1 = np.arange(16)
    y = y * np.exp(-2j * np.pi * np.outer(lanes, k1) / 512)
    # step 3: smem[k1*34 + l]
    smem = np.zeros(16 * 34, complex)
    for l in lanes:
        for k in range(16):
            smem[k * 34 + l] = y[l, k]
    reg2 = np.zeros((32, 16), complex)
    for lp in lanes:
        kk, h = lp // 2, lp % 2
        for m in range(16):
            reg2[lp, m] = smem[kk * 34 + h + 2 * m]
    e = dft16(reg2)                                                         # E_h[q]
    # step 5: odd lanes multiply by W32^q, exchange with lane^1
    q = np.arange(16)
    w32 = np.exp(-2j * np.pi * q / 32)
    send = np.where((lanes % 2 == 1)[:, None], e * w32, e)
    recv = send[lanes ^ 1]
    res = np.where((lanes % 2 == 0)[:, None], send + recv, recv - send)     # lane (k1,h) holds k = k1+16q+256h
    # step 6: natural-order scatter with pad p(k) = k + 16*(k>>8)
    zs = np.zeros(544, complex)
    for lp in lanes:
        kk, h = lp // 2, lp % 2
        for qq in range(16):
            k = kk + 16 * qq + 256 * h
            zs[k + 16 * (k >> 8)] = res[lp, qq]
    power = np.zeros(513)
    for l in lanes:
        for j in range(8):
            k = l + 32 * j
            a = zs[k]
            kb = (512 - k) % 512
            b = np.conj(zs[kb + 16 * (kb >> 8)])
            ev = 0.5 * (a + b)
            od = -0.5j * (a - b)
            w = np.exp(-2j * np.pi * k / 1024)
            power[k] = abs(ev + w * od) ** 2
            power[512 - k] = abs(ev - w * od) ** 2
    z256 = zs[256 + 16]
    power[256] = abs(z256) ** 2
    return power, zs


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    f = rng.standard_normal(640)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(640) / 640)
    p, zs = warp_rfft_power(f, win)
    ref = np.abs(np.fft.rfft(f * win, n=1024)) ** 2
    print("max rel err", np.max(np.abs(p - ref) / ref.max()))
    assert np.allclose(p, ref, rtol=1e-10, atol=1e-10 * ref.max())
    print("ok")
