// Log-mel / MFCC front-end: one warp per frame, whole pipeline fused in one kernel.
//
// The reference has no feature-extraction code (features are loaded pre-computed,
// nsga_penalty.py:64-71, sa_nsga_penalty.py:42-63); the spec is oracle/mfcc_ref.py:
//   frame (no padding) -> periodic Hann -> |rfft(.,1024)|^2 -> 40 Slaney mel bands
//   -> 10*log10(max(.,1e-10)) -> [orthonormal DCT-II] -> [per-feature standardise]
// Output layout (clip, frame, feature) is what prepare_dataset expects
// (nsga_penalty.py:104-114).
//
// Kernel design (tools/fft_lane_model.py is the NumPy model of the data movement):
//   * a warp owns a frame; lane l loads float2 samples z[32a+l] = x[2n] + i x[2n+1]
//     straight from global memory (256 B per warp-load, fully coalesced; each sample
//     is touched by two frames and the second touch is an L1/L2 hit, so HBM sees the
//     waveform once) and multiplies by the window from shared memory
//   * 1024-point real FFT = 512-point complex FFT (16 x 32 Cooley-Tukey):
//       radix-16 in registers over a (inputs a >= frame_length/64 are structural zeros
//       and pruned at compile time) -> W512 twiddle -> transpose through a padded,
//       conflict-free shared-memory tile -> radix-16 over even/odd halves -> W32
//       combine with the neighbour lane via one shuffle per value
//   * real-FFT unpack: Z[k], Z[512-k] pairs from a padded natural-order tile give
//     |X[k]|^2 and |X[512-k]|^2 together
//   * mel: lane l owns bins [17l, 17l+17) (odd stride: conflict-free); each bin feeds the rising
//     edge of band s and the falling edge of band s-1 with per-bin weights; partial sums are
//     flushed at mel-segment ends and combined per band (12 adds); log10
//   * DCT-II folded by its even/odd symmetry (half the multiply-adds), lane c owns
//     coefficient c, the last 8 coefficients use 4 lanes each; optional (x-mean)*inv_scale;
//     160-B coalesced stores per frame
//   * persistent grid (multiple of the SM count), all tables staged once per CTA
// Bound: HBM by contract (71 840 B/clip) but ~1.6 MFLOP/clip of fp32 butterflies puts
// it between the HBM and the fp32-pipe roofs; bench.py reports the HBM fraction.
#include <math.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "mfcc_generic.cuh"

namespace {

constexpr int kWarps = 8;            // warps (frames in flight) per CTA
constexpr int kNfft = 1024;
constexpr int kHalf = kNfft / 2;     // complex FFT length
constexpr int kBins = kHalf + 1;
constexpr int kMaxMel = 64;
constexpr int kMaxOut = 64;
constexpr int kTileStride = 34;      // 16 x 34 floats per plane: conflict-free transpose
constexpr int kZPlane = 544;         // natural-order plane with 16 floats of padding at k>=256
constexpr int kChunk = 17;           // mel: bins per lane (odd -> conflict-free strided reads), 31 lanes cover 513 bins
constexpr int kPowFloats = 32 * kChunk;   // power spectrum padded so every lane can read a full chunk
constexpr int kMaxFlush = 6;         // mel: segment partials per lane

__host__ __device__ constexpr float cos32(int j) {
    constexpr float t[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
                             0.f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
                             -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f,
                             -0.98078528040323043f};
    return t[j];
}
__host__ __device__ constexpr float sin32(int j) {
    constexpr float t[16] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                             0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                             1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
                             0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};
    return t[j];
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

// a * exp(-2*pi*i*J/32), J a compile-time constant in [0,16)
template <int J>
__device__ __forceinline__ float2 mul_w32(float2 a) {
    if constexpr (J == 0) {
        return a;
    } else if constexpr (J == 8) {
        return make_float2(a.y, -a.x);
    } else if constexpr (J == 4) {
        constexpr float s = 0.70710678118654757f;
        return make_float2((a.x + a.y) * s, (a.y - a.x) * s);
    } else if constexpr (J == 12) {
        constexpr float s = 0.70710678118654757f;
        return make_float2((a.y - a.x) * s, -(a.x + a.y) * s);
    } else {
        constexpr float c = cos32(J), s = sin32(J);      // w = c - i s
        return make_float2(fmaf(a.y, s, a.x * c), fmaf(-a.x, s, a.y * c));
    }
}

__host__ __device__ constexpr int brev4(int i) {
    return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
}

// In-place 16-point DIF FFT; result X[brev4(i)] = v[i].  Inputs v[NZ..15] are known zeros.
template <int NZ>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    static_for<0, 8>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        if constexpr (J + 8 < NZ) {
            const float2 u = v[J], t = v[J + 8];
            v[J] = cadd(u, t);
            v[J + 8] = mul_w32<2 * J>(csub(u, t));
        } else {
            v[J + 8] = mul_w32<2 * J>(v[J]);
        }
    });
    static_for<0, 2>([&](auto bc) {
        constexpr int B = decltype(bc)::value * 8;
        static_for<0, 4>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            const float2 u = v[B + J], t = v[B + J + 4];
            v[B + J] = cadd(u, t);
            v[B + J + 4] = mul_w32<4 * J>(csub(u, t));
        });
    });
    static_for<0, 4>([&](auto bc) {
        constexpr int B = decltype(bc)::value * 4;
        static_for<0, 2>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            const float2 u = v[B + J], t = v[B + J + 2];
            v[B + J] = cadd(u, t);
            v[B + J + 2] = mul_w32<8 * J>(csub(u, t));
        });
    });
    static_for<0, 8>([&](auto bc) {
        constexpr int B = decltype(bc)::value * 2;
        const float2 u = v[B], t = v[B + 1];
        v[B] = cadd(u, t);
        v[B + 1] = csub(u, t);
    });
}

// Device-resident tables, one contiguous float buffer (offsets in floats).
struct Tables {
    int window2;   // float2 [half_len]           (w[2n], w[2n+1])
    int tw512;     // float2 [16][32]             exp(-2 pi i l k1 / 512)
    int tw1024;    // float2 [256]                exp(-2 pi i k / 1024)
    int mel_w;     // float  [mel_taps]           (mel_mode 0: per-band sparse taps)
    int mel_meta;  // int2   [n_mels]             (first bin, tap offset) ; count in mel_cnt
    int mel_cnt;   // int    [n_mels]
    int mel_wt;    // float2 [32*kChunk]          (mel_mode 1: per-bin rising / falling weight)
    int mel_flush; // int    [32]                 per-lane flush mask
    int mel_comb;  // int2   [n_mels]             packed (lane | k<<8 | n<<16) for the R and F partial runs
    int dct;       // float  [n_mels][n_out]      (b-major; dct_mode 0)
    int dcth;      // float  [n_mels/2][n_out]    (b-major; dct_mode 1: even/odd folded)
    int mean;      // float  [n_out]
    int inv_scale; // float  [n_out]
    int total;
};

struct Params {
    const float* wave;
    float* out;
    const float* tables;
    Tables t;
    long long n_frames_total;
    int n_samples, frames_per_clip, frame_length, hop, half_len;
    int n_mels, n_out, use_dct, vec2, mel_mode, dct_mode;
    int runs_per_clip, frames_per_run;   // FAST path: a warp walks a run of consecutive frames of one clip
    long long total_runs;
    float log_floor;
};

constexpr int kWarpFloats = 2 * 16 * kTileStride + kPowFloats;   // tile planes (re-used for mel partials, logmel, folded DCT input) | power
static_assert(2 * 16 * kTileStride >= 2 * kZPlane, "natural-order planes must fit in the transpose tile");
static_assert(2 * 16 * kTileStride >= 2 * kMaxFlush * 32 + 2 * kMaxMel, "mel partials + logmel + folded input must fit in the tile");
static_assert(kPowFloats >= kBins, "power buffer too small");

// FAST = 0: any supported configuration, one independent frame per warp iteration.
// FAST = 1 / 2: the default 640/320/40-mel configuration without / with the 40-point DCT, everything that
// depends on the configuration folded at compile time; a warp walks consecutive frames of one clip, keeps the
// overlapping half frame in registers (frame t+1 re-uses rows a+5 of frame t as its rows a) and prefetches the
// next half frame while the current FFT runs, so every sample is loaded exactly once and never waited for.
template <int NZ, int FAST>
__global__ void __launch_bounds__(kWarps * 32, FAST ? 2 : 3) mfcc_kernel(Params p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_mels = FAST ? 40 : p.n_mels, n_out = FAST ? 40 : p.n_out;
    const int use_dct = FAST ? (FAST == 2) : p.use_dct, mel_mode = FAST ? 1 : p.mel_mode, dct_mode = FAST ? 1 : p.dct_mode;
    // ---- stage tables once per CTA
    for (int i = threadIdx.x; i < p.t.total; i += blockDim.x) smem[i] = p.tables[i];
    __syncthreads();
    const float2* s_win = reinterpret_cast<const float2*>(smem + p.t.window2);
    const float2* s_tw512 = reinterpret_cast<const float2*>(smem + p.t.tw512);
    const float2* s_tw1024 = reinterpret_cast<const float2*>(smem + p.t.tw1024);
    const float* s_melw = smem + p.t.mel_w;
    const int2* s_meta = reinterpret_cast<const int2*>(smem + p.t.mel_meta);
    const int* s_cnt = reinterpret_cast<const int*>(smem + p.t.mel_cnt);
    const float2* s_melwt = reinterpret_cast<const float2*>(smem + p.t.mel_wt);
    const int2* s_comb = reinterpret_cast<const int2*>(smem + p.t.mel_comb);
    const float* s_dct = smem + p.t.dct;
    const float* s_dcth = smem + p.t.dcth;
    const unsigned flush_mask = mel_mode ? reinterpret_cast<const unsigned*>(smem + p.t.mel_flush)[lane] : 0u;
    const float* s_mean = smem + p.t.mean;
    const float* s_inv = smem + p.t.inv_scale;
    float* w_base = smem + ((p.t.total + 3) & ~3) + warp * kWarpFloats;
    float* t_re = w_base;                       // transpose tile, later natural-order planes
    float* t_im = w_base + 16 * kTileStride;
    float* z_re = w_base;
    float* z_im = w_base + kZPlane;
    float* s_pow = w_base + 2 * 16 * kTileStride;
    float* part_r = w_base;                                  // [kMaxFlush][32]   (tile is dead once the power spectrum exists)
    float* part_f = w_base + kMaxFlush * 32;
    float* s_lm = w_base + 2 * kMaxFlush * 32;               // [kMaxMel]
    float* s_sd = s_lm + kMaxMel;                            // [kMaxMel] folded DCT input
    for (int i = kBins + lane; i < kPowFloats; i += 32) s_pow[i] = 0.f;   // padding read with zero weights

    const int h = lane & 1, k1o = lane >> 1;
    const long long stride = (long long)gridDim.x * kWarps;
    const long long first = (long long)blockIdx.x * kWarps + warp;

    // everything after the (windowed) samples are in registers: FFT -> power -> mel -> log -> [DCT] -> store
    auto compute = [&](float2 (&v)[16], long long f) {
        // ---- radix-16 over a, W512 twiddle, transpose
        dft16<NZ>(v);
        static_for<0, 16>([&](auto ic) {
            constexpr int I = decltype(ic)::value;
            constexpr int K1 = brev4(I);
            float2 y = v[I];
            if constexpr (K1 != 0) y = cmul(y, s_tw512[K1 * 32 + lane]);
            t_re[K1 * kTileStride + lane] = y.x;
            t_im[K1 * kTileStride + lane] = y.y;
        });
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            v[m].x = t_re[k1o * kTileStride + h + 2 * m];
            v[m].y = t_im[k1o * kTileStride + h + 2 * m];
        }
        __syncwarp();
        // ---- radix-16 over m, W32 combine with the neighbour lane
        dft16<16>(v);
        static_for<0, 16>([&](auto ic) {
            constexpr int I = decltype(ic)::value;
            constexpr int Q = brev4(I);
            float2 s = v[I];
            if (h) s = mul_w32<Q>(s);
            float2 r;
            r.x = __shfl_xor_sync(0xffffffffu, s.x, 1);
            r.y = __shfl_xor_sync(0xffffffffu, s.y, 1);
            const float sg = h ? -1.f : 1.f;
            const float2 zk = make_float2(fmaf(sg, s.x, r.x), fmaf(sg, s.y, r.y));
            const int k = k1o + 16 * Q + 256 * h;        // natural-order bin of the 512-pt FFT
            z_re[k + 16 * h] = zk.x;
            z_im[k + 16 * h] = zk.y;
        });
        __syncwarp();
        // ---- real-FFT unpack -> power spectrum
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = lane + 32 * j;
            const int kb = (kHalf - k) & (kHalf - 1);
            const int pb = kb + ((kb >> 8) << 4);
            const float ar = z_re[k], ai = z_im[k];
            const float br = z_re[pb], bi = -z_im[pb];
            const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
            const float orr = 0.5f * (ai - bi), oi = -0.5f * (ar - br);     // -i/2 * (A - B)
            const float2 w = s_tw1024[k];
            const float pr = fmaf(orr, w.x, -oi * w.y), pi = fmaf(orr, w.y, oi * w.x);
            const float xr = er + pr, xi = ei + pi, yr = er - pr, yi = ei - pi;
            s_pow[k] = fmaf(xr, xr, xi * xi);
            s_pow[kHalf - k] = fmaf(yr, yr, yi * yi);
        }
        if (lane == 0) {
            const float r = z_re[256 + 16], i = z_im[256 + 16];
            s_pow[256] = fmaf(r, r, i * i);
        }
        __syncwarp();
        // ---- mel + log
        if (mel_mode) {
            // lane owns bins [17*lane, 17*lane+17): conflict-free; every bin feeds the rising edge of band s and
            // the falling edge of band s-1 (s = its mel segment); partial sums are flushed at segment ends
            float accr = 0.f, accf = 0.f;
            int nfl = 0;
            const int k0 = kChunk * lane;
#pragma unroll
            for (int i = 0; i < kChunk; ++i) {
                const float pw = s_pow[k0 + i];
                const float2 w = s_melwt[k0 + i];
                accr = fmaf(w.x, pw, accr);
                accf = fmaf(w.y, pw, accf);
                const bool fl = (flush_mask >> i) & 1u;
                if (fl) part_r[nfl * 32 + lane] = accr;
                if (fl) part_f[nfl * 32 + lane] = accf;
                accr = fl ? 0.f : accr;
                accf = fl ? 0.f : accf;
                nfl += fl ? 1 : 0;
            }
            __syncwarp();
            float mel[2];
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd) {
                const int b = lane + 32 * rnd;
                float acc = 0.f;
                if (b < n_mels) {
                    const int2 cb = s_comb[b];
                    int n = cb.x >> 16;
                    if (n) {
                        const int la = cb.x & 0xff;
                        acc = part_r[((cb.x >> 8) & 0xff) * 32 + la];
                        for (int j = 1; j < n; ++j) acc += part_r[la + j];
                    }
                    n = cb.y >> 16;
                    if (n) {
                        const int la = cb.y & 0xff;
                        acc += part_f[((cb.y >> 8) & 0xff) * 32 + la];
                        for (int j = 1; j < n; ++j) acc += part_f[la + j];
                    }
                }
                mel[rnd] = acc;
            }
            __syncwarp();                                     // partials are read; logmel aliases nothing they use
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd) {
                const int b = lane + 32 * rnd;
                if (b < n_mels) s_lm[b] = 10.f * log10f(fmaxf(mel[rnd], p.log_floor));
            }
        } else {
            for (int b = lane; b < n_mels; b += 32) {
                const int2 meta = s_meta[b];
                const int cnt = s_cnt[b];
                float acc = 0.f;
                for (int i = 0; i < cnt; ++i) acc = fmaf(s_melw[meta.y + i], s_pow[meta.x + i], acc);
                s_lm[b] = 10.f * log10f(fmaxf(acc, p.log_floor));
            }
        }
        __syncwarp();
        // ---- DCT / standardise / store
        float* dst = p.out + (size_t)f * n_out;
        if (use_dct && dct_mode) {
            // D[c][n-1-b] = (-1)^c D[c][b]: fold the input once, halve the multiply-adds
            const int half = n_mels >> 1;
            for (int b = lane; b < half; b += 32) {
                const float a = s_lm[b], z = s_lm[n_mels - 1 - b];
                s_sd[b] = a + z;
                s_sd[half + b] = a - z;
            }
            __syncwarp();
            if (lane < n_out) {
                const float* sd = s_sd + (lane & 1) * half;
                float acc = 0.f;
                for (int b = 0; b < half; ++b) acc = fmaf(s_dcth[b * n_out + lane], sd[b], acc);
                dst[lane] = (acc - s_mean[lane]) * s_inv[lane];
            }
            const int rem = n_out - 32;
            if (rem > 0 && rem <= 8) {
                // the last <= 8 coefficients: 4 lanes per coefficient, strided terms, two shuffles
                const int c = 32 + (lane >> 2), part = lane & 3;
                float acc = 0.f;
                if (c < n_out) {
                    const float* sd = s_sd + (c & 1) * half;
                    for (int b = part; b < half; b += 4) acc = fmaf(s_dcth[b * n_out + c], sd[b], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (part == 0 && c < n_out) dst[c] = (acc - s_mean[c]) * s_inv[c];
            } else if (rem > 0) {
                const int c = 32 + lane;
                if (c < n_out) {
                    const float* sd = s_sd + (c & 1) * half;
                    float acc = 0.f;
                    for (int b = 0; b < half; ++b) acc = fmaf(s_dcth[b * n_out + c], sd[b], acc);
                    dst[c] = (acc - s_mean[c]) * s_inv[c];
                }
            }
        } else {
            for (int c = lane; c < n_out; c += 32) {
                float acc;
                if (use_dct) {
                    acc = 0.f;
                    for (int b = 0; b < n_mels; ++b) acc = fmaf(s_dct[b * n_out + c], s_lm[b], acc);
                } else {
                    acc = s_lm[c];
                }
                dst[c] = (acc - s_mean[c]) * s_inv[c];
            }
        }
        __syncwarp();
    };

    if constexpr (FAST == 0) {
        for (long long f = first; f < p.n_frames_total; f += stride) {
            const long long clip = f / p.frames_per_clip;
            const int t = (int)(f - clip * p.frames_per_clip);
            const float* src = p.wave + (size_t)clip * p.n_samples + (size_t)t * p.hop;
            // ---- load + window: lane holds z[32a + lane]
            float2 v[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                if (a < NZ) {
                    const int n = 32 * a + lane;
                    float2 x = make_float2(0.f, 0.f);
                    if (n < p.half_len) {
                        if (p.vec2) {
                            x = __ldg(reinterpret_cast<const float2*>(src) + n);
                        } else {
                            x.x = __ldg(src + 2 * n);
                            x.y = __ldg(src + 2 * n + 1);
                        }
                        const float2 w = s_win[n];
                        x.x *= w.x;
                        x.y *= w.y;
                    }
                    v[a] = x;
                } else {
                    v[a] = make_float2(0.f, 0.f);
                }
            }
            compute(v, f);
        }
    } else {
        constexpr int HZ = NZ / 2;            // rows of 32 float2 per hop (hop == frame_length / 2 == 64 * HZ samples)
        for (long long run = first; run < p.total_runs; run += stride) {
            const long long clip = run / p.runs_per_clip;
            const int part = (int)(run - clip * p.runs_per_clip);
            const int t0 = part * p.frames_per_run;
            const int t1 = min(p.frames_per_clip, t0 + p.frames_per_run);
            if (t0 >= t1) continue;
            const float2* src = reinterpret_cast<const float2*>(p.wave + (size_t)clip * p.n_samples + (size_t)t0 * p.hop);
            float2 keep[HZ], nxt[HZ];
#pragma unroll
            for (int a = 0; a < HZ; ++a) {
                keep[a] = __ldg(src + 32 * a + lane);
                nxt[a] = __ldg(src + 32 * (a + HZ) + lane);
            }
            for (int t = t0; t < t1; ++t) {
                float2 v[16];
#pragma unroll
                for (int a = 0; a < HZ; ++a) {
                    const float2 w0 = s_win[32 * a + lane], w1 = s_win[32 * (a + HZ) + lane];
                    v[a] = make_float2(keep[a].x * w0.x, keep[a].y * w0.y);
                    v[a + HZ] = make_float2(nxt[a].x * w1.x, nxt[a].y * w1.y);
                    keep[a] = nxt[a];
                }
#pragma unroll
                for (int a = NZ; a < 16; ++a) v[a] = make_float2(0.f, 0.f);
                if (t + 1 < t1) {
                    const float2* nsrc = src + (size_t)(t + 1 - t0) * (32 * HZ);
#pragma unroll
                    for (int a = 0; a < HZ; ++a) nxt[a] = __ldg(nsrc + 32 * (a + HZ) + lane);
                }
                compute(v, clip * p.frames_per_clip + t);
            }
        }
    }
}

// =====================================================================================================
// Frame-pair kernel (default 640/320/40-mel configuration).  The v3 kernel above sits on the shared-memory
// data pipe (406 wavefronts per frame, 1 wavefront/clk/SM; shuffles share that pipe) and on instruction issue.
// Here a warp processes frames (t, t+1) of one clip together as packed f32x2 values (.x = frame t, .y = frame
// t+1): every butterfly is one FADD2/FMUL2/FFMA2 (twiddle constants are broadcast immediates), every
// shared-memory exchange is a 64-bit access, and every table read (window, mel weights, DCT) is shared by the
// two frames.  Exchanges that v3 made through shared memory and that can be made cheaper are:
//   * W512 twiddles: power tree (depth <= 4) from W512^lane instead of a 15-row table
//   * real-FFT unpack: partner bins fetched with shuffles (32 per frame instead of 80 wavefronts);
//     lane = k1o + 16 h holds Z[k1o + 16 Q + 256 h]; the partner 512 - k is register 15 - Q of lane
//     ((16 - k1o) & 15) + 16 (1 - h); lanes with k1o == 0 pair their own register Q + 1 instead (SEL) and keep
//     the self-paired bins 0 / 256 / 512 for the last round; W1024^k = W1024^kb(lane) * W64^Q (immediates)
//   * the power spectrum aliases the transpose tile; mel/log/DCT as in v3, on packed values
//   * waveform rows reach shared memory by 1-D bulk TMA (cp.async.bulk + mbarrier, one 3 840-B copy per pair issued
//     by lane 0 as soon as the previous pair's rows are windowed), so the next pair's samples land during the
//     whole FFT/mel/DCT of the current one without holding registers (a register prefetch was spilled by ptxas)
// tools/fft_pair_model.py is the NumPy model of the index maps.  Everything is kept at twice the true
// amplitude until the power spectrum (the 1/2 of the unpack is folded into the window while staging it).
// =====================================================================================================
typedef float2 P;   // (frame t, frame t+1)

__device__ __forceinline__ uint32_t pair_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pair_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(pair_smem_u32(bar)), "r"(parity) : "memory");
}
// one lane: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void pair_tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy reads of dst are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pair_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pair_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pair_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ P padd(P a, P b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ P pneg(P a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ P psub(P a, P b) { return __fadd2_rn(a, pneg(b)); }
__device__ __forceinline__ P pmul(P a, P b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ P pfma(P a, P b, P c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ P pdup(float c) { return make_float2(c, c); }

struct PC {   // packed complex
    P re, im;
};
__device__ __forceinline__ PC pc_add(PC a, PC b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ PC pc_sub(PC a, PC b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
// a * (wr + i wi), wr / wi duplicated pairs
__device__ __forceinline__ PC pc_mul(PC a, P wr, P wi) {
    return {pfma(a.re, wr, pneg(pmul(a.im, wi))), pfma(a.re, wi, pmul(a.im, wr))};
}

// a * exp(-2*pi*i*J/32), J a compile-time constant in [0,16)
template <int J>
__device__ __forceinline__ PC pc_mul_w32(PC a) {
    if constexpr (J == 0) {
        return a;
    } else if constexpr (J == 8) {
        return {a.im, pneg(a.re)};
    } else if constexpr (J == 4) {
        constexpr float s = 0.70710678118654757f;
        return {pmul(padd(a.re, a.im), pdup(s)), pmul(psub(a.im, a.re), pdup(s))};
    } else if constexpr (J == 12) {
        constexpr float s = 0.70710678118654757f;
        return {pmul(psub(a.im, a.re), pdup(s)), pmul(padd(a.re, a.im), pdup(-s))};
    } else {
        constexpr float c = cos32(J), s = sin32(J);      // w = c - i s
        return {pfma(a.im, pdup(s), pmul(a.re, pdup(c))), pfma(a.re, pdup(-s), pmul(a.im, pdup(c)))};
    }
}

template <int NZ>
__device__ __forceinline__ void pc_dft16(PC (&v)[16]) {
    static_for<0, 8>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        if constexpr (J + 8 < NZ) {
            const PC u = v[J], t = v[J + 8];
            v[J] = pc_add(u, t);
            v[J + 8] = pc_mul_w32<2 * J>(pc_sub(u, t));
        } else {
            v[J + 8] = pc_mul_w32<2 * J>(v[J]);
        }
    });
    static_for<0, 2>([&](auto bc) {
        constexpr int B = decltype(bc)::value * 8;
        static_for<0, 4>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            const PC u = v[B + J], t = v[B + J + 4];
            v[B + J] = pc_add(u, t);
            v[B + J + 4] = pc_mul_w32<4 * J>(pc_sub(u, t));
        });
    });
    static_for<0, 4>([&](auto bc) {
        constexpr int B = decltype(bc)::value * 4;
        static_for<0, 2>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            const PC u = v[B + J], t = v[B + J + 2];
            v[B + J] = pc_add(u, t);
            v[B + J + 2] = pc_mul_w32<8 * J>(pc_sub(u, t));
        });
    });
    static_for<0, 8>([&](auto bc) {
        constexpr int B = decltype(bc)::value * 2;
        const PC u = v[B], t = v[B + 1];
        v[B] = pc_add(u, t);
        v[B + 1] = pc_sub(u, t);
    });
}

__host__ __device__ constexpr float cos64(int q) {
    constexpr float t[16] = {1.f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f,
                             0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f,
                             0.70710678118654757f, 0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f,
                             0.38268343236508984f, 0.29028467725446239f, 0.19509032201612833f, 0.09801714032956077f};
    return t[q];
}
__host__ __device__ constexpr float sin64(int q) {
    constexpr float t[16] = {0.f, 0.09801714032956060f, 0.19509032201612825f, 0.29028467725446233f,
                             0.38268343236508978f, 0.47139673682599764f, 0.55557023301960218f, 0.63439328416364549f,
                             0.70710678118654746f, 0.77301045336273699f, 0.83146961230254524f, 0.88192126434835494f,
                             0.92387953251128674f, 0.95694033573220894f, 0.98078528040323043f, 0.99518472667219682f};
    return t[q];
}

constexpr int kPairWarps = 16;                    // warps per CTA, one CTA per SM
constexpr int kPTile = 33;                        // transpose tile stride in P units (odd: conflict-free 64-bit reads)
constexpr int kPairTileP = 2 * 16 * kPTile;       // re + im planes; the power spectrum (32 * 17 P) aliases them
constexpr int kPairMiscP = 2 * 40;                // log-mel | folded DCT input
constexpr int kPairStageP = 3 * 5 * 32;            // 15 rows of 32 float2: the samples of one frame pair, landed by TMA
constexpr int kPairWarpP = kPairTileP + kPairMiscP + kPairStageP;
static_assert(kPairTileP >= kPowFloats + 2 * kMaxFlush * 32, "power spectrum + mel partials must fit in the transpose tile");

template <int DCT>
__global__ void __launch_bounds__(kPairWarps * 32, 1) mfcc_pair_kernel(Params p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NZ = 10, HZ = 5, n_mels = 40, n_out = 40, half = 20;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- stage tables once per CTA (window scaled by 1/2: the unpack below works at twice the amplitude)
    for (int i = threadIdx.x; i < p.t.total; i += blockDim.x) {
        const float s = (i >= p.t.window2 && i < p.t.window2 + 2 * p.half_len) ? 0.5f : 1.f;
        smem[i] = p.tables[i] * s;
    }
    __syncthreads();
    const float2* s_win = reinterpret_cast<const float2*>(smem + p.t.window2);
    const float2* s_tw512 = reinterpret_cast<const float2*>(smem + p.t.tw512);
    const float2* s_tw1024 = reinterpret_cast<const float2*>(smem + p.t.tw1024);
    const float2* s_melwt = reinterpret_cast<const float2*>(smem + p.t.mel_wt);
    const int2* s_comb = reinterpret_cast<const int2*>(smem + p.t.mel_comb);
    const float* s_dcth = smem + p.t.dcth;
    const unsigned flush_mask = reinterpret_cast<const unsigned*>(smem + p.t.mel_flush)[lane];
    P* w_base = reinterpret_cast<P*>(smem + ((p.t.total + 3) & ~3)) + warp * kPairWarpP;
    P* t_re = w_base;
    P* t_im = w_base + 16 * kPTile;
    P* s_pow = w_base;                                   // aliases the tile (dead once stage 2 has read it)
    P* part_r = w_base + kPowFloats;                     // [kMaxFlush][32], in the tile's upper half (dead during mel)
    P* part_f = part_r + kMaxFlush * 32;
    P* s_lm = w_base + kPairTileP;                       // [40]
    P* s_sd = s_lm + 40;                                 // [40]
    const float2* s_stage = s_sd + 40;                   // [15][32] samples of the current pair
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + ((p.t.total + 3) & ~3) + 2 * kPairWarps * kPairWarpP) + warp;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pair_smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0;

    const int k1o = lane & 15, h = lane >> 4;
    const bool edge = (k1o == 0);                        // lanes that own the self-paired bins
    const int src_lane = ((16 - k1o) & 15) + 16 * (1 - h);
    const int kb = (edge ? 16 : k1o) + 256 * h;          // first bin this lane unpacks (own register 0, or 1 on edge lanes)
    const int k15 = edge ? 256 * h : kb + 240;           // bin of the last round
    float cb, sb;                                        // W1024^kb = cb - i sb
    {
        const float2 w = s_tw1024[kb & 255];             // (cos, -sin) of 2 pi k / 1024
        cb = h ? w.y : w.x;                              // +pi/2: cos -> -sin, sin -> cos
        sb = h ? w.x : -w.y;
    }
    const P sg = pdup(h ? -1.f : 1.f);
    float mean0 = smem[p.t.mean + lane], inv0 = smem[p.t.inv_scale + lane];
    float mean1 = 0.f, inv1 = 1.f;
    if (lane < 4 * (n_out - 32)) {
        mean1 = smem[p.t.mean + 32 + (lane >> 2)];
        inv1 = smem[p.t.inv_scale + 32 + (lane >> 2)];
    }

    const long long stride = (long long)gridDim.x * kPairWarps;
    for (long long run = (long long)blockIdx.x * kPairWarps + warp; run < p.total_runs; run += stride) {
        const long long clip = run / p.runs_per_clip;
        const int part = (int)(run - clip * p.runs_per_clip);
        const int t0 = part * p.frames_per_run;                       // frames_per_run is even
        const int t1 = min(p.frames_per_clip, t0 + p.frames_per_run);
        if (t0 >= t1) continue;
        const float2* clip_src = reinterpret_cast<const float2*>(p.wave + (size_t)clip * p.n_samples);
        // rows of 32 float2 (64 samples): frame t starts at row 5 t; the pair (t, t+1) needs rows 5t .. 5t+14
        auto issue_rows = [&](int t) {
            const bool has_b = (t + 1 < p.frames_per_clip);                // never read past the clip
            pair_tma_load(const_cast<float2*>(s_stage), clip_src + (size_t)t * (32 * HZ),
                          (has_b ? 3 * HZ : 2 * HZ) * 32 * (uint32_t)sizeof(float2), mbar);
        };
        if (lane == 0) issue_rows(t0);
        for (int t = t0; t < t1; t += 2) {
            pair_mbar_wait(mbar, phase);
            phase ^= 1u;
            PC v[16];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                const float2 w = s_win[32 * a + lane];
                const float2 xa = s_stage[32 * a + lane], xb = s_stage[32 * (a + HZ) + lane];
                v[a].re = make_float2(xa.x * w.x, xb.x * w.x);
                v[a].im = make_float2(xa.y * w.y, xb.y * w.y);
            }
            __syncwarp();
            if (lane == 0 && t + 2 < t1) issue_rows(t + 2);
#pragma unroll
            for (int a = NZ; a < 16; ++a) v[a] = {pdup(0.f), pdup(0.f)};
            // ---- radix-16 over a, W512^(lane*k1) twiddle (power tree), transpose
            pc_dft16<NZ>(v);
            {
                const float2 w1s = s_tw512[32 + lane];
                float wr[16], wi[16];                  // scalars: the packed ops broadcast them
                wr[1] = w1s.x;
                wi[1] = w1s.y;
                auto cm = [&](int a, int b, int c) {
                    wr[c] = fmaf(wr[a], wr[b], -wi[a] * wi[b]);
                    wi[c] = fmaf(wr[a], wi[b], wi[a] * wr[b]);
                };
                cm(1, 1, 2); cm(2, 1, 3); cm(2, 2, 4); cm(4, 1, 5); cm(4, 2, 6); cm(4, 3, 7); cm(4, 4, 8);
                cm(8, 1, 9); cm(8, 2, 10); cm(8, 3, 11); cm(8, 4, 12); cm(8, 5, 13); cm(8, 6, 14); cm(8, 7, 15);
                static_for<0, 16>([&](auto ic) {
                    constexpr int I = decltype(ic)::value;
                    constexpr int K1 = brev4(I);
                    PC y = v[I];
                    if constexpr (K1 != 0) y = pc_mul(y, pdup(wr[K1]), pdup(wi[K1]));
                    t_re[K1 * kPTile + lane] = y.re;
                    t_im[K1 * kPTile + lane] = y.im;
                });
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                v[m].re = t_re[k1o * kPTile + h + 2 * m];
                v[m].im = t_im[k1o * kPTile + h + 2 * m];
            }
            __syncwarp();
            // ---- radix-16 over m, W32 combine with lane ^ 16: z[Q] = Z[k1o + 16 Q + 256 h]
            pc_dft16<16>(v);
            PC z[16];
            static_for<0, 16>([&](auto ic) {
                constexpr int I = decltype(ic)::value;
                constexpr int Q = brev4(I);
                PC s = v[I];
                if constexpr (Q != 0) {
                    // lanes with h == 0 multiply by 1: pick the twiddle, not the product (2 selects instead of 4)
                    const float tc = h ? cos32(Q) : 1.f, ts = h ? sin32(Q) : 0.f;      // w = tc - i ts
                    s = pc_mul(s, pdup(tc), pdup(-ts));
                }
                PC r;
                r.re.x = __shfl_xor_sync(0xffffffffu, s.re.x, 16);
                r.re.y = __shfl_xor_sync(0xffffffffu, s.re.y, 16);
                r.im.x = __shfl_xor_sync(0xffffffffu, s.im.x, 16);
                r.im.y = __shfl_xor_sync(0xffffffffu, s.im.y, 16);
                z[Q].re = pfma(sg, s.re, r.re);
                z[Q].im = pfma(sg, s.im, r.im);
            });
            // ---- real-FFT unpack with shuffles -> power spectrum (aliases the tile)
            static_for<0, 16>([&](auto qc) {
                constexpr int Q = decltype(qc)::value;
                PC own = z[Q];
                if (edge) own = z[(Q + 1) & 15];
                PC rcv;
                rcv.re.x = __shfl_sync(0xffffffffu, z[15 - Q].re.x, src_lane);
                rcv.re.y = __shfl_sync(0xffffffffu, z[15 - Q].re.y, src_lane);
                rcv.im.x = __shfl_sync(0xffffffffu, z[15 - Q].im.x, src_lane);
                rcv.im.y = __shfl_sync(0xffffffffu, z[15 - Q].im.y, src_lane);
                // W1024^(kb + 16 Q) = (cb - i sb)(cq - i sq)
                constexpr float cq = cos64(Q), sq = sin64(Q);
                float c1 = fmaf(sb, -sq, cb * cq), s1 = fmaf(cb, sq, sb * cq);
                if constexpr (Q == 15) {
                    if (edge) {
                        rcv = own;
                        c1 = h ? 0.f : 1.f;
                        s1 = h ? 1.f : 0.f;
                    }
                }
                const P c2 = pdup(c1), s2 = pdup(s1);
                const P er = padd(own.re, rcv.re), ei = psub(own.im, rcv.im);
                const P od = padd(own.im, rcv.im), oi = psub(rcv.re, own.re);
                const P xr = pfma(c2, od, pfma(s2, oi, er));
                const P xi = pfma(c2, oi, pfma(pneg(s2), od, ei));
                const P pw = pfma(xr, xr, pmul(xi, xi));
                if constexpr (Q == 15) {
                    s_pow[k15] = pw;
                    if (lane == 0) {
                        const P d = psub(er, od);
                        s_pow[512] = pmul(d, d);
                    }
                } else {
                    s_pow[kb + 16 * Q] = pw;
                }
            });
            __syncwarp();
            // ---- mel: lane owns bins [17 lane, 17 lane + 17)
            {
                P accr = pdup(0.f), accf = pdup(0.f);
                int nfl = 0;
                const int k0 = kChunk * lane;
#pragma unroll
                for (int i = 0; i < kChunk; ++i) {
                    const P pw = s_pow[k0 + i];
                    const float2 w = s_melwt[k0 + i];
                    accr = pfma(pdup(w.x), pw, accr);        // FFMA2 with a scalar-broadcast register operand
                    accf = pfma(pdup(w.y), pw, accf);
                    const bool fl = (flush_mask >> i) & 1u;
                    if (fl) {
                        part_r[nfl * 32 + lane] = accr;
                        part_f[nfl * 32 + lane] = accf;
                        accr = pdup(0.f);
                        accf = pdup(0.f);
                        ++nfl;
                    }
                }
            }
            __syncwarp();
            P mel[2];
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd) {
                const int b = lane + 32 * rnd;
                P acc = pdup(0.f);
                if (b < n_mels) {
                    const int2 cbn = s_comb[b];
                    int n = cbn.x >> 16;
                    if (n) {
                        const int la = cbn.x & 0xff;
                        acc = part_r[((cbn.x >> 8) & 0xff) * 32 + la];
                        for (int j = 1; j < n; ++j) acc = padd(acc, part_r[la + j]);
                    }
                    n = cbn.y >> 16;
                    if (n) {
                        const int la = cbn.y & 0xff;
                        acc = padd(acc, part_f[((cbn.y >> 8) & 0xff) * 32 + la]);
                        for (int j = 1; j < n; ++j) acc = padd(acc, part_f[la + j]);
                    }
                }
                mel[rnd] = acc;
            }
            const long long fa = clip * p.frames_per_clip + t;
            float* dst_a = p.out + (size_t)fa * n_out;
            const bool has_b = (t + 1 < t1);
            if constexpr (!DCT) {
#pragma unroll
                for (int rnd = 0; rnd < 2; ++rnd) {
                    const int b = lane + 32 * rnd;
                    if (b < n_mels) {
                        const float la = 10.f * log10f(fmaxf(mel[rnd].x, p.log_floor));
                        const float lb = 10.f * log10f(fmaxf(mel[rnd].y, p.log_floor));
                        const float mu = smem[p.t.mean + b], is = smem[p.t.inv_scale + b];
                        dst_a[b] = (la - mu) * is;
                        if (has_b) dst_a[n_out + b] = (lb - mu) * is;
                    }
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int rnd = 0; rnd < 2; ++rnd) {
                    const int b = lane + 32 * rnd;
                    if (b < n_mels)
                        s_lm[b] = make_float2(10.f * log10f(fmaxf(mel[rnd].x, p.log_floor)),
                                              10.f * log10f(fmaxf(mel[rnd].y, p.log_floor)));
                }
                __syncwarp();
                // D[c][n-1-b] = (-1)^c D[c][b]: fold the input once, halve the multiply-adds
                if (lane < half) {
                    const P a = s_lm[lane], zz = s_lm[n_mels - 1 - lane];
                    s_sd[lane] = padd(a, zz);
                    s_sd[half + lane] = psub(a, zz);
                }
                __syncwarp();
                {
                    const P* sd = s_sd + (lane & 1) * half;
                    P acc = pdup(0.f);
#pragma unroll
                    for (int b = 0; b < half; b += 2) {
                        const float4 d2 = *reinterpret_cast<const float4*>(sd + b);      // broadcast 16 B
                        const float c0 = s_dcth[b * n_out + lane], c1 = s_dcth[(b + 1) * n_out + lane];
                        acc = pfma(pdup(c0), make_float2(d2.x, d2.y), acc);
                        acc = pfma(pdup(c1), make_float2(d2.z, d2.w), acc);
                    }
                    dst_a[lane] = (acc.x - mean0) * inv0;
                    if (has_b) dst_a[n_out + lane] = (acc.y - mean0) * inv0;
                }
                {
                    // the last 8 coefficients: 4 lanes per coefficient, strided terms, two shuffle rounds
                    const int c = 32 + (lane >> 2), prt = lane & 3;
                    const P* sd = s_sd + (c & 1) * half;
                    P acc = pdup(0.f);
#pragma unroll
                    for (int b = 0; b < half / 4; ++b) {
                        const P d = sd[prt + 4 * b];
                        const float cf = s_dcth[(prt + 4 * b) * n_out + c];
                        acc = pfma(pdup(cf), d, acc);
                    }
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
                    if (prt == 0) {
                        dst_a[c] = (acc.x - mean1) * inv1;
                        if (has_b) dst_a[n_out + c] = (acc.y - mean1) * inv1;
                    }
                }
                __syncwarp();
            }
        }
    }
}

// ---- host-side table construction (fp64, mirrors oracle/mfcc_ref.py)
double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

}  // namespace

struct cmoop_mfcc {
    cmoop::GenericMfcc* generic = nullptr;   // set when the configuration runs on the generic kernel
    cmoop_mfcc_config cfg;
    int n_out = 0, half_len = 0, nz = 0, mel_mode = 0, dct_mode = 0;
    Tables t{};
    std::vector<float> host_tables;
    float* d_tables = nullptr;
    size_t smem_bytes = 0;
    int sm_count = 0;
};

namespace {

int upload_tables(cmoop_mfcc* h) {
    CMOOP_CUDA_OK(cmoop::copy_sync(h->d_tables, h->host_tables.data(), h->host_tables.size() * sizeof(float),
                             cudaMemcpyHostToDevice));
    return CMOOP_OK;
}

template <int NZ, int FAST = 0>
int launch_nz(const Params& p, int grid, size_t smem, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        CMOOP_CUDA_OK(cudaFuncSetAttribute(mfcc_kernel<NZ, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    mfcc_kernel<NZ, FAST><<<grid, kWarps * 32, smem, st>>>(p);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    return CMOOP_OK;
}

// int16 PCM -> fp32 in [-1, 1) (exact: x / 32768), 8 samples per thread
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const int16_t* __restrict__ src, float* __restrict__ dst, long long n) {
    const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;
    if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(src + i) & 15u) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[2 * j] = (float)(short)(w[j] & 0xffffu) * (1.f / 32768.f);
            f[2 * j + 1] = (float)(short)(w[j] >> 16) * (1.f / 32768.f);
        }
        *reinterpret_cast<float4*>(dst + i) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
        for (long long j = i; j < n && j < i + 8; ++j) dst[j] = (float)src[j] * (1.f / 32768.f);
    }
}

template <int DCT>
int launch_pair(const Params& p, int grid, size_t smem, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        CMOOP_CUDA_OK(cudaFuncSetAttribute(mfcc_pair_kernel<DCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    mfcc_pair_kernel<DCT><<<grid, kPairWarps * 32, smem, st>>>(p);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    return CMOOP_OK;
}

}  // namespace

extern "C" {

int cmoop_mfcc_destroy(cmoop_mfcc_handle h) {
    if (!h) return CMOOP_OK;
    if (h->generic) cmoop::generic_mfcc_destroy(h->generic);
    if (h->d_tables) cudaFree(h->d_tables);
    delete h;
    return CMOOP_OK;
}

int cmoop_mfcc_create(const cmoop_mfcc_config* cfg, cmoop_mfcc_handle* out) {
    CMOOP_REQUIRE(cfg && out, "mfcc_create: null pointer");
    *out = nullptr;
    CMOOP_REQUIRE(cfg->hop > 0, "mfcc_create: hop must be positive");
    CMOOP_REQUIRE(cfg->n_mfcc >= 0 && cfg->n_mfcc <= cfg->n_mels, "mfcc_create: n_mfcc=%d outside [0,n_mels]",
                  cfg->n_mfcc);
    CMOOP_REQUIRE(cfg->sample_rate > 0 && cfg->f_min >= 0.f && cfg->f_max > cfg->f_min &&
                      cfg->f_max <= 0.5f * cfg->sample_rate,
                  "mfcc_create: need 0 <= f_min < f_max <= sample_rate/2");
    const bool specialised = cfg->n_fft == kNfft && !cfg->center && cfg->n_mels <= kMaxMel && cfg->frame_length % 2 == 0;
    if (!specialised) {
        if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
        cmoop::GenericMfcc* g = nullptr;
        int rc = cmoop::generic_mfcc_create(cfg, &g);
        if (rc != CMOOP_OK) return rc;
        cmoop_mfcc* h = new cmoop_mfcc();
        h->cfg = *cfg;
        h->generic = g;
        h->n_out = cmoop::generic_mfcc_n_out(g);
        *out = h;
        return CMOOP_OK;
    }
    CMOOP_REQUIRE(cfg->frame_length > 0 && cfg->frame_length <= kNfft,
                  "mfcc_create: frame_length=%d must be in (0,%d]", cfg->frame_length, kNfft);
    CMOOP_REQUIRE(cfg->n_mels > 0, "mfcc_create: n_mels must be positive");
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;

    cmoop_mfcc* h = new cmoop_mfcc();
    h->cfg = *cfg;
    h->n_out = cfg->n_mfcc > 0 ? cfg->n_mfcc : cfg->n_mels;
    h->half_len = cfg->frame_length / 2;
    h->nz = (h->half_len + 31) / 32;
    const int n_mels = cfg->n_mels, n_out = h->n_out;
    const double pi = 3.14159265358979323846;

    // mel filterbank (Slaney scale + area normalisation), sparse taps
    std::vector<double> hz(n_mels + 2);
    {
        const double m_lo = hz_to_mel(cfg->f_min), m_hi = hz_to_mel(cfg->f_max);
        for (int i = 0; i < n_mels + 2; ++i) hz[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    }
    std::vector<float> taps;
    std::vector<int> first(n_mels), offset(n_mels), count(n_mels);
    for (int b = 0; b < n_mels; ++b) {
        const double norm = 2.0 / (hz[b + 2] - hz[b]);
        int lo = -1, hi = -1;
        std::vector<double> w(kBins);
        for (int k = 0; k < kBins; ++k) {
            const double fk = 0.5 * cfg->sample_rate * k / (kBins - 1);
            const double rising = (fk - hz[b]) / (hz[b + 1] - hz[b]);
            const double falling = (hz[b + 2] - fk) / (hz[b + 2] - hz[b + 1]);
            double v = rising < falling ? rising : falling;
            v = v > 0.0 ? v : 0.0;
            w[k] = v * norm;
            if (v > 0.0) {
                if (lo < 0) lo = k;
                hi = k;
            }
        }
        first[b] = lo < 0 ? 0 : lo;
        count[b] = lo < 0 ? 0 : hi - lo + 1;
        offset[b] = (int)taps.size();
        for (int k = 0; k < count[b]; ++k) taps.push_back((float)w[first[b] + k]);
    }

    // mel_mode 1 tables: per-bin (rising weight for band s, falling weight for band s-1) with s the mel
    // segment [hz[s], hz[s+1]) that holds the bin; lanes own 17 consecutive bins and flush a partial at
    // every segment end.  Falls back to the per-band sparse taps (mel_mode 0) if the layout does not fit.
    std::vector<int> seg(32 * kChunk, -1);
    std::vector<float> wt(2 * 32 * kChunk, 0.f);
    for (int k = 0; k < kBins; ++k) {
        const double fk = 0.5 * cfg->sample_rate * k / (kBins - 1);
        int sfound = -1;
        for (int sidx = 0; sidx <= n_mels; ++sidx)
            if (fk >= hz[sidx] && fk < hz[sidx + 1]) sfound = sidx;
        seg[k] = sfound;
        if (sfound < 0) continue;
        const double width = hz[sfound + 1] - hz[sfound];
        if (sfound < n_mels) wt[2 * k] = (float)((fk - hz[sfound]) / width * (2.0 / (hz[sfound + 2] - hz[sfound])));
        if (sfound >= 1) wt[2 * k + 1] = (float)((hz[sfound + 1] - fk) / width * (2.0 / (hz[sfound + 1] - hz[sfound - 1])));
    }
    std::vector<unsigned> flush(32, 0u);
    // run[s] = (first lane, flush index in that lane, number of lanes) of the partials of segment s
    std::vector<int> run_lane(n_mels + 1, 0), run_k(n_mels + 1, 0), run_n(n_mels + 1, 0);
    bool mel_fast = true;
    for (int l = 0; l < 32 && mel_fast; ++l) {
        int nfl = 0;
        for (int i = 0; i < kChunk; ++i) {
            const int k = l * kChunk + i;
            const bool last = (i == kChunk - 1) || seg[k + 1] != seg[k];
            if (!last) continue;
            flush[l] |= 1u << i;
            const int sidx = seg[k];
            if (sidx >= 0) {
                if (run_n[sidx] == 0) {
                    run_lane[sidx] = l;
                    run_k[sidx] = nfl;
                } else if (nfl != 0 || run_lane[sidx] + run_n[sidx] != l) {
                    mel_fast = false;      // a continuation must be the next lane's first partial
                }
                run_n[sidx] += 1;
            }
            ++nfl;
        }
        if (nfl > kMaxFlush) mel_fast = false;
    }
    const bool dct_fast = (n_mels % 2 == 0);

    Tables& t = h->t;
    int off = 0;
    auto take = [&](int floats) {
        const int at = off;
        off += (floats + 3) & ~3;
        return at;
    };
    t.window2 = take(2 * h->half_len);
    t.tw512 = take(2 * 16 * 32);
    t.tw1024 = take(2 * 256);
    t.mel_w = t.mel_meta = t.mel_cnt = t.mel_wt = t.mel_flush = t.mel_comb = 0;
    if (mel_fast) {
        t.mel_wt = take(2 * 32 * kChunk);
        t.mel_flush = take(32);
        t.mel_comb = take(2 * n_mels);
    } else {
        t.mel_w = take((int)taps.size());
        t.mel_meta = take(2 * n_mels);
        t.mel_cnt = take(n_mels);
    }
    t.dct = t.dcth = 0;
    if (dct_fast)
        t.dcth = take((n_mels / 2) * n_out);
    else
        t.dct = take(n_mels * n_out);
    t.mean = take(n_out);
    t.inv_scale = take(n_out);
    t.total = off;
    h->mel_mode = mel_fast ? 1 : 0;
    h->dct_mode = dct_fast ? 1 : 0;
    h->host_tables.assign(off, 0.f);
    float* T = h->host_tables.data();
    for (int n = 0; n < cfg->frame_length; ++n)
        T[t.window2 + n] = (float)(0.5 - 0.5 * cos(2.0 * pi * n / cfg->frame_length));
    for (int k1 = 0; k1 < 16; ++k1)
        for (int l = 0; l < 32; ++l) {
            const double ang = -2.0 * pi * (double)(l * k1) / 512.0;
            T[t.tw512 + 2 * (k1 * 32 + l)] = (float)cos(ang);
            T[t.tw512 + 2 * (k1 * 32 + l) + 1] = (float)sin(ang);
        }
    for (int k = 0; k < 256; ++k) {
        const double ang = -2.0 * pi * k / 1024.0;
        T[t.tw1024 + 2 * k] = (float)cos(ang);
        T[t.tw1024 + 2 * k + 1] = (float)sin(ang);
    }
    if (mel_fast) {
        for (size_t i = 0; i < wt.size(); ++i) T[t.mel_wt + i] = wt[i];
        unsigned* fm = reinterpret_cast<unsigned*>(T + t.mel_flush);
        for (int l = 0; l < 32; ++l) fm[l] = flush[l];
        int* comb = reinterpret_cast<int*>(T + t.mel_comb);
        for (int b = 0; b < n_mels; ++b) {
            comb[2 * b] = run_lane[b] | (run_k[b] << 8) | (run_n[b] << 16);                    // rising: segment b
            comb[2 * b + 1] = run_lane[b + 1] | (run_k[b + 1] << 8) | (run_n[b + 1] << 16);    // falling: segment b+1
        }
    } else {
        for (size_t i = 0; i < taps.size(); ++i) T[t.mel_w + i] = taps[i];
        int* meta = reinterpret_cast<int*>(T + t.mel_meta);
        int* cnt = reinterpret_cast<int*>(T + t.mel_cnt);
        for (int b = 0; b < n_mels; ++b) {
            meta[2 * b] = first[b];
            meta[2 * b + 1] = offset[b];
            cnt[b] = count[b];
        }
    }
    for (int b = 0; b < (dct_fast ? n_mels / 2 : n_mels); ++b)
        for (int c = 0; c < n_out; ++c) {
            double v = cos(pi * c * (2 * b + 1) / (2.0 * n_mels)) * sqrt(2.0 / n_mels);
            if (c == 0) v *= sqrt(0.5);
            T[(dct_fast ? t.dcth : t.dct) + b * n_out + c] = (float)v;
        }
    for (int c = 0; c < n_out; ++c) {
        T[t.mean + c] = 0.f;
        T[t.inv_scale + c] = 1.f;
    }
    if (cudaMalloc((void**)&h->d_tables, (size_t)off * sizeof(float)) != cudaSuccess) {
        cmoop::set_error("mfcc_create: cudaMalloc failed");
        delete h;
        return CMOOP_ERR_CUDA;
    }
    int rc = upload_tables(h);
    if (rc != CMOOP_OK) {
        cmoop_mfcc_destroy(h);
        return rc;
    }
    h->smem_bytes = ((size_t)((off + 3) & ~3) + (size_t)kWarps * kWarpFloats) * sizeof(float);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    *out = h;
    return CMOOP_OK;
}

int cmoop_mfcc_n_frames(cmoop_mfcc_handle h, int n_samples) {
    if (h && h->generic) return cmoop::generic_mfcc_n_frames(h->generic, n_samples);
    if (!h || n_samples < h->cfg.frame_length) return 0;
    return 1 + (n_samples - h->cfg.frame_length) / h->cfg.hop;
}

int cmoop_mfcc_n_out(cmoop_mfcc_handle h) { return h ? h->n_out : 0; }

int cmoop_mfcc_set_standardise(cmoop_mfcc_handle h, const float* mean, const float* scale) {
    CMOOP_REQUIRE(h != nullptr, "mfcc_set_standardise: null handle");
    CMOOP_REQUIRE((mean == nullptr) == (scale == nullptr), "mfcc_set_standardise: pass both or neither");
    if (h->generic) return cmoop::generic_mfcc_set_standardise(h->generic, mean, scale);
    float* T = h->host_tables.data();
    for (int c = 0; c < h->n_out; ++c) {
        T[h->t.mean + c] = mean ? mean[c] : 0.f;
        T[h->t.inv_scale + c] = scale ? 1.f / scale[c] : 1.f;
    }
    CMOOP_CUDA_OK(cudaDeviceSynchronize());
    return upload_tables(h);
}

int cmoop_mfcc_fwd_dev(cmoop_mfcc_handle h, const float* wave, int64_t n_clips, int n_samples, float* out,
                       void* stream) {
    CMOOP_REQUIRE(h != nullptr, "mfcc_fwd: null handle");
    CMOOP_REQUIRE(n_clips >= 0 && n_samples >= 0, "mfcc_fwd: negative size");
    const int frames = cmoop_mfcc_n_frames(h, n_samples);
    if (n_clips == 0 || frames == 0) return CMOOP_OK;
    CMOOP_REQUIRE(wave && out, "mfcc_fwd: null pointer");
    if (h->generic) return cmoop::generic_mfcc_fwd(h->generic, wave, n_clips, n_samples, out, stream);
    Params p{};
    p.wave = wave;
    p.out = out;
    p.tables = h->d_tables;
    p.t = h->t;
    p.n_frames_total = (long long)n_clips * frames;
    p.n_samples = n_samples;
    p.frames_per_clip = frames;
    p.frame_length = h->cfg.frame_length;
    p.hop = h->cfg.hop;
    p.half_len = h->half_len;
    p.n_mels = h->cfg.n_mels;
    p.n_out = h->n_out;
    p.use_dct = h->cfg.n_mfcc > 0;
    p.mel_mode = h->mel_mode;
    p.dct_mode = h->dct_mode;
    p.vec2 = (n_samples % 2 == 0) && (h->cfg.hop % 2 == 0) && (((uintptr_t)wave & 7u) == 0);
    p.log_floor = h->cfg.log_floor;
    cudaStream_t st = (cudaStream_t)stream;
    const auto& c = h->cfg;
    const bool fast = c.frame_length == 640 && c.hop == 320 && c.n_mels == 40 && (c.n_mfcc == 0 || c.n_mfcc == 40) &&
                      h->mel_mode == 1 && h->dct_mode == 1 && p.vec2;
    const bool tma_ok = (((uintptr_t)wave & 15u) == 0) && (n_samples % 4 == 0);
    static const bool use_v3 = [] {
        const char* e = getenv("CMOOP_MFCC_KERNEL");       // A/B switch for profiling: "v3" = warp-per-frame kernel
        return e && strcmp(e, "v3") == 0;
    }();
    if (fast && tma_ok && !use_v3) {
        // frame-pair kernel: runs of consecutive frame pairs, about two runs per resident warp, one CTA per SM
        const long long warps = (long long)h->sm_count * kPairWarps;
        long long rpc = (2 * warps + n_clips - 1) / n_clips;
        rpc = rpc < 1 ? 1 : (rpc > frames ? frames : rpc);
        int fpr = (int)((frames + rpc - 1) / rpc);
        fpr += fpr & 1;                                     // pairs never straddle two runs
        p.frames_per_run = fpr;
        p.runs_per_clip = (frames + fpr - 1) / fpr;
        p.total_runs = (long long)n_clips * p.runs_per_clip;
        const long long blocks_needed = (p.total_runs + kPairWarps - 1) / kPairWarps;
        const int grid = (int)(blocks_needed < h->sm_count ? blocks_needed : h->sm_count);
        const size_t smem = ((size_t)((h->t.total + 3) & ~3)) * sizeof(float) + (size_t)kPairWarps * kPairWarpP * sizeof(P) + kPairWarps * sizeof(uint64_t);
        return c.n_mfcc == 0 ? launch_pair<0>(p, grid, smem, st) : launch_pair<1>(p, grid, smem, st);
    }
    if (fast) {
        // runs of consecutive frames: about two runs per resident warp, at most one clip per run
        const long long warps = (long long)h->sm_count * 2 * kWarps;
        long long rpc = (2 * warps + n_clips - 1) / n_clips;
        rpc = rpc < 1 ? 1 : (rpc > frames ? frames : rpc);
        p.frames_per_run = (int)((frames + rpc - 1) / rpc);
        p.runs_per_clip = (frames + p.frames_per_run - 1) / p.frames_per_run;
        p.total_runs = (long long)n_clips * p.runs_per_clip;
        const long long blocks_needed = (p.total_runs + kWarps - 1) / kWarps;
        const long long persistent = (long long)h->sm_count * 2;
        const int grid = (int)(blocks_needed < persistent ? blocks_needed : persistent);
        return c.n_mfcc == 0 ? launch_nz<10, 1>(p, grid, h->smem_bytes, st) : launch_nz<10, 2>(p, grid, h->smem_bytes, st);
    }
    const long long blocks_needed = (p.n_frames_total + kWarps - 1) / kWarps;
    const long long persistent = (long long)h->sm_count * 3;
    const int grid = (int)(blocks_needed < persistent ? blocks_needed : persistent);
    switch (h->nz) {
        case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 8:
            return launch_nz<8>(p, grid, h->smem_bytes, st);
        case 9: case 10:
            return launch_nz<10>(p, grid, h->smem_bytes, st);
        case 11: case 12: case 13:
            return launch_nz<13>(p, grid, h->smem_bytes, st);
        default:
            return launch_nz<16>(p, grid, h->smem_bytes, st);
    }
}

int cmoop_mfcc_fwd_host(cmoop_mfcc_handle h, const float* wave, int64_t n_clips, int n_samples, float* out) {
    CMOOP_REQUIRE(h != nullptr, "mfcc_fwd: null handle");
    CMOOP_REQUIRE(n_clips >= 0 && n_samples >= 0, "mfcc_fwd: negative size");
    const int frames = cmoop_mfcc_n_frames(h, n_samples);
    if (n_clips == 0 || frames == 0) return CMOOP_OK;
    CMOOP_REQUIRE(wave && out, "mfcc_fwd: null pointer");
    // Chunked double-buffered pipeline: H2D(i+1) overlaps kernel(i) overlaps D2H(i-1).
    const int64_t chunk = 2048;
    const size_t in_bytes = (size_t)chunk * n_samples * sizeof(float);
    const size_t out_bytes = (size_t)chunk * frames * h->n_out * sizeof(float);
    char* d = (char*)cmoop::device_scratch(4, 2 * (cmoop::align_up(in_bytes, 256) + cmoop::align_up(out_bytes, 256)));
    if (!d) return CMOOP_ERR_CUDA;
    static cudaStream_t streams[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i)
        if (!streams[i]) CMOOP_CUDA_OK(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
    float* d_in[2] = {(float*)d, (float*)(d + cmoop::align_up(in_bytes, 256))};
    float* d_out[2] = {(float*)(d + 2 * cmoop::align_up(in_bytes, 256)),
                       (float*)(d + 2 * cmoop::align_up(in_bytes, 256) + cmoop::align_up(out_bytes, 256))};
    int slot = 0;
    for (int64_t c0 = 0; c0 < n_clips; c0 += chunk, slot ^= 1) {
        const int64_t nc = (n_clips - c0) < chunk ? (n_clips - c0) : chunk;
        cudaStream_t st = streams[slot];
        CMOOP_CUDA_OK(cmoop::copy_async(d_in[slot], wave + (size_t)c0 * n_samples, (size_t)nc * n_samples * sizeof(float),
                                      cudaMemcpyHostToDevice, st));
        int rc = cmoop_mfcc_fwd_dev(h, d_in[slot], nc, n_samples, d_out[slot], st);
        if (rc != CMOOP_OK) return rc;
        CMOOP_CUDA_OK(cmoop::copy_async(out + (size_t)c0 * frames * h->n_out, d_out[slot],
                                      (size_t)nc * frames * h->n_out * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CMOOP_CUDA_OK(cudaStreamSynchronize(streams[0]));
    CMOOP_CUDA_OK(cudaStreamSynchronize(streams[1]));
    return CMOOP_OK;
}

// 16-bit PCM input (what a wav file holds): x = sample / 32768.  Device form: converted chunk-wise into library scratch on
// `stream`, then the fp32 kernels run unchanged.  Host form: half the PCIe bytes of the fp32 entry point.
int cmoop_mfcc_fwd_dev_i16(cmoop_mfcc_handle h, const int16_t* wave, int64_t n_clips, int n_samples, float* out, void* stream) {
    CMOOP_REQUIRE(h != nullptr, "mfcc_fwd: null handle");
    CMOOP_REQUIRE(n_clips >= 0 && n_samples >= 0, "mfcc_fwd: negative size");
    const int frames = cmoop_mfcc_n_frames(h, n_samples);
    if (n_clips == 0 || frames == 0) return CMOOP_OK;
    CMOOP_REQUIRE(wave && out, "mfcc_fwd: null pointer");
    // The widened chunk lives in a STREAM-ORDERED allocation of the caller's stream (cudaMallocAsync: served from the
    // device's memory pool, no synchronisation), not in process-global scratch: two calls on different streams -- or from
    // two front-end handles, e.g. the train and validation splits on separate torch streams -- never share a buffer.
    const int64_t chunk = n_clips < 4096 ? n_clips : 4096;
    cudaStream_t st = (cudaStream_t)stream;
    float* d_f = nullptr;
    CMOOP_CUDA_OK(cudaMallocAsync((void**)&d_f, (size_t)chunk * n_samples * sizeof(float), st));
    int rc = CMOOP_OK;
    for (int64_t c0 = 0; c0 < n_clips && rc == CMOOP_OK; c0 += chunk) {
        const int64_t nc = (n_clips - c0) < chunk ? (n_clips - c0) : chunk;
        const long long n = (long long)nc * n_samples;
        pcm16_to_f32_kernel<<<(unsigned)((n + 2047) / 2048), 256, 0, st>>>(wave + (size_t)c0 * n_samples, d_f, n);
        cmoop::count_launch();
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cmoop::set_error("pcm16_to_f32_kernel: %s", cudaGetErrorString(e));
            rc = CMOOP_ERR_CUDA;
            break;
        }
        rc = cmoop_mfcc_fwd_dev(h, d_f, nc, n_samples, out + (size_t)c0 * frames * h->n_out, stream);
    }
    (void)cudaFreeAsync(d_f, st);
    return rc;
}

int cmoop_mfcc_fwd_host_i16(cmoop_mfcc_handle h, const int16_t* wave, int64_t n_clips, int n_samples, float* out) {
    CMOOP_REQUIRE(h != nullptr, "mfcc_fwd: null handle");
    CMOOP_REQUIRE(n_clips >= 0 && n_samples >= 0, "mfcc_fwd: negative size");
    const int frames = cmoop_mfcc_n_frames(h, n_samples);
    if (n_clips == 0 || frames == 0) return CMOOP_OK;
    CMOOP_REQUIRE(wave && out, "mfcc_fwd: null pointer");
    // same double-buffered pipeline as the fp32 form; the staged bytes are int16 and widened on the device
    const int64_t chunk = 2048;
    const size_t pcm_bytes = cmoop::align_up((size_t)chunk * n_samples * sizeof(int16_t), 256);
    const size_t in_bytes = cmoop::align_up((size_t)chunk * n_samples * sizeof(float), 256);
    const size_t out_bytes = cmoop::align_up((size_t)chunk * frames * h->n_out * sizeof(float), 256);
    char* d = (char*)cmoop::device_scratch(6, 2 * (pcm_bytes + in_bytes + out_bytes));
    if (!d) return CMOOP_ERR_CUDA;
    static cudaStream_t streams[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i)
        if (!streams[i]) CMOOP_CUDA_OK(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
    int slot = 0;
    for (int64_t c0 = 0; c0 < n_clips; c0 += chunk, slot ^= 1) {
        const int64_t nc = (n_clips - c0) < chunk ? (n_clips - c0) : chunk;
        cudaStream_t st = streams[slot];
        int16_t* d_pcm = (int16_t*)(d + slot * (pcm_bytes + in_bytes + out_bytes));
        float* d_in = (float*)((char*)d_pcm + pcm_bytes);
        float* d_out = (float*)((char*)d_in + in_bytes);
        const long long n = (long long)nc * n_samples;
        CMOOP_CUDA_OK(cmoop::copy_async(d_pcm, wave + (size_t)c0 * n_samples, (size_t)n * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        pcm16_to_f32_kernel<<<(unsigned)((n + 2047) / 2048), 256, 0, st>>>(d_pcm, d_in, n);
        cmoop::count_launch();
        CMOOP_CUDA_OK(cudaGetLastError());
        int rc = cmoop_mfcc_fwd_dev(h, d_in, nc, n_samples, d_out, st);
        if (rc != CMOOP_OK) return rc;
        CMOOP_CUDA_OK(cmoop::copy_async(out + (size_t)c0 * frames * h->n_out, d_out, (size_t)nc * frames * h->n_out * sizeof(float),
                                      cudaMemcpyDeviceToHost, st));
    }
    CMOOP_CUDA_OK(cudaStreamSynchronize(streams[0]));
    CMOOP_CUDA_OK(cudaStreamSynchronize(streams[1]));
    return CMOOP_OK;
}

}  // extern "C"
