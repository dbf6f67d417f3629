// Generic log-mel / MFCC path: any power-of-two n_fft in [64, 4096], optional librosa-style centring
// (reflect padding), up to 256 mel bands -- e.g. the BirdCLEF-shaped front-end of BASELINE configs[3]
// (32 kHz, n_fft 2048, hop 512, centred, 128 mel x 313 frames for a 5 s clip).  Same spec as the
// specialised kernel (oracle/mfcc_ref.py); one CTA per frame:
//   window -> n_fft/2-point complex Stockham radix-2 FFT in shared memory (ping-pong buffers) -> real-FFT
//   unpack -> power -> sparse triangular mel (thread per band) -> 10 log10 -> [DCT-II] -> [standardise].
// This path favours generality over speed; the 1024-point / 40-band configuration the benchmark measures
// runs in mfcc.cu.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "mfcc_generic.cuh"

namespace cmoop {

namespace {

constexpr int kThreads = 256;

struct GParams {
    const float* wave;
    float* out;
    const float* window;    // [frame_length]
    const float* mel_w;     // sparse taps
    const int* mel_first;   // [n_mels]
    const int* mel_off;     // [n_mels]
    const int* mel_cnt;     // [n_mels]
    const float* dct;       // [n_out][n_mels] (row per coefficient) or null
    const float* mean;      // [n_out]
    const float* inv_scale; // [n_out]
    long long n_frames_total;
    int n_samples, frames_per_clip, frame_length, hop, n_fft, n_mels, n_out, center;
    float log_floor;
};

__device__ __forceinline__ int reflect(int s, int n) {
    if (s < 0) s = -s;
    if (s >= n) s = 2 * (n - 1) - s;
    return s;
}

__global__ void __launch_bounds__(kThreads) mfcc_generic_kernel(GParams p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.n_fft / 2;                 // complex FFT length
    float2* bufa = reinterpret_cast<float2*>(smem);
    float2* bufb = bufa + M;
    float* pw = reinterpret_cast<float*>(bufb + M);   // [M + 1]
    float* lm = pw + M + 1;                            // [n_mels]
    const int tid = threadIdx.x;
    for (long long f = blockIdx.x; f < p.n_frames_total; f += gridDim.x) {
        const long long clip = f / p.frames_per_clip;
        const int t = (int)(f - clip * p.frames_per_clip);
        const float* src = p.wave + (size_t)clip * p.n_samples;
        const int start = t * p.hop - (p.center ? p.n_fft / 2 : 0);
        for (int n = tid; n < M; n += kThreads) {
            float2 z = make_float2(0.f, 0.f);
            const int i0 = 2 * n, i1 = 2 * n + 1;
            if (i0 < p.frame_length) {
                const int s0 = p.center ? reflect(start + i0, p.n_samples) : start + i0;
                z.x = __ldg(src + s0) * __ldg(p.window + i0);
            }
            if (i1 < p.frame_length) {
                const int s1 = p.center ? reflect(start + i1, p.n_samples) : start + i1;
                z.y = __ldg(src + s1) * __ldg(p.window + i1);
            }
            bufa[n] = z;
        }
        __syncthreads();
        // Stockham autosort, decimation in frequency: natural order in, natural order out
        float2* x = bufa;
        float2* y = bufb;
        for (int l = M >> 1, m = 1; l >= 1; l >>= 1, m <<= 1) {
            for (int q = tid; q < (M >> 1); q += kThreads) {
                const int j = q / m, k = q - j * m;
                float sn, cs;
                sincospif(-(float)j / (float)l, &sn, &cs);
                const float2 c0 = x[k + j * m], c1 = x[k + j * m + l * m];
                const float dr = c0.x - c1.x, di = c0.y - c1.y;
                y[k + 2 * j * m] = make_float2(c0.x + c1.x, c0.y + c1.y);
                y[k + 2 * j * m + m] = make_float2(dr * cs - di * sn, dr * sn + di * cs);
            }
            __syncthreads();
            float2* tmp = x;
            x = y;
            y = tmp;
        }
        // real-FFT unpack: X[k] = E + W^k O, X[M-k] = conj(E - W^k O)
        for (int k = tid; k <= (M >> 1); k += kThreads) {
            const int kb = (M - k) & (M - 1);
            const float2 a = x[k], bq = x[kb];
            const float br = bq.x, bi = -bq.y;
            const float er = 0.5f * (a.x + br), ei = 0.5f * (a.y + bi);
            const float orr = 0.5f * (a.y - bi), oi = -0.5f * (a.x - br);
            float sn, cs;
            sincospif(-2.f * (float)k / (float)p.n_fft, &sn, &cs);
            const float pr = orr * cs - oi * sn, pi = orr * sn + oi * cs;
            const float xr = er + pr, xi = ei + pi, yr = er - pr, yi = ei - pi;
            pw[k] = xr * xr + xi * xi;
            pw[M - k] = yr * yr + yi * yi;
        }
        __syncthreads();
        for (int b = tid; b < p.n_mels; b += kThreads) {
            const int first = p.mel_first[b], off = p.mel_off[b], cnt = p.mel_cnt[b];
            float acc = 0.f;
            for (int i = 0; i < cnt; ++i) acc = fmaf(__ldg(p.mel_w + off + i), pw[first + i], acc);
            lm[b] = 10.f * log10f(fmaxf(acc, p.log_floor));
        }
        __syncthreads();
        float* dst = p.out + (size_t)f * p.n_out;
        for (int c = tid; c < p.n_out; c += kThreads) {
            float acc;
            if (p.dct) {
                acc = 0.f;
                const float* row = p.dct + (size_t)c * p.n_mels;
                for (int b = 0; b < p.n_mels; ++b) acc = fmaf(__ldg(row + b), lm[b], acc);
            } else {
                acc = lm[c];
            }
            dst[c] = (acc - p.mean[c]) * p.inv_scale[c];
        }
        __syncthreads();
    }
}

double g_hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double g_mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

}  // namespace

struct GenericMfcc {
    cmoop_mfcc_config cfg;
    int n_out = 0, sm_count = 0;
    float* d_float = nullptr;   // window | taps | dct | mean | inv_scale
    int* d_int = nullptr;       // first | off | cnt
    size_t o_taps = 0, o_dct = 0, o_mean = 0, o_inv = 0;
    std::vector<float> h_float;
};

int generic_mfcc_create(const cmoop_mfcc_config* cfg, GenericMfcc** out) {
    const int n_fft = cfg->n_fft;
    CMOOP_REQUIRE(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "mfcc_create: n_fft=%d must be a power of two in [64,4096]", n_fft);
    CMOOP_REQUIRE(cfg->frame_length > 0 && cfg->frame_length <= n_fft, "mfcc_create: frame_length=%d outside (0,n_fft]", cfg->frame_length);
    CMOOP_REQUIRE(!cfg->center || cfg->frame_length == n_fft, "mfcc_create: center=1 needs frame_length == n_fft");
    CMOOP_REQUIRE(cfg->n_mels > 0 && cfg->n_mels <= 256, "mfcc_create: n_mels=%d outside [1,256]", cfg->n_mels);
    GenericMfcc* g = new GenericMfcc();
    g->cfg = *cfg;
    g->n_out = cfg->n_mfcc > 0 ? cfg->n_mfcc : cfg->n_mels;
    const int n_mels = cfg->n_mels, bins = n_fft / 2 + 1;
    const double pi = 3.14159265358979323846;
    std::vector<double> hz(n_mels + 2);
    const double m_lo = g_hz_to_mel(cfg->f_min), m_hi = g_hz_to_mel(cfg->f_max);
    for (int i = 0; i < n_mels + 2; ++i) hz[i] = g_mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<float> taps;
    std::vector<int> ints(3 * n_mels);
    for (int b = 0; b < n_mels; ++b) {
        const double norm = 2.0 / (hz[b + 2] - hz[b]);
        int lo = -1, hi = -1;
        std::vector<double> w(bins);
        for (int k = 0; k < bins; ++k) {
            const double fk = 0.5 * cfg->sample_rate * k / (bins - 1);
            const double rising = (fk - hz[b]) / (hz[b + 1] - hz[b]), falling = (hz[b + 2] - fk) / (hz[b + 2] - hz[b + 1]);
            double v = rising < falling ? rising : falling;
            v = v > 0.0 ? v : 0.0;
            w[k] = v * norm;
            if (v > 0.0) {
                if (lo < 0) lo = k;
                hi = k;
            }
        }
        ints[b] = lo < 0 ? 0 : lo;
        ints[n_mels + b] = (int)taps.size();
        ints[2 * n_mels + b] = lo < 0 ? 0 : hi - lo + 1;
        for (int k = 0; k < ints[2 * n_mels + b]; ++k) taps.push_back((float)w[ints[b] + k]);
    }
    std::vector<float>& F = g->h_float;
    for (int n = 0; n < cfg->frame_length; ++n) F.push_back((float)(0.5 - 0.5 * cos(2.0 * pi * n / cfg->frame_length)));
    g->o_taps = F.size();
    F.insert(F.end(), taps.begin(), taps.end());
    g->o_dct = F.size();
    if (cfg->n_mfcc > 0)
        for (int c = 0; c < g->n_out; ++c)
            for (int b = 0; b < n_mels; ++b) {
                double v = cos(pi * c * (2 * b + 1) / (2.0 * n_mels)) * sqrt(2.0 / n_mels);
                if (c == 0) v *= sqrt(0.5);
                F.push_back((float)v);
            }
    g->o_mean = F.size();
    F.insert(F.end(), g->n_out, 0.f);
    g->o_inv = F.size();
    F.insert(F.end(), g->n_out, 1.f);
    if (cudaMalloc((void**)&g->d_float, F.size() * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void**)&g->d_int, ints.size() * sizeof(int)) != cudaSuccess ||
        cmoop::copy_sync(g->d_float, F.data(), F.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cmoop::copy_sync(g->d_int, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("mfcc_create: device allocation failed");
        generic_mfcc_destroy(g);
        return CMOOP_ERR_CUDA;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, dev);
    *out = g;
    return CMOOP_OK;
}

void generic_mfcc_destroy(GenericMfcc* g) {
    if (!g) return;
    if (g->d_float) cudaFree(g->d_float);
    if (g->d_int) cudaFree(g->d_int);
    delete g;
}

int generic_mfcc_n_out(const GenericMfcc* g) { return g->n_out; }

int generic_mfcc_n_frames(const GenericMfcc* g, int n_samples) {
    if (g->cfg.center) return n_samples > g->cfg.n_fft / 2 ? 1 + n_samples / g->cfg.hop : 0;   // reflect padding needs n > n_fft/2
    return n_samples < g->cfg.frame_length ? 0 : 1 + (n_samples - g->cfg.frame_length) / g->cfg.hop;
}

int generic_mfcc_set_standardise(GenericMfcc* g, const float* mean, const float* scale) {
    for (int c = 0; c < g->n_out; ++c) {
        g->h_float[g->o_mean + c] = mean ? mean[c] : 0.f;
        g->h_float[g->o_inv + c] = scale ? 1.f / scale[c] : 1.f;
    }
    CMOOP_CUDA_OK(cudaDeviceSynchronize());
    CMOOP_CUDA_OK(cmoop::copy_sync(g->d_float + g->o_mean, g->h_float.data() + g->o_mean, 2 * g->n_out * sizeof(float),
                             cudaMemcpyHostToDevice));
    return CMOOP_OK;
}

int generic_mfcc_fwd(GenericMfcc* g, const float* wave, int64_t n_clips, int n_samples, float* out, void* stream) {
    const int frames = generic_mfcc_n_frames(g, n_samples);
    if (n_clips == 0 || frames == 0) return CMOOP_OK;
    GParams p{};
    p.wave = wave;
    p.out = out;
    p.window = g->d_float;
    p.mel_w = g->d_float + g->o_taps;
    p.mel_first = g->d_int;
    p.mel_off = g->d_int + g->cfg.n_mels;
    p.mel_cnt = g->d_int + 2 * g->cfg.n_mels;
    p.dct = g->cfg.n_mfcc > 0 ? g->d_float + g->o_dct : nullptr;
    p.mean = g->d_float + g->o_mean;
    p.inv_scale = g->d_float + g->o_inv;
    p.n_frames_total = (long long)n_clips * frames;
    p.n_samples = n_samples;
    p.frames_per_clip = frames;
    p.frame_length = g->cfg.frame_length;
    p.hop = g->cfg.hop;
    p.n_fft = g->cfg.n_fft;
    p.n_mels = g->cfg.n_mels;
    p.n_out = g->n_out;
    p.center = g->cfg.center;
    p.log_floor = g->cfg.log_floor;
    const size_t smem = (size_t)(2 * (g->cfg.n_fft / 2) * 2 + g->cfg.n_fft / 2 + 1 + g->cfg.n_mels + 3) * sizeof(float);
    const long long persistent = (long long)g->sm_count * 4;
    const int grid = (int)(p.n_frames_total < persistent ? p.n_frames_total : persistent);
    mfcc_generic_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    return CMOOP_OK;
}

}  // namespace cmoop
