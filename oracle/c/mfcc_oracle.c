/* Plain-C fp64 restatement of oracle/mfcc_ref.py (the front-end spec).  TEST INFRASTRUCTURE ONLY:
 * used as the multi-threaded CPU baseline by bench.py (cpu_baseline / --impl reference) and checked
 * against the NumPy oracle in tests/test_oracle_c.py.  The reference has no feature code
 * (nsga_penalty.py:64-71 loads pre-computed .npy), so this is a "port" baseline of the declared spec:
 * frames of frame_length @ hop, periodic Hann, |rfft(n_fft)|^2, Slaney mel, 10*log10, DCT-II ortho.
 * OpenMP over clips; real FFT = half-length complex radix-2 FFT + unpack. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 15.0, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 15.0, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static void fft_inplace(double* re, double* im, int n, const double* cs, const double* sn) {
    for (int i = 1, j = 0; i < n; ++i) {          /* bit reversal */
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const double wr = cs[k * step], wi = -sn[k * step];
                const int a = i + k, b = i + k + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
    }
}

int cmoop_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* out [n_clips][n_frames][n_out] fp64; returns n_frames (or <0 on bad arguments) */
long cmoop_oracle_mfcc(const float* wave, long n_clips, int n_samples, int sample_rate, int frame_length, int hop,
                       int n_fft, int n_mels, int n_mfcc, double f_min, double f_max, double log_floor, double* out) {
    if (n_fft <= 0 || (n_fft & (n_fft - 1)) || frame_length > n_fft || frame_length <= 0 || hop <= 0) return -1;
    const int half = n_fft / 2, bins = half + 1;
    const int n_out = n_mfcc > 0 ? n_mfcc : n_mels;
    const long frames = n_samples < frame_length ? 0 : 1 + (n_samples - frame_length) / hop;
    if (frames == 0 || n_clips == 0) return frames;
    double* win = malloc(sizeof(double) * frame_length);
    double* cs = malloc(sizeof(double) * half);
    double* sn = malloc(sizeof(double) * half);
    double* ucs = malloc(sizeof(double) * bins);
    double* usn = malloc(sizeof(double) * bins);
    double* fb = calloc((size_t)n_mels * bins, sizeof(double));
    int* lo = malloc(sizeof(int) * n_mels);
    int* hi = malloc(sizeof(int) * n_mels);
    double* dct = malloc(sizeof(double) * n_out * n_mels);
    double* hz = malloc(sizeof(double) * (n_mels + 2));
    for (int n = 0; n < frame_length; ++n) win[n] = 0.5 - 0.5 * cos(2.0 * PI * n / frame_length);
    for (int k = 0; k < half; ++k) { cs[k] = cos(2.0 * PI * k / half); sn[k] = sin(2.0 * PI * k / half); }
    for (int k = 0; k < bins; ++k) { ucs[k] = cos(2.0 * PI * k / n_fft); usn[k] = sin(2.0 * PI * k / n_fft); }
    const double m_lo = hz_to_mel(f_min), m_hi = hz_to_mel(f_max);
    for (int i = 0; i < n_mels + 2; ++i) hz[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    for (int b = 0; b < n_mels; ++b) {
        lo[b] = bins; hi[b] = -1;
        const double norm = 2.0 / (hz[b + 2] - hz[b]);
        for (int k = 0; k < bins; ++k) {
            const double fk = 0.5 * sample_rate * k / (bins - 1);
            const double up = (fk - hz[b]) / (hz[b + 1] - hz[b]), down = (hz[b + 2] - fk) / (hz[b + 2] - hz[b + 1]);
            double v = up < down ? up : down;
            if (v > 0.0) { fb[(size_t)b * bins + k] = v * norm; if (k < lo[b]) lo[b] = k; hi[b] = k; }
        }
    }
    for (int c = 0; c < n_out; ++c)
        for (int b = 0; b < n_mels; ++b) {
            double v = cos(PI * c * (2 * b + 1) / (2.0 * n_mels)) * sqrt(2.0 / n_mels);
            if (c == 0) v *= sqrt(0.5);
            dct[c * n_mels + b] = v;
        }
#pragma omp parallel
    {
        double* re = malloc(sizeof(double) * half);
        double* im = malloc(sizeof(double) * half);
        double* pw = malloc(sizeof(double) * bins);
        double* lm = malloc(sizeof(double) * n_mels);
#pragma omp for schedule(static)
        for (long c = 0; c < n_clips; ++c) {
            for (long t = 0; t < frames; ++t) {
                const float* src = wave + (size_t)c * n_samples + (size_t)t * hop;
                for (int n = 0; n < half; ++n) {
                    const int i0 = 2 * n, i1 = 2 * n + 1;
                    re[n] = i0 < frame_length ? (double)src[i0] * win[i0] : 0.0;
                    im[n] = i1 < frame_length ? (double)src[i1] * win[i1] : 0.0;
                }
                fft_inplace(re, im, half, cs, sn);
                for (int k = 0; k <= half / 2; ++k) {
                    const int kb = (half - k) % half;
                    const double ar = re[k], ai = im[k], br = re[kb], bi = -im[kb];
                    const double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);
                    const double orr = 0.5 * (ai - bi), oi = -0.5 * (ar - br);
                    const double wr = ucs[k], wi = -usn[k];
                    const double pr = orr * wr - oi * wi, pi_ = orr * wi + oi * wr;
                    const double xr = er + pr, xi = ei + pi_, yr = er - pr, yi = ei - pi_;
                    pw[k] = xr * xr + xi * xi;
                    pw[half - k] = yr * yr + yi * yi;
                }
                for (int b = 0; b < n_mels; ++b) {
                    double acc = 0.0;
                    for (int k = lo[b]; k <= hi[b]; ++k) acc += fb[(size_t)b * bins + k] * pw[k];
                    lm[b] = 10.0 * log10(acc > log_floor ? acc : log_floor);
                }
                double* dst = out + ((size_t)c * frames + t) * n_out;
                if (n_mfcc > 0) {
                    for (int q = 0; q < n_out; ++q) {
                        double acc = 0.0;
                        for (int b = 0; b < n_mels; ++b) acc += dct[q * n_mels + b] * lm[b];
                        dst[q] = acc;
                    }
                } else {
                    memcpy(dst, lm, sizeof(double) * n_mels);
                }
            }
        }
        free(re); free(im); free(pw); free(lm);
    }
    free(win); free(cs); free(sn); free(ucs); free(usn); free(fb); free(lo); free(hi); free(dct); free(hz);
    return frames;
}
