// Stem convolution (Cin = 1) of the candidate CNNs: forward and weight gradient as dedicated HBM-bound kernels.
//
// The first Conv2D of both model families (nsga_penalty.py:255, sa_nsga_penalty.py:150) reads the 1-channel
// feature map and writes the largest activation of the network ([64*H*W][F] fp32); its GEMM has K = k*k (+1 bias)
// = 10 or 26, so the generic 64x64x16 SIMT GEMM spent its time on im2col bookkeeping (1.0 ms forward / 1.8 ms
// weight gradient per grouped launch of 32 candidates against ~0.1 ms of HBM traffic).  Here a block owns 1024
// consecutive output pixels of one candidate: the feature-map rows they touch (<= 2 samples) are staged once in
// shared memory with their zero padding, a thread owns 4 (k = 3) or 2 (k = 5) output channels of a pixel with their weights in registers, taps are immediate
// offsets into the staged tile, stores / gradient loads are 16-byte and coalesced.  Contracts are those of
// conv_gemm_kernel / conv_wgrad_kernel (ConvTask / WgradTask): same BN partial-sum layout (64-row tiles), same
// [split][K+1][Cout] partial gradients reduced by reduce_kernel, deterministic (no atomics).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

struct Segs {
    int n_first, ha[2];
};

// Stages the weights-independent part: the rows of the (<= 2) samples touched by output pixels [m0, m1).
template <int KS>
__device__ __forceinline__ void stage_rows(float* tile, int seg_floats, const float* xbase, const int* gather, int H, int W,
                                           int m0, int m1, Segs& sg) {
    constexpr int P = (KS - 1) / 2;
    const int HW = H * W, Wp = W + 2 * P;
    sg.n_first = m0 / HW;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int n = sg.n_first + s;
        const int lo = max(m0, n * HW), hi = min(m1, (n + 1) * HW);
        sg.ha[s] = 0;
        if (lo >= hi) continue;
        const int ha = (lo - n * HW) / W, hb = (hi - 1 - n * HW) / W;
        sg.ha[s] = ha;
        const float* src = xbase + (long long)(gather ? gather[n] : n) * HW;
        float* dst = tile + s * seg_floats;
        const int cnt = (hb - ha + 1 + 2 * P) * Wp;
        for (int i = threadIdx.x; i < cnt; i += kThreads) {
            const int r = i / Wp, c = i - r * Wp;
            const int hi_ = ha - P + r, wi_ = c - P;
            float v = 0.f;
            if ((unsigned)hi_ < (unsigned)H && (unsigned)wi_ < (unsigned)W) v = __ldg(src + hi_ * W + wi_);
            dst[i] = v;
        }
    }
}

// y[m][co] = relu?(b[co] + sum_taps x[m + tap] w[tap][co]);  BN partial sums per 64-row tile.
// A thread keeps the weights of its CPT channels for every tap in registers (its channel group never changes), so a
// tap costs one shared-memory read of x (broadcast among the threads of a pixel) and CPT FMAs.
template <int KS, int CPT>
__device__ __forceinline__ void stem_conv_body(const ConvTask& T, float* sm, int chunk, int rows_per_block, int seg_floats,
                                               int n_b, int step) {
    constexpr int P = (KS - 1) / 2, TAPS = KS * KS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = T.H, W = T.W, HW = H * W, Wp = W + 2 * P, Cout = T.Cout;
    const int M = n_b * HW, m0 = chunk * rows_per_block, m1 = min(M, m0 + rows_per_block);
    if (m0 >= M) return;
    float* tile = sm;                                 // [2][seg_floats]
    Segs sg;
    stage_rows<KS>(tile, seg_floats, T.x + T.x_step * step, T.gather ? T.gather + T.gather_step * step : nullptr, H, W, m0,
                   m1, sg);
    // A warp owns whole 64-row BN tiles (tiles warp, warp + 8 of the block's 16), so the per-tile statistics are a
    // warp-shuffle reduction and the main loop has no block barrier.
    const int CG = Cout / CPT, PW = 32 / CG;          // threads per pixel (<= 32), pixels per warp pass
    const int cg = lane % CG, pl = lane / CG;
    // weights as packed pairs: one FFMA2 (x broadcast) updates two channels
    float2 wr[TAPS + 1][CPT / 2];
#pragma unroll
    for (int t = 0; t <= TAPS; ++t)
#pragma unroll
        for (int j = 0; j < CPT / 2; ++j)
            wr[t][j] = __ldg(reinterpret_cast<const float2*>(T.w + t * Cout + cg * CPT) + j);
    __syncthreads();

    const int tiles = (m1 - m0 + 63) >> 6;
    for (int tl = warp; tl < tiles; tl += kThreads / 32) {
        float2 s1[CPT / 2], s2[CPT / 2];
#pragma unroll
        for (int j = 0; j < CPT / 2; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
        // (sample, row, column) of this thread's first pixel of the tile, then advanced by PW pixels per pass
        int m = m0 + tl * 64 + pl;
        int n = m / HW, rem = m - n * HW;
        int h = rem / W, w = rem - h * W;
        for (int r = pl; r < 64; r += PW, m += PW) {
            if (m < m1) {
                const int s = n - sg.n_first;
                const float* tp = tile + s * seg_floats + (h - sg.ha[s]) * Wp + w;
                float2 acc[CPT / 2];
#pragma unroll
                for (int j = 0; j < CPT / 2; ++j) acc[j] = wr[TAPS][j];
#pragma unroll
                for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                    for (int kw = 0; kw < KS; ++kw) {
                        const float x = tp[kh * Wp + kw];
#pragma unroll
                        for (int j = 0; j < CPT / 2; ++j) acc[j] = __ffma2_rn(make_float2(x, x), wr[kh * KS + kw][j], acc[j]);
                    }
                if (T.relu) {
#pragma unroll
                    for (int j = 0; j < CPT / 2; ++j) acc[j] = make_float2(fmaxf(acc[j].x, 0.f), fmaxf(acc[j].y, 0.f));
                }
                const long long o = (long long)m * Cout + cg * CPT;
                if (T.y) {
                    if constexpr (CPT == 4) {
                        *reinterpret_cast<float4*>(T.y + o) = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
                    } else {
                        *reinterpret_cast<float2*>(T.y + o) = acc[0];
                    }
                }
                if (T.yh) {
                    __nv_bfloat162 hb[CPT / 2];
#pragma unroll
                    for (int j = 0; j < CPT / 2; ++j) hb[j] = __floats2bfloat162_rn(acc[j].x, acc[j].y);
                    if constexpr (CPT == 4) {
                        *reinterpret_cast<uint2*>(T.yh + o) = make_uint2(*reinterpret_cast<uint32_t*>(&hb[0]), *reinterpret_cast<uint32_t*>(&hb[1]));
                    } else {
                        *reinterpret_cast<__nv_bfloat162*>(T.yh + o) = hb[0];
                    }
                    if (!T.y) {                       // bf16-only output: the BN statistics are those of the stored values
#pragma unroll
                        for (int j = 0; j < CPT / 2; ++j) acc[j] = __bfloat1622float2(hb[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < CPT / 2; ++j) {
                    s1[j] = __fadd2_rn(s1[j], acc[j]);
                    s2[j] = __ffma2_rn(acc[j], acc[j], s2[j]);
                }
            }
            w += PW;                                      // PW <= 8 < W: at most one row wrap per pass
            if (w >= W) {
                w -= W;
                if (++h == H) { h = 0; ++n; }
            }
        }
        if (T.stat_part) {
            __syncwarp();
            float* dst = T.stat_part + (long long)((m0 >> 6) + tl) * 2 * Cout + cg * CPT;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                float a1 = (j & 1) ? s1[j >> 1].y : s1[j >> 1].x, a2 = (j & 1) ? s2[j >> 1].y : s2[j >> 1].x;
                for (int msk = CG; msk < 32; msk <<= 1) {
                    a1 += __shfl_xor_sync(0xffffffffu, a1, msk);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, msk);
                }
                if (lane < CG) {
                    dst[j] = a1;
                    dst[Cout + j] = a2;
                }
            }
        }
    }
}

// the two kernel sizes of the genotype space (nsga_penalty.py:189) share one grouped launch
__global__ void __launch_bounds__(kThreads, 3) stem_conv_kernel(const ConvTask* __restrict__ tasks, int blocks_per_task,
                                                             int rows_per_block, int seg_floats, int n_b, int step) {
    extern __shared__ __align__(16) float sm[];
    __shared__ ConvTask T;
    const int task = blockIdx.x / blocks_per_task, chunk = blockIdx.x - task * blocks_per_task;
    if (threadIdx.x == 0) T = tasks[task];
    __syncthreads();
    if (T.k == 3)
        stem_conv_body<3, 4>(T, sm, chunk, rows_per_block, seg_floats, n_b, step);
    else
        stem_conv_body<5, 2>(T, sm, chunk, rows_per_block, seg_floats, n_b, step);
}

// out[split][tap][co] = sum_{m in split} x[m + tap] dy[m][co]  (row TAPS = bias gradient = sum of dy)
template <int KS, int CPT>
__device__ __forceinline__ void stem_wgrad_body(const WgradTask& T, float* sm, int split, int seg_floats, int n_b, int step) {
    constexpr int P = (KS - 1) / 2, TAPS = KS * KS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = T.H, W = T.W, HW = H * W, Wp = W + 2 * P, Cout = T.Cout;
    const int M = n_b * HW, m0 = split * T.m_chunk, m1 = min(M, m0 + T.m_chunk);
    float* tile = sm;                                 // [2][seg_floats]
    float* red = tile + 2 * seg_floats;               // [8 warps][(TAPS+1)][Cout]
    const int n_out = (TAPS + 1) * Cout;
    float* out = T.out + (long long)split * n_out;
    if (m0 >= M) {                                    // empty split: its partial must still be defined
        for (int i = tid; i < n_out; i += kThreads) out[i] = 0.f;
        return;
    }
    Segs sg;
    stage_rows<KS>(tile, seg_floats, T.x + T.x_step * step, T.gather ? T.gather + T.gather_step * step : nullptr, H, W, m0,
                   m1, sg);
    __syncthreads();

    const int CG = Cout / CPT, PP = kThreads / CG;
    const int cg = tid % CG, pl = tid / CG;
    float acc[TAPS + 1][CPT];
#pragma unroll
    for (int t = 0; t <= TAPS; ++t)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[t][j] = 0.f;
    constexpr int UN = 4;                             // gradient loads in flight per thread (HBM latency)
    for (int mb = m0 + pl; mb < m1; mb += UN * PP) {
        float g[UN][CPT];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int m = mb + u * PP;
#pragma unroll
            for (int j = 0; j < CPT; ++j) g[u][j] = 0.f;
            if (m < m1) {
                if (T.dyh) {                          // bf16-only gradient (precision bf16)
                    if constexpr (CPT == 4) {
                        const uint2 r = *reinterpret_cast<const uint2*>(T.dyh + (long long)m * Cout + cg * 4);
                        g[u][0] = __uint_as_float(r.x << 16); g[u][1] = __uint_as_float(r.x & 0xffff0000u);
                        g[u][2] = __uint_as_float(r.y << 16); g[u][3] = __uint_as_float(r.y & 0xffff0000u);
                    } else {
                        const uint32_t r = *reinterpret_cast<const uint32_t*>(T.dyh + (long long)m * Cout + cg * 2);
                        g[u][0] = __uint_as_float(r << 16); g[u][1] = __uint_as_float(r & 0xffff0000u);
                    }
                } else if constexpr (CPT == 4) {
                    const float4 v = ld4(T.dy + (long long)m * Cout + cg * 4);
                    g[u][0] = v.x; g[u][1] = v.y; g[u][2] = v.z; g[u][3] = v.w;
                } else {
                    const float2 v = *reinterpret_cast<const float2*>(T.dy + (long long)m * Cout + cg * 2);
                    g[u][0] = v.x; g[u][1] = v.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int m = min(mb + u * PP, m1 - 1);   // past the end: g == 0, any valid tile position will do
            const int n = m / HW, rem = m - n * HW;
            const int h = rem / W, w = rem - h * W;
            const int s = n - sg.n_first;
            const float* tp = tile + s * seg_floats + (h - sg.ha[s]) * Wp + w;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                for (int kw = 0; kw < KS; ++kw) {
                    const float x = tp[kh * Wp + kw];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) acc[kh * KS + kw][j] = fmaf(x, g[u][j], acc[kh * KS + kw][j]);
                }
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[TAPS][j] += g[u][j];
        }
    }
    // lanes with the same channel group, then the 8 warps
#pragma unroll
    for (int t = 0; t <= TAPS; ++t)
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            float v = acc[t][j];
            for (int msk = CG; msk < 32; msk <<= 1) v += __shfl_xor_sync(0xffffffffu, v, msk);
            if (CG >= 32 || lane < CG) red[(warp * (TAPS + 1) + t) * Cout + cg * CPT + j] = v;
        }
    __syncthreads();
    for (int i = tid; i < n_out; i += kThreads) {
        const int c = i % Cout;
        float a = 0.f;
        if (CG <= 32) {
#pragma unroll
            for (int wq = 0; wq < kThreads / 32; ++wq) a += red[wq * n_out + i];
        } else {
            // CG in {64, 128}: channel group cg lives in the warps with (warp % (CG / 32)) == cg / 32
            const int per = CG / 32, mine = (c / CPT) / 32;
            for (int wq = mine; wq < kThreads / 32; wq += per) a += red[wq * n_out + i];
        }
        out[i] = a;
    }
}

__global__ void __launch_bounds__(kThreads) stem_wgrad_kernel(const WgradTask* __restrict__ tasks, int blocks_per_task,
                                                              int seg_floats, int n_b, int step) {
    extern __shared__ __align__(16) float sm[];
    __shared__ WgradTask T;
    const int task = blockIdx.x / blocks_per_task, split = blockIdx.x - task * blocks_per_task;
    if (threadIdx.x == 0) T = tasks[task];
    __syncthreads();
    if (T.k == 3)
        stem_wgrad_body<3, 4>(T, sm, split, seg_floats, n_b, step);
    else
        stem_wgrad_body<5, 2>(T, sm, split, seg_floats, n_b, step);
}

int seg_floats_for(int W, int k, int rows) {
    const int P = (k - 1) / 2;
    return ((rows / W + 2 + 2 * P) * (W + 2 * P) + 3) & ~3;
}

template <class K>
cudaError_t opt_in(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem < 48 * 1024 ? 48 * 1024 : smem));
}

}  // namespace

bool Launch::stem_ok(int H, int W, int Cin, int Cout, int k, int stride, int n_b) {
    if (Cin != 1 || stride != 1 || (k != 3 && k != 5)) return false;
    if (Cout < 16 || (Cout & (Cout - 1)) != 0) return false;
    if (Cout > (k == 3 ? 128 : 64)) return false;                  // a pixel's channel groups must fit one warp
    if ((long long)H * W < kStemRows || W < 8) return false;        // a block may touch at most two samples
    (void)n_b;
    const size_t smem = ((size_t)(k * k + 1) * Cout + 2 * (size_t)seg_floats_for(W, k, kStemRows) +
                         (size_t)8 * (k * k + 1) * Cout) * sizeof(float);
    return smem <= 200 * 1024;
}

int Launch::stem_conv(const ConvTask* tasks, int n_tasks, int max_k, int W, int max_cout, long long M, int n_b, int step,
                      void* stream) {
    const int bpt = (int)((M + kStemRows - 1) / kStemRows);
    const int seg = seg_floats_for(W, max_k, kStemRows);
    const size_t smem = (2 * (size_t)seg + 8 * 2 * (size_t)max_cout) * sizeof(float);
    cudaError_t e = opt_in(stem_conv_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    stem_conv_kernel<<<n_tasks * bpt, kThreads, smem, (cudaStream_t)stream>>>(tasks, bpt, kStemRows, seg, n_b, step);
    return (int)cudaGetLastError();
}

int Launch::stem_wgrad(const WgradTask* tasks, int n_tasks, int max_k, int W, int max_cout, int splits, int n_b, int step,
                       void* stream) {
    const int seg = seg_floats_for(W, max_k, kStemRows);
    const size_t smem = (2 * (size_t)seg + (size_t)8 * (max_k * max_k + 1) * max_cout) * sizeof(float);
    cudaError_t e = opt_in(stem_wgrad_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    stem_wgrad_kernel<<<n_tasks * splits, kThreads, smem, (cudaStream_t)stream>>>(tasks, splits, seg, n_b, step);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
