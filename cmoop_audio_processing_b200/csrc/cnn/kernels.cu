// fp32 SIMT kernels of the population-batched candidate CNN (exact path): grouped implicit-GEMM
// convolution (forward, data-gradient via pre-transposed weights, weight-gradient with deterministic
// split-M), BatchNorm statistics / apply / backward, ReLU / 2x2 'same' max-pool / residual add,
// global average pooling, dropout, Keras-style sparse cross-entropy, Adam, initialisation.
//
// Semantics follow the reference call sites nsga_penalty.py:250-332 (variant A),
// sa_nsga_penalty.py:150-176 (variant B) and the recipe nsga_penalty.py:377-386 with Keras-default
// numerics (see oracle/cnn_ref.py, which is the CPU statement these kernels are tested against).
// Every kernel is grouped over candidates (cnn.cuh) and deterministic: no floating-point atomics.
#include <cuda_runtime.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <class T, class F>
__device__ __forceinline__ int find_task(const T* tasks, int n, int block, F begin_of) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (begin_of(tasks[mid]) <= block)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// One binary search per block instead of one per thread (the elementwise kernels move 16 bytes per thread, so the
// ~8 dependent loads of a per-thread search were most of their instruction stream).
template <class T, class F>
__device__ __forceinline__ int block_find_task(const T* tasks, int n, int block, F begin_of, const int* block_task = nullptr) {
    if (block_task) return __ldg(block_task + block);      // direct table: one load
    __shared__ int s_task;
    if (threadIdx.x == 0) s_task = find_task(tasks, n, block, begin_of);
    __syncthreads();
    return s_task;
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// y[M][Cout] = im2col(x)[M][K(+1)] * w[K(+1)][Cout]   (M = n_b*Ho*Wo, bias = last row of w)
__global__ void __launch_bounds__(256) conv_gemm_kernel(const ConvTask* __restrict__ tasks, int n_tasks, int n_b,
                                                        int step) {
    __shared__ ConvTask T;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    if (tid == 0) T = tasks[find_task(tasks, n_tasks, blockIdx.x, [](const ConvTask& t) { return t.tile_begin; })];
    __syncthreads();
    const int local = blockIdx.x - T.tile_begin;
    const int tm = local / T.tiles_n, tn = local - tm * T.tiles_n;
    const int HoWo = T.Ho * T.Wo, M = n_b * HoWo, m0 = tm * BM, n0 = tn * BN;
    if (m0 >= M) return;
    const int K = T.k * T.k * T.Cin, Kext = K + T.use_bias;
    const float* xbase = T.x + T.x_step * step;
    const int* gather = T.gather ? T.gather + T.gather_step * step : nullptr;

    const int row = tid & 63, kq = tid >> 6;
    const int m = m0 + row;
    const bool mvalid = m < M;
    int hi0 = 0, wi0 = 0;
    const float* xin = xbase;
    if (mvalid) {
        const int n = m / HoWo, r = m - n * HoWo;
        const int ho = r / T.Wo, wo = r - ho * T.Wo;
        hi0 = ho * T.stride - T.pad;
        wi0 = wo * T.stride - T.pad;
        xin = xbase + (long long)(gather ? gather[n] : n) * ((long long)T.H * T.W * T.Cin);
    }
    const bool vec_a = (T.Cin & 3) == 0 && aligned16(xbase);
    const int bk = tid >> 4, bn4 = (tid & 15) * 4;
    const bool vec_b = (T.Cout & 3) == 0 && aligned16(T.w);
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // register double-buffering: the global loads of K block k0 + BK are in flight while block k0 is multiplied
    // (same accumulation order as a plain loop, so results are bit-identical)
    auto load_a = [&](int k0, float (&a)[4]) {
        a[0] = a[1] = a[2] = a[3] = 0.f;
        const int kb = k0 + kq * 4;
        if (mvalid) {
            if (vec_a && kb + 3 < K) {
                const int kk = kb / T.Cin, ci = kb - kk * T.Cin;
                const int kh = kk / T.k, kw = kk - kh * T.k;
                const int hi = hi0 + kh, wi = wi0 + kw;
                if ((unsigned)hi < (unsigned)T.H && (unsigned)wi < (unsigned)T.W) {
                    const float4 v = *reinterpret_cast<const float4*>(xin + ((long long)hi * T.W + wi) * T.Cin + ci);
                    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kj = kb + j;
                    if (kj < K) {
                        const int kk = kj / T.Cin, ci = kj - kk * T.Cin;
                        const int kh = kk / T.k, kw = kk - kh * T.k;
                        const int hi = hi0 + kh, wi = wi0 + kw;
                        if ((unsigned)hi < (unsigned)T.H && (unsigned)wi < (unsigned)T.W)
                            a[j] = xin[((long long)hi * T.W + wi) * T.Cin + ci];
                    } else if (kj == K && T.use_bias) {
                        a[j] = 1.f;
                    }
                }
            }
        }
    };
    auto load_b = [&](int k0, float (&b)[4]) {
        b[0] = b[1] = b[2] = b[3] = 0.f;
        const int kB = k0 + bk;
        if (kB < Kext) {
            const float* wr = T.w + (long long)kB * T.Cout + n0 + bn4;
            if (vec_b && n0 + bn4 + 3 < T.Cout) {
                const float4 v = *reinterpret_cast<const float4*>(wr);
                b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + bn4 + j < T.Cout) b[j] = wr[j];
            }
        }
    };
    float a[4], b[4];
    load_a(0, a);
    load_b(0, b);
    for (int k0 = 0; k0 < Kext; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) As[kq * 4 + j][row] = a[j];
        *reinterpret_cast<float4*>(&Bs[bk][bn4]) = make_float4(b[0], b[1], b[2], b[3]);
        __syncthreads();
        if (k0 + BK < Kext) {
            load_a(k0 + BK, a);
            load_b(k0 + BK, b);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }

    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mi = m0 + ty * 4 + i;
        if (mi >= M) continue;
        long long base;
        if (T.out_s == 0) {
            base = (long long)mi * T.Cout;
        } else {
            const int n = mi / HoWo, r = mi - n * HoWo;
            const int ho = r / T.Wo, wo = r - ho * T.Wo;
            base = (((long long)n * T.out_h + ho * T.out_s) * T.out_w + wo * T.out_s) * T.Cout;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx * 4 + j;
            if (col >= T.Cout) continue;
            float v = acc[i][j];
            if (T.relu) v = fmaxf(v, 0.f);
            if (!T.y) v = __bfloat162float(__float2bfloat16_rn(v));     // bf16-only output: the stored value is THE value
            if (T.y) {
                if (T.accumulate)
                    T.y[base + col] += v;
                else
                    T.y[base + col] = v;
            }
            if (T.yh) T.yh[base + col] = __float2bfloat16_rn(v);
            s1[j] += v;
            s2[j] = fmaf(v, v, s2[j]);
        }
    }
    if (T.stat_part) {
        float(*red1)[BM + 4] = As;
        float(*red2)[BN + 4] = Bs;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red1[ty][tx * 4 + j] = s1[j];
            red2[ty][tx * 4 + j] = s2[j];
        }
        __syncthreads();
        if (tid < BN && n0 + tid < T.Cout) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                a1 += red1[r][tid];
                a2 += red2[r][tid];
            }
            T.stat_part[((long long)tm * 2 + 0) * T.Cout + n0 + tid] = a1;
            T.stat_part[((long long)tm * 2 + 1) * T.Cout + n0 + tid] = a2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// out[split][K+1][Cout] = sum_{m in split} im2col(x)[m][K+1]^T * dy[m][Cout]   (row K = bias gradient)
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradTask* __restrict__ tasks, int n_tasks, int n_b,
                                                         int step) {
    __shared__ WgradTask T;
    __shared__ __align__(16) float As[BK][BM + 4];   // [m][kk]
    __shared__ __align__(16) float Bs[BK][BN + 4];   // [m][co]
    const int tid = threadIdx.x;
    if (tid == 0) T = tasks[find_task(tasks, n_tasks, blockIdx.x, [](const WgradTask& t) { return t.tile_begin; })];
    __syncthreads();
    int local = blockIdx.x - T.tile_begin;
    const int per_split = T.tiles_k * T.tiles_n;
    const int split = local / per_split;
    local -= split * per_split;
    const int tk = local / T.tiles_n, tn = local - tk * T.tiles_n;
    const int HoWo = T.Ho * T.Wo, M = n_b * HoWo;
    const int K = T.k * T.k * T.Cin, Kext = K + 1;
    const int kk0 = tk * BM, n0 = tn * BN;
    const int m_begin = split * T.m_chunk;
    const int m_end = min(M, m_begin + T.m_chunk);
    const float* xbase = T.x + T.x_step * step;
    const int* gather = T.gather ? T.gather + T.gather_step * step : nullptr;
    const long long img = (long long)T.H * T.W * T.Cin;

    const int mm = tid >> 4, q4 = (tid & 15) * 4;
    // fixed per thread: the four kk it loads
    const int kb = kk0 + q4;
    const bool vec_a = (T.Cin & 3) == 0 && aligned16(xbase) && kb + 3 < K;
    int kh[4], kw[4], ci[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int kj = kb + j;
        if (kj < K) {
            const int kk = kj / T.Cin;
            ci[j] = kj - kk * T.Cin;
            kh[j] = kk / T.k;
            kw[j] = kk - kh[j] * T.k;
        } else {
            kh[j] = kw[j] = 0;
            ci[j] = (kj == K) ? -1 : -2;     // -1: ones column, -2: outside
        }
    }
    const bool vec_b = !T.dyh && (T.Cout & 3) == 0 && aligned16(T.dy);
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int ms = m_begin; ms < m_end; ms += BK) {
        const int m = ms + mm;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < m_end) {
            const int n = m / HoWo, r = m - n * HoWo;
            const int ho = r / T.Wo, wo = r - ho * T.Wo;
            const int hi0 = ho * T.stride - T.pad, wi0 = wo * T.stride - T.pad;
            const float* xin = xbase + (long long)(gather ? gather[n] : n) * img;
            if (vec_a) {
                const int hi = hi0 + kh[0], wi = wi0 + kw[0];
                if ((unsigned)hi < (unsigned)T.H && (unsigned)wi < (unsigned)T.W) {
                    const float4 v = *reinterpret_cast<const float4*>(xin + ((long long)hi * T.W + wi) * T.Cin + ci[0]);
                    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (ci[j] >= 0) {
                        const int hi = hi0 + kh[j], wi = wi0 + kw[j];
                        if ((unsigned)hi < (unsigned)T.H && (unsigned)wi < (unsigned)T.W)
                            a[j] = xin[((long long)hi * T.W + wi) * T.Cin + ci[j]];
                    } else if (ci[j] == -1) {
                        a[j] = 1.f;
                    }
                }
            }
            const float* dr = T.dy + (long long)m * T.Cout + n0 + q4;
            if (T.dyh) {                               // bf16-only gradient (precision bf16, Cin = 1 convolution)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + q4 + j < T.Cout) b[j] = __bfloat162float(T.dyh[(long long)m * T.Cout + n0 + q4 + j]);
            } else if (vec_b && n0 + q4 + 3 < T.Cout) {
                const float4 v = *reinterpret_cast<const float4*>(dr);
                b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + q4 + j < T.Cout) b[j] = dr[j];
            }
        }
        *reinterpret_cast<float4*>(&As[mm][q4]) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(&Bs[mm][q4]) = make_float4(b[0], b[1], b[2], b[3]);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < BK; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = T.out + (long long)split * Kext * T.Cout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kk = kk0 + ty * 4 + i;
        if (kk >= Kext) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx * 4 + j;
            if (col < T.Cout) out[(long long)kk * T.Cout + col] = acc[i][j];
        }
    }
}

__global__ void __launch_bounds__(256) reduce_kernel(const ReduceTask* __restrict__ tasks, int n_tasks) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const ReduceTask& r) { return r.block_begin; });
    const ReduceTask T = tasks[t];
    const int i = (blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    if (i >= T.n) return;
    float s = 0.f;
    int k = 0;
    for (; k + 3 < T.splits; k += 4) {          // four independent loads per iteration, same summation order
        const float a0 = T.part[(long long)k * T.n + i], a1 = T.part[(long long)(k + 1) * T.n + i];
        const float a2 = T.part[(long long)(k + 2) * T.n + i], a3 = T.part[(long long)(k + 3) * T.n + i];
        s += a0; s += a1; s += a2; s += a3;
    }
    for (; k < T.splits; ++k) s += T.part[(long long)k * T.n + i];
    T.out[i] = s;
}

__global__ void __launch_bounds__(256) wt_kernel(const WtTask* __restrict__ tasks, int n_tasks) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const WtTask& r) { return r.block_begin; });
    const WtTask T = tasks[t];
    const int e = (blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    const int total = T.k * T.k * T.Cin * T.Cout;
    if (e >= total) return;
    const int ci = e % T.Cin;
    int r = e / T.Cin;
    const int co = r % T.Cout;
    r /= T.Cout;
    const int kw = r % T.k, kh = r / T.k;
    const int src = (((T.k - 1 - kh) * T.k + (T.k - 1 - kw)) * T.Cin + ci) * T.Cout + co;
    T.wt[e] = T.w[src];
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics -> (mean, invstd, scale, shift); grid (task, 8-channel chunk): 16 tile lanes x 8 channels sum
// the tile partials in double and are combined in lane order (deterministic)
__global__ void __launch_bounds__(128) bn_finalize_kernel(const PostTask* __restrict__ tasks, int n_b, int training,
                                                          float momentum, float eps) {
    __shared__ double red[2][128];
    const PostTask T = tasks[blockIdx.x];
    const int c = blockIdx.y * 8 + (threadIdx.x & 7);
    if (!T.has_bn || blockIdx.y * 8 >= T.C) return;
    const int t_lane = threadIdx.x >> 3;
    const double count = (double)n_b * T.H * T.W;
    const int tiles = (n_b * T.H * T.W + BM - 1) / BM;    // tiles the conv epilogue / bn_stats wrote for THIS batch size
    double s1 = 0.0, s2 = 0.0;
    if (training && c < T.C) {
        // four tiles per iteration: the eight loads are independent (the stem has ~2 000 tiles: 122 dependent round trips per
        // thread otherwise), the summation order stays t_lane, t_lane + 16, ...
        const float* sp = T.stat_part + c;
        const long long stride = 2LL * T.C;
        int t = t_lane;
        for (; t + 48 < tiles; t += 64) {
            const float a0 = sp[(long long)t * stride], b0 = sp[(long long)t * stride + T.C];
            const float a1 = sp[(long long)(t + 16) * stride], b1 = sp[(long long)(t + 16) * stride + T.C];
            const float a2 = sp[(long long)(t + 32) * stride], b2 = sp[(long long)(t + 32) * stride + T.C];
            const float a3 = sp[(long long)(t + 48) * stride], b3 = sp[(long long)(t + 48) * stride + T.C];
            s1 += (double)a0; s2 += (double)b0;
            s1 += (double)a1; s2 += (double)b1;
            s1 += (double)a2; s2 += (double)b2;
            s1 += (double)a3; s2 += (double)b3;
        }
        for (; t < tiles; t += 16) {
            s1 += (double)sp[(long long)t * stride];
            s2 += (double)sp[(long long)t * stride + T.C];
        }
    }
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (t_lane == 0 && c < T.C) {
        float mean, var;
        if (training) {
            double a1 = 0.0, a2 = 0.0;
            for (int l = 0; l < 16; ++l) {
                a1 += red[0][l * 8 + (threadIdx.x & 7)];
                a2 += red[1][l * 8 + (threadIdx.x & 7)];
            }
            const double mu = a1 / count;
            double vv = a2 / count - mu * mu;
            vv = vv > 0.0 ? vv : 0.0;
            mean = (float)mu;
            var = (float)vv;
            T.mov_mean[c] = T.mov_mean[c] * momentum + mean * (1.f - momentum);
            T.mov_var[c] = T.mov_var[c] * momentum + var * (1.f - momentum);
        } else {
            mean = T.mov_mean[c];
            var = T.mov_var[c];
        }
        const float invstd = rsqrtf(var + eps);
        const float scale = T.gamma[c] * invstd;
        T.bn[0 * T.C + c] = mean;
        T.bn[1 * T.C + c] = invstd;
        T.bn[2 * T.C + c] = scale;
        T.bn[3 * T.C + c] = T.beta[c] - mean * scale;
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint2 bf16x4(float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// eight consecutive channels from the fp32 tensor or from its bf16-only form (exact widening)
__device__ __forceinline__ void load8(const float* f, const __nv_bfloat16* h, long long e, float (&v)[8]) {
    if (h) {
        const uint4 r = *reinterpret_cast<const uint4*>(h + e);
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
        v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
        v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
    } else {
        const float4 a = ld4(f + e), b = ld4(f + e + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
__device__ __forceinline__ void load4(const float* f, const __nv_bfloat16* h, long long e, float (&v)[4]) {
    if (h) {
        const uint2 r = *reinterpret_cast<const uint2*>(h + e);
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    } else {
        const float4 a = ld4(f + e);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
}
__device__ __forceinline__ float load1(const float* f, const __nv_bfloat16* h, long long e) {
    return h ? __bfloat162float(h[e]) : f[e];
}
// compile-time storage choice (the post kernels are instantiated per precision: no per-load branch, so the compiler can
// issue all loads of an iteration back to back)
template <bool HALF>
__device__ __forceinline__ void load8t(const float* __restrict__ f, const __nv_bfloat16* __restrict__ h, long long e, float (&v)[8]) {
    if constexpr (HALF) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(h + e));
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
        v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
        v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(f + e)), b = __ldg(reinterpret_cast<const float4*>(f + e + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
template <bool HALF>
__device__ __forceinline__ void store8t(float* __restrict__ f, __nv_bfloat16* __restrict__ h, long long e, const float (&z)[8]) {
    if constexpr (HALF) {
        const uint2 h0 = bf16x4(make_float4(z[0], z[1], z[2], z[3])), h1 = bf16x4(make_float4(z[4], z[5], z[6], z[7]));
        *reinterpret_cast<uint4*>(h + e) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    } else {
        *reinterpret_cast<float4*>(f + e) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(f + e + 4) = make_float4(z[4], z[5], z[6], z[7]);
    }
}

// Thread mapping of the two elementwise post kernels: a thread OWNS 8 consecutive channels (its BN constants and the task
// fields stay in registers) and walks over output pixels -- kPostIter per thread, 256 / (C / 8) pixel lanes per block.  The
// first version gave every thread one (pixel, 8 channels) element: each thread then re-loaded the task record and twelve
// 16-byte BN vectors for 16 bytes of output and the kernels ran at a third of the HBM rate whatever the activation width.
constexpr int kPostIter = 8;
__host__ __device__ inline int post_pixels_per_block(int C) {
    const int cg = C >> 3;                               // 8-channel groups: 2 .. 64
    return (256 / cg) * kPostIter;
}
__device__ __forceinline__ void store8(float* f, __nv_bfloat16* h, long long e, const float (&z)[8]) {
    if (f) {
        *reinterpret_cast<float4*>(f + e) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(f + e + 4) = make_float4(z[4], z[5], z[6], z[7]);
    }
    if (h) {
        const uint2 h0 = bf16x4(make_float4(z[0], z[1], z[2], z[3])), h1 = bf16x4(make_float4(z[4], z[5], z[6], z[7]));
        *reinterpret_cast<uint4*>(h + e) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    }
}

// [BN] -> [ReLU] -> [2x2/s2 'same' max-pool] -> [+skip, ReLU].  HALF: activations are stored in bf16 only (precision bf16).
// The (up to) four window loads of a pooled pixel are UNCONDITIONAL (coordinates clamped into the image, validity applied
// in the compare chain): with per-load branches the compiler kept them in four separate reconvergence regions and every
// iteration paid four dependent memory round trips.
template <bool HALF>
__global__ void __launch_bounds__(256, 2) post_fwd_kernel(const PostTask* __restrict__ tasks, int n_tasks, int n_b,
                                                          const int* __restrict__ block_task) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const PostTask& r) { return r.block_begin; }, block_task);
    const PostTask* __restrict__ Tp = tasks + t;
    const int C = Tp->C, cgs = C >> 3, lanes = 256 / cgs;
    const int cg = threadIdx.x % cgs, pl = threadIdx.x / cgs;
    if (pl >= lanes) return;
    const int c = cg * 8;
    const int H = Tp->H, W = Tp->W, Ho = Tp->Ho, Wo = Tp->Wo;
    const int pool = Tp->pool, relu_mid = Tp->relu_mid, add_skip = Tp->add_skip, has_bn = Tp->has_bn;
    const float* __restrict__ u = Tp->u;
    const __nv_bfloat16* __restrict__ uh = Tp->uh;
    const float* __restrict__ skip = Tp->skip;
    const __nv_bfloat16* __restrict__ skiph = Tp->skiph;
    float* __restrict__ v = Tp->v;
    __nv_bfloat16* __restrict__ vh = Tp->vh;
    uint8_t* __restrict__ idx = Tp->idx;
    const int n_pix = n_b * Ho * Wo;
    const int pixb = lanes * kPostIter;
    const int p0 = (blockIdx.x - Tp->block_begin) * pixb;
    const int p1 = min(n_pix, p0 + pixb);
    float sc[8], sh[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { sc[q] = 1.f; sh[q] = 0.f; }
    if (has_bn) {
        const float* bn = Tp->bn;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const float4 a = ld4(bn + 2 * C + c + 4 * h2), b = ld4(bn + 3 * C + c + 4 * h2);
            sc[4 * h2 + 0] = a.x; sc[4 * h2 + 1] = a.y; sc[4 * h2 + 2] = a.z; sc[4 * h2 + 3] = a.w;
            sh[4 * h2 + 0] = b.x; sh[4 * h2 + 1] = b.y; sh[4 * h2 + 2] = b.z; sh[4 * h2 + 3] = b.w;
        }
    }
    for (int pix = p0 + pl; pix < p1; pix += lanes) {
        const long long e = (long long)pix * C + c;
        float z[8];
        float sk[8];
        if (add_skip) load8t<HALF>(skip, skiph, e, sk);
        if (pool) {
            const int r1 = pix / Wo, wo = pix - r1 * Wo;
            const int n = r1 / Ho, ho = r1 - n * Ho;
            const int h0 = 2 * ho, w0 = 2 * wo;
            const bool okh = h0 + 1 < H, okw = w0 + 1 < W;             // (h0, w0) itself is always inside
            const long long row0 = ((long long)n * H + h0) * W, row1 = ((long long)n * H + (okh ? h0 + 1 : h0)) * W;
            const int wb = okw ? w0 + 1 : w0;
            float x4[4][8];
            load8t<HALF>(u, uh, (row0 + w0) * C + c, x4[0]);
            load8t<HALF>(u, uh, (row0 + wb) * C + c, x4[1]);
            load8t<HALF>(u, uh, (row1 + w0) * C + c, x4[2]);
            load8t<HALF>(u, uh, (row1 + wb) * C + c, x4[3]);
            const bool ok[4] = {true, okw, okh, okh && okw};
            int code[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) code[q] = 0;
            // the max in the reference's scan order (first maximum wins); position 0 is always valid
#pragma unroll
            for (int d = 0; d < 4; ++d) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float x = x4[d][q];
                    if (has_bn) x = fmaf(x, sc[q], sh[q]);
                    if (relu_mid) x = fmaxf(x, 0.f);
                    if (d == 0 || (ok[d] && x > z[q])) {
                        z[q] = x;
                        code[q] = d;
                    }
                }
            }
            uint2 pk;
            pk.x = (unsigned)code[0] | ((unsigned)code[1] << 8) | ((unsigned)code[2] << 16) | ((unsigned)code[3] << 24);
            pk.y = (unsigned)code[4] | ((unsigned)code[5] << 8) | ((unsigned)code[6] << 16) | ((unsigned)code[7] << 24);
            *reinterpret_cast<uint2*>(idx + e) = pk;
        } else {
            float x8[8];
            load8t<HALF>(u, uh, e, x8);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float x = x8[q];
                if (has_bn) x = fmaf(x, sc[q], sh[q]);
                if (relu_mid) x = fmaxf(x, 0.f);
                z[q] = x;
            }
        }
        if (add_skip) {
#pragma unroll
            for (int q = 0; q < 8; ++q) z[q] = fmaxf(z[q] + sk[q], 0.f);
        }
        store8t<HALF>(v, vh, e, z);
        if (!HALF && vh) store8t<true>(nullptr, vh, e, z);
    }
}

// BN backward, pass 1: per-channel partial sums of g and g*xhat over the unit's output elements.
// A CTA covers T.bwd_pix output pixels (128 on the large maps, 32 on the small ones: the deep blocks have 768-2 240 pooled
// pixels per batch, and a thread that walks 128 of them alone is a ~200 us chain of dependent gathers); thread = (pixel
// lane, 4 channels); the pixel lanes are summed in lane order through shared memory: ONE partial row per CTA.
template <bool HALF>
__global__ void __launch_bounds__(256) post_bwd_reduce_kernel(const PostTask* __restrict__ tasks, int n_tasks,
                                                              int n_b, const int* __restrict__ block_task) {
    __shared__ float red[256 * 8];
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const PostTask& r) { return r.block_begin_bwd; }, block_task);
    const PostTask T = tasks[t];
    const int blk = blockIdx.x - T.block_begin_bwd;
    const int C4 = T.C >> 2;
    const int cb = C4 < 128 ? C4 : 128;
    const int lanes = 256 / cb;
    const int p_lane = threadIdx.x / cb, c_lane = threadIdx.x - p_lane * cb;
    const bool active = p_lane < lanes;
    const int n_pix = n_b * T.Ho * T.Wo;
    const int pix0 = blk * T.bwd_pix, pix1 = pix0 + T.bwd_pix < n_pix ? pix0 + T.bwd_pix : n_pix;
    if (pix0 >= n_pix) return;                          // uniform for the block (short last batch)
    for (int cbase = 0; cbase < C4; cbase += cb) {      // same trip count for every thread (block barriers inside)
        const int c4 = cbase + c_lane;
        const bool live = active && c4 < C4;
        const int c = (live ? c4 : 0) * 4;
        const float4 mean = ld4(T.bn + 0 * T.C + c), invstd = ld4(T.bn + 1 * T.C + c);
        const float4 scale = ld4(T.bn + 2 * T.C + c), shift = ld4(T.bn + 3 * T.C + c);
        const float mu[4] = {mean.x, mean.y, mean.z, mean.w}, is[4] = {invstd.x, invstd.y, invstd.z, invstd.w};
        const float sc[4] = {scale.x, scale.y, scale.z, scale.w}, sh[4] = {shift.x, shift.y, shift.z, shift.w};
        float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
        for (int pix = pix0 + p_lane; live && pix < pix1; pix += lanes) {
            const long long e = (long long)pix * T.C + c;
            const float4 gv = ld4(T.dv + e);
            float g[4] = {gv.x, gv.y, gv.z, gv.w};
            if (T.add_skip) {
                float vv[4];
                if constexpr (HALF) load4(nullptr, T.vh, e, vv); else load4(T.v, nullptr, e, vv);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (!(vv[q] > 0.f)) g[q] = 0.f;
            }
            float u[4];
            if (T.pool) {
                const int r = pix / T.Wo, wo = pix - r * T.Wo;
                const int n = r / T.Ho, ho = r - n * T.Ho;
                const uchar4 cd = *reinterpret_cast<const uchar4*>(T.idx + e);
                const unsigned char code[4] = {cd.x, cd.y, cd.z, cd.w};
                const long long base = (((long long)n * T.H + 2 * ho) * T.W + 2 * wo) * T.C + c;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    u[q] = HALF ? __bfloat162float(T.uh[base + ((long long)(code[q] >> 1) * T.W + (code[q] & 1)) * T.C + q])
                                : T.u[base + ((long long)(code[q] >> 1) * T.W + (code[q] & 1)) * T.C + q];
            } else {
                if constexpr (HALF) load4(nullptr, T.uh, e, u); else load4(T.u, nullptr, e, u);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float gq = g[q];
                if (T.relu_mid && !(fmaf(u[q], sc[q], sh[q]) > 0.f)) gq = 0.f;
                sg[q] += gq;
                sgx[q] = fmaf(gq, (u[q] - mu[q]) * is[q], sgx[q]);
            }
        }
        // pixel lanes -> one row: lanes that share a warp (cb < 32) are combined by shuffles first, then the per-warp (or
        // per-lane) partials go through shared memory and the first cb threads add them in a fixed order
        float a[8] = {sg[0], sg[1], sg[2], sg[3], sgx[0], sgx[1], sgx[2], sgx[3]};
        for (int msk = cb; msk < 32; msk <<= 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], msk);
        }
        const int groups = cb < 32 ? 8 : lanes;                 // partial rows in shared memory: per warp / per pixel lane
        const int grp = cb < 32 ? (threadIdx.x >> 5) : p_lane;
        if (cb >= 32 || (threadIdx.x & 31) < cb) {
            float* mine = red + (grp * cb + c_lane) * 8;
#pragma unroll
            for (int q = 0; q < 8; ++q) mine[q] = a[q];
        }
        __syncthreads();
        if (live && p_lane == 0) {
            float r8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int l = 0; l < groups; ++l) {
                const float* r = red + (l * cb + c_lane) * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) r8[q] += r[q];
            }
            *reinterpret_cast<float4*>(T.bwd_part + ((long long)blk * 2 + 0) * T.C + c) = make_float4(r8[0], r8[1], r8[2], r8[3]);
            *reinterpret_cast<float4*>(T.bwd_part + ((long long)blk * 2 + 1) * T.C + c) = make_float4(r8[4], r8[5], r8[6], r8[7]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const PostTask* __restrict__ tasks, int n_b) {
    __shared__ double red[2][128];
    const PostTask T = tasks[blockIdx.x];
    const int c = blockIdx.y * 8 + (threadIdx.x & 7);
    if (!T.has_bn || blockIdx.y * 8 >= T.C) return;
    const int r_lane = threadIdx.x >> 3;
    const long long n_pix = (long long)n_b * T.Ho * T.Wo;
    const long long rows = (n_pix + T.bwd_pix - 1) / T.bwd_pix;      // one partial row per CTA of post_bwd_reduce
    const double count = (double)n_b * T.H * T.W;
    double sg = 0.0, sgx = 0.0;
    if (c < T.C) {
        // four rows per iteration: the loads are independent, the summation order stays r_lane, r_lane + 16, ...
        long long r = r_lane;
        for (; r + 48 < rows; r += 64) {
            const float a0 = T.bwd_part[(r * 2 + 0) * T.C + c], b0 = T.bwd_part[(r * 2 + 1) * T.C + c];
            const float a1 = T.bwd_part[((r + 16) * 2 + 0) * T.C + c], b1 = T.bwd_part[((r + 16) * 2 + 1) * T.C + c];
            const float a2 = T.bwd_part[((r + 32) * 2 + 0) * T.C + c], b2 = T.bwd_part[((r + 32) * 2 + 1) * T.C + c];
            const float a3 = T.bwd_part[((r + 48) * 2 + 0) * T.C + c], b3 = T.bwd_part[((r + 48) * 2 + 1) * T.C + c];
            sg += (double)a0; sgx += (double)b0;
            sg += (double)a1; sgx += (double)b1;
            sg += (double)a2; sgx += (double)b2;
            sg += (double)a3; sgx += (double)b3;
        }
        for (; r < rows; r += 16) {
            sg += (double)T.bwd_part[(r * 2 + 0) * T.C + c];
            sgx += (double)T.bwd_part[(r * 2 + 1) * T.C + c];
        }
    }
    red[0][threadIdx.x] = sg;
    red[1][threadIdx.x] = sgx;
    __syncthreads();
    if (r_lane == 0 && c < T.C) {
        double a1 = 0.0, a2 = 0.0;
        for (int l = 0; l < 16; ++l) {
            a1 += red[0][l * 8 + (threadIdx.x & 7)];
            a2 += red[1][l * 8 + (threadIdx.x & 7)];
        }
        T.dbeta[c] = (float)a1;
        T.dgamma[c] = (float)a2;
        T.bn[4 * T.C + c] = (float)(a1 / count);
        T.bn[5 * T.C + c] = (float)(a2 / count);
    }
}

// backward of the whole post stage.  Same thread mapping as post_fwd_kernel: a thread owns 8 channels (six BN vectors in
// registers) and walks over OUTPUT pixels; for a pooled unit it reads the pooled gradient / mask / argmax code once and
// writes the gradient of all (up to four) input pixels of the 2x2 window (unconditional, clamped loads as above).
template <bool HALF>
__global__ void __launch_bounds__(256, 2) post_bwd_apply_kernel(const PostTask* __restrict__ tasks, int n_tasks, int n_b,
                                                                const int* __restrict__ block_task) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const PostTask& r) { return r.block_begin_apply; }, block_task);
    const PostTask* __restrict__ Tp = tasks + t;
    const int C = Tp->C, cgs = C >> 3, lanes = 256 / cgs;
    const int cg = threadIdx.x % cgs, pl = threadIdx.x / cgs;
    if (pl >= lanes) return;
    const int c = cg * 8;
    const int H = Tp->H, W = Tp->W, Ho = Tp->Ho, Wo = Tp->Wo;
    const int pool = Tp->pool, relu_mid = Tp->relu_mid, add_skip = Tp->add_skip, has_bn = Tp->has_bn, relu_in = Tp->relu_in;
    const float* __restrict__ u = Tp->u;
    const __nv_bfloat16* __restrict__ uh = Tp->uh;
    const float* __restrict__ v = Tp->v;
    const __nv_bfloat16* __restrict__ vh = Tp->vh;
    const float* __restrict__ dv = Tp->dv;
    const uint8_t* __restrict__ idx = Tp->idx;
    float* __restrict__ du = Tp->du;
    __nv_bfloat16* __restrict__ duh = Tp->duh;
    float* __restrict__ dskip = Tp->dskip;
    __nv_bfloat16* __restrict__ dskiph = Tp->dskiph;
    const int n_pix = n_b * Ho * Wo;
    const int pixb = lanes * kPostIter;
    const int p0 = (blockIdx.x - Tp->block_begin_apply) * pixb;
    const int p1 = min(n_pix, p0 + pixb);
    float mu[8], is[8], sc[8], sh[8], mg[8], mgx[8];
    if (has_bn) {
        const float* bn = Tp->bn;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int ch = c + 4 * h2;
            const float4 a0 = ld4(bn + 0 * C + ch), a1 = ld4(bn + 1 * C + ch), a2 = ld4(bn + 2 * C + ch);
            const float4 a3 = ld4(bn + 3 * C + ch), a4 = ld4(bn + 4 * C + ch), a5 = ld4(bn + 5 * C + ch);
            mu[4 * h2] = a0.x; mu[4 * h2 + 1] = a0.y; mu[4 * h2 + 2] = a0.z; mu[4 * h2 + 3] = a0.w;
            is[4 * h2] = a1.x; is[4 * h2 + 1] = a1.y; is[4 * h2 + 2] = a1.z; is[4 * h2 + 3] = a1.w;
            sc[4 * h2] = a2.x; sc[4 * h2 + 1] = a2.y; sc[4 * h2 + 2] = a2.z; sc[4 * h2 + 3] = a2.w;
            sh[4 * h2] = a3.x; sh[4 * h2 + 1] = a3.y; sh[4 * h2 + 2] = a3.z; sh[4 * h2 + 3] = a3.w;
            mg[4 * h2] = a4.x; mg[4 * h2 + 1] = a4.y; mg[4 * h2 + 2] = a4.z; mg[4 * h2 + 3] = a4.w;
            mgx[4 * h2] = a5.x; mgx[4 * h2 + 1] = a5.y; mgx[4 * h2 + 2] = a5.z; mgx[4 * h2 + 3] = a5.w;
        }
    }
    for (int pix = p0 + pl; pix < p1; pix += lanes) {
        const long long oe = (long long)pix * C + c;
        float g[8];
        {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(dv + oe)), g1 = __ldg(reinterpret_cast<const float4*>(dv + oe + 4));
            g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        }
        float vv[8];
        if (add_skip) load8t<HALF>(v, vh, oe, vv);
        long long ew[4] = {oe, oe, oe, oe};
        bool ok[4] = {true, false, false, false};
        unsigned code_lo = 0, code_hi = 0;
        if (pool) {
            const int r1 = pix / Wo, wo = pix - r1 * Wo;
            const int n = r1 / Ho, ho = r1 - n * Ho;
            const int h0 = 2 * ho, w0 = 2 * wo;
            const bool okh = h0 + 1 < H, okw = w0 + 1 < W;
            const long long row0 = ((long long)n * H + h0) * W, row1 = ((long long)n * H + (okh ? h0 + 1 : h0)) * W;
            const int wb = okw ? w0 + 1 : w0;
            ew[0] = (row0 + w0) * C + c; ew[1] = (row0 + wb) * C + c; ew[2] = (row1 + w0) * C + c; ew[3] = (row1 + wb) * C + c;
            ok[1] = okw; ok[2] = okh; ok[3] = okh && okw;
            const uint2 cd = __ldg(reinterpret_cast<const uint2*>(idx + oe));
            code_lo = cd.x;
            code_hi = cd.y;
        }
        float uw[4][8];
        load8t<HALF>(u, uh, ew[0], uw[0]);
        if (pool) {
            load8t<HALF>(u, uh, ew[1], uw[1]);
            load8t<HALF>(u, uh, ew[2], uw[2]);
            load8t<HALF>(u, uh, ew[3], uw[3]);
        }
        if (add_skip) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (!(vv[q] > 0.f)) g[q] = 0.f;
            if (HALF) store8t<true>(nullptr, dskiph, oe, g);      // the tensor-core consumers read only the bf16 form
            else {
                if (dskip) store8t<false>(dskip, nullptr, oe, g);
                if (dskiph) store8t<true>(nullptr, dskiph, oe, g);
            }
        }
        const int n_win = pool ? 4 : 1;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d >= n_win || !ok[d]) continue;
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned cq = ((q < 4 ? code_lo : code_hi) >> (8 * (q & 3))) & 0xffu;
                float gq = (!pool || cq == (unsigned)d) ? g[q] : 0.f;
                const float uq = uw[d][q];
                float r;
                if (has_bn) {
                    if (relu_mid && !(fmaf(uq, sc[q], sh[q]) > 0.f)) gq = 0.f;
                    const float xhat = (uq - mu[q]) * is[q];
                    r = sc[q] * (gq - mg[q] - xhat * mgx[q]);
                } else {
                    if (relu_mid && !(uq > 0.f)) gq = 0.f;
                    r = gq;
                }
                if (relu_in && !(uq > 0.f)) r = 0.f;
                o[q] = r;
            }
            if (HALF) store8t<true>(nullptr, duh, ew[d], o);
            else {
                if (du) store8t<false>(du, nullptr, ew[d], o);
                if (duh) store8t<true>(nullptr, duh, ew[d], o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// global average pooling: a thread owns 4 consecutive channels (C is a multiple of 16 for every conv stack of the space)
__global__ void __launch_bounds__(256) gap_fwd_kernel(const HeadTask* __restrict__ tasks, int n_tasks, int n_b) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const HeadTask& r) { return r.block_begin_fwd; });
    const HeadTask T = tasks[t];
    const int e4 = (blockIdx.x - T.block_begin_fwd) * 256 + threadIdx.x;
    const int C4 = T.C >> 2;
    if (e4 >= n_b * C4) return;
    const int n = e4 / C4, c = (e4 - n * C4) * 4;
    const int hw = T.Hf * T.Wf;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const long long base = (long long)n * hw * T.C + c;
    if (T.vh) {
#pragma unroll 5
        for (int p = 0; p < hw; ++p) {
            const uint2 r = __ldg(reinterpret_cast<const uint2*>(T.vh + base + (long long)p * T.C));
            s[0] += __uint_as_float(r.x << 16); s[1] += __uint_as_float(r.x & 0xffff0000u);
            s[2] += __uint_as_float(r.y << 16); s[3] += __uint_as_float(r.y & 0xffff0000u);
        }
    } else {
#pragma unroll 5
        for (int p = 0; p < hw; ++p) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(T.v + base + (long long)p * T.C));
            s[0] += r.x; s[1] += r.y; s[2] += r.z; s[3] += r.w;
        }
    }
    const float inv = (float)hw;
    *reinterpret_cast<float4*>(T.gap + (long long)n * T.C + c) = make_float4(s[0] / inv, s[1] / inv, s[2] / inv, s[3] / inv);
}

__global__ void __launch_bounds__(256) gap_bwd_kernel(const HeadTask* __restrict__ tasks, int n_tasks, int n_b) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const HeadTask& r) { return r.block_begin; });
    const HeadTask T = tasks[t];
    const int e4 = (blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    const int C4 = T.C >> 2, hw = T.Hf * T.Wf;
    if (e4 >= n_b * hw * C4) return;
    const int pix = e4 / C4, c = (e4 - pix * C4) * 4;
    const int n = pix / hw;
    const float4 g = __ldg(reinterpret_cast<const float4*>(T.dgap + n * T.C + c));
    const float inv = (float)hw;
    *reinterpret_cast<float4*>(T.dv + (long long)e4 * 4) = make_float4(g.x / inv, g.y / inv, g.z / inv, g.w / inv);
}

__global__ void __launch_bounds__(256) drop_fwd_kernel(const DropTask* __restrict__ tasks, int n_tasks, int n_b, int step,
                                                       int training, float rate) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const DropTask& r) { return r.block_begin; });
    const DropTask T = tasks[t];
    const int e = (blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    if (e >= n_b * T.units) return;
    float v = T.u[e];
    if (training && T.use_dropout) {
        const bool keep = hash_uniform(T.seed, (unsigned)T.layer, (unsigned)step, (unsigned)e) >= rate;
        v = keep ? v / (1.f - rate) : 0.f;
    }
    T.v[e] = v;
}

__global__ void __launch_bounds__(256) drop_bwd_kernel(const DropTask* __restrict__ tasks, int n_tasks, int n_b, int step,
                                                       float rate) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const DropTask& r) { return r.block_begin; });
    const DropTask T = tasks[t];
    const int e = (blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    if (e >= n_b * T.units) return;
    float g = T.dv[e];
    if (T.use_dropout) {
        const bool keep = hash_uniform(T.seed, (unsigned)T.layer, (unsigned)step, (unsigned)e) >= rate;
        g = keep ? g / (1.f - rate) : 0.f;
    }
    if (!(T.u[e] > 0.f)) g = 0.f;
    T.dz[e] = g;
}

// Keras sparse_categorical_crossentropy on softmax probabilities (clip 1e-7) + gradient w.r.t. logits.
// One CTA per candidate, one warp per sample (8 warps x 8 rounds); losses are summed in sample order.
__global__ void __launch_bounds__(256) ce_kernel(const CeTask* __restrict__ tasks, int n_b, int step, int training) {
    const CeTask T = tasks[blockIdx.x];
    __shared__ float s_loss[kBatch];
    __shared__ int s_ok[kBatch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = T.n_classes;
    const float lo = 1e-7f, hi = 1.f - 1e-7f;
    for (int r = warp; r < n_b; r += 8) {
        const int sidx = T.gather ? T.gather[T.gather_step * step + r] : (int)(T.label_step * step) + r;
        const int y = T.labels[sidx];
        const float* z = T.logits + (long long)r * C;
        float mx = -3.4e38f;
        int arg = 0x7fffffff;
        for (int j = lane; j < C; j += 32) {
            const float v = z[j];
            if (v > mx) {
                mx = v;
                arg = j;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) {
                mx = om;
                arg = oa;
            }
        }
        float se = 0.f;
        for (int j = lane; j < C; j += 32) se += expf(z[j] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float inv = 1.f / se;
        float sc = 0.f, sgp = 0.f;   // sum of clipped probs ; sum_j G_j p_j
        for (int j = lane; j < C; j += 32) sc += fminf(fmaxf(expf(z[j] - mx) * inv, lo), hi);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
        if (training) {
            for (int j = lane; j < C; j += 32) {
                const float p = expf(z[j] - mx) * inv;
                const float pc = fminf(fmaxf(p, lo), hi);
                const bool inside = p >= lo && p <= hi;
                const float G = inside ? (pc / sc - (j == y ? 1.f : 0.f)) / pc : 0.f;
                sgp = fmaf(G, p, sgp);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sgp += __shfl_xor_sync(0xffffffffu, sgp, o);
            const float invn = 1.f / (float)n_b;
            for (int j = lane; j < C; j += 32) {
                const float p = expf(z[j] - mx) * inv;
                const float pc = fminf(fmaxf(p, lo), hi);
                const bool inside = p >= lo && p <= hi;
                const float G = inside ? (pc / sc - (j == y ? 1.f : 0.f)) / pc : 0.f;
                T.dlogits[(long long)r * C + j] = p * (G - sgp) * invn;
            }
        }
        if (lane == 0) {
            const float py = fminf(fmaxf(expf(z[y] - mx) * inv, lo), hi);
            s_loss[r] = -(logf(py) - logf(sc));
            s_ok[r] = (arg == y) ? 1 : 0;
            if (T.pred) T.pred[sidx] = arg;
            if (T.confusion) atomicAdd(&T.confusion[(T.y_true_zero ? 0 : y) * C + arg], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ls = 0.0;
        int ok = 0;
        for (int r = 0; r < n_b; ++r) {
            ls += (double)s_loss[r];
            ok += s_ok[r];
        }
        T.acc[0] += ls;
        T.acc[1] += (double)n_b;
        T.acc[2] += (double)ok;
    }
}

// confusion_matrix(y_true, y_pred, labels=range(C)) of calculate_fpr (nsga_penalty.py:355): integer atomics, samples with a
// label outside [0, C) are dropped exactly as scikit-learn drops labels that are not in `labels`.
__global__ void __launch_bounds__(256) confusion_kernel(const int* __restrict__ y_true, const int* __restrict__ y_pred, int n,
                                                        int C, int* __restrict__ cm) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int t = y_true[i], p = y_pred[i];
        if (t >= 0 && t < C && p >= 0 && p < C) atomicAdd(&cm[(long long)t * C + p], 1);
    }
}

// 1 024 parameters per block: four consecutive elements per thread as 16-byte accesses when the candidate's four flat
// buffers are 16-byte aligned (they are: engine.cu takes them from a 256-byte-aligned arena), scalar otherwise / on the tail
__global__ void __launch_bounds__(256) adam_kernel(const AdamTask* __restrict__ tasks, int n_tasks, float alpha, float b1,
                                                   float b2, float eps) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const AdamTask& r) { return r.block_begin; });
    const AdamTask T = tasks[t];
    const int i = ((blockIdx.x - T.block_begin) * 256 + threadIdx.x) * 4;
    if (i >= T.n) return;
    auto upd = [&](float g, float& m, float& v, float& p) {
        m = b1 * m + (1.f - b1) * g;
        v = b2 * v + (1.f - b2) * g * g;
        p -= alpha * m / (sqrtf(v) + eps);
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) | reinterpret_cast<uintptr_t>(T.m) |
                           reinterpret_cast<uintptr_t>(T.v)) & 15) == 0;
    if (aligned && i + 3 < T.n) {
        const float4 g = *reinterpret_cast<const float4*>(T.g + i);
        float4 m = *reinterpret_cast<float4*>(T.m + i), v = *reinterpret_cast<float4*>(T.v + i), p = *reinterpret_cast<float4*>(T.p + i);
        upd(g.x, m.x, v.x, p.x);
        upd(g.y, m.y, v.y, p.y);
        upd(g.z, m.z, v.z, p.z);
        upd(g.w, m.w, v.w, p.w);
        *reinterpret_cast<float4*>(T.m + i) = m;
        *reinterpret_cast<float4*>(T.v + i) = v;
        *reinterpret_cast<float4*>(T.p + i) = p;
    } else {
        for (int j = i; j < min(i + 4, T.n); ++j) {
            float m = T.m[j], v = T.v[j], p = T.p[j];
            upd(T.g[j], m, v, p);
            T.m[j] = m;
            T.v[j] = v;
            T.p[j] = p;
        }
    }
}

// Per-epoch shuffle on the device: the harness-imposed "Keras shuffle" is a Fisher-Yates walk driven by the fmix32 counter
// hash (engine.cu make_permutation; cmoop_cnn_debug_permutation exports the same stream to the oracle).  The walk is
// sequential by construction (swap i with hash(i) % (i + 1), i = n-1 .. 1), so one thread per candidate runs it in shared
// memory (n = 3 072: ~0.2 ms, all candidates of the wave in parallel) and the warp writes the result out coalesced --
// no host-generated index arrays, no per-candidate upload.
__global__ void __launch_bounds__(32) perm_kernel(const PermTask* __restrict__ tasks, int epoch, int n) {
    extern __shared__ int s_perm[];
    const PermTask T = tasks[blockIdx.x];
    for (int i = threadIdx.x; i < n; i += 32) s_perm[i] = i;
    __syncwarp();
    if (threadIdx.x == 0) {
        const unsigned s = fmix32(T.seed_lo ^ 0x5bd1e995u * (unsigned)(epoch + 1)) ^ T.seed_hi;
        const unsigned s2 = fmix32(s ^ 0x27d4eb2fu);
        for (int i = n - 1; i > 0; --i) {
            const unsigned r = fmix32(s2 ^ (unsigned)i);
            const int j = (int)(r % (unsigned)(i + 1));
            const int t = s_perm[i];
            s_perm[i] = s_perm[j];
            s_perm[j] = t;
        }
    }
    __syncwarp();
    for (int i = threadIdx.x; i < n; i += 32) T.perm[i] = s_perm[i];
}

__global__ void __launch_bounds__(256) init_kernel(const InitTask* __restrict__ tasks, int n_tasks) {
    const int t = block_find_task(tasks, n_tasks, blockIdx.x, [](const InitTask& r) { return r.block_begin; });
    const InitTask T = tasks[t];
    const int i = (blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    if (i >= T.n) return;
    T.p[i] = T.kind == 0 ? (2.f * hash_uniform(T.seed, (unsigned)T.tensor, 0u, (unsigned)i) - 1.f) * T.limit : T.value;
}

inline int check() { return (int)cudaGetLastError(); }

}  // namespace

int Launch::conv(const ConvTask* tasks, int n, int tiles, int n_b, int step, void* st) {
    if (n == 0 || tiles == 0) return 0;
    conv_gemm_kernel<<<tiles, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, step);
    return check();
}
int Launch::wgrad(const WgradTask* tasks, int n, int tiles, int n_b, int step, void* st) {
    if (n == 0 || tiles == 0) return 0;
    conv_wgrad_kernel<<<tiles, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, step);
    return check();
}
int Launch::reduce(const ReduceTask* tasks, int n, int blocks, void* st) {
    if (n == 0 || blocks == 0) return 0;
    reduce_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n);
    return check();
}
int Launch::wt(const WtTask* tasks, int n, int blocks, void* st) {
    if (n == 0 || blocks == 0) return 0;
    wt_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n);
    return check();
}
int Launch::bn_finalize(const PostTask* tasks, int n, int max_c, int n_b, int training, float momentum, float eps, void* st) {
    if (n == 0) return 0;
    bn_finalize_kernel<<<dim3(n, (max_c + 7) / 8), 128, 0, (cudaStream_t)st>>>(tasks, n_b, training, momentum, eps);
    return check();
}
int Launch::post_blocks(long long out_pixels, int C) {
    const int pixb = post_pixels_per_block(C);
    return (int)((out_pixels + pixb - 1) / pixb);
}
int Launch::post_fwd(const PostTask* tasks, int n, int blocks, int n_b, void* st, const int* bt, bool half) {
    if (n == 0 || blocks == 0) return 0;
    if (half)
        post_fwd_kernel<true><<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, bt);
    else
        post_fwd_kernel<false><<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, bt);
    return check();
}
int Launch::bwd_pix(long long npix) { return npix >= 16384 ? 256 : (npix >= 4096 ? 128 : 32); }
int Launch::post_bwd_reduce(const PostTask* tasks, int n, int blocks, int n_b, void* st, const int* bt, bool half) {
    if (n == 0 || blocks == 0) return 0;
    if (half)
        post_bwd_reduce_kernel<true><<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, bt);
    else
        post_bwd_reduce_kernel<false><<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, bt);
    return check();
}
int Launch::bn_bwd_finalize(const PostTask* tasks, int n, int max_c, int n_b, void* st) {
    if (n == 0) return 0;
    bn_bwd_finalize_kernel<<<dim3(n, (max_c + 7) / 8), 128, 0, (cudaStream_t)st>>>(tasks, n_b);
    return check();
}
int Launch::post_bwd_apply(const PostTask* tasks, int n, int blocks, int n_b, void* st, const int* bt, bool half) {
    if (n == 0 || blocks == 0) return 0;
    if (half)
        post_bwd_apply_kernel<true><<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, bt);
    else
        post_bwd_apply_kernel<false><<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, bt);
    return check();
}
int Launch::gap_blocks(long long elems) { return (int)((elems / 4 + 255) / 256); }
int Launch::gap_fwd(const HeadTask* tasks, int n, int blocks, int n_b, void* st) {
    if (n == 0 || blocks == 0) return 0;
    gap_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b);
    return check();
}
int Launch::gap_bwd(const HeadTask* tasks, int n, int blocks, int n_b, void* st) {
    if (n == 0 || blocks == 0) return 0;
    gap_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b);
    return check();
}
int Launch::drop_fwd(const DropTask* tasks, int n, int blocks, int n_b, int step, int training, float rate, void* st) {
    if (n == 0 || blocks == 0) return 0;
    drop_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, step, training, rate);
    return check();
}
int Launch::drop_bwd(const DropTask* tasks, int n, int blocks, int n_b, int step, float rate, void* st) {
    if (n == 0 || blocks == 0) return 0;
    drop_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, n_b, step, rate);
    return check();
}
int Launch::ce(const CeTask* tasks, int n, int n_b, int step, int training, void* st) {
    if (n == 0) return 0;
    ce_kernel<<<n, 256, 0, (cudaStream_t)st>>>(tasks, n_b, step, training);
    return check();
}
int Launch::confusion(const int* y_true, const int* y_pred, int n, int C, int* cm, void* st) {
    const int blocks = n < 256 * 592 ? (n + 255) / 256 : 592;
    confusion_kernel<<<blocks > 0 ? blocks : 1, 256, 0, (cudaStream_t)st>>>(y_true, y_pred, n, C, cm);
    return (int)cudaGetLastError();
}

int Launch::adam(const AdamTask* tasks, int n, int blocks, float alpha, float b1, float b2, float eps, void* st) {
    if (n == 0 || blocks == 0) return 0;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n, alpha, b1, b2, eps);
    return check();
}
bool Launch::perm_ok(int n) { return n >= 1 && (size_t)n * sizeof(int) <= 200 * 1024; }
int Launch::perm(const PermTask* tasks, int n_tasks, int epoch, int n, void* st) {
    if (n_tasks == 0) return 0;
    const size_t smem = (size_t)n * sizeof(int);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    perm_kernel<<<n_tasks, 32, smem, (cudaStream_t)st>>>(tasks, epoch, n);
    return (int)cudaGetLastError();
}

int Launch::init(const InitTask* tasks, int n, int blocks, void* st) {
    if (n == 0 || blocks == 0) return 0;
    init_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n);
    return check();
}

}  // namespace cmoop_cnn
