// tcgen05 implicit-GEMM convolution (forward and data-gradient) for the population-batched CNN.
//
// Replaces the cuDNN/cuBLAS library calls TensorFlow makes for layers.Conv2D inside model.fit /
// model.predict (nsga_penalty.py:255-330, 383-386) with a hand-written sm_100a kernel:
//   D[128 x BN] (fp32, TMEM) += A[128 x 64] (bf16, smem, K-major, SWIZZLE_128B) * B[BN x 64]^T (bf16, smem)
//   * A = im2col tile: 128 output pixels x 64 consecutive (kh,kw,ci) taps, gathered from the fp32 NHWC
//     shadow copy of the activations by 4 producer warps with 16-byte cp.async (zero-fill for padding), two
//     stages in flight per thread, written with the 128-byte XOR swizzle the UMMA descriptor expects
//   * B = pre-transposed bf16 weights [Cout][K_pad] (refreshed once per optimiser step by wt_bf16_kernel;
//     the data-gradient uses the spatially flipped, channel-transposed copy [Cin][K'_pad])
//   * one elected thread of warp 4 issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1), stages are
//     recycled through tcgen05.commit -> mbarrier; 3-stage ring, 2 CTAs per SM, 8 producer warps per CTA
//   * epilogue: the 8 producer warps (lane group = warp % 4, column half = warp / 4) read TMEM with tcgen05.ld (32x32b.x16), add the fp32
//     bias, apply ReLU and store fp32 rows (optionally strided/accumulating for the 1x1/s2 skip dgrad)
// Grouped over candidates exactly like the SIMT kernels (cnn.cuh): blockIdx.x -> (task, tile).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;        // 16 KiB
constexpr int TC_B_BYTES = 128 * TC_BK * 2;          // 16 KiB (BN <= 128)
constexpr int TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_PRODUCERS = 256;          // 8 producer / epilogue warps
constexpr int TC_THREADS = TC_PRODUCERS + 32;   // + the MMA-issuing warp
constexpr uint32_t TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// rows of 128 B, 8-row groups 1024 B apart (SBO), LBO = 1 (unused for swizzled K-major).
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 1u << 7;                       // a_format = BF16
    d |= 1u << 10;                      // b_format = BF16
    d |= (uint32_t)(bn >> 3) << 17;     // n_dim
    d |= (uint32_t)(TC_BM >> 4) << 24;  // m_dim
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous global->shared copy; src_bytes == 0 zero-fills (padding / out-of-range rows)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int TC_LOOKAHEAD = 2;   // stages whose copies are in flight per producer thread

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const TcConvTask* __restrict__ tasks, int n_tasks,
                                                                int n_b, int step, const int* __restrict__ block_task) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ TcConvTask T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < (int)(sizeof(TcConvTask) / 4)) {
        int lo = 0;
        if (block_task) {                        // direct block -> task table (one load instead of a binary search)
            lo = __ldg(block_task + blockIdx.x);
        } else {
            int hi = n_tasks - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (tasks[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
            }
        }
        reinterpret_cast<uint32_t*>(&T)[tid] = reinterpret_cast<const uint32_t*>(tasks + lo)[tid];
    }
    __syncthreads();
    const int local = blockIdx.x - T.tile_begin;
    const int tm = local / T.tiles_n, tn = local - tm * T.tiles_n;
    const int HoWo = T.Ho * T.Wo, M = n_b * HoWo, m0 = tm * TC_BM, n0 = tn * T.bn;
    if (m0 >= M) return;
    const int K = T.k * T.k * T.Cin;
    const int num_kb = T.K_pad / TC_BK;

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_smem = base;
    uint8_t* b_smem = base + TC_STAGES * TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));
    uint64_t* full_bar = bars;                // [TC_STAGES], 128 producer arrivals
    uint64_t* empty_bar = bars + TC_STAGES;   // [TC_STAGES], 1 arrival (tcgen05.commit)
    uint64_t* accum_bar = bars + 2 * TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], TC_PRODUCERS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCERS / 32) {
        // ================= producers: im2col gather + weight tile, cp.async straight into the swizzled stage ====
        const __nv_bfloat16* xbase = T.xh + T.x_step * step;
        const int j = tid & 7;                 // 16-byte chunk (8 bf16) inside the 64-wide K block
        const int r_lo = tid >> 3;             // rows r_lo + 32*i
        const long long img = (long long)T.H * T.W * T.Cin;
        const __nv_bfloat16* xrow[4];
        int hi0[4], wi0[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + r_lo + 32 * i;
            if (m < M) {
                const int n = m / HoWo, r = m - n * HoWo;
                const int ho = r / T.Wo, wo = r - ho * T.Wo;
                hi0[i] = ho * T.stride - T.pad;
                wi0[i] = wo * T.stride - T.pad;
                xrow[i] = xbase + (long long)n * img;
            } else {
                xrow[i] = nullptr;
                hi0[i] = wi0[i] = 0;
            }
        }
        const int b_rows = (T.bn + 31) / 32;   // weight rows handled per thread (rows r_lo + 32*i < bn)
        for (int kb = 0; kb < num_kb + TC_LOOKAHEAD; ++kb) {
            if (kb < num_kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t* a_st = a_smem + s * TC_A_BYTES;
                uint8_t* b_st = b_smem + s * TC_B_BYTES;
                const int k = kb * TC_BK + j * 8;
                int kh = 0, kw = 0, ci = 0;
                const bool kvalid = k < K;
                if (kvalid) {
                    const int pos = k / T.Cin;
                    ci = k - pos * T.Cin;
                    kh = pos / T.k;
                    kw = pos - kh * T.k;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r_lo + 32 * i;
                    const __nv_bfloat16* src = xbase;
                    uint32_t bytes = 0;
                    if (kvalid && xrow[i] != nullptr) {
                        const int hi = hi0[i] + kh, wi = wi0[i] + kw;
                        if ((unsigned)hi < (unsigned)T.H && (unsigned)wi < (unsigned)T.W) {
                            src = xrow[i] + ((long long)hi * T.W + wi) * T.Cin + ci;
                            bytes = 16;
                        }
                    }
                    cp_async16(a_st + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4), src, bytes);
                }
                for (int i = 0; i < b_rows; ++i) {
                    const int r = r_lo + 32 * i;
                    if (r < T.bn)
                        cp_async16(b_st + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4),
                                   T.wt + (long long)(n0 + r) * T.K_pad + kb * TC_BK + j * 8, 16);
                }
            }
            cp_async_commit();
            if (kb >= TC_LOOKAHEAD) {
                cp_async_wait<TC_LOOKAHEAD>();                               // the copies of stage kb - LOOKAHEAD have landed
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // ... and are visible to the MMA (async proxy)
                mbar_arrive(&full_bar[(kb - TC_LOOKAHEAD) % TC_STAGES]);
            }
        }
        // ================= epilogue: TMEM -> registers -> global =================
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lane_grp = warp & 3;                          // TMEM lanes 32*lane_grp .. +31 (warp-id mod 4 rule)
        const int c_begin = T.bn >= 32 ? (warp >> 2) * (T.bn / 2) : 0;
        const int c_end = T.bn >= 32 ? c_begin + T.bn / 2 : ((warp >> 2) == 0 ? T.bn : 0);
        const int row = lane_grp * 32 + lane;
        const int mi = m0 + row;
        long long obase = 0;
        if (mi < M) {
            if (T.out_s == 0) {
                obase = (long long)mi * T.Cout;
            } else {
                const int n = mi / HoWo, r = mi - n * HoWo;
                const int ho = r / T.Wo, wo = r - ho * T.Wo;
                obase = (((long long)n * T.out_h + ho * T.out_s) * T.out_w + wo * T.out_s) * T.Cout;
            }
        }
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (mi < M) {
                float* dst = T.y + obase + n0 + c0;
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    float4 o;
                    o.x = __uint_as_float(v[q + 0]);
                    o.y = __uint_as_float(v[q + 1]);
                    o.z = __uint_as_float(v[q + 2]);
                    o.w = __uint_as_float(v[q + 3]);
                    if (T.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(T.bias + n0 + c0 + q));
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    if (T.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    if (T.y) {
                        float4* d4 = reinterpret_cast<float4*>(dst + q);
                        if (T.accumulate) {
                            const float4 old = *d4;
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *d4 = o;
                    }
                    if (T.yh)
                        *reinterpret_cast<uint2*>(T.yh + obase + n0 + c0 + q) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ================= MMA issuer (warp 4, one elected lane) =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(T.bn);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(a_smem + s * TC_A_BYTES);
                const uint32_t b_addr = smem_u32(b_smem + s * TC_B_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                    const uint64_t ad = make_desc_k_sw128(a_addr + k4 * 32);
                    const uint64_t bd = make_desc_k_sw128(b_addr + k4 * 32);
                    umma_bf16(tmem_base, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);     // frees the stage when these MMAs have read it
            }
            umma_commit(accum_bar);             // accumulator complete
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == TC_PRODUCERS / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// MN-major, SWIZZLE_128B descriptor: 64-element (128-byte) groups along M/N `lbo` bytes apart, 8-row groups
// along K 1024 bytes apart (rows = K index, 128 bytes each).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Weight gradient on tcgen05:  dW[kk][co] (+ bias row kk == K) = sum_m im2col(x)[m][kk] * dy[m][co].
// The reduction runs over pixels, so both operands are MN-major: a stage holds 64 pixels (UMMA K) as rows of
// 128 bytes -- A: two 64-wide kk groups (UMMA M = 128), B: bn/64 co groups -- which is exactly the natural
// [pixel][channel] order of the im2col row and of dy.  Split-M partials are written per split (deterministic).
__global__ void __launch_bounds__(TC_THREADS, 2) wgrad_tc_kernel(const TcWgradTask* __restrict__ tasks, int n_tasks,
                                                                 int n_b, const int* __restrict__ block_task) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ TcWgradTask T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < (int)(sizeof(TcWgradTask) / 4)) {
        int lo = 0;
        if (block_task) {                        // direct block -> task table (one load instead of a binary search)
            lo = __ldg(block_task + blockIdx.x);
        } else {
            int hi = n_tasks - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (tasks[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
            }
        }
        reinterpret_cast<uint32_t*>(&T)[tid] = reinterpret_cast<const uint32_t*>(tasks + lo)[tid];
    }
    __syncthreads();
    int local = blockIdx.x - T.tile_begin;
    const int per_split = T.tiles_k * T.tiles_n;
    const int split = local / per_split;
    local -= split * per_split;
    const int tk = local / T.tiles_n, tn = local - tk * T.tiles_n;
    const int HoWo = T.Ho * T.Wo, M = n_b * HoWo;
    const int K = T.k * T.k * T.Cin, Kext = K + 1;
    const int kk0 = tk * 128, n0 = tn * T.bn;
    const int m_begin = split * T.m_chunk;
    const int m_end = min(M, m_begin + T.m_chunk);
    float* out = T.out + (long long)split * Kext * T.Cout;
    const int num_st = m_end > m_begin ? (m_end - m_begin + 63) / 64 : 0;
    if (num_st == 0) {                      // empty split (partial last batch): its partial must still be zero
        for (int e = tid; e < 128 * T.bn; e += TC_THREADS) {
            const int r = e / T.bn, c = e - r * T.bn;
            if (kk0 + r < Kext) out[(long long)(kk0 + r) * T.Cout + n0 + c] = 0.f;
        }
        return;
    }

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_smem = base;
    uint8_t* b_smem = base + TC_STAGES * TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + TC_STAGES;
    uint64_t* accum_bar = bars + 2 * TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], TC_PRODUCERS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCERS / 32) {
        // ---- producers: thread owns chunk column c16 (8 consecutive kk / co) and rows r_lo + 8 i of the stage
        const int c16 = tid & 15, r_lo = tid >> 4;      // rows r_lo + 16*i
        const long long img = (long long)T.H * T.W * T.Cin;
        const int kk = kk0 + c16 * 8;
        int kh = 0, kw = 0, ci = 0, a_kind;             // a_kind 0: zeros, 1: im2col taps, 2: chunk containing the ones column
        if (kk + 7 < K) {
            a_kind = 1;
            const int pos = kk / T.Cin;
            ci = kk - pos * T.Cin;
            kh = pos / T.k;
            kw = pos - kh * T.k;
        } else {
            a_kind = (kk <= K && K < kk + 8) ? 2 : 0;   // K is a multiple of 8, so the ones column starts a chunk
        }
        const bool b_on = c16 * 8 < T.bn;
        const uint32_t a_off = (uint32_t)(c16 >> 3) * 8192u, a_chunk = (uint32_t)(c16 & 7);
        for (int st = 0; st < num_st + TC_LOOKAHEAD; ++st) {
            if (st < num_st) {
                const int s = st % TC_STAGES;
                const uint32_t ph = (uint32_t)(st / TC_STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t* a_st = a_smem + s * TC_A_BYTES;
                uint8_t* b_st = b_smem + s * TC_B_BYTES;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r_lo + 16 * i;
                    const int m = m_begin + st * 64 + r;
                    const uint32_t row_off = (uint32_t)r * 128u + ((a_chunk ^ (uint32_t)(r & 7)) << 4);
                    const __nv_bfloat16* asrc = T.xh;
                    uint32_t abytes = 0;
                    const bool mvalid = m < m_end;
                    if (mvalid && a_kind == 1) {
                        const int n = m / HoWo, rr = m - n * HoWo;
                        const int ho = rr / T.Wo, wo = rr - ho * T.Wo;
                        const int hi = ho * T.stride - T.pad + kh, wi = wo * T.stride - T.pad + kw;
                        if ((unsigned)hi < (unsigned)T.H && (unsigned)wi < (unsigned)T.W) {
                            asrc = T.xh + (long long)n * img + ((long long)hi * T.W + wi) * T.Cin + ci;
                            abytes = 16;
                        }
                    }
                    if (a_kind == 2)       // the ones column (bias-gradient row): bf16(1.0) in the first element
                        *reinterpret_cast<uint4*>(a_st + a_off + row_off) = make_uint4(mvalid ? 0x00003F80u : 0u, 0u, 0u, 0u);
                    else
                        cp_async16(a_st + a_off + row_off, asrc, abytes);
                    if (b_on)
                        cp_async16(b_st + a_off + row_off, mvalid ? T.dyh + (long long)m * T.Cout + n0 + c16 * 8 : T.dyh,
                                   mvalid ? 16u : 0u);
                }
            }
            cp_async_commit();
            if (st >= TC_LOOKAHEAD) {
                cp_async_wait<TC_LOOKAHEAD>();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&full_bar[(st - TC_LOOKAHEAD) % TC_STAGES]);
            }
        }
        // ---- epilogue: D rows = kk, columns = co
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lane_grp = warp & 3;
        const int c_begin = T.bn >= 32 ? (warp >> 2) * (T.bn / 2) : 0;
        const int c_end = T.bn >= 32 ? c_begin + T.bn / 2 : ((warp >> 2) == 0 ? T.bn : 0);
        const int row = lane_grp * 32 + lane;
        const int krow = kk0 + row;
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (krow < Kext) {
                float* dst = out + (long long)krow * T.Cout + n0 + c0;
#pragma unroll
                for (int q = 0; q < 16; q += 4)
                    *reinterpret_cast<float4*>(dst + q) = make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]),
                                                                      __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        if (lane == 0) {
            uint32_t idesc = make_idesc_bf16(T.bn);
            idesc |= (1u << 15) | (1u << 16);            // A and B are MN-major
            for (int st = 0; st < num_st; ++st) {
                const int s = st % TC_STAGES;
                const uint32_t ph = (uint32_t)(st / TC_STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(a_smem + s * TC_A_BYTES);
                const uint32_t b_addr = smem_u32(b_smem + s * TC_B_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {           // 16 pixels (rows) per MMA: 2 KiB further down
                    const uint64_t ad = make_desc_mn_sw128(a_addr + k4 * 2048, 8192);
                    const uint64_t bd = make_desc_mn_sw128(b_addr + k4 * 2048, 8192);
                    umma_bf16(tmem_base, ad, bd, idesc, (st | k4) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == TC_PRODUCERS / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// bf16 K-major weight copies: mode 0 forward [Cout][K_pad], mode 1 data-gradient [Cin][K'_pad]
__global__ void __launch_bounds__(256) wt_bf16_kernel(const WtBf16Task* __restrict__ tasks, int n_tasks) {
    int lo = 0, hi = n_tasks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tasks[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const WtBf16Task T = tasks[lo];
    const long long e = (long long)(blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    const int rows = T.mode == 0 ? T.Cout : T.Cin;
    if (e >= (long long)rows * T.K_pad) return;
    const int kk = (int)(e % T.K_pad);
    const int row = (int)(e / T.K_pad);
    float v = 0.f;
    if (T.mode == 0) {
        if (kk < T.k * T.k * T.Cin) v = T.w[(long long)kk * T.Cout + row];
    } else {
        if (kk < T.k * T.k * T.Cout) {
            const int co = kk % T.Cout;
            const int pos = kk / T.Cout;
            const int kh = pos / T.k, kw = pos - kh * T.k;
            v = T.w[((long long)((T.k - 1 - kh) * T.k + (T.k - 1 - kw)) * T.Cin + row) * T.Cout + co];
        }
    }
    T.out[e] = __float2bfloat16_rn(v);
}

// per-64-row-tile column sums of y and y^2 (same [tile][2][C] layout the SIMT conv epilogue writes).
// bf16-only activations (the only caller today): a thread owns 8 consecutive channels of some rows of the tile (one 16-byte
// load per row instead of one 2-byte load per element), row lanes are combined through shared memory in lane order.
__global__ void __launch_bounds__(128) bn_stats_kernel(const StatTask* __restrict__ tasks, int n_tasks, int n_b) {
    __shared__ float red[2][128 * 8];
    int lo = 0, hi = n_tasks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tasks[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const StatTask T = tasks[lo];
    const int tile = blockIdx.x - T.block_begin;
    const long long M = (long long)n_b * T.rows_per_sample;
    const long long r0 = (long long)tile * 64;
    if (r0 >= M) return;
    const long long r1 = r0 + 64 < M ? r0 + 64 : M;
    if (T.yh && (T.C & 7) == 0) {
        const int cgs = T.C >> 3;                         // 8-channel groups
        const int per = cgs < 128 ? cgs : 128;            // groups handled per pass
        const int lanes = 128 / per;
        const int g_lane = threadIdx.x % per, p_lane = threadIdx.x / per;
        for (int g0 = 0; g0 < cgs; g0 += per) {
            const int c = (g0 + g_lane) * 8;
            float s1[8], s2[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) s1[q] = s2[q] = 0.f;
            if (p_lane < lanes)
                for (long long r = r0 + p_lane; r < r1; r += lanes) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(T.yh + r * T.C + c));
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float a = __uint_as_float(w[q] << 16), b = __uint_as_float(w[q] & 0xffff0000u);
                        s1[2 * q] += a; s2[2 * q] = fmaf(a, a, s2[2 * q]);
                        s1[2 * q + 1] += b; s2[2 * q + 1] = fmaf(b, b, s2[2 * q + 1]);
                    }
                }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                red[0][threadIdx.x * 8 + q] = s1[q];
                red[1][threadIdx.x * 8 + q] = s2[q];
            }
            __syncthreads();
            if (p_lane == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float a1 = 0.f, a2 = 0.f;
                    for (int l = 0; l < lanes; ++l) {
                        a1 += red[0][(l * per + g_lane) * 8 + q];
                        a2 += red[1][(l * per + g_lane) * 8 + q];
                    }
                    T.part[((long long)tile * 2 + 0) * T.C + c + q] = a1;
                    T.part[((long long)tile * 2 + 1) * T.C + c + q] = a2;
                }
            }
            __syncthreads();
        }
        return;
    }
    const int cb = T.C < 128 ? T.C : 128;
    const int lanes = 128 / cb;
    const int p_lane = threadIdx.x / cb, c_lane = threadIdx.x - p_lane * cb;
    for (int c0 = 0; c0 < T.C; c0 += cb) {
        const int c = c0 + c_lane;
        float s1 = 0.f, s2 = 0.f;
        if (p_lane < lanes && c < T.C)
            for (long long r = r0 + p_lane; r < r1; r += lanes) {
                const float v = T.yh ? __bfloat162float(T.yh[r * T.C + c]) : T.y[r * T.C + c];
                s1 += v;
                s2 = fmaf(v, v, s2);
            }
        red[0][threadIdx.x] = s1;
        red[1][threadIdx.x] = s2;
        __syncthreads();
        if (p_lane == 0 && c < T.C) {
            float a1 = 0.f, a2 = 0.f;
            for (int l = 0; l < lanes; ++l) {
                a1 += red[0][l * cb + c_lane];
                a2 += red[1][l * cb + c_lane];
            }
            T.part[((long long)tile * 2 + 0) * T.C + c] = a1;
            T.part[((long long)tile * 2 + 1) * T.C + c] = a2;
        }
        __syncthreads();
    }
}

}  // namespace

int Launch::conv_tc(const TcConvTask* tasks, int n, int tiles, int n_b, int step, void* st, const int* block_task) {
    if (n == 0 || tiles == 0) return 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    conv_tc_kernel<<<tiles, TC_THREADS, TC_SMEM, (cudaStream_t)st>>>(tasks, n, n_b, step, block_task);
    return (int)cudaGetLastError();
}
int Launch::wgrad_tc(const TcWgradTask* tasks, int n, int tiles, int n_b, void* st, const int* block_task) {
    if (n == 0 || tiles == 0) return 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    wgrad_tc_kernel<<<tiles, TC_THREADS, TC_SMEM, (cudaStream_t)st>>>(tasks, n, n_b, block_task);
    return (int)cudaGetLastError();
}
int Launch::wt_bf16(const WtBf16Task* tasks, int n, int blocks, void* st) {
    if (n == 0 || blocks == 0) return 0;
    wt_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n);
    return (int)cudaGetLastError();
}
int Launch::bn_stats(const StatTask* tasks, int n, int blocks, int n_b, void* st) {
    if (n == 0 || blocks == 0) return 0;
    bn_stats_kernel<<<blocks, 128, 0, (cudaStream_t)st>>>(tasks, n, n_b);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
