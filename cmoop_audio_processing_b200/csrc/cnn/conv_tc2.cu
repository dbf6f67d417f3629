// tcgen05 convolution, patch-resident form (stride 1, 'same' padding; forward and data gradient).
//
// conv_tc.cu stages an im2col tile per K block: every input pixel travels from L2 to shared memory k*k times in
// 16-byte pieces (2.2 GB of staging for 76 MB of unique activations on the 25x20 residual block of a 32-candidate
// launch, i.e. the kernel ran at L2 sector throughput with the tensor pipe 10-13 % active).  Here the GEMM rows are
// *padded-linear* output positions q = (n*Hp + h + p)*Wp + (w + p) (Hp = H + 2p, Wp = W + 2p), so the input of tap
// (kh, kw) for row q is simply position q + (kh - p)*Wp + (kw - p): one shared-memory patch serves all k*k taps.
//   * patch: for a 16-channel sub-slab, positions [q0 - S, q0 + 256 + S) (S = p*Wp + p) as two 8-channel planes
//     [plane][position][16 B] -- the canonical K-major *no-swizzle* UMMA layout (core matrix = 8 positions x 16 B,
//     contiguous 128 B; SBO = 128 B, LBO = plane stride), whose start address may sit at ANY position, so the A
//     operand of a tap is the same buffer with the start address advanced by (kh*Wp + kw)*16 bytes.  Loaded once per
//     sub-slab by 256 producer threads with zero-filling 16-byte cp.async (padding, batch tail), 4-deep ring.
//   * weights: pre-arranged per (16-channel sub-slab, tap) as [plane][cout][16 B] by wt_bf16_kernel (modes 2 / 3),
//     one contiguous 32*bn-byte stage -> one 1-D bulk TMA (cp.async.bulk + mbarrier), 8-deep ring, own warp.
//   * one elected thread issues, per (sub-slab, tap), two UMMAs 128 x bn x 16 (the CTA's two 128-row tiles) into
//     two TMEM accumulators; tcgen05.commit frees the weight stage / the patch buffer.
//   * epilogue as in conv_tc.cu (tcgen05.ld, bias, ReLU, fp32 store, optional bf16 shadow); rows that are padding
//     positions or beyond the batch are skipped.
// Rows on padding positions cost (Hp*Wp)/(H*W) - 1 extra tensor work (16 % at 25x20, k = 3) in exchange for a k*k-fold
// cut of the staging traffic.  Grouped over candidates like every kernel here (cnn.cuh).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int P2_MT = 2;                    // 128-row tiles per CTA
constexpr int P2_ROWS = P2_MT * 128;
constexpr int P2_PB = 4;                    // patch sub-slab buffers
constexpr int P2_WS = 12;                   // weight stages of 4 KB = 128 / bn (sub-slab, tap) entries each: 48 KB in flight per
                                            // CTA, about the L2 latency at the 32 B/clk the tensor pipe consumes (the first
                                            // version's 8 single-tap stages left the MMA thread waiting on weights)
constexpr int P2_LOOK = 2;                  // sub-slabs whose copies are in flight per producer thread
constexpr int P2_PRODUCERS = 256;
constexpr int P2_THREADS = P2_PRODUCERS + 64;     // + MMA warp + weight-TMA warp
constexpr uint32_t P2_TMEM_COLS = 256;
constexpr int P2_W_STAGE = 32 * 128;        // bytes of a weight stage at bn = 128

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
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// K-major, no swizzle (cute::UMMA::LayoutType::SWIZZLE_NONE): core matrices of 8 rows x 16 B (128 contiguous bytes);
// LBO = byte distance between the two 8-element K chunks of a UMMA (plane stride), SBO = between 8-row groups.
__device__ __forceinline__ uint64_t make_desc_k_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
    return d;                           // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // D = F32
    d |= 1u << 7;                       // A = BF16
    d |= 1u << 10;                      // B = BF16
    d |= (uint32_t)(bn >> 3) << 17;     // N
    d |= (uint32_t)(128 >> 4) << 24;    // M
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
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(P2_THREADS, 2) conv_tc2_kernel(const TcConvTask* __restrict__ tasks, int n_tasks, int n_b,
                                                                 int step, int q_max) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ TcConvTask T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        int lo = 0, hi = n_tasks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tasks[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        T = tasks[lo];
    }
    __syncthreads();
    const int local = blockIdx.x - T.tile_begin;
    const int tm = local / T.tiles_n, tn = local - tm * T.tiles_n;
    const int p = T.pad, Hp = T.H + 2 * p, Wp = T.W + 2 * p, HpWp = Hp * Wp;
    const int Mq = n_b * HpWp;                         // < 2^31 for every supported shape (n_b <= 64)
    const int q0 = tm * P2_ROWS;
    if (q0 >= Mq) return;
    const int S = p * Wp + p;
    const int Q = P2_ROWS + 2 * S;                     // patch positions of this task (<= q_max)
    const int bn = T.bn, n0 = tn * bn;
    const int taps = T.k * T.k, n_cs = T.Cin >> 4;     // 16-channel sub-slabs
    const uint32_t w_stage_bytes = 32u * (uint32_t)bn;

    uint8_t* patch = smem_raw;                                          // [P2_PB][2][Q][16 B] (stride 32*q_max)
    uint8_t* wsm = patch + (size_t)P2_PB * 32 * q_max;                  // [P2_WS][P2_W_STAGE]
    int* src_off = reinterpret_cast<int*>(wsm + P2_WS * P2_W_STAGE);    // [q_max] element offset of a position's pixel, -1 = zero
    float* bias_s = reinterpret_cast<float*>(src_off + ((q_max + 3) & ~3));   // [128] this N tile's bias (0 without bias)
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);
    uint64_t* pfull = bars;                     // [P2_PB]  256 producer arrivals
    uint64_t* pempty = bars + P2_PB;            // [P2_PB]  tcgen05.commit
    uint64_t* wfull = bars + 2 * P2_PB;         // [P2_WS]  TMA transaction
    uint64_t* wempty = wfull + P2_WS;           // [P2_WS]  tcgen05.commit
    uint64_t* accum_bar = wempty + P2_WS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    if (tid == 0) {
        for (int s = 0; s < P2_PB; ++s) {
            mbar_init(&pfull[s], P2_PRODUCERS);
            mbar_init(&pempty[s], 1);
        }
        for (int s = 0; s < P2_WS; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == P2_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(P2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < bn) bias_s[tid] = T.bias ? __ldg(T.bias + n0 + tid) : 0.f;
    // element offsets of the patch positions (shared by every sub-slab)
    for (int i = tid; i < Q; i += P2_THREADS) {
        const int q = q0 - S + i;
        int off = -1;
        if (q >= 0 && q < Mq) {
            const int n = q / HpWp, rem = q - n * HpWp;
            const int hp = rem / Wp, wp = rem - hp * Wp;
            if (hp >= p && hp < T.H + p && wp >= p && wp < T.W + p)
                off = ((n * T.H + hp - p) * T.W + wp - p) * T.Cin;
        }
        src_off[i] = off;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < P2_PRODUCERS / 32) {
        // ================= producers: one 16-channel sub-slab of the patch per ring slot =================
        const __nv_bfloat16* xbase = T.xh + T.x_step * step;
        const int plane = tid & 1, pix0 = tid >> 1;
        for (int cs = 0; cs < n_cs + P2_LOOK; ++cs) {
            if (cs < n_cs) {
                const int b = cs % P2_PB;
                mbar_wait(&pempty[b], (((uint32_t)(cs / P2_PB)) & 1u) ^ 1u);
                uint8_t* dst = patch + (size_t)b * 32 * q_max + (size_t)plane * 16 * Q;
                const int coff = cs * 16 + plane * 8;
                for (int i = pix0; i < Q; i += 128) {
                    const int off = src_off[i];
                    cp_async16(dst + (size_t)i * 16, off >= 0 ? xbase + off + coff : xbase, off >= 0 ? 16u : 0u);
                }
            }
            cp_async_commit();
            if (cs >= P2_LOOK) {
                cp_async_wait<P2_LOOK>();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&pfull[(cs - P2_LOOK) % P2_PB]);
            }
        }
        // ================= epilogue =================
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lane_grp = warp & 3;
        const int c_begin = bn >= 32 ? (warp >> 2) * (bn / 2) : 0;
        const int c_end = bn >= 32 ? c_begin + bn / 2 : ((warp >> 2) == 0 ? bn : 0);
#pragma unroll 1
        for (int mt = 0; mt < P2_MT; ++mt) {
            const int q = q0 + mt * 128 + lane_grp * 32 + lane;
            long long obase = -1;
            if (q < Mq) {
                const int n = q / HpWp, rem = q - n * HpWp;
                const int hp = rem / Wp, wp = rem - hp * Wp;
                if (hp >= p && hp < T.H + p && wp >= p && wp < T.W + p)
                    obase = ((long long)(n * T.H + hp - p) * T.W + wp - p) * T.Cout;
            }
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(mt * 128 + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (obase >= 0) {
                    float* dst = T.y + obase + n0 + c0;
#pragma unroll
                    for (int qd = 0; qd < 16; qd += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + qd);
                        float4 o;
                        o.x = __uint_as_float(v[qd + 0]) + bb.x;
                        o.y = __uint_as_float(v[qd + 1]) + bb.y;
                        o.z = __uint_as_float(v[qd + 2]) + bb.z;
                        o.w = __uint_as_float(v[qd + 3]) + bb.w;
                        if (T.relu) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        float4* d4 = reinterpret_cast<float4*>(dst + qd);
                        if (T.accumulate) {
                            const float4 old = *d4;
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *d4 = o;
                        if (T.yh)
                            *reinterpret_cast<uint2*>(T.yh + obase + n0 + c0 + qd) =
                                make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == P2_PRODUCERS / 32) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(bn);
            const uint32_t lbo_a = 16u * (uint32_t)Q, lbo_b = 16u * (uint32_t)bn;
            const int tps = 128 / bn;                  // (sub-slab, tap) entries per 4 KB weight stage
            const int total = n_cs * taps;
            int st = 0, e = 0;                         // stage counter; entry = cs * taps + tap in weight-stream order
            for (int cs = 0; cs < n_cs; ++cs) {
                const int b = cs % P2_PB;
                mbar_wait(&pfull[b], ((uint32_t)(cs / P2_PB)) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base = smem_u32(patch + (size_t)b * 32 * q_max);
                int kh = 0, kw = 0;
                for (int t = 0; t < taps; ++t, ++e) {
                    const int sub = e % tps, ws = st % P2_WS;
                    if (sub == 0) {
                        mbar_wait(&wfull[ws], ((uint32_t)(st / P2_WS)) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t shift = (uint32_t)(kh * Wp + kw) * 16u;
                    const uint64_t bd = make_desc_k_none(smem_u32(wsm + ws * P2_W_STAGE) + (uint32_t)sub * w_stage_bytes, lbo_b, 128);
#pragma unroll
                    for (int mt = 0; mt < P2_MT; ++mt) {
                        const uint64_t ad = make_desc_k_none(a_base + shift + (uint32_t)mt * 128u * 16u, lbo_a, 128);
                        umma_bf16(tmem_base + (uint32_t)(mt * 128), ad, bd, idesc, (cs | t) != 0 ? 1u : 0u);
                    }
                    if (sub == tps - 1 || e == total - 1) {        // last entry of the stage (or of the tile)
                        umma_commit(&wempty[ws]);
                        ++st;
                    }
                    if (++kw == T.k) { kw = 0; ++kh; }
                }
                umma_commit(&pempty[b]);
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        // ================= weight stages: one bulk copy per (sub-slab, tap) =================
        if (lane == 0) {
            const uint32_t total_bytes = (uint32_t)(n_cs * taps) * w_stage_bytes;
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(T.wt) + (size_t)tn * total_bytes;
            int st = 0;
            for (uint32_t o = 0; o < total_bytes; o += P2_W_STAGE, ++st) {
                const int ws = st % P2_WS;
                const uint32_t bytes = total_bytes - o < (uint32_t)P2_W_STAGE ? total_bytes - o : (uint32_t)P2_W_STAGE;
                mbar_wait(&wempty[ws], (((uint32_t)(st / P2_WS)) & 1u) ^ 1u);
                tma_load_1d(wsm + ws * P2_W_STAGE, wsrc + o, bytes, &wfull[ws]);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == P2_PRODUCERS / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P2_TMEM_COLS) : "memory");
    }
}

// bf16 weight stages for conv_tc2_kernel: out[tn][cs][tap][plane][n][8] (n < bn = min(Cout_gemm, 128)).
//   mode 2 (forward):       B[n = co][k = (tap, ci)]        = w[(kh, kw), ci, co]
//   mode 3 (data gradient): B[n = ci][k = (tap, co)]        = w[(k-1-kh, k-1-kw), ci, co]
__global__ void __launch_bounds__(256) wt_bf16_v2_kernel(const WtBf16Task* __restrict__ tasks, int n_tasks) {
    int lo = 0, hi = n_tasks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tasks[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const WtBf16Task T = tasks[lo];
    const int gi = T.mode == 2 ? T.Cin : T.Cout, go = T.mode == 2 ? T.Cout : T.Cin;     // GEMM input / output channels
    const int taps = T.k * T.k;
    const long long chunks = (long long)taps * gi * go / 8;             // one thread = one 16-byte chunk (8 input channels)
    const long long c = (long long)(blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    if (c >= chunks) return;
    const int bn = go < 128 ? go : 128, n_cs = gi >> 4;
    // c -> (tn, cs, tap, plane, n): consecutive threads take consecutive output channels, so the forward reads
    // (stride Cout between a thread's 8 values) coalesce across the warp and the 16-byte stores are contiguous
    long long r = c;
    const int n = (int)(r % bn); r /= bn;
    const int plane = (int)(r % 2); r /= 2;
    const int tap = (int)(r % taps); r /= taps;
    const int cs = (int)(r % n_cs); r /= n_cs;
    const int tn = (int)r;
    const int ch_in0 = cs * 16 + plane * 8, ch_out = tn * bn + n;
    const int kh = tap / T.k, kw = tap - kh * T.k;
    float v[8];
    if (T.mode == 2) {
        const float* src = T.w + ((long long)(kh * T.k + kw) * T.Cin + ch_in0) * T.Cout + ch_out;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (long long)j * T.Cout);
    } else {
        const float* src = T.w + ((long long)((T.k - 1 - kh) * T.k + (T.k - 1 - kw)) * T.Cin + ch_out) * T.Cout + ch_in0;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(src)), a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    reinterpret_cast<uint4*>(T.out)[c] = o;
}

size_t p2_smem_bytes(int q_max) {
    return (size_t)P2_PB * 32 * q_max + (size_t)P2_WS * P2_W_STAGE + (size_t)((q_max + 3) & ~3) * 4 + 512 + 512;
}

}  // namespace

int Launch::tc2_q(int W, int k) {
    const int p = (k - 1) / 2;
    return P2_ROWS + 2 * (p * (W + 2 * p) + p);
}
int Launch::tc2_rows() { return P2_ROWS; }
bool Launch::tc2_ok(int H, int W, int Cin, int Cout, int k, int stride) {
    (void)H;
    if (stride != 1 || (k & 1) == 0 || Cin % 16 != 0 || Cout % 16 != 0) return false;
    if (Cout > 128 && Cout % 128 != 0) return false;
    return p2_smem_bytes(tc2_q(W, k)) <= 110 * 1024;          // two CTAs per SM
}

int Launch::conv_tc2(const TcConvTask* tasks, int n, int tiles, int n_b, int step, int q_max, void* st) {
    if (n == 0 || tiles == 0) return 0;
    const size_t smem = p2_smem_bytes(q_max);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    conv_tc2_kernel<<<tiles, P2_THREADS, smem, (cudaStream_t)st>>>(tasks, n, n_b, step, q_max);
    return (int)cudaGetLastError();
}

int Launch::wt_bf16_v2(const WtBf16Task* tasks, int n, int blocks, void* st) {
    if (n == 0 || blocks == 0) return 0;
    wt_bf16_v2_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
