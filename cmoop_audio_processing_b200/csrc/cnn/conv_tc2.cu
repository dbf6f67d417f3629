// tcgen05 convolution, patch-resident form (stride 1, 'same' padding; forward and data gradient).
//
// conv_tc.cu stages an im2col tile per K block: every input pixel travels from L2 to shared memory k*k times in
// 16-byte pieces (2.2 GB of staging for 76 MB of unique activations on the 25x20 residual block of a 32-candidate
// launch, i.e. the kernel ran at L2 sector throughput with the tensor pipe 10-13 % active).  Here the GEMM rows are
// *padded-linear* output positions q = (n*Hp + h + p)*Wp + (w + p) (Hp = H + 2p, Wp = W + 2p), so the input of tap
// (kh, kw) for row q is simply position q + (kh - p)*Wp + (kw - p): one shared-memory patch serves all k*k taps.
//   * patch: for a 64-channel slab, positions [q0 - S, q0 + 256 + S) (S = p*Wp + p) as rows of 128 B in the K-major
//     SWIZZLE_128B UMMA layout (16-byte chunk c of row i at i*128 + ((c ^ (i & 7)) << 4), buffer 1 KB aligned).  The A
//     operand of tap (kh, kw) is the same buffer with the start address advanced by (kh*Wp + kw) rows (the swizzle is
//     a function of the absolute address, so any row offset is valid), so every tap reads the one resident patch.
//     (A first version used the no-swizzle layout, whose start address is unconstrained; measured with the in-kernel
//     phase clocks below it fetched the A operand at ~16 B/clk: 270 clk per UMMA whatever N.)  Loaded once per slab by
//     TILED TMA: the input is a dense NHWC bf16 tensor described by a 4-D tensor map {C, W, H, N}; one
//     cp.async.bulk.tensor.4d with box {64, Wp, 1, 1} at (slab * 64, -p, hp - p, n) delivers one whole zero-padded image
//     row (out-of-bounds w / h / n / c read as zeros = the padding, the batch tail and the channel tail) already in the
//     SWIZZLE_128B layout -- the TMA unit swizzles by absolute shared-memory address exactly as the UMMA descriptor reads
//     it (tools/microbench/tma_probe.cu), so a row box may land at any 128-byte row of the buffer.  One thread issues the
//     ~15 row boxes of a patch; the former 256 cp.async producer threads only run the epilogue now.
//   * weights: pre-arranged and pre-swizzled per (slab, tap) as [cout][64 ch] rows of 128 B by wt_bf16_v2_kernel,
//     16 KB stages of 128 / bn entries -> one 1-D bulk TMA each (cp.async.bulk + mbarrier), P2_WS-deep ring, own warp.
//   * one elected thread issues, per (slab, tap, 16-channel step), two UMMAs 128 x bn x 16 (the CTA's two 128-row
//     tiles) into two TMEM accumulators; tcgen05.commit frees the weight stage / the patch buffer.
//   * epilogue as in conv_tc.cu (tcgen05.ld, bias, ReLU, fp32 store, optional bf16 shadow); rows that are padding
//     positions or beyond the batch are skipped.
// Rows on padding positions cost (Hp*Wp)/(H*W) - 1 extra tensor work (16 % at 25x20, k = 3) in exchange for a k*k-fold
// cut of the staging traffic.  Grouped over candidates like every kernel here (cnn.cuh).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int P2_MT = 2;                    // 128-row tiles per CTA
constexpr int P2_ROWS = P2_MT * 128;
constexpr int P2_SLAB = 64;                 // channels per patch slab (one 128-byte swizzled row per position)
constexpr int P2_WS = 2;                    // weight stages of 16 KB = 128 / bn (slab, tap) entries each; with two CTAs per SM
                                            // 64 KB of weights are in flight per SM
constexpr int P2_PRODUCERS = 256;
constexpr int P2_THREADS = P2_PRODUCERS + 64;     // + MMA warp + weight-TMA warp
constexpr uint32_t P2_TMEM_COLS = 256;
constexpr int P2_W_STAGE = 128 * 128;       // bytes of a weight stage: 128 rows x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t* v, uint32_t taddr) {      // 16 consecutive accumulator columns of this lane's row
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
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
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one zero-padded image row of 64 channels: box {64, Wp, 1, 1} of the 4-D map at (c, w, h, n)
__device__ __forceinline__ void tma_load_row(void* dst, const void* tmap, int c, int w, int h, int n, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(c), "r"(w), "r"(h), "r"(n), "r"(smem_u32(bar))
        : "memory");
}
// The tensor maps live in GLOBAL memory and are rewritten by the host between launches (new buffers per wave): the TMA unit
// caches descriptors by address, so the issuing thread acquires the current contents before its first use -- without this
// a launch could run with the previous wave's (freed) addresses.
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Operand descriptors (K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart = SBO, LBO field 1, version 1) are
// assembled in the issue loop from a constant high word and the start-address field.  The start address may be ANY
// 128-byte row of a buffer that holds the absolute-address swizzle (chunk ^ (row & 7)): measured here, the hardware XORs
// address bits [4:6] with bits [7:9] of the absolute shared-memory address, so a tap shift needs no base-offset field
// (setting base_offset = (start >> 7) & 7 gave wrong results; 0 is exact for every shift).
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

// prof (CMOOP_TC2_PROF=1): per-launch sums of clock64 spans -- [0] CTAs, [1] prologue, [2] MMA thread: start -> first
// operands ready, [3] MMA issue loop, [4] producers: start -> accumulator complete, [5] epilogue, [6] whole CTA
// OUT_BF16: every task of the launch stores bf16 (the forward of precision bf16) / fp32 (data gradient) -- one epilogue per
// instantiation keeps the kernel small
template <bool OUT_BF16>
__global__ void __launch_bounds__(P2_THREADS, 2) conv_tc2_kernel(const TcConvTask* __restrict__ tasks, int n_tasks, int n_b,
                                                                 int step, int q_max, int pb, unsigned long long* prof,
                                                                 const int* __restrict__ block_task) {
    const long long t_start = clock64();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ TcConvTask T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < (int)(sizeof(TcConvTask) / 4)) {
        // block -> task: one load from the direct table (a binary search over the task list is ~log2(n) dependent global
        // loads, several thousand cycles of a CTA that lives ~30 k); the record is then copied by 4-byte lanes
        int lo = 0;
        if (block_task) {
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
    const int p = T.pad, Hp = T.H + 2 * p, Wp = T.W + 2 * p, HpWp = Hp * Wp;
    const int Mq = n_b * HpWp;                         // < 2^31 for every supported shape (n_b <= 64)
    const int q0 = tm * P2_ROWS;
    if (q0 >= Mq) return;
    const int S = p * Wp + p;
    // the patch holds whole padded image rows: global padded rows [R0, R1] cover positions [q0 - S, q0 + P2_ROWS + S)
    const int Hp_rows = Hp;
    const int R0 = floor_div(q0 - S, Wp), R1 = (q0 + P2_ROWS - 1 + S) / Wp;
    const int n_rows = R1 - R0 + 1;                    // n_rows * Wp <= q_max positions
    const int row0_off = (q0 - S) - R0 * Wp;           // patch row of position q0 - S (tap (0, 0) of GEMM row 0)
    const int bn = T.bn, n0 = tn * bn;
    const int taps = T.k * T.k, n_slab = (T.Cin + P2_SLAB - 1) / P2_SLAB;
    const uint32_t entry_bytes = 128u * (uint32_t)bn;  // one (slab, tap) weight block

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const size_t patch_stride = ((size_t)q_max * 128 + 1023) & ~size_t(1023);
    uint8_t* patch = base;                                              // [pb][Q rows][128 B], swizzled
    uint8_t* wsm = patch + (size_t)pb * patch_stride;                   // [P2_WS][P2_W_STAGE]
    float* bias_s = reinterpret_cast<float*>(wsm + P2_WS * P2_W_STAGE);       // [128] this N tile's bias (0 without bias)
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);
    uint64_t* pfull = bars;                     // [2]  one arrive.expect_tx + the row boxes' bytes
    uint64_t* pempty = bars + 2;                // [2]  tcgen05.commit
    uint64_t* wfull = bars + 4;                 // [P2_WS]  TMA transaction
    uint64_t* wempty = wfull + P2_WS;           // [P2_WS]  tcgen05.commit
    uint64_t* accum_bar = wempty + P2_WS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&pfull[s], 1);
            mbar_init(&pempty[s], 1);
        }
        for (int s = 0; s < P2_WS; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&pfull[0], (uint32_t)n_rows * (uint32_t)Wp * 128u);
    }
    if (warp == 0) {
        // slab 0 of the patch is requested right here, by the warp whose lane 0 initialised the barriers: the ~2 k clk of TMA
        // latency run under the TMEM allocation, the bias load and the block barrier below instead of after them; one row
        // box per lane (a single thread spends ~120 clk per box: 1.8 k clk for the ~15 rows of a patch)
        __syncwarp();
        const uint32_t row_bytes = (uint32_t)Wp * 128u;
        tmap_acquire(T.tmap);
        for (int r = lane; r < n_rows; r += 32) {
            const int R = R0 + r, n = floor_div(R, Hp_rows), hp = R - n * Hp_rows;
            tma_load_row(patch + (size_t)r * row_bytes, T.tmap, 0, -p, hp - p, n, &pfull[0]);
        }
    }
    if (warp == P2_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(P2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < bn) bias_s[tid] = T.bias ? __ldg(T.bias + n0 + tid) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const long long t_pro = clock64();
    if (prof && tid == 0) {
        atomicAdd(prof + 0, 1ull);
        atomicAdd(prof + 1, (unsigned long long)(t_pro - t_start));
    }

    if (warp < P2_PRODUCERS / 32) {
        // ================= patch loads: one 64-channel slab per buffer, one row box per padded image row =================
        if (tid == 0) {
            const uint32_t row_bytes = (uint32_t)Wp * 128u;
            for (int sl = 1; sl < n_slab; ++sl) {          // slab 0 was requested in the prologue
                const int b = sl % pb;
                mbar_wait(&pempty[b], (((uint32_t)(sl / pb)) & 1u) ^ 1u);
                uint8_t* dst = patch + (size_t)b * patch_stride;
                mbar_expect_tx(&pfull[b], (uint32_t)n_rows * row_bytes);
                int n = floor_div(R0, Hp_rows), hp = R0 - n * Hp_rows;
                for (int r = 0; r < n_rows; ++r) {
                    tma_load_row(dst + (size_t)r * row_bytes, T.tmap, sl * P2_SLAB, -p, hp - p, n, &pfull[b]);
                    if (++hp == Hp_rows) { hp = 0; ++n; }
                }
            }
        }
        __syncwarp();
        // ================= epilogue =================
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t_acc = clock64();
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
            if constexpr (OUT_BF16) {
                // bf16-only output (forward): 32 accumulator columns per pass (two tcgen05.ld in flight, one wait); a lane owns
                // one GEMM row, so 8 consecutive channels leave as ONE 16-byte store (round 2: 2 539 -> 2 255 us on the 25x20 block)
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    uint32_t v[32];
                    const int ncol = min(32, c_end - c0);
                    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(mt * 128 + c0);
                    tmem_ld16(v, taddr);
                    if (ncol > 16) tmem_ld16(v + 16, taddr + 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (obase >= 0) {
#pragma unroll
                        for (int qd = 0; qd < 32; qd += 8) {
                            if (qd >= ncol) break;
                            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0 + qd);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + qd + 4);
                            float o[8];
                            o[0] = __uint_as_float(v[qd + 0]) + b0.x; o[1] = __uint_as_float(v[qd + 1]) + b0.y;
                            o[2] = __uint_as_float(v[qd + 2]) + b0.z; o[3] = __uint_as_float(v[qd + 3]) + b0.w;
                            o[4] = __uint_as_float(v[qd + 4]) + b1.x; o[5] = __uint_as_float(v[qd + 5]) + b1.y;
                            o[6] = __uint_as_float(v[qd + 6]) + b1.z; o[7] = __uint_as_float(v[qd + 7]) + b1.w;
                            if (T.relu) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
                            }
                            if (T.y) {                       // fp32 copy next to the bf16 one (not used by the engine today)
                                float4* d4 = reinterpret_cast<float4*>(T.y + obase + n0 + c0 + qd);
                                d4[0] = make_float4(o[0], o[1], o[2], o[3]);
                                d4[1] = make_float4(o[4], o[5], o[6], o[7]);
                            }
                            *reinterpret_cast<uint4*>(T.yh + obase + n0 + c0 + qd) =
                                make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
                        }
                    }
                }
            } else {
                // fp32 output (data gradient, precision-fp32 callers): 16-column passes, the stores of one pass overlap the load
                // of the next (32-column passes measured 4 % slower here: the store count does not shrink)
                for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                    uint32_t v[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(mt * 128 + c0);
                    tmem_ld16(v, taddr);
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
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (prof && tid == 0) {
            atomicAdd(prof + 4, (unsigned long long)(t_acc - t_pro));
            atomicAdd(prof + 5, (unsigned long long)(clock64() - t_acc));
        }
    } else if (warp == P2_PRODUCERS / 32) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(bn);
            const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            // Lean issue loop (round 2): the per-tap bookkeeping runs on ONE thread, so every dependent integer instruction is
            // ~5 clk of the tile's critical path.  No runtime division / modulo (tps is a power of two, the ring has two
            // stages), running byte addresses instead of kh * Wp + kw products, task fields read once from shared memory.
            // Measured with CMOOP_TC2_PROF=1 before: ~300 clk per tap of a 16-filter tile for 2 UMMAs of 16 clk each.
            const uint32_t tps = 128u / (uint32_t)bn;  // (slab, tap) entries per 16 KB weight stage: 8 / 4 / 2 / 1
            const int kk = T.k, cin = T.Cin;
            const uint32_t wsm_s = smem_u32(wsm);
            const uint32_t row_skip = (uint32_t)(Wp - kk) * 128u;     // from the end of one kernel row to the start of the next
            uint32_t st = 0, sub = 0;                  // stage counter; entry within the stage
            uint32_t b_addr = wsm_s;                   // byte address of the current weight entry
            uint32_t acc = 0;                          // 0 for the very first UMMA of the tile
            long long t_first = 0, t_wwait = 0, t_pwait = 0;
            for (int sl = 0; sl < n_slab; ++sl) {
                const int b = pb == 2 ? (sl & 1) : 0;
                const uint32_t pph = pb == 2 ? ((uint32_t)sl >> 1) & 1u : (uint32_t)sl & 1u;
                const long long tp0 = prof ? clock64() : 0;
                mbar_wait(&pfull[b], pph);
                if (prof && sl > 0) t_pwait += clock64() - tp0;
                if (sl == 0) {
                    mbar_wait(&wfull[0], 0);
                    t_first = clock64();
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // descriptors: the high word is constant (SBO 1024 B, version 1, SWIZZLE_128B); the issue loop only adds to
                // the 14-bit start-address field of the low word (a single issuing thread that rebuilds 64-bit descriptors
                // per UMMA is slower than the tensor pipe: tools/microbench/umma_layouts.cu)
                uint32_t a_tap = smem_u32(patch + (size_t)b * patch_stride) + (uint32_t)row0_off * 128u;   // tap (0, 0)
                const int ksteps = min(cin - sl * P2_SLAB, P2_SLAB) >> 4;
                const bool last_slab = sl == n_slab - 1;
                int kw = 0;
                for (int t = 0; t < taps; ++t) {
                    const uint32_t ws = st & 1u;
                    if (sub == 0) {
                        const long long tw0 = prof ? clock64() : 0;
                        mbar_wait(&wfull[ws], (st >> 1) & 1u);
                        if (prof) t_wwait += clock64() - tw0;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t a_lo = ((a_tap >> 4) & 0x3FFFu) | (1u << 16);
                    const uint32_t b_lo = ((b_addr >> 4) & 0x3FFFu) | (1u << 16);
                    for (int k4 = 0; k4 < ksteps; ++k4) {
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + (uint32_t)k4 * 2u);
#pragma unroll
                        for (int mt = 0; mt < P2_MT; ++mt) {
                            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)mt * 1024u + (uint32_t)k4 * 2u);
                            umma_bf16(tmem_base + (uint32_t)(mt * 128), ad, bd, idesc, acc);
                        }
                        acc = 1u;
                    }
                    ++sub;
                    b_addr += entry_bytes;
                    if (sub == tps || (last_slab && t == taps - 1)) {      // last entry of the stage (or of the tile)
                        umma_commit(&wempty[ws]);
                        ++st;
                        sub = 0;
                        b_addr = wsm_s + ((st & 1u) ? (uint32_t)P2_W_STAGE : 0u);
                    }
                    a_tap += 128u;
                    if (++kw == kk) { kw = 0; a_tap += row_skip; }
                }
                umma_commit(&pempty[b]);
            }
            umma_commit(accum_bar);
            if (prof) {
                atomicAdd(prof + 2, (unsigned long long)(t_first - t_pro));
                atomicAdd(prof + 3, (unsigned long long)(clock64() - t_first));
                atomicAdd(prof + 7, (unsigned long long)(t_wwait + t_pwait));
            }
        }
        __syncwarp();
    } else {
        // ================= weight stages: one bulk copy per (sub-slab, tap) =================
        if (lane == 0) {
            const uint32_t total_bytes = (uint32_t)(n_slab * taps) * entry_bytes;
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
    if (prof && tid == 0) atomicAdd(prof + 6, (unsigned long long)(clock64() - t_start));
}

// bf16 weight blocks for conv_tc2_kernel: out[tn][slab][tap][n (bn rows)][64 ch], every row 128 B, 16-byte chunk c of
// row n stored at chunk position c ^ (n & 7) (SWIZZLE_128B as the UMMA descriptor reads it; channels past Cin are zero).
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
    const int taps = T.k * T.k, n_slab = (gi + P2_SLAB - 1) / P2_SLAB;
    const long long chunks = (long long)taps * n_slab * go * 8;         // one thread = one 16-byte chunk
    const long long c = (long long)(blockIdx.x - T.block_begin) * 256 + threadIdx.x;
    if (c >= chunks) return;
    const int bn = go < 128 ? go : 128;
    // c -> (tn, slab, tap, chunk, n): consecutive threads take consecutive output channels (coalesced forward reads)
    long long r = c;
    const int n = (int)(r % bn); r /= bn;
    const int ch = (int)(r % 8); r /= 8;
    const int tap = (int)(r % taps); r /= taps;
    const int sl = (int)(r % n_slab); r /= n_slab;
    const int tn = (int)r;
    const int ch_in0 = sl * P2_SLAB + ch * 8, ch_out = tn * bn + n;
    const int kh = tap / T.k, kw = tap - kh * T.k;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ch_in0 < gi) {
        if (T.mode == 2) {
            const float* src = T.w + ((long long)(kh * T.k + kw) * T.Cin + ch_in0) * T.Cout + ch_out;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (long long)j * T.Cout);
        } else {
            const float* src = T.w + ((long long)((T.k - 1 - kh) * T.k + (T.k - 1 - kw)) * T.Cin + ch_out) * T.Cout + ch_in0;
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(src)), a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        }
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    const long long entry = ((long long)tn * n_slab + sl) * taps + tap;      // block of bn rows x 128 B
    reinterpret_cast<uint4*>(T.out)[(entry * bn + n) * 8 + (ch ^ (n & 7))] = o;
}

size_t p2_smem_bytes(int q_max, int pb) {
    const size_t patch_stride = ((size_t)q_max * 128 + 1023) & ~size_t(1023);
    return 1024 /*align*/ + (size_t)pb * patch_stride + (size_t)P2_WS * P2_W_STAGE + 512 + 512;
}
// two patch buffers (the next slab loads while the current one is multiplied) only when two CTAs still fit one SM
int p2_buffers(int cin, int q_max) { return cin > P2_SLAB && p2_smem_bytes(q_max, 2) <= 113 * 1024 ? 2 : 1; }

}  // namespace

// positions a patch buffer must hold: whole padded rows covering [q0 - S, q0 + P2_ROWS + S) for any q0
int Launch::tc2_q(int W, int k) {
    const int p = (k - 1) / 2, Wp = W + 2 * p, S = p * Wp + p;
    return ((P2_ROWS + 2 * S + Wp - 2) / Wp + 1) * Wp;
}
int Launch::tc2_rows() { return P2_ROWS; }
long long Launch::tc2_weight_elems(int Cin, int Cout, int k) {
    return (long long)k * k * ((Cin + P2_SLAB - 1) / P2_SLAB) * P2_SLAB * Cout;
}
bool Launch::tc2_ok(int H, int W, int Cin, int Cout, int k, int stride) {
    (void)H;
    if (stride != 1 || (k & 1) == 0 || Cin % 16 != 0 || Cout % 16 != 0) return false;
    if (Cout > 128 && Cout % 128 != 0) return false;
    return p2_smem_bytes(tc2_q(W, k), 1) <= 227 * 1024;
}

int Launch::conv_tc2(const TcConvTask* tasks, int n, int tiles, int n_b, int step, int q_max, int max_cin, void* st,
                     const int* block_task, bool out_bf16) {
    if (n == 0 || tiles == 0) return 0;
    const int pb = p2_buffers(max_cin, q_max);
    const size_t smem = p2_smem_bytes(q_max, pb);
    static size_t configured = 0;
    if (smem > configured) {
        for (int v = 0; v < 2; ++v) {
            const void* fn = v ? (const void*)conv_tc2_kernel<true> : (const void*)conv_tc2_kernel<false>;
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            // all of the SM's unified L1/shared storage as shared memory: two ~100 KB CTAs must be co-resident
            e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return (int)e;
        }
        configured = smem;
    }
    // optional phase profile (development aid): CMOOP_TC2_PROF=1 prints per-launch clock sums at process exit
    static unsigned long long* prof = nullptr;
    static bool prof_checked = false;
    if (!prof_checked) {
        prof_checked = true;
        if (getenv("CMOOP_TC2_PROF")) {
            cudaMalloc((void**)&prof, 64 * 8 * sizeof(unsigned long long));
            cudaMemset(prof, 0, 64 * 8 * sizeof(unsigned long long));
            atexit([] {
                unsigned long long h[64 * 8];
                cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
                for (int i = 0; i < 64; ++i) {
                    const unsigned long long* r = h + i * 8;
                    if (!r[0]) continue;
                    fprintf(stderr, "tc2 prof slot %2d: ctas %6llu  per-CTA clk: prologue %6llu  first-operands %6llu  mma-loop %6llu  "
                            "to-accum %6llu  epilogue %6llu  total %6llu  loop-waits %6llu\n", i, r[0], r[1] / r[0], r[2] / r[0], r[3] / r[0],
                            r[4] / r[0], r[5] / r[0], r[6] / r[0], r[7] / r[0]);
                }
            });
        }
    }
    static int launch_no = 0;
    unsigned long long* slot = prof ? prof + (size_t)(launch_no++ % 64) * 8 : nullptr;
    if (out_bf16)
        conv_tc2_kernel<true><<<tiles, P2_THREADS, smem, (cudaStream_t)st>>>(tasks, n, n_b, step, q_max, pb, slot, block_task);
    else
        conv_tc2_kernel<false><<<tiles, P2_THREADS, smem, (cudaStream_t)st>>>(tasks, n, n_b, step, q_max, pb, slot, block_task);
    return (int)cudaGetLastError();
}

int Launch::wt_bf16_v2(const WtBf16Task* tasks, int n, int blocks, void* st) {
    if (n == 0 || blocks == 0) return 0;
    wt_bf16_v2_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(tasks, n);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
