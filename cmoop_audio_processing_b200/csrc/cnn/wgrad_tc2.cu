// tcgen05 weight gradient on the patch layout of conv_tc2.cu (stride-1 convolutions), operands by tiled TMA.
//
//   dW[tap][ci][co] = sum_q X[q + shift(tap)][ci] * dY[q][co]         q = padded-linear output position
//
// wgrad_tc_kernel (conv_tc.cu) stages an im2col tile per 64 pixels: every activation travels k*k times through
// 16-byte cp.async gathers (and dY once per 128-row K tile), which makes it L2-bandwidth-bound.  Here the reduction runs
// over positions, so both operands are MN-major and the activations of ALL taps come from one resident patch:
//   * X patch: rows = positions, 128 B = 64 input channels, SWIZZLE_128B -- byte-for-byte the patch of conv_tc2.cu.
//     As an MN-major A operand its 64-element group is a row; the UMMA's M = 128 is TWO groups LBO bytes apart, which
//     here are two TAPS: the same buffer at shift(a) and shift(b) rows (LBO = (shift(b) - shift(a)) * 128; the
//     swizzle depends on absolute addresses, so any row offset is valid).
//   * dY tile: rows = positions, 128 B = 64 output channels per slab (two slabs LBO apart at bn = 128).
//   * both are loaded by TILED TMA from the dense NHWC bf16 tensors: 4-D tensor maps {C, W, H, N}, one
//     cp.async.bulk.tensor.4d with box {64, Wp, 1, 1} at (c0, -p, hp - p, n) per zero-padded image row -- padding
//     positions, the batch tail (tail-batch map: N = n_b) and channels past C arrive as zeros, already swizzled.  One
//     thread issues the row boxes of a chunk; a W2_ST-deep ring keeps the next chunk in flight while this one multiplies.
//   * the bias gradient sum_q dY[q][co] is the tap "after the last": a constant buffer whose rows are (1, 0, ..., 0)
//     paired with the last real tap (k*k is odd), so it costs no extra UMMA.
//   * one UMMA 128 x bn x 16 per (16 positions, tap pair) into one TMEM accumulator per pair (<= 512 columns; the
//     pairs are split into balanced groups when they do not fit), issued by ONE thread whose loop only adds to the
//     descriptors' address field (the first version rebuilt them -- integer divisions included -- per UMMA and measured
//     ~250 clk per UMMA, which was misread as a slow MN-major operand fetch: tools/microbench/umma_layouts.cu shows the
//     UMMA rate is independent of operand layout, start-row alignment and LBO).
//   * a CTA walks the chunks of its split and writes its [taps][64][bn] block of the split's partial gradient once;
//     splits are summed by reduce_kernel (deterministic).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int W2_QC = 128;                   // positions per chunk (8 UMMA K steps)
constexpr int W2_ST = 2;                     // chunk buffers (X patch + dY tile) in the ring
constexpr int W2_EPI = 256;                  // 8 epilogue warps
constexpr int W2_THREADS = W2_EPI + 64;      // + MMA warp + TMA warp
constexpr uint32_t W2_TMEM_COLS = 512;
constexpr int W2_MAX_PAIRS = 13;             // (25 + 1) / 2

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();      // a protocol error traps instead of hanging the GPU
    }
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
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
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

// Work item of a CTA: (task, split, input-channel slab, output-channel tile, tap-pair group).
// qx_max / qy_max: positions (whole padded rows) a patch / dY buffer must hold for the widest task of the launch.
__global__ void __launch_bounds__(W2_THREADS, 1) wgrad_tc2_kernel(const TcWgradTask* __restrict__ tasks, int n_tasks, int n_b,
                                                                  int qx_max, int qy_max, const int* __restrict__ block_task) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
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
    // ---- decode: local = ((split * n_slab + slab) * tiles_n + tn) * n_grp + grp
    const int bn = T.bn, taps = T.k * T.k, n_pairs = (taps + 1) >> 1;
    const int gp_max = 512 / bn < W2_MAX_PAIRS ? 512 / bn : W2_MAX_PAIRS;      // accumulators that fit TMEM
    const int n_grp = (n_pairs + gp_max - 1) / gp_max, gp = (n_pairs + n_grp - 1) / n_grp;     // balanced groups
    const int n_slab = (T.Cin + 63) >> 6;
    int local = blockIdx.x - T.tile_begin;
    const int grp = local % n_grp; local /= n_grp;
    const int tn = local % T.tiles_n; local /= T.tiles_n;
    const int slab = local % n_slab;
    const int split = local / n_slab;
    const int pair0 = grp * gp, my_pairs = min(gp, n_pairs - pair0);
    const int n0 = tn * bn, c0 = slab * 64, cw = min(T.Cin - c0, 64);      // this CTA's output / input channels
    const int p = T.pad, Wp = T.W + 2 * p, Hp = T.H + 2 * p, HpWp = Hp * Wp, S = p * Wp + p;
    const int Mq = n_b * HpWp;
    const int cps = T.m_chunk / W2_QC;                                      // chunks per split (full batch)
    const int chunk_begin = split * cps;
    const int n_chunks = max(0, min(cps, (Mq + W2_QC - 1) / W2_QC - chunk_begin));
    const int Kext = taps * T.Cin + 1;
    float* out = T.out + (long long)split * Kext * T.Cout;

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const size_t xstride = ((size_t)qx_max * 128 + 1023) & ~size_t(1023);
    const size_t yslab = ((size_t)qy_max * 128 + 1023) & ~size_t(1023);     // one 64-channel slab of a dY tile
    const int yslabs = (bn + 63) >> 6;                                       // MN-major atoms (64 channels) per dY row
    uint8_t* xbuf = base;                                        // [W2_ST][qx_max rows][128 B]
    uint8_t* ybuf = xbuf + W2_ST * xstride;                      // [W2_ST][2][qy_max rows][128 B]
    uint8_t* ones = ybuf + W2_ST * 2 * yslab;                    // [QC + 8 rows][128 B]: element 0 of every row = 1
    uint32_t* pair_sa = reinterpret_cast<uint32_t*>(ones + (size_t)(W2_QC + 8) * 128);   // [16] row shift of a pair's first tap
    uint32_t* pair_lbo = pair_sa + 16;                           // [16] LBO field of the pair (0 = the ones pair)
    uint64_t* bars = reinterpret_cast<uint64_t*>(pair_lbo + 16);
    uint64_t* full = bars;               // [W2_ST]  one arrive.expect_tx + the row boxes' bytes
    uint64_t* empty = bars + W2_ST;      // [W2_ST]  tcgen05.commit
    uint64_t* accum_bar = bars + 2 * W2_ST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * W2_ST + 1);

    if (n_chunks == 0) {             // nothing of the (short) batch falls into this split: its partial is zero
        for (int pr = 0; pr < my_pairs; ++pr)
            for (int i = tid; i < 128 * bn; i += W2_THREADS) {
                const int L = i / bn, c = i - L * bn, tap = 2 * (pair0 + pr) + (L >> 6), ci = L & 63;
                if (tap < taps && ci < cw) out[(long long)(tap * T.Cin + c0 + ci) * T.Cout + n0 + c] = 0.f;
                if (tap == taps && ci == 0 && slab == 0) out[(long long)(Kext - 1) * T.Cout + n0 + c] = 0.f;
            }
        return;
    }
    if (tid == 0) {
        for (int s = 0; s < W2_ST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W2_EPI / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(W2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the ones buffer: row r, chunk 0 lives at chunk position (0 ^ (r & 7)); bf16(1.0) in its first element
    for (int i = tid; i < (W2_QC + 8) * 8; i += W2_THREADS) {
        const int r = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(ones + (size_t)r * 128 + (ch << 4)) = make_uint4(ch == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    }
    // per tap pair: row shift of its first tap and the LBO field to the second tap of the same patch (constant)
    if (tid < my_pairs) {
        const int ta = 2 * (pair0 + tid), tb = ta + 1;
        const uint32_t sa = (uint32_t)((ta / T.k) * Wp + (ta % T.k));
        pair_sa[tid] = sa;
        pair_lbo[tid] = tb < taps ? ((((uint32_t)((tb / T.k) * Wp + (tb % T.k)) - sa) * 128u) >> 4) << 16 : 0u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t row_bytes = (uint32_t)Wp * 128u;

    if (warp < W2_EPI / 32) {
        // ================= epilogue: accumulator lanes = (tap of the pair, ci), columns = co =================
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lane_grp = warp & 3, half = warp >> 2;
        const int cb = bn >= 32 ? half * (bn / 2) : 0;
        const int ce = bn >= 32 ? cb + bn / 2 : (half == 0 ? bn : 0);
        const int L = lane_grp * 32 + lane, ci = L & 63;
        for (int pr = 0; pr < my_pairs; ++pr) {
            const int tap = 2 * (pair0 + pr) + (L >> 6);
            long long row = -1;                                   // row of the [K+1][Cout] gradient this lane holds
            if (tap < taps) {
                if (ci < cw) row = (long long)tap * T.Cin + c0 + ci;
            } else if (ci == 0 && slab == 0) {
                row = Kext - 1;                                   // the ones "tap": bias gradient
            }
            for (int cc = cb; cc < ce; cc += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(pr * bn + cc);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row >= 0) {
                    float* dst = out + row * T.Cout + n0 + cc;
#pragma unroll
                    for (int qd = 0; qd < 16; qd += 4)
                        *reinterpret_cast<float4*>(dst + qd) = make_float4(__uint_as_float(v[qd]), __uint_as_float(v[qd + 1]),
                                                                           __uint_as_float(v[qd + 2]), __uint_as_float(v[qd + 3]));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == W2_EPI / 32) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t idesc = 0;
            idesc |= 1u << 4;                       // D = F32
            idesc |= 1u << 7;                       // A = BF16
            idesc |= 1u << 10;                      // B = BF16
            idesc |= (1u << 15) | (1u << 16);       // A and B are MN-major
            idesc |= (uint32_t)(bn >> 3) << 17;     // N
            idesc |= (uint32_t)(128 >> 4) << 24;    // M
            // descriptor high word: SBO = 1024 B (8-row groups along K), version 1, SWIZZLE_128B
            const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            const uint32_t y_lbo = (uint32_t)((yslab >> 4) & 0x3FFF) << 16;               // second 64-channel slab of dY
            const uint32_t ones_addr = smem_u32(ones);
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c % W2_ST;
                const int q0 = (chunk_begin + c) * W2_QC;
                // position q0 - S sits dx rows into the X buffer, position q0 dy rows into the dY buffer (whole padded
                // image rows are loaded); the ones buffer is entered at the row with the same (row & 7) as the A start,
                // so that LBO stays a multiple of the 8-row swizzle period?  No such need: the swizzle is by absolute
                // address on both sides -- any row offset is exact.
                const uint32_t dx = (uint32_t)((q0 - S) - floor_div(q0 - S, Wp) * Wp);
                const uint32_t dy = (uint32_t)(q0 - (q0 / Wp) * Wp);
                mbar_wait(&full[buf], ((uint32_t)(c / W2_ST)) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t xa = smem_u32(xbuf + (size_t)buf * xstride) + dx * 128u;
                const uint32_t y_lo = (((smem_u32(ybuf + (size_t)buf * 2 * yslab) + dy * 128u) >> 4) & 0x3FFFu) | y_lbo;
#pragma unroll 1
                for (int pr = 0; pr < my_pairs; ++pr) {
                    const uint32_t a0 = xa + pair_sa[pr] * 128u;
                    uint32_t lbo = pair_lbo[pr];
                    if (lbo == 0u) lbo = (((ones_addr - a0) >> 4) & 0x3FFFu) << 16;       // second group: the ones buffer
                    const uint32_t a_lo = ((a0 >> 4) & 0x3FFFu) | lbo;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(pr * bn);
#pragma unroll
                    for (int ks = 0; ks < W2_QC / 16; ++ks) {      // 16 positions = 16 rows of 128 B further down
                        const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)ks * 128u);
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(y_lo + (uint32_t)ks * 128u);
                        umma_bf16(d_tmem, ad, bd, idesc, (c | ks) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty[buf]);
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        // ================= TMA issuer: the padded image rows of the X patch and of the dY tile =================
        if (lane == 0) {
            const char* maps = reinterpret_cast<const char*>(T.tmaps) + (n_b == T.n_full ? 0 : 2 * kTmapBytes);
            const void* tm_x = maps;
            const void* tm_y = maps + kTmapBytes;
            tmap_acquire(tm_x);
            tmap_acquire(tm_y);
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c % W2_ST;
                const int q0 = (chunk_begin + c) * W2_QC;
                const int r0x = floor_div(q0 - S, Wp), r1x = (q0 + W2_QC - 1 + S) / Wp;
                const int r0y = q0 / Wp, r1y = (q0 + W2_QC - 1) / Wp;
                const int nx = r1x - r0x + 1, ny = r1y - r0y + 1;
                mbar_wait(&empty[buf], (((uint32_t)(c / W2_ST)) & 1u) ^ 1u);
                mbar_expect_tx(&full[buf], (uint32_t)(nx + ny * yslabs) * row_bytes);
                uint8_t* xb = xbuf + (size_t)buf * xstride;
                uint8_t* yb = ybuf + (size_t)buf * 2 * yslab;
                {
                    int n = floor_div(r0x, Hp), hp = r0x - n * Hp;
                    for (int r = 0; r < nx; ++r) {
                        tma_load_row(xb + (size_t)r * row_bytes, tm_x, c0, -p, hp - p, n, &full[buf]);
                        if (++hp == Hp) { hp = 0; ++n; }
                    }
                }
                for (int j = 0; j < yslabs; ++j) {
                    int n = r0y / Hp, hp = r0y - n * Hp;
                    for (int r = 0; r < ny; ++r) {
                        tma_load_row(yb + (size_t)j * yslab + (size_t)r * row_bytes, tm_y, n0 + 64 * j, -p, hp - p, n, &full[buf]);
                        if (++hp == Hp) { hp = 0; ++n; }
                    }
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == W2_EPI / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(W2_TMEM_COLS) : "memory");
    }
}

int w2_xrows(int W, int k) {
    const int p = (k - 1) / 2, Wp = W + 2 * p, S = p * Wp + p;
    return ((W2_QC + 2 * S + Wp - 2) / Wp + 1) * Wp;
}
int w2_yrows(int W, int k) {
    const int p = (k - 1) / 2, Wp = W + 2 * p;
    return ((W2_QC + Wp - 2) / Wp + 1) * Wp;
}
size_t w2_smem_bytes(int qx, int qy) {
    const size_t xstride = ((size_t)qx * 128 + 1023) & ~size_t(1023), yslab = ((size_t)qy * 128 + 1023) & ~size_t(1023);
    return 1024 + W2_ST * (xstride + 2 * yslab) + (size_t)(W2_QC + 8) * 128 + 512;
}

}  // namespace

int Launch::wg2_bn(int Cout) { return Cout < 128 ? Cout : 128; }
// positions (whole padded rows) of the X patch in the low 16 bits, of the dY tile in the high 16 bits
int Launch::wg2_q(int W, int k) { return w2_xrows(W, k) | (w2_yrows(W, k) << 16); }
// CTAs per split: slabs x output tiles x tap-pair groups
int Launch::wg2_items(int Cin, int Cout, int k) {
    const int bn = wg2_bn(Cout), pairs = (k * k + 1) / 2, gp = 512 / bn < W2_MAX_PAIRS ? 512 / bn : W2_MAX_PAIRS;
    return ((Cin + 63) / 64) * (Cout / bn) * ((pairs + gp - 1) / gp);       // the kernel balances the pairs over these groups
}
// split geometry over the padded-linear positions of a full batch: chunk rows per split (multiple of W2_QC).  A split
// costs one [K+1][Cout] fp32 partial (written, then re-read by reduce_kernel), so wide layers -- whose (slab, tile, pair
// group) items already give the task ~32 CTAs -- are split less: about 32 CTAs per task, at least 4 chunks per CTA.
void Launch::wg2_splits(long long Mq, int items, int* splits, int* m_chunk) {
    const long long chunks = (Mq + W2_QC - 1) / W2_QC;
    long long s = std::min<long long>(std::max(1, 32 / std::max(1, items)), chunks / 4);
    s = s < 1 ? 1 : (s > 16 ? 16 : s);
    const long long cps = (chunks + s - 1) / s;
    *splits = (int)((chunks + cps - 1) / cps);
    *m_chunk = (int)(cps * W2_QC);
}
bool Launch::wg2_ok(int H, int W, int Cin, int Cout, int k, int stride) {
    (void)H;
    if (stride != 1 || (k != 3 && k != 5) || Cin % 16 != 0 || Cout % 16 != 0) return false;
    if (Cout > 128 && Cout % 128 != 0) return false;
    if (W + k - 1 > 256) return false;               // a padded row is one TMA box
    return w2_smem_bytes(w2_xrows(W, k), w2_yrows(W, k)) <= 220 * 1024;
}

// q_max: the launch's largest wg2_q() components (engine: max of the low and of the high halves)
int Launch::wgrad_tc2(const TcWgradTask* tasks, int n, int tiles, int n_b, int q_max, void* st, const int* block_task) {
    if (n == 0 || tiles == 0) return 0;
    const int qx = q_max & 0xffff, qy = q_max >> 16;
    const size_t smem = w2_smem_bytes(qx, qy);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    wgrad_tc2_kernel<<<tiles, W2_THREADS, smem, (cudaStream_t)st>>>(tasks, n, n_b, qx, qy, block_task);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
