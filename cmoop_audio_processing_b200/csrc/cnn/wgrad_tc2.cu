// tcgen05 weight gradient on the patch layout of conv_tc2.cu (stride-1 convolutions).  OPT-IN (CMOOP_CNN_WG2=1), kept as a
// measured negative result: correct (tests/test_gpu_conv_tc.py::test_patch_weight_gradient) but slower than wgrad_tc_kernel.
// MN-major UMMA operands are fetched at ~16 B/clk (~250 clk per 128 x 64 x 16 UMMA with 8-row-aligned starts) and a start
// row that is not a multiple of 8 -- every tap shift here -- costs another 3x (297 / 181 / 163 us vs 100 / 68 / 62 us with
// artificially aligned shifts, tools/diag_wg2.py), whereas K-major SWIZZLE_128B operands (conv_tc2.cu) take any row offset
// at full speed.  A K-major formulation needs position-contiguous (transposed) activations, where a one-position tap
// shift is a 2-byte address offset the descriptor cannot express.
//
//   dW[tap][ci][co] = sum_q X[q + shift(tap)][ci] * dY[q][co]         q = padded-linear output position
//
// wgrad_tc_kernel (conv_tc.cu) stages an im2col tile per 64 pixels: every activation travels k*k times through
// 16-byte cp.async gathers and the GEMM has 42 flop per staged byte.  Here the reduction runs over positions, so both
// operands are MN-major and the activations of ALL taps come from one resident patch:
//   * X patch: rows = positions [q0 - S, q0 + 256 + S), 128 B = 64 input channels, SWIZZLE_128B -- byte-for-byte the
//     patch of conv_tc2.cu.  As an MN-major A operand its 64-element group is a row; the UMMA's M = 128 is TWO
//     groups LBO bytes apart, which here are two TAPS: the same buffer at shift(a) and shift(b) rows
//     (LBO = (shift(b) - shift(a)) * 128; the swizzle depends on absolute addresses, so any row offset is valid).
//   * dY tile: rows = positions, 128 B = 64 output channels (zero rows at padding positions / past the batch).
//   * the bias gradient sum_q dY[q][co] is the tap "after the last": a constant buffer whose rows are (1, 0, ..., 0)
//     paired with the last real tap (k*k is odd), so it costs no extra UMMA.
//   * one UMMA 128 x bn x 16 per (16 positions, tap pair) into one TMEM accumulator per pair (<= 512 columns: all 5
//     pairs of a 3x3, two groups of 8 + 5 for a 5x5 at bn = 64); a CTA walks the chunks of its split with two
//     shared-memory buffers (loads of chunk c+1 overlap the UMMAs of chunk c) and writes its [taps][64][bn] block of
//     the split's partial gradient once; splits are summed by reduce_kernel (deterministic).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int W2_QC = 128;                   // positions per chunk (8 UMMA K steps)
constexpr int W2_PRODUCERS = 256;            // 8 producer / epilogue warps
constexpr int W2_THREADS = W2_PRODUCERS + 32;
constexpr uint32_t W2_TMEM_COLS = 512;
constexpr int W2_MAX_PAIRS = 13;             // (25 + 1) / 2
constexpr int W2_ST = 3;                     // chunk buffers (X patch + dY tile) in the ring
constexpr int W2_LOOK = 2;                   // chunks whose cp.async groups are in flight per producer thread

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();      // a protocol error traps instead of hanging the GPU
    }
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
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
// Work item of a CTA: (task, split, input-channel slab, output-channel tile, tap-pair group).
__global__ void __launch_bounds__(W2_THREADS, 1) wgrad_tc2_kernel(const TcWgradTask* __restrict__ tasks, int n_tasks, int n_b,
                                                                  int q_max) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ TcWgradTask T;
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
    const int p = T.pad, Wp = T.W + 2 * p, HpWp = (T.H + 2 * p) * Wp, S = p * Wp + p;
    const int Mq = n_b * HpWp;
    const int QX = W2_QC + 2 * S;                                           // rows of the X patch (<= q_max)
    const int cps = T.m_chunk / W2_QC;                                      // chunks per split (full batch)
    const int chunk_begin = split * cps;
    const int n_chunks = max(0, min(cps, (Mq + W2_QC - 1) / W2_QC - chunk_begin));
    const int Kext = taps * T.Cin + 1;
    float* out = T.out + (long long)split * Kext * T.Cout;

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const size_t xstride = ((size_t)q_max * 128 + 1023) & ~size_t(1023);
    uint8_t* xbuf = base;                                        // [W2_ST][QX rows][128 B]
    const int yslabs = (bn + 63) >> 6;                           // 64-channel groups of the dY tile (MN-major atoms)
    const size_t ystride = (size_t)yslabs * W2_QC * 128;
    uint8_t* ybuf = xbuf + W2_ST * xstride;                      // [W2_ST][yslabs][QC rows][128 B]
    uint8_t* ones = ybuf + W2_ST * ystride;                      // [QC rows][128 B]: element 0 of every row = 1
    uint32_t* pair_lo = reinterpret_cast<uint32_t*>(ones + (size_t)W2_QC * 128);   // [W2_ST][16] A-descriptor low words
    uint64_t* bars = reinterpret_cast<uint64_t*>(pair_lo + W2_ST * 16);
    uint64_t* full = bars;               // [W2_ST]
    uint64_t* empty = bars + W2_ST;      // [W2_ST]
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
            mbar_init(&full[s], W2_PRODUCERS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W2_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(W2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the ones buffer: row r, chunk 0 lives at chunk position (0 ^ (r & 7)); bf16(1.0) in its first element
    for (int i = tid; i < W2_QC * 8; i += W2_THREADS) {
        const int r = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(ones + (size_t)r * 128 + (ch << 4)) = make_uint4(ch == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    }
    // A-operand descriptors of every tap pair, built ONCE (the issue loop below only adds the K step to the address
    // field): low word = start address >> 4 | (LBO >> 4) << 16 at K step 0 of buffer `b`.  Group 0 of M is tap 2*pair at
    // its row shift, group 1 the next tap of the same patch (LBO = shift difference) or, after the last tap, the ones
    // buffer (bias gradient).  A single issuing thread that rebuilt them per UMMA (integer divisions by k included) was
    // the whole cost of the first version of this kernel: ~250 clk per UMMA against 62 clk for the lean loop
    // (tools/microbench/umma_layouts.cu: operand layout, start-row alignment and LBO do NOT change the UMMA rate).
    if (tid < W2_ST * my_pairs) {
        const int b = tid / my_pairs, pr = tid - b * my_pairs;
        const int ta = 2 * (pair0 + pr), tb = ta + 1;
        const uint32_t xa = smem_u32(xbuf + (size_t)b * xstride);
        const uint32_t a0 = xa + (uint32_t)((ta / T.k) * Wp + (ta % T.k)) * 128u;
        const uint32_t a1 = tb < taps ? xa + (uint32_t)((tb / T.k) * Wp + (tb % T.k)) * 128u : smem_u32(ones);
        pair_lo[b * 16 + pr] = ((a0 >> 4) & 0x3FFFu) | ((((a1 - a0) >> 4) & 0x3FFFu) << 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < W2_PRODUCERS / 32) {
        // ================= producers =================
        // Thread -> (16-byte chunk column, first row) of the X patch and of the dY tile; every further row of the thread
        // is `rstep` rows on, so its padded coordinates (n, hp, wp) advance incrementally (two integer divisions per
        // thread and chunk instead of two per row).  W2_LOOK chunks of cp.async groups stay in flight per thread.
        const int xcs = cw >> 3;                                   // 16-byte chunks per X row
        const int ycs = bn >> 3;                                   // 16-byte chunks per dY row
        const int x_chunk = tid % xcs, x_r0 = tid / xcs, x_step = W2_PRODUCERS / xcs;
        const int y_chunk = tid % ycs, y_r0 = tid / ycs, y_step = W2_PRODUCERS / ycs;
        const int x_dh = x_step / Wp, x_dw = x_step - x_dh * Wp, y_dh = y_step / Wp, y_dw = y_step - y_dh * Wp;
        const int Hp = T.H + 2 * p;
        const __nv_bfloat16* xsrc = T.xh + c0 + x_chunk * 8;
        const __nv_bfloat16* ysrc = T.dyh + n0 + y_chunk * 8;
        const uint32_t y_col = (uint32_t)(y_chunk >> 3) * (uint32_t)(W2_QC * 128), y_c7 = (uint32_t)(y_chunk & 7);
        for (int c = 0; c < n_chunks + W2_LOOK; ++c) {
            if (c < n_chunks) {
                const int buf = c % W2_ST;
                const int q0 = (chunk_begin + c) * W2_QC;
                mbar_wait(&empty[buf], (((uint32_t)(c / W2_ST)) & 1u) ^ 1u);
                uint8_t* xb = xbuf + (size_t)buf * xstride;
                uint8_t* yb = ybuf + (size_t)buf * ystride;
                if (x_r0 < x_step) {   // X patch rows q0 - S + i
                    int q = q0 - S + x_r0;
                    int n = 0, hp = 0, wp = 0;
                    if (q >= 0) { n = q / HpWp; const int rem = q - n * HpWp; hp = rem / Wp; wp = rem - hp * Wp; }
                    else { // rows before the first sample: walk up from a negative position (zero rows until q >= 0)
                        const int qq = q + HpWp; n = -1; hp = qq / Wp; wp = qq - hp * Wp;
                    }
                    for (int i = x_r0; i < QX; i += x_step, q += x_step) {
                        const bool ok = n >= 0 && q < Mq && hp >= p && hp < T.H + p && wp >= p && wp < T.W + p;
                        const long long off = ok ? ((long long)(n * T.H + hp - p) * T.W + wp - p) * T.Cin : 0;
                        cp_async16(xb + (size_t)i * 128 + ((x_chunk ^ (i & 7)) << 4), xsrc + off, ok ? 16u : 0u);
                        wp += x_dw; hp += x_dh;
                        if (wp >= Wp) { wp -= Wp; ++hp; }
                        while (hp >= Hp) { hp -= Hp; ++n; }
                    }
                }
                if (y_r0 < y_step) {   // dY rows q0 + i
                    int q = q0 + y_r0;
                    int n = q / HpWp;
                    const int rem = q - n * HpWp;
                    int hp = rem / Wp, wp = rem - hp * Wp;
                    for (int i = y_r0; i < W2_QC; i += y_step, q += y_step) {
                        const bool ok = q < Mq && hp >= p && hp < T.H + p && wp >= p && wp < T.W + p;
                        const long long off = ok ? ((long long)(n * T.H + hp - p) * T.W + wp - p) * T.Cout : 0;
                        cp_async16(yb + y_col + (size_t)i * 128 + ((y_c7 ^ (uint32_t)(i & 7)) << 4), ysrc + off, ok ? 16u : 0u);
                        wp += y_dw; hp += y_dh;
                        if (wp >= Wp) { wp -= Wp; ++hp; }
                        while (hp >= Hp) { hp -= Hp; ++n; }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (c >= W2_LOOK) {
                asm volatile("cp.async.wait_group %0;" ::"n"(W2_LOOK) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&full[(c - W2_LOOK) % W2_ST]);
            }
        }
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
    } else {
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
            const uint32_t y_lbo = (uint32_t)(((W2_QC * 128) >> 4) & 0x3FFF) << 16;        // second 64-channel group of dY
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c % W2_ST;
                mbar_wait(&full[buf], ((uint32_t)(c / W2_ST)) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t y_lo = ((smem_u32(ybuf + (size_t)buf * ystride) >> 4) & 0x3FFFu) | y_lbo;
                const uint32_t* plo = pair_lo + buf * 16;
#pragma unroll 1
                for (int pr = 0; pr < my_pairs; ++pr) {
                    const uint32_t a_lo = plo[pr];
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
    }
    __syncthreads();
    if (warp == W2_PRODUCERS / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(W2_TMEM_COLS) : "memory");
    }
}

size_t w2_smem_bytes(int q_max) {
    const size_t xstride = ((size_t)q_max * 128 + 1023) & ~size_t(1023);
    // X patches + dY tiles (two 64-channel slabs at bn = 128) + ones buffer + offsets + pair descriptors / barriers
    return 1024 + W2_ST * xstride + (W2_ST * 2 + 1) * (size_t)W2_QC * 128 + 1024;
}

}  // namespace

int Launch::wg2_bn(int Cout) { return Cout < 128 ? Cout : 128; }
int Launch::wg2_q(int W, int k) {
    const int p = (k - 1) / 2;
    return W2_QC + 2 * (p * (W + 2 * p) + p);
}
// CTAs per split: slabs x output tiles x tap-pair groups
int Launch::wg2_items(int Cin, int Cout, int k) {
    const int bn = wg2_bn(Cout), pairs = (k * k + 1) / 2, gp = 512 / bn < W2_MAX_PAIRS ? 512 / bn : W2_MAX_PAIRS;
    return ((Cin + 63) / 64) * (Cout / bn) * ((pairs + gp - 1) / gp);       // the kernel balances the pairs over these groups
}
// split geometry over the padded-linear positions of a full batch: chunk rows per split (multiple of 256)
void Launch::wg2_splits(long long Mq, int* splits, int* m_chunk) {
    const long long chunks = (Mq + W2_QC - 1) / W2_QC;
    long long s = chunks / 8;                        // >= 8 chunks per CTA amortise its prologue / epilogue
    s = s < 1 ? 1 : (s > 16 ? 16 : s);
    const long long cps = (chunks + s - 1) / s;
    *splits = (int)((chunks + cps - 1) / cps);
    *m_chunk = (int)(cps * W2_QC);
}
bool Launch::wg2_ok(int H, int W, int Cin, int Cout, int k, int stride) {
    (void)H;
    if (stride != 1 || (k != 3 && k != 5) || Cin % 16 != 0 || Cout % 16 != 0) return false;
    if (Cout > 128 && Cout % 128 != 0) return false;
    return w2_smem_bytes(wg2_q(W, k)) <= 220 * 1024;
}

int Launch::wgrad_tc2(const TcWgradTask* tasks, int n, int tiles, int n_b, int q_max, void* st) {
    if (n == 0 || tiles == 0) return 0;
    const size_t smem = w2_smem_bytes(q_max);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    wgrad_tc2_kernel<<<tiles, W2_THREADS, smem, (cudaStream_t)st>>>(tasks, n, n_b, q_max);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
