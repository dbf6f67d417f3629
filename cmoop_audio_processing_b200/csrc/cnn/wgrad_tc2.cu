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

constexpr int W2_QC = 256;                   // positions per chunk (16 UMMA K steps)
constexpr int W2_PRODUCERS = 256;            // 8 producer / epilogue warps
constexpr int W2_THREADS = W2_PRODUCERS + 32;
constexpr uint32_t W2_TMEM_COLS = 512;
constexpr int W2_MAX_PAIRS = 13;             // (25 + 1) / 2

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
// MN-major, SWIZZLE_128B: 64-element (128-byte) groups along M/N `lbo` bytes apart, 8-row groups along K 1024 B apart
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
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
__device__ __forceinline__ void named_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
    const int bn = T.bn, taps = T.k * T.k, n_pairs = (taps + 1) >> 1, gp = 512 / bn < W2_MAX_PAIRS ? 512 / bn : W2_MAX_PAIRS;
    const int n_grp = (n_pairs + gp - 1) / gp, n_slab = (T.Cin + 63) >> 6;
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
    uint8_t* xbuf = base;                                        // [2][QX rows][128 B]
    uint8_t* ybuf = xbuf + 2 * xstride;                          // [2][256 rows][128 B]
    uint8_t* ones = ybuf + 2 * (size_t)W2_QC * 128;              // [256 rows][128 B]: element 0 of every row = 1
    int* xoff = reinterpret_cast<int*>(ones + (size_t)W2_QC * 128);      // [q_max] element offset into xh, -1 = zero row
    int* yoff = xoff + ((q_max + 3) & ~3);                               // [256]   element offset into dyh, -1 = zero row
    uint64_t* bars = reinterpret_cast<uint64_t*>(yoff + W2_QC);
    uint64_t* full = bars;           // [2]
    uint64_t* empty = bars + 2;      // [2]
    uint64_t* accum_bar = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

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
        for (int s = 0; s < 2; ++s) {
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < W2_PRODUCERS / 32) {
        // ================= producers =================
        const int xcs = cw >> 3;                                   // 16-byte chunks per X row
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1;
            const int q0 = (chunk_begin + c) * W2_QC;
            mbar_wait(&empty[buf], (((uint32_t)(c >> 1)) & 1u) ^ 1u);
            named_bar(1, W2_PRODUCERS);                            // everyone is done with the previous offsets
            for (int i = tid; i < QX + W2_QC; i += W2_PRODUCERS) {
                const bool is_x = i < QX;
                const int q = is_x ? q0 - S + i : q0 + (i - QX);
                int off = -1;
                if (q >= 0 && q < Mq) {
                    const int n = q / HpWp, rem = q - n * HpWp;
                    const int hp = rem / Wp, wp = rem - hp * Wp;
                    if (hp >= p && hp < T.H + p && wp >= p && wp < T.W + p)
                        off = ((n * T.H + hp - p) * T.W + wp - p) * (is_x ? T.Cin : T.Cout);
                }
                if (is_x) xoff[i] = off; else yoff[i - QX] = off;
            }
            named_bar(1, W2_PRODUCERS);
            uint8_t* xb = xbuf + (size_t)buf * xstride;
            uint8_t* yb = ybuf + (size_t)buf * W2_QC * 128;
            {   // X patch: cw / 8 chunks per row
                const int chunk = tid % xcs, r0 = tid / xcs, rstep = W2_PRODUCERS / xcs;
                if (r0 < rstep)
                    for (int i = r0; i < QX; i += rstep) {
                        const int off = xoff[i];
                        cp_async16(xb + (size_t)i * 128 + ((chunk ^ (i & 7)) << 4), off >= 0 ? T.xh + off + c0 + chunk * 8 : T.xh,
                                   off >= 0 ? 16u : 0u);
                    }
            }
            {   // dY tile: bn / 8 chunks per row
                const int ycs = bn >> 3;
                const int chunk = tid % ycs, r0 = tid / ycs, rstep = W2_PRODUCERS / ycs;
                if (r0 < rstep)
                    for (int i = r0; i < W2_QC; i += rstep) {
                        const int off = yoff[i];
                        cp_async16(yb + (size_t)i * 128 + ((chunk ^ (i & 7)) << 4), off >= 0 ? T.dyh + off + n0 + chunk * 8 : T.dyh,
                                   off >= 0 ? 16u : 0u);
                    }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full[buf]);
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
            const uint32_t ones_addr = smem_u32(ones);
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c & 1;
                mbar_wait(&full[buf], ((uint32_t)(c >> 1)) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t xa = smem_u32(xbuf + (size_t)buf * xstride), ya = smem_u32(ybuf + (size_t)buf * W2_QC * 128);
                for (int ks = 0; ks < W2_QC / 16; ++ks) {
                    const uint64_t bd = make_desc_mn_sw128(ya + (uint32_t)ks * 2048u, 8192);
                    for (int pr = 0; pr < my_pairs; ++pr) {
                        const int ta = 2 * (pair0 + pr), tb = ta + 1;
                        const uint32_t sa = (uint32_t)((ta / T.k) * Wp + (ta % T.k));
                        const uint32_t a0 = xa + (sa + (uint32_t)ks * 16u) * 128u;
                        // second 64-row group of M: the next tap of the same patch, or the ones buffer after the last tap
                        const uint32_t a1 = tb < taps ? xa + ((uint32_t)((tb / T.k) * Wp + (tb % T.k)) + (uint32_t)ks * 16u) * 128u
                                                      : ones_addr + (uint32_t)ks * 2048u;
                        const uint64_t ad = make_desc_mn_sw128(a0, a1 - a0);
                        umma_bf16(tmem_base + (uint32_t)(pr * bn), ad, bd, idesc, (c | ks) != 0 ? 1u : 0u);
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
    return 1024 + 2 * xstride + 3 * (size_t)W2_QC * 128 + (size_t)(((q_max + 3) & ~3) + W2_QC) * 4 + 256;
}

}  // namespace

int Launch::wg2_bn(int Cout) { return Cout < 64 ? Cout : 64; }
int Launch::wg2_q(int W, int k) {
    const int p = (k - 1) / 2;
    return W2_QC + 2 * (p * (W + 2 * p) + p);
}
// CTAs per split: slabs x output tiles x tap-pair groups
int Launch::wg2_items(int Cin, int Cout, int k) {
    const int bn = wg2_bn(Cout), pairs = (k * k + 1) / 2, gp = 512 / bn < W2_MAX_PAIRS ? 512 / bn : W2_MAX_PAIRS;
    return ((Cin + 63) / 64) * (Cout / bn) * ((pairs + gp - 1) / gp);
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
    if (Cout > 64 && Cout % 64 != 0) return false;
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
