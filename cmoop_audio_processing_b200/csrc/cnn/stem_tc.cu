// Stem convolution (Cin = 1) of the candidate CNNs on the warp-level tensor-core path (mma.sync m16n8k16, bf16 operands,
// fp32 accumulation) for precision bf16, where the convolution output and its gradient are STORED in bf16.
//
// The first Conv2D of both model families (nsga_penalty.py:255, sa_nsga_penalty.py:150) has K = k*k (+1 bias) = 10 or 26:
// far too thin for tcgen05 (a 128-row UMMA would be 80-90 % padding) and, in fp32 SIMT form (stem.cu), bound by FMA issue at
// 6-8x the time its HBM traffic needs (2.35 ms forward / 3.2 ms weight gradient per grouped launch of 256 candidates against
// ~0.4 ms of bf16 activation traffic).  Here the contraction runs on mma.sync with the fp32 inputs carried as bf16 hi + lo
// pairs (x = hi + lo to 2^-17, three / two products per term), so the arithmetic stays fp32-grade (the rounding statement of
// oracle/cnn_ref.py -- "the Cin = 1 stem's forward arithmetic stays fp32" -- is unchanged: the only bf16 rounding is the
// stored output / the stored output gradient) while the instruction count per output drops ~5x:
//
//  * stem_conv_tc:  D[channel][pixel] = W^T[channel][tap] * X^T[tap][pixel].  K slots are (kw, kw+1) tap PAIRS of one kernel
//    row, so a B-fragment register is ONE aligned 32-bit shared-memory load from a bf16 copy of the padded feature-map rows
//    (two copies, shifted by one element, make every pair aligned); the bias is the pair (1, 0).  Weights live in registers
//    as A fragments for the whole block.  The M rows are a permutation of the channels chosen so that, after one movmatrix
//    transposition per 8x8 block, a thread owns 8 (4) consecutive channels of a pixel and stores them as one 16-byte (8-byte)
//    word.  BN batch statistics of the STORED values come from the tensor core as well: the packed bf16 output fragment is
//    both the A operand of a sum (times ones) and the A and B operand of a Gram product whose diagonal is the sum of squares.
//  * stem_wgrad_tc: dW[tap][channel] = X^T[tap][pixel] * dY[pixel][channel] over the pixels of a split.  K slots are pixel
//    pairs (even/odd column of one image row; odd widths get a zero pad column), A fragments are 32-bit loads of the same
//    shifted bf16 copies (hi + lo), B fragments come from a 3-stage cp.async ring of dY rows through ldmatrix.trans.
//    Per-split partials [split][K+1][Cout] and their fixed-order reduction are those of stem.cu (deterministic).
//
// Contracts (ConvTask / WgradTask, 1 024 pixels per block, 64-row BN tiles) are those of stem.cu.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int kThreads = 256;
constexpr int kWgStageRows = 128;     // pixels per cp.async stage of the weight gradient (one k-step per warp)
constexpr int kWgStages = 3;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint16_t bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi = __bfloat16_as_ushort(h);
    lo = bf16_bits(v - __bfloat162float(h));
}
__device__ __forceinline__ uint32_t pack_relu(float a, float b, int relu) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    if (relu) v = __hmax2(v, __floats2bfloat162_rn(0.f, 0.f));
    return *reinterpret_cast<uint32_t*>(&v);
}

struct Segs {
    int n_first, ha[2];
};

// x / d for 0 <= x < 2^24 without the integer-division sequence (float reciprocal + one correction step)
struct FastDiv {
    float inv;
    int d;
    __device__ __forceinline__ explicit FastDiv(int d_) : inv(1.0f / (float)d_), d(d_) {}
    __device__ __forceinline__ int div(int x) const {
        int q = (int)((float)x * inv);
        const int r = x - q * d;
        if (r < 0) --q;
        if (r >= d) ++q;
        return q;
    }
};

// The feature-map rows (<= 2 samples) touched by output pixels [m0, m1), zero padding included, as four bf16 copies of
// pitch Wr (even): hi / lo parts, each at element offset 0 ("even") and shifted down by one element ("odd"), so that the
// element pair starting at ANY offset i is the aligned 32-bit word i >> 1 of the copy selected by i & 1.
// Layout (ushort units): hiE @0, hiO @CE, loE @2CE, loO @3CE, CE = 2*seg + 8.  A thread produces whole 32-bit words: word
// wi of the even copies holds elements (2wi, 2wi+1), word wi of the odd copies (2wi+1, 2wi+2).
template <int KS>
__device__ __forceinline__ void stage_x(uint16_t* sm, int seg, const float* xbase, const int* gather, int H, int W, int m0,
                                        int m1, Segs& sg) {
    constexpr int P = (KS - 1) / 2;
    const int HW = H * W, Wr = (W + 2 * P + 1) & ~1, CE = 2 * seg + 8;
    const FastDiv dW(W), dWr(Wr);
    sg.n_first = m0 / HW;
    // segment 0 / 1: first image row, staged rows (halo included), source sample
    const int n0 = sg.n_first, lo0 = max(m0, n0 * HW), hi0 = min(m1, (n0 + 1) * HW);
    const int lo1 = max(m0, (n0 + 1) * HW), hi1 = m1;
    const int ha0 = dW.div(lo0 - n0 * HW), rows0 = dW.div(hi0 - 1 - n0 * HW) - ha0 + 1 + 2 * P;
    const bool has1 = lo1 < hi1;
    const int rows1 = has1 ? dW.div(hi1 - 1 - (n0 + 1) * HW) + 1 + 2 * P : 0;     // segment 1 starts at image row 0
    const float* src0 = xbase + (long long)(gather ? gather[n0] : n0) * HW;
    const float* src1 = has1 ? xbase + (long long)(gather ? gather[n0 + 1] : n0 + 1) * HW : xbase;
    sg.ha[0] = ha0;
    sg.ha[1] = 0;
    auto value = [&](int idx) -> float {
        const bool s = idx >= seg;
        const int i = idx - (s ? seg : 0);
        const int r = dWr.div(i), c = i - r * Wr;
        const int hi_ = (s ? 0 : ha0) - P + r, wi_ = c - P;
        const bool ok = idx < 2 * seg && r < (s ? rows1 : rows0) && (unsigned)hi_ < (unsigned)H && (unsigned)wi_ < (unsigned)W;
        return ok ? __ldg((s ? src1 : src0) + hi_ * W + wi_) : 0.f;
    };
    uint32_t* w32 = reinterpret_cast<uint32_t*>(sm);
    const int CW = CE >> 1;                              // words per copy
    for (int wi = threadIdx.x; wi < CW; wi += kThreads) {
        const float v0 = value(2 * wi), v1 = value(2 * wi + 1), v2 = 2 * wi + 2 < CE ? value(2 * wi + 2) : 0.f;
        uint16_t h0, l0, h1, l1, h2, l2;
        split_bf16(v0, h0, l0);
        split_bf16(v1, h1, l1);
        split_bf16(v2, h2, l2);
        w32[wi] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        w32[CW + wi] = (uint32_t)h1 | ((uint32_t)h2 << 16);
        w32[2 * CW + wi] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        w32[3 * CW + wi] = (uint32_t)l1 | ((uint32_t)l2 << 16);
    }
}

// ------------------------------------------------------------------------------------------------ forward
// KS: kernel size; MTW: 16-channel m-tiles per warp (a warp owns 16*MTW consecutive channels of its pixels)
template <int KS, int MTW>
__device__ __forceinline__ void stem_conv_tc_body(const ConvTask& T, uint16_t* sm, int chunk0, int n_chunks, int seg, int n_b,
                                                  int step) {
    constexpr int P = (KS - 1) / 2, TAPS = KS * KS, PPR = (KS + 1) / 2, NPAIR = KS * PPR + 1, KST = (NPAIR + 7) / 8;
    constexpr int CPT = 4 * MTW;                       // consecutive channels a thread stores per pixel
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int H = T.H, W = T.W, HW = H * W, Wr = (W + 2 * P + 1) & ~1, Cout = T.Cout, CE = 2 * seg + 8;
    const int M = n_b * HW;
    if (chunk0 * kStemRows >= M) return;
    // weights of this warp's channel group as A fragments (hi / lo), rows permuted: row 8i + 2t' + e of m-tile j is channel
    // cbase + CPT*t' + 4j + 2i + e
    const int groups = Cout / (16 * MTW), grp = warp % groups, cbase = grp * 16 * MTW;
    uint32_t ah[MTW][KST][4], al[MTW][KST][4];
#pragma unroll
    for (int j = 0; j < MTW; ++j)
#pragma unroll
        for (int s = 0; s < KST; ++s)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = r & 1, pair = t + (r >> 1) * 4 + 8 * s;
                const int ch = cbase + CPT * (g >> 1) + 4 * j + 2 * i + (g & 1);
                float wv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kh = pair / PPR, kw = 2 * (pair - kh * PPR) + e;
                    wv[e] = 0.f;
                    if (pair < KS * PPR) {
                        if (kw < KS) wv[e] = __ldg(T.w + (kh * KS + kw) * Cout + ch);
                    } else if (pair == KS * PPR && e == 0) {
                        wv[e] = __ldg(T.w + TAPS * Cout + ch);
                    }
                }
                uint16_t h0, l0, h1, l1;
                split_bf16(wv[0], h0, l0);
                split_bf16(wv[1], h1, l1);
                ah[j][s][r] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                al[j][s][r] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
    // this thread's B-fragment pairs: q = 2s + r -> pair t + 4q; word offsets inside the tile
    int dq[2 * KST];
    bool isb[2 * KST];
#pragma unroll
    for (int q = 0; q < 2 * KST; ++q) {
        const int pair = t + 4 * q, kh = pair / PPR, kw0 = 2 * (pair - kh * PPR);
        isb[q] = pair == KS * PPR;
        dq[q] = pair < KS * PPR ? (kh * Wr + kw0) >> 1 : 0;
    }
    const uint32_t* base32 = reinterpret_cast<const uint32_t*>(sm);
    uint16_t* tab = sm + 4 * CE;                       // [kStemRows] tile offset of every pixel of the chunk
    const int relu = T.relu;
    const bool stats = T.stat_part != nullptr;
    const FastDiv dW(W);
    // the block walks n_chunks consecutive 1 024-pixel chunks with its weights in registers
    for (int cc = 0; cc < n_chunks; ++cc) {
    const int m0 = (chunk0 + cc) * kStemRows, m1 = min(M, m0 + kStemRows);
    if (m0 >= M) break;
    if (cc) __syncthreads();                           // the previous chunk's tile is fully consumed
    Segs sg;
    stage_x<KS>(sm, seg, T.x + T.x_step * step, T.gather ? T.gather + T.gather_step * step : nullptr, H, W, m0, m1, sg);
    for (int i = tid; i < kStemRows; i += kThreads) {
        const int m = min(m0 + i, m1 - 1);
        const int s = m >= (sg.n_first + 1) * HW ? 1 : 0, rem = m - (sg.n_first + s) * HW;
        const int h = dW.div(rem), w = rem - h * W;
        tab[i] = (uint16_t)(s * seg + (h - (s ? 0 : sg.ha[0])) * Wr + w);
    }
    __syncthreads();
    const int tiles = (m1 - m0 + 63) >> 6;
    for (int tl = warp / groups; tl < tiles; tl += (kThreads / 32) / groups) {
        float s1[MTW][4], s2[MTW][4];
#pragma unroll
        for (int j = 0; j < MTW; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) s1[j][r] = s2[j][r] = 0.f;
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {
            float c[MTW][2][4];
#pragma unroll
            for (int j = 0; j < MTW; ++j)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) c[j][nt][r] = 0.f;
            const int pl0 = tl * 64 + sub * 16 + g;           // block-local pixel of n-tile 0 (n-tile 1: + 8)
            uint32_t bh[2][2 * KST], bl[2][2 * KST];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int pl = pl0 + nt * 8;
                const bool valid = m0 + pl < m1;
                const int off = tab[pl];
                const uint32_t* ph = base32 + ((off & 1) ? (CE >> 1) : 0) + (off >> 1);
                const uint32_t* pw = ph + CE;
#pragma unroll
                for (int q = 0; q < 2 * KST; ++q) {
                    bh[nt][q] = ph[dq[q]];
                    bl[nt][q] = pw[dq[q]];
                    if (isb[q]) { bh[nt][q] = 0x00003F80u; bl[nt][q] = 0u; }
                    if (!valid) { bh[nt][q] = 0u; bl[nt][q] = 0u; }
                }
            }
            // three product terms per k-step; consecutive MMAs go to different accumulators (no dependent issue)
#pragma unroll
            for (int s = 0; s < KST; ++s) {
#pragma unroll
                for (int j = 0; j < MTW; ++j)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma16816(c[j][nt], al[j][s], bh[nt][2 * s], bh[nt][2 * s + 1]);
#pragma unroll
                for (int j = 0; j < MTW; ++j)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma16816(c[j][nt], ah[j][s], bl[nt][2 * s], bl[nt][2 * s + 1]);
#pragma unroll
                for (int j = 0; j < MTW; ++j)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma16816(c[j][nt], ah[j][s], bh[nt][2 * s], bh[nt][2 * s + 1]);
            }
            // epilogue: ReLU + bf16 rounding, statistics of the stored values, transposition, one vector store per pixel
            uint32_t tr[2][MTW][2];
#pragma unroll
            for (int j = 0; j < MTW; ++j) {
                uint32_t y[4];                                  // A fragment [16 channels][16 pixels] of the packed output
                y[0] = pack_relu(c[j][0][0], c[j][0][1], relu);
                y[1] = pack_relu(c[j][0][2], c[j][0][3], relu);
                y[2] = pack_relu(c[j][1][0], c[j][1][1], relu);
                y[3] = pack_relu(c[j][1][2], c[j][1][3], relu);
                if (stats) {
                    // sum: Y * ones; sum of squares: diagonal of Y * Y^T, rows 0-7 against channels 0-7 and rows 8-15
                    // against channels 8-15 (the other row half of A zeroed) into ONE accumulator
                    mma16816(s1[j], y, 0x3F803F80u, 0x3F803F80u);
                    mma16816(s2[j], y[0], 0u, y[2], 0u, y[0], y[2]);
                    mma16816(s2[j], 0u, y[1], 0u, y[3], y[1], y[3]);
                }
                tr[0][j][0] = movmatrix_trans(y[0]);
                tr[0][j][1] = movmatrix_trans(y[1]);
                tr[1][j][0] = movmatrix_trans(y[2]);
                tr[1][j][1] = movmatrix_trans(y[3]);
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int m = m0 + pl0 + nt * 8;
                if (m < m1) {
                    __nv_bfloat16* dst = T.yh + (long long)m * Cout + cbase + CPT * t;
                    if constexpr (MTW == 2)
                        *reinterpret_cast<uint4*>(dst) = make_uint4(tr[nt][0][0], tr[nt][0][1], tr[nt][1][0], tr[nt][1][1]);
                    else
                        *reinterpret_cast<uint2*>(dst) = make_uint2(tr[nt][0][0], tr[nt][0][1]);
                }
            }
        }
        if (stats) {
            float* dst = T.stat_part + (long long)((m0 >> 6) + tl) * 2 * Cout;
#pragma unroll
            for (int j = 0; j < MTW; ++j) {
                const int ch0 = cbase + CPT * (g >> 1) + 4 * j + (g & 1);     // row g (i = 0); row g + 8 (i = 1): + 2
                if (t == 0) {
                    dst[ch0] = s1[j][0];
                    dst[ch0 + 2] = s1[j][2];
                }
                if (t == (g >> 1)) {                                           // diagonal of the two Gram blocks
                    dst[Cout + ch0] = (g & 1) ? s2[j][1] : s2[j][0];
                    dst[Cout + ch0 + 2] = (g & 1) ? s2[j][3] : s2[j][2];
                }
            }
        }
    }
    }
}

__global__ void __launch_bounds__(kThreads, 2) stem_conv_tc_kernel(const ConvTask* __restrict__ tasks, int blocks_per_task,
                                                                   int chunks_per_block, int seg, int n_b, int step) {
    extern __shared__ __align__(16) uint16_t smh[];
    __shared__ ConvTask T;
    const int task = blockIdx.x / blocks_per_task, chunk0 = (blockIdx.x - task * blocks_per_task) * chunks_per_block;
    if (threadIdx.x == 0) T = tasks[task];
    __syncthreads();
    if (T.k == 3) {
        if (T.Cout == 16) stem_conv_tc_body<3, 1>(T, smh, chunk0, chunks_per_block, seg, n_b, step);
        else stem_conv_tc_body<3, 2>(T, smh, chunk0, chunks_per_block, seg, n_b, step);
    } else {
        if (T.Cout == 16) stem_conv_tc_body<5, 1>(T, smh, chunk0, chunks_per_block, seg, n_b, step);
        else stem_conv_tc_body<5, 2>(T, smh, chunk0, chunks_per_block, seg, n_b, step);
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// out[split][tap][co] = sum_{m in split} x[m + tap] dy[m][co]  (row TAPS = bias gradient = sum of dy).
// K runs over "q" positions: q = R * Wq + w with R the global image row (n*H + h) and Wq = W rounded up to even, so that
// K pairs (q, q+1) are horizontally adjacent pixels; the pad column of an odd W and positions outside [m0, m1) carry a
// zero dY row.
template <int KS, int C16>
__device__ __forceinline__ void stem_wgrad_tc_body(const WgradTask& T, uint16_t* sm, int split, int seg, int tabn, int n_b,
                                                   int step) {
    constexpr int P = (KS - 1) / 2, TAPS = KS * KS, MT = (TAPS + 1 + 15) / 16, NT = 2 * C16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int H = T.H, W = T.W, HW = H * W, Wr = (W + 2 * P + 1) & ~1, Wq = (W + 1) & ~1, Cout = 16 * C16, CE = 2 * seg + 8;
    const int M = n_b * HW, m0 = split * T.m_chunk, m1 = min(M, m0 + T.m_chunk);
    const int n_out = (TAPS + 1) * Cout;
    float* out = T.out + (long long)split * n_out;
    if (m0 >= M) {                                    // empty split: its partial must still be defined
        for (int i = tid; i < n_out; i += kThreads) out[i] = 0.f;
        return;
    }
    // q range of the split (pair aligned)
    const FastDiv dW(W), dWq(Wq), dH(H);
    const int R0 = dW.div(m0), R1 = dW.div(m1 - 1);
    const int qs = (R0 * Wq + (m0 - R0 * W)) & ~1, qe = R1 * Wq + (m1 - 1 - R1 * W) + 1;
    const int n_stage = (qe - qs + kWgStageRows - 1) / kWgStageRows;
    uint16_t* tab = sm + 4 * CE;                       // [tabn] tile offset of the pair (qs + 2i, qs + 2i + 1)
    const int pitch = Cout * 2 + 16;                   // bytes per dY row in the ring (16-byte skew: conflict-free ldmatrix)
    unsigned char* ring = reinterpret_cast<unsigned char*>(sm + 4 * CE + ((tabn + 7) & ~7));
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const int stage_bytes = kWgStageRows * pitch;
    const __nv_bfloat16* dy = T.dyh;
    // cp.async producer: a thread owns one row position of the stage and every (256 / rows)-th 16-byte chunk
    constexpr int CH = 2 * C16;                        // 16-byte chunks per row
    auto issue = [&](int st) {
        if (st < n_stage) {
            const uint32_t dst0 = ring_s + (st % kWgStages) * stage_bytes;
            for (int it = tid; it < kWgStageRows * CH; it += kThreads) {
                const int row = it / CH, cch = it - row * CH;
                const int q = qs + st * kWgStageRows + row, R = dWq.div(q), w = q - R * Wq;
                const long long m = (long long)R * W + w;
                const bool ok = w < W && m >= m0 && m < m1;
                cp_async16(dst0 + row * pitch + cch * 16, dy + (ok ? m * Cout + cch * 8 : 0), ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };
    issue(0);                                          // the first dY stages fly while the x tile is staged
    issue(1);
    Segs sg;
    stage_x<KS>(sm, seg, T.x + T.x_step * step, T.gather ? T.gather + T.gather_step * step : nullptr, H, W, m0, m1, sg);
    for (int i = tid; i < tabn; i += kThreads) {
        const int q = qs + 2 * i, R = dWq.div(q), w = q - R * Wq;
        int off = 0;
        if (q < qe && R >= R0 && R <= R1) {
            const int n = dH.div(R), h = R - n * H, s = n - sg.n_first;
            off = s * seg + (h - (s ? 0 : sg.ha[0])) * Wr + w;
        }
        tab[i] = (uint16_t)off;
    }

    // A-fragment rows of this thread: taps 16*mt + g and 16*mt + 8 + g
    const uint32_t* base32 = reinterpret_cast<const uint32_t*>(sm);
    const uint32_t* rp[MT][2];
    int kind[MT][2];                                   // 0: tap, 1: bias row (ones), 2: zero row
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int tap = 16 * mt + 8 * i + g;
            kind[mt][i] = tap < TAPS ? 0 : (tap == TAPS ? 1 : 2);
            const int kh = tap < TAPS ? tap / KS : 0, kw = tap < TAPS ? tap - kh * KS : 0;
            rp[mt][i] = base32 + ((kw & 1) ? (CE >> 1) : 0) + ((kh * Wr + kw) >> 1);
        }
    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
    // ldmatrix lane address inside a stage: matrices {px 0-7, ch 0-7}, {px 8-15, ch 0-7}, {px 0-7, ch 8-15}, {px 8-15, ch 8-15}
    const int lm_row = warp * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), lm_col = (lane >> 4) * 16;

    for (int st = 0; st < n_stage; ++st) {
        cp_async_wait<1>();
        __syncthreads();                               // stage st landed for every thread; stage st-1 fully consumed
        issue(st + 2);
        const int ip0 = (st * kWgStageRows + warp * 16) >> 1;
        const int offa = tab[ip0 + t] >> 1, offb = tab[ip0 + t + 4] >> 1;
        uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = r & 1, o = (r >> 1) ? offb : offa;
                uint32_t h = rp[mt][i][o], l = rp[mt][i][o + CE];
                if (kind[mt][i] == 1) { h = 0x3F803F80u; l = 0u; }
                if (kind[mt][i] == 2) { h = 0u; l = 0u; }
                ahi[mt][r] = h;
                alo[mt][r] = l;
            }
        const uint32_t sbase = ring_s + (st % kWgStages) * stage_bytes + lm_row * pitch + lm_col;
        // 32 channels at a time; the lo pass over all accumulators, then the hi pass (no dependent back-to-back MMAs)
        constexpr int CG = C16 >= 2 ? 2 : 1;
#pragma unroll
        for (int c0 = 0; c0 < C16; c0 += CG) {
            uint32_t b[CG][4];
#pragma unroll
            for (int cc = 0; cc < CG; ++cc) ldmatrix_x4_trans(b[cc], sbase + (c0 + cc) * 32);
#pragma unroll
            for (int cc = 0; cc < CG; ++cc)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816(acc[mt][2 * (c0 + cc)], alo[mt], b[cc][0], b[cc][1]);
                    mma16816(acc[mt][2 * (c0 + cc) + 1], alo[mt], b[cc][2], b[cc][3]);
                }
#pragma unroll
            for (int cc = 0; cc < CG; ++cc)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816(acc[mt][2 * (c0 + cc)], ahi[mt], b[cc][0], b[cc][1]);
                    mma16816(acc[mt][2 * (c0 + cc) + 1], ahi[mt], b[cc][2], b[cc][3]);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // the 8 warps' partials, summed in warp order (aliases the ring)
    float* red = reinterpret_cast<float*>(ring);       // [8][(TAPS+1)][Cout]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int tap = 16 * mt + 8 * i + g;
                if (tap <= TAPS)
                    *reinterpret_cast<float2*>(red + (warp * (TAPS + 1) + tap) * Cout + nt * 8 + 2 * t) =
                        make_float2(acc[mt][nt][2 * i], acc[mt][nt][2 * i + 1]);
            }
    __syncthreads();
    for (int i = tid; i < n_out; i += kThreads) {
        float a = 0.f;
#pragma unroll
        for (int wq = 0; wq < kThreads / 32; ++wq) a += red[wq * n_out + i];
        out[i] = a;
    }
}

__global__ void __launch_bounds__(kThreads, 2) stem_wgrad_tc_kernel(const WgradTask* __restrict__ tasks, int blocks_per_task,
                                                                    int seg, int tabn, int n_b, int step) {
    extern __shared__ __align__(16) uint16_t smh[];
    __shared__ WgradTask T;
    const int task = blockIdx.x / blocks_per_task, split = blockIdx.x - task * blocks_per_task;
    if (threadIdx.x == 0) T = tasks[task];
    __syncthreads();
    if (T.k == 3) {
        switch (T.Cout) {
            case 16: stem_wgrad_tc_body<3, 1>(T, smh, split, seg, tabn, n_b, step); break;
            case 32: stem_wgrad_tc_body<3, 2>(T, smh, split, seg, tabn, n_b, step); break;
            case 64: stem_wgrad_tc_body<3, 4>(T, smh, split, seg, tabn, n_b, step); break;
            default: stem_wgrad_tc_body<3, 8>(T, smh, split, seg, tabn, n_b, step); break;
        }
    } else {
        switch (T.Cout) {
            case 16: stem_wgrad_tc_body<5, 1>(T, smh, split, seg, tabn, n_b, step); break;
            case 32: stem_wgrad_tc_body<5, 2>(T, smh, split, seg, tabn, n_b, step); break;
            default: stem_wgrad_tc_body<5, 4>(T, smh, split, seg, tabn, n_b, step); break;
        }
    }
}

int seg_elems_for(int W, int k, int rows) {
    const int P = (k - 1) / 2, Wr = (W + 2 * P + 1) & ~1;
    return ((rows / W + 2 + 2 * P) * Wr + 7) & ~7;
}
int wg_tabn(int W) {                                   // pairs of q positions a 1 024-pixel split can span (+ a stage of slack)
    const int Wq = (W + 1) & ~1;
    const int qspan = kStemRows + (kStemRows / W + 2) * (Wq - W) + 2;
    return ((qspan + kWgStageRows - 1) / kWgStageRows * kWgStageRows) / 2 + 8;
}

template <class K>
cudaError_t opt_in(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem < 48 * 1024 ? 48 * 1024 : smem));
}

size_t conv_smem(int W, int k) { return (size_t)(4 * (2 * seg_elems_for(W, k, kStemRows) + 8) + kStemRows) * 2; }
size_t wgrad_smem(int W, int k, int cout) {
    const size_t ring = (size_t)kWgStages * kWgStageRows * (cout * 2 + 16);
    const size_t red = (size_t)8 * (k * k + 1) * cout * 4;
    return (size_t)(4 * (2 * seg_elems_for(W, k, kStemRows) + 8) + ((wg_tabn(W) + 7) & ~7)) * 2 + (ring > red ? ring : red);
}

}  // namespace

// precision bf16 only (bf16-only output / output gradient); the shapes of stem_ok with the channel counts the fragments cover
bool Launch::stem_tc_ok(int H, int W, int Cout, int k) {
    if (k != 3 && k != 5) return false;
    if (Cout != 16 && Cout != 32 && Cout != 64 && !(Cout == 128 && k == 3)) return false;
    if ((long long)H * W < kStemRows || W < 16) return false;       // a block may touch at most two samples
    if (2 * seg_elems_for(W, k, kStemRows) + 8 > 65535) return false;   // 16-bit tile offsets
    return conv_smem(W, k) <= 100 * 1024 && wgrad_smem(W, k, Cout) <= 110 * 1024;
}

int Launch::stem_conv_tc(const ConvTask* tasks, int n_tasks, int max_k, int W, long long M, int n_b, int step, void* stream) {
    const int chunks = (int)((M + kStemRows - 1) / kStemRows);
    // chunks per block: amortise the weight-fragment set-up while keeping >= ~8 blocks per SM slot in the grid
    int cpb = 4;
    while (cpb > 1 && (long long)n_tasks * ((chunks + cpb - 1) / cpb) < 148 * 2 * 8) cpb >>= 1;
    const int bpt = (chunks + cpb - 1) / cpb;
    const int seg = seg_elems_for(W, max_k, kStemRows);
    const size_t smem = conv_smem(W, max_k);
    cudaError_t e = opt_in(stem_conv_tc_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    stem_conv_tc_kernel<<<n_tasks * bpt, kThreads, smem, (cudaStream_t)stream>>>(tasks, bpt, cpb, seg, n_b, step);
    return (int)cudaGetLastError();
}

int Launch::stem_wgrad_tc(const WgradTask* tasks, int n_tasks, int max_k, int W, int max_cout, int splits, int n_b, int step,
                          void* stream) {
    const int seg = seg_elems_for(W, max_k, kStemRows);
    const size_t smem = wgrad_smem(W, max_k, max_cout);
    cudaError_t e = opt_in(stem_wgrad_tc_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    stem_wgrad_tc_kernel<<<n_tasks * splits, kThreads, smem, (cudaStream_t)stream>>>(tasks, splits, seg, wg_tabn(W), n_b, step);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
