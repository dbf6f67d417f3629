// 1x1 / stride-2 skip projection of the residual blocks (nsga_penalty.py:276-279, sa_nsga_penalty.py:171-174) in precision
// bf16: forward and data gradient as a row-gathered GEMM on mma.sync (m16n8k16, bf16 operands, fp32 accumulation).
//
// The projection is a plain [pixels][Cin] x [Cin][Cout] product over every second pixel of every second row: K = Cin is
// 16...256, the output is a few MB per candidate, so it is bound by memory traffic and launch structure, not by math.  On the
// tcgen05 kernel (conv_tc.cu: one 128-row tile per CTA, TMEM allocation + barrier set-up + one K block + epilogue) a grouped
// launch of 256 candidates took 0.5-0.6 ms of CTA fixed cost for ~0.07 ms of traffic.  Here a CTA keeps a 64-channel slice
// of the K-major bf16 weights in shared memory and its 8 warps walk 32-pixel row groups with no block barrier:
//   * A fragments are 16-byte global loads of 8 consecutive channels of a pixel row (thread t of a quad owns channels
//     8t..8t+7 of every 32-channel block: the K slots of two k-steps are a fixed permutation of those channels, the same
//     permutation the B fragments use, so no shuffle or shared-memory staging is needed),
//   * B fragments are 16-byte shared-memory loads of the same 8 channels of one output-channel row,
//   * output columns are permuted (n-tile nt, column c <-> channel 2*NT*(c/2) + 2*nt + c%2) so that a thread stores 2*NT
//     consecutive channels of a pixel: 16-byte bf16 stores (forward) / 16-byte fp32 read-modify-writes into the strided
//     positions of the block-input gradient (data gradient, `accumulate`).
// Contracts are those of conv_tc_kernel for the same TcConvTask (forward: bias + optional ReLU, bf16-only output;
// data gradient: out_s = 2 scatter with accumulation into the fp32 gradient buffer).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cnn.cuh"

namespace cmoop_cnn {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerBlock = 512;      // pixels per CTA (two 32-pixel groups per warp)
constexpr int kNC = 64;                 // output channels per CTA (<= 8 n-tiles)

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NT: n-tiles (8 channels each) of the CTA's channel slice
template <int NT>
__device__ __forceinline__ void skip_body(const TcConvTask& T, unsigned char* sm, int mgroup, int nchunk, int n_b, int step) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int K = T.Cin, N = T.Cout, n0 = nchunk * (8 * NT);
    const int HoWo = T.Ho * T.Wo, M = n_b * HoWo;
    // ---- weight slice [8*NT][K] -> shared memory, row pitch K*2 + 16 bytes (conflict-free 16-byte row reads)
    const int pitch = K * 2 + 16, kch = K / 8;                 // 16-byte chunks per row
    for (int i = tid; i < 8 * NT * kch; i += kThreads) {
        const int r = i / kch, c = i - r * kch;
        *reinterpret_cast<uint4*>(sm + r * pitch + c * 16) =
            __ldg(reinterpret_cast<const uint4*>(T.wt + (long long)(n0 + r) * T.K_pad) + c);
    }
    __syncthreads();
    const __nv_bfloat16* xh = T.xh + T.x_step * step;
    const int m_lo = mgroup * kRowsPerBlock, m_hi = min(M, m_lo + kRowsPerBlock);
    // B-fragment rows of this thread: column g of n-tile nt is channel 2*NT*(g/2) + 2*nt + g%2 (row 2*nt further down)
    const unsigned char* brow0 = sm + (2 * NT * (g >> 1) + (g & 1)) * pitch + t * 16;
    const float* bias = T.bias ? T.bias + n0 + 2 * NT * t : nullptr;

    for (int mb = m_lo + warp * 32; mb < m_hi; mb += (kThreads / 32) * 32) {
        // the four pixel rows this thread reads: m-tile i (16 pixels), rows g and g + 8
        const __nv_bfloat16* arow[2][2];
        int orow[2][2];                                         // output pixel index (dense or scattered), -1: past the end
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = mb + 16 * i + 8 * h + g;
                const int mc = min(m, M - 1);
                const int n = mc / HoWo, rem = mc - n * HoWo, ho = rem / T.Wo, wo = rem - ho * T.Wo;
                const long long pin = ((long long)n * T.H + ho * T.stride) * T.W + wo * T.stride;
                arow[i][h] = xh + pin * K + 8 * t;
                const int pout = T.out_s ? (n * T.out_h + ho * T.out_s) * T.out_w + wo * T.out_s : mc;
                orow[i][h] = m < M ? pout : -1;
            }
        float acc[2][NT][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[i][nt][r] = 0.f;
        int kb = 0;
        for (; kb + 32 <= K; kb += 32) {
            uint4 a[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h) a[i][h] = __ldg(reinterpret_cast<const uint4*>(arow[i][h] + kb));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint4 b = *reinterpret_cast<const uint4*>(brow0 + 2 * nt * pitch + kb * 2);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mma16816(acc[i][nt], a[i][0].x, a[i][1].x, a[i][0].y, a[i][1].y, b.x, b.y);
                    mma16816(acc[i][nt], a[i][0].z, a[i][1].z, a[i][0].w, a[i][1].w, b.z, b.w);
                }
            }
        }
        if (kb < K) {                                           // 16-channel tail: thread t owns channels 4t..4t+3
            uint2 a[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h) a[i][h] = __ldg(reinterpret_cast<const uint2*>(arow[i][h] - 8 * t + kb + 4 * t));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint2 b = *reinterpret_cast<const uint2*>(brow0 + 2 * nt * pitch - t * 16 + (kb + 4 * t) * 2);
#pragma unroll
                for (int i = 0; i < 2; ++i) mma16816(acc[i][nt], a[i][0].x, a[i][1].x, a[i][0].y, a[i][1].y, b.x, b.y);
            }
        }
        // ---- epilogue: thread t holds channels n0 + 2*NT*t + (2*nt + e) of pixels (i, g) and (i, g + 8)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (orow[i][h] < 0) continue;
                float v[2 * NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 bb = bias ? __ldg(reinterpret_cast<const float2*>(bias) + nt) : make_float2(0.f, 0.f);
                    v[2 * nt] = acc[i][nt][2 * h] + bb.x;
                    v[2 * nt + 1] = acc[i][nt][2 * h + 1] + bb.y;
                }
                if (T.relu) {
#pragma unroll
                    for (int j = 0; j < 2 * NT; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                const long long o = (long long)orow[i][h] * N + n0 + 2 * NT * t;
                if (T.yh) {
                    uint32_t pk[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[2 * nt], v[2 * nt + 1]);
                        pk[nt] = *reinterpret_cast<const uint32_t*>(&p2);
                    }
                    if constexpr (NT == 2) {
                        *reinterpret_cast<uint2*>(T.yh + o) = make_uint2(pk[0], pk[1]);
                    } else {
#pragma unroll
                        for (int q = 0; q < NT / 4; ++q)
                            *reinterpret_cast<uint4*>(T.yh + o + 8 * q) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    }
                }
                if (T.y) {
#pragma unroll
                    for (int q = 0; q < NT / 2; ++q) {
                        float4* dst = reinterpret_cast<float4*>(T.y + o + 4 * q);
                        float4 w4 = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        if (T.accumulate) {
                            const float4 old = *dst;
                            w4.x += old.x; w4.y += old.y; w4.z += old.z; w4.w += old.w;
                        }
                        *dst = w4;
                    }
                }
            }
    }
}

__global__ void __launch_bounds__(kThreads, 2) skip_tc_kernel(const TcConvTask* __restrict__ tasks, int n_tasks, int n_b, int step) {
    extern __shared__ __align__(16) unsigned char smb[];
    __shared__ TcConvTask T;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_tasks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tasks[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        T = tasks[lo];
    }
    __syncthreads();
    const int tile = blockIdx.x - T.tile_begin;                 // tile = mgroup * tiles_n + nchunk
    const int mgroup = tile / T.tiles_n, nchunk = tile - mgroup * T.tiles_n;
    if (T.Cout >= kNC) skip_body<8>(T, smb, mgroup, nchunk, n_b, step);
    else if (T.Cout == 32) skip_body<4>(T, smb, mgroup, nchunk, n_b, step);
    else skip_body<2>(T, smb, mgroup, nchunk, n_b, step);
}

}  // namespace

// 1x1 projection, forward (stride 2, dense bf16-only output) or data gradient (stride 1 input, out_s = 2 scatter into fp32)
bool Launch::skip_tc_ok(const TcConvTask& t) {
    if (t.k != 1 || t.pad != 0) return false;
    if (t.Cin % 16 != 0 || t.Cin > 512) return false;
    if (t.Cout != 16 && t.Cout != 32 && t.Cout % kNC != 0) return false;
    if (t.K_pad % 8 != 0) return false;
    if (t.y == nullptr && t.yh == nullptr) return false;
    if (t.y != nullptr && t.yh != nullptr) return false;        // one output form per task
    return true;
}
int Launch::skip_tc_tiles_n(int Cout) { return Cout >= kNC ? Cout / kNC : 1; }
int Launch::skip_tc_tiles_m(long long M) { return (int)((M + kRowsPerBlock - 1) / kRowsPerBlock); }

int Launch::skip_tc(const TcConvTask* tasks, int n_tasks, int total_tiles, int n_b, int step, int max_k, void* stream) {
    if (n_tasks == 0 || total_tiles == 0) return 0;
    const size_t smem = (size_t)kNC * (max_k * 2 + 16);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(skip_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    skip_tc_kernel<<<total_tiles, kThreads, smem, (cudaStream_t)stream>>>(tasks, n_tasks, n_b, step);
    return (int)cudaGetLastError();
}

}  // namespace cmoop_cnn
