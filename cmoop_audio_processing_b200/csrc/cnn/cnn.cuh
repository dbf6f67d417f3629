// Population-batched candidate-CNN training / scoring: device task descriptors shared by the
// kernels (kernels.cu, conv_tc.cu) and the host executor (engine.cu).
//
// Replaces build_model / evaluate_individual / compute_objectives_and_constraints
// (nsga_penalty.py:225-442, sa_nsga_penalty.py:137-253).  Every launch is GROUPED: one grid covers
// the same stage of every active candidate (heterogeneous genotypes), blockIdx.x is mapped to a
// (candidate task, tile) pair through the tasks' tile_begin prefix.
// Layout: activations NHWC fp32, conv kernels HWIO (= GEMM B matrix [K][Cout]) immediately followed by
// the bias row in the flat parameter buffer, so bias is the (K+1)-th row of the GEMM ("ones column").
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace cmoop_cnn {

constexpr int kBatch = 64;
constexpr int kStemRows = 1024;     // output pixels per block of the stem (Cin = 1) kernels = weight-gradient split size

struct ConvTask {
    const float* x;        // input activations (or dataset base)
    const int* gather;     // per-step sample indices into x rows (stem conv on the dataset), or null
    const float* w;        // [K (+1 bias row)][Cout]
    float* y;              // output
    __nv_bfloat16* yh;     // optional bf16 shadow of y for a tensor-core consumer
    float* stat_part;      // [tiles_m][2][Cout] per-tile column sums of y and y^2 (BN batch stats), or null
    long long x_step;      // elements added to x per step (eval: contiguous batches of the split)
    long long gather_step; // elements added to gather per step
    int H, W, Cin, Ho, Wo, Cout, k, stride, pad;
    int relu;              // ReLU in the epilogue
    int use_bias;          // append the virtual ones column (K_ext = K + 1)
    int out_h, out_w, out_s;  // scattered destination grid (strided-conv dgrad); out_s == 0 -> dense [M][Cout]
    int accumulate;        // y += result
    int tiles_n, tile_begin;
};

// tcgen05 implicit GEMM (conv_tc.cu): bf16 operands, fp32 accumulation in TMEM
// A CUtensorMap (cuda.h) is an opaque 128-byte, 64-byte-aligned object; the kernels only pass its address to
// cp.async.bulk.tensor, so the task structs carry it as const void* and this header stays free of the driver API.
constexpr int kTmapBytes = 128;

struct TcConvTask {
    const __nv_bfloat16* xh;     // bf16 NHWC input (shadow copy written by the producing kernel)
    const void* tmap;            // conv_tc2: 4-D tiled tensor map {C, W, H, N} of xh, box {64, W + 2 pad, 1, 1} (device memory)
    const __nv_bfloat16* wt;     // bf16 weights, K-major [Cout][K_pad]
    const float* bias;           // fp32 [Cout] or null
    float* y;
    __nv_bfloat16* yh;           // optional bf16 shadow of y (when y feeds another tensor-core GEMM directly)
    long long x_step;
    int H, W, Cin, Ho, Wo, Cout, k, stride, pad;
    int K_pad;                   // K rounded up to a multiple of 64
    int bn;                      // N tile (UMMA N): min(Cout, 128)
    int relu;
    int out_h, out_w, out_s, accumulate;
    int tiles_n, tile_begin;
};

struct TcWgradTask {
    const void* tmaps;           // wgrad_tc2: 4 consecutive tensor maps {x, dy, x (tail batch), dy (tail batch)} as above
    const __nv_bfloat16* xh;     // forward input of the conv (bf16 NHWC shadow)
    const __nv_bfloat16* dyh;    // [M][Cout] bf16 shadow of the output gradient
    float* out;                  // grad [K+1][Cout] (splits == 1) or workspace [splits][K+1][Cout]
    int H, W, Cin, Ho, Wo, Cout, k, stride, pad;
    int splits, m_chunk;         // m_chunk multiple of 64
    int bn, tiles_k, tiles_n, tile_begin;
    int n_full;                  // wgrad_tc2: samples of a full batch (n_b == n_full uses the first pair of tensor maps)
};

struct WtBf16Task {
    const float* w;              // fp32 HWIO master weights
    __nv_bfloat16* out;
    int k, Cin, Cout, K_pad, mode;   // mode 0: forward copy, 1: flipped/transposed copy for the data gradient;
                                     // 2 / 3: the same two operands in the stage layout of conv_tc2.cu
    int block_begin;
};

struct StatTask {               // BN batch statistics of a conv output produced by the tensor-core path
    const float* y;              // fp32 conv output, or
    const __nv_bfloat16* yh;     // ... the bf16-only one (precision bf16: activations are STORED in bf16, statistics are taken
                                 // from the stored values)
    float* part;
    int C, rows_per_sample, block_begin;
};

struct WgradTask {
    const float* x;        // forward input of the conv (same addressing as ConvTask)
    const int* gather;
    const float* dy;       // [M][Cout] gradient of the conv output (fp32), or
    const __nv_bfloat16* dyh;   // ... its bf16-only form (stem kernel, precision bf16)
    float* out;            // grad buffer [K+1][Cout] (splits == 1) or workspace [splits][K+1][Cout]
    long long x_step, gather_step;
    int H, W, Cin, Ho, Wo, Cout, k, stride, pad;
    int splits, m_chunk;   // rows of M per split (multiple of 16)
    int tiles_k, tiles_n, tile_begin;
};

struct ReduceTask {       // out[i] = sum_s part[s][i]
    const float* part;
    float* out;
    int n, splits, block_begin;
};

struct WtTask {           // dgrad weights: wt[(k-1-kh, k-1-kw, co)][ci] = w[(kh,kw,ci)][co]
    const float* w;
    float* wt;
    int k, Cin, Cout, block_begin;
};

// precision bf16 stores activations in bf16 ONLY: u / v / skip / du / dskip are then null and uh / vh / skiph / duh /
// dskiph are the tensors; precision fp32 uses the fp32 pointers (vh / duh / dskiph optional shadows for tensor-core units)
struct PostTask {
    const float* u;        // conv output [n,H,W,C]
    const __nv_bfloat16* uh;
    const __nv_bfloat16* skiph;
    float* v;              // unit output [n,Ho,Wo,C]
    __nv_bfloat16* vh;     // bf16 form / shadow of v
    __nv_bfloat16* duh;    // optional bf16 shadow of du
    __nv_bfloat16* dskiph; // optional bf16 shadow of dskip
    const float* skip;     // residual branch [n,Ho,Wo,C] or null
    uint8_t* idx;          // 2x2 pool argmax code per output element, or null
    const float* gamma;    // BN parameters (null when the unit has no BN)
    const float* beta;
    float* mov_mean;
    float* mov_var;
    const float* stat_part;   // forward partial sums written by the conv epilogue
    float* bn;             // [6][C]: mean, invstd, scale, shift, mean_g, mean_gx
    const float* dv;       // backward: gradient of v
    float* du;             // gradient of u (dense, [n,H,W,C])
    float* dskip;          // gradient of the residual branch [n,Ho,Wo,C] or null
    float* dgamma;
    float* dbeta;
    float* bwd_part;       // [rows][2][C] partial sums of g and g*xhat
    int H, W, C, Ho, Wo;
    int pool, relu_mid, add_skip, relu_in, has_bn;
    int stat_tiles, bwd_rows;
    int bwd_pix;           // output pixels per CTA of post_bwd_reduce (= per partial row of bwd_part)
    int block_begin;       // grouped offset for post_fwd_kernel
    int block_begin_apply; // grouped offset for post_bwd_apply_kernel
    int block_begin_bwd;   // grouped offset for the backward reduce kernel
};

struct HeadTask {
    const float* v;        // last feature map [n,Hf,Wf,C] (fp32), or
    const __nv_bfloat16* vh;   // ... bf16-only (precision bf16)
    float* gap;            // [n,C]
    const float* dgap;     // [n,C]
    float* dv;             // [n,Hf,Wf,C]
    int Hf, Wf, C;
    int block_begin;       // gap_bwd grid: one thread per 4 consecutive elements of dv
    int block_begin_fwd;   // gap_fwd grid: one thread per 4 consecutive channels of a sample
};

struct DropTask {          // dense ReLU output u -> v = u*keep/(1-rate); backward dz = dv*keep/(1-rate)*(u>0)
    const float* u;
    float* v;
    const float* dv;
    float* dz;
    unsigned seed;
    int layer, units, use_dropout, block_begin;
};

struct CeTask {
    const float* logits;   // [n,C]
    const int* labels;     // dataset labels
    const int* gather;     // sample indices for this step or null (eval: contiguous)
    long long label_step, gather_step;
    float* dlogits;        // [n,C] (training)
    double* acc;           // [4]: loss sum, sample count, correct count, (unused)
    int* pred;             // [n_split] argmax predictions (eval/predict) or null
    int* confusion;        // [C][C] (predict) or null
    int n_classes, y_true_zero;
};

struct AdamTask {
    float* p;
    const float* g;
    float* m;
    float* v;
    int n, block_begin;
};

struct PermTask {          // per-epoch shuffle of one candidate (engine.cu make_permutation, same hash stream)
    int* perm;
    unsigned seed_lo, seed_hi;
};

struct InitTask {          // Glorot-uniform / constant initialisation of one tensor
    float* p;
    int n, tensor, kind;   // kind 0: uniform(-limit, limit), 1: constant value
    float limit, value;
    unsigned seed;
    int block_begin;
};

// dropout / init hash shared with oracle/cnn_ref.py (dropout_keep_mask)
__host__ __device__ inline unsigned fmix32(unsigned h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ inline float hash_uniform(unsigned seed, unsigned stream, unsigned step, unsigned elem) {
    unsigned h = fmix32(seed + 0x9E3779B9u * (stream + 1u));
    h = fmix32(h ^ step);
    h = fmix32(h ^ elem);
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// launchers (kernels.cu); every function returns cudaGetLastError()
struct Launch {
    static int conv(const ConvTask* tasks, int n_tasks, int total_tiles, int n_b, int step, void* stream);
    static int wgrad(const WgradTask* tasks, int n_tasks, int total_tiles, int n_b, int step, void* stream);
    static int reduce(const ReduceTask* tasks, int n_tasks, int total_blocks, void* stream);
    static int wt(const WtTask* tasks, int n_tasks, int total_blocks, void* stream);
    static int bn_finalize(const PostTask* tasks, int n_tasks, int max_c, int n_b, int training, float momentum, float eps, void* stream);
    static int post_blocks(long long out_pixels, int C);     // blocks of post_fwd / post_bwd_apply for one task
    // block_task (optional, device): task index of every block of the grouped grid -- one load instead of a binary search
    // of ~log2(n_tasks) dependent global loads at the start of every CTA
    static int post_fwd(const PostTask* tasks, int n_tasks, int total_blocks, int n_b, void* stream, const int* block_task = nullptr,
                        bool half = false);     // half: the launch's activations are stored in bf16 only
    static int bwd_pix(long long out_pixels);               // pixels per CTA / partial row of post_bwd_reduce for one task
    static int post_bwd_reduce(const PostTask* tasks, int n_tasks, int total_blocks, int n_b, void* stream, const int* block_task = nullptr,
                               bool half = false);
    static int bn_bwd_finalize(const PostTask* tasks, int n_tasks, int max_c, int n_b, void* stream);
    static int post_bwd_apply(const PostTask* tasks, int n_tasks, int total_blocks, int n_b, void* stream, const int* block_task = nullptr,
                              bool half = false);
    static int gap_blocks(long long elems);                 // blocks of gap_fwd (elems = n*C) / gap_bwd (elems = n*Hf*Wf*C) for one task
    static int gap_fwd(const HeadTask* tasks, int n_tasks, int total_blocks, int n_b, void* stream);
    static int gap_bwd(const HeadTask* tasks, int n_tasks, int total_blocks, int n_b, void* stream);
    static int drop_fwd(const DropTask* tasks, int n_tasks, int total_blocks, int n_b, int step, int training, float rate, void* stream);
    static int drop_bwd(const DropTask* tasks, int n_tasks, int total_blocks, int n_b, int step, float rate, void* stream);
    static int ce(const CeTask* tasks, int n_tasks, int n_b, int step, int training, void* stream);
    static int confusion(const int* y_true, const int* y_pred, int n, int C, int* cm, void* stream);
    static int adam(const AdamTask* tasks, int n_tasks, int total_blocks, float alpha, float b1, float b2, float eps, void* stream);
    static int init(const InitTask* tasks, int n_tasks, int total_blocks, void* stream);
    static bool perm_ok(int n);      // the permutation fits the kernel's shared memory
    static int perm(const PermTask* tasks, int n_tasks, int epoch, int n, void* stream);
    static int conv_tc(const TcConvTask* tasks, int n_tasks, int total_tiles, int n_b, int step, void* stream,
                       const int* block_task = nullptr);
    static int wgrad_tc(const TcWgradTask* tasks, int n_tasks, int total_tiles, int n_b, void* stream, const int* block_task = nullptr);
    static int wt_bf16(const WtBf16Task* tasks, int n_tasks, int total_blocks, void* stream);
    static int bn_stats(const StatTask* tasks, int n_tasks, int total_blocks, int n_b, void* stream);
    // conv_tc2.cu: patch-resident tcgen05 convolution (stride 1); tiles = ceil(n_b*Hp*Wp / tc2_rows()) * tiles_n
    static bool tc2_ok(int H, int W, int Cin, int Cout, int k, int stride);
    static int tc2_q(int W, int k);
    static int tc2_rows();
    static long long tc2_weight_elems(int Cin, int Cout, int k);     // bf16 elements of the (slab-padded) weight blocks
    // out_bf16: every task stores the bf16-only form (yh, optionally y next to it); false: fp32 y (data gradient)
    static int conv_tc2(const TcConvTask* tasks, int n_tasks, int total_tiles, int n_b, int step, int q_max, int max_cin,
                        void* stream, const int* block_task = nullptr, bool out_bf16 = false);
    static int wt_bf16_v2(const WtBf16Task* tasks, int n_tasks, int total_blocks, void* stream);
    // host side of the tiled-TMA operand loads (engine.cu): encodes the 4-D map {C, W, H, N} of a dense NHWC bf16 tensor
    // whose box {64 channels, W + 2 pad, 1, 1} at (c0, -pad, h - pad, n) is one zero-padded image row; returns 0 on success
    static int make_row_tmap(void* out128, const void* dptr, int C, int W, int H, int N, int pad);
    // wgrad_tc2.cu: tcgen05 weight gradient on the patch layout; tiles = splits * wg2_items(), TcWgradTask.m_chunk /
    // splits from wg2_splits(n_b * Hp * Wp), bn = wg2_bn(Cout), tiles_n = Cout / bn
    static bool wg2_ok(int H, int W, int Cin, int Cout, int k, int stride);
    static int wg2_bn(int Cout);
    static int wg2_q(int W, int k);
    static int wg2_items(int Cin, int Cout, int k);
    static void wg2_splits(long long Mq, int items, int* splits, int* m_chunk);
    static int wgrad_tc2(const TcWgradTask* tasks, int n_tasks, int total_tiles, int n_b, int q_max, void* stream, const int* block_task = nullptr);
    // stem.cu: dedicated Cin = 1 kernels; every task of a launch has the same M (= n_b*H*W) and W
    static bool stem_ok(int H, int W, int Cin, int Cout, int k, int stride, int n_b);
    static int stem_conv(const ConvTask* tasks, int n_tasks, int max_k, int W, int max_cout, long long M, int n_b, int step,
                         void* stream);
    static int stem_wgrad(const WgradTask* tasks, int n_tasks, int max_k, int W, int max_cout, int splits, int n_b, int step,
                          void* stream);
    // skip_tc.cu: the 1x1 / stride-2 skip projection (forward and data gradient) as a row-gathered mma.sync GEMM;
    // tiles of a task = skip_tc_tiles_m(n_b * Ho * Wo) * skip_tc_tiles_n(Cout), TcConvTask.tiles_n = skip_tc_tiles_n(Cout)
    static bool skip_tc_ok(const TcConvTask& t);
    static int skip_tc_tiles_n(int Cout);
    static int skip_tc_tiles_m(long long M);
    static int skip_tc(const TcConvTask* tasks, int n_tasks, int total_tiles, int n_b, int step, int max_k, void* stream);
    // stem_tc.cu: the same two operations on mma.sync (bf16 hi + lo operands) for precision bf16 -- tasks carry yh / dyh only
    static bool stem_tc_ok(int H, int W, int Cout, int k);
    static int stem_conv_tc(const ConvTask* tasks, int n_tasks, int max_k, int W, long long M, int n_b, int step, void* stream);
    static int stem_wgrad_tc(const WgradTask* tasks, int n_tasks, int max_k, int W, int max_cout, int splits, int n_b, int step,
                             void* stream);
};

}  // namespace cmoop_cnn
