// Per-feature mean / variance of a feature tensor that lives on the device -- the statistics of the reference's
// StandardScaler (prepare_dataset, nsga_penalty.py:102-114: X.reshape(-1, F), population variance) -- so that
// waveforms -> MFCC -> standardised features -> CNN dataset never leaves HBM.
// Two passes in fp64 (mean, then centred sum of squares): deterministic (fixed block partials, no atomics) and within
// 1e-12 of NumPy's float64 mean / var.  HBM-bound: rows * F * 4 bytes per pass.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxF = 256;

// pass 0: partial[b][f] = sum over the block's rows of x[r][f];  pass 1: ... of (x[r][f] - mean[f])^2
__global__ void __launch_bounds__(kThreads) feature_partial_kernel(const float* __restrict__ x, long long rows, int F,
                                                                   const double* __restrict__ mean, double* __restrict__ partial,
                                                                   long long rows_per_block) {
    __shared__ double red[kThreads];
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    // thread -> (feature f = tid % F, row lane = tid / F); requires F <= kThreads
    const int f = threadIdx.x % F, lane = threadIdx.x / F, lanes = kThreads / F;
    double acc = 0.0;
    if (lane < lanes) {
        const double mu = mean ? mean[f] : 0.0;
        for (long long r = r0 + lane; r < r1; r += lanes) {
            const double v = (double)x[r * F + f];
            acc += mean ? (v - mu) * (v - mu) : v;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < F) {
        double s = 0.0;
        for (int l = 0; l < lanes; ++l) s += red[l * F + threadIdx.x];
        partial[(long long)blockIdx.x * F + threadIdx.x] = s;
    }
}

__global__ void feature_finalize_kernel(const double* __restrict__ partial, int blocks, int F, double count, double* out) {
    const int f = threadIdx.x;
    if (f >= F) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += partial[(long long)b * F + f];
    out[f] = s / count;
}

}  // namespace

extern "C" {

int cmoop_feature_stats_dev(const float* feats, int64_t rows, int n_features, double* mean_host, double* var_host,
                            void* stream) {
    CMOOP_REQUIRE(feats && mean_host && var_host, "feature_stats: null pointer");
    CMOOP_REQUIRE(rows > 0 && n_features > 0 && n_features <= kMaxF, "feature_stats: need rows > 0 and 0 < n_features <= %d", kMaxF);
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)(rows < 1184 ? rows : 1184);                  // 8 per SM on a 148-SM part, fixed -> deterministic
    const long long rpb = (rows + blocks - 1) / blocks;
    double* d = (double*)cmoop::device_scratch(7, ((size_t)blocks * n_features + 2 * (size_t)n_features) * sizeof(double));
    if (!d) return CMOOP_ERR_CUDA;
    double* d_mean = d + (size_t)blocks * n_features;
    double* d_var = d_mean + n_features;
    feature_partial_kernel<<<blocks, kThreads, 0, st>>>(feats, rows, n_features, nullptr, d, rpb);
    feature_finalize_kernel<<<1, kMaxF, 0, st>>>(d, blocks, n_features, (double)rows, d_mean);
    feature_partial_kernel<<<blocks, kThreads, 0, st>>>(feats, rows, n_features, d_mean, d, rpb);
    feature_finalize_kernel<<<1, kMaxF, 0, st>>>(d, blocks, n_features, (double)rows, d_var);
    cmoop::count_launch(4);
    CMOOP_CUDA_OK(cudaGetLastError());
    CMOOP_CUDA_OK(cmoop::copy_async(mean_host, d_mean, n_features * sizeof(double), cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cmoop::copy_async(var_host, d_var, n_features * sizeof(double), cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

}  // extern "C"
