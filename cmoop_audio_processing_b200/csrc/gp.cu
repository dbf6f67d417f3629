// Batched Gaussian-process posterior (mean + std), fp64, one warp per (model, query).
//
// Replaces GaussianProcessRegressor.predict(X, return_std=True) as called by
// SurrogateManager.predict (ablation_study/sa_nsga_local.py:212-223,
// sa_nsga_penalty.py:342-363) and predict_gps (mobo_penalty.py:265-273); the
// arithmetic restated is sklearn/gaussian_process/_gpr.py:446-499 with the Matern
// kernel of kernels.py:1713-1743.
//
//   k*[i]   = amp * matern(|| xq/l - x_i/l ||)          lanes own rows i = lane + 32 r
//   mean    = sum_i k*[i] alpha[i]                       warp shuffle reduction
//   v       = L^-1 k*                                    column-oriented forward substitution:
//             the right-hand side stays in registers (R = ceil(n/32) doubles per lane), the
//             solved v_j is broadcast with one shuffle, and column j of L is read as a
//             contiguous row of L^T (uploaded transposed), i.e. fully coalesced
//   std     = sqrt(max(0, amp + noise - sum_j v_j^2))
// n <= 1024 is latency-bound (n <= 288 in the reference: the genotype space has 288
// points and training rows are de-duplicated, sa_nsga_local.py:200-202).
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

struct DevModel {
    int n;
    double amplitude, inv_length, nu, noise, y_scale, y_shift;
    const double* x_train;   // [n][dim]
    const double* alpha;     // [n]
    const double* chol_t;    // [n][n] = L^T row-major (null: mean only)
};

constexpr int kWarpsPerBlock = 8;

template <int R>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) gp_predict_kernel(const DevModel* __restrict__ models,
                                                                         const double* __restrict__ xq, int q, int dim,
                                                                         double* __restrict__ mean_out,
                                                                         double* __restrict__ std_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * kWarpsPerBlock + warp;
    if (qi >= q) return;
    const DevModel mdl = models[blockIdx.y];
    const int n = mdl.n;
    const double* xrow = xq + (size_t)qi * dim;

    double rhs[R];
    double part = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = lane + 32 * r;
        double kv = 0.0;
        if (i < n) {
            double ss = 0.0;
            for (int k = 0; k < dim; ++k) {
                const double df = xrow[k] * mdl.inv_length - mdl.x_train[(size_t)i * dim + k] * mdl.inv_length;
                ss += df * df;
            }
            const double d = sqrt(ss);
            double kval;
            if (mdl.nu == 1.5) {
                const double t = d * 1.7320508075688772;
                kval = (1.0 + t) * exp(-t);
            } else if (mdl.nu == 2.5) {
                const double t = d * 2.23606797749979;
                kval = (1.0 + t + t * t / 3.0) * exp(-t);
            } else {
                kval = exp(-d);
            }
            kv = mdl.amplitude * kval;
            part += kv * mdl.alpha[i];
        }
        rhs[r] = kv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) mean_out[(size_t)blockIdx.y * q + qi] = part * mdl.y_scale + mdl.y_shift;
    if (std_out == nullptr) return;

    double sumsq = 0.0;
    if (mdl.chol_t != nullptr) {
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (32 * s < n) {
                for (int jj = 0; jj < 32; ++jj) {
                    const int j = 32 * s + jj;
                    if (j >= n) break;
                    const double* col = mdl.chol_t + (size_t)j * n;
                    const double mine = rhs[s] / col[j];
                    const double v = __shfl_sync(0xffffffffu, mine, jj);
                    sumsq += v * v;
#pragma unroll
                    for (int r = s; r < R; ++r) {
                        const int i = lane + 32 * r;
                        if (i > j && i < n) rhs[r] -= col[i] * v;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        double var = (mdl.amplitude + mdl.noise) - sumsq;
        if (var < 0.0) var = 0.0;
        std_out[(size_t)blockIdx.y * q + qi] = sqrt(var) * mdl.y_scale;
    }
}

}  // namespace

struct cmoop_gp {
    int n_models = 0, dim = 0, max_n = 0;
    bool has_chol = true;
    DevModel* d_models = nullptr;
    std::vector<void*> owned;
};

extern "C" {

int cmoop_gp_destroy(cmoop_gp_handle h) {
    if (!h) return CMOOP_OK;
    for (void* p : h->owned) cudaFree(p);
    if (h->d_models) cudaFree(h->d_models);
    delete h;
    return CMOOP_OK;
}

int cmoop_gp_create(const cmoop_gp_model* models, int n_models, cmoop_gp_handle* out) {
    CMOOP_REQUIRE(models && out && n_models > 0, "gp_create: bad arguments");
    *out = nullptr;
    const int dim = models[0].dim;
    for (int i = 0; i < n_models; ++i) {
        CMOOP_REQUIRE(models[i].dim == dim && dim > 0, "gp_create: all models must share dim > 0");
        CMOOP_REQUIRE(models[i].n_train > 0 && models[i].n_train <= 1024, "gp_create: n_train=%d outside [1,1024]",
                      models[i].n_train);
        CMOOP_REQUIRE(models[i].x_train && models[i].alpha, "gp_create: null x_train/alpha");
        CMOOP_REQUIRE(models[i].nu == 0.5 || models[i].nu == 1.5 || models[i].nu == 2.5,
                      "gp_create: nu must be 0.5, 1.5 or 2.5");
        CMOOP_REQUIRE(models[i].length_scale > 0.0, "gp_create: length_scale must be positive");
    }
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cmoop_gp* h = new cmoop_gp();
    h->n_models = n_models;
    h->dim = dim;
    std::vector<DevModel> host(n_models);
    auto upload = [&](const double* src, size_t count, const double** dst) -> int {
        void* d = nullptr;
        CMOOP_CUDA_OK(cudaMalloc(&d, count * sizeof(double)));
        h->owned.push_back(d);
        CMOOP_CUDA_OK(cmoop::copy_sync(d, src, count * sizeof(double), cudaMemcpyHostToDevice));
        *dst = (const double*)d;
        return CMOOP_OK;
    };
    for (int i = 0; i < n_models; ++i) {
        const cmoop_gp_model& m = models[i];
        const int n = m.n_train;
        h->max_n = n > h->max_n ? n : h->max_n;
        DevModel& dm = host[i];
        dm.n = n;
        dm.amplitude = m.amplitude;
        dm.inv_length = 1.0 / m.length_scale;
        dm.nu = m.nu;
        dm.noise = m.noise;
        dm.y_scale = m.y_scale;
        dm.y_shift = m.y_shift;
        int rc = upload(m.x_train, (size_t)n * dim, &dm.x_train);
        if (rc == CMOOP_OK) rc = upload(m.alpha, n, &dm.alpha);
        dm.chol_t = nullptr;
        if (rc == CMOOP_OK && m.chol_lower) {
            std::vector<double> t((size_t)n * n);
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) t[(size_t)c * n + r] = m.chol_lower[(size_t)r * n + c];
            rc = upload(t.data(), (size_t)n * n, &dm.chol_t);
        } else {
            h->has_chol = false;
        }
        if (rc != CMOOP_OK) {
            cmoop_gp_destroy(h);
            return rc;
        }
    }
    if (cudaMalloc((void**)&h->d_models, sizeof(DevModel) * n_models) != cudaSuccess ||
        cmoop::copy_sync(h->d_models, host.data(), sizeof(DevModel) * n_models, cudaMemcpyHostToDevice) != cudaSuccess) {
        cmoop::set_error("gp_create: device allocation failed");
        cmoop_gp_destroy(h);
        return CMOOP_ERR_CUDA;
    }
    *out = h;
    return CMOOP_OK;
}

int cmoop_gp_predict_dev(cmoop_gp_handle h, const double* xq, int q, double* mean, double* std, void* stream) {
    CMOOP_REQUIRE(h != nullptr, "gp_predict: null handle");
    CMOOP_REQUIRE(q >= 0, "gp_predict: negative q");
    if (q == 0) return CMOOP_OK;
    CMOOP_REQUIRE(xq && mean, "gp_predict: null pointer");
    if (std) CMOOP_REQUIRE(h->has_chol, "gp_predict: std requested but a model was created without chol_lower");
    dim3 grid((q + kWarpsPerBlock - 1) / kWarpsPerBlock, h->n_models);
    dim3 block(kWarpsPerBlock * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->max_n <= 128)
        gp_predict_kernel<4><<<grid, block, 0, st>>>(h->d_models, xq, q, h->dim, mean, std);
    else if (h->max_n <= 320)
        gp_predict_kernel<10><<<grid, block, 0, st>>>(h->d_models, xq, q, h->dim, mean, std);
    else
        gp_predict_kernel<32><<<grid, block, 0, st>>>(h->d_models, xq, q, h->dim, mean, std);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    return CMOOP_OK;
}

int cmoop_gp_predict_host(cmoop_gp_handle h, const double* xq, int q, double* mean, double* std) {
    CMOOP_REQUIRE(h != nullptr, "gp_predict: null handle");
    CMOOP_REQUIRE(q >= 0, "gp_predict: negative q");
    if (q == 0) return CMOOP_OK;
    CMOOP_REQUIRE(xq && mean, "gp_predict: null pointer");
    cudaStream_t st = cmoop::internal_stream();
    const size_t b_x = cmoop::align_up((size_t)q * h->dim * 8, 256);
    const size_t b_o = cmoop::align_up((size_t)q * h->n_models * 8, 256);
    char* d = (char*)cmoop::device_scratch(2, b_x + 2 * b_o);
    if (!d) return CMOOP_ERR_CUDA;
    double* d_x = (double*)d;
    double* d_mean = (double*)(d + b_x);
    double* d_std = std ? (double*)(d + b_x + b_o) : nullptr;
    CMOOP_CUDA_OK(cmoop::copy_async(d_x, xq, (size_t)q * h->dim * 8, cudaMemcpyHostToDevice, st));
    int rc = cmoop_gp_predict_dev(h, d_x, q, d_mean, d_std, st);
    if (rc != CMOOP_OK) return rc;
    CMOOP_CUDA_OK(cmoop::copy_async(mean, d_mean, (size_t)q * h->n_models * 8, cudaMemcpyDeviceToHost, st));
    if (std) CMOOP_CUDA_OK(cmoop::copy_async(std, d_std, (size_t)q * h->n_models * 8, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

}  // extern "C"
