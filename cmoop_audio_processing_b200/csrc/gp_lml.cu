// Gaussian-process log-marginal likelihood and its gradient on the device, fp64, one CTA per (target, theta) problem.
//
// Replaces the objective that GaussianProcessRegressor.fit hands to L-BFGS-B in SurrogateManager.update
// (ablation_study/sa_nsga_local.py:195-210: 4 models x 11 starts, every generation) and train_gps
// (mobo_penalty.py:252-263): sklearn/gaussian_process/_gpr.py log_marginal_likelihood(theta, eval_gradient=True)
//     K = k_theta(X, X) + alpha I;   L = chol(K);   a = K^-1 y
//     lml = -1/2 y^T a - sum_i log L_ii - n/2 log(2 pi)
//     d lml / d theta_l = 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij / d theta_l
// for the two kernels the reference uses, theta being the log-parameters in scikit-learn's order:
//     kind 0:  C * Matern(l, nu) + WhiteKernel      theta = (log c, log l, log noise)
//     kind 1:  Matern(l, nu)                        theta = (log l)
//
// The gradient needs the whole inverse, so instead of Cholesky + triangular inverse + product (three dependent passes)
// the CTA runs the symmetric SWEEP operator over the lower triangle: step k takes pivot d_k = A_kk (the k-th Schur
// complement, = L_kk^2, so log det K = sum_k log d_k and a non-positive pivot is exactly scikit-learn's "not positive
// definite -> lml = -inf"), sets column k to A_ik / d_k, A_kk = -1 / d_k and subtracts A_ik A_jk / d_k everywhere else;
// after n steps A = -K^-1.  Blocked by 16 pivots: the block's 16 columns (all n rows, by symmetry) are swept in shared
// memory one pivot after the other, keeping each pivot column u_k as it was when used; the rest of the matrix then
// takes the 16 rank-1 updates at once, A_ij -= sum_k u_ik u_jk / d_k -- a lane owns one row (its 16 scaled values in
// registers) and walks 16 columns whose u rows are broadcast reads, 16 DFMA per 16-byte read-modify-write.  (A first
// blocked form multiplied by an explicit inverse of the 16 x 16 pivot block and lost cond(block) digits: 1e-3 relative
// at cond(K) = 3e9 against 1e-9 for this one, which performs the unblocked algorithm's operations.)  The matrix
// (n <= 512; 288 x 288 x 8 B = 663 KB in the reference) lives in an L2-resident scratch slice, column-major, so warps
// stream columns coalesced, and is passed over n / 16 times instead of n.  All reductions are fixed-order (results are
// run-to-run identical).  One problem per CTA: the 44 optimiser starts of a surrogate update are one launch on 44 SMs.
#include <math.h>

#include <mutex>
#include <vector>

#include "common.cuh"

struct cmoop_gp_lml {
    int n = 0, dim = 0, n_targets = 0, kind = 0, n_theta = 0, slots = 0, device = 0;
    double nu = 1.5, jitter = 1e-10;
    double* d_x = nullptr;        // [n][dim]
    double* d_y = nullptr;        // [n_targets][n]
    double* d_a = nullptr;        // [slots][n][n]
    double* d_io = nullptr;       // [slots][8]: theta[3], lml, grad[3]
    int* d_target = nullptr;      // [slots]
    std::vector<cudaStream_t> streams;
};

namespace {

constexpr int kLmlThreads = 512;
constexpr int kLmlWarps = kLmlThreads / 32;
constexpr int kPB = 16;                 // pivots per block
constexpr int kPS = 18;                 // panel row stride in doubles: 144 B keeps rows 16-byte aligned and spreads banks
constexpr int kLmlMaxN = 512;
constexpr int kInFlight = 8;            // matrix columns a lane has in flight in the rank-16 update

struct LmlParams {
    int n, dim, kind, n_theta;
    double nu, jitter;
    const double* x;
    const double* y;
    double* a;
    double* io;
    const int* target;
};

__device__ __forceinline__ void matern(double r, int nu2, double& k, double& dk_dlogl) {
    // r = |x - x'| / l ; nu2 = 2 nu in {1, 3, 5} (sklearn kernels.py:1713-1743 and its eval_gradient branch)
    if (nu2 == 3) {
        const double s = 1.7320508075688772 * r, e = exp(-s);
        k = (1.0 + s) * e;
        dk_dlogl = s * s * e;
    } else if (nu2 == 5) {
        const double s = 2.23606797749979 * r, e = exp(-s);
        k = (1.0 + s + s * s / 3.0) * e;
        dk_dlogl = s * s / 3.0 * (1.0 + s) * e;
    } else {
        const double e = exp(-r);
        k = e;
        dk_dlogl = r * e;
    }
}

// fixed-order block sum of up to 3 values per thread; result valid in every thread
__device__ __forceinline__ void block_sum3(double (&v)[3], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    __syncthreads();
    if (lane == 0) {
        red[warp * 3 + 0] = v[0];
        red[warp * 3 + 1] = v[1];
        red[warp * 3 + 2] = v[2];
    }
    __syncthreads();
    double s[3] = {0.0, 0.0, 0.0};
    for (int w = 0; w < kLmlWarps; ++w) {
        s[0] += red[w * 3 + 0];
        s[1] += red[w * 3 + 1];
        s[2] += red[w * 3 + 2];
    }
    v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
}

// shared memory, in doubles
__host__ __device__ inline size_t lml_panel_doubles(int n) {
    const size_t panels = 2 * (size_t)n * kPS;                 // P and Q
    const size_t partial = (size_t)kLmlWarps * n;              // K^-1 y partial sums (alias the panels)
    return panels > partial ? panels : partial;
}

// MR = ceil(n / 32): panel rows a thread keeps in registers during a block's sweep steps
template <int MR>
__global__ void __launch_bounds__(kLmlThreads) gp_lml_kernel(LmlParams P, int slot0) {
    extern __shared__ __align__(16) double sm[];
    const int n = P.n, dim = P.dim;
    const int slot = slot0 + blockIdx.x;
    double* pp = sm;                                 // [n][kPS]  S: the block's columns, swept in place
    double* qq = pp + (size_t)n * kPS;               // [n][kPS]  U: each column as it was when it became the pivot column
    double* xs = sm + lml_panel_doubles(n);          // [n][dim]
    double* ys = xs + (size_t)n * dim;               // [n]
    double* al = ys + n;                             // [n]
    double* mm = al + n;                             // [kPB]  1 / pivot (kPB * kPB reserved)
    double* red = mm + kPB * kPB;                    // [kLmlWarps * 3]
    __shared__ int s_fail, s_items;
    __shared__ unsigned short items[(kLmlMaxN / 32) * (kLmlMaxN / kPB + 2) / 2 + 16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* A = P.a + (size_t)slot * n * n;
    double* io = P.io + (size_t)slot * 8;
    const double* y = P.y + (size_t)P.target[slot] * n;

    double amp = 1.0, noise = 0.0, inv_l;
    if (P.kind == 0) {
        amp = exp(io[0]);
        inv_l = exp(-io[1]);
        noise = exp(io[2]);
    } else {
        inv_l = exp(-io[0]);
    }
    const int nu2 = (int)(2.0 * P.nu + 0.5);
    for (int e = tid; e < n * dim; e += kLmlThreads) xs[e] = P.x[e];
    for (int i = tid; i < n; i += kLmlThreads) ys[i] = y[i];
    // work items of the rank-16 update: (32-row chunk, 16-column group) pairs that touch the lower triangle
    const int row_chunks = (n + 31) / 32, col_groups = (n + kPB - 1) / kPB;
    if (tid == 0) {
        s_fail = 0;
        int cnt = 0;
        for (int rc = 0; rc < row_chunks; ++rc) {
            const int groups = 2 * rc + 2 < col_groups ? 2 * rc + 2 : col_groups;
            for (int cg = 0; cg < groups; ++cg) items[cnt++] = (unsigned short)(rc << 8 | cg);
        }
        s_items = cnt;
    }
    __syncthreads();
    const int n_items = s_items;

    // ---- K (lower triangle, column-major): warps over columns, lanes over rows
    for (int j = warp; j < n; j += kLmlWarps) {
        for (int i = j + lane; i < n; i += 32) {
            double ss = 0.0;
            for (int d = 0; d < dim; ++d) {
                const double df = xs[i * dim + d] - xs[j * dim + d];
                ss += df * df;
            }
            double k, dk;
            matern(sqrt(ss) * inv_l, nu2, k, dk);
            A[(size_t)j * n + i] = amp * k + (i == j ? noise + P.jitter : 0.0);
        }
    }
    __syncthreads();

    // ---- blocked sweep
    for (int k0 = 0; k0 < n; k0 += kPB) {
        const int bk = n - k0 < kPB ? n - k0 : kPB;
        // (1) panel S[i][c] = A(i, k0 + c) by symmetry (zero beyond bk); U[:, 0] = S[:, 0], other U columns cleared
        if (tid < kPB) mm[tid] = 0.0;
        for (int e = tid; e < k0 * kPB; e += kLmlThreads) {              // rows above the block: row k0 + c of column i
            const int i = e / kPB, c = e - i * kPB;
            const double v = c < bk ? A[(size_t)i * n + k0 + c] : 0.0;
            pp[i * kPS + c] = v;
            qq[i * kPS + c] = c == 0 ? v : 0.0;
        }
        const int below = n - k0;
        for (int e = tid; e < below * kPB; e += kLmlThreads) {            // rows from the block down: column k0 + c
            const int c = e / below, i = k0 + e - c * below;
            double v = 0.0;
            if (c < bk) v = i >= k0 + c ? A[(size_t)(k0 + c) * n + i] : A[(size_t)i * n + k0 + c];
            pp[i * kPS + c] = v;
            qq[i * kPS + c] = c == 0 ? v : 0.0;
        }
        __syncthreads();
        // (2) the block's 16 sweep steps on the panel alone, one after the other (an explicit D^-1 would lose
        // cond(D) digits); U[:, k] keeps column k as it was when it became the pivot column, mm[k] = 1 / d_k
        // A thread keeps its panel elements -- column pc, rows pi0 + 32 m -- in registers across the block's steps
        // and reads only the pivot column from shared memory; the owners of column k + 1 publish it for the next step.
        const int pc = tid & (kPB - 1), pi0 = tid >> 4;
        double sv[MR];
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            const int i = pi0 + 32 * m;
            sv[m] = i < n ? pp[i * kPS + pc] : 0.0;
        }
        bool bad = false;
        for (int k = 0; k < bk; ++k) {
            const int r = k0 + k;
            const double d = qq[r * kPS + k];
            if (!(d > 0.0) || !(d < 1.7e308)) {                           // same value in every thread: uniform exit
                bad = true;
                break;
            }
            const double inv = 1.0 / d;
            if (tid == 0) {
                al[r] = d;                                                // pivots; their logs are summed after the sweep
                mm[k] = inv;
            }
            if (pc < bk) {
                // column pc == k: v = u_i / d (pivot row: -1 / d); other columns: v = s - u_i t with t = u_(k0+pc) / d
                // (pivot row: t).  Written as one fused form v = a * u_i + b with per-thread a and the row-r exception.
                const double t = qq[(k0 + pc) * kPS + k] * inv;
                const bool own = pc == k;
                const double mul = own ? inv : -t;
                const double piv_val = own ? -inv : t;
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int i = pi0 + 32 * m;
                    const double col_i = qq[(i < n ? i : 0) * kPS + k];
                    double v = fma(mul, col_i, own ? 0.0 : sv[m]);
                    if (i == r) v = piv_val;
                    sv[m] = v;
                }
                if (pc == k + 1 || k == bk - 1) {                         // publish the next pivot column / the final panel
                    double* dst = (k == bk - 1 ? pp : qq) + pc;
#pragma unroll
                    for (int m = 0; m < MR; ++m) {
                        const int i = pi0 + 32 * m;
                        if (i < n) dst[i * kPS] = sv[m];
                    }
                }
            }
            __syncthreads();
        }
        if (bad) {
            if (tid == 0) s_fail = 1;
            break;
        }
        // (4) rank-16 update of the lower triangle outside the block; a lane owns a row, a warp item = 32 rows x 16 columns
        const int kb = k0 / kPB;
        for (int item = warp; item < n_items; item += kLmlWarps) {
            {
                const int rc = items[item] >> 8, cg = items[item] & 0xff;
                if (cg == kb) continue;
                const int i = rc * 32 + lane;
                const bool row_ok = i < n && (i < k0 || i >= k0 + kPB);
                double q[kPB];
                {
                    const double2* qrow = reinterpret_cast<const double2*>(qq + (i < n ? i : 0) * kPS);
#pragma unroll
                    for (int c = 0; c < kPB / 2; ++c) {
                        const double2 t = qrow[c];
                        q[2 * c] = t.x * mm[2 * c]; q[2 * c + 1] = t.y * mm[2 * c + 1];
                    }
                }
                const int j0 = cg * kPB;
#pragma unroll
                for (int jj = 0; jj < kPB; jj += kInFlight) {
                    double v[kInFlight];
                    bool ok[kInFlight];
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) {
                        const int j = j0 + jj + u;
                        ok[u] = row_ok && j <= i && j < n;
                        v[u] = ok[u] ? A[(size_t)j * n + i] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) {
                        const int j = j0 + jj + u;
                        const double2* prow = reinterpret_cast<const double2*>(qq + (j < n ? j : 0) * kPS);
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int c = 0; c < kPB / 2; ++c) {
                            const double2 t = prow[c];
                            s0 = fma(q[2 * c], t.x, s0);
                            s1 = fma(q[2 * c + 1], t.y, s1);
                        }
                        if (ok[u]) A[(size_t)j * n + i] = v[u] - (s0 + s1);
                    }
                }
            }
        }
        // (5) the block's own rows and columns take the panel's final values (disjoint from the elements (4) touches)
        for (int e = tid; e < k0 * kPB; e += kLmlThreads) {
            const int i = e / kPB, c = e - i * kPB;
            if (c < bk) A[(size_t)i * n + k0 + c] = pp[i * kPS + c];
        }
        for (int e = tid; e < below * kPB; e += kLmlThreads) {
            const int c = e / below, i = k0 + e - c * below;
            if (c < bk && i >= k0 + c) A[(size_t)(k0 + c) * n + i] = pp[i * kPS + c];
        }
        __syncthreads();
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) {
            io[3] = -INFINITY;
            for (int q = 0; q < P.n_theta; ++q) io[4 + q] = 0.0;
        }
        return;
    }

    // log det K = sum of the logs of the pivots (kept in al[] by the sweep), all threads in parallel
    double ld[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += kLmlThreads) ld[0] += log(al[i]);
    block_sum3(ld, red);
    const double logdet = ld[0];
    __syncthreads();

    // ---- a = K^-1 y with A = -K^-1 (lower triangle): one coalesced pass, warp per column j.  Element (i, j), i > j,
    // adds A_ij y_j to row i (per-warp partial rows in shared memory) and A_ij y_i to row j (warp reduction).
    double* part = sm;                               // [kLmlWarps][n], aliases the panels
    for (int e = tid; e < kLmlWarps * n; e += kLmlThreads) part[e] = 0.0;
    __syncthreads();
    for (int j = warp; j < n; j += kLmlWarps) {
        const double yj = ys[j];
        double tj = 0.0;
        for (int i = j + lane; i < n; i += 32) {
            const double v = A[(size_t)j * n + i];
            part[warp * n + i] += v * yj;            // a lane revisits row i only in later columns of this warp
            if (i > j) tj += v * ys[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tj += __shfl_xor_sync(0xffffffffu, tj, o);
        __syncwarp();
        if (lane == 0) part[warp * n + j] += tj;
        __syncwarp();
    }
    __syncthreads();
    for (int i = tid; i < n; i += kLmlThreads) {
        double s = 0.0;
        for (int w = 0; w < kLmlWarps; ++w) s += part[w * n + i];
        al[i] = -s;
    }
    __syncthreads();

    // ---- lml and gradient: acc[0..2] = sum_ij inner_ij dK_ij/dtheta over the full symmetric matrix
    double acc[3] = {0.0, 0.0, 0.0};
    double yta = 0.0;
    for (int i = tid; i < n; i += kLmlThreads) yta += ys[i] * al[i];
    for (int j = warp; j < n; j += kLmlWarps) {
        const double aj = al[j];
        for (int i = j + lane; i < n; i += 32) {
            const double inner = al[i] * aj + A[(size_t)j * n + i];
            if (i == j) {
                if (P.kind == 0) {
                    acc[0] += inner * amp;
                    acc[2] += inner * noise;
                }
            } else {
                double ss = 0.0;
                for (int d = 0; d < dim; ++d) {
                    const double df = xs[i * dim + d] - xs[j * dim + d];
                    ss += df * df;
                }
                double kk, dk;
                matern(sqrt(ss) * inv_l, nu2, kk, dk);
                if (P.kind == 0) {
                    acc[0] += 2.0 * inner * amp * kk;
                    acc[1] += 2.0 * inner * amp * dk;
                } else {
                    acc[0] += 2.0 * inner * dk;
                }
            }
        }
    }
    block_sum3(acc, red);
    double t3[3] = {yta, 0.0, 0.0};
    block_sum3(t3, red);
    if (tid == 0) {
        io[3] = -0.5 * t3[0] - 0.5 * logdet - 0.5 * (double)n * 1.8378770664093453;      // log(2 pi)
        for (int q = 0; q < P.n_theta; ++q) io[4 + q] = 0.5 * acc[q];
    }
}

size_t lml_smem(int n, int dim) {
    return (lml_panel_doubles(n) + (size_t)n * dim + 2 * (size_t)n + kPB * kPB + kLmlWarps * 3) * sizeof(double);
}

}  // namespace

extern "C" {

int cmoop_gp_lml_destroy(cmoop_gp_lml_handle h) {
    if (!h) return CMOOP_OK;
    for (cudaStream_t s : h->streams) cudaStreamDestroy(s);
    cudaFree(h->d_x);
    cudaFree(h->d_y);
    cudaFree(h->d_a);
    cudaFree(h->d_io);
    cudaFree(h->d_target);
    delete h;
    return CMOOP_OK;
}

int cmoop_gp_lml_create(const double* x, int n, int dim, const double* y, int n_targets, int kind, double nu,
                        double jitter, int slots, cmoop_gp_lml_handle* out) {
    CMOOP_REQUIRE(out != nullptr, "gp_lml_create: null out");
    *out = nullptr;
    CMOOP_REQUIRE(x && y, "gp_lml_create: null pointer");
    CMOOP_REQUIRE(n >= 1 && n <= kLmlMaxN, "gp_lml_create: n_train must be in [1, %d] (got %d)", kLmlMaxN, n);
    CMOOP_REQUIRE(dim >= 1 && n_targets >= 1 && slots >= 1, "gp_lml_create: dim, n_targets and slots must be positive");
    CMOOP_REQUIRE(kind == 0 || kind == 1, "gp_lml_create: kind must be 0 (C*Matern+White) or 1 (Matern)");
    CMOOP_REQUIRE(nu == 0.5 || nu == 1.5 || nu == 2.5, "gp_lml_create: nu must be 0.5, 1.5 or 2.5");
    CMOOP_REQUIRE(lml_smem(n, dim) <= 200 * 1024, "gp_lml_create: n_train x dim does not fit in shared memory");
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(gp_lml_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(gp_lml_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(gp_lml_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(gp_lml_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    auto* h = new cmoop_gp_lml();
    cudaGetDevice(&h->device);
    h->n = n; h->dim = dim; h->n_targets = n_targets; h->kind = kind; h->n_theta = kind == 0 ? 3 : 1;
    h->nu = nu; h->jitter = jitter; h->slots = slots;
    bool ok = cudaMalloc((void**)&h->d_x, (size_t)n * dim * 8) == cudaSuccess &&
              cudaMalloc((void**)&h->d_y, (size_t)n_targets * n * 8) == cudaSuccess &&
              cudaMalloc((void**)&h->d_a, (size_t)slots * n * n * 8) == cudaSuccess &&
              cudaMalloc((void**)&h->d_io, (size_t)slots * 8 * 8) == cudaSuccess &&
              cudaMalloc((void**)&h->d_target, (size_t)slots * 4) == cudaSuccess &&
              cmoop::copy_sync(h->d_x, x, (size_t)n * dim * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
              cmoop::copy_sync(h->d_y, y, (size_t)n_targets * n * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    for (int s = 0; ok && s < slots; ++s) {
        cudaStream_t st;
        ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        if (ok) h->streams.push_back(st);
    }
    if (!ok) {
        cmoop::set_error("gp_lml_create: device allocation failed (%s)", cudaGetErrorString(cudaGetLastError()));
        cmoop_gp_lml_destroy(h);
        return CMOOP_ERR_CUDA;
    }
    *out = h;
    return CMOOP_OK;
}

int cmoop_gp_lml_n_theta(cmoop_gp_lml_handle h) { return h ? h->n_theta : 0; }

int cmoop_gp_lml_eval(cmoop_gp_lml_handle h, int slot, int count, const double* theta, const int* target, double* lml,
                      double* grad) {
    CMOOP_REQUIRE(h != nullptr, "gp_lml_eval: null handle");
    CMOOP_REQUIRE(count >= 0 && slot >= 0 && slot + count <= h->slots, "gp_lml_eval: slots [%d, %d) outside [0, %d)", slot,
                  slot + count, h->slots);
    if (count == 0) return CMOOP_OK;
    CMOOP_REQUIRE(theta && target && lml && grad, "gp_lml_eval: null pointer");
    CMOOP_CUDA_OK(cudaSetDevice(h->device));            // callers are worker threads: the current device is per host thread
    cudaStream_t st = h->streams[slot];
    const int nt = h->n_theta;
    std::vector<double> io((size_t)count * 8, 0.0);
    for (int b = 0; b < count; ++b) {
        CMOOP_REQUIRE(target[b] >= 0 && target[b] < h->n_targets, "gp_lml_eval: target %d outside [0, %d)", target[b],
                      h->n_targets);
        for (int q = 0; q < nt; ++q) io[(size_t)b * 8 + q] = theta[(size_t)b * nt + q];
    }
    CMOOP_CUDA_OK(cmoop::copy_async(h->d_io + (size_t)slot * 8, io.data(), io.size() * 8, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(h->d_target + slot, target, (size_t)count * 4, cudaMemcpyHostToDevice, st));
    LmlParams P{h->n, h->dim, h->kind, nt, h->nu, h->jitter, h->d_x, h->d_y, h->d_a, h->d_io, h->d_target};
    const size_t smem = lml_smem(h->n, h->dim);
    if (h->n <= 64) gp_lml_kernel<2><<<count, kLmlThreads, smem, st>>>(P, slot);
    else if (h->n <= 160) gp_lml_kernel<5><<<count, kLmlThreads, smem, st>>>(P, slot);
    else if (h->n <= 288) gp_lml_kernel<9><<<count, kLmlThreads, smem, st>>>(P, slot);
    else gp_lml_kernel<16><<<count, kLmlThreads, smem, st>>>(P, slot);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    CMOOP_CUDA_OK(cmoop::copy_async(io.data(), h->d_io + (size_t)slot * 8, io.size() * 8, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    for (int b = 0; b < count; ++b) {
        lml[b] = io[(size_t)b * 8 + 3];
        for (int q = 0; q < nt; ++q) grad[(size_t)b * nt + q] = io[(size_t)b * 8 + 4 + q];
    }
    return CMOOP_OK;
}

}  // extern "C"
