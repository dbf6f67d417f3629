// Front-quality kernels: exact 2-D/3-D hypervolume, GD / IGD / Spread, non-dominated
// filter and C-metric.  Replaces the post-hoc metrics of compare.ipynb cell 0
// (sections 5-9: pg.hypervolume(...).compute(r), generational_distance, inverted_gd,
// spread_metric, coverage_metric, dominates_min) and the feasible Pareto filter of
// mobo_penalty.py:478-485.  All fp64, single CTA, latency-bound (fronts <= 4096 points).
//
// Hypervolume: points are filtered (strictly better than ref), ranked by (x,y) and by z
// with counting sorts; thread k owns slab [z_k, z_k+1) and sweeps the x-sorted list
// keeping the staircase of points already "switched on" (z-rank <= k); the slab
// volumes are summed in z order by one thread with separate multiply/add so the
// result is bit-identical to the sequential oracle (oracle/hv_ref.py).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kHvMaxN = 4096;

template <class Pred>
__device__ int compact1024(int n, Pred pred, int* out, int* s_scan) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int total = 0;
    for (int start = 0; start < n; start += kThreads) {
        const int i = start + threadIdx.x;
        const bool f = (i < n) && pred(i);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        const int pre = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_scan[warp] = __popc(b);
        __syncthreads();
        if (warp == 0) {
            int v = s_scan[lane], inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            s_scan[lane] = inc - v;
            if (lane == 31) s_scan[32] = inc;
        }
        __syncthreads();
        if (f) out[total + s_scan[warp] + pre] = i;
        total += s_scan[32];
        __syncthreads();
    }
    return total;
}

__global__ void __launch_bounds__(kThreads, 1) hv_kernel(const double* __restrict__ pts, int n, int m,
                                                         const double* __restrict__ ref, double* __restrict__ out) {
    extern __shared__ __align__(16) char smem[];
    __shared__ int s_scan[34];
    double* xs = (double*)smem;          // x-sorted coordinates
    double* ys = xs + n;
    double* zs = ys + n;
    double* zsorted = zs + n;            // z in ascending order
    double* area = zsorted + n;          // slab areas by z-rank
    int* list = (int*)(area + n);        // filtered original indices
    int* zrank = list + n;               // z-rank of the x-sorted point
    const int tid = threadIdx.x;
    const double rx = ref[0], ry = ref[1], rz = (m == 3) ? ref[2] : 1.0;

    const int nv = compact1024(
        n,
        [&](int i) {
            bool ok = pts[(size_t)i * m] < rx && pts[(size_t)i * m + 1] < ry;
            if (m == 3) ok = ok && pts[(size_t)i * m + 2] < rz;
            return ok;
        },
        list, s_scan);
    if (nv == 0) {
        if (tid == 0) out[0] = 0.0;
        return;
    }
    // rank by (x, y, original order) -> x-sorted arrays
    for (int a = tid; a < nv; a += kThreads) {
        const int i = list[a];
        const double x = pts[(size_t)i * m], y = pts[(size_t)i * m + 1];
        int r = 0;
        for (int b = 0; b < nv; ++b) {
            const int j = list[b];
            const double xb = pts[(size_t)j * m], yb = pts[(size_t)j * m + 1];
            r += (xb < x || (xb == x && (yb < y || (yb == y && b < a)))) ? 1 : 0;
        }
        xs[r] = x;
        ys[r] = y;
        zs[r] = (m == 3) ? pts[(size_t)i * m + 2] : 0.0;
    }
    __syncthreads();
    // z-rank of every x-sorted point (ties by x-order; tie order does not change the volume)
    for (int a = tid; a < nv; a += kThreads) {
        const double z = zs[a];
        int r = 0;
        for (int b = 0; b < nv; ++b) r += (zs[b] < z || (zs[b] == z && b < a)) ? 1 : 0;
        zrank[a] = r;
        zsorted[r] = z;
    }
    __syncthreads();
    for (int k = tid; k < nv; k += kThreads) {
        const double z_lo = zsorted[k];
        const double z_hi = (k + 1 < nv) ? zsorted[k + 1] : rz;
        double acc = 0.0;
        if (z_hi > z_lo) {
            double best = ry;
            for (int i = 0; i < nv; ++i) {
                if (zrank[i] <= k) {
                    const double y = ys[i];
                    if (y < best) {
                        acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(rx, xs[i]), __dsub_rn(best, y)));
                        best = y;
                    }
                }
            }
        }
        area[k] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        double vol = 0.0;
        for (int k = 0; k < nv; ++k) {
            const double z_lo = zsorted[k];
            const double z_hi = (k + 1 < nv) ? zsorted[k + 1] : rz;
            if (z_hi > z_lo) vol = __dadd_rn(vol, __dmul_rn(area[k], __dsub_rn(z_hi, z_lo)));
        }
        out[0] = vol;
    }
}

__device__ __forceinline__ double dist_rn(const double* a, const double* b, int m) {
    double s = 0.0;
    for (int k = 0; k < m; ++k) {
        const double d = __dsub_rn(a[k], b[k]);
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    return sqrt(s);
}

// out3 = {GD, IGD, Spread}; dmin_f [nf], dmin_t [nt] scratch in global memory
__global__ void __launch_bounds__(kThreads, 1) front_metrics_kernel(const double* __restrict__ front, int nf,
                                                                    const double* __restrict__ truef, int nt, int m,
                                                                    double* __restrict__ dmin_f,
                                                                    double* __restrict__ dmin_t,
                                                                    double* __restrict__ out3) {
    __shared__ double s_lo[CMOOP_NDS_MAX_M], s_hi[CMOOP_NDS_MAX_M];
    __shared__ double s_df[kThreads / 32], s_dl[kThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (tid < m) {
        double lo = inf, hi = -inf;
        for (int j = 0; j < nt; ++j) {
            const double v = truef[(size_t)j * m + tid];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        s_lo[tid] = lo;
        s_hi[tid] = hi;
    }
    __syncthreads();
    double df = inf, dl = inf;
    for (int i = tid; i < nf; i += kThreads) {
        double best = inf;
        for (int j = 0; j < nt; ++j) {
            const double d = dist_rn(front + (size_t)i * m, truef + (size_t)j * m, m);
            best = d < best ? d : best;
        }
        dmin_f[i] = best;
        const double a = dist_rn(front + (size_t)i * m, s_lo, m);
        const double b = dist_rn(front + (size_t)i * m, s_hi, m);
        df = a < df ? a : df;
        dl = b < dl ? b : dl;
    }
    for (int j = tid; j < nt; j += kThreads) {
        double best = inf;
        for (int i = 0; i < nf; ++i) {
            const double d = dist_rn(truef + (size_t)j * m, front + (size_t)i * m, m);
            best = d < best ? d : best;
        }
        dmin_t[j] = best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a = __shfl_xor_sync(0xffffffffu, df, o), b = __shfl_xor_sync(0xffffffffu, dl, o);
        df = a < df ? a : df;
        dl = b < dl ? b : dl;
    }
    if (lane == 0) {
        s_df[warp] = df;
        s_dl[warp] = dl;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kThreads / 32; ++w) {
            df = s_df[w] < df ? s_df[w] : df;
            dl = s_dl[w] < dl ? s_dl[w] : dl;
        }
        double sq = 0.0, sum = 0.0;
        for (int i = 0; i < nf; ++i) {
            sq += dmin_f[i] * dmin_f[i];
            sum += dmin_f[i];
        }
        out3[0] = sqrt(sq / nf);
        double sqt = 0.0;
        for (int j = 0; j < nt; ++j) sqt += dmin_t[j] * dmin_t[j];
        out3[1] = sqrt(sqt / nt);
        if (nf < 2) {
            out3[2] = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const double mean = sum / nf;
            double dev = 0.0;
            for (int i = 0; i < nf; ++i) dev += fabs(dmin_f[i] - mean);
            const double num = df + dl + dev;
            const double den = df + dl + (nf - 1) * mean;
            out3[2] = den != 0.0 ? num / den : __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}

__device__ __forceinline__ bool dominates_min(const double* a, const double* b, int m) {
    bool le = true, lt = false;
    for (int k = 0; k < m; ++k) {
        le = le && (a[k] <= b[k]);
        lt = lt || (a[k] < b[k]);
    }
    return le && lt;
}

// covered[i] = 1 if some point of A dominates B[i]  (skip_self: A and B are the same array)
__global__ void dominated_kernel(const double* __restrict__ a, int na, const double* __restrict__ b, int nb, int m,
                                 int skip_self, uint8_t* __restrict__ covered) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    bool hit = false;
    for (int j = 0; j < na && !hit; ++j) {
        if (skip_self && j == i) continue;
        hit = dominates_min(a + (size_t)j * m, b + (size_t)i * m, m);
    }
    covered[i] = hit ? 1 : 0;
}

}  // namespace

extern "C" {

size_t cmoop_hypervolume_workspace_bytes(int) { return 0; }

int cmoop_hypervolume_dev(const double* points, int n, int m, const double* ref, double* out, void*, size_t,
                          void* stream) {
    CMOOP_REQUIRE(m == 2 || m == 3, "hypervolume: m must be 2 or 3 (got %d)", m);
    CMOOP_REQUIRE(n >= 0 && n <= kHvMaxN, "hypervolume: n=%d outside [0,%d]", n, kHvMaxN);
    CMOOP_REQUIRE(out && ref, "hypervolume: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        CMOOP_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double), st));
        return CMOOP_OK;
    }
    CMOOP_REQUIRE(points != nullptr, "hypervolume: null points");
    const size_t smem = (size_t)n * (5 * 8 + 2 * 4);
    static bool configured = false;
    if (!configured) {
        CMOOP_CUDA_OK(cudaFuncSetAttribute(hv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    hv_kernel<<<1, kThreads, smem, st>>>(points, n, m, ref, out);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    return CMOOP_OK;
}

int cmoop_hypervolume_host(const double* points, int n, int m, const double* ref, double* out) {
    CMOOP_REQUIRE(m == 2 || m == 3, "hypervolume: m must be 2 or 3 (got %d)", m);
    CMOOP_REQUIRE(n >= 0 && n <= kHvMaxN, "hypervolume: n=%d outside [0,%d]", n, kHvMaxN);
    CMOOP_REQUIRE(out && ref, "hypervolume: null pointer");
    if (n == 0) {
        *out = 0.0;
        return CMOOP_OK;
    }
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const size_t b_p = cmoop::align_up((size_t)n * m * 8, 256);
    char* d = (char*)cmoop::device_scratch(3, b_p + 512);
    if (!d) return CMOOP_ERR_CUDA;
    CMOOP_CUDA_OK(cmoop::copy_async(d, points, (size_t)n * m * 8, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d + b_p, ref, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    int rc = cmoop_hypervolume_dev((const double*)d, n, m, (const double*)(d + b_p), (double*)(d + b_p + 256), nullptr,
                                   0, st);
    if (rc != CMOOP_OK) return rc;
    CMOOP_CUDA_OK(cmoop::copy_async(out, d + b_p + 256, 8, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

int cmoop_front_metrics_host(const double* front, int nf, const double* true_front, int nt, int m, double* out3) {
    CMOOP_REQUIRE(front && true_front && out3, "front_metrics: null pointer");
    CMOOP_REQUIRE(nf > 0 && nt > 0, "front_metrics: empty front");
    CMOOP_REQUIRE(m >= 1 && m <= CMOOP_NDS_MAX_M, "front_metrics: m=%d outside [1,%d]", m, CMOOP_NDS_MAX_M);
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const size_t b_f = cmoop::align_up((size_t)nf * m * 8, 256), b_t = cmoop::align_up((size_t)nt * m * 8, 256);
    const size_t b_df = cmoop::align_up((size_t)nf * 8, 256), b_dt = cmoop::align_up((size_t)nt * 8, 256);
    char* d = (char*)cmoop::device_scratch(3, b_f + b_t + b_df + b_dt + 256);
    if (!d) return CMOOP_ERR_CUDA;
    CMOOP_CUDA_OK(cmoop::copy_async(d, front, (size_t)nf * m * 8, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d + b_f, true_front, (size_t)nt * m * 8, cudaMemcpyHostToDevice, st));
    double* d_out = (double*)(d + b_f + b_t + b_df + b_dt);
    front_metrics_kernel<<<1, kThreads, 0, st>>>((const double*)d, nf, (const double*)(d + b_f), nt, m,
                                                 (double*)(d + b_f + b_t), (double*)(d + b_f + b_t + b_df), d_out);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    CMOOP_CUDA_OK(cmoop::copy_async(out3, d_out, 24, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

static int dominated_host(const double* a, int na, const double* b, int nb, int m, int skip_self, uint8_t* covered) {
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const size_t b_a = cmoop::align_up((size_t)na * m * 8, 256), b_b = cmoop::align_up((size_t)nb * m * 8, 256);
    char* d = (char*)cmoop::device_scratch(3, b_a + b_b + cmoop::align_up(nb, 256));
    if (!d) return CMOOP_ERR_CUDA;
    CMOOP_CUDA_OK(cmoop::copy_async(d, a, (size_t)na * m * 8, cudaMemcpyHostToDevice, st));
    const double* d_b = (const double*)d;
    if (!skip_self) {
        CMOOP_CUDA_OK(cmoop::copy_async(d + b_a, b, (size_t)nb * m * 8, cudaMemcpyHostToDevice, st));
        d_b = (const double*)(d + b_a);
    }
    uint8_t* d_c = (uint8_t*)(d + b_a + b_b);
    dominated_kernel<<<(nb + 255) / 256, 256, 0, st>>>((const double*)d, na, d_b, nb, m, skip_self, d_c);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    CMOOP_CUDA_OK(cmoop::copy_async(covered, d_c, nb, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

int cmoop_nondominated_mask_host(const double* points, int n, int m, uint8_t* mask) {
    CMOOP_REQUIRE(n >= 0 && m >= 1 && m <= CMOOP_NDS_MAX_M, "nondominated_mask: bad shape");
    if (n == 0) return CMOOP_OK;
    CMOOP_REQUIRE(points && mask, "nondominated_mask: null pointer");
    int rc = dominated_host(points, n, points, n, m, 1, mask);
    if (rc != CMOOP_OK) return rc;
    for (int i = 0; i < n; ++i) mask[i] = mask[i] ? 0 : 1;
    return CMOOP_OK;
}

int cmoop_coverage_host(const double* a, int na, const double* b, int nb, int m, double* out) {
    CMOOP_REQUIRE(out != nullptr && na >= 0 && nb >= 0 && m >= 1 && m <= CMOOP_NDS_MAX_M, "coverage: bad arguments");
    if (nb == 0 || na == 0) {
        *out = 0.0;
        return CMOOP_OK;
    }
    CMOOP_REQUIRE(a && b, "coverage: null pointer");
    uint8_t* covered = (uint8_t*)cmoop::pinned_scratch(3, nb);
    if (!covered) return CMOOP_ERR_CUDA;
    int rc = dominated_host(a, na, b, nb, m, 0, covered);
    if (rc != CMOOP_OK) return rc;
    int hit = 0;
    for (int i = 0; i < nb; ++i) hit += covered[i];
    *out = (double)hit / (double)nb;
    return CMOOP_OK;
}

}  // extern "C"
