// Penalised non-dominated sort + crowding distance, one CTA per problem.
//
// Replaces dominates / fast_non_dominated_sort / crowding_distance of the reference
// (nsga_penalty.py:448-524, sa_nsga_penalty.py:382-442, ablation_study/sa_nsga_local.py:240-277).
//
// Design (latency-bound, N <= 8192, M <= 8):
//   A  P[i][k] = f + lam*CV with __dmul_rn/__dadd_rn (no FMA: a fused multiply-add
//      changes the last bit of ~21% of samples and with it the ranks).
//   B  dominance bit-matrix: warp per row p, lane per column q, one ballot per
//      32 columns; the transposed ballot gives the dominated-by count n[p].
//   C  front 0 = {n[p]==0} by an ordered block compaction (ascending index).
//   D  peeling: for every unranked q count its dominators inside the current front
//      and remember the position of the LAST one; q enters the next front when its
//      count reaches 0, and the reference's discovery order is exactly the order of
//      key = last_pos * n + q  (walk p in front order, q ascending inside S[p]).
//   E  crowding per front on RAW objectives: stable rank-by-counting sort per
//      objective, +inf at the ends, interior += (next-prev)/(max-min) in objective
//      order with IEEE sub/div/add -- bit-identical to the Python floats.
// Everything lives in shared memory for n <= 1024 (<= 216 KiB); larger problems
// use the caller's global workspace with the same code.
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kSmallFront = 64;

struct Arena {
    size_t pen, dist, sortv, wsort, wq, cnt, rank, order, key, tmp, foff, dom, total;
    int words;
};

__host__ __device__ inline Arena make_arena(int n, int m) {
    Arena a;
    a.words = (n + 31) / 32;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t at = off;
        off += (bytes + 15) / 16 * 16;
        return at;
    };
    a.pen = take((size_t)n * m * 8);   // penalised objectives, later re-used for raw objectives
    a.dist = take((size_t)n * 8);
    a.sortv = take((size_t)n * 8);
    a.wsort = take((size_t)kWarps * kSmallFront * 8);
    a.wq = take((size_t)kWarps * kSmallFront * 4);
    a.cnt = take((size_t)n * 4);
    a.rank = take((size_t)n * 4);
    a.order = take((size_t)n * 4);
    a.key = take((size_t)n * 4);
    a.tmp = take((size_t)n * 4);
    a.foff = take((size_t)(n + 1) * 4);
    a.dom = take((size_t)n * a.words * 4);
    a.total = off;
    return a;
}

struct Params {
    const double* objs;
    const double* cv;
    int n, m;
    double lam, eps;
    int crowd_mode;
    int* rank;
    int* order;
    int* front_offsets;
    int* n_fronts;
    double* crowd;
    char* workspace;      // null -> arena in dynamic shared memory
    size_t ws_stride;
    const int* given_front;  // non-null: skip the sort, treat this index list as one front
    int given_len;
};

// Ordered compaction of {i in [0,n) : pred(i)} into out[], ascending i. Block-uniform call.
template <class Pred>
__device__ int block_compact(int n, Pred pred, int* out, int* s_scan /*[34]*/) {
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
            int v = s_scan[lane];
            int inc = v;
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

__device__ __forceinline__ bool range_applies(double span, double eps, int mode) {
    return mode == 0 ? (span > eps) : !(span < eps);
}

__global__ void __launch_bounds__(kThreads, 1) nds_crowding_kernel(Params p) {
    extern __shared__ __align__(16) char smem_raw[];
    __shared__ int s_scan[34];
    const int n = p.n, m = p.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Arena A = make_arena(n, m);
    char* base = p.workspace ? p.workspace + (size_t)blockIdx.x * p.ws_stride : smem_raw;
    double* pen = (double*)(base + A.pen);
    double* dist = (double*)(base + A.dist);
    double* sortv = (double*)(base + A.sortv);
    double* wsort = (double*)(base + A.wsort) + warp * kSmallFront;
    int* wq = (int*)(base + A.wq) + warp * kSmallFront;
    int* cnt = (int*)(base + A.cnt);
    int* rnk = (int*)(base + A.rank);
    int* order = (int*)(base + A.order);
    int* key = (int*)(base + A.key);
    int* tmp = (int*)(base + A.tmp);
    int* foff = (int*)(base + A.foff);
    unsigned* dom = (unsigned*)(base + A.dom);
    const int words = A.words;

    const double* objs = p.objs + (size_t)blockIdx.x * n * m;
    const double* cv = p.cv ? p.cv + (size_t)blockIdx.x * n : nullptr;
    int nf = 0;

    if (p.given_front == nullptr) {
        // ---- A: penalised objectives
        for (int i = tid; i < n * m; i += kThreads) {
            const int row = i / m;
            const double pen_term = cv ? __dmul_rn(p.lam, cv[row]) : 0.0;
            pen[i] = __dadd_rn(objs[i], pen_term);
        }
        __syncthreads();
        // ---- B: dominance bit matrix + dominated-by counts
        for (int r = warp; r < n; r += kWarps) {
            int c = 0;
            for (int w = 0; w < words; ++w) {
                const int q = w * 32 + lane;
                bool r_le = true, r_lt = false, q_le = true, q_lt = false;
                if (q < n) {
                    for (int k = 0; k < m; ++k) {
                        const double a = pen[r * m + k], b = pen[q * m + k];
                        r_le = r_le && !(a > b);
                        r_lt = r_lt || (a < b);
                        q_le = q_le && !(b > a);
                        q_lt = q_lt || (b < a);
                    }
                } else {
                    r_le = false;
                    q_le = false;
                }
                const unsigned fwd = __ballot_sync(0xffffffffu, r_le && r_lt);   // r dominates q
                const unsigned bwd = __ballot_sync(0xffffffffu, q_le && q_lt);   // q dominates r
                if (lane == 0) dom[(size_t)r * words + w] = fwd;
                c += __popc(bwd);
            }
            if (lane == 0) cnt[r] = c;
        }
        for (int i = tid; i < n; i += kThreads) rnk[i] = -1;
        __syncthreads();
        // ---- C: first front
        int F = block_compact(n, [&](int i) { return cnt[i] == 0; }, order, s_scan);
        for (int a = tid; a < F; a += kThreads) rnk[order[a]] = 0;
        if (tid == 0) foff[0] = 0;
        __syncthreads();
        // ---- D: peel
        int off = 0;
        while (F > 0) {
            if (tid == 0) foff[nf + 1] = off + F;
            for (int q = tid; q < n; q += kThreads) {
                int kq = -1;
                if (rnk[q] < 0) {
                    int c = 0, last = -1;
                    const int wq_ = q >> 5;
                    const unsigned bit = 1u << (q & 31);
                    for (int t = 0; t < F; ++t) {
                        const int pr = order[off + t];
                        if (dom[(size_t)pr * words + wq_] & bit) {
                            ++c;
                            last = t;
                        }
                    }
                    if (c > 0) {
                        const int left = cnt[q] - c;
                        cnt[q] = left;
                        if (left == 0) kq = last * n + q;
                    }
                }
                key[q] = kq;
            }
            __syncthreads();
            const int Fn = block_compact(n, [&](int i) { return key[i] >= 0; }, tmp, s_scan);
            for (int a = tid; a < Fn; a += kThreads) {
                const int q = tmp[a];
                const int kq = key[q];
                int r = 0;
                for (int b = 0; b < Fn; ++b) r += (key[tmp[b]] < kq) ? 1 : 0;
                order[off + F + r] = q;
                rnk[q] = nf + 1;
            }
            __syncthreads();
            off += F;
            F = Fn;
            ++nf;
        }
        for (int i = nf + 1 + tid; i <= n; i += kThreads) foff[i] = n;
        __syncthreads();
    } else {
        nf = p.given_len > 0 ? 1 : 0;
        for (int a = tid; a < p.given_len; a += kThreads) order[a] = p.given_front[a];
        if (tid == 0) {
            foff[0] = 0;
            foff[1] = p.given_len;
        }
        __syncthreads();
    }

    // ---- E: crowding on raw objectives (pen region now holds the raw copy)
    double* raw = pen;
    for (int i = tid; i < n * m; i += kThreads) raw[i] = objs[i];
    for (int i = tid; i < n; i += kThreads) dist[i] = 0.0;
    __syncthreads();
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    // small fronts: one warp per front
    for (int f = warp; f < nf; f += kWarps) {
        const int o = foff[f], F = foff[f + 1] - o;
        if (F > kSmallFront) continue;
        for (int k = 0; k < m; ++k) {
            for (int a = lane; a < F; a += 32) {
                const int q = order[o + a];
                const double v = raw[q * m + k];
                int r = 0;
                for (int b = 0; b < F; ++b) {
                    const double vb = raw[order[o + b] * m + k];
                    r += (vb < v || (vb == v && b < a)) ? 1 : 0;
                }
                wsort[r] = v;
                wq[r] = q;
            }
            __syncwarp();
            const double span = __dsub_rn(wsort[F - 1], wsort[0]);
            const bool apply = range_applies(span, p.eps, p.crowd_mode);
            for (int r = lane; r < F; r += 32) {
                const int q = wq[r];
                if (r == 0 || r == F - 1) {
                    dist[q] = inf;
                } else if (apply) {
                    dist[q] = __dadd_rn(dist[q], __ddiv_rn(__dsub_rn(wsort[r + 1], wsort[r - 1]), span));
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // large fronts: whole block per front
    for (int f = 0; f < nf; ++f) {
        const int o = foff[f], F = foff[f + 1] - o;
        if (F <= kSmallFront) continue;
        for (int k = 0; k < m; ++k) {
            for (int a = tid; a < F; a += kThreads) {
                const int q = order[o + a];
                const double v = raw[q * m + k];
                int r = 0;
                for (int b = 0; b < F; ++b) {
                    const double vb = raw[order[o + b] * m + k];
                    r += (vb < v || (vb == v && b < a)) ? 1 : 0;
                }
                sortv[r] = v;
                tmp[r] = q;
            }
            __syncthreads();
            const double span = __dsub_rn(sortv[F - 1], sortv[0]);
            const bool apply = range_applies(span, p.eps, p.crowd_mode);
            for (int r = tid; r < F; r += kThreads) {
                const int q = tmp[r];
                if (r == 0 || r == F - 1) {
                    dist[q] = inf;
                } else if (apply) {
                    dist[q] = __dadd_rn(dist[q], __ddiv_rn(__dsub_rn(sortv[r + 1], sortv[r - 1]), span));
                }
            }
            __syncthreads();
        }
    }

    // ---- outputs
    const size_t ob = (size_t)blockIdx.x * n;
    if (p.given_front) {
        if (p.crowd)
            for (int a = tid; a < p.given_len; a += kThreads) p.crowd[a] = dist[order[a]];
        return;
    }
    for (int i = tid; i < n; i += kThreads) {
        if (p.rank) p.rank[ob + i] = rnk[i];
        if (p.order) p.order[ob + i] = order[i];
        if (p.crowd) p.crowd[ob + i] = dist[i];
    }
    if (p.front_offsets)
        for (int i = tid; i <= n; i += kThreads) p.front_offsets[(size_t)blockIdx.x * (n + 1) + i] = foff[i];
    if (p.n_fronts && tid == 0) p.n_fronts[blockIdx.x] = nf;
}

const size_t kSmemLimit = 226 * 1024;   // 227 KiB per-CTA limit minus the kernel's static shared memory

int launch(Params p, int batch, cudaStream_t stream) {
    const Arena A = make_arena(p.n, p.m);
    size_t smem = 0;
    if (A.total <= kSmemLimit) {
        smem = A.total;
        p.workspace = nullptr;
        p.ws_stride = 0;
    } else {
        CMOOP_REQUIRE(p.workspace != nullptr, "nds: n=%d needs a global workspace of %zu bytes", p.n,
                      cmoop::align_up(A.total, 256) * batch);
        p.ws_stride = cmoop::align_up(A.total, 256);
    }
    static size_t configured = 0;
    if (smem > configured) {
        CMOOP_CUDA_OK(cudaFuncSetAttribute(nds_crowding_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kSmemLimit));
        configured = kSmemLimit;
    }
    nds_crowding_kernel<<<batch, kThreads, smem, stream>>>(p);
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaGetLastError());
    return CMOOP_OK;
}

}  // namespace

extern "C" {

size_t cmoop_nds_workspace_bytes(int n, int m, int batch) {
    if (n <= 0 || m <= 0 || batch <= 0) return 0;
    const Arena A = make_arena(n, m);
    if (A.total <= kSmemLimit) return 0;
    return cmoop::align_up(A.total, 256) * (size_t)batch;
}

int cmoop_nds_crowding_dev(const double* objs, const double* cv, int n, int m, int batch, double lam, double eps,
                           int crowd_mode, int* rank, int* order, int* front_offsets, int* n_fronts,
                           double* crowd, void* workspace, size_t workspace_bytes, void* stream) {
    CMOOP_REQUIRE(n >= 0 && n <= CMOOP_NDS_MAX_N, "nds: n=%d outside [0,%d]", n, CMOOP_NDS_MAX_N);
    CMOOP_REQUIRE(m >= 1 && m <= CMOOP_NDS_MAX_M, "nds: m=%d outside [1,%d]", m, CMOOP_NDS_MAX_M);
    CMOOP_REQUIRE(batch >= 0, "nds: negative batch");
    CMOOP_REQUIRE(crowd_mode == 0 || crowd_mode == 1, "nds: crowd_mode must be 0 or 1");
    if (n == 0 || batch == 0) {
        if (n_fronts && batch > 0) CMOOP_CUDA_OK(cudaMemsetAsync(n_fronts, 0, sizeof(int) * batch, (cudaStream_t)stream));
        if (front_offsets && batch > 0)
            CMOOP_CUDA_OK(cudaMemsetAsync(front_offsets, 0, sizeof(int) * batch, (cudaStream_t)stream));
        return CMOOP_OK;
    }
    CMOOP_REQUIRE(objs != nullptr, "nds: objs is null");
    CMOOP_REQUIRE(workspace_bytes >= cmoop_nds_workspace_bytes(n, m, batch), "nds: workspace too small");
    Params p{};
    p.objs = objs;
    p.cv = cv;
    p.n = n;
    p.m = m;
    p.lam = lam;
    p.eps = eps;
    p.crowd_mode = crowd_mode;
    p.rank = rank;
    p.order = order;
    p.front_offsets = front_offsets;
    p.n_fronts = n_fronts;
    p.crowd = crowd;
    p.workspace = (char*)workspace;
    return launch(p, batch, (cudaStream_t)stream);
}

int cmoop_nds_crowding_host(const double* objs, const double* cv, int n, int m, int batch, double lam, double eps,
                            int crowd_mode, int* rank, int* order, int* front_offsets, int* n_fronts,
                            double* crowd) {
    CMOOP_REQUIRE(n >= 0 && n <= CMOOP_NDS_MAX_N, "nds: n=%d outside [0,%d]", n, CMOOP_NDS_MAX_N);
    CMOOP_REQUIRE(m >= 1 && m <= CMOOP_NDS_MAX_M, "nds: m=%d outside [1,%d]", m, CMOOP_NDS_MAX_M);
    CMOOP_REQUIRE(batch >= 0, "nds: negative batch");
    if (n == 0 || batch == 0) {
        for (int b = 0; b < batch; ++b) {
            if (n_fronts) n_fronts[b] = 0;
            if (front_offsets) front_offsets[b] = 0;
        }
        return CMOOP_OK;
    }
    CMOOP_REQUIRE(objs != nullptr, "nds: objs is null");
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const size_t nb = (size_t)n * batch;
    // device layout: objs | cv | crowd | rank | order | foff | nf
    const size_t b_objs = cmoop::align_up(nb * m * 8, 256), b_cv = cmoop::align_up(nb * 8, 256);
    const size_t b_crowd = b_cv, b_i = cmoop::align_up(nb * 4, 256);
    const size_t b_foff = cmoop::align_up((size_t)(n + 1) * batch * 4, 256), b_nf = cmoop::align_up((size_t)batch * 4, 256);
    const size_t total = b_objs + b_cv + b_crowd + 2 * b_i + b_foff + b_nf;
    char* d = (char*)cmoop::device_scratch(0, total);
    if (!d) return CMOOP_ERR_CUDA;
    double* d_objs = (double*)d;
    double* d_cv = (double*)(d + b_objs);
    double* d_crowd = (double*)(d + b_objs + b_cv);
    int* d_rank = (int*)(d + b_objs + b_cv + b_crowd);
    int* d_order = (int*)((char*)d_rank + b_i);
    int* d_foff = (int*)((char*)d_order + b_i);
    int* d_nf = (int*)((char*)d_foff + b_foff);
    const size_t ws_bytes = cmoop_nds_workspace_bytes(n, m, batch);
    void* ws = nullptr;
    if (ws_bytes) {
        ws = cmoop::device_scratch(1, ws_bytes);
        if (!ws) return CMOOP_ERR_CUDA;
    }
    CMOOP_CUDA_OK(cmoop::copy_async(d_objs, objs, nb * m * 8, cudaMemcpyHostToDevice, st));
    if (cv) CMOOP_CUDA_OK(cmoop::copy_async(d_cv, cv, nb * 8, cudaMemcpyHostToDevice, st));
    int rc = cmoop_nds_crowding_dev(d_objs, cv ? d_cv : nullptr, n, m, batch, lam, eps, crowd_mode, d_rank, d_order,
                                    d_foff, d_nf, d_crowd, ws, ws_bytes, st);
    if (rc != CMOOP_OK) return rc;
    if (rank) CMOOP_CUDA_OK(cmoop::copy_async(rank, d_rank, nb * 4, cudaMemcpyDeviceToHost, st));
    if (order) CMOOP_CUDA_OK(cmoop::copy_async(order, d_order, nb * 4, cudaMemcpyDeviceToHost, st));
    if (front_offsets)
        CMOOP_CUDA_OK(cmoop::copy_async(front_offsets, d_foff, (size_t)(n + 1) * batch * 4, cudaMemcpyDeviceToHost, st));
    if (n_fronts) CMOOP_CUDA_OK(cmoop::copy_async(n_fronts, d_nf, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    if (crowd) CMOOP_CUDA_OK(cmoop::copy_async(crowd, d_crowd, nb * 8, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

int cmoop_crowding_distance_host(const double* objs, int n, int m, const int* front, int front_len, double eps,
                                 int crowd_mode, double* out) {
    CMOOP_REQUIRE(n >= 0 && n <= CMOOP_NDS_MAX_N, "crowding: n=%d outside [0,%d]", n, CMOOP_NDS_MAX_N);
    CMOOP_REQUIRE(m >= 1 && m <= CMOOP_NDS_MAX_M, "crowding: m=%d outside [1,%d]", m, CMOOP_NDS_MAX_M);
    CMOOP_REQUIRE(front_len >= 0 && front_len <= n, "crowding: front_len=%d outside [0,n=%d]", front_len, n);
    CMOOP_REQUIRE(crowd_mode == 0 || crowd_mode == 1, "crowding: crowd_mode must be 0 or 1");
    if (front_len == 0) return CMOOP_OK;
    CMOOP_REQUIRE(objs && front && out, "crowding: null pointer");
    for (int i = 0; i < front_len; ++i)
        CMOOP_REQUIRE(front[i] >= 0 && front[i] < n, "crowding: front[%d]=%d outside [0,%d)", i, front[i], n);
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const size_t b_objs = cmoop::align_up((size_t)n * m * 8, 256), b_f = cmoop::align_up((size_t)front_len * 4, 256);
    const size_t b_out = cmoop::align_up((size_t)front_len * 8, 256);
    char* d = (char*)cmoop::device_scratch(0, b_objs + b_f + b_out);
    if (!d) return CMOOP_ERR_CUDA;
    const size_t ws_bytes = cmoop_nds_workspace_bytes(n, m, 1);
    void* ws = nullptr;
    if (ws_bytes) {
        ws = cmoop::device_scratch(1, ws_bytes);
        if (!ws) return CMOOP_ERR_CUDA;
    }
    CMOOP_CUDA_OK(cmoop::copy_async(d, objs, (size_t)n * m * 8, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d + b_objs, front, (size_t)front_len * 4, cudaMemcpyHostToDevice, st));
    Params p{};
    p.objs = (const double*)d;
    p.n = n;
    p.m = m;
    p.eps = eps;
    p.crowd_mode = crowd_mode;
    p.crowd = (double*)(d + b_objs + b_f);
    p.workspace = (char*)ws;
    p.given_front = (const int*)(d + b_objs);
    p.given_len = front_len;
    int rc = launch(p, 1, st);
    if (rc != CMOOP_OK) return rc;
    CMOOP_CUDA_OK(cmoop::copy_async(out, p.crowd, (size_t)front_len * 8, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    return CMOOP_OK;
}

}  // extern "C"
