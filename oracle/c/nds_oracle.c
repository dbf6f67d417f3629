/* Plain-C restatement of the reference's penalised fast non-dominated sort + crowding distance
 * (sa_nsga_penalty.py:382-442 / nsga_penalty.py:448-524).  TEST INFRASTRUCTURE ONLY: a fast CPU
 * checker/baseline for sizes where the pure-Python reference takes seconds (N=512: 1.05 s).
 * Compile with -ffp-contract=off so f + lam*CV rounds like CPython (multiply, then add). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int dominates(const double* p, int m, int a, int b) {
    int strictly = 0;
    for (int k = 0; k < m; ++k) {
        if (p[a * m + k] > p[b * m + k]) return 0;
        if (p[a * m + k] < p[b * m + k]) strictly = 1;
    }
    return strictly;
}

/* order[n]: indices front by front in the reference's discovery order; front_offsets[n+1]; returns n_fronts */
int cmoop_oracle_nds(const double* objs, const double* cv, int n, int m, double lam, int* rank, int* order,
                     int* front_offsets) {
    if (n <= 0) return 0;
    double* pen = malloc(sizeof(double) * n * m);
    for (int i = 0; i < n; ++i) {
        const double term = lam * cv[i];
        for (int k = 0; k < m; ++k) pen[i * m + k] = objs[i * m + k] + term;
    }
    unsigned char* dom = calloc((size_t)n * n, 1);
    int* cnt = calloc(n, sizeof(int));
    int filled = 0, nf = 0;
    for (int p = 0; p < n; ++p) {
        for (int q = 0; q < n; ++q) {
            if (p == q) continue;
            if (dominates(pen, m, p, q)) dom[(size_t)p * n + q] = 1;
            else if (dominates(pen, m, q, p)) cnt[p]++;
        }
    }
    front_offsets[0] = 0;
    for (int p = 0; p < n; ++p)
        if (cnt[p] == 0) { order[filled++] = p; rank[p] = 0; }
    int start = 0;
    while (filled > start) {
        front_offsets[++nf] = filled;
        const int end = filled;
        for (int a = start; a < end; ++a) {
            const int p = order[a];
            for (int q = 0; q < n; ++q)
                if (dom[(size_t)p * n + q] && --cnt[q] == 0) { order[filled++] = q; rank[q] = nf; }
        }
        start = end;
    }
    for (int i = nf + 1; i <= n; ++i) front_offsets[i] = n;
    free(pen); free(dom); free(cnt);
    return nf;
}

/* crowding distance of front[0..f) on raw objectives; mode 0: apply iff span > eps, mode 1: skip iff span < eps */
void cmoop_oracle_crowding(const double* objs, int m, const int* front, int f, double eps, int mode, double* out) {
    int* idx = malloc(sizeof(int) * (f > 0 ? f : 1));
    for (int i = 0; i < f; ++i) out[i] = 0.0;
    for (int k = 0; k < m; ++k) {
        for (int i = 0; i < f; ++i) idx[i] = i;
        for (int i = 1; i < f; ++i) {                      /* stable insertion sort on objective k */
            const int cur = idx[i];
            const double v = objs[front[cur] * m + k];
            int j = i - 1;
            while (j >= 0 && objs[front[idx[j]] * m + k] > v) { idx[j + 1] = idx[j]; --j; }
            idx[j + 1] = cur;
        }
        if (f == 0) break;
        out[idx[0]] = INFINITY;
        out[idx[f - 1]] = INFINITY;
        const double lo = objs[front[idx[0]] * m + k], hi = objs[front[idx[f - 1]] * m + k];
        const double span = hi - lo;
        const int apply = mode == 0 ? (span > eps) : !(span < eps);
        if (!apply) continue;
        for (int i = 1; i + 1 < f; ++i)
            out[idx[i]] += (objs[front[idx[i + 1]] * m + k] - objs[front[idx[i - 1]] * m + k]) / span;
    }
    free(idx);
}
