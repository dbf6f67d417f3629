// Host executor of the population-batched candidate-CNN training/scoring path and its C ABI.
//
// Replaces the serial `for ind in population: evaluate_individual(ind)` loop of
// compute_objectives_and_constraints (nsga_penalty.py:418-442; sa_nsga_penalty.py:231-253) together with
// build_model / model.fit / model.evaluate / model.predict / calculate_fpr / compute_model_size_mb
// (nsga_penalty.py:225-395).  One process drives one GPU; candidates are packed into "waves" that fit
// the activation arena, and every kernel launch of a training / evaluation step covers the same stage
// of every still-active candidate of the wave (grouped launch, cnn.cuh).
//
// Canonical stage order (a valid topological order for every genotype of both variants):
//   0 stem1 | 1 stem2 (variant A) | per block b: 2+3b skip (1x1/s2), 3+3b conv1, 4+3b conv2 (variant A)
//   | GAP | 11..14 fc0..fc3 | 15 output layer | cross-entropy
// Backward walks the same list in reverse; gradients w.r.t. unit outputs live in buffer A, gradients
// w.r.t. conv outputs in buffer B, the residual-branch gradient in S (see run_backward).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <cuda.h>

#include "../common.cuh"
#include "cnn.cuh"

using namespace cmoop_cnn;

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library is linked against the static runtime
// only, so it still loads on a host without libcuda (tests/test_abi.py) and fails loudly at the first compute call.
int cmoop_cnn::Launch::make_row_tmap(void* out128, const void* dptr, int C, int W, int H, int N, int pad) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            cmoop::set_error("cuTensorMapEncodeTiled is not available from this driver");
            return -1;
        }
        encode = (encode_fn)fn;
    }
    static_assert(sizeof(CUtensorMap) == kTmapBytes, "CUtensorMap size");
    alignas(64) CUtensorMap tm;
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)(W + 2 * pad), 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dptr), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cmoop::set_error("cuTensorMapEncodeTiled(C=%d, W=%d, H=%d, N=%d, pad=%d) -> %d", C, W, H, N, pad, (int)r);
        return -1;
    }
    memcpy(out128, &tm, kTmapBytes);
    return 0;
}

// ---- optional per-kernel-family device timing (cmoop_profile_enable): one cudaEvent pair around every grouped launch of
// the engine's stream, resolved after the call.  Launches are serialised on one stream, so the pairs partition the device
// time of a call; bench.py uses the table for the `roofline.dominant_kernel` object and the step breakdown.  Off by default
// (two event records per launch perturb small launches by a few microseconds).
#include <map>
#include <string>
namespace {
struct ProfAgg { long long launches = 0; double ms = 0.0, flops = 0.0; };
struct ProfRec { const char* name; cudaEvent_t a, b; double flops; int stage; };
int g_prof_stage = -1;        // stage of the launches being recorded (CMOOP_PROF_STAGES=1 splits the table per stage)
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
std::map<std::string, ProfAgg> g_prof_table;
double g_last_device_ms = 0.0;

inline cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
inline void prof_begin(const char* name, double flops, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r{name, prof_event(), prof_event(), flops, g_prof_stage};
    cudaEventRecord(r.a, st);
    g_prof_recs.push_back(r);
}
inline void prof_end(cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEventRecord(g_prof_recs.back().b, st);
}
// CMOOP_CNN_SYNC=1: synchronise after every grouped launch so that an asynchronous fault is reported with the launch that
// caused it (development aid; compute-sanitizer is not available on the GPU pool)
inline bool debug_sync() {
    static const bool on = getenv("CMOOP_CNN_SYNC") != nullptr;
    return on;
}
// call after the stream has been synchronised
void prof_resolve() {
    for (ProfRec& r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            std::string key(r.name);
            const size_t cut = key.find('(');
            if (cut != std::string::npos) key.resize(cut);
            if (key.rfind("Launch::", 0) == 0) key = key.substr(8);
            static const bool per_stage = getenv("CMOOP_PROF_STAGES") != nullptr;
            if (per_stage && r.stage >= 0) key += "@" + std::to_string(r.stage);
            ProfAgg& a = g_prof_table[key];
            a.launches += 1;
            a.ms += ms;
            a.flops += r.flops;
        }
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    (void)cudaGetLastError();
    g_prof_recs.clear();
}
}  // namespace

struct cmoop_cnn_dataset {
    float* x_train = nullptr;
    int* y_train = nullptr;
    float* x_val = nullptr;
    int* y_val = nullptr;
    int n_train = 0, n_val = 0, H = 0, W = 0;
};

namespace {

constexpr int N_STAGES = 16, ST_FC0 = 11, ST_OUT = 15;
const int kFcUnits[5][4] = {{0, 0, 0, 0}, {64, 0, 0, 0}, {128, 64, 0, 0}, {256, 128, 64, 0}, {512, 256, 128, 64}};

struct Unit {
    int stage = 0, cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
    int H = 1, W = 1, Ho = 1, Wo = 1, Po = 1, Qo = 1;   // input grid, conv-output grid, unit-output grid
    bool dense = false, is_skip = false, has_bn = false, relu_epi = false, relu_mid = false, pool = false;
    bool add_skip = false, post_fwd = false, need_dgrad = true, use_dropout = false;
    int input = -1;        // producing unit, -1 = dataset, -2 = GAP output
    int skip_unit = -1;
    int fc_index = -1;
    long long w_off = 0, bn_off = -1;
    long long u_elems = 0, v_elems = 0;
    int stat_tiles = 0, bwd_rows = 0, bwd_pix = 128, wg_splits = 1, wg_chunk = 0;
    float* U = nullptr;
    float* V = nullptr;
    float* stat = nullptr;
    float* bn = nullptr;
    float* bwd_part = nullptr;
    float* wt = nullptr;
    uint8_t* idx = nullptr;
    bool tc = false;                 // convolution runs on the tcgen05 path (bf16 operands)
    bool stem = false;               // Cin = 1 first convolution: dedicated kernels (stem.cu)
    bool tc2 = false;                // tensor-core unit on the patch-resident kernel (conv_tc2.cu)
    bool wg2 = false;                // ... and its weight gradient on wgrad_tc2.cu
    int kpad_f = 0, kpad_d = 0;      // padded K of the forward / data-gradient GEMM
    bool need_vh = false;            // a tensor-core unit consumes this unit's output
    // precision bf16 stores a convolution unit's activations in bf16 ONLY: U / V stay null, Uh / Vh are the tensors
    // (Vh == Uh when the unit has no post stage); precision fp32 keeps U / V and Vh is an optional shadow of V
    __nv_bfloat16* Uh = nullptr;
    __nv_bfloat16* Vh = nullptr;     // bf16 shadow of V
    __nv_bfloat16* wtb = nullptr;    // bf16 [cout][kpad_f]
    __nv_bfloat16* wtd = nullptr;    // bf16 [cin][kpad_d]
    const char* tmaps = nullptr;     // device: 4 tensor maps {input Vh, dU (gBh), the same two for the tail batch} (tc2 units)
};

struct Cand {
    cmoop_genotype g{};
    uint64_t seed = 0;
    int index = 0;
    std::vector<Unit> units;
    int last_conv = 0, n_fc = 0;
    long long n_params = 0;
    size_t arena_bytes = 0;
    float *p = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr, *best = nullptr;
    float *gA = nullptr, *gB = nullptr, *gS = nullptr, *gG = nullptr, *gap = nullptr, *dlogits = nullptr, *wg_ws = nullptr;
    __nv_bfloat16 *gBh = nullptr, *gSh = nullptr;   // bf16 shadows of B / S for the tensor-core dgrad / wgrad
    int* perm = nullptr;
    double* acc = nullptr;     // [3][4]: train, val, predict accumulators
    int* pred = nullptr;
    int* confusion = nullptr;
    // early-stopping state (keras.callbacks.EarlyStopping semantics, see oracle/cnn_ref.py)
    bool active = true, has_best = false;
    int epochs_run = 0, wait = 0;
    double best_loss = INFINITY, last_val_loss = NAN, last_val_acc = NAN, final_acc = NAN, fpr = NAN;
};

long long conv_params(int cin, int cout, int k) { return (long long)k * k * cin * cout + cout; }

// Builds the unit list + flat parameter layout of one genotype (order == oracle/cnn_ref.layer_specs).
void build_units(Cand& c, const cmoop_cnn_config& cfg, int H, int W, int batch) {
    c.units.clear();
    const bool A = cfg.variant == 0;
    const bool bn = c.g.use_bn != 0;
    long long off = 0;
    int f = c.g.filters, k = c.g.kernel_size;
    int h = H, w = W;
    auto add_conv = [&](int stage, int cin, int cout, int ks, int stride, int hin, int win, bool with_bn) -> int {
        Unit u;
        u.stage = stage;
        u.cin = cin;
        u.cout = cout;
        u.k = ks;
        u.stride = stride;
        u.pad = stride == 1 ? (ks - 1) / 2 : 0;
        u.H = hin;
        u.W = win;
        u.Ho = stride == 1 ? hin : (hin + 1) / 2;
        u.Wo = stride == 1 ? win : (win + 1) / 2;
        u.Po = u.Ho;
        u.Qo = u.Wo;
        u.has_bn = with_bn;
        u.w_off = off;
        off += conv_params(cin, cout, ks);
        if (with_bn) {
            u.bn_off = off;
            off += 4LL * cout;
        }
        c.units.push_back(u);
        return (int)c.units.size() - 1;
    };
    auto set_pool = [&](Unit& u) {
        u.pool = true;
        u.Po = (u.Ho + 1) / 2;
        u.Qo = (u.Wo + 1) / 2;
    };
    int prev;
    if (A) {
        int s1 = add_conv(0, 1, f, k, 1, h, w, bn);
        c.units[s1].input = -1;
        c.units[s1].need_dgrad = false;
        c.units[s1].relu_epi = !bn;
        c.units[s1].relu_mid = bn;
        c.units[s1].post_fwd = bn;
        int s2 = add_conv(1, f, f, k, 1, h, w, bn);
        c.units[s2].input = s1;
        c.units[s2].relu_epi = !bn;
        c.units[s2].relu_mid = bn;
        set_pool(c.units[s2]);
        c.units[s2].post_fwd = true;
        prev = s2;
    } else {
        int s1 = add_conv(0, 1, f, k, 1, h, w, bn);
        c.units[s1].input = -1;
        c.units[s1].need_dgrad = false;
        c.units[s1].relu_epi = true;
        set_pool(c.units[s1]);
        c.units[s1].post_fwd = true;
        prev = s1;
    }
    h = c.units[prev].Po;
    w = c.units[prev].Qo;
    for (int b = 0; b < c.g.residual_blocks; ++b) {
        int sk = add_conv(2 + 3 * b, f, 2 * f, 1, 2, h, w, false);
        c.units[sk].input = prev;
        c.units[sk].is_skip = true;
        int last;
        if (A) {
            int c1 = add_conv(3 + 3 * b, f, 2 * f, k, 1, h, w, bn);
            c.units[c1].input = prev;
            c.units[c1].relu_epi = !bn;
            c.units[c1].relu_mid = bn;
            c.units[c1].post_fwd = bn;
            int c2 = add_conv(4 + 3 * b, 2 * f, 2 * f, k, 1, h, w, bn);
            c.units[c2].input = c1;
            set_pool(c.units[c2]);
            c.units[c2].add_skip = true;
            c.units[c2].skip_unit = sk;
            c.units[c2].post_fwd = true;
            last = c2;
        } else {
            int c1 = add_conv(3 + 3 * b, f, 2 * f, k, 1, h, w, bn);
            c.units[c1].input = prev;
            c.units[c1].relu_epi = true;
            set_pool(c.units[c1]);
            c.units[c1].add_skip = true;
            c.units[c1].skip_unit = sk;
            c.units[c1].post_fwd = true;
            last = c1;
        }
        prev = last;
        f *= 2;
        h = c.units[prev].Po;
        w = c.units[prev].Qo;
    }
    c.last_conv = prev;
    int width = f;
    c.n_fc = c.g.fc_layers;
    int in_unit = -2;
    for (int i = 0; i < c.n_fc; ++i) {
        const int units = kFcUnits[c.g.fc_layers][i];
        int d = add_conv(ST_FC0 + i, width, units, 1, 1, 1, 1, false);
        Unit& u = c.units[d];
        u.dense = true;
        u.relu_epi = true;
        u.input = in_unit;
        u.fc_index = i;
        u.use_dropout = c.g.use_dropout != 0;
        in_unit = d;
        width = units;
    }
    int o = add_conv(ST_OUT, width, cfg.n_classes, 1, 1, 1, 1, false);
    c.units[o].dense = true;
    c.units[o].input = in_unit;
    c.n_params = off;
    for (Unit& u : c.units) {
        u.tc = cfg.precision == 1 && !u.dense && u.cin % 16 == 0 && u.cout % 16 == 0;
        static const bool no_tc2 = getenv("CMOOP_CNN_NO_TC2") != nullptr;       // A/B switch: im2col-staging tcgen05 kernel
        u.tc2 = u.tc && !no_tc2 && Launch::tc2_ok(u.H, u.W, u.cin, u.cout, u.k, u.stride) &&
                Launch::tc2_ok(u.H, u.W, u.cout, u.cin, u.k, u.stride);
        u.kpad_f = (u.k * u.k * u.cin + 63) / 64 * 64;
        u.kpad_d = (u.k * u.k * u.cout + 63) / 64 * 64;
        u.u_elems = (long long)batch * u.Ho * u.Wo * u.cout;
        u.v_elems = (long long)batch * u.Po * u.Qo * u.cout;
        const long long M = (long long)batch * u.Ho * u.Wo;
        u.stat_tiles = (int)((M + 63) / 64);
        const long long npix = (long long)batch * u.Po * u.Qo;
        u.bwd_pix = Launch::bwd_pix(npix);                         // post_bwd_reduce: one partial row per CTA
        u.bwd_rows = (int)((npix + u.bwd_pix - 1) / u.bwd_pix);
        int splits = (int)std::min<long long>(32, std::max<long long>(1, M / 2048));
        int chunk = (int)((M + splits - 1) / splits);
        const int gran = u.tc ? 64 : 16;
        chunk = (chunk + gran - 1) / gran * gran;
        splits = (int)((M + chunk - 1) / chunk);
        static const bool no_stem = getenv("CMOOP_CNN_NO_STEM") != nullptr;      // A/B switch: generic SIMT kernels for the stem
        u.stem = !no_stem && !u.dense && u.input == -1 && Launch::stem_ok(u.H, u.W, u.cin, u.cout, u.k, u.stride, batch);
        if (u.stem) {
            chunk = kStemRows;
            splits = (int)((M + chunk - 1) / chunk);
        }
        // weight gradient of the patch-resident units on wgrad_tc2.cu (tiled-TMA operands, two taps per UMMA);
        // CMOOP_CNN_NO_WG2=1 is the A/B switch back to the im2col-staging kernel wgrad_tc_kernel
        static const bool no_wg2 = getenv("CMOOP_CNN_NO_WG2") != nullptr;
        u.wg2 = u.tc2 && !no_wg2 && Launch::wg2_ok(u.H, u.W, u.cin, u.cout, u.k, u.stride);
        if (u.wg2)
            Launch::wg2_splits((long long)batch * (u.H + 2 * u.pad) * (u.W + 2 * u.pad), Launch::wg2_items(u.cin, u.cout, u.k),
                               &splits, &chunk);
        u.wg_splits = splits;
        u.wg_chunk = chunk;
    }
    for (Unit& u : c.units)
        if (u.tc && u.input >= 0) c.units[u.input].need_vh = true;
}

struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    bool dry = false;
    void* take(size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        if (dry) return nullptr;
        return off <= cap ? base + at : nullptr;
    }
};

// Assigns (or, in a dry run, just sizes) all device buffers of one candidate.
void place(Cand& c, Arena& a, const cmoop_cnn_config& cfg, int n_train, int n_val, int batch) {
    const size_t f4 = sizeof(float);
    c.p = (float*)a.take(c.n_params * f4);
    c.grad = (float*)a.take(c.n_params * f4);
    c.m = (float*)a.take(c.n_params * f4);
    c.v = (float*)a.take(c.n_params * f4);
    c.best = cfg.restore_best_weights ? (float*)a.take(c.n_params * f4) : nullptr;
    long long maxV = (long long)batch * 1, maxU = 1, maxUh = 1, maxS = 1, maxWs = 1;
    for (Unit& u : c.units) {
        const bool half_only = cfg.precision == 1 && !u.dense;      // activations of convolution units live in bf16 only
        const bool separate_v = u.post_fwd || (u.dense && u.use_dropout);
        if (half_only) {
            u.U = u.V = nullptr;
            u.Uh = (__nv_bfloat16*)a.take(u.u_elems * 2);
            u.Vh = separate_v ? (__nv_bfloat16*)a.take(u.v_elems * 2) : u.Uh;
        } else {
            u.Uh = nullptr;
            u.U = (float*)a.take(u.u_elems * f4);
            u.V = separate_v ? (float*)a.take(u.v_elems * f4) : u.U;
            u.Vh = u.need_vh ? (__nv_bfloat16*)a.take(u.v_elems * 2) : nullptr;
        }
        u.idx = u.pool ? (uint8_t*)a.take(u.v_elems) : nullptr;
        if (u.has_bn) {
            u.stat = (float*)a.take((size_t)u.stat_tiles * 2 * u.cout * f4);
            u.bn = (float*)a.take((size_t)6 * u.cout * f4);
            u.bwd_part = (float*)a.take((size_t)u.bwd_rows * 2 * u.cout * f4);
        }
        u.wt = (u.need_dgrad && !u.tc) ? (float*)a.take((size_t)u.k * u.k * u.cin * u.cout * f4) : nullptr;
        if (u.tc) {
            const size_t wf = u.tc2 ? (size_t)Launch::tc2_weight_elems(u.cin, u.cout, u.k) : (size_t)u.cout * u.kpad_f;
            const size_t wd = u.tc2 ? (size_t)Launch::tc2_weight_elems(u.cout, u.cin, u.k) : (size_t)u.cin * u.kpad_d;
            u.wtb = (__nv_bfloat16*)a.take(wf * 2);
            u.wtd = u.need_dgrad ? (__nv_bfloat16*)a.take(wd * 2) : nullptr;
        }
        maxV = std::max(maxV, u.v_elems);
        maxV = std::max(maxV, (long long)batch * u.H * u.W * u.cin);
        if (!half_only) maxU = std::max(maxU, u.u_elems);
        else maxUh = std::max(maxUh, u.u_elems);
        if (u.is_skip) maxS = std::max(maxS, u.u_elems);
        if (u.wg_splits > 1)
            maxWs = std::max(maxWs, (long long)u.wg_splits * ((long long)u.k * u.k * u.cin + 1) * u.cout);
    }
    const Unit& lc = c.units[c.last_conv];
    c.gA = (float*)a.take(maxV * f4);
    c.gB = (float*)a.take(maxU * f4);                 // precision bf16: only the dense layers' pre-activation gradients
    c.gS = cfg.precision == 1 ? nullptr : (float*)a.take(maxS * f4);
    c.gBh = cfg.precision == 1 ? (__nv_bfloat16*)a.take(maxUh * 2) : nullptr;
    c.gSh = cfg.precision == 1 ? (__nv_bfloat16*)a.take(maxS * 2) : nullptr;
    c.gG = (float*)a.take((size_t)batch * lc.cout * f4);
    c.gap = (float*)a.take((size_t)batch * lc.cout * f4);
    c.dlogits = (float*)a.take((size_t)batch * cfg.n_classes * f4);
    c.wg_ws = (float*)a.take(maxWs * f4);
    c.perm = (int*)a.take((size_t)n_train * sizeof(int));
    c.acc = (double*)a.take(12 * sizeof(double));
    c.pred = (int*)a.take((size_t)n_val * sizeof(int));
    c.confusion = (int*)a.take((size_t)cfg.n_classes * cfg.n_classes * sizeof(int));
}

template <class T>
struct DevList {
    std::vector<T> h;
    T* d = nullptr;
    int total = 0;   // tiles / blocks of the grouped grid
    size_t blob_off = 0;
    // optional direct block -> task table of the grouped grid (hot kernels: one load instead of a binary search)
    std::vector<int> bt;
    int* d_bt = nullptr;
    size_t bt_off = 0;
};

// fills l.bt from the tasks' begin offsets (begin_of(task) ascending; the last task ends at l.total)
template <class T, class F>
void fill_block_table(DevList<T>& l, F begin_of) {
    l.bt.assign((size_t)l.total, 0);
    for (size_t i = 0; i < l.h.size(); ++i) {
        const int b0 = begin_of(l.h[i]), b1 = i + 1 < l.h.size() ? begin_of(l.h[i + 1]) : l.total;
        for (int b = b0; b < b1; ++b) l.bt[(size_t)b] = (int)i;
    }
}

// CMOOP_CNN_NO_SKIP_TC=1: A/B switch back to the tcgen05 kernel (conv_tc.cu) for the 1x1 / stride-2 projections
inline bool use_skip_tc() {
    static const bool off = getenv("CMOOP_CNN_NO_SKIP_TC") != nullptr;
    return !off;
}

struct StageLists {
    DevList<ConvTask> conv, conv_eval, dgrad;
    DevList<TcConvTask> conv_tc, dgrad_tc, conv_tc2, dgrad_tc2;
    DevList<TcConvTask> skip_fwd, skip_dg;       // 1x1 / stride-2 projections on skip_tc.cu (mma.sync)
    int skip_k = 0, skip_k_d = 0;               // largest GEMM K of those lists (shared-memory size)
    double f_skip = 0, f_skip_d = 0;
    int q_max = 0, tc2_cin = 0, tc2_cin_d = 0;   // largest patch; most GEMM input channels of the conv_tc2 / dgrad_tc2 tasks
    // algorithmic flop per SAMPLE of the stage's tensor-core launches (2*Ho*Wo*K*Cout summed over tasks; profiler only)
    double f_fwd2 = 0, f_fwd1 = 0, f_dg2 = 0, f_dg1 = 0, f_wg = 0, f_simt = 0, f_simt_dg = 0, f_simt_wg = 0;
    DevList<TcWgradTask> wgrad_tc, wgrad_tc2;
    int wg2_q = 0;                   // largest X patch of the stage's wgrad_tc2 tasks
    DevList<StatTask> stat;
    // stem.cu path: every task of conv / conv_eval / wgrad is an eligible Cin = 1 convolution
    bool stem = true;
    int stem_k = 0, stem_w = 0, stem_cout = 0, stem_splits = 0;
    bool stem_tc = true;             // ... and all of them go through the mma.sync kernels of stem_tc.cu (precision bf16)
    bool any = false;
    int max_bn_c = 0;                // widest BN unit of the stage (grid.y of the finalize kernels)
    DevList<PostTask> post_fwd, post_bn, post_bwd;
    DevList<WgradTask> wgrad;
    DevList<ReduceTask> wreduce;
    DevList<DropTask> drop_fwd, drop_bwd;
};

struct Wave {
    std::vector<Cand*> cands;
    int lane = 0;                    // concurrent lane of the wave (own stream, own scratch slots)
    StageLists st[N_STAGES];
    DevList<HeadTask> head;
    int head_fwd_total = 0;          // grid of gap_fwd (head.total is gap_bwd's)
    DevList<CeTask> ce_train, ce_val, ce_pred;
    DevList<AdamTask> adam;
    DevList<WtTask> wt;
    DevList<WtBf16Task> wt_bf16, wt_bf16_v2;
    DevList<PermTask> perm;          // per-epoch shuffles of the active candidates, generated on the device
    // one contiguous block per wave, so an epoch's results reach the host with ONE copy and ONE synchronisation:
    double* d_acc = nullptr;         // [cands][12]: train / validation / predict accumulators (Cand::acc points into it)
    int* d_cm = nullptr;             // [cands][C][C] confusion matrices (Cand::confusion points into it)
    bool weights_dirty = true;
    char* d_tmaps = nullptr;         // [tc2 units of the wave][4][128 B], written once per wave (build_tensor_maps)
    char* d_blob = nullptr;
    size_t blob_cap = 0;
};

template <class T>
void blob_add(std::vector<char>& blob, DevList<T>& l) {
    l.blob_off = (blob.size() + 255) / 256 * 256;
    blob.resize(l.blob_off + l.h.size() * sizeof(T));
    if (!l.h.empty()) memcpy(blob.data() + l.blob_off, l.h.data(), l.h.size() * sizeof(T));
    l.bt_off = (blob.size() + 255) / 256 * 256;
    blob.resize(l.bt_off + l.bt.size() * sizeof(int));
    if (!l.bt.empty()) memcpy(blob.data() + l.bt_off, l.bt.data(), l.bt.size() * sizeof(int));
}

inline int blocks_for(long long elems) { return (int)((elems + 255) / 256); }

struct Engine {
    const cmoop_cnn_dataset* data;
    cmoop_cnn_config cfg;
    int batch;
    cudaStream_t stream;
    int global_step = 0;   // dropout hash stream

    // ---- tiled-TMA descriptors of the wave's tensor-core units, encoded once per wave (buffers are fixed after place()):
    // per unit {input activation Vh, output gradient dU} as 4-D maps over the dense NHWC bf16 tensors, plus the same two
    // with N = the training split's tail batch, so that the samples past a short last batch read as zeros
    int build_tensor_maps(Wave& wv) {
        std::vector<char> host;
        std::vector<Unit*> owners;
        const int tail = data->n_train % batch ? data->n_train % batch : batch;
        for (Cand* cp : wv.cands)
            for (Unit& u : cp->units) {
                u.tmaps = nullptr;
                if (!u.tc2) continue;
                const size_t at = host.size();
                host.resize(at + 4 * kTmapBytes);
                const void* xin = cp->units[u.input].Vh;
                for (int t = 0; t < 2; ++t) {
                    const int n = t == 0 ? batch : tail;
                    if (Launch::make_row_tmap(host.data() + at + (2 * t + 0) * kTmapBytes, xin, u.cin, u.W, u.H, n, u.pad) != 0 ||
                        Launch::make_row_tmap(host.data() + at + (2 * t + 1) * kTmapBytes, cp->gBh, u.cout, u.Wo, u.Ho, n, u.pad) != 0)
                        return CMOOP_ERR_CUDA;
                }
                owners.push_back(&u);
            }
        if (host.empty()) return CMOOP_OK;
        wv.d_tmaps = (char*)cmoop::device_scratch(11 + wv.lane, host.size());
        if (!wv.d_tmaps) return CMOOP_ERR_CUDA;
        CMOOP_CUDA_OK(cmoop::copy_async(wv.d_tmaps, host.data(), host.size(), cudaMemcpyHostToDevice, stream));
        CMOOP_CUDA_OK(cudaStreamSynchronize(stream));          // `host` goes out of scope
        for (size_t i = 0; i < owners.size(); ++i) owners[i]->tmaps = wv.d_tmaps + i * 4 * kTmapBytes;
        return CMOOP_OK;
    }

    // ---- task-list construction for the currently active candidates of a wave
    int build_lists(Wave& wv) {
        for (int s = 0; s < N_STAGES; ++s) wv.st[s] = StageLists();
        wv.head = DevList<HeadTask>();
        wv.head_fwd_total = 0;
        wv.ce_train = wv.ce_val = wv.ce_pred = DevList<CeTask>();
        wv.adam = DevList<AdamTask>();
        wv.wt = DevList<WtTask>();
        wv.wt_bf16 = DevList<WtBf16Task>();
        wv.wt_bf16_v2 = DevList<WtBf16Task>();
        wv.perm = DevList<PermTask>();
        wv.weights_dirty = true;
        const long long img = (long long)data->H * data->W;
        for (Cand* cp : wv.cands) {
            Cand& c = *cp;
            if (!c.active) continue;
            for (size_t ui = 0; ui < c.units.size(); ++ui) {
                Unit& u = c.units[ui];
                StageLists& S = wv.st[u.stage];
                const float* xin = u.input == -1 ? data->x_train : (u.input == -2 ? (c.n_fc ? c.gap : c.gap) : c.units[u.input].V);
                // ---- forward conv
                S.any = true;
                if (u.tc) {
                    TcConvTask t{};
                    t.xh = c.units[u.input].Vh; t.wt = u.wtb; t.bias = c.p + u.w_off + (long long)u.k * u.k * u.cin * u.cout;
                    t.tmap = u.tmaps;                                   // {C = cin, W, H, N} of the producer's bf16 output
                    t.y = nullptr;                                      // bf16-only activation storage
                    t.yh = u.Uh;
                    t.H = u.H; t.W = u.W; t.Cin = u.cin; t.Ho = u.Ho; t.Wo = u.Wo; t.Cout = u.cout;
                    t.k = u.k; t.stride = u.stride; t.pad = u.pad; t.K_pad = u.kpad_f;
                    t.bn = u.cout < 128 ? u.cout : 128;
                    t.relu = u.relu_epi;
                    t.tiles_n = (u.cout + t.bn - 1) / t.bn;
                    if (u.tc2) {
                        const long long mq = (long long)batch * (u.H + 2 * u.pad) * (u.W + 2 * u.pad);
                        t.tile_begin = S.conv_tc2.total;
                        S.conv_tc2.h.push_back(t);
                        S.conv_tc2.total += (int)((mq + Launch::tc2_rows() - 1) / Launch::tc2_rows()) * t.tiles_n;
                        S.f_fwd2 += 2.0 * u.Ho * u.Wo * u.k * u.k * u.cin * u.cout;
                        S.q_max = std::max(S.q_max, Launch::tc2_q(u.W, u.k));
                        S.tc2_cin = std::max(S.tc2_cin, u.cin);
                    } else if (use_skip_tc() && u.is_skip && Launch::skip_tc_ok(t)) {
                        t.tiles_n = Launch::skip_tc_tiles_n(u.cout);
                        t.tile_begin = S.skip_fwd.total;
                        S.skip_fwd.h.push_back(t);
                        S.skip_fwd.total += Launch::skip_tc_tiles_m((long long)batch * u.Ho * u.Wo) * t.tiles_n;
                        S.skip_k = std::max(S.skip_k, u.cin);
                        S.f_skip += 2.0 * u.Ho * u.Wo * u.cin * u.cout;
                    } else {
                        t.tile_begin = S.conv_tc.total;
                        S.conv_tc.h.push_back(t);
                        S.conv_tc.total += (int)(((long long)batch * u.Ho * u.Wo + 127) / 128) * t.tiles_n;
                        S.f_fwd1 += 2.0 * u.Ho * u.Wo * u.k * u.k * u.cin * u.cout;
                    }
                    if (u.has_bn) {
                        StatTask sk{};
                        sk.y = u.U; sk.yh = u.Uh; sk.part = u.stat; sk.C = u.cout; sk.rows_per_sample = u.Ho * u.Wo;
                        sk.block_begin = S.stat.total;
                        S.stat.h.push_back(sk);
                        S.stat.total += u.stat_tiles;
                    }
                    WtBf16Task wb{};
                    wb.w = c.p + u.w_off; wb.out = u.wtb; wb.k = u.k; wb.Cin = u.cin; wb.Cout = u.cout;
                    wb.K_pad = u.kpad_f; wb.mode = u.tc2 ? 2 : 0;
                    if (u.tc2) {
                        wb.block_begin = wv.wt_bf16_v2.total;
                        wv.wt_bf16_v2.h.push_back(wb);
                        wv.wt_bf16_v2.total += blocks_for(Launch::tc2_weight_elems(u.cin, u.cout, u.k) / 8);
                    } else {
                        wb.block_begin = wv.wt_bf16.total;
                        wv.wt_bf16.h.push_back(wb);
                        wv.wt_bf16.total += blocks_for((long long)u.cout * u.kpad_f);
                    }
                } else {
                    ConvTask t{};
                    t.x = xin;
                    t.w = c.p + u.w_off;
                    t.y = u.U;                                          // null for a convolution unit in precision bf16
                    t.yh = u.Uh;
                    t.stat_part = u.has_bn ? u.stat : nullptr;
                    t.H = u.H; t.W = u.W; t.Cin = u.cin; t.Ho = u.Ho; t.Wo = u.Wo; t.Cout = u.cout;
                    t.k = u.k; t.stride = u.stride; t.pad = u.pad;
                    t.relu = u.relu_epi;
                    t.use_bias = 1;
                    t.tiles_n = (u.cout + 63) / 64;
                    t.tile_begin = S.conv.total;
                    if (u.stem && (S.stem_w == 0 || (S.stem_w == u.W && S.stem_splits == u.wg_splits))) {
                        S.stem_k = std::max(S.stem_k, u.k); S.stem_w = u.W; S.stem_cout = std::max(S.stem_cout, u.cout);
                        S.stem_splits = u.wg_splits;
                        // CMOOP_CNN_NO_STEM_TC=1: A/B switch back to the fp32 SIMT stem kernels
                        static const bool no_stem_tc = getenv("CMOOP_CNN_NO_STEM_TC") != nullptr;
                        if (no_stem_tc || cfg.precision != 1 || t.y || !t.yh || !Launch::stem_tc_ok(u.H, u.W, u.cout, u.k))
                            S.stem_tc = false;
                    } else {
                        S.stem = false;
                    }
                    if (u.input == -1) {
                        t.gather = c.perm;
                        t.gather_step = batch;
                        ConvTask e = t;
                        e.x = data->x_val;
                        e.gather = nullptr;
                        e.gather_step = 0;
                        e.x_step = (long long)batch * img;
                        e.tile_begin = S.conv_eval.total;
                        S.conv_eval.h.push_back(e);
                        S.conv_eval.total += u.stat_tiles * t.tiles_n;
                    }
                    S.conv.h.push_back(t);
                    S.conv.total += u.stat_tiles * t.tiles_n;
                    S.f_simt += 2.0 * u.Ho * u.Wo * (u.k * u.k * u.cin + 1) * u.cout;
                }
                // ---- post stage (forward / BN / backward lists)
                if (!u.dense && !u.is_skip) {
                    PostTask p{};
                    p.u = u.U; p.uh = u.Uh; p.v = u.V;
                    p.skip = u.add_skip ? c.units[u.skip_unit].U : nullptr;
                    p.skiph = u.add_skip ? c.units[u.skip_unit].Uh : nullptr;
                    p.idx = u.idx;
                    if (u.has_bn) {
                        float* bnp = c.p + u.bn_off;
                        p.gamma = bnp; p.beta = bnp + u.cout; p.mov_mean = bnp + 2 * u.cout; p.mov_var = bnp + 3 * u.cout;
                        p.dgamma = c.grad + u.bn_off; p.dbeta = c.grad + u.bn_off + u.cout;
                        p.stat_part = u.stat; p.bn = u.bn; p.bwd_part = u.bwd_part;
                    }
                    // precision bf16: every convolution's weight / data gradient reads the bf16 gradient only
                    const bool half = cfg.precision == 1;
                    p.dv = c.gA; p.du = half ? nullptr : c.gB;
                    p.dskip = (u.add_skip && !half) ? c.gS : nullptr;
                    p.vh = u.post_fwd ? u.Vh : nullptr;
                    p.duh = half ? c.gBh : nullptr;
                    p.dskiph = (u.add_skip && half) ? c.gSh : nullptr;
                    p.H = u.Ho; p.W = u.Wo; p.C = u.cout; p.Ho = u.Po; p.Wo = u.Qo;
                    p.pool = u.pool; p.relu_mid = u.relu_mid; p.add_skip = u.add_skip; p.relu_in = u.relu_epi;
                    p.has_bn = u.has_bn;
                    p.stat_tiles = u.stat_tiles; p.bwd_rows = u.bwd_rows; p.bwd_pix = u.bwd_pix;
                    if (u.post_fwd) {
                        PostTask q = p;
                        q.block_begin = S.post_fwd.total;
                        S.post_fwd.h.push_back(q);
                        S.post_fwd.total += Launch::post_blocks((long long)batch * u.Po * u.Qo, u.cout);
                    }
                    if (u.has_bn) {
                        S.max_bn_c = std::max(S.max_bn_c, u.cout);
                        PostTask q = p;
                        q.block_begin_bwd = S.post_bn.total;
                        S.post_bn.h.push_back(q);
                        S.post_bn.total += u.bwd_rows;
                    }
                    PostTask q = p;
                    q.block_begin_apply = S.post_bwd.total;
                    S.post_bwd.h.push_back(q);
                    S.post_bwd.total += Launch::post_blocks((long long)batch * u.Po * u.Qo, u.cout);
                }
                // ---- dense ReLU / dropout
                if (u.dense && u.fc_index >= 0) {
                    DropTask d{};
                    d.u = u.U; d.v = u.V; d.dv = c.gA; d.dz = c.gB;
                    d.seed = (unsigned)(c.seed & 0xffffffffu);
                    d.layer = u.fc_index; d.units = u.cout; d.use_dropout = u.use_dropout;
                    if (u.use_dropout) {
                        d.block_begin = S.drop_fwd.total;
                        S.drop_fwd.h.push_back(d);
                        S.drop_fwd.total += blocks_for((long long)batch * u.cout);
                    }
                    d.block_begin = S.drop_bwd.total;
                    S.drop_bwd.h.push_back(d);
                    S.drop_bwd.total += blocks_for((long long)batch * u.cout);
                }
                // ---- weight gradient (+ bias row)
                {
                    const bool to_ws = u.wg_splits > 1;
                    const int kext = u.k * u.k * u.cin + 1;
                    const float* dy = u.is_skip ? c.gS : (u.stage == ST_OUT ? c.dlogits : c.gB);
                    float* dst = to_ws ? c.wg_ws : c.grad + u.w_off;
                    if (u.tc) {
                        TcWgradTask g{};
                        g.xh = c.units[u.input].Vh; g.dyh = u.is_skip ? c.gSh : c.gBh; g.out = dst;
                        g.tmaps = u.tmaps;
                        g.n_full = batch;
                        g.H = u.H; g.W = u.W; g.Cin = u.cin; g.Ho = u.Ho; g.Wo = u.Wo; g.Cout = u.cout;
                        g.k = u.k; g.stride = u.stride; g.pad = u.pad;
                        g.splits = u.wg_splits; g.m_chunk = u.wg_chunk;
                        g.bn = u.cout < 128 ? u.cout : 128;
                        g.tiles_k = (kext + 127) / 128; g.tiles_n = (u.cout + g.bn - 1) / g.bn;
                        if (u.wg2) {
                            g.bn = Launch::wg2_bn(u.cout);
                            g.tiles_n = u.cout / g.bn;
                            g.tile_begin = S.wgrad_tc2.total;
                            S.wgrad_tc2.h.push_back(g);
                            S.wgrad_tc2.total += g.splits * Launch::wg2_items(u.cin, u.cout, u.k);
                            S.f_wg += 2.0 * u.Ho * u.Wo * kext * u.cout;
                            const int wq = Launch::wg2_q(u.W, u.k);          // X patch | dY tile positions, 16 bits each
                            S.wg2_q = std::max(S.wg2_q & 0xffff, wq & 0xffff) | (std::max(S.wg2_q >> 16, wq >> 16) << 16);
                        } else {
                            g.tile_begin = S.wgrad_tc.total;
                            S.wgrad_tc.h.push_back(g);
                            S.wgrad_tc.total += g.tiles_k * g.tiles_n * g.splits;
                            S.f_wg += 2.0 * u.Ho * u.Wo * kext * u.cout;
                        }
                    } else {
                        WgradTask g{};
                        g.x = xin;
                        if (u.input == -1) { g.gather = c.perm; g.gather_step = batch; }
                        g.dy = dy;
                        g.dyh = (cfg.precision == 1 && !u.dense) ? (u.is_skip ? c.gSh : c.gBh) : nullptr;
                        g.out = dst;
                        g.H = u.H; g.W = u.W; g.Cin = u.cin; g.Ho = u.Ho; g.Wo = u.Wo; g.Cout = u.cout;
                        g.k = u.k; g.stride = u.stride; g.pad = u.pad;
                        g.splits = u.wg_splits; g.m_chunk = u.wg_chunk;
                        g.tiles_k = (kext + 63) / 64; g.tiles_n = (u.cout + 63) / 64;
                        g.tile_begin = S.wgrad.total;
                        S.wgrad.h.push_back(g);
                        S.wgrad.total += g.tiles_k * g.tiles_n * g.splits;
                        S.f_simt_wg += 2.0 * u.Ho * u.Wo * kext * u.cout;
                    }
                    if (to_ws) {
                        ReduceTask r{};
                        r.part = c.wg_ws; r.out = c.grad + u.w_off; r.n = kext * u.cout; r.splits = u.wg_splits;
                        r.block_begin = S.wreduce.total;
                        S.wreduce.h.push_back(r);
                        S.wreduce.total += blocks_for(r.n);
                    }
                }
                // ---- data gradient: conv of dy with flipped/transposed weights into buffer A (or gG below the first FC)
                if (u.need_dgrad && u.tc) {
                    WtBf16Task wb{};
                    wb.w = c.p + u.w_off; wb.out = u.wtd; wb.k = u.k; wb.Cin = u.cin; wb.Cout = u.cout;
                    wb.K_pad = u.kpad_d; wb.mode = u.tc2 ? 3 : 1;
                    if (u.tc2) {
                        wb.block_begin = wv.wt_bf16_v2.total;
                        wv.wt_bf16_v2.h.push_back(wb);
                        wv.wt_bf16_v2.total += blocks_for(Launch::tc2_weight_elems(u.cout, u.cin, u.k) / 8);
                    } else {
                        wb.block_begin = wv.wt_bf16.total;
                        wv.wt_bf16.h.push_back(wb);
                        wv.wt_bf16.total += blocks_for((long long)u.cin * u.kpad_d);
                    }
                    TcConvTask d{};
                    d.xh = u.is_skip ? c.gSh : c.gBh;
                    d.tmap = u.tmaps ? u.tmaps + kTmapBytes : nullptr;  // {C = cout, W, H, N} of this unit's dU
                    d.wt = u.wtd;
                    d.y = c.gA;
                    d.Cin = u.cout; d.Cout = u.cin; d.k = u.k; d.K_pad = u.kpad_d;
                    d.H = u.Ho; d.W = u.Wo; d.Ho = u.Ho; d.Wo = u.Wo;
                    d.stride = 1; d.pad = u.pad;
                    if (u.is_skip) {
                        d.pad = 0;
                        d.out_h = u.H; d.out_w = u.W; d.out_s = 2;
                        d.accumulate = 1;
                    }
                    d.bn = u.cin < 128 ? u.cin : 128;
                    d.tiles_n = (u.cin + d.bn - 1) / d.bn;
                    if (u.tc2) {
                        const long long mq = (long long)batch * (u.H + 2 * u.pad) * (u.W + 2 * u.pad);
                        d.tile_begin = S.dgrad_tc2.total;
                        S.dgrad_tc2.h.push_back(d);
                        S.dgrad_tc2.total += (int)((mq + Launch::tc2_rows() - 1) / Launch::tc2_rows()) * d.tiles_n;
                        S.f_dg2 += 2.0 * u.Ho * u.Wo * u.k * u.k * u.cin * u.cout;
                        S.q_max = std::max(S.q_max, Launch::tc2_q(u.W, u.k));
                        S.tc2_cin_d = std::max(S.tc2_cin_d, u.cout);
                    } else if (use_skip_tc() && u.is_skip && Launch::skip_tc_ok(d)) {
                        d.tiles_n = Launch::skip_tc_tiles_n(u.cin);
                        d.tile_begin = S.skip_dg.total;
                        S.skip_dg.h.push_back(d);
                        S.skip_dg.total += Launch::skip_tc_tiles_m((long long)batch * u.Ho * u.Wo) * d.tiles_n;
                        S.skip_k_d = std::max(S.skip_k_d, u.cout);
                        S.f_skip_d += 2.0 * u.Ho * u.Wo * u.cin * u.cout;
                    } else {
                        d.tile_begin = S.dgrad_tc.total;
                        S.dgrad_tc.h.push_back(d);
                        S.dgrad_tc.total += (int)(((long long)batch * u.Ho * u.Wo + 127) / 128) * d.tiles_n;
                        S.f_dg1 += 2.0 * u.Ho * u.Wo * u.k * u.k * u.cin * u.cout;
                    }
                } else if (u.need_dgrad) {
                    WtTask w{};
                    w.w = c.p + u.w_off; w.wt = u.wt; w.k = u.k; w.Cin = u.cin; w.Cout = u.cout;
                    w.block_begin = wv.wt.total;
                    wv.wt.h.push_back(w);
                    wv.wt.total += blocks_for((long long)u.k * u.k * u.cin * u.cout);
                    ConvTask d{};
                    d.x = u.is_skip ? c.gS : (u.stage == ST_OUT ? c.dlogits : c.gB);
                    d.w = u.wt;
                    d.y = u.input == -2 ? c.gG : c.gA;
                    d.Cin = u.cout; d.Cout = u.cin; d.k = u.k;
                    d.H = u.Ho; d.W = u.Wo; d.Ho = u.Ho; d.Wo = u.Wo;
                    d.stride = 1; d.pad = u.pad;
                    if (u.is_skip) {
                        d.pad = 0;
                        d.out_h = u.H; d.out_w = u.W; d.out_s = 2;
                        d.accumulate = 1;
                    }
                    d.tiles_n = (u.cin + 63) / 64;
                    d.tile_begin = S.dgrad.total;
                    S.dgrad.h.push_back(d);
                    S.dgrad.total += u.stat_tiles * d.tiles_n;
                    S.f_simt_dg += 2.0 * u.Ho * u.Wo * u.k * u.k * u.cin * u.cout;
                }
            }
            const Unit& lc = c.units[c.last_conv];
            HeadTask hd{};
            hd.v = lc.V; hd.vh = lc.V ? nullptr : lc.Vh; hd.gap = c.gap; hd.dgap = c.gG; hd.dv = c.gA;
            hd.Hf = lc.Po; hd.Wf = lc.Qo; hd.C = lc.cout;
            hd.block_begin = wv.head.total;
            hd.block_begin_fwd = wv.head_fwd_total;
            wv.head.h.push_back(hd);
            wv.head.total += Launch::gap_blocks((long long)batch * lc.Po * lc.Qo * lc.cout);
            wv.head_fwd_total += Launch::gap_blocks((long long)batch * lc.cout);
            const Unit& ou = c.units.back();
            CeTask ce{};
            ce.logits = ou.U; ce.dlogits = c.dlogits; ce.n_classes = cfg.n_classes;
            ce.labels = data->y_train; ce.gather = c.perm; ce.gather_step = batch; ce.acc = c.acc;
            wv.ce_train.h.push_back(ce);
            ce.labels = data->y_val; ce.gather = nullptr; ce.gather_step = 0; ce.label_step = batch; ce.acc = c.acc + 4;
            wv.ce_val.h.push_back(ce);
            ce.acc = c.acc + 8; ce.pred = c.pred; ce.confusion = c.confusion; ce.y_true_zero = cfg.y_true_zero;
            wv.ce_pred.h.push_back(ce);
            PermTask pt{};
            pt.perm = c.perm; pt.seed_lo = (unsigned)(c.seed & 0xffffffffu); pt.seed_hi = (unsigned)(c.seed >> 32);
            wv.perm.h.push_back(pt);
            AdamTask ad{};
            ad.p = c.p; ad.g = c.grad; ad.m = c.m; ad.v = c.v; ad.n = (int)c.n_params;
            ad.block_begin = wv.adam.total;
            wv.adam.h.push_back(ad);
            wv.adam.total += (int)((c.n_params + 1023) / 1024);      // adam_kernel: 1 024 parameters per block
        }
        for (int s = 0; s < N_STAGES; ++s) {
            StageLists& S = wv.st[s];
            fill_block_table(S.conv_tc2, [](const TcConvTask& t) { return t.tile_begin; });
            fill_block_table(S.dgrad_tc2, [](const TcConvTask& t) { return t.tile_begin; });
            fill_block_table(S.wgrad_tc2, [](const TcWgradTask& t) { return t.tile_begin; });
            fill_block_table(S.conv_tc, [](const TcConvTask& t) { return t.tile_begin; });
            fill_block_table(S.dgrad_tc, [](const TcConvTask& t) { return t.tile_begin; });
            fill_block_table(S.wgrad_tc, [](const TcWgradTask& t) { return t.tile_begin; });
            // (the elementwise post kernels keep the shared binary search: a direct table measured slower there)
        }
        // ---- one blob upload, then fix the device pointers
        std::vector<char> blob;
        for (int s = 0; s < N_STAGES; ++s) {
            StageLists& S = wv.st[s];
            blob_add(blob, S.conv); blob_add(blob, S.conv_eval); blob_add(blob, S.dgrad);
            blob_add(blob, S.conv_tc); blob_add(blob, S.dgrad_tc); blob_add(blob, S.stat); blob_add(blob, S.wgrad_tc);
            blob_add(blob, S.conv_tc2); blob_add(blob, S.dgrad_tc2); blob_add(blob, S.wgrad_tc2);
            blob_add(blob, S.skip_fwd); blob_add(blob, S.skip_dg);
            blob_add(blob, S.post_fwd); blob_add(blob, S.post_bn); blob_add(blob, S.post_bwd);
            blob_add(blob, S.wgrad); blob_add(blob, S.wreduce); blob_add(blob, S.drop_fwd); blob_add(blob, S.drop_bwd);
        }
        blob_add(blob, wv.head); blob_add(blob, wv.ce_train); blob_add(blob, wv.ce_val); blob_add(blob, wv.ce_pred);
        blob_add(blob, wv.adam); blob_add(blob, wv.wt); blob_add(blob, wv.wt_bf16); blob_add(blob, wv.wt_bf16_v2);
        blob_add(blob, wv.perm);
        if (blob.size() > wv.blob_cap) {
            if (wv.d_blob) {
                CMOOP_CUDA_OK(cudaStreamSynchronize(stream));
                CMOOP_CUDA_OK(cudaFree(wv.d_blob));
            }
            wv.blob_cap = blob.size() * 2 + 4096;
            CMOOP_CUDA_OK(cudaMalloc((void**)&wv.d_blob, wv.blob_cap));
        }
        CMOOP_CUDA_OK(cudaStreamSynchronize(stream));   // previous launches may still read the old lists
        CMOOP_CUDA_OK(cmoop::copy_sync(wv.d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        auto fix = [&](auto& l) {
            l.d = reinterpret_cast<decltype(l.d)>(wv.d_blob + l.blob_off);
            l.d_bt = l.bt.empty() ? nullptr : reinterpret_cast<int*>(wv.d_blob + l.bt_off);
        };
        for (int s = 0; s < N_STAGES; ++s) {
            StageLists& S = wv.st[s];
            fix(S.conv); fix(S.conv_eval); fix(S.dgrad); fix(S.post_fwd); fix(S.post_bn); fix(S.post_bwd);
            fix(S.conv_tc); fix(S.dgrad_tc); fix(S.stat); fix(S.wgrad_tc); fix(S.conv_tc2); fix(S.dgrad_tc2); fix(S.wgrad_tc2);
            fix(S.wgrad); fix(S.wreduce); fix(S.drop_fwd); fix(S.drop_bwd); fix(S.skip_fwd); fix(S.skip_dg);
        }
        fix(wv.head); fix(wv.ce_train); fix(wv.ce_val); fix(wv.ce_pred); fix(wv.adam); fix(wv.wt); fix(wv.wt_bf16); fix(wv.wt_bf16_v2);
        fix(wv.perm);
        return CMOOP_OK;
    }

#define CNN_LAUNCH_N(label, flops, expr)                                                            \
    do {                                                                                           \
        prof_begin(label, (double)(flops), stream);                                                \
        int _e = (expr);                                                                           \
        prof_end(stream);                                                                          \
        cmoop::count_launch();                                                                     \
        if (_e == 0 && debug_sync()) _e = (int)cudaStreamSynchronize(stream);  /* CMOOP_CNN_SYNC=1 */ \
        if (_e != 0) {                                                                             \
            cmoop::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString((cudaError_t)_e)); \
            return CMOOP_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)
#define CNN_LAUNCH(expr) CNN_LAUNCH_N(#expr, 0.0, expr)

    // mode 0: training batch from the permutation; 1: validation loss/accuracy; 2: final predict (+confusion)
    int run_forward(Wave& wv, int mode, int step, int n_b) {
        const int training = mode == 0;
        if (wv.weights_dirty && !wv.wt_bf16.h.empty()) {
            CNN_LAUNCH(Launch::wt_bf16(wv.wt_bf16.d, (int)wv.wt_bf16.h.size(), wv.wt_bf16.total, stream));
        }
        if (wv.weights_dirty && !wv.wt_bf16_v2.h.empty()) {
            CNN_LAUNCH(Launch::wt_bf16_v2(wv.wt_bf16_v2.d, (int)wv.wt_bf16_v2.h.size(), wv.wt_bf16_v2.total, stream));
        }
        wv.weights_dirty = false;
        for (int s = 0; s < N_STAGES; ++s) {
            StageLists& S = wv.st[s];
            g_prof_stage = s;
            if (s == ST_FC0) CNN_LAUNCH(Launch::gap_fwd(wv.head.d, (int)wv.head.h.size(), wv.head_fwd_total, n_b, stream));
            if (!S.any) continue;
            DevList<ConvTask>& cl = (s == 0 && !training) ? S.conv_eval : S.conv;
            if (!cl.h.empty()) {
                if (S.stem && S.stem_w > 0 && S.stem_tc)
                    CNN_LAUNCH_N("stem_conv", S.f_simt * n_b,
                                 Launch::stem_conv_tc(cl.d, (int)cl.h.size(), S.stem_k, S.stem_w,
                                                      (long long)n_b * cl.h[0].H * cl.h[0].W, n_b, step, stream));
                else if (S.stem && S.stem_w > 0)
                    CNN_LAUNCH_N("stem_conv", S.f_simt * n_b,
                                 Launch::stem_conv(cl.d, (int)cl.h.size(), S.stem_k, S.stem_w, S.stem_cout,
                                                   (long long)n_b * cl.h[0].H * cl.h[0].W, n_b, step, stream));
                else
                    CNN_LAUNCH_N(s >= ST_FC0 ? "dense_fwd(simt)" : "conv_fwd(simt)", S.f_simt * n_b,
                                 Launch::conv(cl.d, (int)cl.h.size(), cl.total, n_b, step, stream));
            }
            if (!S.conv_tc.h.empty())
                CNN_LAUNCH_N("conv_tc.fwd", S.f_fwd1 * n_b,
                             Launch::conv_tc(S.conv_tc.d, (int)S.conv_tc.h.size(), S.conv_tc.total, n_b, step, stream, S.conv_tc.d_bt));
            if (!S.skip_fwd.h.empty())
                CNN_LAUNCH_N("skip_tc.fwd", S.f_skip * n_b,
                             Launch::skip_tc(S.skip_fwd.d, (int)S.skip_fwd.h.size(), S.skip_fwd.total, n_b, step, S.skip_k, stream));
            if (!S.conv_tc2.h.empty())
                CNN_LAUNCH_N("conv_tc2.fwd", S.f_fwd2 * n_b,
                             Launch::conv_tc2(S.conv_tc2.d, (int)S.conv_tc2.h.size(), S.conv_tc2.total, n_b, step, S.q_max,
                                              S.tc2_cin, stream, S.conv_tc2.d_bt, /*out_bf16=*/true));
            if (!S.stat.h.empty() && training)
                CNN_LAUNCH(Launch::bn_stats(S.stat.d, (int)S.stat.h.size(), S.stat.total, n_b, stream));
            if (!S.post_bn.h.empty())
                CNN_LAUNCH(Launch::bn_finalize(S.post_bn.d, (int)S.post_bn.h.size(), S.max_bn_c, n_b, training, cfg.bn_momentum,
                                               cfg.bn_eps, stream));
            if (!S.post_fwd.h.empty())
                CNN_LAUNCH(Launch::post_fwd(S.post_fwd.d, (int)S.post_fwd.h.size(), S.post_fwd.total, n_b, stream, nullptr,
                                            cfg.precision == 1));
            if (!S.drop_fwd.h.empty())
                CNN_LAUNCH(Launch::drop_fwd(S.drop_fwd.d, (int)S.drop_fwd.h.size(), S.drop_fwd.total, n_b, global_step,
                                            training, cfg.dropout_rate, stream));
        }
        g_prof_stage = -1;
        DevList<CeTask>& ce = mode == 0 ? wv.ce_train : (mode == 1 ? wv.ce_val : wv.ce_pred);
        CNN_LAUNCH(Launch::ce(ce.d, (int)ce.h.size(), n_b, step, training, stream));
        return CMOOP_OK;
    }

    int run_backward(Wave& wv, int step, int n_b) {
        if (!wv.wt.h.empty()) CNN_LAUNCH(Launch::wt(wv.wt.d, (int)wv.wt.h.size(), wv.wt.total, stream));
        for (int s = N_STAGES - 1; s >= 0; --s) {
            StageLists& S = wv.st[s];
            g_prof_stage = s;
            if (s == ST_FC0 - 1) CNN_LAUNCH(Launch::gap_bwd(wv.head.d, (int)wv.head.h.size(), wv.head.total, n_b, stream));
            if (!S.any) continue;
            if (!S.drop_bwd.h.empty())      // dense: A (grad of the layer output) -> B (grad of the pre-activation)
                CNN_LAUNCH(Launch::drop_bwd(S.drop_bwd.d, (int)S.drop_bwd.h.size(), S.drop_bwd.total, n_b, global_step,
                                            cfg.dropout_rate, stream));
            if (!S.post_bn.h.empty()) {
                CNN_LAUNCH(Launch::post_bwd_reduce(S.post_bn.d, (int)S.post_bn.h.size(), S.post_bn.total, n_b, stream, nullptr,
                                                   cfg.precision == 1));
                CNN_LAUNCH(Launch::bn_bwd_finalize(S.post_bn.d, (int)S.post_bn.h.size(), S.max_bn_c, n_b, stream));
            }
            if (!S.post_bwd.h.empty())
                CNN_LAUNCH(Launch::post_bwd_apply(S.post_bwd.d, (int)S.post_bwd.h.size(), S.post_bwd.total, n_b, stream, nullptr,
                                                  cfg.precision == 1));
            if (!S.wgrad.h.empty()) {
                if (S.stem && S.stem_w > 0 && S.stem_tc)
                    CNN_LAUNCH_N("stem_wgrad", S.f_simt_wg * n_b,
                                 Launch::stem_wgrad_tc(S.wgrad.d, (int)S.wgrad.h.size(), S.stem_k, S.stem_w, S.stem_cout,
                                                       S.stem_splits, n_b, step, stream));
                else if (S.stem && S.stem_w > 0)
                    CNN_LAUNCH_N("stem_wgrad", S.f_simt_wg * n_b,
                                 Launch::stem_wgrad(S.wgrad.d, (int)S.wgrad.h.size(), S.stem_k, S.stem_w, S.stem_cout,
                                                    S.stem_splits, n_b, step, stream));
                else
                    CNN_LAUNCH_N(s >= ST_FC0 ? "dense_wgrad(simt)" : "conv_wgrad(simt)", S.f_simt_wg * n_b,
                                 Launch::wgrad(S.wgrad.d, (int)S.wgrad.h.size(), S.wgrad.total, n_b, step, stream));
            }
            if (!S.wgrad_tc.h.empty())
                CNN_LAUNCH_N("wgrad_tc", S.f_wg * n_b,
                             Launch::wgrad_tc(S.wgrad_tc.d, (int)S.wgrad_tc.h.size(), S.wgrad_tc.total, n_b, stream, S.wgrad_tc.d_bt));
            if (!S.wgrad_tc2.h.empty())
                CNN_LAUNCH_N("wgrad_tc2", S.f_wg * n_b,
                             Launch::wgrad_tc2(S.wgrad_tc2.d, (int)S.wgrad_tc2.h.size(), S.wgrad_tc2.total, n_b, S.wg2_q, stream,
                                               S.wgrad_tc2.d_bt));
            if (!S.wreduce.h.empty())
                CNN_LAUNCH(Launch::reduce(S.wreduce.d, (int)S.wreduce.h.size(), S.wreduce.total, stream));
            if (!S.dgrad.h.empty())
                CNN_LAUNCH_N(s >= ST_FC0 ? "dense_dgrad(simt)" : "conv_dgrad(simt)", S.f_simt_dg * n_b,
                             Launch::conv(S.dgrad.d, (int)S.dgrad.h.size(), S.dgrad.total, n_b, 0, stream));
            if (!S.dgrad_tc.h.empty())
                CNN_LAUNCH_N("conv_tc.dgrad", S.f_dg1 * n_b,
                             Launch::conv_tc(S.dgrad_tc.d, (int)S.dgrad_tc.h.size(), S.dgrad_tc.total, n_b, 0, stream, S.dgrad_tc.d_bt));
            if (!S.skip_dg.h.empty())
                CNN_LAUNCH_N("skip_tc.dgrad", S.f_skip_d * n_b,
                             Launch::skip_tc(S.skip_dg.d, (int)S.skip_dg.h.size(), S.skip_dg.total, n_b, 0, S.skip_k_d, stream));
            if (!S.dgrad_tc2.h.empty())
                CNN_LAUNCH_N("conv_tc2.dgrad", S.f_dg2 * n_b,
                             Launch::conv_tc2(S.dgrad_tc2.d, (int)S.dgrad_tc2.h.size(), S.dgrad_tc2.total, n_b, 0, S.q_max,
                                              S.tc2_cin_d, stream, S.dgrad_tc2.d_bt));
        }
        g_prof_stage = -1;
        return CMOOP_OK;
    }

    int adam_step(Wave& wv, int t) {
        const double b1 = cfg.beta1, b2 = cfg.beta2;
        const double alpha = cfg.learning_rate * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
        CNN_LAUNCH(Launch::adam(wv.adam.d, (int)wv.adam.h.size(), wv.adam.total, (float)alpha, cfg.beta1, cfg.beta2,
                                cfg.adam_eps, stream));
        wv.weights_dirty = true;
        return CMOOP_OK;
    }

    int init_params(const std::vector<Cand*>& cands) {
        std::vector<InitTask> tasks;
        int total = 0;
        for (Cand* cp : cands) {
            Cand& c = *cp;
            int tensor = 0;
            auto add = [&](float* p, long long n, int kind, float limit, float value) {
                InitTask t{};
                t.p = p; t.n = (int)n; t.tensor = tensor++; t.kind = kind; t.limit = limit; t.value = value;
                t.seed = (unsigned)(c.seed & 0xffffffffu) ^ (unsigned)(c.seed >> 32);
                t.block_begin = total;
                total += blocks_for(n);
                tasks.push_back(t);
            };
            for (const Unit& u : c.units) {
                const double fan_in = (double)u.k * u.k * u.cin, fan_out = (double)u.k * u.k * u.cout;
                add(c.p + u.w_off, (long long)u.k * u.k * u.cin * u.cout, 0, (float)sqrt(6.0 / (fan_in + fan_out)), 0.f);
                add(c.p + u.w_off + (long long)u.k * u.k * u.cin * u.cout, u.cout, 1, 0.f, 0.f);
                if (u.has_bn) {
                    add(c.p + u.bn_off, u.cout, 1, 0.f, 1.f);
                    add(c.p + u.bn_off + u.cout, u.cout, 1, 0.f, 0.f);
                    add(c.p + u.bn_off + 2 * u.cout, u.cout, 1, 0.f, 0.f);
                    add(c.p + u.bn_off + 3 * u.cout, u.cout, 1, 0.f, 1.f);
                }
            }
            CMOOP_CUDA_OK(cudaMemsetAsync(c.grad, 0, c.n_params * sizeof(float), stream));
            CMOOP_CUDA_OK(cudaMemsetAsync(c.m, 0, c.n_params * sizeof(float), stream));
            CMOOP_CUDA_OK(cudaMemsetAsync(c.v, 0, c.n_params * sizeof(float), stream));
            CMOOP_CUDA_OK(cudaMemsetAsync(c.acc, 0, 12 * sizeof(double), stream));
        }
        InitTask* d = nullptr;
        CMOOP_CUDA_OK(cudaMalloc((void**)&d, tasks.size() * sizeof(InitTask)));
        CMOOP_CUDA_OK(cmoop::copy_async(d, tasks.data(), tasks.size() * sizeof(InitTask), cudaMemcpyHostToDevice, stream));
        int rc = Launch::init(d, (int)tasks.size(), total, stream);
        cmoop::count_launch();
        CMOOP_CUDA_OK(cudaStreamSynchronize(stream));
        cudaFree(d);
        if (rc != 0) {
            cmoop::set_error("init kernel: %s", cudaGetErrorString((cudaError_t)rc));
            return CMOOP_ERR_CUDA;
        }
        return CMOOP_OK;
    }
};

// Deterministic Fisher-Yates driven by the fmix32 hash (the harness-imposed "Keras shuffle").
void make_permutation(uint64_t seed, int epoch, int n, int* out) {
    for (int i = 0; i < n; ++i) out[i] = i;
    const unsigned s = fmix32((unsigned)(seed & 0xffffffffu) ^ 0x5bd1e995u * (unsigned)(epoch + 1)) ^ (unsigned)(seed >> 32);
    for (int i = n - 1; i > 0; --i) {
        const unsigned r = fmix32(fmix32(s ^ 0x27d4eb2fu) ^ (unsigned)i);
        const int j = (int)(r % (unsigned)(i + 1));
        const int t = out[i];
        out[i] = out[j];
        out[j] = t;
    }
}

int check_config(const cmoop_genotype* g, int n, const cmoop_cnn_config* cfg) {
    CMOOP_REQUIRE(cfg != nullptr, "cnn: null config");
    CMOOP_REQUIRE(cfg->variant == 0 || cfg->variant == 1, "cnn: variant must be 0 (A) or 1 (B)");
    CMOOP_REQUIRE(cfg->n_classes >= 2 && cfg->n_classes <= 4096, "cnn: n_classes=%d outside [2,4096]", cfg->n_classes);
    CMOOP_REQUIRE(cfg->batch_size >= 1 && cfg->batch_size <= kBatch, "cnn: batch_size=%d outside [1,%d]", cfg->batch_size,
                  kBatch);
    CMOOP_REQUIRE(cfg->max_epochs >= 1 && cfg->patience >= 0, "cnn: bad epochs/patience");
    if (cfg->precision != 0 && cfg->precision != 1) {
        cmoop::set_error("cnn: precision=%d unknown (0 = fp32 SIMT, 1 = bf16 tcgen05)", cfg->precision);
        return CMOOP_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < n; ++i) {
        CMOOP_REQUIRE(g[i].filters >= 4 && g[i].filters <= 256 && g[i].filters % 4 == 0,
                      "cnn: genotype %d filters=%d must be a multiple of 4 in [4,256]", i, g[i].filters);
        if (cfg->precision == 1 && g[i].filters % 16 != 0) {
            // precision bf16 stores every convolution's activations in bf16 only and runs every Cin >= 16 convolution
            // on the tensor cores: channel counts must be multiples of 16 (the reference's space is {16, 32, 64})
            cmoop::set_error("cnn: genotype %d filters=%d: precision bf16 needs a multiple of 16", i, g[i].filters);
            return CMOOP_ERR_UNSUPPORTED;
        }
        CMOOP_REQUIRE(g[i].kernel_size == 1 || g[i].kernel_size == 3 || g[i].kernel_size == 5 || g[i].kernel_size == 7,
                      "cnn: genotype %d kernel_size=%d must be odd and <= 7", i, g[i].kernel_size);
        CMOOP_REQUIRE(g[i].residual_blocks >= 0 && g[i].residual_blocks <= 3, "cnn: genotype %d residual_blocks=%d", i,
                      g[i].residual_blocks);
        CMOOP_REQUIRE(g[i].fc_layers >= 1 && g[i].fc_layers <= 4, "cnn: genotype %d fc_layers=%d outside [1,4]", i,
                      g[i].fc_layers);
    }
    return CMOOP_OK;
}

// numpy.mean of a short Python list, bit for bit: np.add.reduce's pairwise summation (8 interleaved accumulators on blocks
// of at most 128 elements, halves split on multiples of 8 above that) followed by one division -- what
// `return np.mean(fpr_vals)` (nsga_penalty.py:364) does to the per-class rates.
double np_pairwise_sum(const double* a, size_t n) {
    if (n < 8) {
        double r = 0.0;
        for (size_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        size_t i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// calculate_fpr from the C x C confusion matrix.  Per class FP / (FP + TN) with FP + TN = total - row_i; classes with an
// empty denominator count as 0.0 (nsga_penalty.py:357-363, init_sa_nsga_local.py:139-143) or are left out of the mean
// (filtered form, sa_nsga_local.py:140-141; 0.0 when no class qualifies).
double fpr_from_confusion(const std::vector<int>& cm, int C, bool filtered) {
    long long total = 0;
    std::vector<long long> row(C, 0), col(C, 0);
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) {
            const int v = cm[(size_t)i * C + j];
            total += v;
            row[i] += v;
            col[j] += v;
        }
    std::vector<double> vals;
    vals.reserve(C);
    for (int i = 0; i < C; ++i) {
        const long long fp = col[i] - cm[(size_t)i * C + i];
        const long long denom = total - row[i];
        if (denom > 0)
            vals.push_back((double)fp / (double)denom);
        else if (!filtered)
            vals.push_back(0.0);
    }
    if (vals.empty()) return 0.0;
    return np_pairwise_sum(vals.data(), vals.size()) / (double)vals.size();
}

}  // namespace

extern "C" {

int cmoop_cnn_dataset_destroy(cmoop_cnn_dataset_handle h) {
    if (!h) return CMOOP_OK;
    cudaFree(h->x_train);
    cudaFree(h->y_train);
    cudaFree(h->x_val);
    cudaFree(h->y_val);
    delete h;
    return CMOOP_OK;
}

int cmoop_cnn_dataset_create_host(const float* x_train, const int* y_train, int n_train, const float* x_val,
                                  const int* y_val, int n_val, int height, int width, cmoop_cnn_dataset_handle* out) {
    CMOOP_REQUIRE(out && x_train && y_train && x_val && y_val, "cnn_dataset: null pointer");
    CMOOP_REQUIRE(n_train > 0 && n_val > 0 && height > 0 && width > 0, "cnn_dataset: empty split or bad shape");
    *out = nullptr;
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cmoop_cnn_dataset* d = new cmoop_cnn_dataset();
    d->n_train = n_train; d->n_val = n_val; d->H = height; d->W = width;
    const size_t img = (size_t)height * width * sizeof(float);
    if (cudaMalloc((void**)&d->x_train, img * n_train) != cudaSuccess || cudaMalloc((void**)&d->y_train, sizeof(int) * n_train) != cudaSuccess ||
        cudaMalloc((void**)&d->x_val, img * n_val) != cudaSuccess || cudaMalloc((void**)&d->y_val, sizeof(int) * n_val) != cudaSuccess) {
        cmoop::set_error("cnn_dataset: cudaMalloc failed");
        cmoop_cnn_dataset_destroy(d);
        return CMOOP_ERR_CUDA;
    }
    CMOOP_CUDA_OK(cmoop::copy_sync(d->x_train, x_train, img * n_train, cudaMemcpyHostToDevice));
    CMOOP_CUDA_OK(cmoop::copy_sync(d->y_train, y_train, sizeof(int) * n_train, cudaMemcpyHostToDevice));
    CMOOP_CUDA_OK(cmoop::copy_sync(d->x_val, x_val, img * n_val, cudaMemcpyHostToDevice));
    CMOOP_CUDA_OK(cmoop::copy_sync(d->y_val, y_val, sizeof(int) * n_val, cudaMemcpyHostToDevice));
    *out = d;
    return CMOOP_OK;
}

int cmoop_cnn_dataset_create_dev(const float* x_train_dev, const int* y_train, int n_train, const float* x_val_dev,
                                 const int* y_val, int n_val, int height, int width, void* stream,
                                 cmoop_cnn_dataset_handle* out) {
    CMOOP_REQUIRE(out && x_train_dev && y_train && x_val_dev && y_val, "cnn_dataset: null pointer");
    CMOOP_REQUIRE(n_train > 0 && n_val > 0 && height > 0 && width > 0, "cnn_dataset: empty split or bad shape");
    *out = nullptr;
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cmoop_cnn_dataset* d = new cmoop_cnn_dataset();
    d->n_train = n_train; d->n_val = n_val; d->H = height; d->W = width;
    const size_t img = (size_t)height * width * sizeof(float);
    if (cudaMalloc((void**)&d->x_train, img * n_train) != cudaSuccess || cudaMalloc((void**)&d->y_train, sizeof(int) * n_train) != cudaSuccess ||
        cudaMalloc((void**)&d->x_val, img * n_val) != cudaSuccess || cudaMalloc((void**)&d->y_val, sizeof(int) * n_val) != cudaSuccess) {
        cmoop::set_error("cnn_dataset: cudaMalloc failed");
        cmoop_cnn_dataset_destroy(d);
        return CMOOP_ERR_CUDA;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CMOOP_CUDA_OK(cmoop::copy_async(d->x_train, x_train_dev, img * n_train, cudaMemcpyDeviceToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d->x_val, x_val_dev, img * n_val, cudaMemcpyDeviceToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d->y_train, y_train, sizeof(int) * n_train, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d->y_val, y_val, sizeof(int) * n_val, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    *out = d;
    return CMOOP_OK;
}

int cmoop_profile_enable(int on) {
    if (on && !g_prof_on) g_prof_table.clear();
    g_prof_on = on != 0;
    return CMOOP_OK;
}

size_t cmoop_profile_read(char* buf, size_t cap) {
    std::string out;
    char line[256];
    for (const auto& kv : g_prof_table) {
        snprintf(line, sizeof(line), "%s %lld %.6f %.6e\n", kv.first.c_str(), kv.second.launches, kv.second.ms, kv.second.flops);
        out += line;
    }
    if (buf && cap > 0) {
        const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size() + 1;
}

double cmoop_cnn_last_device_ms(void) { return g_last_device_ms; }

int cmoop_fpr_from_predictions_host(const int* y_true, const int* y_pred, int n, int n_classes, int mode, double* fpr_out,
                                    int* confusion_out) {
    CMOOP_REQUIRE(fpr_out && (n == 0 || (y_true && y_pred)), "fpr_from_predictions: null pointer");
    CMOOP_REQUIRE(n >= 0 && n_classes >= 1 && n_classes <= 4096, "fpr_from_predictions: n=%d n_classes=%d", n, n_classes);
    CMOOP_REQUIRE(mode >= 0 && mode <= 2, "fpr_from_predictions: mode=%d (0 all classes, 1 filtered, 2 vectorised)", mode);
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const size_t b_lab = cmoop::align_up((size_t)std::max(1, n) * sizeof(int), 256);
    const size_t b_cm = (size_t)n_classes * n_classes * sizeof(int);
    char* d = (char*)cmoop::device_scratch(8, 2 * b_lab + b_cm);
    if (!d) return CMOOP_ERR_CUDA;
    int* d_true = (int*)d;
    int* d_pred = (int*)(d + b_lab);
    int* d_cm = (int*)(d + 2 * b_lab);
    if (n > 0) {
        CMOOP_CUDA_OK(cmoop::copy_async(d_true, y_true, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        CMOOP_CUDA_OK(cmoop::copy_async(d_pred, y_pred, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CMOOP_CUDA_OK(cudaMemsetAsync(d_cm, 0, b_cm, st));
    if (n > 0) {
        const int rc = Launch::confusion(d_true, d_pred, n, n_classes, d_cm, st);
        cmoop::count_launch();
        if (rc != 0) {
            cmoop::set_error("confusion kernel: %s", cudaGetErrorString((cudaError_t)rc));
            return CMOOP_ERR_CUDA;
        }
    }
    std::vector<int> cm((size_t)n_classes * n_classes);
    CMOOP_CUDA_OK(cmoop::copy_async(cm.data(), d_cm, b_cm, cudaMemcpyDeviceToHost, st));
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    *fpr_out = fpr_from_confusion(cm, n_classes, mode == 1);
    if (confusion_out) memcpy(confusion_out, cm.data(), b_cm);
    return CMOOP_OK;
}

long long cmoop_cnn_param_count(const cmoop_genotype* g, const cmoop_cnn_config* cfg) {
    if (!g || check_config(g, 1, cfg) != CMOOP_OK) return -1;
    Cand c;
    c.g = *g;
    build_units(c, *cfg, 49, 40, cfg->batch_size);
    return c.n_params;
}

int cmoop_cnn_debug_permutation(uint64_t seed, int epoch, int n, int* out) {
    CMOOP_REQUIRE(out && n >= 0, "debug_permutation: bad arguments");
    make_permutation(seed, epoch, n, out);
    return CMOOP_OK;
}

}  // extern "C"

namespace {

// host-generated shuffles: A/B switch, or a training split too large for perm_kernel's shared memory
bool host_permutations(int n_train) {
    static const bool host_perm = getenv("CMOOP_CNN_HOST_PERM") != nullptr;
    return host_perm || !Launch::perm_ok(n_train);
}

// Trains and scores one wave of candidates to completion.  The wave's candidates (8 or more) are split over `n_lanes` LANES, each with
// its own task lists and CUDA stream, stepped in lock-step by this one host thread: lane 0's forward / backward / Adam
// launches of a step are enqueued on stream 0, lane 1's on stream 1, ...  The lanes share nothing but the read-only
// dataset, so the device overlaps one lane's latency-bound small kernels and launch tails with the other lane's wide
// kernels (a single stream serialises ~350 grouped launches per step; at 32 candidates per GPU a third of them are
// bound by fixed latency, not by throughput).  Results are independent of the lane split (candidates never interact).
int run_wave(Engine& eng, Wave* lanes, int n_lanes, cudaStream_t* streams, double* out, double* history, int debug_steps,
             float* dbg_losses, float* dbg_grads, float* dbg_params) {
    const cmoop_cnn_dataset* data = eng.data;
    const cmoop_cnn_config& cfg = eng.cfg;
    const int batch = eng.batch;
    int rc = CMOOP_OK;
    // the dropout stream counts a candidate's own optimiser steps (oracle: drop_ctx=(seed, step)), so a candidate's masks
    // must not depend on which wave / lane it lands in
    for (int l = 0; l < n_lanes; ++l) {
        eng.stream = streams[l];
        if ((rc = eng.init_params(lanes[l].cands)) != CMOOP_OK) return rc;
        if ((rc = eng.build_tensor_maps(lanes[l])) != CMOOP_OK) return rc;
        if ((rc = eng.build_lists(lanes[l])) != CMOOP_OK) return rc;
    }
    const int steps_per_epoch = (data->n_train + batch - 1) / batch;
    const int val_steps = (data->n_val + batch - 1) / batch;
    int t_adam = 0, global_step = 0;
    const int max_epochs = debug_steps > 0 ? 1 : cfg.max_epochs;
    std::vector<char> lane_live(n_lanes, 1);
    for (int epoch = 0; epoch < max_epochs; ++epoch) {
        bool any_lane = false;
        for (int l = 0; l < n_lanes; ++l) {
            Wave& wv = lanes[l];
            cudaStream_t st = streams[l];
            eng.stream = st;
            bool any = false;
            for (Cand* c : wv.cands) any = any || c->active;
            lane_live[l] = any ? 1 : 0;
            if (!any) continue;
            any_lane = true;
            // per-epoch shuffles: generated on the device for every active candidate in one launch (same fmix32 stream as
            // make_permutation / cmoop_cnn_debug_permutation); a training split too large for the kernel's shared memory
            // falls back to host-generated index arrays (single lane: the staging buffer is shared)
            cudaStream_t stream = st;
            if (!host_permutations(data->n_train)) {
                CNN_LAUNCH(Launch::perm(wv.perm.d, (int)wv.perm.h.size(), epoch, data->n_train, st));
            } else {
                int n_active = 0;
                for (Cand* c : wv.cands) n_active += c->active ? 1 : 0;
                int* stage = (int*)cmoop::pinned_scratch(5, (size_t)n_active * data->n_train * sizeof(int));
                if (!stage) return CMOOP_ERR_CUDA;
                int slot = 0;
                for (Cand* c : wv.cands) {
                    if (!c->active) continue;
                    int* dst = stage + (size_t)slot++ * data->n_train;
                    make_permutation(c->seed, epoch, data->n_train, dst);
                    CMOOP_CUDA_OK(cmoop::copy_async(c->perm, dst, sizeof(int) * data->n_train, cudaMemcpyHostToDevice, st));
                }
            }
            CMOOP_CUDA_OK(cudaMemsetAsync(wv.d_acc, 0, wv.cands.size() * 12 * sizeof(double), st));
        }
        if (!any_lane) break;
        const int n_steps = debug_steps > 0 ? std::min(debug_steps, steps_per_epoch) : steps_per_epoch;
        for (int s = 0; s < n_steps; ++s) {
            const int n_b = std::min(batch, data->n_train - s * batch);
            ++t_adam;
            for (int l = 0; l < n_lanes; ++l) {
                if (!lane_live[l]) continue;
                Wave& wv = lanes[l];
                cudaStream_t st = streams[l];
                eng.stream = st;
                eng.global_step = global_step;
                if ((rc = eng.run_forward(wv, 0, s, n_b)) != CMOOP_OK) return rc;
                if ((rc = eng.run_backward(wv, s, n_b)) != CMOOP_OK) return rc;
                if (debug_steps > 0) {
                    Cand* c = wv.cands[0];
                    if (s == 0 && dbg_grads)
                        CMOOP_CUDA_OK(cmoop::copy_async(dbg_grads, c->grad, c->n_params * sizeof(float), cudaMemcpyDeviceToHost, st));
                    if (dbg_losses) {
                        double a[4];
                        CMOOP_CUDA_OK(cmoop::copy_async(a, c->acc, sizeof(a), cudaMemcpyDeviceToHost, st));
                        CMOOP_CUDA_OK(cudaStreamSynchronize(st));
                        dbg_losses[s] = (float)(a[0] / a[1]);
                        CMOOP_CUDA_OK(cudaMemsetAsync(c->acc, 0, 4 * sizeof(double), st));
                    }
                }
                if ((rc = eng.adam_step(wv, t_adam)) != CMOOP_OK) return rc;
            }
            ++global_step;
        }
        if (debug_steps > 0) {
            Cand* c = lanes[0].cands[0];
            if (dbg_params)
                CMOOP_CUDA_OK(cmoop::copy_async(dbg_params, c->p, c->n_params * sizeof(float), cudaMemcpyDeviceToHost, streams[0]));
            CMOOP_CUDA_OK(cudaStreamSynchronize(streams[0]));
            return CMOOP_OK;
        }
        for (int s = 0; s < val_steps; ++s) {
            const int n_b = std::min(batch, data->n_val - s * batch);
            for (int l = 0; l < n_lanes; ++l) {
                if (!lane_live[l]) continue;
                eng.stream = streams[l];
                eng.global_step = global_step;
                if ((rc = eng.run_forward(lanes[l], 1, s, n_b)) != CMOOP_OK) return rc;
            }
        }
        // ONE device->host copy and ONE synchronisation per epoch and lane (a lane's accumulators are contiguous); the copies of
        // all lanes are enqueued before the first wait
        std::vector<double*> h_accs(n_lanes, nullptr);
        for (int l = 0; l < n_lanes; ++l) {
            if (!lane_live[l]) continue;
            Wave& wv = lanes[l];
            h_accs[l] = (double*)cmoop::pinned_scratch(6 + 2 * l, wv.cands.size() * 12 * sizeof(double));
            if (!h_accs[l]) return CMOOP_ERR_CUDA;
            CMOOP_CUDA_OK(cmoop::copy_async(h_accs[l], wv.d_acc, wv.cands.size() * 12 * sizeof(double), cudaMemcpyDeviceToHost, streams[l]));
        }
        for (int l = 0; l < n_lanes; ++l) {
            if (!lane_live[l]) continue;
            Wave& wv = lanes[l];
            cudaStream_t st = streams[l];
            eng.stream = st;
            const double* h_acc = h_accs[l];
            CMOOP_CUDA_OK(cudaStreamSynchronize(st));
            bool changed = false;
            for (size_t ci = 0; ci < wv.cands.size(); ++ci) {
                Cand* c = wv.cands[ci];
                if (!c->active) continue;
                const double* acc = h_acc + ci * 12;
                const double train_loss = acc[0] / acc[1], val_loss = acc[4] / acc[5], val_acc = acc[6] / acc[5];
                c->epochs_run = epoch + 1;
                c->last_val_loss = val_loss;
                c->last_val_acc = val_acc;
                if (history) {
                    double* hrow = history + ((size_t)c->index * cfg.max_epochs + epoch) * 3;
                    hrow[0] = train_loss; hrow[1] = val_loss; hrow[2] = val_acc;
                }
                // keras.callbacks.EarlyStopping(monitor='val_loss', patience, restore_best_weights)
                if (cfg.restore_best_weights && !c->has_best) {
                    CMOOP_CUDA_OK(cmoop::copy_async(c->best, c->p, c->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
                    c->has_best = true;
                }
                c->wait += 1;
                if (val_loss < c->best_loss) {
                    c->best_loss = val_loss;
                    c->wait = 0;
                    if (cfg.restore_best_weights)
                        CMOOP_CUDA_OK(cmoop::copy_async(c->best, c->p, c->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
                } else if (c->wait >= cfg.patience && epoch > 0) {
                    c->active = false;
                    changed = true;
                }
            }
            if (changed) {
                bool left = false;
                for (Cand* c : wv.cands) left = left || c->active;
                if (left && (rc = eng.build_lists(wv)) != CMOOP_OK) return rc;
            }
        }
    }
    // ---- final scoring: [restore best] -> predict on the validation split -> accuracy / confusion / FPR
    const size_t cm_elems = (size_t)cfg.n_classes * cfg.n_classes;
    for (int l = 0; l < n_lanes; ++l) {
        Wave& wv = lanes[l];
        cudaStream_t st = streams[l];
        eng.stream = st;
        for (Cand* c : wv.cands) {
            c->active = true;
            if (cfg.restore_best_weights && c->has_best)
                CMOOP_CUDA_OK(cmoop::copy_async(c->p, c->best, c->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        CMOOP_CUDA_OK(cudaMemsetAsync(wv.d_acc, 0, wv.cands.size() * 12 * sizeof(double), st));
        CMOOP_CUDA_OK(cudaMemsetAsync(wv.d_cm, 0, wv.cands.size() * cm_elems * sizeof(int), st));
        if ((rc = eng.build_lists(wv)) != CMOOP_OK) return rc;
    }
    for (int s = 0; s < val_steps; ++s) {
        const int n_b = std::min(batch, data->n_val - s * batch);
        for (int l = 0; l < n_lanes; ++l) {
            eng.stream = streams[l];
            eng.global_step = global_step;
            if ((rc = eng.run_forward(lanes[l], 2, s, n_b)) != CMOOP_OK) return rc;
        }
    }
    for (int l = 0; l < n_lanes; ++l) {
        Wave& wv = lanes[l];
        cudaStream_t st = streams[l];
        double* h_acc = (double*)cmoop::pinned_scratch(6 + 2 * l, wv.cands.size() * 12 * sizeof(double));
        int* h_cm = (int*)cmoop::pinned_scratch(7 + 2 * l, wv.cands.size() * cm_elems * sizeof(int));
        if (!h_acc || !h_cm) return CMOOP_ERR_CUDA;
        CMOOP_CUDA_OK(cmoop::copy_async(h_acc, wv.d_acc, wv.cands.size() * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CMOOP_CUDA_OK(cmoop::copy_async(h_cm, wv.d_cm, wv.cands.size() * cm_elems * sizeof(int), cudaMemcpyDeviceToHost, st));
        CMOOP_CUDA_OK(cudaStreamSynchronize(st));
        std::vector<int> cm(cm_elems);
        for (size_t ci = 0; ci < wv.cands.size(); ++ci) {
            Cand* c = wv.cands[ci];
            const double* acc = h_acc + ci * 12 + 8;
            memcpy(cm.data(), h_cm + ci * cm_elems, cm_elems * sizeof(int));
            const double acc_eval = acc[2] / acc[1];
            c->final_acc = cfg.acc_from_history ? c->last_val_acc : acc_eval;
            c->fpr = fpr_from_confusion(cm, cfg.n_classes, cfg.fpr_filtered != 0);
            double* o = out + (size_t)c->index * 6;
            o[0] = c->final_acc;
            o[1] = (double)c->n_params * 4.0 / (1024.0 * 1024.0);
            o[2] = c->fpr;
            o[3] = (double)c->epochs_run;
            o[4] = c->last_val_loss;
            o[5] = c->best_loss;
        }
    }
    return CMOOP_OK;
}

int run_population(cmoop_cnn_dataset_handle data, const cmoop_genotype* genotypes, const uint64_t* seeds, int P,
                   const cmoop_cnn_config* cfg, double* out, double* history, int debug_steps, float* dbg_losses,
                   float* dbg_grads, float* dbg_params) {
    CMOOP_REQUIRE(data && genotypes && (out || debug_steps > 0), "cnn: null pointer");
    int rc = check_config(genotypes, P, cfg);
    if (rc != CMOOP_OK) return rc;
    if (P == 0) return CMOOP_OK;
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    Engine eng;
    eng.data = data;
    eng.cfg = *cfg;
    eng.batch = cfg->batch_size;
    eng.stream = cmoop::internal_stream();
    static cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    if (!ev_begin) {
        CMOOP_CUDA_OK(cudaEventCreate(&ev_begin));
        CMOOP_CUDA_OK(cudaEventCreate(&ev_end));
    }
    CMOOP_CUDA_OK(cudaEventRecord(ev_begin, eng.stream));
    std::vector<Cand> cands(P);
    for (int i = 0; i < P; ++i) {
        cands[i].g = genotypes[i];
        cands[i].seed = seeds ? seeds[i] : (uint64_t)i;
        cands[i].index = i;
        build_units(cands[i], *cfg, data->H, data->W, eng.batch);
        Arena dry;
        dry.dry = true;
        place(cands[i], dry, *cfg, data->n_train, data->n_val, eng.batch);
        cands[i].arena_bytes = dry.off;
    }
    if (history)
        for (size_t i = 0; i < (size_t)P * cfg->max_epochs * 3; ++i) history[i] = NAN;
    size_t free_b = 0, total_b = 0;
    CMOOP_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = cfg->memory_budget_bytes > 0 ? (size_t)cfg->memory_budget_bytes : (size_t)(0.6 * (double)free_b);
    // ---- waves: consecutive candidates while they fit the arena; a wave runs as up to kMaxLanes concurrent lanes
    int next = 0;
    constexpr int kMaxLanes = 4;              // compiled-in ceiling (CMOOP_CNN_LANES overrides the size-based default)
    static Wave lanes[kMaxLanes];             // task-list blobs are kept across calls (grow-only, like the arena)
    static cudaStream_t lane_streams[kMaxLanes] = {};
    static cudaEvent_t lane_fork = nullptr, lane_join[kMaxLanes] = {};
    cudaStream_t main_stream = eng.stream;
    lane_streams[0] = main_stream;
    for (int l = 1; l < kMaxLanes; ++l)
        if (!lane_streams[l]) CMOOP_CUDA_OK(cudaStreamCreateWithFlags(&lane_streams[l], cudaStreamNonBlocking));
    if (!lane_fork) {
        CMOOP_CUDA_OK(cudaEventCreateWithFlags(&lane_fork, cudaEventDisableTiming));
        for (int l = 0; l < kMaxLanes; ++l) CMOOP_CUDA_OK(cudaEventCreateWithFlags(&lane_join[l], cudaEventDisableTiming));
    }
    // CMOOP_CNN_LANES=1 switches the concurrency off (A/B); per-kernel profiling and the debug hooks need one stream
    static const int lanes_env = getenv("CMOOP_CNN_LANES") ? atoi(getenv("CMOOP_CNN_LANES")) : 0;      // 0: by wave size
    const bool one_lane = g_prof_on || debug_steps > 0 || host_permutations(data->n_train) || debug_sync();
    static char* arena_base = nullptr;        // grow-only, kept across calls (one process drives one GPU)
    static size_t arena_cap = 0;
    if (arena_cap > 0 && !(cfg->memory_budget_bytes > 0)) budget += arena_cap;   // already ours, not in the free figure
    while (next < P) {
        size_t need = 0;
        int end = next;
        while (end < P && (end == next || need + cands[end].arena_bytes <= budget)) need += cands[end++].arena_bytes;
        if (need > arena_cap) {
            if (arena_base) CMOOP_CUDA_OK(cudaFree(arena_base));
            arena_base = nullptr;
            arena_cap = 0;
            // grow with 25 % headroom (never past the budget): successive populations of a search differ a little in
            // size, and re-allocating tens of GB costs hundreds of milliseconds per call
            size_t want = need + need / 4;
            if (want > budget) want = budget > need ? budget : need;
            if (cudaMalloc((void**)&arena_base, want) != cudaSuccess) {
                (void)cudaGetLastError();
                want = need;
                if (cudaMalloc((void**)&arena_base, want) != cudaSuccess) {
                    (void)cudaGetLastError();
                    cmoop::set_error("cnn: cannot allocate a %.2f GB arena for candidates [%d,%d)", need / 1e9, next, end);
                    return CMOOP_ERR_CUDA;
                }
            }
            arena_cap = want;
        }
        Arena a;
        a.base = arena_base;
        a.cap = arena_cap;
        const size_t n_wave = (size_t)(end - next), cm_elems = (size_t)cfg->n_classes * cfg->n_classes;
        double* d_acc = (double*)cmoop::device_scratch(9, n_wave * 12 * sizeof(double));
        int* d_cm = (int*)cmoop::device_scratch(10, n_wave * cm_elems * sizeof(int));
        if (!d_acc || !d_cm) return CMOOP_ERR_CUDA;
        // lanes: the wave's candidates dealt out by descending arena footprint (a proxy of their cost) so that the lanes
        // carry similar work; at least 4 candidates per lane (11 candidates: 0.33 -> 0.29 s for two epochs, 24: 0.46 -> 0.41 s)
        // measured on B200, 2 lanes vs 1 (tools/bench_cnn.py, one epoch of 3 072 clips): 11 candidates 0.33 -> 0.29 s (two epochs),
        // 32: 0.31 -> 0.28 s, 64: 0.46 -> 0.42 s, 128: 0.84 -> 0.80 s; bench.py headline (256 candidates x 4 epochs, two runs
        // each): 43.3 -> 45.0 evals/s; true evaluations of a generation (85 candidates): 0.75-0.88 -> 0.65-0.82 s
        static const size_t lane_min = getenv("CMOOP_CNN_LANE_MIN") ? (size_t)atoi(getenv("CMOOP_CNN_LANE_MIN")) : 8;
        // 4 lanes for waves of up to 64 candidates (32: 124 -> 132 evals/s, 64: 159 -> 165 against 2 lanes), 2 above (85 / 256
        // candidates: no difference between 2 and 4 lanes)
        const int lanes_wanted = one_lane ? 1 : std::max(1, std::min(kMaxLanes, lanes_env > 0 ? lanes_env : (n_wave <= 64 ? 4 : 2)));
        const int n_lanes = n_wave >= lane_min ? std::max(1, std::min(lanes_wanted, (int)(n_wave / 4))) : 1;
        std::vector<int> order((size_t)n_wave);
        for (size_t i = 0; i < n_wave; ++i) order[i] = next + (int)i;
        if (n_lanes > 1)
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cands[x].arena_bytes > cands[y].arena_bytes; });
        for (int l = 0; l < kMaxLanes; ++l) lanes[l].cands.clear();
        std::vector<size_t> lane_load((size_t)n_lanes, 0);
        for (size_t i = 0; i < n_wave; ++i) {
            const int l = (int)(std::min_element(lane_load.begin(), lane_load.end()) - lane_load.begin());
            lanes[l].cands.push_back(&cands[order[i]]);
            lane_load[l] += cands[order[i]].arena_bytes;
        }
        size_t slot = 0;
        for (int l = 0; l < n_lanes; ++l) {
            lanes[l].lane = l;
            lanes[l].d_acc = d_acc + slot * 12;
            lanes[l].d_cm = d_cm + slot * cm_elems;
            for (Cand* c : lanes[l].cands) {
                place(*c, a, *cfg, data->n_train, data->n_val, eng.batch);
                c->acc = d_acc + slot * 12;
                c->confusion = d_cm + slot * cm_elems;
                ++slot;
            }
        }
        if (a.off > a.cap) {
            cmoop::set_error("cnn: arena overflow placing candidates [%d,%d): laid out %zu bytes, sized %zu, capacity %zu", next, end,
                             a.off, need, a.cap);
            return CMOOP_ERR_CUDA;
        }
        // fork: the lanes start after everything enqueued on the main stream so far; join: the main stream continues after them
        CMOOP_CUDA_OK(cudaEventRecord(lane_fork, main_stream));
        for (int l = 1; l < n_lanes; ++l) CMOOP_CUDA_OK(cudaStreamWaitEvent(lane_streams[l], lane_fork, 0));
        rc = run_wave(eng, lanes, n_lanes, lane_streams, out, history, debug_steps, dbg_losses, dbg_grads, dbg_params);
        eng.stream = main_stream;
        for (int l = 1; l < n_lanes; ++l) {
            cudaEventRecord(lane_join[l], lane_streams[l]);
            cudaStreamWaitEvent(main_stream, lane_join[l], 0);
        }
        if (rc != CMOOP_OK) break;
        next = end;
    }
    cudaEventRecord(ev_end, eng.stream);
    cudaStreamSynchronize(eng.stream);
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev_begin, ev_end) == cudaSuccess) g_last_device_ms = ms;
        (void)cudaGetLastError();
    }
    prof_resolve();
    if (rc != CMOOP_OK) cudaDeviceSynchronize();     // a failed wave may have left work on a lane stream
    return rc;
}

}  // namespace

extern "C" {

int cmoop_cnn_pop_train_eval(cmoop_cnn_dataset_handle data, const cmoop_genotype* genotypes, const uint64_t* seeds,
                             int n_candidates, const cmoop_cnn_config* cfg, double* out, double* history) {
    CMOOP_REQUIRE(n_candidates >= 0, "cnn: negative population");
    return run_population(data, genotypes, seeds, n_candidates, cfg, out, history, 0, nullptr, nullptr, nullptr);
}

int cmoop_cnn_debug_train_steps(cmoop_cnn_dataset_handle data, const cmoop_genotype* g, uint64_t seed,
                                const cmoop_cnn_config* cfg, int n_steps, float* losses, float* grads_first,
                                float* params_out) {
    CMOOP_REQUIRE(n_steps >= 1, "debug_train_steps: n_steps must be >= 1");
    return run_population(data, g, &seed, 1, cfg, nullptr, nullptr, n_steps, losses, grads_first, params_out);
}

int cmoop_cnn_debug_init_params(const cmoop_genotype* g, uint64_t seed, const cmoop_cnn_config* cfg, float* out) {
    CMOOP_REQUIRE(g && out, "debug_init_params: null pointer");
    int rc = check_config(g, 1, cfg);
    if (rc != CMOOP_OK) return rc;
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    Engine eng;
    cmoop_cnn_dataset dummy;
    eng.data = &dummy;
    eng.cfg = *cfg;
    eng.batch = cfg->batch_size;
    eng.stream = cmoop::internal_stream();
    Cand c;
    c.g = *g;
    c.seed = seed;
    build_units(c, *cfg, 49, 40, eng.batch);
    float* buf = nullptr;
    CMOOP_CUDA_OK(cudaMalloc((void**)&buf, (size_t)c.n_params * 4 * sizeof(float) + 12 * sizeof(double) + 1024));
    c.p = buf;
    c.grad = buf + c.n_params;
    c.m = buf + 2 * c.n_params;
    c.v = buf + 3 * c.n_params;
    c.acc = (double*)(((uintptr_t)(buf + 4 * c.n_params) + 255) / 256 * 256);
    std::vector<Cand*> one{&c};
    rc = eng.init_params(one);
    if (rc == CMOOP_OK) {
        cudaError_t e = cmoop::copy_sync(out, c.p, c.n_params * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            cmoop::set_error("debug_init_params: %s", cudaGetErrorString(e));
            rc = CMOOP_ERR_CUDA;
        }
    }
    cudaFree(buf);
    return rc;
}


// Test hook: one convolution (mode 0: forward y = conv(x, w) + b; mode 1: data gradient dx = conv^T(dy, w)) through
// the fp32 SIMT kernel (use_tc = 0) or the tcgen05 kernel (use_tc = 1).  Host pointers.
int cmoop_cnn_debug_conv(int mode, int use_tc, const float* in, const float* w, const float* bias, int n, int H, int W,
                         int Cin, int Cout, int k, int stride, int relu, float* out) {
    CMOOP_REQUIRE(in && w && out, "debug_conv: null pointer");
    CMOOP_REQUIRE(n >= 1 && n <= kBatch && (stride == 1 || (stride == 2 && k == 1)), "debug_conv: unsupported shape");
    CMOOP_REQUIRE((use_tc != 1 && use_tc != 3 && use_tc != 5) || (Cin % 16 == 0 && Cout % 16 == 0), "debug_conv: tensor-core path needs Cin, Cout multiples of 16");
    CMOOP_REQUIRE(use_tc != 3 || (Launch::tc2_ok(H, W, Cin, Cout, k, stride) && Launch::tc2_ok(H, W, Cout, Cin, k, stride)),
                  "debug_conv: shape not eligible for the patch-resident tcgen05 kernel");
    CMOOP_REQUIRE(use_tc != 2 || (mode == 0 && Launch::stem_ok(H, W, Cin, Cout, k, stride, n)),
                  "debug_conv: shape not eligible for the stem (Cin = 1) kernel");
    CMOOP_REQUIRE(use_tc != 4 || (mode == 0 && Launch::stem_ok(H, W, Cin, Cout, k, stride, n) && Launch::stem_tc_ok(H, W, Cout, k)),
                  "debug_conv: shape not eligible for the mma.sync stem kernel");
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const int pad = stride == 1 ? (k - 1) / 2 : 0;
    const int Ho = stride == 1 ? H : (H + 1) / 2, Wo = stride == 1 ? W : (W + 1) / 2;
    const long long n_x = (long long)n * H * W * Cin, n_y = (long long)n * Ho * Wo * Cout;
    const long long n_w = (long long)k * k * Cin * Cout;
    const long long n_in = mode == 0 ? n_x : n_y, n_out = mode == 0 ? n_y : n_x;
    const int gi = mode == 0 ? Cin : Cout, go = mode == 0 ? Cout : Cin;      // GEMM input / output channels
    const int K = k * k * gi, K_pad = (K + 63) / 64 * 64;
    float *d_in, *d_w, *d_out, *d_wt;
    __nv_bfloat16 *d_wb, *d_inh;
    void* d_task;
    std::vector<__nv_bfloat16> in_h((size_t)n_in);
    for (long long i = 0; i < n_in; ++i) in_h[i] = __float2bfloat16_rn(in[i]);
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_inh, n_in * 2));
    CMOOP_CUDA_OK(cmoop::copy_async(d_inh, in_h.data(), n_in * 2, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_in, n_in * 4));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_w, (n_w + Cout) * 4));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_out, n_out * 4));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_wt, n_w * 4));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_wb, std::max((size_t)go * K_pad, (size_t)Launch::tc2_weight_elems(gi, go, k)) * 2));
    CMOOP_CUDA_OK(cudaMalloc(&d_task, 1024));
    CMOOP_CUDA_OK(cmoop::copy_async(d_in, in, n_in * 4, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d_w, w, n_w * 4, cudaMemcpyHostToDevice, st));
    if (bias)
        CMOOP_CUDA_OK(cmoop::copy_async(d_w + n_w, bias, Cout * 4, cudaMemcpyHostToDevice, st));
    else
        CMOOP_CUDA_OK(cudaMemsetAsync(d_w + n_w, 0, Cout * 4, st));
    CMOOP_CUDA_OK(cudaMemsetAsync(d_out, 0, n_out * 4, st));
    const int tiles_m64 = (int)(((long long)n * Ho * Wo + 63) / 64), tiles_m128 = (int)(((long long)n * Ho * Wo + 127) / 128);
    int rc = 0;
    if (use_tc != 1 && use_tc != 3 && use_tc != 5) {
        ConvTask t{};
        t.x = d_in; t.y = d_out;
        t.Ho = Ho; t.Wo = Wo; t.k = k;
        if (mode == 0) {
            t.w = d_w; t.use_bias = 1; t.relu = relu;
            t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.stride = stride; t.pad = pad;
        } else {
            WtTask wt{};
            wt.w = d_w; wt.wt = d_wt; wt.k = k; wt.Cin = Cin; wt.Cout = Cout;
            CMOOP_CUDA_OK(cmoop::copy_async(d_task, &wt, sizeof(wt), cudaMemcpyHostToDevice, st));
            rc = Launch::wt((const WtTask*)d_task, 1, (int)((n_w + 255) / 256), st);
            CMOOP_CUDA_OK(cudaStreamSynchronize(st));
            t.w = d_wt; t.use_bias = 0;
            t.H = Ho; t.W = Wo; t.Cin = Cout; t.Cout = Cin; t.stride = 1; t.pad = stride == 1 ? pad : 0;
            if (stride == 2) { t.out_h = H; t.out_w = W; t.out_s = 2; t.accumulate = 1; }
        }
        t.tiles_n = (t.Cout + 63) / 64;
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &t, sizeof(t), cudaMemcpyHostToDevice, st));
        if (rc == 0 && use_tc == 4) {
            // bf16-only output, widened for the caller; the hook also checks the kernel's BN partial sums (64-row tiles of
            // the STORED values) against fp64 sums of what it read back
            __nv_bfloat16* d_yh = nullptr;
            float* d_stat = nullptr;
            const long long M = (long long)n * H * W;
            const int tiles = (int)((M + 63) / 64);
            CMOOP_CUDA_OK(cudaMalloc((void**)&d_yh, n_y * 2));
            CMOOP_CUDA_OK(cudaMalloc((void**)&d_stat, (size_t)tiles * 2 * Cout * 4));
            CMOOP_CUDA_OK(cudaMemsetAsync(d_stat, 0xff, (size_t)tiles * 2 * Cout * 4, st));
            t.y = nullptr; t.yh = d_yh; t.stat_part = d_stat;
            CMOOP_CUDA_OK(cmoop::copy_async(d_task, &t, sizeof(t), cudaMemcpyHostToDevice, st));
            rc = Launch::stem_conv_tc((const ConvTask*)d_task, 1, k, W, M, n, 0, st);
            std::vector<__nv_bfloat16> yh((size_t)n_y);
            std::vector<float> stat((size_t)tiles * 2 * Cout), yf((size_t)n_y);
            cudaError_t e2 = cudaStreamSynchronize(st);
            if (rc == 0 && e2 == cudaSuccess) e2 = cmoop::copy_sync(yh.data(), d_yh, n_y * 2, cudaMemcpyDeviceToHost);
            if (rc == 0 && e2 == cudaSuccess) e2 = cmoop::copy_sync(stat.data(), d_stat, stat.size() * 4, cudaMemcpyDeviceToHost);
            cudaFree(d_yh); cudaFree(d_stat);
            if (rc == 0 && e2 != cudaSuccess) rc = (int)e2;
            for (long long i = 0; i < n_y; ++i) yf[i] = __bfloat162float(yh[i]);
            for (int tl = 0; rc == 0 && tl < tiles; ++tl)
                for (int c = 0; c < Cout; ++c) {
                    double a1 = 0, a2 = 0;
                    for (long long m = 64LL * tl; m < std::min<long long>(M, 64LL * tl + 64); ++m) {
                        const double v = yf[m * Cout + c];
                        a1 += v; a2 += v * v;
                    }
                    const double g1 = stat[(size_t)tl * 2 * Cout + c], g2 = stat[(size_t)tl * 2 * Cout + Cout + c];
                    if (!(fabs(g1 - a1) <= 1e-4 * (fabs(a1) + 64.0) && fabs(g2 - a2) <= 1e-4 * (a2 + 64.0))) {
                        cmoop::set_error("debug_conv: stem BN partial sums differ (tile %d channel %d: %g vs %g, %g vs %g)", tl, c,
                                         g1, a1, g2, a2);
                        cudaFree(d_in); cudaFree(d_w); cudaFree(d_out); cudaFree(d_wt); cudaFree(d_wb); cudaFree(d_task); cudaFree(d_inh);
                        return CMOOP_ERR_CUDA;
                    }
                }
            if (rc == 0) CMOOP_CUDA_OK(cmoop::copy_async(d_out, yf.data(), n_y * 4, cudaMemcpyHostToDevice, st));
            CMOOP_CUDA_OK(cudaStreamSynchronize(st));
        } else if (rc == 0 && use_tc == 2)
            rc = Launch::stem_conv((const ConvTask*)d_task, 1, k, W, Cout, (long long)n * H * W, n, 0, st);
        else if (rc == 0)
            rc = Launch::conv((const ConvTask*)d_task, 1, tiles_m64 * t.tiles_n, n, 0, st);
    } else {
        WtBf16Task wb{};
        wb.w = d_w; wb.out = d_wb; wb.k = k; wb.Cin = Cin; wb.Cout = Cout; wb.K_pad = K_pad; wb.mode = mode + (use_tc == 3 ? 2 : 0);
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &wb, sizeof(wb), cudaMemcpyHostToDevice, st));
        rc = use_tc == 3 ? Launch::wt_bf16_v2((const WtBf16Task*)d_task, 1, (int)((Launch::tc2_weight_elems(gi, go, k) / 8 + 255) / 256), st)
                         : Launch::wt_bf16((const WtBf16Task*)d_task, 1, (int)(((long long)go * K_pad + 255) / 256), st);
        CMOOP_CUDA_OK(cudaStreamSynchronize(st));
        TcConvTask t{};
        t.xh = d_inh; t.wt = d_wb; t.y = d_out; t.k = k; t.K_pad = K_pad;
        t.Ho = Ho; t.Wo = Wo;
        if (mode == 0) {
            t.bias = d_w + n_w; t.relu = relu;
            t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.stride = stride; t.pad = pad;
        } else {
            t.H = Ho; t.W = Wo; t.Cin = Cout; t.Cout = Cin; t.stride = 1; t.pad = stride == 1 ? pad : 0;
            if (stride == 2) { t.out_h = H; t.out_w = W; t.out_s = 2; t.accumulate = 1; }
        }
        t.bn = t.Cout < 128 ? t.Cout : 128;
        t.tiles_n = (t.Cout + t.bn - 1) / t.bn;
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &t, sizeof(t), cudaMemcpyHostToDevice, st));
        if (rc == 0 && use_tc == 3) {
            // tiled-TMA descriptor of the GEMM input (the activation for mode 0, the output gradient for mode 1)
            alignas(64) char tm[kTmapBytes];
            if (Launch::make_row_tmap(tm, d_inh, t.Cin, t.W, t.H, n, t.pad) != 0) return CMOOP_ERR_CUDA;
            char* d_tm = (char*)d_task + 512;
            CMOOP_CUDA_OK(cmoop::copy_async(d_tm, tm, kTmapBytes, cudaMemcpyHostToDevice, st));
            t.tmap = d_tm;
            CMOOP_CUDA_OK(cmoop::copy_async(d_task, &t, sizeof(t), cudaMemcpyHostToDevice, st));
            CMOOP_CUDA_OK(cudaStreamSynchronize(st));
            const long long mq = (long long)n * (H + 2 * pad) * (W + 2 * pad);
            rc = Launch::conv_tc2((const TcConvTask*)d_task, 1, (int)((mq + Launch::tc2_rows() - 1) / Launch::tc2_rows()) * t.tiles_n,
                                  n, 0, Launch::tc2_q(W, k), gi, st, nullptr, t.yh != nullptr);
        } else if (rc == 0 && use_tc == 5) {             // skip_tc.cu: the 1x1 projection on mma.sync (fp32 output form)
            if (!Launch::skip_tc_ok(t)) {
                cmoop::set_error("debug_conv: shape not eligible for the mma.sync 1x1 kernel");
                return CMOOP_ERR_UNSUPPORTED;
            }
            t.tiles_n = Launch::skip_tc_tiles_n(t.Cout);
            CMOOP_CUDA_OK(cmoop::copy_async(d_task, &t, sizeof(t), cudaMemcpyHostToDevice, st));
            CMOOP_CUDA_OK(cudaStreamSynchronize(st));
            rc = Launch::skip_tc((const TcConvTask*)d_task, 1, Launch::skip_tc_tiles_m((long long)n * Ho * Wo) * t.tiles_n, n, 0,
                                 t.Cin, st);
        } else if (rc == 0) {
            rc = Launch::conv_tc((const TcConvTask*)d_task, 1, tiles_m128 * t.tiles_n, n, 0, st);
        }
    }
    cmoop::count_launch();
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == 0 && e == cudaSuccess) e = cmoop::copy_sync(out, d_out, n_out * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_out); cudaFree(d_wt); cudaFree(d_wb); cudaFree(d_task); cudaFree(d_inh);
    if (rc != 0 || e != cudaSuccess) {
        cmoop::set_error("debug_conv: %s", cudaGetErrorString(rc != 0 ? (cudaError_t)rc : e));
        return CMOOP_ERR_CUDA;
    }
    return CMOOP_OK;
}


// Test hook: weight (+bias) gradient out[K+1][Cout] of one convolution from x[n][H][W][Cin] and dy[n][Ho][Wo][Cout].
int cmoop_cnn_debug_wgrad(int use_tc, const float* x, const float* dy, int n, int H, int W, int Cin, int Cout, int k,
                          int stride, int splits, float* out) {
    CMOOP_REQUIRE(x && dy && out, "debug_wgrad: null pointer");
    CMOOP_REQUIRE(n >= 1 && n <= kBatch && (stride == 1 || (stride == 2 && k == 1)) && splits >= 1 && splits <= 32,
                  "debug_wgrad: unsupported shape");
    CMOOP_REQUIRE(use_tc != 1 || (Cin % 16 == 0 && Cout % 16 == 0), "debug_wgrad: tensor-core path needs Cin, Cout multiples of 16");
    CMOOP_REQUIRE(use_tc != 3 || Launch::wg2_ok(H, W, Cin, Cout, k, stride), "debug_wgrad: shape not eligible for wgrad_tc2");
    CMOOP_REQUIRE((use_tc != 2 && use_tc != 4) || Launch::stem_ok(H, W, Cin, Cout, k, stride, n),
                  "debug_wgrad: shape not eligible for the stem (Cin = 1) kernel");
    CMOOP_REQUIRE(use_tc != 4 || Launch::stem_tc_ok(H, W, Cout, k), "debug_wgrad: shape not eligible for the mma.sync stem kernel");
    if (!cmoop::ensure_device()) return CMOOP_ERR_CUDA;
    cudaStream_t st = cmoop::internal_stream();
    const int pad = stride == 1 ? (k - 1) / 2 : 0;
    const int Ho = stride == 1 ? H : (H + 1) / 2, Wo = stride == 1 ? W : (W + 1) / 2;
    const long long n_x = (long long)n * H * W * Cin, n_y = (long long)n * Ho * Wo * Cout;
    const int kext = k * k * Cin + 1;
    const long long n_o = (long long)kext * Cout;
    const long long M = (long long)n * Ho * Wo;
    const int gran = use_tc ? 64 : 16;
    int chunk = (int)((M + splits - 1) / splits);
    chunk = (chunk + gran - 1) / gran * gran;
    if (use_tc == 2 || use_tc == 4) {                                   // the stem kernel fixes the split size
        chunk = kStemRows;
        splits = (int)((M + chunk - 1) / chunk);
    }
    if (use_tc == 3)
        Launch::wg2_splits((long long)n * (H + 2 * pad) * (W + 2 * pad), Launch::wg2_items(Cin, Cout, k), &splits, &chunk);
    float *d_x, *d_y, *d_ws, *d_o;
    __nv_bfloat16 *d_xh, *d_yh;
    void* d_task;
    std::vector<__nv_bfloat16> xh((size_t)n_x), yh((size_t)n_y);
    for (long long i = 0; i < n_x; ++i) xh[i] = __float2bfloat16_rn(x[i]);
    for (long long i = 0; i < n_y; ++i) yh[i] = __float2bfloat16_rn(dy[i]);
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_xh, n_x * 2));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_yh, n_y * 2));
    CMOOP_CUDA_OK(cmoop::copy_async(d_xh, xh.data(), n_x * 2, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d_yh, yh.data(), n_y * 2, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_x, n_x * 4));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_y, n_y * 4));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_ws, n_o * 4 * splits));
    CMOOP_CUDA_OK(cudaMalloc((void**)&d_o, n_o * 4));
    CMOOP_CUDA_OK(cudaMalloc(&d_task, 1024));
    CMOOP_CUDA_OK(cmoop::copy_async(d_x, x, n_x * 4, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cmoop::copy_async(d_y, dy, n_y * 4, cudaMemcpyHostToDevice, st));
    CMOOP_CUDA_OK(cudaMemsetAsync(d_ws, 0xff, n_o * 4 * splits, st));       // poison: every partial must be written
    int rc;
    if (use_tc == 3) {
        TcWgradTask g{};
        g.xh = d_xh; g.dyh = d_yh; g.out = d_ws; g.H = H; g.W = W; g.Cin = Cin; g.Ho = Ho; g.Wo = Wo; g.Cout = Cout;
        g.k = k; g.stride = stride; g.pad = pad; g.splits = splits; g.m_chunk = chunk;
        g.bn = Launch::wg2_bn(Cout); g.tiles_n = Cout / g.bn;
        // tiled-TMA descriptors {x, dy, x (tail batch), dy (tail batch)}; the hook always runs the full-batch pair unless
        // n < kBatch, where both pairs describe the n samples given
        alignas(64) char tm[4 * kTmapBytes];
        for (int t = 0; t < 2; ++t)
            if (Launch::make_row_tmap(tm + (2 * t) * kTmapBytes, d_xh, Cin, W, H, n, pad) != 0 ||
                Launch::make_row_tmap(tm + (2 * t + 1) * kTmapBytes, d_yh, Cout, Wo, Ho, n, pad) != 0)
                return CMOOP_ERR_CUDA;
        char* d_tm = (char*)d_task + 512;
        CMOOP_CUDA_OK(cmoop::copy_async(d_tm, tm, sizeof(tm), cudaMemcpyHostToDevice, st));
        g.tmaps = d_tm;
        g.n_full = n;
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &g, sizeof(g), cudaMemcpyHostToDevice, st));
        CMOOP_CUDA_OK(cudaStreamSynchronize(st));
        rc = Launch::wgrad_tc2((const TcWgradTask*)d_task, 1, splits * Launch::wg2_items(Cin, Cout, k), n, Launch::wg2_q(W, k), st);
    } else if (use_tc == 1) {
        TcWgradTask g{};
        g.xh = d_xh; g.dyh = d_yh; g.out = d_ws; g.H = H; g.W = W; g.Cin = Cin; g.Ho = Ho; g.Wo = Wo; g.Cout = Cout;
        g.k = k; g.stride = stride; g.pad = pad; g.splits = splits; g.m_chunk = chunk;
        g.bn = Cout < 128 ? Cout : 128; g.tiles_k = (kext + 127) / 128; g.tiles_n = (Cout + g.bn - 1) / g.bn;
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &g, sizeof(g), cudaMemcpyHostToDevice, st));
        rc = Launch::wgrad_tc((const TcWgradTask*)d_task, 1, g.tiles_k * g.tiles_n * splits, n, st);
    } else {
        WgradTask g{};
        g.x = d_x; g.dy = d_y; g.out = d_ws; g.H = H; g.W = W; g.Cin = Cin; g.Ho = Ho; g.Wo = Wo; g.Cout = Cout;
        g.k = k; g.stride = stride; g.pad = pad; g.splits = splits; g.m_chunk = chunk;
        g.tiles_k = (kext + 63) / 64; g.tiles_n = (Cout + 63) / 64;
        if (use_tc == 4) { g.dy = nullptr; g.dyh = d_yh; }   // bf16-only output gradient (precision bf16)
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &g, sizeof(g), cudaMemcpyHostToDevice, st));
        rc = use_tc == 4 ? Launch::stem_wgrad_tc((const WgradTask*)d_task, 1, k, W, Cout, splits, n, 0, st)
           : use_tc == 2 ? Launch::stem_wgrad((const WgradTask*)d_task, 1, k, W, Cout, splits, n, 0, st)
                         : Launch::wgrad((const WgradTask*)d_task, 1, g.tiles_k * g.tiles_n * splits, n, 0, st);
    }
    cmoop::count_launch();
    CMOOP_CUDA_OK(cudaStreamSynchronize(st));
    if (rc == 0) {
        ReduceTask r{};
        r.part = d_ws; r.out = d_o; r.n = (int)n_o; r.splits = splits;
        CMOOP_CUDA_OK(cmoop::copy_async(d_task, &r, sizeof(r), cudaMemcpyHostToDevice, st));
        rc = Launch::reduce((const ReduceTask*)d_task, 1, (int)((n_o + 255) / 256), st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == 0 && e == cudaSuccess) e = cmoop::copy_sync(out, d_o, n_o * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_x); cudaFree(d_y); cudaFree(d_ws); cudaFree(d_o); cudaFree(d_task); cudaFree(d_xh); cudaFree(d_yh);
    if (rc != 0 || e != cudaSuccess) {
        cmoop::set_error("debug_wgrad: %s", cudaGetErrorString(rc != 0 ? (cudaError_t)rc : e));
        return CMOOP_ERR_CUDA;
    }
    return CMOOP_OK;
}

}  // extern "C"
