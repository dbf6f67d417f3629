/* cmoop_b200.h -- C ABI of the B200-native population-fitness hot path.
 *
 * Drop-in boundary for sumansamui/CMOOP_Audio_Processing.  The reference has no
 * FFI of its own: its seam is a set of Python functions called by the unmodified
 * drivers nsga2() / run_mobo().  Each entry point below names the reference
 * function(s) it replaces (file:line relative to the reference root); the Python
 * shims in cmoop_audio_processing_b200/ bind them with ctypes under the
 * reference's own names (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary
 *   - *_dev entry points take DEVICE pointers and a cudaStream_t (as void*), are
 *     stream-ordered and never synchronise; *_host entry points take HOST
 *     pointers, stage through library-owned device scratch and return after the
 *     result is in host memory
 *   - every function returns 0 on success or a negative cmoop_status; the
 *     message is available from cmoop_last_error().  There is no CPU fallback:
 *     without a usable sm_100 device every compute call returns CMOOP_ERR_CUDA.
 *   - one host thread per process drives the library (the reference is
 *     single-threaded); handles are not thread-safe.
 */
#ifndef CMOOP_B200_H
#define CMOOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CMOOP_OK = 0,
    CMOOP_ERR_INVALID = -1,      /* bad argument (shape, null pointer, unsupported size) */
    CMOOP_ERR_CUDA = -2,         /* CUDA runtime error or no device */
    CMOOP_ERR_UNSUPPORTED = -3,  /* configuration outside what the kernels implement */
    CMOOP_ERR_NOT_FITTED = -4
} cmoop_status;

int cmoop_abi_version(void);
const char* cmoop_last_error(void);
/* number of CUDA devices visible; 0 (not an error) on a CPU-only host */
int cmoop_device_count(void);
/* bind the calling thread (and the library's scratch / streams / arena) to a device; ONE device per process (one process
 * per GPU): once an entry point has used a device, another device is refused with CMOOP_ERR_UNSUPPORTED */
int cmoop_set_device(int device);
/* kernel launches issued by this library since load (bench.py "gpu_launches") */
uint64_t cmoop_launch_count(void);
/* host->device / device->host bytes copied by this library since load (bench.py "e2e.h2d_bytes_per_step" / "d2h_...") */
void cmoop_copy_bytes(uint64_t* h2d, uint64_t* d2h);
/* per-kernel-family device timing of the candidate-CNN engine (CUDA-event pairs around every grouped launch of its
 * stream; off by default).  enable(1) clears the table.  read() writes one line per family, "name launches ms flops"
 * (flops = algorithmic 2*M*K*N of the contraction kernels, 0 for the others), and returns the bytes needed. */
int cmoop_profile_enable(int on);
size_t cmoop_profile_read(char* buf, size_t cap);
/* device time (CUDA events on the engine's stream) of the most recent cmoop_cnn_pop_train_eval call, milliseconds */
double cmoop_cnn_last_device_ms(void);

/* ------------------------------------------------------------------ (4) NDS + crowding
 * Replaces dominates / fast_non_dominated_sort / crowding_distance
 *   nsga_penalty.py:448-524, sa_nsga_penalty.py:382-442, ablation_study/sa_nsga_local.py:240-277.
 * objs [batch][n][m] row-major fp64 RAW objectives, cv [batch][n] constraint violation.
 * Penalised objectives P = f + lam*CV are formed with a separate multiply and add
 * (no FMA) so ranks are bit-identical to CPython.
 * Outputs per problem:
 *   rank[n]           front index of every individual
 *   order[n]          individuals listed front by front in the REFERENCE'S order
 *                     (front 0 ascending; later fronts in discovery order)
 *   front_offsets[n+1] start of each front inside order[]; entries >= n_fronts hold n
 *   n_fronts[1]
 *   crowd[n]          crowding distance of every individual inside its own front
 *                     (raw objectives, stable sort, +inf at the ends)
 * crowd_mode 0: objective skipped unless (max-min) >  eps   (sa_nsga_penalty.py:437)
 * crowd_mode 1: objective skipped when   (max-min) <  eps   (nsga_penalty.py:518)
 * n <= CMOOP_NDS_MAX_N, 1 <= m <= CMOOP_NDS_MAX_M.  Any output pointer may be NULL.
 */
#define CMOOP_NDS_MAX_N 8192
#define CMOOP_NDS_MAX_M 8
size_t cmoop_nds_workspace_bytes(int n, int m, int batch);
int cmoop_nds_crowding_dev(const double* objs, const double* cv, int n, int m, int batch, double lam,
                           double eps, int crowd_mode, int* rank, int* order, int* front_offsets,
                           int* n_fronts, double* crowd, void* workspace, size_t workspace_bytes,
                           void* stream);
int cmoop_nds_crowding_host(const double* objs, const double* cv, int n, int m, int batch, double lam,
                            double eps, int crowd_mode, int* rank, int* order, int* front_offsets,
                            int* n_fronts, double* crowd);
/* crowding_distance(front, results) for an arbitrary list of distinct indices:
 * out[i] is the distance of front[i]. */
int cmoop_crowding_distance_host(const double* objs, int n, int m, const int* front, int front_len,
                                 double eps, int crowd_mode, double* out);

/* ------------------------------------------------------------------ (3) GP posterior
 * Replaces GaussianProcessRegressor.predict(X, return_std) as called from
 *   SurrogateManager.predict  ablation_study/sa_nsga_local.py:212-223, sa_nsga_penalty.py:342-363
 *   predict_gps               mobo_penalty.py:265-273
 * (sklearn/gaussian_process/_gpr.py:446-499, kernels.py Matern).  Kernel is
 * amplitude * Matern(length_scale, nu in {0.5,1.5,2.5}) [+ WhiteKernel(noise)].
 * mean_out = (K* alpha) * y_scale + y_shift ; std_out = sqrt(max(0, amp+noise - |L^-1 K*^T|^2)) * y_scale.
 */
typedef struct {
    int n_train;
    int dim;
    double amplitude;
    double length_scale;
    double nu;
    double noise;
    double y_scale;
    double y_shift;
    const double* x_train;    /* [n_train][dim] row-major, host */
    const double* alpha;      /* [n_train], host */
    const double* chol_lower; /* [n_train][n_train] row-major lower Cholesky factor, host (may be NULL: mean only) */
} cmoop_gp_model;

typedef struct cmoop_gp* cmoop_gp_handle;
/* uploads n_models models (all with the same dim); the handle owns the device copies */
int cmoop_gp_create(const cmoop_gp_model* models, int n_models, cmoop_gp_handle* out);
int cmoop_gp_destroy(cmoop_gp_handle h);
/* xq [q][dim] ; mean/std [n_models][q] ; std may be NULL */
int cmoop_gp_predict_host(cmoop_gp_handle h, const double* xq, int q, double* mean, double* std);
int cmoop_gp_predict_dev(cmoop_gp_handle h, const double* xq, int q, double* mean, double* std, void* stream);

/* ------------------------------------------------------------------ (3b) GP hyper-parameter objective
 * Replaces GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)
 * (sklearn/gaussian_process/_gpr.py, the objective fit() hands to L-BFGS-B) for the fits in
 *   SurrogateManager.update   ablation_study/sa_nsga_local.py:195-210 (4 models x 11 starts per generation)
 *   train_gps                 mobo_penalty.py:252-263
 * kind 0: C * Matern(l, nu) + WhiteKernel, theta = (log c, log l, log noise); kind 1: Matern(l, nu), theta = (log l).
 * x [n][dim] and y [n_targets][n] (already normalised as scikit-learn's y_train_) are host arrays, copied once;
 * jitter is GaussianProcessRegressor.alpha (1e-10).  A handle has `slots` independent evaluation slots, each with
 * its own stream and scratch: calls on disjoint slot ranges may be issued concurrently from different host threads.
 * lml = -inf and grad = 0 when K is not positive definite (scikit-learn's convention). */
typedef struct cmoop_gp_lml* cmoop_gp_lml_handle;
int cmoop_gp_lml_create(const double* x, int n, int dim, const double* y, int n_targets, int kind, double nu,
                        double jitter, int slots, cmoop_gp_lml_handle* out);
int cmoop_gp_lml_destroy(cmoop_gp_lml_handle h);
int cmoop_gp_lml_n_theta(cmoop_gp_lml_handle h);
/* evaluates `count` problems in slots [slot, slot + count): theta [count][n_theta], target [count] (row of y),
 * lml [count], grad [count][n_theta]; blocks the calling thread until the results are on the host */
int cmoop_gp_lml_eval(cmoop_gp_lml_handle h, int slot, int count, const double* theta, const int* target, double* lml,
                      double* grad);

/* ------------------------------------------------------------------ (4) front quality
 * Exact hypervolume for minimisation, m in {2,3}; replaces pg.hypervolume(points).compute(ref)
 * (compare.ipynb cell 0, section 5).  Points not strictly better than ref in every
 * objective contribute nothing.
 */
int cmoop_hypervolume_host(const double* points, int n, int m, const double* ref, double* out);
int cmoop_hypervolume_dev(const double* points, int n, int m, const double* ref, double* out,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t cmoop_hypervolume_workspace_bytes(int n);
/* generational_distance / inverted_gd / spread_metric of compare.ipynb sections 7-8.
 * front [nf][m], true_front [nt][m]; out[0]=GD out[1]=IGD out[2]=Spread (NaN if nf<2). */
int cmoop_front_metrics_host(const double* front, int nf, const double* true_front, int nt, int m,
                             double* out3);
/* non-dominated filter (compare.ipynb section 6 / mobo_penalty.py:478-485): mask[i]=1 if
 * no other point dominates i; also coverage C(A,B) via cmoop_coverage_host. */
int cmoop_nondominated_mask_host(const double* points, int n, int m, uint8_t* mask);
int cmoop_coverage_host(const double* a, int na, const double* b, int nb, int m, double* out);

/* ------------------------------------------------------------------ (1) log-mel / MFCC front-end
 * No reference code exists for this stage (features are loaded pre-computed:
 * nsga_penalty.py:64-71, sa_nsga_penalty.py:42-63); the spec is oracle/mfcc_ref.py.
 * n_fft == 1024 without centring and n_mels <= 64 runs the specialised warp-per-frame kernel (mfcc.cu); any other
 * power-of-two n_fft in [64, 4096] / centring / up to 256 bands runs the generic CTA-per-frame kernel.
 * wave [n_clips][n_samples] fp32 ; out [n_clips][n_frames][n_out] fp32 with
 * n_out = n_mfcc if n_mfcc > 0 else n_mels  -- the (N, T, F) layout prepare_dataset
 * expects (nsga_penalty.py:104-114).
 */
typedef struct {
    int sample_rate;   /* 16000 */
    int frame_length;  /* 640  */
    int hop;           /* 320  */
    int n_fft;         /* 1024 */
    int n_mels;        /* 40   */
    int n_mfcc;        /* 40 ; 0 = log-mel output */
    float f_min;       /* 0    */
    float f_max;       /* 8000 */
    float log_floor;   /* 1e-10 */
    int center;        /* 0: frames start at t*hop (default); 1: librosa-style centring with reflect padding
                          (needs frame_length == n_fft), e.g. BirdCLEF-shaped 32 kHz / 2048 / 512 / 128 mel */
} cmoop_mfcc_config;

typedef struct cmoop_mfcc* cmoop_mfcc_handle;
int cmoop_mfcc_create(const cmoop_mfcc_config* cfg, cmoop_mfcc_handle* out);
int cmoop_mfcc_destroy(cmoop_mfcc_handle h);
int cmoop_mfcc_n_frames(cmoop_mfcc_handle h, int n_samples);
int cmoop_mfcc_n_out(cmoop_mfcc_handle h);
int cmoop_mfcc_fwd_dev(cmoop_mfcc_handle h, const float* wave, int64_t n_clips, int n_samples, float* out,
                       void* stream);
/* host buffers (pinned or pageable); copies are chunked and overlapped with compute */
int cmoop_mfcc_fwd_host(cmoop_mfcc_handle h, const float* wave, int64_t n_clips, int n_samples, float* out);
/* 16-bit PCM waveforms (x = sample / 32768, the content of a wav file; SURVEY.md section 8a-1 "fp32 (or int16)"): same
 * features as the fp32 entry points on the widened samples, half the host->device bytes */
int cmoop_mfcc_fwd_dev_i16(cmoop_mfcc_handle h, const int16_t* wave, int64_t n_clips, int n_samples, float* out,
                           void* stream);
int cmoop_mfcc_fwd_host_i16(cmoop_mfcc_handle h, const int16_t* wave, int64_t n_clips, int n_samples, float* out);
/* optional fused per-feature standardisation (prepare_dataset, nsga_penalty.py:102-114):
 * out = (out - mean[f]) / scale[f]; pass NULL/NULL to disable. mean/scale are host fp32 [n_out]. */
int cmoop_mfcc_set_standardise(cmoop_mfcc_handle h, const float* mean, const float* scale);
/* StandardScaler statistics (prepare_dataset, nsga_penalty.py:102-114: per feature over feats.reshape(-1, F), population
 * variance) of a DEVICE-resident feature tensor [rows][n_features] fp32; fp64 two-pass, deterministic.  mean/var: host. */
int cmoop_feature_stats_dev(const float* feats, int64_t rows, int n_features, double* mean_host, double* var_host,
                            void* stream);

/* ------------------------------------------------------------------ (2) candidate-CNN train + score
 * Replaces build_model / evaluate_individual / the serial loop of compute_objectives_and_constraints
 *   nsga_penalty.py:225-442 (variant A), sa_nsga_penalty.py:137-253 (variant B),
 *   mobo_penalty.py:128-247, ablation_study/*.py copies.
 * A whole population is trained and scored in one call; every kernel launch is grouped over the
 * candidates that are still training.  The reference is unseeded; here seeds[i] fixes candidate i's
 * Glorot-uniform initialisation, its per-epoch shuffles and its dropout masks.
 * out[i] = { accuracy, size_mb, fpr, epochs_run, last_val_loss, best_val_loss }.
 * history (optional) [P][max_epochs][3] = { train_loss, val_loss, val_accuracy } per epoch, NaN after the stop.
 */
typedef struct {
    int filters;          /* 16 | 32 | 64 */
    int kernel_size;      /* 3 | 5 */
    int use_bn;
    int residual_blocks;  /* 1..3 */
    int fc_layers;        /* 1..4 */
    int use_dropout;
} cmoop_genotype;

typedef struct {
    int variant;               /* 0 = A (nsga_penalty.py / mobo_penalty.py), 1 = B (sa_nsga_*.py) */
    int n_classes;
    int batch_size;            /* 64 (nsga_penalty.py:178); at most 64 */
    int max_epochs;            /* 300 */
    int patience;              /* 5 */
    int restore_best_weights;  /* EarlyStopping(restore_best_weights=...) */
    int acc_from_history;      /* 1: history['val_accuracy'][-1] (nsga_penalty.py:384); 0: model.evaluate after restore */
    int y_true_zero;           /* 1: reproduce argmax(y_val, axis=1) == 0 of nsga_penalty.py:387 */
    int fpr_filtered;          /* 1: mean over classes with FP+TN > 0 (sa_nsga_local.py:138-141) */
    float learning_rate;       /* 1e-3 (Keras Adam default) */
    float beta1, beta2, adam_eps;   /* 0.9, 0.999, 1e-7 */
    float bn_momentum, bn_eps;      /* 0.99, 1e-3 */
    float dropout_rate;        /* 0.3 */
    int precision;             /* 0: fp32 SIMT (exact path), 1: bf16 tcgen05 implicit GEMM for Cin >= 16 convolutions */
    double memory_budget_bytes;/* activation arena per wave of candidates; 0 = 60% of free device memory */
} cmoop_cnn_config;

typedef struct cmoop_cnn_dataset* cmoop_cnn_dataset_handle;
/* x_* [n][height][width] fp32 (single channel, already standardised), y_* [n] int32 class ids; copied to the device */
int cmoop_cnn_dataset_create_host(const float* x_train, const int* y_train, int n_train, const float* x_val,
                                  const int* y_val, int n_val, int height, int width,
                                  cmoop_cnn_dataset_handle* out);
/* the same with DEVICE-resident features (e.g. straight out of cmoop_mfcc_fwd_dev): device-to-device copies on `stream`;
 * labels stay host pointers */
int cmoop_cnn_dataset_create_dev(const float* x_train_dev, const int* y_train, int n_train, const float* x_val_dev,
                                 const int* y_val, int n_val, int height, int width, void* stream,
                                 cmoop_cnn_dataset_handle* out);
int cmoop_cnn_dataset_destroy(cmoop_cnn_dataset_handle h);
long long cmoop_cnn_param_count(const cmoop_genotype* g, const cmoop_cnn_config* cfg);
/* calculate_fpr(y_true, y_pred, num_classes) on fixed label / prediction vectors (host int32 [n]); the scoring tail of
 * cmoop_cnn_pop_train_eval uses the same confusion-matrix accumulation (integer atomics on the device) and the same
 * host reduction, so this entry point pins it bit-exactly against the reference's three textual forms:
 *   mode 0  nsga_penalty.py:351-364              mean over all classes, 0.0 where FP + TN == 0
 *   mode 1  ablation_study/sa_nsga_local.py:138-141   mean over classes with FP + TN > 0 (0.0 if none)
 *   mode 2  ablation_study/init_sa_nsga_local.py:137-143  vectorised statement of mode 0 (same value)
 * The per-class rates are averaged with numpy's pairwise summation order, so the result == np.mean(fpr_vals).
 * Labels outside [0, n_classes) are dropped (sklearn confusion_matrix(labels=range(C))).  confusion_out [C][C] may be NULL. */
int cmoop_fpr_from_predictions_host(const int* y_true, const int* y_pred, int n, int n_classes, int mode, double* fpr_out,
                                    int* confusion_out);
int cmoop_cnn_pop_train_eval(cmoop_cnn_dataset_handle data, const cmoop_genotype* genotypes, const uint64_t* seeds,
                             int n_candidates, const cmoop_cnn_config* cfg, double* out, double* history);
/* test hooks: the harness-imposed random streams, so the CPU oracle can train on identical inputs */
int cmoop_cnn_debug_init_params(const cmoop_genotype* g, uint64_t seed, const cmoop_cnn_config* cfg, float* out);
int cmoop_cnn_debug_permutation(uint64_t seed, int epoch, int n, int* out);
/* run n_steps Adam steps of epoch 0 for one candidate; losses [n_steps]; grads_first / params_out [param_count] may be NULL */
int cmoop_cnn_debug_train_steps(cmoop_cnn_dataset_handle data, const cmoop_genotype* g, uint64_t seed,
                                const cmoop_cnn_config* cfg, int n_steps, float* losses, float* grads_first,
                                float* params_out);

/* one convolution through a chosen kernel; host pointers.  use_tc: 0 = generic fp32 SIMT, 1 = tcgen05 with im2col staging
 * (conv_tc.cu), 2 = dedicated Cin = 1 stem kernel (stem.cu, mode 0 only), 3 = patch-resident tcgen05 (conv_tc2.cu, stride 1),
 * 4 = mma.sync Cin = 1 stem kernel of precision bf16 (stem_tc.cu; the output is the bf16-stored one, widened; the hook also
 * verifies the kernel's BN partial sums).
 * mode 0: out[n][Ho][Wo][Cout] = conv(in[n][H][W][Cin], w[k][k][Cin][Cout]) + bias (optional ReLU)
 * mode 1: out[n][H][W][Cin] = data gradient of that convolution for in = dy[n][Ho][Wo][Cout] */
int cmoop_cnn_debug_conv(int mode, int use_tc, const float* in, const float* w, const float* bias, int n, int H,
                         int W, int Cin, int Cout, int k, int stride, int relu, float* out);

/* weight (+ bias, last row) gradient out[k*k*Cin + 1][Cout] of one convolution; `splits` deterministic split-M partials
 * (use_tc as above; 2, 3 and 4 choose their own split geometry: stem_wgrad_kernel / wgrad_tc2_kernel / stem_wgrad_tc_kernel;
 * 4 reads dy rounded to bf16) */
int cmoop_cnn_debug_wgrad(int use_tc, const float* x, const float* dy, int n, int H, int W, int Cin, int Cout, int k,
                          int stride, int splits, float* out);

#ifdef __cplusplus
}
#endif
#endif /* CMOOP_B200_H */
