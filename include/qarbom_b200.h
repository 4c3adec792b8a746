/*
 * qarbom_b200 -- C ABI of the B200-native classical RBM training hot path
 * (CD-k / PCD Gibbs chains, positive/negative-phase gradient, parameter update,
 * reconstruction-error and free-energy evaluation) of NITeQ/QARBoM.jl.
 *
 * The reference has NO FFI for this path: it is reached through Julia multiple dispatch
 * only.  Each entry point below therefore cites the reference *function* it replaces
 * (paths relative to the reference repository); INTEGRATION.md shows the `ccall` stubs
 * a maintainer adds on the Julia side.
 *
 * Conventions
 *   - every function returns 0 on success, a negative QB_E_* code on failure; the
 *     message is available (per calling thread) from qb_last_error();
 *   - host buffers are borrowed for the duration of the call only; all calls on one
 *     handle must be serialised by the caller; calls block until the device is done;
 *   - parameter arrays use the Julia layout: W is column-major V x H (W[i + V*j]),
 *     U is column-major L x H; datasets are N rows of V (resp. L) contiguous elements
 *     (Vector{Vector{T}} flattened), element type selected by a QB_DT_* code;
 *   - there is no CPU fallback: every entry point that computes fails with
 *     QB_E_CUDA when no sm_100 device is usable.
 */
#ifndef QARBOM_B200_H
#define QARBOM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_VERSION 100

/* status codes */
#define QB_OK 0
#define QB_E_ARG (-1)      /* invalid argument / shape */
#define QB_E_CUDA (-2)     /* CUDA runtime / driver failure, or no usable device */
#define QB_E_STATE (-3)    /* call sequence error (no data, no chains, ...) */
#define QB_E_NCCL (-4)     /* NCCL unavailable or failed */
#define QB_E_UNSUPPORTED (-5)

/* visible unit kind: RBM / RBMClassifier vs GRBM / GRBMClassifier (src/rbm.jl:2-39) */
#define QB_VISIBLE_BERNOULLI 0
#define QB_VISIBLE_GAUSSIAN 1

/* arithmetic mode */
#define QB_PREC_FP32 0  /* fp32 CUDA-core kernels: validation mode (1e-5 rel. vs Float64) */
#define QB_PREC_BF16 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate: production */
#define QB_PREC_FP32X3 2 /* validation mode ON the tensor cores: every fp32 operand is split into three bf16 components
                          * (x = hi + mid + lo, 24 mantissa bits), the component products that matter (i + j <= 2) accumulate
                          * into one fp32 TMEM tile; everything around the contractions is the QB_PREC_FP32 path.  Same
                          * tolerances as QB_PREC_FP32, 3-6x the MMA work of QB_PREC_BF16. */

/* host element types for datasets */
#define QB_DT_F64 0
#define QB_DT_I64 1
#define QB_DT_F32 2
#define QB_DT_U8 3

typedef struct qb_handle qb_handle;

typedef struct qb_config {
    int64_t n_visible;     /* V */
    int64_t n_hidden;      /* H */
    int64_t n_classifiers; /* L, 0 for RBM / GRBM */
    int32_t visible_kind;  /* QB_VISIBLE_* */
    int32_t precision;     /* QB_PREC_* */
    int64_t max_batch;     /* largest number of LOCAL rows (mini-batch rows == chains) per step */
    int32_t device;        /* CUDA device ordinal */
    int32_t rank;          /* data-parallel rank of this handle (0 on one GPU) */
    int32_t world;         /* number of data-parallel ranks (1 on one GPU) */
    int32_t reserved;
} qb_config;

int32_t qb_version(void);
int32_t qb_device_count(int32_t *count);
const char *qb_last_error(void);

/* RBM(...) / GRBM(...) / RBMClassifier(...) constructors own the host arrays
 * (src/rbm.jl:43-99); the handle owns their device mirror. */
int32_t qb_create(qb_handle **out, const qb_config *cfg);
int32_t qb_destroy(qb_handle *h);

/* rbm.W / rbm.a / rbm.b / rbm.U / rbm.c upload and download (fields of src/rbm.jl:2-39;
 * the download is what copy_rbm / copy_rbm! snapshot, src/rbm.jl:231-255).
 * U and c are NULL when n_classifiers == 0. */
int32_t qb_set_params(qb_handle *h, const double *W, const double *a, const double *b, const double *U, const double *c);
int32_t qb_get_params(qb_handle *h, double *W, double *a, double *b, double *U, double *c);

/* best_rbm = copy_rbm(rbm) / copy_rbm!(rbm, best_rbm) and the final copy_rbm!(best_rbm, rbm)
 * (src/rbm.jl:231-255; src/training/train_cd.jl:60,99-102,109-111) kept in HBM: the epoch loop snapshots
 * and restores without moving the parameters to the host. */
int32_t qb_snapshot_params(qb_handle *h);
int32_t qb_restore_params(qb_handle *h);

/* Extension (absent from update_rbm!, src/rbm.jl:152-163): both default to 0, which
 * reproduces the reference update exactly. */
int32_t qb_set_optimizer(qb_handle *h, double momentum, double weight_decay);

/* x_train [, label_train] of train! (src/training/train_cd.jl:45-59,
 * src/training/train_pcd.jl:133-148): the LOCAL rows of the dataset become resident in HBM.
 * With world > 1 the caller passes, for every global mini-batch, this rank's contiguous
 * slice of rows, concatenated in mini-batch order. Y may be NULL when n_classifiers == 0. */
int32_t qb_set_data(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t n_rows);

/* _init_fantasy_data (src/fantasy_data.jl:66-86): n_chains LOCAL persistent chains.  Host
 * arrays (n_chains rows of V / H / L doubles: the reference's rand(V), rand(H), rand(L))
 * or NULL to draw the same U[0,1) floats from the device Philox stream. */
int32_t qb_init_chains(qb_handle *h, int64_t n_chains, const double *v, const double *hid, const double *y);
int32_t qb_get_chains(qb_handle *h, double *v, double *hid, double *y);
/* The same for the LOCAL chains [row0, row0 + n_rows): FantasyData[i].v / .h / .y (src/fantasy_data.jl:1-10) of a
 * few chains without moving all of them (parity checks of row subsets at full problem sizes). */
int32_t qb_get_chain_rows(qb_handle *h, int64_t row0, int64_t n_rows, double *v, double *hid, double *y);

/* Production RNG: Philox4x32-10 keyed by `seed`, counters keyed by (draw, global row,
 * unit) so results do not depend on the number of GPUs. */
int32_t qb_seed(qb_handle *h, uint64_t seed);

/* Validation mode: replace the random stream of the NEXT step call by injected values,
 * dense per Gibbs step: U_h[k][B][H], U_v[k][B][V] (Bernoulli visibles) or Z_v[k][B][V]
 * (standard normals, Gaussian visibles), U_y[k][B][L].  This is the de-interleaved form
 * of the serial rand()/randn() order of src/gibbs.jl:57-72 and src/fantasy_data.jl:12-29
 * (SURVEY.md section 8(a)).  B = local rows. Unused arrays are NULL. */
int32_t qb_inject_streams(qb_handle *h, const double *U_h, const double *U_v, const double *U_y, const double *Z_v,
                          int64_t k, int64_t B);

/* One mini-batch of persistent_contrastive_divergence! (src/training/train_pcd.jl:11-40;
 * classifier :57-92) on resident mini-batch `batch_index` with B local rows.
 * The number of chains is whatever qb_init_chains set: == B is the reference; != B is a declared
 * extension (all chains advance, negative statistics are averaged over all of them).
 * k == 0 leaves the chains untouched and uses them as externally supplied negative samples: with
 * qb_init_chains(1, v_model, h_model, y_model) this is one mini-batch of persistent_qubo!
 * (src/training/train_quantum_pcd.jl:10-37, :55-83) -- the sampler call itself stays in Julia. */
int32_t qb_pcd_step(qb_handle *h, int64_t batch_index, int64_t B, int64_t k, double lr, double llr);
/* One mini-batch of contrastive_divergence! (src/training/train_cd.jl:4-19, :25-41).
 * B == 1 is the reference (online); B > 1 is the declared mini-batch extension. */
int32_t qb_cd_step(qb_handle *h, int64_t batch_index, int64_t B, int64_t k, double lr, double llr);

/* persistent_contrastive_divergence! / contrastive_divergence! for a whole epoch over the
 * resident data (`_set_mini_batches` semantics, src/utils.jl:13-26: trailing remainder
 * dropped).  times[3] = seconds in (sample, gibbs, update), the triple the reference
 * returns to train! (src/training/train_pcd.jl:42), measured with CUDA events. */
int32_t qb_pcd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double llr, double times[3]);
int32_t qb_cd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double llr, double times[3]);
/* n_steps consecutive mini-batch steps of the same two functions on the resident mini-batches first_batch, first_batch + 1,
 * ... (wrapping around after the last full mini-batch); an epoch is first_batch = 0, n_steps = N / B.  pcd != 0:
 * persistent_contrastive_divergence!, else contrastive_divergence!.  On small models (bf16 mode, one GPU, no momentum) all
 * the steps run inside ONE persistent kernel launch (sm100_epoch.cuh); the results are those of n_steps single-step calls. */
int32_t qb_train_steps(qb_handle *h, int32_t pcd, int64_t first_batch, int64_t n_steps, int64_t B, int64_t k, double lr,
                       double llr);

/* One epoch of fast_persistent_contrastive_divergence! (src/training/train_fast_pcd.jl:1-57; classifiers :59-127):
 * inside a mini-batch the slow weights are updated and the fast weights decayed (19/20) after every SAMPLE, then
 * the chains advance with W + W_fast (U + U_fast, ...; src/fantasy_data.jl:31-64).  Fast weights restart at zero
 * every call, as in the reference.  llr / fast_llr (label_learning_rate, fast_label_learning_rate) are used by
 * classifiers only.  Reference quirks kept: (1) classifiers pair EVERY sample of a mini-batch with chain 1
 * (`batch_index` is never advanced, :77-119) -- option "fast_pcd_pair_chains" = 1 pairs sample i with chain i;
 * (2) on Gaussian visibles the fast chain step thresholds the noisy value, v = 1[u < mean + z]
 * (src/gibbs.jl:22-25 + src/rbm.jl:188-189), so the chains turn binary.  Philox stream only.
 * times[3] = (0, chain seconds, online sample+update seconds). */
int32_t qb_fast_pcd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double fast_lr, double llr, double fast_llr,
                          double times[3]);

/* Same step, but the mini-batch comes from HOST memory on every call (end-to-end path:
 * upload, step, and a device->host read of the batch reconstruction error sum: sum over the batch rows and the
 * visible units of (x - reconstruct(x))^2, src/rbm.jl:220-224.  PCD on binary models without label units and with
 * chains == batch rows takes it inside the step from the positive-phase p(h | x) it computes anyway, i.e. with the
 * weights BEFORE the update; every other case reconstructs after the step, with the updated weights). */
int32_t qb_pcd_step_host(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t B,
                         int64_t k, double lr, double llr, double *recon_sq_err);
int32_t qb_cd_step_host(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t B,
                        int64_t k, double lr, double llr, double *recon_sq_err);

/* Pipelining hint for the two calls above: X/Y of the NEXT qb_*_step_host call.  Its upload starts on
 * a copy stream inside the next step call and overlaps that step's kernels; the call after that, given
 * the same X / Y pointers, element types and B, consumes it without a fresh copy (any mismatch falls back to a fresh
 * upload).  The hinted host buffers (pinned memory for a truly asynchronous copy) must stay valid AND UNCHANGED from the
 * step call that starts their upload until the call that consumes them returns: the copy engine reads them in between,
 * a buffer refilled in that window trains on a mix of old and new contents. */
int32_t qb_set_next_host_batch(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype);

/* The epoch loop of train! inside the library (src/training/train_cd.jl:60-115, train_pcd.jl:149-213,
 * train_fast_pcd.jl:185-250 and the classifier variants): initial_evaluation, then per epoch the
 * training epoch with learning_rate[epoch], evaluate on every requested metric, _diverged on the stopping metric
 * (src/evaluation.jl:132-168), patience / early stopping, best-model snapshot in HBM (copy_rbm!), and at the end
 * copy_rbm!(best_rbm, rbm) -- no host round trip of the parameters and no caller-side loop.  The resident dataset
 * (qb_set_data) is the training set; chains are initialised here for PCD / FastPCD (_init_fantasy_data(rbm, batch_size)).
 * history[(n_epochs + 1) x n_metrics], row-major: row 0 = initial_evaluation, row e = after epoch e (rows after an early
 * stop stay untouched) -- merge_metrics' columns.  times[3] = totals of (sample, gibbs, update) seconds.
 * Metrics: MeanSquaredError / FreeEnergy work on every model; Accuracy and CrossEntropy (reference-literal: computed from the
 * rounded prediction, hence NaN) need label units and the labels of the evaluated rows in Y_eval.  X_eval == NULL evaluates
 * the resident training set (Y_eval then holds the training labels); the reference uses the test set for classifiers only
 * when both arrays are given (train_cd.jl:143-147) -- that rule is the caller's.
 * on_epoch (optional) is called after every epoch from the calling thread with the epoch number, that epoch's metric row,
 * its (sample, gibbs, update) seconds and QB_EPOCH_* flags: the place for _log_epoch / _log_metrics. */
#define QB_METHOD_CD 0
#define QB_METHOD_PCD 1
#define QB_METHOD_FAST_PCD 2
#define QB_METRIC_MSE 0
#define QB_METRIC_ACCURACY 1
#define QB_METRIC_CROSS_ENTROPY 2
#define QB_METRIC_FREE_ENERGY 3
#define QB_EPOCH_DIVERGED 1      /* _diverged(metrics_dict, stopping_metric) was true */
#define QB_EPOCH_EARLY_STOP 2    /* ... and patience ran out: this was the last epoch */
#define QB_EPOCH_NEW_BEST 4      /* best_rbm was replaced by the current parameters */
typedef struct qb_train_config {
    int32_t method;                     /* QB_METHOD_* */
    int32_t n_epochs;
    int64_t batch_size, gibbs_steps;
    const double *learning_rate;        /* [n_epochs] */
    const double *label_learning_rate;  /* [n_epochs]; NULL for models without label units */
    double fast_learning_rate, fast_label_learning_rate;   /* FastPCD only */
    int32_t n_metrics; const int32_t *metrics;              /* QB_METRIC_*, in reporting order */
    int32_t stopping_metric;            /* must be one of `metrics` (the reference indexes metrics_dict with it) */
    int32_t early_stopping, store_best_rbm, patience;
    const void *X_eval; int32_t x_dtype; int64_t n_eval;    /* NULL / 0: evaluate the resident training set */
    const void *Y_eval; int32_t y_dtype;                    /* labels of the evaluated rows (Accuracy / CrossEntropy) */
    void (*on_epoch)(void *user, int32_t epoch, const double *metrics, const double times[3], int32_t flags);
    void *user;
} qb_train_config;
int32_t qb_train(qb_handle *h, const qb_train_config *cfg, double *history, double times[3], int32_t *epochs_run);

/* evaluate + _evaluate(MeanSquaredError) + reconstruct (src/evaluation.jl:16-20,52-70;
 * src/rbm.jl:220-224): sum over samples of sum_units (x - reconstruct(x))^2 / n_rows.
 * X == NULL evaluates the resident dataset. Z (n_rows x V standard normals) optionally
 * injects the noise the GRBM reconstruct draws (src/rbm.jl:187). */
int32_t qb_eval_mse(qb_handle *h, const void *X, int32_t x_dtype, int64_t n_rows, const double *Z, double *out);
/* New metric: mean free energy, sign convention of the QUBO objective (src/qubo.jl:43-55). */
int32_t qb_eval_free_energy(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype,
                            int64_t n_rows, double *out);
/* reconstruct(rbm, v) for n_rows rows (src/rbm.jl:220-224). */
int32_t qb_reconstruct(qb_handle *h, const double *v, int64_t n_rows, const double *Z, double *out);
/* conditional_prob_h (src/rbm.jl:165,170-171) for n_rows rows; y may be NULL (2-arg form). */
int32_t qb_conditional_prob_h(qb_handle *h, const double *v, const double *y, int64_t n_rows, double *out);

/* Single sampling half-steps, for unit tests and callers that drive their own chains.
 * qb_sample_hidden = gibbs_sample_hidden (src/gibbs.jl:3-6; with y: :29-32): out[n_rows][H] = 1[u < p(h | v[, y])].
 * qb_sample_visible = gibbs_sample_visible (src/gibbs.jl:13-16; Gaussian visibles :18-20: mean + z, a real value) and,
 * for classifiers, gibbs_sample_label (src/gibbs.jl:46-49) into out_y (may be NULL).
 * U / R / Ry are the caller's random numbers, one per unit in unit order (R: uniforms, or standard normals for
 * Gaussian visibles; Ry: label uniforms, required with R on a classifier); NULL draws from the Philox stream
 * (one draw id per call, row index = global row).  In bf16 mode a Gaussian sample comes back rounded to bf16. */
int32_t qb_sample_hidden(qb_handle *h, const double *v, const double *y, int64_t n_rows, const double *U, double *out);
int32_t qb_sample_visible(qb_handle *h, const double *hid, int64_t n_rows, const double *R, const double *Ry,
                          double *out_v, double *out_y);

/* conditional_prob_y_given_v (src/rbm.jl:191-194) for n_rows rows -> out[n_rows][L]; X == NULL uses the
 * resident dataset.  `classify` (src/rbm.jl:226-229) and the Accuracy metric (src/evaluation.jl:9-14) are
 * the caller's arg-max / comparison over these probabilities. */
int32_t qb_conditional_prob_y_given_v(qb_handle *h, const void *X, int32_t x_dtype, int64_t n_rows, double *out);

/* Data-parallel gradient exchange over NCCL (new; the reference is single-process).
 * Rank 0 creates the id, the caller ships the 128 bytes to the other ranks.  fp32 mode: one allreduce of the flat
 * gradient per step, replicated masters.  bf16 mode with n_hidden % world == 0 (option "sharded_update", default 1):
 * reduce-scatter, update of the owned hidden rows, all-gather of the bf16 shadow; the fp32 master is then sharded and
 * qb_get_params / qb_snapshot_params gather it -- they are collective, call them on every rank at the same point. */
int32_t qb_comm_unique_id(void *id128);
int32_t qb_comm_init(qb_handle *h, const void *id128);

/* Optional NVLink peer-memory path (bf16 mode, 2..8 ranks on one node): after qb_comm_init (still used for nothing
 * but kept as the fallback) every rank exports 5 CUDA IPC handles (320 bytes), the caller all-gathers them
 * (world x 320 bytes, rank-major) and every rank imports the lot.  From then on the gradient exchange is ONE
 * kernel per rank: reduce-scatter by P2P loads, update of the owned shard of the fp32 master, all-gather of the new
 * bf16 shadows by P2P stores.  The master is sharded by hidden rows; qb_get_params / qb_snapshot_params gather it and
 * must therefore be called by all ranks at the same point of the training loop. */
int32_t qb_p2p_export(qb_handle *h, void *handles320);
int32_t qb_p2p_import(qb_handle *h, const void *all_handles);
/* milliseconds accumulated inside the fused exchange kernels since the last call: waiting for the peers' gradients,
 * and reduce + update + push */
int32_t qb_get_p2p_timing(qb_handle *h, double *wait_ms, double *work_ms);

/* Introspection for benchmarks.  Counters since the last reset: kernels launched by this
 * library on this handle, GEMM launches and their algorithmic flops; with option
 * "time_gemms" = 1 every GEMM launch is bracketed by CUDA events on the launching stream and
 * timed_* report the number of launches timed, their summed duration and flops. */
int32_t qb_get_counters(qb_handle *h, int64_t *kernel_launches, int64_t *gemm_launches, double *gemm_flops,
                        int64_t *timed_gemms, double *timed_gemm_ms, double *timed_gemm_flops);
/* GEMM launches since the last reset by kernel specialisation: counts[(pair * 4 + b) * QB_GEMM_KINDS + kind] with
 * pair = 1 for the cta_group::2 kernel, b = log2(tile width / 32) (32, 64, 128, 256), kind = QB_GEMM_KIND_*; the
 * last slot counts launches of the fp32 CUDA-core GEMM.  Tests use it to assert WHICH kernel produced a result. */
#define QB_GEMM_KINDS 11
#define QB_GEMM_HIST_SLOTS (2 * 4 * QB_GEMM_KINDS + 1)
#define QB_GEMM_KIND_GENERIC 0
#define QB_GEMM_KIND_GRAD 1
#define QB_GEMM_KIND_FAST_SAMPLE 2
#define QB_GEMM_KIND_FAST_PROB 3
#define QB_GEMM_KIND_FAST_SAMPLE_PROB 4
#define QB_GEMM_KIND_FAST_GAUSS 5  /* Gaussian visibles: mean + Box-Muller normal in the epilogue */
#define QB_GEMM_KIND_FAST_LABEL 6  /* Bernoulli visibles + label units (softmax + Bernoulli) in the epilogue */
#define QB_GEMM_KIND_CHAIN 7       /* persistent multi-half-step kernel: all half-steps of a chain advance / CD forward pass */
#define QB_GEMM_KIND_EPOCH 8       /* persistent multi-mini-batch kernel: whole training steps incl. gradient + update */
#define QB_GEMM_KIND_GRAD_X3 9     /* gradient-store kernel with more than two products (the component products of QB_PREC_FP32X3) */
#define QB_GEMM_KIND_FAST_SQERR 10 /* reconstruction error reduced in the epilogue, nothing stored (qb_eval_mse, qb_*_step_host) */
int32_t qb_get_gemm_histogram(qb_handle *h, int64_t *counts, int32_t n_slots);
/* Tensor-core flops ISSUED since the last reset: equal to the algorithmic gemm_flops of qb_get_counters except in
 * QB_PREC_FP32X3, where every contraction issues 3..6 bf16 component products (reported separately). */
int32_t qb_get_issued_flops(qb_handle *h, double *issued_flops);
/* with "time_gemms": number and summed CUDA-event duration of the update_w kernel launches */
int32_t qb_get_update_timing(qb_handle *h, int64_t *timed_updates, double *timed_update_ms);
int32_t qb_reset_counters(qb_handle *h);
/* Options: "async_steps" (qb_pcd_step / qb_cd_step return without waiting; use qb_sync), "time_gemms",
 * "stale_chains" (declared extension: exchange + update of step t overlap the chain advance of step t + 1; collective
 * when a communicator exists: it splits off a CTA-limited communicator), "sharded_update" (multi-GPU bf16 exchange:
 * 1 = reduce-scatter + shard update + bf16 all-gather, the default when n_hidden % world == 0; 0 = allreduce),
 * "state_exchange" (multi-GPU PCD on binary models without labels, exact: positive product reduce-scattered behind the
 * chain advance + bit-packed chain-state all-gather; collective, measured neutral, off by default),
 * "online_kernel" (0: batch-size-1 CD through per-sample launches instead of the cooperative epoch kernel),
 * "online_wpb" (warps per block of the cooperative kernels), "fast_pcd_pair_chains" (classifier FastPCD pairs
 * sample i with chain i instead of the reference's chain 1), "p2p_blocks_per_sm",
 * "nvls_exchange" (= 1, COLLECTIVE, after qb_comm_init: the gradient exchange becomes ONE kernel per rank over NVSwitch
 * multicast addresses -- multimem.ld_reduce of the fp32 gradient shard, update of the owned hidden rows, multimem.st of the
 * new bf16 shadow; needs NCCL >= 2.28 at run time (symmetric windows) and multicast-capable GPUs, else QB_E_UNSUPPORTED with
 * the reason and the NCCL exchange stays),
 * "split_exchange" (on the NVLS path, binary PCD without label units: the gradient is exchanged as its two products -- the
 * positive product is computed before the chain advance and reduced behind it, the negative product travels as exact uint16
 * counts summed as packed u64 by the switch; 0 never (default: measured slower), 1 when a half-step
 * leaves SMs idle, 2 whenever possible; same results as the unsplit exchange up to fp32 summation order),
 * "fused_chain" (the half-steps of a step as one persistent launch: 0 never, 1 automatic, 2 always), "chain_bn" (its tile
 * width, 0 = automatic), "chain_pairs" (experiments: CTA pairs it may use, 0 = all), "chain_quad" (its 4-CTA-cluster variant with multicast weights: 0 never (default: measured no faster),
 * 1 one-wave shapes, 2 whenever possible), "epoch_kernel" (many mini-batch steps of a small model as one persistent launch
 * in qb_train_steps / the epoch calls: 0 never, 1 automatic, 2 always),
 * and for experiments / tests "force_bn" (GEMM tile width) and "use_2cta" (0 never, 1 automatic, 2 always). */
int32_t qb_set_option(qb_handle *h, const char *key, double value);
int32_t qb_sync(qb_handle *h);
/* CUDA-event stopwatch on the handle's stream (the stream every kernel is launched on). */
int32_t qb_timer_start(qb_handle *h);
int32_t qb_timer_stop(qb_handle *h, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* QARBOM_B200_H */
