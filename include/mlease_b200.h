/*
 * mlease_b200.h -- C ABI of the B200-native ADMM logistic-regression hot path.
 *
 * Drop-in boundary for ONE path of linkedin/ml-ease: the body of
 * RegressionAdmmTrain.run() (jobs/RegressionAdmmTrain.java:278-501), the reducer it drives
 * (AdmmReducer.reduce, :642-718 -> LibLinear.train, llf/LibLinear.java:200-208), the scoring
 * of RegressionTest/RegressionTestLoglik and the per-key fits of RegressionNaiveTrain.
 * The reference has no FFI seam (it is 100% Java); these are the entry points a JNI shim
 *  binds.  Paths cited below are relative to
 * /root/reference/src/main/java/com/linkedin/mlease/ unless they start with bw/ (= de/bwaldvogel/liblinear/).
 *
 * Conventions: every function returns 0 on success, non-zero on error (message via
 * mlease_last_error(), thread-local).  Plain pointers and sizes only.  "host-or-device"
 * pointers may be either (UVA); everything else says which.  Feature ids are the GLOBAL
 * dictionary 0..num_features-1; the intercept "(INTERCEPT)" is index num_features (last), the
 * bias column the reference appends to every row (regression/liblinearfunc/LibLinearDataset.java:592-614).
 * One host thread drives a session (the reference's driver and reducers are single threaded).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MLEASE_B200_H
#define MLEASE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlease_session mlease_session;
typedef struct mlease_comm mlease_comm;     /* one NCCL communicator rank (multi-GPU jobs) */
typedef struct mlease_world mlease_world;   /* N sessions on N GPUs of this process */

#define MLEASE_OK 0
#define MLEASE_ERR_INVALID 1   /* bad argument / config (reference: IOException in the job) */
#define MLEASE_ERR_CUDA 2      /* CUDA failure (reference: IOException("Model fitting error!"), jobs/RegressionAdmmTrain.java:713-716) */
#define MLEASE_ERR_NUMERIC 3   /* a fit did not converge / Hessian not SPD (same mapping) */
#define MLEASE_ERR_STATE 4     /* call order (e.g. partitions missing: RuntimeException("Some models failed!"), utils/LinearModelUtils.java:80-83) */

const char* mlease_last_error(void);
int mlease_abi_version(void);

/* ---------------------------------------------------------------------------------------
 * Session = one RegressionAdmmTrain job on one GPU (partitions are sharded over sessions; the
 * all-reduce between them is the library's, see "Multi-GPU" below, or a caller-supplied callback).
 * Config keys mirrored (jobs/RegressionAdmmTrain.java:78-122,138-185):
 *   num.blocks, lambda (list, Float.parseFloat), rho (list or NULL -> 1 if lambda<=100 else 10),
 *   regularizer (2 = L2 z-update :377-404, 1 = L1 thresholded z-update :406-451, anything else -> "Only L1 and L2
 *   regularization supported!" :144-147), penalize.intercept, epsilon, rho.adapt.coefficient,
 *   aggressive.liblinear.epsilon.decay (pure control: only moves the stop rule :493-496).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int32_t device;              /* CUDA ordinal */
  int32_t num_blocks;          /* P, partitions over ALL processes */
  int32_t num_features;        /* global dictionary size, intercept excluded */
  int32_t num_lambdas;         /* L */
  const float* lambdas;        /* [L] host */
  const float* rhos;           /* [L] host or NULL */
  const float* lambda_map;     /* [num_features] host or NULL; >0 entries override lambda per feature (:382-386) */
  int32_t regularizer;         /* 1 or 2 */
  int32_t penalize_intercept;  /* default 0 */
  int32_t aggressive_decay;    /* default 0 */
  int32_t binary_feature;      /* binary.feature: ignore values, use 1 (regression/liblinearfunc/LibLinearBinaryDataset.java) */
  double epsilon;              /* outer stop (:473); < 0 -> default 1e-4, 0 = never stop early */
  float rho_adapt_coefficient; /* default 0 (:323-327) */
  /* solver knobs (no reference equivalent: the inner solve is exact Newton, not TRON) */
  double newton_xtol;          /* stop when |dir|_inf <= xtol*max(|beta|_inf,1e-2); 0 -> 2e-7 (float32 lattice of the data path) */
  int32_t max_newton;          /* max accepted Newton steps per x-update; 0 -> 50 */
  int32_t hessian_policy;      /* 0 adaptive chord (refresh when contraction is poor), 1 every step */
  void* stream;                /* cudaStream_t to run on (NULL = legacy default stream) */
} mlease_admm_config;

int mlease_session_create(const mlease_admm_config* cfg, mlease_session** out);
int mlease_session_destroy(mlease_session* s);

/* Replaces the per-iteration re-ingest LibLinearDataset.addInstanceAvro/finish
 * (regression/liblinearfunc/LibLinearDataset.java:413-484,586-658; jobs/RegressionAdmmTrain.java:677-690):
 * records are uploaded ONCE and stay resident in HBM.  Rows are RegressionPrepareOutput records
 * (src/main/avro/RegressionPrepareOutput.avsc): response in {1,0,-1} (0 -> -1), weight >= 0, float32
 * values.  Pointers are host-or-device; the library copies.
 * dense: X row-major [nrows x num_features], leading dimension ldx (floats).
 * csr:   rowptr [nrows+1] (int64), colidx (global ids, any order, duplicates add), vals.  The call returns when the arrays
 *        have been copied (the caller's buffers are free again); the checks on colidx and the derived lists of partition p
 *        are built while partition p+1 is being copied, so "partition <id>: feature index out of range" is reported by the
 *        NEXT call on the session (add_partition / begin / fit / objective).  response / weight errors are immediate. */
int mlease_add_partition_dense(mlease_session* s, int32_t partition_id, int64_t nrows, const float* X, int64_t ldx,
                               const int32_t* response, const float* weight, const float* offset);
int mlease_add_partition_csr(mlease_session* s, int32_t partition_id, int64_t nrows, const int64_t* rowptr,
                             const int32_t* colidx, const float* vals, const int32_t* response, const float* weight,
                             const float* offset);

/* ADMM loop, one iteration = jobs/RegressionAdmmTrain.java:281-497 :
 *   begin      : z = {}, u = {}  (:155-185, :312), schedule reset (:278-279)
 *   local_step : x-update of every local (partition, lambda) (AdmmReducer.reduce :642-718) and
 *                exchange[l][k] = sum over LOCAL partitions of float(x_p)[k] + u_p[k]   (device, double,
 *                [L][num_features+1]) -- the only data that crosses GPUs
 *   consensus  : given the exchange buffer summed over ALL processes (one all-reduce), the z-update
 *                (:362-404), convergence (:456-472), schedule + stop rule (:338-346,:493-496) and next u
 *                (computeU :736-765).  *stop is 1 when the reference would break.
 * mlease_admm_run drives these for a single-process job (allreduce == NULL) or with a caller-supplied
 * all-reduce (sum, double, in place on `buf` which is DEVICE memory, ordered on `stream`). */
typedef int (*mlease_allreduce_fn)(void* ctx, double* buf, size_t count, void* stream);
int mlease_admm_begin(mlease_session* s);
/* initialize.boost.rate > 0 (jobs/RegressionAdmmTrain.java:236-266, 313-316): the run starts from z0 ([num_lambdas][num_features+1]
 * doubles on the host, intercept last: the mean of per-partition RegressionNaiveTrain fits, which the caller obtains with
 * mlease_fit_partition) instead of z = {}, u is empty, and the reducers of ITERATION 1 use rho * boost_rate: the driver builds a
 * new JobConf every iteration (:286-291), so rho.adapt.rate is back at the reducers' default 1.0f (:621) from iteration 2 on
 * (or follows rho_adapt_coefficient, :323-327).  The factors built with the boosted rho are invalidated at iteration 2. */
int mlease_admm_begin_initialized(mlease_session* s, const double* z0, float boost_rate);
int mlease_admm_local_step(mlease_session* s, double* exchange_dev);
int mlease_admm_consensus(mlease_session* s, const double* exchange_sum_dev, double* maxdiff, int32_t* stop);
int mlease_admm_run(mlease_session* s, int32_t num_iters, mlease_allreduce_fn allreduce, void* ctx, int32_t* iters_done);
/* One iteration with the exchange inside the library: local_step, all-reduce over the attached communicator (none for a
 * single-process job holding all num_blocks partitions), consensus.  Lets a host job write the reference's iter-<i>/ files
 * between iterations.  A failed fit on any rank returns MLEASE_ERR_NUMERIC on every rank ("Model fitting error!", :713-716). */
int mlease_admm_iterate(mlease_session* s, double* maxdiff, int32_t* stop);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: the path shards exactly where ADMM does -- partitions are independent in the x-update and meet in ONE
 * all-reduce (sum, fp64, [L][num_features+1] (+1 failure counter)) per iteration, the mean the reference's driver takes over
 * the reducer outputs (:362-364, cons/MeanLinearModelConsumer.java:44-70).  NCCL is loaded at run time (libnccl.so.2).
 *  (a) one process per GPU: rank 0 calls mlease_comm_unique_id and ships the 128 bytes to the other ranks by any means;
 *      every rank creates its communicator, attaches it to its session (which holds the partitions p with p % nranks == rank)
 *      and calls mlease_admm_run(s, iters, NULL, NULL, &done): the whole RegressionAdmmTrain loop runs in C.
 *  (b) one process, several GPUs (a single-JVM driver, the C++ job layer): mlease_world_* below -- same calls as a session,
 *      partitions routed to GPU  partition_id % ndev, one worker thread per GPU, ncclCommInitAll inside.
 * ------------------------------------------------------------------------------------- */
#define MLEASE_COMM_ID_BYTES 128
int mlease_comm_unique_id(void* id128);
int mlease_comm_create(const void* id128, int32_t rank, int32_t nranks, int32_t device, mlease_comm** out);
int mlease_comm_destroy(mlease_comm* c);
int mlease_comm_info(const mlease_comm* c, int32_t* rank, int32_t* nranks, int32_t* nccl_version);
int mlease_session_set_comm(mlease_session* s, mlease_comm* comm);   /* not owned; NULL detaches */

int mlease_world_create(const mlease_admm_config* cfg /* .device/.stream ignored */, const int32_t* devices /* NULL = 0..ndev-1 */,
                        int32_t ndev, mlease_world** out);
int mlease_world_destroy(mlease_world* w);
int mlease_world_num_devices(const mlease_world* w);
int mlease_world_add_partition_dense(mlease_world* w, int32_t partition_id, int64_t nrows, const float* X, int64_t ldx,
                                     const int32_t* response, const float* weight, const float* offset);
int mlease_world_add_partition_csr(mlease_world* w, int32_t partition_id, int64_t nrows, const int64_t* rowptr, const int32_t* colidx,
                                   const float* vals, const int32_t* response, const float* weight, const float* offset);
int mlease_world_begin(mlease_world* w);
int mlease_world_begin_initialized(mlease_world* w, const double* z0, float boost_rate);
int mlease_world_iterate(mlease_world* w, double* maxdiff, int32_t* stop);
int mlease_world_run(mlease_world* w, int32_t num_iters, int32_t* iters_done);
int mlease_world_get_z(mlease_world* w, int32_t lambda_idx, double* out);
int mlease_world_get_final_model(mlease_world* w, int32_t lambda_idx, float* out);
int mlease_world_get_x(mlease_world* w, int32_t partition_id, int32_t lambda_idx, double* out);
int mlease_world_get_u(mlease_world* w, int32_t partition_id, int32_t lambda_idx, float* out);
int mlease_world_get_uplusx(mlease_world* w, int32_t partition_id, int32_t lambda_idx, float* out);
int mlease_world_fit_partition(mlease_world* w, int32_t partition_id, double* x, const double* m, const double* q, int32_t* newton_steps);

/* State readback (host buffers).  z: driver-side double z (:365-404); final model = float(z)
 * (models/LinearModel.java:697-720 toAvro).  After consensus of iteration i: x = the double x_p of iteration i
 * (float(x) is iter-<i>/model), uplusx = float(u+x) of iteration i (:706-711), u = float(uplusx - z), i.e. the
 * iter-<i+1>/u file computeU (:736-765) writes.  Length num_features+1, intercept last. */
int mlease_get_z(mlease_session* s, int32_t lambda_idx, double* out);
int mlease_get_final_model(mlease_session* s, int32_t lambda_idx, float* out);
int mlease_get_x(mlease_session* s, int32_t partition_id, int32_t lambda_idx, double* out);
int mlease_get_u(mlease_session* s, int32_t partition_id, int32_t lambda_idx, float* out);
int mlease_get_uplusx(mlease_session* s, int32_t partition_id, int32_t lambda_idx, float* out);

typedef struct {
  int64_t k1_passes;        /* fused score/reweight/gradient passes over X (all problems) */
  int64_t gram_builds;      /* Gram + Cholesky refreshes */
  int64_t newton_steps;     /* accepted Newton steps */
  int64_t rejected_steps;   /* line-search rejections */
  int64_t kernel_launches;  /* kernels launched by this session */
  int32_t not_converged;    /* x-updates that hit max_newton */
  int32_t last_iter_slots;  /* evaluation slots used by the last local_step */
  double last_maxdiff;
  float liblinear_epsilon;  /* schedule variable (:279,338-346), control only */
  int32_t k1_fused;         /* 1: the fused multi-lambda CSR K1 (csrc/k1_csr_fused.cu) serves this session's ADMM problems */
  double k1_shared_bytes;   /* CSR sessions: K1 bytes when the lambdas of a partition count as ONE read of its rows
                               (8*nnz + 9*n per partition pass + 8*n per lambda served); k1_bytes of mlease_profile counts every
                               (partition, lambda) pass separately, as SURVEY 8d defines the unit */
} mlease_stats;
int mlease_get_stats(mlease_session* s, mlease_stats* out);
int mlease_world_get_stats(mlease_world* w, mlease_stats* out);   /* counters summed over the devices */
/* Per-kernel device timing for roofline reporting (CUDA events on the session stream around every launch of
 * the Newton slot; categories: 0 = K1 fused pass, 1 = small kernels (reduce/decide, solve, poll), 2 = Gram (tcgen05),
 * 3 = Cholesky).  enable: 1 on, 0 off, 2 on + reset accumulators, -1 read only.  Outputs (any may be NULL) are the
 * accumulators BEFORE this call's reset: ms4[4], count4[4], and the algorithmic work done by the session so far:
 * k1_bytes (dense n*(4*ldx+9) per pass), k1_emit_bytes (bf16 operand writes), gram_flops (n*D'*(D'+1)). */
int mlease_profile(mlease_session* s, int32_t enable, double* ms4, int64_t* count4, double* k1_bytes, double* k1_emit_bytes,
                   double* gram_flops);

/* ---------------------------------------------------------------------------------------
 * Function-level entry points (parity tests against the oracle's fun/grad/hessian):
 * LogisticRegressionL2.fun/grad/hessian (regression/liblinearfunc/LogisticRegressionL2.java:156-297)
 * evaluated on a resident partition at host vector w, prior mean m, prior precision q (=1/priorVar),
 * all of length num_features+1.  Any output may be NULL.  H is [Dt x Dt] row-major (full, symmetric).
 * tensor != 0 builds H with the tcgen05 Gram kernel (bf16 operands for dense partitions, e4m3 operands assembled from the rows
 * for CSR partitions with sorted unique rows), 0 with the fp32 SIMT debug kernel (dense bf16 operand only).
 * tensor == 2 returns in H the INVERSE the Newton direction is computed with (tcgen05 Gram + diag(q) -> fp64 blocked
 * Cholesky -> explicit inverse), so that tests can check H^-1 * H = I for every factorisation path.
 * ------------------------------------------------------------------------------------- */
int mlease_objective(mlease_session* s, int32_t partition_id, const double* w, const double* m, const double* q,
                     double* f, double* g, double* H, int32_t tensor);
/* LibLinear.train(dataset, init, priorMean, priorVar...) (regression/liblinearfunc/LibLinear.java:200-208) for one
 * resident partition: exact Newton solve of the same objective.  x: in = init, out = minimiser. */
int mlease_fit_partition(mlease_session* s, int32_t partition_id, double* x, const double* m, const double* q,
                         int32_t* newton_steps);

/* Posterior variance of the model w of one resident partition under prior precision q (= 1/priorVar), the
 * computePosteriorVar / computeFullPostVar tail of LibLinear.train (regression/liblinearfunc/LibLinear.java:315-334) that
 * ItemModelTrain reports as "posteriorVar" (jobs/ItemModelTrain.java:257-266):
 *   full = 0: var[k] = 1 / (q[k] + sum_i weight_i p_i (1-p_i) x_ik^2)         (hessianDiagonal, LogisticRegressionL2.java:304-327)
 *   full = 1: var = diag(H^-1), H = LogisticRegressionL2.hessian (:258-297) accumulated in fp64 (NOT the bf16 tensor-core
 *             Gram, which only preconditions), Cholesky + explicit inverse in fp64; cov (may be NULL) receives H^-1,
 *             [Dt x Dt] row-major.  Needs rows with strictly increasing column ids, as the reference's hessian() does (:277).
 * var has num_features+1 entries, intercept last; a feature absent from the partition gets its prior variance 1/q[k]
 * (the reference lists only the features present in the dataset). */
int mlease_posterior_variance(mlease_session* s, int32_t partition_id, const double* w, const double* q, int32_t full,
                              double* var, double* cov);

/* ---------------------------------------------------------------------------------------
 * RegressionNaiveTrain (jobs/RegressionNaiveTrain.java:302-415): num_keys x num_lambdas independent fits ("lambda#key"
 * reducers, :228-241).  Key k owns rows [key_rowstart[k], key_rowstart[k+1]) of ONE matrix, uploaded once for all lambdas:
 *   CSR   (rowptr != NULL): rowptr [nrows+1] int64, colidx (global ids), vals -- the reference's per-key sparse datasets
 *         (:360-398).  A feature that no row of a key lists is not in that key's dataset, hence not in its model
 *         (regression/liblinearfunc/LibLinear.java:343-350): its output coefficient is 0, whatever prior.mean is.
 *         binary_feature: every listed feature counts as 1 (LibLinearBinaryDataset).
 *   dense (rowptr == NULL): vals = X row-major [nrows x num_features], leading dimension ldx; every feature is present.
 * priorVar = 1/lambda, 1/lambda_map[k] for listed features (lambda_map [num_features] or NULL, entries > 0), intercept variance
 * 100000 unless penalize_intercept (:333-343), prior.mean, has.intercept, data.size.threshold (skipped keys -> skipped[k]=1,
 * model 0, :379-382).  out_model [num_lambdas][num_keys][num_features+1] double, intercept last.  All pointers host-or-device
 * except out_model / skipped (host).  A fit that does not converge -> MLEASE_ERR_NUMERIC ("Model fitting error!", :400-412).
 * ------------------------------------------------------------------------------------- */
int mlease_naive_train(int32_t device, void* stream, int32_t num_keys, int32_t num_features, const int64_t* key_rowstart,
                       const int64_t* rowptr, const int32_t* colidx, const float* vals, int64_t ldx, const int32_t* response,
                       const float* weight, const float* offset, int32_t num_lambdas, const float* lambdas, const float* lambda_map,
                       float prior_mean, int32_t penalize_intercept, int32_t has_intercept, int32_t data_size_threshold,
                       int32_t binary_feature, double* out_model, int32_t* skipped);
/* single-lambda dense form of the above (kept from ABI version 1) */
int mlease_naive_train_dense(int32_t device, void* stream, int32_t num_keys, int32_t num_features,
                             const int64_t* key_rowstart, const float* X, int64_t ldx, const int32_t* response,
                             const float* weight, const float* offset, float lambda, const float* lambda_map,
                             float prior_mean, int32_t penalize_intercept, int32_t has_intercept,
                             int32_t data_size_threshold, double* out_model, int32_t* skipped);

/* ---------------------------------------------------------------------------------------
 * RegressionTest / RegressionTestLoglik.
 * score: pred = float(offset + interceptTerm + sum beta_k x_k), interceptTerm = -log(n-1+n*exp(-b)),
 *        n = num.click.replicates (models/LinearModel.java:241-257,491-554; jobs/RegressionTest.java:163).
 *        dense (colidx==NULL: X = vals, ld = ldx) or CSR.  All data pointers host-or-device; pred host-or-device.
 * test_loglik: mapper float cast, combiner partial sums cast to float per `combiner_block` records
 *        (<=0: no combiner), reducer float(sum/count) (jobs/RegressionTestLoglik.java:124-200).
 * ------------------------------------------------------------------------------------- */
int mlease_score(int32_t device, void* stream, int32_t num_features, int64_t nrows, const int64_t* rowptr,
                 const int32_t* colidx, const float* vals, int64_t ldx, const float* offset, const double* model,
                 int32_t num_click_replicates, int32_t binary_feature, float* pred);
int mlease_test_loglik(int32_t device, void* stream, int64_t nrows, const int32_t* response, const float* pred,
                       const float* weight, int64_t combiner_block, float* out_loglik, double* out_count);

/* Bench / profiling hooks (not part of the reference surface): time one fused K1 pass or one Gram build
 * on a resident partition with CUDA events on the session stream, `reps` launches, returns avg ms. */
int mlease_time_kernel(mlease_session* s, int32_t partition_id, int32_t which /*1=K1,2=Gram tcgen05,3=cholesky*/,
                       int32_t reps, int32_t emit_scaled, float* avg_ms);

#ifdef __cplusplus
}
#endif
#endif /* MLEASE_B200_H */
