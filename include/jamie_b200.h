/* jamie_b200 -- C ABI of the B200-native JAMIE hot path (coupled-VAE train step + modal_predict / transform).
 *
 * The reference (Oafish1/JAMIE v4.4.5) is pure Python and has no FFI; these entry points are what a binding for its
 * hot path would call.  Each one names the reference code it replaces (paths relative to the reference root).
 * Binding stub (ctypes) and the Python-side mirror of the reference API: INTEGRATION.md, jamie_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; jb_last_error() gives the message (thread-local).
 *     The Python wrapper raises RuntimeError / AssertionError with the reference's own messages where it has one.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). All device work of a call is
 *     enqueued on it; calls do not synchronise the host unless stated.
 *   - the handle owns parameters, Adam moments, BatchNorm statistics, datasets and workspaces; the
 *     caller owns every buffer it passes in. A handle is bound to one device and is not thread-safe.
 *   - "packed" parameter order = edModelVar.named_parameters() order (jamie/model.py:147-220):
 *       sigma[2]; encoders.{0,1}.{0.weight,0.bias,1.weight,1.bias,4.weight,4.bias,5.weight,5.bias};
 *       fc_mus.{0,1}.{weight,bias}; fc_vars.{0,1}.{weight,bias};
 *       decoders.{0,1}.{0.weight,0.bias,1.weight,1.bias,4.weight,4.bias,5.weight,5.bias,8.weight,8.bias}
 *     "packed" BatchNorm order = encoders.0.1, encoders.0.5, encoders.1.1, encoders.1.5, decoders.0.1, decoders.0.5,
 *     decoders.1.1, decoders.1.5; for each layer running_mean[w] then running_var[w].
 */
#ifndef JAMIE_B200_H
#define JAMIE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jb_engine jb_engine;

typedef struct jb_config {
  int dims[2];          /* post-preprocessing widths D0, D1 (self.col, jamie/jamie.py:469) */
  int latent;           /* output_dim (jamie/jamie.py:475) */
  int max_batch;        /* largest batch size that will be stepped (batch_size, jamie/jamie.py:511-514) */
  float dropout;        /* resolved dropout probability (jamie/model.py:144-145) */
  float lr;             /* model_lr (jamie/jamie.py:481) */
  float beta1, beta2, adam_eps;   /* torch.optim.Adam defaults .9, .999, 1e-8 */
  float max_grad_norm;  /* clip_grad_norm_(..., 1) (jamie/jamie.py:739) */
  float loss_w[4];      /* loss_weights or 1s: KL, Rec, CosSim, F (jamie/jamie.py:723-728) */
  float pf_ratio;       /* PF_Ratio (None -> 1) (jamie/jamie.py:517, 604) */
  unsigned long long seed;   /* Philox key for dropout masks / eps when none are injected */
  int device;           /* CUDA device ordinal */
  int world_size;       /* data-parallel ranks; gradients are divided by it in jb_step_update */
} jb_config;

const char* jb_last_error(void);
int jb_version(void);

/* edModelVar.__init__ + optim.Adam(...) (jamie/model.py:121-220, jamie/jamie.py:472-481): allocates everything. */
int jb_create(const jb_config* cfg, jb_engine** out);
void jb_destroy(jb_engine* e);

/* number of packed parameters / packed BatchNorm floats */
long long jb_num_params(const jb_engine* e);
long long jb_num_bn_floats(const jb_engine* e);

/* state_dict I/O for save_model / load_model (jamie/jamie.py:967-972). Host pointers, packed order. Synchronous. */
int jb_set_params(jb_engine* e, const float* packed, long long n);
int jb_get_params(jb_engine* e, float* packed, long long n);
int jb_set_bn_stats(jb_engine* e, const float* packed, long long n, const long long num_batches_tracked[8]);
int jb_get_bn_stats(jb_engine* e, float* packed, long long n, long long num_batches_tracked[8]);
/* gradients of the last backward pass, packed order (parity tests; the reference reads p.grad) */
int jb_get_grads(jb_engine* e, float* packed, long long n);
/* Adam moments (packed order) and step count -- optional resume side-car; the reference never saves them. */
int jb_get_adam_state(jb_engine* e, float* m_packed, float* v_packed, long long n, long long* t);
int jb_set_adam_state(jb_engine* e, const float* m_packed, const float* v_packed, long long n, long long t);

/* self.dataset[i] after preprocessing (jamie/jamie.py:468, 507-508): n rows of dims[mod] floats, row pitch ld.
 * on_device != 0: X is a device pointer. The engine keeps its own copy (fp32). */
int jb_set_dataset(jb_engine* e, int mod, const float* X, long long n, long long ld, int on_device, void* stream);

/* Correspondence prior P (jamie/jamie.py:423-431) and F = match_result[0] (jamie/jamie.py:431).
 *   diag : P = diag(m), m[n] (identity, partially matched, or zeros) -- no n x n object is ever built
 *   dense: row-major [n0, n1] host matrix (small n only)
 *   F    : dense or absent (use_f_tilde=False -> zeros, jamie/jamie.py:172-173) */
int jb_set_prior_diag(jb_engine* e, const float* m, long long n);
int jb_set_prior_dense(jb_engine* e, const float* P, long long n0, long long n1);
int jb_set_f_dense(jb_engine* e, const float* F, long long n0, long long n1);   /* F == NULL: zeros */

/* Upload the sampling plan of the next `nsteps` optimizer steps (one epoch or more): batch indices per modality
 * (random_batch, jamie/jamie.py:553-583) and the KL anneal factor per step (jamie/jamie.py:630-631; the 0.032 and
 * loss weight are applied by the engine). Resets the plan cursor. idx*: [nsteps, batch] int64 host arrays. */
int jb_upload_plan(jb_engine* e, const long long* idx0, const long long* idx1, const double* kl_anneal,
                   int nsteps, int batch, void* stream);

/* Inject the randomness of the NEXT step only (parity tests): eps[i] = [batch, latent] fp32 (Normal.rsample,
 * jamie/model.py:239-240), masks[k] = [batch, width_k] uint8 keep-masks in draw order enc0 d1,d2; enc1 d1,d2;
 * dec0 d1,d2; dec1 d1,d2 (jamie/model.py:154,164,195,200). Host pointers. Without it the engine draws Philox. */
int jb_inject_randomness(jb_engine* e, const float* eps0, const float* eps1, const unsigned char* const masks[8],
                         void* stream);

/* Run `nsteps` full optimizer steps from the plan cursor: gather, P/F blocks, forward, 4 losses, backward,
 * clip_grad_norm_, Adam, zero_grad (jamie/jamie.py:549-742 with batch_step=True). Asynchronous; ONE launch of the
 * persistent step kernel for all `nsteps` steps (csrc/stepk.cuh), no host synchronisation. */
int jb_train_steps(jb_engine* e, int nsteps, void* stream);
/* Data-parallel split of one step: backward leaves the summed-loss gradients in the flat buffer returned by
 * jb_grad_buffer (the caller all-reduces it, e.g. NCCL sum), update divides by world_size, clips and applies Adam. */
int jb_step_backward(jb_engine* e, void* stream);
int jb_step_update(jb_engine* e, void* stream);
int jb_grad_buffer(jb_engine* e, float** dev_ptr, long long* n_floats);
/* Make the engine write its gradients into a caller-owned device buffer of at least the jb_grad_buffer size (16-byte
 * aligned), e.g. one allocated in NVLink symmetric / multicast memory so that the exchange can be a multimem all-reduce
 * instead of an NCCL call. dev_ptr = NULL returns to the engine's own buffer. The caller keeps the buffer alive. */
int jb_set_grad_buffer(jb_engine* e, float* dev_ptr, long long n_floats);
/* In-kernel gradient exchange over peer memory (NVLink / NVSwitch): grad_ptrs[q] = rank q's gradient buffer (each rank
 * passed its own to jb_set_grad_buffer; all of them mapped into this process, e.g. NVLink symmetric memory), flag_ptrs[q]
 * = rank q's zero-initialised scratch block of at least jb_exchange_scratch_bytes(). With an exchange configured the
 * step kernel sums the gradients across the ranks itself (reduce-scatter + all-gather + clip-norm partials inside the
 * persistent kernel, between WGRAD and ADAM): jb_train_steps(n) is ONE launch for n data-parallel optimizer steps and the
 * caller must not all-reduce jb_grad_buffer. All ranks must run the same number of steps in lockstep. world <= 1 or
 * NULL pointers turn it off. grad_multicast = the NVSwitch multicast address bound to all ranks' gradient buffers, or
 * NULL: with it the sum is taken inside the switch (multimem.ld_reduce) and broadcast with multimem.st. */
int jb_set_exchange(jb_engine* e, int rank, int world, float* const* grad_ptrs, unsigned int* const* flag_ptrs,
                    float* grad_multicast);
long long jb_exchange_scratch_bytes(void);
/* dist_method of sim_diff_func (jamie/jamie.py:484-502): 0 = 'euclidean' (the default: the "CosSim" loss is the row-wise
 * squared distance |z_i - c_i|^2), 1 = 'cosine' ((1 - cos(z_i, c_i))^2). Takes effect with the next step. */
int jb_set_dist_method(jb_engine* e, int method);
/* batch_step=False (jamie/jamie.py:744-749): accumulate gradients over several jb_step_backward calls, one
 * jb_step_update per epoch. accumulate != 0 makes the following backward passes add into the gradient buffer instead of
 * overwriting it (a device-side flag: no rebuild, no synchronisation). The optimizer step count (Adam bias correction)
 * advances per jb_step_update, the Philox stream per backward pass. */
int jb_set_grad_accumulate(jb_engine* e, int accumulate);

/* One optimizer step whose batch arrives from HOST memory (the reference keeps self.dataset wherever `device` says;
 * this is the end-to-end form for host-resident data): x0/x1 = the gathered rows data[i][idx_i] ([batch, dims[i]],
 * packed, ideally pinned), idx0/idx1 = their global cell ids (needed for the P/F blocks), kl_anneal as in
 * jb_upload_plan. Copies the batch in, runs the step graph, copies the 8 loss scalars out and synchronises. */
int jb_train_step_hostbatch(jb_engine* e, const float* x0, const float* x1, const long long* idx0, const long long* idx1,
                            int batch, double kl_anneal, float out_losses[8], void* stream);

/* Data-parallel form of the host-batch step: copy the batch in and run forward + backward only; the caller then
 * all-reduces jb_grad_buffer and calls jb_step_update. Host-batch steps alternate between two plan rows (slots): the
 * k-th host-batch step (k = 0, 1, ...) leaves its losses in row k & 1 of jb_read_losses(e, out, 2, stream). */
int jb_step_backward_hostbatch(jb_engine* e, const float* x0, const float* x1, const long long* idx0,
                               const long long* idx1, int batch, double kl_anneal, void* stream);

/* Asynchronous form of jb_train_step_hostbatch for a host loop that prepares batch k + 1 (the reference's
 * dataset[i][random_batch[i]] gather, jamie/jamie.py:583) while step k runs: submit enqueues the copies on an internal
 * copy stream and the step on `stream` and returns at once; wait blocks until the OLDEST submitted step has finished and
 * returns its 8 loss scalars. At most two steps may be in flight; x0 / x1 must stay valid (pinned) until the step's
 * jb_hostbatch_wait has returned. */
int jb_hostbatch_submit(jb_engine* e, const float* x0, const float* x1, const long long* idx0, const long long* idx1, int batch,
                        double kl_anneal, void* stream);
int jb_hostbatch_wait(jb_engine* e, float out_losses[8]);

/* Phases of the step kernel (csrc/stepk.cuh: StepPhase), in execution order. */
int jb_num_phases(void);
const char* jb_phase_name(int phase);

/* Benchmark hook: `iters` launches of the step kernel restricted to ONE phase, timed with CUDA events on `stream`.
 * Returns the average microseconds per launch (kernel set-up included: an upper bound of the phase's share of a step)
 * and, for GEMM phases, the FLOPs of one launch (0 otherwise). */
int jb_bench_stage(jb_engine* e, int phase, int iters, float* avg_us, double* flops, void* stream);

/* Profiling hook: runs `iters` (+1 warm-up) training steps in ONE launch of the persistent step kernel while CTA 0 records
 * the GPU global timer at every phase boundary; out_us[p] = average microseconds of phase p inside the kernel (its grid
 * barrier included), *n_launches = jb_num_phases(). Rewinds the plan cursor first (needs >= 2 plan rows) and takes
 * optimizer steps like jb_train_steps. */
int jb_profile_step(jb_engine* e, int iters, float* out_us, int cap, int* n_launches, void* stream);
/* Detail of the last jb_profile_step: out[4 p + k] for phase p, k = 0 total us (as above), 1 longest CTA work,
 * 2 mean CTA work, 3 barrier tail (end of the grid barrier - end of the last CTA's work). */
int jb_profile_detail(jb_engine* e, float* out, int cap);

/* Per-step results of the steps run since the last jb_upload_plan, in plan order. Synchronises `stream`.
 * out[s*8 + k]: k=0..3 the reference's `losses` list (KL incl. 0.032*anneal, Rec, 32*CosSim, F; unweighted by
 * loss_weights), k=4 weighted total (batch_loss), k=5 pre-clip gradient norm, k=6,7 reserved. */
int jb_read_losses(jb_engine* e, float* out, int nsteps, void* stream);

/* Eval-mode paths (BatchNorm running statistics folded into the weights, dropout off, z = mu).
 *   jb_encode : fc_mus[mod](encoders[mod](x))                      transform_one / transform / final encode
 *               (jamie/jamie.py:794-799, 817-837)
 *   jb_predict: decoders[to](fc_mus[from](encoders[from](x)))      edModelVar.impute (jamie/model.py:277-282)
 * X: [n, dims[from]] fp32 with row pitch ldx; out: [n, latent] or [n, dims[to]] with row pitch ldo.
 * on_device != 0: both are device pointers (no copies); else host pointers (pinned for full overlap) that are
 * streamed through the device in row chunks. Synchronous for host pointers, asynchronous for device pointers. */
int jb_encode(jb_engine* e, int mod, const float* X, long long n, long long ldx, float* out, long long ldo,
              int on_device, void* stream);
int jb_predict(jb_engine* e, int from, int to, const float* X, long long n, long long ldx, float* out,
               long long ldo, int on_device, void* stream);

/* preclass.transform / inverse_transform with PCA (jamie/utilities.py:660-678): out = ((X - mean) @ comp^T - m)/s
 * and out = (Z*s + m) @ comp + mean, fp32 on device (3xTF32 split for fp32-level accuracy).
 * comp: [k, d] row-major host matrix (pca.components_), mean: [d] (pca.mean_), m/s: scalar standardisation.
 * X/out host or device pointers as above. */
int jb_pca_project(jb_engine* e, const float* X, long long n, long long d, const float* comp, const float* mean,
                   int k, float m, float s, float* out, int on_device, void* stream);
int jb_pca_inverse(jb_engine* e, const float* Z, long long n, int k, const float* comp, const float* mean,
                   long long d, float m, float s, float* out, int on_device, void* stream);

/* PCA fit (jamie/jamie.py:436-452, sklearn PCA(n_components).fit with the covariance_eigh solver): the two O(n d) /
 * O(n d^2) passes over the [n, d] matrix run on the GPU -- column sums, then the centred Gram matrix
 * (X - mean)^T (X - mean) as a split (fp32-class) tensor-core GEMM accumulated over row chunks in float64. The caller
 * divides / all-reduces (rows sharded over ranks: sum colsum and n, then sum the Gram matrices) and solves the d x d
 * symmetric eigenproblem (host LAPACK: independent of n). X host or device pointer (packed rows); colsum [d], gram [d, d]
 * host float64. Synchronous. */
int jb_pca_colsum(jb_engine* e, const float* X, long long n, long long d, double* colsum, int on_device, void* stream);
int jb_pca_gram(jb_engine* e, const float* X, long long n, long long d, const double* mean, double* gram, int on_device,
                void* stream);

/* Evaluation metrics of the reference on the GPU (no engine handle needed; host pointers in, synchronous).
 *   jb_metric_foscttm: test_closer (jamie/evaluation.py:65-85, jamie/jamie.py:892-913) on two [n, L] embeddings with the
 *     euclidean distance: raw_count_closer = #{j: d(a_i, b_j) < d(a_i, b_i)} + #{j: d(b_i, a_j) < d(b_i, a_i)} summed over
 *     i; the caller divides by 2 n^2. Distances in float64.
 *   jb_metric_knn_vote: the kNN classifier of test_LabelTA (jamie/evaluation.py:114-132, jamie/jamie.py:943-961): for
 *     every query row the majority class (uniform votes; ties to the lowest class index) of its k nearest reference rows
 *     (euclidean; equal distances by the lowest row index). ref_class: class indices in [0, n_classes).
 *   jb_metric_feature_pearson: per-feature Pearson r of two [n, d] matrices (jamie/evaluation.py:491-513, sklearn
 *     r_regression); a constant feature gives nan. */
int jb_metric_foscttm(const float* emb0, const float* emb1, long long n, int L, int device,
                      unsigned long long* raw_count_closer);
int jb_metric_knn_vote(const float* query, long long nq, const float* ref, const int* ref_class, long long nr, int L,
                       int k, int n_classes, int device, int* pred_class);
int jb_metric_feature_pearson(const float* x, const float* y, long long n, long long d, int device, double* r);

/* Debug / parity taps: copy a named intermediate of the last step to the host (synchronises).
 * Names: "x0","y1_0","h1_0","y2_0","h2_0","mulv0","z0","c0","xhat0", ... (see csrc/engine.cu: tap table),
 * "corr", "fblk", "grad" (padded flat), "theta". Returns the number of floats written, or -1. */
long long jb_debug_read(jb_engine* e, const char* name, float* out, long long cap);
/* number of kernel launches issued by the handle so far */
long long jb_launch_count(const jb_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* JAMIE_B200_H */
