/* gpnet_b200 -- C-ABI of the B200-native dense-GP hot path of filippocastelli/GPnet.
 *
 * The reference is pure Python and has no FFI: the boundary it offers is the Python class API
 * (GPnetRegressor / GPnetClassifier).  This header is the C-ABI introduced BENEATH that API; each entry point
 * names the reference code it replaces (file:line in the reference repository).  The Python drop-in classes in
 * gpnet_b200/ bind it with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - one context per process per GPU; every call is stream-ordered on the context's stream and returns after
 *    the work is enqueued, except calls with `_h` output pointers, which synchronise before returning;
 *  - pointers are DEVICE pointers unless the parameter name ends in `_h` (host);
 *  - matrices are row-major fp64 with an explicit leading dimension (elements); leading dimensions and any
 *    sub-matrix column offset must be even and base pointers 16-byte aligned (TMA stride rule);
 *  - return value: 0 ok; >0 LAPACK-style info (order of the leading minor that is not positive definite);
 *    <0 bad argument (-i = i-th argument); <= -1000 CUDA/driver error (gpn_last_error() has the text);
 *  - never throws, never calls back into the host language;
 *  - hyper-parameters cross the boundary EXPONENTIATED: every `etheta_h` (and the landscape axes) holds np.exp(theta), the
 *    vector the reference itself computes first (`theta = np.exp(theta)`, GPnet.py:303).  The host evaluates that expression
 *    with numpy so that the p-step order int(np.exp(theta)[1]) (GPnet.py:347, quirk A-4) and the domain asserts
 *    (GPnet.py:339,344) see exactly the reference's values: libm's exp differs from numpy's by an ulp on some inputs;
 *  - every entry point that enqueues device work opens an NVTX range named after itself (visible to nsys / ncu --nvtx).
 */
#ifndef GPNET_B200_H
#define GPNET_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpn_ctx gpn_ctx;

/* kernel types, GPnet.py:333 kernel_list */
enum { GPN_KERNEL_DIFFUSION = 0, GPN_KERNEL_REGULARIZED_LAPLACIAN = 1, GPN_KERNEL_PSTEP_WALK = 2 };

/* ---- context ------------------------------------------------------------------------------------------- */
gpn_ctx* gpn_create(int device, void* cuda_stream /* NULL: the library creates its own stream */);
void gpn_destroy(gpn_ctx* ctx);
const char* gpn_last_error(void);
int gpn_sync(gpn_ctx* ctx);
/* counters: kernels launched through this context and tensor flops executed by the GEMM launcher */
long long gpn_launch_count(gpn_ctx* ctx);
double gpn_flop_count(gpn_ctx* ctx);
void gpn_reset_counters(gpn_ctx* ctx);

/* measurement hooks (bench.py): sustained FP64 tensor (DMMA.8x8x4) issue-rate ceiling of this GPU, measured for
 * about `seconds`; per-launch CUDA-event profiling of the flop-carrying kernels (out_h[0..2]: ms, executed flops, launches
 * of all GEMM launches since gpn_set_profile(ctx,1); out_h[3..5]: the same for the 128x128 GEMM configuration only;
 * out_h[6..8]: the same for the dataflow Cholesky kernel; out_h[9] / out_h[10]: length in ms of the UNION of the launch
 * intervals of the dataflow Cholesky / of the GEMM launches -- launches of one kernel that run concurrently on several
 * streams share the chip, so chip-level throughput is executed flops / union; out_h has 11 entries). */
int gpn_fp64_peak(gpn_ctx* ctx, double seconds, double* tflops_h);
void gpn_set_profile(gpn_ctx* ctx, int on);
int gpn_profile_summary(gpn_ctx* ctx, double* out_h);
/* the recorded launch intervals themselves, for merging across contexts that evaluate concurrently (lanes): row i of out_h =
 * (start ms, end ms, executed flops, kind: 0 GEMM, 2 dataflow Cholesky) relative to the first recorded launch of `ref`
 * (NULL: ctx); *rows_h = rows written, at most max_rows */
int gpn_profile_intervals(gpn_ctx* ctx, gpn_ctx* ref, double* out_h, int max_rows, int* rows_h);
/* development aid: device buffer of 32 int64 receiving clock64() stamps of the phases of the next leaf-kernel launch */
void gpn_debug_leaf_prof(gpn_ctx* ctx, void* dev_buf32);
/* development aid: run the 128x128 leaf on `copies` stacked blocks (lda 128), one CTA each; block 0 stamps its phases */
int gpn_debug_leaf_bench(gpn_ctx* ctx, void* A_dev, int copies, void* prof32_dev);
/* development aid: the 32 device reductions the last evaluation read back (layout private to the route that ran) */
int gpn_debug_scalars(gpn_ctx* ctx, double* out32_h);
/* development aid: layout of the whole-graph dissection, out40 = [K, s, so, ND, off[9], nk[8], ok[8], c0[8]] */
int gpn_debug_d6_layout(gpn_ctx* ctx, int* out40);
/* development aid: device pointer / capacity in bytes of work buffer `which` (0..11) of the whole-graph dissection route */
void* gpn_debug_d6_buffer(gpn_ctx* ctx, int which, long long* bytes);
/* development aid: device buffer receiving 8 int64 %globaltimer stamps per tile of the next dataflow-Cholesky launches
 * (tile (i, j) at slot j * tile_rows + i; tools/gpu_check.py chain / dfbreak / augbreak read it); NULL switches it off */
void gpn_debug_df_trace(gpn_ctx* ctx, void* dev_buf);

/* ---- graph and node sets (host -> device once; resident afterwards) -------------------------------------- */
/* CSR adjacency in graph-position order (row i = i-th node of list(G)); weights_h may be NULL (all ones).
 * Replaces the per-call nx.normalized_laplacian_matrix(self.Graph) of GPnet.py:329: degrees and D^-1/2 are
 * computed once here. */
int gpn_set_graph(gpn_ctx* ctx, int N, const int* indptr_h, const int* indices_h, const double* weights_h);
/* sorted graph positions of the training / test nodes (GPnet.py:305-323 set algebra, quirk A-5) */
int gpn_set_nodes(gpn_ctx* ctx, int n_train, const int* train_h, int n_test, const int* test_h);

/* makes the n_train training targets resident on the device; GP-regression calls with t_h == NULL use them
 * (GPnetBase.set_training_values, GPnet.py:554-556) */
int gpn_set_targets(gpn_ctx* ctx, const double* t_h, int n);

/* ---- device-level primitives (unit-testable; device pointers) ------------------------------------------- */
/* out = gamma*I + alpha*L~ with L~ = D^-1/2 (D-A) D^-1/2 (GPnet.py:329 and the operands of :340,:342,:346) */
int gpn_laplacian_dense(gpn_ctx* ctx, double alpha, double gamma, double* out, long ld);
/* C = alpha*op(A)*op(B) + beta*C; transX != 0 means the operand is stored transposed ([K][M] resp. [N][K]).
 * Replaces the dgemm behind np.dot / np.linalg.matrix_power (GPnet.py:346). */
int gpn_gemm_f64(gpn_ctx* ctx, int transa, int transb, int M, int N, int K, double alpha, const double* A, long lda,
                 const double* B, long ldb, double beta, double* C, long ldc);
/* lower Cholesky in place (strict upper zeroed); replaces np.linalg.cholesky (GPnetRegressor.py:123,141;
 * GPnetClassifier.py:110,159,208).  info_h receives 0 or the failing minor. */
int gpn_potrf_f64(gpn_ctx* ctx, int n, double* A, long lda, int* info_h);
/* out = inverse of the SPD matrix A (A is destroyed: it holds L afterwards); replaces scipy.linalg.inv (GPnet.py:342)
 * and the inv(K) of GPnetClassifier.py:126.  work: n*ld doubles of scratch. */
int gpn_spd_inverse_f64(gpn_ctx* ctx, int n, double* A, long lda, double* out, long ldo, double* work, int* info_h);
/* C = alpha*A*A^T + beta*C for the n x k row-major A, full symmetric result (lower tiles computed, mirrored): the SYRK
 * behind (K^2)[tr,tr] and the expm / inverse products of symmetric operands. */
int gpn_syrk_f64(gpn_ctx* ctx, int n, int k, double alpha, const double* A, long lda, double beta, double* C, long ldc);
/* Cholesky with right-hand sides in ONE dataflow kernel (LAPACK posv-like): A = L L^T in place (strict upper zeroed) and,
 * if B != NULL, B <- B L^-T for the brows x n row-major B (whose rows up to the next multiple of 128 must exist and be
 * zero), and, if Winv_t != NULL, Winv_t <- L^-T (round128(n) x n buffer; its upper 128-tiles are written, the lower ones
 * left alone).  Replaces np.linalg.cholesky followed by the np.linalg.solve(L, .) calls of GPnetRegressor.py:123-126,
 * 141-144 and GPnetClassifier.py:110-119 in one launch; the appended rows fill the SMs the Cholesky's dependency chain
 * leaves idle.  All leading dimensions even, bases 16-byte aligned. */
int gpn_posv_f64(gpn_ctx* ctx, int n, double* A, long lda, double* B, long ldb, int brows, double* Winv_t, long ldw, int* info_h);
/* B <- B L^-T (brows x n, row-major) with the factor L that the gpn_potrf_f64 / gpn_posv_f64 call IMMEDIATELY preceding on
 * this context left in place (its inverted diagonal blocks are still resident; -2 otherwise): the triangular solves
 * np.linalg.solve(L, .) of GPnetRegressor.py:124,126,144 and GPnetClassifier.py:117-119,161-162,211 as a TRSM. */
int gpn_trsm_f64(gpn_ctx* ctx, int n, const double* L, long ldl, double* B, long ldb, int brows);
/* out = (L L^T)^-1, full symmetric, from the factor of the immediately preceding gpn_potrf_f64 / gpn_posv_f64 (LAPACK potri):
 * the explicit K^-1 of the intended gradient (old_implementation/GPR.py:56) and the inv(K) of GPnetClassifier.py:126. */
int gpn_potri_f64(gpn_ctx* ctx, int n, const double* L, long ldl, double* out, long ldo);
/* out = expm(A) for a symmetric negative semi-definite N x N matrix (Pade 3/5/7/9/13 with scaling and squaring, all products
 * on the DMMA GEMM): the scipy.linalg.expm of GPnet.py:340 as a primitive.  m_used_h / s_used_h (may be NULL): Pade order and
 * number of squarings.  Synchronises.  Uses the context's N x N scratch (a resident full kernel is dropped). */
int gpn_expm_sym_f64(gpn_ctx* ctx, int N, const double* A, long lda, double* out, long ldo, int* m_used_h, int* s_used_h);
/* out = B^p for a symmetric N x N matrix in numpy's binary-powering order (np.linalg.matrix_power, GPnet.py:346); out_pm1
 * (may be NULL) receives B^(p-1), the by-product of the p-step kernel's derivative.  Synchronises; scratch as above. */
int gpn_matpow_f64(gpn_ctx* ctx, int N, const double* B, long ldb, int p, double* out, long ldo, double* out_pm1, long ldo1);
/* Upper bound of the device memory (bytes) a context allocates for a model with N graph nodes, n training and m test nodes
 * and the given kernel (host-only arithmetic, no context required): lets a caller size shards for the 180 GB of one B200
 * (C5: N = 32768, n = m = 16384, diffusion -> 1.2e11; measured 1.06e11 for the regressor alone). */
long long gpn_workspace_bytes(int N, int n, int m, int kind);
/* out[i][j] = K[rows[i]][cols[j]] + addc  (GPnet.py:362-365: np.delete x2 + scalar noise on every entry) */
int gpn_gather_block_add(gpn_ctx* ctx, const double* K, long ld, const int* rows, int nr, const int* cols, int nc,
                         double addc, double* out, long ldo);

/* ---- full kernel (GPnet.py:335-349) ------------------------------------------------------------------------ */
/* Builds the N x N kernel for the parameters etheta_h[0..P) = np.exp(theta) and keeps it resident.  Parameter-domain
 * violations return -100 (diffusion: exp(theta0) >= 1, GPnet.py:339) / -101 (p-step: exp(theta0) < 2, GPnet.py:344) so the
 * host can raise the reference's AssertionError.  want_grad keeps the derivative by-products. */
int gpn_kernel_build(gpn_ctx* ctx, int kind, const double* etheta_h, int P, int want_grad);
/* device pointer / leading dimension of the resident full kernel */
const double* gpn_kernel_ptr(gpn_ctx* ctx, long* ld);
/* Pade order and number of squarings used by the last diffusion build */
void gpn_expm_info(gpn_ctx* ctx, int* m, int* s);
/* K[rows][cols] + addc copied to the host (the return value of GPnetBase.kernel, GPnet.py:362-367); positions outside
 * [0, N) return -2 (rows) / -4 (columns): the reference raises IndexError there */
int gpn_kernel_block_h(gpn_ctx* ctx, const int* rows_h, int nr, const int* cols_h, int nc, double addc, double* out_h);

/* ---- GP regression (GPnetRegressor.py:79-150) ---------------------------------------------------------------- */
/* out_h[0] = -logp (GPnetRegressor.py:136-150, quirk A-8: *info_h>0 means Cholesky failed and the caller returns
 * -inf); with want_grad, out_h[1..P] = d(-logp)/d theta_j (analytic form of old_implementation/GPR.py:52-64).
 * Rebuilds the full kernel from theta (as the reference does on every call). t_h: n_train un-shifted targets
 * (NULL: the resident targets of gpn_set_targets). */
int gpn_gpr_logp_grad(gpn_ctx* ctx, int kind, const double* etheta_h, int P, const double* t_h, int want_grad,
                      double* out_h, int* info_h);
/* Asynchronous form of gpn_gpr_logp_grad for batches of independent evaluations (multi-start, finite-difference probes,
 * landscape rows: the loops of GPnet.py:259-274,539-552): _begin enqueues the evaluation on this context's streams and returns,
 * _end waits for it and delivers out_h / info_h as gpn_gpr_logp_grad does.  On the whole-graph dissection route (level 6)
 * nothing synchronises in between, so several contexts bound to the same graph ("lanes") evaluate different thetas at the
 * same time and fill the SMs that one latency-bound evaluation leaves idle; every other route completes inside _begin.
 * gpn_set_sm_budget limits the SMs a lane fills with co-resident factorisation kernels (0: the whole chip): concurrent lanes
 * must not oversubscribe the chip, or their cooperative launches wait for one another. */
int gpn_gpr_logp_grad_begin(gpn_ctx* ctx, int kind, const double* etheta_h, int P, const double* t_h, int want_grad);
int gpn_gpr_logp_grad_end(gpn_ctx* ctx, double* out_h, int* info_h);
void gpn_set_sm_budget(gpn_ctx* ctx, int sms);
int gpn_num_sms(gpn_ctx* ctx);

/* ---- spectral fast path for sweeps (the reference's write-up, eq. 55-56: every Kondor kernel of GPnet.py:338-349
 * is V r(lam) V^T with L~ = V diag(lam) V^T) and batched small-n evaluation --------------------------------------
 * gpn_spectral_prepare: eigendecomposition of the bound graph's normalized Laplacian, once per graph (two-sided block Jacobi:
 * 64 x 64 pivot blocks diagonalised in shared memory, rotations applied with batched DMMA GEMMs); *sweeps_h = Jacobi sweeps,
 * *offdiag_rel_h = ||offdiag||_F / ||L~||_F reached.  gpn_spectral_enable(ctx, 1): every later kernel build of this context
 * (gpn_kernel_build and all entry points that build a kernel: predict, logp, Newton loop, landscape) forms K = V r(lam) V^T with
 * ONE symmetric GEMM (N^3 flops) instead of the Pade / inverse / power algorithms; results agree with them to ~1e-12 (own
 * parity tests).  gpn_spectral_get copies lam (N, ascending) and V^T (N x N row-major: row k = eigenvector k) to the host. */
int gpn_spectral_prepare(gpn_ctx* ctx, int* sweeps_h, double* offdiag_rel_h);
void gpn_spectral_enable(gpn_ctx* ctx, int on);
int gpn_spectral_get(gpn_ctx* ctx, double* lam_h, double* VT_h);
/* GPnetRegressor.logPosterior (+ gradient) for B parameter vectors in ONE launch, one CTA per theta (training block in shared
 * memory: n_train <= GPN_SPECTRAL_BATCH_MAX_N, else status -20): the loops of GPnet.py:259-274 / 539-552 around
 * GPnetRegressor.py:136-150 for the small-n regime where one evaluation is launch- and latency-bound.  etheta_h: B x P
 * exponentiated parameters, row-major; t_h: targets (NULL: resident); out_h: B x (1+P) rows (-logp, -dlogp/dtheta) like
 * gpn_gpr_logp_grad; info_h[b] > 0: K_y of row b not positive definite.  Status -100 / -101: parameter domain (GPnet.py:339,344). */
#define GPN_SPECTRAL_BATCH_MAX_N 112
int gpn_gpr_logp_grad_batch(gpn_ctx* ctx, int kind, const double* etheta_h, int B, int P, const double* t_h, int want_grad,
                            double* out_h, int* info_h);
/* GPnetClassifier.NRiteration (GPnetClassifier.py:90-147; the loop GPnet.py:539-552 runs per grid point for a classifier) for B
 * parameter vectors in ONE launch, one CTA per theta, same size limit and statuses as gpn_gpr_logp_grad_batch: y_h = the +-1
 * labels of the training nodes, tol / phif0 / scale = the reference's defaults 0.1 / 1e100 / 1.0; logq_h[b], iters_h[b] (Newton
 * iterations), info_h[b] (> 0: B not positive definite in some iteration, as np.linalg.cholesky raising inside the loop). */
int gpn_gpc_newton_batch(gpn_ctx* ctx, int kind, const double* etheta_h, int B, int P, const double* y_h, double tol, double phif0,
                         double scale, double* logq_h, int* iters_h, int* info_h);
/* GPnetRegressor.predict body (:86-128).  Outputs (host): alpha[n], fstar[m], s[m] (sqrt of the predictive
 * variance), kss_diag[m]; info_h[0]: Cholesky info of k, info_h[1]: info of kstarstar (the two PD gates, A-7). */
int gpn_gpr_predict(gpn_ctx* ctx, int kind, const double* etheta_h, int P, const double* t_h, double* alpha_h,
                    double* fstar_h, double* s_h, double* kss_diag_h, int* info_h);
/* after gpn_gpr_predict: which = 0 k (n x n), 1 kstar (m x n), 2 L (n x n), 3 v (n x m), 4 kstarstar (m x m) */
int gpn_gpr_fetch(gpn_ctx* ctx, int which, double* out_h);

/* Route of gpn_gpr_logp_grad for the regularized-Laplacian kernel (all agree to rounding):
 *   0  full kernel: K = (I + beta L~)^-1 formed, block gathered, K_y factored (the reference's own order of operations);
 *   1  block path: training nodes numbered last, only K[tr,:] formed from the triangular inverse;
 *   2  Schur route: K_tt^-1 = L_tt L_tt^T is read off the Cholesky factor of I + beta L~, the all-entries noise term is a
 *      rank-one update, and the gradient traces reduce to norms of the bottom block-row of L^-1;
 *   3  the same, and for gradient evaluations the other nodes are eliminated through the SPARSE coupling block of
 *      I + beta L~: Z = M_oo^-1, X = M_to Z (sparse x dense), S = M_tt - X M_ot, W_to = -W_tt X;
 *   4  as 3, with the triangular inverse / triangular product that follow the two Cholesky factorisations carried as
 *      appended row blocks of the dataflow Cholesky kernel itself ([M_oo; I] and [S; X^T; I]);
 *   5  as 4, with M_oo^-1 assembled from two independent halves of a level-set dissection of the other nodes (two
 *      concurrent half-size factorisations and a rank-|separator| correction); falls back to 4 where no thin separator exists;
 *   6  (default) no dense Schur complement at all: log det S = log det M - log det M_oo, the quadratic forms through two
 *      solves with M_oo, the gradient trace as tr(M^-1) - tr(M_oo^-1) - n, all from ONE augmented factorisation [M_kk; I] per
 *      part of a K-way level-set dissection of the whole graph (K concurrent dataflow Cholesky kernels) plus the small
 *      separator complements; falls back to 5 where the graph has no thin separators. */
void gpn_set_fast_paths(gpn_ctx* ctx, int level);
/* Number of parts of the whole-graph dissection of level 6: 0 (default) chooses it from the graph size (about 2048 nodes per
 * part, at most 8) and requires N >= 2048; a value in 2..8 forces it and relaxes the size thresholds (tests, tuning). */
void gpn_set_dissection_parts(gpn_ctx* ctx, int parts);
/* Route the last regressor likelihood evaluation (gpn_gpr_logp_grad, landscape rows) actually took, after the fallbacks a
 * level applies on its own (e.g. 5 -> 4 where the other nodes have no thin separator); -1 before the first evaluation.
 * parts_h (may be NULL; room for 16 ints): [parts of the node dissection used (0: none), separator size, part sizes ...]. */
int gpn_last_route(gpn_ctx* ctx, int* parts_h);

/* ---- GP classification, Laplace approximation (GPnetClassifier.py:90-147,184-240) ---------------------------- */
/* Newton loop on K = kernel(train,train) (built from theta unless K_dev != NULL, in which case the n x n device
 * matrix is used as is -- the notebook KAT path).  variant 0 = GPnetClassifier.py:138-145 logq, 1 = the older
 * expression of scripts/prova.py:197-200.  Outputs (host): f[n], a[n], logq, iterations. */
int gpn_gpc_newton(gpn_ctx* ctx, int kind, const double* etheta_h, int P, const double* y_h, int n, const double* K_dev,
                   long ldk, double tol, double phif0, double scale, int variant, double* f_h, double* a_h, double* logq_h,
                   int* iters_h, int* info_h);
/* GPnetClassifier.predict (:184-233): Newton + fstar[m], V[m], pi_star[m] */
int gpn_gpc_predict(gpn_ctx* ctx, int kind, const double* etheta_h, int P, const double* y_h, double* f_h, double* a_h,
                    double* logq_h, int* iters_h, double* fstar_h, double* V_h, double* pistar_h, int* info_h);

/* GPnetClassifier.gradLogPosterior (:149-182, RW Alg. 5.1 as written there) with the derivative kernels of SURVEY App. C:
 * out_h[0] = -logq (as logPosterior), out_h[1..P] = -gradZ. */
int gpn_gpc_grad(gpn_ctx* ctx, int kind, const double* etheta_h, int P, const double* y_h, double* out_h, int* iters_h,
                 int* info_h);

/* ---- landscape (GPnet.py:539-552) ------------------------------------------------------------------------------ */
/* For the `count` flat grid indices idx_h[q] (idx = i1*n2 + i2): etheta[ax0] = ax1_h[i1], etheta[ax1] = ax2_h[i2] (in this
 * order, as the reference writes them; the axes are exponentiated by the host like etheta_h); out_h[q*(1+P) ..] = (lml = +logp, dlogp/dtheta) (gradient only if want_grad).
 * classifier != 0 evaluates the Laplace loop's logq instead (y in t_h; no gradient).
 * Points that differ only in the noise entry theta[P-1] share their kernel (the noise is added to every entry, quirk A-3):
 * they are evaluated back to back with ONE kernel build, and for the regressor with ONE factorisation of K_tt -- K_y is a
 * rank-one update of it, so every noise value of a grid row is host arithmetic on seven device reductions.
 * gpn_set_fast_paths(ctx, 0) restores one independent evaluation per point.  Used by the multi-GPU sharded landscape. */
int gpn_landscape_points(gpn_ctx* ctx, int kind, int classifier, const double* etheta_h, int P, int ax0, int ax1,
                         const double* ax1_h, int n1, const double* ax2_h, int n2, const double* t_h, const long* idx_h,
                         long count, int want_grad, double* out_h);
/* The same for idx = idx_begin + q*idx_stride < idx_end. */
int gpn_landscape(gpn_ctx* ctx, int kind, int classifier, const double* etheta_h, int P, int ax0, int ax1, const double* ax1_h,
                  int n1, const double* ax2_h, int n2, const double* t_h, long idx_begin, long idx_end, long idx_stride,
                  int want_grad, double* out_h);

#ifdef __cplusplus
}
#endif
#endif /* GPNET_B200_H */
