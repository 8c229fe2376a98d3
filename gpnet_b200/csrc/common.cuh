// gpnet_b200 -- shared declarations of the sm_100a library (internal; the public C-ABI is include/gpnet_b200.h).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "../../include/gpnet_b200.h"

// Status convention: 0 ok, >0 LAPACK-style info, <0 bad argument, <=-1000 CUDA error.
#define GPN_CUDA_ERR(e) (-1000 - (int)(e))
#define GPN_CK(call)                                                                          \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            gpn_set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));                   \
            return GPN_CUDA_ERR(_e);                                                          \
        }                                                                                     \
    } while (0)
#define GPN_TRY(call)                                                                         \
    do {                                                                                      \
        int _s = (call);                                                                      \
        if (_s != 0) return _s;                                                               \
    } while (0)

void gpn_set_last_error(const char* file, int line, const char* msg);

// NVTX range per C-ABI call: header-only NVTX3, a no-op unless a profiler injects itself
#include <nvtx3/nvToolsExt.h>
struct GpnRange {
    explicit GpnRange(const char* name) { nvtxRangePushA(name); }
    ~GpnRange() { nvtxRangePop(); }
};
#define GPN_RANGE() GpnRange _gpn_range(__func__)

// A growable device buffer owned by the context.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// GEMM flags (tile-level structure exploitation; see gemm_f64.cu)
enum : int {
    GF_LOWER = 1,    // compute only tiles with tile_m >= tile_n (square C, BM == BN)
    GF_MIRROR = 2,   // with GF_LOWER: also store the transposed tile (symmetric result)
    GF_KLO_M = 4,    // contributions only for k >= m0   (op(A)[m][k] == 0 for k < m)
    GF_KLO_N = 8,    // contributions only for k >= n0   (op(B)[k][n] == 0 for k < n)
    GF_KHI_M = 16,   // contributions only for k <  m0+BM (op(A)[m][k] == 0 for k > m)
    GF_KHI_N = 32,   // contributions only for k <  n0+BN (op(B)[k][n] == 0 for k > n)
    GF_FORCE_SMALL = 64,   // force the 64x64 tile configuration
    GF_FORCE_BIG = 128,    // force the 128x128 tile configuration
    // debugging aids (GPNET_B200_GEMM_DEBUG, tools/gemm_race.py)
    GF_DBG_PROD_FENCE = 0x100, GF_DBG_CONS_FENCE = 0x200, GF_DBG_ONE_CTA = 0x400, GF_DBG_PROD_STAY = 0x800,
};

struct ProfEvent {
    cudaEvent_t e0, e1;
    double flops;
    bool big;
    int kind = 0;   // 0 GEMM launch, 2 dataflow Cholesky
};

// K-way level-set dissection of the WHOLE graph (level 6 of the regularized-Laplacian regressor likelihood, dissect.cu):
// the nodes are numbered [P_1 | ... | P_K | Sep], no edge joins two different parts, every part is numbered [others | training]
// and so is the separator.  M = I + beta L~ is then block-arrow, and the leading block of every M_kk (and of Sep) is the
// corresponding block of M_oo: ONE factorisation per part serves both M and M_oo.
constexpr int D6_MAXP = 12;      // parts of the whole-graph dissection (each has its own stream, flag slot and info slot)
constexpr int D6_SLOTS = 16;     // streams / flag slots: parts 0..11, 15 = other-node separator complement
struct Dissect6 {
    bool built = false;          // the analysis below was run for the current graph / node set
    bool usable = false;         // ... and found thin separators
    int K = 0;                   // parts
    int s = 0, so = 0;           // separator nodes, of which other (non-training) nodes (numbered first)
    int ND = 0;                  // N - s: nodes in the parts
    int off[D6_MAXP + 1] = {0};  // first row of part k (even), off[K] = ND
    int nk[D6_MAXP] = {0}, ok[D6_MAXP] = {0};   // nodes of part k, of which others (numbered first)
    int strips = 0, cells = 0;   // the parts are `strips` BFS strips of the graph, each cut again into `cells` pieces (strips x cells = K)
    int c0[D6_MAXP] = {0};             // first row of part k (multiple of 128) that has a neighbour in the separator: inside the other-node
                                 // group and inside the training group the interior nodes are numbered before the boundary nodes,
                                 // so G = W C is zero above row c0 of the part
    long long sep_nnz = 0;
    DevBuf indptr, indices, weights, dinv;   // the permuted graph (CSR, D^-1/2 and degrees)
    DevBuf trank;                // int[N]: rank of the permuted node among the training nodes (index into t), -1 for others
    DevBuf sep_ptr, sep_idx, sep_w;          // CSR of the (Sep rows) x (part columns) block, columns ascending; sep_w = dinv_r (-w) dinv_c
    DevBuf sep_pp;               // int[s x (K+1)]: entry range of part k inside row r is [sep_pp[r(K+1)+k], sep_pp[r(K+1)+k+1])
    DevBuf work[12];             // G^T, G_o^T, U^T, U_o^T, T, T_o and their factors / inverses, vectors
};

// Spectral fast path (spectral.cu): L~ = V diag(lam) V^T of the bound graph, computed once (two-sided block Jacobi); every
// kernel of the reference is then V r(lam) V^T.
struct Spectral {
    bool ready = false;          // VT / lam hold the eigendecomposition of the bound graph
    bool on = false;             // kernel_build goes through V r(lam) V^T (gpn_spectral_enable)
    bool vt_valid = false;       // Vt holds the training columns of VT for the bound node set
    int sweeps = 0;              // Jacobi sweeps the decomposition took
    double off_rel = 0.0;        // ||offdiag||_F / ||A||_F it stopped at
    std::vector<double> h_lam;   // eigenvalues, ascending
    DevBuf VT;                   // N x pad128(N): row k = eigenvector k
    DevBuf lam, rbuf;            // eigenvalues; r(lam) staging (2 x pad128(N))
    DevBuf Vt;                   // N x n4: Vt[k][i] = VT[k][train_i] (batched evaluator)
    DevBuf bt, bth, bout;        // batched evaluator: targets, thetas, results
};

struct gpn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t stream2 = nullptr;      // side stream: independent GEMM work overlapped with a latency-bound Cholesky
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    long flags_rows = 0;                 // layout of `flags`: [ready flags (flags_rows ints) | task counters]
    int potrf_grid_cap = 0;              // > 0: the next dataflow Cholesky may use at most this many CTAs
    int sm_budget = 0;                   // > 0: SMs this context may fill with co-resident (cooperative) factorisations: several
                                         // contexts ("lanes") evaluating different thetas at once share the chip (gpn_set_sm_budget)
    cudaEvent_t ev_done = nullptr;       // end of an evaluation started by gpn_gpr_logp_grad_begin
    int pend_state = 0;                  // 0 none, 1 asynchronous level-6 evaluation in flight, 2 result already on the host
    int pend_P = 0, pend_grad = 0, pend_info = 0, pend_status = 0;
    double pend_out[9] = {0};
    int num_sms = 148;
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    long long launches = 0;        // kernels launched through this context (bench "gpu_launches")
    double flops = 0.0;            // executed tensor flops accounted by the GEMM launcher
    bool profile = false;          // record a CUDA-event pair around every GEMM launch (bench roofline leg)
    std::vector<ProfEvent> prof_events;

    // graph (resident)
    int N = 0;
    long long nnz = 0;
    DevBuf indptr, indices, weights, dinv;

    // training / test sets (resident)
    int n_train = 0, n_test = 0;
    DevBuf train_idx, test_idx, tvec;

    // host copies (graph CSR, sorted training positions) and the "training nodes last" permuted graph of the
    // regularized-Laplacian fast path (built lazily; see gpr_logp_grad_reglap_fast in gp.cu)
    std::vector<int> h_indptr, h_indices, h_train;
    std::vector<double> h_weights;
    bool perm_valid = false;
    bool train_unique = false;
    int last_route = -1;                 // route the last regressor likelihood evaluation actually took (gpn_last_route)
    const void* winv_for = nullptr;      // matrix whose inverted diagonal blocks `winv` holds (set by gpn_potrf_f64 / gpn_posv_f64
    int winv_n = 0;                      // for gpn_trsm_f64 / gpn_potri_f64; any other factorisation clears it)
    int fast_reglap = 6;                 // regressor, regularized Laplacian: 0 full kernel, 1 block route, 2 Schur route, 3 Schur route with
                                         // sparse coupling, 4 the same with the triangular work appended to the
                                         // two Cholesky kernels, 5 the same with M_oo^-1 from two independent halves
                                         // of a dissection of the other nodes, 6 no dense Schur complement at all: log-determinants,
                                         // traces and solves of M and M_oo from one factorisation per part of a K-way dissection
                                         // of the whole graph (GPNET_B200_REGLAP_FAST)
    DevBuf p_indptr, p_indices, p_weights, p_dinv, iota;
    // level-set dissection of the other nodes (ensure_perm): they are numbered [A | B | Sep], no edge joins A and B
    int dis_a = 0, dis_b = 0, dis_s = 0;   // block sizes (dis_s == 0: not dissected)
    DevBuf sa_ptr, sa_idx, sa_val;         // CSR of the (Sep rows) x (A columns) block; values per theta
    DevBuf sb_ptr, sb_idx, sb_idx_rel, sb_val;   // (Sep rows) x (B columns): absolute column ids and ids relative to B
    DevBuf to_indptr, to_idx, to_val;    // CSR of the (training rows) x (other columns) block of the permuted graph; values per theta
    long long to_nnz = 0;

    // full-kernel state
    int kind = -1;
    int P = 0;
    double theta_exp[8] = {0};
    int pstep_p = 0;
    int expm_m = 0, expm_s = 0;
    bool have_full = false;
    bool have_aux = false;       // derivative by-product present (A = -beta*L~ for diffusion, B^(p-1) for p-step)
    long ldN = 0;

    // named work buffers (N x N unless noted)
    DevBuf mat[10];
    DevBuf nb[16];               // n-sized / n x n / n x m buffers of the GP algebra
    DevBuf winv;                 // inverted diagonal blocks of the running Cholesky (N x 128)
    DevBuf flags;                // dataflow Cholesky: epoch-stamped tile-ready flags (T x T ints)
    int epoch = 0;
    DevBuf flags2;               // the same for a second dataflow Cholesky running concurrently on the side stream (slot 1)
    int epoch2 = 0;
    long flags_rows2 = 0;
    // slots 2..9: further concurrent factorisations (one per part of the whole-graph dissection), each on its own side stream
    DevBuf flagsK[D6_SLOTS];
    int epochK[D6_SLOTS] = {0};
    long flags_rowsK[D6_SLOTS] = {0};
    cudaStream_t side[D6_SLOTS] = {nullptr};
    cudaEvent_t ev_side[D6_SLOTS] = {nullptr};
    // a second stream per part: the U_k products run beside the G_k^T G_k products of the same part (both only need G_k)
    cudaStream_t side2[D6_SLOTS] = {nullptr};
    cudaEvent_t ev_g[D6_SLOTS] = {nullptr}, ev_u[D6_SLOTS] = {nullptr};
    Spectral spec;
    Dissect6 d6;
    int d6_parts = 0;                    // > 0: number of parts forced by gpn_set_dissection_parts (0: chosen from N)
    void* df_trace = nullptr;    // optional device buffer: 8 x int64 globaltimer stamps per dataflow-Cholesky task
    void* leaf_prof = nullptr;   // optional device buffer (32 x int64) receiving clock64 stamps of the leaf phases
    int potrf_mode = 0;          // 0: dataflow kernel (default), 1: recursive multi-launch (GPNET_B200_POTRF=recursive)
    DevBuf scal;                 // small device scalars (reductions, info)
    DevBuf partial;              // reduction partials
    void* pinned = nullptr;      // pinned host staging for scalar read-back
    size_t pinned_cap = 0;
};

int gpn_buf_ensure(gpn_ctx* c, DevBuf& b, size_t bytes);
int gpn_winv_ensure(gpn_ctx* c, size_t bytes);

// ---- internal device-level primitives (all stream-ordered on c->stream, device pointers) -----------------
// C[M x N] = alpha * op(A) op(B) + beta * C.   a_kmajor: A stored [M][K] (row-major, K contiguous); else [K][M].
// b_kmajor: B stored [N][K]; else [K][N].  Batched over `batch` problems with element strides sA/sB/sC.
int gpn_gemm(gpn_ctx* c, int M, int N, int K, double alpha, const double* A, long lda, bool a_kmajor, const double* B,
             long ldb, bool b_kmajor, double beta, double* C, long ldc, int flags, int batch = 1, long sA = 0,
             long sB = 0, long sC = 0, double* CT = nullptr, long ldct = 0, long sCT = 0,    // CT: optional transposed copy
             int klo_shift = 0);   // GF_KLO_M relative to row m0 - klo_shift (trapezoidal op(A))

// Cholesky (lower, in place, row-major), triangular inverse and SPD inverse.  info: device int (0 = ok).
int gpn_potrf(gpn_ctx* c, int n, double* A, long lda, double* winv, int* d_info, int info_offset = 0);
// X (m x k, ldx) <- X * L^-T for the factored k x k lower L (ldl) with its inverted diagonal blocks in winv
int gpn_trsm(gpn_ctx* c, int m, int k, double* X, long ldx, const double* L, long ldl, const double* winv);
int gpn_leaf_bench(gpn_ctx* c, double* A, int copies, double* winv, int* d_info, long long* prof);
int gpn_potrf_dataflow(gpn_ctx* c, int n, double* A, long lda, double* winv, int* d_info, int info_offset, double* B = nullptr, long ldb = 0,
                       int brows = 0, double* Wt = nullptr, long ldwt = 0, int slot = 0);   // B <- B L^-T, Wt <- L^-T (upper tiles): see potrf_dataflow.cu
// W = L^-1 (normal, lower) and WT = W^T (upper) from L and the inverted diagonal blocks; TT: n x n scratch.  All ld = ldw.
int gpn_trtri(gpn_ctx* c, int n, const double* L, long ldl, const double* winv, double* W, double* WT, double* TT, long ldw);
// out (lower [+ mirrored upper]) = W^T W computed from WT = W^T (both operands K-major)
int gpn_lauum_t(gpn_ctx* c, int n, const double* WT, long ldw, double* out, long ldo, bool mirror);

// elementwise / reductions (elementwise.cu)
int gpn_k_fill_zero(gpn_ctx* c, double* A, long ld, int rows, int cols);
int gpn_k_laplacian_scatter(gpn_ctx* c, int N, const int* indptr, const int* indices, const double* w, const double* dinv,
                            double alpha, double gamma, double* out, long ld);
int gpn_k_degree_dinv(gpn_ctx* c, int N, const int* indptr, const int* indices, const double* w, double* dinv);
// pad_r/pad_c < 0: write zeros up to the next multiple of 128 rows / min(next multiple of 128, ldo) columns
int gpn_k_gather(gpn_ctx* c, const double* K, long ld, const int* rows, int nr, const int* cols, int nc, double addc,
                 double* out, long ldo, int pad_r = -1, int pad_c = -1);
int gpn_k_gather_rows(gpn_ctx* c, const double* K, long ld, const int* rows, int nr, int ncols, double* out, long ldo);
int gpn_k_zero_upper(gpn_ctx* c, double* A, long ld, int n);
int gpn_k_lincomb(gpn_ctx* c, int n, double* out, long ldo, double diag, int nterms, const double* const* X, const long* ldx,
                  const double* coef);
int gpn_k_onenorm(gpn_ctx* c, int n, const double* A, long ld, double* d_out /*device scalar*/);
int gpn_k_copy(gpn_ctx* c, int rows, int cols, const double* A, long lda, double* B, long ldb);
int gpn_k_gemv(gpn_ctx* c, int rows, int cols, const double* A, long lda, const double* x, double* y, int mode);
int gpn_k_gemv_t(gpn_ctx* c, int rows, int cols, const double* A, long lda, const double* x, double* y, int mode);
int gpn_k_dot(gpn_ctx* c, int n, const double* x, const double* y, double* d_out);
int gpn_k_sum(gpn_ctx* c, int n, const double* x, double* d_out);
int gpn_k_sumlogdiag(gpn_ctx* c, int n, const double* A, long ld, double* d_out);
int gpn_k_fill_vec(gpn_ctx* c, int n, double* y, double v);
int gpn_k_rowdot_pass(gpn_ctx* c, int rows, int cols, const double* Y, long ld, const double* ya, const double* yb, double* za, double* zb,
                      double* d_fro, bool upper_only, bool zero_fro);
int gpn_k_z_border(gpn_ctx* c, int D, int s, const double* V, long ldv, const double* Sigma, long ldsg, double* Z, long ldz);
int gpn_k_csr_values(gpn_ctx* c, int rows, int row0, const int* indptr, const int* idx, const double* M, long ld, double* val);
int gpn_k_spmm_rows(gpn_ctx* c, int rows, int cols, const int* indptr, const int* idx, const double* val, const double* Z, long ldz, double* X,
                    long ldx, const double* base = nullptr, long ldb = 0);   // base: X = base - (sparse rows) Z
int gpn_k_spmm_rows_t(gpn_ctx* c, int rows, int cols, const int* indptr, const int* idx, const double* val, const double* Z, long ldz,
                      double* XT, long ldxt);   // XT = (sparse rows * Z)^T
int gpn_k_schur_update(gpn_ctx* c, int n, const double* Mtt, long ldm, const int* indptr, const int* idx, const double* val, const double* X,
                       long ldx, double* S, long lds);
int gpn_k_fill_zero_lower(gpn_ctx* c, double* A, long ld, int rows, int skip_row0 = -1, int skip_cols = 0);
int gpn_k_gemv_t2(gpn_ctx* c, int rows, int cols, const double* A, long lda, const double* x1, const double* x2, double* y1, double* y2,
                  int mode);
int gpn_k_bilinear_trace(gpn_ctx* c, int n, const double* S, long lds, const double* G, long ldg, const double* Ktt, long ldk, double g_scale,
                         double k_scale, const double* x1, const double* xt, double* d_out4);
int gpn_k_trap_pass(gpn_ctx* c, int rows, int cols, int off, const double* Wb, long ld, const double* ya, const double* yb, double* za,
                    double* zb, double* d_fro, bool zero_fro = true);
int gpn_k_multi_dot(gpn_ctx* c, int count, const double* const* x, const double* const* y, const int* n, double* d_out);
int gpn_k_colnorm2(gpn_ctx* c, int rows, int cols, const double* V, long ld, double* out);
int gpn_k_rownorm2(gpn_ctx* c, int rows, int cols, const double* V, long ld, double* out);
int gpn_k_scale_cols(gpn_ctx* c, int rows, int cols, double* A, long lda, const double* s);
int gpn_k_diag_gather(gpn_ctx* c, const double* K, long ld, const int* idx, int m, double addc, double* out);
int gpn_k_grad_trace(gpn_ctx* c, int n, const double* Kinv, long ldi, const double* G, long ldg, const double* Ktt, long ldk,
                     double g_scale, double k_scale, double k_shift, const double* alpha, double* d_out3);
int gpn_k_form_B(gpn_ctx* c, int n, const double* K, long ldk, const double* sw, double* B, long ldb);
int gpn_k_scale_rows(gpn_ctx* c, int rows, int cols, const double* A, long lda, const double* s, double* out, long ldo);

enum { GEMV_FULL = 0, GEMV_LOWER = 1 /* treat entries with col > row as zero */, GEMV_UPPER = 2 /* ... with col < row as zero */ };

// spectral path (spectral.cu)
int gpn_spec_invalidate(gpn_ctx* c, bool graph_changed);
int gpn_spec_prepare(gpn_ctx* c);
int gpn_spec_build_kernel(gpn_ctx* c, int kind, double th0, int p, double* Kout, long ld, double* pm1, double* scratch);
const double* gpn_resident_targets(gpn_ctx* c);      // device copy of the targets of gpn_set_targets (gp.cu)

// whole-graph dissection route (dissect.cu)
int gpn_d6_ensure(gpn_ctx* c);
int gpn_d6_fill_zero(gpn_ctx* c, double* M, long ld);
int gpn_d6_coupling(gpn_ctx* c, const double* t, double beta, double* b1, double* bt, double* d_out5);
int gpn_d6_factor_reduce(gpn_ctx* c, const double* L, const double* WT, long ld, double* d_out4);
int gpn_d6_sep_spmm(gpn_ctx* c, int k, double beta, const double* WT, long ld, double* GT, double* GTo, long ldg);
int gpn_d6_t_reduce(gpn_ctx* c, int s, const double* E, long lde, const double* Cb, long lds, int nb, double* T);
int gpn_d6_sep_reduce(gpn_ctx* c, int s, const double* Lt, long ldl, const double* WTt, long ldw, const double* Y, long ldy, int rows,
                      double* d_out2);
int gpn_d6_tri_solve(gpn_ctx* c, int n, const double* WTf, long ld, const double* xa, const double* xb, double* va, double* vb, double* ua,
                     double* ub);
int gpn_d6_block_solve(gpn_ctx* c, const double* WT, long ld, const double* xa, const double* xb, double* va, double* vb, double* ua, double* ub);
int gpn_d6_sep_gather(gpn_ctx* c, double beta, const double* ba, const double* bb, const double* ya, const double* yb, double* ra, double* rb);
int gpn_d6_sep_scatter(gpn_ctx* c, double beta, const double* ua, const double* ub, double* wa, double* wb);
