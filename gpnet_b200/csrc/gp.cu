// GP drivers behind the C-ABI: context, graph upload, full-kernel construction (diffusion / regularized Laplacian /
// p-step walk), GP-regression log marginal likelihood + analytic gradient + predictive mean/variance, the Laplace
// Newton loop of the classifier and its predictive probabilities, and the landscape slice evaluator.
// Reference lines are cited per function; the numerical kernels live in gemm_f64.cu / factor.cu / elementwise.cu.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <algorithm>


static thread_local char g_last_error[512] = "";
void gpn_set_last_error(const char* file, int line, const char* msg) {
    snprintf(g_last_error, sizeof g_last_error, "%s:%d: %s", file, line, msg);
}

// winv must be zero above the diagonal of every 128x128 block (the leaf writes lower triangles only)
int gpn_winv_ensure(gpn_ctx* c, size_t bytes) {
    if (bytes <= c->winv.cap) return 0;
    GPN_TRY(gpn_buf_ensure(c, c->winv, bytes));
    GPN_CK(cudaMemsetAsync(c->winv.p, 0, c->winv.cap, c->stream));
    return 0;
}

int gpn_buf_ensure(gpn_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) {
        GPN_CK(cudaStreamSynchronize(c->stream));
        GPN_CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    bytes = (bytes + 255) & ~(size_t)255;
    GPN_CK(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

namespace {

constexpr double LOG_2PI = 1.8378770664093454835606594728112;

inline long pad128(long n) { return (n + 127) / 128 * 128; }

// n-sized work buffers
enum NB {
    NB_KY = 0, NB_KC, NB_W, NB_T, NB_ROWS1, NB_ROWS2, NB_G, NB_KSTAR, NB_KSS, NB_V, NB_S, NB_B, NB_WB, NB_WK, NB_WT, NB_COUNT
};
static_assert(NB_COUNT <= 16, "nb[] too small");

struct Scal {           // layout of c->scal (doubles), followed by ints
    enum { DOT = 0, LOGDET, SUM, TR0, TR1, TR2, CONV, Q, LIK, AF, NORM0, NORM1, NORM2, NORM3, SCH0 /* up to 8 reductions */, FRO = SCH0 + 8, NDBL = 32 };
};

double* dscal(gpn_ctx* c, int i) { return (double*)c->scal.p + i; }
int* iscal(gpn_ctx* c, int i) { return (int*)((double*)c->scal.p + Scal::NDBL) + i; }

int read_scalars(gpn_ctx* c, double* out_d, int nd, int* out_i, int ni) {
    // one D2H of the whole scalar block through pinned memory, then a stream sync
    const size_t bytes = sizeof(double) * Scal::NDBL + sizeof(int) * 16;
    GPN_CK(cudaMemcpyAsync(c->pinned, c->scal.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    GPN_CK(cudaStreamSynchronize(c->stream));
    if (out_d) memcpy(out_d, c->pinned, sizeof(double) * nd);
    if (out_i) memcpy(out_i, (char*)c->pinned + sizeof(double) * Scal::NDBL, sizeof(int) * ni);
    return 0;
}
int clear_info(gpn_ctx* c) {
    GPN_CK(cudaMemsetAsync(iscal(c, 0), 0, sizeof(int) * 16, c->stream));
    return 0;
}
int upload(gpn_ctx* c, void* dst, const void* src_h, size_t bytes) {
    // small host vectors: stage through pinned memory when they fit, so the copy is truly asynchronous
    GPN_CK(cudaMemcpyAsync(dst, src_h, bytes, cudaMemcpyHostToDevice, c->stream));
    GPN_CK(cudaStreamSynchronize(c->stream));   // the host buffer may be a temporary of the caller
    return 0;
}
int download(gpn_ctx* c, void* dst_h, const void* src, size_t bytes) {
    GPN_CK(cudaMemcpyAsync(dst_h, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
int download_matrix(gpn_ctx* c, double* dst_h, const double* src, long ld, int rows, int cols) {
    GPN_CK(cudaMemcpy2DAsync(dst_h, sizeof(double) * cols, src, sizeof(double) * ld, sizeof(double) * cols, rows,
                             cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

double* mat(gpn_ctx* c, int i) { return (double*)c->mat[i].p; }
int ensure_mat(gpn_ctx* c, int i) {
    const long np = pad128(c->N);
    return gpn_buf_ensure(c, c->mat[i], sizeof(double) * (size_t)np * np);
}
int ensure_nb(gpn_ctx* c, int i, long rows, long cols) { return gpn_buf_ensure(c, c->nb[i], sizeof(double) * (size_t)rows * cols); }
double* nbp(gpn_ctx* c, int i) { return (double*)c->nb[i].p; }

// ---- vector kernels of the Laplace Newton loop (GPnetClassifier.py:106-112,124-145,204-231) -------------------------
__global__ void newton_pre_kernel(int n, const double* __restrict__ f, const double* __restrict__ y, double* __restrict__ w,
                                  double* __restrict__ sw, double* __restrict__ p, double* __restrict__ b, double* __restrict__ g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double fi = f[i];
    const double e = exp(-fabs(fi));
    const double d = 1.0 + e;
    const double pi = (fi >= 0.0) ? 1.0 / d : e / d;     // sigma(f)           (:111 / :209)
    const double wi = e / (d * d);                        // sigma(1-sigma)     (:106 / :204-206)
    w[i] = wi;
    sw[i] = sqrt(wi);
    p[i] = pi;
    const double gi = 0.5 * (y[i] + 1.0) - pi;
    b[i] = wi * fi + gi;                                  // :112
    g[i] = gi;                                            // (y+1)/2 - p        (:210)
}
__global__ void vec_mul_kernel(int n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i] * y[i];
}
// a = scale * (b - sw * sol)                                                  (:113-121)
__global__ void newton_a_kernel(int n, double scale, const double* __restrict__ b, const double* __restrict__ sw,
                                const double* __restrict__ sol, double* __restrict__ a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = scale * (b[i] - sw[i] * sol[i]);
}
// phif_i = log p_i - q/2 - logdet/2 - n/2 log 2pi; conv = sum (old - new)^2 (:124-132).  Single block.
__global__ void newton_phif_kernel(int n, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ q_alt,
                                   const int* __restrict__ info_k, const double* __restrict__ logdet, double* __restrict__ phif, int first,
                                   double phif0, double* __restrict__ conv) {
    __shared__ double sh[32];
    // f^T K^-1 f through the Cholesky of K; if K is not numerically positive definite: f^T a (the same quantity, f = K a)
    const double qq = (*info_k == 0) ? *q : *q_alt;
    const double common = -0.5 * qq - 0.5 * (*logdet) - 0.5 * n * LOG_2PI;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = log(p[i]) + common;
        const double old = first ? phif0 : phif[i];
        const double d = old - v;
        s += d * d;
        phif[i] = v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *conv = s;
    }
}
// lik = sum log(1 + log(1 + exp(-s))), s = -y f  (:138-143, quirk A-10); variant 1: scripts/prova.py:197-200
__global__ void newton_lik_kernel(int n, const double* __restrict__ f, const double* __restrict__ y, int variant,
                                  double* __restrict__ out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double s = -y[i] * f[i];
        if (variant == 0) {
            acc += log(1.0 + log(1.0 + exp(-s)));
        } else {
            const double ps = s > 0.0 ? s : 0.0;
            acc += log(ps + log(exp(-ps) + exp(s - ps)));
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) *out = acc;
    }
}
// K12: V = kss - ||v||^2 and the 5-erf averaged sigmoid (GPnetClassifier.py:217-231, constants GPnet.py:59-62)
__global__ void pistar_kernel(int m, const double* __restrict__ fstar, const double* __restrict__ kss, const double* __restrict__ vn,
                              double* __restrict__ V, double* __restrict__ pistar) {
    const double LAM[5] = {0.41, 0.4, 0.37, 0.44, 0.39};
    const double COEF[5] = {-1854.8214151, 3516.89893646, 221.29346712, 128.12323805, -2010.49422654};
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double v = kss[i] - vn[i];
    V[i] = v;
    const double alpha = 1.0 / (2.0 * v);
    const double fs = fstar[i];
    double acc = 0.0, csum = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double gamma = LAM[k] * fs;
        const double integ = sqrt(M_PI / alpha) * erf(gamma * sqrt(alpha / (alpha + LAM[k] * LAM[k]))) / (2.0 * sqrt(v * 2.0 * M_PI));
        acc += COEF[k] * integ;
        csum += COEF[k];
    }
    pistar[i] = acc + 0.5 * csum;
}

// ---- vector kernels of the classifier gradient (GPnetClassifier.py:156-182) ------------------------------------------
__global__ void diag_extract_kernel(int n, const double* __restrict__ A, long ld, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = A[(long)i * ld + i];
}
// s2 = -0.5 (diag(K) - diag(C^T C)) * 2 hess (0.5 - p), hess = -w                                     (:163-168)
__global__ void gpc_s2_kernel(int n, const double* __restrict__ kdiag, const double* __restrict__ cn, const double* __restrict__ w,
                              const double* __restrict__ p, double* __restrict__ s2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s2[i] = -0.5 * (kdiag[i] - cn[i]) * (2.0 * (-w[i]) * (0.5 - p[i]));
}
// out = x - y
__global__ void vec_sub_kernel(int n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i] - y[i];
}
__global__ void vec_fill_kernel(int n, double v, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
// out = sum_ij sw_i Binv_ij sw_j D_ij   (= tr(R D) for symmetric D), one block per row
__global__ void gpc_trace_kernel(int n, const double* __restrict__ Binv, long ldb, const double* __restrict__ sw, const double* __restrict__ D,
                                 long ldd, double* __restrict__ out) {
    __shared__ double sh[32];
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) s += Binv[(long)r * ldb + c] * sw[c] * D[(long)r * ldd + c];
    s *= sw[r];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

inline int g1(long n) { return (int)((n + 255) / 256); }

// register-resident DMMA.8x8x4 loop: the FP64 tensor issue-rate ceiling the roofline is reported against
// (MEASURED_PEAKS.json has no FP64 entry).  8 independent accumulator pairs per warp, 16 warps per SM.
__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters, double a0, double b0) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
    const double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// full kernel construction (GPnet.py:335-349)
// ---------------------------------------------------------------------------------------------------------------------
int laplacian_into(gpn_ctx* c, double alpha, double gamma, double* out, long ld) {
    const long np = pad128(c->N);
    GPN_TRY(gpn_k_fill_zero(c, out, ld, (int)np, (int)np));
    return gpn_k_laplacian_scatter(c, c->N, (const int*)c->indptr.p, (const int*)c->indices.p,
                                   c->weights.p ? (const double*)c->weights.p : nullptr, (const double*)c->dinv.p, alpha, gamma, out,
                                   ld);
}

// symmetric product of two symmetric (commuting) matrices: lower tiles + mirrored store
int symm_mul(gpn_ctx* c, const double* A, const double* B, double* C) {
    const long ld = c->ldN;
    return gpn_gemm(c, c->N, c->N, c->N, 1.0, A, ld, true, B, ld, true, 0.0, C, ld, GF_LOWER | GF_MIRROR);
}

// SPD inverse of the N x N matrix in `a` (destroyed -> L); result (full symmetric) in `out`; w/t scratch
int spd_inverse(gpn_ctx* c, int n, double* a, long ld, double* w, double* wt, double* tt, double* out, int info_slot) {
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(n) * 128));
    static const int aug_max = getenv("GPNET_B200_INV_AUG_MAX") ? atoi(getenv("GPNET_B200_INV_AUG_MAX")) : 3072;
    if (n <= aug_max && c->potrf_mode == 0) {
        // small n: the factorisation is bound by its dependency chain, so W^T = L^-T rides along as the identity block of the same
        // dataflow kernel ([A; I]) instead of a dozen triangular-inverse launches behind it
        GPN_TRY(gpn_potrf_dataflow(c, n, a, ld, (double*)c->winv.p, iscal(c, info_slot), 0, nullptr, 0, 0, wt, ld, 0));
        return gpn_lauum_t(c, n, wt, ld, out, ld, true);
    }
    GPN_TRY(gpn_potrf(c, n, a, ld, (double*)c->winv.p, iscal(c, info_slot)));
    GPN_TRY(gpn_trtri(c, n, a, ld, (const double*)c->winv.p, w, wt, tt, ld));
    return gpn_lauum_t(c, n, wt, ld, out, ld, true);
}

int build_reglap(gpn_ctx* c, double beta) {
    for (int i = 0; i < 4; ++i) GPN_TRY(ensure_mat(c, i));
    GPN_TRY(laplacian_into(c, beta, 1.0, mat(c, 0), c->ldN));                       // M = I + beta L~   (:342)
    GPN_TRY(spd_inverse(c, c->N, mat(c, 0), c->ldN, mat(c, 1), mat(c, 2), mat(c, 3), mat(c, 0), 0));
    c->mat[9] = c->mat[0];   // alias: slot 9 designates the resident kernel
    return 0;
}

// np.linalg.matrix_power structure (binary powering: z squared every step, multiplied into the result on set bits);
// all factors are polynomials in B => symmetric products.  Z0/Z1 and R0/R1 are ping-pong scratch matrices.
int matpow(gpn_ctx* c, const double* src, int p, double* Z0, double* Z1, double* R0, double* R1, const double** res) {
    const double* z = src;
    const double* r = nullptr;
    double* Zb[2] = {Z0, Z1};
    double* Rb[2] = {R0, R1};
    int zi = 0, ri = 0;
    bool first = true;
    for (int e = p; e > 0; e >>= 1) {
        if (!first) {
            GPN_TRY(symm_mul(c, z, z, Zb[zi]));
            z = Zb[zi];
            zi ^= 1;
        }
        first = false;
        if (e & 1) {
            if (r == nullptr) {
                if ((e >> 1) == 0) {
                    r = z;                       // last bit: no further squaring can clobber z
                } else {
                    GPN_TRY(gpn_k_copy(c, c->N, c->N, z, c->ldN, Rb[ri], c->ldN));
                    r = Rb[ri];
                    ri ^= 1;
                }
            } else {
                GPN_TRY(symm_mul(c, r, z, Rb[ri]));
                r = Rb[ri];
                ri ^= 1;
            }
        }
    }
    *res = r;
    return 0;
}

int build_pstep(gpn_ctx* c, double a, int p, int want_grad) {
    for (int i = 0; i < 8; ++i) GPN_TRY(ensure_mat(c, i));
    const long ld = c->ldN;
    GPN_TRY(laplacian_into(c, -1.0, a, mat(c, 0), ld));                              // B = a I - L~      (:346)
    c->have_aux = false;
    if (p <= 0) {   // matrix_power(B, 0) = I
        GPN_TRY(gpn_k_fill_zero(c, mat(c, 5), ld, (int)pad128(c->N), (int)pad128(c->N)));
        const double* X[1] = {mat(c, 0)};
        const long l1[1] = {ld};
        const double cf[1] = {0.0};
        GPN_TRY(gpn_k_lincomb(c, c->N, mat(c, 5), ld, 1.0, 1, X, l1, cf));
        c->mat[9] = c->mat[5];
        return 0;
    }
    const double* res = nullptr;
    if (want_grad && p >= 2) {
        // K = B^(p-1) * B, keeping B^(p-1) for dK/dtheta0 = a p B^(p-1)
        GPN_TRY(matpow(c, mat(c, 0), p - 1, mat(c, 1), mat(c, 2), mat(c, 3), mat(c, 6), &res));
        if (res != mat(c, 4)) GPN_TRY(gpn_k_copy(c, c->N, c->N, res, ld, mat(c, 4), ld));
        GPN_TRY(symm_mul(c, mat(c, 4), mat(c, 0), mat(c, 5)));
        c->have_aux = true;       // aux = mat 4
        c->mat[9] = c->mat[5];
        return 0;
    }
    GPN_TRY(matpow(c, mat(c, 0), p, mat(c, 1), mat(c, 2), mat(c, 3), mat(c, 6), &res));
    if (res != mat(c, 5)) GPN_TRY(gpn_k_copy(c, c->N, c->N, res, ld, mat(c, 5), ld));
    c->mat[9] = c->mat[5];
    return 0;
}

// expm(-beta L~): scaling-and-squaring Pade (Al-Mohy & Higham 2009 order selection with exact 1-norms of the
// even powers; the odd-power "ell" refinement of scipy is not needed for symmetric A), all
// products symmetric, and Q X = P solved through the SPD inverse of Q (GPnet.py:340).
int expm_of_mat0(gpn_ctx* c);
int build_diffusion(gpn_ctx* c, double beta) {
    for (int i = 0; i < 8; ++i) GPN_TRY(ensure_mat(c, i));
    GPN_TRY(laplacian_into(c, -beta, 0.0, mat(c, 0), c->ldN));
    return expm_of_mat0(c);
}
// expm of the symmetric negative semi-definite N x N matrix in mat 0 (zero padded to ldN); result in mat[9]
int expm_of_mat0(gpn_ctx* c) {
    const long ld = c->ldN;
    const int N = c->N;
    double *A = mat(c, 0), *A2 = mat(c, 1), *A4 = mat(c, 2), *A6 = mat(c, 3), *Z = mat(c, 4), *U = mat(c, 5), *V = mat(c, 6),
           *X8 = mat(c, 7);
    // Order selection with as few powers as possible (each power is an N^3 product): bounds ||A^(p+q)|| <= ||A^p|| ||A^q||
    // stand in for the powers that have not been formed yet, so the choice can only err towards a HIGHER order than
    // scipy's (which uses estimated norms), never a lower one.
    double sc[Scal::NDBL];
    int m, s = 0;
    GPN_TRY(symm_mul(c, A, A, A2));
    GPN_TRY(gpn_k_onenorm(c, N, A2, ld, dscal(c, Scal::NORM1)));
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, nullptr, 0));
    const double n2 = sc[Scal::NORM1];
    bool have4 = false, have6 = false;
    if (sqrt(n2) < 1.495585217958292e-002) {          // d4, d6 <= ||A^2||^(1/2)
        m = 3;
    } else {
        GPN_TRY(symm_mul(c, A2, A2, A4));
        have4 = true;
        GPN_TRY(gpn_k_onenorm(c, N, A4, ld, dscal(c, Scal::NORM2)));
        GPN_TRY(read_scalars(c, sc, Scal::NDBL, nullptr, 0));
        const double n4 = sc[Scal::NORM2];
        const double d4 = pow(n4, 0.25);
        if (fmax(d4, pow(n2 * n4, 1.0 / 6.0)) < 2.539398330063230e-001) {   // d6 <= (||A^2|| ||A^4||)^(1/6)
            m = 5;
        } else {
            GPN_TRY(symm_mul(c, A4, A2, A6));
            have6 = true;
            GPN_TRY(gpn_k_onenorm(c, N, A6, ld, dscal(c, Scal::NORM3)));
            GPN_TRY(read_scalars(c, sc, Scal::NDBL, nullptr, 0));
            const double n6 = sc[Scal::NORM3];
            const double d6 = pow(n6, 1.0 / 6.0);
            const double d8 = fmin(d4, pow(n2 * n6, 0.125)), d10 = pow(n4 * n6, 0.1);
            const double eta1 = fmax(d4, d6), eta3 = fmax(d6, d8);
            if (eta1 < 2.539398330063230e-001) m = 5;
            else if (eta3 < 9.504178996162932e-001) m = 7;
            else if (eta3 < 2.097847961257068e+000) m = 9;
            else {
                m = 13;
                const double eta4 = fmax(d8, d10), eta5 = fmin(eta3, eta4);
                if (eta5 > 0) s = (int)fmax(ceil(log2(eta5 / 4.25)), 0.0);
            }
        }
    }
    (void)have4;
    (void)have6;
    c->expm_m = m;
    c->expm_s = s;
    static const double b3[] = {120., 60., 12., 1.};
    static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
    static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                                 10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};
    const double* X3[3] = {A2, A4, A6};
    const long l3[3] = {ld, ld, ld};
    if (m <= 9) {
        const double* b = (m == 3) ? b3 : (m == 5) ? b5 : (m == 7) ? b7 : b9;
        const int nt = (m - 1) / 2;          // number of even powers A2.. used
        const double* X4[4] = {A2, A4, A6, X8};
        const long l4[4] = {ld, ld, ld, ld};
        if (m == 9) GPN_TRY(symm_mul(c, A6, A2, X8));
        double cu[4], cv[4];
        for (int i = 0; i < nt; ++i) {
            cu[i] = b[2 * i + 3];
            cv[i] = b[2 * i + 2];
        }
        GPN_TRY(gpn_k_lincomb(c, N, Z, ld, b[1], nt, X4, l4, cu));       // Z = b1 I + b3 A2 + ...
        GPN_TRY(symm_mul(c, A, Z, U));                                     // U = A Z
        GPN_TRY(gpn_k_lincomb(c, N, V, ld, b[0], nt, X4, l4, cv));        // V = b0 I + b2 A2 + ...
    } else {
        const double* b = b13;
        const double s2 = ldexp(1.0, -2 * s), s4 = ldexp(1.0, -4 * s), s6 = ldexp(1.0, -6 * s), s1 = ldexp(1.0, -s);
        {   // U2 = B6 (b13 B6 + b11 B4 + b9 B2);  U = B (U2 + b7 B6 + b5 B4 + b3 B2 + b1 I)
            const double cf[3] = {b[9] * s2, b[11] * s4, b[13] * s6};
            GPN_TRY(gpn_k_lincomb(c, N, Z, ld, 0.0, 3, X3, l3, cf));
            GPN_TRY(gpn_gemm(c, N, N, N, s6, A6, ld, true, Z, ld, true, 0.0, X8, ld, GF_LOWER | GF_MIRROR));
            const double* X4[4] = {A2, A4, A6, X8};
            const long l4[4] = {ld, ld, ld, ld};
            const double c2[4] = {b[3] * s2, b[5] * s4, b[7] * s6, 1.0};
            GPN_TRY(gpn_k_lincomb(c, N, Z, ld, b[1], 4, X4, l4, c2));
            GPN_TRY(gpn_gemm(c, N, N, N, s1, A, ld, true, Z, ld, true, 0.0, U, ld, GF_LOWER | GF_MIRROR));
        }
        {   // V2 = B6 (b12 B6 + b10 B4 + b8 B2);  V = V2 + b6 B6 + b4 B4 + b2 B2 + b0 I
            const double cf[3] = {b[8] * s2, b[10] * s4, b[12] * s6};
            GPN_TRY(gpn_k_lincomb(c, N, Z, ld, 0.0, 3, X3, l3, cf));
            GPN_TRY(gpn_gemm(c, N, N, N, s6, A6, ld, true, Z, ld, true, 0.0, X8, ld, GF_LOWER | GF_MIRROR));
            const double* X4[4] = {A2, A4, A6, X8};
            const long l4[4] = {ld, ld, ld, ld};
            const double c2[4] = {b[2] * s2, b[4] * s4, b[6] * s6, 1.0};
            GPN_TRY(gpn_k_lincomb(c, N, V, ld, b[0], 4, X4, l4, c2));
        }
    }
    {   // Q = V - U -> Z (zero padded);  P = V + U -> X8
        GPN_TRY(gpn_k_fill_zero(c, Z, ld, (int)pad128(N), (int)pad128(N)));
        GPN_TRY(gpn_k_fill_zero(c, X8, ld, (int)pad128(N), (int)pad128(N)));       // P is an appended row block of [Q; P; I]: zero padded rows
        const double* X2[2] = {V, U};
        const long l2[2] = {ld, ld};
        const double cq[2] = {1.0, -1.0}, cp[2] = {1.0, 1.0};
        GPN_TRY(gpn_k_lincomb(c, N, Z, ld, 0.0, 2, X2, l2, cq));
        GPN_TRY(gpn_k_lincomb(c, N, X8, ld, 0.0, 2, X2, l2, cp));
    }
    // X = Q^-1 P  (Q SPD: q_m(x) = p_m(-x) has positive coefficients and spec(A) <= 0)
    static const int aug_max = getenv("GPNET_B200_EXPM_AUG_MAX") ? atoi(getenv("GPNET_B200_EXPM_AUG_MAX")) : 3072;
    if (N <= aug_max) {
        // small N: the Cholesky of Q is bound by its dependency chain, so the solve rides along in the SAME dataflow kernel:
        // [Q; P; I] -> L, Y = P L^-T, W^T = L^-T, then X = Y W = P Q^-1 = Q^-1 P (polynomials in A commute) as one product with a
        // triangular k-range.  Replaces potrf + triangular inverse (a dozen launches) + W^T W + Q^-1 P.
        GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(N) * 128));
        GPN_TRY(gpn_potrf_dataflow(c, N, Z, ld, (double*)c->winv.p, iscal(c, 0), 0, X8, ld, N, A2, ld, 0));
        GPN_TRY(gpn_gemm(c, N, N, N, 1.0, X8, ld, true, A2, ld, true, 0.0, U, ld, GF_LOWER | GF_MIRROR | GF_KLO_N));
    } else {
        GPN_TRY(spd_inverse(c, N, Z, ld, A2, A4, V, A6, 0));       // Qinv -> A6 (A2, A4, V scratch: U and V are dead once Q, P exist)
        GPN_TRY(symm_mul(c, A6, X8, U));                            // X -> U
    }
    double* cur = U;
    double* oth = V;
    for (int i = 0; i < s; ++i) {
        GPN_TRY(symm_mul(c, cur, cur, oth));
        double* t = cur; cur = oth; oth = t;
    }
    c->mat[9] = (cur == U) ? c->mat[5] : c->mat[6];
    c->have_aux = true;     // aux = A (mat 0) for dK/dtheta0 = A K
    return 0;
}

const double* kfull(gpn_ctx* c) { return (const double*)c->mat[9].p; }

int kernel_build(gpn_ctx* c, int kind, const double* theta_h, int P, int want_grad) {
    if (c->N <= 0) return -1;
    if (P < 1 || P > 8) return -4;
    c->have_full = false;
    c->kind = kind;
    c->P = P;
    for (int i = 0; i < P; ++i) c->theta_exp[i] = theta_h[i];          // = np.exp(theta)[i] of GPnet.py:303, exponentiated by the HOST (A-4)
    c->ldN = pad128(c->N);
    GPN_TRY(clear_info(c));
    int st;
    if (c->spec.on) {
        // spectral route (spectral.cu): K = V r(lam) V^T from the eigendecomposition of L~ computed once per graph
        if (kind != GPN_KERNEL_DIFFUSION && kind != GPN_KERNEL_REGULARIZED_LAPLACIAN && kind != GPN_KERNEL_PSTEP_WALK) return -2;
        if (kind == GPN_KERNEL_DIFFUSION && !(c->theta_exp[0] < 1.0)) return -100;
        if (kind == GPN_KERNEL_PSTEP_WALK) {
            if (!(c->theta_exp[0] >= 2.0)) return -101;
            if (P < 2) return -4;
            c->pstep_p = (int)c->theta_exp[1];
        }
        for (int i : {0, 1, 4, 5}) GPN_TRY(ensure_mat(c, i));
        const bool pm1 = kind == GPN_KERNEL_PSTEP_WALK && want_grad && c->pstep_p >= 2;
        c->have_aux = false;
        st = gpn_spec_build_kernel(c, kind, c->theta_exp[0], c->pstep_p, mat(c, 5), c->ldN, pm1 ? mat(c, 4) : nullptr, mat(c, 1));
        if (st != 0) return st;
        c->mat[9] = c->mat[5];
        if (kind == GPN_KERNEL_DIFFUSION) {        // the gradient reads A = -beta L~ from mat 0 (gpr_grad_kernel_block)
            GPN_TRY(laplacian_into(c, -c->theta_exp[0], 0.0, mat(c, 0), c->ldN));
            c->have_aux = true;
        } else if (pm1) {
            c->have_aux = true;
        }
        c->expm_m = 0;
        c->expm_s = 0;
        c->have_full = true;
        return 0;
    }
    if (kind == GPN_KERNEL_DIFFUSION) {
        if (!(c->theta_exp[0] < 1.0)) return -100;                        // GPnet.py:339
        st = build_diffusion(c, c->theta_exp[0]);
    } else if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN) {
        st = build_reglap(c, c->theta_exp[0]);
    } else if (kind == GPN_KERNEL_PSTEP_WALK) {
        if (!(c->theta_exp[0] >= 2.0)) return -101;                       // GPnet.py:344
        if (P < 2) return -4;
        c->pstep_p = (int)c->theta_exp[1];                                // int(theta[1]) after exp, quirk A-4
        st = build_pstep(c, c->theta_exp[0], c->pstep_p, want_grad);
    } else {
        return -2;                                                        // "kerneltype not implemented" (:337)
    }
    if (st == 0) c->have_full = true;
    return st;
}

// ---------------------------------------------------------------------------------------------------------------------
// GP regression
// ---------------------------------------------------------------------------------------------------------------------
struct GprFactor {
    int n;
    long ldn;
    double *Ky, *W, *T, *WT;
};

// Ky = K[tr,tr] + c (NB_KY); keep_copy: NB_KC = Ky before factorisation.
int gpr_gather_ky(gpn_ctx* c, bool keep_copy) {
    const int n = c->n_train;
    const long ldn = pad128(n);
    GPN_TRY(ensure_nb(c, NB_KY, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_W, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_T, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_WT, ldn, ldn));
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(std::max(n, c->N)) * 128));
    const double noise = c->theta_exp[c->P - 1];                           // measnoise * theta[-1], quirk A-3
    const int* tr = (const int*)c->train_idx.p;
    GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, tr, n, tr, n, noise, nbp(c, NB_KY), ldn));
    if (keep_copy) {
        GPN_TRY(ensure_nb(c, NB_KC, ldn, ldn));
        GPN_TRY(gpn_k_copy(c, n, n, nbp(c, NB_KY), ldn, nbp(c, NB_KC), ldn));
    }
    return 0;
}
// L = chol(Ky) in place (NB_KY); W = L^-1 (NB_W), W^T (NB_WT)
int gpr_factor_ky(gpn_ctx* c, GprFactor* f) {
    const int n = c->n_train;
    const long ldn = pad128(n);
    GPN_TRY(gpn_potrf(c, n, nbp(c, NB_KY), ldn, (double*)c->winv.p, iscal(c, 1)));
    GPN_TRY(gpn_trtri(c, n, nbp(c, NB_KY), ldn, (const double*)c->winv.p, nbp(c, NB_W), nbp(c, NB_WT), nbp(c, NB_T), ldn));
    f->n = n; f->ldn = ldn; f->Ky = nbp(c, NB_KY); f->W = nbp(c, NB_W); f->T = nbp(c, NB_T); f->WT = nbp(c, NB_WT);
    return 0;
}
int gpr_factor(gpn_ctx* c, bool keep_copy, GprFactor* f) {
    GPN_TRY(gpr_gather_ky(c, keep_copy));
    return gpr_factor_ky(c, f);
}

// vectors: partition of c->tvec (doubles): slot k at offset k*vstride
double* vecp(gpn_ctx* c, int k) {
    const long vs = pad128(std::max(std::max(c->n_train, c->n_test), 1)) + 128;
    return (double*)c->tvec.p + (long)k * vs;
}
int ensure_vecs(gpn_ctx* c) {
    const long vs = pad128(std::max(std::max(c->n_train, c->n_test), 1)) + 128;
    return gpn_buf_ensure(c, c->tvec, sizeof(double) * (size_t)vs * 24);
}
enum VEC { V_TRES = 23, V_T = 0, V_Z, V_ALPHA, V_F, V_A, V_B, V_P, V_W, V_SW, V_G, V_KB, V_RHS, V_Z2, V_SOL, V_Y, V_PHIF, V_U, V_FSTAR, V_VN, V_KSS, V_VV, V_PI };

// Number of CTAs a latency-bound dataflow Cholesky of n rows needs to keep up with its per-column dependency chain
// (~100 us): the busiest tile column has (T/2)^2 k-blocks of ~17 us.  The rest of the GPU is left to the side stream.
int chain_bound_ctas(int n) {
    const int T = (n + 127) / 128;
    static const double scale = getenv("GPNET_B200_CHAIN_SCALE") ? atof(getenv("GPNET_B200_CHAIN_SCALE")) : 0.7;   // measured flat 0.5-1.0
    const int need = (int)(scale * 0.25 * T * T * 17.0 / 90.0) + 8;
    return need < 16 ? 16 : need;
}

// (d K / d theta_0)[tr,tr] pieces that are plain GEMM work and independent of the factorisation of K_y
int gpr_grad_kernel_block(gpn_ctx* c, int kind, long ldn, const double** G, double* gs, bool full = false) {
    const int n = c->n_train, N = c->N;
    const int* tr = (const int*)c->train_idx.p;
    *G = nullptr;
    *gs = 0.0;
    GPN_TRY(ensure_nb(c, NB_G, ldn, ldn));
    if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN) {
        // dK/dtheta0 = K^2 - K  => (K^2)[tr,tr] = K[tr,:] K[tr,:]^T  (SYRK n x n x N)
        GPN_TRY(ensure_nb(c, NB_ROWS1, n, c->ldN));
        GPN_TRY(gpn_k_gather_rows(c, kfull(c), c->ldN, tr, n, N, nbp(c, NB_ROWS1), c->ldN));
        GPN_TRY(gpn_gemm(c, n, n, N, 1.0, nbp(c, NB_ROWS1), c->ldN, true, nbp(c, NB_ROWS1), c->ldN, true, 0.0, nbp(c, NB_G), ldn,
                         GF_LOWER | (full ? GF_MIRROR : 0)));
        *G = nbp(c, NB_G); *gs = 1.0;
    } else if (kind == GPN_KERNEL_DIFFUSION) {
        // dK/dtheta0 = -beta L~ K = A K  => (A K)[tr,tr] = A[tr,:] K[tr,:]^T (symmetric: A and K commute)
        GPN_TRY(ensure_nb(c, NB_ROWS1, n, c->ldN));
        GPN_TRY(ensure_nb(c, NB_ROWS2, n, c->ldN));
        GPN_TRY(gpn_k_gather_rows(c, mat(c, 0), c->ldN, tr, n, N, nbp(c, NB_ROWS1), c->ldN));
        GPN_TRY(gpn_k_gather_rows(c, kfull(c), c->ldN, tr, n, N, nbp(c, NB_ROWS2), c->ldN));
        GPN_TRY(gpn_gemm(c, n, n, N, 1.0, nbp(c, NB_ROWS1), c->ldN, true, nbp(c, NB_ROWS2), c->ldN, true, 0.0, nbp(c, NB_G), ldn,
                         GF_LOWER | (full ? GF_MIRROR : 0)));
        *G = nbp(c, NB_G); *gs = 1.0;
    } else {
        // dK/dtheta0 = a p B^(p-1); p = 1: B^0 = I; p <= 0: K = I, derivative 0
        const int p = c->pstep_p;
        if (p >= 2) {
            GPN_TRY(gpn_k_gather(c, mat(c, 4), c->ldN, tr, n, tr, n, 0.0, nbp(c, NB_G), ldn));
            *G = nbp(c, NB_G); *gs = c->theta_exp[0] * p;
        } else if (p == 1) {
            GPN_TRY(gpn_k_fill_zero(c, nbp(c, NB_G), ldn, n, n));
            const double* X[1] = {nbp(c, NB_G)};
            const long l1[1] = {ldn};
            const double cf[1] = {0.0};
            GPN_TRY(gpn_k_lincomb(c, n, nbp(c, NB_G), ldn, 1.0, 1, X, l1, cf));
            *G = nbp(c, NB_G); *gs = c->theta_exp[0];
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Regularized-Laplacian fast path: never forms the full N x N inverse.
//
// logPosterior / its gradient only need K[tr,tr] and K[tr,:] of K = (I + beta L~)^-1.  With the training nodes numbered
// LAST (a symmetric permutation of the graph, done once per node set), M = L L^T and W = L^-1 give
//     K[tr,:] = W_tt^T [ W_to  W_tt ]          (W_tt = trailing block of W, W_to the block to its left)
// so the N^3/3 flops of the full W^T W shrink to n^3 (K_to) + n^3/3 (K_tt), and without the gradient only the trailing
// block of W is needed at all (n^3/3 instead of N^3/3 for the triangular inverse).  The block boundary is rounded down to a
// multiple of 128 (o2) so that every operand stays aligned with the Cholesky's tiles; the up to 127 extra rows are computed
// and ignored.  Values agree with the full-inverse path to rounding (different elimination order).
// Level-set dissection of the subgraph induced on the other (non-training) nodes: BFS levels from a pseudo-peripheral node; a
// middle level Sep separates the levels below it (A) from the levels above it (B) exactly, because edges only join equal or
// adjacent levels.  With the others numbered [A | B | Sep], M_oo = I + beta L~_oo is block-arrow: M_AA and M_BB are independent
// (two Cholesky chains of half the length, run concurrently) and Sep couples them through a rank-|Sep| correction.
// *a = number of others, *b = *s = 0 when the dissection is not worth it (small blocks, fat separator, dense graphs).
void dissect_others(const gpn_ctx* c, const std::vector<char>& is_train, std::vector<int>& order, int* a, int* b, int* s) {
    const int N = c->N;
    std::vector<int> others;
    for (int v = 0; v < N; ++v)
        if (!is_train[v]) others.push_back(v);
    const int o = (int)others.size();
    order = others;
    *a = o; *b = 0; *s = 0;
    if (o < 1536) return;
    std::vector<int> lev(N), queue;
    auto bfs = [&](int root) {
        std::fill(lev.begin(), lev.end(), -1);
        queue.clear();
        queue.push_back(root);
        lev[root] = 0;
        for (size_t h = 0; h < queue.size(); ++h) {
            const int u = queue[h];
            for (int q = c->h_indptr[u]; q < c->h_indptr[u + 1]; ++q) {
                const int v = c->h_indices[q];
                if (!is_train[v] && lev[v] < 0) { lev[v] = lev[u] + 1; queue.push_back(v); }
            }
        }
        return queue.back();                                   // a node of the last level
    };
    int root = others[0];
    for (int sweep = 0; sweep < 2; ++sweep) root = bfs(root);  // pseudo-peripheral node
    bfs(root);
    const int reached = (int)queue.size();
    int nlev = 0;
    for (int v : queue) nlev = std::max(nlev, lev[v] + 1);
    if (nlev < 5) return;
    std::vector<int> cnt(nlev, 0);
    for (int v : queue) cnt[lev[v]]++;
    int mid = 0, cum = 0;
    while (mid < nlev - 1 && cum + cnt[mid] < reached / 2) cum += cnt[mid++];
    std::vector<int> A, B, S;
    for (int v : others) {
        if (lev[v] < 0 || lev[v] > mid) B.push_back(v);        // other components have no edge to A either
        else if (lev[v] < mid) A.push_back(v);
        else S.push_back(v);
    }
    if (A.size() & 1) { S.push_back(A.back()); A.pop_back(); } // block B starts at an even offset (16-byte aligned sub-matrix)
    const int na = (int)A.size(), nb = (int)B.size(), ns = (int)S.size();
    if (na < 512 || nb < 512 || ns < 1 || ns > std::min(na, nb) / 2) return;
    order.clear();
    order.insert(order.end(), A.begin(), A.end());
    order.insert(order.end(), B.begin(), B.end());
    order.insert(order.end(), S.begin(), S.end());
    *a = na; *b = nb; *s = ns;
}

int ensure_perm(gpn_ctx* c) {
    if (c->perm_valid) return 0;
    const int N = c->N, n = c->n_train;
    std::vector<char> is_train(N, 0);
    for (int v : c->h_train) is_train[v] = 1;
    std::vector<int> perm(N), inv(N);
    // the other nodes first -- as [A | B | Sep] when a level-set dissection of their induced subgraph is worth it --, then the
    // training nodes in their given (ascending) order
    std::vector<int> others_order;
    dissect_others(c, is_train, others_order, &c->dis_a, &c->dis_b, &c->dis_s);
    {
        int a = 0, b = N - n;
        for (int v : others_order) { perm[v] = a; inv[a] = v; ++a; }
        for (int v = 0; v < N; ++v)
            if (is_train[v]) { perm[v] = b; inv[b] = v; ++b; }
    }
    std::vector<int> pptr(N + 1, 0), pidx(c->h_indices.size());
    std::vector<double> pw(c->h_weights.size());
    size_t e = 0;
    for (int r = 0; r < N; ++r) {
        const int v = inv[r];
        for (int q = c->h_indptr[v]; q < c->h_indptr[v + 1]; ++q, ++e) {
            pidx[e] = perm[c->h_indices[q]];
            if (!pw.empty()) pw[e] = c->h_weights[q];
        }
        pptr[r + 1] = (int)e;
    }
    std::vector<int> iota(N);
    for (int v = 0; v < N; ++v) iota[v] = v;
    GPN_TRY(gpn_buf_ensure(c, c->p_indptr, sizeof(int) * (size_t)(N + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->p_indices, sizeof(int) * (pidx.size() + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->p_dinv, sizeof(double) * (size_t)(2 * N)));
    GPN_TRY(gpn_buf_ensure(c, c->iota, sizeof(int) * (size_t)N));
    GPN_TRY(upload(c, c->p_indptr.p, pptr.data(), sizeof(int) * (size_t)(N + 1)));
    if (!pidx.empty()) GPN_TRY(upload(c, c->p_indices.p, pidx.data(), sizeof(int) * pidx.size()));
    GPN_TRY(upload(c, c->iota.p, iota.data(), sizeof(int) * (size_t)N));
    const double* wdev = nullptr;
    if (!pw.empty()) {
        GPN_TRY(gpn_buf_ensure(c, c->p_weights, sizeof(double) * pw.size()));
        GPN_TRY(upload(c, c->p_weights.p, pw.data(), sizeof(double) * pw.size()));
        wdev = (const double*)c->p_weights.p;
    }
    GPN_TRY(gpn_k_degree_dinv(c, N, (const int*)c->p_indptr.p, (const int*)c->p_indices.p, wdev, (double*)c->p_dinv.p));
    // coupling block (training rows) x (other columns) of the permuted graph: what the sparse-coupling route multiplies with
    const int o = N - n;
    std::vector<int> tptr(n + 1, 0), tidx;
    for (int r = 0; r < n; ++r) {
        for (int q = pptr[o + r]; q < pptr[o + r + 1]; ++q)
            if (pidx[q] < o) tidx.push_back(pidx[q]);
        tptr[r + 1] = (int)tidx.size();
    }
    c->to_nnz = (long long)tidx.size();
    GPN_TRY(gpn_buf_ensure(c, c->to_indptr, sizeof(int) * (size_t)(n + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->to_idx, sizeof(int) * (tidx.size() + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->to_val, sizeof(double) * (tidx.size() + 1)));
    GPN_TRY(upload(c, c->to_indptr.p, tptr.data(), sizeof(int) * (size_t)(n + 1)));
    if (!tidx.empty()) GPN_TRY(upload(c, c->to_idx.p, tidx.data(), sizeof(int) * tidx.size()));
    if (c->dis_s > 0) {   // the two coupling blocks of the separator rows
        const int da = c->dis_a, db = c->dis_b, ds = c->dis_s;
        std::vector<int> aptr(ds + 1, 0), aidx, bptr(ds + 1, 0), bidx, bidr;
        for (int r = 0; r < ds; ++r) {
            for (int q = pptr[da + db + r]; q < pptr[da + db + r + 1]; ++q) {
                const int k = pidx[q];
                if (k < da) aidx.push_back(k);
                else if (k < da + db) { bidx.push_back(k); bidr.push_back(k - da); }
            }
            aptr[r + 1] = (int)aidx.size();
            bptr[r + 1] = (int)bidx.size();
        }
        GPN_TRY(gpn_buf_ensure(c, c->sa_ptr, sizeof(int) * (size_t)(ds + 1)));
        GPN_TRY(gpn_buf_ensure(c, c->sb_ptr, sizeof(int) * (size_t)(ds + 1)));
        GPN_TRY(gpn_buf_ensure(c, c->sa_idx, sizeof(int) * (aidx.size() + 1)));
        GPN_TRY(gpn_buf_ensure(c, c->sb_idx, sizeof(int) * (bidx.size() + 1)));
        GPN_TRY(gpn_buf_ensure(c, c->sb_idx_rel, sizeof(int) * (bidx.size() + 1)));
        GPN_TRY(gpn_buf_ensure(c, c->sa_val, sizeof(double) * (aidx.size() + 1)));
        GPN_TRY(gpn_buf_ensure(c, c->sb_val, sizeof(double) * (bidx.size() + 1)));
        GPN_TRY(upload(c, c->sa_ptr.p, aptr.data(), sizeof(int) * (size_t)(ds + 1)));
        GPN_TRY(upload(c, c->sb_ptr.p, bptr.data(), sizeof(int) * (size_t)(ds + 1)));
        if (!aidx.empty()) GPN_TRY(upload(c, c->sa_idx.p, aidx.data(), sizeof(int) * aidx.size()));
        if (!bidx.empty()) {
            GPN_TRY(upload(c, c->sb_idx.p, bidx.data(), sizeof(int) * bidx.size()));
            GPN_TRY(upload(c, c->sb_idx_rel.p, bidr.data(), sizeof(int) * bidr.size()));
        }
    }
    c->perm_valid = true;
    return 0;
}

int gpr_logp_grad_reglap_fast(gpn_ctx* c, const double* theta_h, int P, const double* t_h, int want_grad, double* out_h, int* info_h) {
    const int N = c->N, n = c->n_train;
    GPN_TRY(ensure_perm(c));
    c->last_route = 1;
    GPN_TRY(ensure_vecs(c));
    c->have_full = false;                       // the resident full kernel is NOT formed on this path
    c->kind = GPN_KERNEL_REGULARIZED_LAPLACIAN;
    c->P = P;
    for (int i = 0; i < P; ++i) c->theta_exp[i] = theta_h[i];
    c->ldN = pad128(N);
    GPN_TRY(clear_info(c));
    const long ld = c->ldN, ldn = pad128(n);
    const double beta = c->theta_exp[0], noise = c->theta_exp[P - 1];
    const int o = N - n, o2 = (o / 128) * 128, d = o - o2, n2 = N - o2;
    for (int i = 0; i < 4; ++i) GPN_TRY(ensure_mat(c, i));
    double *M = mat(c, 0), *W = mat(c, 1), *TT = mat(c, 2), *WT = mat(c, 3);
    if (t_h) GPN_TRY(upload(c, vecp(c, V_T), t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(vecp(c, V_T), vecp(c, V_TRES), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    // M = I + beta L~ on the permuted graph; L = chol(M)
    GPN_TRY(gpn_k_fill_zero(c, M, ld, (int)ld, (int)ld));
    GPN_TRY(gpn_k_laplacian_scatter(c, N, (const int*)c->p_indptr.p, (const int*)c->p_indices.p,
                                    c->h_weights.empty() ? nullptr : (const double*)c->p_weights.p, (const double*)c->p_dinv.p, beta, 1.0, M, ld));
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(N) * 128));
    GPN_TRY(gpn_potrf(c, N, M, ld, (double*)c->winv.p, iscal(c, 0)));
    const long sub = (long)o2 * ld + o2;
    GPN_TRY(ensure_nb(c, NB_ROWS1, n2, ld));
    double* R = nbp(c, NB_ROWS1);                 // rows o2 .. N-1 of K (n2 x N)
    if (want_grad) {
        GPN_TRY(gpn_trtri(c, N, M, ld, (const double*)c->winv.p, W, WT, TT, ld));
        // K[o2:, :o2] = W_tt^T W_to : A[i][k] = W_tt[k][i] = WT[o2+i][o2+k] (k >= i), B[j][k] = W_to[k][j] = WT[j][o2+k]
        if (o2 > 0) GPN_TRY(gpn_gemm(c, n2, o2, n2, 1.0, WT + sub, ld, true, WT + o2, ld, true, 0.0, R, ld, GF_KLO_M));
    } else {
        // only the trailing block of W is needed
        GPN_TRY(gpn_trtri(c, n2, M + sub, ld, (const double*)c->winv.p + (long)(o2 / 128) * 128 * 128, W + sub, WT + sub, TT + sub, ld));
    }
    // K[o2:, o2:] = W_tt^T W_tt (full symmetric block)
    GPN_TRY(gpn_gemm(c, n2, n2, n2, 1.0, WT + sub, ld, true, WT + sub, ld, true, 0.0, R + o2, ld, GF_LOWER | GF_MIRROR | GF_KLO_M | GF_KLO_N));
    // K_y = K[tr,tr] + c
    GPN_TRY(ensure_nb(c, NB_KY, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_W, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_T, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_WT, ldn, ldn));
    const int* iota = (const int*)c->iota.p;
    GPN_TRY(gpn_k_gather(c, R + o2, ld, iota + d, n, iota + d, n, noise, nbp(c, NB_KY), ldn));
    const double* Rtr = R + (long)d * ld;         // the n training rows of K (n x N, K-major)
    GprFactor f;
    const double* G = nullptr;
    bool forked = false;
    int st;
    if (want_grad) {
        cudaStream_t main_stream = c->stream;
        GPN_CK(cudaEventRecord(c->ev_fork, main_stream));
        c->potrf_grid_cap = chain_bound_ctas(n);
        st = gpr_factor_ky(c, &f);
        c->potrf_grid_cap = 0;
        if (st != 0) return st;
        GPN_CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        c->stream = c->stream2;
        st = ensure_nb(c, NB_G, ldn, ldn);       // (K^2)[tr,tr] = K[tr,:] K[tr,:]^T
        if (st == 0) st = gpn_gemm(c, n, n, N, 1.0, Rtr, ld, true, Rtr, ld, true, 0.0, nbp(c, NB_G), ldn, GF_LOWER);
        c->stream = main_stream;
        if (st != 0) return st;
        GPN_CK(cudaEventRecord(c->ev_join, c->stream2));
        G = nbp(c, NB_G);
        forked = true;
    } else {
        GPN_TRY(gpr_factor_ky(c, &f));
    }
    GPN_TRY(gpn_k_gemv(c, n, n, f.W, f.ldn, vecp(c, V_T), vecp(c, V_Z), GEMV_LOWER));
    GPN_TRY(gpn_k_dot(c, n, vecp(c, V_Z), vecp(c, V_Z), dscal(c, Scal::DOT)));
    GPN_TRY(gpn_k_sumlogdiag(c, n, f.Ky, f.ldn, dscal(c, Scal::LOGDET)));
    if (want_grad) {
        GPN_TRY(gpn_k_gemv_t(c, n, n, f.W, f.ldn, vecp(c, V_Z), vecp(c, V_ALPHA), GEMV_LOWER));
        GPN_TRY(gpn_lauum_t(c, n, f.WT, f.ldn, f.T, f.ldn, false));
        if (forked) GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        // D0 = (K^2 - K)[tr,tr]: K[tr,tr] is read in place from the training rows (no noise in there)
        GPN_TRY(gpn_k_grad_trace(c, n, f.T, f.ldn, G, f.ldn, Rtr + o2 + d, ld, 1.0, -1.0, 0.0, vecp(c, V_ALPHA), dscal(c, Scal::TR0)));
    }
    double sc[Scal::NDBL];
    int info[16];
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    if (info_h) *info_h = info[1] ? info[1] : info[0];
    out_h[0] = 0.5 * sc[Scal::DOT] + sc[Scal::LOGDET] + 0.5 * n * LOG_2PI;
    if (want_grad) {
        for (int j = 0; j < P; ++j) out_h[1 + j] = 0.0;
        const double sinv = sc[Scal::TR0], tr0 = sc[Scal::TR1], sa = sc[Scal::TR2];
        out_h[1 + 0] += -0.5 * tr0;
        out_h[1 + (P - 1)] += -0.5 * noise * (sa * sa - sinv);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Regularized-Laplacian Schur path: no kernel block is formed at all.
//
// With the training nodes numbered last and M = I + beta L~ = L L^T, the trailing block of the factor IS the factor of the
// inverse training block: K_tt^-1 = M_tt - M_to M_oo^-1 M_ot = L_tt L_tt^T.  K_y = K_tt + c 1 1^T (noise on every entry,
// quirk A-3) is a rank-one update of it, so with r_1 = L_tt^T 1, r_t = L_tt^T t, s = r_1.r_1, a = r_1.r_t:
//     K_y^-1 = S - c (S1)(S1)^T / (1 + c s),  S = L_tt L_tt^T;     t^T K_y^-1 t = r_t.r_t - c a^2 / (1 + c s)
//     (1/2) log det K_y = -sum log diag(L_tt) + (1/2) log1p(c s)
// i.e. -logp costs ONE Cholesky of M plus two triangular GEMVs.  The gradient needs the bottom block-row Wb = (L^-1)[tr,:]:
//     D = dK_y/dtheta_0 = (K^2 - K)[tr,tr],  K[:,tr] x = Wb^T (W_tt x),  K_tt = W_tt^T W_tt,  W_tt L_tt = I
//     x^T D x' = z_x.z_x' - y_x.y_x'   with  y_x = W_tt x, z_x = Wb^T y_x;   for x = S t: y = r_t;  for x = S 1: y = r_1
//     tr(S D) = ||Wb||_F^2 - n,   tr(K_y^-1 D) = tr(S D) - c/(1 + c s) (S1)^T D (S1)
// so after the triangular inverse one HBM pass over Wb (two transposed GEMVs + a Frobenius norm) finishes the job:
// N^3/3 + N^3/3 flops instead of N^3 + n^2 N + n^3.  All scalar algebra happens on the host from nine reductions.

// Noise-independent reductions of one kernel: everything -logp and its gradient need for ANY noise level c, because
// K_y = K_tt + c 1 1^T is a rank-one update of K_tt (S = K_tt^-1, D = dK_tt/dtheta_0).
struct RowScalars {
    int n = 0;
    double half_logdet = 0.0;                    // (1/2) log det K_tt
    double s = 0.0, a = 0.0, q = 0.0;            // 1^T S 1, 1^T S t, t^T S t
    double trSD = 0.0, D11 = 0.0, D1t = 0.0, Dtt = 0.0;   // tr(S D), (S1)^T D (S1), (S1)^T D (St), (St)^T D (St)
};
// out[0] = -logp, out[1..P] = -dlogp/dtheta (GPnetRegressor.py:136-150 and the gradient of old_implementation/GPR.py:52-64)
void row_eval(const RowScalars& r, double noise, int P, int want_grad, double* out) {
    const double den = 1.0 + noise * r.s, mu = noise * r.a / den;
    out[0] = 0.5 * (r.q - mu * r.a) + (r.half_logdet + 0.5 * log1p(noise * r.s)) + 0.5 * r.n * LOG_2PI;
    if (!want_grad) return;
    for (int j = 0; j < P; ++j) out[1 + j] = 0.0;
    const double aDa = r.Dtt - 2.0 * mu * r.D1t + mu * mu * r.D11;          // alpha = S t - mu S 1
    const double trKD = r.trSD - noise / den * r.D11;                       // K_y^-1 = S - c (S1)(S1)^T / (1 + c s)
    out[1 + 0] += -(0.5 * aDa - 0.5 * trKD);
    const double sa = r.a / den, sinv = r.s / den;                           // 1^T alpha, 1^T K_y^-1 1
    out[1 + (P - 1)] += -0.5 * noise * (sa * sa - sinv);                     // D_last = c 1 1^T
}

int gpr_reglap_schur_scalars(gpn_ctx* c, const double* theta_h, int P, const double* t_h, int want_grad, RowScalars* rs, int* info_h) {
    const int N = c->N, n = c->n_train;
    GPN_TRY(ensure_perm(c));
    c->last_route = 2;
    GPN_TRY(ensure_vecs(c));
    c->have_full = false;
    c->kind = GPN_KERNEL_REGULARIZED_LAPLACIAN;
    c->P = P;
    for (int i = 0; i < P; ++i) c->theta_exp[i] = theta_h[i];
    c->ldN = pad128(N);
    GPN_TRY(clear_info(c));
    const long ld = c->ldN;
    const double beta = c->theta_exp[0];
    const int o = N - n;
    for (int i = 0; i < (want_grad ? 4 : 1); ++i) GPN_TRY(ensure_mat(c, i));
    double* M = mat(c, 0);
    if (t_h) GPN_TRY(upload(c, vecp(c, V_T), t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(vecp(c, V_T), vecp(c, V_TRES), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    GPN_TRY(gpn_k_fill_zero_lower(c, M, ld, (int)ld));       // nothing on this route reads above the diagonal tiles
    GPN_TRY(gpn_k_laplacian_scatter(c, N, (const int*)c->p_indptr.p, (const int*)c->p_indices.p,
                                    c->h_weights.empty() ? nullptr : (const double*)c->p_weights.p, (const double*)c->p_dinv.p, beta, 1.0, M, ld));
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(N) * 128));
    GPN_TRY(gpn_potrf(c, N, M, ld, (double*)c->winv.p, iscal(c, 0)));
    const double* Ltt = M + (long)o * ld + o;
    double *ones = vecp(c, V_U), *r1 = vecp(c, V_Z), *rt = vecp(c, V_Z2);
    GPN_TRY(gpn_k_fill_vec(c, n, ones, 1.0));
    GPN_TRY(gpn_k_gemv_t2(c, n, n, Ltt, ld, ones, vecp(c, V_T), r1, rt, GEMV_LOWER));
    GPN_TRY(gpn_k_sumlogdiag(c, n, Ltt, ld, dscal(c, Scal::LOGDET)));
    const double *dx[6] = {r1, r1, rt}, *dy[6] = {r1, rt, rt};
    int dn[6] = {n, n, n, N, N, N};
    if (want_grad) {
        double *W = mat(c, 1), *TT = mat(c, 2), *WT = mat(c, 3);
        GPN_TRY(gpn_trtri(c, N, M, ld, (const double*)c->winv.p, W, WT, TT, ld));
        GPN_TRY(ensure_nb(c, NB_S, 2, ld));
        double *z1 = nbp(c, NB_S), *zt = z1 + ld;
        GPN_TRY(gpn_k_trap_pass(c, n, N, o, W + (long)o * ld, ld, r1, rt, z1, zt, dscal(c, Scal::FRO)));
        dx[3] = z1; dy[3] = z1; dx[4] = z1; dy[4] = zt; dx[5] = zt; dy[5] = zt;
    }
    GPN_TRY(gpn_k_multi_dot(c, want_grad ? 6 : 3, dx, dy, dn, dscal(c, Scal::SCH0)));
    double sc[Scal::NDBL];
    int info[16];
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    if (info_h) *info_h = info[0];
    rs->n = n;
    rs->half_logdet = -sc[Scal::LOGDET];
    rs->s = sc[Scal::SCH0]; rs->a = sc[Scal::SCH0 + 1]; rs->q = sc[Scal::SCH0 + 2];
    if (want_grad) {
        rs->D11 = sc[Scal::SCH0 + 3] - rs->s; rs->D1t = sc[Scal::SCH0 + 4] - rs->a; rs->Dtt = sc[Scal::SCH0 + 5] - rs->q;
        rs->trSD = sc[Scal::FRO] - n;
    }
    return 0;
}
// Z = M_oo^-1 through the dissection of the other nodes (ensure_perm numbered them [A | B | Sep], no edge joins A and B):
//   M_oo = [D C; C^T E],  D = blkdiag(M_AA, M_BB),  C = [M_AS; M_BS] (sparse),  E = M_SS
//   Z_A = M_AA^-1 and Z_B = M_BB^-1: two INDEPENDENT augmented factorisations [M; I] (half the tile columns each: half the
//   dependency chain) on the two streams, each on half of the SMs;  U = D^-1 C = [Z_A M_AS; Z_B M_BS] (sparse x dense),
//   Sigma = (E - C^T U)^-1 (|Sep| x |Sep|),  V = U Sigma,
//   Z = [D^-1 + V U^T, -V; -V^T, Sigma]                                   (a rank-|Sep| correction of blkdiag(Z_A, Z_B)).
// a^3 + b^3 flops instead of o^3, and the chain of 32 tile columns at C3 becomes two concurrent chains of 16.
// In: M holds I + beta L~ (lower tiles) with its (B, A) block zero.  Out: the full symmetric Z in its leading o x o block.
int reglap_inverse_oo_dissected(gpn_ctx* c, double* M, long ld, int* info_slot_used) {
    const int a = c->dis_a, b = c->dis_b, s = c->dis_s, D = a + b;
    const long lds = pad128(s);
    const int Ta = (a + 127) / 128, Tb = (b + 127) / 128;
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)(Ta + Tb + 2) * 128 * 128));
    double* WT = mat(c, 2);
    double* U = mat(c, 1);                                  // (a+b) x s, then V, S_sep, Sigma and the scratch of its inverse
    double* V = U + (size_t)pad128(D) * lds;
    double* Ssep = V + (size_t)pad128(D) * lds;
    double* Sg = Ssep + lds * lds;
    double *w1 = Sg + lds * lds, *w2 = w1 + lds * lds, *w3 = w2 + lds * lds;
    if ((size_t)(2 * pad128(D) * lds + 6 * lds * lds) > (size_t)c->ldN * c->ldN) return -9;
    // separator coupling values of this theta (the entries the scatter wrote)
    GPN_TRY(gpn_k_csr_values(c, s, D, (const int*)c->sa_ptr.p, (const int*)c->sa_idx.p, M, ld, (double*)c->sa_val.p));
    GPN_TRY(gpn_k_csr_values(c, s, D, (const int*)c->sb_ptr.p, (const int*)c->sb_idx.p, M, ld, (double*)c->sb_val.p));
    // the two halves concurrently
    cudaStream_t main_stream = c->stream;
    double* MB = M + (long)a * ld + a;
    GPN_CK(cudaEventRecord(c->ev_fork, main_stream));
    GPN_CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    c->potrf_grid_cap = c->num_sms / 2;
    int st = gpn_potrf_dataflow(c, a, M, ld, (double*)c->winv.p, iscal(c, 0), 0, nullptr, 0, 0, WT, ld, 0);
    if (st == 0) {
        c->stream = c->stream2;
        st = gpn_potrf_dataflow(c, b, MB, ld, (double*)c->winv.p + (size_t)Ta * 128 * 128, iscal(c, 2), 0, nullptr, 0, 0, WT + (long)a * ld + a, ld, 1);
        if (st == 0) st = gpn_lauum_t(c, b, WT + (long)a * ld + a, ld, MB, ld, true);
        c->stream = main_stream;
    }
    c->potrf_grid_cap = 0;
    if (st != 0) return st;
    GPN_CK(cudaEventRecord(c->ev_join, c->stream2));
    GPN_TRY(gpn_lauum_t(c, a, WT, ld, M, ld, true));
    GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    // U = [Z_A M_AS; Z_B M_BS] = ([M_SA Z_A, M_SB Z_B])^T, written transposed by the SpMM
    GPN_TRY(gpn_k_spmm_rows_t(c, s, a, (const int*)c->sa_ptr.p, (const int*)c->sa_idx.p, (const double*)c->sa_val.p, M, ld, U, lds));
    GPN_TRY(gpn_k_spmm_rows_t(c, s, b, (const int*)c->sb_ptr.p, (const int*)c->sb_idx_rel.p, (const double*)c->sb_val.p, MB, ld, U + (long)a * lds, lds));
    // S_sep = M_SS - M_SA U_A - M_SB U_B  (lower triangle), Sigma = S_sep^-1
    GPN_TRY(gpn_k_spmm_rows(c, s, s, (const int*)c->sa_ptr.p, (const int*)c->sa_idx.p, (const double*)c->sa_val.p, U, lds, Ssep, lds,
                            M + (long)D * ld + D, ld));
    GPN_TRY(gpn_k_spmm_rows(c, s, s, (const int*)c->sb_ptr.p, (const int*)c->sb_idx_rel.p, (const double*)c->sb_val.p, U + (long)a * lds, lds, Ssep, lds,
                            Ssep, lds));
    {   // spd_inverse works with one leading dimension for everything
        GPN_TRY(gpn_potrf(c, s, Ssep, lds, (double*)c->winv.p, iscal(c, 3)));
        GPN_TRY(gpn_trtri(c, s, Ssep, lds, (const double*)c->winv.p, w1, w2, w3, lds));
        GPN_TRY(gpn_lauum_t(c, s, w2, lds, Sg, lds, true));
    }
    // V = U Sigma;  Z_DD = blkdiag(Z_A, Z_B) + V U^T;  borders
    GPN_TRY(gpn_gemm(c, D, s, s, 1.0, U, lds, true, Sg, lds, true, 0.0, V, lds, 0));
    GPN_TRY(gpn_gemm(c, D, D, s, 1.0, V, lds, true, U, lds, true, 1.0, M, ld, GF_LOWER | GF_MIRROR));
    GPN_TRY(gpn_k_z_border(c, D, s, V, lds, Sg, lds, M, ld));
    *info_slot_used = 3;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Level 6: whole-graph K-way dissection (dissect.cu has the derivation).  No dense Schur complement: log det S, the quadratic
// forms and the gradient traces come from log det / tr(inverse) / two solves with M and M_oo, which K concurrent augmented
// factorisations [M_kk; I] (one per part) and the small separator complements T, T_o deliver.
// ---------------------------------------------------------------------------------------------------------------------
int d6_second_stream(gpn_ctx* c, int k) {
    if (!c->side2[k]) {
        GPN_CK(cudaStreamCreateWithFlags(&c->side2[k], cudaStreamNonBlocking));
        GPN_CK(cudaEventCreateWithFlags(&c->ev_g[k], cudaEventDisableTiming));
        GPN_CK(cudaEventCreateWithFlags(&c->ev_u[k], cudaEventDisableTiming));
    }
    return 0;
}
int d6_side_stream(gpn_ctx* c, int k) {
    if (!c->side[k]) {
        GPN_CK(cudaStreamCreateWithFlags(&c->side[k], cudaStreamNonBlocking));
        GPN_CK(cudaEventCreateWithFlags(&c->ev_side[k], cudaEventDisableTiming));
    }
    return 0;
}

int gpr_reglap_d6_enqueue(gpn_ctx* c, const double* theta_h, int P, const double* t_h, int want_grad) {
    Dissect6& q = c->d6;
    const int N = c->N, n = c->n_train, K = q.K, s = q.s, so = q.so, ND = q.ND;
    GPN_TRY(ensure_vecs(c));
    c->have_full = false;
    c->kind = GPN_KERNEL_REGULARIZED_LAPLACIAN;
    c->P = P;
    for (int i = 0; i < P; ++i) c->theta_exp[i] = theta_h[i];
    c->ldN = pad128(N);
    c->last_route = 6;
    GPN_TRY(clear_info(c));
    const long ld = c->ldN, ldg = pad128(ND), lds = pad128(s), ldso = pad128(std::max(so, 1)), rowsU = pad128(ND);
    const int nb = K;                                           // partial products G_k^T G_k, one per part
    const double beta = c->theta_exp[0];
    GPN_TRY(ensure_mat(c, 0));
    GPN_TRY(ensure_mat(c, 2));
    double *M = mat(c, 0), *WT = mat(c, 2);
    if (t_h) GPN_TRY(upload(c, vecp(c, V_T), t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(vecp(c, V_T), vecp(c, V_TRES), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    GPN_TRY(gpn_buf_ensure(c, q.work[0], sizeof(double) * (size_t)lds * ldg));             // G^T   (s x ND; only columns >= c0 of each part are used)
    GPN_TRY(gpn_buf_ensure(c, q.work[1], sizeof(double) * (size_t)ldso * ldg));            // G_o^T (so x ND, training columns zero)
    GPN_TRY(gpn_buf_ensure(c, q.work[4], sizeof(double) * (size_t)lds * lds * 2));         // T, W_T^T
    GPN_TRY(gpn_buf_ensure(c, q.work[5], sizeof(double) * (size_t)ldso * ldso * 2));       // T_o, W_To^T
    GPN_TRY(gpn_buf_ensure(c, q.work[6], sizeof(double) * (size_t)ld * 16));               // vectors of length N
    GPN_TRY(gpn_buf_ensure(c, q.work[7], sizeof(double) * (size_t)nb * lds * lds));        // split-K partial products of G^T G
    GPN_TRY(gpn_buf_ensure(c, q.work[8], sizeof(double) * (size_t)nb * ldso * ldso));
    double *GT = (double*)q.work[0].p, *GTo = (double*)q.work[1].p;
    double *Tb = (double*)q.work[4].p, *WTt = Tb + lds * lds;
    double *Tob = (double*)q.work[5].p, *WTto = Tob + ldso * ldso;
    double *U = nullptr, *Uo = nullptr;
    if (want_grad) {   // U = D^-1 C (ND x s) and its other-node counterpart (rows of training nodes stay zero)
        const size_t ub = sizeof(double) * (size_t)rowsU * lds, uob = sizeof(double) * (size_t)rowsU * ldso;
        const bool fresh = q.work[2].cap < ub;
        GPN_TRY(gpn_buf_ensure(c, q.work[2], ub));
        GPN_TRY(gpn_buf_ensure(c, q.work[3], uob));
        U = (double*)q.work[2].p;
        Uo = (double*)q.work[3].p;
        if (fresh) GPN_CK(cudaMemsetAsync(U, 0, q.work[2].cap, c->stream));                // rows [ND, rowsU) must read as zero
        GPN_TRY(gpn_k_fill_zero(c, Uo, ldso, (int)rowsU, (int)ldso));
    }
    auto vec = [&](int i) { return (double*)q.work[6].p + (long)i * ld; };
    double *b1 = vec(0), *bt = vec(1), *v1 = vec(2), *vt = vec(3), *y1 = vec(4), *yt = vec(5), *w1 = vec(6), *wt = vec(7), *u1 = vec(8),
           *ut = vec(9), *r1 = vec(10), *rt = vec(11), *us1 = vec(12), *ust = vec(13);
    // development aid (GPNET_B200_D6_TIMING=1): CUDA-event time of every phase on the main stream, printed to stderr
    static const bool timing = getenv("GPNET_B200_D6_TIMING") != nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* name) {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, c->stream);
        marks.push_back({name, e});
    };
    mark("start");
    // scalar block of this route: [0..4] coupling sums, [5..8] factor reductions, [9,10] tr / log det of T, [11,12] of T_o, [13..18] dots
    GPN_CK(cudaMemsetAsync(dscal(c, 0), 0, sizeof(double) * Scal::NDBL, c->stream));
    // ---- assembly: M on the permuted graph (only the diagonal blocks and the separator block are read as dense) ----
    GPN_TRY(gpn_d6_fill_zero(c, M, ld));
    GPN_TRY(gpn_k_laplacian_scatter(c, N, (const int*)q.indptr.p, (const int*)q.indices.p, c->h_weights.empty() ? nullptr : (const double*)q.weights.p,
                                    (const double*)q.dinv.p, beta, 1.0, M, ld));
    GPN_TRY(gpn_d6_coupling(c, vecp(c, V_T), beta, b1, bt, dscal(c, 0)));
    mark("assembly");
    // (all products below use the 128x128 GEMM configuration even when they have few tiles: the parts run concurrently, so the chip is
    //  filled across streams, and the big tiles are the efficient ones: 513.7 -> 555.4 evals/s with two lanes)
    // ---- one stream per part: [M_kk; I] -> L_k, W_k^T;  G_k^T = (W_k C_k)^T;  U_k = W_k^T G_k (gradient only) ----
    int Tk[D6_MAXP], woff[D6_MAXP + 1] = {0};
    for (int k = 0; k < K; ++k) { Tk[k] = (q.nk[k] + 127) / 128; woff[k + 1] = woff[k] + Tk[k]; }
    const int Ts = (s + 127) / 128, Tso = (std::max(so, 1) + 127) / 128;
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)(woff[K] + Ts + Tso + 2) * 128 * 128));
    double* winv = (double*)c->winv.p;
    cudaStream_t main_stream = c->stream;
    const int sms = c->sm_budget > 0 ? std::min(c->sm_budget, c->num_sms) : c->num_sms;    // co-resident CTAs this context may use
    {
        double wsum = 0.0;
        for (int k = 0; k < K; ++k) wsum += (double)q.nk[k] * q.nk[k] * q.nk[k];
        int caps[D6_MAXP], used = 0;
        const int cap_min = std::max(4, std::min(12, sms / (2 * K)));
        for (int k = 0; k < K; ++k) { caps[k] = std::max(cap_min, (int)(sms * ((double)q.nk[k] * q.nk[k] * q.nk[k] / wsum))); used += caps[k]; }
        for (int k = 0; used > sms; k = (k + 1) % K)
            if (caps[k] > cap_min) { --caps[k]; --used; }
        GPN_CK(cudaEventRecord(c->ev_fork, main_stream));
        int st = 0;
        bool forked_u[D6_MAXP] = {false};
        for (int k = 0; k < K && st == 0; ++k) {
            if (k > 0) {
                st = d6_side_stream(c, k);
                if (st != 0) break;
                GPN_CK(cudaStreamWaitEvent(c->side[k], c->ev_fork, 0));
                c->stream = c->side[k];
            }
            c->potrf_grid_cap = caps[k];
            double *Mk = M + (long)q.off[k] * ld + q.off[k], *WTk = WT + (long)q.off[k] * ld + q.off[k];
            st = gpn_potrf_dataflow(c, q.nk[k], Mk, ld, winv + (size_t)woff[k] * 128 * 128, iscal(c, k), 0, nullptr, 0, 0, WTk, ld, k);
            c->potrf_grid_cap = 0;
            if (st == 0) st = gpn_d6_sep_spmm(c, k, beta, WT, ld, GT, GTo, ldg);
            // G_k is zero above row c0 of the part (interior nodes come first): every product below starts there
            const int c0 = q.c0[k], kr = q.nk[k] - c0, c0o = std::min(c0, q.ok[k]), kro = q.ok[k] - c0o;
            const double* GTk = GT + q.off[k] + c0;
            bool u_forked = false;
            static const bool second = !(getenv("GPNET_B200_D6_SECOND") && atoi(getenv("GPNET_B200_D6_SECOND")) == 0);
            if (st == 0 && want_grad && kr > 0 && !second) {      // A/B switch: the U products behind the G^T G products on the part's stream
                double* Uk = U + (long)q.off[k] * lds;
                st = gpn_gemm(c, q.nk[k], s, kr, 1.0, WTk + c0, ld, true, GTk, ldg, true, 0.0, Uk, lds, GF_KLO_M | GF_FORCE_BIG, 1, 0, 0, 0, nullptr, 0, 0,
                              c0);
                if (st == 0 && so > 0 && kro > 0)
                    st = gpn_gemm(c, q.ok[k], so, kro, 1.0, WTk + c0o, ld, true, GTo + q.off[k] + c0o, ldg, true, 0.0, Uo + (long)q.off[k] * ldso, ldso,
                                  GF_KLO_M | GF_FORCE_BIG, 1, 0, 0, 0, nullptr, 0, 0, c0o);
            }
            if (st == 0 && want_grad && kr > 0 && second) {
                // U_k = W_k^T G_k (and its other-node counterpart) on the part's SECOND stream, beside the G_k^T G_k products below:
                // both only need G_k, and a launch of 20-70 tiles of ~115 us leaves most of the chip to the other one
                cudaStream_t part_stream = c->stream;
                st = d6_second_stream(c, k);
                if (st == 0) {
                    cudaError_t e = cudaEventRecord(c->ev_g[k], part_stream);
                    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->side2[k], c->ev_g[k], 0);
                    if (e != cudaSuccess) st = GPN_CUDA_ERR(e);
                }
                if (st == 0) {
                    c->stream = c->side2[k];
                    // U_k[m][r] = sum_{j >= max(m, c0)} WT_k[m][j] G^T[r][j]: ONE launch, the operand W_k^T[:, c0:] is trapezoidal (c0 is a
                    // multiple of the tile size: GF_KLO_M shifted by c0)
                    double* Uk = U + (long)q.off[k] * lds;
                    st = gpn_gemm(c, q.nk[k], s, kr, 1.0, WTk + c0, ld, true, GTk, ldg, true, 0.0, Uk, lds, GF_KLO_M | GF_FORCE_BIG, 1, 0, 0, 0, nullptr, 0,
                                  0, c0);
                    if (st == 0 && so > 0 && kro > 0) {
                        double* Uok = Uo + (long)q.off[k] * ldso;
                        const double* GTok = GTo + q.off[k] + c0o;
                        st = gpn_gemm(c, q.ok[k], so, kro, 1.0, WTk + c0o, ld, true, GTok, ldg, true, 0.0, Uok, ldso, GF_KLO_M | GF_FORCE_BIG, 1, 0, 0, 0,
                                      nullptr, 0, 0, c0o);
                    }
                    if (st == 0) {
                        cudaError_t e = cudaEventRecord(c->ev_u[k], c->side2[k]);
                        if (e != cudaSuccess) st = GPN_CUDA_ERR(e);
                        else u_forked = true;
                    }
                    c->stream = part_stream;
                }
            }
            forked_u[k] = u_forked;
            double *Cb = (double*)q.work[7].p + (size_t)k * lds * lds, *Cbo = (double*)q.work[8].p + (size_t)k * ldso * ldso;
            if (st == 0) {
                if (kr > 0) st = gpn_gemm(c, s, s, kr, 1.0, GTk, ldg, true, GTk, ldg, true, 0.0, Cb, lds, GF_LOWER | GF_FORCE_BIG);
                else st = gpn_k_fill_zero(c, Cb, lds, s, (int)lds);
            }
            if (st == 0 && so > 0) {
                if (kro > 0) st = gpn_gemm(c, so, so, kro, 1.0, GTo + q.off[k] + c0o, ldg, true, GTo + q.off[k] + c0o, ldg, true, 0.0, Cbo, ldso, GF_LOWER | GF_FORCE_BIG);
                else st = gpn_k_fill_zero(c, Cbo, ldso, so, (int)ldso);
            }
            if (k > 0 && st == 0) {
                cudaError_t e = cudaEventRecord(c->ev_side[k], c->side[k]);
                if (e != cudaSuccess) st = GPN_CUDA_ERR(e);
            }
            c->stream = main_stream;
        }
        c->potrf_grid_cap = 0;
        if (st != 0) return st;
        for (int k = 1; k < K; ++k) GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_side[k], 0));
        for (int k = 0; k < K; ++k)
            if (forked_u[k]) GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_u[k], 0));
    }
    mark("parts: factor, G, U");
    GPN_TRY(gpn_d6_factor_reduce(c, M, WT, ld, dscal(c, 5)));
    // ---- separator complements T = E - sum_k G_k^T G_k (partial products from the part streams) and T_o; their augmented
    //      factorisations [T; I; U] give L_T, W_T^T and Y = U L_T^-T:  tr(T^-1 (I + U^T U)) = ||W_T||_F^2 + ||Y||_F^2 ----
    const double* E = M + (long)ND * ld + ND;
    double* winvT = winv + (size_t)woff[K] * 128 * 128;
    constexpr int TS = D6_SLOTS - 1;                            // stream / flag slot of the other-node separator complement
    if (so > 0) {   // the other-node complement on a side stream, next to the full one
        GPN_TRY(d6_side_stream(c, TS));
        GPN_CK(cudaEventRecord(c->ev_fork, main_stream));
        GPN_CK(cudaStreamWaitEvent(c->side[TS], c->ev_fork, 0));
        c->stream = c->side[TS];
        int st = gpn_d6_t_reduce(c, so, E, ld, (const double*)q.work[8].p, ldso, nb, Tob);
        c->potrf_grid_cap = std::max(2, sms / 3);
        if (st == 0) st = gpn_potrf_dataflow(c, so, Tob, ldso, winvT + (size_t)Ts * 128 * 128, iscal(c, 15), 0, want_grad ? Uo : nullptr, ldso,
                                             want_grad ? ND : 0, WTto, ldso, TS);
        c->potrf_grid_cap = 0;
        if (st == 0) st = gpn_d6_sep_reduce(c, so, Tob, ldso, want_grad ? WTto : nullptr, ldso, want_grad ? Uo : nullptr, ldso, ND, dscal(c, 11));
        // u_x = M_oo^-1 (M_ot x) for x = 1, t: block-arrow solve through the factors
        if (st == 0) st = gpn_d6_block_solve(c, WT, ld, b1, bt, v1, vt, y1, yt);
        if (st == 0) st = gpn_d6_sep_gather(c, beta, b1, bt, y1, yt, r1, rt);
        if (st == 0) st = gpn_d6_tri_solve(c, so, WTto, ldso, r1, rt, v1, vt, us1, ust);
        c->stream = main_stream;
        if (st != 0) return st;
        GPN_CK(cudaMemcpyAsync(w1, b1, sizeof(double) * ND, cudaMemcpyDeviceToDevice, c->side[TS]));
        GPN_CK(cudaMemcpyAsync(wt, bt, sizeof(double) * ND, cudaMemcpyDeviceToDevice, c->side[TS]));
        GPN_CK(cudaMemsetAsync(u1, 0, sizeof(double) * 2 * ld, c->side[TS]));               // u1, ut are adjacent
        c->stream = c->side[TS];
        st = gpn_d6_sep_scatter(c, beta, us1, ust, w1, wt);
        if (st == 0) st = gpn_d6_block_solve(c, WT, ld, w1, wt, v1, vt, u1, ut);
        c->stream = main_stream;
        if (st != 0) return st;
        GPN_CK(cudaMemcpyAsync(u1 + ND, us1, sizeof(double) * so, cudaMemcpyDeviceToDevice, c->side[TS]));
        GPN_CK(cudaMemcpyAsync(ut + ND, ust, sizeof(double) * so, cudaMemcpyDeviceToDevice, c->side[TS]));
        GPN_CK(cudaEventRecord(c->ev_side[TS], c->side[TS]));
    } else {
        GPN_CK(cudaMemsetAsync(u1, 0, sizeof(double) * 2 * ld, c->stream));
        GPN_TRY(gpn_d6_block_solve(c, WT, ld, b1, bt, v1, vt, u1, ut));
    }
    {
        GPN_TRY(gpn_d6_t_reduce(c, s, E, ld, (const double*)q.work[7].p, lds, nb, Tb));
        mark("T");
        c->potrf_grid_cap = so > 0 ? sms - std::max(2, sms / 3) : sms;
        int st = gpn_potrf_dataflow(c, s, Tb, lds, winvT, iscal(c, 14), 0, want_grad ? U : nullptr, lds, want_grad ? ND : 0,
                                    want_grad ? WTt : nullptr, lds, 0);
        c->potrf_grid_cap = 0;
        if (st != 0) return st;
        mark("[T; I; U]");
        GPN_TRY(gpn_d6_sep_reduce(c, s, Tb, lds, want_grad ? WTt : nullptr, lds, want_grad ? U : nullptr, lds, ND, dscal(c, 9)));
    }
    if (so > 0) GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_side[TS], 0));
    {
        const double *dx[6] = {b1, b1, bt, u1, u1, ut}, *dy[6] = {u1, ut, ut, u1, ut, ut};
        const int dn[6] = {N, N, N, N, N, N};
        GPN_TRY(gpn_k_multi_dot(c, 6, dx, dy, dn, dscal(c, 13)));
    }
    mark("reductions, join, dots");
    // the scalar block travels to pinned host memory behind everything above; gpr_reglap_d6_finish waits for it
    GPN_CK(cudaMemcpyAsync(c->pinned, c->scal.p, sizeof(double) * Scal::NDBL + sizeof(int) * 16, cudaMemcpyDeviceToHost, c->stream));
    GPN_CK(cudaEventRecord(c->ev_done, c->stream));
    c->pend_grad = want_grad;
    if (timing) {
        GPN_CK(cudaStreamSynchronize(c->stream));
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, "d6 %-26s %8.3f ms\n", marks[i].first, ms);
        }
        for (auto& m : marks) cudaEventDestroy(m.second);
    }
    return 0;
}

int gpr_reglap_d6_finish(gpn_ctx* c, RowScalars* rs, int* info_h) {
    GPN_CK(cudaEventSynchronize(c->ev_done));
    const int n = c->n_train;
    const double* sc = (const double*)c->pinned;
    const int* info = (const int*)((const char*)c->pinned + sizeof(double) * Scal::NDBL);
    if (info_h) {
        *info_h = 0;
        for (int i = 0; i < 16 && *info_h == 0; ++i) *info_h = info[i];      // parts 0..11, 14 / 15 = separator complements
    }
    rs->n = n;
    rs->half_logdet = -((sc[7] + sc[10]) - (sc[8] + sc[12]));          // (1/2) log det K_tt = -(1/2) (log det M - log det M_oo)
    rs->s = sc[0] - sc[13];
    rs->a = sc[1] - sc[14];
    rs->q = sc[2] - sc[15];
    if (c->pend_grad) {
        const double trM = sc[5] + sc[9], trMoo = sc[6] + sc[11];
        rs->trSD = trM - trMoo - n;
        rs->D11 = sc[16] + n - rs->s;
        rs->D1t = sc[17] + sc[3] - rs->a;
        rs->Dtt = sc[18] + sc[4] - rs->q;
    }
    return 0;
}

int gpr_reglap_d6_scalars(gpn_ctx* c, const double* theta_h, int P, const double* t_h, int want_grad, RowScalars* rs, int* info_h) {
    GPN_TRY(gpr_reglap_d6_enqueue(c, theta_h, P, t_h, want_grad));
    return gpr_reglap_d6_finish(c, rs, info_h);
}

// level 6 applies when the whole graph dissects into parts with thin separators; otherwise the caller takes level 5 / 4
bool d6_route(gpn_ctx* c) {
    if (c->fast_reglap < 6) return false;
    if (gpn_d6_ensure(c) != 0) return false;
    return c->d6.usable;
}

// Which evaluations take the sparse-coupling route: gradients from level 3 on; values alone only when the inverse of the
// other-node block comes from the dissected halves (level 5): otherwise the dense Cholesky of M of the Schur route is as fast.
bool coupled_route(gpn_ctx* c, int want_grad) {
    if (want_grad) return c->fast_reglap >= 3;
    const int o = c->N - c->n_train;
    if (c->fast_reglap < 5 || o < 256 || c->n_train < 256) return false;
    if (!c->perm_valid && ensure_perm(c) != 0) return false;
    return c->dis_s > 0;
}

// Schur route with sparse coupling (gradient evaluations): the only link between the other nodes (o) and the training nodes (t)
// in M = I + beta L~ is the SPARSE block M_to (graph edges between the two sets), so the elimination of the o-block does not
// need the dense N x N factorisation:
//     Z = M_oo^-1 (dense SPD inverse, o^3),   X = M_to Z (sparse x dense, memory bound),   S = M_tt - X M_ot = L_tt L_tt^T,
//     W_tt = L_tt^-1,   W_to = -W_tt L_to W_oo = -W_tt M_to M_oo^-1 = -W_tt X            (one triangular n x o x n product)
// and the reductions of the Schur route run over [W_to  W_tt] as before: o^3 + n^2 o + 2 n^3/3 flops instead of 2 N^3/3
// (half at n = N/2).  S, L_tt and the reductions are the same quantities as on level 2, so results agree to rounding.
int gpr_reglap_coupled_scalars(gpn_ctx* c, const double* theta_h, int P, const double* t_h, int want_grad, RowScalars* rs, int* info_h) {
    const int N = c->N, n = c->n_train, o = N - n;
    GPN_TRY(ensure_perm(c));
    GPN_TRY(ensure_vecs(c));
    c->have_full = false;
    c->kind = GPN_KERNEL_REGULARIZED_LAPLACIAN;
    c->P = P;
    for (int i = 0; i < P; ++i) c->theta_exp[i] = theta_h[i];
    c->ldN = pad128(N);
    GPN_TRY(clear_info(c));
    const long ld = c->ldN, ldn = pad128(n);
    const double beta = c->theta_exp[0];
    for (int i = 0; i < 4; ++i) GPN_TRY(ensure_mat(c, i));
    double* M = mat(c, 0);
    if (t_h) GPN_TRY(upload(c, vecp(c, V_T), t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(vecp(c, V_T), vecp(c, V_TRES), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    GPN_TRY(gpn_k_fill_zero_lower(c, M, ld, (int)ld, o, o));   // M_oo and M_tt (lower tiles); of M_to only the scattered entries are read
    GPN_TRY(gpn_k_laplacian_scatter(c, N, (const int*)c->p_indptr.p, (const int*)c->p_indices.p,
                                    c->h_weights.empty() ? nullptr : (const double*)c->p_weights.p, (const double*)c->p_dinv.p, beta, 1.0, M, ld));
    const int *tptr = (const int*)c->to_indptr.p, *tidx = (const int*)c->to_idx.p;
    double* tval = (double*)c->to_val.p;
    GPN_TRY(gpn_k_csr_values(c, n, o, tptr, tidx, M, ld, tval));
    GPN_TRY(ensure_nb(c, NB_ROWS1, n, ld));
    double* X = nbp(c, NB_ROWS1);
    GPN_TRY(ensure_nb(c, NB_KY, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_WT, ldn, ldn));
    double* S = nbp(c, NB_KY);
    double *ones = vecp(c, V_U), *r1 = vecp(c, V_Z), *rt = vecp(c, V_Z2);
    GPN_TRY(gpn_k_fill_vec(c, n, ones, 1.0));
    GPN_TRY(ensure_nb(c, NB_S, 2, ld));
    double *z1 = nbp(c, NB_S), *zt = z1 + ld;
    const bool augmented = c->fast_reglap >= 4 && o >= 256 && n >= 256;
    c->last_route = !augmented ? 3 : (c->fast_reglap >= 5 && c->dis_s > 0 ? 5 : 4);
    if (augmented) {
        // Both factorisations are small enough to be bound by their per-column dependency chain, so the triangular work that
        // follows them rides along in the same dataflow kernel as appended row blocks (potrf_dataflow.cu):
        //   [M_oo; I]      -> L_oo and W_oo^T            then Z = W_oo^T W_oo
        //   [S; X^T; I]    -> L_tt, Y = X^T L_tt^-T = (W_tt X)^T = -W_to^T, and W_tt^T
        GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(std::max(n, o)) * 128));
        if (c->fast_reglap >= 5 && c->dis_s > 0) {
            int used = 0;
            GPN_TRY(reglap_inverse_oo_dissected(c, M, ld, &used));                                // Z from two independent halves
        } else {
            GPN_TRY(gpn_potrf_dataflow(c, o, M, ld, (double*)c->winv.p, iscal(c, 0), 0, nullptr, 0, 0, mat(c, 2), ld));
            GPN_TRY(gpn_lauum_t(c, o, mat(c, 2), ld, M, ld, true));                               // Z, full symmetric, in place of M_oo
        }
        const long po = pad128(o);
        GPN_TRY(ensure_nb(c, NB_ROWS2, po, ldn));
        double* Y = nbp(c, NB_ROWS2);                                                             // X^T (o x n), rows padded with zeros
        if (po > o) GPN_TRY(gpn_k_fill_zero(c, Y + (po - 128) * ldn, ldn, 128, (int)ldn));
        GPN_TRY(gpn_k_spmm_rows_t(c, n, o, tptr, tidx, tval, M, ld, Y, ldn));                    // X = M_to Z, written transposed
        // S = M_tt - M_to (Z M_ot) = M_tt - M_to X^T, row by row: the sparse rows of M_to times the dense X^T (lower triangle)
        GPN_TRY(gpn_k_spmm_rows(c, n, n, tptr, tidx, tval, Y, ldn, S, ldn, M + (long)o * ld + o, ld));
        if (!want_grad) {
            // -logp alone: L_tt = chol(S), two triangular GEMVs and the log-determinant; no right-hand sides, no inverse
            GPN_TRY(gpn_potrf_dataflow(c, n, S, ldn, (double*)c->winv.p, iscal(c, 1), 0));
            GPN_TRY(gpn_k_gemv_t2(c, n, n, S, ldn, ones, vecp(c, V_T), r1, rt, GEMV_LOWER));
            GPN_TRY(gpn_k_sumlogdiag(c, n, S, ldn, dscal(c, Scal::LOGDET)));
            const double *dx3[3] = {r1, r1, rt}, *dy3[3] = {r1, rt, rt};
            const int dn3[3] = {n, n, n};
            GPN_TRY(gpn_k_multi_dot(c, 3, dx3, dy3, dn3, dscal(c, Scal::SCH0)));
            double sc0[Scal::NDBL];
            int info0[16];
            GPN_TRY(read_scalars(c, sc0, Scal::NDBL, info0, 16));
            if (info_h) *info_h = info0[1] ? info0[1] : (info0[0] ? info0[0] : (info0[2] ? info0[2] : info0[3]));
            rs->n = n;
            rs->half_logdet = -sc0[Scal::LOGDET];
            rs->s = sc0[Scal::SCH0]; rs->a = sc0[Scal::SCH0 + 1]; rs->q = sc0[Scal::SCH0 + 2];
            return 0;
        }
        GPN_TRY(gpn_potrf_dataflow(c, n, S, ldn, (double*)c->winv.p, iscal(c, 1), 0, Y, ldn, o, nbp(c, NB_WT), ldn));
        GPN_TRY(gpn_k_gemv_t2(c, n, n, S, ldn, ones, vecp(c, V_T), r1, rt, GEMV_LOWER));
        GPN_TRY(gpn_k_sumlogdiag(c, n, S, ldn, dscal(c, Scal::LOGDET)));
        // z = Wb^T y = [-Y y ; W_tt^T y] and W_tt^T L_tt^T x = x exactly: the training part of z is 1 resp. t, only Y is swept
        GPN_TRY(gpn_k_rowdot_pass(c, o, n, Y, ldn, r1, rt, z1, zt, dscal(c, Scal::FRO), false, true));
        GPN_TRY(gpn_k_rowdot_pass(c, n, n, nbp(c, NB_WT), ldn, r1, rt, nullptr, nullptr, dscal(c, Scal::FRO), true, false));
        const double *dx[8] = {r1, r1, rt, z1, z1, zt, ones, vecp(c, V_T)}, *dy[8] = {r1, rt, rt, z1, zt, zt, vecp(c, V_T), vecp(c, V_T)};
        const int dn[8] = {n, n, n, o, o, o, n, n};
        GPN_TRY(gpn_k_multi_dot(c, 8, dx, dy, dn, dscal(c, Scal::SCH0)));
    } else {
        // Z = M_oo^-1 in place (full symmetric, leading o x o block of M); rows >= o of M keep M_to and the lower part of M_tt
        GPN_TRY(spd_inverse(c, o, M, ld, mat(c, 1), mat(c, 2), mat(c, 3), M, 0));
        GPN_TRY(gpn_k_spmm_rows(c, n, o, tptr, tidx, tval, M, ld, X, ld));
        GPN_TRY(ensure_nb(c, NB_W, ldn, ldn));
        GPN_TRY(ensure_nb(c, NB_T, ldn, ldn));
        GPN_TRY(gpn_k_schur_update(c, n, M + (long)o * ld + o, ld, tptr, tidx, tval, X, ld, S, ldn));
        GprFactor f;
        GPN_TRY(gpr_factor_ky(c, &f));                                                              // L_tt (in S), W_tt, W_tt^T
        GPN_TRY(gpn_k_gemv_t2(c, n, n, S, ldn, ones, vecp(c, V_T), r1, rt, GEMV_LOWER));
        GPN_TRY(gpn_k_sumlogdiag(c, n, S, ldn, dscal(c, Scal::LOGDET)));
        // P = W_tt X = -W_to  (W_tt lower triangular: only k < m0 + 128 contributes)
        GPN_TRY(ensure_nb(c, NB_ROWS2, n, ld));
        double* Pm = nbp(c, NB_ROWS2);
        if (o > 0) GPN_TRY(gpn_gemm(c, n, o, n, 1.0, f.W, f.ldn, true, X, ld, false, 0.0, Pm, ld, GF_KHI_M));
        GPN_TRY(gpn_k_trap_pass(c, n, o, N, Pm, ld, r1, rt, z1, zt, dscal(c, Scal::FRO), true));
        GPN_TRY(gpn_k_trap_pass(c, n, n, 0, f.W, f.ldn, r1, rt, z1 + o, zt + o, dscal(c, Scal::FRO), false));
        const double *dx[6] = {r1, r1, rt, z1, z1, zt}, *dy[6] = {r1, rt, rt, z1, zt, zt};
        const int dn[6] = {n, n, n, N, N, N};
        GPN_TRY(gpn_k_multi_dot(c, 6, dx, dy, dn, dscal(c, Scal::SCH0)));
    }
    double sc[Scal::NDBL];
    int info[16];
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    if (info_h) *info_h = info[1] ? info[1] : (info[0] ? info[0] : (info[2] ? info[2] : info[3]));
    rs->n = n;
    rs->half_logdet = -sc[Scal::LOGDET];
    rs->s = sc[Scal::SCH0]; rs->a = sc[Scal::SCH0 + 1]; rs->q = sc[Scal::SCH0 + 2];
    rs->D11 = sc[Scal::SCH0 + 3] - rs->s; rs->D1t = sc[Scal::SCH0 + 4] - rs->a; rs->Dtt = sc[Scal::SCH0 + 5] - rs->q;
    if (augmented) {        // the training part of z1 is the ones vector, that of zt is t
        rs->D11 += n; rs->D1t += sc[Scal::SCH0 + 6]; rs->Dtt += sc[Scal::SCH0 + 7];
    }
    rs->trSD = sc[Scal::FRO] - n;
    return 0;
}

int gpr_logp_grad_reglap_schur(gpn_ctx* c, const double* theta_h, int P, const double* t_h, int want_grad, double* out_h, int* info_h) {
    RowScalars rs;
    if (d6_route(c)) GPN_TRY(gpr_reglap_d6_scalars(c, theta_h, P, t_h, want_grad, &rs, info_h));
    else if (coupled_route(c, want_grad)) GPN_TRY(gpr_reglap_coupled_scalars(c, theta_h, P, t_h, want_grad, &rs, info_h));
    else GPN_TRY(gpr_reglap_schur_scalars(c, theta_h, P, t_h, want_grad, &rs, info_h));
    row_eval(rs, c->theta_exp[P - 1], P, want_grad, out_h);
    return 0;
}

// The same reductions for any kernel from the resident full kernel: L = chol(K_tt) (NO noise), W = L^-1, u = W 1, v = W t,
// s = u.u, a = u.v, q = v.v; S1 = W^T u, St = W^T v, S = W^T W.  Used by the landscape driver, where every noise value of a
// grid row shares them.  info > 0: K_tt itself is not positive definite (the caller falls back to factoring K_y per point).
int gpr_row_scalars(gpn_ctx* c, int kind, const double* theta_h, int P, const double* t_h, int want_grad, RowScalars* rs, int* info_h) {
    if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN && c->fast_reglap >= 2 && c->train_unique && c->n_train < c->N) {
        if (d6_route(c)) return gpr_reglap_d6_scalars(c, theta_h, P, t_h, want_grad, rs, info_h);
        return coupled_route(c, want_grad) ? gpr_reglap_coupled_scalars(c, theta_h, P, t_h, want_grad, rs, info_h)
                                                  : gpr_reglap_schur_scalars(c, theta_h, P, t_h, want_grad, rs, info_h);
    }
    int st = kernel_build(c, kind, theta_h, P, want_grad);
    if (st != 0) return st;
    GPN_TRY(ensure_vecs(c));
    const int n = c->n_train;
    const long ldn = pad128(n);
    if (t_h) GPN_TRY(upload(c, vecp(c, V_T), t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(vecp(c, V_T), vecp(c, V_TRES), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    const bool reglap = kind == GPN_KERNEL_REGULARIZED_LAPLACIAN;
    const double noise_saved = c->theta_exp[P - 1];
    c->theta_exp[P - 1] = 0.0;                                   // gather K_tt without the noise term
    st = gpr_gather_ky(c, want_grad && reglap);
    c->theta_exp[P - 1] = noise_saved;
    if (st != 0) return st;
    GprFactor f;
    GPN_TRY(gpr_factor_ky(c, &f));
    double *ones = vecp(c, V_U), *u = vecp(c, V_Z), *v = vecp(c, V_Z2), *x1 = vecp(c, V_ALPHA), *xt = vecp(c, V_SOL);
    GPN_TRY(gpn_k_fill_vec(c, n, ones, 1.0));
    GPN_TRY(gpn_k_gemv(c, n, n, f.W, f.ldn, ones, u, GEMV_LOWER));
    GPN_TRY(gpn_k_gemv(c, n, n, f.W, f.ldn, vecp(c, V_T), v, GEMV_LOWER));
    GPN_TRY(gpn_k_sumlogdiag(c, n, f.Ky, f.ldn, dscal(c, Scal::LOGDET)));
    const double *dx[3] = {u, u, v}, *dy[3] = {u, v, v};
    const int dn[3] = {n, n, n};
    GPN_TRY(gpn_k_multi_dot(c, 3, dx, dy, dn, dscal(c, Scal::SCH0)));
    if (want_grad) {
        const double* G = nullptr;
        double gs = 0.0;
        GPN_TRY(gpr_grad_kernel_block(c, kind, ldn, &G, &gs));
        GPN_TRY(gpn_k_gemv_t(c, n, n, f.W, f.ldn, u, x1, GEMV_LOWER));
        GPN_TRY(gpn_k_gemv_t(c, n, n, f.W, f.ldn, v, xt, GEMV_LOWER));
        GPN_TRY(gpn_lauum_t(c, n, f.WT, f.ldn, f.T, f.ldn, false));            // S = W^T W (lower) -> NB_T
        GPN_TRY(gpn_k_bilinear_trace(c, n, f.T, f.ldn, G, f.ldn, reglap ? nbp(c, NB_KC) : nullptr, f.ldn, gs, reglap ? -1.0 : 0.0, x1, xt,
                                     dscal(c, Scal::SCH0 + 3)));
    }
    double sc[Scal::NDBL];
    int info[16];
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    if (info_h) *info_h = info[1] ? info[1] : info[0];
    rs->n = n;
    rs->half_logdet = sc[Scal::LOGDET];
    rs->s = sc[Scal::SCH0]; rs->a = sc[Scal::SCH0 + 1]; rs->q = sc[Scal::SCH0 + 2];
    if (want_grad) { rs->trSD = sc[Scal::SCH0 + 3]; rs->D11 = sc[Scal::SCH0 + 4]; rs->D1t = sc[Scal::SCH0 + 5]; rs->Dtt = sc[Scal::SCH0 + 6]; }
    return 0;
}

int gpr_logp_grad(gpn_ctx* c, int kind, const double* theta_h, int P, const double* t_h, int want_grad, double* out_h, int* info_h) {
    if (c->n_train <= 0) return -1;
    if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN && c->fast_reglap && c->train_unique && P >= 1 && P <= 8 && c->n_train < c->N)
        return c->fast_reglap >= 2 ? gpr_logp_grad_reglap_schur(c, theta_h, P, t_h, want_grad, out_h, info_h)
                                   : gpr_logp_grad_reglap_fast(c, theta_h, P, t_h, want_grad, out_h, info_h);
    int st = kernel_build(c, kind, theta_h, P, want_grad);
    if (st != 0) return st;
    c->last_route = 0;
    GPN_TRY(ensure_vecs(c));
    const int n = c->n_train;
    const long ldn = pad128(n);
    if (t_h) GPN_TRY(upload(c, vecp(c, V_T), t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(vecp(c, V_T), vecp(c, V_TRES), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    const double noise = c->theta_exp[P - 1];
    const double* G = nullptr;
    double gs = 0.0;
    bool forked = false;
    GPN_TRY(gpr_gather_ky(c, want_grad && kind == GPN_KERNEL_REGULARIZED_LAPLACIAN));
    GprFactor f;
    if (want_grad) {
        // The Cholesky of K_y is bound by its per-column dependency chain, not by flops: give it only the CTAs it needs and
        // run the (independent, flop-bound) gradient GEMM on the side stream at the same time.  The side stream is released
        // by an event recorded right before the Cholesky launch so that the cooperative kernel is dispatched first.
        cudaStream_t main_stream = c->stream;
        GPN_CK(cudaEventRecord(c->ev_fork, main_stream));
        c->potrf_grid_cap = chain_bound_ctas(n);
        st = gpr_factor_ky(c, &f);
        c->potrf_grid_cap = 0;
        if (st != 0) return st;
        GPN_CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        c->stream = c->stream2;
        st = gpr_grad_kernel_block(c, kind, ldn, &G, &gs);
        c->stream = main_stream;
        if (st != 0) return st;
        GPN_CK(cudaEventRecord(c->ev_join, c->stream2));
        forked = true;
    } else {
        GPN_TRY(gpr_factor_ky(c, &f));
    }
    // z = L^-1 t, alpha = L^-T z, t^T alpha = z^T z                         (GPnetRegressor.py:144-148)
    GPN_TRY(gpn_k_gemv(c, n, n, f.W, f.ldn, vecp(c, V_T), vecp(c, V_Z), GEMV_LOWER));
    GPN_TRY(gpn_k_dot(c, n, vecp(c, V_Z), vecp(c, V_Z), dscal(c, Scal::DOT)));
    GPN_TRY(gpn_k_sumlogdiag(c, n, f.Ky, f.ldn, dscal(c, Scal::LOGDET)));
    if (want_grad) {
        GPN_TRY(gpn_k_gemv_t(c, n, n, f.W, f.ldn, vecp(c, V_Z), vecp(c, V_ALPHA), GEMV_LOWER));
        GPN_TRY(gpn_lauum_t(c, n, f.WT, f.ldn, f.T, f.ldn, false));            // Kinv = W^T W (lower) -> NB_T
        const double* Ktt = nullptr;
        double ks = 0.0, kshift = 0.0;
        if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN) { Ktt = nbp(c, NB_KC); ks = -1.0; kshift = noise; }
        if (forked) GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        GPN_TRY(gpn_k_grad_trace(c, n, f.T, f.ldn, G, f.ldn, Ktt, f.ldn, gs, ks, kshift, vecp(c, V_ALPHA), dscal(c, Scal::TR0)));
    }
    double sc[Scal::NDBL];
    int info[16];
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    if (info_h) *info_h = info[1] ? info[1] : info[0];
    out_h[0] = 0.5 * sc[Scal::DOT] + sc[Scal::LOGDET] + 0.5 * n * LOG_2PI;     // -logp (GPnetRegressor.py:145-150)
    if (want_grad) {
        for (int j = 0; j < P; ++j) out_h[1 + j] = 0.0;
        const double sinv = sc[Scal::TR0], tr0 = sc[Scal::TR1], sa = sc[Scal::TR2];
        out_h[1 + 0] += -0.5 * tr0;                                              // -(1/2)(a^T D a - tr(K^-1 D))
        out_h[1 + (P - 1)] += -0.5 * noise * (sa * sa - sinv);                   // D_last = c 1 1^T
    }
    return 0;
}

int gpr_predict(gpn_ctx* c, int kind, const double* theta_h, int P, const double* t_h, double* alpha_h, double* fstar_h, double* s_h,
                double* kss_diag_h, int* info_h) {
    if (c->n_train <= 0) return -1;
    int st = kernel_build(c, kind, theta_h, P, 0);
    if (st != 0) return st;
    GPN_TRY(ensure_vecs(c));
    const int n = c->n_train, m = c->n_test;
    const long ldm = pad128(m);
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += t_h[i];
    mean /= n;                                                  // np.mean: pairwise in numpy; difference is O(ulp), see DESIGN.md
    std::vector<double> ts(n);
    for (int i = 0; i < n; ++i) ts[i] = t_h[i] - mean;          // GPnetRegressor.py:86-87
    GPN_TRY(upload(c, vecp(c, V_T), ts.data(), sizeof(double) * n));
    GprFactor f;
    GPN_TRY(gpr_factor(c, true, &f));
    const int* tr = (const int*)c->train_idx.p;
    const int* te = (const int*)c->test_idx.p;
    const double noise = c->theta_exp[P - 1];
    GPN_TRY(gpn_k_gemv(c, n, n, f.W, f.ldn, vecp(c, V_T), vecp(c, V_Z), GEMV_LOWER));
    GPN_TRY(gpn_k_gemv_t(c, n, n, f.W, f.ldn, vecp(c, V_Z), vecp(c, V_ALPHA), GEMV_LOWER));      // :124
    if (m > 0) {
        GPN_TRY(ensure_nb(c, NB_KSTAR, ldm, f.ldn));
        GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, te, m, tr, n, 0.0, nbp(c, NB_KSTAR), f.ldn));   // kstar, measnoise=False (:96-102)
        GPN_TRY(gpn_k_gemv(c, m, n, nbp(c, NB_KSTAR), f.ldn, vecp(c, V_ALPHA), vecp(c, V_FSTAR), GEMV_FULL));   // :125
        // v = L^-1 kstar^T = W kstar^T  (n x m)                                                   (:126)
        GPN_TRY(ensure_nb(c, NB_V, f.ldn, ldm));
        GPN_TRY(gpn_gemm(c, n, m, n, 1.0, f.W, f.ldn, true, nbp(c, NB_KSTAR), f.ldn, true, 0.0, nbp(c, NB_V), ldm, GF_KHI_M));
        GPN_TRY(gpn_k_colnorm2(c, n, m, nbp(c, NB_V), ldm, vecp(c, V_VN)));                         // diag(v^T v) (:127-128)
        GPN_TRY(gpn_k_diag_gather(c, kfull(c), c->ldN, te, m, noise, vecp(c, V_KSS)));             // kstarstar_diag (:110)
        // PD gate on kstarstar (:117-121) through a Cholesky instead of eigvals (quirk A-7)
        GPN_TRY(ensure_nb(c, NB_KSS, ldm, ldm));
        GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, te, m, te, m, noise, nbp(c, NB_KSS), ldm));
        GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(std::max(std::max(n, m), c->N)) * 128));
        GPN_TRY(gpn_potrf(c, m, nbp(c, NB_KSS), ldm, (double*)c->winv.p, iscal(c, 2)));
    }
    std::vector<double> fs(m), vn(m), kss(m);
    GPN_TRY(download(c, alpha_h, vecp(c, V_ALPHA), sizeof(double) * n));
    if (m > 0) {
        GPN_TRY(download(c, fs.data(), vecp(c, V_FSTAR), sizeof(double) * m));
        GPN_TRY(download(c, vn.data(), vecp(c, V_VN), sizeof(double) * m));
        GPN_TRY(download(c, kss.data(), vecp(c, V_KSS), sizeof(double) * m));
    }
    int info[16];
    GPN_TRY(read_scalars(c, nullptr, 0, info, 16));
    info_h[0] = info[1] ? info[1] : info[0];
    info_h[1] = info[2];
    for (int i = 0; i < m; ++i) {
        fstar_h[i] = fs[i] + mean;
        kss_diag_h[i] = kss[i];
        s_h[i] = sqrt(kss[i] - vn[i]);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// GP classification: Laplace Newton loop (GPnetClassifier.py:90-147)
// ---------------------------------------------------------------------------------------------------------------------
struct NewtonState {
    int n;
    long ldn;
    const double* K;
    long ldk;
    bool have_wk;
};

// factor B = I + sw K sw from the current V_SW; leaves L_B in NB_B, W_B in NB_WB
int factor_B(gpn_ctx* c, const NewtonState& s) {
    GPN_TRY(gpn_k_form_B(c, s.n, s.K, s.ldk, vecp(c, V_SW), nbp(c, NB_B), s.ldn));
    GPN_TRY(gpn_potrf(c, s.n, nbp(c, NB_B), s.ldn, (double*)c->winv.p, iscal(c, 3)));
    return gpn_trtri(c, s.n, nbp(c, NB_B), s.ldn, (const double*)c->winv.p, nbp(c, NB_WB), nbp(c, NB_WT), nbp(c, NB_T), s.ldn);
}

// The same from ONE augmented dataflow factorisation [B; I]: L_B in NB_B and W_B^T = L_B^-T (upper tiles) in
// NB_WT; the triangular inverse rides along in the Cholesky kernel instead of ten batched GEMM launches.  NB_WB is NOT produced.
int factor_B_fast(gpn_ctx* c, const NewtonState& s, double* rhs_rows = nullptr, int nrhs = 0) {
    GPN_TRY(gpn_k_form_B(c, s.n, s.K, s.ldk, vecp(c, V_SW), nbp(c, NB_B), s.ldn));
    return gpn_potrf_dataflow(c, s.n, nbp(c, NB_B), s.ldn, (double*)c->winv.p, iscal(c, 3), 0, rhs_rows, s.ldn, nrhs,
                              rhs_rows ? nullptr : nbp(c, NB_WT), s.ldn, 0);
}

int gpc_newton(gpn_ctx* c, const double* K, long ldk, int n, const double* y_h, double tol, double phif0, double scale, int variant,
               double* f_h, double* a_h, double* logq_h, int* iters_h, int* info_h) {
    GPN_TRY(ensure_vecs(c));
    NewtonState s;
    s.n = n; s.ldn = pad128(n); s.K = K; s.ldk = ldk;
    GPN_TRY(ensure_nb(c, NB_B, s.ldn, s.ldn));
    GPN_TRY(ensure_nb(c, NB_WB, s.ldn, s.ldn));
    GPN_TRY(ensure_nb(c, NB_T, s.ldn, s.ldn));
    GPN_TRY(ensure_nb(c, NB_WK, s.ldn, s.ldn));
    GPN_TRY(ensure_nb(c, NB_KC, s.ldn, s.ldn));
    GPN_TRY(ensure_nb(c, NB_WT, s.ldn, s.ldn));
    const size_t wblk = (size_t)pad128(std::max(n, c->N)) * 128;
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * 2 * wblk));
    GPN_TRY(upload(c, vecp(c, V_Y), y_h, sizeof(double) * n));
    GPN_CK(cudaMemsetAsync(vecp(c, V_F), 0, sizeof(double) * n, c->stream));                    // f = 0 (:96)
    // f^T K^-1 f of the convergence test (:126) through ONE Cholesky of K (K is fixed during the loop): [K; I] in one dataflow
    // kernel gives W_K^T = L_K^-T (upper tiles) -> NB_WK.  It runs on the side stream, on half of the SMs, next to the first
    // factorisation of B (own flag slot, own block of inverted diagonal tiles); the loop joins it before its first use.
    int info[16];
    cudaStream_t main_stream = c->stream;
    {
        GPN_CK(cudaEventRecord(c->ev_fork, main_stream));
        GPN_CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        c->stream = c->stream2;
        int st = gpn_k_fill_zero(c, nbp(c, NB_KC), s.ldn, (int)s.ldn, (int)s.ldn);
        if (st == 0) st = gpn_k_copy(c, n, n, K, ldk, nbp(c, NB_KC), s.ldn);
        c->potrf_grid_cap = c->num_sms / 2;
        if (st == 0) st = gpn_potrf_dataflow(c, n, nbp(c, NB_KC), s.ldn, (double*)c->winv.p + wblk, iscal(c, 4), 0, nullptr, 0, 0, nbp(c, NB_WK), s.ldn, 1);
        c->potrf_grid_cap = 0;
        c->stream = main_stream;
        if (st != 0) return st;
        GPN_CK(cudaEventRecord(c->ev_join, c->stream2));
    }
    s.have_wk = true;      // decided on the device from the Cholesky's info (newton_phif_kernel)
    bool joined = false;
    int count = 0, iters = 0;
    double sc[Scal::NDBL];
    while (true) {
        ++count;
        ++iters;
        newton_pre_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_F), vecp(c, V_Y), vecp(c, V_W), vecp(c, V_SW), vecp(c, V_P),
                                                         vecp(c, V_B), vecp(c, V_G));
        c->launches++;
        if (!joined) c->potrf_grid_cap = c->num_sms - c->num_sms / 2;                             // next to [K; I]
        {
            const int stf = factor_B_fast(c, s);                                                 // :110
            c->potrf_grid_cap = 0;
            if (stf != 0) return stf;
        }
        GPN_TRY(gpn_k_gemv(c, n, n, K, ldk, vecp(c, V_B), vecp(c, V_KB), GEMV_FULL));            // K b
        vec_mul_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_SW), vecp(c, V_KB), vecp(c, V_RHS));
        c->launches++;
        // B^-1 rhs = W^T (W rhs) with W^T = L_B^-T stored upper: W x = (W^T)^T x
        GPN_TRY(gpn_k_gemv_t(c, n, n, nbp(c, NB_WT), s.ldn, vecp(c, V_RHS), vecp(c, V_Z2), GEMV_UPPER));
        GPN_TRY(gpn_k_gemv(c, n, n, nbp(c, NB_WT), s.ldn, vecp(c, V_Z2), vecp(c, V_SOL), GEMV_UPPER));
        newton_a_kernel<<<g1(n), 256, 0, c->stream>>>(n, scale, vecp(c, V_B), vecp(c, V_SW), vecp(c, V_SOL), vecp(c, V_A));
        c->launches++;
        GPN_TRY(gpn_k_gemv(c, n, n, K, ldk, vecp(c, V_A), vecp(c, V_F), GEMV_FULL));             // f = K a (:122)
        if (!joined) {
            GPN_CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
            joined = true;
        }
        GPN_TRY(gpn_k_gemv_t(c, n, n, nbp(c, NB_WK), s.ldn, vecp(c, V_F), vecp(c, V_U), GEMV_UPPER));          // u = W_K f
        GPN_TRY(gpn_k_dot(c, n, vecp(c, V_U), vecp(c, V_U), dscal(c, Scal::Q)));
        GPN_TRY(gpn_k_dot(c, n, vecp(c, V_F), vecp(c, V_A), dscal(c, Scal::AF)));                              // the fall-back: f^T a
        GPN_TRY(gpn_k_sumlogdiag(c, n, nbp(c, NB_B), s.ldn, dscal(c, Scal::LOGDET)));
        newton_phif_kernel<<<1, 1024, 0, c->stream>>>(n, vecp(c, V_P), dscal(c, Scal::Q), dscal(c, Scal::AF), iscal(c, 4),
                                                       dscal(c, Scal::LOGDET), vecp(c, V_PHIF), iters == 1, phif0, dscal(c, Scal::CONV));
        c->launches++;
        GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
        if (info[3] != 0) {     // np.linalg.cholesky raised inside the loop: propagate (GPnetClassifier.py:110)
            *info_h = info[3];
            *iters_h = iters;
            return 0;
        }
        if (sc[Scal::CONV] < tol) break;                                                         // :132
        if (count > 100) {                                                                       // :134-136
            count = 0;
            scale /= 2.0;
        }
        if (iters > 100000) return -95;
    }
    // logq (:138-145): uses L of the LAST iteration (computed from the previous f)
    GPN_TRY(gpn_k_dot(c, n, vecp(c, V_A), vecp(c, V_F), dscal(c, Scal::AF)));
    newton_lik_kernel<<<1, 1024, 0, c->stream>>>(n, vecp(c, V_F), vecp(c, V_Y), variant, dscal(c, Scal::LIK));
    c->launches++;
    if (f_h) GPN_TRY(download(c, f_h, vecp(c, V_F), sizeof(double) * n));
    if (a_h) GPN_TRY(download(c, a_h, vecp(c, V_A), sizeof(double) * n));
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    *logq_h = -0.5 * sc[Scal::AF] - sc[Scal::LIK] - sc[Scal::LOGDET];
    *iters_h = iters;
    *info_h = 0;
    return 0;
}

int gpc_train_block(gpn_ctx* c) {   // K = kernel(train, train) with noise on every entry -> NB_KY
    const int n = c->n_train;
    const long ldn = pad128(n);
    GPN_TRY(ensure_nb(c, NB_KY, ldn, ldn));
    const int* tr = (const int*)c->train_idx.p;
    return gpn_k_gather(c, kfull(c), c->ldN, tr, n, tr, n, c->theta_exp[c->P - 1], nbp(c, NB_KY), ldn);
}

// GPnetClassifier.gradLogPosterior (GPnetClassifier.py:149-182): RW Alg. 5.1 as written there, fed by the derivative
// kernels of SURVEY Appendix C (the reference's own kernel() no longer returns them, quirk A-12).  out_h[0] = -logq,
// out_h[1..P] = -gradZ.  Every expression is linear in dK_d, so the kernel-parameter part (d = 0) and the all-entries
// noise part c 1 1^T (d = P-1) are accumulated separately; the rank-one part needs only vector sums.
int gpc_grad(gpn_ctx* c, int kind, const double* theta_h, int P, const double* y_h, double* out_h, int* iters_h, int* info_h) {
    int st = kernel_build(c, kind, theta_h, P, 1);
    if (st != 0) return st;
    const int n = c->n_train;
    const long ldn = pad128(n);
    GPN_TRY(gpc_train_block(c));
    const double* K = nbp(c, NB_KY);
    double logq = 0.0;
    GPN_TRY(gpc_newton(c, K, ldn, n, y_h, 0.1, 1e100, 1.0, 0, nullptr, nullptr, &logq, iters_h, info_h));   // :155
    out_h[0] = -logq;
    for (int j = 0; j < P; ++j) out_h[1 + j] = 0.0;
    if (*info_h != 0) return 0;
    const double noise = c->theta_exp[P - 1];
    const int* tr = (const int*)c->train_idx.p;
    NewtonState s;
    s.n = n; s.ldn = ldn; s.K = K; s.ldk = ldn; s.have_wk = false;
    // W, p from the converged f (:156-158, :163); B refactored (:159); Binv = W_B^T W_B (full)
    newton_pre_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_F), vecp(c, V_Y), vecp(c, V_W), vecp(c, V_SW), vecp(c, V_P), vecp(c, V_B),
                                                     vecp(c, V_G));
    c->launches++;
    GPN_TRY(factor_B(c, s));
    GPN_TRY(ensure_nb(c, NB_KSS, ldn, ldn));
    double* Binv = nbp(c, NB_KSS);
    GPN_TRY(gpn_lauum_t(c, n, nbp(c, NB_WT), ldn, Binv, ldn, true));
    // C = L^-1 sqrtW K = W_B (sw o K); only diag(C^T C) is needed (:162-166)
    GPN_TRY(ensure_nb(c, NB_S, ldn, ldn));
    GPN_TRY(ensure_nb(c, NB_V, ldn, ldn));
    GPN_TRY(gpn_k_scale_rows(c, n, n, K, ldn, vecp(c, V_SW), nbp(c, NB_S), ldn));
    GPN_TRY(gpn_gemm(c, n, n, n, 1.0, nbp(c, NB_WB), ldn, true, nbp(c, NB_S), ldn, false, 0.0, nbp(c, NB_V), ldn, GF_KHI_M));
    GPN_TRY(gpn_k_colnorm2(c, n, n, nbp(c, NB_V), ldn, vecp(c, V_VN)));
    diag_extract_kernel<<<g1(n), 256, 0, c->stream>>>(n, K, ldn, vecp(c, V_KSS));
    gpc_s2_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_KSS), vecp(c, V_VN), vecp(c, V_W), vecp(c, V_P), vecp(c, V_VV));   // s2 -> V_VV
    c->launches += 2;
    // s3(b) = b - K (R b), R b = sw o (Binv (sw o b));  result dotted with s2 into dscal(slot)
    auto s2_dot_s3 = [&](const double* b, int slot) -> int {
        vec_mul_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_SW), b, vecp(c, V_RHS));
        c->launches++;
        GPN_TRY(gpn_k_gemv(c, n, n, Binv, ldn, vecp(c, V_RHS), vecp(c, V_Z2), GEMV_FULL));
        vec_mul_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_SW), vecp(c, V_Z2), vecp(c, V_SOL));
        c->launches++;
        GPN_TRY(gpn_k_gemv(c, n, n, K, ldn, vecp(c, V_SOL), vecp(c, V_U), GEMV_FULL));
        vec_sub_kernel<<<g1(n), 256, 0, c->stream>>>(n, b, vecp(c, V_U), vecp(c, V_PI));
        c->launches++;
        return gpn_k_dot(c, n, vecp(c, V_VV), vecp(c, V_PI), dscal(c, slot));
    };
    // ---- d = 0: D0 = dK/dtheta0 [tr,tr] (full symmetric) ----
    const double* G = nullptr;
    double gs = 0.0;
    GPN_TRY(gpr_grad_kernel_block(c, kind, ldn, &G, &gs, true));
    bool have_d0 = (G != nullptr);
    if (have_d0) {
        double* D0 = nbp(c, NB_G);
        if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN) {      // D0 = (K^2)[tr,tr] - K[tr,tr]  (K without the noise)
            GPN_TRY(ensure_nb(c, NB_KC, ldn, ldn));
            GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, tr, n, tr, n, 0.0, nbp(c, NB_KC), ldn));
            const double* X[2] = {D0, nbp(c, NB_KC)};
            const long l2[2] = {ldn, ldn};
            const double cf[2] = {1.0, -1.0};
            GPN_TRY(gpn_k_lincomb(c, n, D0, ldn, 0.0, 2, X, l2, cf));
        } else if (gs != 1.0) {
            const double* X[1] = {D0};
            const long l1[1] = {ldn};
            const double cf[1] = {gs};
            GPN_TRY(gpn_k_lincomb(c, n, D0, ldn, 0.0, 1, X, l1, cf));
        }
        GPN_TRY(gpn_k_gemv(c, n, n, D0, ldn, vecp(c, V_A), vecp(c, V_KB), GEMV_FULL));
        GPN_TRY(gpn_k_dot(c, n, vecp(c, V_A), vecp(c, V_KB), dscal(c, Scal::TR0)));                 // a^T D0 a
        GPN_CK(cudaMemsetAsync(dscal(c, Scal::TR1), 0, sizeof(double), c->stream));
        gpc_trace_kernel<<<n, 256, 0, c->stream>>>(n, Binv, ldn, vecp(c, V_SW), D0, ldn, dscal(c, Scal::TR1));   // tr(R D0)
        c->launches++;
        GPN_TRY(gpn_k_gemv(c, n, n, D0, ldn, vecp(c, V_G), vecp(c, V_KB), GEMV_FULL));              // b = D0 ((t+1)/2 - p)
        GPN_TRY(s2_dot_s3(vecp(c, V_KB), Scal::TR2));
    }
    // ---- d = P-1: D = c 1 1^T ----
    GPN_TRY(gpn_k_sum(c, n, vecp(c, V_A), dscal(c, Scal::NORM0)));                                  // sum a
    GPN_TRY(gpn_k_sum(c, n, vecp(c, V_G), dscal(c, Scal::NORM1)));                                  // sum ((t+1)/2 - p)
    GPN_TRY(gpn_k_gemv(c, n, n, Binv, ldn, vecp(c, V_SW), vecp(c, V_Z), GEMV_FULL));
    GPN_TRY(gpn_k_dot(c, n, vecp(c, V_SW), vecp(c, V_Z), dscal(c, Scal::NORM2)));                   // sw^T Binv sw = sum_ij R_ij
    vec_fill_kernel<<<g1(n), 256, 0, c->stream>>>(n, 1.0, vecp(c, V_KB));
    c->launches++;
    GPN_TRY(s2_dot_s3(vecp(c, V_KB), Scal::NORM3));                                                 // s2 . s3(1)
    double sc[Scal::NDBL];
    int info[16];
    GPN_TRY(read_scalars(c, sc, Scal::NDBL, info, 16));
    if (info[3] != 0) { *info_h = info[3]; return 0; }
    double g0 = 0.0;
    if (have_d0) g0 = 0.5 * sc[Scal::TR0] - 0.5 * sc[Scal::TR1] + sc[Scal::TR2];
    const double sa = sc[Scal::NORM0], sg = sc[Scal::NORM1];
    const double gn = 0.5 * noise * sa * sa - 0.5 * noise * sc[Scal::NORM2] + noise * sg * sc[Scal::NORM3];
    out_h[1 + 0] += -g0;
    out_h[1 + (P - 1)] += -gn;
    return 0;
}

// the N x N scratch matrices belong to the bound graph's size: run `body` with the context temporarily sized for N
template <class F>
int with_matrix_size(gpn_ctx* c, int N, F body) {
    const int N0 = c->N;
    const long ld0 = c->ldN;
    c->N = N;
    c->ldN = pad128(N);
    c->have_full = false;
    const int st = body();
    c->N = N0;
    c->ldN = ld0;
    return st;
}

}   // namespace

const double* gpn_resident_targets(gpn_ctx* c) { return vecp(c, V_TRES); }

// =====================================================================================================================
// C-ABI
// =====================================================================================================================
extern "C" {

const char* gpn_last_error(void) { return g_last_error; }

gpn_ctx* gpn_create(int device, void* cuda_stream) {
    GPN_RANGE();
    if (cudaSetDevice(device) != cudaSuccess) {
        gpn_set_last_error(__FILE__, __LINE__, "cudaSetDevice failed (no CUDA device? this library has no CPU fallback)");
        return nullptr;
    }
    gpn_ctx* c = new gpn_ctx();
    c->device = device;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            gpn_set_last_error(__FILE__, __LINE__, "cudaStreamCreate failed");
            delete c;
            return nullptr;
        }
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        gpn_set_last_error(__FILE__, __LINE__, "cudaGetDeviceProperties failed");
        delete c;
        return nullptr;
    }
    c->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        gpn_set_last_error(__FILE__, __LINE__, "gpnet_b200 requires an sm_100a device (B200)");
        delete c;
        return nullptr;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
        gpn_set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available from the driver");
        delete c;
        return nullptr;
    }
    c->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    if (cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess) {
        gpn_set_last_error(__FILE__, __LINE__, "side stream / event creation failed");
        delete c;
        return nullptr;
    }
    if (const char* m = getenv("GPNET_B200_POTRF")) c->potrf_mode = (strcmp(m, "recursive") == 0) ? 1 : 0;
    if (const char* m = getenv("GPNET_B200_REGLAP_FAST")) c->fast_reglap = atoi(m);
    if (cudaMalloc(&c->scal.p, 1024) != cudaSuccess || cudaMallocHost(&c->pinned, 4096) != cudaSuccess) {
        gpn_set_last_error(__FILE__, __LINE__, "initial allocations failed");
        delete c;
        return nullptr;
    }
    c->scal.cap = 1024;
    c->pinned_cap = 4096;
    cudaMemsetAsync(c->scal.p, 0, 1024, c->stream);
    return c;
}

void gpn_destroy(gpn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->mat[9] = DevBuf();   // alias slot
    DevBuf* all[] = {&c->indptr, &c->indices, &c->weights, &c->dinv, &c->train_idx, &c->test_idx, &c->tvec, &c->winv, &c->scal, &c->partial,
                     &c->flags, &c->p_indptr, &c->p_indices, &c->p_weights, &c->p_dinv, &c->iota, &c->to_indptr, &c->to_idx, &c->to_val, &c->flags2, &c->sa_ptr, &c->sa_idx, &c->sa_val, &c->sb_ptr, &c->sb_idx,
                     &c->sb_idx_rel, &c->sb_val};
    for (DevBuf* b : all)
        if (b->p) cudaFree(b->p);
    for (auto& b : c->mat)
        if (b.p) cudaFree(b.p);
    for (auto& b : c->nb)
        if (b.p) cudaFree(b.p);
    {
        Spectral& sp = c->spec;
        DevBuf* spb[] = {&sp.VT, &sp.lam, &sp.rbuf, &sp.Vt, &sp.bt, &sp.bth, &sp.bout};
        for (DevBuf* b : spb)
            if (b->p) cudaFree(b->p);
    }
    {
        Dissect6& q = c->d6;
        DevBuf* d6b[] = {&q.indptr, &q.indices, &q.weights, &q.dinv, &q.trank, &q.sep_ptr, &q.sep_idx, &q.sep_w, &q.sep_pp};
        for (DevBuf* b : d6b)
            if (b->p) cudaFree(b->p);
        for (auto& b : q.work)
            if (b.p) cudaFree(b.p);
        for (auto& b : c->flagsK)
            if (b.p) cudaFree(b.p);
        for (int k = 0; k < D6_SLOTS; ++k) {
            if (c->side[k]) { cudaStreamSynchronize(c->side[k]); cudaStreamDestroy(c->side[k]); }
            if (c->ev_side[k]) cudaEventDestroy(c->ev_side[k]);
            if (c->side2[k]) { cudaStreamSynchronize(c->side2[k]); cudaStreamDestroy(c->side2[k]); }
            if (c->ev_g[k]) cudaEventDestroy(c->ev_g[k]);
            if (c->ev_u[k]) cudaEventDestroy(c->ev_u[k]);
        }
    }
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int gpn_sync(gpn_ctx* c) {
    GPN_RANGE();
    GPN_CK(cudaStreamSynchronize(c->stream));
    return 0;
}
long long gpn_launch_count(gpn_ctx* c) { return c->launches; }
double gpn_flop_count(gpn_ctx* c) { return c->flops; }
void gpn_reset_counters(gpn_ctx* c) {
    c->launches = 0;
    c->flops = 0.0;
}

int gpn_set_graph(gpn_ctx* c, int N, const int* indptr_h, const int* indices_h, const double* weights_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (N <= 0) return -2;
    if (!indptr_h) return -3;
    GPN_CK(cudaSetDevice(c->device));
    const long long nnz = indptr_h[N];
    if (nnz < 0) return -3;
    if (nnz > 0 && !indices_h) return -4;
    for (long long e = 0; e < nnz; ++e)
        if (indices_h[e] < 0 || indices_h[e] >= N) return -4;
    GPN_TRY(gpn_buf_ensure(c, c->indptr, sizeof(int) * (size_t)(N + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->indices, sizeof(int) * (size_t)(nnz + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->dinv, sizeof(double) * (size_t)(2 * N)));
    GPN_TRY(upload(c, c->indptr.p, indptr_h, sizeof(int) * (size_t)(N + 1)));
    if (nnz > 0) GPN_TRY(upload(c, c->indices.p, indices_h, sizeof(int) * (size_t)nnz));
    if (weights_h && nnz > 0) {
        GPN_TRY(gpn_buf_ensure(c, c->weights, sizeof(double) * (size_t)nnz));
        GPN_TRY(upload(c, c->weights.p, weights_h, sizeof(double) * (size_t)nnz));
    } else if (c->weights.p) {
        GPN_CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->weights.p);
        c->weights = DevBuf();
    }
    c->N = N;
    c->nnz = nnz;
    c->ldN = pad128(N);
    c->have_full = false;
    c->n_train = c->n_test = 0;
    c->h_indptr.assign(indptr_h, indptr_h + N + 1);
    c->h_indices.assign(indices_h, indices_h + nnz);
    if (weights_h) c->h_weights.assign(weights_h, weights_h + nnz);
    else c->h_weights.clear();
    c->h_train.clear();
    c->perm_valid = false;
    c->d6.built = false;
    gpn_spec_invalidate(c, true);
    return gpn_k_degree_dinv(c, N, (const int*)c->indptr.p, (const int*)c->indices.p, (const double*)c->weights.p, (double*)c->dinv.p);
}

int gpn_set_nodes(gpn_ctx* c, int n_train, const int* train_h, int n_test, const int* test_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (n_train < 0 || n_train > c->N) return -2;
    if (n_test < 0 || n_test > c->N) return -4;
    for (int i = 0; i < n_train; ++i)
        if (train_h[i] < 0 || train_h[i] >= c->N) return -3;
    for (int i = 0; i < n_test; ++i)
        if (test_h[i] < 0 || test_h[i] >= c->N) return -5;
    GPN_TRY(gpn_buf_ensure(c, c->train_idx, sizeof(int) * (size_t)(n_train + 1)));
    GPN_TRY(gpn_buf_ensure(c, c->test_idx, sizeof(int) * (size_t)(n_test + 1)));
    if (n_train) GPN_TRY(upload(c, c->train_idx.p, train_h, sizeof(int) * (size_t)n_train));
    if (n_test) GPN_TRY(upload(c, c->test_idx.p, test_h, sizeof(int) * (size_t)n_test));
    c->n_train = n_train;
    c->n_test = n_test;
    c->h_train.assign(train_h, train_h + n_train);
    std::sort(c->h_train.begin(), c->h_train.end());
    c->train_unique = std::adjacent_find(c->h_train.begin(), c->h_train.end()) == c->h_train.end();
    c->perm_valid = false;
    c->d6.built = false;
    gpn_spec_invalidate(c, false);
    return ensure_vecs(c);
}

int gpn_fp64_peak(gpn_ctx* c, double seconds, double* tflops_h) {
    GPN_RANGE();
    if (!c) return -1;
    GPN_TRY(gpn_buf_ensure(c, c->partial, sizeof(double) * (size_t)c->num_sms * 512));
    cudaEvent_t e0, e1;
    GPN_CK(cudaEventCreate(&e0));
    GPN_CK(cudaEventCreate(&e1));
    const int iters = 16384;
    const double flops_per_launch = 512.0 * 8 * iters * 16 * c->num_sms;
    dmma_peak_kernel<<<c->num_sms, 512, 0, c->stream>>>((double*)c->partial.p, 64, 1.0000001, 0.9999999);   // warm-up
    GPN_CK(cudaStreamSynchronize(c->stream));
    GPN_CK(cudaEventRecord(e0, c->stream));
    int reps = 0;
    float ms = 0.f;
    do {
        for (int r = 0; r < 4; ++r) dmma_peak_kernel<<<c->num_sms, 512, 0, c->stream>>>((double*)c->partial.p, iters, 1.0000001, 0.9999999);
        reps += 4;
        GPN_CK(cudaEventRecord(e1, c->stream));
        GPN_CK(cudaEventSynchronize(e1));
        GPN_CK(cudaEventElapsedTime(&ms, e0, e1));
    } while (ms < seconds * 1e3);
    *tflops_h = flops_per_launch * reps / (ms * 1e-3) * 1e-12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

void gpn_debug_leaf_prof(gpn_ctx* c, void* dev_buf32) {
    if (c) c->leaf_prof = dev_buf32;
}
int gpn_debug_leaf_bench(gpn_ctx* c, void* A_dev, int copies, void* prof32_dev) {
    GPN_RANGE();
    if (!c) return -1;
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)copies * 128 * 128));
    GPN_TRY(clear_info(c));
    return gpn_leaf_bench(c, (double*)A_dev, copies, (double*)c->winv.p, iscal(c, 0), (long long*)prof32_dev);
}
void gpn_set_fast_paths(gpn_ctx* c, int level) {
    if (c) c->fast_reglap = level < 0 ? 0 : (level > 6 ? 6 : level);
}
void gpn_set_dissection_parts(gpn_ctx* c, int parts) {
    if (!c) return;
    c->d6_parts = parts < 0 ? 0 : (parts > 8 ? 8 : parts);
    c->d6.built = false;
}
int gpn_debug_scalars(gpn_ctx* c, double* out32_h) {
    if (!c || !c->pinned) return -1;
    memcpy(out32_h, c->pinned, sizeof(double) * 32);
    return 0;
}
/* development aid: layout of the whole-graph dissection: out40 = [K, s, so, ND, off[9], nk[8], ok[8], c0[8]] */
int gpn_debug_d6_layout(gpn_ctx* c, int* out40) {
    if (!c || !c->d6.built) return -1;
    const Dissect6& q = c->d6;
    int p = 0;
    out40[p++] = q.K; out40[p++] = q.s; out40[p++] = q.so; out40[p++] = q.ND;
    for (int k = 0; k < 9; ++k) out40[p++] = q.off[k];      // (the first eight parts only: development aid)
    for (int k = 0; k < 8; ++k) out40[p++] = q.nk[k];
    for (int k = 0; k < 8; ++k) out40[p++] = q.ok[k];
    for (int k = 0; k < 8; ++k) out40[p++] = q.c0[k];
    return 0;
}
/* development aid: device pointer and capacity (bytes) of work buffer `which` of the whole-graph dissection route */
void* gpn_debug_d6_buffer(gpn_ctx* c, int which, long long* bytes) {
    if (!c || which < 0 || which >= 12) return nullptr;
    if (bytes) *bytes = (long long)c->d6.work[which].cap;
    return c->d6.work[which].p;
}
void gpn_debug_df_trace(gpn_ctx* c, void* dev_buf) {
    if (c) c->df_trace = dev_buf;
}

void gpn_set_profile(gpn_ctx* c, int on) {
    if (!c) return;
    c->profile = on != 0;
    if (!on) return;
    for (auto& p : c->prof_events) {
        cudaEventDestroy(p.e0);
        cudaEventDestroy(p.e1);
    }
    c->prof_events.clear();
}

/* out_h[0..2] = summed duration (ms), executed flops and count of the GEMM launches recorded since gpn_set_profile(ctx,1);
 * out_h[3..5] = the same restricted to the 128x128 GEMM configuration; out_h[6..8] = the same for the dataflow Cholesky;
 * out_h[9] / out_h[10] = length (ms) of the UNION of the launch intervals of the dataflow Cholesky / of all GEMM launches:
 * launches of one kernel that run concurrently on several streams share the chip, so chip-level throughput is flops / union. */
int gpn_profile_summary(gpn_ctx* c, double* out_h) {
    GPN_RANGE();
    if (!c) return -1;
    GPN_CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 11; ++i) out_h[i] = 0.0;
    if (c->prof_events.empty()) return 0;
    std::vector<std::pair<float, float>> iv[2];
    cudaEvent_t ref = c->prof_events.front().e0;
    for (auto& p : c->prof_events) {
        GPN_CK(cudaEventSynchronize(p.e1));
        float ms = 0.f, t0 = 0.f, t1 = 0.f;
        GPN_CK(cudaEventElapsedTime(&ms, p.e0, p.e1));
        GPN_CK(cudaEventElapsedTime(&t0, ref, p.e0));
        GPN_CK(cudaEventElapsedTime(&t1, ref, p.e1));
        iv[p.kind == 2 ? 0 : 1].push_back({t0, t1});
        if (p.kind == 2) { out_h[6] += ms; out_h[7] += p.flops; out_h[8] += 1; continue; }
        out_h[0] += ms; out_h[1] += p.flops; out_h[2] += 1;
        if (p.big) { out_h[3] += ms; out_h[4] += p.flops; out_h[5] += 1; }
    }
    for (int f = 0; f < 2; ++f) {
        std::sort(iv[f].begin(), iv[f].end());
        double total = 0.0;
        float cur0 = 0.f, cur1 = -1.f;
        for (auto& x : iv[f]) {
            if (cur1 < cur0 || x.first > cur1) {
                if (cur1 >= cur0) total += cur1 - cur0;
                cur0 = x.first;
                cur1 = x.second;
            } else if (x.second > cur1) {
                cur1 = x.second;
            }
        }
        if (cur1 >= cur0) total += cur1 - cur0;
        out_h[9 + f] = total;
    }
    return 0;
}

/* Launch intervals recorded since gpn_set_profile(ctx,1), for merging ACROSS contexts (lanes that evaluate concurrently): row i of
 * out_h = (start ms, end ms, executed flops, kind: 0 GEMM / 2 dataflow Cholesky), times relative to the first recorded launch of
 * `ref` (a context on the same device; NULL: this one).  *rows_h = rows written (at most max_rows). */
int gpn_profile_intervals(gpn_ctx* c, gpn_ctx* ref, double* out_h, int max_rows, int* rows_h) {
    GPN_RANGE();
    if (!c || !rows_h) return -1;
    *rows_h = 0;
    if (!ref) ref = c;
    GPN_CK(cudaStreamSynchronize(c->stream));
    if (c->prof_events.empty() || ref->prof_events.empty()) return 0;
    cudaEvent_t e_ref = ref->prof_events.front().e0;
    GPN_CK(cudaEventSynchronize(e_ref));
    int r = 0;
    for (auto& p : c->prof_events) {
        if (r >= max_rows) break;
        GPN_CK(cudaEventSynchronize(p.e1));
        float t0 = 0.f, t1 = 0.f;
        GPN_CK(cudaEventElapsedTime(&t0, e_ref, p.e0));
        GPN_CK(cudaEventElapsedTime(&t1, e_ref, p.e1));
        out_h[4 * r + 0] = t0; out_h[4 * r + 1] = t1; out_h[4 * r + 2] = p.flops; out_h[4 * r + 3] = p.kind;
        ++r;
    }
    *rows_h = r;
    return 0;
}

int gpn_set_targets(gpn_ctx* c, const double* t_h, int n) {
    GPN_RANGE();
    if (!c) return -1;
    if (n != c->n_train || n <= 0) return -3;
    GPN_TRY(ensure_vecs(c));
    return upload(c, vecp(c, V_TRES), t_h, sizeof(double) * n);
}

int gpn_laplacian_dense(gpn_ctx* c, double alpha, double gamma, double* out, long ld) {
    GPN_RANGE();
    if (!c || c->N <= 0) return -1;
    if (ld < c->N || (ld & 1)) return -5;
    GPN_TRY(gpn_k_fill_zero(c, out, ld, c->N, c->N));
    return gpn_k_laplacian_scatter(c, c->N, (const int*)c->indptr.p, (const int*)c->indices.p,
                                   c->weights.p ? (const double*)c->weights.p : nullptr, (const double*)c->dinv.p, alpha, gamma, out,
                                   ld);
}

int gpn_gemm_f64(gpn_ctx* c, int transa, int transb, int M, int N, int K, double alpha, const double* A, long lda, const double* B,
                 long ldb, double beta, double* C, long ldc) {
    GPN_RANGE();
    if (!c) return -1;
    return gpn_gemm(c, M, N, K, alpha, A, lda, transa == 0, B, ldb, transb != 0, beta, C, ldc, 0);
}

int gpn_potrf_f64(gpn_ctx* c, int n, double* A, long lda, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    GPN_TRY(clear_info(c));
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(n) * 128));
    GPN_TRY(gpn_potrf(c, n, A, lda, (double*)c->winv.p, iscal(c, 0)));
    GPN_TRY(gpn_k_zero_upper(c, A, lda, n));
    c->winv_for = A;             // gpn_trsm_f64 / gpn_potri_f64 may follow
    c->winv_n = n;
    if (!info_h) return 0;       // asynchronous use: no read-back
    int info[16];
    GPN_TRY(read_scalars(c, nullptr, 0, info, 16));
    *info_h = info[0];
    return 0;
}

int gpn_syrk_f64(gpn_ctx* c, int n, int k, double alpha, const double* A, long lda, double beta, double* C, long ldc) {
    GPN_RANGE();
    if (!c) return -1;
    // C = alpha A A^T + beta C: lower tiles computed, mirrored into the upper ones (the result is the full symmetric matrix)
    return gpn_gemm(c, n, n, k, alpha, A, lda, true, A, lda, true, beta, C, ldc, GF_LOWER | GF_MIRROR);
}

int gpn_posv_f64(gpn_ctx* c, int n, double* A, long lda, double* B, long ldb, int brows, double* Winv_t, long ldw, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (n <= 0) return -2;
    if (B && brows <= 0) return -7;
    GPN_TRY(clear_info(c));
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(n) * 128));
    GPN_TRY(gpn_potrf_dataflow(c, n, A, lda, (double*)c->winv.p, iscal(c, 0), 0, B, ldb, brows, Winv_t, ldw));
    GPN_TRY(gpn_k_zero_upper(c, A, lda, n));
    c->winv_for = A;
    c->winv_n = n;
    if (!info_h) return 0;
    int info[16];
    GPN_TRY(read_scalars(c, nullptr, 0, info, 16));
    *info_h = info[0];
    return 0;
}

/* B <- B L^-T for the brows x n row-major B and the factor L left in A by the gpn_potrf_f64 / gpn_posv_f64 call that
 * immediately precedes on this context (the inverted diagonal blocks it produced are still resident). */
int gpn_trsm_f64(gpn_ctx* c, int n, const double* L, long ldl, double* B, long ldb, int brows) {
    GPN_RANGE();
    if (!c) return -1;
    if (n <= 0 || c->winv_for != (const void*)L || c->winv_n != n) return -2;
    if (brows <= 0) return -7;
    const void* keep = c->winv_for;
    int st = gpn_trsm(c, brows, n, B, ldb, L, ldl, (const double*)c->winv.p);
    c->winv_for = keep;
    return st;
}

/* out = (L L^T)^-1 (full symmetric) from the factor L of the immediately preceding gpn_potrf_f64 / gpn_posv_f64 (LAPACK potri). */
int gpn_potri_f64(gpn_ctx* c, int n, const double* L, long ldl, double* out, long ldo) {
    GPN_RANGE();
    if (!c) return -1;
    if (n <= 0 || c->winv_for != (const void*)L || c->winv_n != n) return -2;
    const long ldw = pad128(n);
    GPN_TRY(ensure_nb(c, NB_W, ldw, ldw));
    GPN_TRY(ensure_nb(c, NB_WT, ldw, ldw));
    GPN_TRY(ensure_nb(c, NB_T, ldw, ldw));
    GPN_TRY(gpn_trtri(c, n, L, ldl, (const double*)c->winv.p, nbp(c, NB_W), nbp(c, NB_WT), nbp(c, NB_T), ldw));
    return gpn_lauum_t(c, n, nbp(c, NB_WT), ldw, out, ldo, true);
}


/* out = expm(A) for a symmetric negative semi-definite N x N device matrix A (scaling-and-squaring Pade, GPnet.py:340);
 * m_used / s_used (host, may be NULL) receive the Pade order and the number of squarings. */
int gpn_expm_sym_f64(gpn_ctx* c, int N, const double* A, long lda, double* out, long ldo, int* m_used, int* s_used) {
    GPN_RANGE();
    if (!c) return -1;
    if (N <= 0) return -2;
    const int st = with_matrix_size(c, N, [&]() -> int {
        for (int i = 0; i < 8; ++i) GPN_TRY(ensure_mat(c, i));
        GPN_TRY(clear_info(c));
        GPN_TRY(gpn_k_fill_zero(c, mat(c, 0), c->ldN, (int)c->ldN, (int)c->ldN));
        GPN_TRY(gpn_k_copy(c, N, N, A, lda, mat(c, 0), c->ldN));
        GPN_TRY(expm_of_mat0(c));
        GPN_TRY(gpn_k_copy(c, N, N, kfull(c), c->ldN, out, ldo));
        return gpn_sync(c);
    });
    if (m_used) *m_used = c->expm_m;
    if (s_used) *s_used = c->expm_s;
    return st;
}

/* out = B^p for a symmetric N x N device matrix B (numpy's binary powering order, GPnet.py:346); out_pm1 (may be NULL)
 * receives B^(p-1), the by-product the p-step derivative needs. */
int gpn_matpow_f64(gpn_ctx* c, int N, const double* B, long ldb, int p, double* out, long ldo, double* out_pm1, long ldo1) {
    GPN_RANGE();
    if (!c) return -1;
    if (N <= 0) return -2;
    if (p < 0) return -5;
    return with_matrix_size(c, N, [&]() -> int {
        for (int i = 0; i < 7; ++i) GPN_TRY(ensure_mat(c, i));
        const long ld = c->ldN;
        GPN_TRY(gpn_k_fill_zero(c, mat(c, 0), ld, (int)ld, (int)ld));
        GPN_TRY(gpn_k_copy(c, N, N, B, ldb, mat(c, 0), ld));
        auto identity = [&](double* dst, long ldd) -> int {
            GPN_TRY(gpn_k_fill_zero(c, mat(c, 5), ld, (int)ld, (int)ld));
            const double* X[1] = {mat(c, 0)};
            const long l1[1] = {ld};
            const double cf[1] = {0.0};
            GPN_TRY(gpn_k_lincomb(c, N, mat(c, 5), ld, 1.0, 1, X, l1, cf));
            return gpn_k_copy(c, N, N, mat(c, 5), ld, dst, ldd);
        };
        const double* res = nullptr;
        if (p == 0) {
            GPN_TRY(identity(out, ldo));
            if (out_pm1) return -8;                               // B^-1 is not a power
            return gpn_sync(c);
        }
        if (out_pm1 && p >= 2) {
            GPN_TRY(matpow(c, mat(c, 0), p - 1, mat(c, 1), mat(c, 2), mat(c, 3), mat(c, 6), &res));
            if (res != mat(c, 4)) GPN_TRY(gpn_k_copy(c, N, N, res, ld, mat(c, 4), ld));
            GPN_TRY(symm_mul(c, mat(c, 4), mat(c, 0), mat(c, 5)));
            GPN_TRY(gpn_k_copy(c, N, N, mat(c, 4), ld, out_pm1, ldo1));
            GPN_TRY(gpn_k_copy(c, N, N, mat(c, 5), ld, out, ldo));
            return gpn_sync(c);
        }
        GPN_TRY(matpow(c, mat(c, 0), p, mat(c, 1), mat(c, 2), mat(c, 3), mat(c, 6), &res));
        GPN_TRY(gpn_k_copy(c, N, N, res, ld, out, ldo));
        if (out_pm1) GPN_TRY(identity(out_pm1, ldo1));           // p == 1: B^0
        return gpn_sync(c);
    });
}

/* Route the last gpn_gpr_logp_grad / landscape evaluation took (see gpn_set_fast_paths; -1: none yet).  parts_h (may be NULL,
 * room for 16 ints): [number of parts of the node dissection used (0: none), separator size, part sizes...]. */
int gpn_last_route(gpn_ctx* c, int* parts_h) {
    if (!c) return -1;
    if (parts_h) {
        for (int i = 0; i < 16; ++i) parts_h[i] = 0;
        if (c->last_route == 5) {
            parts_h[0] = 2; parts_h[1] = c->dis_s; parts_h[2] = c->dis_a; parts_h[3] = c->dis_b;
        } else if (c->last_route == 6) {
            parts_h[0] = c->d6.K; parts_h[1] = c->d6.s;
            for (int k = 0; k < c->d6.K; ++k) parts_h[2 + k] = c->d6.nk[k];
        }
    }
    return c->last_route;
}

long long gpn_workspace_bytes(int N, int n, int m, int kind) {
    // what the context allocates on the device for one model of this size (see ensure_mat / ensure_nb in this file):
    // N x N operands (padded to 128): 4 for the regularized Laplacian and the Cholesky-based inverses, 6 for the p-step powers,
    // 10 for the Pade matrix exponential; block buffers K_y, W, T, W^T, G, a copy (n x n), two row panels (n x N), kstar
    // (m x n), kstarstar, V (m x m), Newton scratch (3 n x n); inverted diagonal blocks; vectors
    if (N <= 0 || n < 0 || m < 0) return -1;
    const long long Np = pad128(N), np_ = pad128(n > 0 ? n : 1), mp = pad128(m > 0 ? m : 1);
    const long long mats = kind == GPN_KERNEL_DIFFUSION ? 10 : (kind == GPN_KERNEL_PSTEP_WALK ? 6 : 4);
    long long d = mats * Np * Np;
    d += 6 * np_ * np_ + 2 * np_ * Np + mp * np_ + 2 * mp * mp + 3 * np_ * np_;
    d += Np * 128 + 24 * (std::max(np_, mp) + 128) + 2 * Np;
    return 8 * d;
}

int gpn_spd_inverse_f64(gpn_ctx* c, int n, double* A, long lda, double* out, long ldo, double* work, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (lda != ldo) return -6;
    GPN_TRY(clear_info(c));
    GPN_TRY(gpn_winv_ensure(c, sizeof(double) * (size_t)pad128(n) * 128));
    GPN_TRY(gpn_potrf(c, n, A, lda, (double*)c->winv.p, iscal(c, 0)));
    // W -> out, W^T -> internal buffer, T^T -> work; then K = W^T W overwrites out (W itself is not an operand any more)
    GPN_TRY(ensure_nb(c, NB_WT, n, ldo));
    GPN_TRY(gpn_trtri(c, n, A, lda, (const double*)c->winv.p, out, nbp(c, NB_WT), work, ldo));
    GPN_TRY(gpn_lauum_t(c, n, nbp(c, NB_WT), ldo, out, ldo, true));
    int info[16];
    GPN_TRY(read_scalars(c, nullptr, 0, info, 16));
    if (info_h) *info_h = info[0];
    return 0;
}

int gpn_gather_block_add(gpn_ctx* c, const double* K, long ld, const int* rows, int nr, const int* cols, int nc, double addc,
                         double* out, long ldo) {
    GPN_RANGE();
    if (!c) return -1;
    return gpn_k_gather(c, K, ld, rows, nr, cols, nc, addc, out, ldo, nr, nc);
}

int gpn_kernel_build(gpn_ctx* c, int kind, const double* theta_h, int P, int want_grad) {
    GPN_RANGE();
    if (!c) return -1;
    return kernel_build(c, kind, theta_h, P, want_grad);
}
const double* gpn_kernel_ptr(gpn_ctx* c, long* ld) {
    if (!c || !c->have_full) return nullptr;
    if (ld) *ld = c->ldN;
    return kfull(c);
}
void gpn_expm_info(gpn_ctx* c, int* m, int* s) {
    if (m) *m = c->expm_m;
    if (s) *s = c->expm_s;
}

int gpn_kernel_block_h(gpn_ctx* c, const int* rows_h, int nr, const int* cols_h, int nc, double addc, double* out_h) {
    GPN_RANGE();
    if (!c || !c->have_full) return -1;
    if (nr <= 0 || nc <= 0) return 0;
    for (int i = 0; i < nr; ++i)
        if (rows_h[i] < 0 || rows_h[i] >= c->N) return -2;      // the reference's np.delete / indexing raises IndexError
    for (int i = 0; i < nc; ++i)
        if (cols_h[i] < 0 || cols_h[i] >= c->N) return -4;
    DevBuf& ib = c->partial;
    const long ldo = (nc + 1) & ~1L;
    GPN_TRY(gpn_buf_ensure(c, ib, sizeof(int) * (size_t)(nr + nc + 2)));
    GPN_TRY(ensure_nb(c, NB_S, nr, ldo));
    int* dr = (int*)ib.p;
    int* dc = dr + nr;
    GPN_TRY(upload(c, dr, rows_h, sizeof(int) * (size_t)nr));
    GPN_TRY(upload(c, dc, cols_h, sizeof(int) * (size_t)nc));
    GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, dr, nr, dc, nc, addc, nbp(c, NB_S), ldo, nr, nc));
    GPN_TRY(download_matrix(c, out_h, nbp(c, NB_S), ldo, nr, nc));
    return gpn_sync(c);
}

int gpn_gpr_logp_grad(gpn_ctx* c, int kind, const double* theta_h, int P, const double* t_h, int want_grad, double* out_h, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    return gpr_logp_grad(c, kind, theta_h, P, t_h, want_grad, out_h, info_h);
}

/* Asynchronous form of gpn_gpr_logp_grad: _begin enqueues the evaluation on this context's streams and returns; _end waits for
 * it and delivers the result.  On the whole-graph dissection route (level 6) nothing synchronises in between, so several
 * contexts bound to the same graph ("lanes", each with an SM budget) evaluate different thetas concurrently; every other
 * route runs to completion inside _begin. */
int gpn_gpr_logp_grad_begin(gpn_ctx* c, int kind, const double* theta_h, int P, const double* t_h, int want_grad) {
    GPN_RANGE();
    if (!c) return -1;
    if (c->n_train <= 0) return -1;
    if (P < 1 || P > 8) return -4;
    c->pend_state = 0;
    c->pend_P = P;
    if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN && c->fast_reglap >= 6 && c->train_unique && c->n_train < c->N && d6_route(c)) {
        int st = gpr_reglap_d6_enqueue(c, theta_h, P, t_h, want_grad);
        if (st != 0) return st;
        c->pend_state = 1;
        return 0;
    }
    c->pend_status = gpr_logp_grad(c, kind, theta_h, P, t_h, want_grad, c->pend_out, &c->pend_info);
    c->pend_state = 2;
    c->pend_grad = want_grad;
    return c->pend_status;
}
int gpn_gpr_logp_grad_end(gpn_ctx* c, double* out_h, int* info_h) {
    GPN_RANGE();
    if (!c || c->pend_state == 0) return -1;
    const int P = c->pend_P;
    if (c->pend_state == 1) {
        RowScalars rs;
        int info = 0;
        c->pend_state = 0;
        GPN_TRY(gpr_reglap_d6_finish(c, &rs, &info));
        row_eval(rs, c->theta_exp[P - 1], P, c->pend_grad, out_h);
        if (info_h) *info_h = info;
        return 0;
    }
    c->pend_state = 0;
    for (int j = 0; j < 1 + (c->pend_grad ? P : 0); ++j) out_h[j] = c->pend_out[j];
    if (info_h) *info_h = c->pend_info;
    return c->pend_status;
}
/* SMs this context may fill with co-resident factorisations (0: the whole chip).  Lanes that run concurrently must not
 * oversubscribe the chip: cooperative launches that do not fit wait for the others to finish. */
void gpn_set_sm_budget(gpn_ctx* c, int sms) {
    if (c) c->sm_budget = sms < 0 ? 0 : sms;
}
int gpn_num_sms(gpn_ctx* c) { return c ? c->num_sms : 0; }

int gpn_gpr_predict(gpn_ctx* c, int kind, const double* theta_h, int P, const double* t_h, double* alpha_h, double* fstar_h, double* s_h,
                    double* kss_diag_h, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    return gpr_predict(c, kind, theta_h, P, t_h, alpha_h, fstar_h, s_h, kss_diag_h, info_h);
}

int gpn_gpr_fetch(gpn_ctx* c, int which, double* out_h) {
    GPN_RANGE();
    if (!c || !c->have_full) return -1;
    const int n = c->n_train, m = c->n_test;
    const long ldn = pad128(n), ldm = pad128(m);
    switch (which) {
        case 0: GPN_TRY(download_matrix(c, out_h, nbp(c, NB_KC), ldn, n, n)); break;
        case 1: GPN_TRY(download_matrix(c, out_h, nbp(c, NB_KSTAR), ldn, m, n)); break;
        case 2:
            GPN_TRY(gpn_k_zero_upper(c, nbp(c, NB_KY), ldn, n));
            GPN_TRY(download_matrix(c, out_h, nbp(c, NB_KY), ldn, n, n));
            break;
        case 3: GPN_TRY(download_matrix(c, out_h, nbp(c, NB_V), ldm, n, m)); break;
        case 4: {
            const int* te = (const int*)c->test_idx.p;
            GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, te, m, te, m, c->theta_exp[c->P - 1], nbp(c, NB_KSS), ldm, m, m));
            GPN_TRY(download_matrix(c, out_h, nbp(c, NB_KSS), ldm, m, m));
            break;
        }
        default: return -2;
    }
    return gpn_sync(c);
}

int gpn_gpc_newton(gpn_ctx* c, int kind, const double* theta_h, int P, const double* y_h, int n, const double* K_dev, long ldk, double tol,
                   double phif0, double scale, int variant, double* f_h, double* a_h, double* logq_h, int* iters_h, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    const double* K = K_dev;
    if (K == nullptr) {
        int st = kernel_build(c, kind, theta_h, P, 0);
        if (st != 0) return st;
        if (n != c->n_train) return -6;
        GPN_TRY(gpc_train_block(c));
        K = nbp(c, NB_KY);
        ldk = pad128(n);
    } else {
        // external K: vectors are sized from the node sets, make sure they can hold n
        if (c->n_train < n) c->n_train = n;
        GPN_TRY(clear_info(c));
    }
    return gpc_newton(c, K, ldk, n, y_h, tol, phif0, scale, variant, f_h, a_h, logq_h, iters_h, info_h);
}

int gpn_gpc_predict(gpn_ctx* c, int kind, const double* theta_h, int P, const double* y_h, double* f_h, double* a_h, double* logq_h,
                    int* iters_h, double* fstar_h, double* V_h, double* pistar_h, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    int st = kernel_build(c, kind, theta_h, P, 0);
    if (st != 0) return st;
    const int n = c->n_train, m = c->n_test;
    const long ldn = pad128(n), ldm = pad128(m);
    GPN_TRY(gpc_train_block(c));
    const double* K = nbp(c, NB_KY);
    GPN_TRY(gpc_newton(c, K, ldn, n, y_h, 0.1, 1e100, 1.0, 0, f_h, a_h, logq_h, iters_h, info_h));   // :200-202 defaults
    if (*info_h != 0 || m <= 0) return 0;
    NewtonState s;
    s.n = n; s.ldn = ldn; s.K = K; s.ldk = ldn; s.have_wk = false;
    // W, p from the FINAL f (:203-209); B refactored (:208)
    newton_pre_kernel<<<g1(n), 256, 0, c->stream>>>(n, vecp(c, V_F), vecp(c, V_Y), vecp(c, V_W), vecp(c, V_SW), vecp(c, V_P), vecp(c, V_B),
                                                     vecp(c, V_G));
    c->launches++;
    const int* tr = (const int*)c->train_idx.p;
    const int* te = (const int*)c->test_idx.p;
    // kstar^T = K[te, tr] (m x n, measnoise = 0, :193-199) with its rows padded to a multiple of 128 (appended block of the
    // factorisation below); fstar = kstar^T g (:210)
    GPN_TRY(ensure_nb(c, NB_KSTAR, ldm, ldn));
    double* KsT = nbp(c, NB_KSTAR);
    GPN_TRY(gpn_k_gather(c, kfull(c), c->ldN, te, m, tr, n, 0.0, KsT, ldn));
    GPN_TRY(gpn_k_gemv(c, m, n, KsT, ldn, vecp(c, V_G), vecp(c, V_FSTAR), GEMV_FULL));
    // v = L_B^-1 (sqrtW kstar) (:211): its transpose (kstar^T sqrtW) L_B^-T is the appended right-hand-side block of the
    // SAME dataflow kernel that factors B (:208) -- no triangular inverse, no separate product; ||v_i||^2 = row norms
    GPN_TRY(gpn_k_scale_cols(c, m, n, KsT, ldn, vecp(c, V_SW)));
    GPN_TRY(factor_B_fast(c, s, KsT, m));
    GPN_TRY(gpn_k_rownorm2(c, m, n, KsT, ldn, vecp(c, V_VN)));
    GPN_TRY(gpn_k_diag_gather(c, kfull(c), c->ldN, te, m, 0.0, vecp(c, V_KSS)));                          // :212-214
    pistar_kernel<<<g1(m), 256, 0, c->stream>>>(m, vecp(c, V_FSTAR), vecp(c, V_KSS), vecp(c, V_VN), vecp(c, V_VV), vecp(c, V_PI));
    c->launches++;
    GPN_TRY(download(c, fstar_h, vecp(c, V_FSTAR), sizeof(double) * m));
    GPN_TRY(download(c, V_h, vecp(c, V_VV), sizeof(double) * m));
    GPN_TRY(download(c, pistar_h, vecp(c, V_PI), sizeof(double) * m));
    int info[16];
    GPN_TRY(read_scalars(c, nullptr, 0, info, 16));
    if (info[3] != 0) *info_h = info[3];
    return 0;
}

int gpn_gpc_grad(gpn_ctx* c, int kind, const double* theta_h, int P, const double* y_h, double* out_h, int* iters_h, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (c->n_train <= 0) return -1;
    return gpc_grad(c, kind, theta_h, P, y_h, out_h, iters_h, info_h);
}

int gpn_landscape_points(gpn_ctx* c, int kind, int classifier, const double* theta_h, int P, int ax0, int ax1, const double* ax1_h, int n1,
                         const double* ax2_h, int n2, const double* t_h, const long* idx_h, long count, int want_grad, double* out_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (P < 1 || P > 8) return -5;
    if (ax0 < 0 || ax0 >= P) return -6;
    if (ax1 < 0 || ax1 >= P) return -7;
    if (count < 0) return -14;
    const int rec = 1 + P;
    // Points whose theta differs only in the noise entry (the LAST one, added to every kernel entry: quirk A-3) share the
    // kernel; they are evaluated back to back so that it is built once per group.  The p-step kernel with P = 2 reads its
    // order from the same entry (A-4), and P = 1 makes beta the noise: no sharing there.
    const int last = P - 1;
    const bool share = c->fast_reglap >= 1 && P >= 2 && !(kind == GPN_KERNEL_PSTEP_WALK && P == 2) && (ax0 == last || ax1 == last);
    std::vector<long> order(count), key(count);
    for (long q = 0; q < count; ++q) {
        const long idx = idx_h[q];
        if (idx < 0 || idx >= (long)n1 * n2) return -13;
        order[q] = q;
        if (!share) key[q] = q;
        else if (ax1 == last) key[q] = (ax0 == last) ? 0 : idx / n2;        // the second axis is written last (GPnet.py:545-546)
        else key[q] = idx % n2;
    }
    std::stable_sort(order.begin(), order.end(), [&](long x, long y) { return key[x] < key[y]; });
    double th[8];
    RowScalars rs;
    long cur_key = -1;
    bool have_row = false, row_ok = false;
    for (long k = 0; k < count; ++k) {
        const long q = order[k], idx = idx_h[q];
        for (int j = 0; j < P; ++j) th[j] = theta_h[j];
        th[ax0] = ax1_h[idx / n2];          // GPnet.py:544-546 (ax0 written first, then ax1, as the reference)
        th[ax1] = ax2_h[idx % n2];
        double* o = out_h + q * rec;
        for (int j = 0; j < rec; ++j) o[j] = 0.0;
        const bool fresh = !share || !have_row || key[q] != cur_key;
        cur_key = key[q];
        int info = 0, st;
        double tmp[9];
        if (!classifier) {
            if (share) {
                if (fresh) {
                    st = gpr_row_scalars(c, kind, th, P, t_h, want_grad, &rs, &info);
                    if (st != 0) return st;
                    have_row = true;
                    row_ok = info == 0 && std::isfinite(rs.half_logdet) && std::isfinite(rs.s);
                }
                if (row_ok) {
                    row_eval(rs, th[last], P, want_grad, tmp);
                    o[0] = -tmp[0];
                    if (want_grad)
                        for (int j = 0; j < P; ++j) o[1 + j] = -tmp[1 + j];
                    continue;
                }
                have_row = false;           // K_tt not PD: factor K_y itself for this point (and rebuild the row state next time)
            }
            st = gpr_logp_grad(c, kind, th, P, t_h, want_grad, tmp, &info);
            if (st != 0) return st;
            if (info > 0) {
                o[0] = INFINITY;            // logPosterior returned -inf (A-8) => lml = +inf
            } else {
                o[0] = -tmp[0];
                if (want_grad)
                    for (int j = 0; j < P; ++j) o[1 + j] = -tmp[1 + j];
            }
        } else {
            double logq = 0.0;
            int iters = 0;
            if (fresh || !c->have_full) {
                st = kernel_build(c, kind, th, P, 0);
                if (st != 0) return st;
                have_row = true;
            } else {
                c->theta_exp[last] = th[last];          // same kernel, new noise
                GPN_TRY(clear_info(c));
            }
            GPN_TRY(gpc_train_block(c));
            st = gpc_newton(c, nbp(c, NB_KY), pad128(c->n_train), c->n_train, t_h, 0.1, 1e100, 1.0, 0, nullptr, nullptr, &logq, &iters,
                            &info);
            if (st != 0) return st;
            o[0] = (info > 0) ? NAN : logq;
        }
    }
    return 0;
}

int gpn_landscape(gpn_ctx* c, int kind, int classifier, const double* theta_h, int P, int ax0, int ax1, const double* ax1_h, int n1,
                  const double* ax2_h, int n2, const double* t_h, long idx_begin, long idx_end, long idx_stride, int want_grad,
                  double* out_h) {
    GPN_RANGE();
    if (idx_stride <= 0) return -15;
    std::vector<long> idx;
    for (long i = idx_begin; i < idx_end && i < (long)n1 * n2; i += idx_stride) idx.push_back(i);
    return gpn_landscape_points(c, kind, classifier, theta_h, P, ax0, ax1, ax1_h, n1, ax2_h, n2, t_h, idx.data(), (long)idx.size(), want_grad,
                                out_h);
}

}   // extern "C"
