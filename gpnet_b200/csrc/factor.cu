// Cholesky / triangular inverse / SPD inverse built on the DMMA GEMM (gemm_f64.cu).
//
// Replaces np.linalg.cholesky + the np.linalg.solve(L, .) chains of the reference (GPnetRegressor.py:123-126,141-144;
// GPnetClassifier.py:110-121,208-211) and scipy.linalg.inv (GPnet.py:342).
//
//  * gpn_potrf   : recursive blocked Cholesky (lower, row-major, in place).  Leaves are 128x128 diagonal blocks
//                  factored AND inverted by one CTA in shared memory (potf2_inv_kernel); everything else is GEMM:
//                  TRSM is "multiply by the inverted diagonal block" (recursive), the trailing update is a
//                  lower-tiles-only SYRK.  The inverted diagonal blocks are kept (winv) for the triangular inverse.
//  * gpn_trtri_from_winv : W = L^-1 level by level; all pairs of one level run as ONE batched GEMM launch
//                  (W21 = -W22 * (L21 * W11)), triangular structure skipped at tile granularity.
//  * gpn_lauum   : W^T W as a single structure-aware GEMM (k >= max(m0,n0), lower tiles + mirrored store).
#include "common.cuh"

namespace {

constexpr int LB = 128;          // leaf block
constexpr int LDS_ = LB + 1;     // padded shared-memory leading dimension (conflict-free column access)

// One CTA (256 threads): factor the nv x nv (nv <= 128) diagonal block at A (lower Cholesky, written back, strict upper
// zeroed) and write inv(L) (identity-padded to 128x128, upper zero) to winv (ld 128).
__global__ void __launch_bounds__(256, 1) potf2_inv_kernel(double* __restrict__ A, long lda, int nv, double* __restrict__ winv,
                                                            int* __restrict__ info, int info_off) {
    extern __shared__ double sm[];
    double* S = sm;                       // [128][129]
    double* T = S + LB * LDS_;            // [96][33] temp block column
    double* rinv = T + 96 * 33;           // [128]
    __shared__ double sh_l[8][9];
    __shared__ double sh_piv;
    const int tid = threadIdx.x;

    for (int idx = tid; idx < LB * LB; idx += 256) {
        const int r = idx >> 7, c = idx & 127;
        double v = 0.0;
        if (c <= r) {
            if (r < nv) v = A[(long)r * lda + c];
            else if (r == c) v = 1.0;
        }
        S[r * LDS_ + c] = v;
    }
    __syncthreads();

    // ---------------- factorization: 8-column micro panels, right looking ----------------
    for (int kb = 0; kb < LB; kb += 8) {
        double x[8];
        const bool owner = (tid < LB) && (tid >= kb);
        if (owner) {
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = S[tid * LDS_ + kb + c];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (tid == kb + c) {
                double d = x[c];
                if (!(d > 0.0)) atomicCAS(info, 0, info_off + kb + c + 1);
                d = sqrt(d);
                x[c] = d;
                sh_piv = d;
            }
            __syncthreads();
            if (owner && tid > kb + c) {
                x[c] = x[c] / sh_piv;
                if (tid < kb + 8) sh_l[tid - kb][c] = x[c];
            }
            __syncthreads();
            if (owner && tid > kb + c) {
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2) x[c2] -= x[c] * sh_l[c2][c];
            }
        }
        if (owner) {
#pragma unroll
            for (int c = 0; c < 8; ++c) S[tid * LDS_ + kb + c] = (kb + c <= tid) ? x[c] : 0.0;
        }
        __syncthreads();
        // trailing update S[r][c'] -= sum_k S[r][kb+k] * S[c'][kb+k]   (r >= c' >= kb+8), 4x4 register tiles
        const int t0 = kb + 8;
        const int rem = LB - t0;
        const int nt = (rem + 3) >> 2;
        for (int ti = tid; ti < nt * nt; ti += 256) {
            const int tr = ti / nt, tc = ti - tr * nt;
            if (tc > tr) continue;
            const int r0 = t0 + 4 * tr, c0 = t0 + 4 * tc;
            double a[4][8], b[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a[i][k] = S[(r0 + i) * LDS_ + kb + k];
                    b[i][k] = S[(c0 + i) * LDS_ + kb + k];
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) s += a[i][k] * b[j][k];
                    if (c0 + j <= r0 + i) S[(r0 + i) * LDS_ + c0 + j] -= s;
                }
        }
        __syncthreads();
    }

    // ---------------- write L back (valid region; strict upper zeroed) ----------------
    for (int idx = tid; idx < LB * LB; idx += 256) {
        const int r = idx >> 7, c = idx & 127;
        if (r < nv && c < nv) A[(long)r * lda + c] = S[r * LDS_ + c];
    }
    if (tid < LB) rinv[tid] = 1.0 / S[tid * LDS_ + tid];
    __syncthreads();

    // ---------------- in-place inverse, 32x32 blocks ----------------
    // (1) diagonal blocks: warp b inverts block b, lane j owns column j (forward substitution in registers)
    {
        const int b = tid >> 5, j = tid & 31;
        double x[32];
        if (b < 4) {
            const double* Sb = S + (b * 32) * LDS_ + b * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k + 1 < i; k += 2) {
                    s0 -= Sb[i * LDS_ + k] * x[k];
                    s1 -= Sb[i * LDS_ + k + 1] * x[k + 1];
                }
                if (i & 1) s0 -= Sb[i * LDS_ + i - 1] * x[i - 1];
                x[i] = (s0 + s1) * rinv[b * 32 + i];
            }
        }
        __syncthreads();
        if (b < 4) {
            double* Sb = S + (b * 32) * LDS_ + b * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i >= j) Sb[i * LDS_ + j] = x[i];
        }
        __syncthreads();
    }
    // (2) block columns right to left: X = -W_trail * (L_col * W_kk)
    for (int k = 2; k >= 0; --k) {
        const int r0 = 32 * (k + 1);
        const int nr = LB - r0;
        const int j = tid & 31;
        for (int r = tid >> 5; r < nr; r += 8) {
            double s = 0.0;
            const double* Lr = S + (r0 + r) * LDS_ + 32 * k;
            const double* Wk = S + (32 * k) * LDS_ + 32 * k;
            for (int q = j; q < 32; ++q) s += Lr[q] * Wk[q * LDS_ + j];
            T[r * 33 + j] = s;
        }
        __syncthreads();
        for (int r = tid >> 5; r < nr; r += 8) {
            double s = 0.0;
            const double* Wr = S + (r0 + r) * LDS_ + r0;
            for (int q = 0; q <= r; ++q) s += Wr[q] * T[q * 33 + j];
            S[(r0 + r) * LDS_ + 32 * k + j] = -s;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < LB * LB; idx += 256) {
        const int r = idx >> 7, c = idx & 127;
        winv[idx] = (c <= r) ? S[r * LDS_ + c] : 0.0;
    }
}

constexpr size_t LEAF_SMEM = (size_t)(LB * LDS_ + 96 * 33 + LB) * sizeof(double);

int launch_leaf(gpn_ctx* c, double* A, long lda, int nv, double* winv, int* d_info, int info_off) {
    static bool configured = false;
    if (!configured) {
        GPN_CK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM));
        configured = true;
    }
    potf2_inv_kernel<<<1, 256, LEAF_SMEM, c->stream>>>(A, lda, nv, winv, d_info, info_off);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}

inline int split_point(int n) {   // first-half size: a multiple of LB, at least half
    int blocks = (n + LB - 1) / LB;
    return ((blocks + 1) / 2) * LB;
}

// X (m x k, ldx) <- X * L^-T, L (k x k lower, ldl) with inverted diagonal blocks in winv
int trsm_rec(gpn_ctx* c, int m, int k, double* X, long ldx, const double* L, long ldl, const double* winv) {
    if (m <= 0 || k <= 0) return 0;
    if (k <= LB) {
        // in place is safe: one N tile (BN = 128 >= k) => every CTA reads only its own rows of X before writing them
        return gpn_gemm(c, m, k, k, 1.0, X, ldx, true, winv, LB, true, 0.0, X, ldx, GF_FORCE_BIG);
    }
    const int k1 = split_point(k), k2 = k - k1;
    GPN_TRY(trsm_rec(c, m, k1, X, ldx, L, ldl, winv));
    // X2 -= X1 * Lb^T, Lb = L[k1:, :k1]
    GPN_TRY(gpn_gemm(c, m, k2, k1, -1.0, X, ldx, true, L + (long)k1 * ldl, ldl, true, 1.0, X + k1, ldx, 0));
    return trsm_rec(c, m, k2, X + k1, ldx, L + (long)k1 * ldl + k1, ldl, winv + (long)k1 * LB);
}

int potrf_rec(gpn_ctx* c, int n, double* A, long lda, double* winv, int* d_info, int off) {
    if (n <= LB) return launch_leaf(c, A, lda, n, winv, d_info, off);
    const int n1 = split_point(n), n2 = n - n1;
    GPN_TRY(potrf_rec(c, n1, A, lda, winv, d_info, off));
    double* A21 = A + (long)n1 * lda;
    double* A22 = A21 + n1;
    GPN_TRY(trsm_rec(c, n2, n1, A21, lda, A, lda, winv));
    GPN_TRY(gpn_gemm(c, n2, n2, n1, -1.0, A21, lda, true, A21, lda, true, 1.0, A22, lda, GF_LOWER));
    return potrf_rec(c, n2, A22, lda, winv + (long)n1 * LB, d_info, off + n1);
}

__global__ void copy_diag_blocks_kernel(const double* __restrict__ winv, double* __restrict__ W, long ldw, int n) {
    const int b = blockIdx.x;
    for (int idx = threadIdx.x; idx < LB * LB; idx += blockDim.x) {
        const int r = idx >> 7, cc = idx & 127;
        const int gr = b * LB + r, gc = b * LB + cc;
        if (gr < n && gc < n) W[(long)gr * ldw + gc] = winv[(long)b * LB * LB + idx];
    }
}

}   // namespace

int gpn_potrf(gpn_ctx* c, int n, double* A, long lda, double* winv, int* d_info, int info_offset) {
    if (n <= 0) return 0;
    return potrf_rec(c, n, A, lda, winv, d_info, info_offset);
}

// W (n x n, ldw; fully overwritten incl. zero upper triangle) = L^-1.  `tmp` is an n x n scratch with leading dim ldw.
int gpn_trtri_from_winv_tmp(gpn_ctx* c, int n, const double* L, long ldl, const double* winv, double* W, long ldw, double* tmp) {
    if (n <= 0) return 0;
    GPN_TRY(gpn_k_fill_zero(c, W, ldw, n, n));
    const int m = (n + LB - 1) / LB;
    copy_diag_blocks_kernel<<<m, 256, 0, c->stream>>>(winv, W, ldw, n);
    c->launches++;
    GPN_CK(cudaGetLastError());
    for (long b = LB; b < n; b *= 2) {       // b = half size of the pairs of this level
        const long pair = 2 * b;
        const int full = (int)(n / pair);    // pairs entirely inside the valid n x n region
        if (full > 0) {
            const long sL = pair * ldl + pair, sW = pair * ldw + pair;
            // T = L21 * W11       (W11 lower: only k >= n0 contributes)
            GPN_TRY(gpn_gemm(c, (int)b, (int)b, (int)b, 1.0, L + b * ldl, ldl, true, W, ldw, false, 0.0, tmp + b * ldw, ldw, GF_KLO_N,
                             full, sL, sW, sW));
            // W21 = -W22 * T      (W22 lower: only k < m0+BM contributes)
            GPN_TRY(gpn_gemm(c, (int)b, (int)b, (int)b, -1.0, W + b * ldw + b, ldw, true, tmp + b * ldw, ldw, false, 0.0, W + b * ldw,
                             ldw, GF_KHI_M, full, sW, sW, sW));
        }
        const long o = (long)full * pair;    // ragged remainder: [o, o+b) full first half, [o+b, n) partial second half
        const long r = n - o - b;
        if (r > 0) {
            const double* L21 = L + (o + b) * ldl + o;
            double* T21 = tmp + (o + b) * ldw + o;
            double* W21 = W + (o + b) * ldw + o;
            GPN_TRY(gpn_gemm(c, (int)r, (int)b, (int)b, 1.0, L21, ldl, true, W + o * ldw + o, ldw, false, 0.0, T21, ldw, GF_KLO_N));
            GPN_TRY(gpn_gemm(c, (int)r, (int)b, (int)r, -1.0, W + (o + b) * ldw + (o + b), ldw, true, T21, ldw, false, 0.0, W21, ldw,
                             GF_KHI_M));
        }
    }
    return 0;
}

// out (lower [+ mirrored upper]) = W^T W for lower-triangular W.
int gpn_lauum(gpn_ctx* c, int n, const double* W, long ldw, double* out, long ldo, bool mirror) {
    return gpn_gemm(c, n, n, n, 1.0, W, ldw, false, W, ldw, false, 0.0, out, ldo,
                    GF_LOWER | (mirror ? GF_MIRROR : 0) | GF_KLO_M | GF_KLO_N);
}
