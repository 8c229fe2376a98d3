// Cholesky / triangular inverse / SPD inverse built on the DMMA GEMM (gemm_f64.cu).
//
// Replaces np.linalg.cholesky + the np.linalg.solve(L, .) chains of the reference (GPnetRegressor.py:123-126,141-144;
// GPnetClassifier.py:110-121,208-211) and scipy.linalg.inv (GPnet.py:342).
//
//  * gpn_potrf   : n > 128: the single-kernel dataflow Cholesky (potrf_dataflow.cu).  n <= 128 (and the recursive multi-launch
//                  fallback selected by GPNET_B200_POTRF=recursive, kept for A/B timing): leaves are 128x128 diagonal blocks
//                  factored AND inverted by one CTA in shared memory (leaf.cuh); TRSM is "multiply by the inverted diagonal
//                  block", the trailing update a lower-tiles-only SYRK.  The inverted diagonal blocks are kept (winv).
//  * gpn_trtri   : W = L^-1 (and W^T) level by level; all pairs of one level run as ONE batched GEMM launch
//                  (W21 = -W22 * (L21 * W11)), triangular structure skipped at tile granularity, K-major operands only.
//  * gpn_lauum_t : W^T W as a single structure-aware GEMM (k >= max(m0,n0), lower tiles + mirrored store).
#include "common.cuh"
#include "leaf.cuh"

namespace {

constexpr int LB = leaf::LB;

// One CTA (256 threads): factor the nv x nv (nv <= 128) diagonal block at A (lower Cholesky, written back, strict upper
// zeroed) and write inv(L) (identity-padded to 128x128, upper zero) to winv (ld 128).  Body: leaf.cuh.
__global__ void __launch_bounds__(256, 1) potf2_inv_kernel(double* __restrict__ A, long lda, int nv, double* __restrict__ winv,
                                                            int* __restrict__ info, int info_off, long long* prof) {
    // blockIdx.x > 0 only in the timing harness (gpn_debug_leaf_bench): every block factors its own copy so that the leaf is
    // timed on a full grid (single-CTA launches of kernels larger than the instruction cache are not representative)
    A += (long)blockIdx.x * LB * lda;
    winv += (long)blockIdx.x * LB * LB;
    if (blockIdx.x != 0) prof = nullptr;
    extern __shared__ double sm[];
    double* S = sm + leaf::GUARD;
    double* Wd = S + LB * leaf::LDS_;
    double* T = Wd + 4 * 32 * leaf::WLD;
    const int tid = threadIdx.x;
    if (tid < leaf::GUARD) sm[tid] = 0.0;
    for (int idx = tid; idx < LB * LB; idx += 256) {
        const int r = idx >> 7, c = idx & 127;
        double v = 0.0;
        if (c <= r) {
            if (r < nv) v = A[(long)r * lda + c];
            else if (r == c) v = 1.0;
        }
        S[r * leaf::LDS_ + c] = v;
    }
    __syncthreads();
    if (prof && tid == 0) prof[31] = clock64();
    leaf::potf2_inverse<0>(S, Wd, T, A, lda, nv, winv, info, info_off, tid, prof);
}

constexpr size_t LEAF_SMEM = leaf::SMEM_BYTES;

int launch_leaf(gpn_ctx* c, double* A, long lda, int nv, double* winv, int* d_info, int info_off) {
    static bool configured = false;
    if (!configured) {
        GPN_CK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM));
        configured = true;
    }
    potf2_inv_kernel<<<1, 256, LEAF_SMEM, c->stream>>>(A, lda, nv, winv, d_info, info_off, (long long*)c->leaf_prof);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}

}   // namespace

// Timing harness: `copies` independent 128x128 blocks stacked in A (lda = 128), one CTA each; block 0 records clock64 stamps.
int gpn_leaf_bench(gpn_ctx* c, double* A, int copies, double* winv, int* d_info, long long* prof) {
    GPN_CK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM));
    potf2_inv_kernel<<<copies, 256, LEAF_SMEM, c->stream>>>(A, LB, LB, winv, d_info, 0, prof);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}

namespace {

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

__global__ void copy_diag_blocks_kernel(const double* __restrict__ winv, double* __restrict__ W, double* __restrict__ WT, long ldw, int n) {
    const int b = blockIdx.x;
    for (int idx = threadIdx.x; idx < LB * LB; idx += blockDim.x) {
        const int r = idx >> 7, cc = idx & 127;
        const int gr = b * LB + r, gc = b * LB + cc;
        if (gr < n && gc < n) {
            const double v = winv[(long)b * LB * LB + idx];
            W[(long)gr * ldw + gc] = v;
            WT[(long)gc * ldw + gr] = v;
        }
    }
}

}   // namespace

int gpn_trsm(gpn_ctx* c, int m, int k, double* X, long ldx, const double* L, long ldl, const double* winv) {
    return trsm_rec(c, m, k, X, ldx, L, ldl, winv);
}

int gpn_potrf(gpn_ctx* c, int n, double* A, long lda, double* winv, int* d_info, int info_offset) {
    c->winv_for = nullptr;
    if (n <= 0) return 0;
    if (n > LB && c->potrf_mode == 0) return gpn_potrf_dataflow(c, n, A, lda, winv, d_info, info_offset);
    return potrf_rec(c, n, A, lda, winv, d_info, info_offset);
}

// W = L^-1 (lower, upper part zeroed) and WT = W^T, level by level; every GEMM reads K-major operands only (the 99 %
// DMMA path): T^T is kept instead of T, and each W21 block is stored both ways by the GEMM epilogue.
int gpn_trtri(gpn_ctx* c, int n, const double* L, long ldl, const double* winv, double* W, double* WT, double* TT, long ldw) {
    if (n <= 0) return 0;
    // No zero fill of W / WT: every consumer restricts its k-range to the diagonal 128-block of the tile it reads (structure
    // flags, GEMV_LOWER), the diagonal blocks are written completely (with their zero halves) right here, and the strictly
    // upper (resp. lower) off-diagonal blocks are never read.
    const int m = (n + LB - 1) / LB;
    copy_diag_blocks_kernel<<<m, 256, 0, c->stream>>>(winv, W, WT, ldw, n);
    c->launches++;
    GPN_CK(cudaGetLastError());
    for (long b = LB; b < n; b *= 2) {       // b = half size of the pairs of this level
        const long pair = 2 * b;
        const int full = (int)(n / pair);    // pairs entirely inside the valid n x n region
        if (full > 0) {
            const long sL = pair * ldl + pair, sW = pair * ldw + pair;
            // T = L21 * W11 (W11 lower: only k >= n0 contributes), stored transposed: TT[o+c][o+b+r]
            GPN_TRY(gpn_gemm(c, (int)b, (int)b, (int)b, 1.0, L + b * ldl, ldl, true, WT, ldw, true, 0.0, nullptr, 0, GF_KLO_N, full, sL, sW, 0,
                             TT + b, ldw, sW));
            // W21 = -W22 * T (W22 lower: only k < m0+BM contributes); normal copy into W, transposed copy into WT
            GPN_TRY(gpn_gemm(c, (int)b, (int)b, (int)b, -1.0, W + b * ldw + b, ldw, true, TT + b, ldw, true, 0.0, W + b * ldw, ldw, GF_KHI_M,
                             full, sW, sW, sW, WT + b, ldw, sW));
        }
        const long o = (long)full * pair;    // ragged remainder: [o, o+b) full first half, [o+b, n) partial second half
        const long r = n - o - b;
        if (r > 0) {
            GPN_TRY(gpn_gemm(c, (int)r, (int)b, (int)b, 1.0, L + (o + b) * ldl + o, ldl, true, WT + o * ldw + o, ldw, true, 0.0, nullptr, 0,
                             GF_KLO_N, 1, 0, 0, 0, TT + o * ldw + (o + b), ldw, 0));
            GPN_TRY(gpn_gemm(c, (int)r, (int)b, (int)r, -1.0, W + (o + b) * ldw + (o + b), ldw, true, TT + o * ldw + (o + b), ldw, true, 0.0,
                             W + (o + b) * ldw + o, ldw, GF_KHI_M, 1, 0, 0, 0, WT + o * ldw + (o + b), ldw, 0));
        }
    }
    return 0;
}

// out (lower [+ mirrored upper]) = W^T W = sum_k WT[i][k] WT[j][k], k >= max(i, j)
int gpn_lauum_t(gpn_ctx* c, int n, const double* WT, long ldw, double* out, long ldo, bool mirror) {
    return gpn_gemm(c, n, n, n, 1.0, WT, ldw, true, WT, ldw, true, 0.0, out, ldo, GF_LOWER | (mirror ? GF_MIRROR : 0) | GF_KLO_M | GF_KLO_N);
}
