// 128x128 Cholesky leaf: factor AND invert one diagonal block in shared memory with 256 threads (8 warps).
// Used by the stand-alone leaf kernel (factor.cu) and inside the dataflow Cholesky (potrf_dataflow.cu).
//
//  * 32x32 diagonal sub-blocks: ONE warp, rows in registers, column sweep with warp shuffles (no block barrier on
//    the per-column critical chain), followed by the register-resident triangular inverse of the sub-block;
//  * panel (L21 = A21 W11^T) and trailing update (A22 -= L21 L21^T) and the off-diagonal blocks of the inverse:
//    4x4 register-tiled shared-memory GEMMs over all 8 warps.
// Shared memory: S[128][129] (padded, conflict-free column access), Wd[4][32][33], T[96][33].
#pragma once
#include "common.cuh"

namespace leaf {

constexpr int LB = 128;
constexpr int LDS_ = LB + 1;
constexpr int WLD = 33;
constexpr size_t SMEM_DOUBLES = (size_t)LB * LDS_ + 4 * 32 * WLD + 96 * WLD + LB;   // S, Wd, T, reciprocal diagonal
constexpr size_t SMEM_BYTES = SMEM_DOUBLES * sizeof(double);

template <int BAR_ID>
__device__ __forceinline__ void bar256() {
    if (BAR_ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, 256;" ::"n"(BAR_ID) : "memory");
}

// acc[i][j] = sum_{k < klim(tile)} A[(r0+i)*sAr + k] * B[(c0+j)*sBc + k*sBk] for the 4x4 tile (r0, c0).
// KTRI: the A operand is lower triangular in (row, k) => k <= r0+3 suffices.
template <bool KTRI>
__device__ __forceinline__ void tile4x4(const double* __restrict__ A, int sAr, const double* __restrict__ B, int sBc, int sBk, int K,
                                        int r0, int c0, double (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int klim = KTRI ? min(K, r0 + 4) : K;
    const double* a0 = A + r0 * sAr;
    const double* b0 = B + c0 * sBc;
#pragma unroll 4
    for (int k = 0; k < klim; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = a0[i * sAr + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = b0[j * sBc + k * sBk];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
}

// Loop over the 4x4 tiles of an R x Cn output (R multiple of 32, Cn multiple of 16): warp-level super tiles of
// 32 rows x 16 cols, lane -> (tile row lane>>2, tile col lane&3).  f(r0, c0) is called for every tile this thread owns.
template <typename F>
__device__ __forceinline__ void for_tiles(int R, int Cn, int tid, bool lower_only, F f) {
    const int warp = tid >> 5, lane = tid & 31;
    const int superC = Cn >> 4, superR = R >> 5;
    if (!lower_only) {
        for (int st = warp; st < superR * superC; st += 8) {
            const int sr = st / superC, sc = st - sr * superC;
            f((sr * 8 + (lane >> 2)) * 4, (sc * 4 + (lane & 3)) * 4);
        }
    } else {
        // super-tile row sr owns 2(sr+1) super tiles that touch the lower triangle: enumerate only those
        const int total = superR * (superR + 1);
        for (int st = warp; st < total; st += 8) {
            int sr = 0;
            while ((sr + 1) * (sr + 2) <= st) ++sr;
            const int sc = st - sr * (sr + 1);
            const int r0 = (sr * 8 + (lane >> 2)) * 4, c0 = (sc * 4 + (lane & 3)) * 4;
            if (c0 > r0 + 3) continue;
            f(r0, c0);
        }
    }
}

// FP64 arithmetic has ~36 cycles of dependent-issue latency on B200 (DFMA at ILP16 on 1 warp/SMSP reaches 88% of peak),
// so the leaf is organised around the length of its dependent chains, not around its flop count.
//
// One warp: factor the 32x32 diagonal block at S[b0.., b0..] (lower Cholesky, in place, strict upper zeroed) and store
// 1/L_cc in rdiag[c].  Lane r keeps row r in registers; y[q] is element (r, c+q) so the current column is always y[0]
// (rolled loop, static register indices).  Software pipelined: the rank-1 update of column c-1 on columns c+1.. is
// issued while rsqrt(d_c) is in flight; only the update of the next pivot column sits on the per-pivot chain
// shfl -> rsqrt -> mul -> shfl -> fma.
__device__ __forceinline__ void diag32_factor(double* __restrict__ S, int b0, double* __restrict__ rdiag, int* info, int info_off, int lane) {
    const unsigned full = 0xffffffffu;
    double* Sd = S + b0 * LDS_ + b0;
    double y[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) y[q] = Sd[lane * LDS_ + q];
    double lp = 0.0;                       // this row's multiplier of the previous column
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
        const double d = __shfl_sync(full, y[0], c);
        if (!(d > 0.0) && lane == 0) atomicCAS(info, 0, info_off + b0 + c + 1);
        const double rd = rsqrt(d);
#pragma unroll
        for (int q = 1; q < 32; ++q) {     // column c-1 applied to columns c+1 .. (wrapped lanes hold lp == 0)
            const double lq = __shfl_sync(full, lp, (c + q) & 31);
            y[q] -= lp * lq;
        }
        const double l = (lane > c) ? y[0] * rd : 0.0;
        Sd[lane * LDS_ + c] = (lane == c) ? d * rd : l;
        if (lane == c) rdiag[c] = rd;
        const double l1 = __shfl_sync(full, l, (c + 1) & 31);
        y[0] = y[1] - l * l1;              // next pivot column, fully updated
#pragma unroll
        for (int q = 2; q < 32; ++q) y[q - 1] = y[q];
        lp = l;
    }
    __syncwarp();
}

// One thread per row below the diagonal block: row <- row * L_kk^-T by a column sweep against L_kk in shared memory
// (broadcast reads), i.e. the panel solve without forming the inverse of the diagonal block.  Chain per column: mul -> fma.
__device__ __forceinline__ void panel_sweep(double* __restrict__ S, int b0, int row, const double* __restrict__ rdiag) {
    const double* Sd = S + b0 * LDS_ + b0;
    double* Sr = S + row * LDS_ + b0;
    double y[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) y[q] = Sr[q];
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
        const double x = y[0] * rdiag[c];
        Sr[c] = x;
#pragma unroll
        for (int q = 1; q < 32; ++q) y[q - 1] = y[q] - x * Sd[((c + q) & 31) * LDS_ + c];   // wrapped rows read zeros (upper part)
    }
}

// One warp: W = L_kk^-1 by forward substitution, lane j owns column j; X kept in shared memory (Wb, ld WLD).
// Dot products on 8 independent accumulators; the division is a multiply by the stored reciprocal diagonal.
__device__ __forceinline__ void diag32_inverse(const double* __restrict__ S, int b0, double* __restrict__ Wb, const double* __restrict__ rdiag,
                                               int lane) {
    const double* Sd = S + b0 * LDS_ + b0;
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
        double s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = 0.0;
        s[0] = (i == lane) ? 1.0 : 0.0;
        const double* Li = Sd + i * LDS_;
        const int i8 = i & ~7;
#pragma unroll 1
        for (int k = 0; k < i8; k += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] -= Li[k + u] * Wb[(k + u) * WLD + lane];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i8 + u < i) s[u] -= Li[i8 + u] * Wb[(i8 + u) * WLD + lane];
        const double tot = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
        Wb[i * WLD + lane] = (i >= lane) ? tot * rdiag[i] : 0.0;
    }
    __syncwarp();
}

// S holds the lower triangle of an SPD 128x128 block (identity padded beyond the valid size by the caller).
// On return S holds inv(L) (lower); L has been written to Lout (valid nv x nv region, strict upper zeroed) and
// inv(L) to winv (128x128, ld 128, upper zero, identity padding).
template <int BAR_ID>
__device__ __forceinline__ void potf2_inverse(double* __restrict__ S, double* __restrict__ Wd, double* __restrict__ T,
                                              double* __restrict__ Lout, long ldl, int nv, double* __restrict__ winv, int* info,
                                              int info_off, int tid, long long* prof = nullptr) {
    const int warp = tid >> 5, lane = tid & 31;
    int pidx = 0;
#define LEAF_STAMP() do { if (prof && tid == 0) prof[pidx++] = clock64(); } while (0)
    LEAF_STAMP();
    // ---------------- factorisation, 32-wide block columns ----------------
    double* rdiag = T + 96 * WLD;
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb) {
        const int b0 = 32 * kb, r1 = b0 + 32, R = LB - r1;
        if (warp == 0) diag32_factor(S, b0, rdiag + b0, info, info_off, lane);
        bar256<BAR_ID>();
        LEAF_STAMP();
        if (R > 0) {
            if (tid < R) panel_sweep(S, b0, r1 + tid, rdiag + b0);
            bar256<BAR_ID>();
            LEAF_STAMP();
            // trailing update: A22 -= P P^T (lower tiles)
            for_tiles(R, R, tid, true, [&](int r0, int c0) {
                double acc[4][4];
                tile4x4<false>(S + r1 * LDS_ + b0, LDS_, S + r1 * LDS_ + b0, LDS_, 1, 32, r0, c0, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) S[(r1 + r0 + i) * LDS_ + r1 + c0 + j] -= acc[i][j];
            });
            bar256<BAR_ID>();
            LEAF_STAMP();
        }
    }
    // ---------------- write L back ----------------
    for (int idx = tid; idx < LB * LB / 2; idx += 256) {
        const int r = idx >> 6, c = (idx & 63) * 2;
        if (r < nv && c < nv) {
            const double v0 = (c <= r) ? S[r * LDS_ + c] : 0.0, v1 = (c + 1 <= r) ? S[r * LDS_ + c + 1] : 0.0;
            if (c + 1 < nv) *reinterpret_cast<double2*>(Lout + (long)r * ldl + c) = make_double2(v0, v1);
            else Lout[(long)r * ldl + c] = v0;
        }
    }
    bar256<BAR_ID>();
    LEAF_STAMP();
    // ---------------- inverse: the four diagonal blocks in parallel (one warp each), then block columns right to left ------
    if (warp < 4) diag32_inverse(S, 32 * warp, Wd + warp * 32 * WLD, rdiag + 32 * warp, lane);
    bar256<BAR_ID>();
    LEAF_STAMP();
    for (int idx = tid; idx < 4 * 32 * 32; idx += 256) {
        const int b = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
        S[(32 * b + r) * LDS_ + 32 * b + c] = Wd[b * 32 * WLD + r * WLD + c];
    }
    bar256<BAR_ID>();
#pragma unroll 1
    for (int k = 2; k >= 0; --k) {
        const int r1 = 32 * (k + 1), R = LB - r1;
        // T = L_col * W_kk           (B[c=j][q] = W_kk[q][j])
        for_tiles(R, 32, tid, false, [&](int r0, int c0) {
            double acc[4][4];
            tile4x4<false>(S + r1 * LDS_ + 32 * k, LDS_, Wd + k * 32 * WLD, 1, WLD, 32, r0, c0, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) T[(r0 + i) * WLD + c0 + j] = acc[i][j];
        });
        bar256<BAR_ID>();
        // X = -W_trail * T           (W_trail lower => k <= r)
        for_tiles(R, 32, tid, false, [&](int r0, int c0) {
            double acc[4][4];
            tile4x4<true>(S + r1 * LDS_ + r1, LDS_, T, 1, WLD, R, r0, c0, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) S[(r1 + r0 + i) * LDS_ + 32 * k + c0 + j] = -acc[i][j];
        });
        bar256<BAR_ID>();
        LEAF_STAMP();
    }
    for (int idx = tid; idx < LB * LB / 2; idx += 256) {
        const int r = idx >> 6, c = (idx & 63) * 2;
        reinterpret_cast<double2*>(winv)[idx] = make_double2((c <= r) ? S[r * LDS_ + c] : 0.0, (c + 1 <= r) ? S[r * LDS_ + c + 1] : 0.0);
    }
    LEAF_STAMP();
#undef LEAF_STAMP
}

}   // namespace leaf
