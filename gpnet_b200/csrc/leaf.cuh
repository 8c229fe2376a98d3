// 128x128 Cholesky leaf: factor AND invert one diagonal block in shared memory with 256 threads (8 warps).
// Used by the stand-alone leaf kernel (factor.cu) and inside the dataflow Cholesky (potrf_dataflow.cu).
//
//  * 32x32 diagonal sub-blocks: ONE warp, rows in registers, software-pipelined column sweep (no block barrier on the
//    per-pivot dependency chain); a partner warp on another SMSP inverts the block in lockstep, one pivot step behind;
//  * panel below a diagonal block, trailing updates, block rows of the inverse: DMMA shared-memory products;
//  * global write-backs and everything that is not on the chain run next to the factorisation of the next diagonal block.
// Shared memory: guard band, S[128][129] (padded, conflict-free column access; its unused strictly-upper 32-blocks hold the
// off-diagonal blocks of the inverse), Wd[4][32][33] (diagonal blocks of the inverse), T (32 x 97 scratch), 1/diag, column
// broadcast buffer, 32 step mbarriers.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "tile_primitives.cuh"

namespace leaf {

constexpr int LB = 128;
constexpr int LDS_ = LB + 1;
constexpr int WLD = 33;
constexpr int GUARD = 32;                   // doubles in front of S (kept zero; absorbs reads just before a row start)
constexpr int TLD = 97;                     // row stride of T when it holds a 32 x 96 block row (v2)
constexpr size_t SMEM_DOUBLES = GUARD + (size_t)LB * LDS_ + 4 * 32 * WLD + 96 * WLD + LB + 128 + 32;   // guard, S, Wd, T, 1/diag, column broadcast, 32 step mbarriers
constexpr size_t SMEM_BYTES = SMEM_DOUBLES * sizeof(double);

template <int BAR_ID>
__device__ __forceinline__ void bar256() {
    if (BAR_ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, 256;" ::"n"(BAR_ID) : "memory");
}

// 1/sqrt(d) for normal positive d without the library's special-case branch (a branch in the pivot loop is a scheduling
// barrier: the independent bulk update could not be issued under the latency of this chain).  MUFU.RSQ64H seed (~2^-22)
// + one cubic Newton step (error ~ e^3), the same arithmetic the CUDA math library uses on its fast path.
// d <= 0 / NaN give garbage; the caller records those pivots in `info`.
__device__ __forceinline__ double rsqrt_nobranch(double d) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double e = fma(-d, y0 * y0, 1.0);
    const double p = fma(e, 0.375, 0.5);
    return fma(p, y0 * e, y0);
}

// 1/d for normal positive d: MUFU.RCP64H seed + one cubic Newton step, branch free (see rsqrt_nobranch).
__device__ __forceinline__ double rcp_nobranch(double d) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    const double e = fma(-d, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// Shared-memory GEMM on the FP64 tensor pipe for the small products of the leaf:
//     C(m, n) = sum_k A[m][k] * Bop(k, n),   A row-major with stride SA,   Bop(k, n) = BT ? B[n*SB + k] : B[k*SB + n].
// Work unit = 16 rows x 32 columns (2 x 4 DMMA blocks = 8 independent accumulators: a DMMA has ~140 cycles of latency, so
// a warp needs about that many in flight), dealt round-robin to the 8 warps.  TRI = 1: A is lower triangular in (m, k)
// (k-steps beyond the row blocks are skipped); TRI = 2: only units touching the lower triangle of a square C.
// epi(row, col, value) receives the values a lane owns (C fragment: row g, cols 2t, 2t+1 of each 8x8 block).
// Only the warps in `wmask` take part; skip_lead drops the leading 32x32 block (already updated ahead of the others).
template <int SA, int SB, bool BT, int TRI, typename Epi>
__device__ __forceinline__ void smem_dmma(const double* __restrict__ A, const double* __restrict__ B, int munits, int nunits, int K, int tid,
                                          Epi epi, unsigned wmask = 0xffu, bool skip_lead = false) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (!((wmask >> warp) & 1u)) return;
    const int nw = __popc(wmask), wrank = __popc(wmask & ((1u << warp) - 1u));
    for (int u = wrank; u < munits * nunits; u += nw) {
        const int mu = u / nunits, nu = u - mu * nunits;
        if (TRI == 2 && 32 * nu > 16 * mu + 15) continue;
        if (skip_lead && nu == 0 && mu < 2) continue;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const int kend = (TRI == 1) ? min(K, 16 * mu + 16) : K;
        const double* ap = A + (16 * mu + g) * SA + t;
        const double* bp = BT ? B + (32 * nu + g) * SB + t : B + t * SB + 32 * nu + g;
#pragma unroll 2
        for (int k0 = 0; k0 < kend; k0 += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = ap[i * 8 * SA + k0];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = BT ? bp[j * 8 * SB + k0] : bp[k0 * SB + 8 * j];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) epi(16 * mu + 8 * i + g, 32 * nu + 8 * j + 2 * t + e, acc[i][j][e]);
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
__device__ __forceinline__ void diag32_factor(double* __restrict__ S, int b0, double* __restrict__ rdiag, double* __restrict__ colbuf,
                                              int* info, int info_off, int lane, uint32_t step_bars = 0) {
    const unsigned full = 0xffffffffu;
    double* Sd = S + b0 * LDS_ + b0;
    double y[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) y[q] = Sd[lane * LDS_ + q];
    // colbuf: 2 x 64 doubles; entries 32..63 of each half stay zero so that "row c+q" needs no wrap-around arithmetic
    colbuf[lane] = 0.0; colbuf[32 + lane] = 0.0; colbuf[64 + lane] = 0.0; colbuf[96 + lane] = 0.0;
    __syncwarp();
    double lp = 0.0;                       // this row's multiplier of the previous column
    double* sp = Sd + lane * LDS_;         // &L[lane][c]
    int first_bad = 32;                    // first column whose pivot is not positive (LAPACK info), reported after the loop
    auto pivot_step = [&](int c, auto qm_tag) {
        constexpr int QM = decltype(qm_tag)::value;        // live slots: columns c .. c+QM
        // Per-pivot dependent chain (FP64 ops cost ~40 cycles each): shfl -> 1/d -> fma.  The next pivot column needs
        // l_r l_{c+1} = (y_r y_{c+1}) / d, so it does not wait for rsqrt(d); the multipliers themselves (l = y rsqrt(d))
        // are only consumed by the NEXT iteration's bulk update and by the stores, off the chain.
        const double d = __shfl_sync(full, y[0], c);
        const double yn = __shfl_sync(full, y[0], (c + 1) & 31);        // element (c+1, c) before scaling
        first_bad = (!(d > 0.0) && first_bad > c) ? c : first_bad;      // select, not a branch
        const double id = rcp_nobranch(d);
        const double rd = rsqrt_nobranch(d);
        const double w = (lane > c) ? y[0] * yn : 0.0;
        // column c-1 applied to columns c+1 .. : multipliers of rows c+1.. are broadcast from shared memory (pointer +
        // immediate loads); results are written one slot down so that column c+1 becomes y[0] for the next pivot
        const double* lc = colbuf + ((c & 1) << 6) + c;      // lc[q] = l_{c+q, c-1}  (zero beyond row 31)
        const double t1 = y[1] - lp * lc[1];
#pragma unroll
        for (int q = 2; q <= QM; ++q) y[q - 1] = y[q] - lp * lc[q];
        const double l = (lane > c) ? y[0] * rd : 0.0;
        *sp = (lane == c) ? d * rd : l;
        if (lane == c) rdiag[c] = rd;
        colbuf[(((c + 1) & 1) << 6) + lane] = l;             // published for the next iteration's bulk update
        y[0] = fma(-w, id, t1);            // next pivot column, fully updated
        lp = l;
        ++sp;
        __syncwarp();
        // column c of L and 1/L_cc are in shared memory: release the partner warp (diag32_inverse_lockstep) for step c.
        // One mbarrier per step: an arrive is a fire-and-forget shared-memory atomic (a st.release here costs a fence on the chain).
        if (step_bars != 0 && lane == 0) tp::mbar_arrive(step_bars + 8 * c);
    };
#pragma unroll 1
    for (int c = 0; c < 16; ++c) pivot_step(c, std::integral_constant<int, 31>{});
#pragma unroll 1
    for (int c = 16; c < 32; ++c) pivot_step(c, std::integral_constant<int, 16>{});   // columns >= 32 do not exist
    if (first_bad < 32 && lane == 0) atomicCAS(info, 0, info_off + b0 + first_bad + 1);
}

// Lookahead piece of the trailing update: only the NEXT diagonal block, D -= P P^T with P = its 32 rows of the freshly solved
// panel (K = 32).  The ten lower 8x8 DMMA blocks go to warps 0..7 (+ blocks 8, 9 to warps 0, 1); every block splits K in
// two halves so that all of a warp's dependent DMMA chains (4 long) are in flight at once.
__device__ __forceinline__ void trail_diag(double* __restrict__ S, int b0, int tid) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r1 = b0 + 32;
    const double* P = S + r1 * LDS_ + b0;
    double* D = S + r1 * LDS_ + r1;
    double acc[2][2][2];
    int mb[2], nb[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
        const int L = warp + 8 * s2;                                   // lower blocks in row-major order: (0,0) (1,0) (1,1) (2,0) ..
        mb[s2] = (L >= 6) ? 3 : (L >= 3) ? 2 : (L >= 1) ? 1 : 0;
        nb[s2] = L - mb[s2] * (mb[s2] + 1) / 2;
        acc[s2][0][0] = acc[s2][0][1] = acc[s2][1][0] = acc[s2][1][1] = 0.0;
    }
    const bool two = warp < 2;
    const double* ap0 = P + (8 * mb[0] + g) * LDS_ + t;
    const double* bp0 = P + (8 * nb[0] + g) * LDS_ + t;
    const double* ap1 = two ? P + (8 * mb[1] + g) * LDS_ + t : ap0;
    const double* bp1 = two ? P + (8 * nb[1] + g) * LDS_ + t : bp0;
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {
        dmma884(acc[0][0], ap0[k0], bp0[k0]);
        dmma884(acc[0][1], ap0[k0 + 16], bp0[k0 + 16]);
        dmma884(acc[1][0], ap1[k0], bp1[k0]);
        dmma884(acc[1][1], ap1[k0 + 16], bp1[k0 + 16]);
    }
    D[(8 * mb[0] + g) * LDS_ + 8 * nb[0] + 2 * t] -= acc[0][0][0] + acc[0][1][0];
    D[(8 * mb[0] + g) * LDS_ + 8 * nb[0] + 2 * t + 1] -= acc[0][0][1] + acc[0][1][1];
    if (two) {
        D[(8 * mb[1] + g) * LDS_ + 8 * nb[1] + 2 * t] -= acc[1][0][0] + acc[1][1][0];
        D[(8 * mb[1] + g) * LDS_ + 8 * nb[1] + 2 * t + 1] -= acc[1][0][1] + acc[1][1][1];
    }
}

// One warp, IN LOCKSTEP with diag32_factor running on another warp: W = L_kk^-1 by right-looking forward substitution
// (L W = I).  Lane c owns COLUMN c of W; z[q] is the pending value of row k+q, so the current row is z[0]:
//     step k (once column k of L and 1/L_kk are published):  w = z[0] / L_kk = W[k][c];   z[q-1] = z[q] - L[k+q][k] w
// Multipliers are broadcast reads of L's column k in shared memory; chain per step: mul -> fma, well under the ~237 cycles
// of a pivot step, so the inverse of a diagonal block is complete about one step after its factor.
__device__ __forceinline__ void diag32_inverse_lockstep(const double* __restrict__ S, int b0, double* __restrict__ Wb,
                                                        const double* __restrict__ rdiag, uint32_t step_bars, uint32_t parity, int lane) {
    const double* lp = S + b0 * LDS_ + b0;          // &L_kk[k][k], walked down the diagonal
    double* wp = Wb + lane;                         // &W[k][lane]
    double z[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) z[q] = (lane == q) ? 1.0 : 0.0;
    auto step = [&](int k, auto qm_tag) {
        constexpr int QM = decltype(qm_tag)::value;
        tp::mbar_wait(step_bars + 8 * k, parity);   // column k of L and 1/L_kk are published
        const double w = z[0] * rdiag[k];
        *wp = w;                                    // zero for k < lane by construction
#pragma unroll
        for (int q = 1; q <= QM; ++q) z[q - 1] = z[q] - lp[q * LDS_] * w;      // L[k+q][k]; rows >= 32 feed slots that are never used
        lp += LDS_ + 1;
        wp += WLD;
    };
#pragma unroll 1
    for (int k = 0; k < 16; ++k) step(k, std::integral_constant<int, 31>{});
#pragma unroll 1
    for (int k = 16; k < 32; ++k) step(k, std::integral_constant<int, 15>{});
    __syncwarp();
}

// v2: panel below diagonal block kb, P <- P * W_kk^T in place (R x 32, K = 32).  W_kk is lower triangular, so the 8-column
// block j only needs k < 8 j + 8.  Unit = 8 rows x 32 columns; a warp keeps up to two units in flight and splits K in two
// halves per accumulator, so the dependent DMMA chains are at most 4 long (a DMMA has ~140 cycles of latency).
__device__ __forceinline__ void panel_product(double* __restrict__ S, int b0, const double* __restrict__ Wkk, int tid) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r1 = b0 + 32, units = (LB - r1) / 8;
    if (warp >= units) return;
    const bool two = warp + 8 < units;
    double* P0 = S + (r1 + 8 * warp + g) * LDS_ + b0;
    double* P1 = two ? P0 + 64 * LDS_ : P0;
    const double* bp = Wkk + g * WLD + t;           // B(k, n) = W[n][k], n = 8 j + g, k = k0 + t
    double acc[2][4][2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[u][j][0][0] = acc[u][j][0][1] = acc[u][j][1][0] = acc[u][j][1][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
        const double a0 = P0[t + k0], a1 = P1[t + k0];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 < 8 * j + 8) {
                const double b = bp[j * 8 * WLD + k0];
                dmma884(acc[0][j][k0 >> 4], a0, b);
                dmma884(acc[1][j][k0 >> 4], a1, b);
            }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            P0[8 * j + 2 * t + e] = acc[0][j][0][e] + acc[0][j][1][e];
            if (two) P1[8 * j + 2 * t + e] = acc[1][j][0][e] + acc[1][j][1][e];
        }
}

// v2 helper: T_i = L(i, 0:i) * W(0:i, 0:i) (32 x 32i) for block row i of the inverse, from the blocks of W computed so far:
// diagonal blocks W_jj live in Wd[j], off-diagonal blocks W(k, j) (k > j) in the otherwise unused UPPER block (j, k) of S.
// One 16 x 32 unit per (half, j), dealt to the warps of `wmask`.
__device__ __forceinline__ void inv_row_T(const double* __restrict__ S, const double* __restrict__ Wd, double* __restrict__ T, int i, int tid,
                                          unsigned wmask) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (!((wmask >> warp) & 1u)) return;
    const int nw = __popc(wmask), wrank = __popc(wmask & ((1u << warp) - 1u));
    for (int u = wrank; u < 2 * i; u += nw) {
        const int h = u & 1, j = u >> 1;
        double acc[2][4][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        for (int k = j; k < i; ++k) {
            const double* ap = S + (32 * i + 16 * h + g) * LDS_ + 32 * k + t;                 // L(i, k)
            const double* bp = (k == j) ? Wd + j * 32 * WLD + t * WLD + g                     // W_jj[kk][n]
                                        : S + (32 * j + t) * LDS_ + 32 * k + g;               // W(k, j)[kk][n] at block (j, k)
            const int sb = (k == j) ? WLD : LDS_;
#pragma unroll 2
            for (int k0 = 0; k0 < 32; k0 += 4) {
                double a[2], b[4];
#pragma unroll
                for (int x = 0; x < 2; ++x) a[x] = ap[x * 8 * LDS_ + k0];
#pragma unroll
                for (int y = 0; y < 4; ++y) b[y] = bp[k0 * sb + 8 * y];
#pragma unroll
                for (int x = 0; x < 2; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) dmma884(acc[x][y], a[x], b[y]);
            }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
#pragma unroll
                for (int e = 0; e < 2; ++e) T[(16 * h + 8 * x + g) * TLD + 32 * j + 8 * y + 2 * t + e] = acc[x][y][e];
    }
}

// v2 helper: global write-back of what block step kb has finalised, by the warps in `wmask` (half a warp per 32-double
// row segment, double2 per lane): block column kb of L (rows 32 kb .., lower part; only the diagonal block if diag_only) and
// block row kb of the inverse (mirror blocks of S + Wd[kb]).
__device__ __forceinline__ void write_back_step(const double* __restrict__ S, const double* __restrict__ Wd, int kb, double* __restrict__ Lout,
                                                long ldl, int nv, double* __restrict__ winv, int tid, unsigned wmask, bool diag_only) {
    const int warp = tid >> 5, lane = tid & 31;
    if (!((wmask >> warp) & 1u)) return;
    const int nhw = 2 * __popc(wmask), hw = 2 * __popc(wmask & ((1u << warp) - 1u)) + (lane >> 4);
    const int b0 = 32 * kb, c2 = 2 * (lane & 15);
    const int lend = diag_only ? b0 + 32 : LB;
#pragma unroll 4
    for (int r = b0 + hw; r < lend; r += nhw) {
        const int c = b0 + c2;
        if (r < nv && c <= r) {
            const double v0 = S[r * LDS_ + c], v1 = (c + 1 <= r) ? S[r * LDS_ + c + 1] : 0.0;
            if (c + 1 < nv) *reinterpret_cast<double2*>(Lout + (long)r * ldl + c) = make_double2(v0, v1);
            else Lout[(long)r * ldl + c] = v0;
        }
    }
    const int nseg = 32 * (kb + 1);                   // (row, 32-column segment) pairs of block row kb of the inverse
#pragma unroll 4
    for (int q = hw; q < nseg; q += nhw) {
        const int rr = q & 31, j = q >> 5;
        double v0, v1;
        if (j < kb) {                                  // W(kb, j) lives in the upper block (j, kb) of S
            v0 = S[(32 * j + rr) * LDS_ + b0 + c2];
            v1 = S[(32 * j + rr) * LDS_ + b0 + c2 + 1];
        } else {
            v0 = (c2 <= rr) ? Wd[kb * 32 * WLD + rr * WLD + c2] : 0.0;
            v1 = (c2 + 1 <= rr) ? Wd[kb * 32 * WLD + rr * WLD + c2 + 1] : 0.0;
        }
        *reinterpret_cast<double2*>(winv + (b0 + rr) * LB + 32 * j + c2) = make_double2(v0, v1);
    }
}

// S holds the lower triangle of an SPD 128x128 block (identity padded beyond the valid size by the caller).
// On return L has been written to Lout (valid nv x nv region, lower part) and inv(L) to winv (128x128, ld 128, lower part,
// identity padding); S is scratch.
//
// Organised so that little but the 128 pivot steps is left on the critical path:
//  * while warp 0 factors a 32x32 diagonal block, warp 1 (another SMSP) inverts it in lockstep (diag32_inverse_lockstep);
//  * with W_kk at hand the panel below the block is ONE small DMMA product (panel * W_kk^T) instead of a 32-step sweep;
//  * block row kb of the inverse, W(kb, :) = -W_kk (L(kb, :kb) W(:kb, :kb)), is finished right after W_kk: the product
//    T = L(kb, :) W is formed during the factorisation of block kb by the warps that are not on the chain, and so are the
//    global write-backs of everything step kb - 1 finalised.
// Chain: D0, then 3 x [panel product + next diagonal update, D(kb+1)], then the last -W_33 T and 32 rows of write-back.
template <int BAR_ID>
__device__ __forceinline__ void potf2_inverse(double* __restrict__ S, double* __restrict__ Wd, double* __restrict__ T,
                                                 double* __restrict__ Lout, long ldl, int nv, double* __restrict__ winv, int* info,
                                                 int info_off, int tid, long long* prof = nullptr) {
    const int warp = tid >> 5, lane = tid & 31;
    int pidx = 0;
#define LEAF_STAMP() do { if (prof && tid == 0) prof[pidx++] = clock64(); } while (0)
    LEAF_STAMP();
    double* rdiag = T + 96 * WLD;
    // one mbarrier per pivot step of a 32-block; each completes one phase per block, i.e. four per call: parity = kb & 1
    const uint32_t bars = tp::smem_u32(rdiag + LB + 128);
    constexpr unsigned OFFCHAIN = 0xECu;        // warps 2, 3, 5, 6, 7: not warp 0 (factor), 1 (lockstep inverse), 4 (shares warp 0's SMSP)
    if (tid < 32) tp::mbar_init(bars + 8 * tid, 1);
    bar256<BAR_ID>();
    if (warp == 0) diag32_factor(S, 0, rdiag, rdiag + LB, info, info_off, lane, bars);
    else if (warp == 1) diag32_inverse_lockstep(S, 0, Wd, rdiag, bars, 0, lane);
    bar256<BAR_ID>();
    LEAF_STAMP();
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb) {
        const int b0 = 32 * kb, r1 = b0 + 32, R = LB - r1;
        // ---- (a) everything that was waiting for W_kk: the panel below block kb and block row kb of the inverse ----
        if (kb < 3) panel_product(S, b0, Wd + kb * 32 * WLD, tid);
        if (kb > 0) {
            const unsigned wm = kb == 1 ? 0xC0u : (kb == 2 ? 0xF0u : 0xFFu);      // the warps with the least panel work
            smem_dmma<WLD, TLD, false, 1>(Wd + kb * 32 * WLD, T, 2, kb, 32, tid,
                                          [&](int r, int cc, double v) { S[(32 * (cc >> 5) + r) * LDS_ + b0 + (cc & 31)] = -v; }, wm);
        }
        bar256<BAR_ID>();
        LEAF_STAMP();
        if (kb == 3) break;
        // ---- (b) next diagonal block first ----
        trail_diag(S, b0, tid);
        bar256<BAR_ID>();
        if (prof && tid == 0) prof[12 + kb] = clock64();
        // ---- (c) chain: factor D(kb+1) + lockstep inverse.  Off the chain (warps 2, 3, 5, 6, 7; warp 4 only moves memory next
        //      to the factor warp): what the NEXT step needs of the trailing matrix, left-looking -- block column kb+1 below its
        //      diagonal block and diagonal block kb+2, each with all panels so far (K = 32 (kb+1)); the newest panel's
        //      contribution to a diagonal block is always added by trail_diag --, T for block row kb+1 of the inverse, and the
        //      write-back of block column kb of L / block row kb of the inverse shared by all six warps ----
        if (warp == 0) {
            diag32_factor(S, r1, rdiag + r1, rdiag + LB, info, info_off, lane, bars);
            if (prof && lane == 0) prof[16 + kb] = clock64();
        } else if (warp == 1) {
            diag32_inverse_lockstep(S, r1, Wd + (kb + 1) * 32 * WLD, rdiag + r1, bars, (kb + 1) & 1, lane);
            if (prof && lane == 0) prof[20 + kb] = clock64();
        } else {
            if (warp != 4) {
                if (R > 32) {
                    const double* rows2 = S + (r1 + 32) * LDS_;           // rows below diagonal block kb+1, all panels 0..kb
                    double* Cp = S + (r1 + 32) * LDS_ + r1;               // block column kb+1 below its diagonal block
                    smem_dmma<LDS_, LDS_, true, 0>(rows2, S + r1 * LDS_, (R - 32) / 16, 1, r1, tid,
                                                   [&](int r, int cc, double v) { Cp[r * LDS_ + cc] -= v; }, OFFCHAIN);
                    double* Dn = S + (r1 + 32) * LDS_ + r1 + 32;          // diagonal block kb+2
                    smem_dmma<LDS_, LDS_, true, 2>(rows2, rows2, 2, 1, r1, tid, [&](int r, int cc, double v) { Dn[r * LDS_ + cc] -= v; },
                                                   kb == 0 ? 0xC0u : 0x60u);     // warps 6, 7 resp. 5, 6: the first product's least loaded
                }
                inv_row_T(S, Wd, T, kb + 1, tid, OFFCHAIN);
                if (prof && tid == 64) prof[24 + kb] = clock64();
            }
            write_back_step(S, Wd, kb, Lout, ldl, nv, winv, tid, OFFCHAIN | 0x10u, false);
            if (prof && tid == 128) prof[28 + kb] = clock64();
        }
        bar256<BAR_ID>();
        LEAF_STAMP();
    }
    // ---- what the last step finalised: diagonal block 3 of L, block row 3 of the inverse ----
    write_back_step(S, Wd, 3, Lout, ldl, nv, winv, tid, 0xFFu, true);
    if (tid < 32) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bars + 8 * tid) : "memory");   // the region is reused (TMA ring)
    LEAF_STAMP();
#undef LEAF_STAMP
}

}   // namespace leaf
