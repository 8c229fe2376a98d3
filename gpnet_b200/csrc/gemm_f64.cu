// FP64 GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> DMMA.8x8x4 consumers.
//
// Replaces the dense products behind the reference's numpy/scipy calls: scipy.linalg.expm's Pade products
// (GPnet.py:340), np.linalg.matrix_power (GPnet.py:346), the GEMM-shaped parts of scipy.linalg.inv (GPnet.py:342)
// and of np.linalg.cholesky / np.linalg.solve (GPnetRegressor.py:123-126,141-144; GPnetClassifier.py:110-121).
//
// tcgen05 has no FP64 kind (ptxas rejects .kind::f64), so the FP64 tensor path on Blackwell is the warp-level
// mma.sync m8n8k4 (SASS DMMA.8x8x4, 512 flop / warp instruction, measured 37.06 TFLOP/s chip peak).  Design:
//   * CTA tile BM x BN = (64*WM) x (32*WN), k-step 16 doubles (one 128-byte swizzle row), STAGES-deep ring.
//   * one producer warp (one elected lane) issues TMA loads and arms `full` mbarriers with expect_tx;
//     WM*WN consumer warps each own a 64 x 32 accumulator tile (64 doubles / lane) and release stages through
//     `empty` mbarriers.
//   * operands may be K-major ([mn][k], k contiguous: fragments come from conflict-free LDS.128 thanks to a
//     row permutation inside every 8-row group) or MN-major ([k][mn]: conflict-free LDS.64), so all four
//     NN/NT/TN/TT forms read straight from TMA-swizzled tiles without a transpose pass.
//   * tile-level structure flags skip work that is known to be zero (triangular operands) or redundant
//     (symmetric results: lower tiles only + mirrored store).
#include "tile_primitives.cuh"

namespace {

using namespace tp;

struct GemmArgs {
    int M, N, K;
    double* C;      // normal output (may be null when only the transposed copy is wanted)
    double* CT;     // optional transposed output: CT[col][row] = value
    long ldc, sC, ldct, sCT;
    double alpha, beta;
    int flags;
    int tiles_m, tiles_n;
    int klo_shift;   // GF_KLO_M counts from m0 - klo_shift (>= 0): op(A)[m][k] == 0 for k < m - klo_shift (a trapezoidal operand)
};

// 8-consumer configuration: 12 warps = 2 consumer warpgroups + 1 producer warpgroup, registers rebalanced with
// setmaxnreg (producer 40, consumers 232) because the register file is allocated in 4-warp granules.
// Minimum CTAs per SM the 64x64 configuration is compiled for (4 caps it at 168 registers per thread; the 128x128
// configuration owns its SM).  Compiled for one CTA per SM it took 215 registers and, while CTAs of OTHER kernels were co-resident
// on the SM under heavy memory traffic (a strided copy on another stream, several contexts evaluating at once), sporadically
// produced wrong rows 40..63 of its tiles -- the fragments held in the highest registers.  tools/gemm_race.py and
// tests/test_gpu_primitives.py::test_small_gemm_is_bit_stable_next_to_a_strided_copy reproduce it (9-12 wrong runs of 12 before,
// none with the cap); the mechanism is not understood, the cap is empirical.
#ifndef GEMM_SMALL_MIN_CTAS
#define GEMM_SMALL_MIN_CTAS 4
#endif
template <int WM, int WN>
constexpr int gemm_threads() {
    return (WM * WN == 8) ? 384 : (WM * WN + 1) * 32;
}


template <int WM, int WN, int STAGES, bool AK, bool BK>
__global__ void __launch_bounds__((WM * WN == 8) ? 384 : (WM * WN + 1) * 32, (WM * WN == 8) ? 1 : GEMM_SMALL_MIN_CTAS)
gemm_f64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    constexpr int BM = 64 * WM, BN = 32 * WN;
    constexpr int NCONS = WM * WN;
    constexpr bool REBALANCE = (NCONS == 8);
    constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = smem0 + STAGES * STAGE_BYTES;   // STAGES x 8 bytes
    const uint32_t bar_empty = bar_full + STAGES * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int batch = blockIdx.y;

    // ---- tile decode -------------------------------------------------------------------------------------
    int tm, tn;
    {
        const int tix = blockIdx.x;
        if (g.flags & GF_LOWER) {
            // lower-triangular tile enumeration, heaviest rows last by default
            int i = (int)((sqrt(8.0 * (double)tix + 1.0) - 1.0) * 0.5);
            while ((long)i * (i + 1) / 2 > tix) --i;
            while ((long)(i + 1) * (i + 2) / 2 <= tix) ++i;
            tm = i;
            tn = tix - (int)((long)i * (i + 1) / 2);
        } else {
            // grouped raster: 8 tile-rows per group so that concurrently running CTAs share A/B panels in L2
            const int GROUP = 8;
            const int per_group = GROUP * g.tiles_n;
            const int gid = tix / per_group;
            const int first_m = gid * GROUP;
            const int gsz = min(g.tiles_m - first_m, GROUP);
            const int r = tix - gid * per_group;
            tm = first_m + (r % gsz);
            tn = r / gsz;
            // longest-first: with a lower-triangular A (k < m0+BM) the bottom tile rows carry the long k-loops
            if ((g.flags & GF_KHI_M) && !(g.flags & (GF_KLO_M | GF_KLO_N))) tm = g.tiles_m - 1 - tm;
        }
    }
    const int m0 = tm * BM, n0 = tn * BN;
    int k_lo = 0, k_hi = g.K;
    if (g.flags & GF_KLO_M) k_lo = max(k_lo, m0 - g.klo_shift);
    if (g.flags & GF_KLO_N) k_lo = max(k_lo, n0);
    if (g.flags & GF_KHI_M) k_hi = min(k_hi, m0 + BM);
    if (g.flags & GF_KHI_N) k_hi = min(k_hi, n0 + BN);
    k_lo &= ~(KSTEP - 1);
    const int niter = (k_hi > k_lo) ? (k_hi - k_lo + KSTEP - 1) / KSTEP : 0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCONS) {
        // ===================== TMA producer (one elected lane of the first producer warp) =====================
        if (REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == NCONS && lane == 0) {
            for (int it = 0; it < niter; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(bar_empty + s * 8, ph ^ 1);
                if (g.flags & GF_DBG_PROD_FENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t full = bar_full + s * 8;
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t sa = smem0 + s * STAGE_BYTES, sb = sa + A_BYTES;
                const int k = k_lo + it * KSTEP;
                const CUtensorMap *pA = &tmA, *pB = &tmB;
                if (AK) {
                    tma_load_3d(sa, pA, full, k, m0, batch);
                } else {
#pragma unroll
                    for (int j = 0; j < BM / 16; ++j) tma_load_3d(sa + j * 2048, pA, full, m0 + 16 * j, k, batch);
                }
                if (BK) {
                    tma_load_3d(sb, pB, full, k, n0, batch);
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 16; ++j) tma_load_3d(sb + j * 2048, pB, full, n0 + 16 * j, k, batch);
                }
            }
            if (g.flags & GF_DBG_PROD_STAY) {      // leave only after the consumers have released every stage
                for (int it = niter; it < niter + STAGES && it - STAGES >= 0; ++it) mbar_wait(bar_empty + (it % STAGES) * 8, ((it / STAGES) & 1) ^ 1);
            }
        }
        return;
    }

    // ===================== DMMA consumers =====================
    if (REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm = warp / WN, wn = warp % WN;
    const int gq = lane >> 2, t = lane & 3;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int it = 0; it < niter; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        const uint32_t sa = smem0 + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double af[8][2], bf[4][2];
            load_frags<8, AK>(af, sa, wm * 64, h, gq, t);
            load_frags<4, BK>(bf, sb, wn * 32, h, gq, t);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j], af[i][u], bf[j][u]);
        }
        if (g.flags & GF_DBG_CONS_FENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + s * 8);
    }

    // ---- epilogue: C = alpha*acc + beta*C (+ mirrored store) ---------------------------------------------
    double* __restrict__ C = g.C ? g.C + (long)batch * g.sC : nullptr;
    double* __restrict__ CT = g.CT ? g.CT + (long)batch * g.sCT : nullptr;
    const bool mirror = (g.flags & GF_MIRROR) && (tm != tn);
    const int pg = (gq >> 1) | ((gq & 1) << 2);
    const int rsel = AK ? pg : gq;
    const double alpha = g.alpha, beta = g.beta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm * 64 + i * 8 + rsel;
        if (row >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn * 32 + j * 8 + (BK ? (t + 4 * e) : (2 * t + e));
                if (col < g.N) {
                    double v = alpha * acc[i][j][e];
                    if (C) {
                        double* p = C + (long)row * g.ldc + col;
                        if (beta != 0.0) v += beta * (*p);
                        *p = v;
                        if (mirror) C[(long)col * g.ldc + row] = v;
                    }
                    if (CT) CT[(long)col * g.ldct + row] = v;
                }
            }
        }
    }
}

template <int WM, int WN, int STAGES>
constexpr size_t gemm_smem_bytes() {
    return (size_t)STAGES * (64 * WM + 32 * WN) * 128 + 2 * STAGES * 8 + 1024;
}

int encode_operand(gpn_ctx* c, CUtensorMap* tm, const double* ptr, long ld, bool kmajor, int mn_extent, int k_extent,
                   int tile_mn, int batch, long stride) {
    if (((uintptr_t)ptr & 15) || (ld & 1) || (batch > 1 && (stride & 1))) return -3;
    cuuint64_t gdim[3], gstr[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (k_extent < 1) k_extent = 1;   // K == 0: no k-steps are issued, the map only has to be encodable
    if (kmajor) {   // [mn][k]
        gdim[0] = (cuuint64_t)k_extent;
        gdim[1] = (cuuint64_t)mn_extent;
        box[0] = KSTEP;
        box[1] = (cuuint32_t)tile_mn;
    } else {        // [k][mn]
        gdim[0] = (cuuint64_t)mn_extent;
        gdim[1] = (cuuint64_t)k_extent;
        box[0] = 16;
        box[1] = KSTEP;
    }
    gdim[2] = (cuuint64_t)batch;
    box[2] = 1;
    gstr[0] = (cuuint64_t)ld * 8;
    gstr[1] = (cuuint64_t)((batch > 1) ? stride : (long)gdim[1] * ld) * 8;
    if (gstr[1] == 0) gstr[1] = 16;
    CUresult r = c->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): ptr=%p ld=%ld mn=%d k=%d batch=%d stride=%ld", (int)r,
                 (const void*)ptr, ld, mn_extent, k_extent, batch, stride);
        gpn_set_last_error(__FILE__, __LINE__, msg);
        return -1100 - (int)r;
    }
    return 0;
}

template <int WM, int WN, int STAGES, bool AK, bool BK>
int launch_cfg(gpn_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int ntiles, int batch) {
    auto kern = gemm_f64_kernel<WM, WN, STAGES, AK, BK>;
    constexpr size_t smem = gemm_smem_bytes<WM, WN, STAGES>();
    static bool configured = false;
    if (!configured) {
        GPN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(ntiles, batch), block(gemm_threads<WM, WN>());
    size_t smem_req = smem;
    if ((a.flags & GF_DBG_ONE_CTA) && smem < 120 * 1024) {      // debug: one CTA per SM
        static bool big_configured = false;
        if (!big_configured) {
            GPN_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
            big_configured = true;
        }
        smem_req = 120 * 1024;
    }
    kern<<<grid, block, smem_req, c->stream>>>(ta, tb, a);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}

template <int WM, int WN, int STAGES>
int launch_layout(gpn_ctx* c, bool ak, bool bk, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int ntiles,
                  int batch) {
    if (ak && bk) return launch_cfg<WM, WN, STAGES, true, true>(c, ta, tb, a, ntiles, batch);
    if (ak && !bk) return launch_cfg<WM, WN, STAGES, true, false>(c, ta, tb, a, ntiles, batch);
    if (!ak && bk) return launch_cfg<WM, WN, STAGES, false, true>(c, ta, tb, a, ntiles, batch);
    return launch_cfg<WM, WN, STAGES, false, false>(c, ta, tb, a, ntiles, batch);
}

}   // namespace

int gpn_gemm(gpn_ctx* c, int M, int N, int K, double alpha, const double* A, long lda, bool a_kmajor, const double* B, long ldb,
             bool b_kmajor, double beta, double* C, long ldc, int flags, int batch, long sA, long sB, long sC, double* CT, long ldct,
             long sCT, int klo_shift) {
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    if (K < 0) return -4;
    if ((flags & GF_LOWER) && M != N) return -2;
    // tile configuration: 128x128 (8 consumer warps) when it fills the chip, else 64x64 (2 consumer warps, 3 CTAs/SM)
    auto count_tiles = [&](int bm, int bn, int& tmn, int& tnn) {
        tmn = (M + bm - 1) / bm;
        tnn = (N + bn - 1) / bn;
        return (flags & GF_LOWER) ? (long)tmn * (tmn + 1) / 2 : (long)tmn * tnn;
    };
    int tm_b, tn_b, tm_s, tn_s;
    const long big = count_tiles(128, 128, tm_b, tn_b) * batch;
    const long small = count_tiles(64, 64, tm_s, tn_s) * batch;
    bool use_big = big >= (long)c->num_sms * 3 / 4;
    static const int dbg = getenv("GPNET_B200_GEMM_DEBUG") ? (int)strtol(getenv("GPNET_B200_GEMM_DEBUG"), nullptr, 0) : 0;
    flags |= dbg & (GF_DBG_PROD_FENCE | GF_DBG_CONS_FENCE | GF_DBG_ONE_CTA | GF_DBG_PROD_STAY | GF_FORCE_BIG);
    if (flags & GF_FORCE_SMALL) use_big = false;
    if (flags & GF_FORCE_BIG) use_big = true;
    const int bm = use_big ? 128 : 64, bn = bm;
    GemmArgs a;
    a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.sC = sC; a.CT = CT; a.ldct = ldct; a.sCT = sCT; a.alpha = alpha; a.beta = beta; a.flags = flags; a.klo_shift = klo_shift;
    a.tiles_m = use_big ? tm_b : tm_s;
    a.tiles_n = use_big ? tn_b : tn_s;
    const long ntiles = (use_big ? big : small) / batch;
    CUtensorMap ta, tb;
    GPN_TRY(encode_operand(c, &ta, A, lda, a_kmajor, M, K, bm, batch, sA));
    GPN_TRY(encode_operand(c, &tb, B, ldb, b_kmajor, N, K, bn, batch, sB));
    // executed-flop accounting: exact sum over the tiles of the k-steps each one runs (structure flags respected)
    double ksum = 0.0;
    if (flags & (GF_KLO_M | GF_KLO_N | GF_KHI_M | GF_KHI_N)) {
        for (int tm = 0; tm < a.tiles_m; ++tm)
            for (int tn = 0; tn < ((flags & GF_LOWER) ? tm + 1 : a.tiles_n); ++tn) {
                int k_lo = 0, k_hi = K;
                if (flags & GF_KLO_M) k_lo = k_lo > tm * bm - klo_shift ? k_lo : tm * bm - klo_shift;
                if (flags & GF_KLO_N) k_lo = k_lo > tn * bn ? k_lo : tn * bn;
                if (flags & GF_KHI_M) k_hi = k_hi < (tm + 1) * bm ? k_hi : (tm + 1) * bm;
                if (flags & GF_KHI_N) k_hi = k_hi < (tn + 1) * bn ? k_hi : (tn + 1) * bn;
                k_lo &= ~(KSTEP - 1);
                if (k_hi > k_lo) ksum += (double)((k_hi - k_lo + KSTEP - 1) / KSTEP * KSTEP);
            }
    } else {
        ksum = (double)ntiles * ((K + KSTEP - 1) / KSTEP * KSTEP);
    }
    const double fl = 2.0 * bm * bn * ksum * batch;
    c->flops += fl;
    ProfEvent pe;
    if (c->profile) {
        cudaEventCreate(&pe.e0);
        cudaEventCreate(&pe.e1);
        pe.flops = fl;
        pe.big = use_big;
        cudaEventRecord(pe.e0, c->stream);
    }
    const int st = use_big ? launch_layout<2, 4, 6>(c, a_kmajor, b_kmajor, ta, tb, a, (int)ntiles, batch)
                           : launch_layout<1, 2, 4>(c, a_kmajor, b_kmajor, ta, tb, a, (int)ntiles, batch);
    if (c->profile) {
        cudaEventRecord(pe.e1, c->stream);
        c->prof_events.push_back(pe);
    }
    return st;
}
