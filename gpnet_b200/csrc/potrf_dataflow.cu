// Dataflow (task-parallel) Cholesky for sm_100a: ONE persistent kernel factors the whole matrix.
//
// Replaces np.linalg.cholesky (GPnetRegressor.py:123,141; GPnetClassifier.py:110,159,208) and the Cholesky inside the
// SPD inverse that stands in for scipy.linalg.inv (GPnet.py:342).
//
// The matrix is cut into 128x128 tiles; a task is one output tile (i, j), i >= j, processed left-looking:
//     C   = A_ij - sum_{k<j} L_ik L_jk^T        TMA -> smem ring -> DMMA.8x8x4, accumulators stay in registers (K = 128 j)
//     i == j : L_jj = chol(C), W_j = L_jj^-1     in shared memory by the same CTA (leaf.cuh)
//     i >  j : L_ij = C W_j^T                    second in-CTA DMMA GEMM, C staged through shared memory
// Tasks are enumerated column-major and dealt round-robin to a co-resident grid (one CTA per SM, cooperative launch);
// a task only waits on tasks that precede it in that order (tiles (i,k), (j,k), k < j, and (j,j)), so the schedule is
// deadlock-free, and several tile columns are in flight at once: the serial chain chol(j) -> trsm(j+1,j) -> chol(j+1)
// overlaps the bulk updates of later columns.  Dependencies are epoch-stamped flags in global memory
// (st.release / ld.acquire at gpu scope, fence.proxy.async before the TMA reads of tiles written by other SMs).
//
// Appended row blocks ("augmented" factorisation): below the square SPD block the task list may carry further tile rows whose
// tiles obey the same recurrence with A_ij taken from another buffer,
//     Y_ij = (B_ij - sum_{k<j} Y_ik L_jk^T) W_j^T        i.e.  Y = B L^-T  (a triangular solve with many right-hand sides),
// and an identity block (B = I, tiles left of its diagonal are zero and skipped, k starts at the tile's own row) that yields
// L^-T itself.  These tasks have no chain of their own: they fill the SMs that the per-column dependency chain of a small
// Cholesky leaves idle, and they replace the separate triangular inverse / triangular product launches that would follow.
#include "leaf.cuh"
#include "tile_primitives.cuh"

namespace {

using namespace tp;

constexpr int TB = 128;
constexpr int STAGES = 6;
constexpr uint32_t TILE_BYTES = TB * 128;               // one operand tile per k-step: 128 rows x 16 doubles
constexpr uint32_t STAGE_BYTES = 2 * TILE_BYTES;
constexpr uint32_t RING_BYTES = STAGES * STAGE_BYTES;   // 192 KB
constexpr uint32_t CREG_BYTES = 8 * TILE_BYTES;         // C staged as 8 K-major sub-tiles (stages 0..3 of the ring)
constexpr int WSTAGES = 4;                              // W_j ring in stages 4..5 of the ring
constexpr size_t DF_SMEM = RING_BYTES + 256 + 1024;

static_assert(leaf::SMEM_BYTES <= RING_BYTES, "leaf scratch must fit in the ring");

struct DfSeg {          // one block of tile rows: [row0, row0 + rows) in tile units, stored at ptr with leading dimension ld
    double* ptr;
    long ld;
    int row0, rows;
};
struct DfArgs {
    double* A;
    long lda;
    int n, T, ntasks;
    int Trows;            // tile rows in total (T for a plain factorisation)
    int id0;              // first tile row of the identity block (Trows if there is none)
    int Gc;               // CTAs [0, Gc) own the square block's tasks (the dependency chain); all CTAs if nothing is appended
    long nspd, napp;      // tasks of the square block; appended tile slots (column-major, Trows - T per column)
    int* counters;        // two task counters (zeroed per launch): square block, appended slots.  Tasks are GRABBED in list order
                          // by whichever CTA is free (atomicAdd), not dealt round robin: the balance adapts to the real task
                          // durations and to the moment the chain group runs out of square-block tasks
    DfSeg seg[2];         // appended blocks: seg[0] = general right-hand sides, seg[1] = identity block
    double* winv;
    int* flags;
    int epoch;
    int* info;
    int info_off;
    long long* trace;     // optional: 8 x int64 globaltimer stamps per task (development aid)
};

__device__ __forceinline__ void wait_flag(const int* f, int epoch) {
    while (ld_acquire(f) != epoch) __nanosleep(32);
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DF_STAMP(slot) do { if (g.trace && tid == 0) g.trace[((long)j * g.Trows + i) * 8 + (slot)] = gtime(); } while (0)
__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// Task schedule.  Two lists, each in column-major order: the square block's tiles and the appended tiles.
//   * CTAs [0, Gc) take square-block tasks until that list is exhausted (the per-column dependency chain never queues behind
//     a long appended task), then appended tasks;
//   * CTAs [Gc, grid) take appended tasks from the start.
// Every task only waits on tasks that precede it in its own list or on square-block tasks (appended tiles depend on the
// square block and on earlier tiles of their own row, never the other way round); the square-block list makes progress on
// its own, so the schedule cannot deadlock.
struct TaskIter {
    int phase = 0;
    int j = 0;            // column tracking of the square block's list (only ever moves forward)
    long q0 = 0;
};
// Called by ONE thread per CTA (the producer); the result is handed to the consumer warps through shared memory.
__device__ __forceinline__ bool next_task(const DfArgs& g, TaskIter& s, int& i, int& j) {
    const int R = g.Trows - g.T;
    for (;;) {
        if (s.phase == 0) {
            if ((int)blockIdx.x < g.Gc) {
                const long q = atomicAdd(g.counters + 0, 1);
                if (q < g.nspd) {
                    while (q >= s.q0 + (g.T - s.j)) { s.q0 += g.T - s.j; ++s.j; }
                    i = s.j + (int)(q - s.q0);
                    j = s.j;
                    return true;
                }
            }
            s.phase = 1;
            continue;
        }
        const long p = atomicAdd(g.counters + 1, 1);
        if (p >= g.napp) return false;
        j = (int)(p / R);
        i = g.T + (int)(p - (long)j * R);
        if (i >= g.id0 && j < i - g.id0) continue;        // identity block: zero tiles left of its diagonal
        return true;
    }
}

// One k-step of a DIAGONAL tile for the warp at (WM_, WN_): only the 8x8 blocks touching the lower triangle are issued,
// selected at compile time (no per-DMMA predicate arithmetic on the dependency chain's last k-block).
template <int WM_, int WN_>
__device__ __forceinline__ void diag_kstep(double (&acc)[8][4][2], uint32_t sa, uint32_t sb, int gq, int t) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double af[8][2], bf[4][2];
        load_frags<8, true>(af, sa, WM_ * 64, h, gq, t);
        load_frags<4, true>(bf, sb, WN_ * 32, h, gq, t);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (32 * WN_ + 8 * b <= 64 * WM_ + 8 * a + 7) dmma884(acc[a][b], af[a][u], bf[b][u]);
    }
}

__global__ void __launch_bounds__(384, 1)
potrf_dataflow_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmB2, const DfArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t bar_full = smem0 + RING_BYTES, bar_empty = bar_full + STAGES * 8;
    const uint32_t bar_wfull = bar_empty + STAGES * 8, bar_wempty = bar_wfull + WSTAGES * 8;
    const uint32_t bar_mdone = bar_wempty + WSTAGES * 8, bar_fdone = bar_mdone + 8;
    const uint32_t bar_task = bar_fdone + 8;                                   // two barriers: the producer has posted task t (slot t & 1)
    volatile int2* task_slot = reinterpret_cast<volatile int2*>(smem_gen + RING_BYTES + 224);   // two (i, j) slots

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int T = g.T;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 8);
        }
#pragma unroll
        for (int s = 0; s < WSTAGES; ++s) {
            mbar_init(bar_wfull + s * 8, 1);
            mbar_init(bar_wempty + s * 8, 8);
        }
        mbar_init(bar_mdone, 8);
        mbar_init(bar_fdone, 8);
        mbar_init(bar_task, 1);
        mbar_init(bar_task + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp >= 8) {
        // ============================== TMA producer ==============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8 && lane == 0) {
            uint32_t it = 0, wit = 0;
            int task_idx = 0, i, j;
            TaskIter ts;
            for (;;) {
                const bool more = next_task(g, ts, i, j);
                // slot t & 1 was last read for task t - 2, whose consumers are long past it (bar_fdone(t - 2) was waited for)
                task_slot[task_idx & 1].x = more ? i : -1;
                task_slot[task_idx & 1].y = j;
                mbar_arrive(bar_task + (task_idx & 1) * 8);
                if (!more) break;
                const int kstart = i >= g.id0 ? i - g.id0 : 0;                    // identity block: k starts at the tile's own row
                // A-operand rows of this task: the square block or one of the appended blocks
                const CUtensorMap* tmA = i < T ? &tmL : (i < g.id0 ? &tmB1 : &tmB2);
                const int arow = (i < T ? i : (i < g.id0 ? i - g.seg[0].row0 : i - g.seg[1].row0)) * TB;
                if (task_idx > 0) mbar_wait(bar_fdone, (task_idx - 1) & 1);      // consumers left the C region / leaf scratch
                for (int kb = kstart; kb < j; ++kb) {
                    wait_flag(g.flags + (long)i * T + kb, g.epoch);
                    if (i != j) wait_flag(g.flags + (long)j * T + kb, g.epoch);
                    fence_proxy_async();
#pragma unroll 1
                    for (int s = 0; s < 8; ++s, ++it) {
                        const int st = it % STAGES;
                        mbar_wait(bar_empty + st * 8, ((it / STAGES) & 1) ^ 1);
                        const uint32_t full = bar_full + st * 8;
                        mbar_expect_tx(full, STAGE_BYTES);
                        const uint32_t sa = smem0 + st * STAGE_BYTES;
                        tma_load_3d(sa, tmA, full, kb * TB + s * KSTEP, arow, 0);
                        tma_load_3d(sa + TILE_BYTES, &tmL, full, kb * TB + s * KSTEP, j * TB, 0);
                    }
                }
                if (i != j) {
                    wait_flag(g.flags + (long)j * T + j, g.epoch);                // L_jj factored, W_j written
                    fence_proxy_async();
                    mbar_wait(bar_mdone, task_idx & 1);                           // ring drained by every consumer warp
#pragma unroll 1
                    for (int s = 0; s < 8; ++s, ++wit) {
                        const int ws = wit % WSTAGES;
                        mbar_wait(bar_wempty + ws * 8, ((wit / WSTAGES) & 1) ^ 1);
                        const uint32_t full = bar_wfull + ws * 8;
                        mbar_expect_tx(full, TILE_BYTES);
                        tma_load_3d(smem0 + CREG_BYTES + ws * TILE_BYTES, &tmW, full, s * KSTEP, j * TB, 0);
                    }
                }
                ++task_idx;
            }
        }
        return;
    }

    // ============================== DMMA consumers (8 warps, 64 x 32 each) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // warp w runs on SMSP w % 4.  Column groups are dealt so that every SMSP hosts one "early" and one "late" column group
    // (wn = w for warps 0-3, 7 - w for warps 4-7): the triangular skips below (diagonal tiles need only their lower
    // triangle, W_j is lower triangular) then remove ~45 % of the DMMA work from EVERY SMSP instead of idling two of them.
    const int wm = warp >> 2, wn = (warp < 4) ? warp : 7 - warp;
    const int gq = lane >> 2, t = lane & 3;
    const int pg = (gq >> 1) | ((gq & 1) << 2);
    // bit (a*4+b): 8x8 block (a, b) of this warp's 64x32 tile touches the lower triangle of a diagonal tile
    uint32_t diag_mask = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (32 * wn + 8 * b <= 64 * wm + 8 * a + 7) diag_mask |= 1u << (a * 4 + b);
    double* S = reinterpret_cast<double*>(smem_gen) + leaf::GUARD;
    double* Wd = S + leaf::LB * leaf::LDS_;
    double* Tt = Wd + 4 * 32 * leaf::WLD;

    uint32_t it = 0, wit = 0;
    for (int task_idx = 0;; ++task_idx) {
        mbar_wait(bar_task + (task_idx & 1) * 8, (task_idx >> 1) & 1);
        const int i = task_slot[task_idx & 1].x, j = task_slot[task_idx & 1].y;
        if (i < 0) break;
        const int kstart = i >= g.id0 ? i - g.id0 : 0;
        // where this task's tile lives: the square block or an appended block (all of whose rows exist in memory)
        double* tile_ptr;
        long tile_ld;
        long row_lim;                              // rows of the block that may be read / written
        if (i < T) { tile_ptr = g.A + (long)i * TB * g.lda; tile_ld = g.lda; row_lim = (long)g.n - (long)i * TB; }
        else {
            const DfSeg& sg = g.seg[i < g.id0 ? 0 : 1];
            tile_ptr = sg.ptr + (long)(i - sg.row0) * TB * sg.ld; tile_ld = sg.ld; row_lim = TB;
        }
        tile_ptr += (long)j * TB;

        double acc[8][4][2];
        // The accumulators start from -A_ij, so that after the k-loop they hold -(A_ij - sum L_ik L_jk^T) and the epilogue on
        // the dependency chain needs no global loads: these 64 independent loads per lane are issued while the task is still
        // waiting for its operands.  Rows / columns beyond n are identity padded (diagonal tile) or zero.
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int r = wm * 64 + a * 8 + pg;
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cc = wn * 32 + b * 8 + t + 4 * e;
                    const long gc = (long)j * TB + cc;
                    double v = (i == j && r == cc) ? -1.0 : 0.0;
                    if (i >= g.id0) v = (j == i - g.id0 && r == cc) ? -1.0 : 0.0;        // identity block: nothing to read
                    else if (r < row_lim && gc < g.n) v = -__ldcg(tile_ptr + (long)r * tile_ld + cc);   // L2: the tile was written by earlier kernels on other SMs
                    acc[a][b][e] = v;
                }
        }
        DF_STAMP(0);
        const int niter = (j - kstart) * 8;
        const uint32_t mask = (i == j) ? diag_mask : 0xffffffffu;
        for (int n = 0; n < niter; ++n, ++it) {
            const int st = it % STAGES;
            mbar_wait(bar_full + st * 8, (it / STAGES) & 1);
            const uint32_t sa = smem0 + st * STAGE_BYTES, sb = sa + TILE_BYTES;
            if (mask == 0xffffffffu) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double af[8][2], bf[4][2];
                    load_frags<8, true>(af, sa, wm * 64, h, gq, t);
                    load_frags<4, true>(bf, sb, wn * 32, h, gq, t);
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int a = 0; a < 8; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) dmma884(acc[a][b], af[a][u], bf[b][u]);
                }
            } else {
                switch (warp) {      // (wm, wn) = (w >> 2, w < 4 ? w : 7 - w); warps 2 and 3 own no lower-triangle block
                    case 0: diag_kstep<0, 0>(acc, sa, sb, gq, t); break;
                    case 1: diag_kstep<0, 1>(acc, sa, sb, gq, t); break;
                    case 4: diag_kstep<1, 3>(acc, sa, sb, gq, t); break;
                    case 5: diag_kstep<1, 2>(acc, sa, sb, gq, t); break;
                    case 6: diag_kstep<1, 1>(acc, sa, sb, gq, t); break;
                    case 7: diag_kstep<1, 0>(acc, sa, sb, gq, t); break;
                    default: break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + st * 8);
        }
        bar_consumers();                        // every consumer warp has finished reading the ring
        if (lane == 0) mbar_arrive(bar_mdone);
        DF_STAMP(1);

        if (i == j) {
            // ---- C = -acc -> S (identity padded); factor + invert in shared memory ----
            if (tid < leaf::GUARD) reinterpret_cast<double*>(smem_gen)[tid] = 0.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int r = wm * 64 + a * 8 + pg;
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) S[r * leaf::LDS_ + wn * 32 + b * 8 + t + 4 * e] = -acc[a][b][e];
            }
            bar_consumers();
            DF_STAMP(2);
            const int nv = min(TB, g.n - j * TB);
            leaf::potf2_inverse<1>(S, Wd, Tt, g.A + ((long)j * TB) * g.lda + (long)j * TB, g.lda, nv, g.winv + (long)j * TB * TB, g.info,
                                   g.info_off + j * TB, tid);
            __threadfence();
            bar_consumers();
            if (tid == 0) {
                fence_proxy_async();
                st_release(g.flags + (long)j * T + j, g.epoch);
            }
            DF_STAMP(4);
        } else {
            // ---- C = -acc -> K-major swizzled sub-tiles; L_ij = C W_j^T ----
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int r = wm * 64 + a * 8 + pg;
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = wn * 32 + b * 8 + t + 4 * e;
                        const double v = -acc[a][b][e];
                        const int cw = cc & 15;
                        sts64(smem0 + (cc >> 4) * TILE_BYTES + r * 128 + ((((cw >> 1) ^ (r & 7)) << 4) | ((cw & 1) << 3)), v);
                        acc[a][b][e] = 0.0;
                    }
            }
            bar_consumers();
            DF_STAMP(2);
            for (int s = 0; s < 8; ++s, ++wit) {
                const int ws = wit % WSTAGES;
                mbar_wait(bar_wfull + ws * 8, (wit / WSTAGES) & 1);
                if (s == 0) DF_STAMP(3);
                const uint32_t sa = smem0 + s * TILE_BYTES, sb = smem0 + CREG_BYTES + ws * TILE_BYTES;
                // W_j[n][k] == 0 for k > n: k-step s only reaches the 8-column blocks with 16 s <= 32 wn + 8 b + 7
                if (16 * s <= 32 * wn + 31) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double af[8][2], bf[4][2];
                        load_frags<8, true>(af, sa, wm * 64, h, gq, t);
                        load_frags<4, true>(bf, sb, wn * 32, h, gq, t);
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (16 * s <= 32 * wn + 8 * b + 7) {
#pragma unroll
                                for (int u = 0; u < 2; ++u)
#pragma unroll
                                    for (int a = 0; a < 8; ++a) dmma884(acc[a][b], af[a][u], bf[b][u]);
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_wempty + ws * 8);
            }
            DF_STAMP(5);
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int r = wm * 64 + a * 8 + pg;
                if (r < row_lim) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int cc = wn * 32 + b * 8 + t + 4 * e;
                            if ((long)j * TB + cc < g.n) tile_ptr[(long)r * tile_ld + cc] = acc[a][b][e];    // ragged last tile column
                        }
                }
            }
            __threadfence();
            bar_consumers();
            if (tid == 0) {
                fence_proxy_async();
                st_release(g.flags + (long)i * T + j, g.epoch);
            }
            DF_STAMP(4);
        }
        fence_proxy_async();                    // generic smem traffic of this task before the next task's TMA writes
        if (lane == 0) mbar_arrive(bar_fdone);
    }
}

int encode_tiles(gpn_ctx* c, CUtensorMap* tm, const double* ptr, long ld, long rows, long cols) {
    if (((uintptr_t)ptr & 15) || (ld & 1)) return -3;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * 8 * (cuuint64_t)rows};
    cuuint32_t box[3] = {KSTEP, TB, 1}, estr[3] = {1, 1, 1};
    CUresult r = c->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        gpn_set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed (dataflow potrf)");
        return -1100 - (int)r;
    }
    return 0;
}

}   // namespace

// Lower Cholesky of the n x n matrix A in place (tiles strictly above the diagonal are left untouched, the strict upper
// triangle of the diagonal tiles is zeroed); inverted diagonal blocks go to winv (ceil(n/128) blocks of 128x128).
// Optional appended blocks (see the header comment): B (brows x n, overwritten by B L^-T; its rows beyond brows up to the
// next multiple of 128 must exist and be zero) and Wt (pad128(n) x n, any content: overwritten by L^-T in its upper tiles,
// i.e. the transposed inverse of the factor, lower tiles untouched; its previous content is never read).
int gpn_potrf_dataflow(gpn_ctx* c, int n, double* A, long lda, double* winv, int* d_info, int info_offset, double* B, long ldb, int brows,
                       double* Wt, long ldwt, int slot) {
    // slot 1: own flag / counter buffer, for a factorisation that runs concurrently with another one (side stream)
    c->winv_for = nullptr;
    if (slot < 0 || slot >= D6_SLOTS) return -13;
    DevBuf& fbuf = slot >= 2 ? c->flagsK[slot] : (slot ? c->flags2 : c->flags);
    int& fepoch = slot >= 2 ? c->epochK[slot] : (slot ? c->epoch2 : c->epoch);
    long& frows = slot >= 2 ? c->flags_rowsK[slot] : (slot ? c->flags_rows2 : c->flags_rows);
    const int T = (n + TB - 1) / TB;
    const int R1 = B ? (brows + TB - 1) / TB : 0, R2 = Wt ? T : 0;
    const int Trows = T + R1 + R2;
    long nt = 0;
    for (int j = 0; j < T; ++j) nt += Trows - j;
    const int ntasks = (int)nt;
    if (fbuf.cap < sizeof(int) * ((size_t)Trows * T + 4) || frows != (long)Trows * T) {     // ready flags + task counters
        GPN_TRY(gpn_buf_ensure(c, fbuf, sizeof(int) * ((size_t)Trows * T + 4)));
        frows = (long)Trows * T;
        GPN_CK(cudaMemsetAsync(fbuf.p, 0, fbuf.cap, c->stream));
        // the epoch keeps counting across shape changes: a flag of an earlier launch can never match, whatever the order in which
        // the clearing memset and earlier kernels on other streams touch the buffer
    }
    DfArgs a;
    a.A = A; a.lda = lda; a.n = n; a.T = T; a.ntasks = ntasks; a.winv = winv; a.flags = (int*)fbuf.p;
    a.Trows = Trows; a.id0 = T + R1;
    a.seg[0] = DfSeg{B, ldb, T, R1};
    a.seg[1] = DfSeg{Wt, ldwt, T + R1, R2};
    a.epoch = ++fepoch; a.info = d_info; a.info_off = info_offset; a.trace = (long long*)c->df_trace;
    CUtensorMap tl, tw, tb1, tb2;
    GPN_TRY(encode_tiles(c, &tl, A, lda, n, n));
    GPN_TRY(encode_tiles(c, &tw, winv, TB, (long)T * TB, TB));
    tb1 = tl; tb2 = tl;
    if (B) GPN_TRY(encode_tiles(c, &tb1, B, ldb, (long)R1 * TB, n));
    if (Wt) GPN_TRY(encode_tiles(c, &tb2, Wt, ldwt, (long)T * TB, n));
    static bool configured = false;
    if (!configured) {
        GPN_CK(cudaFuncSetAttribute(potrf_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM));
        configured = true;
    }
    int grid = ntasks < c->num_sms ? ntasks : c->num_sms;
    if (c->potrf_grid_cap > 0 && grid > c->potrf_grid_cap) grid = c->potrf_grid_cap;   // leave SMs to concurrent work
    // CTA groups (see next_task): with appended rows and enough tile columns for the chain to matter, a group of CTAs sized for
    // the square block's work within the chain's duration (x 1.5: a sweep at C3 is flat between 72 and 96 of 148 CTAs) keeps
    // the dependency chain to itself until the square block is done
    const int R = Trows - T;
    a.nspd = (long)T * (T + 1) / 2;
    a.napp = (long)T * R;
    a.Gc = grid;
    a.counters = (int*)fbuf.p + (size_t)Trows * T;
    GPN_CK(cudaMemsetAsync(a.counters, 0, 4 * sizeof(int), c->stream));
    if (R > 0 && T >= 8 && grid >= 32) {
        static const int gc_env = getenv("GPNET_B200_DF_CHAIN_CTAS") ? atoi(getenv("GPNET_B200_DF_CHAIN_CTAS")) : 0;
        const double kb_us = 17.0, col_us = 65.0;                          // one k-block on one SM; one step of the chain
        double spd_kb = 0.0;
        for (int j = 0; j < T; ++j) spd_kb += (double)(T - j - 1) * j + 0.55 * j + 2.0 * (T - j);   // + epilogue / leaf equivalents
        int gc = (int)(spd_kb * kb_us / (T * col_us) * 1.5) + 2;
        if (gc < 24) gc = 24;
        if (gc > (2 * grid) / 3) gc = (2 * grid) / 3;
        if (gc_env > 0 && gc_env < grid) gc = gc_env;
        a.Gc = gc;
    }
    void* args[] = {(void*)&tl, (void*)&tw, (void*)&tb1, (void*)&tb2, (void*)&a};
    // executed flops: an off-diagonal task runs j k-blocks of 128^3 MACs, a diagonal one only the 136 of 256 8x8 blocks that
    // touch the lower triangle; the in-CTA trsm skips the zero k-steps of W_j (36 of 64); leaf = factor + inverse; appended rows
    // add their own tasks (identity block: k from the tile's own row)
    double kb = 0.0;
    for (int j = 0; j < T; ++j) {
        kb += (double)(T - j - 1) * j + j * (136.0 / 256.0) + (T - j - 1) * (36.0 / 64.0) + 1.0 / 3.0;
        kb += (double)R1 * (j + 36.0 / 64.0);
        if (R2) kb += 0.5 * j * (j + 1.0) + (j + 1) * (36.0 / 64.0);          // identity rows r <= j: (j - r) k-blocks each
    }
    const double fl = 2.0 * kb * TB * TB * TB;
    ProfEvent pe;
    if (c->profile) {
        cudaEventCreate(&pe.e0);
        cudaEventCreate(&pe.e1);
        pe.flops = fl;
        pe.big = false;
        pe.kind = 2;
        cudaEventRecord(pe.e0, c->stream);
    }
    static const bool noncoop = getenv("GPNET_B200_DF_NONCOOP") != nullptr;
    if (noncoop) GPN_CK(cudaLaunchKernel((const void*)potrf_dataflow_kernel, dim3(grid), dim3(384), args, DF_SMEM, c->stream));
    else GPN_CK(cudaLaunchCooperativeKernel((const void*)potrf_dataflow_kernel, dim3(grid), dim3(384), args, DF_SMEM, c->stream));
    if (c->profile) {
        cudaEventRecord(pe.e1, c->stream);
        c->prof_events.push_back(pe);
    }
    c->launches++;
    c->flops += fl;
    return 0;
}
