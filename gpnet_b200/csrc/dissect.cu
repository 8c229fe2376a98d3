// Whole-graph K-way dissection for the regularized-Laplacian regressor likelihood (route level 6; driver: gp.cu).
//
// Replaces, for graphs with thin separators, everything GPnetRegressor.logPosterior (GPnetRegressor.py:136-150) and the
// intended gradient (old_implementation/GPR.py:52-64) need from K = (I + beta L~)^-1 (GPnet.py:341-342):
//     K_tt^-1 = S = M_tt - M_to M_oo^-1 M_ot                       (M = I + beta L~, t = training nodes, o = the others)
//     log det S = log det M - log det M_oo
//     x^T S y   = x^T M_tt y - (M_ot x)^T M_oo^-1 (M_ot y)
//     tr(S dK_tt/dtheta_0) = tr(M^-1) - tr(M_oo^-1) - n ,   (S x)^T (K^2)_tt (S y) = u_x . u_y + x . y,  u_x = M_oo^-1 M_ot x
// so that NO dense n x n Schur complement is formed or factored.  With the nodes numbered [P_1 | .. | P_K | Sep] (level-set
// dissection: no edge joins two parts) M = [D C; C^T E], D = blkdiag(M_kk), and with every part numbered [others | training] the
// leading block of M_kk is the matching block of M_oo: ONE augmented factorisation [M_kk; I] per part (K concurrent dataflow
// Cholesky kernels, each a K-times shorter dependency chain) yields L_k and W_k = L_k^-1 for both M and M_oo:
//     G = W C (sparse columns),  T = E - G^T G,   log det M = 2 sum log diag L_k + log det T,
//     tr(M^-1) = sum ||W_k||_F^2 + tr(T^-1 (I + U^T U)),   U = D^-1 C = W^T G
// and the same with every operand cut to its leading (other-node) block for M_oo.
//
// This file: the host-side analysis (ensure), the HBM-bound kernels of the route and their launchers.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int TPB = 256;

struct D6Desc {
    int K, s, so, ND;
    int off[D6_MAXP + 1];
    int nk[D6_MAXP], ok[D6_MAXP];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v) {      // result valid in thread 0
    __shared__ double sh[32];
    __syncthreads();
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
        v = warp_sum(v);
    }
    return v;
}
__device__ __forceinline__ int part_of(const D6Desc& d, int r) {
    int k = 0;
    while (k + 1 < d.K && r >= d.off[k + 1]) ++k;
    return k;
}

// Zero fill of what the factorisations read as dense: the lower 128-tiles of every diagonal block M_kk and the separator
// block E (everything else of M is only ever read at the entries the scatter writes).  One row per blockIdx.y.
__global__ void d6_fill_zero_kernel(D6Desc d, double* __restrict__ M, long ld) {
    const int r = blockIdx.y;
    int c0, c1;
    if (r < d.ND) {
        const int k = part_of(d, r);
        const int i = r - d.off[k];
        c0 = d.off[k];
        c1 = d.off[k] + min(d.nk[k], (i / 128 + 1) * 128);
    } else {
        c0 = d.ND;
        c1 = d.ND + d.s;
    }
    double* row = M + (long)r * ld;
    const int nv = (c1 - c0) / 2;                       // c0 is even
    double2* v = reinterpret_cast<double2*>(row + c0);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nv; x += gridDim.x * blockDim.x) v[x] = make_double2(0.0, 0.0);
    if (((c1 - c0) & 1) && blockIdx.x == 0 && threadIdx.x == 0) row[c1 - 1] = 0.0;
}

// One warp per permuted node i.  Other node: b_1[i] = sum_{j in t} M_ij, b_t[i] = sum_{j in t} M_ij t_j (the rows of M_ot x).
// Training node: its row of M_tt times 1 and t, reduced into out[0..4] = 1^T M_tt 1, 1^T M_tt t, t^T M_tt t, 1^T t, t^T t.
// Entries as the scatter writes them: M_ij = beta (dinv_i ((-w_ij) dinv_j)), M_ii = 1 + beta (dinv_i ((deg_i - w_ii) dinv_i)).
__global__ void d6_coupling_kernel(int N, const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ w,
                                   const double* __restrict__ dinv, const int* __restrict__ trank, const double* __restrict__ t, double beta,
                                   double* __restrict__ b1, double* __restrict__ bt, double* __restrict__ out) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const double di = dinv[i], deg = dinv[N + i];
    const int ri = trank[i];
    double a1 = 0.0, at = 0.0, self = 0.0;
    for (int e = indptr[i] + lane; e < indptr[i + 1]; e += 32) {
        const int j = indices[e];
        const double a = w ? w[e] : 1.0;
        if (j == i) { self = a; continue; }
        const int rj = trank[j];
        if (rj < 0) continue;
        const double m = beta * (di * ((-a) * dinv[j]));
        a1 += m;
        at += m * t[rj];
    }
    a1 = warp_sum(a1);
    at = warp_sum(at);
    self = warp_sum(self);
    if (lane != 0) return;
    if (ri < 0) {
        b1[i] = a1;
        bt[i] = at;
    } else {
        b1[i] = 0.0;
        bt[i] = 0.0;
        const double ti = t[ri];
        const double mii = 1.0 + beta * (di * ((deg - self) * di));
        const double r1 = mii + a1, rt = mii * ti + at;
        atomicAdd(out + 0, r1);
        atomicAdd(out + 1, rt);
        atomicAdd(out + 2, ti * rt);
        atomicAdd(out + 3, ti);
        atomicAdd(out + 4, ti * ti);
    }
}

// One block per row r < ND of the parts: ||W_k||_F^2 (upper triangle of W_k^T = WT_k) and sum log diag(L_k), each also for the
// leading other-node block.  out[0..3] += fro, fro_o, logdet, logdet_o.
__global__ void d6_factor_reduce_kernel(D6Desc d, const double* __restrict__ L, const double* __restrict__ WT, long ld, double* __restrict__ out) {
    const int r = blockIdx.x;
    const int k = part_of(d, r);
    const int i = r - d.off[k], n = d.nk[k], o = d.ok[k];
    const double* row = WT + (long)r * ld + d.off[k];
    double f = 0.0, fo = 0.0;
    for (int j = i + threadIdx.x; j < n; j += blockDim.x) {
        const double v = row[j];
        f += v * v;
        if (j < o) fo += v * v;          // i <= j < o
    }
    f = block_sum(f);
    fo = block_sum(fo);
    if (threadIdx.x == 0) {
        const double lg = log(L[(long)r * ld + r]);
        atomicAdd(out + 0, f);
        atomicAdd(out + 2, lg);
        if (i < o) {
            atomicAdd(out + 1, fo);
            atomicAdd(out + 3, lg);
        }
    }
}

// G^T = (W C)^T: row r of G^T is sum_e C[c_e, r] * (row c_e of W_k^T) over the entries of separator node r inside part k
// (W_k^T upper triangular: row c contributes to columns j >= c only, and every c_e >= c0_k: the columns below c0_k are zero
// and never stored).  GTo = the same with other-node rows / columns only.  One launch per part, so that it can follow the
// part's factorisation on the part's own stream.
__global__ void d6_sep_spmm_kernel(D6Desc d, int k, int c0, const int* __restrict__ pp, const int* __restrict__ idx, const double* __restrict__ sw,
                                   double beta, const double* __restrict__ WT, long ld, const int* __restrict__ trank, double* __restrict__ GT,
                                   double* __restrict__ GTo, long ldg) {
    const int r = blockIdx.y;
    const int j = d.off[k] + c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.off[k + 1]) return;
    const int e0 = pp[r * (d.K + 1) + k], e1 = pp[r * (d.K + 1) + k + 1];
    double acc = 0.0;
    for (int e = e0; e < e1; ++e) {
        const int c = idx[e];
        if (c <= j) acc += (beta * sw[e]) * WT[(long)c * ld + j];
    }
    GT[(long)r * ldg + j] = acc;
    if (r < d.so) GTo[(long)r * ldg + j] = trank[j] < 0 ? acc : 0.0;
}

// T = E - sum_b C_b on the lower 128-tiles (the split-K partial products G^T G were computed on their lower tiles only)
__global__ void d6_t_reduce_kernel(int s, const double* __restrict__ E, long lde, const double* __restrict__ Cb, long lds, int nb,
                                   double* __restrict__ T) {
    const int r = blockIdx.y;
    const int lim = min((r / 128 + 1) * 128, s);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < lim; c += gridDim.x * blockDim.x) {
        const int rr = max(r, c), cc = min(r, c);          // partial products are valid on and below the diagonal (any tile size)
        double acc = E[(long)rr * lde + cc];
        for (int b = 0; b < nb; ++b) acc -= Cb[((long)b * lds + rr) * lds + cc];
        T[(long)r * lds + c] = acc;
    }
}

// After the augmented factorisation [T; I; U]: out[0] += ||W_T||_F^2 (upper triangle of WTt) + ||Y||_F^2 (Y = U L_T^-T, rows x s),
// out[1] += sum log diag(L_T).  Blocks [0, s): rows of WTt; blocks [s, s + rows): rows of Y.
__global__ void d6_sep_reduce_kernel(int s, const double* __restrict__ Lt, long ldl, const double* __restrict__ WTt, long ldw,
                                     const double* __restrict__ Y, long ldy, int rows, double* __restrict__ out) {
    const int b = blockIdx.x;
    double acc = 0.0;
    if (b < s) {
        if (WTt)
            for (int c = b + threadIdx.x; c < s; c += blockDim.x) { const double v = WTt[(long)b * ldw + c]; acc += v * v; }
    } else {
        const double* row = Y + (long)(b - s) * ldy;
        for (int c = threadIdx.x; c < s; c += blockDim.x) { const double v = row[c]; acc += v * v; }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) {
        if (acc != 0.0) atomicAdd(out + 0, acc);
        if (b < s) atomicAdd(out + 1, log(Lt[(long)b * ldl + b]));
    }
}

// y = W_k^o x for every part (W^o = leading other-node block of W_k = L_k^-1, read from its transpose WT_k):
// y_i = sum_{j <= i} WT[j][i] x_j.  Two right-hand sides.  Block (32, 8): 32 columns, rows split over 8 warps.
__global__ void d6_w_mul_kernel(D6Desc d, const double* __restrict__ WT, long ld, const double* __restrict__ xa, const double* __restrict__ xb,
                                double* __restrict__ ya, double* __restrict__ yb) {
    __shared__ double sa[8][33], sb[8][33];
    const int k = blockIdx.y;
    const int o = d.ok[k], base = d.off[k];
    const int i = blockIdx.x * 32 + threadIdx.x;
    double a = 0.0, b = 0.0;
    if (blockIdx.x * 32 < o) {
        const int ilim = min(i, o - 1);
        const int jmax = min(blockIdx.x * 32 + 31, o - 1);
        for (int j = threadIdx.y; j <= jmax; j += 8) {
            if (j <= ilim && i < o) {
                const double v = WT[(long)(base + j) * ld + base + i];
                a += v * xa[base + j];
                b += v * xb[base + j];
            }
        }
    }
    sa[threadIdx.y][threadIdx.x] = a;
    sb[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0 && i < o) {
#pragma unroll
        for (int q = 1; q < 8; ++q) { a += sa[q][threadIdx.x]; b += sb[q][threadIdx.x]; }
        ya[base + i] = a;
        yb[base + i] = b;
    }
}
// y = (W_k^o)^T v for every part: y_j = sum_{j <= i < o} WT[j][i] v_i.  One warp per row; training rows get 0.
__global__ void d6_wt_mul_kernel(D6Desc d, const double* __restrict__ WT, long ld, const double* __restrict__ va, const double* __restrict__ vb,
                                 double* __restrict__ ya, double* __restrict__ yb) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= d.ND) return;
    const int k = part_of(d, r);
    const int j = r - d.off[k], o = d.ok[k], base = d.off[k];
    double a = 0.0, b = 0.0;
    if (j < o) {
        const double* row = WT + (long)r * ld + base;
        for (int i = j + lane; i < o; i += 32) {
            const double v = row[i];
            a += v * va[base + i];
            b += v * vb[base + i];
        }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { ya[r] = a; yb[r] = b; }
}
// r_S = b_S - C_o^T y for the other-node separator rows (one warp per row)
__global__ void d6_sep_gather_kernel(D6Desc d, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ sw, double beta,
                                     const int* __restrict__ trank, const double* __restrict__ ba, const double* __restrict__ bb,
                                     const double* __restrict__ ya, const double* __restrict__ yb, double* __restrict__ ra, double* __restrict__ rb) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= d.so) return;
    double a = 0.0, b = 0.0;
    for (int e = ptr[r] + lane; e < ptr[r + 1]; e += 32) {
        const int c = idx[e];
        if (trank[c] >= 0) continue;
        const double m = beta * sw[e];
        a += m * ya[c];
        b += m * yb[c];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { ra[r] = ba[d.ND + r] - a; rb[r] = bb[d.ND + r] - b; }
}
// w = b_D - C_o u_S on the other-node rows of the parts: w starts as b, every separator row scatters its contribution
__global__ void d6_sep_scatter_kernel(D6Desc d, const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ sw, double beta,
                                      const int* __restrict__ trank, const double* __restrict__ ua, const double* __restrict__ ub,
                                      double* __restrict__ wa, double* __restrict__ wb) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= d.so) return;
    const double xa = ua[r], xb = ub[r];
    for (int e = ptr[r] + lane; e < ptr[r + 1]; e += 32) {
        const int c = idx[e];
        if (trank[c] >= 0) continue;
        const double m = beta * sw[e];
        atomicAdd(wa + c, -m * xa);
        atomicAdd(wb + c, -m * xb);
    }
}

template <class T>
int up(gpn_ctx* c, DevBuf& b, const std::vector<T>& v) {
    GPN_TRY(gpn_buf_ensure(c, b, sizeof(T) * (v.size() + 2)));
    if (!v.empty()) {
        GPN_CK(cudaMemcpyAsync(b.p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, c->stream));
        GPN_CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

D6Desc make_desc(const Dissect6& q) {
    D6Desc d;
    d.K = q.K; d.s = q.s; d.so = q.so; d.ND = q.ND;
    for (int k = 0; k <= D6_MAXP; ++k) d.off[k] = q.off[k];
    for (int k = 0; k < D6_MAXP; ++k) { d.nk[k] = q.nk[k]; d.ok[k] = q.ok[k]; }
    return d;
}

}   // namespace

// Host analysis, once per (graph, training set): BFS level structure of the whole graph from a pseudo-peripheral node, K - 1 cut
// levels at the quantiles of the node count (a level separates the levels below it from those above it exactly: edges only
// join equal or adjacent levels), separator nodes without a neighbour in the next level pushed back into the part below.
// usable = false (the caller falls back to level 5) when the graph is too small or its separators are fat.
// Minimum vertex separator between BFS level l and level l + 1 of the subgraph whose nodes satisfy scope(v): the edges between
// the two levels form a bipartite graph, and by Koenig's theorem its minimum vertex cover (from a maximum matching: Kuhn's
// augmenting paths, a few hundred nodes) is the smallest node set whose removal disconnects everything at or below level l
// from everything at or above level l + 1.  Typically 25-40 % fewer nodes than the boundary nodes of one level.
template <class Scope>
static void level_cut_cover(const gpn_ctx* c, const std::vector<int>& level, const std::vector<int>& nodes_l, Scope scope, int l,
                            std::vector<int>& cover) {
    cover.clear();
    std::vector<int> A, B;
    std::vector<int> bidx;                       // node -> index in B (lazy, via map vector sized on demand)
    std::vector<std::vector<int>> adj;
    std::vector<std::pair<int, int>> bmap;       // (node, index) sorted for lookup
    for (int u : nodes_l) {
        std::vector<int> nb;
        for (int e = c->h_indptr[u]; e < c->h_indptr[u + 1]; ++e) {
            const int v = c->h_indices[e];
            if (scope(v) && level[v] == l + 1) nb.push_back(v);
        }
        if (!nb.empty()) { A.push_back(u); adj.push_back(nb); }
    }
    for (auto& nb : adj)
        for (int v : nb) bmap.push_back({v, 0});
    std::sort(bmap.begin(), bmap.end());
    bmap.erase(std::unique(bmap.begin(), bmap.end()), bmap.end());
    for (size_t i = 0; i < bmap.size(); ++i) { bmap[i].second = (int)i; B.push_back(bmap[i].first); }
    auto bindex = [&](int v) { return std::lower_bound(bmap.begin(), bmap.end(), std::make_pair(v, 0))->second; };
    for (auto& nb : adj)
        for (int& v : nb) v = bindex(v);
    const int na = (int)A.size(), nbn = (int)B.size();
    std::vector<int> matchA(na, -1), matchB(nbn, -1), seen(nbn, -1);
    // Kuhn's algorithm (iterative DFS would be overkill at these sizes: recursion depth <= na)
    struct Aug {
        const std::vector<std::vector<int>>& adj;
        std::vector<int>&matchA, &matchB, &seen;
        bool run(int a, int stamp) {
            for (int b : adj[a]) {
                if (seen[b] == stamp) continue;
                seen[b] = stamp;
                if (matchB[b] < 0 || run(matchB[b], stamp)) { matchA[a] = b; matchB[b] = a; return true; }
            }
            return false;
        }
    } aug{adj, matchA, matchB, seen};
    for (int a = 0; a < na; ++a) aug.run(a, a);
    // Koenig: Z = vertices reachable from unmatched A-vertices by alternating paths; cover = (A \ Z) + (B & Z)
    std::vector<char> za(na, 0), zb(nbn, 0);
    std::vector<int> stack;
    for (int a = 0; a < na; ++a)
        if (matchA[a] < 0) { za[a] = 1; stack.push_back(a); }
    while (!stack.empty()) {
        const int a = stack.back();
        stack.pop_back();
        for (int b : adj[a]) {
            if (zb[b] || matchA[a] == b) continue;
            zb[b] = 1;
            const int a2 = matchB[b];
            if (a2 >= 0 && !za[a2]) { za[a2] = 1; stack.push_back(a2); }
        }
    }
    for (int a = 0; a < na; ++a)
        if (!za[a]) cover.push_back(A[a]);
    for (int b = 0; b < nbn; ++b)
        if (zb[b]) cover.push_back(B[b]);
}

static int d6_analyse(gpn_ctx* c, bool allow_cells);
int gpn_d6_ensure(gpn_ctx* c) {
    Dissect6& q = c->d6;
    if (q.built) return 0;
    q.built = true;
    q.usable = false;
    GPN_TRY(d6_analyse(c, true));
    if (!q.usable) GPN_TRY(d6_analyse(c, false));      // the strips do not split into cells (or the result is poor): plain strips
    return 0;
}
static int d6_analyse(gpn_ctx* c, bool allow_cells) {
    Dissect6& q = c->d6;
    q.usable = false;
    const int N = c->N, n = c->n_train;
    static const int k_env = getenv("GPNET_B200_D6_PARTS") ? atoi(getenv("GPNET_B200_D6_PARTS")) : 0;
    const int forced = c->d6_parts > 0 ? c->d6_parts : k_env;      // forced: tests exercise the route on small graphs too
    // Parts = `strips` BFS strips of the whole graph, each cut again into `cells` pieces by the level sets of a BFS INSIDE the strip
    // (its levels run along the strip): on a planar-like graph the separator of a strips x cells grid of ~N / (strips cells)-node
    // parts is hardly larger than that of strips x 1, while the dependency chain of a part's factorisation shrinks with its size:
    // N = 8192: 5 x 1 parts of ~1500 nodes + 870 separator nodes = 13 + 7 tile columns on the critical path, 4 x 2 parts of ~950
    // + 636 = 8-10 + 5 (and half of the flops).  Forced part counts (tests, GPNET_B200_D6_PARTS) are plain strips.
    static const int cells_env = getenv("GPNET_B200_D6_CELLS") ? atoi(getenv("GPNET_B200_D6_CELLS")) : -1;
    int K, cells = 1;
    if (forced > 0) {
        K = forced;
    } else {
        // ~1000 nodes per part, two pieces per strip when that gives at least four parts.  Measured at N = 8192 with the final separator
        // rule and 32 hardware queues (evals/s with one / three lanes): strips x cells = 3x2: 610 / 907, 2x3: 580 / 916, 4x2: 657 / 981
        // (default), 3x3: 635 / 931, 5x2: 600 / 837, 4x3: 622 / 834, 6x2: 553 / 758 -- smaller parts shorten the chains, but every extra
        // cut adds separator nodes, and the separator's work grows with its square
        const int want = std::max(2, (N + 512) / 1024);
        cells = want >= 4 ? 2 : 1;
        K = std::max(2, (int)std::lround((double)want / cells));
        if (cells > 1 && K * cells > D6_MAXP) K = D6_MAXP / cells;
        static const int strips_env = getenv("GPNET_B200_D6_STRIPS") ? atoi(getenv("GPNET_B200_D6_STRIPS")) : -1;
        if (strips_env >= 2) K = strips_env;
        if (cells_env >= 1) cells = cells_env;
        if (!allow_cells) cells = 1;
        while (K * cells > D6_MAXP && cells > 1) --cells;
        if (cells == 1 && !(strips_env >= 2 && allow_cells)) K = (N + 820) / 1640;                    // strips only: ~1600 nodes per part (flat between 5 and 6 at N = 8192)
    }
    if (K > (cells > 1 ? D6_MAXP / cells : 8)) K = cells > 1 ? D6_MAXP / cells : 8;
    const int min_nodes = forced > 0 ? 64 : 2048, min_part = forced > 0 ? 8 : 128;
    if (K < 2 || N < min_nodes || n <= 0 || n >= N) return 0;
    std::vector<char> is_train(N, 0);
    for (int v : c->h_train) is_train[v] = 1;
    std::vector<int> lev(N), queue;
    auto bfs = [&](int root) {
        std::fill(lev.begin(), lev.end(), -1);
        queue.clear();
        queue.push_back(root);
        lev[root] = 0;
        for (size_t h = 0; h < queue.size(); ++h) {
            const int u = queue[h];
            for (int e = c->h_indptr[u]; e < c->h_indptr[u + 1]; ++e) {
                const int v = c->h_indices[e];
                if (lev[v] < 0) { lev[v] = lev[u] + 1; queue.push_back(v); }
            }
        }
        return queue.back();
    };
    int root = 0;
    {   // start in the largest component: the one a BFS from the highest-degree node reaches is a good guess
        int best = 0;
        for (int v = 0; v < N; ++v)
            if (c->h_indptr[v + 1] - c->h_indptr[v] > c->h_indptr[best + 1] - c->h_indptr[best]) best = v;
        root = best;
    }
    for (int sweep = 0; sweep < 2; ++sweep) root = bfs(root);
    bfs(root);
    const int reached = (int)queue.size();
    int nlev = 0;
    for (int v : queue) nlev = std::max(nlev, lev[v] + 1);
    if (nlev < (forced > 0 ? 2 * K + 1 : 3 * K)) return 0;
    std::vector<int> cnt(nlev, 0);
    for (int v : queue) cnt[lev[v]]++;
    // Cut k separates the levels <= l_k from the levels > l_k; l_k is the quantile level of the node count (optionally the thinnest cut in a
    // window around it), and the separator itself is the
    // minimum vertex cover of the edges between level l_k and l_k + 1 (level_cut_cover), not a whole boundary layer.
    // The separator of a cut is the minimum vertex cover of the edges between level l_k and level l_k + 1 (level_cut_cover), not the
    // whole boundary layer of level l_k (GPNET_B200_D6_THIN=0: the first round-2 version): 674 -> 562 separator nodes and 26.8 ->
    // 24.1 GFLOP per evaluation at N = 8192; 822 -> 875 evals/s with three lanes, 569 -> 588 with one (with 8 hardware queues it had
    // measured slower, like every configuration with more concurrent work).
    static const bool thin = !(getenv("GPNET_B200_D6_THIN") && atoi(getenv("GPNET_B200_D6_THIN")) == 0);
    // Levels searched on either side of the quantile level for the THINNEST cut (smallest vertex cover), as long as the cut stays within
    // 20 % of a part's node count of the quantile: the work of the separator phase and of the per-part products falls with the
    // separator (N = 8192: 562 -> 502 nodes, 24.1 -> 21.8 GFLOP per evaluation), which is what counts once three evaluations in flight
    // fill the chip (873 -> 926 evals/s); a single evaluation alone is bound by its longest part instead (parts of 1066..1698 nodes
    // instead of 1212..1356) and pays for it.  GPNET_B200_D6_WINDOW=0 keeps the balanced cuts.
    static const int window = getenv("GPNET_B200_D6_WINDOW") ? atoi(getenv("GPNET_B200_D6_WINDOW")) : 2;
    std::vector<std::vector<int>> bylev(nlev);
    for (int v : queue) bylev[lev[v]].push_back(v);
    std::vector<int> cuts;
    std::vector<char> in_sep(N, 0);
    {
        int l = 0;
        long cum = 0;
        std::vector<int> cover, best;
        auto whole = [](int) { return true; };
        for (int k = 1; k < K; ++k) {
            const long target = (long)reached * k / K;
            while (l < nlev - 1 && cum + cnt[l] < target) cum += cnt[l++];
            const int lo_ok = cuts.empty() ? 1 : cuts.back() + 2;
            if (l < lo_ok || l <= 0 || l >= nlev - 1) return 0;
            int lbest = l;
            if (thin) level_cut_cover(c, lev, bylev[l], whole, l, best);
            if (thin && forced <= 0) {
                for (int dl = -window; dl <= window; ++dl) {
                    const int lc = l + dl;
                    if (dl == 0 || lc < lo_ok || lc <= 0 || lc >= nlev - 1) continue;
                    long below = 0;
                    for (int x = 0; x < lc; ++x) below += cnt[x];
                    if (std::labs(below - target) * 5 > reached / K) continue;          // keep the parts within 20 % of their share
                    level_cut_cover(c, lev, bylev[lc], whole, lc, cover);
                    if (cover.size() < best.size()) { best = cover; lbest = lc; }
                }
            }
            if (!thin) {      // round-2 first version: the boundary nodes of level l (those with a neighbour in level l + 1)
                best.clear();
                for (int v : bylev[l]) {
                    bool up_nb = false;
                    for (int e = c->h_indptr[v]; e < c->h_indptr[v + 1] && !up_nb; ++e) up_nb = lev[c->h_indices[e]] == l + 1;
                    if (up_nb) best.push_back(v);
                }
            }
            cuts.push_back(lbest);
            for (int v : best) in_sep[v] = 1;
            // the running count continues from the quantile level (the window may have moved the cut by up to two levels)
        }
    }
    // part of every node (unreached components join the last part: they have no edge to anything else); -1 = separator
    std::vector<int> part(N);
    for (int v = 0; v < N; ++v) {
        if (lev[v] < 0) { part[v] = K - 1; continue; }
        if (in_sep[v]) { part[v] = -1; continue; }
        int p = 0;
        for (int l : cuts)
            if (lev[v] > l) ++p;
        part[v] = p;
    }
    const int strips = K;
    if (cells > 1) {
        // second stage: inside every strip, BFS levels of the INDUCED subgraph from a pseudo-peripheral node of the strip; cells - 1
        // cut levels at the quantiles; cut nodes without a neighbour in the next level stay in the piece below.  Pieces of one
        // strip are separated by the cut levels (edges of the induced subgraph join equal or adjacent levels only), different
        // strips by the first stage.  Nodes the BFS does not reach (other components of the strip) join the strip's last piece.
        std::vector<int> lev2(N, -1), newpart(N, -1), q2;
        bool ok = true;
        for (int p = 0; p < strips && ok; ++p) {
            std::vector<int> nodes;
            for (int v = 0; v < N; ++v)
                if (part[v] == p) nodes.push_back(v);
            if (nodes.empty()) { ok = false; break; }
            auto bfs2 = [&](int root) {
                for (int v : nodes) lev2[v] = -1;
                q2.clear();
                q2.push_back(root);
                lev2[root] = 0;
                for (size_t h = 0; h < q2.size(); ++h) {
                    const int u = q2[h];
                    for (int e = c->h_indptr[u]; e < c->h_indptr[u + 1]; ++e) {
                        const int v = c->h_indices[e];
                        if (part[v] == p && lev2[v] < 0) { lev2[v] = lev2[u] + 1; q2.push_back(v); }
                    }
                }
                return q2.back();
            };
            int r2 = nodes[0];
            for (int v : nodes)
                if (c->h_indptr[v + 1] - c->h_indptr[v] > c->h_indptr[r2 + 1] - c->h_indptr[r2]) r2 = v;
            for (int sweep = 0; sweep < 2; ++sweep) r2 = bfs2(r2);
            bfs2(r2);
            const int reached2 = (int)q2.size();
            int nl2 = 0;
            for (int v : q2) nl2 = std::max(nl2, lev2[v] + 1);
            if (nl2 < 3 * cells) { ok = false; break; }
            std::vector<int> cnt2(nl2, 0);
            for (int v : q2) cnt2[lev2[v]]++;
            std::vector<std::vector<int>> bylev2(nl2);
            for (int v : q2) bylev2[lev2[v]].push_back(v);
            auto in_strip = [&](int v) { return part[v] == p; };
            std::vector<int> cuts2, cover, best;
            std::vector<int> sep2;
            int l = 0;
            long cum = 0;
            for (int k = 1; k < cells && ok; ++k) {
                const long target = (long)reached2 * k / cells;
                while (l < nl2 - 1 && cum + cnt2[l] < target) cum += cnt2[l++];
                const int lo_ok = cuts2.empty() ? 1 : cuts2.back() + 2;
                if (l < lo_ok || l <= 0 || l >= nl2 - 1) { ok = false; break; }
                int lbest = l;
                if (thin) {
                    level_cut_cover(c, lev2, bylev2[l], in_strip, l, best);
                    for (int dl = -window; dl <= window; ++dl) {
                        const int lc = l + dl;
                        if (dl == 0 || lc < lo_ok || lc <= 0 || lc >= nl2 - 1) continue;
                        long below = 0;
                        for (int x = 0; x < lc; ++x) below += cnt2[x];
                        if (std::labs(below - target) * 5 > reached2 / cells) continue;
                        level_cut_cover(c, lev2, bylev2[lc], in_strip, lc, cover);
                        if (cover.size() < best.size()) { best = cover; lbest = lc; }
                    }
                } else {
                    best.clear();
                    for (int v : bylev2[l]) {
                        bool up_nb = false;
                        for (int e = c->h_indptr[v]; e < c->h_indptr[v + 1] && !up_nb; ++e) {
                            const int u = c->h_indices[e];
                            up_nb = part[u] == p && lev2[u] == l + 1;
                        }
                        if (up_nb) best.push_back(v);
                    }
                }
                cuts2.push_back(lbest);
                sep2.insert(sep2.end(), best.begin(), best.end());
            }
            if (!ok) break;
            for (int v : nodes) {
                if (lev2[v] < 0) { newpart[v] = p * cells + cells - 1; continue; }
                int piece = 0;
                for (int lc : cuts2)
                    if (lev2[v] > lc) ++piece;
                newpart[v] = p * cells + piece;
            }
            for (int v : sep2) newpart[v] = -1;
        }
        if (ok) {
            for (int v = 0; v < N; ++v) part[v] = part[v] < 0 ? -1 : newpart[v];
            K = strips * cells;
        } else {
            return 0;             // the strips do not split: the caller retries with plain strips
        }
    }
    std::vector<std::vector<int>> po(K), pt(K);
    std::vector<int> so_nodes, st_nodes;
    for (int v = 0; v < N; ++v) {
        if (part[v] < 0) (is_train[v] ? st_nodes : so_nodes).push_back(v);
        else (is_train[v] ? pt[part[v]] : po[part[v]]).push_back(v);
    }
    for (int k = 0; k < K; ++k) {     // even part sizes: every block starts at a 16-byte aligned column
        if ((po[k].size() + pt[k].size()) & 1) {
            if (!pt[k].empty()) { st_nodes.push_back(pt[k].back()); pt[k].pop_back(); }
            else { so_nodes.push_back(po[k].back()); po[k].pop_back(); }
        }
    }
    std::sort(so_nodes.begin(), so_nodes.end());
    std::sort(st_nodes.begin(), st_nodes.end());
    {   // inside each group: interior nodes first, then the nodes with a neighbour in the separator (ascending ids in both)
        std::vector<char> in_sep(N, 0), bnd(N, 0);
        for (int v : so_nodes) in_sep[v] = 1;
        for (int v : st_nodes) in_sep[v] = 1;
        for (int v = 0; v < N; ++v)
            for (int e = c->h_indptr[v]; e < c->h_indptr[v + 1] && !bnd[v]; ++e) bnd[v] = in_sep[c->h_indices[e]];
        for (int k = 0; k < K; ++k) {
            std::stable_partition(po[k].begin(), po[k].end(), [&](int v) { return !bnd[v]; });
            std::stable_partition(pt[k].begin(), pt[k].end(), [&](int v) { return !bnd[v]; });
            int first = (int)(po[k].size() + pt[k].size());
            for (size_t i = 0; i < po[k].size(); ++i)
                if (bnd[po[k][i]]) { first = (int)i; break; }
            if (first == (int)(po[k].size() + pt[k].size()))
                for (size_t i = 0; i < pt[k].size(); ++i)
                    if (bnd[pt[k][i]]) { first = (int)(po[k].size() + i); break; }
            q.c0[k] = first / 128 * 128;
        }
    }
    const int s = (int)(so_nodes.size() + st_nodes.size());
    int minpart = N;
    for (int k = 0; k < K; ++k) minpart = std::min(minpart, (int)(po[k].size() + pt[k].size()));
    if (s < 1 || s > N / (forced > 0 ? 2 : 4) || minpart < min_part) return 0;
    q.K = K; q.s = s; q.so = (int)so_nodes.size(); q.ND = N - s;
    q.strips = strips; q.cells = cells;
    std::vector<int> perm(N), inv(N);
    int pos = 0;
    for (int k = 0; k < K; ++k) {
        q.off[k] = pos;
        q.ok[k] = (int)po[k].size();
        q.nk[k] = (int)(po[k].size() + pt[k].size());
        for (int v : po[k]) { perm[v] = pos; inv[pos++] = v; }
        for (int v : pt[k]) { perm[v] = pos; inv[pos++] = v; }
    }
    for (int k = K; k <= D6_MAXP; ++k) q.off[k] = pos;
    for (int k = K; k < D6_MAXP; ++k) q.nk[k] = q.ok[k] = 0;
    for (int v : so_nodes) { perm[v] = pos; inv[pos++] = v; }
    for (int v : st_nodes) { perm[v] = pos; inv[pos++] = v; }
    // permuted graph, training ranks
    std::vector<int> pptr(N + 1, 0), pidx(c->h_indices.size()), trank(N, -1);
    std::vector<double> pw(c->h_weights.size());
    size_t e = 0;
    for (int r = 0; r < N; ++r) {
        const int v = inv[r];
        for (int g = c->h_indptr[v]; g < c->h_indptr[v + 1]; ++g, ++e) {
            pidx[e] = perm[c->h_indices[g]];
            if (!pw.empty()) pw[e] = c->h_weights[g];
        }
        pptr[r + 1] = (int)e;
    }
    for (int i = 0; i < n; ++i) trank[perm[c->h_train[i]]] = i;      // h_train is sorted: rank == position in t
    // D^-1/2 on the host exactly as degree_dinv_kernel computes it (same summation order, IEEE sqrt and division)
    std::vector<double> hd(N);
    for (int r = 0; r < N; ++r) {
        double d = 0.0;
        for (int g = pptr[r]; g < pptr[r + 1]; ++g) d += pw.empty() ? 1.0 : pw[g];
        double x = 1.0 / sqrt(d);
        if (std::isinf(x)) x = 0.0;
        hd[r] = x;
    }
    // separator rows x part columns, columns ascending, with the per-part entry ranges
    std::vector<int> sptr(s + 1, 0), sidx, spp((size_t)s * (K + 1), 0);
    std::vector<double> sw;
    std::vector<std::pair<int, double>> rowbuf;
    for (int r = 0; r < s; ++r) {
        rowbuf.clear();
        const int row = q.ND + r;
        for (int g = pptr[row]; g < pptr[row + 1]; ++g)
            if (pidx[g] < q.ND) rowbuf.push_back({pidx[g], hd[row] * ((-(pw.empty() ? 1.0 : pw[g])) * hd[pidx[g]])});
        std::sort(rowbuf.begin(), rowbuf.end());
        int k = 0;
        spp[(size_t)r * (K + 1)] = (int)sidx.size();
        for (auto& pr : rowbuf) {
            while (k + 1 < K && pr.first >= q.off[k + 1]) { ++k; spp[(size_t)r * (K + 1) + k] = (int)sidx.size(); }
            sidx.push_back(pr.first);
            sw.push_back(pr.second);
        }
        while (k < K) { ++k; spp[(size_t)r * (K + 1) + k] = (int)sidx.size(); }
        sptr[r + 1] = (int)sidx.size();
    }
    q.sep_nnz = (long long)sidx.size();
    GPN_TRY(up(c, q.indptr, pptr));
    GPN_TRY(up(c, q.indices, pidx));
    if (!pw.empty()) GPN_TRY(up(c, q.weights, pw));
    GPN_TRY(up(c, q.trank, trank));
    GPN_TRY(up(c, q.sep_ptr, sptr));
    GPN_TRY(up(c, q.sep_idx, sidx));
    GPN_TRY(up(c, q.sep_w, sw));
    GPN_TRY(up(c, q.sep_pp, spp));
    GPN_TRY(gpn_buf_ensure(c, q.dinv, sizeof(double) * (size_t)(2 * N)));
    GPN_TRY(gpn_k_degree_dinv(c, N, (const int*)q.indptr.p, (const int*)q.indices.p, pw.empty() ? nullptr : (const double*)q.weights.p,
                              (double*)q.dinv.p));
    q.usable = true;
    return 0;
}

int gpn_d6_fill_zero(gpn_ctx* c, double* M, long ld) {
    const Dissect6& q = c->d6;
    d6_fill_zero_kernel<<<dim3(2, c->N), TPB, 0, c->stream>>>(make_desc(q), M, ld);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
int gpn_d6_coupling(gpn_ctx* c, const double* t, double beta, double* b1, double* bt, double* d_out5) {
    const Dissect6& q = c->d6;
    GPN_CK(cudaMemsetAsync(d_out5, 0, 5 * sizeof(double), c->stream));
    d6_coupling_kernel<<<(c->N * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(c->N, (const int*)q.indptr.p, (const int*)q.indices.p,
                                                                         c->h_weights.empty() ? nullptr : (const double*)q.weights.p,
                                                                         (const double*)q.dinv.p, (const int*)q.trank.p, t, beta, b1, bt, d_out5);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
int gpn_d6_factor_reduce(gpn_ctx* c, const double* L, const double* WT, long ld, double* d_out4) {
    const Dissect6& q = c->d6;
    GPN_CK(cudaMemsetAsync(d_out4, 0, 4 * sizeof(double), c->stream));
    d6_factor_reduce_kernel<<<q.ND, TPB, 0, c->stream>>>(make_desc(q), L, WT, ld, d_out4);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
int gpn_d6_sep_spmm(gpn_ctx* c, int k, double beta, const double* WT, long ld, double* GT, double* GTo, long ldg) {
    const Dissect6& q = c->d6;
    const int cols = q.nk[k] - q.c0[k];
    if (cols <= 0) return 0;
    d6_sep_spmm_kernel<<<dim3((cols + TPB - 1) / TPB, q.s), TPB, 0, c->stream>>>(make_desc(q), k, q.c0[k], (const int*)q.sep_pp.p,
                                                                               (const int*)q.sep_idx.p, (const double*)q.sep_w.p, beta, WT, ld,
                                                                               (const int*)q.trank.p, GT, GTo, ldg);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
int gpn_d6_t_reduce(gpn_ctx* c, int s, const double* E, long lde, const double* Cb, long lds, int nb, double* T) {
    if (s <= 0) return 0;
    d6_t_reduce_kernel<<<dim3(2, s), TPB, 0, c->stream>>>(s, E, lde, Cb, lds, nb, T);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
// d_out2[0] += ||W_T||_F^2 + ||Y||_F^2 (either may be NULL / rows == 0), d_out2[1] += sum log diag(Lt); the caller zeroes d_out2
int gpn_d6_sep_reduce(gpn_ctx* c, int s, const double* Lt, long ldl, const double* WTt, long ldw, const double* Y, long ldy, int rows,
                      double* d_out2) {
    if (s <= 0) return 0;
    if (!Y) rows = 0;
    d6_sep_reduce_kernel<<<s + rows, TPB, 0, c->stream>>>(s, Lt, ldl, WTt, ldw, Y, ldy, rows, d_out2);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
// u = (L L^T)^-1 x for ONE dense n x n factor given as WTf = L^-T (upper): u = W^T (W x), two right-hand sides; v = scratch
int gpn_d6_tri_solve(gpn_ctx* c, int n, const double* WTf, long ld, const double* xa, const double* xb, double* va, double* vb, double* ua,
                     double* ub) {
    if (n <= 0) return 0;
    D6Desc d = {};
    d.K = 1; d.ND = n; d.nk[0] = n; d.ok[0] = n; d.off[0] = 0;
    for (int k = 1; k <= D6_MAXP; ++k) d.off[k] = n;
    d6_w_mul_kernel<<<dim3((n + 31) / 32, 1), dim3(32, 8), 0, c->stream>>>(d, WTf, ld, xa, xb, va, vb);
    d6_wt_mul_kernel<<<(n * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(d, WTf, ld, va, vb, ua, ub);
    c->launches += 2;
    GPN_CK(cudaGetLastError());
    return 0;
}
// u = blkdiag(M_kk^o)^-1 x on the other-node rows of the parts (training rows of u become 0): u = W^T (W x); v = scratch
int gpn_d6_block_solve(gpn_ctx* c, const double* WT, long ld, const double* xa, const double* xb, double* va, double* vb, double* ua, double* ub) {
    const Dissect6& q = c->d6;
    int omax = 1;
    for (int k = 0; k < q.K; ++k) omax = std::max(omax, q.ok[k]);
    d6_w_mul_kernel<<<dim3((omax + 31) / 32, q.K), dim3(32, 8), 0, c->stream>>>(make_desc(q), WT, ld, xa, xb, va, vb);
    d6_wt_mul_kernel<<<(q.ND * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(make_desc(q), WT, ld, va, vb, ua, ub);
    c->launches += 2;
    GPN_CK(cudaGetLastError());
    return 0;
}
int gpn_d6_sep_gather(gpn_ctx* c, double beta, const double* ba, const double* bb, const double* ya, const double* yb, double* ra, double* rb) {
    const Dissect6& q = c->d6;
    if (q.so <= 0) return 0;
    d6_sep_gather_kernel<<<(q.so * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(make_desc(q), (const int*)q.sep_ptr.p, (const int*)q.sep_idx.p,
                                                                            (const double*)q.sep_w.p, beta, (const int*)q.trank.p, ba, bb, ya, yb, ra, rb);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
int gpn_d6_sep_scatter(gpn_ctx* c, double beta, const double* ua, const double* ub, double* wa, double* wb) {
    const Dissect6& q = c->d6;
    if (q.so <= 0) return 0;
    d6_sep_scatter_kernel<<<(q.so * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(make_desc(q), (const int*)q.sep_ptr.p, (const int*)q.sep_idx.p,
                                                                             (const double*)q.sep_w.p, beta, (const int*)q.trank.p, ua, ub, wa, wb);
    c->launches++;
    GPN_CK(cudaGetLastError());
    return 0;
}
