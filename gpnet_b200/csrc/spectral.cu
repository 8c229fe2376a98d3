// Spectral fast path for hyper-parameter sweeps and batched small-n evaluation (8f-3).
//
// Every Kondor kernel of the reference is a function of the normalized Laplacian (GPnet.py:338-349; write-up eq. 55-56):
//     L~ = V diag(lam) V^T   =>   K(theta) = V diag(r(lam; theta)) V^T,
//     r = exp(-beta lam)  (diffusion, :340),   1 / (1 + beta lam)  (regularized Laplacian, :342),   (a - lam)^p  (p-step walk, :346)
// so ONE eigendecomposition per graph turns every later kernel build (12.3 N^3 flops for a Pade-9 expm) into one symmetric
// product (N^3), and for small training sets a whole logp+gradient evaluation into n^2 N flops that fit one CTA.
//
//  * gpn_spectral_prepare: two-sided block Jacobi.  The matrix is cut into 32-row blocks; a round pairs all blocks (round-robin
//    tournament), every pair's 64 x 64 pivot block is diagonalised by cyclic Jacobi rotations inside one CTA (shared memory,
//    jacobi64_kernel), and the block-diagonal rotation J of the round is applied as A <- J^T A J, V^T <- J^T V^T with three batched
//    DMMA GEMMs (gemm_f64.cu).  The blocks of a pair are made adjacent by a row gather with a FIXED permutation per round (the
//    symmetric second half re-uses the transposed output of the first product), so the batched GEMMs have uniform strides.
//    Jacobi is chosen for robustness: the spectra of grid-like graphs are heavily degenerate, which is harmless here (any
//    orthonormal basis of an eigenspace gives the same K) and needs no reorthogonalisation logic.
//  * gpn_spec_build_kernel: r(lam) on the host, K = V^T' diag(r) V^T as one symmetric GEMM (lower tiles + mirror).
//  * spectral_batch_kernel: one CTA per theta: K_tt and dK_tt/dtheta_0 accumulated from the training columns of V^T (register
//    tiles, DFMA: on B200 the vector FP64 pipe has the DMMA peak), Cholesky, inverse and the trace terms in shared memory.
#include "common.cuh"

#include <algorithm>
#include <numeric>

namespace {

constexpr int TPB = 256;
constexpr double LOG_2PI = 1.8378770664093454835606594728112;
inline long pad128(long n) { return (n + 127) / 128 * 128; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the block, result broadcast to every thread (red: 33 doubles of shared memory)
__device__ __forceinline__ double block_sum_all(double v, double* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.0;
        x = warp_sum(x);
        if (l == 0) red[32] = x;
    }
    __syncthreads();
    return red[32];
}

// ---- Jacobi eigensolver -----------------------------------------------------------------------------------------------
// pads: A[i][i] = pad0 + 0.01 (i - N) for N <= i < Np (rows / columns >= N are otherwise zero: they never couple), VT = I
__global__ void spec_init_kernel(int N, int Np, long ld, double* __restrict__ A, double* __restrict__ VT, double pad0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    if (i >= N) A[(long)i * ld + i] = pad0 + 0.01 * (i - N);
    VT[(long)i * ld + i] = 1.0;
}

// One CTA per block pair: S = the 64 x 64 pivot block of the row-gathered matrix (columns through colmap), diagonalised by
// cyclic two-sided Jacobi (round-robin order: 32 disjoint rotations per step, each applied by 8 threads); Q (S_new = Q^T S Q)
// is written row-major.  Rotations with |a_pq| <= thresh are skipped.
__global__ void __launch_bounds__(256) jacobi64_kernel(const double* __restrict__ AP, long ld, const int* __restrict__ colmap,
                                                        double* __restrict__ Qbuf, double thresh, int max_sweeps) {
    extern __shared__ double sm[];
    double(*S)[65] = reinterpret_cast<double(*)[65]>(sm);
    double(*Q)[65] = reinterpret_cast<double(*)[65]>(sm + 64 * 65);
    __shared__ int any_rot;
    const int tid = threadIdx.x, g = tid >> 3, sub = tid & 7;
    const int p0 = blockIdx.x * 64;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int r = idx >> 6, cc = idx & 63;
        S[r][cc] = AP[(long)(p0 + r) * ld + colmap[p0 + cc]];
        Q[r][cc] = (r == cc) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < 4096; idx += 256) {      // exact symmetry: the lower triangle is the reference
        const int r = idx >> 6, cc = idx & 63;
        if (cc > r) S[r][cc] = S[cc][r];
    }
    __syncthreads();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        if (tid == 0) any_rot = 0;
        int rotated = 0;
        for (int step = 0; step < 63; ++step) {
            int p, q;
            if (g == 0) { p = 63; q = step; }
            else { p = (step + g) % 63; q = (step - g + 63) % 63; }
            if (p > q) { const int t = p; p = q; q = t; }
            const double app = S[p][p], aqq = S[q][q], apq = S[p][q];
            double c = 1.0, s = 0.0;
            if (fabs(apq) > thresh) {
                // t = tan(theta) of the smaller rotation angle, without the division by a_pq: with d = a_qq - a_pp and
                // h = sqrt(d^2 + 4 a_pq^2), t = 2 a_pq / (d + sign(d) h)   (one sqrt, one division, one rsqrt on the chain)
                const double d = aqq - app, b2 = 2.0 * apq;
                const double h = sqrt(fma(d, d, b2 * b2));
                const double t = b2 / (d + copysign(h, d));
                c = rsqrt(fma(t, t, 1.0));
                s = t * c;
                rotated = 1;
            }
            if (s != 0.0) {       // columns p, q of S and of Q (all rows)
                for (int i = sub; i < 64; i += 8) {
                    const double x = S[i][p], y = S[i][q];
                    S[i][p] = c * x - s * y;
                    S[i][q] = s * x + c * y;
                    const double u = Q[i][p], v = Q[i][q];
                    Q[i][p] = c * u - s * v;
                    Q[i][q] = s * u + c * v;
                }
            }
            __syncthreads();
            if (s != 0.0) {       // rows p, q of S (all columns)
                for (int j = sub; j < 64; j += 8) {
                    const double x = S[p][j], y = S[q][j];
                    S[p][j] = c * x - s * y;
                    S[q][j] = s * x + c * y;
                }
            }
            __syncthreads();
        }
        if (rotated && sub == 0) any_rot = 1;
        __syncthreads();
        const int go_on = any_rot;
        __syncthreads();
        if (!go_on) break;
    }
    double* Qo = Qbuf + (long)blockIdx.x * 4096;
    for (int idx = tid; idx < 4096; idx += 256) Qo[idx] = Q[idx >> 6][idx & 63];
}

// out2[0] += sum of squares off the diagonal, out2[1] += sum of squares on it
__global__ void offnorm_kernel(int Np, const double* __restrict__ A, long ld, double* __restrict__ out2) {
    __shared__ double red[33];
    const int r = blockIdx.x;
    double off = 0.0;
    for (int c = threadIdx.x; c < Np; c += blockDim.x) {
        const double v = A[(long)r * ld + c];
        if (c != r) off += v * v;
    }
    off = block_sum_all(off, red);
    if (threadIdx.x == 0) {
        const double d = A[(long)r * ld + r];
        atomicAdd(out2 + 0, off);
        atomicAdd(out2 + 1, d * d);
    }
}
__global__ void diag_kernel(int Np, const double* __restrict__ A, long ld, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Np) out[i] = A[(long)i * ld + i];
}
// out[k][i] = VT[k][cols[i]]
__global__ void gather_cols_kernel(int rows, int nc, const double* __restrict__ VT, long ld, const int* __restrict__ cols,
                                   double* __restrict__ out, long ldo) {
    const int k = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) out[(long)k * ldo + i] = VT[(long)k * ld + cols[i]];
}

int sync_read(gpn_ctx* c, void* dst_h, const void* src_d, size_t bytes) {
    GPN_CK(cudaMemcpyAsync(dst_h, src_d, bytes, cudaMemcpyDeviceToHost, c->stream));
    GPN_CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int sync_write(gpn_ctx* c, void* dst_d, const void* src_h, size_t bytes) {
    GPN_CK(cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, c->stream));
    GPN_CK(cudaStreamSynchronize(c->stream));
    return 0;
}
void buf_free(gpn_ctx* c, DevBuf& b) {
    if (b.p) {
        cudaStreamSynchronize(c->stream);
        cudaFree(b.p);
    }
    b = DevBuf();
}

// Round-robin tournament on physical slots: after every round the two blocks of pair i sit in slots 2i, 2i+1, so from the second
// round on the gather permutation is the same every round.  rm[new row] = old row.
void round_maps(int nb, std::vector<int>& rm0, std::vector<int>& rm1) {
    std::vector<int> arr(nb);
    std::iota(arr.begin(), arr.end(), 0);
    auto emit = [&](std::vector<int>& rm) {
        rm.resize((size_t)nb * 32);
        for (int i = 0; i < nb / 2; ++i) {
            const int a = arr[i], b = arr[nb - 1 - i];
            for (int t = 0; t < 32; ++t) {
                rm[(size_t)(2 * i) * 32 + t] = a * 32 + t;
                rm[(size_t)(2 * i + 1) * 32 + t] = b * 32 + t;
            }
        }
        std::vector<int> na(nb);
        for (int i = 0; i < nb / 2; ++i) { na[i] = 2 * i; na[nb - 1 - i] = 2 * i + 1; }
        arr[0] = na[0];
        arr[1] = na[nb - 1];
        for (int i = 2; i < nb; ++i) arr[i] = na[i - 1];
    };
    emit(rm0);
    emit(rm1);
}

// Eigendecomposition of the symmetric Np x Np matrix in A (Np a multiple of 64; rows / columns >= N hold only the pad diagonal).
// On return A is diagonal to working precision and the rows of VT are the eigenvectors, in the (permuted) order of diag(A).
int jacobi_run(gpn_ctx* c, int Np, long ld, double* A, double* AP, double* BT, double* VT, double* VP, double* Qb, int* rm0_d, int* rm1_d,
               double* sc2_d, int* sweeps_out, double* offrel_out) {
    const int nb = Np / 32, npair = nb / 2;
    const size_t jsm = sizeof(double) * 2 * 64 * 65;
    static bool configured = false;
    if (!configured) {
        GPN_CK(cudaFuncSetAttribute(jacobi64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jsm));
        configured = true;
    }
    double sc[2];
    GPN_CK(cudaMemsetAsync(sc2_d, 0, 2 * sizeof(double), c->stream));
    offnorm_kernel<<<Np, TPB, 0, c->stream>>>(Np, A, ld, sc2_d);
    c->launches++;
    GPN_TRY(sync_read(c, sc, sc2_d, sizeof sc));
    const double total2 = sc[0] + sc[1];
    const double rms = sqrt(total2 / Np);
    const double thresh = 1e-16 * rms;
    double off_rel = total2 > 0 ? sqrt(sc[0] / total2) : 0.0, prev = off_rel;
    int sweeps = 0;
    bool first = true;
    static const int max_sweeps = getenv("GPNET_B200_JACOBI_SWEEPS") ? atoi(getenv("GPNET_B200_JACOBI_SWEEPS")) : 40;
    // cyclic sweeps over a 64 x 64 pivot block per visit: one is enough (more did not reduce the number of outer sweeps)
    static const int inner = getenv("GPNET_B200_JACOBI_INNER") ? atoi(getenv("GPNET_B200_JACOBI_INNER")) : 1;
    while (off_rel > 1e-14 && sweeps < max_sweeps) {
        for (int r = 0; r < nb - 1; ++r) {
            const int* rm = first ? rm0_d : rm1_d;
            first = false;
            GPN_TRY(gpn_k_gather_rows(c, A, ld, rm, Np, Np, AP, ld));
            jacobi64_kernel<<<npair, 256, jsm, c->stream>>>(AP, ld, rm, Qb, thresh, inner);
            c->launches++;
            GPN_CK(cudaGetLastError());
            // B = J^T (P A): only its transpose is kept, B^T = A P^T J
            GPN_TRY(gpn_gemm(c, 64, Np, 64, 1.0, Qb, 64, false, AP, ld, false, 0.0, nullptr, 0, GF_FORCE_SMALL, npair, 4096, 64 * ld, 0, BT, ld, 64));
            GPN_TRY(gpn_k_gather_rows(c, BT, ld, rm, Np, Np, AP, ld));
            GPN_TRY(gpn_gemm(c, 64, Np, 64, 1.0, Qb, 64, false, AP, ld, false, 0.0, A, ld, GF_FORCE_SMALL, npair, 4096, 64 * ld, 64 * ld));
            GPN_TRY(gpn_k_gather_rows(c, VT, ld, rm, Np, Np, VP, ld));
            GPN_TRY(gpn_gemm(c, 64, Np, 64, 1.0, Qb, 64, false, VP, ld, false, 0.0, VT, ld, GF_FORCE_SMALL, npair, 4096, 64 * ld, 64 * ld));
        }
        ++sweeps;
        GPN_CK(cudaMemsetAsync(sc2_d, 0, 2 * sizeof(double), c->stream));
        offnorm_kernel<<<Np, TPB, 0, c->stream>>>(Np, A, ld, sc2_d);
        c->launches++;
        GPN_TRY(sync_read(c, sc, sc2_d, sizeof sc));
        off_rel = sqrt(sc[0] / (sc[0] + sc[1]));
        if (off_rel < 1e-12 && off_rel > 0.5 * prev) break;      // rounding floor reached
        prev = off_rel;
    }
    *sweeps_out = sweeps;
    *offrel_out = off_rel;
    return 0;
}

// ---- batched small-n evaluation ------------------------------------------------------------------------------------------
struct SpecBatchArgs {
    int N, n, ldv, P, kind, want_grad, B;
    const double *Vt, *lam, *t, *thetas;
    double* out;
    int* info;
};

__device__ __forceinline__ double ipow(double b, int p) {      // binary powering like np.linalg.matrix_power
    double r = 1.0;
    while (p > 0) {
        if (p & 1) r *= b;
        b *= b;
        p >>= 1;
    }
    return r;
}
__device__ __forceinline__ void spectral_weights(int kind, double lam, double th0, int p, double& r, double& rd) {
    if (kind == GPN_KERNEL_DIFFUSION) {
        r = exp(-th0 * lam);
        rd = -th0 * lam * r;                        // d/dtheta0 exp(-e^theta0 lam)
    } else if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN) {
        r = 1.0 / (1.0 + th0 * lam);
        rd = r * r - r;                             // K^2 - K
    } else {
        const double b = th0 - lam;
        if (p <= 0) { r = 1.0; rd = 0.0; }
        else { r = ipow(b, p); rd = th0 * p * ipow(b, p - 1); }      // a p B^(p-1)
    }
}

constexpr int KC = 32;       // eigenvectors per shared-memory chunk

// One CTA per theta.  Shared memory: X[n][n+1]: K_y -> L in the lower triangle, dK/dtheta0 (lower, entry (i,j) at X[j][i+1]);
// Y[n][n+1]: W = L^-1 (lower), K_y^-1 (entry (i,j), j <= i, at Y[j][i+1]); the chunk of V_t rows aliases Y while K is accumulated.
__global__ void __launch_bounds__(256) spectral_batch_kernel(SpecBatchArgs a) {
    extern __shared__ double sm[];
    const int n = a.n, ldx = n + 1, n4 = (n + 3) / 4 * 4;
    double* X = sm;
    double* Y = X + (size_t)n * ldx;
    const size_t ysz = max((size_t)n * ldx, (size_t)KC * n4 + 2 * KC);
    double* tv = Y + ysz;
    double* alpha = tv + n4;
    double* rsum = alpha + n4;
    double* red = rsum + n4;       // 33 doubles
    __shared__ int fail;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nb4 = n4 / 4, nlb = nb4 * (nb4 + 1) / 2;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double* th = a.thetas + (long)b * a.P;
        const double th0 = th[0], noise = th[a.P - 1];
        const int pw = (a.kind == GPN_KERNEL_PSTEP_WALK && a.P >= 2) ? (int)th[1] : 0;
        if (tid == 0) fail = 0;
        for (int i = tid; i < n4; i += 256) tv[i] = i < n ? a.t[i] : 0.0;

        // ---- K_tt and D = dK_tt/dtheta0 on 4x4 register tiles of the lower triangle ----
        int bi[2], bj[2];
        double accK[2][4][4], accD[2][4][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = tid + 256 * u;
            int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
            while ((long)i * (i + 1) / 2 > e) --i;
            while ((long)(i + 1) * (i + 2) / 2 <= e) ++i;
            bi[u] = e < nlb ? i : -1;
            bj[u] = e - i * (i + 1) / 2;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) accK[u][x][y] = accD[u][x][y] = 0.0;
        }
        double* chunk = Y;                  // KC x n4
        double* rk = Y + (size_t)KC * n4;   // KC weights, then KC derivative weights
        for (int k0 = 0; k0 < a.N; k0 += KC) {
            __syncthreads();
            const int kc = min(KC, a.N - k0);
            for (int idx = tid; idx < KC * n4; idx += 256) {
                const int kk = idx / n4, i = idx - kk * n4;
                chunk[idx] = (kk < kc && i < n) ? a.Vt[(long)(k0 + kk) * a.ldv + i] : 0.0;
            }
            if (tid < KC) {
                double r = 0.0, rd = 0.0;
                if (tid < kc) spectral_weights(a.kind, a.lam[k0 + tid], th0, pw, r, rd);
                rk[tid] = r;
                rk[KC + tid] = rd;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (bi[u] < 0) continue;
                const double* vi = chunk + 4 * bi[u];
                const double* vj = chunk + 4 * bj[u];
                for (int kk = 0; kk < KC; ++kk) {
                    const double r = rk[kk], rd = rk[KC + kk];
                    double xi[4], xj[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) { xi[x] = vi[kk * n4 + x]; xj[x] = vj[kk * n4 + x]; }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double ur = r * xi[x], ud = rd * xi[x];
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            accK[u][x][y] = fma(ur, xj[y], accK[u][x][y]);
                            if (a.want_grad) accD[u][x][y] = fma(ud, xj[y], accD[u][x][y]);
                        }
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (bi[u] < 0) continue;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = 4 * bi[u] + x, j = 4 * bj[u] + y;
                    if (i < n && j <= i) {
                        X[(size_t)i * ldx + j] = accK[u][x][y] + noise;          // noise on EVERY entry (quirk A-3)
                        X[(size_t)j * ldx + i + 1] = accD[u][x][y];
                    }
                }
        }
        __syncthreads();

        // ---- Cholesky of K_y in place (lower) ----
        for (int k = 0; k < n; ++k) {
            const double d = X[(size_t)k * ldx + k];
            if (!(d > 0.0)) {
                if (tid == 0) fail = k + 1;
                break;
            }
            const double lkk = sqrt(d), inv = 1.0 / lkk;
            __syncthreads();
            for (int i = k + 1 + tid; i < n; i += 256) X[(size_t)i * ldx + k] *= inv;
            if (tid == 0) X[(size_t)k * ldx + k] = lkk;
            __syncthreads();
            for (int i = k + 1 + ty; i < n; i += 16) {
                const double lik = X[(size_t)i * ldx + k];
                for (int j = k + 1 + tx; j <= i; j += 16) X[(size_t)i * ldx + j] -= lik * X[(size_t)j * ldx + k];
            }
            __syncthreads();
        }
        __syncthreads();
        if (fail) {
            if (tid == 0) {
                a.info[b] = fail;
                for (int j = 0; j <= a.P; ++j) a.out[(long)b * (1 + a.P) + j] = 0.0;
            }
            __syncthreads();
            continue;
        }
        // ---- W = L^-1 (Y lower) by row-wise elimination ----
        for (int idx = tid; idx < n * ldx; idx += 256) {
            const int i = idx / ldx, j = idx - i * ldx;
            Y[idx] = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            const double inv = 1.0 / X[(size_t)k * ldx + k];
            for (int j = tid; j <= k; j += 256) Y[(size_t)k * ldx + j] *= inv;
            __syncthreads();
            for (int i = k + 1 + ty; i < n; i += 16) {
                const double lik = X[(size_t)i * ldx + k];
                for (int j = tx; j <= k; j += 16) Y[(size_t)i * ldx + j] -= lik * Y[(size_t)k * ldx + j];
            }
            __syncthreads();
        }
        // ---- K_y^-1 = W^T W: entry (i, j), j <= i, at Y[j][i+1] ----
        for (int i = ty; i < n; i += 16)
            for (int j = tx; j <= i; j += 16) {
                double s0 = 0.0, s1 = 0.0;
                int k = i;
                for (; k + 1 < n; k += 2) {
                    s0 = fma(Y[(size_t)k * ldx + i], Y[(size_t)k * ldx + j], s0);
                    s1 = fma(Y[(size_t)(k + 1) * ldx + i], Y[(size_t)(k + 1) * ldx + j], s1);
                }
                if (k < n) s0 = fma(Y[(size_t)k * ldx + i], Y[(size_t)k * ldx + j], s0);
                Y[(size_t)j * ldx + i + 1] = s0 + s1;
            }
        __syncthreads();
        // ---- alpha = K_y^-1 t, row sums of K_y^-1 ----
        for (int i = tid; i < n; i += 256) {
            double s = 0.0, rs = 0.0;
            for (int j = 0; j < n; ++j) {
                const double kij = j <= i ? Y[(size_t)j * ldx + i + 1] : Y[(size_t)i * ldx + j + 1];
                s = fma(kij, tv[j], s);
                rs += kij;
            }
            alpha[i] = s;
            rsum[i] = rs;
        }
        __syncthreads();
        double q = 0.0, s1 = 0.0, ld_ = 0.0, tk1 = 0.0, trKD = 0.0, aDa = 0.0;
        for (int i = tid; i < n; i += 256) {
            q += tv[i] * alpha[i];
            s1 += alpha[i];
            ld_ += log(X[(size_t)i * ldx + i]);
            tk1 += rsum[i];
        }
        if (a.want_grad) {
            for (int i = ty; i < n; i += 16)
                for (int j = tx; j <= i; j += 16) {
                    const double w = (i == j) ? 1.0 : 2.0;
                    const double dij = X[(size_t)j * ldx + i + 1];
                    trKD += w * Y[(size_t)j * ldx + i + 1] * dij;
                    aDa += w * alpha[i] * alpha[j] * dij;
                }
        }
        q = block_sum_all(q, red);
        s1 = block_sum_all(s1, red);
        ld_ = block_sum_all(ld_, red);
        tk1 = block_sum_all(tk1, red);
        trKD = block_sum_all(trKD, red);
        aDa = block_sum_all(aDa, red);
        if (tid == 0) {
            double* o = a.out + (long)b * (1 + a.P);
            a.info[b] = 0;
            o[0] = 0.5 * q + ld_ + 0.5 * n * LOG_2PI;                 // -logp (GPnetRegressor.py:145-150)
            for (int j = 1; j <= a.P; ++j) o[j] = 0.0;
            if (a.want_grad) {
                o[1] += -0.5 * (aDa - trKD);                           // -(1/2)(a^T D a - tr(K^-1 D)), old_implementation/GPR.py:52-64
                o[a.P] += -0.5 * noise * (s1 * s1 - tk1);              // D_last = c 1 1^T
            }
        }
        __syncthreads();
    }
}


// ---- batched small-n Laplace Newton loop (GPnetClassifier.NRiteration, GPnetClassifier.py:90-147) ------------------------------
// One CTA per theta: K = V_t r(lam) V_t^T + c in shared memory, then the reference's loop line for line (the same statements as
// gpc_newton in gp.cu): W, p, b from f; B = I + sqrtW K sqrtW; L = chol(B); a = scale (b - sqrtW (B^-1 (sqrtW (K b)))); f = K a;
// phif from f^T K^-1 f (one Cholesky of K per theta; f^T a if K is not numerically positive definite), log det B and log p;
// stop when sum (phif_old - phif)^2 < tol; logq = -a^T f / 2 - sum log(1 + log(1 + exp(y f))) - sum log diag L   (quirk A-10).
// Shared memory: X[n][n+1]: K (lower), L_K (entry (i,j), j <= i, at X[j][i+1]);  Y[n][n+1]: B -> L_B (lower).
struct SpecNewtonArgs {
    int N, n, ldv, P, kind, B;
    const double *Vt, *lam, *y, *thetas;
    double tol, phif0, scale0;
    double* logq;      // B
    int* iters;        // B
    int* info;         // B
};

// in-place lower Cholesky of the n x n matrix A (row stride ldx) by the whole CTA; returns 0 or the failing minor (uniform)
__device__ __forceinline__ int cta_cholesky(double* A, int ldx, int n, int tid, int tx, int ty) {
    for (int k = 0; k < n; ++k) {
        const double d = A[(size_t)k * ldx + k];
        if (!(d > 0.0)) return k + 1;
        const double lkk = sqrt(d), inv = 1.0 / lkk;
        __syncthreads();
        for (int i = k + 1 + tid; i < n; i += 256) A[(size_t)i * ldx + k] *= inv;
        if (tid == 0) A[(size_t)k * ldx + k] = lkk;
        __syncthreads();
        for (int i = k + 1 + ty; i < n; i += 16) {
            const double lik = A[(size_t)i * ldx + k];
            for (int j = k + 1 + tx; j <= i; j += 16) A[(size_t)i * ldx + j] -= lik * A[(size_t)j * ldx + k];
        }
        __syncthreads();
    }
    return 0;
}
// v <- L^-1 v (forward) for L lower in A;  TRANS_STORE: L[i][k] is read from A[k][i+1] (the packed upper half)
template <bool TRANS_STORE>
__device__ __forceinline__ void cta_forward(const double* A, int ldx, int n, double* v, int tid) {
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        const double lkk = TRANS_STORE ? A[(size_t)k * ldx + k + 1] : A[(size_t)k * ldx + k];
        const double vk = v[k] / lkk;
        __syncthreads();
        if (tid == 0) v[k] = vk;
        for (int i = k + 1 + tid; i < n; i += 256) v[i] -= (TRANS_STORE ? A[(size_t)k * ldx + i + 1] : A[(size_t)i * ldx + k]) * vk;
    }
    __syncthreads();
}
// v <- L^-T v (backward) for L lower in A
__device__ __forceinline__ void cta_backward(const double* A, int ldx, int n, double* v, int tid) {
    for (int k = n - 1; k >= 0; --k) {
        __syncthreads();
        const double vk = v[k] / A[(size_t)k * ldx + k];
        __syncthreads();
        if (tid == 0) v[k] = vk;
        for (int j = tid; j < k; j += 256) v[j] -= A[(size_t)k * ldx + j] * vk;
    }
    __syncthreads();
}
// out = K x for the symmetric K held in the lower triangle of X
__device__ __forceinline__ void cta_symv(const double* X, int ldx, int n, const double* x, double* out, int tid) {
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(j <= i ? X[(size_t)i * ldx + j] : X[(size_t)j * ldx + i], x[j], s);
        out[i] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) spectral_newton_kernel(SpecNewtonArgs a) {
    extern __shared__ double sm[];
    const int n = a.n, ldx = n + 1, n4 = (n + 3) / 4 * 4;
    double* X = sm;
    double* Y = X + (size_t)n * ldx;
    const size_t ysz = max((size_t)n * ldx, (size_t)KC * n4 + 2 * KC);
    double* yv = Y + ysz;          // labels
    double* f = yv + n4;
    double* av = f + n4;
    double* bv = av + n4;
    double* pv = bv + n4;
    double* sw = pv + n4;
    double* tmp = sw + n4;
    double* phif = tmp + n4;
    double* red = phif + n4;       // 33 doubles
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nb4 = n4 / 4, nlb = nb4 * (nb4 + 1) / 2;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double* th = a.thetas + (long)b * a.P;
        const double th0 = th[0], noise = th[a.P - 1];
        const int pw = (a.kind == GPN_KERNEL_PSTEP_WALK && a.P >= 2) ? (int)th[1] : 0;
        __syncthreads();
        for (int i = tid; i < n4; i += 256) { yv[i] = i < n ? a.y[i] : 0.0; f[i] = 0.0; }
        // ---- K_tt on 4x4 register tiles of the lower triangle (see spectral_batch_kernel) ----
        int bi[2], bj[2];
        double accK[2][4][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = tid + 256 * u;
            int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
            while ((long)i * (i + 1) / 2 > e) --i;
            while ((long)(i + 1) * (i + 2) / 2 <= e) ++i;
            bi[u] = e < nlb ? i : -1;
            bj[u] = e - i * (i + 1) / 2;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) accK[u][x][y] = 0.0;
        }
        double* chunk = Y;
        double* rk = Y + (size_t)KC * n4;
        for (int k0 = 0; k0 < a.N; k0 += KC) {
            __syncthreads();
            const int kc = min(KC, a.N - k0);
            for (int idx = tid; idx < KC * n4; idx += 256) {
                const int kk = idx / n4, i = idx - kk * n4;
                chunk[idx] = (kk < kc && i < n) ? a.Vt[(long)(k0 + kk) * a.ldv + i] : 0.0;
            }
            if (tid < KC) {
                double r = 0.0, rd = 0.0;
                if (tid < kc) spectral_weights(a.kind, a.lam[k0 + tid], th0, pw, r, rd);
                rk[tid] = r;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (bi[u] < 0) continue;
                const double* vi = chunk + 4 * bi[u];
                const double* vj = chunk + 4 * bj[u];
                for (int kk = 0; kk < KC; ++kk) {
                    const double r = rk[kk];
                    double xi[4], xj[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) { xi[x] = vi[kk * n4 + x]; xj[x] = vj[kk * n4 + x]; }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double ur = r * xi[x];
#pragma unroll
                        for (int y = 0; y < 4; ++y) accK[u][x][y] = fma(ur, xj[y], accK[u][x][y]);
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (bi[u] < 0) continue;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = 4 * bi[u] + x, j = 4 * bj[u] + y;
                    if (i < n && j <= i) {
                        const double v = accK[u][x][y] + noise;       // kernel(train, train): noise on EVERY entry (A-3)
                        X[(size_t)i * ldx + j] = v;
                        Y[(size_t)i * ldx + j] = v;
                    }
                }
        }
        __syncthreads();
        // ---- L_K = chol(K) (in Y), parked in the upper half of X; K not positive definite: q = f^T a instead ----
        const int info_k = cta_cholesky(Y, ldx, n, tid, tx, ty);
        __syncthreads();
        if (info_k == 0)
            for (int i = ty; i < n; i += 16)
                for (int j = tx; j <= i; j += 16) X[(size_t)j * ldx + i + 1] = Y[(size_t)i * ldx + j];
        __syncthreads();
        // ---- Newton loop ----
        double scale = a.scale0, logdet = 0.0;
        int count = 0, iters = 0, fail = 0;
        for (;;) {
            ++count;
            ++iters;
            for (int i = tid; i < n; i += 256) {
                const double fi = f[i];
                const double e = exp(-fabs(fi));
                const double d = 1.0 + e;
                const double pi = (fi >= 0.0) ? 1.0 / d : e / d;
                const double wi = e / (d * d);
                pv[i] = pi;
                sw[i] = sqrt(wi);
                bv[i] = wi * fi + (0.5 * (yv[i] + 1.0) - pi);
            }
            __syncthreads();
            for (int i = ty; i < n; i += 16)
                for (int j = tx; j <= i; j += 16) Y[(size_t)i * ldx + j] = (i == j ? 1.0 : 0.0) + sw[i] * X[(size_t)i * ldx + j] * sw[j];
            __syncthreads();
            fail = cta_cholesky(Y, ldx, n, tid, tx, ty);
            __syncthreads();
            if (fail) break;
            cta_symv(X, ldx, n, bv, tmp, tid);                                  // K b
            for (int i = tid; i < n; i += 256) tmp[i] *= sw[i];
            cta_forward<false>(Y, ldx, n, tmp, tid);
            cta_backward(Y, ldx, n, tmp, tid);                                   // B^-1 (sqrtW K b)
            for (int i = tid; i < n; i += 256) av[i] = scale * (bv[i] - sw[i] * tmp[i]);
            cta_symv(X, ldx, n, av, f, tid);                                     // f = K a
            double q;
            if (info_k == 0) {
                for (int i = tid; i < n; i += 256) tmp[i] = f[i];
                cta_forward<true>(X, ldx, n, tmp, tid);                          // u = L_K^-1 f
                double s = 0.0;
                for (int i = tid; i < n; i += 256) s += tmp[i] * tmp[i];
                q = block_sum_all(s, red);
            } else {
                double s = 0.0;
                for (int i = tid; i < n; i += 256) s += f[i] * av[i];
                q = block_sum_all(s, red);
            }
            double ld_ = 0.0;
            for (int i = tid; i < n; i += 256) ld_ += log(Y[(size_t)i * ldx + i]);
            logdet = block_sum_all(ld_, red);
            const double common = -0.5 * q - 0.5 * logdet - 0.5 * n * LOG_2PI;
            double cs = 0.0;
            for (int i = tid; i < n; i += 256) {
                const double v = log(pv[i]) + common;
                const double old = iters == 1 ? a.phif0 : phif[i];
                const double dd = old - v;
                cs += dd * dd;
                phif[i] = v;
            }
            const double conv = block_sum_all(cs, red);
            if (conv < a.tol) break;
            if (count > 100) { count = 0; scale *= 0.5; }
            if (iters > 100000) { fail = -1; break; }
        }
        double af = 0.0, lik = 0.0;
        if (!fail) {
            for (int i = tid; i < n; i += 256) {
                af += av[i] * f[i];
                const double sv = -yv[i] * f[i];
                lik += log(1.0 + log(1.0 + exp(-sv)));
            }
        }
        af = block_sum_all(af, red);
        lik = block_sum_all(lik, red);
        if (tid == 0) {
            a.info[b] = fail;
            a.iters[b] = iters;
            a.logq[b] = -0.5 * af - lik - logdet;
        }
        __syncthreads();
    }
}

}   // namespace

// =====================================================================================================================
// internal entry points (kernel_build in gp.cu) and the C-ABI of the spectral path
// =====================================================================================================================
int gpn_spec_invalidate(gpn_ctx* c, bool graph_changed) {
    if (graph_changed) c->spec.ready = false;
    c->spec.vt_valid = false;
    return 0;
}

// eigendecomposition of the bound graph's normalized Laplacian (once per graph)
int gpn_spec_prepare(gpn_ctx* c) {
    Spectral& sp = c->spec;
    if (sp.ready) return 0;
    const int N = c->N;
    if (N <= 0) return -1;
    const int Np = (N + 63) / 64 * 64, nb = Np / 32;
    const long ld = pad128(Np);
    DevBuf work[5], qb, rm0, rm1, sc, dg, idx;
    auto cleanup = [&]() {
        for (auto& w : work) buf_free(c, w);
        buf_free(c, qb); buf_free(c, rm0); buf_free(c, rm1); buf_free(c, sc); buf_free(c, dg); buf_free(c, idx);
    };
    int st = 0;
    auto run = [&]() -> int {
        for (auto& w : work) GPN_TRY(gpn_buf_ensure(c, w, sizeof(double) * (size_t)Np * ld));
        GPN_TRY(gpn_buf_ensure(c, qb, sizeof(double) * (size_t)(nb / 2) * 4096));
        GPN_TRY(gpn_buf_ensure(c, rm0, sizeof(int) * (size_t)Np));
        GPN_TRY(gpn_buf_ensure(c, rm1, sizeof(int) * (size_t)Np));
        GPN_TRY(gpn_buf_ensure(c, sc, sizeof(double) * 4));
        GPN_TRY(gpn_buf_ensure(c, dg, sizeof(double) * (size_t)Np));
        GPN_TRY(gpn_buf_ensure(c, idx, sizeof(int) * (size_t)Np));
        double *A = (double*)work[0].p, *AP = (double*)work[1].p, *BT = (double*)work[2].p, *VT = (double*)work[3].p, *VP = (double*)work[4].p;
        GPN_TRY(gpn_k_fill_zero(c, A, ld, Np, (int)ld));
        GPN_TRY(gpn_k_fill_zero(c, VT, ld, Np, (int)ld));
        GPN_TRY(gpn_k_laplacian_scatter(c, N, (const int*)c->indptr.p, (const int*)c->indices.p, c->weights.p ? (const double*)c->weights.p : nullptr,
                                        (const double*)c->dinv.p, 1.0, 0.0, A, ld));
        spec_init_kernel<<<(Np + TPB - 1) / TPB, TPB, 0, c->stream>>>(N, Np, ld, A, VT, 4.0);      // spec(L~) lies in [0, 2]
        c->launches++;
        std::vector<int> h0, h1;
        round_maps(nb, h0, h1);
        GPN_TRY(sync_write(c, rm0.p, h0.data(), sizeof(int) * h0.size()));
        GPN_TRY(sync_write(c, rm1.p, h1.data(), sizeof(int) * h1.size()));
        GPN_TRY(jacobi_run(c, Np, ld, A, AP, BT, VT, VP, (double*)qb.p, (int*)rm0.p, (int*)rm1.p, (double*)sc.p, &sp.sweeps, &sp.off_rel));
        diag_kernel<<<(Np + TPB - 1) / TPB, TPB, 0, c->stream>>>(Np, A, ld, (double*)dg.p);
        c->launches++;
        std::vector<double> lam(Np);
        GPN_TRY(sync_read(c, lam.data(), dg.p, sizeof(double) * Np));
        std::vector<int> order;
        for (int i = 0; i < Np; ++i)
            if (lam[i] < 3.0) order.push_back(i);
        if ((int)order.size() != N) {
            gpn_set_last_error(__FILE__, __LINE__, "Jacobi eigensolver: pad eigenvalues mixed with the spectrum");
            return -96;
        }
        std::sort(order.begin(), order.end(), [&](int x, int y) { return lam[x] < lam[y]; });
        sp.h_lam.resize(N);
        for (int i = 0; i < N; ++i) sp.h_lam[i] = lam[order[i]];
        GPN_TRY(sync_write(c, idx.p, order.data(), sizeof(int) * N));
        const long ldN = pad128(N);
        GPN_TRY(gpn_buf_ensure(c, sp.VT, sizeof(double) * (size_t)ldN * ldN));
        GPN_TRY(gpn_buf_ensure(c, sp.lam, sizeof(double) * (size_t)ldN));
        GPN_TRY(gpn_buf_ensure(c, sp.rbuf, sizeof(double) * (size_t)ldN * 2));
        GPN_TRY(gpn_k_fill_zero(c, (double*)sp.VT.p, ldN, (int)ldN, (int)ldN));
        GPN_TRY(gpn_k_gather_rows(c, VT, ld, (const int*)idx.p, N, N, (double*)sp.VT.p, ldN));
        GPN_TRY(sync_write(c, sp.lam.p, sp.h_lam.data(), sizeof(double) * N));
        return 0;
    };
    st = run();
    cleanup();
    if (st != 0) return st;
    sp.ready = true;
    sp.vt_valid = false;
    return 0;
}

static void host_weights(const gpn_ctx* c, int kind, double th0, int p, std::vector<double>& r, std::vector<double>* r_pm1) {
    const std::vector<double>& lam = c->spec.h_lam;
    const size_t N = lam.size();
    r.resize(N);
    if (r_pm1) r_pm1->resize(N);
    auto ipw = [](double b, int e) {
        double x = 1.0;
        while (e > 0) { if (e & 1) x *= b; b *= b; e >>= 1; }
        return x;
    };
    for (size_t k = 0; k < N; ++k) {
        if (kind == GPN_KERNEL_DIFFUSION) r[k] = exp(-th0 * lam[k]);
        else if (kind == GPN_KERNEL_REGULARIZED_LAPLACIAN) r[k] = 1.0 / (1.0 + th0 * lam[k]);
        else {
            r[k] = p <= 0 ? 1.0 : ipw(th0 - lam[k], p);
            if (r_pm1) (*r_pm1)[k] = p <= 1 ? 1.0 : ipw(th0 - lam[k], p - 1);
        }
    }
}

// K = V diag(r(lam)) V^T into Kout (full symmetric, ld); pm1 (p-step only, may be NULL): B^(p-1) for the derivative; scratch: N x ld
int gpn_spec_build_kernel(gpn_ctx* c, int kind, double th0, int p, double* Kout, long ld, double* pm1, double* scratch) {
    GPN_TRY(gpn_spec_prepare(c));
    Spectral& sp = c->spec;
    const int N = c->N;
    const long ldN = pad128(N);
    std::vector<double> r, r1;
    host_weights(c, kind, th0, p, r, pm1 ? &r1 : nullptr);
    double* rb = (double*)sp.rbuf.p;
    const double* VT = (const double*)sp.VT.p;
    GPN_TRY(sync_write(c, rb, r.data(), sizeof(double) * N));
    GPN_TRY(gpn_k_scale_rows(c, N, N, VT, ldN, rb, scratch, ld));
    GPN_TRY(gpn_gemm(c, N, N, N, 1.0, VT, ldN, false, scratch, ld, false, 0.0, Kout, ld, GF_LOWER | GF_MIRROR));
    if (pm1) {
        GPN_TRY(sync_write(c, rb + ldN, r1.data(), sizeof(double) * N));
        GPN_TRY(gpn_k_scale_rows(c, N, N, VT, ldN, rb + ldN, scratch, ld));
        GPN_TRY(gpn_gemm(c, N, N, N, 1.0, VT, ldN, false, scratch, ld, false, 0.0, pm1, ld, GF_LOWER | GF_MIRROR));
    }
    return 0;
}

// the training columns of V^T for the bound node set (once per node set): Vt[k][i] = VT[k][train_i]
static int spec_training_columns(gpn_ctx* c) {
    GPN_TRY(gpn_spec_prepare(c));
    Spectral& sp = c->spec;
    if (sp.vt_valid) return 0;
    const int n = c->n_train, N = c->N;
    const long ldN = pad128(N);
    const int ldv = (n + 3) / 4 * 4;
    GPN_TRY(gpn_buf_ensure(c, sp.Vt, sizeof(double) * (size_t)N * ldv));
    dim3 grid(1, N);
    gather_cols_kernel<<<grid, 128, 0, c->stream>>>(N, n, (const double*)sp.VT.p, ldN, (const int*)c->train_idx.p, (double*)sp.Vt.p, ldv);
    c->launches++;
    GPN_CK(cudaGetLastError());
    sp.vt_valid = true;
    return 0;
}

extern "C" {

int gpn_spectral_prepare(gpn_ctx* c, int* sweeps_h, double* offdiag_rel_h) {
    GPN_RANGE();
    if (!c) return -1;
    GPN_CK(cudaSetDevice(c->device));
    GPN_TRY(gpn_spec_prepare(c));
    if (sweeps_h) *sweeps_h = c->spec.sweeps;
    if (offdiag_rel_h) *offdiag_rel_h = c->spec.off_rel;
    return 0;
}

void gpn_spectral_enable(gpn_ctx* c, int on) {
    if (c) c->spec.on = on != 0;
}

int gpn_spectral_get(gpn_ctx* c, double* lam_h, double* VT_h) {
    GPN_RANGE();
    if (!c || !c->spec.ready) return -1;
    const int N = c->N;
    const long ldN = pad128(N);
    if (lam_h) memcpy(lam_h, c->spec.h_lam.data(), sizeof(double) * N);
    if (VT_h) {
        GPN_CK(cudaMemcpy2DAsync(VT_h, sizeof(double) * N, c->spec.VT.p, sizeof(double) * ldN, sizeof(double) * N, N, cudaMemcpyDeviceToHost, c->stream));
        GPN_CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int gpn_gpr_logp_grad_batch(gpn_ctx* c, int kind, const double* thetas_h, int B, int P, const double* t_h, int want_grad, double* out_h,
                            int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (B <= 0) return 0;
    if (P < 1 || P > 8) return -5;
    if (kind < 0 || kind > 2) return -2;
    const int n = c->n_train, N = c->N;
    if (n <= 0 || n > GPN_SPECTRAL_BATCH_MAX_N) return -20;      // the training block must fit one CTA's shared memory
    if (kind == GPN_KERNEL_PSTEP_WALK && P < 2) return -5;
    for (int b = 0; b < B; ++b) {      // parameter domains as in kernel_build (GPnet.py:339,344)
        if (kind == GPN_KERNEL_DIFFUSION && !(thetas_h[(long)b * P] < 1.0)) return -100;
        if (kind == GPN_KERNEL_PSTEP_WALK && !(thetas_h[(long)b * P] >= 2.0)) return -101;
    }
    GPN_CK(cudaSetDevice(c->device));
    GPN_TRY(spec_training_columns(c));
    Spectral& sp = c->spec;
    const int ldv = (n + 3) / 4 * 4;
    GPN_TRY(gpn_buf_ensure(c, sp.bt, sizeof(double) * (size_t)(n + 8)));
    GPN_TRY(gpn_buf_ensure(c, sp.bth, sizeof(double) * (size_t)B * P));
    GPN_TRY(gpn_buf_ensure(c, sp.bout, sizeof(double) * (size_t)B * (1 + P) + sizeof(int) * (size_t)B));
    if (t_h) GPN_TRY(sync_write(c, sp.bt.p, t_h, sizeof(double) * n));
    else GPN_CK(cudaMemcpyAsync(sp.bt.p, gpn_resident_targets(c), sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    GPN_TRY(sync_write(c, sp.bth.p, thetas_h, sizeof(double) * (size_t)B * P));
    SpecBatchArgs a;
    a.N = N; a.n = n; a.ldv = ldv; a.P = P; a.kind = kind; a.want_grad = want_grad; a.B = B;
    a.Vt = (const double*)sp.Vt.p; a.lam = (const double*)sp.lam.p; a.t = (const double*)sp.bt.p; a.thetas = (const double*)sp.bth.p;
    a.out = (double*)sp.bout.p;
    a.info = (int*)((double*)sp.bout.p + (size_t)B * (1 + P));
    const int n4 = ldv;
    const size_t ysz = std::max((size_t)n * (n + 1), (size_t)KC * n4 + 2 * KC);
    const size_t smem = sizeof(double) * ((size_t)n * (n + 1) + ysz + 3 * (size_t)n4 + 40);
    static size_t configured = 0;
    if (smem > configured) {
        GPN_CK(cudaFuncSetAttribute(spectral_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = std::min(B, 4 * c->num_sms);
    spectral_batch_kernel<<<grid, 256, smem, c->stream>>>(a);
    c->launches++;
    GPN_CK(cudaGetLastError());
    c->flops += (double)B * (want_grad ? 4.0 : 2.0) * 0.5 * n * (double)n * N;      // DFMA flops of the K / D accumulation (lower halves)
    GPN_TRY(sync_read(c, out_h, a.out, sizeof(double) * (size_t)B * (1 + P)));
    GPN_TRY(sync_read(c, info_h, a.info, sizeof(int) * (size_t)B));
    c->last_route = 7;
    return 0;
}

int gpn_gpc_newton_batch(gpn_ctx* c, int kind, const double* thetas_h, int B, int P, const double* y_h, double tol, double phif0, double scale,
                         double* logq_h, int* iters_h, int* info_h) {
    GPN_RANGE();
    if (!c) return -1;
    if (B <= 0) return 0;
    if (P < 1 || P > 8) return -5;
    if (kind < 0 || kind > 2) return -2;
    const int n = c->n_train, N = c->N;
    if (n <= 0 || n > GPN_SPECTRAL_BATCH_MAX_N) return -20;
    if (!y_h) return -6;
    if (kind == GPN_KERNEL_PSTEP_WALK && P < 2) return -5;
    for (int b = 0; b < B; ++b) {
        if (kind == GPN_KERNEL_DIFFUSION && !(thetas_h[(long)b * P] < 1.0)) return -100;
        if (kind == GPN_KERNEL_PSTEP_WALK && !(thetas_h[(long)b * P] >= 2.0)) return -101;
    }
    GPN_CK(cudaSetDevice(c->device));
    GPN_TRY(spec_training_columns(c));
    Spectral& sp = c->spec;
    const int ldv = (n + 3) / 4 * 4, n4 = ldv;
    GPN_TRY(gpn_buf_ensure(c, sp.bt, sizeof(double) * (size_t)(n + 8)));
    GPN_TRY(gpn_buf_ensure(c, sp.bth, sizeof(double) * (size_t)B * P));
    GPN_TRY(gpn_buf_ensure(c, sp.bout, sizeof(double) * (size_t)B + sizeof(int) * (size_t)B * 2 + 64));
    GPN_TRY(sync_write(c, sp.bt.p, y_h, sizeof(double) * n));
    GPN_TRY(sync_write(c, sp.bth.p, thetas_h, sizeof(double) * (size_t)B * P));
    SpecNewtonArgs a;
    a.N = N; a.n = n; a.ldv = ldv; a.P = P; a.kind = kind; a.B = B;
    a.Vt = (const double*)sp.Vt.p; a.lam = (const double*)sp.lam.p; a.y = (const double*)sp.bt.p; a.thetas = (const double*)sp.bth.p;
    a.tol = tol; a.phif0 = phif0; a.scale0 = scale;
    a.logq = (double*)sp.bout.p;
    a.iters = (int*)(a.logq + B);
    a.info = a.iters + B;
    const size_t ysz = std::max((size_t)n * (n + 1), (size_t)KC * n4 + 2 * KC);
    const size_t smem = sizeof(double) * ((size_t)n * (n + 1) + ysz + 8 * (size_t)n4 + 40);
    static size_t configured = 0;
    if (smem > configured) {
        GPN_CK(cudaFuncSetAttribute(spectral_newton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = std::min(B, 4 * c->num_sms);
    spectral_newton_kernel<<<grid, 256, smem, c->stream>>>(a);
    c->launches++;
    GPN_CK(cudaGetLastError());
    c->flops += (double)B * n * (double)n * N;
    GPN_TRY(sync_read(c, logq_h, a.logq, sizeof(double) * (size_t)B));
    GPN_TRY(sync_read(c, iters_h, a.iters, sizeof(int) * (size_t)B));
    GPN_TRY(sync_read(c, info_h, a.info, sizeof(int) * (size_t)B));
    c->last_route = 7;
    return 0;
}

}   // extern "C"
