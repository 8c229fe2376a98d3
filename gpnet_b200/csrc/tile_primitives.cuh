// Device-side building blocks shared by the GEMM (gemm_f64.cu) and the dataflow Cholesky (potrf_dataflow.cu):
// mbarrier / TMA / DMMA wrappers and the bank-conflict-free fragment loaders for TMA 128B-swizzled tiles.
#pragma once
#include "common.cuh"

namespace tp {

constexpr int KSTEP = 16;   // doubles per k-step == one 128-byte swizzled row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Fragment loader for one operand.  NB = number of 8-wide blocks this warp needs (8 for the 64-row A side,
// 4 for the 32-column B side); `mn0` = first tile row/col of the warp.  Result f[b][u], u = k sub-slot.
// k index carried by lane t in half h, sub-slot u: k = 2t + 8h + u  (same for both operands).
template <int NB, bool KMAJOR>
__device__ __forceinline__ void load_frags(double (&f)[NB][2], uint32_t base, int mn0, int h, int g, int t) {
    if (KMAJOR) {
        const int pg = (g >> 1) | ((g & 1) << 2);   // row permutation inside an 8-row group (bank-conflict-free LDS.128)
        const uint32_t a0 = base + (uint32_t)(mn0 + pg) * 128u + (uint32_t)(((t + 4 * h) ^ pg) << 4);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double2 v = lds128(a0 + b * 8 * 128);
            f[b][0] = v.x;
            f[b][1] = v.y;
        }
    } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = 2 * t + 8 * h + u;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int col = mn0 + b * 8 + g;          // tile-local mn index
                const int box = col >> 4, cw = col & 15;  // 16-column TMA boxes of 16 k-rows x 128 B
                const uint32_t a = base + box * 2048 + k * 128 + ((((cw >> 1) ^ (k & 7)) << 4) | ((cw & 1) << 3));
                f[b][u] = lds64(a);
            }
        }
    }
}


__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// barrier over the 256 consumer threads only (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

}   // namespace tp
