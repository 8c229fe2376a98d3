// FP64 ceiling microbenchmark for B200 (sm_100a): register-resident DMMA.8x8x4 and DFMA loops.
// Measures the sustained FP64 issue-rate ceiling that bench.py's roofline is reported against
// (MEASURED_PEAKS.json carries no FP64 entry).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a0, double b0) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double a = a0, b = b0 + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 4096;
    int warps_list[] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; ++wi) {
        int warps = warps_list[wi]; int threads = warps * 32;
        {
            double ms = time_ms([&] { dmma_kernel<8><<<sms, threads>>>(out, iters, 1.0000001, 0.9999999); }, 5);
            double fl = 512.0 * 8 * iters * warps * sms;
            printf("DMMA884 ILP8  warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        }
        {
            double ms = time_ms([&] { dmma_kernel<16><<<sms, threads>>>(out, iters, 1.0000001, 0.9999999); }, 5);
            double fl = 512.0 * 16 * iters * warps * sms;
            printf("DMMA884 ILP16 warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        }
        {
            double ms = time_ms([&] { dfma_kernel<16><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 32 * 16 * iters * warps * sms;
            printf("DFMA    ILP16 warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        }
    }
    // sustained (4 s) DMMA loop, 8 warps/SM ILP16, to see the power-capped rate
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int reps = 0; CK(cudaEventRecord(e0));
        double fl1 = 512.0 * 16 * iters * 8 * 8 * sms;
        float ms = 0;
        do {
            for (int r = 0; r < 20; ++r) dmma_kernel<16><<<sms, 256>>>(out, iters * 8, 1.0000001, 0.9999999);
            reps += 20; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        } while (ms < 4000.f);
        printf("DMMA884 sustained %.1f s: %7.2f TFLOP/s\n", ms * 1e-3, fl1 * reps / ms * 1e-9);
    }
    return 0;
}
