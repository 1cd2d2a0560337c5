// ubench_fp64.cu — do DMMA (mma.sync.m8n8k4.f64) and DFMA share one fp64 datapath on B200?
// Times (a) DMMA only, (b) DFMA only, (c) both interleaved in every warp, (d) half the warps each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
    double c[8][2], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-9 + i;
    const bool do_mma = (MODE == 0) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_fma = (MODE == 1) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
        }
        if (do_fma) {
            // 8 DMMA = 8*256 FMA per warp = 64 per lane -> 4 rounds of 16 independent DFMA
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int MODE> float run(double *out, int grid, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, iters / 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int grid = p.multiProcessorCount * 4, iters = 20000;
    double *out; cudaMalloc(&out, 8);
    double warps = (double)grid * 8;
    double fl = warps * iters * 8.0 * 512.0;   // flops of one "unit" (8 DMMA or 64 DFMA/lane) per warp-iter
    float a = run<0>(out, grid, iters), b = run<1>(out, grid, iters), c = run<2>(out, grid, iters), d = run<3>(out, grid, iters);
    printf("SMs %d clock %d kHz\n", p.multiProcessorCount, p.clockRate);
    printf("DMMA only        : %8.3f ms  %6.2f TFLOP/s\n", a, fl / a / 1e9);
    printf("DFMA only        : %8.3f ms  %6.2f TFLOP/s\n", b, fl / b / 1e9);
    printf("both, same warp  : %8.3f ms  %6.2f TFLOP/s (2x work)  -> concurrent if ~max(a,b)=%.3f, serial if ~a+b=%.3f\n",
           c, 2 * fl / c / 1e9, a > b ? a : b, a + b);
    printf("half warps each  : %8.3f ms  %6.2f TFLOP/s (1x work)  -> concurrent if ~%.3f, serial if ~%.3f\n",
           d, fl / d / 1e9, (a > b ? a : b) / 2, (a + b) / 2);
    return 0;
}
