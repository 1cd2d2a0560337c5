// ubench_dmma_occ.cu — DMMA.8x8x4 throughput per SM vs resident warps and independent chains per warp.
// Answers: how many warps / how much ILP does it take to saturate the fp64 tensor pipe on B200?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k(double *out, int iters, double a, double b) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16 / CH; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int CH> void run(double *out, int sms, int warps, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH><<<sms, warps * 32>>>(out, iters / 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<CH><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dm = (double)sms * warps * iters * 16.0;          // DMMAs
    double cyc_per_sm = ms * 1e-3 * 1.965e9;
    printf("chains %2d warps/SM %2d : %7.3f ms  %6.2f TFLOP/s  %6.2f cycles/DMMA/SM  (%.1f cyc per warp-DMMA)\n",
           CH, warps, ms, dm * 512 / ms / 1e9, cyc_per_sm / (warps * iters * 16.0), cyc_per_sm / (iters * 16.0));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, 8);
    int sms = p.multiProcessorCount, iters = 20000;
    int ws[] = {1, 2, 4, 8, 16, 32};
    for (int w : ws) { run<1>(out, sms, w, iters); }
    for (int w : ws) { run<2>(out, sms, w, iters); }
    for (int w : ws) { run<4>(out, sms, w, iters); }
    for (int w : ws) { run<8>(out, sms, w, iters); }
    for (int w : ws) { run<16>(out, sms, w, iters); }
    return 0;
}
