// dfma_operand_probe -- DFMA issue rate vs. how many of its source operands are fresh registers.
//   P1: fma(A, B, o[i])      A, B fixed registers (operand reuse), 1 fresh operand
//   P2: fma(A, b[i%4], o[i]) 2 fresh operands
//   P3: fma(a[i/4], b[i%4], o[i])  the outer-product pattern of the filter (a reused 4x)
//   P4: P3 with negated addend (-o[i]) as in the filter
// 32 independent DFMAs per iteration, results never consumed (volatile asm keeps them).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int P>
__global__ void __launch_bounds__(128, 3) k(double *out, int iters, const double *src) {
    double o[32], a[8], b[4];
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = src[(tid + i) & 1023];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = src[(tid * 3 + i) & 1023];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = src[(tid * 7 + i) & 1023];
    double sink = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            double d;
            if (P == 1) asm volatile("fma.rm.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a[0]), "d"(b[0]), "d"(o[i]));
            if (P == 2) asm volatile("fma.rm.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a[0]), "d"(b[i & 3]), "d"(o[i]));
            if (P == 3) asm volatile("fma.rm.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a[i >> 2]), "d"(b[i & 3]), "d"(o[i]));
            if (P == 4) asm volatile("{ .reg .f64 t; neg.f64 t, %3; fma.rm.f64 %0, %1, %2, t; }" : "=d"(d) : "d"(a[i >> 2]), "d"(b[i & 3]), "d"(o[i]));
            if (i == 31) sink += d;
        }
    }
    out[blockIdx.x * blockDim.x + tid] = sink;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *out, *src;
    CK(cudaMalloc(&out, sizeof(double) * sms * 12 * 128));
    double h[1024]; srand(1);
    for (int i = 0; i < 1024; ++i) h[i] = 0.9 + 0.2 * (rand() / (double)RAND_MAX);
    CK(cudaMalloc(&src, sizeof(h))); CK(cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice));
    const int it = 16384, g = sms * 12;
    const double n = (double)g * 128 * 32 * it;
    printf("{\"gpu\": \"%s\"", p.name);
    printf(", \"P1_one_fresh\": %.4e", n / (time_ms([&] { k<1><<<g, 128>>>(out, it, src); }, 3) * 1e-3));
    printf(", \"P2_two_fresh\": %.4e", n / (time_ms([&] { k<2><<<g, 128>>>(out, it, src); }, 3) * 1e-3));
    printf(", \"P3_outer_product\": %.4e", n / (time_ms([&] { k<3><<<g, 128>>>(out, it, src); }, 3) * 1e-3));
    printf(", \"P4_outer_product_neg\": %.4e", n / (time_ms([&] { k<4><<<g, 128>>>(out, it, src); }, 3) * 1e-3));
    printf(", \"unit\": \"DFMA/s (peak 1.69e13)\"}\n");
    return 0;
}
