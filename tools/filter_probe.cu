// filter_probe -- which way of combining the DFMA sign words is cheapest on sm_100a?
// Same structure as fw_bulk_kernel's fast path (8x4 micro-tile, operands from shared memory,
// one warp vote per k step), only the reduction differs:
//   MODE 0  no reduction except one word per step (DFMA issue ceiling of this loop shape)
//   MODE 1  3-input LOP3 tree (what the kernel did in round 1a)
//   MODE 2  serial 2-input AND into 4 accumulators
//   MODE 3  predicate chain  setp.ge.or  on each high word
//   MODE 4  serial 2-input signed max (IMNMX) into 4 accumulators
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(128, 3) k_filter(double *out, int *outm, int iters, const double *src) {
    __shared__ __align__(16) double As[2][16][64];
    __shared__ __align__(16) double Bs[2][16][64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = tid; i < 2 * 16 * 64; i += 128) { (&As[0][0][0])[i] = src[i & 1023]; (&Bs[0][0][0])[i] = src[1024 + (i & 1023)]; }
    __syncthreads();
    double o[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = 1.3 + 1e-4 * (r + c + tx);
    int fired = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
#pragma unroll 2
        for (int kk = 0; kk < 16; ++kk) {
            double av[8], bv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const double2 v = *reinterpret_cast<const double2 *>(&As[buf][kk][ty * 8 + q * 2]); av[q * 2] = v.x; av[q * 2 + 1] = v.y; }
#pragma unroll
            for (int q = 0; q < 2; ++q) { const double2 v = *reinterpret_cast<const double2 *>(&Bs[buf][kk][q * 32 + tx * 2]); bv[q * 2] = v.x; bv[q * 2 + 1] = v.y; }
            int hi[8][4];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) hi[r][c] = __double2hiint(__fma_rd(av[r], bv[c], -o[r][c]));
            bool cand;
            if (MODE == 0) {
                // every DFMA result is consumed by an empty volatile asm (forces the DFMA, emits nothing)
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) asm volatile("" ::"r"(hi[r][c]));
                cand = (hi[0][0] & hi[7][3]) >= 0;
            } else if (MODE == 1) {
                int accr[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) accr[r] = (hi[r][0] & hi[r][1]) & (hi[r][2] & hi[r][3]);
                const int acc = ((accr[0] & accr[1]) & (accr[2] & accr[3])) & ((accr[4] & accr[5]) & (accr[6] & accr[7]));
                cand = acc >= 0;
            } else if (MODE == 2) {
                int a0 = -1, a1 = -1, a2 = -1, a3 = -1;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    asm volatile("and.b32 %0, %0, %1;" : "+r"(a0) : "r"(hi[r][0]));
                    asm volatile("and.b32 %0, %0, %1;" : "+r"(a1) : "r"(hi[r][1]));
                    asm volatile("and.b32 %0, %0, %1;" : "+r"(a2) : "r"(hi[r][2]));
                    asm volatile("and.b32 %0, %0, %1;" : "+r"(a3) : "r"(hi[r][3]));
                }
                cand = ((a0 & a1) & (a2 & a3)) >= 0;
            } else if (MODE == 3) {
                int p = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        asm volatile("{ .reg .pred q; setp.ne.s32 q, %0, 0; setp.ge.or.s32 q, %1, 0, q; selp.s32 %0, 1, 0, q; }" : "+r"(p) : "r"(hi[r][c]));
                cand = p != 0;
            } else {
                int a0 = (int)0x80000000, a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(a0) : "r"(hi[r][0]));
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(a1) : "r"(hi[r][1]));
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(a2) : "r"(hi[r][2]));
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(a3) : "r"(hi[r][3]));
                }
                cand = max(max(a0, a1), max(a2, a3)) >= 0;
            }
            if (__builtin_expect(__any_sync(0xffffffffu, cand), 0)) {
                fired++;
#pragma unroll
                for (int r = 0; r < 8; ++r) o[r][0] += 1e-9;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s += o[r][c];
    out[blockIdx.x * blockDim.x + tid] = s;
    outm[blockIdx.x * blockDim.x + tid] = fired;
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
    double *out; int *outm; double *src;
    CK(cudaMalloc(&out, sizeof(double) * sms * 12 * 128)); CK(cudaMalloc(&outm, sizeof(int) * sms * 12 * 128));
    double h[2048]; srand(1);
    for (int i = 0; i < 2048; ++i) h[i] = 0.9 + 0.2 * (rand() / (double)RAND_MAX);
    CK(cudaMalloc(&src, sizeof(h))); CK(cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice));
    const int it = 2048, g = sms * 12;
    const double n = (double)g * 128 * 32 * 16 * it;
    printf("{\"gpu\": \"%s\"", p.name);
    printf(", \"mode0_dfma_only\": %.4e", n / (time_ms([&] { k_filter<0><<<g, 128>>>(out, outm, it, src); }, 3) * 1e-3));
    printf(", \"mode1_lop3_tree\": %.4e", n / (time_ms([&] { k_filter<1><<<g, 128>>>(out, outm, it, src); }, 3) * 1e-3));
    printf(", \"mode2_and2_chain\": %.4e", n / (time_ms([&] { k_filter<2><<<g, 128>>>(out, outm, it, src); }, 3) * 1e-3));
    printf(", \"mode3_setp_chain\": %.4e", n / (time_ms([&] { k_filter<3><<<g, 128>>>(out, outm, it, src); }, 3) * 1e-3));
    printf(", \"mode4_imnmx_chain\": %.4e", n / (time_ms([&] { k_filter<4><<<g, 128>>>(out, outm, it, src); }, 3) * 1e-3));
    printf(", \"unit\": \"relaxations/s\"}\n");
    return 0;
}
