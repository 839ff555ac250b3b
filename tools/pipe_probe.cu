// pipe_probe -- isolates the per-instruction cost of the relaxation's pieces on sm_100a.
// Each kernel runs `iters` iterations of 16 independent chains per thread; output = ops/s
// and cycles per warp-instruction per SM sub-partition (at the reported clock).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;

// A: DSETP only: p = x[i] < t ; cnt += p   (DSETP + predicated IADD)
__global__ void __launch_bounds__(256) k_dsetp(int *out, int iters, const double *src) {
    double x[16]; int cnt[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = src[(threadIdx.x + i) & 1023]; cnt[i] = 0; }
    double t = src[threadIdx.x & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("{ .reg .pred p; setp.lt.f64 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt[i]) : "d"(x[i]), "d"(t));
        }
        t += 1e-9;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// B: 64-bit integer compare: setp.lt.s64 + predicated add
__global__ void __launch_bounds__(256) k_isetp64(int *out, int iters, const u64 *src) {
    u64 x[16]; int cnt[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = src[(threadIdx.x + i) & 1023]; cnt[i] = 0; }
    u64 t = src[threadIdx.x & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("{ .reg .pred p; setp.lt.s64 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt[i]) : "l"(x[i]), "l"(t));
        }
        t += 3;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// C: predicated-add only (ALU baseline): 32-bit compare + add
__global__ void __launch_bounds__(256) k_isetp32(int *out, int iters, const u64 *src) {
    int x[16]; int cnt[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = (int)src[(threadIdx.x + i) & 1023]; cnt[i] = 0; }
    int t = (int)src[threadIdx.x & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("{ .reg .pred p; setp.lt.s32 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt[i]) : "r"(x[i]), "r"(t));
        }
        t += 3;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// D: relax chains in registers, fp compare:  n = o*a (DMUL); if (o2 < n) o2 = n, m = it
template <int MODE>  // 0 = fp64 compare, 1 = s64 compare of the bit patterns, 2 = no compare (DMUL + unconditional max by hi word)
__global__ void __launch_bounds__(256) k_relax_reg(double *out, int iters, const double *src) {
    double o[16]; int m[16];
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i] = src[(threadIdx.x + i) & 1023]; m[i] = -1; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = src[(threadIdx.x * 3 + i) & 1023]; b[i] = src[(threadIdx.x * 7 + i) & 1023]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double n = a[r] * b[c];
                if (MODE == 0) {
                    if (o[r * 4 + c] < n) { o[r * 4 + c] = n; m[r * 4 + c] = it; }
                } else if (MODE == 1) {
                    if (__double_as_longlong(o[r * 4 + c]) < __double_as_longlong(n)) { o[r * 4 + c] = n; m[r * 4 + c] = it; }
                } else {
                    if (__double2hiint(o[r * 4 + c]) < __double2hiint(n)) { o[r * 4 + c] = n; m[r * 4 + c] = it; }
                }
            }
        // perturb operands a little so nothing is loop invariant
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = __longlong_as_double(__double_as_longlong(a[i]) ^ (it & 1)); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += o[i] + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// E: FSEL/SEL throughput: x = p ? y : x with a precomputed predicate pattern
__global__ void __launch_bounds__(256) k_sel(int *out, int iters, const u64 *src) {
    int x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = (int)src[(threadIdx.x + i) & 1023]; y[i] = (int)src[(threadIdx.x * 5 + i) & 1023]; }
    int t = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        bool p = ((it + t) & 64) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.b32 %0, %1, %0, q; }" : "+r"(x[i]) : "r"(y[i]), "r"((int)p));
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *outd; int *outi; double *src;
    CK(cudaMalloc(&outd, sizeof(double) * sms * 8 * 256));
    CK(cudaMalloc(&outi, sizeof(int) * sms * 8 * 256));
    double h[1024]; srand(2);
    for (int i = 0; i < 1024; ++i) h[i] = 0.9 + 0.2 * (rand() / (double)RAND_MAX);
    CK(cudaMalloc(&src, sizeof(h))); CK(cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice));
    const int g = sms * 8, it = 4096;
    const double nops = (double)g * 256 * 16 * it;       // per-thread ops of one kind
    const double warps_per_smsp = (double)g * 8 / (sms * 4);
    auto report = [&](const char *name, double ms, double ops_per_iter_elem) {
        const double tot = nops * ops_per_iter_elem;
        const double cyc = ms * 1e-3 * clk * 1e3;         // SM cycles elapsed (at max clock)
        const double winstr_per_smsp = warps_per_smsp * 16.0 * it * ops_per_iter_elem;
        printf("  \"%s\": {\"ms\": %.3f, \"ops_per_s\": %.4e, \"cycles_per_warp_op_per_smsp_at_max_clock\": %.2f},\n",
               name, ms, tot / (ms * 1e-3), cyc / winstr_per_smsp);
    };
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f,\n", p.name, sms, clk / 1000.0);
    report("dsetp_plus_padd", time_ms([&] { k_dsetp<<<g, 256>>>(outi, it, src); }, 3), 1);
    report("isetp64_plus_padd", time_ms([&] { k_isetp64<<<g, 256>>>(outi, it, (const u64 *)src); }, 3), 1);
    report("isetp32_plus_padd", time_ms([&] { k_isetp32<<<g, 256>>>(outi, it, (const u64 *)src); }, 3), 1);
    report("sel32", time_ms([&] { k_sel<<<g, 256>>>(outi, it, (const u64 *)src); }, 3), 1);
    report("relax_reg_fpcmp", time_ms([&] { k_relax_reg<0><<<g, 256>>>(outd, it, src); }, 3), 1);
    report("relax_reg_s64cmp", time_ms([&] { k_relax_reg<1><<<g, 256>>>(outd, it, src); }, 3), 1);
    report("relax_reg_hi32cmp", time_ms([&] { k_relax_reg<2><<<g, 256>>>(outd, it, src); }, 3), 1);
    printf("  \"note\": \"cycles computed at max clock; divide by (actual/max) clock ratio\"\n}\n");
    return 0;
}
