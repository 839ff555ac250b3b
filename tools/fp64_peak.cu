// fp64_peak -- measures the FP64 CUDA-core ceilings the roofline is quoted against
// (MEASURED_PEAKS.json carries only HBM and bf16 numbers; SURVEY.md 8d asks for
// a DFMA chain and a DMUL+DSETP pair measured on the same box).
//   dfma      : independent DFMA chains            -> FLOP/s (2 per DFMA)
//   dmul      : independent DMUL chains            -> DMUL/s
//   relax     : mul + strict compare + select of value and mid, operands from
//               shared memory, 8x4 register micro-tile (the bulk kernel's inner
//               loop without global traffic)       -> relaxations/s
//   relax_val : same without the mid select
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmul(double *out, int iters, double a) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 1.0 + threadIdx.x * 1e-6 + i * 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __dmul_rn(x[i], a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma_rm(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fma_rd(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the bulk kernel's fast path alone: 32 DFMA.RM + sign-word AND tree + warp vote per k step
template <int CTAS>
__global__ void __launch_bounds__(128, CTAS) k_filter(double *out, int *outm, int iters, const double *src) {
    __shared__ __align__(16) double As[2][64][18];
    __shared__ __align__(16) double Bs[2][16][64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = tid; i < 2 * 64 * 16; i += 128) { As[i >> 10][(i >> 4) & 63][i & 15] = src[i & 1023]; Bs[i >> 10][(i >> 6) & 15][i & 63] = src[1024 + (i & 1023)]; }
    __syncthreads();
    double o[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = 1.3 + 1e-4 * (r + c + tx);   // above every product: the filter never fires
    int fired = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
            double av[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) av[r] = As[buf][r * 8 + ty][kk];
            const double2 b01 = *reinterpret_cast<const double2 *>(&Bs[buf][kk][tx * 2]);
            const double2 b23 = *reinterpret_cast<const double2 *>(&Bs[buf][kk][32 + tx * 2]);
            int hi[8][4];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                hi[r][0] = __double2hiint(__fma_rd(av[r], b01.x, -o[r][0]));
                hi[r][1] = __double2hiint(__fma_rd(av[r], b01.y, -o[r][1]));
                hi[r][2] = __double2hiint(__fma_rd(av[r], b23.x, -o[r][2]));
                hi[r][3] = __double2hiint(__fma_rd(av[r], b23.y, -o[r][3]));
            }
            int accr[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) accr[r] = (hi[r][0] & hi[r][1]) & (hi[r][2] & hi[r][3]);
            const int acc = ((accr[0] & accr[1]) & (accr[2] & accr[3])) & ((accr[4] & accr[5]) & (accr[6] & accr[7]));
            if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                fired++;
#pragma unroll
                for (int r = 0; r < 8; ++r) o[r][0] += 1e-9;   // keep o live/mutable
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

template <bool MID>
__global__ void __launch_bounds__(128, 3) k_relax(double *out, int *outm, int iters, const double *src) {
    __shared__ __align__(16) double As[2][64][18];
    __shared__ __align__(16) double Bs[2][16][64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = tid; i < 2 * 64 * 16; i += 128) { As[i >> 10][(i >> 4) & 63][i & 15] = src[i & 1023]; Bs[i >> 10][(i >> 6) & 15][i & 63] = src[1024 + (i & 1023)]; }
    __syncthreads();
    double o[8][4];
    int m[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { o[r][c] = 0.97 + 1e-4 * (r + c + tx); m[r][c] = -1; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            double2 a2[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) a2[r] = *reinterpret_cast<const double2 *>(&As[buf][r * 8 + ty][k2 * 2]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = k2 * 2 + h;
                const double2 b01 = *reinterpret_cast<const double2 *>(&Bs[buf][kk][tx * 2]);
                const double2 b23 = *reinterpret_cast<const double2 *>(&Bs[buf][kk][32 + tx * 2]);
                const int kloc = it * 16 + kk;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double av = h ? a2[r].y : a2[r].x;
                    double n;
                    n = av * b01.x; if (o[r][0] < n) { o[r][0] = n; if (MID) m[r][0] = kloc; }
                    n = av * b01.y; if (o[r][1] < n) { o[r][1] = n; if (MID) m[r][1] = kloc; }
                    n = av * b23.x; if (o[r][2] < n) { o[r][2] = n; if (MID) m[r][2] = kloc; }
                    n = av * b23.y; if (o[r][3] < n) { o[r][3] = n; if (MID) m[r][3] = kloc; }
                }
            }
        }
    }
    double s = 0; int ms = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { s += o[r][c]; ms += m[r][c]; }
    out[blockIdx.x * blockDim.x + tid] = s;
    outm[blockIdx.x * blockDim.x + tid] = ms;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *out; int *outm; double *src;
    CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 256));
    CK(cudaMalloc(&outm, sizeof(int) * sms * 32 * 256));
    double h[2048];
    srand(1);
    for (int i = 0; i < 2048; ++i) h[i] = 0.9 + 0.2 * (rand() / (double)RAND_MAX);
    CK(cudaMalloc(&src, sizeof(h)));
    CK(cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice));

    const int it1 = 8192;
    const int g1 = sms * 8;
    double t_dfma = time_ms([&] { k_dfma<<<g1, 256>>>(out, it1, 0.999999, 1e-9); }, 5);
    double t_dmul = time_ms([&] { k_dmul<<<g1, 256>>>(out, it1, 0.9999999); }, 5);
    double t_dfma_rm = time_ms([&] { k_dfma_rm<<<g1, 256>>>(out, it1, 0.999999, 1e-9); }, 5);
    const double n1 = (double)g1 * 256 * 16 * it1;
    const int it2 = 2048;
    const int g2 = sms * 3 * 4;
    double t_rel = time_ms([&] { k_relax<true><<<g2, 128>>>(out, outm, it2, src); }, 5);
    double t_relv = time_ms([&] { k_relax<false><<<g2, 128>>>(out, outm, it2, src); }, 5);
    const double n2 = (double)g2 * 128 * 32 * 16 * it2;
    double t_f2 = time_ms([&] { k_filter<2><<<sms * 2 * 4, 128>>>(out, outm, it2, src); }, 5);
    double t_f3 = time_ms([&] { k_filter<3><<<sms * 3 * 4, 128>>>(out, outm, it2, src); }, 5);
    double t_f4 = time_ms([&] { k_filter<4><<<sms * 4 * 4, 128>>>(out, outm, it2, src); }, 5);
    const double nf = (double)128 * 32 * 16 * it2 * sms * 4;
    CK(cudaDeviceSynchronize());
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, "
           "\"dfma_tflops\": %.3f, \"dmul_tops\": %.3f, "
           "\"dfma_rm_tflops\": %.3f, \"filter_relax_per_s_2cta\": %.4e, \"filter_relax_per_s_3cta\": %.4e, \"filter_relax_per_s_4cta\": %.4e, "
           "\"relax_per_s\": %.4e, \"relax_val_only_per_s\": %.4e, "
           "\"nominal_fp64_fma_tflops\": %.2f, \"nominal_relax_ceiling_per_s\": %.4e}\n",
           p.name, sms, clk / 1000.0, 2.0 * n1 / (t_dfma * 1e-3) / 1e12, n1 / (t_dmul * 1e-3) / 1e12,
           2.0 * n1 / (t_dfma_rm * 1e-3) / 1e12, nf * 2 / (t_f2 * 1e-3), nf * 3 / (t_f3 * 1e-3), nf * 4 / (t_f4 * 1e-3),
           n2 / (t_rel * 1e-3), n2 / (t_relv * 1e-3), sms * 64 * 2 * (clk / 1e6) / 1e3 ,
           sms * 32.0 * clk * 1e3);
    return 0;
}
