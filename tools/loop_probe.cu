// loop_probe -- decomposes the cost of the bulk kernel's fast path on sm_100a, one ingredient at
// a time.  Every variant runs the 8x4 micro-tile step structure (32 DFMA.RM per k step per thread,
// 128-thread CTAs, 3 CTAs/SM unless stated) and reports cycles per warp-step per SM sub-partition
// (at the max clock) plus relaxations/s.  The instruction mix of every variant is checked in SASS
// (cuobjdump -sass loop_probe | tools/count_sass.sh) before its number is trusted.
//
//   V0  chain DFMA only (o = fma(a,b,-o)), operands in registers          FP64-pipe floor
//   V1  V0 + the 6 LDS.128 operand fetches per step
//   V2  V1 + 16 LOP3 per step on registers that do NOT depend on the DFMAs  (pure pipe overlap)
//   V3  filter DFMAs (o constant) + 16 LOP3 sign tree, accumulated, one vote per 16 steps
//   V4  V3 with a vote + branch every step                                 (one-level filter)
//   V5  two-level tree (20 LOP3) + vote + branch every step                (what the kernel ships)
//   V6  sign words combined on the FMA pipe: acc = fma(hi_as_float, 0.0f, acc)  (-0 + -0 = -0)
//   V7  V5, sign words consumed two steps late (software pipelined reduction)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ int lop_and3(int a, int b, int c) {
    int d; asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

template <int V, int CTAS>
__global__ void __launch_bounds__(128, CTAS) k_loop(double *out, int *outm, int iters, const double *src) {
    __shared__ __align__(16) double As[2][16][64];
    __shared__ __align__(16) double Bs[2][16][64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = tid; i < 2 * 16 * 64; i += 128) { (&As[0][0][0])[i] = src[i & 1023]; (&Bs[0][0][0])[i] = src[1024 + (i & 1023)]; }
    __syncthreads();
    double o[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = 1.3 + 1e-4 * (r * 4 + c) + 1e-6 * tx;
    int fired = 0;
    int junk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) junk[i] = tid * 7 + i;
    double av[8], bv[4];
#pragma unroll
    for (int r = 0; r < 8; ++r) av[r] = As[0][0][ty * 8 + r];
#pragma unroll
    for (int c = 0; c < 4; ++c) bv[c] = Bs[0][0][(c >> 1) * 32 + tx * 2 + (c & 1)];
    int acc_all = -1;
    float facc[4] = {-0.0f, -0.0f, -0.0f, -0.0f};
    int hprev[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
#pragma unroll 2
        for (int kk = 0; kk < 16; ++kk) {
            if (V >= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { const double2 v = *reinterpret_cast<const double2 *>(&As[buf][kk][ty * 8 + q * 2]); av[q * 2] = v.x; av[q * 2 + 1] = v.y; }
#pragma unroll
                for (int q = 0; q < 2; ++q) { const double2 v = *reinterpret_cast<const double2 *>(&Bs[buf][kk][q * 32 + tx * 2]); bv[q * 2] = v.x; bv[q * 2 + 1] = v.y; }
            }
            if (V <= 2 || V == 9 || V == 11 || V == 12) {
                if (V == 12) {
                    // like V9, but the LOP3s read the words the PREVIOUS step's DFMAs wrote (one step old)
                    int w[8], x[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        w[r] = lop_and3(__double2hiint(o[r][0]), __double2hiint(o[r][1]), __double2hiint(o[r][2]));
                        x[r] = __double2hiint(o[r][3]);
                    }
                    int u0 = lop_and3(w[0], x[0], x[1]), u1 = lop_and3(w[1], x[2], x[3]);
                    int u2 = lop_and3(w[2], x[4], x[5]), u3 = lop_and3(w[3], x[6], x[7]);
                    int w0 = lop_and3(u0, u1, w[4]), w1 = lop_and3(u2, u3, w[5]);
                    int x0 = lop_and3(w0, w1, w[6]);
                    acc_all = lop_and3(acc_all, x0, w[7]);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[r][c] = __fma_rd(av[r], bv[c], -o[r][c]);
                if (V == 9 || V == 11) {
                    // 16 LOP3 per step reading the hi (V9) / lo (V11) words the in-place DFMAs just wrote
                    int w[8], x[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        w[r] = (V == 9) ? lop_and3(__double2hiint(o[r][0]), __double2hiint(o[r][1]), __double2hiint(o[r][2]))
                                        : lop_and3(__double2loint(o[r][0]), __double2loint(o[r][1]), __double2loint(o[r][2]));
                        x[r] = (V == 9) ? __double2hiint(o[r][3]) : __double2loint(o[r][3]);
                    }
                    int u0 = lop_and3(w[0], x[0], x[1]), u1 = lop_and3(w[1], x[2], x[3]);
                    int u2 = lop_and3(w[2], x[4], x[5]), u3 = lop_and3(w[3], x[6], x[7]);
                    int w0 = lop_and3(u0, u1, w[4]), w1 = lop_and3(u2, u3, w[5]);
                    int x0 = lop_and3(w0, w1, w[6]);
                    acc_all = lop_and3(acc_all, x0, w[7]);
                }
                if (V == 2) {
                    // 16 LOP3 per step on integer state that only depends on the loaded operands
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        junk[i] = lop_and3(junk[i], __double2loint(av[i]), __double2loint(bv[i & 3]) | 0x55);
                        junk[i] = lop_and3(junk[i], __double2hiint(av[i]) | 3, kk + it);
                    }
                }
            } else {
                int hi[8][4], lo[8][4];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double d = __fma_rd(av[r], bv[c], -o[r][c]);
                        hi[r][c] = __double2hiint(d); lo[r][c] = __double2loint(d);
                    }
                if (V == 3 || V == 4) {
                    int t[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) t[r] = lop_and3(hi[r][0], hi[r][1], hi[r][2]);
                    int u0 = lop_and3(t[0], hi[0][3], hi[1][3]), u1 = lop_and3(t[1], hi[2][3], hi[3][3]);
                    int u2 = lop_and3(t[2], hi[4][3], hi[5][3]), u3 = lop_and3(t[3], hi[6][3], hi[7][3]);
                    int w0 = lop_and3(u0, u1, t[4]), w1 = lop_and3(u2, u3, t[5]);
                    int x0 = lop_and3(w0, w1, t[6]);
                    if (V == 3) {
                        acc_all = lop_and3(acc_all, x0, t[7]);
                        if ((kk & 15) == 15) {
                            if (__builtin_expect(__any_sync(0xffffffffu, acc_all >= 0), 0)) {
                                fired++;
#pragma unroll
                                for (int r = 0; r < 8; ++r) o[r][0] += 1e-9;
                            }
                            acc_all = -1;
                        }
                    } else {
                        const int acc = x0 & t[7];
                        if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                            fired++;
#pragma unroll
                            for (int r = 0; r < 8; ++r) o[r][0] += 1e-9;
                        }
                    }
                } else if (V == 5) {
                    int accr[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) accr[r] = (hi[r][0] & hi[r][1]) & (hi[r][2] & hi[r][3]);
                    const int acc = ((accr[0] & accr[1]) & (accr[2] & accr[3])) & ((accr[4] & accr[5]) & (accr[6] & accr[7]));
                    if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                        fired++;
#pragma unroll
                        for (int r = 0; r < 8; ++r) if (accr[r] >= 0) o[r][0] += 1e-9;
                    }
                } else if (V == 6) {
                    // FMA pipe: (+-0 from hi*0) + acc; stays -0 only while every sign bit is set
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            asm("fma.rn.f32 %0, %1, 0f00000000, %0;" : "+f"(facc[c]) : "f"(__int_as_float(hi[r][c])));
                    const int acc = (__float_as_int(facc[0]) & __float_as_int(facc[1])) & (__float_as_int(facc[2]) & __float_as_int(facc[3]));
                    if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                        fired++;
#pragma unroll
                        for (int r = 0; r < 8; ++r) o[r][0] += 1e-9;
#pragma unroll
                        for (int c = 0; c < 4; ++c) facc[c] = -0.0f;
                    }
                } else if (V == 10) {
                    // chain form: 8 accumulators, acc_j = acc_j & h1 & h2 (one operand is never a DFMA result)
                    int a8[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) a8[r] = lop_and3(-1 - (kk & 0), hi[r][0], hi[r][1]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) a8[r] = lop_and3(a8[r], hi[r][2], hi[r][3]);
                    const int acc = lop_and3(lop_and3(a8[0], a8[1], a8[2]), lop_and3(a8[3], a8[4], a8[5]), a8[6] & a8[7]);
                    if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                        fired++;
#pragma unroll
                        for (int r = 0; r < 8; ++r) o[r][0] += 1e-9;
                    }
                } else if (V == 7) {
                    // reduce this step's words to 8 row words now, finish LAST step's reduction + vote
                    int accr[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) accr[r] = (hi[r][0] & hi[r][1]) & (hi[r][2] & hi[r][3]);
                    const int acc = ((hprev[0] & hprev[1]) & (hprev[2] & hprev[3])) & ((hprev[4] & hprev[5]) & (hprev[6] & hprev[7]));
                    if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                        fired++;
#pragma unroll
                        for (int r = 0; r < 8; ++r) if (hprev[r] >= 0) o[r][0] += 1e-9;
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) hprev[r] = accr[r];
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s += o[r][c];
    int js = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) js += junk[i] + hprev[i];
    out[blockIdx.x * blockDim.x + tid] = s;
    outm[blockIdx.x * blockDim.x + tid] = fired + js + acc_all;
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

static double *g_out; static int *g_outm; static double *g_src; static int g_sms, g_clk;

template <int V, int CTAS>
static void run(const char *name, bool last = false) {
    const int it = 1024, g = g_sms * CTAS * 4;       // 4 waves of resident CTAs
    const double ms = time_ms([&] { k_loop<V, CTAS><<<g, 128>>>(g_out, g_outm, it, g_src); }, 3);
    CK(cudaGetLastError());
    const double steps_per_smsp = (double)g * 4 /*warps*/ * 16.0 * it / (g_sms * 4.0);
    const double cyc = ms * 1e-3 * g_clk * 1e3;
    printf("  \"%s\": {\"ms\": %.3f, \"cycles_per_warp_step_per_smsp\": %.1f, \"relax_per_s\": %.4e}%s\n", name, ms,
           cyc / steps_per_smsp, (double)g * 128 * 32 * 16 * it / (ms * 1e-3), last ? "" : ",");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    cudaDeviceGetAttribute(&g_clk, cudaDevAttrClockRate, 0);
    CK(cudaMalloc(&g_out, sizeof(double) * g_sms * 32 * 128)); CK(cudaMalloc(&g_outm, sizeof(int) * g_sms * 32 * 128));
    double h[2048]; srand(1);
    for (int i = 0; i < 2048; ++i) h[i] = 0.9 + 0.2 * (rand() / (double)RAND_MAX);
    CK(cudaMalloc(&g_src, sizeof(h))); CK(cudaMemcpy(g_src, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f,\n", p.name, g_sms, g_clk / 1000.0);
    run<0, 3>("v0_dfma_chain");
    run<1, 3>("v1_dfma_lds");
    run<2, 3>("v2_dfma_lds_16lop3_independent");
    run<3, 3>("v3_filter_tree16_vote_per_16");
    run<4, 3>("v4_filter_tree16_vote_per_step");
    run<5, 3>("v5_filter_two_level_vote_per_step");
    run<5, 2>("v5_2cta");
    run<5, 4>("v5_4cta");
    run<6, 3>("v6_filter_ffma_zero");
    run<7, 3>("v7_two_level_pipelined");
    run<9, 3>("v9_inplace_chain_16lop3_on_hi_words");
    run<11, 3>("v11_inplace_chain_16lop3_on_lo_words");
    run<12, 3>("v12_inplace_chain_16lop3_on_one_step_old_hi_words");
    run<10, 3>("v10_filter_chain_form_vote_per_step", true);
    printf("}\n");
    return 0;
}
