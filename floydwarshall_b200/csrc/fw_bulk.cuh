// Phase 3 of the exact-order blocked solve: every entry with row and column
// outside the k-block [b0, b0+B) takes the B steps of the block in ascending
// order against the two read-only snapshot panels
//     n = Cp[i][kk] * Rw[kk][j];   if (R[i][j] < n) { R[i][j] = n; mid = kk; }
// and resolves its next-hop once, at the end:  NX[i][j] = NCp[i][mid].
// This is >99 % of the work of a large solve.  Relaxations fire rarely (about
// 7 times per ENTRY over a whole solve, i.e. ~7/N of them), so each k step
// first runs a 1-DFMA-per-relaxation filter that proves "nothing fires here"
// for the whole warp and only falls into the exact mul/compare/select path
// when some lane has a candidate.  The fast path is FP64-pipe bound (DFMA).
//
// CTA = 128 threads, 64x64 output tile held in registers (8 rows x 4 columns
// per thread + 32 mids); the panels stream through shared memory in k-chunks
// of 16 with cp.async double buffering.  Thread (ty,tx): rows r*8+ty (r<8) so
// the two ty of a warp hit different banks; columns cq*32 + tx*2 + e so global
// and shared accesses are contiguous 16-byte pieces across the 16 tx lanes.
#pragma once
#include "fw_common.cuh"

namespace fw {

struct BulkArgs {
    double *rate;
    int32_t *next;
    int32_t *mid;          // nullable
    long long ld;
    int npad;
    int b0;
    const double *Cp;      // N x B
    const int32_t *NCp;    // N x B
    const double *Rw;      // B x N
    long long ldw;
};

constexpr int BULK_T = 64;    // tile edge
constexpr int BULK_KC = 16;   // k-chunk
constexpr int BULK_AP = 18;   // shared pitch of the A chunk rows (doubles)

__global__ void __launch_bounds__(128, 3) fw_bulk_kernel(BulkArgs a) {
    __shared__ __align__(16) double As[2][BULK_T][BULK_AP];
    __shared__ __align__(16) double Bs[2][BULK_KC][BULK_T];

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int tb = a.b0 / BULK_T;  // first of the two tile indices covered by the k-block
    int ti = blockIdx.y, tj = blockIdx.x;
    ti = ti < tb ? ti : ti + FW_B / BULK_T;
    tj = tj < tb ? tj : tj + FW_B / BULK_T;
    const int i0 = ti * BULK_T, j0 = tj * BULK_T;
    const long long ld = a.ld;

    auto load_chunk = [&](int ch, int buf) {
        const int kk0 = ch * BULK_KC;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int p = tid + 128 * t;
            const int row = p >> 3, part = p & 7;
            cp_async16(&As[buf][row][part * 2], a.Cp + (long long)(i0 + row) * FW_B + kk0 + part * 2);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int p = tid + 128 * t;
            const int kk = p >> 5, part = p & 31;
            cp_async16(&Bs[buf][kk][part * 2], a.Rw + (long long)(kk0 + kk) * a.ldw + j0 + part * 2);
        }
    };

    load_chunk(0, 0);
    cp_async_commit();

    double o[8][4];
    int m[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long ro = (long long)(i0 + r * 8 + ty) * ld + j0 + tx * 2;
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
            const double2 v = *reinterpret_cast<const double2 *>(a.rate + ro + cq * 32);
            o[r][cq * 2] = v.x; o[r][cq * 2 + 1] = v.y;
            m[r][cq * 2] = -1; m[r][cq * 2 + 1] = -1;
        }
    }
    if (ti == tj) {  // the matrix diagonal crosses this tile: hold it as NaN (see fw_common.cuh)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (r * 8 + ty == (c >> 1) * 32 + tx * 2 + (c & 1)) o[r][c] = qnan();
    }

    constexpr int NCH = FW_B / BULK_KC;
    for (int ch = 0; ch < NCH; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < NCH) {
            load_chunk(ch + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll 2
        for (int k2 = 0; k2 < BULK_KC / 2; ++k2) {
            double2 a2[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) a2[r] = *reinterpret_cast<const double2 *>(&As[buf][r * 8 + ty][k2 * 2]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = k2 * 2 + h;
                const double2 b01 = *reinterpret_cast<const double2 *>(&Bs[buf][kk][tx * 2]);
                const double2 b23 = *reinterpret_cast<const double2 *>(&Bs[buf][kk][32 + tx * 2]);
                // Filter (1 DFMA per relaxation): with round-toward-minus-infinity,
                //   sign(fma(a, b, -o)) is clear  <=>  exact(a*b) > o  (or a positive-signed NaN),
                // and exact(a*b) <= o implies RN(a*b) <= o, i.e. the reference's strict test
                // o < a*b (Algorithms.hs:55,61) cannot fire.  An exact tie gives -0 under RM.
                // So "all sign bits set" proves that none of these 32 relaxations fires.
                int acc = -1;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double av = h ? a2[r].y : a2[r].x;
                    const double d0 = __fma_rd(av, b01.x, -o[r][0]);
                    const double d1 = __fma_rd(av, b01.y, -o[r][1]);
                    const double d2 = __fma_rd(av, b23.x, -o[r][2]);
                    const double d3 = __fma_rd(av, b23.y, -o[r][3]);
                    acc &= __double2hiint(d0) & __double2hiint(d1);
                    acc &= __double2hiint(d2) & __double2hiint(d3);
                }
                if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                    // exact path: one rounded multiply, strict compare, ascending k
                    const int kloc = ch * BULK_KC + kk;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const double av = h ? a2[r].y : a2[r].x;
                        double n;
                        n = av * b01.x; if (o[r][0] < n) { o[r][0] = n; m[r][0] = kloc; }
                        n = av * b01.y; if (o[r][1] < n) { o[r][1] = n; m[r][1] = kloc; }
                        n = av * b23.x; if (o[r][2] < n) { o[r][2] = n; m[r][2] = kloc; }
                        n = av * b23.y; if (o[r][3] < n) { o[r][3] = n; m[r][3] = kloc; }
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue: values (vector stores), next-hops / mids only where a relaxation fired ----
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = i0 + r * 8 + ty;
        const long long ro = (long long)row * ld + j0 + tx * 2;
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
            const int m0 = m[r][cq * 2], m1 = m[r][cq * 2 + 1];
            if (m0 >= 0 || m1 >= 0) {
                const long long eo = ro + cq * 32;
                if (m0 >= 0 && m1 >= 0) {
                    *reinterpret_cast<double2 *>(a.rate + eo) = make_double2(o[r][cq * 2], o[r][cq * 2 + 1]);
                } else if (m0 >= 0) {
                    a.rate[eo] = o[r][cq * 2];
                } else {
                    a.rate[eo + 1] = o[r][cq * 2 + 1];
                }
                if (m0 >= 0) {
                    a.next[eo] = a.NCp[(long long)row * FW_B + m0];
                    if (a.mid) a.mid[eo] = a.b0 + m0;
                }
                if (m1 >= 0) {
                    a.next[eo + 1] = a.NCp[(long long)row * FW_B + m1];
                    if (a.mid) a.mid[eo + 1] = a.b0 + m1;
                }
            }
        }
    }
}

}  // namespace fw
