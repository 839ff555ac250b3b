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
    int npad;              // number of columns
    int b0;                // first pivot of the k-block (global)
    int rows;              // rows held by this shard
    int row0;              // global index of local row 0
    int blk_r0;            // LOCAL row of pivot b0, INT_MAX if the shard does not hold it
    const double *Cp;      // N x B
    const int32_t *NCp;    // N x B
    const double *Rw;      // B x N
    long long ldw;
};

constexpr int BULK_T = 64;    // tile edge
constexpr int BULK_KC = 16;   // k-chunk
constexpr int BULK_AP = 18;   // shared pitch of the A chunk rows (doubles)
constexpr size_t bulk_smem_bytes() {
    return sizeof(double) * 2 * (BULK_T * BULK_AP + BULK_KC * BULK_T) + sizeof(int) * 32 * 128;
}

template <int WINDOW>
__global__ void __launch_bounds__(128, 3) fw_bulk_kernel(BulkArgs a) {
    extern __shared__ __align__(16) unsigned char bulk_smem[];
    typedef double (*AsT)[BULK_T][BULK_AP];
    typedef double (*BsT)[BULK_KC][BULK_T];
    typedef int (*MsT)[128];
    AsT As = reinterpret_cast<AsT>(bulk_smem);                                         // [2][64][18]
    BsT Bs = reinterpret_cast<BsT>(bulk_smem + sizeof(double) * 2 * BULK_T * BULK_AP); // [2][16][64]
    // mid (k of the last replacement) per entry; written only on the rare exact path, so it
    // lives in shared memory ([entry][thread], conflict-free) and leaves the registers to the
    // DFMA results in flight.
    MsT Ms = reinterpret_cast<MsT>(bulk_smem + sizeof(double) * 2 * (BULK_T * BULK_AP + BULK_KC * BULK_T));  // [32][128]

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int tbc = a.b0 / BULK_T;       // first of the two tile columns covered by the k-block
    const int tbr = a.blk_r0 / BULK_T;   // first of the two LOCAL tile rows covered (huge if none)
    int ti = blockIdx.y, tj = blockIdx.x;
    ti = ti < tbr ? ti : ti + FW_B / BULK_T;
    tj = tj < tbc ? tj : tj + FW_B / BULK_T;
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
    unsigned chg = 0;   // bit e set <=> entry e of this thread was replaced (Ms[e][tid] is then valid)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long ro = (long long)(i0 + r * 8 + ty) * ld + j0 + tx * 2;
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
            const double2 v = *reinterpret_cast<const double2 *>(a.rate + ro + cq * 32);
            o[r][cq * 2] = v.x; o[r][cq * 2 + 1] = v.y;
        }
    }
    if (a.row0 + i0 == j0) {  // the matrix diagonal crosses this tile: hold it as NaN (see fw_common.cuh)
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
        if constexpr (WINDOW == 2) {
#pragma unroll 2
        for (int k2 = 0; k2 < BULK_KC / 2; ++k2) {
            // operands of TWO consecutive steps kk = 2*k2, 2*k2+1
            double2 a2[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) a2[r] = *reinterpret_cast<const double2 *>(&As[buf][r * 8 + ty][k2 * 2]);
            const double2 p01 = *reinterpret_cast<const double2 *>(&Bs[buf][k2 * 2][tx * 2]);
            const double2 p23 = *reinterpret_cast<const double2 *>(&Bs[buf][k2 * 2][32 + tx * 2]);
            const double2 q01 = *reinterpret_cast<const double2 *>(&Bs[buf][k2 * 2 + 1][tx * 2]);
            const double2 q23 = *reinterpret_cast<const double2 *>(&Bs[buf][k2 * 2 + 1][32 + tx * 2]);
            // Filter (1 DFMA per relaxation): with round-toward-minus-infinity,
            //   sign(fma(a, b, -o)) is clear  <=>  exact(a*b) > o  (or a positive-signed NaN),
            // and exact(a*b) <= o implies RN(a*b) <= o, i.e. the reference's strict test
            // o < a*b (Algorithms.hs:55,61) cannot fire.  An exact tie gives -0 under RM.
            // Both steps are filtered against the o of the first one: o only grows, so a stale
            // (smaller) o can only add candidates, never hide one.  accr[r] covers the 8
            // relaxations of micro-tile row r; a warp-uniform vote on their AND skips the
            // pair of steps (the common case), otherwise only rows with a candidate in some
            // lane replay the two steps exactly, in ascending k.
            int accr[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int h0 = __double2hiint(__fma_rd(a2[r].x, p01.x, -o[r][0]));
                const int h1 = __double2hiint(__fma_rd(a2[r].x, p01.y, -o[r][1]));
                const int h2 = __double2hiint(__fma_rd(a2[r].x, p23.x, -o[r][2]));
                const int h3 = __double2hiint(__fma_rd(a2[r].x, p23.y, -o[r][3]));
                const int h4 = __double2hiint(__fma_rd(a2[r].y, q01.x, -o[r][0]));
                const int h5 = __double2hiint(__fma_rd(a2[r].y, q01.y, -o[r][1]));
                const int h6 = __double2hiint(__fma_rd(a2[r].y, q23.x, -o[r][2]));
                const int h7 = __double2hiint(__fma_rd(a2[r].y, q23.y, -o[r][3]));
                accr[r] = ((h0 & h1 & h2) & (h3 & h4 & h5)) & (h6 & h7);
            }
            const int acc = ((accr[0] & accr[1] & accr[2]) & (accr[3] & accr[4] & accr[5])) & (accr[6] & accr[7]);
            if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                const int kloc = ch * BULK_KC + k2 * 2;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (__any_sync(0xffffffffu, accr[r] >= 0)) {
                        // exact path: one rounded multiply, strict compare (Algorithms.hs:55,61)
                        double n;
                        n = a2[r].x * p01.x; if (o[r][0] < n) { o[r][0] = n; Ms[r * 4 + 0][tid] = kloc; chg |= 1u << (r * 4 + 0); }
                        n = a2[r].x * p01.y; if (o[r][1] < n) { o[r][1] = n; Ms[r * 4 + 1][tid] = kloc; chg |= 1u << (r * 4 + 1); }
                        n = a2[r].x * p23.x; if (o[r][2] < n) { o[r][2] = n; Ms[r * 4 + 2][tid] = kloc; chg |= 1u << (r * 4 + 2); }
                        n = a2[r].x * p23.y; if (o[r][3] < n) { o[r][3] = n; Ms[r * 4 + 3][tid] = kloc; chg |= 1u << (r * 4 + 3); }
                        n = a2[r].y * q01.x; if (o[r][0] < n) { o[r][0] = n; Ms[r * 4 + 0][tid] = kloc + 1; chg |= 1u << (r * 4 + 0); }
                        n = a2[r].y * q01.y; if (o[r][1] < n) { o[r][1] = n; Ms[r * 4 + 1][tid] = kloc + 1; chg |= 1u << (r * 4 + 1); }
                        n = a2[r].y * q23.x; if (o[r][2] < n) { o[r][2] = n; Ms[r * 4 + 2][tid] = kloc + 1; chg |= 1u << (r * 4 + 2); }
                        n = a2[r].y * q23.y; if (o[r][3] < n) { o[r][3] = n; Ms[r * 4 + 3][tid] = kloc + 1; chg |= 1u << (r * 4 + 3); }
                    }
                }
            }
        }
        } else {
            // operands of step kk are fetched one step ahead so that their shared-memory
            // latency hides behind the previous step's DFMAs and vote
            double av[8];
            double2 b01, b23;
#pragma unroll
            for (int r = 0; r < 8; ++r) av[r] = As[buf][r * 8 + ty][0];
            b01 = *reinterpret_cast<const double2 *>(&Bs[buf][0][tx * 2]);
            b23 = *reinterpret_cast<const double2 *>(&Bs[buf][0][32 + tx * 2]);
#pragma unroll 4
            for (int kk = 0; kk < BULK_KC; ++kk) {
                double avn[8];
                double2 b01n, b23n;
                const int kn = (kk + 1 < BULK_KC) ? kk + 1 : kk;
#pragma unroll
                for (int r = 0; r < 8; ++r) avn[r] = As[buf][r * 8 + ty][kn];
                b01n = *reinterpret_cast<const double2 *>(&Bs[buf][kn][tx * 2]);
                b23n = *reinterpret_cast<const double2 *>(&Bs[buf][kn][32 + tx * 2]);
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
                    const int kloc = ch * BULK_KC + kk;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (__any_sync(0xffffffffu, accr[r] >= 0)) {
                            double n;
                            n = av[r] * b01.x; if (o[r][0] < n) { o[r][0] = n; Ms[r * 4 + 0][tid] = kloc; chg |= 1u << (r * 4 + 0); }
                            n = av[r] * b01.y; if (o[r][1] < n) { o[r][1] = n; Ms[r * 4 + 1][tid] = kloc; chg |= 1u << (r * 4 + 1); }
                            n = av[r] * b23.x; if (o[r][2] < n) { o[r][2] = n; Ms[r * 4 + 2][tid] = kloc; chg |= 1u << (r * 4 + 2); }
                            n = av[r] * b23.y; if (o[r][3] < n) { o[r][3] = n; Ms[r * 4 + 3][tid] = kloc; chg |= 1u << (r * 4 + 3); }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) av[r] = avn[r];
                b01 = b01n; b23 = b23n;
            }
        }
        __syncthreads();
    }

    // ---- epilogue: values (vector stores), next-hops / mids only where a relaxation fired ----
    if (chg == 0) return;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = i0 + r * 8 + ty;
        const long long ro = (long long)row * ld + j0 + tx * 2;
#pragma unroll
        for (int cq = 0; cq < 2; ++cq) {
            const unsigned cb = (chg >> (r * 4 + cq * 2)) & 3u;
            if (cb) {
                const int m0 = (cb & 1u) ? Ms[r * 4 + cq * 2][tid] : -1;
                const int m1 = (cb & 2u) ? Ms[r * 4 + cq * 2 + 1][tid] : -1;
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
