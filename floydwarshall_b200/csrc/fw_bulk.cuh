// Phase 3 of the exact-order blocked solve: every entry with row and column
// outside the k-block [b0, b0+B) takes the B steps of the block in ascending
// order against the two read-only snapshot panels
//     n = CpT[kk][i] * Rw[kk][j];   if (R[i][j] < n) { R[i][j] = n; mid = kk; }
// and resolves its next-hop once, at the end:  NX[i][j] = NCp[i][mid].
// This is >97 % of the time of a large solve.
//
// Relaxations fire rarely (about 7 replacements per ENTRY over a whole solve,
// i.e. ~7/N of the relaxations), so each k step first runs a filter that costs
// ONE DFMA per relaxation and proves "nothing fires here" for the whole warp;
// only when some lane has a candidate do the affected micro-tile rows replay
// the step with the exact mul / strict-compare / select sequence.
//
//   d = fma_rd(a, b, -o)            (round toward minus infinity)
//   sign(d) clear  <=>  exact(a*b) > o   (or a positive-signed NaN)
//   exact(a*b) <= o  ==>  RN(a*b) <= o,  so a set sign bit proves that the
//   reference's test  o < a*b  (Algorithms.hs:55,61) is false.  An exact tie
//   gives -0 under RM (hence round-down rather than round-to-nearest).
// The filter is conservative for every IEEE input, so results stay bit-exact.
//
// CTA = 128 threads = 8 (ty) x 16 (tx).  Output tile 64 rows x (32*CQ) columns in
// registers: thread rows ty*8 + r (r < 8), columns cq*32 + tx*2 + e, so global
// and shared accesses are contiguous 16-byte pieces across the 16 tx lanes and
// the a-operands of a thread are four LDS.128.  The panels stream through
// shared memory in k-chunks of 16 with cp.async double buffering.  CQ = 2:
// 8x4 micro-tile, 3 CTAs/SM;  CQ = 4: 8x8 micro-tile, 2 CTAs/SM.
#pragma once
#include "fw_common.cuh"

namespace fw {

struct BulkArgs {
    double *rate;
    int32_t *next;
    int32_t *mid;          // nullable
    long long ld;
    int b0;                // first pivot of the (first) k-block (global)
    // Row shards: this launch sees local rows of a shard starting at local row `row0` of it; local row l of
    // the shard is global row ((l / cbr) * P + r) * cbr + l % cbr  (cyclic blocks of cbr rows over P ranks;
    // unsharded: row0 = 0, cbr huge, P = 1, r = 0).  Only the diagonal test needs global rows.
    int row0, cbr, P, r;
    // One launch applies nb (1 .. BULK_MAXNB) CONSECUTIVE k-blocks [b0, b0 + nb*128) to every selected tile,
    // from per-block snapshot panels:
    int nb;
    const double *CpT[8];  // B x rows  column snapshots, transposed: CpT[kk*ldc + i]
    const int32_t *NCp[8]; // rows x B  next-hop snapshots: NCp[i*B + kk]
    const double *Rw[8];   // B x N     row snapshots: Rw[kk*ldw + j]
    long long ldc, ldw;
    // Tile selection, in tile units (64 local rows / TW columns): grid (x, y) -> tile
    //   tj = col_lo + x, += cskipn if tj >= cskip0;   ti = row_lo + y, += rskipn if ti >= rskip0
    int row_lo, rskip0, rskipn;
    int col_lo, cskip0, cskipn;   // in units of 64 columns (scaled by 64/TW inside the kernel)
    // nb > 1: a tile in the row or column strip of the group's block i (i < nb-1; strips of consecutive
    // blocks are adjacent, 2 tile rows / 128 columns each) already took blocks 0..i (phase 2 of block i and
    // the strip launches before it) and starts at block i+1.  half_r0 / half_c0 = first tile row / first
    // 64-column unit of the group's FIRST block; huge if none.
    int half_r0, half_c0;
    // 1-D grid of gx*gy CTAs, rasterised in column bands of `band` tile columns: a band's slice of the
    // row-snapshot panels (band*TW columns x 256 steps x 8 B, a few MB) stays hot in L2 while the CTAs
    // sweep all tile rows, instead of the whole 64 MB panel being re-streamed for every tile row.
    int gx, gy, band;
};

constexpr int BULK_TR = 64;   // tile rows
constexpr int BULK_MAXNB = 8; // most k-blocks per fused launch
constexpr int BULK_KC = 16;   // k-chunk
constexpr int BULK_ST = 3;    // cp.async pipeline stages (one __syncthreads per chunk)
#ifndef FW_BULK_UNROLL
#define FW_BULK_UNROLL 2      // k steps per loop body of the bulk kernel (experiment knob)
#endif
constexpr int BULK_UNROLL = FW_BULK_UNROLL;
#ifndef FW_BULK_MINCTAS
#define FW_BULK_MINCTAS 3      // resident CTAs/SM the 8x4 variant is compiled for (register cap 168)
#endif
#ifndef FW_BULK_PREFETCH
#define FW_BULK_PREFETCH 1     // fetch the operands of step k+1 during step k
#endif
#ifndef FW_BULK_TMA
#define FW_BULK_TMA 0          // 1: panels staged by cp.async.bulk (TMA, 1-D) + mbarriers instead of cp.async (LDGSTS)
#endif
#ifndef FW_BULK_ONELEVEL
#define FW_BULK_ONELEVEL 0     // 1: one flat AND tree over the 32 sign words of a step; the per-row words the replay selects rows by are recomputed on the (rare) exact path
#endif
#ifndef FW_BULK_LATEVOTE
#define FW_BULK_LATEVOTE 0     // 1: vote on step k's sign words after step k+1's DFMAs (loop_probe V7: +5 % in isolation, -3.5 % in the solve)
#endif
template <int CQ>
constexpr size_t bulk_smem_bytes() {
    return sizeof(double) * BULK_ST * BULK_KC * (BULK_TR + 32 * CQ) + sizeof(int) * (16 * CQ) * 128 +
           (FW_BULK_TMA ? 2 * BULK_ST * sizeof(unsigned long long) : 0);
}

#ifdef FW_BULK_STATS
// experiment build only (make VARIANT=_stats EXTRA=-DFW_BULK_STATS): [0] warp-steps, [1] warp-steps that left
// the fast path, [2] micro-tile rows replayed, [3] entries replaced.  Read with fw_debug_bulk_stats().
__device__ unsigned long long fw_bulk_stats[4];
#endif

#if FW_BULK_TMA
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#endif

template <int N>
__device__ __forceinline__ int and_tree(const int *h) {
    if constexpr (N == 1) return h[0];
    else if constexpr (N == 2) return h[0] & h[1];
    else if constexpr (N == 3) return h[0] & h[1] & h[2];
    else return and_tree<N / 2>(h) & and_tree<N - N / 2>(h + N / 2);
}

template <int CQ>
__global__ void __launch_bounds__(128, (CQ == 2 ? FW_BULK_MINCTAS : 2)) fw_bulk_kernel(BulkArgs a) {
    constexpr int TW = 32 * CQ;      // tile width (columns)
    constexpr int NC = 2 * CQ;       // columns per thread
    extern __shared__ __align__(16) unsigned char bulk_smem[];
    typedef double (*AsT)[BULK_KC][BULK_TR];
    typedef double (*BsT)[BULK_KC][TW];
    typedef int (*MsT)[128];
    AsT As = reinterpret_cast<AsT>(bulk_smem);                                                    // [ST][16][64]
    BsT Bs = reinterpret_cast<BsT>(bulk_smem + sizeof(double) * BULK_ST * BULK_KC * BULK_TR);     // [ST][16][TW]
    // mid (k of the last replacement) per entry: written only on the rare exact path, so it lives in
    // shared memory ([entry][thread], conflict-free) and leaves the registers to the DFMA results.
    MsT Ms = reinterpret_cast<MsT>(bulk_smem + sizeof(double) * BULK_ST * BULK_KC * (BULK_TR + TW)); // [8*NC][128]

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    constexpr int CU = TW / 64;             // 64-column units per tile column
    int bx, by;
    {
        const int lin = (int)blockIdx.x;
        const int per_band = a.band * a.gy;
        const int bi = lin / per_band, rem = lin - bi * per_band;
        const int w = min(a.band, a.gx - bi * a.band);       // width of this (possibly last, narrower) band
        by = rem / w;
        bx = bi * a.band + (rem - by * w);
    }
    int ti = a.row_lo + by, tj = a.col_lo / CU + bx;
    if (ti >= a.rskip0) ti += a.rskipn;
    if (tj >= a.cskip0 / CU) tj += a.cskipn / CU;
    const int i0 = ti * BULK_TR, j0 = tj * TW;
    int kstart = 0;
    if (a.nb > 1) {
        const int dr = ti - a.half_r0, dc = tj - a.half_c0 / CU;     // distance from the first block's strips
        const int br = (dr >= 0 && dr < (a.nb - 1) * (FW_B / BULK_TR)) ? dr / (FW_B / BULK_TR) + 1 : 0;
        const int bc = (dc >= 0 && dc < (a.nb - 1) * (FW_B / TW)) ? dc / (FW_B / TW) + 1 : 0;
        kstart = max(br, bc) * FW_B;
    }
    const long long ld = a.ld;

#if FW_BULK_TMA
    // TMA staging: lane l of warp 0 owns one 1-D bulk copy per chunk -- l < 16: row l of the A chunk
    // (64 doubles of CpT), l >= 16: row l-16 of the B chunk (TW doubles of Rw).  full[s] completes when the
    // 16 + 16 rows of stage s have landed (expect_tx), empty[s] when all four warps are done reading it.
    constexpr unsigned bufA = BULK_KC * BULK_TR * 8, bufB = BULK_KC * TW * 8;   // bytes per stage
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        bulk_smem + sizeof(double) * BULK_ST * BULK_KC * (BULK_TR + TW) + sizeof(int) * (8 * NC) * 128);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);          // full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (ST + s)
    const int lane = tid & 31;
    const bool producer = tid < 32;
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < BULK_ST; ++st) { mbar_init(bar0 + 8 * st, 1); mbar_init(bar0 + 8 * (BULK_ST + st), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const bool isA = lane < 16;
    const int prow = lane & 15;
    const long long poff = isA ? ((long long)prow * a.ldc + i0) : ((long long)prow * a.ldw + j0);
    const unsigned pdst = isA ? (unsigned)__cvta_generic_to_shared(&As[0][prow][0]) : (unsigned)__cvta_generic_to_shared(&Bs[0][prow][0]);
    auto load_chunk = [&](int c, int stage) {       // warp 0 only
        const int set = c >> 3, kk0 = (c & 7) * BULK_KC;   // TMA variant
        const unsigned full = bar0 + 8 * stage;
        if (lane == 0) mbar_expect_tx(full, bufA + bufB);
        __syncwarp();
        const double *src = isA ? a.CpT[set] + ((long long)kk0 * a.ldc + poff) : a.Rw[set] + ((long long)kk0 * a.ldw + poff);
        bulk_g2s(pdst + stage * (isA ? bufA : bufB), src, isA ? BULK_TR * 8 : TW * 8, full);
    };
    const int ch0 = kstart / BULK_KC, ch1 = a.nb * (FW_B / BULK_KC);
    if (producer) {
        load_chunk(ch0, 0);
        load_chunk(ch0 + 1, 1);
    }
#else
    // cp.async staging.  A chunk is 16 panel rows of 512 B (A: CpT, 64 doubles) and 16 of TW*8 B (B: Rw).
    // Thread tid owns row kk = tid >> 3 of both and, inside the row, the 16-byte pieces at byte offsets
    // (tid & 7)*16 + t*128: all pieces of a thread hang off ONE pointer with constant offsets, and the eight
    // lanes of a row cover 128 contiguous bytes per instruction.  The two pointers advance by 16 panel rows
    // per chunk (two 64-bit adds) instead of being rebuilt from the kernel arguments (which cost ~60
    // integer instructions per chunk, 4 % of the issue slots of a 16-step chunk).
    constexpr unsigned bufA = BULK_KC * BULK_TR * 8, bufB = BULK_KC * TW * 8;   // bytes per stage
    const int prow = tid >> 3, pseg = (tid & 7) * 2;                              // row, first double of piece 0
    const unsigned dstA = (unsigned)__cvta_generic_to_shared(&As[0][prow][pseg]);
    const unsigned dstB = (unsigned)__cvta_generic_to_shared(&Bs[0][prow][pseg]);
    const long long stepA = (long long)BULK_KC * a.ldc, stepB = (long long)BULK_KC * a.ldw;   // doubles per chunk
    const int ch0 = kstart / BULK_KC, ch1 = a.nb * (FW_B / BULK_KC);
    // chunk c (0 .. nb*8-1) = steps [16c, 16c+16) relative to b0; chunks 8.. come from the second block's panels
    const double *pa = a.CpT[ch0 >> 3] + ((long long)((ch0 & 7) * BULK_KC + prow) * a.ldc + i0 + pseg);
    const double *pb = a.Rw[ch0 >> 3] + ((long long)((ch0 & 7) * BULK_KC + prow) * a.ldw + j0 + pseg);
    int cnext = ch0;                          // the chunk pa / pb point at
    auto load_chunk = [&](int buf) {          // loads chunk `cnext` into stage `buf`, then advances
#pragma unroll
        for (int t = 0; t < 4; ++t)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dstA + buf * bufA + t * 128), "l"(pa + t * 16));
#pragma unroll
        for (int t = 0; t < TW / 16; ++t)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dstB + buf * bufB + t * 128), "l"(pb + t * 16));
        ++cnext;
        if ((cnext & (FW_B / BULK_KC - 1)) == 0) {        // first chunk of the next k-block: switch panel sets
            const int set = min(cnext / (FW_B / BULK_KC), BULK_MAXNB - 1);   // (one past the last: never loaded)
            pa = a.CpT[set] + ((long long)prow * a.ldc + i0 + pseg);
            pb = a.Rw[set] + ((long long)prow * a.ldw + j0 + pseg);
        } else {
            pa += stepA;
            pb += stepB;
        }
    };

    load_chunk(0);
    cp_async_commit();
    load_chunk(1);
    cp_async_commit();

#endif

    double o[8][NC];
    unsigned long long chg = 0;   // bit e set <=> entry e of this thread was replaced (Ms[e][tid] valid)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long ro = (long long)(i0 + ty * 8 + r) * ld + j0 + tx * 2;
#pragma unroll
        for (int cq = 0; cq < CQ; ++cq) {
            // streamed once per launch: evict-first, so that the snapshot panels keep their L2 lines
            o[r][cq * 2] = __ldcs(a.rate + ro + cq * 32);
            o[r][cq * 2 + 1] = __ldcs(a.rate + ro + cq * 32 + 1);
        }
    }
    {   // the matrix diagonal crosses this tile: hold it as +inf -- o < n is false for every n, it is never
        // written back (chg stays clear), and fma(a, b, -inf) = -inf keeps the filter on its fast path
        const int li0 = a.row0 + i0;                                          // a 64-row tile never straddles a cyclic block
        const int gi0 = ((li0 / a.cbr) * a.P + a.r) * a.cbr + li0 % a.cbr;
        if (gi0 + BULK_TR > j0 && gi0 < j0 + TW) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (gi0 + ty * 8 + r == j0 + (c >> 1) * 32 + tx * 2 + (c & 1)) o[r][c] = pinf();
        }
    }

    static_assert(BULK_ST == 3, "buffer rotation below assumes 3 stages");
    int buf = 0;
    for (int ch = ch0; ch < ch1; ++ch) {
#if FW_BULK_TMA
        const int it = ch - ch0;                         // chunk it uses stage it % 3 in its (it / 3)-th round
        if (producer && ch + 2 < ch1) {
            const int rs = buf == 0 ? 2 : buf - 1;       // (buf + 2) % 3 == the stage chunk it-1 used
            if (it >= 1) mbar_wait(bar0 + 8 * (BULK_ST + rs), ((it - 1) / BULK_ST) & 1);   // all warps done with it
            load_chunk(ch + 2, rs);
        }
        mbar_wait(bar0 + 8 * buf, (it / BULK_ST) & 1);   // chunk ch has landed
#else
        // chunk ch has landed once at most one younger group is still in flight
        if (ch + 1 < ch1) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();   // (a) chunk ch visible to all; (b) everyone is done with chunk ch-1's buffer
        if (ch + 2 < ch1) {
            load_chunk(buf == 0 ? 2 : buf - 1);   // chunk ch+2 into (buf + 2) % 3 == the buffer chunk ch-1 used
            cp_async_commit();
        }
#endif
        // operands of step kk are fetched one step ahead so that their shared-memory latency
        // hides behind the previous step's DFMAs and vote
        double av[8], bv[NC];
        auto fetch = [&](int kk, double (&ax)[8], double (&bx)[NC]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 v = *reinterpret_cast<const double2 *>(&As[buf][kk][ty * 8 + q * 2]);
                ax[q * 2] = v.x; ax[q * 2 + 1] = v.y;
            }
#pragma unroll
            for (int cq = 0; cq < CQ; ++cq) {
                const double2 v = *reinterpret_cast<const double2 *>(&Bs[buf][kk][cq * 32 + tx * 2]);
                bx[cq * 2] = v.x; bx[cq * 2 + 1] = v.y;
            }
        };
        // exact path for step `ks` of this chunk (operands re-read from shared memory: it is rare):
        // one rounded multiply, strict compare (Algorithms.hs:55,61), rows selected by their sign words
#if FW_BULK_ONELEVEL
        auto vote_and_replay_flat = [&](const int acc, int ks) {
            if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                const int kloc = ch * BULK_KC + ks;
                double ax[8], bx[NC];
                fetch(ks, ax, bx);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    int h[NC];
#pragma unroll
                    for (int c = 0; c < NC; ++c) h[c] = __double2hiint(__fma_rd(ax[r], bx[c], -o[r][c]));
                    if (__any_sync(0xffffffffu, and_tree<NC>(h) >= 0)) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const double n = ax[r] * bx[c];
                            if (o[r][c] < n) {
                                o[r][c] = n;
                                Ms[r * NC + c][tid] = kloc;
                                chg |= 1ull << (r * NC + c);
                            }
                        }
                    }
                }
            }
        };
#endif
        auto vote_and_replay = [&](const int (&rw)[8], int ks) {
            const int acc = and_tree<8>(rw);
#ifdef FW_BULK_STATS
            if ((tid & 31) == 0) atomicAdd(&fw_bulk_stats[0], 1ull);
#endif
            if (__builtin_expect(__any_sync(0xffffffffu, acc >= 0), 0)) {
                const int kloc = ch * BULK_KC + ks;   // step index relative to b0 (0..255)
#ifdef FW_BULK_STATS
                if ((tid & 31) == 0) atomicAdd(&fw_bulk_stats[1], 1ull);
#endif
                double ax[8], bx[NC];
                fetch(ks, ax, bx);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (__any_sync(0xffffffffu, rw[r] >= 0)) {
#ifdef FW_BULK_STATS
                        if ((tid & 31) == 0) atomicAdd(&fw_bulk_stats[2], 1ull);
#endif
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const double n = ax[r] * bx[c];
                            if (o[r][c] < n) {
#ifdef FW_BULK_STATS
                                atomicAdd(&fw_bulk_stats[3], 1ull);
#endif
                                o[r][c] = n;
                                Ms[r * NC + c][tid] = kloc;
                                chg |= 1ull << (r * NC + c);
                            }
                        }
                    }
                }
            }
        };
#if FW_BULK_LATEVOTE
        int hprev[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) hprev[r] = -1;      // "no candidate": the vote of step -1 is a no-op
#endif
#if FW_BULK_PREFETCH
        fetch(0, av, bv);
#endif
#pragma unroll BULK_UNROLL
        for (int kk = 0; kk < BULK_KC; ++kk) {
#if FW_BULK_PREFETCH
            double avn[8], bvn[NC];
            fetch((kk + 1 < BULK_KC) ? kk + 1 : kk, avn, bvn);
#else
            fetch(kk, av, bv);
#endif
            // all DFMAs of the step first, then the integer reduction of their sign words
            int hi[8][NC];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c) hi[r][c] = __double2hiint(__fma_rd(av[r], bv[c], -o[r][c]));
#if FW_BULK_ONELEVEL
            vote_and_replay_flat(and_tree<8 * NC>(&hi[0][0]), kk);
#else
            // two levels: accr[r] covers micro-tile row r, acc the whole step
            int accr[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) accr[r] = and_tree<NC>(hi[r]);
#if FW_BULK_LATEVOTE
            // the second level and the vote run one step late, behind the next step's DFMAs: the row words of
            // step kk-1 (hprev) are voted on here.  A replay of step kk-1 after the filter of step kk is
            // still exact-order: the filter only PROVES "nothing fires", and it stays conservative when it
            // saw entries that a later replay then raises (a smaller o can only add candidates).
            vote_and_replay(hprev, kk - 1);
#pragma unroll
            for (int r = 0; r < 8; ++r) hprev[r] = accr[r];
#else
            vote_and_replay(accr, kk);
#endif
#endif
#if FW_BULK_PREFETCH
#pragma unroll
            for (int r = 0; r < 8; ++r) av[r] = avn[r];
#pragma unroll
            for (int c = 0; c < NC; ++c) bv[c] = bvn[c];
#endif
        }
#if FW_BULK_LATEVOTE
        vote_and_replay(hprev, BULK_KC - 1);   // the chunk's last step, before its buffer can be recycled
#endif
#if FW_BULK_TMA
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (BULK_ST + buf));   // this warp is done with the stage
#endif
        buf = (buf == BULK_ST - 1) ? 0 : buf + 1;
    }

    // ---- epilogue: only entries that were replaced are written back ----
    if (chg == 0) return;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = i0 + ty * 8 + r;
        const long long ro = (long long)row * ld + j0 + tx * 2;
#pragma unroll
        for (int cq = 0; cq < CQ; ++cq) {
            const unsigned cb = (unsigned)(chg >> (r * NC + cq * 2)) & 3u;
            if (cb) {
                const long long eo = ro + cq * 32;
                if (cb == 3u) {
                    *reinterpret_cast<double2 *>(a.rate + eo) = make_double2(o[r][cq * 2], o[r][cq * 2 + 1]);
                } else if (cb == 1u) {
                    a.rate[eo] = o[r][cq * 2];
                } else {
                    a.rate[eo + 1] = o[r][cq * 2 + 1];
                }
                if (cb & 1u) {
                    const int m0 = Ms[r * NC + cq * 2][tid];
                    a.next[eo] = a.NCp[m0 >> 7][(long long)row * FW_B + (m0 & (FW_B - 1))];
                    if (a.mid) a.mid[eo] = a.b0 + m0;
                }
                if (cb & 2u) {
                    const int m1 = Ms[r * NC + cq * 2 + 1][tid];
                    a.next[eo + 1] = a.NCp[m1 >> 7][(long long)row * FW_B + (m1 & (FW_B - 1))];
                    if (a.mid) a.mid[eo + 1] = a.b0 + m1;
                }
            }
        }
    }
}

}  // namespace fw
