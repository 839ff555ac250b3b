// fw_tile_kernel -- one CTA runs the whole k loop on one 128x128 tile.
//
// Two uses:
//  * phase 1 of the blocked solve: the pivot diagonal tile of k-block b0, which
//    additionally emits the step-k snapshots the other phases consume
//      Cp [kk*ldc + i]    = R[i][b0+kk]  as of step b0+kk   (column snapshot, transposed)
//      NCp[(b0+i)*B + kk] = NX[i][b0+kk] as of step b0+kk
//      Rw [kk*ldw + b0+j] = R[b0+kk][j]  as of step b0+kk   (row snapshot)
//    (the reference reads row k / column k AS OF STEP k, so the
//    textbook "finished diagonal tile" is not result-equivalent.)
//  * batched mode: one CTA per independent graph with n <= 128 (the FSM replay,
//    reference src/lib/ProcessRequests.hs:82-84): blockIdx.x selects the graph.
//
// Layout: 512 threads = 32 (ty) x 16 (tx); thread owns rows ty*4..+3 and
// columns tx*8..+7 of the tile in REGISTERS (64 regs of fp64 state).  Per step
// k the owners of column k / row k publish them through double-buffered shared
// vectors; everybody else reads 4 + 8 operands and runs a 4x8 micro-tile of
// mul / compare / select.  next-hops (and mids) live in shared memory and are
// written with predicated stores only when a relaxation fires (1-2 % of them).
//
// Step-to-step synchronisation (FW_TILE_PIPE): a full barrier per step left the SM idle for a third of the
// time (ncu: barrier stalls 0.48 per issue at 49 % issue activity, profiles/r01d_pivot_phases_ncu_summary.json).
// Every thread now relaxes FIRST the entries of its micro-tile that lie in row k+1 / column k+1 (the entries
// of a step are independent, so their order is free), the owners publish them, and the warp ARRIVES on the
// mbarrier that guards step k+1 before it relaxes the other 21 entries.  A warp that starts step k+1 waits
// only for those arrivals, i.e. for the next operands -- not for the slowest warp's whole step.  Warps are
// never more than one step apart, which is what the double-buffered row / column vectors allow; column k+1
// and row k+1 of next / mid are final before the arrival as well, so the snapshot reads stay exact.
#pragma once
#include "fw_common.cuh"

namespace fw {

struct TileArgs {
    double *rate;           // matrix (or graph 0) base
    int32_t *next;
    int32_t *mid;           // PATHS only
    int32_t *csT;           // PATHS only
    int32_t *rs;            // PATHS only
    long long ld;           // leading dimension, elements
    long long batch_stride; // elements between graphs (batched mode), else 0
    int b0;                 // tile origin: global column (== global row of the pivots)
    int r0;                 // tile origin: LOCAL row inside this shard (== b0 when unsharded)
    int nv;                 // valid rows/cols inside the tile (1..128)
    double *Cp;             // snapshot outputs; null in batched mode.  Cp is TRANSPOSED: Cp[kk*ldc + row]
    long long ldc;
    int32_t *NCp;
    double *Rw;
    long long ldw;
};

constexpr int TILE_NXP = 132;  // padded pitch of the shared next/mid tiles (ints)
#ifndef FW_TILE_PIPE
#define FW_TILE_PIPE 1   // 1: warps run up to one step ahead of each other (mbarrier arrive / wait); 0: one __syncthreads per step
#endif
constexpr size_t tile_smem_bytes(bool paths) {
    return (size_t)(paths ? 2 : 1) * 128 * TILE_NXP * 4 + 4 * 128 * 8 + 16;
}

template <bool PATHS>
__global__ void __launch_bounds__(512, 1) fw_tile_kernel(TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *NXs = reinterpret_cast<int32_t *>(smem_raw);
    int32_t *MIDs = NXs + 128 * TILE_NXP;  // only touched when PATHS
    double *colbuf = reinterpret_cast<double *>(smem_raw + (size_t)(PATHS ? 2 : 1) * 128 * TILE_NXP * 4);
    double *rowbuf = colbuf + 256;

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int nv = a.nv;
    const long long ld = a.ld;
    const long long goff = (long long)blockIdx.x * a.batch_stride + (long long)a.r0 * ld + a.b0;
    double *R = a.rate + goff;
    int32_t *NX = a.next + goff;

    // ---- load: values to registers, next/mid to shared ----
    double o[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = ty * 4 + r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = tx * 8 + c;
            const bool valid = (i < nv) && (j < nv) && (i != j);
            o[r][c] = valid ? R[(long long)i * ld + j] : qnan();
        }
    }
    for (int idx = tid; idx < 128 * 128; idx += 512) {
        const int i = idx >> 7, j = idx & 127;
        const bool valid = (i < nv) && (j < nv);
        NXs[i * TILE_NXP + j] = valid ? NX[(long long)i * ld + j] : -1;
        if (PATHS) MIDs[i * TILE_NXP + j] = valid ? a.mid[goff + (long long)i * ld + j] : -1;
    }
    // publish column 0 / row 0
    if (tx == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) colbuf[ty * 4 + r] = o[r][0];
    }
    if (ty == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) rowbuf[swz128(tx * 8 + c)] = o[0][c];
    }

    int32_t *nxp = NXs + (ty * 4) * TILE_NXP + tx * 8;
    int32_t *mdp = MIDs + (ty * 4) * TILE_NXP + tx * 8;
#if FW_TILE_PIPE
    // bars[b] guards the operands of the steps k = b (mod 2), k >= 1; 16 warps arrive once per use
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(rowbuf + 256);
    if (tid == 0) {
        mbar_init(bar0, 16);
        mbar_init(bar0 + 8, 16);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();     // tile loaded, column 0 / row 0 published, barriers ready
#endif

    for (int kt = 0; kt < 16; ++kt) {
        if (kt * 8 >= nv) break;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int k = kt * 8 + c;
            const int buf = c & 1;
#if FW_TILE_PIPE
            if (k >= 1 && k < nv) mbar_wait(bar0 + 8 * buf, ((k - 1) >> 1) & 1);   // row k / column k are published
#else
            __syncthreads();
#endif
            if (k < nv) {  // uniform
                const double *cb = colbuf + buf * 128;
                const double *rb = rowbuf + buf * 128;
                // snapshots for the other phases (diag mode only)
                if (a.Cp != nullptr) {
                    // The pivot itself (held as NaN in here, see fw_common.cuh) is EXPORTED as 0.0: the panel
                    // kernels put the DFMA filter in front of their exact path, which a NaN factor would
                    // defeat, and a zero factor skips i == k / j == k just as well for entries >= 0
                    // (s * 0 is 0 or NaN, never larger) -- the domain fw_validate_kernel enforces.
                    if (tid < 128) {
                        a.Cp[(long long)k * a.ldc + a.r0 + tid] = (tid == k) ? 0.0 : cb[tid];
                        a.NCp[(long long)(a.r0 + tid) * FW_B + k] = NXs[tid * TILE_NXP + k];
                    } else if (tid < 256) {
                        const int j = tid - 128;
                        a.Rw[(long long)k * a.ldw + a.b0 + j] = (j == k) ? 0.0 : rb[swz128(j)];
                    }
                }
                if (PATHS) {
                    if (tid >= 256 && tid < 384) {
                        const int i = tid - 256;
                        if (i < nv) a.csT[goff + (long long)i * ld + k] = MIDs[i * TILE_NXP + k];
                    } else if (tid >= 384) {
                        const int j = tid - 384;
                        if (j < nv) a.rs[goff + (long long)k * ld + j] = MIDs[k * TILE_NXP + j];
                    }
                }
                // operands
                double av[4], bv[8];
                int an[4];
                {
                    const double2 a01 = *reinterpret_cast<const double2 *>(cb + ty * 4);
                    const double2 a23 = *reinterpret_cast<const double2 *>(cb + ty * 4 + 2);
                    av[0] = a01.x; av[1] = a01.y; av[2] = a23.x; av[3] = a23.y;
#pragma unroll
                    for (int r = 0; r < 4; ++r) an[r] = NXs[(ty * 4 + r) * TILE_NXP + k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 b2 = *reinterpret_cast<const double2 *>(rb + q * 32 + tx * 2);
                        bv[q * 2] = b2.x; bv[q * 2 + 1] = b2.y;
                    }
                }
                const int kabs = a.b0 + k;
                auto relax1 = [&](int r, int cc) {
                    const double n = av[r] * bv[cc];          // one rounded multiply (Algorithms.hs:61)
                    if (o[r][cc] < n) {                       // strict (Algorithms.hs:55)
                        o[r][cc] = n;
                        nxp[r * TILE_NXP + cc] = an[r];
                        if (PATHS) mdp[r * TILE_NXP + cc] = kabs;
                    }
                };
                const int cn = (c + 1) & 7;       // micro-tile column that holds matrix column k+1 (for tx == ktn)
                const int rn = (c + 1) & 3;       // micro-tile row that holds matrix row k+1 (for ty == tyn)
#if FW_TILE_PIPE
                // the entries next step's operands come from go first ...
#pragma unroll
                for (int r = 0; r < 4; ++r) relax1(r, cn);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                    if (cc != cn) relax1(rn, cc);
#else
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) relax1(r, cc);
#endif
                // publish column k+1 / row k+1 (values after step k) into the other buffer
                const int ktn = kt + ((c == 7) ? 1 : 0);
                if (tx == ktn) {
                    double *cbn = colbuf + (buf ^ 1) * 128 + ty * 4;
                    *reinterpret_cast<double2 *>(cbn) = make_double2(o[0][cn], o[1][cn]);
                    *reinterpret_cast<double2 *>(cbn + 2) = make_double2(o[2][cn], o[3][cn]);
                }
                const int tyn = kt * 2 + ((c + 1) >> 2);
                if (ty == tyn) {
                    double *rbn = rowbuf + (buf ^ 1) * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<double2 *>(rbn + q * 32 + tx * 2) =
                            make_double2(o[rn][q * 2], o[rn][q * 2 + 1]);
                }
#if FW_TILE_PIPE
                // ... then the warp arrives for step k+1 and relaxes the rest
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar0 + 8 * (buf ^ 1));
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc)
                        if (r != rn && cc != cn) relax1(r, cc);
#endif
            }
        }
    }
    __syncthreads();

    // ---- store ----
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = ty * 4 + r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = tx * 8 + c;
            if ((i < nv) && (j < nv) && (i != j)) R[(long long)i * ld + j] = o[r][c];
        }
    }
    for (int idx = tid; idx < 128 * 128; idx += 512) {
        const int i = idx >> 7, j = idx & 127;
        if ((i < nv) && (j < nv)) {
            NX[(long long)i * ld + j] = NXs[i * TILE_NXP + j];
            if (PATHS) a.mid[goff + (long long)i * ld + j] = MIDs[i * TILE_NXP + j];
        }
    }
}

}  // namespace fw
