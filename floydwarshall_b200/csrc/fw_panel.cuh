// Phase 2 of the exact-order blocked solve: the pivot COLUMN panel (rows outside
// the k-block, columns inside) and the pivot ROW panel (rows inside, columns
// outside).  Both are "B sequential steps on a B-vector":
//
//   column panel, one job per matrix row i:   y[j] = R[i][b0+j]
//       step kk:  s = y[kk]  (snapshot -> Cp[i][kk], NCp[i][kk])
//                 y[j] <- s * Rd[kk][j]  if larger      (Rd = diagonal tile row
//                 nx[j] <- nx[kk]                         snapshots = Rw[:, b0..])
//   row panel, one job per matrix column j:   x[i] = R[b0+i][j]
//       step kk:  s = x[kk]  (snapshot -> Rw[kk][j])
//                 x[i] <- Cd[i][kk] * s  if larger      (Cd = diagonal tile column
//                 nx[i] <- NCd[i][kk]                     snapshots = Cp[b0.., :])
//
// A job is spread over 16 lanes x 8 elements; the pivot element of step kk is
// broadcast with warp shuffles, the factor row comes from a swizzled shared
// copy of the B x B snapshot matrix (conflict-free LDS.128).  The diagonal of
// that matrix is 0.0 (fw_tile.cuh publishes "never a factor" entries as zero), which
// is what makes element kk skip itself at step kk: s * 0 never beats an entry >= 0.
// Every step runs the bulk kernel's one-DFMA filter first and the exact
// mul / compare / select only when some lane of the warp has a candidate.
// 512 threads = 32 jobs per CTA pass; persistent CTAs stride over the jobs.
#pragma once
#include "fw_common.cuh"

namespace fw {

struct PanelArgs {
    double *rate;
    int32_t *next;
    int32_t *mid;   // PATHS only
    int32_t *csT;   // PATHS only
    int32_t *rs;    // PATHS only
    long long ld;
    int npad;       // padded matrix order = number of columns (multiple of FW_B)
    int b0;         // first pivot of the k-block (global index)
    int rows;       // rows held by this shard (== npad when unsharded)
    int blk_r0;     // LOCAL row of pivot b0 if this shard holds the k-block rows (row panel), else INT_MAX
    int skip_r0;    // column panel: local rows [skip_r0, skip_r0 + skipn) are left out (INT_MAX: none)
    int skipn;
    double *Cp;     // B x rows column snapshots, TRANSPOSED: Cp[kk*ldc + i]
    long long ldc;
    int32_t *NCp;
    double *Rw;     // B x N row snapshots
    long long ldw;
};

constexpr int PANEL_FP = 130;  // shared pitch (doubles): 16-byte aligned rows, 4-way max conflict on transposed fill
constexpr size_t panel_smem_bytes() { return (size_t)128 * PANEL_FP * 8; }
// Jobs per half-warp.  A step is one dependent chain (shuffle -> 8 DFMA -> sign words -> vote), about 150
// cycles long, and a 512-thread CTA has only 4 warps per SM sub-partition to hide it: with one job per
// half-warp the kernels ran at a quarter of their issue rate.  NJ independent jobs per half-warp share the
// factor loads and interleave their chains.  Measured (N=16384, per launch): NJ=2 is 20 % faster once there
// are enough jobs to fill every SM twice over, and slower below that (N=4096: 53 vs 35 us) -- panel_nj().
__host__ __device__ constexpr int panel_nj(int jobs, int sm_count) { return jobs >= 64 * sm_count ? 2 : 1; }

__device__ __forceinline__ int and8(const int (&h)[8]) {
    return ((h[0] & h[1] & h[2]) & (h[3] & h[4] & h[5])) & (h[6] & h[7]);
}

template <bool PATHS, int NJ>
__device__ __forceinline__ void colpanel_body(const PanelArgs &a, const int bid, const int nctas) {
    constexpr int PANEL_JOBS = 32 * NJ;   // jobs per CTA pass
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Fs = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int b0 = a.b0;
    // factor matrix F[kk][j] = Rd[kk][j] = Rw[kk][b0+j]
    for (int idx = tid; idx < 128 * 128; idx += 512) {
        const int kk = idx >> 7, col = idx & 127;
        Fs[kk * PANEL_FP + swz128(col)] = a.Rw[(long long)kk * a.ldw + b0 + col];
    }
    __syncthreads();

    const int job = tid >> 4, l = tid & 15;
    const int nrows = a.rows - (a.skip_r0 < a.rows ? a.skipn : 0);
    for (int g = bid; g * PANEL_JOBS < nrows; g += nctas) {
        long long off[NJ];
        int irow[NJ];
        double y[NJ][8];
        int nx[NJ][8], md[NJ][8];
#pragma unroll
        for (int u = 0; u < NJ; ++u) {
            const int rp = g * PANEL_JOBS + job * NJ + u;
            irow[u] = rp < a.skip_r0 ? rp : rp + a.skipn;   // local row, skipping the k-block rows
            off[u] = (long long)irow[u] * a.ld + b0 + l * 8;
            const double2 *p = reinterpret_cast<const double2 *>(a.rate + off[u]);
#pragma unroll
            for (int q = 0; q < 4; ++q) { double2 v = p[q]; y[u][q * 2] = v.x; y[u][q * 2 + 1] = v.y; }
            const int4 *pn = reinterpret_cast<const int4 *>(a.next + off[u]);
            int4 n0 = pn[0], n1 = pn[1];
            nx[u][0] = n0.x; nx[u][1] = n0.y; nx[u][2] = n0.z; nx[u][3] = n0.w;
            nx[u][4] = n1.x; nx[u][5] = n1.y; nx[u][6] = n1.z; nx[u][7] = n1.w;
            if (PATHS) {
                const int4 *pm = reinterpret_cast<const int4 *>(a.mid + off[u]);
                int4 m0 = pm[0], m1 = pm[1];
                md[u][0] = m0.x; md[u][1] = m0.y; md[u][2] = m0.z; md[u][3] = m0.w;
                md[u][4] = m1.x; md[u][5] = m1.y; md[u][6] = m1.z; md[u][7] = m1.w;
            }
        }

        for (int lt = 0; lt < 16; ++lt) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int kk = lt * 8 + c;
                double s[NJ];
#pragma unroll
                for (int u = 0; u < NJ; ++u) s[u] = __shfl_sync(0xffffffffu, y[u][c], lt, 16);
                if (l == lt) {
                    // the step-kk snapshots of column b0+kk, written as they are taken
#pragma unroll
                    for (int u = 0; u < NJ; ++u) {
                        a.Cp[(long long)kk * a.ldc + irow[u]] = s[u];
                        a.NCp[(long long)irow[u] * FW_B + kk] = nx[u][c];
                        if (PATHS) a.csT[off[u] + c] = md[u][c];
                    }
                }
                const double2 *fr = reinterpret_cast<const double2 *>(Fs + kk * PANEL_FP) + l;
                double f[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) { double2 v = fr[q * 16]; f[q * 2] = v.x; f[q * 2 + 1] = v.y; }
                // filter (fw_bulk.cuh): a set sign bit of fma_rd(s, f, -y) proves y < RN(s*f) is false
                int hu[NJ];
#pragma unroll
                for (int u = 0; u < NJ; ++u) {
                    int h[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[e] = __double2hiint(__fma_rd(s[u], f[e], -y[u][e]));
                    hu[u] = and8(h);
                }
                int acc = hu[0];
#pragma unroll
                for (int u = 1; u < NJ; ++u) acc &= hu[u];
                if (__any_sync(0xffffffffu, acc >= 0)) {
#pragma unroll
                    for (int u = 0; u < NJ; ++u) {
                        if (__any_sync(0xffffffffu, hu[u] >= 0)) {
                            // exact path: one rounded multiply, strict compare (Algorithms.hs:55,61)
                            const int snx = __shfl_sync(0xffffffffu, nx[u][c], lt, 16);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const double n = s[u] * f[e];
                                if (y[u][e] < n) {
                                    y[u][e] = n; nx[u][e] = snx;
                                    if (PATHS) md[u][e] = b0 + kk;
                                }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NJ; ++u) {
            double2 *p = reinterpret_cast<double2 *>(a.rate + off[u]);
#pragma unroll
            for (int q = 0; q < 4; ++q) p[q] = make_double2(y[u][q * 2], y[u][q * 2 + 1]);
            int4 *pn = reinterpret_cast<int4 *>(a.next + off[u]);
            pn[0] = make_int4(nx[u][0], nx[u][1], nx[u][2], nx[u][3]);
            pn[1] = make_int4(nx[u][4], nx[u][5], nx[u][6], nx[u][7]);
            if (PATHS) {
                int4 *pm = reinterpret_cast<int4 *>(a.mid + off[u]);
                pm[0] = make_int4(md[u][0], md[u][1], md[u][2], md[u][3]);
                pm[1] = make_int4(md[u][4], md[u][5], md[u][6], md[u][7]);
            }
        }
    }
}

template <bool PATHS, int NJ>
__global__ void __launch_bounds__(512, 1) fw_colpanel_kernel(PanelArgs a) {
    colpanel_body<PATHS, NJ>(a, (int)blockIdx.x, (int)gridDim.x);
}

template <bool PATHS, int NJ>
__device__ __forceinline__ void rowpanel_body(const PanelArgs &a, const int bid, const int nctas) {
    constexpr int PANEL_JOBS = 32 * NJ;   // jobs per CTA pass
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Fs = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int b0 = a.b0;
    // factor matrix F[kk][i] = Cd[i][kk] = Cp[kk*ldc + blk_r0 + i]   (owner shard only)
    for (int idx = tid; idx < 128 * 128; idx += 512) {
        const int kk = idx >> 7, i = idx & 127;
        Fs[kk * PANEL_FP + swz128(i)] = a.Cp[(long long)kk * a.ldc + a.blk_r0 + i];
    }
    __syncthreads();

    const int job = tid >> 4, l = tid & 15;
    const int ncols = a.npad - FW_B;
    for (int g = bid; g * PANEL_JOBS < ncols; g += nctas) {
        // NJ adjacent columns per half-warp (never straddling the k-block: b0 and NJ divide 128)
        const int jp = g * PANEL_JOBS + job * NJ;
        const int j = jp < b0 ? jp : jp + FW_B;
        const long long off = (long long)(a.blk_r0 + l * 8) * a.ld + j;  // + c*ld + u
        double x[NJ][8];
        int m[NJ][8], mo[NJ][8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int u = 0; u < NJ; ++u) {
                x[u][c] = a.rate[off + (long long)c * a.ld + u];
                m[u][c] = -1;
                if (PATHS) mo[u][c] = a.mid[off + (long long)c * a.ld + u];
            }
        for (int lt = 0; lt < 16; ++lt) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int kk = lt * 8 + c;
                double s[NJ];
#pragma unroll
                for (int u = 0; u < NJ; ++u) s[u] = __shfl_sync(0xffffffffu, x[u][c], lt, 16);
                if (l == lt) {
                    // the step-kk snapshot of row b0+kk, written as it is taken
#pragma unroll
                    for (int u = 0; u < NJ; ++u) {
                        a.Rw[(long long)kk * a.ldw + j + u] = s[u];
                        if (PATHS) a.rs[off + (long long)c * a.ld + u] = (m[u][c] >= 0) ? b0 + m[u][c] : mo[u][c];
                    }
                }
                const double2 *fr = reinterpret_cast<const double2 *>(Fs + kk * PANEL_FP) + l;
                double f[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) { double2 v = fr[q * 16]; f[q * 2] = v.x; f[q * 2 + 1] = v.y; }
                int hu[NJ];
#pragma unroll
                for (int u = 0; u < NJ; ++u) {
                    int h[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[e] = __double2hiint(__fma_rd(f[e], s[u], -x[u][e]));
                    hu[u] = and8(h);
                }
                int acc = hu[0];
#pragma unroll
                for (int u = 1; u < NJ; ++u) acc &= hu[u];
                if (__any_sync(0xffffffffu, acc >= 0)) {
#pragma unroll
                    for (int u = 0; u < NJ; ++u) {
                        if (__any_sync(0xffffffffu, hu[u] >= 0)) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const double n = f[e] * s[u];
                                if (x[u][e] < n) { x[u][e] = n; m[u][e] = kk; }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int u = 0; u < NJ; ++u) {
                const long long o = off + (long long)c * a.ld + u;
                if (m[u][c] >= 0) {
                    a.rate[o] = x[u][c];
                    a.next[o] = a.NCp[(long long)(a.blk_r0 + l * 8 + c) * FW_B + m[u][c]];
                    if (PATHS) a.mid[o] = b0 + m[u][c];
                }
            }
    }
}

template <bool PATHS, int NJ>
__global__ void __launch_bounds__(512, 1) fw_rowpanel_kernel(PanelArgs a) {
    rowpanel_body<PATHS, NJ>(a, (int)blockIdx.x, (int)gridDim.x);
}

// Both panels of a k-block in ONE launch: they depend only on the diagonal tile's snapshots and write disjoint
// strips, so the first col_ctas CTAs take the column panel and the others the row panel.  At small N (a panel
// fills a fraction of the SMs) the two then run side by side instead of one after the other.
template <bool PATHS, int NJ>
__global__ void __launch_bounds__(512, 1) fw_panels_kernel(PanelArgs a, int col_ctas) {
    if ((int)blockIdx.x < col_ctas) colpanel_body<PATHS, NJ>(a, (int)blockIdx.x, col_ctas);
    else rowpanel_body<PATHS, NJ>(a, (int)blockIdx.x - col_ctas, (int)gridDim.x - col_ctas);
}

}  // namespace fw
