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

template <bool PATHS>
__global__ void __launch_bounds__(512, 1) fw_colpanel_kernel(PanelArgs a) {
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
    for (int g = blockIdx.x; g * 32 < nrows; g += gridDim.x) {
        const int rp = g * 32 + job;
        const int i = rp < a.skip_r0 ? rp : rp + a.skipn;   // local row, skipping the k-block rows
        const long long off = (long long)i * a.ld + b0 + l * 8;
        double y[8], cs[8];
        int nx[8], ncs[8], md[8], mcs[8];
        {
            const double2 *p = reinterpret_cast<const double2 *>(a.rate + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) { double2 v = p[q]; y[q * 2] = v.x; y[q * 2 + 1] = v.y; }
            const int4 *pn = reinterpret_cast<const int4 *>(a.next + off);
            int4 n0 = pn[0], n1 = pn[1];
            nx[0] = n0.x; nx[1] = n0.y; nx[2] = n0.z; nx[3] = n0.w;
            nx[4] = n1.x; nx[5] = n1.y; nx[6] = n1.z; nx[7] = n1.w;
            if (PATHS) {
                const int4 *pm = reinterpret_cast<const int4 *>(a.mid + off);
                int4 m0 = pm[0], m1 = pm[1];
                md[0] = m0.x; md[1] = m0.y; md[2] = m0.z; md[3] = m0.w;
                md[4] = m1.x; md[5] = m1.y; md[6] = m1.z; md[7] = m1.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) { cs[c] = 0.0; ncs[c] = -1; mcs[c] = -1; }

        for (int lt = 0; lt < 16; ++lt) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int kk = lt * 8 + c;
                const double s = __shfl_sync(0xffffffffu, y[c], lt, 16);
                if (l == lt) {
                    cs[c] = s; ncs[c] = nx[c];
                    if (PATHS) mcs[c] = md[c];
                }
                const double2 *fr = reinterpret_cast<const double2 *>(Fs + kk * PANEL_FP) + l;
                double f[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) { double2 v = fr[q * 16]; f[q * 2] = v.x; f[q * 2 + 1] = v.y; }
                // filter (fw_bulk.cuh): a set sign bit of fma_rd(s, f, -y) proves y < RN(s*f) is false
                int h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = __double2hiint(__fma_rd(s, f[e], -y[e]));
                const int acc = ((h[0] & h[1] & h[2]) & (h[3] & h[4] & h[5])) & (h[6] & h[7]);
                if (__any_sync(0xffffffffu, acc >= 0)) {
                    // exact path: one rounded multiply, strict compare (Algorithms.hs:55,61)
                    const int snx = __shfl_sync(0xffffffffu, nx[c], lt, 16);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const double n = s * f[e];
                        if (y[e] < n) {
                            y[e] = n; nx[e] = snx;
                            if (PATHS) md[e] = b0 + kk;
                        }
                    }
                }
            }
        }
        {
            double2 *p = reinterpret_cast<double2 *>(a.rate + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) p[q] = make_double2(y[q * 2], y[q * 2 + 1]);
            int4 *pn = reinterpret_cast<int4 *>(a.next + off);
            pn[0] = make_int4(nx[0], nx[1], nx[2], nx[3]);
            pn[1] = make_int4(nx[4], nx[5], nx[6], nx[7]);
            const long long poff = (long long)i * FW_B + l * 8;
#pragma unroll
            for (int c = 0; c < 8; ++c) a.Cp[(long long)(l * 8 + c) * a.ldc + i] = cs[c];
            int4 *pnc = reinterpret_cast<int4 *>(a.NCp + poff);
            pnc[0] = make_int4(ncs[0], ncs[1], ncs[2], ncs[3]);
            pnc[1] = make_int4(ncs[4], ncs[5], ncs[6], ncs[7]);
            if (PATHS) {
                int4 *pm = reinterpret_cast<int4 *>(a.mid + off);
                pm[0] = make_int4(md[0], md[1], md[2], md[3]);
                pm[1] = make_int4(md[4], md[5], md[6], md[7]);
                int4 *pcs = reinterpret_cast<int4 *>(a.csT + off);
                pcs[0] = make_int4(mcs[0], mcs[1], mcs[2], mcs[3]);
                pcs[1] = make_int4(mcs[4], mcs[5], mcs[6], mcs[7]);
            }
        }
    }
}

template <bool PATHS>
__global__ void __launch_bounds__(512, 1) fw_rowpanel_kernel(PanelArgs a) {
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
    for (int g = blockIdx.x; g * 32 < ncols; g += gridDim.x) {
        const int jp = g * 32 + job;
        const int j = jp < b0 ? jp : jp + FW_B;
        const long long off = (long long)(a.blk_r0 + l * 8) * a.ld + j;  // + c*ld
        double x[8], snap[8];
        int m[8], mo[8], msnap[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            x[c] = a.rate[off + (long long)c * a.ld];
            m[c] = -1; snap[c] = 0.0; msnap[c] = -1;
            if (PATHS) mo[c] = a.mid[off + (long long)c * a.ld];
        }
        for (int lt = 0; lt < 16; ++lt) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int kk = lt * 8 + c;
                const double s = __shfl_sync(0xffffffffu, x[c], lt, 16);
                if (l == lt) {
                    snap[c] = s;
                    if (PATHS) msnap[c] = (m[c] >= 0) ? b0 + m[c] : mo[c];
                }
                const double2 *fr = reinterpret_cast<const double2 *>(Fs + kk * PANEL_FP) + l;
                double f[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) { double2 v = fr[q * 16]; f[q * 2] = v.x; f[q * 2 + 1] = v.y; }
                int h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = __double2hiint(__fma_rd(f[e], s, -x[e]));
                const int acc = ((h[0] & h[1] & h[2]) & (h[3] & h[4] & h[5])) & (h[6] & h[7]);
                if (__any_sync(0xffffffffu, acc >= 0)) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const double n = f[e] * s;
                        if (x[e] < n) { x[e] = n; m[e] = kk; }
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const long long o = off + (long long)c * a.ld;
            if (m[c] >= 0) {
                a.rate[o] = x[c];
                a.next[o] = a.NCp[(long long)(a.blk_r0 + l * 8 + c) * FW_B + m[c]];
                if (PATHS) a.mid[o] = b0 + m[c];
            }
            a.Rw[(long long)(l * 8 + c) * a.ldw + j] = snap[c];
            if (PATHS) a.rs[o] = msnap[c];
        }
    }
}

}  // namespace fw
