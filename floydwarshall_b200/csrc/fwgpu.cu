// libfwgpu.so -- C ABI (include/fwgpu.h) over the sm_100a kernels.
//
// Replaces runAlgo (reference src/lib/Algorithms.hs:42-61) behind
// floydWarshall (Algorithms.hs:19-20).  Host orchestration of the exact-order
// blocked solve, per k-block of FW_B pivots:
//   1. fw_tile_kernel      pivot diagonal tile + step-k snapshots
//   2. fw_colpanel_kernel  pivot column panel  -> Cp / NCp snapshots (N x B)
//      fw_rowpanel_kernel  pivot row panel     -> Rw snapshots      (B x N)
//   3. fw_bulk_kernel      everything else against the two snapshot panels
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <numeric>
#include <vector>

#include "../../include/fwgpu.h"
struct fw_ctx;
#include "fw_bulk.cuh"
#include "fw_common.cuh"
#include "fw_panel.cuh"
#include "fw_paths.cuh"
#include "fw_tile.cuh"

static_assert(FW_B == FW_TILE, "kernel block size and public tile size must agree");

namespace {

// Error text: kept per OS thread (fw_last_error) AND per context (fw_ctx_last_error).  A GHC `safe` foreign
// call and the fw_last_error call after it may run on different OS threads, so bindings read the
// context's copy; every entry point names the context it works for with an ErrScope.
thread_local std::string g_err;
thread_local fw_ctx *g_err_ctx = nullptr;
void note_ctx_error(fw_ctx *c, const std::string &msg);

int fail(int code, const std::string &msg) {
    g_err = msg;
    note_ctx_error(g_err_ctx, msg);
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    note_ctx_error(g_err_ctx, g_err);
    return (e == cudaErrorMemoryAllocation) ? FW_ERR_NOMEM : FW_ERR_CUDA;
}
#define FW_ENTER(c) ErrScope es__(c); std::lock_guard<std::recursive_mutex> lk__((c)->mu)
struct ErrScope {
    fw_ctx *prev;
    explicit ErrScope(fw_ctx *c) : prev(g_err_ctx) { g_err_ctx = c; }
    ~ErrScope() { g_err_ctx = prev; }
};
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
    } while (0)

// ---------------------------------------------------------------- helper kernels
// flag bit 0: negative rate; bit 1: rate > 0 (off-diagonal) with next < 0
// Each of `batch` graphs is rows x n (rows == n unless it is a row shard: local row l is then global row
// ((l / cbr) * P + r) * cbr + l % cbr; rows at or beyond nvalid are padding and not checked).
__global__ void fw_validate_kernel(const double *rate, const int32_t *next, long long ld, long long stride,
                                   int rows, int n, int cbr, int P, int r, int nvalid, long long total, int *flag) {
    int bad = 0;
    const long long nn = (long long)rows * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long g = e / nn, rem = e - g * nn;
        const int i = (int)(rem / n), j = (int)(rem - (long long)i * n);
        const int gi = ((i / cbr) * P + r) * cbr + i % cbr;
        if (gi == j || gi >= nvalid || j >= nvalid) continue;  // the diagonal is never read (Algorithms.hs:50,54)
        const long long off = g * stride + (long long)i * ld + j;
        const double v = rate[off];
        if (v < 0.0) bad |= 1;
        if (v > 0.0 && next[off] < 0) bad |= 2;
    }
    if (bad) atomicOr(flag, bad);
}

// pad region of an npad x npad matrix that holds an n x n problem: NaN / -1
__global__ void fw_fill_pad_kernel(double *rate, int32_t *next, int32_t *mid, int32_t *csT, int32_t *rs,
                                   long long ld, int n, int npad) {
    const long long total = (long long)npad * npad;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / npad), j = (int)(e - (long long)i * npad);
        if (i < n && j < n) continue;
        const long long off = (long long)i * ld + j;
        rate[off] = fw::qnan();
        next[off] = -1;
        if (mid) { mid[off] = -1; csT[off] = -1; rs[off] = -1; }
    }
}

template <typename T>
__global__ void fw_copy2d_kernel(T *dst, long long dld, const T *src, long long sld, int rows, int cols) {
    const long long total = (long long)rows * cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / cols), j = (int)(e - (long long)i * cols);
        dst[(long long)i * dld + j] = src[(long long)i * sld + j];
    }
}

// buildMatrix (Algorithms.hs:26-40) on the device, step 1: diagonal and same-currency rule
__global__ void fw_build_base_kernel(double *rate, int32_t *next, long long ld, int n, const int32_t *ccy) {
    const long long total = (long long)n * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / n), j = (int)(e - (long long)i * n);
        const bool same = (i != j) && (ccy[i] == ccy[j]);          // :33 i == j first, :34 same currency
        rate[(long long)i * ld + j] = same ? 1.0 : 0.0;
        next[(long long)i * ld + j] = same ? j : -1;
    }
}
// step 2: the map entries (:35-36) -- never override the diagonal or a same-currency pair
__global__ void fw_build_edges_kernel(double *rate, int32_t *next, long long ld, int n, const int32_t *ccy,
                                      int m, const int32_t *src, const int32_t *dst, const double *val, int *flag) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < m; e += gridDim.x * blockDim.x) {
        const int i = src[e], j = dst[e];
        if (i < 0 || j < 0 || i >= n || j >= n) { atomicOr(flag, 4); continue; }
        if (i == j || ccy[i] == ccy[j]) continue;
        rate[(long long)i * ld + j] = val[e];
        next[(long long)i * ld + j] = j;
    }
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    int ensure(size_t n) {
        if (n <= cap) return FW_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc"); }
        cap = n;
        return FW_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct PinnedBuf {       // pinned (and mapped) host memory, grown on demand
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return FW_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc((void **)&p, n * sizeof(T), cudaHostAllocMapped);
        if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaHostAlloc"); }
        cap = n;
        return FW_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace


// scratch of the path entry points, pooled per context (no cudaMalloc / cudaFree per call)
struct fw_path_scratch {
    DevBuf<int32_t> q, verts;
    DevBuf<long long> len, off;
    DevBuf<unsigned long long> gstack;
    PinnedBuf<long long> h_len;
    PinnedBuf<unsigned char> h_opt;
    void release() { q.release(); verts.release(); len.release(); off.release(); gstack.release(); h_len.release(); h_opt.release(); }
};

struct fw_state;
struct fw_ctx {
    int device = 0;
    fw_state *edge_state = nullptr;   // cached device state of fw_solve_edges (buffers reused across calls)
    struct fw_path_scratch *pscratch = nullptr;   // pooled scratch of the path entry points
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // high priority: pivot phases of the NEXT k-blocks overlap the bulk kernel
    cudaStream_t cur = nullptr;           // stream the launch helpers use right now (stream or side_stream)
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
    bool overlap = true;                  // knob FW_OVERLAP=0: everything on one stream
    cudaStream_t stream = nullptr;
    // A context serialises its own calls (include/fwgpu.h).  Recursive: composite entry points
    // (fw_solve_edges, fw_solve_batched, fw_paths, fw_state_optimum) hold it across the calls they make.
    std::recursive_mutex mu;
    std::mutex err_mu;
    std::string err, err_ret;      // text of the last failure on this context (fw_ctx_last_error)
    // Lifetime: fw_state / fw_tables objects keep a reference, so a context destroyed before them (a garbage
    // collector may finalise the objects in any order) stays allocated as a closed shell until the last goes.
    std::atomic<int> refs{1};
    bool closed = false;
    int64_t launches = 0;
    // snapshot panels
    DevBuf<double> Cp[16], Rw[16];   // panel sets: k-blocks go in groups of up to 8, and the next group is factored ahead
    DevBuf<int32_t> NCp[16];
    int fuse_group = 0;            // k-blocks per fused bulk launch, 0 = by size: knob FW_FUSE_GROUP=1|2|4|8 (forces it for every
    bool fuse_forced = false;      // size); FW_FUSE_PAIRS=0 / =2 kept as aliases of FW_FUSE_GROUP=1 / =2
    int panel_nj = 0;              // knob FW_PANEL_NJ=1|2 forces the jobs per half-warp of the panel kernels (0: by size)
    bool merge_panels = true;      // knob FW_MERGE_PANELS=0: column and row panel as two launches
    // padded working copy (n not a multiple of FW_B) and host-API staging
    DevBuf<double> w_rate;
    DevBuf<int32_t> w_next, w_mid, w_csT, w_rs;
    // device staging of the host-buffer batched API
    DevBuf<double> s_rate;
    DevBuf<int32_t> s_next, s_mid, s_csT, s_rs;
    int *d_flag = nullptr;
    int *h_flag = nullptr;
    bool attrs_set = false;
    // verification hook (fw_ctx_set_row_snapshot_sink): every k-block's row-snapshot panel is also copied here
    double *snap_sink = nullptr;
    long long snap_ld = 0;
    int bulk_cq = 2;       // fw_bulk_kernel tile width / 32 (2: 8x4 micro-tile, 4: 8x8); knob FW_BULK_CQ
    int bulk_band = 64;    // tile columns per raster band of fw_bulk_kernel (L2 locality); knob FW_BULK_BAND
    // optional per-phase timing (CUDA events on the launching stream)
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int phase; };
    std::vector<Span> spans;      // spans of the last solve
    std::vector<Span> pool;       // recycled events
};

namespace {

extern std::atomic<fw_ctx *> g_default;
void note_ctx_error(fw_ctx *c, const std::string &msg) {
    if (!c) c = g_default.load();     // calls made with ctx == NULL belong to the default context
    if (!c) return;
    std::lock_guard<std::mutex> lk(c->err_mu);
    c->err = msg;
}

int set_kernel_attrs(fw_ctx *c) {
    if (c->attrs_set) return FW_OK;
    CU(cudaFuncSetAttribute(fw::fw_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::tile_smem_bytes(false)));
    CU(cudaFuncSetAttribute(fw::fw_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::tile_smem_bytes(true)));
    CU(cudaFuncSetAttribute(fw::fw_colpanel_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_colpanel_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_colpanel_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_colpanel_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_rowpanel_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_rowpanel_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_rowpanel_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_rowpanel_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::bulk_smem_bytes<2>()));
    CU(cudaFuncSetAttribute(fw::fw_bulk_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CU(cudaFuncSetAttribute(fw::fw_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fw::bulk_smem_bytes<4>()));
    CU(cudaFuncSetAttribute(fw::fw_bulk_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (const char *e = getenv("FW_BULK_CQ")) c->bulk_cq = atoi(e) == 4 ? 4 : 2;
    if (const char *e = getenv("FW_FUSE_PAIRS")) { c->fuse_group = atoi(e) != 0 ? 2 : 1; c->fuse_forced = atoi(e) == 2; }
    if (const char *e = getenv("FW_FUSE_GROUP")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) { c->fuse_group = v; c->fuse_forced = true; }
    }
    if (const char *e = getenv("FW_PANEL_NJ")) c->panel_nj = atoi(e);
    if (const char *e = getenv("FW_MERGE_PANELS")) c->merge_panels = atoi(e) != 0;
    CU(cudaFuncSetAttribute(fw::fw_panels_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_panels_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_panels_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw::panel_smem_bytes()));
    CU(cudaFuncSetAttribute(fw::fw_panels_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw::panel_smem_bytes()));
    if (const char *e = getenv("FW_BULK_BAND")) c->bulk_band = atoi(e) > 0 ? atoi(e) : 64;
    c->attrs_set = true;
    return FW_OK;
}

// phase ids: 0 tile (phase 1), 1 column panel, 2 row panel, 3 bulk (phase 3)
struct PhaseTimer {
    fw_ctx *c; fw_ctx::Span s; bool on;
    PhaseTimer(fw_ctx *c_, int phase) : c(c_), on(c_->profiling) {
        if (!on) return;
        if (!c->pool.empty()) { s = c->pool.back(); c->pool.pop_back(); }
        else { cudaEventCreate(&s.a); cudaEventCreate(&s.b); }
        s.phase = phase;
        cudaEventRecord(s.a, c->cur ? c->cur : c->stream);
    }
    ~PhaseTimer() {
        if (!on) return;
        cudaEventRecord(s.b, c->cur ? c->cur : c->stream);
        c->spans.push_back(s);
    }
};
void recycle_spans(fw_ctx *c) {
    for (auto &sp : c->spans) c->pool.push_back(sp);
    c->spans.clear();
}

int grid_for(long long total, int sm_count) {
    long long g = (total + 255) / 256;
    long long cap = (long long)sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Tile ranges are given in 64-row / 64-column units (see fw::BulkArgs); ncu = units of columns, nru = of rows.
void launch_bulk(fw_ctx *c, fw::BulkArgs g, int ncu, int nru) {
    if (ncu <= 0 || nru <= 0) return;
    PhaseTimer pt(c, 3);
    g.gy = nru;
    g.band = c->bulk_band;
    if (c->bulk_cq == 4) {
        g.gx = ncu / 2;
        if (g.band > g.gx) g.band = g.gx;
        fw::fw_bulk_kernel<4><<<g.gx * g.gy, 128, fw::bulk_smem_bytes<4>(), (c->cur ? c->cur : c->stream)>>>(g);
    } else {
        g.gx = ncu;
        if (g.band > g.gx) g.band = g.gx;
        fw::fw_bulk_kernel<2><<<g.gx * g.gy, 128, fw::bulk_smem_bytes<2>(), (c->cur ? c->cur : c->stream)>>>(g);
    }
    c->launches++;
}

constexpr int NOSKIP = 0x3fffffff;

// One phase-2 panel launch: `jobs` matrix rows (column panel) or columns (row panel), NJ per half-warp.
template <bool COL>
void launch_panel(fw_ctx *c, const fw::PanelArgs &p, int jobs, bool paths, cudaStream_t st) {
    const int nj = (c->panel_nj == 1 || c->panel_nj == 2) ? c->panel_nj : fw::panel_nj(jobs, c->sm_count);
    const int passes = jobs / (32 * nj);
    const int grid = passes < c->sm_count ? passes : c->sm_count;
    const size_t sm = fw::panel_smem_bytes();
    if (COL) {
        if (paths) { if (nj == 2) fw::fw_colpanel_kernel<true, 2><<<grid, 512, sm, st>>>(p); else fw::fw_colpanel_kernel<true, 1><<<grid, 512, sm, st>>>(p); }
        else       { if (nj == 2) fw::fw_colpanel_kernel<false, 2><<<grid, 512, sm, st>>>(p); else fw::fw_colpanel_kernel<false, 1><<<grid, 512, sm, st>>>(p); }
    } else {
        if (paths) { if (nj == 2) fw::fw_rowpanel_kernel<true, 2><<<grid, 512, sm, st>>>(p); else fw::fw_rowpanel_kernel<true, 1><<<grid, 512, sm, st>>>(p); }
        else       { if (nj == 2) fw::fw_rowpanel_kernel<false, 2><<<grid, 512, sm, st>>>(p); else fw::fw_rowpanel_kernel<false, 1><<<grid, 512, sm, st>>>(p); }
    }
    c->launches++;
}


// Column and row panel of one k-block in one launch (both have `jobs` jobs).
void launch_both_panels(fw_ctx *c, const fw::PanelArgs &p, int jobs, bool paths, cudaStream_t st) {
    const int nj = (c->panel_nj == 1 || c->panel_nj == 2) ? c->panel_nj : fw::panel_nj(jobs, c->sm_count);
    const int passes = jobs / (32 * nj);
    const int each = passes < c->sm_count ? passes : c->sm_count;
    const size_t sm = fw::panel_smem_bytes();
    if (paths) { if (nj == 2) fw::fw_panels_kernel<true, 2><<<2 * each, 512, sm, st>>>(p, each); else fw::fw_panels_kernel<true, 1><<<2 * each, 512, sm, st>>>(p, each); }
    else       { if (nj == 2) fw::fw_panels_kernel<false, 2><<<2 * each, 512, sm, st>>>(p, each); else fw::fw_panels_kernel<false, 1><<<2 * each, 512, sm, st>>>(p, each); }
    c->launches++;
}

// Domain check (synchronises the stream once).
int validate_device(fw_ctx *c, const double *rate, const int32_t *next, long long ld, long long stride,
                    int batch, int n, int rows = -1, int cbr = 1 << 30, int P = 1, int r = 0, int nvalid = -1,
                    bool sync = true) {
    if (rows < 0) rows = n;
    if (nvalid < 0) nvalid = n;
    CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
    const long long total = (long long)batch * rows * n;
    fw_validate_kernel<<<grid_for(total, c->sm_count), 256, 0, (c->cur ? c->cur : c->stream)>>>(rate, next, ld, stride, rows, n, cbr, P, r,
                                                                             nvalid, total, c->d_flag);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (!sync) return FW_OK;            // the caller collects h_flag of several devices after one round of syncs
    CU(cudaStreamSynchronize(c->stream));
    if (*c->h_flag & 1) return fail(FW_ERR_DOMAIN, "rate matrix holds a negative entry");
    if (*c->h_flag & 2) return fail(FW_ERR_DOMAIN, "rate > 0 with next < 0 (inconsistent next-hop matrix)");
    return FW_OK;
}

// Phase 1 + 2 of k-block b0 into panel set `set`: diagonal tile, column panel, row panel.
int launch_pivot_phases(fw_ctx *c, int npad, long long ld, double *rate, int32_t *next, int32_t *mid, int32_t *csT,
                        int32_t *rs, int b0, int set, bool panels) {
    const bool paths = (mid != nullptr);
    fw::TileArgs t;
    t.rate = rate; t.next = next; t.mid = mid; t.csT = csT; t.rs = rs;
    t.ld = ld; t.batch_stride = 0; t.b0 = b0; t.r0 = b0; t.nv = FW_B;
    t.Cp = c->Cp[set].p; t.ldc = npad; t.NCp = c->NCp[set].p; t.Rw = c->Rw[set].p; t.ldw = npad;
    {
        PhaseTimer pt(c, 0);
        if (paths)
            fw::fw_tile_kernel<true><<<1, 512, fw::tile_smem_bytes(true), (c->cur ? c->cur : c->stream)>>>(t);
        else
            fw::fw_tile_kernel<false><<<1, 512, fw::tile_smem_bytes(false), (c->cur ? c->cur : c->stream)>>>(t);
    }
    c->launches++;
    if (panels) {
        fw::PanelArgs p;
        p.rate = rate; p.next = next; p.mid = mid; p.csT = csT; p.rs = rs;
        p.ld = ld; p.npad = npad; p.b0 = b0; p.rows = npad; p.blk_r0 = b0; p.skip_r0 = b0; p.skipn = FW_B;
        p.Cp = c->Cp[set].p; p.ldc = npad; p.NCp = c->NCp[set].p; p.Rw = c->Rw[set].p; p.ldw = npad;
        if (c->merge_panels) {
            PhaseTimer pt(c, 1);          // column + row panel, one launch (reported under "col_panel")
            launch_both_panels(c, p, npad - FW_B, paths, (c->cur ? c->cur : c->stream));
        } else {
            {
                PhaseTimer pt(c, 1);
                launch_panel<true>(c, p, npad - FW_B, paths, (c->cur ? c->cur : c->stream));
            }
            {
                PhaseTimer pt(c, 2);
                launch_panel<false>(c, p, npad - FW_B, paths, (c->cur ? c->cur : c->stream));
            }
        }
    }
    CU(cudaGetLastError());
    if (c->snap_sink)   // rows b0 .. b0+127 as their own steps began (the pivot entry itself is exported as 0.0)
        CU(cudaMemcpy2DAsync(c->snap_sink + (long long)b0 * c->snap_ld, (size_t)c->snap_ld * 8, c->Rw[set].p,
                             (size_t)npad * 8, (size_t)npad * 8, FW_B, cudaMemcpyDeviceToDevice,
                             (c->cur ? c->cur : c->stream)));
    return FW_OK;
}

// Blocked solve on a padded (npad % FW_B == 0, pads = NaN) device matrix.
//
// k-blocks are taken in GROUPS of G = 1, 2 or 4 consecutive blocks so that the bulk kernel loads every
// tile once per G*128 steps, and the pivot phases of the NEXT group run on a high-priority side stream
// while the bulk of the current group is still busy:
//
//   main stream : bulk(G) on the strips of G'  | record E1 |  bulk(G) on everything else
//   side stream :                    wait E1 -> for every block b'+j of G' in turn: the earlier blocks of G'
//                                               on the strips of b'+j (one fused launch of j blocks per
//                                               strip direction), then phases 1+2 of b'+j | record E2
//   next group  : main waits E2
//
// bulk(G) gives tiles outside all strips of G all G*128 steps, and a tile in the strips of the group's
// block i only the blocks after i (phase 2 of block i and the strip launches gave it blocks 0..i); the
// strips of the group's last block are complete and left out.  The order of relaxations seen by every
// entry is unchanged (ascending k), so results are identical to the plain loop.
int solve_blocked(fw_ctx *c, int npad, long long ld, double *rate, int32_t *next, int32_t *mid,
                  int32_t *csT, int32_t *rs, int kb0 = 0, int kb1 = -1) {
    constexpr int GMAX = fw::BULK_MAXNB;
    int rc;
    const int nblk = npad / FW_B;
    if (kb1 < 0) kb1 = nblk;            // k-blocks [kb0, kb1) only (fw_solve_device_range); the group size follows the full solve
    const int nu = npad / 64;                 // 64-row / 64-column units
    // grouping splits every bulk launch in three and adds strip launches; below ~48 k-blocks the extra
    // launches cost more than the saved tile loads (N=1024: 2.26 ms ungrouped vs 2.50 ms in pairs)
    // Measured (ms per solve, groups of 1 / 2 / 4 / 8): N=8192 90.5 / 87.2 / 84.6 / 86.5, N=16384 - / 607 / 589 / 581,
    // N=32768 - / 4560 / 4441 / -.  Larger groups move more work into the small strip launches.
    const int gsz = c->fuse_forced ? c->fuse_group
                                   : (c->fuse_group == 1 ? 1 : (nblk >= 128 ? 8 : (nblk >= 48 ? 4 : 1)));
    for (int set = 0; set < (gsz == 1 ? 2 : 2 * gsz); ++set) {
        if ((rc = c->Cp[set].ensure((size_t)npad * FW_B)) != FW_OK) return rc;
        if ((rc = c->NCp[set].ensure((size_t)npad * FW_B)) != FW_OK) return rc;
        if ((rc = c->Rw[set].ensure((size_t)npad * FW_B)) != FW_OK) return rc;
    }
    cudaStream_t S = c->stream;
    cudaStream_t T = (c->overlap && c->side_stream) ? c->side_stream : c->stream;
    const bool two = (T != S);
    fw::BulkArgs g;
    g.rate = rate; g.next = next; g.mid = mid; g.ld = ld; g.row0 = 0; g.cbr = 1 << 30; g.P = 1; g.r = 0; g.ldc = npad; g.ldw = npad;
    auto use_sets = [&](int s0) {
        for (int i = 0; i < GMAX; ++i) {
            const int s = s0 + (i < gsz ? i : 0);
            g.CpT[i] = c->Cp[s].p; g.NCp[i] = c->NCp[s].p; g.Rw[i] = c->Rw[s].p;
        }
    };
    // phases 1+2 of the group starting at block b (gn blocks) into panel sets s0 .. s0+gn-1 -- on c->cur
    auto pivot_group = [&](int b, int gn, int s0) -> int {
        const int b0 = b * FW_B, u0 = b0 / 64;
        int r = launch_pivot_phases(c, npad, ld, rate, next, mid, csT, rs, b0, s0, nblk > 1);
        for (int j = 1; j < gn && r == FW_OK; ++j) {
            // blocks b .. b+j-1 on the strips of block b+j (its pivot rows / columns must be current);
            // the strips of b+j-1 are complete, tiles in the strips of an earlier block start after it
            const int uj = u0 + 2 * j;
            use_sets(s0);
            g.b0 = b0; g.nb = j; g.half_r0 = u0; g.half_c0 = u0;
            g.row_lo = uj; g.rskip0 = NOSKIP; g.rskipn = 0;          // the 2 tile rows of b+j ...
            g.col_lo = 0; g.cskip0 = uj - 2; g.cskipn = 2;           // ... x all columns outside b+j-1
            launch_bulk(c, g, nu - 2, 2);
            g.row_lo = 0; g.rskip0 = uj - 2; g.rskipn = 4;           // all rows outside b+j-1 and b+j ...
            g.col_lo = uj; g.cskip0 = NOSKIP; g.cskipn = 0;          // ... x the 2 tile columns of b+j
            launch_bulk(c, g, 2, nu - 4);
            r = launch_pivot_phases(c, npad, ld, rate, next, mid, csT, rs, b0 + j * FW_B, s0 + j, true);
        }
        return r;
    };

    if (two) { CU(cudaEventRecord(c->ev_main, S)); CU(cudaStreamWaitEvent(T, c->ev_main, 0)); }
    int b = kb0, sbase = 0;
    int gn = (kb1 - b < gsz) ? kb1 - b : gsz;
    c->cur = T;
    if ((rc = pivot_group(b, gn, sbase)) != FW_OK) { c->cur = nullptr; return rc; }
    if (two) CU(cudaEventRecord(c->ev_side, T));
    while (b < kb1 && nblk > 1) {
        const int b0 = b * FW_B, u0 = b0 / 64;
        const int bn = b + gn;                                          // first block of the next group
        const int gnn = (kb1 - bn < gsz) ? kb1 - bn : gsz;              // its size (0: none)
        const int uL = u0 + 2 * (gn - 1);                               // units of the LAST block of this group
        c->cur = S;
        if (two) CU(cudaStreamWaitEvent(S, c->ev_side, 0));             // panels of this group are ready
        use_sets(sbase);
        g.b0 = b0; g.nb = gn;
        g.half_r0 = (gn > 1) ? u0 : NOSKIP; g.half_c0 = (gn > 1) ? u0 : NOSKIP;
        if (gnn > 0) {
            const int uN = uL + 2;                                      // units of the next group: [uN, uN + 2*gnn)
            // (1) the next group's pivot rows / columns first
            g.row_lo = uN; g.rskip0 = NOSKIP; g.rskipn = 0;
            g.col_lo = 0; g.cskip0 = uL; g.cskipn = 2;
            launch_bulk(c, g, nu - 2, 2 * gnn);
            g.row_lo = 0; g.rskip0 = uL; g.rskipn = 2 + 2 * gnn;
            g.col_lo = uN; g.cskip0 = NOSKIP; g.cskipn = 0;
            launch_bulk(c, g, 2 * gnn, nu - 2 - 2 * gnn);
            if (two) { CU(cudaEventRecord(c->ev_main, S)); CU(cudaStreamWaitEvent(T, c->ev_main, 0)); }
            // (2) everything else of this group's bulk ...
            g.row_lo = 0; g.rskip0 = uL; g.rskipn = 2 + 2 * gnn;
            g.col_lo = 0; g.cskip0 = uL; g.cskipn = 2 + 2 * gnn;
            launch_bulk(c, g, nu - 2 - 2 * gnn, nu - 2 - 2 * gnn);
            // ... while the side stream factors the next group
            c->cur = T;
            if ((rc = pivot_group(bn, gnn, sbase ^ gsz)) != FW_OK) { c->cur = nullptr; return rc; }
            if (two) CU(cudaEventRecord(c->ev_side, T));
        } else {
            g.row_lo = 0; g.rskip0 = uL; g.rskipn = 2;
            g.col_lo = 0; g.cskip0 = uL; g.cskipn = 2;
            launch_bulk(c, g, nu - 2, nu - 2);
        }
        CU(cudaGetLastError());
        b = bn; gn = gnn; sbase ^= gsz;
    }
    c->cur = nullptr;
    if (two) { CU(cudaEventRecord(c->ev_side, T)); CU(cudaStreamWaitEvent(S, c->ev_side, 0)); }
    return FW_OK;
}

int fill_minus1_2d(fw_ctx *c, int32_t *p, long long ld, int rows, int cols) {
    CU(cudaMemset2DAsync(p, (size_t)ld * 4, 0xFF, (size_t)cols * 4, (size_t)rows, c->stream));
    return FW_OK;
}

// n <= FW_B graphs, one CTA each, directly on the caller's layout.
int solve_tiles(fw_ctx *c, int batch, int n, long long ld, long long stride, double *rate, int32_t *next,
                int32_t *mid, int32_t *csT, int32_t *rs) {
    const bool paths = (mid != nullptr);
    fw::TileArgs t;
    t.rate = rate; t.next = next; t.mid = mid; t.csT = csT; t.rs = rs;
    t.ld = ld; t.batch_stride = stride; t.b0 = 0; t.r0 = 0; t.nv = n;
    t.Cp = nullptr; t.ldc = 0; t.NCp = nullptr; t.Rw = nullptr; t.ldw = 0;
    {
        PhaseTimer pt(c, 0);
        if (paths)
            fw::fw_tile_kernel<true><<<batch, 512, fw::tile_smem_bytes(true), c->stream>>>(t);
        else
            fw::fw_tile_kernel<false><<<batch, 512, fw::tile_smem_bytes(false), c->stream>>>(t);
    }
    c->launches++;
    CU(cudaGetLastError());
    return FW_OK;
}

int solve_device_locked(fw_ctx *c, int n, long long ld, double *rate, int32_t *next, int32_t *mid,
                        int32_t *csT, int32_t *rs, bool validate) {
    int rc;
    c->cur = nullptr;
    if ((rc = set_kernel_attrs(c)) != FW_OK) return rc;
    if (validate && (rc = validate_device(c, rate, next, ld, 0, 1, n)) != FW_OK) return rc;
    const bool paths = (mid != nullptr);
    if (n <= FW_B) {
        if (paths) {
            if ((rc = fill_minus1_2d(c, mid, ld, n, n)) != FW_OK) return rc;
            if ((rc = fill_minus1_2d(c, csT, ld, n, n)) != FW_OK) return rc;
            if ((rc = fill_minus1_2d(c, rs, ld, n, n)) != FW_OK) return rc;
        }
        return solve_tiles(c, 1, n, ld, 0, rate, next, mid, csT, rs);
    }
    const bool aligned = (ld % 4 == 0) && !((uintptr_t)rate & 15) && !((uintptr_t)next & 15) &&
                         (!paths || (!((uintptr_t)mid & 15) && !((uintptr_t)csT & 15) && !((uintptr_t)rs & 15)));
    if (n % FW_B == 0 && aligned) {
        if (paths) {
            if ((rc = fill_minus1_2d(c, mid, ld, n, n)) != FW_OK) return rc;
            if ((rc = fill_minus1_2d(c, csT, ld, n, n)) != FW_OK) return rc;
            if ((rc = fill_minus1_2d(c, rs, ld, n, n)) != FW_OK) return rc;
        }
        return solve_blocked(c, n, ld, rate, next, mid, csT, rs);
    }
    // padded working copy
    const int npad = (n + FW_B - 1) / FW_B * FW_B;
    const size_t tot = (size_t)npad * npad;
    if ((rc = c->w_rate.ensure(tot)) != FW_OK) return rc;
    if ((rc = c->w_next.ensure(tot)) != FW_OK) return rc;
    if (paths) {
        if ((rc = c->w_mid.ensure(tot)) != FW_OK) return rc;
        if ((rc = c->w_csT.ensure(tot)) != FW_OK) return rc;
        if ((rc = c->w_rs.ensure(tot)) != FW_OK) return rc;
        CU(cudaMemsetAsync(c->w_mid.p, 0xFF, tot * 4, c->stream));
        CU(cudaMemsetAsync(c->w_csT.p, 0xFF, tot * 4, c->stream));
        CU(cudaMemsetAsync(c->w_rs.p, 0xFF, tot * 4, c->stream));
    }
    const int g = grid_for((long long)n * n, c->sm_count);
    fw_copy2d_kernel<double><<<g, 256, 0, c->stream>>>(c->w_rate.p, npad, rate, ld, n, n);
    fw_copy2d_kernel<int32_t><<<g, 256, 0, c->stream>>>(c->w_next.p, npad, next, ld, n, n);
    fw_fill_pad_kernel<<<grid_for((long long)tot, c->sm_count), 256, 0, c->stream>>>(
        c->w_rate.p, c->w_next.p, nullptr, nullptr, nullptr, npad, n, npad);
    c->launches += 3;
    CU(cudaGetLastError());
    rc = solve_blocked(c, npad, npad, c->w_rate.p, c->w_next.p, paths ? c->w_mid.p : nullptr,
                       paths ? c->w_csT.p : nullptr, paths ? c->w_rs.p : nullptr);
    if (rc != FW_OK) return rc;
    fw_copy2d_kernel<double><<<g, 256, 0, c->stream>>>(rate, ld, c->w_rate.p, npad, n, n);
    fw_copy2d_kernel<int32_t><<<g, 256, 0, c->stream>>>(next, ld, c->w_next.p, npad, n, n);
    c->launches += 2;
    if (paths) {
        fw_copy2d_kernel<int32_t><<<g, 256, 0, c->stream>>>(mid, ld, c->w_mid.p, npad, n, n);
        fw_copy2d_kernel<int32_t><<<g, 256, 0, c->stream>>>(csT, ld, c->w_csT.p, npad, n, n);
        fw_copy2d_kernel<int32_t><<<g, 256, 0, c->stream>>>(rs, ld, c->w_rs.p, npad, n, n);
        c->launches += 3;
    }
    CU(cudaGetLastError());
    return FW_OK;
}

std::atomic<fw_ctx *> g_default{nullptr};
std::mutex g_default_mu;

int get_ctx(fw_ctx *in, fw_ctx **out) {
    if (in) { *out = in; return FW_OK; }
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default.load()) {
        fw_ctx *d = nullptr;
        int rc = fw_ctx_create(0, &d);
        if (rc != FW_OK) return rc;
        g_default.store(d);
    }
    *out = g_default.load();
    return FW_OK;
}

bool paths_args_ok(const int32_t *mid, const int32_t *csT, const int32_t *rs) {
    const int k = (mid != nullptr) + (csT != nullptr) + (rs != nullptr);
    return k == 0 || k == 3;
}

}  // namespace

extern "C" {

const char *fw_version(void) { return "fwgpu 0.1 (sm_100a, exact-order blocked max-times Floyd-Warshall)"; }
const char *fw_last_error(void) { return g_err.c_str(); }

// Text of the last failure of a call made on `ctx` (NULL: the process-wide default context), valid until the
// next failing call on that context.  Unlike fw_last_error it does not depend on the calling OS thread.
const char *fw_ctx_last_error(fw_ctx *c) {
    if (!c) {
        c = g_default.load();
        if (!c) return g_err.c_str();   // the default context could not even be created: this thread's text
    }
    std::lock_guard<std::mutex> lk(c->err_mu);
    c->err_ret = c->err;                 // stable copy handed to the caller
    return c->err_ret.c_str();
}

int fw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fw_ctx_create(int device, fw_ctx **out) {
    if (!out || device < 0) return fail(FW_ERR_INVALID, "fw_ctx_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(FW_ERR_CUDA, "no CUDA device available (libfwgpu has no CPU fallback)");
    if (device >= ndev) return fail(FW_ERR_INVALID, "fw_ctx_create: device index out of range");
    CU(cudaSetDevice(device));
    fw_ctx *c = new (std::nothrow) fw_ctx();
    if (!c) return fail(FW_ERR_NOMEM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c; return cuda_fail(e, "cudaStreamCreate");
    }
    c->stream = c->own_stream;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi) != cudaSuccess) c->side_stream = nullptr;
        cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming);
        if (const char *e = getenv("FW_OVERLAP")) c->overlap = atoi(e) != 0;
    }
    if ((e = cudaMalloc(&c->d_flag, sizeof(int))) != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc"); }
    if ((e = cudaMallocHost(&c->h_flag, sizeof(int))) != cudaSuccess) { delete c; return cuda_fail(e, "cudaMallocHost"); }
    *out = c;
    return FW_OK;
}

static void ctx_unref(fw_ctx *c) {
    if (c->refs.fetch_sub(1) == 1) delete c;
}

void fw_ctx_destroy(fw_ctx *c) {
    if (!c) return;
    {
        std::lock_guard<std::recursive_mutex> lk(c->mu);
        if (c->closed) return;
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->edge_state) { fw_state_destroy(c->edge_state); c->edge_state = nullptr; }
        if (c->pscratch) { c->pscratch->release(); delete c->pscratch; c->pscratch = nullptr; }
        for (int i = 0; i < 16; ++i) { c->Cp[i].release(); c->Rw[i].release(); c->NCp[i].release(); }
        c->w_rate.release(); c->w_next.release(); c->w_mid.release(); c->w_csT.release(); c->w_rs.release();
        c->s_rate.release(); c->s_next.release(); c->s_mid.release(); c->s_csT.release(); c->s_rs.release();
        if (c->d_flag) cudaFree(c->d_flag);
        if (c->h_flag) cudaFreeHost(c->h_flag);
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        if (c->side_stream) cudaStreamDestroy(c->side_stream);
        if (c->ev_main) cudaEventDestroy(c->ev_main);
        if (c->ev_side) cudaEventDestroy(c->ev_side);
        c->d_flag = nullptr; c->h_flag = nullptr; c->own_stream = c->side_stream = c->stream = nullptr;
        c->ev_main = c->ev_side = nullptr;
        c->closed = true;
    }
    ctx_unref(c);
}

int fw_ctx_set_stream(fw_ctx *c, void *cuda_stream, int external) {
    ErrScope es0__(c);
    if (!c) return fail(FW_ERR_INVALID, "fw_ctx_set_stream: null context");
    FW_ENTER(c);
    c->stream = external ? (cudaStream_t)cuda_stream : c->own_stream;
    return FW_OK;
}

int64_t fw_ctx_last_launches(const fw_ctx *c) { return c ? c->launches : 0; }

int fw_ctx_set_profiling(fw_ctx *c, int on) {
    ErrScope es0__(c);
    if (!c) return fail(FW_ERR_INVALID, "fw_ctx_set_profiling: null context");
    FW_ENTER(c);
    c->profiling = (on != 0);
    return FW_OK;
}

int64_t fw_ctx_phase_spans(fw_ctx *c, int phase, double *ms, int64_t cap) {
    ErrScope es0__(c);
    if (!c || phase < 0 || phase > 3 || (!ms && cap > 0)) return fail(FW_ERR_INVALID, "fw_ctx_phase_spans: bad argument");
    FW_ENTER(c);
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)
        return fail(FW_ERR_CUDA, "fw_ctx_phase_spans: stream synchronize failed");
    int64_t n = 0;
    for (auto &sp : c->spans) {
        if (sp.phase != phase) continue;
        if (n < cap) {
            float t = 0.f;
            cudaEventElapsedTime(&t, sp.a, sp.b);
            ms[n] = t;
        }
        ++n;
    }
    return n;
}

int fw_ctx_phase_ms(fw_ctx *c, double ms[4], int64_t count[4]) {
    ErrScope es0__(c);
    if (!c || !ms || !count) return fail(FW_ERR_INVALID, "fw_ctx_phase_ms: bad argument");
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 4; ++i) { ms[i] = 0.0; count[i] = 0; }
    for (auto &sp : c->spans) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, sp.a, sp.b));
        ms[sp.phase] += t;
        count[sp.phase]++;
    }
    return FW_OK;
}

int fw_ctx_synchronize(fw_ctx *c) {
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return FW_OK;
}

int fw_solve_device(fw_ctx *c, int32_t n, int64_t ld, double *d_rate, int32_t *d_next, int32_t *d_mid,
                    int32_t *d_csT, int32_t *d_rs) {
    ErrScope es0__(c);
    if (n < 0) return fail(FW_ERR_INVALID, "fw_solve_device: n < 0");
    if (n == 0) return FW_OK;
    if (!d_rate || !d_next || ld < n) return fail(FW_ERR_INVALID, "fw_solve_device: bad argument");
    if (!paths_args_ok(d_mid, d_csT, d_rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    recycle_spans(c);
    return solve_device_locked(c, n, ld, d_rate, d_next, d_mid, d_csT, d_rs, true);
}

int fw_solve_device_range(fw_ctx *c, int32_t n, int64_t ld, double *d_rate, int32_t *d_next, int32_t *d_mid,
                          int32_t *d_csT, int32_t *d_rs, int32_t kb0, int32_t kb1) {
    if (n <= 0 || n % FW_B) return fail(FW_ERR_INVALID, "fw_solve_device_range: n must be a positive multiple of 128");
    if (!d_rate || !d_next || ld < n || ld % 4 || ((uintptr_t)d_rate & 15) || ((uintptr_t)d_next & 15))
        return fail(FW_ERR_INVALID, "fw_solve_device_range: bad argument (ld % 4 == 0, 16-byte aligned)");
    if (!paths_args_ok(d_mid, d_csT, d_rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    if (kb0 < 0 || kb1 < kb0 || kb1 > n / FW_B) return fail(FW_ERR_INVALID, "fw_solve_device_range: bad k-block range");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    recycle_spans(c);
    c->cur = nullptr;
    if ((rc = set_kernel_attrs(c)) != FW_OK) return rc;
    if (kb0 == 0) {
        if ((rc = validate_device(c, d_rate, d_next, ld, 0, 1, n)) != FW_OK) return rc;
        if (d_mid) {
            if ((rc = fill_minus1_2d(c, d_mid, ld, n, n)) != FW_OK) return rc;
            if ((rc = fill_minus1_2d(c, d_csT, ld, n, n)) != FW_OK) return rc;
            if ((rc = fill_minus1_2d(c, d_rs, ld, n, n)) != FW_OK) return rc;
        }
    }
    if (kb0 == kb1) return FW_OK;
    if (n == FW_B) return solve_tiles(c, 1, n, ld, 0, d_rate, d_next, d_mid, d_csT, d_rs);
    return solve_blocked(c, n, ld, d_rate, d_next, d_mid, d_csT, d_rs, kb0, kb1);
}

int fw_ctx_set_row_snapshot_sink(fw_ctx *c, double *d_sink, int64_t ld) {
    if (!c || (d_sink && ld <= 0)) return fail(FW_ERR_INVALID, "fw_ctx_set_row_snapshot_sink: bad argument");
    FW_ENTER(c);
    c->snap_sink = d_sink;
    c->snap_ld = ld;
    return FW_OK;
}

int fw_solve(fw_ctx *c, int32_t n, double *rate, int32_t *next, int32_t *mid, int32_t *csT, int32_t *rs) {
    ErrScope es0__(c);
    if (n < 0) return fail(FW_ERR_INVALID, "fw_solve: n < 0");
    if (n == 0) return FW_OK;  // floydWarshall M.empty == V.empty
    if (!rate || !next) return fail(FW_ERR_INVALID, "fw_solve: null buffer");
    if (!paths_args_ok(mid, csT, rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    recycle_spans(c);
    if ((rc = set_kernel_attrs(c)) != FW_OK) return rc;
    const bool paths = (mid != nullptr);
    const int npad = (n <= FW_B) ? n : (n + FW_B - 1) / FW_B * FW_B;
    const size_t tot = (size_t)npad * npad;
    if ((rc = c->w_rate.ensure(tot)) != FW_OK) return rc;
    if ((rc = c->w_next.ensure(tot)) != FW_OK) return rc;
    if (paths) {
        if ((rc = c->w_mid.ensure(tot)) != FW_OK) return rc;
        if ((rc = c->w_csT.ensure(tot)) != FW_OK) return rc;
        if ((rc = c->w_rs.ensure(tot)) != FW_OK) return rc;
    }
    CU(cudaMemcpy2DAsync(c->w_rate.p, (size_t)npad * 8, rate, (size_t)n * 8, (size_t)n * 8, n,
                         cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpy2DAsync(c->w_next.p, (size_t)npad * 4, next, (size_t)n * 4, (size_t)n * 4, n,
                         cudaMemcpyHostToDevice, c->stream));
    if ((rc = validate_device(c, c->w_rate.p, c->w_next.p, npad, 0, 1, n)) != FW_OK) return rc;
    if (paths) {
        CU(cudaMemsetAsync(c->w_mid.p, 0xFF, tot * 4, c->stream));
        CU(cudaMemsetAsync(c->w_csT.p, 0xFF, tot * 4, c->stream));
        CU(cudaMemsetAsync(c->w_rs.p, 0xFF, tot * 4, c->stream));
    }
    if (npad != n) {
        fw_fill_pad_kernel<<<grid_for((long long)tot, c->sm_count), 256, 0, c->stream>>>(
            c->w_rate.p, c->w_next.p, nullptr, nullptr, nullptr, npad, n, npad);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (n <= FW_B)
        rc = solve_tiles(c, 1, n, npad, 0, c->w_rate.p, c->w_next.p, paths ? c->w_mid.p : nullptr,
                         paths ? c->w_csT.p : nullptr, paths ? c->w_rs.p : nullptr);
    else
        rc = solve_blocked(c, npad, npad, c->w_rate.p, c->w_next.p, paths ? c->w_mid.p : nullptr,
                           paths ? c->w_csT.p : nullptr, paths ? c->w_rs.p : nullptr);
    if (rc != FW_OK) return rc;
    CU(cudaMemcpy2DAsync(rate, (size_t)n * 8, c->w_rate.p, (size_t)npad * 8, (size_t)n * 8, n,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpy2DAsync(next, (size_t)n * 4, c->w_next.p, (size_t)npad * 4, (size_t)n * 4, n,
                         cudaMemcpyDeviceToHost, c->stream));
    if (paths) {
        CU(cudaMemcpy2DAsync(mid, (size_t)n * 4, c->w_mid.p, (size_t)npad * 4, (size_t)n * 4, n,
                             cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpy2DAsync(csT, (size_t)n * 4, c->w_csT.p, (size_t)npad * 4, (size_t)n * 4, n,
                             cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpy2DAsync(rs, (size_t)n * 4, c->w_rs.p, (size_t)npad * 4, (size_t)n * 4, n,
                             cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return FW_OK;
}

int fw_solve_batched_device(fw_ctx *c, int32_t batch, int32_t n, double *d_rate, int32_t *d_next,
                            int32_t *d_mid, int32_t *d_csT, int32_t *d_rs) {
    ErrScope es0__(c);
    if (n < 0 || batch < 0) return fail(FW_ERR_INVALID, "fw_solve_batched_device: negative size");
    if (n == 0 || batch == 0) return FW_OK;
    if (!d_rate || !d_next) return fail(FW_ERR_INVALID, "fw_solve_batched_device: null buffer");
    if (!paths_args_ok(d_mid, d_csT, d_rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    recycle_spans(c);
    if ((rc = set_kernel_attrs(c)) != FW_OK) return rc;
    const long long stride = (long long)n * n;
    if ((rc = validate_device(c, d_rate, d_next, n, stride, batch, n)) != FW_OK) return rc;
    if (n <= FW_B) {
        if (d_mid) {
            CU(cudaMemsetAsync(d_mid, 0xFF, (size_t)batch * stride * 4, c->stream));
            CU(cudaMemsetAsync(d_csT, 0xFF, (size_t)batch * stride * 4, c->stream));
            CU(cudaMemsetAsync(d_rs, 0xFF, (size_t)batch * stride * 4, c->stream));
        }
        return solve_tiles(c, batch, n, n, stride, d_rate, d_next, d_mid, d_csT, d_rs);
    }
    for (int g = 0; g < batch; ++g) {
        const long long o = (long long)g * stride;
        rc = solve_device_locked(c, n, n, d_rate + o, d_next + o, d_mid ? d_mid + o : nullptr,
                                 d_csT ? d_csT + o : nullptr, d_rs ? d_rs + o : nullptr, false);
        if (rc != FW_OK) return rc;
    }
    return FW_OK;
}

int fw_solve_batched(fw_ctx *c, int32_t batch, int32_t n, double *rate, int32_t *next, int32_t *mid,
                     int32_t *csT, int32_t *rs) {
    ErrScope es0__(c);
    if (n < 0 || batch < 0) return fail(FW_ERR_INVALID, "fw_solve_batched: negative size");
    if (n == 0 || batch == 0) return FW_OK;
    if (!rate || !next) return fail(FW_ERR_INVALID, "fw_solve_batched: null buffer");
    if (!paths_args_ok(mid, csT, rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    const size_t tot = (size_t)batch * n * n;
    const bool paths = (mid != nullptr);
    FW_ENTER(c);   // held across upload, solve and download (the staging buffers belong to the context)
    {
        CU(cudaSetDevice(c->device));
        if ((rc = c->s_rate.ensure(tot)) != FW_OK) return rc;
        if ((rc = c->s_next.ensure(tot)) != FW_OK) return rc;
        if (paths) {
            if ((rc = c->s_mid.ensure(tot)) != FW_OK) return rc;
            if ((rc = c->s_csT.ensure(tot)) != FW_OK) return rc;
            if ((rc = c->s_rs.ensure(tot)) != FW_OK) return rc;
        }
        CU(cudaMemcpyAsync(c->s_rate.p, rate, tot * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->s_next.p, next, tot * 4, cudaMemcpyHostToDevice, c->stream));
    }
    rc = fw_solve_batched_device(c, batch, n, c->s_rate.p, c->s_next.p, paths ? c->s_mid.p : nullptr,
                                 paths ? c->s_csT.p : nullptr, paths ? c->s_rs.p : nullptr);
    if (rc != FW_OK) return rc;
    CU(cudaMemcpyAsync(rate, c->s_rate.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(next, c->s_next.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
    if (paths) {
        CU(cudaMemcpyAsync(mid, c->s_mid.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(csT, c->s_csT.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(rs, c->s_rs.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return FW_OK;
}

/* ---- exact `_path` expansion (Algorithms.hs:55) --------------------------- */
}  // extern "C"

namespace {

fw_path_scratch *path_scratch(fw_ctx *c) {
    if (!c->pscratch) c->pscratch = new (std::nothrow) fw_path_scratch();
    return c->pscratch;
}

fw::PathTables one_shard_tables(int n, long long ld, const int32_t *init_next, const int32_t *mid, const int32_t *csT,
                                const int32_t *rs) {
    fw::PathTables t;
    for (int i = 0; i < fw::PATH_MAXSHARD; ++i) { t.init_next[i] = init_next; t.mid[i] = mid; t.csT[i] = csT; t.rs[i] = rs; }
    t.ld = ld; t.n = n; t.cbr = n > 0 ? n : 1; t.P = 1;
    return t;
}

constexpr long long PATH_MAX_LEN = 1LL << 24;   // hard per-path bound (arbitrage cycles make paths grow exponentially)

// Paths of nq (src, dst) pairs against device tables `t` (possibly row-sharded, peers mapped).  Context lock held.
int paths_locked(fw_ctx *c, const fw::PathTables &t, int32_t nq, const int32_t *queries, int64_t *offsets,
                 int32_t *verts, int64_t cap) {
    fw_path_scratch *ps = path_scratch(c);
    if (!ps) return fail(FW_ERR_NOMEM, "out of host memory");
    int rc;
    if ((rc = ps->q.ensure(2 * (size_t)nq)) != FW_OK || (rc = ps->len.ensure(nq)) != FW_OK ||
        (rc = ps->off.ensure((size_t)nq + 1)) != FW_OK || (rc = ps->h_len.ensure((size_t)nq + 1)) != FW_OK)
        return rc;
    fw::PathArgs a;
    a.t = t; a.nq = nq; a.queries = ps->q.p; a.lengths = ps->len.p; a.offsets = ps->off.p; a.verts = nullptr;
    a.max_len = PATH_MAX_LEN; a.gstack = nullptr; a.gcap = 0; a.flag = c->d_flag;
    CU(cudaMemcpyAsync(ps->q.p, queries, sizeof(int32_t) * 2 * (size_t)nq, cudaMemcpyHostToDevice, c->stream));
    // pass 0 (lengths); a second attempt with a global overflow stack if some walk went deeper than the local one
    int q_chunk = nq;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int q0 = 0; q0 < nq; q0 += q_chunk) {
            const int qn = (nq - q0 < q_chunk) ? nq - q0 : q_chunk;
            fw::PathArgs b = a;
            b.nq = qn; b.queries = a.queries + 2 * (size_t)q0; b.lengths = a.lengths + q0;
            if (q0 == 0) CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
            fw::fw_paths_kernel<0><<<(qn + 63) / 64, 64, 0, c->stream>>>(b);
            c->launches++;
        }
        CU(cudaMemcpyAsync(ps->h_len.p, ps->len.p, sizeof(long long) * (size_t)nq, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (!(*c->h_flag & 1) || attempt == 1) break;
        // deep recursion (arbitrage): n + 2 overflow slots per walk always suffice; bound the scratch to ~256 MB
        a.gcap = (long long)t.n + 2;
        const long long per = a.gcap * 8;
        q_chunk = (int)std::max<long long>(1, std::min<long long>(nq, (256LL << 20) / per));
        if ((rc = ps->gstack.ensure((size_t)q_chunk * (size_t)a.gcap)) != FW_OK) return rc;
        a.gstack = ps->gstack.p;
    }
    if (*c->h_flag & 4) return fail(FW_ERR_INVALID, "fw_paths: query vertex out of range");
    if (*c->h_flag & 1) return fail(FW_ERR_CAP, "fw_paths: path recursion deeper than the device stack");
    long long *h_off = ps->h_len.p;      // lengths -> exclusive prefix sums, in place (back to front is not needed: use a carry)
    long long run = 0;
    offsets[0] = 0;
    for (int i = 0; i < nq; ++i) { const long long l = h_off[i]; h_off[i] = run; run += l; offsets[i + 1] = run; }
    h_off[nq] = run;
    if (*c->h_flag & 2) return fail(FW_ERR_CAP, "fw_paths: a path is longer than 2^24 hops");
    if (run > cap) return fail(FW_ERR_CAP, "fw_paths: output capacity too small (offsets[nq] holds the size needed)");
    if (run > 0) {
        if (!verts) return fail(FW_ERR_INVALID, "fw_paths: verts is null");
        if ((rc = ps->verts.ensure((size_t)run)) != FW_OK) return rc;
        a.verts = ps->verts.p;
        CU(cudaMemcpyAsync(ps->off.p, h_off, sizeof(long long) * ((size_t)nq + 1), cudaMemcpyHostToDevice, c->stream));
        for (int q0 = 0; q0 < nq; q0 += q_chunk) {
            const int qn = (nq - q0 < q_chunk) ? nq - q0 : q_chunk;
            fw::PathArgs b = a;
            b.nq = qn; b.queries = a.queries + 2 * (size_t)q0; b.offsets = a.offsets + q0;
            fw::fw_paths_kernel<1><<<(qn + 63) / 64, 64, 0, c->stream>>>(b);
            c->launches++;
        }
        CU(cudaMemcpyAsync(verts, ps->verts.p, sizeof(int32_t) * (size_t)run, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return FW_OK;
}

// optimum for one pair in one launch (fw_optimum_kernel) -> mapped pinned record.  Context lock held.
int optimum_locked(fw_ctx *c, const fw::PathTables &t, const double *const *rate_shards, long long rate_ld, int32_t src,
                   int32_t dst, double *rate, int32_t *path, int32_t cap, int32_t *path_len) {
    fw_path_scratch *ps = path_scratch(c);
    if (!ps) return fail(FW_ERR_NOMEM, "out of host memory");
    int rc;
    const int kcap = cap < 4096 ? cap : 4096;        // vertices returned through the mapped record; longer paths take fw_paths
    if ((rc = ps->h_opt.ensure(16 + 4 * (size_t)4096)) != FW_OK) return rc;
    if ((rc = ps->gstack.ensure((size_t)t.n + 2)) != FW_OK) return rc;
    fw::OptimumArgs a;
    a.t = t;
    for (int i = 0; i < fw::PATH_MAXSHARD; ++i) a.rate[i] = rate_shards[i];
    a.rate_ld = rate_ld; a.src = src; a.dst = dst; a.cap = kcap; a.max_len = PATH_MAX_LEN;
    a.gstack = ps->gstack.p; a.gcap = (long long)t.n + 2;
    void *dview = nullptr;
    CU(cudaHostGetDevicePointer(&dview, ps->h_opt.p, 0));
    a.out = static_cast<unsigned char *>(dview);
    fw::fw_optimum_kernel<<<1, 32, 0, c->stream>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    double r; long long len;
    memcpy(&r, ps->h_opt.p, 8);
    memcpy(&len, ps->h_opt.p + 8, 8);
    *rate = r;
    if (len == -1) return fail(FW_ERR_CAP, "fw_state_optimum: path recursion deeper than the device stack");
    if (len == -2) { *path_len = 0; return fail(FW_ERR_CAP, "fw_state_optimum: the path is longer than 2^24 hops"); }
    *path_len = (int32_t)len;
    if (len > cap) return fail(FW_ERR_CAP, "fw_state_optimum: path capacity too small (*path_len holds the size needed)");
    if (len <= kcap) {
        if (len > 0) memcpy(path, ps->h_opt.p + 16, 4 * (size_t)len);
        return FW_OK;
    }
    const int32_t q[2] = {src, dst};
    int64_t off[2] = {0, 0};
    return paths_locked(c, t, 1, q, off, path, cap);
}

}  // namespace

extern "C" {

int fw_paths_device(fw_ctx *c, int32_t n, int64_t ld, const int32_t *d_init_next, const int32_t *d_mid,
                    const int32_t *d_csT, const int32_t *d_rs, int32_t nq, const int32_t *queries,
                    int64_t *offsets, int32_t *verts, int64_t cap) {
    ErrScope es0__(c);
    if (n < 0 || nq < 0 || !offsets) return fail(FW_ERR_INVALID, "fw_paths: bad argument");
    offsets[0] = 0;
    if (nq == 0 || n == 0) { for (int i = 0; i < nq; ++i) offsets[i + 1] = 0; return FW_OK; }
    if (!d_init_next || !d_mid || !d_csT || !d_rs || !queries || ld < n)
        return fail(FW_ERR_INVALID, "fw_paths: null table");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    return paths_locked(c, one_shard_tables(n, ld, d_init_next, d_mid, d_csT, d_rs), nq, queries, offsets, verts, cap);
}

/* The four n x n tables of one optimised matrix kept on the device, so that `_path` thunks of the same
 * Matrix RateEntry are expanded without re-uploading them (one upload per matrix, not per entry). */
struct fw_tables {
    fw_ctx *ctx = nullptr;
    int n = 0;
    DevBuf<int32_t> t[4];
};

int fw_tables_create(fw_ctx *c, int32_t n, const int32_t *init_next, const int32_t *mid, const int32_t *csT,
                     const int32_t *rs, fw_tables **out) {
    ErrScope es0__(c);
    if (!out || n < 0) return fail(FW_ERR_INVALID, "fw_tables_create: bad argument");
    *out = nullptr;
    if (n > 0 && (!init_next || !mid || !csT || !rs)) return fail(FW_ERR_INVALID, "fw_tables_create: null table");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    if (c->closed) return fail(FW_ERR_INVALID, "fw_tables_create: the context has been destroyed");
    fw_tables *t = new (std::nothrow) fw_tables();
    if (!t) return fail(FW_ERR_NOMEM, "out of host memory");
    t->ctx = c; t->n = n;
    const size_t tot = (size_t)n * n;
    const int32_t *h[4] = {init_next, mid, csT, rs};
    for (int i = 0; i < 4 && tot > 0; ++i) {
        rc = t->t[i].ensure(tot);
        cudaError_t e = cudaSuccess;
        if (rc == FW_OK) e = cudaMemcpyAsync(t->t[i].p, h[i], tot * 4, cudaMemcpyHostToDevice, c->stream);
        if (rc == FW_OK && e != cudaSuccess) rc = cuda_fail(e, "fw_tables_create upload");
        if (rc != FW_OK) { for (int j = 0; j < 4; ++j) t->t[j].release(); delete t; return rc; }
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);   // the host tables may be freed by the caller after this returns
    if (e != cudaSuccess) { for (int j = 0; j < 4; ++j) t->t[j].release(); delete t; return cuda_fail(e, "fw_tables_create"); }
    c->refs.fetch_add(1);
    *out = t;
    return FW_OK;
}

void fw_tables_destroy(fw_tables *t) {
    if (!t) return;
    fw_ctx *c = t->ctx;
    {
        FW_ENTER(c);
        cudaSetDevice(c->device);
        if (!c->closed) cudaStreamSynchronize(c->stream);
        for (int j = 0; j < 4; ++j) t->t[j].release();
        delete t;
    }
    ctx_unref(c);
}

int fw_tables_paths(fw_tables *t, int32_t nq, const int32_t *queries, int64_t *offsets, int32_t *verts, int64_t cap) {
    if (!t || nq < 0 || !offsets) return fail(FW_ERR_INVALID, "fw_tables_paths: bad argument");
    if (t->ctx->closed) return fail(FW_ERR_INVALID, "fw_tables_paths: the context of these tables has been destroyed");
    return fw_paths_device(t->ctx, t->n, t->n, t->t[0].p, t->t[1].p, t->t[2].p, t->t[3].p, nq, queries, offsets, verts, cap);
}

int fw_paths(fw_ctx *c, int32_t n, const int32_t *init_next, const int32_t *mid, const int32_t *csT,
             const int32_t *rs, int32_t nq, const int32_t *queries, int64_t *offsets, int32_t *verts,
             int64_t cap) {
    ErrScope es0__(c);
    if (n < 0 || nq < 0 || !offsets) return fail(FW_ERR_INVALID, "fw_paths: bad argument");
    if (n == 0 || nq == 0) return fw_paths_device(c, n, n, nullptr, nullptr, nullptr, nullptr, nq, queries, offsets, verts, cap);
    if (!init_next || !mid || !csT || !rs) return fail(FW_ERR_INVALID, "fw_paths: null table");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);      // held across upload and expansion
    fw_tables *t = nullptr;
    if ((rc = fw_tables_create(c, n, init_next, mid, csT, rs, &t)) != FW_OK) return rc;
    rc = fw_tables_paths(t, nq, queries, offsets, verts, cap);
    fw_tables_destroy(t);
    return rc;
}

/* ---- device-side buildMatrix and the resident "InSync" matrix ------------------------------ */
struct fw_state {
    fw_ctx *ctx = nullptr;
    int n = 0;
    bool synced = false;
    bool want_paths = true;   // keep mid/csT/rs (needed by fw_state_optimum)
    DevBuf<double> rate;
    DevBuf<int32_t> next, init_next, mid, csT, rs, ccy, src, dst;
    DevBuf<double> val;
};

static int build_matrix_locked(fw_ctx *c, int n, long long ld, const int32_t *d_ccy, int m, const int32_t *d_src,
                               const int32_t *d_dst, const double *d_val, double *d_rate, int32_t *d_next) {
    CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
    fw_build_base_kernel<<<grid_for((long long)n * n, c->sm_count), 256, 0, c->stream>>>(d_rate, d_next, ld, n, d_ccy);
    c->launches++;
    if (m > 0) {
        fw_build_edges_kernel<<<grid_for(m, c->sm_count), 256, 0, c->stream>>>(d_rate, d_next, ld, n, d_ccy, m, d_src,
                                                                                d_dst, d_val, c->d_flag);
        c->launches++;
    }
    CU(cudaGetLastError());
    return FW_OK;
}

int fw_build_matrix_device(fw_ctx *c, int32_t n, int64_t ld, const int32_t *ccy, int32_t m, const int32_t *src,
                           const int32_t *dst, const double *val, double *d_rate, int32_t *d_next) {
    ErrScope es0__(c);
    if (n < 0 || m < 0) return fail(FW_ERR_INVALID, "fw_build_matrix_device: negative size");
    if (n == 0) return FW_OK;
    if (!ccy || !d_rate || !d_next || ld < n || (m > 0 && (!src || !dst || !val)))
        return fail(FW_ERR_INVALID, "fw_build_matrix_device: bad argument");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    DevBuf<int32_t> dccy, dsrc, ddst; DevBuf<double> dval;
    auto rel = [&]() { dccy.release(); dsrc.release(); ddst.release(); dval.release(); };
    if ((rc = dccy.ensure(n)) != FW_OK || (rc = dsrc.ensure(m > 0 ? m : 1)) != FW_OK ||
        (rc = ddst.ensure(m > 0 ? m : 1)) != FW_OK || (rc = dval.ensure(m > 0 ? m : 1)) != FW_OK) { rel(); return rc; }
    cudaError_t e = cudaMemcpyAsync(dccy.p, ccy, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(dsrc.p, src, sizeof(int32_t) * m, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(ddst.p, dst, sizeof(int32_t) * m, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(dval.p, val, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { rel(); return cuda_fail(e, "fw_build_matrix_device upload"); }
    rc = build_matrix_locked(c, n, ld, dccy.p, m, dsrc.p, ddst.p, dval.p, d_rate, d_next);
    if (rc == FW_OK) {
        e = cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "fw_build_matrix_device");
        else if (*c->h_flag & 4) rc = fail(FW_ERR_INVALID, "fw_build_matrix_device: edge endpoint out of range");
    }
    rel();
    return rc;
}

int fw_state_create(fw_ctx *c, fw_state **out) {
    ErrScope es0__(c);
    if (!out) return fail(FW_ERR_INVALID, "fw_state_create: null output");
    *out = nullptr;
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    if (c->closed) return fail(FW_ERR_INVALID, "fw_state_create: the context has been destroyed");
    fw_state *s = new (std::nothrow) fw_state();
    if (!s) return fail(FW_ERR_NOMEM, "out of host memory");
    s->ctx = c;
    c->refs.fetch_add(1);
    *out = s;
    return FW_OK;
}

void fw_state_destroy(fw_state *s) {
    if (!s) return;
    fw_ctx *c = s->ctx;
    {
        FW_ENTER(c);
        cudaSetDevice(c->device);
        if (!c->closed) cudaStreamSynchronize(c->stream);
        s->rate.release(); s->next.release(); s->init_next.release(); s->mid.release(); s->csT.release();
        s->rs.release(); s->ccy.release(); s->src.release(); s->dst.release(); s->val.release();
        delete s;
    }
    ctx_unref(c);
}

int fw_state_sync(fw_state *s, int32_t n, const int32_t *ccy, int32_t m, const int32_t *src, const int32_t *dst,
                  const double *val) {
    if (!s || n < 0 || m < 0) return fail(FW_ERR_INVALID, "fw_state_sync: bad argument");
    fw_ctx *c = s->ctx;
    FW_ENTER(c);
    if (c->closed) return fail(FW_ERR_INVALID, "fw_state_sync: the context of this state has been destroyed");
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    recycle_spans(c);
    s->synced = false;
    s->n = n;
    if (n == 0) { s->synced = true; return FW_OK; }
    if (!ccy || (m > 0 && (!src || !dst || !val))) return fail(FW_ERR_INVALID, "fw_state_sync: null input");
    int rc;
    const size_t tot = (size_t)n * n;
    if ((rc = set_kernel_attrs(c)) != FW_OK) return rc;
    const bool wp = s->want_paths;
    if ((rc = s->rate.ensure(tot)) != FW_OK || (rc = s->next.ensure(tot)) != FW_OK ||
        (wp && ((rc = s->init_next.ensure(tot)) != FW_OK || (rc = s->mid.ensure(tot)) != FW_OK ||
                (rc = s->csT.ensure(tot)) != FW_OK || (rc = s->rs.ensure(tot)) != FW_OK)) ||
        (rc = s->ccy.ensure(n)) != FW_OK || (rc = s->src.ensure(m > 0 ? m : 1)) != FW_OK ||
        (rc = s->dst.ensure(m > 0 ? m : 1)) != FW_OK || (rc = s->val.ensure(m > 0 ? m : 1)) != FW_OK)
        return rc;
    CU(cudaMemcpyAsync(s->ccy.p, ccy, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    if (m > 0) {
        CU(cudaMemcpyAsync(s->src.p, src, sizeof(int32_t) * m, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(s->dst.p, dst, sizeof(int32_t) * m, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(s->val.p, val, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = build_matrix_locked(c, n, n, s->ccy.p, m, s->src.p, s->dst.p, s->val.p, s->rate.p, s->next.p)) != FW_OK)
        return rc;
    if (wp) CU(cudaMemcpyAsync(s->init_next.p, s->next.p, tot * 4, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (*c->h_flag & 4) return fail(FW_ERR_INVALID, "fw_state_sync: edge endpoint out of range");
    if ((rc = solve_device_locked(c, n, n, s->rate.p, s->next.p, wp ? s->mid.p : nullptr, wp ? s->csT.p : nullptr,
                                  wp ? s->rs.p : nullptr, true)) != FW_OK)
        return rc;
    CU(cudaStreamSynchronize(c->stream));
    s->synced = true;
    return FW_OK;
}

int fw_state_optimum(fw_state *s, int32_t src, int32_t dst, double *rate, int32_t *path, int32_t cap,
                     int32_t *path_len) {
    if (!s || !rate || !path_len || cap < 0 || (cap > 0 && !path))
        return fail(FW_ERR_INVALID, "fw_state_optimum: bad argument");
    fw_ctx *c = s->ctx;
    FW_ENTER(c);
    if (c->closed) return fail(FW_ERR_INVALID, "fw_state_optimum: the context of this state has been destroyed");
    if (!s->synced) return fail(FW_ERR_INVALID, "fw_state_optimum: state is not in sync (call fw_state_sync)");
    if (!s->want_paths) return fail(FW_ERR_INVALID, "fw_state_optimum: state was synced without path tables");
    if (src < 0 || dst < 0 || src >= s->n || dst >= s->n) return fail(FW_ERR_INVALID, "fw_state_optimum: vertex index out of range");
    CU(cudaSetDevice(c->device));
    const double *rs[fw::PATH_MAXSHARD];
    for (int i = 0; i < fw::PATH_MAXSHARD; ++i) rs[i] = s->rate.p;
    return optimum_locked(c, one_shard_tables(s->n, s->n, s->init_next.p, s->mid.p, s->csT.p, s->rs.p), rs, s->n,
                          src, dst, rate, path, cap, path_len);
}

int fw_state_download(fw_state *s, double *rate, int32_t *next) {
    if (!s || !s->synced || s->ctx->closed) return fail(FW_ERR_INVALID, "fw_state_download: state is not in sync");
    if (s->n == 0) return FW_OK;
    fw_ctx *c = s->ctx;
    FW_ENTER(c);
    CU(cudaSetDevice(c->device));
    const size_t tot = (size_t)s->n * s->n;
    if (rate) CU(cudaMemcpyAsync(rate, s->rate.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    if (next) CU(cudaMemcpyAsync(next, s->next.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FW_OK;
}

/* floydWarshall :: Map (Vertex, Vertex) Double -> Matrix RateEntry in one call: the cache goes up
 * as COO (a few MB), buildMatrix + runAlgo run on the device, the dense result comes back. */
int fw_solve_edges(fw_ctx *c, int32_t n, const int32_t *ccy, int32_t m, const int32_t *src, const int32_t *dst,
                   const double *val, double *rate, int32_t *next, int32_t *init_next, int32_t *mid, int32_t *csT,
                   int32_t *rs) {
    ErrScope es0__(c);
    if (n < 0 || m < 0) return fail(FW_ERR_INVALID, "fw_solve_edges: negative size");
    if (n == 0) return FW_OK;
    if (!rate || !next) return fail(FW_ERR_INVALID, "fw_solve_edges: null output");
    if (!paths_args_ok(mid, csT, rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    int rc = get_ctx(c, &c);
    if (rc != FW_OK) return rc;
    FW_ENTER(c);   // one lock from the creation of the cached state to the end of the download
    if (!c->edge_state && (rc = fw_state_create(c, &c->edge_state)) != FW_OK) return rc;
    fw_state *st = c->edge_state;
    st->want_paths = (mid != nullptr) || (init_next != nullptr);
    rc = fw_state_sync(st, n, ccy, m, src, dst, val);
    if (rc != FW_OK) return rc;
    const size_t tot = (size_t)n * n;
    CU(cudaMemcpyAsync(rate, st->rate.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(next, st->next.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
    if (init_next) CU(cudaMemcpyAsync(init_next, st->init_next.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
    if (mid) {
        CU(cudaMemcpyAsync(mid, st->mid.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(csT, st->csT.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(rs, st->rs.p, tot * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return FW_OK;
}

}  // extern "C"

#include "fw_multi.cuh"

#ifdef FW_BULK_STATS
// experiment build only: read (and optionally clear) the bulk kernel's fast-path counters
extern "C" int fw_debug_bulk_stats(unsigned long long out[4], int reset) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out, fw::fw_bulk_stats, 4 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        if (cudaMemcpyToSymbol(fw::fw_bulk_stats, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
