// fw_multi -- the row-sharded multi-GPU solve behind the C ABI (include/fwgpu.h, "multi-GPU solve").
//
// Replaces, for matrices too large (or too slow) for one GPU, the same reference call as fw_solve*:
// floydWarshall = runAlgo 0 . buildMatrix (reference src/lib/Algorithms.hs:19-20), triggered by
// findBestRate.syncMatrix (src/lib/ProcessRequests.hs:82-84).  The reference's boundary is ONE in-process
// call, so the schedule (fw_plan.hpp), the two stream lanes per GPU and the pivot-panel transport live here,
// not in a Python launcher.  Included at the end of fwgpu.cu (one translation unit: shares its kernels,
// launch helpers and fw_ctx).
#pragma once
#include <cuda.h>    // types of the stream memory operations (entry points come from cudaGetDriverEntryPoint)
#include <dlfcn.h>
#include <nccl.h>   // types only; the library is dlopen'ed so that libfwgpu.so loads on boxes without NCCL

#include "fw_plan.hpp"

namespace {

// ---------------------------------------------------------------- NCCL, loaded on first use
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return FW_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) return fail(FW_ERR_CUDA, std::string("cannot load libnccl (") + dlerror() + "); use FW_MULTI_TRANSPORT=p2p in one process");
#define FW_NCCL_SYM(field, name)                                                         \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));             \
    if (!g_nccl.field) { dlclose(h); return fail(FW_ERR_CUDA, std::string("libnccl lacks ") + name); }
    FW_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    FW_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    FW_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    FW_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FW_NCCL_SYM(Broadcast, "ncclBroadcast")
    FW_NCCL_SYM(AllReduce, "ncclAllReduce")
    FW_NCCL_SYM(AllGather, "ncclAllGather")
    FW_NCCL_SYM(GroupStart, "ncclGroupStart")
    FW_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FW_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FW_NCCL_SYM
    g_nccl.h = h;
    return FW_OK;
}
int nccl_fail(ncclResult_t r, const char *what) {
    return fail(FW_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
#define NC(call)                                                     \
    do {                                                             \
        ncclResult_t r__ = (call);                                   \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);        \
    } while (0)

// ---------------------------------------------------------------- shard-side helper kernels
__device__ __forceinline__ int glob_row(int l, int cbr, int P, int r) { return ((l / cbr) * P + r) * cbr + l % cbr; }

// buildMatrix (Algorithms.hs:26-40) for the local rows of one shard; rows / columns at or beyond n are padding
__global__ void fw_build_shard_base_kernel(double *rate, int32_t *next, long long ld, int rows, int npad, int n,
                                           int cbr, int P, int r, const int32_t *ccy) {
    const long long total = (long long)rows * npad;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(e / npad), j = (int)(e - (long long)l * npad);
        const int i = glob_row(l, cbr, P, r);
        double v; int32_t x;
        if (i >= n || j >= n) { v = fw::qnan(); x = -1; }                       // padding: never replaced, never a factor
        else {
            const bool same = (i != j) && (ccy[i] == ccy[j]);                  // :33 i == j first, :34 same currency
            v = same ? 1.0 : 0.0; x = same ? j : -1;
        }
        rate[(long long)l * ld + j] = v;
        next[(long long)l * ld + j] = x;
    }
}
__global__ void fw_build_shard_edges_kernel(double *rate, int32_t *next, long long ld, int n, int cbr, int P, int r,
                                            const int32_t *ccy, int m, const int32_t *src, const int32_t *dst,
                                            const double *val, int *flag) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < m; e += gridDim.x * blockDim.x) {
        const int i = src[e], j = dst[e];
        if (i < 0 || j < 0 || i >= n || j >= n) { atomicOr(flag, 4); continue; }
        const int cb = i / cbr;
        if (cb % P != r) continue;                                             // another shard's row
        if (i == j || ccy[i] == ccy[j]) continue;                              // :33-34 come before the map lookup
        const long long off = (long long)((cb / P) * cbr + i % cbr) * ld + j;
        rate[off] = val[e];
        next[off] = j;
    }
}
// padding of a dense upload: local rows whose global row >= n, and columns >= n
__global__ void fw_pad_shard_kernel(double *rate, int32_t *next, long long ld, int rows, int npad, int n, int cbr,
                                    int P, int r) {
    const long long total = (long long)rows * npad;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(e / npad), j = (int)(e - (long long)l * npad);
        if (glob_row(l, cbr, P, r) < n && j < n) continue;
        rate[(long long)l * ld + j] = fw::qnan();
        next[(long long)l * ld + j] = -1;
    }
}

struct Shard {
    int device = 0, rank = 0;
    fw_ctx *ctx = nullptr;            // lane A = ctx->own_stream, lane B = ctx->side_stream (high priority)
    cudaStream_t sA = nullptr, sB = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr, evBc = nullptr, evStart = nullptr, evStop = nullptr;
    DevBuf<double> rate, sink, val;
    DevBuf<int32_t> next, init_next, mid, csT, rs, ccy, src, dst;
    DevBuf<double> Rw[2 * fw::BULK_MAXNB];
    ncclComm_t comm = nullptr;
    // one process per GPU, copy-engine transport: peers' panel buffers and flag words mapped through CUDA IPC
    DevBuf<int> flags;                // [0, 16): sequence number of the last panel landed in Rw[b];  [16, 16 + world): rank t's "my buffers are free" counter
    DevBuf<unsigned char> ipc_stage;
    struct Peer { void *Rw[2 * fw::BULK_MAXNB]; int *flags; };
    std::vector<Peer> peers;
};

constexpr int FLAG_ARRIVE = 0, FLAG_FREE = 2 * fw::BULK_MAXNB;

// the flag words are written by a peer over NVLink after its copy has completed in stream order
__global__ void fw_signal_kernel(int *remote, int v) {
    __threadfence_system();
    *reinterpret_cast<volatile int *>(remote) = v;
    __threadfence_system();
}
// fallback wait when the driver offers no stream memory operations: one thread polls a local word
__global__ void fw_spin_wait_kernel(const int *flag, int v) {
    while ((int)(*reinterpret_cast<const volatile int *>(flag) - v) < 0) __nanosleep(200);
    __threadfence_system();
}

typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
std::atomic<StreamWaitValue32Fn> g_wait_value32{nullptr};
std::once_flag g_wait_value32_once;

int stream_wait_geq(cudaStream_t st, const int *flag, int v) {
    std::call_once(g_wait_value32_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess && !getenv("FW_MULTI_SPINWAIT"))
            g_wait_value32.store(reinterpret_cast<StreamWaitValue32Fn>(fn));
        cudaGetLastError();
    });
    if (StreamWaitValue32Fn wait = g_wait_value32.load()) {
        CUresult r = wait((CUstream)st, (CUdeviceptr)(uintptr_t)flag, (cuuint32_t)v, CU_STREAM_WAIT_VALUE_GEQ);
        if (r == CUDA_SUCCESS) return FW_OK;
        g_wait_value32.store(nullptr);  // not supported on this stream / device: poll instead
    }
    fw_spin_wait_kernel<<<1, 1, 0, st>>>(flag, v);
    CU(cudaGetLastError());
    return FW_OK;
}

}  // namespace

struct fw_multi {
    int world = 1;
    bool rank_mode = false;
    bool use_nccl = false;            // panels by ncclBroadcast
    bool use_ipc = false;             // one process per GPU: panels by copy-engine copies into IPC-mapped peer buffers
    bool ipc_ready = false;
    int bseq = 0, gseq = 0;           // sequence numbers of broadcasts / factored groups (identical on every rank)
    std::vector<Shard> sh;            // local shards (single process: all `world`; rank mode: one)
    std::recursive_mutex mu;
    std::mutex err_mu;
    std::string err, err_ret;
    // current problem
    int n = 0;
    fwplan::Layout L;
    int n_edges = 0;
    bool allocated = false, want_paths = false, solved = false, record_snap = false, profiling = false;
    bool coo_resident = false;        // the COO of the last fw_multi_sync is still on the devices (fw_multi_resolve)
    double last_ms = 0.0;
    int64_t last_launches = 0;
};

namespace {

thread_local fw_multi *g_err_multi = nullptr;
struct MultiScope {                   // errors raised while a fw_multi call runs are kept on the object too
    fw_multi *m, *prev;
    explicit MultiScope(fw_multi *m_) : m(m_), prev(g_err_multi) { g_err_multi = m_; }
    ~MultiScope() {
        if (m && !g_err.empty()) { std::lock_guard<std::mutex> lk(m->err_mu); m->err = g_err; }
        g_err_multi = prev;
    }
};
#define FW_MENTER(m) g_err.clear(); MultiScope ms__(m); std::lock_guard<std::recursive_mutex> mlk__((m)->mu)

int multi_sync_all(fw_multi *m) {
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        CU(cudaStreamSynchronize(s.sA));
        CU(cudaStreamSynchronize(s.sB));
    }
    return FW_OK;
}

void shard_release(Shard &s) {
    cudaSetDevice(s.device);
    if (s.sA) cudaStreamSynchronize(s.sA);
    if (s.sB) cudaStreamSynchronize(s.sB);
    for (auto &pr : s.peers) {
        for (void *q : pr.Rw) if (q) cudaIpcCloseMemHandle(q);
        if (pr.flags) cudaIpcCloseMemHandle(pr.flags);
    }
    s.peers.clear();
    if (s.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.comm);
    s.flags.release(); s.ipc_stage.release();
    s.rate.release(); s.sink.release(); s.val.release(); s.next.release(); s.init_next.release();
    s.mid.release(); s.csT.release(); s.rs.release(); s.ccy.release(); s.src.release(); s.dst.release();
    for (auto &b : s.Rw) b.release();
    for (cudaEvent_t e : {s.evA, s.evB, s.evBc, s.evStart, s.evStop}) if (e) cudaEventDestroy(e);
    if (s.ctx) fw_ctx_destroy(s.ctx);
    s = Shard();
}

int shard_init(Shard &s, int device, int rank) {
    s.device = device; s.rank = rank;
    int rc = fw_ctx_create(device, &s.ctx);
    if (rc != FW_OK) return rc;
    CU(cudaSetDevice(device));
    s.sA = s.ctx->own_stream;
    s.sB = s.ctx->side_stream ? s.ctx->side_stream : s.ctx->own_stream;
    CU(cudaEventCreateWithFlags(&s.evA, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.evB, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.evBc, cudaEventDisableTiming));
    CU(cudaEventCreate(&s.evStart));
    CU(cudaEventCreate(&s.evStop));
    return set_kernel_attrs(s.ctx);
}

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

// Layout for order n on `world` ranks: k-blocks per group by size (the single-GPU policy: tile loads saved vs
// extra strip launches), cyclic blocks of one group unless FW_MULTI_CYCLIC=0, n padded to whole cyclic rounds.
fwplan::Layout choose_layout(int n, int world) {
    fwplan::Layout L;
    L.world = world; L.B = FW_B;
    const int nblk = (n + FW_B - 1) / FW_B;
    int G = nblk >= 128 ? 8 : (nblk >= 48 ? 4 : (nblk >= 4 * world && nblk >= 8 ? 2 : 1));
    const int forced = env_int("FW_MULTI_GROUP", 0);
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) G = forced;
    L.G = G;
    const int unit = G * FW_B * world;
    L.n = (n + unit - 1) / unit * unit;
    L.cbr = env_int("FW_MULTI_CYCLIC", 1) != 0 ? G * FW_B : L.n / world;
    return L;
}

// One process per GPU: every rank publishes IPC handles of its 2G panel buffers and of its flag words, gathers
// everybody's through NCCL and maps them.  Collective; called whenever the layout (hence the buffers) changed.
int multi_ipc_setup(fw_multi *m) {
    Shard &s = m->sh[0];
    const int P = m->world, NB = 2 * fw::BULK_MAXNB, nh = NB + 1;
    const size_t hb = sizeof(cudaIpcMemHandle_t), mine = nh * hb;
    CU(cudaSetDevice(s.device));
    CU(cudaStreamSynchronize(s.sA));
    CU(cudaStreamSynchronize(s.sB));
    for (auto &pr : s.peers) {
        for (void *&q : pr.Rw) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
        if (pr.flags) { cudaIpcCloseMemHandle(pr.flags); pr.flags = nullptr; }
    }
    s.peers.assign(P, Shard::Peer());
    for (auto &pr : s.peers) { for (void *&q : pr.Rw) q = nullptr; pr.flags = nullptr; }
    int rc;
    if (!s.flags.p) {
        if ((rc = s.flags.ensure(FLAG_FREE + P)) != FW_OK) return rc;
        CU(cudaMemsetAsync(s.flags.p, 0, sizeof(int) * (FLAG_FREE + P), s.sA));
    }
    if ((rc = s.ipc_stage.ensure(mine * (P + 1))) != FW_OK) return rc;
    std::vector<cudaIpcMemHandle_t> hs(nh * (size_t)(P + 1));
    memset(hs.data(), 0, hs.size() * hb);
    for (int b = 0; b < NB; ++b)
        if (s.Rw[b].p) CU(cudaIpcGetMemHandle(&hs[b], s.Rw[b].p));
    CU(cudaIpcGetMemHandle(&hs[NB], s.flags.p));
    CU(cudaMemcpyAsync(s.ipc_stage.p, hs.data(), mine, cudaMemcpyHostToDevice, s.sA));
    NC(g_nccl.AllGather(s.ipc_stage.p, s.ipc_stage.p + mine, mine, ncclInt8, s.comm, s.sA));
    CU(cudaMemcpyAsync(hs.data() + nh, s.ipc_stage.p + mine, mine * P, cudaMemcpyDeviceToHost, s.sA));
    CU(cudaStreamSynchronize(s.sA));
    for (int t = 0; t < P; ++t) {
        if (t == s.rank) continue;
        const cudaIpcMemHandle_t *th = &hs[nh * (size_t)(1 + t)];
        for (int b = 0; b < 2 * m->L.G; ++b)
            CU(cudaIpcOpenMemHandle(&s.peers[t].Rw[b], th[b], cudaIpcMemLazyEnablePeerAccess));
        void *fp = nullptr;
        CU(cudaIpcOpenMemHandle(&fp, th[NB], cudaIpcMemLazyEnablePeerAccess));
        s.peers[t].flags = static_cast<int *>(fp);
    }
    m->ipc_ready = true;
    return FW_OK;
}

int multi_alloc(fw_multi *m, int n, bool want_paths) {
    if (n <= 0) return fail(FW_ERR_INVALID, "fw_multi: n must be positive");
    const fwplan::Layout L = choose_layout(n, m->world);
    if (!L.valid()) return fail(FW_ERR_INVALID, "fw_multi: no valid layout for this n / world");
    const size_t rows = (size_t)L.rows_local(), tot = rows * (size_t)L.n;
    int rc;
    if (m->use_ipc && m->ipc_ready) {
        // A panel buffer that peers have mapped must not be freed under them: if this layout needs bigger buffers
        // (the same decision on every rank: the ranks share their allocation history), every rank first closes
        // its mappings and the ranks meet before anybody reallocates.
        Shard &s = m->sh[0];
        bool grow = false;
        for (int b = 0; b < 2 * L.G; ++b) grow |= s.Rw[b].cap < (size_t)FW_B * L.n;
        if (grow) {
            CU(cudaSetDevice(s.device));
            CU(cudaStreamSynchronize(s.sA));
            CU(cudaStreamSynchronize(s.sB));
            for (auto &pr : s.peers) {
                for (void *&q : pr.Rw) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
                if (pr.flags) { cudaIpcCloseMemHandle(pr.flags); pr.flags = nullptr; }
            }
            m->ipc_ready = false;
            NC(g_nccl.AllReduce(s.ctx->d_flag, s.ctx->d_flag, 1, ncclInt32, ncclMax, s.comm, s.sA));   // rendezvous
            CU(cudaStreamSynchronize(s.sA));
        }
    }
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        if ((rc = s.rate.ensure(tot)) != FW_OK || (rc = s.next.ensure(tot)) != FW_OK) return rc;
        if (want_paths && ((rc = s.init_next.ensure(tot)) != FW_OK || (rc = s.mid.ensure(tot)) != FW_OK ||
                           (rc = s.csT.ensure(tot)) != FW_OK || (rc = s.rs.ensure(tot)) != FW_OK))
            return rc;
        if (m->record_snap && (rc = s.sink.ensure(tot)) != FW_OK) return rc;
        for (int b = 0; b < 2 * L.G; ++b)
            if ((rc = s.Rw[b].ensure((size_t)FW_B * L.n)) != FW_OK) return rc;
        for (int set = 0; set < L.G; ++set)
            if ((rc = s.ctx->Cp[set].ensure(rows * FW_B)) != FW_OK || (rc = s.ctx->NCp[set].ensure(rows * FW_B)) != FW_OK)
                return rc;
    }
    const bool same_layout = m->allocated && m->L.n == L.n && m->L.G == L.G && m->L.cbr == L.cbr;
    if (!(same_layout && m->n == n)) m->coo_resident = false;
    m->n = n; m->L = L; m->want_paths = want_paths; m->allocated = true; m->solved = false;
    if (m->use_ipc && !(same_layout && m->ipc_ready) && (rc = multi_ipc_setup(m)) != FW_OK) return rc;
    return FW_OK;
}

// ---- executor of one plan operation on one local shard ------------------------------------------------
struct ShardView {
    Shard *s;
    const fwplan::Layout *L;
    bool paths;
    int rows() const { return L->rows_local(); }
};

int exec_pivot(const ShardView &v, const fw_plan_op &op, bool record_snap) {
    Shard &s = *v.s;
    fw_ctx *c = s.ctx;
    const int npad = v.L->n, rows = v.rows(), lr = op.row_lo, set = op.buf % v.L->G;
    const long long ld = npad;
    fw::TileArgs t;
    t.rate = s.rate.p; t.next = s.next.p;
    t.mid = v.paths ? s.mid.p : nullptr; t.csT = v.paths ? s.csT.p : nullptr; t.rs = v.paths ? s.rs.p : nullptr;
    t.ld = ld; t.batch_stride = 0; t.b0 = op.b0; t.r0 = lr; t.nv = FW_B;
    t.Cp = c->Cp[set].p; t.ldc = rows; t.NCp = c->NCp[set].p; t.Rw = s.Rw[op.buf].p; t.ldw = npad;
    {
        PhaseTimer pt(c, 0);
        if (v.paths) fw::fw_tile_kernel<true><<<1, 512, fw::tile_smem_bytes(true), c->cur>>>(t);
        else         fw::fw_tile_kernel<false><<<1, 512, fw::tile_smem_bytes(false), c->cur>>>(t);
    }
    c->launches++;
    if (npad > FW_B) {
        fw::PanelArgs p;
        p.rate = t.rate; p.next = t.next; p.mid = t.mid; p.csT = t.csT; p.rs = t.rs;
        p.ld = ld; p.npad = npad; p.b0 = op.b0; p.rows = rows; p.blk_r0 = lr; p.skip_r0 = lr; p.skipn = FW_B;
        p.Cp = c->Cp[set].p; p.ldc = rows; p.NCp = c->NCp[set].p; p.Rw = s.Rw[op.buf].p; p.ldw = npad;
        PhaseTimer pt(c, 2);
        launch_panel<false>(c, p, npad - FW_B, v.paths, c->cur);
    }
    CU(cudaGetLastError());
    if (record_snap)
        CU(cudaMemcpyAsync(s.sink.p + (long long)lr * ld, s.Rw[op.buf].p, sizeof(double) * (size_t)FW_B * npad,
                           cudaMemcpyDeviceToDevice, c->cur));
    return FW_OK;
}

// merges two local row ranges that must be adjacent (or one empty) into one skip range
bool merge_skip(int a0, int an, int b0, int bn, int &s0, int &sn) {
    if (an <= 0) { s0 = bn > 0 ? b0 : NOSKIP; sn = bn > 0 ? bn : 0; return true; }
    if (bn <= 0) { s0 = a0; sn = an; return true; }
    if (a0 + an == b0) { s0 = a0; sn = an + bn; return true; }
    if (b0 + bn == a0) { s0 = b0; sn = an + bn; return true; }
    return false;
}

int exec_apply(const ShardView &v, const fw_plan_op &op) {
    Shard &s = *v.s;
    fw_ctx *c = s.ctx;
    const fwplan::Layout &L = *v.L;
    const int npad = L.n, rows = v.rows(), nb = op.nb, b0 = op.b0;
    const long long ld = npad;
    const int v0 = op.row_lo, vrows = op.row_n;
    if (vrows <= 0 || npad <= FW_B) return FW_OK;
    // the blocks' own rows inside this view (local rows [g0, g0 + nb*B)), if this shard holds them
    const bool gin = op.grp_lo >= 0 && op.grp_lo >= v0 && op.grp_lo + nb * FW_B <= v0 + vrows;
    if (op.grp_lo >= 0 && !gin && op.grp_lo < v0 + vrows && v0 < op.grp_lo + nb * FW_B)
        return fail(FW_ERR_INVALID, "fw_multi: APPLY rows cut through the k-blocks' own rows");
    const int g0 = gin ? op.grp_lo : 0;
    const int ex0 = op.ex_n > 0 ? op.ex_lo : 0, exn = op.ex_n > 0 ? op.ex_n : 0;
    double *rate_v = s.rate.p + (long long)v0 * ld;
    int32_t *next_v = s.next.p + (long long)v0 * ld;
    int32_t *mid_v = v.paths ? s.mid.p + (long long)v0 * ld : nullptr;
    int32_t *csT_v = v.paths ? s.csT.p + (long long)v0 * ld : nullptr;
    fw::BulkArgs g;
    g.rate = rate_v; g.next = next_v; g.mid = mid_v; g.ld = ld; g.b0 = b0;
    g.row0 = v0; g.cbr = L.cbr; g.P = L.world; g.r = s.rank;
    g.ldc = rows; g.ldw = npad;
    for (int i = 0; i < fw::BULK_MAXNB; ++i) {
        const int set = i < nb ? i : nb - 1;
        g.CpT[i] = c->Cp[set].p + v0;                         // CpT[kk*ldc + i]
        g.NCp[i] = c->NCp[set].p + (long long)v0 * FW_B;      // NCp[i*B + kk]
        g.Rw[i] = s.Rw[op.buf + set].p;
    }
    g.half_r0 = gin ? (g0 - v0) / 64 : NOSKIP;                // rows of the group's block i start at block i+1
    auto rows_minus = [&](int tail_lo, int &s0, int &sn) -> bool {   // skip = [tail_lo, end of own rows) + excluded rows
        const int t0 = gin ? tail_lo : 0, tn = gin ? g0 + nb * FW_B - tail_lo : 0;
        return merge_skip(t0, tn, ex0, exn, s0, sn);
    };
    for (int blk = 0; blk < nb; ++blk) {
        int s0, sn;
        if (blk > 0) {
            // blocks 0 .. blk-1 on the column strip of block blk, so that its column panel can run
            if (!rows_minus(g0 + (blk - 1) * FW_B, s0, sn)) return fail(FW_ERR_INVALID, "fw_multi: APPLY skip ranges are not adjacent");
            const int out = vrows - sn;
            if (out > 0) {
                g.nb = blk; g.half_c0 = NOSKIP;
                g.row_lo = 0; g.rskip0 = (s0 == NOSKIP) ? NOSKIP : (s0 - v0) / 64; g.rskipn = sn / 64;
                g.col_lo = (b0 + blk * FW_B) / 64; g.cskip0 = NOSKIP; g.cskipn = 0;
                launch_bulk(c, g, 2, out / 64);
            }
        }
        if (!rows_minus(g0 + blk * FW_B, s0, sn)) return fail(FW_ERR_INVALID, "fw_multi: APPLY skip ranges are not adjacent");
        const int out = vrows - sn;
        if (out > 0) {
            fw::PanelArgs p;
            p.rate = rate_v; p.next = next_v; p.mid = mid_v; p.csT = csT_v; p.rs = nullptr;
            p.ld = ld; p.npad = npad; p.b0 = b0 + blk * FW_B; p.rows = vrows; p.blk_r0 = NOSKIP;
            p.skip_r0 = (s0 == NOSKIP) ? NOSKIP : s0 - v0; p.skipn = sn;
            p.Cp = const_cast<double *>(g.CpT[blk]); p.ldc = rows; p.NCp = const_cast<int32_t *>(g.NCp[blk]);
            p.Rw = s.Rw[op.buf + blk].p; p.ldw = npad;
            PhaseTimer pt(c, 1);
            launch_panel<true>(c, p, out, v.paths, c->cur);
        }
    }
    {
        // all nb blocks for every other tile from one load of the tile
        int s0, sn;
        if (!rows_minus(g0 + (nb - 1) * FW_B, s0, sn)) return fail(FW_ERR_INVALID, "fw_multi: APPLY skip ranges are not adjacent");
        const int out = vrows - sn;
        if (out > 0) {
            g.nb = nb; g.half_c0 = (nb > 1) ? b0 / 64 : NOSKIP;
            g.row_lo = 0; g.rskip0 = (s0 == NOSKIP) ? NOSKIP : (s0 - v0) / 64; g.rskipn = sn / 64;
            g.col_lo = 0; g.cskip0 = (b0 + (nb - 1) * FW_B) / 64; g.cskipn = 2;
            launch_bulk(c, g, npad / 64 - 2, out / 64);
        }
    }
    CU(cudaGetLastError());
    return FW_OK;
}

int exec_bcast(fw_multi *m, const fw_plan_op &op) {
    const size_t count = (size_t)FW_B * m->L.n;
    if (m->use_nccl) {
        if (m->sh.size() > 1) NC(g_nccl.GroupStart());
        for (auto &s : m->sh) {
            if (m->sh.size() == 1) CU(cudaSetDevice(s.device));
            NC(g_nccl.Broadcast(s.Rw[op.buf].p, s.Rw[op.buf].p, count, ncclDouble, op.rank, s.comm, s.sB));
        }
        if (m->sh.size() > 1) NC(g_nccl.GroupEnd());
        return FW_OK;
    }
    if (m->use_ipc) {
        // one process per GPU, copy engines: flag words in device memory order the ranks (no host round trips).
        // First panel of a group: every other rank tells the owner "my buffers of this parity are free" (its
        // look-ahead lane has waited for its main lane by then); the owner waits for all of them, pushes the
        // panel into every peer's buffer and raises the peer's arrival word; the peers' look-ahead lanes wait on it.
        Shard &s = m->sh[0];
        CU(cudaSetDevice(s.device));
        const int P = m->world;
        const bool first = (op.buf % m->L.G) == 0;
        if (first) ++m->gseq;
        ++m->bseq;
        int rc;
        if (s.rank == op.rank) {
            if (first)
                for (int t = 0; t < P; ++t)
                    if (t != s.rank && (rc = stream_wait_geq(s.sB, s.flags.p + FLAG_FREE + t, m->gseq)) != FW_OK) return rc;
            for (int t = 0; t < P; ++t)
                if (t != s.rank)
                    CU(cudaMemcpyAsync(s.peers[t].Rw[op.buf], s.Rw[op.buf].p, count * 8, cudaMemcpyDeviceToDevice, s.sB));
            for (int t = 0; t < P; ++t)
                if (t != s.rank) fw_signal_kernel<<<1, 1, 0, s.sB>>>(s.peers[t].flags + FLAG_ARRIVE + op.buf, m->bseq);
            CU(cudaGetLastError());
        } else {
            if (first) {
                fw_signal_kernel<<<1, 1, 0, s.sB>>>(s.peers[op.rank].flags + FLAG_FREE + s.rank, m->gseq);
                CU(cudaGetLastError());
            }
            if ((rc = stream_wait_geq(s.sB, s.flags.p + FLAG_ARRIVE + op.buf, m->bseq)) != FW_OK) return rc;
        }
        return FW_OK;
    }
    // copy-engine transport (single process): the owner's look-ahead lane pushes the panel into every peer's
    // buffer once the peer's lanes are done with what that buffer held, then the peers' look-ahead lanes wait
    Shard &o = m->sh[op.rank];
    CU(cudaSetDevice(o.device));
    for (auto &t : m->sh) {
        if (t.rank == o.rank) continue;
        CU(cudaStreamWaitEvent(o.sB, t.evA, 0));
        CU(cudaStreamWaitEvent(o.sB, t.evB, 0));
    }
    for (auto &t : m->sh) {
        if (t.rank == o.rank) continue;
        if (t.device == o.device)
            CU(cudaMemcpyAsync(t.Rw[op.buf].p, o.Rw[op.buf].p, count * 8, cudaMemcpyDeviceToDevice, o.sB));
        else
            CU(cudaMemcpyPeerAsync(t.Rw[op.buf].p, t.device, o.Rw[op.buf].p, o.device, count * 8, o.sB));
    }
    CU(cudaEventRecord(o.evBc, o.sB));
    for (auto &t : m->sh) {
        if (t.rank == o.rank) continue;
        CU(cudaSetDevice(t.device));
        CU(cudaStreamWaitEvent(t.sB, o.evBc, 0));
    }
    return FW_OK;
}

Shard *local_shard(fw_multi *m, int rank) {
    if (!m->rank_mode) return (rank >= 0 && rank < (int)m->sh.size()) ? &m->sh[rank] : nullptr;
    return (m->sh[0].rank == rank) ? &m->sh[0] : nullptr;
}

int multi_validate(fw_multi *m) {
    const fwplan::Layout &L = m->L;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        s.ctx->cur = s.sA; s.ctx->stream = s.sA;
        int rc = validate_device(s.ctx, s.rate.p, s.next.p, L.n, 0, 1, L.n, L.rows_local(), L.cbr, L.world, s.rank, m->n, false);
        if (rc != FW_OK) return rc;
    }
    int flags = 0;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        CU(cudaStreamSynchronize(s.sA));
        flags |= *s.ctx->h_flag;
    }
    if (m->rank_mode && m->world > 1) {
        // every rank must take the same decision, or the ranks that go on would wait in the first broadcast
        Shard &s = m->sh[0];
        *s.ctx->h_flag = flags;
        CU(cudaMemcpyAsync(s.ctx->d_flag, s.ctx->h_flag, sizeof(int), cudaMemcpyHostToDevice, s.sA));
        NC(g_nccl.AllReduce(s.ctx->d_flag, s.ctx->d_flag, 1, ncclInt32, ncclMax, s.comm, s.sA));   // flags are 0..3: max keeps "some error"
        CU(cudaMemcpyAsync(s.ctx->h_flag, s.ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s.sA));
        CU(cudaStreamSynchronize(s.sA));
        flags = *s.ctx->h_flag;
    }
    if (flags & 1) return fail(FW_ERR_DOMAIN, "rate matrix holds a negative entry");
    if (flags & 2) return fail(FW_ERR_DOMAIN, "rate > 0 with next < 0 (inconsistent next-hop matrix)");
    return FW_OK;
}

int multi_solve_resident(fw_multi *m, bool clock_started = false) {
    if (!m->allocated) return fail(FW_ERR_INVALID, "fw_multi: no matrix allocated (fw_multi_alloc / fw_multi_sync)");
    const fwplan::Layout &L = m->L;
    int rc;
    m->solved = false;
    for (auto &s : m->sh) {
        s.ctx->launches = 0;
        s.ctx->profiling = m->profiling;
        recycle_spans(s.ctx);
    }
    if ((rc = multi_validate(m)) != FW_OK) return rc;
    const size_t tot = (size_t)L.rows_local() * L.n;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        if (m->record_snap && (rc = s.sink.ensure(tot)) != FW_OK) return rc;
        if (!clock_started) CU(cudaEventRecord(s.evStart, s.sA));
        if (m->want_paths) {
            CU(cudaMemsetAsync(s.mid.p, 0xFF, tot * 4, s.sA));
            CU(cudaMemsetAsync(s.csT.p, 0xFF, tot * 4, s.sA));
            CU(cudaMemsetAsync(s.rs.p, 0xFF, tot * 4, s.sA));
        }
        CU(cudaEventRecord(s.evA, s.sA));
        CU(cudaStreamWaitEvent(s.sB, s.evA, 0));       // the look-ahead lane starts after the reset
    }
    const std::vector<fw_plan_op> plan = fwplan::make_plan(L);
    for (const fw_plan_op &op : plan) {
        if (op.kind == FW_OP_BCAST) {
            if ((rc = exec_bcast(m, op)) != FW_OK) return rc;
            continue;
        }
        Shard *s = local_shard(m, op.rank);
        if (!s) continue;                               // another process's operation
        CU(cudaSetDevice(s->device));
        fw_ctx *c = s->ctx;
        c->cur = op.lane ? s->sB : s->sA;
        c->stream = c->cur;
        ShardView v{s, &L, m->want_paths};
        switch (op.kind) {
            case FW_OP_PIVOT: rc = exec_pivot(v, op, m->record_snap); break;
            case FW_OP_APPLY: rc = exec_apply(v, op); break;
            case FW_OP_A_DONE: CU(cudaEventRecord(s->evA, s->sA)); rc = FW_OK; break;
            case FW_OP_WAIT_A: CU(cudaStreamWaitEvent(s->sB, s->evA, 0)); rc = FW_OK; break;
            case FW_OP_B_DONE: CU(cudaEventRecord(s->evB, s->sB)); rc = FW_OK; break;
            case FW_OP_WAIT_B: CU(cudaStreamWaitEvent(s->sA, s->evB, 0)); rc = FW_OK; break;
            default: rc = fail(FW_ERR_INVALID, "fw_multi: unknown plan operation");
        }
        if (rc != FW_OK) return rc;
    }
    m->last_launches = 0;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        CU(cudaEventRecord(s.evB, s.sB));
        CU(cudaStreamWaitEvent(s.sA, s.evB, 0));
        CU(cudaEventRecord(s.evStop, s.sA));
        s.ctx->cur = nullptr; s.ctx->stream = s.sA;
        m->last_launches += s.ctx->launches;
    }
    if ((rc = multi_sync_all(m)) != FW_OK) return rc;
    m->last_ms = 0.0;
    for (auto &s : m->sh) {
        float t = 0.f;
        CU(cudaSetDevice(s.device));
        CU(cudaEventElapsedTime(&t, s.evStart, s.evStop));
        if (t > m->last_ms) m->last_ms = t;
    }
    m->solved = true;
    return FW_OK;
}

// Host rows [row0, row0+rows) <-> the shards that hold them, one 2-D copy per cyclic block piece.
// dir: 0 host -> device, 1 device -> host.  `dev_of` picks the shard buffer (nullptr result: skip).
template <typename T, typename Pick>
int copy_rows(fw_multi *m, int row0, int rows, T *host, int dir, Pick pick) {
    const fwplan::Layout &L = m->L;
    const int n = m->n;
    int g = row0;
    while (g < row0 + rows) {
        const int piece = std::min(row0 + rows - g, L.cbr - g % L.cbr);
        Shard *s = local_shard(m, L.owner_of_row(g));
        if (s) {
            T *dev = pick(*s);
            if (dev) {
                CU(cudaSetDevice(s->device));
                T *d = dev + (long long)L.local_of_row(g) * L.n;
                T *h = host + (long long)(g - row0) * n;
                if (dir == 0)
                    CU(cudaMemcpy2DAsync(d, (size_t)L.n * sizeof(T), h, (size_t)n * sizeof(T), (size_t)n * sizeof(T), piece,
                                         cudaMemcpyHostToDevice, s->sA));
                else
                    CU(cudaMemcpy2DAsync(h, (size_t)n * sizeof(T), d, (size_t)L.n * sizeof(T), (size_t)n * sizeof(T), piece,
                                         cudaMemcpyDeviceToHost, s->sA));
            }
        }
        g += piece;
    }
    return FW_OK;
}

// COO of the cache -> every shard's device memory (each shard filters the rows it holds)
int multi_upload_coo(fw_multi *m, int n, const int32_t *ccy, int ne, const int32_t *src, const int32_t *dst, const double *val) {
    int rc;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        if ((rc = s.ccy.ensure(n)) != FW_OK || (rc = s.src.ensure(ne > 0 ? ne : 1)) != FW_OK ||
            (rc = s.dst.ensure(ne > 0 ? ne : 1)) != FW_OK || (rc = s.val.ensure(ne > 0 ? ne : 1)) != FW_OK)
            return rc;
        CU(cudaMemcpyAsync(s.ccy.p, ccy, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s.sA));
        if (ne > 0) {
            CU(cudaMemcpyAsync(s.src.p, src, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, s.sA));
            CU(cudaMemcpyAsync(s.dst.p, dst, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, s.sA));
            CU(cudaMemcpyAsync(s.val.p, val, sizeof(double) * ne, cudaMemcpyHostToDevice, s.sA));
        }
    }
    m->n_edges = ne;
    return FW_OK;
}

// buildMatrix (Algorithms.hs:26-40) per shard from the COO already on the device; starts the solve's clock
int multi_build(fw_multi *m) {
    const fwplan::Layout &L = m->L;
    const int n = m->n, ne = m->n_edges;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        fw_ctx *c = s.ctx;
        CU(cudaEventRecord(s.evStart, s.sA));
        CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int), s.sA));
        const long long tot = (long long)L.rows_local() * L.n;
        fw_build_shard_base_kernel<<<grid_for(tot, c->sm_count), 256, 0, s.sA>>>(s.rate.p, s.next.p, L.n, L.rows_local(), L.n, n,
                                                                                  L.cbr, L.world, s.rank, s.ccy.p);
        if (ne > 0)
            fw_build_shard_edges_kernel<<<grid_for(ne, c->sm_count), 256, 0, s.sA>>>(s.rate.p, s.next.p, L.n, n, L.cbr, L.world,
                                                                                      s.rank, s.ccy.p, ne, s.src.p, s.dst.p,
                                                                                      s.val.p, c->d_flag);
        CU(cudaGetLastError());
        if (m->want_paths) CU(cudaMemcpyAsync(s.init_next.p, s.next.p, (size_t)tot * 4, cudaMemcpyDeviceToDevice, s.sA));
        CU(cudaMemcpyAsync(c->h_flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s.sA));
    }
    int flags = 0;
    for (auto &s : m->sh) {
        CU(cudaSetDevice(s.device));
        CU(cudaStreamSynchronize(s.sA));
        flags |= *s.ctx->h_flag;
    }
    if (flags & 4) return fail(FW_ERR_INVALID, "fw_multi_sync: edge endpoint out of range");
    m->coo_resident = true;
    return FW_OK;
}

int multi_common_create(fw_multi *m) {
    // transport: NCCL in rank mode; in one process copy-engine peer copies unless FW_MULTI_TRANSPORT=nccl
    const char *tr = getenv("FW_MULTI_TRANSPORT");
    bool want_nccl = m->rank_mode ? (m->world > 1 && tr && std::string(tr) == "nccl")
                                  : (tr && std::string(tr) == "nccl" && m->world > 1);
    m->use_ipc = m->rank_mode && m->world > 1 && !want_nccl;
    if (!m->rank_mode && m->world > 1) {
        // peer access for the copies and for fw_multi_optimum's cross-shard table walk
        bool all_peers = true;
        for (auto &a : m->sh)
            for (auto &b : m->sh) {
                if (a.device == b.device) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a.device, b.device);
                if (!can) { all_peers = false; continue; }
                cudaSetDevice(a.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) all_peers = false;
                cudaGetLastError();
            }
        bool dup = false;
        for (size_t i = 0; i < m->sh.size(); ++i)
            for (size_t j = i + 1; j < m->sh.size(); ++j) dup |= (m->sh[i].device == m->sh[j].device);
        if (!all_peers && !dup && !(tr && std::string(tr) == "p2p")) want_nccl = true;
        if (dup && want_nccl) return fail(FW_ERR_INVALID, "fw_multi_create: NCCL cannot put two ranks on one device (use p2p)");
    }
    m->use_nccl = want_nccl;
    return FW_OK;
}

}  // namespace

extern "C" {

int fw_multi_unique_id(void *id128) {
    if (!id128) return fail(FW_ERR_INVALID, "fw_multi_unique_id: null output");
    int rc = load_nccl();
    if (rc != FW_OK) return rc;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return FW_OK;
}

int fw_multi_create(int32_t ndev, const int32_t *devices, fw_multi **out) {
    if (!out || ndev < 1 || ndev > 64) return fail(FW_ERR_INVALID, "fw_multi_create: bad argument");
    *out = nullptr;
    fw_multi *m = new (std::nothrow) fw_multi();
    if (!m) return fail(FW_ERR_NOMEM, "out of host memory");
    m->world = ndev; m->rank_mode = false;
    m->sh.resize(ndev);
    int rc = FW_OK;
    for (int i = 0; i < ndev && rc == FW_OK; ++i) rc = shard_init(m->sh[i], devices ? devices[i] : i, i);
    if (rc == FW_OK) rc = multi_common_create(m);
    if (rc == FW_OK && m->use_nccl) {
        rc = load_nccl();
        if (rc == FW_OK) {
            std::vector<ncclComm_t> comms(ndev);
            std::vector<int> devs(ndev);
            for (int i = 0; i < ndev; ++i) devs[i] = m->sh[i].device;
            ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, devs.data());
            if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitAll");
            else for (int i = 0; i < ndev; ++i) m->sh[i].comm = comms[i];
        }
    }
    if (rc != FW_OK) { fw_multi_destroy(m); return rc; }
    *out = m;
    return FW_OK;
}

int fw_multi_create_rank(int32_t device, int32_t rank, int32_t world, const void *nccl_id128, fw_multi **out) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_id128))
        return fail(FW_ERR_INVALID, "fw_multi_create_rank: bad argument");
    *out = nullptr;
    fw_multi *m = new (std::nothrow) fw_multi();
    if (!m) return fail(FW_ERR_NOMEM, "out of host memory");
    m->world = world; m->rank_mode = true;
    m->sh.resize(1);
    int rc = shard_init(m->sh[0], device, rank);
    if (rc == FW_OK) rc = multi_common_create(m);
    if (rc == FW_OK && world > 1) {      // NCCL always bootstraps the ranks (handle exchange, agreement on validation)
        rc = load_nccl();
        if (rc == FW_OK) {
            ncclUniqueId id;
            memcpy(&id, nccl_id128, sizeof(id));
            cudaSetDevice(device);
            ncclResult_t r = g_nccl.CommInitRank(&m->sh[0].comm, world, id, rank);
            if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitRank");
        }
    }
    if (rc != FW_OK) { fw_multi_destroy(m); return rc; }
    *out = m;
    return FW_OK;
}

void fw_multi_destroy(fw_multi *m) {
    if (!m) return;
    for (auto &s : m->sh) shard_release(s);
    delete m;
}

const char *fw_multi_last_error(fw_multi *m) {
    if (!m) return g_err.c_str();
    std::lock_guard<std::mutex> lk(m->err_mu);
    m->err_ret = m->err;
    return m->err_ret.c_str();
}

int fw_multi_alloc(fw_multi *m, int32_t n, int32_t want_paths) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_alloc: null object");
    FW_MENTER(m);
    return multi_alloc(m, n, want_paths != 0);
}

int fw_multi_upload(fw_multi *m, int32_t row0, int32_t rows, const double *rate, const int32_t *next) {
    if (!m || !rate || !next) return fail(FW_ERR_INVALID, "fw_multi_upload: bad argument");
    FW_MENTER(m);
    if (!m->allocated || row0 < 0 || rows < 0 || row0 + rows > m->n) return fail(FW_ERR_INVALID, "fw_multi_upload: bad row range");
    int rc;
    if ((rc = copy_rows(m, row0, rows, const_cast<double *>(rate), 0, [](Shard &s) { return s.rate.p; })) != FW_OK) return rc;
    if ((rc = copy_rows(m, row0, rows, const_cast<int32_t *>(next), 0, [](Shard &s) { return s.next.p; })) != FW_OK) return rc;
    m->solved = false;
    return multi_sync_all(m);
}

int fw_multi_solve_resident(fw_multi *m) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_solve_resident: null object");
    FW_MENTER(m);
    if (!m->allocated) return fail(FW_ERR_INVALID, "fw_multi_solve_resident: nothing allocated");
    const fwplan::Layout &L = m->L;
    if (L.n != m->n) {      // the caller filled rows < n only: pad rows / columns
        for (auto &s : m->sh) {
            CU(cudaSetDevice(s.device));
            const long long tot = (long long)L.rows_local() * L.n;
            fw_pad_shard_kernel<<<grid_for(tot, s.ctx->sm_count), 256, 0, s.sA>>>(s.rate.p, s.next.p, L.n, L.rows_local(), L.n, m->n,
                                                                                   L.cbr, L.world, s.rank);
            CU(cudaGetLastError());
        }
    }
    if (m->want_paths)
        for (auto &s : m->sh) {
            CU(cudaSetDevice(s.device));
            CU(cudaMemcpyAsync(s.init_next.p, s.next.p, (size_t)L.rows_local() * L.n * 4, cudaMemcpyDeviceToDevice, s.sA));
        }
    return multi_solve_resident(m);
}

int fw_multi_sync(fw_multi *m, int32_t n, const int32_t *ccy, int32_t ne, const int32_t *src, const int32_t *dst,
                  const double *val, int32_t want_paths) {
    if (!m || n < 0 || ne < 0) return fail(FW_ERR_INVALID, "fw_multi_sync: bad argument");
    FW_MENTER(m);
    if (n == 0) { m->n = 0; m->allocated = false; m->solved = true; return FW_OK; }
    if (!ccy || (ne > 0 && (!src || !dst || !val))) return fail(FW_ERR_INVALID, "fw_multi_sync: null input");
    int rc;
    if ((rc = multi_alloc(m, n, want_paths != 0)) != FW_OK) return rc;
    m->coo_resident = false;
    if ((rc = multi_upload_coo(m, n, ccy, ne, src, dst, val)) != FW_OK) return rc;
    if ((rc = multi_build(m)) != FW_OK) return rc;
    return multi_solve_resident(m, true);
}

int fw_multi_resolve(fw_multi *m) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_resolve: null object");
    FW_MENTER(m);
    if (!m->allocated || !m->coo_resident) return fail(FW_ERR_INVALID, "fw_multi_resolve: no fw_multi_sync came before");
    int rc;
    if ((rc = multi_build(m)) != FW_OK) return rc;
    return multi_solve_resident(m, true);
}

int fw_multi_download_local(fw_multi *m, int32_t i, double *rate, int32_t *next) {
    if (!m || i < 0 || i >= (int)m->sh.size()) return fail(FW_ERR_INVALID, "fw_multi_download_local: bad argument");
    FW_MENTER(m);
    if (!m->allocated) return fail(FW_ERR_INVALID, "fw_multi_download_local: nothing allocated");
    Shard &s = m->sh[i];
    const size_t tot = (size_t)m->L.rows_local() * m->L.n;
    CU(cudaSetDevice(s.device));
    if (rate) CU(cudaMemcpyAsync(rate, s.rate.p, tot * 8, cudaMemcpyDeviceToHost, s.sA));
    if (next) CU(cudaMemcpyAsync(next, s.next.p, tot * 4, cudaMemcpyDeviceToHost, s.sB));   // second lane: both copy engines busy
    CU(cudaStreamSynchronize(s.sA));
    CU(cudaStreamSynchronize(s.sB));
    return FW_OK;
}

int fw_multi_download(fw_multi *m, int32_t row0, int32_t rows, double *rate, int32_t *next, int32_t *init_next,
                      int32_t *mid, int32_t *csT, int32_t *rs) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_download: null object");
    FW_MENTER(m);
    if (m->n == 0 && rows == 0) return FW_OK;
    if (!m->allocated || row0 < 0 || rows < 0 || row0 + rows > m->n) return fail(FW_ERR_INVALID, "fw_multi_download: bad row range");
    if ((init_next || mid || csT || rs) && !m->want_paths) return fail(FW_ERR_INVALID, "fw_multi_download: no path tables were kept");
    int rc;
    if (rate && (rc = copy_rows(m, row0, rows, rate, 1, [](Shard &s) { return s.rate.p; })) != FW_OK) return rc;
    if (next && (rc = copy_rows(m, row0, rows, next, 1, [](Shard &s) { return s.next.p; })) != FW_OK) return rc;
    if (init_next && (rc = copy_rows(m, row0, rows, init_next, 1, [](Shard &s) { return s.init_next.p; })) != FW_OK) return rc;
    if (mid && (rc = copy_rows(m, row0, rows, mid, 1, [](Shard &s) { return s.mid.p; })) != FW_OK) return rc;
    if (csT && (rc = copy_rows(m, row0, rows, csT, 1, [](Shard &s) { return s.csT.p; })) != FW_OK) return rc;
    if (rs && (rc = copy_rows(m, row0, rows, rs, 1, [](Shard &s) { return s.rs.p; })) != FW_OK) return rc;
    return multi_sync_all(m);
}

int fw_multi_record_row_snapshots(fw_multi *m, int32_t on) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_record_row_snapshots: null object");
    FW_MENTER(m);
    m->record_snap = (on != 0);
    if (!m->record_snap) for (auto &s : m->sh) { cudaSetDevice(s.device); s.sink.release(); }
    return FW_OK;
}

int fw_multi_download_sink(fw_multi *m, int32_t row0, int32_t rows, double *out) {
    if (!m || !out) return fail(FW_ERR_INVALID, "fw_multi_download_sink: bad argument");
    FW_MENTER(m);
    if (!m->allocated || !m->solved || !m->record_snap || row0 < 0 || rows < 0 || row0 + rows > m->n)
        return fail(FW_ERR_INVALID, "fw_multi_download_sink: no recorded solve / bad row range");
    int rc = copy_rows(m, row0, rows, out, 1, [](Shard &s) { return s.sink.p; });
    return rc != FW_OK ? rc : multi_sync_all(m);
}

int fw_multi_solve_edges(fw_multi *m, int32_t n, const int32_t *ccy, int32_t ne, const int32_t *src, const int32_t *dst,
                         const double *val, double *rate, int32_t *next, int32_t *init_next, int32_t *mid, int32_t *csT,
                         int32_t *rs) {
    if (!m || n < 0) return fail(FW_ERR_INVALID, "fw_multi_solve_edges: bad argument");
    if (n == 0) return FW_OK;
    if (!rate || !next) return fail(FW_ERR_INVALID, "fw_multi_solve_edges: null output");
    if (!paths_args_ok(mid, csT, rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    FW_MENTER(m);
    int rc = fw_multi_sync(m, n, ccy, ne, src, dst, val, (mid || init_next) ? 1 : 0);
    if (rc != FW_OK) return rc;
    return fw_multi_download(m, 0, n, rate, next, init_next, mid, csT, rs);
}

int fw_multi_solve(fw_multi *m, int32_t n, double *rate, int32_t *next, int32_t *mid, int32_t *csT, int32_t *rs) {
    if (!m || n < 0) return fail(FW_ERR_INVALID, "fw_multi_solve: bad argument");
    if (n == 0) return FW_OK;
    if (!rate || !next) return fail(FW_ERR_INVALID, "fw_multi_solve: null buffer");
    if (!paths_args_ok(mid, csT, rs)) return fail(FW_ERR_INVALID, "mid/csT/rs: pass all three or none");
    FW_MENTER(m);
    int rc;
    if ((rc = multi_alloc(m, n, mid != nullptr)) != FW_OK) return rc;
    if ((rc = fw_multi_upload(m, 0, n, rate, next)) != FW_OK) return rc;
    if ((rc = fw_multi_solve_resident(m)) != FW_OK) return rc;
    return fw_multi_download(m, 0, n, rate, next, nullptr, mid, csT, rs);
}

int fw_multi_optimum(fw_multi *m, int32_t src, int32_t dst, double *rate, int32_t *path, int32_t cap, int32_t *path_len) {
    if (!m || !rate || !path_len || cap < 0 || (cap > 0 && !path)) return fail(FW_ERR_INVALID, "fw_multi_optimum: bad argument");
    FW_MENTER(m);
    if (!m->allocated || !m->solved) return fail(FW_ERR_INVALID, "fw_multi_optimum: state is not in sync (call fw_multi_sync)");
    if (!m->want_paths) return fail(FW_ERR_INVALID, "fw_multi_optimum: state was synced without path tables");
    if (m->rank_mode && m->world > 1) return fail(FW_ERR_INVALID, "fw_multi_optimum: needs the single-process mode (fw_multi_create)");
    if (m->world > fw::PATH_MAXSHARD) return fail(FW_ERR_INVALID, "fw_multi_optimum: more than 8 shards");
    if (src < 0 || dst < 0 || src >= m->n || dst >= m->n) return fail(FW_ERR_INVALID, "fw_multi_optimum: vertex index out of range");
    const fwplan::Layout &L = m->L;
    fw::PathTables t;
    const double *rsh[fw::PATH_MAXSHARD];
    for (int i = 0; i < fw::PATH_MAXSHARD; ++i) {
        const Shard &s = m->sh[i < m->world ? i : 0];
        t.init_next[i] = s.init_next.p; t.mid[i] = s.mid.p; t.csT[i] = s.csT.p; t.rs[i] = s.rs.p;
        rsh[i] = s.rate.p;
    }
    t.ld = L.n; t.n = m->n; t.cbr = L.cbr; t.P = L.world;
    Shard &q = m->sh[L.owner_of_row(src)];             // the walk starts in the shard that holds row src
    CU(cudaSetDevice(q.device));
    q.ctx->stream = q.sA;
    return optimum_locked(q.ctx, t, rsh, L.n, src, dst, rate, path, cap, path_len);
}

int fw_multi_download_locals(fw_multi *m, double *const *rate, int32_t *const *next) {
    if (!m || (!rate && !next)) return fail(FW_ERR_INVALID, "fw_multi_download_locals: bad argument");
    FW_MENTER(m);
    if (!m->allocated) return fail(FW_ERR_INVALID, "fw_multi_download_locals: nothing allocated");
    const size_t tot = (size_t)m->L.rows_local() * m->L.n;
    for (size_t i = 0; i < m->sh.size(); ++i) {     // every device's two copy engines at once
        Shard &s = m->sh[i];
        CU(cudaSetDevice(s.device));
        if (rate && rate[i]) CU(cudaMemcpyAsync(rate[i], s.rate.p, tot * 8, cudaMemcpyDeviceToHost, s.sA));
        if (next && next[i]) CU(cudaMemcpyAsync(next[i], s.next.p, tot * 4, cudaMemcpyDeviceToHost, s.sB));
    }
    return multi_sync_all(m);
}

const char *fw_multi_transport(fw_multi *m) {
    if (!m) return "";
    if (m->world == 1) return "none (one shard)";
    if (m->use_nccl) return "ncclBroadcast";
    if (m->use_ipc) return "copy-engine copies into CUDA-IPC-mapped peer buffers, flag words + stream memory operations";
    return "copy-engine peer copies, CUDA events";
}

int32_t fw_multi_local_shards(fw_multi *m) { return m ? (int32_t)m->sh.size() : 0; }

int fw_multi_shard(fw_multi *m, int32_t i, fw_shard_info *out) {
    if (!m || !out || i < 0 || i >= (int)m->sh.size()) return fail(FW_ERR_INVALID, "fw_multi_shard: bad argument");
    FW_MENTER(m);
    if (!m->allocated) return fail(FW_ERR_INVALID, "fw_multi_shard: nothing allocated");
    const Shard &s = m->sh[i];
    out->device = s.device; out->rank = s.rank; out->world = m->world; out->rows = m->L.rows_local();
    out->n_padded = m->L.n; out->cyclic_rows = m->L.cbr; out->group = m->L.G; out->ld = m->L.n;
    out->d_rate = s.rate.p; out->d_next = s.next.p;
    return FW_OK;
}

int fw_multi_last_solve_ms(fw_multi *m, double *ms, int64_t *launches) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_last_solve_ms: null object");
    if (ms) *ms = m->last_ms;
    if (launches) *launches = m->last_launches;
    return FW_OK;
}

int fw_multi_set_profiling(fw_multi *m, int32_t on) {
    if (!m) return fail(FW_ERR_INVALID, "fw_multi_set_profiling: null object");
    FW_MENTER(m);
    m->profiling = (on != 0);
    return FW_OK;
}

int fw_multi_phase_ms(fw_multi *m, double ms[4], int64_t count[4]) {
    if (!m || !ms || !count) return fail(FW_ERR_INVALID, "fw_multi_phase_ms: bad argument");
    FW_MENTER(m);
    return fw_ctx_phase_ms(m->sh[0].ctx, ms, count);    // local shard 0: a per-GPU figure
}

int64_t fw_multi_plan(int32_t n, int32_t world, int32_t block, int32_t group, int32_t cyclic_rows, fw_plan_op *ops,
                      int64_t cap) {
    fwplan::Layout L;
    L.n = n; L.world = world; L.B = block; L.G = group; L.cbr = cyclic_rows;
    if (!L.valid() || (cap > 0 && !ops)) return fail(FW_ERR_INVALID, "fw_multi_plan: bad layout");
    const std::vector<fw_plan_op> plan = fwplan::make_plan(L);
    for (int64_t i = 0; i < (int64_t)plan.size() && i < cap; ++i) ops[i] = plan[i];
    return (int64_t)plan.size();
}

}  // extern "C"
