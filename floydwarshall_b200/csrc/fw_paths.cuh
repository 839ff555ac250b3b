// Device-side expansion of the reference's `_path` lists.
//
// The reference concatenates ikPath ++ kjPath as they were AT STEP k
// (src/lib/Algorithms.hs:55), so the final next-hop matrix is not enough; the
// solve records mid / csT / rs and these kernels replay the recursion
//   path(i,j)  = edge(i,j)                 if mid[i][j] < 0
//              = col(i,m) ++ row(m,j)      m = mid[i][j]
//   col(a,k)   = edge(a,k) if csT[a][k]<0 else col(a,m') ++ row(m',k),  m' = csT[a][k]
//   row(k,b)   = edge(k,b) if rs[k][b]<0  else col(k,m') ++ row(m',b),  m' = rs[k][b]
//   edge(a,b)  = [b] iff init_next[a][b] >= 0   (the pre-solve matrix), else []
// with an explicit per-thread stack.  The first index of the pending items strictly
// decreases from the bottom of the stack to its top (every expansion replaces an
// item by two items with a smaller pivot), so n + 1 slots always suffice: the
// first PATH_LSTACK slots live in local memory, deeper ones in a global overflow
// area the host provides when a first attempt reports "stack overflow".
//
// The tables may be ROW-SHARDED (multi-GPU solve, cyclic blocks of rows: fw_plan.hpp);
// peers are read through NVLink (peer access enabled).
#pragma once
#include "fw_common.cuh"

namespace fw {

constexpr int PATH_LSTACK = 64;
constexpr int PATH_MAXSHARD = 8;

struct PathTables {
    const int32_t *init_next[PATH_MAXSHARD], *mid[PATH_MAXSHARD], *csT[PATH_MAXSHARD], *rs[PATH_MAXSHARD];
    long long ld;
    int n;
    int cbr, P;                 // global row x lives on shard (x / cbr) % P at local row (x / (cbr*P)) * cbr + x % cbr
};                              // (unsharded: cbr >= n, P = 1)

struct PathArgs {
    PathTables t;
    int nq;
    const int32_t *queries;     // 2*nq (src, dst)
    long long *lengths;         // nq   (pass 0 output)
    const long long *offsets;   // nq+1 (pass 1 input)
    int32_t *verts;             // pass 1 output
    long long max_len;          // per-path cap
    unsigned long long *gstack; // overflow stack: gcap slots per thread (nullable)
    long long gcap;
    int *flag;                  // bit 0: stack overflow, bit 1: path longer than max_len, bit 2: bad query
};

// Walks path(src, dst); emits vertex y through `emit(pos, y)` for pos < limit and keeps counting beyond it.
// Returns the length, or -1 (stack overflow) / -2 (longer than max_len).
template <typename Emit>
__device__ __forceinline__ long long walk_path(const PathTables &t, int src, int dst, long long max_len,
                                               unsigned long long *gst, long long gcap, Emit emit) {
    // stack item: kind (0 final, 1 col, 2 row) << 62 | a << 31 | b
    unsigned long long st[PATH_LSTACK];
    long long sp = 0;
    auto push = [&](unsigned long long v) -> bool {
        if (sp < PATH_LSTACK) st[sp] = v;
        else if (gst != nullptr && sp - PATH_LSTACK < gcap) gst[sp - PATH_LSTACK] = v;
        else return false;
        ++sp;
        return true;
    };
    push(((unsigned long long)src << 31) | (unsigned long long)dst);
    long long len = 0;
    while (sp > 0) {
        --sp;
        const unsigned long long it = (sp < PATH_LSTACK) ? st[sp] : gst[sp - PATH_LSTACK];
        const int kind = (int)(it >> 62);
        const int x = (int)((it >> 31) & 0x7fffffffu), y = (int)(it & 0x7fffffffu);
        const int cb = x / t.cbr, sh = cb % t.P;
        const long long off = (long long)((cb / t.P) * t.cbr + x % t.cbr) * t.ld + y;
        const int m = (kind == 0) ? t.mid[sh][off] : (kind == 1 ? t.csT[sh][off] : t.rs[sh][off]);
        if (m < 0) {
            if (t.init_next[sh][off] >= 0) {
                if (len >= max_len) return -2;
                emit(len, y);
                ++len;
            }
            continue;
        }
        if (!push((2ull << 62) | ((unsigned long long)m << 31) | (unsigned long long)y)) return -1;   // row(m, y) later
        if (!push((1ull << 62) | ((unsigned long long)x << 31) | (unsigned long long)m)) return -1;   // col(x, m) first
    }
    return len;
}

// One thread per query; run twice (count, then fill at the prefix-summed offsets).
template <int PASS>
__global__ void fw_paths_kernel(PathArgs a) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.nq) return;
    const int src = a.queries[2 * q], dst = a.queries[2 * q + 1];
    if (src < 0 || dst < 0 || src >= a.t.n || dst >= a.t.n) {
        atomicOr(a.flag, 4);
        if (PASS == 0) a.lengths[q] = 0;
        return;
    }
    int32_t *out = (PASS == 1) ? a.verts + a.offsets[q] : nullptr;
    unsigned long long *gst = a.gstack ? a.gstack + (long long)q * a.gcap : nullptr;
    long long len = walk_path(a.t, src, dst, a.max_len, gst, a.gcap,
                              [&](long long pos, int y) { if (PASS == 1) out[pos] = y; });
    if (len == -1) { atomicOr(a.flag, 1); len = 0; }
    if (len == -2) { atomicOr(a.flag, 2); len = a.max_len; }
    if (PASS == 0) a.lengths[q] = len;
}

// `optimum` (Algorithms.hs:74-75) for ONE pair in ONE launch: rate, path length and the first `cap`
// path vertices go straight into a mapped pinned host record  { double rate; long long len; int32 verts[cap] }.
struct OptimumArgs {
    PathTables t;
    const double *rate[PATH_MAXSHARD];   // rate shards (leading dimension rate_ld)
    long long rate_ld;
    int src, dst, cap;
    long long max_len;
    unsigned long long *gstack;          // n + 2 slots
    long long gcap;
    unsigned char *out;                  // device view of the mapped host record
};

__global__ void fw_optimum_kernel(OptimumArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double *o_rate = reinterpret_cast<double *>(a.out);
    long long *o_len = reinterpret_cast<long long *>(a.out + 8);
    int32_t *o_verts = reinterpret_cast<int32_t *>(a.out + 16);
    const int cb = a.src / a.t.cbr, sh = cb % a.t.P;
    *o_rate = a.rate[sh][(long long)((cb / a.t.P) * a.t.cbr + a.src % a.t.cbr) * a.rate_ld + a.dst];
    const int cap = a.cap;
    *o_len = walk_path(a.t, a.src, a.dst, a.max_len, a.gstack, a.gcap,
                       [&](long long pos, int y) { if (pos < cap) o_verts[pos] = y; });
    __threadfence_system();
}

}  // namespace fw
