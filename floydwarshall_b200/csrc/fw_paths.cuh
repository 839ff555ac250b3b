// Device-side expansion of the reference's `_path` lists.
//
// The reference concatenates ikPath ++ kjPath as they were AT STEP k
// (src/lib/Algorithms.hs:55), so the final next-hop matrix is not enough; the
// solve records mid / csT / rs and this kernel replays the recursion
//   path(i,j)  = edge(i,j)                 if mid[i][j] < 0
//              = col(i,m) ++ row(m,j)      m = mid[i][j]
//   col(a,k)   = edge(a,k) if csT[a][k]<0 else col(a,m') ++ row(m',k),  m' = csT[a][k]
//   row(k,b)   = edge(k,b) if rs[k][b]<0  else col(k,m') ++ row(m',b),  m' = rs[k][b]
//   edge(a,b)  = [b] iff init_next[a][b] >= 0   (the pre-solve matrix), else []
// with an explicit per-thread stack.  One thread per query; run twice
// (count, then fill at the prefix-summed offsets).
#pragma once
#include "fw_common.cuh"

namespace fw {

constexpr int PATH_STACK = 512;

struct PathArgs {
    const int32_t *init_next, *mid, *csT, *rs;
    long long ld;
    int n;
    int nq;
    const int32_t *queries;     // 2*nq (src, dst)
    long long *lengths;         // nq   (pass 0 output)
    const long long *offsets;   // nq+1 (pass 1 input)
    int32_t *verts;             // pass 1 output
    long long max_len;          // per-path cap
    int *flag;                  // bit 0: stack overflow, bit 1: path longer than max_len, bit 2: bad query
};

template <int PASS>
__global__ void fw_paths_kernel(PathArgs a) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.nq) return;
    const int src = a.queries[2 * q], dst = a.queries[2 * q + 1];
    if (src < 0 || dst < 0 || src >= a.n || dst >= a.n) {
        atomicOr(a.flag, 4);
        if (PASS == 0) a.lengths[q] = 0;
        return;
    }
    // stack item: kind (0 final, 1 col, 2 row) << 62 | a << 31 | b
    unsigned long long st[PATH_STACK];
    int sp = 0;
    st[sp++] = ((unsigned long long)src << 31) | (unsigned long long)dst;
    long long len = 0;
    int32_t *out = (PASS == 1) ? a.verts + a.offsets[q] : nullptr;
    while (sp > 0) {
        const unsigned long long it = st[--sp];
        const int kind = (int)(it >> 62);
        const int x = (int)((it >> 31) & 0x7fffffffu), y = (int)(it & 0x7fffffffu);
        const long long off = (long long)x * a.ld + y;
        const int m = (kind == 0) ? a.mid[off] : (kind == 1 ? a.csT[off] : a.rs[off]);
        if (m < 0) {
            if (a.init_next[off] >= 0) {
                if (len >= a.max_len) { atomicOr(a.flag, 2); break; }
                if (PASS == 1) out[len] = y;
                ++len;
            }
            continue;
        }
        if (sp + 2 > PATH_STACK) { atomicOr(a.flag, 1); break; }
        st[sp++] = (2ull << 62) | ((unsigned long long)m << 31) | (unsigned long long)y;   // row(m, y) later
        st[sp++] = (1ull << 62) | ((unsigned long long)x << 31) | (unsigned long long)m;   // col(x, m) first
    }
    if (PASS == 0) a.lengths[q] = len;
}

}  // namespace fw
