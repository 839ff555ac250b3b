// Shared definitions for the sm_100a max-times Floyd-Warshall kernels.
//
// Semantics being implemented (reference src/lib/Algorithms.hs:42-61):
//   for k ascending: for every (i,j) with i != k, j != k, j != i:
//       n = R[i][k] * R[k][j]          one rounded binary64 multiply  (:61)
//       if (R[i][j] < n)               strict, ordered                (:55)
//           R[i][j] = n;  NX[i][j] = NX[i][k];  MID[i][j] = k
//
// Two facts shape every kernel here:
//  * Step k writes neither row k nor column k, so in-place == the reference's
//    generation-per-k form.
//  * The three skip rules (i != k, j != k, j != i) need no branches:
//      - an entry that must never be replaced (the diagonal, padding) is held on
//        chip as NaN (tile kernel) or +inf (bulk kernel, so that the DFMA filter
//        stays quiet): o < n is false for every n;
//      - the pivot R[k][k] is held as NaN inside the tile kernel and exported to
//        the panel kernels as 0.0: every product of step k with i == k or j == k
//        is then NaN or 0, and neither beats an entry >= 0.  The diagonal is never
//        a legitimate factor (i != k for R[i][k], j != k for R[k][j]), so no other
//        product changes.  NaN is exact for every IEEE input; 0.0 is exact on the
//        documented domain (entries >= 0, NaN and +inf tolerated; include/fwgpu.h),
//        which fw_validate_kernel enforces before any kernel runs.
//    Padding rows and columns (n not a multiple of the tile) are NaN in memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef FW_B
#define FW_B 128            // k-block size == tile edge of the diagonal kernel
#endif

namespace fw {

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double pinf() { return __longlong_as_double(0x7ff0000000000000LL); }

// one relaxation on register state; returns whether it fired
__device__ __forceinline__ bool relax(double &o, double n) {
    bool p = o < n;
    o = p ? n : o;
    return p;
}

// cp.async 16-byte global->shared (LDGSTS), L2-only (.cg): panels are streamed
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// shared-memory mbarriers (split arrive / wait: the tile kernel lets warps run one step ahead of each other)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @!p bra W;\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Swizzled position of column `col` (0..127) inside a 128-wide shared row so
// that a thread owning 8 consecutive columns [l*8, l*8+8) reads them as four
// conflict-free LDS.128: position = q*32 + l*2 + e  for col = l*8 + q*2 + e.
__device__ __forceinline__ int swz128(int col) {
    int l = col >> 3, q = (col >> 1) & 3, e = col & 1;
    return q * 32 + l * 2 + e;
}

}  // namespace fw
