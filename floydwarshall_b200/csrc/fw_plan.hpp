// Schedule of the row-sharded multi-GPU solve as DATA: a list of semantic operations that the CUDA
// executor (fw_multi.cuh) issues and that tests replay with a numpy model of the same operations
// (tests/np_plan_backend.py) -- one schedule, two executors, so a schedule bug cannot hide in a re-typed copy.
//
// Pure host C++ (no CUDA): compiled into libfwgpu.so and exported through fw_multi_plan().
//
// Sharding (generalised): the n rows are cut into CYCLIC BLOCKS of cbr rows; cyclic block c
// lives on rank c % world at local rows [(c / world) * cbr, +cbr).  cbr = n / world is the contiguous
// row-block sharding of the survey; cbr = one k-block group makes the ownership of the pivot rows rotate
// from group to group, so that the pivot work (diagonal tile + 128 x n row panel) is spread evenly over
// the ranks instead of falling on one rank for n / world consecutive pivots.
//
// k-blocks of B = 128 pivots are taken in GROUPS of G consecutive blocks (a group never straddles two cyclic
// blocks).  Per group p, with panels Rw[G * (p & 1) + j] for its block j:
//   look-ahead lane (B), owner of the group:   for j = 0 .. G-1:
//        rows of block j take the group's blocks 0 .. j-1            APPLY (rows of block j only)
//        diagonal tile + row panel of block j  -> Rw[..+j]           PIVOT
//        broadcast Rw[..+j] from the owner                           BCAST (every rank)
//   main lane (A), every rank:  all local rows take the group's G blocks from ONE load of each tile; a row
//        lying in the group's own block i already has blocks 0 .. i and takes i+1 .. G-1 only       APPLY
//   The owner of the NEXT group first brings that group's rows up to date on lane B (APPLY restricted to
//   those rows) and factors it there while lane A is still busy with the current group; lane A leaves those
//   rows out.  Every entry still sees its relaxations in ascending k (Algorithms.hs:44), so results are
//   identical to the plain loop.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/fwgpu.h"

namespace fwplan {

struct Layout {
    int n = 0;        // padded matrix order
    int world = 1;
    int B = 128;      // k-block
    int G = 1;        // k-blocks per group
    int cbr = 0;      // rows per cyclic block (multiple of G*B; n % (cbr*world) == 0)
    int rows_local() const { return n / world; }
    int owner_of_row(int g) const { return (g / cbr) % world; }
    int local_of_row(int g) const { return (g / (cbr * world)) * cbr + g % cbr; }
    int global_of_local(int rank, int l) const { return ((l / cbr) * world + rank) * cbr + l % cbr; }
    bool valid() const {
        return n > 0 && world > 0 && B > 0 && G > 0 && cbr > 0 && cbr % (G * B) == 0 && n % (cbr * world) == 0;
    }
};

inline fw_plan_op mk(int kind, int rank, int lane) {
    fw_plan_op o;
    o.kind = kind; o.rank = rank; o.lane = lane; o.b0 = 0; o.nb = 0; o.buf = 0;
    o.row_lo = 0; o.row_n = 0; o.ex_lo = -1; o.ex_n = 0; o.grp_lo = -1;
    return o;
}

// The whole job's operations in issue order (every rank's; a rank-mode executor skips other ranks' PIVOT /
// APPLY / event operations and takes part in every BCAST).
inline std::vector<fw_plan_op> make_plan(const Layout &L) {
    std::vector<fw_plan_op> ops;
    if (!L.valid()) return ops;
    const int B = L.B, G = L.G, GB = G * B, ngrp = L.n / GB, P = L.world;
    auto group_owner = [&](int p) { return L.owner_of_row(p * GB); };
    auto group_local = [&](int p) { return L.local_of_row(p * GB); };
    auto factor = [&](int p) {
        const int b0 = p * GB, s = G * (p & 1), ow = group_owner(p), g0 = group_local(p);
        for (int j = 0; j < G; ++j) {
            if (j > 0) {
                fw_plan_op a = mk(FW_OP_APPLY, ow, 1);
                a.b0 = b0; a.nb = j; a.buf = s; a.row_lo = g0 + j * B; a.row_n = B; a.grp_lo = g0;
                ops.push_back(a);
            }
            fw_plan_op pv = mk(FW_OP_PIVOT, ow, 1);
            pv.b0 = b0 + j * B; pv.nb = 1; pv.buf = s + j; pv.row_lo = g0 + j * B; pv.row_n = B;
            ops.push_back(pv);
            if (P > 1) {
                fw_plan_op bc = mk(FW_OP_BCAST, ow, 1);
                bc.b0 = b0 + j * B; bc.nb = 1; bc.buf = s + j;
                ops.push_back(bc);
            }
        }
    };
    factor(0);
    for (int r = 0; r < P; ++r) ops.push_back(mk(FW_OP_B_DONE, r, 1));
    for (int p = 0; p < ngrp; ++p) {
        const int b0 = p * GB, s = G * (p & 1);
        const bool nxt = p + 1 < ngrp;
        const int ow = group_owner(p), own = nxt ? group_owner(p + 1) : -1;
        for (int r = 0; r < P; ++r) ops.push_back(mk(FW_OP_WAIT_B, r, 0));      // all panels of group p are here
        if (nxt) {
            for (int r = 0; r < P; ++r) ops.push_back(mk(FW_OP_WAIT_A, r, 1));  // group p-1 is finished: rows of p+1 and its buffers are free
            fw_plan_op a = mk(FW_OP_APPLY, own, 1);                             // only the next group's rows
            a.b0 = b0; a.nb = G; a.buf = s; a.row_lo = group_local(p + 1); a.row_n = GB;
            a.grp_lo = (own == ow) ? group_local(p) : -1;
            ops.push_back(a);
            factor(p + 1);
            for (int r = 0; r < P; ++r) ops.push_back(mk(FW_OP_B_DONE, r, 1));
        }
        for (int r = 0; r < P; ++r) {
            fw_plan_op a = mk(FW_OP_APPLY, r, 0);
            a.b0 = b0; a.nb = G; a.buf = s; a.row_lo = 0; a.row_n = L.rows_local();
            if (r == own) { a.ex_lo = group_local(p + 1); a.ex_n = GB; }
            a.grp_lo = (r == ow) ? group_local(p) : -1;
            ops.push_back(a);
            ops.push_back(mk(FW_OP_A_DONE, r, 0));
        }
    }
    return ops;
}

}  // namespace fwplan
