// algorithms.hpp -- C++ host-side mirror of the reference's Algorithms module
// (/root/reference/src/lib/Algorithms.hs:2-5: buildMatrix, floydWarshall, optimum) and of the types
// it works on (/root/reference/src/lib/Types.hs:13-39), written over libfwgpu's C ABI
// (include/fwgpu.h).  The reference is Haskell and GHC is not available here, so this header is the
// compiled-language host side: same names, same argument meaning, same error texts.
// floydWarshall has NO CPU fallback: it calls fw_solve_edges (buildMatrix + runAlgo on the GPU) and
// fw_paths (exact `_path` lists) and throws FwGpuError when the library reports a failure.
#pragma once
#include <algorithm>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fwgpu.h"

namespace fwhost {

// Types.hs:13-20
struct Vertex {
    std::string exch, ccy;
    bool operator<(const Vertex &o) const { return exch != o.exch ? exch < o.exch : ccy < o.ccy; }   // derived Ord
    bool operator==(const Vertex &o) const { return exch == o.exch && ccy == o.ccy; }
    bool operator!=(const Vertex &o) const { return !(*this == o); }
    std::string show() const { return "(" + exch + ", " + ccy + ")"; }
};

// Types.hs:24-29
struct RateEntry {
    double bestRate;
    Vertex start;
    std::vector<Vertex> path;
    bool operator==(const RateEntry &o) const { return bestRate == o.bestRate && start == o.start && path == o.path; }
};

using Matrix = std::vector<std::vector<RateEntry>>;                 // Types.hs:39
using ExRates = std::map<std::pair<Vertex, Vertex>, double>;        // M.Map (Vertex, Vertex) Double

struct AlgoOptimumError : std::runtime_error {                      // Types.hs:62-63
    using std::runtime_error::runtime_error;
};
struct FwGpuError : std::runtime_error {
    int code;
    FwGpuError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline RateEntry isolatedEntry(const Vertex &start) { return RateEntry{0.0, start, {}}; }   // Utils.hs:13-14

// Algorithms.hs:29  sort . nub $ keys >>= \(k1,k2) -> [k1,k2]
inline std::vector<Vertex> sortedVertices(const ExRates &ex) {
    std::set<Vertex> s;
    for (const auto &kv : ex) { s.insert(kv.first.first); s.insert(kv.first.second); }
    return std::vector<Vertex>(s.begin(), s.end());
}

// Algorithms.hs:26-40
inline Matrix buildMatrix(const ExRates &ex) {
    const std::vector<Vertex> v = sortedVertices(ex);
    const size_t n = v.size();
    Matrix m(n);
    for (size_t i = 0; i < n; ++i) {
        m[i].reserve(n);
        for (size_t j = 0; j < n; ++j) {
            RateEntry e = isolatedEntry(v[i]);
            if (i != j) {
                if (v[i].ccy == v[j].ccy) { e.bestRate = 1.0; e.path = {v[j]}; }          // :35 before the lookup
                else {
                    auto it = ex.find({v[i], v[j]});
                    if (it != ex.end()) { e.bestRate = it->second; e.path = {v[j]}; }      // :36-37
                }
            }
            m[i].push_back(std::move(e));
        }
    }
    return m;
}

// Algorithms.hs:19-20  floydWarshall = runAlgo 0 . buildMatrix   (both on the GPU)
inline Matrix floydWarshall(const ExRates &ex, fw_ctx *ctx = nullptr) {
    const std::vector<Vertex> v = sortedVertices(ex);
    const int32_t n = (int32_t)v.size();
    if (n == 0) return Matrix{};                                     // floydWarshall M.empty == V.empty
    std::map<Vertex, int32_t> index;
    for (int32_t i = 0; i < n; ++i) index[v[i]] = i;
    std::map<std::string, int32_t> ccyId;
    std::vector<int32_t> ccy(n), src, dst;
    std::vector<double> val;
    for (int32_t i = 0; i < n; ++i) ccy[i] = ccyId.emplace(v[i].ccy, (int32_t)ccyId.size()).first->second;
    for (const auto &kv : ex) { src.push_back(index[kv.first.first]); dst.push_back(index[kv.first.second]); val.push_back(kv.second); }
    const size_t nn = (size_t)n * n;
    std::vector<double> rate(nn);
    std::vector<int32_t> next(nn), init(nn), mid(nn), csT(nn), rs(nn);
    int rc = fw_solve_edges(ctx, n, ccy.data(), (int32_t)src.size(), src.data(), dst.data(), val.data(), rate.data(),
                            next.data(), init.data(), mid.data(), csT.data(), rs.data());
    if (rc != FW_OK) throw FwGpuError(rc, fw_last_error());
    // every `_path` (Algorithms.hs:55), expanded on the device
    std::vector<int32_t> q(2 * nn);
    for (int32_t i = 0; i < n; ++i) for (int32_t j = 0; j < n; ++j) { q[2 * ((size_t)i * n + j)] = i; q[2 * ((size_t)i * n + j) + 1] = j; }
    std::vector<int64_t> off(nn + 1);
    std::vector<int32_t> verts(std::max<size_t>(64, 16 * nn));
    rc = fw_paths(ctx, n, init.data(), mid.data(), csT.data(), rs.data(), (int32_t)nn, q.data(), off.data(), verts.data(), (int64_t)verts.size());
    if (rc == FW_ERR_CAP && off[nn] > (int64_t)verts.size()) {
        verts.resize((size_t)off[nn]);
        rc = fw_paths(ctx, n, init.data(), mid.data(), csT.data(), rs.data(), (int32_t)nn, q.data(), off.data(), verts.data(), (int64_t)verts.size());
    }
    if (rc != FW_OK) throw FwGpuError(rc, fw_last_error());
    Matrix m(n);
    for (int32_t i = 0; i < n; ++i) {
        m[i].reserve(n);
        for (int32_t j = 0; j < n; ++j) {
            const size_t e = (size_t)i * n + j;
            RateEntry re{rate[e], v[i], {}};
            for (int64_t t = off[e]; t < off[e + 1]; ++t) re.path.push_back(v[verts[t]]);
            m[i].push_back(std::move(re));
        }
    }
    return m;
}

// Algorithms.hs:65-78 -- same checks, same order, same texts
inline RateEntry optimum(const Vertex &src, const Vertex &dest, const Matrix &matrix) {
    std::vector<Vertex> starts;
    for (const auto &row : matrix) {
        if (row.empty()) throw AlgoOptimumError("The matrix is empty");
        starts.push_back(row[0].start);
    }
    auto idx = [&](const Vertex &x) -> size_t {
        auto it = std::find(starts.begin(), starts.end(), x);
        if (it == starts.end()) throw AlgoOptimumError(x.show() + " is not entered before");
        return (size_t)(it - starts.begin());
    };
    const size_t si = idx(src), di = idx(dest);
    const std::string notReachable = "There is no exchange between " + src.show() + " and " + dest.show();
    if (di >= matrix[si].size()) throw AlgoOptimumError(notReachable);
    const RateEntry &e = matrix[si][di];
    if (e.path.empty()) throw AlgoOptimumError(notReachable);
    return e;
}

}  // namespace fwhost
