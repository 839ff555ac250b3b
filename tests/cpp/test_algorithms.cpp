// The reference's AlgorithmsTest.hs (src/test/AlgorithmsTest.hs:45-110) against the C++ host mirror
// floydwarshall_b200/host/algorithms.hpp.  `cpu` runs the host-only cases, `gpu` everything.
#include <cstdio>
#include <cstring>
#include <string>

#include "../../floydwarshall_b200/host/algorithms.hpp"

using namespace fwhost;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static const Vertex gdax_btc{"GDAX", "BTC"}, gdax_usd{"GDAX", "USD"}, kraken_btc{"KRAKEN", "BTC"},
    kraken_usd{"KRAKEN", "USD"}, kraken_stc{"KRAKEN", "STC"};

static ExRates rates4() {   // MockData.hs:47-57
    return {{{kraken_btc, kraken_usd}, 1000.0}, {{kraken_usd, kraken_btc}, 0.0009},
            {{gdax_btc, gdax_usd}, 1001.0}, {{gdax_usd, gdax_btc}, 0.0008}};
}

struct Cell { double r; std::vector<int> p; };
static Matrix rateMatrixForTest(const std::vector<Vertex> &v, const std::vector<std::vector<Cell>> &rows) {   // TestUtils.hs:11-21
    Matrix m;
    for (size_t i = 0; i < rows.size(); ++i) {
        m.emplace_back();
        for (const auto &c : rows[i]) {
            RateEntry e{c.r, v[i], {}};
            for (int k : c.p) e.path.push_back(v[k]);
            m.back().push_back(e);
        }
    }
    return m;
}

static std::string errOf(const Vertex &s, const Vertex &d, const Matrix &m) {
    try { optimum(s, d, m); } catch (const AlgoOptimumError &e) { return e.what(); }
    return "";
}

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    const std::vector<Vertex> v = {gdax_btc, gdax_usd, kraken_btc, kraken_usd};
    // buildMatrix_emptyMatrix / buildMatrix_4x4Matrix
    CHECK(buildMatrix({}).empty());
    CHECK(sortedVertices(rates4()) == v);
    CHECK(buildMatrix(rates4()) == rateMatrixForTest(v, {
        {{0, {}}, {1001, {1}}, {1, {2}}, {0, {}}},
        {{0.0008, {0}}, {0, {}}, {0, {}}, {1, {3}}},
        {{1, {0}}, {0, {}}, {0, {}}, {1000, {3}}},
        {{0, {}}, {1, {1}}, {0.0009, {2}}, {0, {}}}}));
    // optimum on hand-made matrices (matrix may be empty / rows empty)
    CHECK(errOf(kraken_btc, kraken_usd, {}) == "(KRAKEN, BTC) is not entered before");
    CHECK(errOf(kraken_btc, kraken_usd, Matrix{{}, {}}) == "The matrix is empty");
    // floydWarshall_emptyMatrix needs no device
    CHECK(floydWarshall({}).empty());
    if (gpu) {
        if (fw_device_count() == 0) { std::printf("no CUDA device\n"); return 2; }
        const Matrix m = floydWarshall(rates4());
        CHECK(m == rateMatrixForTest(v, {                                        // AlgorithmsTest.hs:66-77
            {{0, {}}, {1001, {1}}, {1, {2}}, {1001, {1, 3}}},
            {{0.0009, {3, 2, 0}}, {0, {}}, {0.0009, {3, 2}}, {1, {3}}},
            {{1, {0}}, {1001, {0, 1}}, {0, {}}, {1001, {0, 1, 3}}},
            {{0.0009, {2, 0}}, {1, {1}}, {0.0009, {2}}, {0, {}}}}));
        CHECK(errOf(kraken_stc, kraken_usd, m) == "(KRAKEN, STC) is not entered before");   // :82-91
        CHECK(errOf(kraken_btc, kraken_stc, m) == "(KRAKEN, STC) is not entered before");
        Matrix m2 = m;                                                           // :93-110
        m2[3][0] = isolatedEntry(kraken_usd);
        CHECK(errOf(kraken_usd, gdax_btc, m2) == "There is no exchange between (KRAKEN, USD) and (GDAX, BTC)");
        CHECK(optimum(kraken_btc, gdax_usd, m2) == (RateEntry{1001.0, kraken_btc, {gdax_btc, gdax_usd}}));
        CHECK(optimum(gdax_usd, gdax_btc, m2) == (RateEntry{0.0009, gdax_usd, {kraken_usd, kraken_btc, gdax_btc}}));
        // ProcessRequestsTest.hs:154-162: KRAKEN BTC -> KRAKEN USD = 1001.0 via GDAX
        CHECK(optimum(kraken_btc, kraken_usd, m) == (RateEntry{1001.0, kraken_btc, {gdax_btc, gdax_usd, kraken_usd}}));
    } else {
        // without a device the GPU path must fail loudly (no CPU fallback)
        if (fw_device_count() == 0) {
            bool threw = false;
            try { floydWarshall(rates4()); } catch (const FwGpuError &e) { threw = (e.code == FW_ERR_CUDA); }
            CHECK(threw);
        }
    }
    std::printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
