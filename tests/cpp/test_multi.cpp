// fw_multi_* (the multi-GPU entry of the C ABI, include/fwgpu.h) from plain C++: the reference's golden
// answers (src/test/AlgorithmsTest.hs:66-110, src/test/ProcessRequestsTest.hs:154-162) through
// fw_multi_sync + fw_multi_optimum, and bit-exact agreement of fw_multi_solve_edges with the single-GPU
// fw_solve_edges on a generated graph.  `cpu`: the schedule-as-data entry (fw_multi_plan) and argument
// checks, no device needed;  `gpu [ndev]`: ndev shards (default 2) -- virtual ranks on device 0 when the
// box has fewer GPUs.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/fwgpu.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    // ---- the schedule as data: every k-block pivoted once, by the rank that holds its rows
    {
        const int n = 4096, world = 4, B = 128, G = 8, cbr = G * B;
        const int64_t cnt = fw_multi_plan(n, world, B, G, cbr, nullptr, 0);
        CHECK(cnt > 0);
        std::vector<fw_plan_op> ops((size_t)cnt);
        CHECK(fw_multi_plan(n, world, B, G, cbr, ops.data(), cnt) == cnt);
        std::vector<int> pivoted(n / B, 0);
        for (const auto &o : ops)
            if (o.kind == FW_OP_PIVOT) { pivoted[o.b0 / B]++; CHECK((o.b0 / cbr) % world == o.rank); }
        for (int c : pivoted) CHECK(c == 1);
        CHECK(fw_multi_plan(n, world, B, G, cbr + B, nullptr, 0) < 0);      // cyclic block not a whole number of groups
        fw_multi *bad = nullptr;
        CHECK(fw_multi_create(0, nullptr, &bad) == FW_ERR_INVALID);
    }
    if (gpu) {
        if (fw_device_count() == 0) { std::printf("no CUDA device\n"); return 2; }
        const int want = argc > 2 ? std::atoi(argv[2]) : 2;
        std::vector<int32_t> devs(want);
        for (int i = 0; i < want; ++i) devs[i] = i % fw_device_count();
        fw_multi *m = nullptr;
        CHECK(fw_multi_create(want, devs.data(), &m) == FW_OK);
        // the 4-vertex graph of MockData.hs:47-57; vertex order GDAX-BTC, GDAX-USD, KRAKEN-BTC, KRAKEN-USD
        const int32_t ccy[4] = {0, 1, 0, 1};
        const int32_t src[4] = {2, 3, 0, 1}, dst[4] = {3, 2, 1, 0};
        const double val[4] = {1000.0, 0.0009, 1001.0, 0.0008};
        CHECK(fw_multi_sync(m, 4, ccy, 4, src, dst, val, 1) == FW_OK);
        double r = 0; int32_t path[16], len = 0;
        CHECK(fw_multi_optimum(m, 2, 1, &r, path, 16, &len) == FW_OK);            // KRAKEN BTC -> GDAX USD
        CHECK(r == 1001.0 && len == 2 && path[0] == 0 && path[1] == 1);            // AlgorithmsTest.hs:105-107
        CHECK(fw_multi_optimum(m, 1, 0, &r, path, 16, &len) == FW_OK);            // GDAX USD -> GDAX BTC
        CHECK(r == 0.0009 && len == 3 && path[0] == 3 && path[1] == 2 && path[2] == 0);   // :108-110
        CHECK(fw_multi_optimum(m, 2, 3, &r, path, 16, &len) == FW_OK);            // ProcessRequestsTest.hs:154-162
        CHECK(r == 1001.0 && len == 3 && path[0] == 0 && path[1] == 1 && path[2] == 3);
        CHECK(fw_multi_optimum(m, 0, 0, &r, path, 16, &len) == FW_OK && len == 0); // diagonal: empty path
        CHECK(fw_multi_optimum(m, 9, 0, &r, path, 16, &len) == FW_ERR_INVALID);
        CHECK(std::strlen(fw_multi_last_error(m)) > 0);
        // a generated 40-exchange x 12-currency graph (n = 480): multi == single, bit for bit, tables included
        const int E = 40, C = 12, n = E * C;
        std::mt19937_64 rng(7);
        std::uniform_real_distribution<double> price(0.01, 5e4), spread(5e-4, 1e-2), coin(0, 1);
        std::vector<double> p(C);
        for (auto &x : p) x = price(rng);
        std::vector<int32_t> cc(n), es, ed;
        std::vector<double> ev;
        for (int i = 0; i < n; ++i) cc[i] = i % C;
        for (int e = 0; e < E; ++e)
            for (int a = 0; a < C; ++a)
                for (int b = a + 1; b < C; ++b) {
                    if (coin(rng) > 0.7) continue;
                    const double s = spread(rng);
                    es.push_back(e * C + a); ed.push_back(e * C + b); ev.push_back(p[a] / p[b] * (1 - s));
                    es.push_back(e * C + b); ed.push_back(e * C + a); ev.push_back(p[b] / p[a] * (1 - s));
                }
        const size_t tot = (size_t)n * n;
        std::vector<double> r1(tot), r2(tot);
        std::vector<int32_t> t1[5], t2[5];
        for (int i = 0; i < 5; ++i) { t1[i].resize(tot); t2[i].resize(tot); }
        CHECK(fw_solve_edges(nullptr, n, cc.data(), (int32_t)es.size(), es.data(), ed.data(), ev.data(), r1.data(),
                             t1[0].data(), t1[1].data(), t1[2].data(), t1[3].data(), t1[4].data()) == FW_OK);
        CHECK(fw_multi_solve_edges(m, n, cc.data(), (int32_t)es.size(), es.data(), ed.data(), ev.data(), r2.data(),
                                   t2[0].data(), t2[1].data(), t2[2].data(), t2[3].data(), t2[4].data()) == FW_OK);
        CHECK(std::memcmp(r1.data(), r2.data(), tot * 8) == 0);
        for (int i = 0; i < 5; ++i) CHECK(t1[i] == t2[i]);
        double ms = 0; int64_t launches = 0;
        CHECK(fw_multi_last_solve_ms(m, &ms, &launches) == FW_OK && ms > 0 && launches > 0);
        fw_multi_destroy(m);
    }
    std::printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
