/*
 * fw_oracle.c -- CPU restatement of the reference's matrix optimisation.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path.  It
 * may be built/loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Nothing under floydwarshall_b200/
 * links, imports or falls back to it.
 *
 * Parity status: PINNED.  The reference (Haskell, GHC 8.6.5) cannot be built
 * in this image (no ghc/cabal/stack/nix), so this restatement is pinned
 * against every golden vector the reference's own tests hold for the path
 * (tests/golden/reference_vectors.json, transcribed from
 * src/test/AlgorithmsTest.hs:49-110, src/test/ProcessRequestsTest.hs:83-95,
 * 154-162 and README.md:210-246) by tests/test_oracle_golden.py.
 *
 * What is restated (all paths relative to /root/reference):
 *   src/lib/Algorithms.hs:42-61   runAlgo / updateRow / updateCol
 *   src/lib/Utils.hs:13-14        isolatedEntry = RateEntry 0.0 start []
 *
 * Dense encoding used everywhere in this repo:
 *   rate[i*n+j] = _bestRate (matrix ! i ! j)                      (binary64)
 *   next[i*n+j] = index of (head _path), -1 when _path == []      (int32)
 *   mid [i*n+j] = k of the last step that replaced the entry, -1 = never
 *   csT [i*n+k] = mid of entry (i,k) as it stood when step k began
 *   rs  [k*n+j] = mid of entry (k,j) as it stood when step k began
 * mid/csT/rs are what an exact `_path` needs: the reference concatenates the
 * two sub-paths *as they were at step k* (Algorithms.hs:55, ikPath ++ kjPath),
 * which the final next-hop matrix alone cannot reproduce.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp  (never -ffast-math).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#ifdef _OPENMP
#include <omp.h>
#else  /* single-threaded build (no OpenMP runtime available) */
static int omp_get_max_threads(void) { return 1; }
static void omp_set_num_threads(int n) { (void)n; }
#endif
#endif

/* threads > 0: that many OpenMP threads; threads <= 0: every core of the host.  (The setting is sticky in
 * OpenMP, so "all cores" has to be requested explicitly -- an earlier single-threaded call would otherwise
 * leave every later call single-threaded.) */
static void fw_oracle_set_threads(int32_t threads)
{
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#else
    (void)threads;
#endif
}

/*
 * fw_oracle_run_generations -- the literal form.
 *
 * Algorithms.hs:44  runAlgo k matrix | k < matrixSize = runAlgo (k+1) newMatrix
 *   every k builds a NEW matrix from the OLD one: all reads below come from
 *   `cur`, all writes go to `nxt`; the buffers swap per k.
 * Algorithms.hs:50  i == k           -> row copied unchanged
 * Algorithms.hs:54  any (== j) [i,k] -> entry copied unchanged
 * Algorithms.hs:55  _bestRate origEntry < newRate -> replace (strict <)
 * Algorithms.hs:61  newRate = ikRate * kjRate  (one rounded binary64 multiply)
 * Algorithms.hs:55  _path = ikPath ++ kjPath   => head = head ikPath
 *   (ikPath is non-empty whenever the replace fires on in-domain data because
 *   an empty path carries rate 0.0 and 0 < 0*x is false; for out-of-domain
 *   data -- negative rates -- ikPath may be empty and the head is then the
 *   head of kjPath, which is handled below so that the oracle stays literal.)
 *
 * Returns the number of replacements performed (for the update-rate stats),
 * or -1 on allocation failure.
 */
int64_t fw_oracle_run_generations(int32_t n, double *rate, int32_t *next,
                                  int32_t *mid, int32_t *csT, int32_t *rs)
{
    if (n <= 0) return 0;
    size_t nn = (size_t)n * (size_t)n;
    double *ra = rate, *rb = (double *)malloc(nn * sizeof(double));
    int32_t *na = next, *nb = (int32_t *)malloc(nn * sizeof(int32_t));
    int32_t *ma = NULL, *mb = NULL;
    if (!rb || !nb) { free(rb); free(nb); return -1; }
    if (mid) {
        ma = mid; mb = (int32_t *)malloc(nn * sizeof(int32_t));
        if (!mb) { free(rb); free(nb); return -1; }
        for (size_t e = 0; e < nn; ++e) ma[e] = -1;
    }
    int64_t updates = 0;
    for (int32_t k = 0; k < n; ++k) {
        if (mid && csT) for (int32_t i = 0; i < n; ++i) csT[(size_t)i * n + k] = ma[(size_t)i * n + k];
        if (mid && rs)  for (int32_t j = 0; j < n; ++j) rs[(size_t)k * n + j]  = ma[(size_t)k * n + j];
        for (int32_t i = 0; i < n; ++i) {
            const size_t ro = (size_t)i * n;
            if (i == k) {                                   /* Algorithms.hs:50 */
                memcpy(rb + ro, ra + ro, (size_t)n * sizeof(double));
                memcpy(nb + ro, na + ro, (size_t)n * sizeof(int32_t));
                if (mid) memcpy(mb + ro, ma + ro, (size_t)n * sizeof(int32_t));
                continue;
            }
            const double ik_rate = ra[ro + k];              /* Algorithms.hs:59 */
            const int32_t ik_next = na[ro + k];
            for (int32_t j = 0; j < n; ++j) {
                double orig = ra[ro + j];                   /* Algorithms.hs:58 */
                int32_t onext = na[ro + j];
                int32_t omid = mid ? ma[ro + j] : -1;
                if (j != i && j != k) {                     /* Algorithms.hs:54 */
                    const double kj_rate = ra[(size_t)k * n + j];   /* :60 */
                    const double new_rate = ik_rate * kj_rate;      /* :61 */
                    if (orig < new_rate) {                          /* :55 */
                        orig = new_rate;
                        /* head (ikPath ++ kjPath) */
                        onext = (ik_next >= 0) ? ik_next : na[(size_t)k * n + j];
                        omid = k;
                        ++updates;
                    }
                }
                rb[ro + j] = orig;
                nb[ro + j] = onext;
                if (mid) mb[ro + j] = omid;
            }
        }
        { double *t = ra; ra = rb; rb = t; }
        { int32_t *t = na; na = nb; nb = t; }
        if (mid) { int32_t *t = ma; ma = mb; mb = t; }
    }
    if (ra != rate) {
        memcpy(rate, ra, nn * sizeof(double));
        memcpy(next, na, nn * sizeof(int32_t));
        if (mid) memcpy(mid, ma, nn * sizeof(int32_t));
        free(ra); free(na); if (mid) free(ma);
    } else {
        free(rb); free(nb); if (mid) free(mb);
    }
    return updates;
}

/*
 * fw_oracle_run_inplace -- same loop, one buffer, optional OpenMP over i.
 *
 * Legal because step k never writes row k (Algorithms.hs:50) nor column k
 * (Algorithms.hs:54), and those are the only cells step k reads besides the
 * cell it replaces.  tests/test_oracle_golden.py checks it bit-for-bit
 * against fw_oracle_run_generations.  `threads` <= 0 means all cores.
 * This is the variant used as the CPU baseline.
 */
int64_t fw_oracle_run_inplace(int32_t n, double *rate, int32_t *next,
                              int32_t *mid, int32_t *csT, int32_t *rs,
                              int32_t threads)
{
    if (n <= 0) return 0;
    size_t nn = (size_t)n * (size_t)n;
    if (mid) for (size_t e = 0; e < nn; ++e) mid[e] = -1;
    fw_oracle_set_threads(threads);
    int64_t updates = 0;
    for (int32_t k = 0; k < n; ++k) {
        if (mid && csT) for (int32_t i = 0; i < n; ++i) csT[(size_t)i * n + k] = mid[(size_t)i * n + k];
        if (mid && rs)  for (int32_t j = 0; j < n; ++j) rs[(size_t)k * n + j]  = mid[(size_t)k * n + j];
        const double *rk = rate + (size_t)k * n;
        const int32_t *nk = next + (size_t)k * n;
#pragma omp parallel for schedule(static) reduction(+ : updates)
        for (int32_t i = 0; i < n; ++i) {
            if (i == k) continue;
            double *ri = rate + (size_t)i * n;
            int32_t *ni = next + (size_t)i * n;
            int32_t *mi = mid ? mid + (size_t)i * n : NULL;
            const double ik_rate = ri[k];
            const int32_t ik_next = ni[k];
            for (int32_t j = 0; j < n; ++j) {
                if (j == i || j == k) continue;
                const double new_rate = ik_rate * rk[j];
                if (ri[j] < new_rate) {
                    ri[j] = new_rate;
                    ni[j] = (ik_next >= 0) ? ik_next : nk[j];
                    if (mi) mi[j] = k;
                    ++updates;
                }
            }
        }
    }
    return updates;
}

/*
 * fw_oracle_run_ksteps -- in-place loop restricted to k in [k0, k1).
 * Used to time a BOUNDED sample of a large workload (bench.py cpu_baseline):
 * (k1-k0)*n*n relaxations of the full-size matrix.
 */
int64_t fw_oracle_run_ksteps(int32_t n, double *rate, int32_t *next,
                             int32_t k0, int32_t k1, int32_t threads)
{
    if (n <= 0) return 0;
    fw_oracle_set_threads(threads);
    int64_t updates = 0;
    if (k1 > n) k1 = n;
    for (int32_t k = k0; k < k1; ++k) {
        const double *rk = rate + (size_t)k * n;
        const int32_t *nk = next + (size_t)k * n;
#pragma omp parallel for schedule(static) reduction(+ : updates)
        for (int32_t i = 0; i < n; ++i) {
            if (i == k) continue;
            double *ri = rate + (size_t)i * n;
            int32_t *ni = next + (size_t)i * n;
            const double ik_rate = ri[k];
            const int32_t ik_next = ni[k];
            for (int32_t j = 0; j < n; ++j) {
                if (j == i || j == k) continue;
                const double new_rate = ik_rate * rk[j];
                if (ri[j] < new_rate) {
                    ri[j] = new_rate;
                    ni[j] = (ik_next >= 0) ? ik_next : nk[j];
                    ++updates;
                }
            }
        }
    }
    return updates;
}

/*
 * fw_oracle_run_batched -- `batch` independent graphs of order n, batch-major
 * contiguous; every graph goes through fw_oracle_run_inplace's loop on its own
 * (OpenMP over the graphs).  The FSM replay (ProcessRequests.hs:82-84: one full
 * floydWarshall per OutSync snapshot) at BASELINE config C3's full batch.
 */
int64_t fw_oracle_run_batched(int32_t batch, int32_t n, double *rate, int32_t *next,
                              int32_t *mid, int32_t *csT, int32_t *rs, int32_t threads)
{
    if (n <= 0 || batch <= 0) return 0;
    fw_oracle_set_threads(threads);
    const size_t nn = (size_t)n * (size_t)n;
    int64_t updates = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : updates)
    for (int32_t g = 0; g < batch; ++g) {
        double *R = rate + g * nn;
        int32_t *X = next + g * nn;
        int32_t *M = mid ? mid + g * nn : NULL, *C = csT ? csT + g * nn : NULL, *S = rs ? rs + g * nn : NULL;
        if (M) for (size_t e = 0; e < nn; ++e) M[e] = -1;
        for (int32_t k = 0; k < n; ++k) {
            if (M && C) for (int32_t i = 0; i < n; ++i) C[(size_t)i * n + k] = M[(size_t)i * n + k];
            if (M && S) for (int32_t j = 0; j < n; ++j) S[(size_t)k * n + j] = M[(size_t)k * n + j];
            const double *rk = R + (size_t)k * n;
            const int32_t *nk = X + (size_t)k * n;
            for (int32_t i = 0; i < n; ++i) {
                if (i == k) continue;                               /* Algorithms.hs:50 */
                double *ri = R + (size_t)i * n;
                int32_t *ni = X + (size_t)i * n;
                const double ik_rate = ri[k];
                const int32_t ik_next = ni[k];
                for (int32_t j = 0; j < n; ++j) {
                    if (j == i || j == k) continue;                 /* Algorithms.hs:54 */
                    const double new_rate = ik_rate * rk[j];        /* :61 */
                    if (ri[j] < new_rate) {                         /* :55 */
                        ri[j] = new_rate;
                        ni[j] = (ik_next >= 0) ? ik_next : nk[j];
                        if (M) M[(size_t)i * n + j] = k;
                        ++updates;
                    }
                }
            }
        }
    }
    return updates;
}

/*
 * fw_oracle_replay_rows -- the reference loop restricted to a SAMPLE of matrix
 * rows, for sizes where the full O(n^3) loop is out of reach of a test.
 *
 * In step k the loop (Algorithms.hs:49-61) touches row i using only row i itself
 * and row k as it stands when step k begins (row k is not written in step k,
 * :50).  Given the sequence of those pivot rows  S[k][:] = R_k[k][:]  (recorded
 * by the implementation under test), the whole history of any single row i
 * follows from its initial contents:
 *     for k ascending, k != i:  for j != i, j != k:
 *         n = row[k] * S[k][j];  if (row[j] < n) { row[j] = n; nx[j] = nx[k]; mid[j] = k }
 * and at k == i the row must itself equal S[i] -- which ties the recorded
 * sequence back to the loop.  If every row were replayed this is the full loop
 * (induction over k); a sample checks the sampled rows' final rates, next-hops
 * and mids, their csT rows (mid[k] as step k begins) and, at step i, the recorded
 * pivot row and its rs row, bit for bit.
 *
 * rows[r] = sampled row index; row/nx: nrows x n, initial contents in, final out;
 * mid, csT (nullable): nrows x n out; at_i / mid_at_i (nullable): nrows x n out =
 * the row and its mids as step i begins.  Returns replacements, -1 - r if row r
 * met a replacement through an empty path (out of domain: needs row k's next-hops).
 */
int64_t fw_oracle_replay_rows(int32_t n, int32_t nrows, const int32_t *rows, const double *S, int64_t ldS,
                              double *row, int32_t *nx, int32_t *mid, int32_t *csT, double *at_i,
                              int32_t *mid_at_i, int32_t threads)
{
    if (n <= 0 || nrows <= 0) return 0;
    fw_oracle_set_threads(threads);
    int64_t updates = 0;
    int32_t bad = -1;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : updates)
    for (int32_t r = 0; r < nrows; ++r) {
        const int32_t i = rows[r];
        double *ri = row + (size_t)r * n;
        int32_t *ni = nx + (size_t)r * n;
        int32_t *mi = (int32_t *)malloc((size_t)n * sizeof(int32_t));
        if (!mi) { bad = r; continue; }
        for (int32_t j = 0; j < n; ++j) mi[j] = -1;
        for (int32_t k = 0; k < n; ++k) {
            if (csT) csT[(size_t)r * n + k] = mi[k];
            if (k == i) {                                           /* Algorithms.hs:50 */
                if (at_i) memcpy(at_i + (size_t)r * n, ri, (size_t)n * sizeof(double));
                if (mid_at_i) memcpy(mid_at_i + (size_t)r * n, mi, (size_t)n * sizeof(int32_t));
                continue;
            }
            const double *sk = S + (size_t)k * (size_t)ldS;
            const double ik_rate = ri[k];
            const int32_t ik_next = ni[k];
            for (int32_t j = 0; j < n; ++j) {
                if (j == i || j == k) continue;                     /* Algorithms.hs:54 */
                const double new_rate = ik_rate * sk[j];            /* :61 */
                if (ri[j] < new_rate) {                             /* :55 */
                    if (ik_next < 0) bad = r;
                    ri[j] = new_rate;
                    ni[j] = ik_next;
                    mi[j] = k;
                    ++updates;
                }
            }
        }
        if (mid) memcpy(mid + (size_t)r * n, mi, (size_t)n * sizeof(int32_t));
        free(mi);
    }
    return bad >= 0 ? -1 - (int64_t)bad : updates;
}

int32_t fw_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
