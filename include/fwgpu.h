/*
 * fwgpu.h -- C ABI of libfwgpu.so, the B200 (sm_100a) implementation of the
 * matrix-optimisation hot path of jinilover/floydWarshall.
 *
 * The reference has no FFI today; the boundary this library sits under is the
 * pure Haskell function
 *     floydWarshall :: M.Map (Vertex, Vertex) Double -> Matrix RateEntry
 *     (reference src/lib/Algorithms.hs:19-20  = runAlgo 0 . buildMatrix)
 * whose body (runAlgo, src/lib/Algorithms.hs:42-61) is what fw_solve* replace.
 * INTEGRATION.md shows the `foreign import ccall` stubs a maintainer adds.
 *
 * Dense encoding (row-major, n x n, caller-owned):
 *   rate[i*n+j] = _bestRate (m ! i ! j)                   binary64
 *   next[i*n+j] = index of (head _path), -1 if _path==[]  int32
 * optional exact-path side tables (all int32, n x n, -1 = "initial edge"):
 *   mid[i*n+j]  = k of the last step that replaced entry (i,j)
 *   csT[i*n+k]  = mid of entry (i,k) when step k began
 *   rs [k*n+j]  = mid of entry (k,j) when step k began
 * (the reference's `_path = ikPath ++ kjPath`, Algorithms.hs:55, concatenates
 * the sub-paths as they were AT STEP k; mid/csT/rs is the minimal record that
 * reproduces it -- fw_paths expands it.)
 *
 * Semantics are exactly the reference loop: for k ascending, every entry with
 * i != k, j != k, j != i is replaced iff  rate[i][j] < rate[i][k]*rate[k][j]
 * (strict, one rounded binary64 multiply), taking next[i][k].  Results are
 * bit-identical to that loop for every input in the domain below.
 *
 * Domain: rate entries must not be negative (NaN and +inf are tolerated and
 * behave as in the reference); wherever rate[i][j] > 0 (i != j) next[i][j]
 * must be >= 0 -- both hold for anything buildMatrix (Algorithms.hs:26-40)
 * can produce from parser-validated input (Parsers.hs:40: rate > 0).
 * Violations return FW_ERR_DOMAIN and leave the buffers untouched.
 * The diagonal is never read nor written (Algorithms.hs:50,54).
 *
 * Errors: 0 = OK, negative = error; text via fw_ctx_last_error(ctx) (kept per
 * context: safe when the failing call and the query run on different OS
 * threads, as unbound GHC threads do) or fw_last_error() (per OS thread).
 * No exceptions cross this boundary.  There is no CPU fallback: without a
 * CUDA device every compute entry point returns FW_ERR_CUDA.
 *
 * Threading: a context serialises its own calls -- every entry point holds the
 * context's lock from its first to its last touch of the context's buffers
 * (composite calls such as fw_solve_edges included); distinct contexts may be
 * used from distinct threads.  Entry points call cudaSetDevice themselves, so
 * they may be called from any OS thread (GHC `safe` foreign calls).
 */
#ifndef FWGPU_H
#define FWGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FW_OK 0
#define FW_ERR_INVALID (-1) /* bad argument (null pointer, negative size, ...) */
#define FW_ERR_CUDA (-2)    /* CUDA runtime failure / no device               */
#define FW_ERR_DOMAIN (-3)  /* input outside the documented domain            */
#define FW_ERR_NOMEM (-4)   /* device or host allocation failed               */
#define FW_ERR_CAP (-5)     /* output capacity too small (fw_paths)           */

#define FW_TILE 128 /* k-block size B and the largest "single tile" graph */

typedef struct fw_ctx fw_ctx; /* opaque: device id, stream, workspace */

/* ---- library / device ------------------------------------------------- */
const char *fw_version(void);
const char *fw_last_error(void);
/* Text of the last failure of a call made on ctx (NULL: the default context). */
const char *fw_ctx_last_error(fw_ctx *ctx);
int fw_device_count(void);

/* Create a context on `device` (>= 0).  Workspace grows on demand. */
int fw_ctx_create(int device, fw_ctx **out);
void fw_ctx_destroy(fw_ctx *ctx);
/* external != 0: launch on the caller's cudaStream_t `cuda_stream` (a NULL
 * handle is the legacy default stream, which is what torch's current stream
 * usually is); external == 0: back to the context's own stream. */
int fw_ctx_set_stream(fw_ctx *ctx, void *cuda_stream, int external);
/* Kernel launches issued by the last solve on this context. */
int64_t fw_ctx_last_launches(const fw_ctx *ctx);

/* Optional per-phase timing: when on, every kernel launch of a solve is
 * bracketed by CUDA events on the launching stream.  fw_ctx_phase_ms waits for
 * the stream and returns, for the last solve, the summed device time and the
 * launch count of: [0] diagonal-tile kernel, [1] column-panel kernel,
 * [2] row-panel kernel, [3] bulk (phase 3) kernel. */
int fw_ctx_set_profiling(fw_ctx *ctx, int on);
int fw_ctx_phase_ms(fw_ctx *ctx, double ms[4], int64_t count[4]);
/* Per-launch device times (ms, launch order) of one phase of the last solve;
 * returns the number of launches of that phase (may exceed cap), < 0 on error. */
int64_t fw_ctx_phase_spans(fw_ctx *ctx, int phase, double *ms, int64_t cap);

/* ---- replaces runAlgo (Algorithms.hs:42-61) ---------------------------- */
/* Host buffers, in place.  mid/csT/rs may be NULL (all three or none).
 * n == 0 succeeds and touches nothing (floydWarshall M.empty == V.empty,
 * reference src/test/AlgorithmsTest.hs:62-64).  ctx may be NULL: a
 * process-wide default context on device 0 is used. */
int fw_solve(fw_ctx *ctx, int32_t n, double *rate, int32_t *next,
             int32_t *mid, int32_t *csT, int32_t *rs);

/* Same on DEVICE-resident buffers with leading dimension ld (elements),
 * asynchronous on the context's stream.  Zero-copy when n is a multiple of
 * FW_TILE (or n <= FW_TILE); otherwise the library works on a padded copy. */
int fw_solve_device(fw_ctx *ctx, int32_t n, int64_t ld, double *d_rate,
                    int32_t *d_next, int32_t *d_mid, int32_t *d_csT,
                    int32_t *d_rs);

/* Verification entry: only the k-blocks [kb0, kb1) (pivots kb0*128 .. kb1*128-1) of the solve, with the
 * schedule (k-blocks per fused launch, look-ahead) the FULL solve of this n would use.  After the call the
 * matrix is the reference loop's state as step kb1*128 begins (Algorithms.hs:44), so a test can compare a
 * window of the shipped schedule at BASELINE size with a few CPU steps.  n % 128 == 0, ld % 4 == 0,
 * 16-byte aligned buffers; validation and the mid/csT/rs reset happen only when kb0 == 0. */
int fw_solve_device_range(fw_ctx *ctx, int32_t n, int64_t ld, double *d_rate,
                          int32_t *d_next, int32_t *d_mid, int32_t *d_csT,
                          int32_t *d_rs, int32_t kb0, int32_t kb1);
/* Verification hook: while set (d_sink != NULL), every solve on ctx also stores row k of the rate matrix AS
 * STEP k BEGINS into d_sink[k*ld .. k*ld+n) (device memory, n_padded x ld doubles; entry [k][k] is stored as
 * 0.0).  These are the pivot rows the loop reads (Algorithms.hs:60); with them a CPU oracle can replay any
 * single row's whole history (oracle/fw_oracle.c: fw_oracle_replay_rows). */
int fw_ctx_set_row_snapshot_sink(fw_ctx *ctx, double *d_sink, int64_t ld);

/* `batch` independent graphs of the same n, batch-major contiguous
 * (the FSM replay: one full solve per OutSync snapshot,
 * reference src/lib/ProcessRequests.hs:82-84,97-102). */
int fw_solve_batched(fw_ctx *ctx, int32_t batch, int32_t n, double *rate,
                     int32_t *next, int32_t *mid, int32_t *csT, int32_t *rs);
int fw_solve_batched_device(fw_ctx *ctx, int32_t batch, int32_t n,
                            double *d_rate, int32_t *d_next, int32_t *d_mid,
                            int32_t *d_csT, int32_t *d_rs);

/* ---- replaces the `_path` field of RateEntry (Algorithms.hs:55) ------------
 * Expands, on the device, the exact reference path (start excluded, destination
 * included) of nq (src,dst) index pairs from the side tables of a paths-enabled
 * solve plus the PRE-solve next matrix (`edge(a,b)` exists iff
 * init_next[a*n+b] >= 0).  offsets[nq+1] and verts[cap] are host outputs in CSR
 * form.  If the paths need more than `cap` entries the call returns FW_ERR_CAP
 * with offsets filled (offsets[nq] = entries needed) so the caller can re-size.
 * A single path longer than 2^24 hops (arbitrage cycles) is FW_ERR_CAP too; the
 * recursion depth is not limited (walks deeper than the 64-slot on-chip stack
 * continue in a global overflow area, n + 2 slots always suffice).
 * Unreachable pairs yield empty paths.  fw_paths takes HOST tables (n x n),
 * fw_paths_device DEVICE tables with leading dimension ld. */
int fw_paths(fw_ctx *ctx, int32_t n, const int32_t *init_next, const int32_t *mid,
             const int32_t *csT, const int32_t *rs, int32_t nq, const int32_t *queries,
             int64_t *offsets, int32_t *verts, int64_t cap);
int fw_paths_device(fw_ctx *ctx, int32_t n, int64_t ld, const int32_t *d_init_next,
                    const int32_t *d_mid, const int32_t *d_csT, const int32_t *d_rs,
                    int32_t nq, const int32_t *queries, int64_t *offsets, int32_t *verts,
                    int64_t cap);

/* The four tables of ONE optimised matrix uploaded once and kept on the device, so that the lazily
 * evaluated `_path` fields of a Matrix RateEntry (one thunk per entry) cost one small call each
 * instead of re-uploading 16 n^2 bytes. */
typedef struct fw_tables fw_tables;
int fw_tables_create(fw_ctx *ctx, int32_t n, const int32_t *init_next, const int32_t *mid,
                     const int32_t *csT, const int32_t *rs, fw_tables **out);
void fw_tables_destroy(fw_tables *t);
int fw_tables_paths(fw_tables *t, int32_t nq, const int32_t *queries, int64_t *offsets,
                    int32_t *verts, int64_t cap);

/* ---- replaces buildMatrix (Algorithms.hs:26-40) on the device ---------------
 * The cache in COO form: n vertices in the reference's sorted order
 * (Algorithms.hs:29), ccy[i] = any integer id of vertex i's currency, m map
 * entries (src[e], dst[e]) -> val[e] with unique keys (it is a Map).  Rules as in
 * the reference, in its order: i == j -> (0.0, []); same currency -> (1.0, [j])
 * BEFORE the map lookup; map hit -> (val, [j]); else (0.0, []).  All inputs are
 * host arrays; outputs are device matrices with leading dimension ld. */
int fw_build_matrix_device(fw_ctx *ctx, int32_t n, int64_t ld, const int32_t *ccy, int32_t m,
                           const int32_t *src, const int32_t *dst, const double *val,
                           double *d_rate, int32_t *d_next);

/* floydWarshall (Algorithms.hs:19-20) in ONE call, map in / dense matrix out: the
 * cache goes up in COO form (as fw_build_matrix_device), buildMatrix + runAlgo
 * run on the device, the dense result lands in the HOST outputs rate[n*n],
 * next[n*n] and, if non-NULL, init_next (the buildMatrix next-hops) and the
 * exact-path tables mid/csT/rs (all three or none). */
int fw_solve_edges(fw_ctx *ctx, int32_t n, const int32_t *ccy, int32_t m, const int32_t *src,
                   const int32_t *dst, const double *val, double *rate, int32_t *next,
                   int32_t *init_next, int32_t *mid, int32_t *csT, int32_t *rs);

/* ---- the InSync state kept on the device (Types.hs:35-37; ProcessRequests.hs:78-85)
 * fw_state_sync = syncMatrix on an OutSync state: buildMatrix + runAlgo on the
 * device; the optimised matrix (rate, next and the exact-path tables) STAYS in
 * HBM.  fw_state_optimum = the read-out of `optimum` (Algorithms.hs:74-75) for
 * one (src, dst) index pair: *rate = _bestRate, path[0..*path_len) = `_path` as
 * vertex indices (start excluded); an empty path means "no exchange between".
 * Only the answer crosses PCIe, never the matrix. */
typedef struct fw_state fw_state;
int fw_state_create(fw_ctx *ctx, fw_state **out);
void fw_state_destroy(fw_state *st);
int fw_state_sync(fw_state *st, int32_t n, const int32_t *ccy, int32_t m, const int32_t *src,
                  const int32_t *dst, const double *val);
int fw_state_optimum(fw_state *st, int32_t src, int32_t dst, double *rate, int32_t *path,
                     int32_t cap, int32_t *path_len);
/* Optional full read-back (tests): host rate[n*n] and/or next[n*n]. */
int fw_state_download(fw_state *st, double *rate, int32_t *next);

/* ---- row-sharded building blocks (one shard per GPU) ----------------------
 * The multi-GPU solve keeps rows [row0, row0+rows) of the n x n
 * matrix on each GPU (n, row0, rows multiples of FW_TILE; ld % 4 == 0).  For
 * every k-block b0 = 0, 128, ...:
 *   owner of rows [b0, b0+128):  fw_shard_pivot  -> fills d_Rw (128 x n, the
 *        step-k snapshots of the pivot rows) from its diagonal tile + row panel
 *   caller broadcasts d_Rw from the owner to all ranks (NCCL)
 *   every rank:                  fw_shard_update -> column panel + bulk on its rows
 * Next-hops never cross ranks (NX[i][j] <- NX[i][k] is row-local).  All calls
 * are asynchronous on the context's stream.  No exact-path tables here. */
int fw_shard_validate(fw_ctx *ctx, int32_t n, int32_t row0, int32_t rows, int64_t ld,
                      const double *d_rate, const int32_t *d_next);
int fw_shard_pivot(fw_ctx *ctx, int32_t n, int32_t row0, int32_t rows, int64_t ld,
                   double *d_rate, int32_t *d_next, int32_t b0, double *d_Rw);
int fw_shard_update(fw_ctx *ctx, int32_t n, int32_t row0, int32_t rows, int64_t ld,
                    double *d_rate, int32_t *d_next, int32_t b0, const double *d_Rw);
/* Look-ahead variant: mode 0 = fw_shard_update; mode 1 = ONLY the 128 local rows starting at
 * lr0 (the next k-block's pivot rows, so its owner can factor them early on a second stream);
 * mode 2 = everything mode 0 does EXCEPT those 128 rows (lr0 must be adjacent to the k-block
 * rows when the shard owns them). */
int fw_shard_update_ex(fw_ctx *ctx, int32_t n, int32_t row0, int32_t rows, int64_t ld,
                       double *d_rate, int32_t *d_next, int32_t b0, const double *d_Rw,
                       int32_t mode, int32_t lr0);
/* Two CONSECUTIVE k-blocks b0, b0+128 in one call, so that the bulk kernel loads every tile of the
 * shard once per 256 steps (the single-GPU solve's pairing): column panel of b0, bulk(b0) on the
 * column strip of b0+128, column panel of b0+128, then one fused bulk launch.  d_Rw0 / d_Rw1 are the
 * row-snapshot panels of the two blocks (both pivoted and broadcast before the call).  Needs row0,
 * rows, b0 multiples of 256.  Rows: mode 0 = every local row outside the pair's own 256 rows;
 * mode 1 = ONLY the lrn local rows starting at lr0; mode 2 = mode 0 minus those rows (adjacent to
 * the pair's rows when the shard owns them).  The pair's own rows are the owner's business:
 * fw_shard_pivot(b0), fw_shard_update_ex(b0, mode 1, rows of b0+128), fw_shard_pivot(b0+128) before
 * the broadcasts, fw_shard_update_ex(b0+128, mode 1, rows of b0) after them. */
int fw_shard_update_pair(fw_ctx *ctx, int32_t n, int32_t row0, int32_t rows, int64_t ld,
                         double *d_rate, int32_t *d_next, int32_t b0, const double *d_Rw0,
                         const double *d_Rw1, int32_t mode, int32_t lr0, int32_t lrn);
/* nb (1..8) CONSECUTIVE k-blocks b0, b0+128, ... in one call (fw_shard_update_pair is nb = 2): for each
 * block in turn its column panel, preceded by one fused launch that brings the block's column strip up
 * to date with the earlier blocks of the call, then one fused bulk launch of all nb blocks.  d_Rw[i] is
 * the row-snapshot panel of block i (HOST array of nb device pointers).  Rows: mode 0 = every local row
 * outside [b0, b0 + nb*128); mode 1 = ONLY the lrn local rows starting at lr0 (any rows that do not
 * intersect the blocks' own); mode 2 = mode 0 minus those rows (adjacent to the blocks' rows when the
 * shard owns them).  The blocks' own rows are the owner's business, before and after the broadcasts
 * (floydwarshall_b200/sharded.py: run_schedule_lookahead_groups). */
int fw_shard_update_group(fw_ctx *ctx, int32_t n, int32_t row0, int32_t rows, int64_t ld,
                          double *d_rate, int32_t *d_next, int32_t b0, int32_t nb,
                          const double *const *d_Rw, int32_t mode, int32_t lr0, int32_t lrn);

/* Block until everything queued on the context's stream has finished and
 * report any asynchronous failure (incl. FW_ERR_DOMAIN of *_device calls). */
int fw_ctx_synchronize(fw_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FWGPU_H */
