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

/* ---- multi-GPU solve (config C5) ------------------------------
 * ONE object drives the row-sharded solve; the k-block schedule, the streams and
 * the pivot-panel broadcast all live inside the library, so that the reference's
 * single in-process call (Algorithms.hs:19-20, sole caller ProcessRequests.hs:82-84)
 * reaches every GPU of the box.  Two ways to build it:
 *   fw_multi_create       one process drives ndev GPUs (devices[] or NULL = 0..ndev-1).
 *                         The same device may be named several times ("virtual ranks":
 *                         tests on a one-GPU box; transport is then device copies).
 *   fw_multi_create_rank  one process per GPU (torchrun / MPI style): this process
 *                         holds shard `rank` of `world` on `device`; nccl_id = the
 *                         128 bytes fw_multi_unique_id() produced on rank 0.
 * Transport of the 128 x n row-snapshot panel (measured in DESIGN.md section 6): by
 * default the copy engines -- cudaMemcpyPeerAsync ordered by CUDA events in one
 * process; with one process per GPU, copies into CUDA-IPC-mapped peer buffers ordered
 * by flag words in device memory and stream memory operations (no host round trip).
 * FW_MULTI_TRANSPORT=nccl selects ncclBroadcast instead.  libnccl is dlopen'ed on first
 * use: one process needs it only for that transport, rank mode also to bootstrap the
 * ranks (exchange of the IPC handles, agreement on a validation error).
 *
 * Sharding: rows in cyclic blocks of one k-block group (ownership of the pivot
 * rows rotates over the ranks; FW_MULTI_CYCLIC=0: contiguous row blocks); k-blocks
 * in groups of up to 8 per fused bulk launch (FW_MULTI_GROUP=1|2|4|8).  n is padded
 * internally; any n >= 1 works.
 *
 * State: fw_multi_sync is syncMatrix on an OutSync state (ProcessRequests.hs:82-84):
 * the cache goes up in COO form, every shard runs buildMatrix for its rows and the
 * solve; the optimised matrix and its exact-path tables STAY sharded in HBM.
 * fw_multi_optimum reads one entry + its `_path` (Algorithms.hs:74-75) across the
 * shards (single-process mode; peers are read over NVLink).  fw_multi_download
 * copies any row range to the host (rank mode: local rows only are written, other
 * rows of the range are left untouched). */
typedef struct fw_multi fw_multi;
int fw_multi_create(int32_t ndev, const int32_t *devices, fw_multi **out);
int fw_multi_unique_id(void *id128);
int fw_multi_create_rank(int32_t device, int32_t rank, int32_t world, const void *nccl_id128,
                         fw_multi **out);
/* Rank mode: the objects of all ranks are created, used and destroyed together (every call that touches the
 * matrix is collective, like the solve itself); keep a rank's process alive until every rank has destroyed its
 * object, because peers hold CUDA-IPC mappings of its panel buffers until then. */
void fw_multi_destroy(fw_multi *m);
const char *fw_multi_last_error(fw_multi *m);
int fw_multi_sync(fw_multi *m, int32_t n, const int32_t *ccy, int32_t n_edges, const int32_t *src,
                  const int32_t *dst, const double *val, int32_t want_paths);
/* The last fw_multi_sync once more from the COO that is still on the devices (buildMatrix + runAlgo, nothing
 * crosses PCIe): the benchmark's "inputs resident in HBM" step. */
int fw_multi_resolve(fw_multi *m);
int fw_multi_optimum(fw_multi *m, int32_t src, int32_t dst, double *rate, int32_t *path,
                     int32_t cap, int32_t *path_len);
int fw_multi_download(fw_multi *m, int32_t row0, int32_t rows, double *rate, int32_t *next,
                      int32_t *init_next, int32_t *mid, int32_t *csT, int32_t *rs);
/* floydWarshall in one call on all GPUs (single process): map in COO form in, dense host matrices out. */
int fw_multi_solve_edges(fw_multi *m, int32_t n, const int32_t *ccy, int32_t n_edges,
                         const int32_t *src, const int32_t *dst, const double *val, double *rate,
                         int32_t *next, int32_t *init_next, int32_t *mid, int32_t *csT, int32_t *rs);
/* Dense host matrices in place (runAlgo only): rows go up to their shards, the solve runs, rows come back. */
int fw_multi_solve(fw_multi *m, int32_t n, double *rate, int32_t *next, int32_t *mid, int32_t *csT,
                   int32_t *rs);
/* Resident workflow for benchmarks and tests: allocate the shards for order n, let the caller fill them
 * (fw_multi_upload: host rows -> shards; or device pointers from fw_multi_shard), then solve in place. */
int fw_multi_alloc(fw_multi *m, int32_t n, int32_t want_paths);
int fw_multi_upload(fw_multi *m, int32_t row0, int32_t rows, const double *rate, const int32_t *next);
int fw_multi_solve_resident(fw_multi *m);
/* Local shard i (0 .. nlocal-1): its device, rank, local rows, leading dimension and device pointers.
 * Local row l of rank r is global row  ((l / cyclic_rows) * world + r) * cyclic_rows + l % cyclic_rows. */
typedef struct fw_shard_info {
    int32_t device, rank, world, rows, n_padded, cyclic_rows, group;
    int64_t ld;
    double *d_rate;
    int32_t *d_next;
} fw_shard_info;
int32_t fw_multi_local_shards(fw_multi *m);
/* Local shard i as it lies in HBM (rows x n_padded, local row order) into host buffers; either may be NULL. */
int fw_multi_download_local(fw_multi *m, int32_t i, double *rate, int32_t *next);
/* All local shards at once (rate[i] / next[i] = host buffer of local shard i; arrays or entries may be NULL):
 * the copies of all devices run concurrently. */
int fw_multi_download_locals(fw_multi *m, double *const *rate, int32_t *const *next);
/* How the pivot-row panels travel in this object (a static description). */
const char *fw_multi_transport(fw_multi *m);
int fw_multi_shard(fw_multi *m, int32_t i, fw_shard_info *out);
/* Device time of the last solve (CUDA events, max over the local shards), kernel launches issued, and the
 * bulk kernel's share (per-launch events; only when profiling was on during the solve). */
int fw_multi_last_solve_ms(fw_multi *m, double *ms, int64_t *launches);
int fw_multi_set_profiling(fw_multi *m, int32_t on);
int fw_multi_phase_ms(fw_multi *m, double ms[4], int64_t count[4]);
/* Verification hook (as fw_ctx_set_row_snapshot_sink): when on, every shard keeps, for the pivot rows it
 * owns, the row as its step began; fw_multi_download_sink copies a row range of them to the host. */
int fw_multi_record_row_snapshots(fw_multi *m, int32_t on);
int fw_multi_download_sink(fw_multi *m, int32_t row0, int32_t rows, double *rows_out);

/* The schedule as data: operations of the WHOLE job in issue order, for a matrix of (padded) order n on
 * `world` ranks with k-blocks of `block` pivots in groups of `group`, rows in cyclic blocks of
 * cyclic_rows.  This is what the executor inside fw_multi_* issues; tests replay it with a CPU model of
 * the same operations.  Returns the number of operations (may exceed cap; none written beyond cap), < 0 on a
 * bad layout.  Pure host code: works without a GPU. */
#define FW_OP_PIVOT 1   /* rank's local rows [row_lo, +row_n) = pivot rows b0..: diagonal tile + row panel -> Rw[buf] */
#define FW_OP_APPLY 2   /* rank's local rows [row_lo, +row_n) minus [ex_lo, +ex_n) take k-blocks b0 .. b0+nb*block from Rw[buf..];
                           a row in the blocks' own rows (local rows from grp_lo, block i) takes only blocks i+1.. */
#define FW_OP_BCAST 3   /* Rw[buf] goes from `rank` to every rank (lane B) */
#define FW_OP_A_DONE 4  /* rank: record "main lane done" */
#define FW_OP_WAIT_A 5  /* rank: look-ahead lane waits for the last A_DONE */
#define FW_OP_B_DONE 6  /* rank: record "look-ahead lane done" */
#define FW_OP_WAIT_B 7  /* rank: main lane waits for the last B_DONE */
typedef struct fw_plan_op {
    int32_t kind, rank, lane; /* lane 0 = main (A), 1 = look-ahead (B) */
    int32_t b0, nb, buf;
    int32_t row_lo, row_n, ex_lo, ex_n, grp_lo;
} fw_plan_op;
int64_t fw_multi_plan(int32_t n, int32_t world, int32_t block, int32_t group, int32_t cyclic_rows,
                      fw_plan_op *ops, int64_t cap);

/* Block until everything queued on the context's stream has finished and
 * report any asynchronous failure (incl. FW_ERR_DOMAIN of *_device calls). */
int fw_ctx_synchronize(fw_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FWGPU_H */
