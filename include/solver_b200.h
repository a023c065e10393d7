/* solver_b200.h -- C-ABI of the B200-native sparse direct solver backend.
 *
 * Two layers, both plain C (no torch / C++ types in any signature):
 *
 *  (1) the plugin callbacks, one row in solver_lookup[]; the prototypes mirror
 *      the reference's src/solver_cholmod.h:26-30 / src/solver_umfpack.h and
 *      are called from src/solvers.c:451-452 (init), :479-480 (analyze),
 *      :493-494 (factorize), :564 (evaluate), :461-462 (finalize);
 *
 *  (2) the shim below the wrapper (SURVEY.md 8b "Shim below the wrapper"):
 *      host-C symbolic analysis (b200_analyze*) and the CUDA numeric context
 *      (b200_ctx_*, b200_factor*, b200_solve*) that the callbacks drive and
 *      that tests / bench.py bind with ctypes.
 *
 * There is no CPU fallback for factor/solve: if no CUDA device is usable the
 * numeric entry points return B200_ERR_CUDA and the plugin callbacks abort().
 */
#ifndef SOLVER_B200_H
#define SOLVER_B200_H

#include <stdint.h>
#include "mc_solver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ (1) --
 * plugin callbacks (replace src/solver_cholmod.c:33-342 / solver_umfpack.c) */
void solver_init_b200( solver_state_t* s );
void solver_analyze_b200( solver_state_t* s, matrix_t* A );
void solver_factorize_b200( solver_state_t* s, matrix_t* A );
void solver_evaluate_b200( solver_state_t* s, matrix_t* b, matrix_t* x );
void solver_finalize_b200( solver_state_t* s );

/* capability word advertised in the table row */
#define SOLVER_B200_CAPABILITIES ( SOLVES_FORMAT_CSC | SOLVES_BASE_ZERO | SOLVES_SQUARE_ONLY | \
    SOLVES_UNSYMMETRIC | SOLVES_SYMMETRIC_UPPER_TRIANGULAR | SOLVES_DATA_TYPE_REAL_DOUBLE | SOLVES_RHS_DCOL )

/* ------------------------------------------------------------------ (2) -- */
enum { B200_OK = 0, B200_ERR_ARG = -1, B200_ERR_NOMEM = -2, B200_ERR_CUDA = -3,
       B200_ERR_NOT_SPD = -4, B200_ERR_SINGULAR = -5, B200_ERR_STATE = -6 };

/* factorization kinds */
enum { B200_KIND_CHOLESKY = 0,   /* P A P' = L L'  (symmetric, upper-CSC input)        */
       B200_KIND_LU = 1,         /* P A P' = L U on the pattern of A+A', pivoting in fronts */
       B200_KIND_LDLT = 2 };     /* P A P' = L D L' (symmetric indefinite, upper-CSC input): 1x1 / 2x2 Bunch-Kaufman pivots inside the fronts */

/* ordering methods for b200_analyze_opts */
enum { B200_ORDER_ND = 0, B200_ORDER_NATURAL = 1, B200_ORDER_GIVEN = 2 };

typedef struct b200_options {
  int ordering;          /* B200_ORDER_*                                          */
  const int* user_perm;  /* B200_ORDER_GIVEN: perm[k] = original index of pivot k */
  int nd_leaf;           /* ND recursion stops at this many vertices (default 96) */
  int relax;             /* 1 = relaxed amalgamation (default), 0 = fundamental    */
  int nrelax[3];         /* column thresholds (default 4,16,48)                    */
  double zrelax[3];      /* zero-fraction thresholds (default 0.8,0.1,0.05)        */
  int threads;           /* host threads for the ordering (0 = all)                */
} b200_options_t;

void b200_default_options( b200_options_t* o );

/* Symbolic analysis result: host-resident, plain arrays so that tests can
 * check etree / column counts / supernodes bit-exactly against the oracle. */
typedef struct b200_symbolic {
  int n;
  int kind;              /* B200_KIND_*                                           */
  int64_t nnzA;          /* stored entries of the input CSC                        */
  int* perm;             /* [n]  new -> old                                        */
  int* iperm;            /* [n]  old -> new                                        */
  int* parent;           /* [n]  elimination tree of the permuted matrix (-1 root) */
  int* colcount;         /* [n]  nnz(L(:,j)) incl. diagonal, exact (no relaxation) */
  int64_t nnzL;          /* sum colcount                                           */
  double flops;          /* sum colcount^2 (Cholesky; LU on the same pattern = 2x) */
  int nfund;             /* fundamental supernodes                                 */
  int ns;                /* supernodes after relaxed amalgamation                  */
  int* sn_start;         /* [ns+1] first column of each supernode                  */
  int* sn_parent;        /* [ns]   assembly tree (-1 root)                         */
  int* sn_depth;         /* [ns]   distance to its root (roots: 0)                 */
  int64_t* rowptr;       /* [ns+1] into rowidx                                     */
  int* rowidx;           /* front row lists, sorted; first k_s entries = pivots    */
  int* relidx;           /* same indexing; for i>=k_s: position in parent's front  */
  int64_t* lptr;         /* [ns+1] panel offsets: f_s x k_s, column-major, leading dimension
                            ld_s = f_s rounded up to even (16-byte aligned columns)  */
  int64_t* uptr;         /* [ns+1] LU only: offsets of the U' panels (same layout as L: uptr==lptr) */
  int64_t* amap;         /* [nnzA] destination of every stored input entry: offset
                            into L storage; for LU, entries of the strict upper
                            part map into the U' panels as (offset | 1<<62);
                            -1 = entry ignored                                     */
  int64_t stored_nnzL;   /* lptr[ns]: entries actually stored (with relaxation)    */
  double stored_flops;   /* flops executed on the relaxed structure                */
  int maxfront;          /* largest front order                                    */
  int maxdepth;          /* deepest level                                          */
  int64_t cb_arena[2];   /* doubles needed by the two contribution-block arenas    */
  int64_t* cbptr;        /* [ns] offset of the contribution block in its arena (u_s x u_s,
                            leading dimension u_s rounded up to even)               */
  int64_t* spmv_ptr;     /* [n+1] row pointers of the FULL matrix in the original numbering  */
  int* spmv_col;         /* column of each entry                                              */
  int* spmv_src;         /* index of its value in the input array (residual for refinement)   */
  double t_order, t_symbolic;  /* seconds spent in ordering / rest of the analysis */
} b200_symbolic_t;

/* A: n x n CSC, base 0, unsorted rows allowed, duplicates summed.
 * kind CHOLESKY: only entries with row <= col are used (upper triangle, as the
 * dispatcher delivers for SOLVES_SYMMETRIC_UPPER_TRIANGULAR, src/solvers.c:42-92);
 * kind LU: full matrix. Values are not read ("doesn't care about the actual
 * values", src/solvers.h:95). Returns NULL on failure. */
b200_symbolic_t* b200_analyze( int n, const unsigned int* colptr, const unsigned int* rowidx,
                               int kind, const b200_options_t* opts );
void b200_symbolic_free( b200_symbolic_t* S );

/* standalone pieces of the analysis (exported for parity tests) */
int b200_order_nd( int n, const int* xadj, const int* adj, int leaf, int threads, int* perm );
int b200_etree( int n, const int* colptr, const int* rowidx, int* parent ); /* upper-CSC pattern */

/* ---- numeric context (CUDA, sm_100a) */
typedef struct b200_ctx b200_ctx_t;

b200_ctx_t* b200_ctx_create( int device );       /* NULL if no usable device */
void b200_ctx_destroy( b200_ctx_t* c );
const char* b200_last_error( void );

/* copy the symbolic structure to the device and build the static launch schedule */
int b200_upload_symbolic( b200_ctx_t* c, const b200_symbolic_t* S );

/* numeric factorization; vals = the nnzA stored values in input order.
 * _host: pageable/pinned host pointer (H2D inside); _device: already in HBM. */
int b200_factor_host( b200_ctx_t* c, const double* vals );
int b200_factor_device( b200_ctx_t* c, const double* d_vals );

/* x = A^{-1} b for p right-hand sides, column-major n x p, ld = n */
int b200_solve_host( b200_ctx_t* c, const double* b, int p, double* x );
int b200_solve_device( b200_ctx_t* c, const double* d_b, int p, double* d_x );

/* device memory helpers for callers without a CUDA runtime binding (bench.py, tests) */
void* b200_dev_alloc( b200_ctx_t* c, size_t bytes );
void b200_dev_free( b200_ctx_t* c, void* p );
int b200_dev_upload( b200_ctx_t* c, void* dst, const void* src, size_t bytes );
int b200_dev_download( b200_ctx_t* c, void* dst, const void* src, size_t bytes );
void* b200_host_alloc_pinned( size_t bytes );
void b200_host_free_pinned( void* p );
int b200_dev_sync( b200_ctx_t* c );
int b200_flush_l2( b200_ctx_t* c );             /* write a >L2 buffer between timed steps */


/* ---- one factorization on several GPUs (csrc/host/b200_plan.c; Cholesky only in this round)
 * Every rank runs the same analysis and derives the same plan; rank r then executes only the
 * work it owns.  Supernode s belongs to the contiguous rank group [grp_first[s], +grp_size[s])
 * (proportional mapping of subtree flops); if dist[s] its front columns are dealt out in blocks
 * of level_width[depth] columns, block b to rank grp_first + b mod grp_size, otherwise rank
 * grp_first owns all of it. */
typedef struct b200_plan {
  int world, ns, nlevels;
  int* grp_first;      /* [ns] */
  int* grp_size;       /* [ns] */
  int* dist;           /* [ns] 1: front distributed column-block-cyclically over its group */
  double* work;        /* [ns] flops of the subtree rooted at s (relaxed structure)        */
  int* level_width;    /* [nlevels] outer panel width = column block width of the level    */
} b200_plan_t;
typedef struct b200_xfer { int src, dst, sn; int64_t offset, count; } b200_xfer_t;   /* doubles, inside the child's contribution-block arena */

int b200_outer_width( int maxfront );
void b200_set_outer_width_cap( int cap );   /* 0 = none (default); experiment hook, env B200_NBO_CAP at b200_mg_init */
b200_plan_t* b200_plan_create( const b200_symbolic_t* S, int world, int min_dist_front );
void b200_plan_free( b200_plan_t* pl );
int b200_plan_col_owner( const b200_plan_t* pl, const b200_symbolic_t* S, int s, int c );
int b200_plan_member( const b200_plan_t* pl, int s, int r );
int64_t b200_plan_cb_transfers( const b200_plan_t* pl, const b200_symbolic_t* S, int depth, b200_xfer_t* out, int64_t cap );

/* multi-process mode (one process per GPU, e.g. under torchrun): rank 0 creates two NCCL ids (256 bytes: one communicator for panels, one for bulk factor replication),
 * the launcher hands it to every rank (bench.py: torch.distributed broadcast), and each rank joins
 * BEFORE b200_upload_symbolic.  Panels of distributed fronts and contribution-block columns at subtree
 * boundaries then move with NCCL over NVLink inside b200_factor_*; afterwards every rank holds the
 * whole factor and b200_solve_* works on any of them. */
int b200_mg_unique_id( char* out256 );
int b200_mg_init( b200_ctx_t* c, int rank, int world, const char* id256 );
int b200_mg_rank( b200_ctx_t* c, int* rank, int* world );

/* iterative refinement: r = b - A x is accumulated in double-double on the device, then one more
 * pair of sweeps corrects x.  steps >= 0: exactly that many; steps < 0 (default): automatic -- one
 * step for small systems (n <= 65536), otherwise only while the componentwise backward error
 * max_i |r_i| / (|A||x|+|b|)_i is above 1e-14 (at most 2 steps); in automatic mode the last residual is always formed on the
 * FINAL x, its backward error is reported in b200_stats_t.backward_error, and a non-finite x fails with B200_ERR_SINGULAR */
int b200_set_refinement( b200_ctx_t* c, int steps );

/* statistics of the last factor / solve (device-event milliseconds) */
typedef struct b200_stats {
  double factor_ms;        /* whole factor on the device (events)            */
  double solve_ms;         /* whole solve on the device (events)             */
  double gemm_ms;          /* time in the DMMA trailing-update kernel        */
  double gemm_flops;       /* flops executed by that kernel (whole tiles)    */
  double gemm_alg_flops;   /* its share of the exact factor flops (no tile padding, no relaxation zeros) */
  int64_t launches;        /* kernels launched by the last factor            */
  int64_t solve_launches;  /* kernels launched by the last solve             */
  int64_t gemm_launches;
  double h2d_ms, d2h_ms;
  int64_t device_bytes;    /* bytes of HBM held by the context               */
  int npiv_perturbed;      /* LU: statically perturbed pivots                */
  int refine_steps;        /* refinement steps taken by the last solve       */
  double backward_error;   /* componentwise backward error seen before the last decision (automatic mode), else -1 */
  double first_backward_error;  /* the same quantity for the unrefined solution (automatic mode), else -1 */
  double comm_bytes;       /* multi-GPU: bytes this rank sends + receives over NCCL per factorization (0 on one GPU) */
  double lu_max_multiplier;       /* LU: largest |l(i,j)| over the fully-summed rows of all fronts (1 inside a diagonal block by construction) */
  double pivot_threshold;         /* LU: the threshold u the multipliers are checked against (default 0.1, env B200_PIVOT_THRESHOLD)            */
  int npiv_threshold_violations;  /* LU: multipliers above 1/u, i.e. pivots that threshold partial pivoting over the whole front would have refused */
} b200_stats_t;
int b200_get_stats( b200_ctx_t* c, b200_stats_t* out );
/* per-kernel timing of the next factor (serialises launches; diagnostics only) */
int b200_set_profile( b200_ctx_t* c, int on );
int b200_get_profile( b200_ctx_t* c, char* buf, size_t len );

/* schedule-only context (no device, no CUDA call) and export of the static factor schedule of one rank: lets the CPU
 * tests replay all ranks' streams / events / messages and prove the multi-GPU schedule deadlock-free
 * (tests/test_multi_gpu_schedule.py).  stream: 0 update, 1 panel, 2 NCCL (panels, contribution blocks), 3 NCCL (bulk). */
b200_ctx_t* b200_ctx_create_dry( int rank, int world );
typedef struct b200_sched_entry { int kind, stream, wait_ev, record_ev, depth; int64_t first_msg, nmsg, count;
  double flops, kavg, bytes;   /* schedule-only contexts: executed flops and flop-weighted K of a DMMA launch, bytes written by memset / extend-add (cost models) */
} b200_sched_entry_t;
typedef struct b200_msg { int src, dst, which; int64_t offset, count; } b200_msg_t;   /* which: 0 L, 1 inverse blocks, 2/3 contribution-block arena 0/1 */
int64_t b200_debug_schedule( b200_ctx_t* c, b200_sched_entry_t* out, int64_t cap );
int64_t b200_debug_messages( b200_ctx_t* c, b200_msg_t* out, int64_t cap );
const char* b200_debug_kind_name( int kind );
typedef struct b200_access { int launch, which, write; int64_t offset, count; } b200_access_t;   /* conservative interval a compute launch reads (0) or writes (1) */
int64_t b200_debug_accesses( b200_ctx_t* c, b200_access_t* out, int64_t cap );
/* the ordered unit lists of the persistent sweeps (p <= 4; csrc/cuda/b200_sweep.cuh), schedule-only contexts: dir 0 forward / 1 backward,
 * 16 ints per unit (kind, task, o0, aux, sn, w1, n1, v1, w2, v2, sig, up, 0, 0, 0, 0); tasks: 8 int64 each (vsel, voff, K, dsel, doff,
 * N, tri, mode); number of dependency counters.  tests/sweep_sim.py replays them on the CPU (deadlock freedom, data races). */
int64_t b200_debug_sweep_units( b200_ctx_t* c, int dir, int* out, int64_t cap );
int64_t b200_debug_sweep_tasks( b200_ctx_t* c, int64_t* out, int64_t cap );
int64_t b200_debug_sweep_counters( b200_ctx_t* c );
/* per-unit time stamps of the last persistent sweep (context created with B200_SWEEP_TRACE set): 5 uint64 per unit (globaltimer ns at the
 * start, clock64 at start / inputs ready / done, SM id) and the device unit list (16 ints per unit; small supernodes in batches of 8) */
int64_t b200_debug_sweep_trace( b200_ctx_t* c, int dir, unsigned long long* trace, int* units, int64_t cap );

/* checksum of the factor: colsq[j] = sum_i L(i,j)^2 over the stored rows of column j (permuted numbering), n doubles on the host;
 * fixed reduction order, so identical factors give identical bits (Cholesky: the total equals trace(A)) */
int b200_factor_colnorms( b200_ctx_t* c, double* colsq );

/* diagnostics for the parity tests: read back factor storage (0 = L panels, 1 = U' panels, 2 = inverse blocks) */
int b200_debug_read( b200_ctx_t* c, int which, double* dst, int64_t offset, int64_t count );

/* micro-benchmarks used to establish the FP64 roof on the box */
double b200_bench_dgemm_tflops( b200_ctx_t* c, int m, int n, int k, int iters ); /* our DMMA kernel */
double b200_bench_dgemm_bulk_variant( b200_ctx_t* c, int variant, int m, int n, int k, int iters );   /* the same for the bulk-copy kernel */
double b200_bench_dmma_occupancy( b200_ctx_t* c, int warps_per_sm, int ilp );   /* DMMA rate vs resident warps / independent accumulators */
double b200_bench_fp64_peak_tflops( b200_ctx_t* c, int which, int iters );  /* 0: DMMA.8x8x4 issue roof, 1: DFMA roof */
double b200_bench_stream_gbs( b200_ctx_t* c, size_t bytes, int iters );
double b200_bench_potrf_us( b200_ctx_t* c, int which, int nb, int ntasks, int iters );   /* diagonal-block kernels: 0 unrolled/registers, 1 rolled/shared memory */

#ifdef __cplusplus
}
#endif
#endif
