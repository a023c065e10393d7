/* b200_numeric.cu -- CUDA context, static launch schedule, factor and solve drivers.
 *
 * extern "C" shim below the plugin wrapper: the host-C analysis
 * hands over a b200_symbolic_t; everything that depends only on the pattern
 * (task lists, tile lists, launch order) is built ONCE here at upload time, so
 * that factorize (src/solvers.c:485-495 -> table.factorize) and evaluate
 * (src/solvers.c:500-611 -> table.evaluate) are pure replays of kernel launches
 * with no host-side decisions and no device->host round trips inside the loop.
 *
 * Data layout in HBM (all FP64, column-major):
 *   L      : one f_s x k_s panel per supernode at lptr[s]            (permanent)
 *   U'     : LU only, same layout as L (upper factor stored transposed)
 *   CB[2]  : contribution blocks u_s x u_s of one tree level, double-buffered
 *            by level parity; level d reads CB[(d+1)&1], writes CB[d&1]
 *   inv    : inverse of every <=64x64 diagonal block (and of U's for LU)
 */
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <dlfcn.h>
#include <nccl.h>      /* types only: the library is bound at run time (the copy torch has already loaded, else the system one) */
#include "solver_b200.h"
#include "b200_kernels.cuh"
#include "b200_sweep.cuh"

using namespace b200;

static thread_local char g_err[512] = "";
static int set_err( const char* fmt, const char* a = "", const char* b = "" ) { snprintf( g_err, sizeof( g_err ), fmt, a, b ); return B200_ERR_CUDA; }
extern "C" const char* b200_last_error( void ) { return g_err; }

#define CK( call ) do { cudaError_t e_ = ( call ); if ( e_ != cudaSuccess ) { set_err( "%s: %s", #call, cudaGetErrorString( e_ ) ); return B200_ERR_CUDA; } } while ( 0 )
#define CKN( call ) do { cudaError_t e_ = ( call ); if ( e_ != cudaSuccess ) { set_err( "%s: %s", #call, cudaGetErrorString( e_ ) ); return nullptr; } } while ( 0 )


/* ------------------------------------------------------------ NCCL binding
 * Resolved with dlopen/dlsym so that libmeagre_crowd_b200.so has no link-time dependency on NCCL:
 * a single-GPU user never loads it, and under torchrun the soname resolves to the copy torch
 * brought in (one NCCL per process). */
struct NcclApi {
  void* h = nullptr;
  ncclResult_t ( *GetUniqueId )( ncclUniqueId* ) = nullptr;
  ncclResult_t ( *CommInitRank )( ncclComm_t*, int, ncclUniqueId, int ) = nullptr;
  ncclResult_t ( *CommDestroy )( ncclComm_t ) = nullptr;
  ncclResult_t ( *Send )( const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t ) = nullptr;
  ncclResult_t ( *Recv )( void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t ) = nullptr;
  ncclResult_t ( *GroupStart )() = nullptr;
  ncclResult_t ( *GroupEnd )() = nullptr;
  const char* ( *GetErrorString )( ncclResult_t ) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if ( g_nccl.h ) return 0;
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  void* h = nullptr;
  for ( const char* nm : names ) { h = dlopen( nm, RTLD_NOW | RTLD_GLOBAL ); if ( h ) break; }
  if ( !h ) { set_err( "cannot load NCCL: %s", dlerror() ); return -1; }
#define B200_SYM( field, name ) g_nccl.field = reinterpret_cast<decltype( g_nccl.field )>( dlsym( h, name ) ); if ( !g_nccl.field ) { set_err( "NCCL lacks %s", name ); return -1; }
  B200_SYM( GetUniqueId, "ncclGetUniqueId" ) B200_SYM( CommInitRank, "ncclCommInitRank" ) B200_SYM( CommDestroy, "ncclCommDestroy" )
  B200_SYM( Send, "ncclSend" ) B200_SYM( Recv, "ncclRecv" ) B200_SYM( GroupStart, "ncclGroupStart" ) B200_SYM( GroupEnd, "ncclGroupEnd" )
  B200_SYM( GetErrorString, "ncclGetErrorString" )
#undef B200_SYM
  g_nccl.h = h;
  return 0;
}
#define NK( call ) do { ncclResult_t r_ = ( call ); if ( r_ != ncclSuccess ) { set_err( "%s: %s", #call, g_nccl.GetErrorString( r_ ) ); return B200_ERR_CUDA; } } while ( 0 )

/* outer panel width = K of the trailing update; chosen per tree level from its largest front */
static inline int outer_width( int maxfront ) { return b200_outer_width( maxfront ); }
constexpr int ASM_COLS = 32;  /* destination columns per extend-add CTA */
constexpr int GEMM_STAGES = 4;
/* production tile of the DMMA update kernel: 64x64 per CTA, 8 warps of 32x16, 4-stage bulk-copy pipeline.
 * 69 KB of shared memory and <=80 registers keep 3 CTAs (24 warps) resident per SM, which measured
 * 28-31 TFLOP/s against 17-23 for a 128x128 tile with one CTA per SM (profiles/r1_gemm_variants.txt). */
#define GEMM_CFG 64, 64, 32, 16, GEMM_STAGES
constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_THREADS = 256;
constexpr int BULK_SMEM = GEMM_STAGES * 16 * ( 68 + 68 ) * 8 + 2 * GEMM_STAGES * 8;   /* k-major A and B tiles of every stage + full/empty mbarriers */
constexpr int BULK_THREADS = GEMM_THREADS + 32;             /* + the producer warp */
constexpr int SVG_SMEM_NT = GEMM_STAGES * 16 * ( 68 + 68 ) * 8 + 2 * GEMM_STAGES * 8;                 /* gemm_solve_dmma_kernel<false> */
constexpr int SVG_SMEM_BT = 3 * ( 32 * 68 + 64 * 36 ) * 8 + 2 * 3 * 8;                                /* gemm_solve_dmma_kernel<true>: BK = 32, 3 stages */

enum LaunchKind { LK_MEMSET_CB = 0, LK_ASM, LK_DIAG, LK_TRSM, LK_GEMM_BIG, LK_TRSM_MMA, LK_NOP, LK_DIAG_S, LK_SMALL_FRONT, LK_XFER_CB, LK_BCAST, LK_GATHER, LK_COUNT };
static const char* kind_name[LK_COUNT] = { "memset_cb", "extend_add", "diag_block", "trsm_rows", "gemm_dmma", "trsm_dmma", "nop", "diag_block_small", "small_front", "nccl_cb_columns", "nccl_panel_bcast", "nccl_factor_gather" };
/* one point-to-point message of the multi-GPU schedule: which = 0 L storage, 1 inverse blocks, 2/3 contribution-block arena 0/1 */
struct Msg { int src, dst, which; int64_t off, cnt; };

/* stream 0 = update stream (also level prologues), stream 1 = panel stream (look-ahead), stream 2 = NCCL stream, 3 = bulk NCCL stream, 4 = arena zeroing;
 * wait_ev / record_ev index ctx->events (-1: none) */
struct Launch { int kind; int64_t first; int64_t count; int depth; double flops; int stream = 0; int wait_ev = -1; int record_ev = -1; };
static double g_alg_flops = 0;   /* accumulates minimal flops of the gemm tasks while a schedule is built */

/* solve schedule */
enum SolveKind { SK_FWD_ZERO_W = 0, SK_FWD_GATHER, SK_FWD_TRI, SK_FWD_UPD, SK_BWD_GATHER, SK_BWD_UPD, SK_BWD_TRI, SK_FWD_SMALL, SK_BWD_SMALL, SK_SWEEP_FWD, SK_SWEEP_BWD, SK_COUNT };
/* one launch of a sweep: gather / zero launches use first,count; product launches carry their tiles in the FMA list (t0,nt: 32 outputs per
 * tile for gemv-type tasks, 32 for gemvT-type) and in the tensor-path list (g0,ng: 64 outputs per tile) */
struct SLaunch { int kind; int depth; int64_t first, count; int64_t t0, nt; int64_t g0, ng; };

static thread_local bool g_dry = false;   /* schedule-only mode: no CUDA call at all (b200_ctx_create_dry) */
struct DryGuard { bool old; explicit DryGuard( bool v ) : old( g_dry ) { g_dry = v; } ~DryGuard() { g_dry = old; } };
template <class T> struct DevVec {
  T* d = nullptr; size_t n = 0;
  int upload( const std::vector<T>& h ) {
    n = h.size();
    if ( n == 0 || g_dry ) return 0;
    CK( cudaMalloc( &d, n * sizeof( T ) ) );
    CK( cudaMemcpy( d, h.data(), n * sizeof( T ), cudaMemcpyHostToDevice ) );
    return 0;
  }
  void release() { if ( d && !g_dry ) cudaFree( d ); d = nullptr; n = 0; }
  size_t bytes() const { return n * sizeof( T ); }
};

struct b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr, stream4 = nullptr, stream5 = nullptr;   /* stream5: zeroing of the next level's contribution-block arena, under the current level's updates */
  int rank = 0, world = 1; ncclComm_t comm = nullptr, comm2 = nullptr;   /* comm/stream3: latency-critical panel and contribution-block traffic; comm2/stream4: bulk factor replication */ b200_plan_t* plan = nullptr; std::vector<Msg> msgs; double comm_bytes = 0;
  std::vector<cudaEvent_t> events; int nevents = 0;
  cudaDeviceProp prop;
  /* symbolic copy */
  int n = 0, ns = 0, kind = 0, maxdepth = 0;
  int64_t nnzA = 0, lsize = 0, invsize = 0;
  int64_t cb_arena[2] = { 0, 0 };
  std::vector<int64_t> level_cb;          /* doubles of CB used by each level */
  std::vector<int64_t> level_w;           /* solve workspace rows per level */
  int64_t w_rows[2] = { 0, 0 };
  double flops_exact = 0, flops_stored = 0, gemm_alg_flops = 0;
  /* device arrays */
  double *dL = nullptr, *dU = nullptr, *dCB[2] = { nullptr, nullptr }, *dInv = nullptr, *dInv2 = nullptr, *dInv3 = nullptr, *dVals = nullptr;
  int* dPiv = nullptr; int* dErr = nullptr; int* dPert = nullptr; double* dZeros = nullptr; double* dTiny = nullptr; unsigned long long* dMaxBits = nullptr; int* dViol = nullptr; double pivot_threshold = 0.1; bool use_small_front = true, replicate = false; DevVec<int> small_sns;
  DevVec<int64_t> amap, spmv_ptr; DevVec<int> perm, rowidx, relidx, child_list, spmv_col, spmv_src;
  int refine = -1;            /* -1: automatic (see b200_set_refinement) */
  unsigned long long* dOmega = nullptr;
  double *dR = nullptr, *dD = nullptr;
  DevVec<SnMeta> sn;
  DevVec<GemmTask> gemm_tasks; DevVec<Tile> tiles_big, tiles_small;
  DevVec<DiagTask> diag_tasks; DevVec<TrsmTask> trsm_tasks; DevVec<RowTile> trsm_tiles;
  DevVec<AsmTask> asm_tasks;
  std::vector<Launch> launches;
  /* solve */
  DevVec<SvTask> sv_tasks; DevVec<SvTile> sv_tiles_v, sv_tiles_t, sv_tiles_g; DevVec<GatherTask> bgather_tasks; DevVec<int> sv_small;
  SvGemm* dSvGemm = nullptr; int sv_gemm_p = 0; int solve_wide_min = SB_WIDE_MIN;
  std::vector<SLaunch> slaunches;
  /* persistent sweeps (p <= 4): ordered unit lists with device-side dependencies, b200_sweep.cuh */
  DevVec<SvUnit> sw_units[2]; /* forward, backward */ DevVec<SvTask> sw_tasks; DevVec<SvUnit> sw_inner; int64_t sw_w_rows = 0; int64_t sw_ncnt = 0; int* dSwCnt = nullptr; double* dWg = nullptr; unsigned long long* dSwTrace = nullptr; bool use_sweep = true; int sweep_pmax = 1;      /* measured on 128^3: 14.0 ms against 14.4 (level-set launches) at p = 1, but 17.4 / 27 against 14.5 / 20.4 at p = 2 / 4 -- B200_SWEEP_PMAX widens it for the tests */ int64_t sw_head = 0; int sw_occ[3] = { 0, 0, 0 }; int sw_err_h = 0; std::vector<SvUnit> h_sw_f, h_sw_b; std::vector<SvTask> h_sw_t;
  int64_t sinvsize = 0;
  DevVec<GatherTask> gather_tasks;
  double* dZ = nullptr;
  double *dSinvF = nullptr, *dSinvB = nullptr, *dSinvTmp = nullptr;   /* solve-block inverses: X (forward), X' (backward) */
  DevVec<SinvTask> sinv_tasks;
  double *dY = nullptr, *dW[2] = { nullptr, nullptr }, *dB = nullptr, *dX = nullptr;
  int solve_p_cap = 0, io_p_cap = 0;
  void* flush_buf = nullptr; size_t flush_bytes = 0;
  bool factored = false;
  bool profile = false;
  bool dry = false;           /* no device: the context only builds (and exports) the static schedule */
  /* host copies of the task lists, kept only by a schedule-only context (b200_debug_accesses) */
  std::vector<GemmTask> h_gt; std::vector<Tile> h_tb; std::vector<DiagTask> h_dt; std::vector<TrsmTask> h_tt; std::vector<RowTile> h_ttl; std::vector<AsmTask> h_at; std::vector<int> h_small, h_child; std::vector<SnMeta> h_meta;
  double prof_ms[LK_COUNT] = { 0 }; int64_t prof_n[LK_COUNT] = { 0 };
  double sprof_ms[SK_COUNT] = { 0 }; int64_t sprof_n[SK_COUNT] = { 0 };
  b200_stats_t stats;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr, evs = nullptr, evs2 = nullptr;
  std::vector<cudaEvent_t> lvl_ev; std::vector<float> lvl_ms;   /* update-stream time stamps at level boundaries (diagnostics) */
  int64_t device_bytes = 0;
};

/* ------------------------------------------------------------ context */
extern "C" b200_ctx_t* b200_ctx_create( int device ) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount( &ndev );
  if ( e != cudaSuccess || ndev == 0 ) { set_err( "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString( e ) : "device count is 0" ); return nullptr; }
  if ( device < 0 || device >= ndev ) { set_err( "bad device index%s", "" ); return nullptr; }
  CKN( cudaSetDevice( device ) );
  b200_ctx* c = new b200_ctx();
  c->device = device;
  CKN( cudaGetDeviceProperties( &c->prop, device ) );
  CKN( cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) ); { int lo = 0, hi = 0; CKN( cudaDeviceGetStreamPriorityRange( &lo, &hi ) );   /* the panel chain is short and latency-critical: it must win free SM slots */
    CKN( cudaStreamCreateWithPriority( &c->stream2, cudaStreamNonBlocking, hi ) ); CKN( cudaStreamCreateWithPriority( &c->stream3, cudaStreamNonBlocking, hi ) ); CKN( cudaStreamCreateWithPriority( &c->stream4, cudaStreamNonBlocking, lo ) ); CKN( cudaStreamCreateWithPriority( &c->stream5, cudaStreamNonBlocking, lo ) ); }
  CKN( cudaEventCreate( &c->ev0 ) ); CKN( cudaEventCreate( &c->ev1 ) ); CKN( cudaEventCreate( &c->evp0 ) ); CKN( cudaEventCreate( &c->evp1 ) ); CKN( cudaEventCreateWithFlags( &c->evs, cudaEventDisableTiming ) ); CKN( cudaEventCreateWithFlags( &c->evs2, cudaEventDisableTiming ) );
  CKN( cudaMalloc( &c->dErr, sizeof( int ) ) ); CKN( cudaMalloc( &c->dPert, sizeof( int ) ) ); CKN( cudaMalloc( &c->dOmega, sizeof( unsigned long long ) ) ); CKN( cudaMalloc( &c->dTiny, 8 ) ); CKN( cudaMalloc( &c->dMaxBits, 8 ) ); CKN( cudaMalloc( &c->dViol, sizeof( int ) ) );
  { const char* e = getenv( "B200_PIVOT_THRESHOLD" ); if ( e && atof( e ) > 0.0 && atof( e ) <= 1.0 ) c->pivot_threshold = atof( e ); }
  CKN( cudaFuncSetAttribute( gemm_nt_dmma_bulk_kernel<GEMM_CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, BULK_SMEM ) );
  CKN( cudaMalloc( &c->dZeros, 128 * 8 ) ); CKN( cudaMemset( c->dZeros, 0, 128 * 8 ) );
  CKN( cudaFuncSetAttribute( getrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GETRF_SMEM ) );
  CKN( cudaFuncSetAttribute( ldlt_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LDLT_SMEM ) );
  CKN( cudaFuncSetAttribute( potrf_diag_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_S_SMEM ) );
  CKN( cudaFuncSetAttribute( small_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM ) );
  c->use_small_front = getenv( "B200_NO_SMALL_FRONT" ) == nullptr;
  c->replicate = getenv( "B200_REPLICATE_FACTOR" ) != nullptr;
  { const char* e = getenv( "B200_SWEEP_PMAX" ); if ( e ) c->sweep_pmax = std::max( 0, std::min( 4, atoi( e ) ) ); }
  CKN( cudaFuncSetAttribute( trsm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM ) );
  CKN( cudaFuncSetAttribute( build_sinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SINV_SMEM ) );
  CKN( cudaFuncSetAttribute( sv_gemv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sv_gemv_smem<4>() ) );
  CKN( cudaFuncSetAttribute( sv_gemvT_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sv_gemvT_smem<4>() ) );
  CKN( cudaFuncSetAttribute( gemm_solve_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SVG_SMEM_NT ) );
  CKN( cudaFuncSetAttribute( gemm_solve_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SVG_SMEM_BT ) );
  { const char* e = getenv( "B200_SOLVE_WIDE_MIN" ); if ( e ) c->solve_wide_min = std::max( 2 * SB, atoi( e ) ); }     /* test hook: wide solve blocks on small problems */
  CKN( cudaFuncSetAttribute( sv_sweep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_smem<1>() ) ); CKN( cudaFuncSetAttribute( sv_sweep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_smem<2>() ) ); CKN( cudaFuncSetAttribute( sv_sweep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_smem<4>() ) );
  CKN( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &c->sw_occ[0], sv_sweep_kernel<1>, SW_THREADS, sw_smem<1>() ) ); CKN( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &c->sw_occ[1], sv_sweep_kernel<2>, SW_THREADS, sw_smem<2>() ) ); CKN( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &c->sw_occ[2], sv_sweep_kernel<4>, SW_THREADS, sw_smem<4>() ) );
  memset( &c->stats, 0, sizeof( c->stats ) );
  return c;
}


/* ------------------------------------------------------------ multi-process mode */
extern "C" int b200_mg_unique_id( char* out256 ) {
  if ( !out256 ) return B200_ERR_ARG;
  if ( load_nccl() ) return B200_ERR_CUDA;
  static_assert( sizeof( ncclUniqueId ) == 128, "NCCL id is 128 bytes" );
  for ( int k = 0; k < 2; k++ ) { ncclUniqueId id; NK( g_nccl.GetUniqueId( &id ) ); memcpy( out256 + 128 * k, &id, 128 ); }   /* one id per communicator */
  return B200_OK;
}
extern "C" int b200_mg_init( b200_ctx_t* c, int rank, int world, const char* id256 ) {
  if ( !c || world < 1 || rank < 0 || rank >= world ) return B200_ERR_ARG;
  if ( c->dL ) { set_err( "b200_mg_init must precede b200_upload_symbolic%s", "" ); return B200_ERR_STATE; }
  CK( cudaSetDevice( c->device ) );
  if ( c->comm ) { g_nccl.CommDestroy( c->comm ); c->comm = nullptr; }
  if ( c->comm2 ) { g_nccl.CommDestroy( c->comm2 ); c->comm2 = nullptr; }
  c->rank = rank; c->world = world;
  { const char* e = getenv( "B200_NBO_CAP" ); b200_set_outer_width_cap( e ? atoi( e ) : 0 ); }     /* measured at 4 GPUs: 256 -> 273 ms, 512 -> 266 ms: the chain is bound by per-block costs, not by the width */
  if ( world == 1 ) return B200_OK;
  if ( !id256 ) return B200_ERR_ARG;
  if ( load_nccl() ) return B200_ERR_CUDA;
  ncclUniqueId id; memcpy( &id, id256, 128 );
  NK( g_nccl.CommInitRank( &c->comm, world, id, rank ) );
  memcpy( &id, id256 + 128, 128 );
  NK( g_nccl.CommInitRank( &c->comm2, world, id, rank ) );
  return B200_OK;
}
extern "C" int b200_mg_rank( b200_ctx_t* c, int* rank, int* world ) { if ( !c ) return B200_ERR_ARG; if ( rank ) *rank = c->rank; if ( world ) *world = c->world; return B200_OK; }

/* A context without a device: b200_upload_symbolic then builds the complete launch / message schedule of rank `rank`
 * of `world` with placeholder addresses and no CUDA call, and b200_debug_schedule / b200_debug_messages export it.
 * tests/test_multi_gpu_schedule.py replays the schedules of all ranks on the CPU (streams, events, message matching)
 * to prove that they cannot deadlock.  Factor / solve entry points refuse to run on such a context. */
extern "C" b200_ctx_t* b200_ctx_create_dry( int rank, int world ) {
  if ( world < 1 || rank < 0 || rank >= world ) return nullptr;
  b200_ctx* c = new b200_ctx();
  c->dry = true; c->rank = rank; c->world = world; c->replicate = getenv( "B200_REPLICATE_FACTOR" ) != nullptr;
  memset( &c->stats, 0, sizeof( c->stats ) );
  memset( &c->prop, 0, sizeof( c->prop ) ); c->prop.multiProcessorCount = 148;
  c->use_small_front = getenv( "B200_NO_SMALL_FRONT" ) == nullptr;
  { const char* e = getenv( "B200_SOLVE_WIDE_MIN" ); if ( e ) c->solve_wide_min = std::max( 2 * SB, atoi( e ) ); }
  return c;
}

static void release_symbolic( b200_ctx* c ) {
  DryGuard dg( c->dry );
  if ( c->dry ) {
    c->dL = c->dU = c->dCB[0] = c->dCB[1] = c->dInv = c->dInv2 = c->dInv3 = c->dVals = nullptr; c->dPiv = nullptr; c->dSinvF = c->dSinvB = c->dSinvTmp = nullptr;
    c->launches.clear(); c->slaunches.clear(); c->nevents = 0; c->msgs.clear(); if ( c->plan ) { b200_plan_free( c->plan ); c->plan = nullptr; }
    return;
  }
  cudaFree( c->dL ); if ( c->dU ) cudaFree( c->dU ); cudaFree( c->dCB[0] ); cudaFree( c->dCB[1] ); cudaFree( c->dInv ); cudaFree( c->dInv2 ); cudaFree( c->dInv3 ); c->dInv3 = nullptr; cudaFree( c->dVals ); cudaFree( c->dPiv );
  cudaFree( c->dSinvF ); cudaFree( c->dSinvB ); cudaFree( c->dSinvTmp ); c->dSinvF = c->dSinvB = c->dSinvTmp = nullptr; c->sinv_tasks.release();
  c->dL = c->dU = c->dCB[0] = c->dCB[1] = c->dInv = c->dInv2 = c->dVals = nullptr; c->dPiv = nullptr;
  c->amap.release(); c->spmv_ptr.release(); c->spmv_col.release(); c->spmv_src.release(); c->perm.release(); c->rowidx.release(); c->relidx.release(); c->child_list.release(); c->sn.release();
  c->gemm_tasks.release(); c->tiles_big.release(); c->tiles_small.release(); c->diag_tasks.release(); c->trsm_tasks.release(); c->trsm_tiles.release(); c->asm_tasks.release();
  c->small_sns.release(); c->sv_tasks.release(); c->sv_tiles_v.release(); c->sv_tiles_t.release(); c->sv_tiles_g.release(); c->gather_tasks.release(); c->bgather_tasks.release(); c->sv_small.release();
  cudaFree( c->dY ); cudaFree( c->dW[0] ); cudaFree( c->dW[1] ); cudaFree( c->dB ); cudaFree( c->dX ); cudaFree( c->dR ); cudaFree( c->dD ); cudaFree( c->dZ ); cudaFree( c->dSvGemm );
  c->dR = c->dD = c->dZ = nullptr; c->dSvGemm = nullptr; c->sv_gemm_p = 0;
  c->sw_tasks.release(); for ( int i = 0; i < 2; i++ ) c->sw_units[i].release(); c->sw_inner.release(); cudaFree( c->dSwCnt ); cudaFree( c->dWg ); cudaFree( c->dSwTrace ); c->dSwCnt = nullptr; c->dWg = nullptr; c->dSwTrace = nullptr;
  c->dY = c->dW[0] = c->dW[1] = c->dB = c->dX = nullptr; c->solve_p_cap = 0; c->io_p_cap = 0;
  c->launches.clear(); c->slaunches.clear(); c->nevents = 0; c->msgs.clear(); if ( c->plan ) { b200_plan_free( c->plan ); c->plan = nullptr; }
  c->factored = false; c->device_bytes = 0;
}

extern "C" void b200_ctx_destroy( b200_ctx_t* c ) {
  if ( !c ) return;
  if ( c->dry ) { release_symbolic( c ); delete c; return; }
  cudaSetDevice( c->device );
  cudaStreamSynchronize( c->stream ); cudaStreamSynchronize( c->stream2 ); cudaStreamSynchronize( c->stream3 ); cudaStreamSynchronize( c->stream4 ); cudaStreamSynchronize( c->stream5 );   /* pending NCCL sends read dL / dInv */
  release_symbolic( c );
  cudaFree( c->dErr ); cudaFree( c->dPert ); cudaFree( c->dOmega ); cudaFree( c->dTiny ); cudaFree( c->dMaxBits ); cudaFree( c->dViol ); cudaFree( c->dZeros ); if ( c->flush_buf ) cudaFree( c->flush_buf );
  cudaEventDestroy( c->ev0 ); cudaEventDestroy( c->ev1 ); cudaEventDestroy( c->evp0 ); cudaEventDestroy( c->evp1 ); cudaEventDestroy( c->evs ); cudaEventDestroy( c->evs2 );
  for ( cudaEvent_t e : c->events ) cudaEventDestroy( e );
  for ( cudaEvent_t e : c->lvl_ev ) cudaEventDestroy( e );
  if ( c->comm ) g_nccl.CommDestroy( c->comm );
  if ( c->comm2 ) g_nccl.CommDestroy( c->comm2 );
  if ( c->plan ) b200_plan_free( c->plan );
  cudaStreamDestroy( c->stream ); cudaStreamDestroy( c->stream2 ); cudaStreamDestroy( c->stream3 ); cudaStreamDestroy( c->stream4 ); cudaStreamDestroy( c->stream5 );
  delete c;
}

/* ------------------------------------------------- schedule construction */
static void add_gemm( std::vector<GemmTask>& tasks, std::vector<Tile>& big, std::vector<Tile>& small, double& flops,
                      double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int N, int K, int lower, int diag ) {
  if ( M <= 0 || N <= 0 || K <= 0 ) return;
  GemmTask t; t.C = C; t.A = A; t.B = B; t.ldc = ldc; t.lda = lda; t.ldb = ldb; t.M = M; t.N = N; t.K = K; t.lower = lower; t.diag = diag; t.mode = 0;
  const int id = ( int )tasks.size();
  tasks.push_back( t );
  if ( lower ) { const double nn = std::min( N, M + diag ); g_alg_flops += ( double )K * ( nn * ( nn + 1 ) + 2.0 * ( ( double )M + diag - nn ) * nn ); }
  else g_alg_flops += 2.0 * M * N * K;
  const int bm = GEMM_BM, bn = GEMM_BN;
  std::vector<Tile>& tl = big; ( void )small;
  for ( int tn = 0; tn * bn < N; tn++ )
    for ( int tm = 0; tm * bm < M; tm++ ) {
      if ( lower ) { /* skip tiles strictly above the stored triangle: max row + diag < min col */
        const int rmax = std::min( M, ( tm + 1 ) * bm ) - 1, cmin = tn * bn;
        if ( rmax + diag < cmin ) continue;
      }
      Tile x; x.task = id; x.tm = ( unsigned short )tm; x.tn = ( unsigned short )tn; tl.push_back( x );
      const int mm = std::min( bm, M - tm * bm ), nn = std::min( bn, N - tn * bn );
      flops += 2.0 * mm * nn * K;
    }
}

extern "C" int b200_upload_symbolic( b200_ctx_t* c, const b200_symbolic_t* S ) {
  if ( !c || !S ) return B200_ERR_ARG;
  DryGuard dg( c->dry );
  if ( !c->dry ) { CK( cudaSetDevice( c->device ) ); CK( cudaStreamSynchronize( c->stream ) ); }
  release_symbolic( c );
  const int ns = S->ns, n = S->n;
  const bool LU = S->kind == B200_KIND_LU;
  const bool LD = S->kind == B200_KIND_LDLT;      /* L D L': Cholesky's structure and updates, but a second panel W = L D in the place of U' and pivoting diagonal blocks */
  const bool TWO = LU || LD;                       /* two panels per supernode */
  c->n = n; c->ns = ns; c->kind = S->kind; c->maxdepth = S->maxdepth; c->nnzA = S->nnzA;
  c->lsize = S->stored_nnzL; c->cb_arena[0] = S->cb_arena[0]; c->cb_arena[1] = S->cb_arena[1];
  c->flops_exact = S->flops * ( LU ? 2.0 : 1.0 ); c->flops_stored = S->stored_flops * ( LU ? 2.0 : 1.0 );

  /* ---- per-supernode metadata, children lists, level lists */
  std::vector<SnMeta> meta( ns );
  std::vector<int> child_cnt( ns + 1, 0 ), child_list;
  for ( int s = 0; s < ns; s++ ) if ( S->sn_parent[s] >= 0 ) child_cnt[S->sn_parent[s] + 1]++;
  for ( int s = 0; s < ns; s++ ) child_cnt[s + 1] += child_cnt[s];
  child_list.resize( child_cnt[ns] );
  { std::vector<int> w( child_cnt.begin(), child_cnt.end() - 1 ); for ( int s = 0; s < ns; s++ ) if ( S->sn_parent[s] >= 0 ) child_list[w[S->sn_parent[s]]++] = s; }
  std::vector<std::vector<int>> levels( S->maxdepth + 1 );
  c->level_cb.assign( S->maxdepth + 1, 0 ); c->level_w.assign( S->maxdepth + 1, 0 );
  int64_t invoff = 0, sinvoff = 0; c->sw_w_rows = 0;
  for ( int s = 0; s < ns; s++ ) {
    SnMeta& m = meta[s];
    m.lptr = S->lptr[s]; m.uptr = S->uptr[s]; m.cbptr = S->cbptr[s]; m.rowptr = S->rowptr[s];
    m.c0 = S->sn_start[s]; m.k = S->sn_start[s + 1] - S->sn_start[s]; m.f = ( int )( S->rowptr[s + 1] - S->rowptr[s] );
    m.parent = S->sn_parent[s]; m.depth = S->sn_depth[s]; m.child_begin = child_cnt[s]; m.child_end = child_cnt[s + 1]; m.ld = ( m.f + 1 ) & ~1;
    m.invptr = invoff; invoff += ( int64_t )( m.k / NB ) * NB * NB + ( int64_t )( m.k % NB ) * ( m.k % NB );
    m.sbw = m.k >= c->solve_wide_min ? SB_WIDE : SB;
    { const int64_t full = m.k / m.sbw, rem = m.k % m.sbw;      /* block t of the supernode at m.sinv + t * sbw^2, leading dimension = its width rounded up to even */
      m.sinv = sinvoff; sinvoff += full * m.sbw * m.sbw + ( ( rem + 1 ) & ~( int64_t )1 ) * rem; }
    const int d = m.depth; const int64_t u = m.f - m.k;
    levels[d].push_back( s );
    c->level_cb[d] = std::max( c->level_cb[d], m.cbptr + ( ( u + 1 ) & ~( int64_t )1 ) * u );
    m.wptr = c->level_w[d]; c->level_w[d] += u;
    m.wgl = c->sw_w_rows; c->sw_w_rows += u;
  }
  c->invsize = invoff; c->sinvsize = sinvoff;
  if ( c->world > 1 ) {
    if ( TWO ) { set_err( "multi-GPU factorization is built for Cholesky only (LU / LDL': run on one GPU)%s", "" ); return B200_ERR_ARG; }
    const char* md = getenv( "B200_MIN_DIST_FRONT" );
    c->plan = b200_plan_create( S, c->world, md ? atoi( md ) : 2048 );
    if ( !c->plan ) return B200_ERR_NOMEM;
  }
  c->comm_bytes = 0;
  c->w_rows[0] = c->w_rows[1] = 0;
  for ( int d = 0; d <= S->maxdepth; d++ ) c->w_rows[d & 1] = std::max( c->w_rows[d & 1], c->level_w[d] );

  /* ---- device storage */
  if ( c->dry ) { /* placeholder bases: only offsets matter for the exported schedule */
    c->dL = reinterpret_cast<double*>( ( uintptr_t )1 << 40 ); c->dCB[0] = reinterpret_cast<double*>( ( uintptr_t )2 << 40 ); c->dCB[1] = reinterpret_cast<double*>( ( uintptr_t )3 << 40 );
    c->dInv = reinterpret_cast<double*>( ( uintptr_t )4 << 40 ); c->dSinvF = reinterpret_cast<double*>( ( uintptr_t )5 << 40 ); c->dSinvB = reinterpret_cast<double*>( ( uintptr_t )6 << 40 );
    c->dVals = reinterpret_cast<double*>( ( uintptr_t )7 << 40 );
    if ( TWO ) { c->dU = reinterpret_cast<double*>( ( uintptr_t )8 << 40 ); c->dInv2 = reinterpret_cast<double*>( ( uintptr_t )9 << 40 ); if ( LU ) c->dInv3 = reinterpret_cast<double*>( ( uintptr_t )12 << 40 ); c->dSinvTmp = reinterpret_cast<double*>( ( uintptr_t )10 << 40 ); c->dPiv = reinterpret_cast<int*>( ( uintptr_t )11 << 40 ); }
  }
  const double need = ( double )c->lsize * 8.0 * ( TWO ? 2 : 1 ) + ( double )( c->cb_arena[0] + c->cb_arena[1] ) * 8.0 + ( double )invoff * 8.0 * ( LU ? 3 : TWO ? 2 : 1 ) + ( double )sinvoff * 8.0 * ( TWO ? 3 : 2 ) + ( double )S->nnzA * 16.0;
  if ( !c->dry ) {
  size_t freeb = 0, totb = 0; cudaMemGetInfo( &freeb, &totb );
  if ( need > ( double )freeb * 0.97 ) { char a[64], b[64]; snprintf( a, 64, "%.1f", need / 1e9 ); snprintf( b, 64, "%.1f", freeb / 1e9 ); set_err( "factor needs %s GB of HBM, %s GB free", a, b ); return B200_ERR_NOMEM; }
  CK( cudaMalloc( &c->dL, ( std::max<int64_t>( c->lsize, 1 ) + 32 ) * 8 ) );     /* + slack: the row-run copies of the backward tensor-path sweep read up to 17 rows past a slab */
  CK( cudaMemset( c->dL + std::max<int64_t>( c->lsize, 1 ), 0, 32 * 8 ) );
  if ( TWO ) { CK( cudaMalloc( &c->dU, ( std::max<int64_t>( c->lsize, 1 ) + 32 ) * 8 ) ); CK( cudaMemset( c->dU + std::max<int64_t>( c->lsize, 1 ), 0, 32 * 8 ) ); }
  for ( int a = 0; a < 2; a++ ) CK( cudaMalloc( &c->dCB[a], std::max<int64_t>( c->cb_arena[a], 2 ) * 8 ) );
  CK( cudaMalloc( &c->dInv, ( std::max<int64_t>( invoff, 1 ) + 16 ) * 8 ) );     /* + slack: the bulk copies of an odd-aligned block read one double past it */
  if ( TWO ) { CK( cudaMalloc( &c->dInv2, ( std::max<int64_t>( invoff, 1 ) + 16 ) * 8 ) ); CK( cudaMalloc( &c->dPiv, sizeof( int ) * n ) ); }
  if ( LU ) CK( cudaMalloc( &c->dInv3, ( std::max<int64_t>( invoff, 1 ) + 16 ) * 8 ) );
  CK( cudaMalloc( &c->dVals, std::max<int64_t>( S->nnzA, 1 ) * 8 ) );
  CK( cudaMalloc( &c->dSinvF, ( std::max<int64_t>( sinvoff, 1 ) + 16 ) * 8 ) ); CK( cudaMalloc( &c->dSinvB, ( std::max<int64_t>( sinvoff, 1 ) + 16 ) * 8 ) );
  CK( cudaMemset( c->dSinvF, 0, ( std::max<int64_t>( sinvoff, 1 ) + 16 ) * 8 ) ); CK( cudaMemset( c->dSinvB, 0, ( std::max<int64_t>( sinvoff, 1 ) + 16 ) * 8 ) );     /* padding rows of odd blocks stay zero */
  if ( TWO ) { CK( cudaMalloc( &c->dSinvTmp, ( std::max<int64_t>( sinvoff, 1 ) + 16 ) * 8 ) ); CK( cudaMemset( c->dSinvTmp, 0, ( std::max<int64_t>( sinvoff, 1 ) + 16 ) * 8 ) ); }
  }
  c->device_bytes = ( int64_t )need;
  { std::vector<int64_t> v( S->amap, S->amap + S->nnzA ); if ( c->amap.upload( v ) ) return B200_ERR_CUDA; }
  { std::vector<int> v( S->perm, S->perm + n ); if ( c->perm.upload( v ) ) return B200_ERR_CUDA; }
  { std::vector<int> v( S->rowidx, S->rowidx + S->rowptr[ns] ); if ( c->rowidx.upload( v ) ) return B200_ERR_CUDA; }
  { std::vector<int> v( S->relidx, S->relidx + S->rowptr[ns] ); if ( c->relidx.upload( v ) ) return B200_ERR_CUDA; }
  { std::vector<int64_t> v( S->spmv_ptr, S->spmv_ptr + n + 1 ); if ( c->spmv_ptr.upload( v ) ) return B200_ERR_CUDA;
    std::vector<int> a( S->spmv_col, S->spmv_col + S->spmv_ptr[n] ), b( S->spmv_src, S->spmv_src + S->spmv_ptr[n] );
    if ( c->spmv_col.upload( a ) || c->spmv_src.upload( b ) ) return B200_ERR_CUDA; }
  if ( c->child_list.upload( child_list ) ) return B200_ERR_CUDA;
  if ( c->sn.upload( meta ) ) return B200_ERR_CUDA;

  /* ---- factor schedule.  With more than one rank (b200_mg_init) every rank walks the same loops and keeps
   * only the work it owns (b200_plan.c); messages are emitted in the same global order on every rank. */
  g_alg_flops = 0;
  std::vector<GemmTask> gt; std::vector<Tile> tb, ts; std::vector<DiagTask> dt; std::vector<TrsmTask> tt; std::vector<RowTile> ttl; std::vector<AsmTask> at;
  std::vector<Launch>& LS = c->launches;
  std::vector<Msg>& MS = c->msgs;
  std::vector<int> smalls;
  double* Lb = c->dL; double* Ub = c->dU;
  const bool MG = c->world > 1;
  const int me = c->rank;
  const b200_plan_t* pl = c->plan;
  auto is_dist = [&]( int s ) { return MG && pl->dist[s] != 0; };
  auto member = [&]( int s ) { return !MG || b200_plan_member( pl, s, me ) != 0; };
  auto owns = [&]( int s, int col ) { return !MG || b200_plan_col_owner( pl, S, s, col ) == me; };
  auto new_event = [&]() { return c->nevents++; };
  /* join helpers: make stream `to` wait for everything issued so far on stream `from` */
  auto push_nop = [&]( int d, int stream ) { Launch l = { LK_NOP, 0, 0, d, 0.0 }; l.stream = stream; LS.push_back( l ); };
  /* runs of supernodes with the same holders (contiguous in L storage) and the level after which they are final */
  struct Run { int s0, s1, h0, hn, ready; };
  std::vector<Run> runs; int ev_bulk_last = -1;
  if ( MG ) {
    int s0 = 0;
    while ( s0 < ns ) {
      if ( pl->dist[s0] ) { runs.push_back( { s0, s0, pl->grp_first[s0], pl->grp_size[s0], meta[s0].depth } ); s0++; continue; }
      int s1 = s0, dmin = meta[s0].depth;
      while ( s1 + 1 < ns && !pl->dist[s1 + 1] && pl->grp_first[s1 + 1] == pl->grp_first[s0] ) { s1++; dmin = std::min( dmin, meta[s1].depth ); }
      runs.push_back( { s0, s1, pl->grp_first[s0], 1, dmin } );
      s0 = s1 + 1;
    }
  }
  int ev_asm_done = -1;                      /* recorded on the update stream once the extend-add of the level below has read its children's arena */
  for ( int d = S->maxdepth; d >= 0; d-- ) {
    std::vector<int> lv;                                     /* the supernodes of this level I take part in */
    int maxf_all = 0;
    for ( int s : levels[d] ) { maxf_all = std::max( maxf_all, meta[s].f ); if ( member( s ) ) lv.push_back( s ); }
    double* CB = c->dCB[d & 1];
    { /* The contribution blocks of this level are zeroed before its extend-add accumulates into them.  The arena (parity d & 1) was last read by the
       * extend-add of level d+1, so the zeroing -- pure HBM writes -- runs on a stream of its own UNDER the compute-bound updates of level d+1 instead of
       * in front of this level (9.5 ms of the 128^3 factorization when it ran in line): it waits for ev_asm_done (recorded after that extend-add) and
       * this level's first launch waits for it. */
      const size_t z0 = LS.size();
      if ( !MG ) { if ( c->level_cb[d] > 0 ) LS.push_back( { LK_MEMSET_CB, 0, c->level_cb[d], d, 0.0 } ); }
      else { /* zero only the contribution blocks I will write: merge neighbouring ranges of the arena */
        int64_t lo = -1, hi = -1;
        for ( int s : lv ) {
          const int64_t u = meta[s].f - meta[s].k; if ( u == 0 ) continue;
          const int64_t a = meta[s].cbptr, b = a + ( ( u + 1 ) & ~( int64_t )1 ) * u;
          if ( lo >= 0 && a <= hi + 4096 ) { hi = std::max( hi, b ); continue; }
          if ( lo >= 0 ) LS.push_back( { LK_MEMSET_CB, lo, hi - lo, d, 0.0 } );
          lo = a; hi = b;
        }
        if ( lo >= 0 ) LS.push_back( { LK_MEMSET_CB, lo, hi - lo, d, 0.0 } );
      }
      if ( LS.size() > z0 ) {
        for ( size_t q = z0; q < LS.size(); q++ ) LS[q].stream = 4;
        LS[z0].wait_ev = ev_asm_done;                    /* -1 at the deepest level: nothing has used the arena yet */
        const int ez = new_event(); LS.back().record_ev = ez;
        push_nop( d, 0 ); LS.back().wait_ev = ez;
      }
    }
    if ( MG ) { /* contribution-block columns computed elsewhere that my part of this level's fronts needs, and vice versa */
      const int64_t nx = b200_plan_cb_transfers( pl, S, d, nullptr, 0 );
      if ( nx > 0 ) {
        std::vector<b200_xfer_t> xs( nx ); b200_plan_cb_transfers( pl, S, d, xs.data(), nx );
        const int64_t m0 = ( int64_t )MS.size();
        for ( const b200_xfer_t& x : xs ) if ( x.src == me || x.dst == me ) { MS.push_back( { x.src, x.dst, 2 + ( ( d + 1 ) & 1 ), x.offset, x.count } ); c->comm_bytes += ( double )x.count * 8.0; }
        if ( ( int64_t )MS.size() > m0 ) {
          const int ea = new_event(), eb = new_event();
          push_nop( d, 0 ); LS.back().record_ev = ea;
          Launch l = { LK_XFER_CB, m0, ( int64_t )MS.size() - m0, d, 0.0 }; l.stream = 2; l.wait_ev = ea; l.record_ev = eb; LS.push_back( l );
          push_nop( d, 0 ); LS.back().wait_ev = eb;
        }
      }
    }
    { /* extend-add of the children (they live at depth d+1) into the destination columns I own */
      const int64_t first = ( int64_t )at.size();
      for ( int s : lv ) {
        if ( meta[s].child_begin == meta[s].child_end ) continue;
        bool any = false;
        for ( int ci = meta[s].child_begin; ci < meta[s].child_end; ci++ ) { const SnMeta& ch = meta[child_list[ci]]; if ( ch.f > ch.k ) { any = true; break; } }
        if ( !any ) continue;
        const int acols = meta[s].f >= 4096 ? ASM_COLS / 2 : ASM_COLS;      /* large fronts: finer tasks, so that the single partial wave of CTAs balances */
        for ( int c0 = 0; c0 < meta[s].f; c0 += acols ) if ( owns( s, c0 ) ) at.push_back( { s, c0, std::min( acols, meta[s].f - c0 ) } );
      }
      if ( ( int64_t )at.size() > first ) LS.push_back( { LK_ASM, first, ( int64_t )at.size() - first, d, 0.0 } );
      /* from here on the children's arena (parity (d+1) & 1) is free: level d-1 may zero it */
      push_nop( d, 0 ); ev_asm_done = new_event(); LS.back().record_ev = ev_asm_done;
    }
    if ( c->use_small_front && !TWO ) { /* small fronts: one CTA each, a single launch for the level (small_front_kernel) */
      const int64_t s0 = ( int64_t )smalls.size();
      std::vector<int> rest;
      for ( int s : lv ) { if ( meta[s].k <= SF_KMAX && meta[s].f <= SF_FMAX && !is_dist( s ) ) smalls.push_back( s ); else rest.push_back( s ); }
      if ( ( int64_t )smalls.size() > s0 ) LS.push_back( { LK_SMALL_FRONT, s0, ( int64_t )smalls.size() - s0, d, 0.0 } );
      lv.swap( rest );
    }
    int maxk = 0; bool any_dist = false;
    for ( int s : lv ) { maxk = std::max( maxk, meta[s].k ); any_dist = any_dist || is_dist( s ); }
    const int NBO = outer_width( maxf_all );
    /* look-ahead: with more than one outer panel, the next panel is factored on a second stream while
     * the bulk of the trailing update of the current one still runs on the first */
    const bool look = maxk > NBO || any_dist;
    int ev_next = -1;                       /* recorded after update_next(J-1): panel(J) may start */
    if ( look ) {                           /* level prologue (and everything before it) done */
      if ( LS.empty() || LS.back().record_ev >= 0 || LS.back().stream != 0 ) push_nop( d, 0 );
      LS.back().record_ev = c->nevents; ev_next = c->nevents++;
    }
    for ( int J = 0; J < maxk; J += NBO ) {
      const size_t panel_first = LS.size();
      int ev_ring_last = -1;                /* record event of the last ring step in which a piece of this panel reached me */
      /* A factored panel travels round its group as a RING, 64-column block by block: the owner hands block b to the owner of the
       * next panel at step b (right after the block's row update, while the rest of the panel is still being factored), that rank
       * hands it on at step b+1, and so on -- hop h carries block (step - h).  Every rank sends a panel once; the owner of panel
       * J+h is hop h of the ring, so the ranks receive it in the order in which they need it.  (Round 1 sent the whole panel from its
       * owner to every member: 7 x 77 MB out of one GPU per root panel at 8 GPUs, the largest item of the critical path.)
       * The ring always closes over the whole group. */
      auto ring_step = [&]( int sg, int Jp, bool after_compute ) {
        const int64_t m0 = ( int64_t )MS.size(); bool i_recv = false, i_send_own = false;
        for ( int s : lv ) {
          const SnMeta& m = meta[s];
          if ( !is_dist( s ) || Jp >= m.k ) continue;
          const int gf = pl->grp_first[s], gs = pl->grp_size[s];
          const int root = b200_plan_col_owner( pl, S, s, Jp );
          const int nbp = ( std::min( NBO, m.k - Jp ) + NB - 1 ) / NB;                 /* blocks of this panel */
          const int hops = gs - 1;                                                      /* every member ends up with every panel of the front (the collection of the finished factor relies on it) */
          for ( int h = 0; h < hops; h++ ) {
            const int b = sg - h; if ( b < 0 || b >= nbp ) continue;
            const int src = gf + ( root - gf + h ) % gs, dst = gf + ( root - gf + h + 1 ) % gs;
            if ( me != src && me != dst ) continue;
            const int jb = Jp + b * NB, nb = std::min( NB, m.k - jb );
            MS.push_back( { src, dst, 0, m.lptr + ( int64_t )jb * m.ld, ( int64_t )nb * m.ld } );
            MS.push_back( { src, dst, 1, m.invptr + ( int64_t )( jb / NB ) * NB * NB, ( int64_t )nb * nb } );
            c->comm_bytes += ( double )nb * m.ld * 8.0;
            if ( me == dst ) i_recv = true;
            if ( me == src && h == 0 ) i_send_own = true;
          }
        }
        if ( ( int64_t )MS.size() == m0 ) return;
        int wait = -1;
        if ( i_send_own && after_compute ) { wait = new_event(); LS.back().record_ev = wait; }     /* LS.back(): my last compute launch of this block (row update, or the diagonal block) */
        Launch l = { LK_BCAST, m0, ( int64_t )MS.size() - m0, d, 0.0 }; l.stream = 2; l.wait_ev = wait; l.record_ev = new_event();
        LS.push_back( l );
        if ( i_recv ) ev_ring_last = l.record_ev;
      };
      for ( int j0 = J; j0 < std::min( maxk, J + NBO ); j0 += NB ) {
        if ( j0 > J ) { /* left-looking inside the outer panel: the 64 columns about to be factored receive the update of ALL earlier
                         * blocks of the panel in one product (K = j0 - J) -- one read-modify-write of 64 columns instead of one
                         * K = 64 pass over the rest of the panel after every block (3.5x less traffic, a K the DMMA pipeline can fill) */
          const int64_t b0 = ( int64_t )tb.size();
          double fb = 0;
          for ( int s : lv ) {
            const SnMeta& m = meta[s];
            if ( j0 >= m.k || !owns( s, J ) ) continue;
            const int nb = std::min( NB, m.k - j0 ); const int kk = j0 - J;
            double* Lp = Lb + m.lptr; double* Up = TWO ? Ub + m.lptr : nullptr;
            double fl = 0;
            if ( LD ) add_gemm( gt, tb, ts, fl, Lp + j0 + ( int64_t )j0 * m.ld, m.ld, Lp + j0 + ( int64_t )J * m.ld, m.ld, Up + j0 + ( int64_t )J * m.ld, m.ld, m.f - j0, nb, kk, 1, 0 );      /* A -= L (L D)' */
            else if ( !LU ) add_gemm( gt, tb, ts, fl, Lp + j0 + ( int64_t )j0 * m.ld, m.ld, Lp + j0 + ( int64_t )J * m.ld, m.ld, Lp + j0 + ( int64_t )J * m.ld, m.ld, m.f - j0, nb, kk, 1, 0 );
            else {
              add_gemm( gt, tb, ts, fl, Lp + j0 + ( int64_t )j0 * m.ld, m.ld, Lp + j0 + ( int64_t )J * m.ld, m.ld, Up + j0 + ( int64_t )J * m.ld, m.ld, m.f - j0, nb, kk, 1, 0 );
              add_gemm( gt, tb, ts, fl, Up + j0 + ( int64_t )j0 * m.ld, m.ld, Up + j0 + ( int64_t )J * m.ld, m.ld, Lp + j0 + ( int64_t )J * m.ld, m.ld, m.f - j0, nb, kk, 1, -1 );
            }
            fb += fl;
          }
          if ( ( int64_t )tb.size() > b0 ) { LS.push_back( { LK_GEMM_BIG, b0, ( int64_t )tb.size() - b0, d, fb } ); }
        }
        const int64_t d0 = ( int64_t )dt.size(); int64_t d1 = d0;
        for ( int pass = 0; pass < 2; pass++ ) {           /* full 64-column blocks first (unrolled register kernel), then the narrower ones */
          for ( int s : lv ) {
            const SnMeta& m = meta[s];
            if ( j0 >= m.k || !owns( s, J ) ) continue;
            const int nb = std::min( NB, m.k - j0 );
            if ( !TWO && ( pass == 0 ) != ( nb == NB ) ) continue;
            if ( TWO && pass == 1 ) continue;
            DiagTask x; x.lda = m.ld; x.nb = nb; x.col = m.c0 + j0;
            x.A = Lb + m.lptr + j0 + ( int64_t )j0 * m.ld; x.AU = TWO ? Ub + m.lptr + j0 + ( int64_t )j0 * m.ld : nullptr;
            x.inv = c->dInv + m.invptr + ( int64_t )( j0 / NB ) * NB * NB; x.inv2 = TWO ? c->dInv2 + m.invptr + ( int64_t )( j0 / NB ) * NB * NB : nullptr;
            x.piv = LU ? c->dPiv + m.c0 + j0 : nullptr; x.inv3 = LU ? c->dInv3 + m.invptr + ( int64_t )( j0 / NB ) * NB * NB : nullptr;
            dt.push_back( x );
          }
          if ( pass == 0 ) d1 = ( int64_t )dt.size();
        }
        const bool have_diag = ( int64_t )dt.size() > d0;
        if ( !have_diag && !any_dist ) break;
        if ( d1 > d0 ) LS.push_back( { LK_DIAG, d0, d1 - d0, d, 0.0 } );
        if ( ( int64_t )dt.size() > d1 ) LS.push_back( { LK_DIAG_S, d1, ( int64_t )dt.size() - d1, d, 0.0 } );
        /* rows below the diagonal block */
        const int64_t tl0 = ( int64_t )ttl.size(); std::vector<Tile> tbt, tbt2; std::vector<RowTile> ttl2;
        double tflops = 0;
        for ( int s : lv ) {
          const SnMeta& m = meta[s];
          if ( j0 >= m.k || !owns( s, J ) ) continue;
          const int nb = std::min( NB, m.k - j0 ); const int below = m.f - j0 - nb;
          if ( below <= 0 ) continue;
          const double* inv = c->dInv + m.invptr + ( int64_t )( j0 / NB ) * NB * NB;
          const double* inv2 = TWO ? c->dInv2 + m.invptr + ( int64_t )( j0 / NB ) * NB * NB : nullptr;
          double* Bp = Lb + m.lptr + ( j0 + nb ) + ( int64_t )j0 * m.ld;
          double* Wp = TWO ? Ub + m.lptr + ( j0 + nb ) + ( int64_t )j0 * m.ld : nullptr;
          if ( LD ) { /* W21 = A21 T' into the second panel first, then L21 = A21 T2' in place (two launches: the second overwrites what the first reads) */
            if ( nb == NB ) {
              double dummy = 0; size_t ntask = gt.size();
              add_gemm( gt, tbt, ts, dummy, Wp, m.ld, Bp, m.ld, inv, nb, below, nb, nb, 0, 0 );
              if ( gt.size() > ntask ) { gt.back().mode = 1; g_alg_flops -= 2.0 * below * nb * nb; }
              ntask = gt.size();
              add_gemm( gt, tbt2, ts, dummy, Bp, m.ld, Bp, m.ld, inv2, nb, below, nb, nb, 0, 0 );
              if ( gt.size() > ntask ) { gt.back().mode = 1; g_alg_flops -= 2.0 * below * nb * nb; }
            } else {
              TrsmTask x; x.B = Bp; x.ldb = m.ld; x.m = below; x.nb = nb; x.T = inv; x.piv = nullptr; x.out = Wp; x.full = 1;
              int id = ( int )tt.size(); tt.push_back( x );
              for ( int r0 = 0; r0 < below; r0 += 128 ) ttl.push_back( { id, r0 } );
              TrsmTask y = x; y.T = inv2; y.out = nullptr;
              id = ( int )tt.size(); tt.push_back( y );
              for ( int r0 = 0; r0 < below; r0 += 128 ) ttl2.push_back( { id, r0 } );
            }
            tflops += 2.0 * below * nb * nb;
            continue;
          }
          if ( nb == NB ) {   /* X = B T' on the tensor pipe: C = A B' with C = A = the rows below, B = the inverse diagonal block (in place: one CTA owns whole rows) */
            double dummy = 0; const size_t ntask = gt.size();
            add_gemm( gt, tbt, ts, dummy, Bp, m.ld, Bp, m.ld, LU ? inv2 : inv, nb, below, nb, nb, 0, 0 );
            if ( gt.size() > ntask ) { gt.back().mode = 1; g_alg_flops -= 2.0 * below * nb * nb; }
            tflops += ( double )below * nb * nb;
          } else {
          TrsmTask x; x.B = Bp; x.ldb = m.ld; x.m = below; x.nb = nb; x.T = LU ? inv2 : inv; x.piv = nullptr; x.out = nullptr; x.full = 0;
          int id = ( int )tt.size(); tt.push_back( x );
          for ( int r0 = 0; r0 < below; r0 += 128 ) ttl.push_back( { id, r0 } );
          tflops += ( double )below * nb * nb;
          }
          if ( LU ) { /* rows of the U' panel below the block:  U21' = (A12' P') inv(L)' .  Full blocks: the interchanges are folded into inv3 (getrf_diag), so this
                       * is the same in-place tensor-pipe product as the L rows (round 2: 51 ms of the 27-pt 96^3 factorization sat in the register-tile kernel) */
            if ( nb == NB ) {
              double dummy = 0; const size_t ntask = gt.size();
              add_gemm( gt, tbt, ts, dummy, Wp, m.ld, Wp, m.ld, c->dInv3 + m.invptr + ( int64_t )( j0 / NB ) * NB * NB, nb, below, nb, nb, 0, 0 );
              if ( gt.size() > ntask ) { gt.back().mode = 1; g_alg_flops -= 2.0 * below * nb * nb; }
            } else {
              TrsmTask y; y.B = Wp; y.ldb = m.ld; y.m = below; y.nb = nb; y.T = inv; y.piv = c->dPiv + m.c0 + j0; y.out = nullptr; y.full = 0;
              const int id = ( int )tt.size(); tt.push_back( y );
              for ( int r0 = 0; r0 < below; r0 += 128 ) ttl.push_back( { id, r0 } );
            }
            tflops += ( double )below * nb * nb;
          }
        }
        if ( ( int64_t )ttl.size() > tl0 ) LS.push_back( { LK_TRSM, tl0, ( int64_t )ttl.size() - tl0, d, tflops } );
        if ( !tbt.empty() ) { LS.push_back( { LK_TRSM_MMA, ( int64_t )tb.size(), ( int64_t )tbt.size(), d, tflops } ); tb.insert( tb.end(), tbt.begin(), tbt.end() ); }
        if ( !ttl2.empty() ) { LS.push_back( { LK_TRSM, ( int64_t )ttl.size(), ( int64_t )ttl2.size(), d, 0.0 } ); ttl.insert( ttl.end(), ttl2.begin(), ttl2.end() ); }
        if ( !tbt2.empty() ) { LS.push_back( { LK_TRSM_MMA, ( int64_t )tb.size(), ( int64_t )tbt2.size(), d, 0.0 } ); tb.insert( tb.end(), tbt2.begin(), tbt2.end() ); }
        if ( any_dist ) ring_step( ( j0 - J ) / NB, J, true );
        if ( !have_diag ) continue;
      }
      int ev_panel = -1;
      if ( look ) { /* my compute launches of this panel go to the panel stream */
        int qf = -1, ql = -1;
        for ( size_t q = panel_first; q < LS.size(); q++ ) if ( LS[q].stream == 0 ) { LS[q].stream = 1; if ( qf < 0 ) qf = ( int )q; ql = ( int )q; }
        if ( qf >= 0 ) {
          LS[qf].wait_ev = ev_next;
          if ( LS[ql].record_ev < 0 ) LS[ql].record_ev = c->nevents++;
          ev_panel = LS[ql].record_ev;
        }
      }
      if ( any_dist ) { /* the steps of the ring that are still due once the panel's last block has been factored */
        const int nsteps_done = ( std::min( maxk, J + NBO ) - J + NB - 1 ) / NB;
        for ( int sg = nsteps_done; sg < NBO / NB + c->world - 2; sg++ ) ring_step( sg, J, false );
        if ( ev_ring_last >= 0 ) { /* a panel I did not factor is complete here once its last block has arrived; the update stream must wait for that AND for
                                    * my own panels of this step (other fronts of the level): join the NCCL stream into the panel stream */
          push_nop( d, 1 ); LS.back().wait_ev = ev_ring_last; LS.back().record_ev = new_event(); ev_panel = LS.back().record_ev;
        }
      }
      /* trailing update with the whole outer panel, in two launches: first the columns of the NEXT
       * outer panel (all the next panel factorization depends on), then the remaining pivot columns
       * and the contribution block */
      for ( int part = 0; part < 2; part++ ) {
        const int64_t b0 = ( int64_t )tb.size();
        double fl = 0;
        for ( int s : lv ) {
          const SnMeta& m = meta[s];
          if ( J >= m.k ) continue;
          const int kb = std::min( NBO, m.k - J ); const int Jn = J + kb; const int u = m.f - m.k;
          const int Jn2 = std::min( m.k, Jn + NBO );           /* end of the next outer panel */
          double* Lp = Lb + m.lptr; double* Up = TWO ? Ub + m.lptr : nullptr;
          double* cb = CB + m.cbptr;
          if ( is_dist( s ) ) { /* my column blocks only (Cholesky) */
            const int ldu = ( u + 1 ) & ~1;
            if ( part == 0 ) {
              if ( Jn < Jn2 && owns( s, Jn ) ) { double* Cl = Lp + Jn + ( int64_t )Jn * m.ld; const double* Al = Lp + Jn + ( int64_t )J * m.ld; add_gemm( gt, tb, ts, fl, Cl, m.ld, Al, m.ld, Al, m.ld, m.f - Jn, Jn2 - Jn, kb, 1, 0 ); }
            } else {
              for ( int blk = Jn2 / NBO; blk * NBO < m.f; blk++ ) {
                if ( !owns( s, blk * NBO ) ) continue;
                const int c_lo = std::max( blk * NBO, Jn2 ), c_hi = std::min( ( blk + 1 ) * NBO, m.f );
                if ( c_lo >= c_hi ) continue;
                const int p_hi = std::min( c_hi, m.k );
                if ( c_lo < p_hi ) { double* Cl = Lp + c_lo + ( int64_t )c_lo * m.ld; const double* Al = Lp + c_lo + ( int64_t )J * m.ld; add_gemm( gt, tb, ts, fl, Cl, m.ld, Al, m.ld, Al, m.ld, m.f - c_lo, p_hi - c_lo, kb, 1, 0 ); }
                const int j_lo = std::max( c_lo, m.k ) - m.k, j_hi = c_hi - m.k;
                if ( j_lo < j_hi ) { const double* Al = Lp + ( m.k + j_lo ) + ( int64_t )J * m.ld; add_gemm( gt, tb, ts, fl, cb + j_lo + ( int64_t )j_lo * ldu, ldu, Al, m.ld, Al, m.ld, u - j_lo, j_hi - j_lo, kb, 1, 0 ); }
              }
            }
            continue;
          }
          const int c_lo = part == 0 ? Jn : Jn2, c_hi = part == 0 ? Jn2 : m.k;   /* pivot columns updated by this part */
          if ( c_lo < c_hi ) {
            double* Cl = Lp + c_lo + ( int64_t )c_lo * m.ld; const double* Al = Lp + c_lo + ( int64_t )J * m.ld;
            if ( LD ) add_gemm( gt, tb, ts, fl, Cl, m.ld, Al, m.ld, Up + c_lo + ( int64_t )J * m.ld, m.ld, m.f - c_lo, c_hi - c_lo, kb, 1, 0 );
            else if ( !LU ) add_gemm( gt, tb, ts, fl, Cl, m.ld, Al, m.ld, Al, m.ld, m.f - c_lo, c_hi - c_lo, kb, 1, 0 );
            else {
              double* Cu = Up + c_lo + ( int64_t )c_lo * m.ld; const double* Au = Up + c_lo + ( int64_t )J * m.ld;
              add_gemm( gt, tb, ts, fl, Cl, m.ld, Al, m.ld, Au, m.ld, m.f - c_lo, c_hi - c_lo, kb, 1, 0 );
              add_gemm( gt, tb, ts, fl, Cu, m.ld, Au, m.ld, Al, m.ld, m.f - c_lo, c_hi - c_lo, kb, 1, -1 );
            }
          }
          if ( part == 1 && u > 0 ) {
            const double* Al = Lp + m.k + ( int64_t )J * m.ld;
            if ( LD ) add_gemm( gt, tb, ts, fl, cb, ( u + 1 ) & ~1, Al, m.ld, Up + m.k + ( int64_t )J * m.ld, m.ld, u, u, kb, 1, 0 );
            else if ( !LU ) add_gemm( gt, tb, ts, fl, cb, ( u + 1 ) & ~1, Al, m.ld, Al, m.ld, u, u, kb, 1, 0 );
            else add_gemm( gt, tb, ts, fl, cb, ( u + 1 ) & ~1, Al, m.ld, Up + m.k + ( int64_t )J * m.ld, m.ld, u, u, kb, 0, 0 );
          }
        }
        const bool any = ( int64_t )tb.size() > b0;
        if ( any ) { LS.push_back( { LK_GEMM_BIG, b0, ( int64_t )tb.size() - b0, d, fl } ); }
        if ( look ) {
          if ( part == 0 ) {
            if ( !any ) push_nop( d, 0 );
            LS.back().wait_ev = ev_panel;                               /* update stream waits for the panel */
            LS.back().record_ev = c->nevents; ev_next = c->nevents++;   /* next panel may start */
          }
        }
      }
    }
    if ( MG ) { /* collect what has become final on this level on the SOLVE ROOT (rank 0): it ends up with the whole factor (L panels and
                 * diagonal-block inverses) and runs the sweeps.  (Round 1 replicated the factor on every rank: 72 % of all NVLink bytes
                 * at 8 GPUs, in competition with the panel broadcasts of the critical path; B200_REPLICATE_FACTOR=1 restores that.)
                 * Bulk traffic has its own communicator and stream so that it never queues in front of a panel. */
      const int64_t m0 = ( int64_t )MS.size();
      for ( const Run& rn : runs ) {
        if ( rn.ready != d ) continue;
        const int64_t lo = meta[rn.s0].lptr, lcnt = S->lptr[rn.s1 + 1] - lo;
        const int64_t io = meta[rn.s0].invptr, icnt = ( rn.s1 + 1 < ns ? meta[rn.s1 + 1].invptr : c->invsize ) - io;
        for ( int r = 0; r < ( c->replicate ? c->world : 1 ); r++ ) {
          if ( r >= rn.h0 && r < rn.h0 + rn.hn ) continue;
          if ( me != rn.h0 && me != r ) continue;
          if ( lcnt > 0 ) { MS.push_back( { rn.h0, r, 0, lo, lcnt } ); c->comm_bytes += ( double )lcnt * 8.0; }
          if ( icnt > 0 ) MS.push_back( { rn.h0, r, 1, io, icnt } );
        }
      }
      if ( ( int64_t )MS.size() > m0 ) {
        const int ea = new_event();
        push_nop( d, 0 ); LS.back().record_ev = ea;
        Launch l = { LK_GATHER, m0, ( int64_t )MS.size() - m0, d, 0.0 }; l.stream = 3; l.wait_ev = ea; l.record_ev = new_event(); ev_bulk_last = l.record_ev;
        LS.push_back( l );
      }
    }
  }
  if ( MG && ev_bulk_last >= 0 ) { push_nop( 0, 0 ); LS.back().wait_ev = ev_bulk_last; }     /* the factor is complete on this rank once the bulk channel has drained */
  if ( MG ) { /* ... and once every panel / contribution-block message has left: the sender side of the last broadcasts is otherwise
               * never joined, and the next factorization would zero L under sends that are still queued */
    const int e2 = new_event(); push_nop( 0, 2 ); LS.back().record_ev = e2;
    push_nop( 0, 0 ); LS.back().wait_ev = e2;
    const int e1 = new_event(); push_nop( 0, 1 ); LS.back().record_ev = e1;
    push_nop( 0, 0 ); LS.back().wait_ev = e1;
  }
  if ( c->dry ) { c->h_gt = gt; c->h_tb = tb; c->h_dt = dt; c->h_tt = tt; c->h_ttl = ttl; c->h_at = at; c->h_small = smalls; c->h_child = child_list; c->h_meta = meta; }
  if ( c->small_sns.upload( smalls ) ) return B200_ERR_CUDA;
  if ( c->gemm_tasks.upload( gt ) || c->tiles_big.upload( tb ) || c->tiles_small.upload( ts ) || c->diag_tasks.upload( dt ) ||
       c->trsm_tasks.upload( tt ) || c->trsm_tiles.upload( ttl ) || c->asm_tasks.upload( at ) ) return B200_ERR_CUDA;

  c->gemm_alg_flops = g_alg_flops;
  while ( !c->dry && ( int )c->events.size() < c->nevents ) { cudaEvent_t e; CK( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) ); c->events.push_back( e ); }
  /* ---- solve-block inverses: one task per (solve block, 64-column block column) */
  {
    std::vector<SinvTask> sv;
    for ( int s = 0; s < ns; s++ ) {
      const SnMeta& m = meta[s];
      for ( int j0 = 0; j0 < m.k; j0 += m.sbw ) {
        const int nb = std::min( m.sbw, m.k - j0 );
        const int64_t off = m.sinv + ( int64_t )( j0 / m.sbw ) * m.sbw * m.sbw;
        for ( int jb = 0; jb * NB < nb; jb++ ) {
          SinvTask x; x.P = Lb + m.lptr; x.ld = m.ld; x.j0 = j0; x.nb = nb; x.jb = jb; x.inv64 = c->dInv + m.invptr; x.ldx = ( nb + 1 ) & ~1;
          x.piv = LU ? c->dPiv + m.c0 : nullptr; x.X = c->dSinvF + off; x.XT = TWO ? nullptr : c->dSinvB + off;
          sv.push_back( x );
          if ( TWO ) { SinvTask y = x; y.P = Ub + m.lptr; y.inv64 = c->dInv2 + m.invptr; y.piv = nullptr; y.X = c->dSinvTmp + off; y.XT = c->dSinvB + off; sv.push_back( y ); }
        }
      }
    }
    if ( c->sinv_tasks.upload( sv ) ) return B200_ERR_CUDA;
  }
  /* ---- solve schedule: forward deepest->root, backward root->deepest.  Per level and block step two launches: the triangular product
   * with the block's explicit inverse, then the update of everything the block feeds (kernels: sv_gemv / sv_gemvT, or the tensor path) */
  {
    std::vector<SvTask> tk; std::vector<SvTile> tv, tt, tg; std::vector<GatherTask> gp, bg; std::vector<int> sm;
    const bool use_small = getenv( "B200_NO_SMALL_SOLVE" ) == nullptr;
    auto is_small = [&]( int s ) { return use_small && meta[s].k <= KC && meta[s].f <= 4 * SVS_THREADS; };      /* one canonical chunk, one solve block: the fused per-supernode kernels */
    std::vector<std::vector<int>> big( S->maxdepth + 1 ); std::vector<std::pair<int64_t, int64_t>> small_rng( S->maxdepth + 1 );
    for ( int d = 0; d <= S->maxdepth; d++ ) { small_rng[d].first = ( int64_t )sm.size(); for ( int s : levels[d] ) { if ( is_small( s ) ) sm.push_back( s ); else big[d].push_back( s ); } small_rng[d].second = ( int64_t )sm.size() - small_rng[d].first; }
    std::vector<SLaunch>& SS = c->slaunches;
    const double* Fb = c->dL;                       /* forward sweep: L */
    const double* Mb = TWO ? c->dU : c->dL;         /* backward sweep: L' x, or U x through the U' panels (L D L': U' = W = L D) */
    struct Mark { int64_t t, g; };
    auto mark = [&]( bool T ) { return Mark{ ( int64_t )( T ? tt.size() : tv.size() ), ( int64_t )tg.size() }; };
    auto add_task = [&]( const SvTask& x, bool T ) {
      if ( x.N <= 0 || x.K <= 0 ) return;
      const int id = ( int )tk.size(); tk.push_back( x );
      if ( T ) for ( int o = 0; o < x.N; o += GT_COLS ) tt.push_back( { id, o } ); else for ( int o = 0; o < x.N; o += GV_ROWS ) tv.push_back( { id, o } );
      for ( int o = 0; o < x.N; o += 64 ) tg.push_back( { id, o } );
    };
    auto close = [&]( int kind, int d, const Mark& m0, bool T ) {
      const int64_t nt = ( int64_t )( T ? tt.size() : tv.size() ) - m0.t, ng = ( int64_t )tg.size() - m0.g;
      if ( nt > 0 ) SS.push_back( { kind, d, 0, 0, m0.t, nt, m0.g, ng } );
    };
    for ( int d = S->maxdepth; d >= 0; d-- ) {
      const std::vector<int>& lv = levels[d];
      const int wsel = 2 + ( d & 1 );
      if ( c->level_w[d] > 0 ) SS.push_back( { SK_FWD_ZERO_W, d, 0, c->level_w[d], 0, 0, 0, 0 } );
      { const int64_t g0 = ( int64_t )gp.size();
        for ( int s : lv ) { bool any = false; for ( int ci = meta[s].child_begin; ci < meta[s].child_end; ci++ ) { const SnMeta& ch = meta[child_list[ci]]; if ( ch.f > ch.k ) { any = true; break; } }
          if ( any ) for ( int lo = 0; lo < meta[s].f; lo += GATHER_ROWS ) gp.push_back( { s, lo, std::min( meta[s].f, lo + GATHER_ROWS ) } ); }
        if ( ( int64_t )gp.size() > g0 ) SS.push_back( { SK_FWD_GATHER, d, g0, ( int64_t )gp.size() - g0, 0, 0, 0, 0 } ); }
      if ( small_rng[d].second > 0 ) SS.push_back( { SK_FWD_SMALL, d, small_rng[d].first, small_rng[d].second, 0, 0, 0, 0 } );
      const std::vector<int>& bl = big[d];
      int nblk = 0; for ( int s : bl ) nblk = std::max( nblk, ( meta[s].k + meta[s].sbw - 1 ) / meta[s].sbw );
      for ( int t = 0; t < nblk; t++ ) {
        Mark m0 = mark( false );
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 );
          SvTask x; x.M = c->dSinvF + m.sinv + ( int64_t )t * m.sbw * m.sbw; x.ld = ( nb + 1 ) & ~1; x.K = nb; x.N = nb; x.tri = 1; x.mode = 1;
          x.vsel = 0; x.voff = m.c0 + j0; x.dsel = 1; x.doff = m.c0 + j0; add_task( x, false ); }
        close( SK_FWD_TRI, d, m0, false );
        m0 = mark( false );
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 );
          SvTask x; x.ld = m.ld; x.K = nb; x.tri = 0; x.mode = 0; x.vsel = 1; x.voff = m.c0 + j0;
          x.M = Fb + m.lptr + ( j0 + nb ) + ( int64_t )j0 * m.ld; x.N = m.k - j0 - nb; x.dsel = 0; x.doff = m.c0 + j0 + nb; add_task( x, false );      /* pivot rows below the block */
          x.M = Fb + m.lptr + m.k + ( int64_t )j0 * m.ld; x.N = m.f - m.k; x.dsel = wsel; x.doff = m.wptr; add_task( x, false ); }                     /* update rows */
        close( SK_FWD_UPD, d, m0, false );
      }
    }
    for ( int d = 0; d <= S->maxdepth; d++ ) {
      const std::vector<int>& lv = levels[d];
      const int wsel = 2 + ( d & 1 );
      if ( small_rng[d].second > 0 ) SS.push_back( { SK_BWD_SMALL, d, small_rng[d].first, small_rng[d].second, 0, 0, 0, 0 } );
      const std::vector<int>& bl = big[d];
      { const int64_t g0 = ( int64_t )bg.size();
        for ( int s : bl ) { const int u = meta[s].f - meta[s].k; for ( int lo = 0; lo < u; lo += GATHER_ROWS ) bg.push_back( { s, lo, std::min( u, lo + GATHER_ROWS ) } ); }
        if ( ( int64_t )bg.size() > g0 ) SS.push_back( { SK_BWD_GATHER, d, g0, ( int64_t )bg.size() - g0, 0, 0, 0, 0 } ); }
      { Mark m0 = mark( true );
        for ( int s : bl ) { const SnMeta& m = meta[s];
          SvTask x; x.M = Mb + m.lptr + m.k; x.ld = m.ld; x.K = m.f - m.k; x.N = m.k; x.tri = 0; x.mode = 0; x.vsel = wsel; x.voff = m.wptr; x.dsel = 1; x.doff = m.c0; add_task( x, true ); }
        close( SK_BWD_UPD, d, m0, true ); }
      int nblk = 0; for ( int s : bl ) nblk = std::max( nblk, ( meta[s].k + meta[s].sbw - 1 ) / meta[s].sbw );
      for ( int t = nblk - 1; t >= 0; t-- ) {
        Mark m0 = mark( false );
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 );
          SvTask x; x.M = c->dSinvB + m.sinv + ( int64_t )t * m.sbw * m.sbw; x.ld = ( nb + 1 ) & ~1; x.K = nb; x.N = nb; x.tri = 2; x.mode = 1;
          x.vsel = 1; x.voff = m.c0 + j0; x.dsel = 0; x.doff = m.c0 + j0; add_task( x, false ); }
        close( SK_BWD_TRI, d, m0, false );
        if ( t == 0 ) break;
        m0 = mark( true );
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 );
          SvTask x; x.M = Mb + m.lptr + j0; x.ld = m.ld; x.K = nb; x.N = j0; x.tri = 0; x.mode = 0; x.vsel = 0; x.voff = m.c0 + j0; x.dsel = 1; x.doff = m.c0; add_task( x, true ); }
        close( SK_BWD_UPD, d, m0, true );
      }
    }
    if ( c->sv_tasks.upload( tk ) || c->sv_tiles_v.upload( tv ) || c->sv_tiles_t.upload( tt ) || c->sv_tiles_g.upload( tg ) || c->gather_tasks.upload( gp ) || c->bgather_tasks.upload( bg ) || c->sv_small.upload( sm ) ) return B200_ERR_CUDA;
    if ( !c->dry && tk.size() ) CK( cudaMalloc( &c->dSvGemm, tk.size() * sizeof( SvGemm ) ) );
    c->sv_gemm_p = 0;
  }
  /* ---- persistent sweeps (b200_sweep.cuh): the products of the schedule above as ordered units with device-side dependencies.
   * Counters (one int each, zeroed before a forward sweep): per supernode KD units of its children done, GD gather units done, FT / BT
   * triangular-product tiles done (forward / backward), BD its units done in the backward sweep; rt0[s] + tile: block updates applied to 32
   * pivot rows (then to 32 update rows) of s; ct0[s] + tile: products applied to 64 pivot columns of s in the backward sweep. */
  {
    std::vector<SvTask> tk; std::vector<SvUnit> uf, ub;
    const bool use_small = getenv( "B200_NO_SMALL_SOLVE" ) == nullptr;
    auto is_small = [&]( int s ) { return use_small && meta[s].k <= KC && meta[s].f <= 4 * SVS_THREADS; };
    int64_t off = 0;
    const int64_t KD = off; off += ns; const int64_t GD = off; off += ns; const int64_t FT = off; off += ns;
    const int64_t BD = off; off += ns; const int64_t BT = off; off += ns;
    std::vector<int64_t> rt0( ns ), ct0( ns );
    for ( int s = 0; s < ns; s++ ) { const int k = meta[s].k, u = meta[s].f - meta[s].k; rt0[s] = off; off += ( k + 31 ) / 32 + ( u + 31 ) / 32; ct0[s] = off; off += ( k + 63 ) / 64; }
    c->sw_head = off; off += 4;                      /* forward ticket, backward ticket, error flag */
    c->sw_ncnt = off;
    c->use_sweep = getenv( "B200_SOLVE_LEVELSET" ) == nullptr && off < ( int64_t )2000000000;
    auto unit = [&]( int kind, int task, int o0, int aux, int sn ) { SvUnit u; memset( &u, 0, sizeof( u ) ); u.kind = kind; u.task = task; u.o0 = o0; u.aux = aux; u.sn = sn; u.w2 = -1; u.sig = -1; u.up = -1; return u; };
    const double* Fb = c->dL; const double* Mb = TWO ? c->dU : c->dL;
    std::vector<int> Tf( ns, 0 ), Tb( ns, 0 ), ngath( ns, 0 ), cum( ns, 0 );
    if ( c->use_sweep ) {
    for ( int d = S->maxdepth; d >= 0; d-- ) {
      std::vector<int> bl;
      for ( int s : levels[d] ) {
        if ( !is_small( s ) ) { bl.push_back( s ); continue; }
        SvUnit U = unit( UK_FWD_SMALL, 0, 0, 0, s ); const int nch = meta[s].child_end - meta[s].child_begin;
        if ( nch > 0 ) { U.w2 = ( int )( KD + s ); U.v2 = -1; }      /* value filled in below: the units of all its children */
        uf.push_back( U ); Tf[s]++;
      }
      for ( int s : bl ) {
        const SnMeta& m = meta[s]; bool any = false;
        for ( int ci = m.child_begin; ci < m.child_end; ci++ ) { const SnMeta& ch = meta[child_list[ci]]; if ( ch.f > ch.k ) { any = true; break; } }
        if ( !any ) continue;
        for ( int lo = 0; lo < m.f; lo += GATHER_ROWS ) {
          SvUnit U = unit( UK_FWD_GATHER, 0, lo, std::min( m.f, lo + GATHER_ROWS ), s ); U.w2 = ( int )( KD + s ); U.v2 = -1; U.sig = ( int )( GD + s );
          uf.push_back( U ); Tf[s]++; ngath[s]++;
        }
      }
      int nblk = 0; for ( int s : bl ) nblk = std::max( nblk, ( meta[s].k + meta[s].sbw - 1 ) / meta[s].sbw );
      for ( int t = 0; t < nblk; t++ ) {
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 );
          SvTask x; x.M = c->dSinvF + m.sinv + ( int64_t )t * m.sbw * m.sbw; x.ld = ( nb + 1 ) & ~1; x.K = nb; x.N = nb; x.tri = 1; x.mode = 1;
          x.vsel = 0; x.voff = m.c0 + j0; x.dsel = 1; x.doff = m.c0 + j0;
          const int id = ( int )tk.size(); tk.push_back( x );
          for ( int o0 = 0; o0 < nb; o0 += GV_ROWS ) {
            SvUnit U = unit( UK_GEMV, id, o0, 0, s );
            if ( t > 0 ) { U.w1 = ( int )( rt0[s] + j0 / 32 ); U.n1 = ( nb + 31 ) / 32; U.v1 = t; }
            else if ( ngath[s] > 0 ) { U.w2 = ( int )( GD + s ); U.v2 = ngath[s]; }
            U.sig = ( int )( FT + s ); uf.push_back( U ); Tf[s]++; cum[s]++;
          } }
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 );
          for ( int part = 0; part < 2; part++ ) {
            SvTask x; x.ld = m.ld; x.K = nb; x.tri = 0; x.mode = 0; x.vsel = 1; x.voff = m.c0 + j0;
            int tile0;
            if ( part == 0 ) { x.M = Fb + m.lptr + ( j0 + nb ) + ( int64_t )j0 * m.ld; x.N = m.k - j0 - nb; x.dsel = 0; x.doff = m.c0 + j0 + nb; tile0 = ( j0 + nb ) / 32; }      /* pivot rows below the block (the next block's rows first) */
            else { x.M = Fb + m.lptr + m.k + ( int64_t )j0 * m.ld; x.N = m.f - m.k; x.dsel = 2; x.doff = m.wgl; tile0 = ( m.k + 31 ) / 32; }                                    /* update rows */
            if ( x.N <= 0 ) continue;
            const int id = ( int )tk.size(); tk.push_back( x );
            for ( int o0 = 0; o0 < x.N; o0 += GV_ROWS ) {
              SvUnit U = unit( UK_GEMV, id, o0, 0, s ); const int tile = tile0 + o0 / 32;
              if ( t > 0 ) { U.w1 = ( int )( rt0[s] + tile ); U.n1 = 1; U.v1 = t; }
              U.w2 = ( int )( FT + s ); U.v2 = cum[s]; U.sig = ( int )( rt0[s] + tile ); uf.push_back( U ); Tf[s]++;
            }
          } }
      }
    }
    { /* completion: every unit of s raises its parent's KD counter; the parent's first units wait for the units of ALL its children */
      std::vector<int> kids( ns, 0 );
      for ( int s = 0; s < ns; s++ ) if ( meta[s].parent >= 0 ) kids[meta[s].parent] += Tf[s];
      for ( SvUnit& U : uf ) { U.up = meta[U.sn].parent >= 0 ? ( int )( KD + meta[U.sn].parent ) : -1; if ( U.w2 == ( int )( KD + U.sn ) ) U.v2 = kids[U.sn]; }
    }
    std::fill( cum.begin(), cum.end(), 0 );
    for ( int d = 0; d <= S->maxdepth; d++ ) {
      std::vector<int> bl;
      for ( int s : levels[d] ) {
        if ( !is_small( s ) ) { bl.push_back( s ); continue; }
        SvUnit U = unit( UK_BWD_SMALL, 0, 0, 0, s );
        if ( meta[s].parent >= 0 ) { U.w2 = ( int )( BD + meta[s].parent ); U.v2 = -1; }
        ub.push_back( U ); Tb[s]++;
      }
      for ( int s : bl ) { const SnMeta& m = meta[s]; const int u = m.f - m.k; if ( u <= 0 ) continue;       /* z[pivots] -= M[update rows, :]' x[ancestors], x through the row list */
        SvTask x; x.M = Mb + m.lptr + m.k; x.ld = m.ld; x.K = u; x.N = m.k; x.tri = 0; x.mode = 0; x.vsel = 0; x.voff = 0; x.dsel = 1; x.doff = m.c0;
        const int id = ( int )tk.size(); tk.push_back( x );
        for ( int o0 = ( ( m.k - 1 ) / SW_GT_COLS ) * SW_GT_COLS; o0 >= 0; o0 -= SW_GT_COLS ) {      /* the last columns first: the first triangular product waits for them */
          SvUnit U = unit( UK_GEMVT, id, o0, 1, s );
          if ( m.parent >= 0 ) { U.w2 = ( int )( BD + m.parent ); U.v2 = -1; }
          U.sig = ( int )( ct0[s] + o0 / 64 ); ub.push_back( U ); Tb[s]++;
        } }
      int nblk = 0; for ( int s : bl ) nblk = std::max( nblk, ( meta[s].k + meta[s].sbw - 1 ) / meta[s].sbw );
      for ( int t = nblk - 1; t >= 0; t-- ) {
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 ); const int need = ( m.f > m.k ? 1 : 0 ) + ( ( m.k + m.sbw - 1 ) / m.sbw - 1 - t );
          SvTask x; x.M = c->dSinvB + m.sinv + ( int64_t )t * m.sbw * m.sbw; x.ld = ( nb + 1 ) & ~1; x.K = nb; x.N = nb; x.tri = 2; x.mode = 1;
          x.vsel = 1; x.voff = m.c0 + j0; x.dsel = 0; x.doff = m.c0 + j0;
          const int id = ( int )tk.size(); tk.push_back( x );
          for ( int o0 = 0; o0 < nb; o0 += GV_ROWS ) {
            SvUnit U = unit( UK_GEMV, id, o0, 0, s );
            if ( need > 0 ) { U.w1 = ( int )( ct0[s] + j0 / 64 ); U.n1 = ( nb + 63 ) / 64; U.v1 = need; }
            else if ( m.parent >= 0 ) { U.w2 = ( int )( BD + m.parent ); U.v2 = -1; }
            U.sig = ( int )( BT + s ); ub.push_back( U ); Tb[s]++; cum[s]++;
          } }
        if ( t == 0 ) break;
        for ( int s : bl ) { const SnMeta& m = meta[s]; const int j0 = t * m.sbw; if ( j0 >= m.k ) continue;
          const int nb = std::min( m.sbw, m.k - j0 ); const int need = ( m.f > m.k ? 1 : 0 ) + ( ( m.k + m.sbw - 1 ) / m.sbw - 1 - t );
          SvTask x; x.M = Mb + m.lptr + j0; x.ld = m.ld; x.K = nb; x.N = j0; x.tri = 0; x.mode = 0; x.vsel = 0; x.voff = m.c0 + j0; x.dsel = 1; x.doff = m.c0;
          const int id = ( int )tk.size(); tk.push_back( x );
          for ( int o0 = j0 - SW_GT_COLS; o0 >= 0; o0 -= SW_GT_COLS ) {      /* the previous block's columns first */
            SvUnit U = unit( UK_GEMVT, id, o0, 0, s ); const int tile = o0 / 64;
            if ( need > 0 ) { U.w1 = ( int )( ct0[s] + tile ); U.n1 = 1; U.v1 = need; }
            U.w2 = ( int )( BT + s ); U.v2 = cum[s]; U.sig = ( int )( ct0[s] + tile ); ub.push_back( U ); Tb[s]++;
          } }
      }
    }
    for ( SvUnit& U : ub ) { U.up = ( int )( BD + U.sn ); if ( meta[U.sn].parent >= 0 && U.w2 == ( int )( BD + meta[U.sn].parent ) ) U.v2 = Tb[meta[U.sn].parent]; }      /* x of all ancestors is final once every unit of the parent is done */
    }
    if ( c->dry ) { c->h_sw_f = uf; c->h_sw_b = ub; c->h_sw_t = tk; }      /* exported one unit per tile / supernode: that is what tests/sweep_sim.py checks */
    /* device lists: runs of latency-bound units (small supernodes, gemv tiles with K <= SW_WARP_K) become UK_BATCH units of up to SW_BATCH,
     * one inner unit per warp (waits and signals stay per inner unit) */
    std::vector<SvUnit> inner;
    auto warp_unit = [&]( const SvUnit& U ) { return U.kind == UK_FWD_SMALL || U.kind == UK_BWD_SMALL || ( U.kind == UK_GEMV && tk[U.task].K <= SW_WARP_K ); };
    auto pack = [&]( std::vector<SvUnit>& in ) {
      std::vector<SvUnit> out; out.reserve( in.size() / 4 + 16 );
      for ( size_t i = 0; i < in.size(); ) {
        if ( !warp_unit( in[i] ) ) { out.push_back( in[i++] ); continue; }
        SvUnit B = unit( UK_BATCH, ( int )inner.size(), 0, 0, in[i].sn );
        while ( i < in.size() && warp_unit( in[i] ) && B.o0 < SW_BATCH ) { inner.push_back( in[i] ); B.o0++; i++; }
        out.push_back( B );
      }
      in.swap( out );
    };
    pack( uf ); pack( ub );
    if ( c->sw_tasks.upload( tk ) || c->sw_inner.upload( inner ) || c->sw_units[0].upload( uf ) || c->sw_units[1].upload( ub ) ) return B200_ERR_CUDA;
    if ( !c->dry && c->use_sweep && getenv( "B200_SWEEP_TRACE" ) ) { const size_t ntr = uf.size() + ub.size() + 1; CK( cudaMalloc( &c->dSwTrace, ntr * 5 * sizeof( unsigned long long ) ) ); CK( cudaMemset( c->dSwTrace, 0, ntr * 5 * sizeof( unsigned long long ) ) ); }
    if ( !c->dry && c->use_sweep ) { CK( cudaMalloc( &c->dSwCnt, ( size_t )c->sw_ncnt * sizeof( int ) ) ); CK( cudaMalloc( &c->dWg, ( ( size_t )c->sw_w_rows + 2 ) * 4 * 8 ) ); }
    c->device_bytes += c->sw_tasks.bytes() + c->sw_units[0].bytes() + c->sw_units[1].bytes() + c->sw_inner.bytes() + c->sw_ncnt * 4 + ( c->sw_w_rows + 2 ) * 32;
  }
  c->device_bytes += c->gemm_tasks.bytes() + c->tiles_big.bytes() + c->tiles_small.bytes() + c->diag_tasks.bytes() + c->trsm_tasks.bytes() + c->trsm_tiles.bytes() +
                     c->asm_tasks.bytes() + c->sv_tasks.bytes() * 2 + c->sv_tiles_v.bytes() + c->sv_tiles_t.bytes() + c->sv_tiles_g.bytes() + c->rowidx.bytes() * 2 + c->sn.bytes();
  return B200_OK;
}

/* ------------------------------------------------------------- factor */
static int run_factor( b200_ctx* c ) {
  const bool LU = c->kind == B200_KIND_LU, LD = c->kind == B200_KIND_LDLT, TWO = LU || LD;
  cudaStream_t st = c->stream;
  CK( cudaEventRecord( c->ev0, st ) );
  CK( cudaMemsetAsync( c->dErr, 0, sizeof( int ), st ) );
  CK( cudaMemsetAsync( c->dPert, 0, sizeof( int ), st ) );
  CK( cudaMemsetAsync( c->dL, 0, ( size_t )c->lsize * 8, st ) );
  if ( TWO ) CK( cudaMemsetAsync( c->dU, 0, ( size_t )c->lsize * 8, st ) );
  if ( c->nnzA > 0 ) {
    const int blocks = ( int )std::min<int64_t>( ( c->nnzA + 255 ) / 256, 148 * 16 );
    scatter_A_kernel<<<blocks, 256, 0, st>>>( c->dVals, c->amap.d, c->nnzA, c->dL, c->dU );
  }
  if ( TWO ) { /* static-pivot threshold sqrt(eps) * max|a_ij| */
    CK( cudaMemsetAsync( c->dMaxBits, 0, 8, st ) );
    if ( c->nnzA > 0 ) maxabs_kernel<<<( int )std::min<int64_t>( ( c->nnzA + 255 ) / 256, 148 * 8 ), 256, 0, st>>>( c->dVals, c->nnzA, c->dMaxBits );
    tiny_from_max_kernel<<<1, 1, 0, st>>>( c->dMaxBits, c->dTiny );
  }
  if ( c->world > 1 ) { CK( cudaEventRecord( c->evs, st ) ); CK( cudaStreamWaitEvent( c->stream3, c->evs, 0 ) ); CK( cudaStreamWaitEvent( c->stream4, c->evs, 0 ) ); }   /* nothing may land in L before it has been zeroed and scattered */
  int64_t nlaunch = 3 + ( TWO ? 3 : 0 );
  double gemm_flops = 0; int64_t gemm_launches = 0;
  for ( int k = 0; k < LK_COUNT; k++ ) { c->prof_ms[k] = 0; c->prof_n[k] = 0; }
  int cur_depth = -1; size_t nlev = 0;
  for ( const Launch& l : c->launches ) {
    if ( l.stream == 0 && l.depth != cur_depth ) {   /* time stamp on the update stream when a new tree level begins */
      if ( c->lvl_ev.size() <= nlev ) { cudaEvent_t e; CK( cudaEventCreate( &e ) ); c->lvl_ev.push_back( e ); }
      CK( cudaEventRecord( c->lvl_ev[nlev++], c->stream ) ); cur_depth = l.depth;
    }
    cudaStream_t ls = l.stream == 0 ? c->stream : l.stream == 1 ? c->stream2 : l.stream == 2 ? c->stream3 : l.stream == 3 ? c->stream4 : c->stream5;
    if ( l.wait_ev >= 0 ) CK( cudaStreamWaitEvent( ls, c->events[l.wait_ev], 0 ) );
    if ( c->profile ) CK( cudaEventRecord( c->evp0, ls ) );
    switch ( l.kind ) {
      case LK_MEMSET_CB: CK( cudaMemsetAsync( c->dCB[l.depth & 1] + l.first, 0, ( size_t )l.count * 8, ls ) ); break;
      case LK_ASM:
        if ( LU ) extend_add_kernel<true><<<( unsigned )l.count, 256, 0, ls>>>( c->asm_tasks.d + l.first, c->sn.d, c->child_list.d, c->relidx.d, c->dL, c->dU, c->dCB[l.depth & 1], c->dCB[( l.depth + 1 ) & 1] );
        else extend_add_kernel<false><<<( unsigned )l.count, 256, 0, ls>>>( c->asm_tasks.d + l.first, c->sn.d, c->child_list.d, c->relidx.d, c->dL, c->dU, c->dCB[l.depth & 1], c->dCB[( l.depth + 1 ) & 1] );
        break;
      case LK_DIAG:
        if ( LD ) ldlt_diag_kernel<<<( unsigned )l.count, 256, LDLT_SMEM, ls>>>( c->diag_tasks.d + l.first, c->dTiny, c->dPert, c->dErr );
        else if ( LU ) getrf_diag_kernel<<<( unsigned )l.count, 256, GETRF_SMEM, ls>>>( c->diag_tasks.d + l.first, c->dTiny, c->dPert, c->dErr );
        else potrf_diag_kernel<<<( unsigned )l.count, POTRF_THREADS, 0, ls>>>( c->diag_tasks.d + l.first, c->dErr );
        break;
      case LK_SMALL_FRONT: small_front_kernel<<<( unsigned )l.count, 256, SF_SMEM, ls>>>( c->small_sns.d + l.first, c->sn.d, c->dL, c->dInv, c->dCB[l.depth & 1], c->dErr ); break;
      case LK_DIAG_S: potrf_diag_small_kernel<<<( unsigned )l.count, 256, POTRF_S_SMEM, ls>>>( c->diag_tasks.d + l.first, c->dErr ); break;
      case LK_TRSM: trsm_rows_kernel<<<( unsigned )l.count, 256, TRSM_SMEM, ls>>>( c->trsm_tasks.d, c->trsm_tiles.d + l.first ); break;
      case LK_TRSM_MMA: gemm_nt_dmma_bulk_kernel<GEMM_CFG><<<( unsigned )l.count, BULK_THREADS, BULK_SMEM, ls>>>( c->gemm_tasks.d, c->tiles_big.d + l.first, c->dZeros ); break;
      case LK_GEMM_BIG:
        gemm_nt_dmma_bulk_kernel<GEMM_CFG><<<( unsigned )l.count, BULK_THREADS, BULK_SMEM, ls>>>( c->gemm_tasks.d, c->tiles_big.d + l.first, c->dZeros );
        gemm_flops += l.flops; gemm_launches++; break;
      case LK_NOP: break;
      case LK_XFER_CB: case LK_BCAST: case LK_GATHER: {
        NK( g_nccl.GroupStart() );
        for ( int64_t i = l.first; i < l.first + l.count; i++ ) {
          const Msg& m = c->msgs[i];
          double* base = m.which == 0 ? c->dL : m.which == 1 ? c->dInv : c->dCB[m.which - 2];
          ncclComm_t cm = l.stream == 3 ? c->comm2 : c->comm;
          if ( m.src == c->rank ) NK( g_nccl.Send( base + m.off, ( size_t )m.cnt, ncclDouble, m.dst, cm, ls ) );
          else NK( g_nccl.Recv( base + m.off, ( size_t )m.cnt, ncclDouble, m.src, cm, ls ) );
        }
        NK( g_nccl.GroupEnd() );
        break; }
    }
    if ( l.record_ev >= 0 ) CK( cudaEventRecord( c->events[l.record_ev], ls ) );
    nlaunch++;
    if ( c->profile ) {
      CK( cudaEventRecord( c->evp1, ls ) ); CK( cudaEventSynchronize( c->evp1 ) );
      float ms = 0; cudaEventElapsedTime( &ms, c->evp0, c->evp1 ); c->prof_ms[l.kind] += ms; c->prof_n[l.kind]++;
      static const bool trace = getenv( "B200_TRACE_GEMM" ) != nullptr;
      static const bool trace_all = getenv( "B200_TRACE_ALL" ) != nullptr;
      if ( trace_all ) fprintf( stderr, "launch %s depth=%d count=%lld stream=%d ms=%.4f flops=%.4e\n", kind_name[l.kind], l.depth, ( long long )l.count, l.stream, ms, l.flops );
      if ( trace && l.kind == LK_GEMM_BIG && ms > 0.5f )
        fprintf( stderr, "gemm depth=%d tiles=%lld flops=%.3e ms=%.3f TF/s=%.2f\n", l.depth, ( long long )l.count, l.flops, ms, l.flops / ( ms * 1e-3 ) / 1e12 );
    }
  }
  /* explicit inverses of the solve blocks (feeds the sweeps; part of the factorization time) */
  if ( c->sinv_tasks.n && ( c->world == 1 || c->rank == 0 || c->replicate ) ) { build_sinv_kernel<<<( unsigned )c->sinv_tasks.n, 256, SINV_SMEM, st>>>( c->sinv_tasks.d ); nlaunch++; }
  int nviol = 0; unsigned long long gbits = 0;
  if ( LU ) { /* threshold partial pivoting verified over all fully-summed rows (part of the factorization time) */
    CK( cudaMemsetAsync( c->dMaxBits, 0, 8, st ) ); CK( cudaMemsetAsync( c->dViol, 0, sizeof( int ), st ) );
    lu_growth_kernel<<<148 * 8, 256, 0, st>>>( c->sn.d, c->ns, c->dL, 1.0 / c->pivot_threshold, c->dMaxBits, c->dViol ); nlaunch++;
  }
  CK( cudaEventRecord( c->ev1, st ) );
  CK( cudaGetLastError() );
  int err = 0, pert = 0;
  CK( cudaMemcpyAsync( &err, c->dErr, sizeof( int ), cudaMemcpyDeviceToHost, st ) );
  CK( cudaMemcpyAsync( &pert, c->dPert, sizeof( int ), cudaMemcpyDeviceToHost, st ) );
  if ( LU ) { CK( cudaMemcpyAsync( &nviol, c->dViol, sizeof( int ), cudaMemcpyDeviceToHost, st ) ); CK( cudaMemcpyAsync( &gbits, c->dMaxBits, 8, cudaMemcpyDeviceToHost, st ) ); }
  CK( cudaStreamSynchronize( st ) );
  { double g; memcpy( &g, &gbits, 8 ); c->stats.lu_max_multiplier = LU ? std::max( g, c->ns > 0 ? 1.0 : 0.0 ) : 0.0; c->stats.npiv_threshold_violations = nviol; c->stats.pivot_threshold = LU ? c->pivot_threshold : 0.0; }
  float ms = 0; cudaEventElapsedTime( &ms, c->ev0, c->ev1 );
  c->lvl_ms.assign( nlev, 0.f );
  for ( size_t i = 0; i < nlev; i++ ) { cudaEvent_t nx = i + 1 < nlev ? c->lvl_ev[i + 1] : c->ev1; cudaEventElapsedTime( &c->lvl_ms[i], c->lvl_ev[i], nx ); }
  c->stats.factor_ms = ms; c->stats.launches = nlaunch; c->stats.gemm_flops = gemm_flops; c->stats.gemm_launches = gemm_launches;
  c->stats.gemm_alg_flops = c->gemm_alg_flops * ( c->flops_stored > 0 ? c->flops_exact / c->flops_stored : 1.0 );
  c->stats.gemm_ms = c->profile ? c->prof_ms[LK_GEMM_BIG] : 0.0;
  c->stats.npiv_perturbed = pert; c->stats.device_bytes = c->device_bytes; c->stats.comm_bytes = c->comm_bytes;
  if ( err < 0 ) { char a[64]; snprintf( a, 64, "%d", -err - 1 ); set_err( "NaN in pivot column %s of the permuted matrix: the matrix is singular or holds non-finite values%s", a, "" ); c->factored = false; return B200_ERR_SINGULAR; }
  if ( err != 0 ) { char a[64]; snprintf( a, 64, "%d", err - 1 ); set_err( "matrix is not positive definite (pivot column %s of the permuted matrix)%s", a, "" ); c->factored = false; return B200_ERR_NOT_SPD; }
  c->factored = true;
  return B200_OK;
}

extern "C" int b200_factor_device( b200_ctx_t* c, const double* d_vals ) {
  if ( !c || !c->dL || c->dry ) return B200_ERR_STATE;
  CK( cudaSetDevice( c->device ) );
  if ( d_vals != c->dVals && c->nnzA > 0 ) CK( cudaMemcpyAsync( c->dVals, d_vals, ( size_t )c->nnzA * 8, cudaMemcpyDeviceToDevice, c->stream ) );
  c->stats.h2d_ms = 0;
  return run_factor( c );
}

extern "C" int b200_factor_host( b200_ctx_t* c, const double* vals ) {
  if ( !c || !c->dL || c->dry ) return B200_ERR_STATE;
  CK( cudaSetDevice( c->device ) );
  CK( cudaEventRecord( c->evp0, c->stream ) );
  if ( c->nnzA > 0 ) CK( cudaMemcpyAsync( c->dVals, vals, ( size_t )c->nnzA * 8, cudaMemcpyHostToDevice, c->stream ) );
  CK( cudaEventRecord( c->evp1, c->stream ) );
  int r = run_factor( c );
  float ms = 0; cudaEventElapsedTime( &ms, c->evp0, c->evp1 ); c->stats.h2d_ms = ms;
  return r;
}

/* -------------------------------------------------------------- solve */
static inline int solve_pitch( int p ) { return p <= 4 ? p : ( p + 1 ) & ~1; }     /* tensor path: even pitch = 16-byte aligned rows for the bulk copies */
static int ensure_solve_ws( b200_ctx* c, int p ) {
  if ( p <= c->solve_p_cap ) return B200_OK;
  cudaFree( c->dY ); cudaFree( c->dW[0] ); cudaFree( c->dW[1] ); cudaFree( c->dR ); cudaFree( c->dD ); cudaFree( c->dZ );
  c->dY = c->dW[0] = c->dW[1] = c->dR = c->dD = c->dZ = nullptr; c->solve_p_cap = 0; c->sv_gemm_p = 0;
  const size_t pitch = ( size_t )solve_pitch( p ) + 2;      /* capacity for any p' <= p */
  CK( cudaMalloc( &c->dY, ( ( size_t )c->n + 2 ) * pitch * 8 ) ); CK( cudaMalloc( &c->dZ, ( ( size_t )c->n + 2 ) * pitch * 8 ) );
  CK( cudaMalloc( &c->dR, ( size_t )c->n * p * 8 ) ); CK( cudaMalloc( &c->dD, ( size_t )c->n * p * 8 ) );
  for ( int a = 0; a < 2; a++ ) CK( cudaMalloc( &c->dW[a], ( ( size_t )std::max<int64_t>( c->w_rows[a], 1 ) + 2 ) * pitch * 8 ) );
  c->solve_p_cap = p;
  return B200_OK;
}

/* one pair of sweeps for p <= 4 as two persistent kernels (b200_sweep.cuh): units are taken from an ordered list and wait on the device
 * for their inputs, so there is no launch per tree level / solve block and independent fronts overlap */
template <int PCT>
static int launch_sweep( b200_ctx* c, const SweepArgs& a, int occ, cudaStream_t st ) {
  if ( a.nunits <= 0 ) return B200_OK;
  const int grid = ( int )std::min<int64_t>( a.nunits, ( int64_t )c->prop.multiProcessorCount * std::max( occ, 1 ) );
  sv_sweep_kernel<PCT><<<grid, SW_THREADS, sw_smem<PCT>(), st>>>( a );
  return B200_OK;
}
static int sweep_persistent( b200_ctx* c, const double* d_b, int p, double* d_x, int64_t& nl ) {
  cudaStream_t st = c->stream;
  const int n = c->n, pitch = p;
  const int pblocks = ( int )std::min<int64_t>( ( ( int64_t )n * pitch + 255 ) / 256, 148 * 32 );
  permute_in_kernel<<<pblocks, 256, 0, st>>>( d_b, c->perm.d, n, p, pitch, c->dY );
  CK( cudaMemsetAsync( c->dSwCnt, 0, ( size_t )( c->sw_head + 2 ) * sizeof( int ), st ) );      /* counters and the two tickets; the error flag behind them is cleared once per solve */
  if ( c->sw_w_rows > 0 ) CK( cudaMemsetAsync( c->dWg, 0, ( size_t )c->sw_w_rows * pitch * 8, st ) );
  SweepArgs a;
  a.cnt = c->dSwCnt; a.err = c->dSwCnt + c->sw_head + 2; a.tasks = c->sw_tasks.d; a.inner = c->sw_inner.d; a.sn = c->sn.d; a.child_list = c->child_list.d; a.relidx = c->relidx.d; a.rowidx = c->rowidx.d;
  a.Lf = c->dL; a.Lb = c->kind == B200_KIND_CHOLESKY ? c->dL : c->dU; a.sinvF = c->dSinvF; a.sinvB = c->dSinvB;
  a.B.b[0] = c->dY; a.B.b[1] = c->dZ; a.B.b[2] = c->dWg; a.B.b[3] = nullptr; a.p = p; a.pitch = pitch;
  const int pi = p <= 1 ? 0 : p <= 2 ? 1 : 2;
  for ( int dir = 0; dir < 2; dir++ ) {
    a.units = c->sw_units[dir].d; a.nunits = ( int )c->sw_units[dir].n; a.head = c->dSwCnt + c->sw_head + dir;
    a.trace = c->dSwTrace ? c->dSwTrace + ( dir == 0 ? 0 : 5 * c->sw_units[0].n ) : nullptr;
    if ( c->profile ) CK( cudaEventRecord( c->evp0, st ) );
    if ( pi == 0 ) launch_sweep<1>( c, a, c->sw_occ[0], st ); else if ( pi == 1 ) launch_sweep<2>( c, a, c->sw_occ[1], st ); else launch_sweep<4>( c, a, c->sw_occ[2], st );
    if ( c->profile ) { CK( cudaEventRecord( c->evp1, st ) ); CK( cudaEventSynchronize( c->evp1 ) ); float ms = 0; cudaEventElapsedTime( &ms, c->evp0, c->evp1 ); c->sprof_ms[SK_SWEEP_FWD + dir] += ms; c->sprof_n[SK_SWEEP_FWD + dir]++; }
  }
  permute_out_kernel<<<pblocks, 256, 0, st>>>( c->dY, c->perm.d, n, p, pitch, d_x );
  CK( cudaMemcpyAsync( &c->sw_err_h, c->dSwCnt + c->sw_head + 2, sizeof( int ), cudaMemcpyDeviceToHost, st ) );
  nl += 7;
  return B200_OK;
}

/* one pair of sweeps: d_x = (LL')^-1 d_b  (or (LU)^-1), launches counted into nl */
static int sweep_once( b200_ctx* c, const double* d_b, int p, double* d_x, int64_t& nl ) {
  if ( p <= c->sweep_pmax && c->use_sweep ) return sweep_persistent( c, d_b, p, d_x, nl );
  cudaStream_t st = c->stream;
  const int n = c->n;
  const int pitch = solve_pitch( p );
  const bool tensor = p > 4;
  const int pct = p <= 1 ? 1 : p <= 2 ? 2 : 4;
  const int pblocks = ( int )std::min<int64_t>( ( ( int64_t )n * pitch + 255 ) / 256, 148 * 32 );
  SvBases B; B.b[0] = c->dY; B.b[1] = c->dZ; B.b[2] = c->dW[0]; B.b[3] = c->dW[1];
  if ( tensor && c->sv_gemm_p != p && c->sv_tasks.n ) {      /* the task list is the same for every p; only the vector addresses / pitch / M change */
    sv_expand_kernel<<<( unsigned )( ( c->sv_tasks.n + 255 ) / 256 ), 256, 0, st>>>( c->sv_tasks.d, ( int64_t )c->sv_tasks.n, B, p, pitch, c->dSvGemm );
    c->sv_gemm_p = p; nl++;
  }
  permute_in_kernel<<<pblocks, 256, 0, st>>>( d_b, c->perm.d, n, p, pitch, c->dY );
  nl += 2;
  const unsigned gy = ( unsigned )( ( p + 63 ) / 64 );
  int tqs = 0; while ( tqs < 5 && ( 1 << tqs ) < pitch ) tqs++;      /* gather kernels: 2^tqs threads along the right-hand sides of a row */
  /* the fused small-supernode kernel of a level is independent of the level's generic launches: it runs on a second stream,
   * forked after the level's gather and joined before the next level (or the turn of the sweep) begins */
  cudaStream_t st2 = c->profile ? st : c->stream2;
  bool forked = false; int cur_depth = -1, cur_dir = 0;
  for ( const SLaunch& l : c->slaunches ) {
    const int a = l.depth & 1, ac = ( l.depth + 1 ) & 1;
    const int dir = ( l.kind == SK_BWD_GATHER || l.kind == SK_BWD_UPD || l.kind == SK_BWD_TRI || l.kind == SK_BWD_SMALL ) ? 1 : 0;
    if ( ( l.depth != cur_depth || dir != cur_dir ) && forked ) { CK( cudaEventRecord( c->evs, st2 ) ); CK( cudaStreamWaitEvent( st, c->evs, 0 ) ); forked = false; }
    cur_depth = l.depth; cur_dir = dir;
    if ( c->profile ) CK( cudaEventRecord( c->evp0, st ) );
    cudaStream_t sst = st;
    if ( ( l.kind == SK_FWD_SMALL || l.kind == SK_BWD_SMALL ) && st2 != st ) { CK( cudaEventRecord( c->evs2, st ) ); CK( cudaStreamWaitEvent( st2, c->evs2, 0 ) ); forked = true; sst = st2; }
    switch ( l.kind ) {
      case SK_FWD_ZERO_W: CK( cudaMemsetAsync( c->dW[a], 0, ( size_t )l.count * pitch * 8, st ) ); break;
      case SK_FWD_GATHER: fwd_gather_kernel<<<( unsigned )l.count, 256, 0, st>>>( c->gather_tasks.d + l.first, c->sn.d, c->child_list.d, c->relidx.d, c->dY, pitch, tqs, c->dW[a], c->dW[ac] ); break;
      case SK_BWD_GATHER: bwd_gather_kernel<<<( unsigned )l.count, 256, 0, st>>>( c->bgather_tasks.d + l.first, c->sn.d, c->rowidx.d, c->dY, pitch, tqs, c->dW[a] ); break;
      case SK_FWD_SMALL: {
        const int pch = tensor ? ( p + 7 ) / 8 : 1; const unsigned g = ( unsigned )( l.count * pch ); const int* sl = c->sv_small.d + l.first;
        if ( tensor ) sv_small_fwd_kernel<8><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, c->dL, c->dSinvF, c->dY, c->dZ, c->dW[a], p, pitch );
        else if ( pct == 1 ) sv_small_fwd_kernel<1><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, c->dL, c->dSinvF, c->dY, c->dZ, c->dW[a], p, pitch );
        else if ( pct == 2 ) sv_small_fwd_kernel<2><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, c->dL, c->dSinvF, c->dY, c->dZ, c->dW[a], p, pitch );
        else sv_small_fwd_kernel<4><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, c->dL, c->dSinvF, c->dY, c->dZ, c->dW[a], p, pitch );
        break; }
      case SK_BWD_SMALL: {
        const int pch = tensor ? ( p + 7 ) / 8 : 1; const unsigned g = ( unsigned )( l.count * pch ); const int* sl = c->sv_small.d + l.first;
        const double* Mb = c->kind == B200_KIND_CHOLESKY ? c->dL : c->dU;
        if ( tensor ) sv_small_bwd_kernel<8><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, Mb, c->dSinvB, c->rowidx.d, c->dY, c->dZ, p, pitch );
        else if ( pct == 1 ) sv_small_bwd_kernel<1><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, Mb, c->dSinvB, c->rowidx.d, c->dY, c->dZ, p, pitch );
        else if ( pct == 2 ) sv_small_bwd_kernel<2><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, Mb, c->dSinvB, c->rowidx.d, c->dY, c->dZ, p, pitch );
        else sv_small_bwd_kernel<4><<<g, SVS_THREADS, 0, sst>>>( sl, pch, c->sn.d, Mb, c->dSinvB, c->rowidx.d, c->dY, c->dZ, p, pitch );
        break; }
      case SK_FWD_TRI: case SK_FWD_UPD: case SK_BWD_TRI:
        if ( tensor ) gemm_solve_dmma_kernel<false><<<dim3( ( unsigned )l.ng, gy ), 288, SVG_SMEM_NT, st>>>( c->dSvGemm, c->sv_tiles_g.d + l.g0, c->dZeros );
        else if ( pct == 1 ) sv_gemv_kernel<1><<<( unsigned )l.nt, 256, sv_gemv_smem<1>(), st>>>( c->sv_tasks.d, c->sv_tiles_v.d + l.t0, B, p, pitch );
        else if ( pct == 2 ) sv_gemv_kernel<2><<<( unsigned )l.nt, 256, sv_gemv_smem<2>(), st>>>( c->sv_tasks.d, c->sv_tiles_v.d + l.t0, B, p, pitch );
        else sv_gemv_kernel<4><<<( unsigned )l.nt, 256, sv_gemv_smem<4>(), st>>>( c->sv_tasks.d, c->sv_tiles_v.d + l.t0, B, p, pitch );
        break;
      case SK_BWD_UPD:
        if ( tensor ) gemm_solve_dmma_kernel<true><<<dim3( ( unsigned )l.ng, gy ), 288, SVG_SMEM_BT, st>>>( c->dSvGemm, c->sv_tiles_g.d + l.g0, c->dZeros );
        else if ( pct == 1 ) sv_gemvT_kernel<1><<<( unsigned )l.nt, 128, sv_gemvT_smem<1>(), st>>>( c->sv_tasks.d, c->sv_tiles_t.d + l.t0, B, p, pitch );
        else if ( pct == 2 ) sv_gemvT_kernel<2><<<( unsigned )l.nt, 128, sv_gemvT_smem<2>(), st>>>( c->sv_tasks.d, c->sv_tiles_t.d + l.t0, B, p, pitch );
        else sv_gemvT_kernel<4><<<( unsigned )l.nt, 128, sv_gemvT_smem<4>(), st>>>( c->sv_tasks.d, c->sv_tiles_t.d + l.t0, B, p, pitch );
        break;
    }
    nl++;
    if ( c->profile ) {
      CK( cudaEventRecord( c->evp1, st ) ); CK( cudaEventSynchronize( c->evp1 ) );
      float ms = 0; cudaEventElapsedTime( &ms, c->evp0, c->evp1 ); c->sprof_ms[l.kind] += ms; c->sprof_n[l.kind]++;
      static const bool trace_all = getenv( "B200_TRACE_ALL" ) != nullptr;
      if ( trace_all ) fprintf( stderr, "slaunch kind=%d depth=%d tiles=%lld gtiles=%lld ms=%.4f\n", l.kind, l.depth, ( long long )l.nt, ( long long )l.ng, ms );
    }
  }
  if ( forked ) { CK( cudaEventRecord( c->evs, st2 ) ); CK( cudaStreamWaitEvent( st, c->evs, 0 ) ); }
  permute_out_kernel<<<pblocks, 256, 0, st>>>( c->dY, c->perm.d, n, p, pitch, d_x );
  return B200_OK;
}

static int run_solve( b200_ctx* c, const double* d_b, int p, double* d_x ) {
  cudaStream_t st = c->stream;
  const int n = c->n;
  int r = ensure_solve_ws( c, p ); if ( r ) return r;
  const int pblocks = ( int )std::min<int64_t>( ( ( int64_t )n * p + 255 ) / 256, 148 * 32 );
  int64_t nl = 0;
  for ( int k = 0; k < SK_COUNT; k++ ) { c->sprof_ms[k] = 0; c->sprof_n[k] = 0; }
  CK( cudaEventRecord( c->ev0, st ) );
  if ( c->dSwCnt ) CK( cudaMemsetAsync( c->dSwCnt + c->sw_head + 2, 0, sizeof( int ), st ) );
  r = sweep_once( c, d_b, p, d_x, nl ); if ( r ) return r;
  /* iterative refinement with a double-double residual.  Automatic mode: the componentwise backward error
   * omega = max_i |r_i| / (|A||x|+|b|)_i is formed after every pair of sweeps; one correction step always when the
   * system is small (the extra sweeps cost microseconds and buy the last digits the reference's known-answer
   * tolerances ask for), otherwise only while omega exceeds 1e-14 -- two orders below the stated 1e-12 bound, so a
   * well-conditioned SPD system (the headline) runs exactly one pair of sweeps.  omega is NaN-propagating: a
   * non-finite x ends the solve with B200_ERR_SINGULAR instead of being reported as converged. */
  const int maxit = c->refine >= 0 ? c->refine : 2;
  c->stats.refine_steps = 0; c->stats.backward_error = -1.0; c->stats.first_backward_error = -1.0;
  for ( int it = 0; it <= maxit; it++ ) {
    if ( c->refine >= 0 && it == maxit ) break;       /* fixed number of steps: no final residual */
    CK( cudaMemsetAsync( c->dOmega, 0, sizeof( unsigned long long ), st ) );
    residual_dd_kernel<<<pblocks, 256, 0, st>>>( c->spmv_ptr.d, c->spmv_col.d, c->spmv_src.d, c->dVals, d_x, d_b, n, p, c->dR, c->dOmega );
    nl++;
    if ( c->refine < 0 ) {
      unsigned long long bits = 0;
      CK( cudaMemcpyAsync( &bits, c->dOmega, sizeof( bits ), cudaMemcpyDeviceToHost, st ) );
      CK( cudaStreamSynchronize( st ) );
      double omega; memcpy( &omega, &bits, sizeof( omega ) );
      c->stats.backward_error = omega; if ( it == 0 ) c->stats.first_backward_error = omega;
      if ( !std::isfinite( omega ) ) { set_err( "the solution is not finite (singular or badly pivoted factor)%s", "" ); return B200_ERR_SINGULAR; }
      const bool small = n <= 65536;     /* depends on n only: the answer must not depend on how columns are batched */
      if ( it == maxit || !( ( small && it == 0 ) || omega > 1e-14 ) ) break;
    }
    r = sweep_once( c, c->dR, p, c->dD, nl ); if ( r ) return r;
    axpy_kernel<<<pblocks, 256, 0, st>>>( d_x, c->dD, ( int64_t )n * p );
    nl++;
    c->stats.refine_steps = it + 1;
  }
  CK( cudaEventRecord( c->ev1, st ) );
  CK( cudaGetLastError() );
  CK( cudaStreamSynchronize( st ) );
  if ( c->sw_err_h != 0 ) { char a_[32]; snprintf( a_, 32, "%d", c->sw_err_h - 1 ); c->sw_err_h = 0; set_err( "persistent sweep: unit %s gave up waiting for its inputs (schedule error)%s", a_, "" ); return B200_ERR_CUDA; }
  float ms = 0; cudaEventElapsedTime( &ms, c->ev0, c->ev1 );
  c->stats.solve_ms = ms; c->stats.solve_launches = nl;
  return B200_OK;
}

static bool holds_factor( const b200_ctx* c ) { return c->world == 1 || c->rank == 0 || c->replicate; }
extern "C" int b200_solve_device( b200_ctx_t* c, const double* d_b, int p, double* d_x ) {
  if ( !c || !c->factored ) { set_err( "solve before a successful factor%s", "" ); return B200_ERR_STATE; }
  if ( !holds_factor( c ) ) { set_err( "multi-GPU: the finished factor is collected on rank 0, which runs the sweeps (B200_REPLICATE_FACTOR=1 keeps a copy on every rank)%s", "" ); return B200_ERR_STATE; }
  if ( p <= 0 ) return B200_ERR_ARG;
  CK( cudaSetDevice( c->device ) );
  return run_solve( c, d_b, p, d_x );
}

extern "C" int b200_solve_host( b200_ctx_t* c, const double* b, int p, double* x ) {
  if ( !c || !c->factored ) { set_err( "solve before a successful factor%s", "" ); return B200_ERR_STATE; }
  if ( !holds_factor( c ) ) { set_err( "multi-GPU: the finished factor is collected on rank 0, which runs the sweeps (B200_REPLICATE_FACTOR=1 keeps a copy on every rank)%s", "" ); return B200_ERR_STATE; }
  if ( p <= 0 || !b || !x ) return B200_ERR_ARG;
  CK( cudaSetDevice( c->device ) );
  if ( p > c->io_p_cap ) {
    cudaFree( c->dB ); cudaFree( c->dX ); c->dB = c->dX = nullptr; c->io_p_cap = 0;
    CK( cudaMalloc( &c->dB, ( size_t )c->n * p * 8 ) ); CK( cudaMalloc( &c->dX, ( size_t )c->n * p * 8 ) );
    c->io_p_cap = p;
  }
  const size_t bytes = ( size_t )c->n * p * 8;
  CK( cudaEventRecord( c->evp0, c->stream ) );
  CK( cudaMemcpyAsync( c->dB, b, bytes, cudaMemcpyHostToDevice, c->stream ) );
  CK( cudaEventRecord( c->evp1, c->stream ) );
  int r = run_solve( c, c->dB, p, c->dX );
  if ( r ) return r;
  float ms = 0; cudaEventElapsedTime( &ms, c->evp0, c->evp1 ); c->stats.h2d_ms = ms;
  CK( cudaEventRecord( c->evp0, c->stream ) );
  CK( cudaMemcpyAsync( x, c->dX, bytes, cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaEventRecord( c->evp1, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  cudaEventElapsedTime( &ms, c->evp0, c->evp1 ); c->stats.d2h_ms = ms;
  return B200_OK;
}

/* ------------------------------------------------------------ helpers */
extern "C" void* b200_dev_alloc( b200_ctx_t* c, size_t bytes ) { if ( !c ) return nullptr; cudaSetDevice( c->device ); void* p = nullptr; if ( cudaMalloc( &p, bytes ? bytes : 8 ) != cudaSuccess ) { set_err( "cudaMalloc failed%s", "" ); return nullptr; } return p; }
extern "C" void b200_dev_free( b200_ctx_t* c, void* p ) { if ( c && p ) { cudaSetDevice( c->device ); cudaFree( p ); } }
extern "C" int b200_dev_upload( b200_ctx_t* c, void* dst, const void* src, size_t bytes ) { if ( !c ) return B200_ERR_ARG; CK( cudaSetDevice( c->device ) ); CK( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyHostToDevice, c->stream ) ); CK( cudaStreamSynchronize( c->stream ) ); return B200_OK; }
extern "C" int b200_dev_download( b200_ctx_t* c, void* dst, const void* src, size_t bytes ) { if ( !c ) return B200_ERR_ARG; CK( cudaSetDevice( c->device ) ); CK( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToHost, c->stream ) ); CK( cudaStreamSynchronize( c->stream ) ); return B200_OK; }
extern "C" void* b200_host_alloc_pinned( size_t bytes ) { void* p = nullptr; if ( cudaHostAlloc( &p, bytes ? bytes : 8, cudaHostAllocDefault ) != cudaSuccess ) return nullptr; return p; }
extern "C" void b200_host_free_pinned( void* p ) { if ( p ) cudaFreeHost( p ); }
extern "C" int b200_dev_sync( b200_ctx_t* c ) { if ( !c ) return B200_ERR_ARG; CK( cudaSetDevice( c->device ) ); CK( cudaStreamSynchronize( c->stream ) ); return B200_OK; }
extern "C" int b200_flush_l2( b200_ctx_t* c ) {
  if ( !c ) return B200_ERR_ARG;
  CK( cudaSetDevice( c->device ) );
  if ( !c->flush_buf ) { c->flush_bytes = ( size_t )256 << 20; CK( cudaMalloc( &c->flush_buf, c->flush_bytes ) ); }
  CK( cudaMemsetAsync( c->flush_buf, 1, c->flush_bytes, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  return B200_OK;
}
extern "C" int b200_get_stats( b200_ctx_t* c, b200_stats_t* out ) { if ( !c || !out ) return B200_ERR_ARG; *out = c->stats; return B200_OK; }
extern "C" int b200_set_refinement( b200_ctx_t* c, int steps ) { if ( !c ) return B200_ERR_ARG; c->refine = steps < 0 ? -1 : steps; return B200_OK; }
extern "C" int b200_set_profile( b200_ctx_t* c, int on ) { if ( !c ) return B200_ERR_ARG; c->profile = on != 0; return B200_OK; }
extern "C" int b200_get_profile( b200_ctx_t* c, char* buf, size_t len ) {
  if ( !c || !buf || len == 0 ) return B200_ERR_ARG;
  std::string s;
  char line[256];
  for ( int k = 0; k < LK_COUNT; k++ ) { snprintf( line, sizeof( line ), "%s launches=%lld ms=%.3f\n", kind_name[k], ( long long )c->prof_n[k], c->prof_ms[k] ); s += line; }
  static const char* sname[SK_COUNT] = { "solve_zero_w", "solve_fwd_gather", "solve_fwd_tri", "solve_fwd_upd", "solve_bwd_gather", "solve_bwd_upd", "solve_bwd_tri", "solve_fwd_small", "solve_bwd_small", "solve_sweep_fwd", "solve_sweep_bwd" };
  for ( int k = 0; k < SK_COUNT; k++ ) { snprintf( line, sizeof( line ), "%s launches=%lld ms=%.3f\n", sname[k], ( long long )c->sprof_n[k], c->sprof_ms[k] ); s += line; }
  { float t0 = 0; if ( !c->lvl_ev.empty() ) cudaEventElapsedTime( &t0, c->ev0, c->lvl_ev[0] ); snprintf( line, sizeof( line ), "levels_ms(deepest first; prologue %.2f):", t0 ); s += line;
    for ( float v : c->lvl_ms ) { snprintf( line, sizeof( line ), " %.2f", v ); s += line; } s += "\n"; }
  snprintf( buf, len, "%s", s.c_str() );
  return B200_OK;
}

/* diagnostics: copy a range of a device array to the host. which: 0 = L panels, 1 = U' panels, 2 = inverse blocks */
extern "C" int b200_debug_read( b200_ctx_t* c, int which, double* dst, int64_t offset, int64_t count ) {
  if ( !c || !dst ) return B200_ERR_ARG;
  CK( cudaSetDevice( c->device ) );
  const double* src = which == 0 ? c->dL : which == 1 ? c->dU : which == 2 ? c->dInv : which == 3 ? c->dSinvF : c->dSinvB;
  if ( !src ) return B200_ERR_STATE;
  CK( cudaMemcpy( dst, src + offset, ( size_t )count * 8, cudaMemcpyDeviceToHost ) );
  return B200_OK;
}

/* checksum of the factor held by this context: colsq[j] = sum of squares of column j of L (permuted numbering), n doubles on the host.
 * Deterministic (fixed reduction order), so equal factors give equal bits on any number of GPUs. */
extern "C" int b200_factor_colnorms( b200_ctx_t* c, double* colsq ) {
  if ( !c || !colsq || !c->factored || c->dry ) return B200_ERR_STATE;
  if ( !holds_factor( c ) ) { for ( int i = 0; i < c->n; i++ ) colsq[i] = 0.0; return B200_OK; }      /* only a rank with the whole factor reports it */
  CK( cudaSetDevice( c->device ) );
  double* d = nullptr; CK( cudaMalloc( &d, ( size_t )c->n * 8 ) );
  CK( cudaMemsetAsync( d, 0, ( size_t )c->n * 8, c->stream ) );
  panel_colnorm_kernel<<<148 * 8, 256, 0, c->stream>>>( c->sn.d, c->ns, c->dL, d );
  CK( cudaMemcpyAsync( colsq, d, ( size_t )c->n * 8, cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  cudaFree( d );
  return B200_OK;
}

/* ---- micro-benchmarks: the FP64 roof of OUR kernel and the HBM copy roof */
extern "C" double b200_bench_dgemm_tflops( b200_ctx_t* c, int m, int n, int k, int iters ) {
  if ( !c || m <= 0 || n <= 0 || k <= 0 || ( m & 1 ) || ( n & 1 ) ) return -1.0;      /* even leading dimensions: 16-byte aligned bulk copies */
  cudaSetDevice( c->device );
  double *A, *B, *C;
  if ( cudaMalloc( &A, ( size_t )m * k * 8 ) || cudaMalloc( &B, ( size_t )n * k * 8 ) || cudaMalloc( &C, ( size_t )m * n * 8 ) ) return -1.0;
  cudaMemset( A, 0, ( size_t )m * k * 8 ); cudaMemset( B, 0, ( size_t )n * k * 8 ); cudaMemset( C, 0, ( size_t )m * n * 8 );
  std::vector<GemmTask> gt; std::vector<Tile> tb, ts; double fl = 0;
  add_gemm( gt, tb, ts, fl, C, m, A, m, B, n, m, n, k, 0, 0 );
  DevVec<GemmTask> dgt; DevVec<Tile> dtb, dts; dgt.upload( gt ); dtb.upload( tb ); dts.upload( ts );
  float best = 1e30f;
  for ( int it = 0; it < iters + 2; it++ ) {
    cudaEventRecord( c->ev0, c->stream );
    if ( dtb.n ) gemm_nt_dmma_bulk_kernel<GEMM_CFG><<<( unsigned )dtb.n, BULK_THREADS, BULK_SMEM, c->stream>>>( dgt.d, dtb.d, c->dZeros );
    cudaEventRecord( c->ev1, c->stream ); cudaEventSynchronize( c->ev1 );
    float ms; cudaEventElapsedTime( &ms, c->ev0, c->ev1 ); if ( it >= 2 && ms < best ) best = ms;
  }
  dgt.release(); dtb.release(); dts.release(); cudaFree( A ); cudaFree( B ); cudaFree( C );
  if ( cudaGetLastError() != cudaSuccess ) return -1.0;
  return 2.0 * m * n * ( double )k / ( best * 1e-3 ) / 1e12;
}

/* DMMA throughput as a function of resident warps per SM and independent accumulators per warp */
extern "C" double b200_bench_dmma_occupancy( b200_ctx_t* c, int warps_per_sm, int ilp ) {
  if ( !c ) return -1.0;
  cudaSetDevice( c->device );
  double* out; if ( cudaMalloc( &out, 8 ) ) return -1.0;
  const int inner = 8192;
  const int threads = std::min( warps_per_sm, 32 ) * 32; const int blocks = 148 * std::max( 1, warps_per_sm / 32 );
  float best = 1e30f;
  for ( int it = 0; it < 4; it++ ) {
    cudaEventRecord( c->ev0, c->stream );
    switch ( ilp ) {
      case 1: dmma_occ_kernel<1><<<blocks, threads, 0, c->stream>>>( out, inner ); break;
      case 2: dmma_occ_kernel<2><<<blocks, threads, 0, c->stream>>>( out, inner ); break;
      case 4: dmma_occ_kernel<4><<<blocks, threads, 0, c->stream>>>( out, inner ); break;
      case 8: dmma_occ_kernel<8><<<blocks, threads, 0, c->stream>>>( out, inner ); break;
      default: dmma_occ_kernel<16><<<blocks, threads, 0, c->stream>>>( out, inner ); ilp = 16; break;
    }
    cudaEventRecord( c->ev1, c->stream ); cudaEventSynchronize( c->ev1 );
    float ms; cudaEventElapsedTime( &ms, c->ev0, c->ev1 ); if ( it >= 1 && ms < best ) best = ms;
  }
  cudaFree( out );
  return ( double )blocks * ( threads / 32 ) * inner * ilp * 512.0 / ( best * 1e-3 ) / 1e12;
}

/* which: 0 = DMMA.8x8x4 issue roof, 1 = DFMA roof; TFLOP/s, best of iters */
extern "C" double b200_bench_fp64_peak_tflops( b200_ctx_t* c, int which, int iters ) {
  if ( !c ) return -1.0;
  cudaSetDevice( c->device );
  double* out; if ( cudaMalloc( &out, 8 ) ) return -1.0;
  const int inner = 4096, blocks = 148 * 8;
  float best = 1e30f;
  for ( int it = 0; it < iters + 2; it++ ) {
    cudaEventRecord( c->ev0, c->stream );
    if ( which == 0 ) dmma_peak_kernel<<<blocks, 256, 0, c->stream>>>( out, inner ); else dfma_peak_kernel<<<blocks, 256, 0, c->stream>>>( out, inner );
    cudaEventRecord( c->ev1, c->stream ); cudaEventSynchronize( c->ev1 );
    float ms; cudaEventElapsedTime( &ms, c->ev0, c->ev1 ); if ( it >= 2 && ms < best ) best = ms;
  }
  cudaFree( out );
  const double flops = which == 0 ? ( double )blocks * 8 * inner * 16 * 512.0 : ( double )blocks * 256 * inner * 16 * 2.0;
  return flops / ( best * 1e-3 ) / 1e12;
}

extern "C" double b200_bench_stream_gbs( b200_ctx_t* c, size_t bytes, int iters ) {
  if ( !c ) return -1.0;
  cudaSetDevice( c->device );
  double2 *a, *b; bytes &= ~( size_t )15;
  if ( cudaMalloc( &a, bytes ) || cudaMalloc( &b, bytes ) ) return -1.0;
  cudaMemset( a, 0, bytes );
  float best = 1e30f;
  for ( int it = 0; it < iters + 2; it++ ) {
    cudaEventRecord( c->ev0, c->stream );
    stream_copy_kernel<<<148 * 16, 512, 0, c->stream>>>( a, b, ( int64_t )( bytes / 16 ) );
    cudaEventRecord( c->ev1, c->stream ); cudaEventSynchronize( c->ev1 );
    float ms; cudaEventElapsedTime( &ms, c->ev0, c->ev1 ); if ( it >= 2 && ms < best ) best = ms;
  }
  cudaFree( a ); cudaFree( b );
  return 2.0 * bytes / ( best * 1e-3 ) / 1e9;
}

/* micro-benchmark of the diagonal-block kernels: which = 0 unrolled register kernel, 1 rolled shared-memory kernel;
 * ntasks blocks of order nb per launch; returns microseconds per launch (best of iters) */
extern "C" double b200_bench_potrf_us( b200_ctx_t* c, int which, int nb, int ntasks, int iters ) {
  if ( !c || nb < 1 || nb > NB || ntasks < 1 ) return -1.0;
  cudaSetDevice( c->device );
  const size_t blk = ( size_t )NB * NB;
  std::vector<double> h( blk * ntasks, 0.0 );
  for ( int t = 0; t < ntasks; t++ ) for ( int cc = 0; cc < nb; cc++ ) for ( int r = cc; r < nb; r++ ) h[t * blk + r + ( size_t )cc * NB] = r == cc ? 12.0 + 0.01 * ( t % 7 ) : -1.0 / ( 1 + r - cc );
  double *A0, *A, *inv; int* err; DiagTask* dtk;
  if ( cudaMalloc( &A0, h.size() * 8 ) || cudaMalloc( &A, h.size() * 8 ) || cudaMalloc( &inv, h.size() * 8 ) || cudaMalloc( &err, 4 ) || cudaMalloc( &dtk, sizeof( DiagTask ) * ntasks ) ) return -1.0;
  cudaMemcpy( A0, h.data(), h.size() * 8, cudaMemcpyHostToDevice ); cudaMemset( err, 0, 4 );
  std::vector<DiagTask> tk( ntasks );
  for ( int t = 0; t < ntasks; t++ ) { tk[t].A = A + t * blk; tk[t].AU = nullptr; tk[t].lda = NB; tk[t].nb = nb; tk[t].inv = inv + t * blk; tk[t].inv2 = nullptr; tk[t].piv = nullptr; tk[t].inv3 = nullptr; tk[t].col = 0; }
  cudaMemcpy( dtk, tk.data(), sizeof( DiagTask ) * ntasks, cudaMemcpyHostToDevice );
  float best = 1e30f;
  for ( int it = 0; it < iters + 2; it++ ) {
    cudaMemcpyAsync( A, A0, h.size() * 8, cudaMemcpyDeviceToDevice, c->stream );
    cudaEventRecord( c->ev0, c->stream );
    if ( which == 0 ) potrf_diag_kernel<<<ntasks, POTRF_THREADS, 0, c->stream>>>( dtk, err );
    else potrf_diag_small_kernel<<<ntasks, 256, POTRF_S_SMEM, c->stream>>>( dtk, err );
    cudaEventRecord( c->ev1, c->stream ); cudaEventSynchronize( c->ev1 );
    float ms; cudaEventElapsedTime( &ms, c->ev0, c->ev1 ); if ( it >= 2 && ms < best ) best = ms;
  }
  int e = 0; cudaMemcpy( &e, err, 4, cudaMemcpyDeviceToHost );
  cudaFree( A0 ); cudaFree( A ); cudaFree( inv ); cudaFree( err ); cudaFree( dtk );
  if ( cudaGetLastError() != cudaSuccess || e != 0 ) return -2.0;
  return best * 1e3;
}

/* tuning aid: the bulk-copy DMMA kernel with other tile configurations (C -= A B', plain rectangular task) */
template <int BM, int BN, int WM, int WN, int ST>
static double bench_bulk_variant( b200_ctx* c, int m, int n, int k, int iters ) {
  constexpr int SMEM = ST * 16 * ( BM + 4 + BN + 4 ) * 8 + 2 * ST * 8;
  constexpr int NT = ( BM / WM ) * ( BN / WN ) * 32 + 32;
  if ( cudaFuncSetAttribute( gemm_nt_dmma_bulk_kernel<BM, BN, WM, WN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM ) != cudaSuccess ) { cudaGetLastError(); return -1.0; }
  double *A, *B, *C;
  if ( cudaMalloc( &A, ( size_t )m * k * 8 ) || cudaMalloc( &B, ( size_t )n * k * 8 ) || cudaMalloc( &C, ( size_t )m * n * 8 ) ) return -1.0;
  cudaMemset( A, 0, ( size_t )m * k * 8 ); cudaMemset( B, 0, ( size_t )n * k * 8 ); cudaMemset( C, 0, ( size_t )m * n * 8 );
  GemmTask t; t.C = C; t.A = A; t.B = B; t.ldc = m; t.lda = m; t.ldb = n; t.M = m; t.N = n; t.K = k; t.lower = 0; t.diag = 0; t.mode = 0;
  std::vector<GemmTask> gt( 1, t ); std::vector<Tile> tl;
  for ( int tn = 0; tn * BN < n; tn++ ) for ( int tm = 0; tm * BM < m; tm++ ) { Tile x; x.task = 0; x.tm = ( unsigned short )tm; x.tn = ( unsigned short )tn; tl.push_back( x ); }
  DevVec<GemmTask> dgt; DevVec<Tile> dtl; dgt.upload( gt ); dtl.upload( tl );
  float best = 1e30f;
  for ( int it = 0; it < iters + 2; it++ ) {
    cudaEventRecord( c->ev0, c->stream );
    gemm_nt_dmma_bulk_kernel<BM, BN, WM, WN, ST><<<( unsigned )dtl.n, NT, SMEM, c->stream>>>( dgt.d, dtl.d, c->dZeros );
    cudaEventRecord( c->ev1, c->stream ); cudaEventSynchronize( c->ev1 );
    float ms; cudaEventElapsedTime( &ms, c->ev0, c->ev1 ); if ( it >= 2 && ms < best ) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  dgt.release(); dtl.release(); cudaFree( A ); cudaFree( B ); cudaFree( C );
  if ( e != cudaSuccess ) { set_err( "bulk variant: %s", cudaGetErrorString( e ) ); return -1.0; }
  return 2.0 * m * n * ( double )k / ( best * 1e-3 ) / 1e12;
}
extern "C" double b200_bench_dgemm_bulk_variant( b200_ctx_t* c, int variant, int m, int n, int k, int iters ) {
  if ( !c ) return -1.0;
  cudaSetDevice( c->device );
  switch ( variant ) {
    case 0: return bench_bulk_variant<64, 64, 32, 16, 4>( c, m, n, k, iters );      /* production */
    case 1: return bench_bulk_variant<128, 64, 32, 32, 4>( c, m, n, k, iters );
    case 2: return bench_bulk_variant<64, 128, 32, 32, 4>( c, m, n, k, iters );
    case 3: return bench_bulk_variant<128, 64, 64, 16, 4>( c, m, n, k, iters );
    case 4: return bench_bulk_variant<64, 64, 32, 16, 6>( c, m, n, k, iters );
    case 5: return bench_bulk_variant<64, 64, 16, 32, 4>( c, m, n, k, iters );
    case 6: return bench_bulk_variant<128, 64, 32, 32, 3>( c, m, n, k, iters );
    case 7: return bench_bulk_variant<128, 128, 64, 32, 3>( c, m, n, k, iters );
    case 8: return bench_bulk_variant<64, 64, 32, 32, 4>( c, m, n, k, iters );
    default: return -1.0;
  }
}

/* ---- export of the static factor schedule (diagnostics; see b200_ctx_create_dry) */
extern "C" int64_t b200_debug_schedule( b200_ctx_t* c, b200_sched_entry_t* out, int64_t cap ) {
  if ( !c ) return -1;
  const int64_t n = ( int64_t )c->launches.size();
  for ( int64_t i = 0; i < n && i < cap; i++ ) {
    const Launch& l = c->launches[i];
    const bool comm = l.kind == LK_XFER_CB || l.kind == LK_BCAST || l.kind == LK_GATHER;
    out[i].kind = l.kind; out[i].stream = l.stream; out[i].wait_ev = l.wait_ev; out[i].record_ev = l.record_ev; out[i].depth = l.depth;
    out[i].first_msg = comm ? l.first : -1; out[i].nmsg = comm ? l.count : 0; out[i].count = l.count;
    out[i].flops = 0; out[i].kavg = 0; out[i].bytes = 0;
    if ( c->dry && ( l.kind == LK_GEMM_BIG || l.kind == LK_TRSM_MMA ) ) {
      double fl = 0, fk = 0;
      for ( int64_t t = l.first; t < l.first + l.count; t++ ) { const Tile& tl = c->h_tb[t]; const GemmTask& g = c->h_gt[tl.task];
        const double mm = std::min( GEMM_BM, g.M - tl.tm * GEMM_BM ), nn = std::min( GEMM_BN, g.N - tl.tn * GEMM_BN ); const double f = 2.0 * mm * nn * g.K; fl += f; fk += f * g.K; }
      out[i].flops = fl; out[i].kavg = fl > 0 ? fk / fl : 0;
    } else if ( c->dry && l.kind == LK_MEMSET_CB ) out[i].bytes = 8.0 * l.count;
    else if ( c->dry && l.kind == LK_ASM ) {
      double b = 0;
      for ( int64_t t = l.first; t < l.first + l.count; t++ ) { const AsmTask& a = c->h_at[t]; const SnMeta& P = c->h_meta[a.parent];
        for ( int ci = P.child_begin; ci < P.child_end; ci++ ) { const SnMeta& C = c->h_meta[c->h_child[ci]]; const double uc = C.f - C.k;
          b += 12.0 * uc * uc * ( double )a.ncols / std::max( 1, P.f ); } }     /* ~ read child + read/write parent over the lower triangle: 24 B x u^2/2, shared by the parent's column tasks */
      out[i].bytes = b;
    }
  }
  return n;
}
extern "C" int64_t b200_debug_messages( b200_ctx_t* c, b200_msg_t* out, int64_t cap ) {
  if ( !c ) return -1;
  const int64_t n = ( int64_t )c->msgs.size();
  for ( int64_t i = 0; i < n && i < cap; i++ ) { const Msg& m = c->msgs[i]; out[i].src = m.src; out[i].dst = m.dst; out[i].which = m.which; out[i].offset = m.off; out[i].count = m.cnt; }
  return n;
}
/* ---- export of the persistent-sweep unit lists (schedule-only contexts; tests/sweep_sim.py replays them on the CPU).
 * dir: 0 forward, 1 backward; every unit is 16 ints (SvUnit).  Tasks: 8 int64 each = vsel, voff, K, dsel, doff, N, tri, mode. */
extern "C" int64_t b200_debug_sweep_units( b200_ctx_t* c, int dir, int* out, int64_t cap ) {
  if ( !c || !c->dry ) return -1;
  const std::vector<SvUnit>& u = dir == 0 ? c->h_sw_f : c->h_sw_b;
  const int64_t n = ( int64_t )u.size();
  if ( out ) for ( int64_t i = 0; i < n && i < cap; i++ ) memcpy( out + 16 * i, &u[i], 64 );
  return n;
}
extern "C" int64_t b200_debug_sweep_tasks( b200_ctx_t* c, int64_t* out, int64_t cap ) {
  if ( !c || !c->dry ) return -1;
  const int64_t n = ( int64_t )c->h_sw_t.size();
  if ( out ) for ( int64_t i = 0; i < n && i < cap; i++ ) { const SvTask& t = c->h_sw_t[i]; int64_t* o = out + 8 * i;
    o[0] = t.vsel; o[1] = t.voff; o[2] = t.K; o[3] = t.dsel; o[4] = t.doff; o[5] = t.N; o[6] = t.tri; o[7] = t.mode; }
  return n;
}
extern "C" int64_t b200_debug_sweep_counters( b200_ctx_t* c ) { return c ? c->sw_ncnt : -1; }
/* diagnostics: per-unit time stamps of the last persistent sweep (context created with B200_SWEEP_TRACE set) and the device unit lists.
 * trace: 5 uint64 per unit (globaltimer ns at start, clock64 at start / inputs ready / done, SM id); units: 16 ints each */
extern "C" int64_t b200_debug_sweep_trace( b200_ctx_t* c, int dir, unsigned long long* trace, int* units, int64_t cap ) {
  if ( !c || c->dry || !c->dSwTrace || dir < 0 || dir > 1 ) return -1;
  cudaSetDevice( c->device );
  const int64_t n = ( int64_t )c->sw_units[dir].n; const int64_t m = std::min( n, cap );
  if ( trace && m > 0 ) cudaMemcpy( trace, c->dSwTrace + ( dir == 0 ? 0 : 5 * c->sw_units[0].n ), ( size_t )m * 5 * 8, cudaMemcpyDeviceToHost );
  if ( units && m > 0 ) cudaMemcpy( units, c->sw_units[dir].d, ( size_t )m * 64, cudaMemcpyDeviceToHost );
  return n;
}
extern "C" const char* b200_debug_kind_name( int kind ) { return kind >= 0 && kind < LK_COUNT ? kind_name[kind] : "?"; }

/* every memory range a compute launch of the factor schedule may read or write, as conservative intervals of doubles inside
 * one of the four buffers (0 L, 1 inverse blocks, 2/3 contribution-block arena 0/1); schedule-only contexts only.
 * tests/sched_sim.py uses it to prove that every message is ordered against the launches that touch the same data. */
extern "C" int64_t b200_debug_accesses( b200_ctx_t* c, b200_access_t* out, int64_t cap ) {
  if ( !c || !c->dry ) return -1;
  int64_t n = 0;
  auto emit = [&]( int launch, const void* p, int64_t cnt, int write ) {
    if ( !p || cnt <= 0 ) return;
    const uintptr_t a = ( uintptr_t )p; const int tag = ( int )( a >> 40 );
    const int which = tag == 1 ? 0 : tag == 4 ? 1 : tag == 2 ? 2 : tag == 3 ? 3 : -1;
    if ( which < 0 ) return;                              /* other buffers (U', pivots, solve inverses) never travel */
    if ( n < cap ) { out[n].launch = launch; out[n].which = which; out[n].write = write; out[n].offset = ( int64_t )( ( a - ( ( uintptr_t )tag << 40 ) ) / 8 ); out[n].count = cnt; }
    n++;
  };
  for ( size_t i = 0; i < c->launches.size(); i++ ) {
    const Launch& l = c->launches[i]; const int li = ( int )i;
    switch ( l.kind ) {
      case LK_MEMSET_CB: emit( li, c->dCB[l.depth & 1] + l.first, l.count, 1 ); break;
      case LK_ASM:
        for ( int64_t t = l.first; t < l.first + l.count; t++ ) {
          const AsmTask& a = c->h_at[t]; const SnMeta& P = c->h_meta[a.parent];
          const int u = P.f - P.k; const int64_t ldu = ( u + 1 ) & ~1;
          const int p_hi = std::min( a.col0 + a.ncols, P.k );
          if ( a.col0 < p_hi ) emit( li, c->dL + P.lptr + ( int64_t )a.col0 * P.ld, ( int64_t )( p_hi - a.col0 ) * P.ld, 1 );
          const int j_lo = std::max( a.col0, P.k ) - P.k, j_hi = a.col0 + a.ncols - P.k;
          if ( j_lo < j_hi ) emit( li, c->dCB[l.depth & 1] + P.cbptr + ( int64_t )j_lo * ldu, ( int64_t )( j_hi - j_lo ) * ldu, 1 );
          for ( int ci = P.child_begin; ci < P.child_end; ci++ ) { const SnMeta& C = c->h_meta[c->h_child[ci]]; const int uc = C.f - C.k; if ( uc > 0 ) emit( li, c->dCB[( l.depth + 1 ) & 1] + C.cbptr, ( int64_t )uc * ( ( uc + 1 ) & ~1 ), 0 ); }
        }
        break;
      case LK_DIAG: case LK_DIAG_S:
        for ( int64_t t = l.first; t < l.first + l.count; t++ ) { const DiagTask& d = c->h_dt[t]; emit( li, d.A, ( int64_t )( d.nb - 1 ) * d.lda + d.nb, 1 ); emit( li, d.inv, ( int64_t )d.nb * d.nb, 1 ); }
        break;
      case LK_TRSM:
        for ( int64_t t = l.first; t < l.first + l.count; t++ ) { const TrsmTask& x = c->h_tt[c->h_ttl[t].task]; emit( li, x.B, ( int64_t )( x.nb - 1 ) * x.ldb + x.m, 1 ); emit( li, x.T, ( int64_t )x.nb * x.nb, 0 ); }
        break;
      case LK_GEMM_BIG: case LK_TRSM_MMA: {
        int last = -1;
        for ( int64_t t = l.first; t < l.first + l.count; t++ ) {
          const int id = c->h_tb[t].task; if ( id == last ) continue; last = id;
          const GemmTask& g = c->h_gt[id];
          emit( li, g.C, ( int64_t )( g.N - 1 ) * g.ldc + g.M, 1 );
          emit( li, g.A, ( int64_t )( g.K - 1 ) * g.lda + g.M, 0 );
          emit( li, g.B, ( int64_t )( g.K - 1 ) * g.ldb + g.N, 0 );
        }
        break; }
      case LK_SMALL_FRONT:
        for ( int64_t t = l.first; t < l.first + l.count; t++ ) {
          const SnMeta& S = c->h_meta[c->h_small[t]]; const int u = S.f - S.k;
          emit( li, c->dL + S.lptr, ( int64_t )S.k * S.ld, 1 ); emit( li, c->dInv + S.invptr, ( int64_t )S.k * S.k, 1 );
          if ( u > 0 ) emit( li, c->dCB[l.depth & 1] + S.cbptr, ( int64_t )u * ( ( u + 1 ) & ~1 ), 1 );
        }
        break;
      default: break;
    }
  }
  return n;
}
