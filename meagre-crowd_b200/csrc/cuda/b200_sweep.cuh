/* b200_sweep.cuh -- the triangular sweeps for 1..4 right-hand sides as ONE persistent kernel per sweep.
 *
 * evaluate (src/solvers.c:500-611 -> table.evaluate; cholmod_solve at src/solver_cholmod.c:303) is HBM bound: every entry
 * of the factor is read once per sweep.  As a sequence of level-set launches (one per tree level and solve block, ~420 of
 * them on the 128^3 factor) it ran at a quarter of the HBM roof: most launches are a handful of tiles, each a serial walk
 * over its slab, and nothing of the next step can start before the slowest tile of the current one has ended.
 *
 * Here the SAME tiles (same tasks, same canonical arithmetic: sv_gemv / sv_gemvT bodies of b200_kernels.cuh) are "units" of one
 * ordered work list.  CTAs take units from the head of the list (atomic ticket) and wait on the device for exactly the
 * data the unit reads:
 *     counters per 32 pivot / update rows of a supernode (forward) and per 64 pivot columns (backward) count the block
 *     updates applied so far; per supernode: triangular-product tiles done, gather units done, children (forward) or
 *     parent (backward) complete.
 * A unit only ever waits for units that precede it in the list, and those have been taken by running CTAs, so the
 * schedule cannot deadlock however many CTAs are resident; a spin bound turns a logic error into an error code instead
 * of a hang.  The triangular product of block b+1 starts as soon as the 32-row tiles of ITS rows have been updated, while
 * the other tiles of block b's update still stream; the fronts of the next tree level start while the tail of this one runs.
 *
 * Vectors written during the sweep (y, z, w) are read with ld.global.cg / written through to L2: L1 is not coherent
 * between SMs.  The factor itself is read-only (ld.global.nc).
 */
#pragma once
#include "b200_kernels.cuh"

namespace b200 {

enum { UK_FWD_GATHER = 0, UK_FWD_SMALL, UK_GEMV, UK_GEMVT, UK_BWD_SMALL, UK_BATCH };

struct SvUnit {          /* 64 bytes */
  int kind;
  int task;              /* UK_GEMV / UK_GEMVT: index into the task list; UK_*_SMALL on the device: first entry of the batch in the small-supernode table */
  int o0;                /* first output of the tile (UK_GEMV: 32 rows, UK_GEMVT: 64 columns); UK_FWD_GATHER: first destination row; small batch: its size */
  int aux;               /* UK_FWD_GATHER: end of the destination rows; UK_GEMVT: 1 = v is x gathered through the row list of sn   */
  int sn;                /* supernode                                                                                             */
  int w1, n1, v1;        /* wait until cnt[w1 .. w1+n1) >= v1  (n1 = 0: nothing to wait for)                                      */
  int w2, v2;            /* ... and cnt[w2] >= v2              (w2 < 0: nothing)                                                  */
  int sig;               /* cnt[sig] += 1 when the unit is done (sig < 0: none)                                                   */
  int up;                /* cnt[up] += 1 as well: the completion counter its consumers wait on -- the parent's "children done" (forward: it waits for
                          * the units of ALL its children), the supernode's own "done" (backward: its children wait for all its units); < 0: none */
  int pad0, pad1, pad2, pad3;
};
static_assert( sizeof( SvUnit ) == 64, "SvUnit is four 16-byte words" );
/* Units that are latency rather than bandwidth (small supernodes; gemv tiles with K <= SW_WARP_K) run one per WARP: the device list holds
 * UK_BATCH units { task = first entry of the batch in a second unit array, o0 = its size <= SW_BATCH }; every warp waits and signals
 * for its own inner unit. */
constexpr int SW_WARP_K = 256;
constexpr int SW_BATCH = 8;

#ifndef SW_MINB
#define SW_MINB 3          /* resident CTAs per SM the sweep kernel is compiled for (80 registers: no spills worth the name) */
#endif
constexpr int SW_THREADS = 256;
constexpr int SW_GT_COLS = 64;                        /* columns per gemvT unit: 8 warps x 8 columns */
constexpr unsigned SW_SPIN_LIMIT = 1u << 22;          /* polls before a unit gives up (a few seconds): a broken schedule must not hang the GPU */
template <int PCT> constexpr int sw_smem() {      /* doubles: the CTA-wide gemv tile, the gemvT tile, or eight warp-level units */
  constexpr int a = ( SB_WIDE + SVG_MAXCH * GV_ROWS ) * PCT, b = GT_XS * PCT + 8 * 8 * ( KC + 1 ), c = SW_BATCH * SW_WARP_K * PCT;
  return ( a > b ? ( a > c ? a : c ) : ( b > c ? b : c ) ) * 8;
}

__device__ __forceinline__ int ld_acquire( const int* p ) { int v; asm volatile( "ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"( v ) : "l"( p ) : "memory" ); return v; }

/* dest (-)= M v over 32 output rows (the body of sv_gemv_kernel; vector reads bypass L1) */
template <int PCT>
__device__ __forceinline__ void sw_gemv( const SvTask& t, int o0, const SvBases& B, int p, int pitch, double* dsm ) {
  double ( *vs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm );
  double ( *part )[GV_ROWS][PCT] = reinterpret_cast<double ( * )[GV_ROWS][PCT]>( dsm + SB_WIDE * PCT );
  const int pcn = min( PCT, p );
  const int tid = threadIdx.x, r = tid & ( GV_ROWS - 1 ), g = tid >> 5;
  const int row = o0 + r;
  const bool live = row < t.N;
  const double* v = B.b[t.vsel] + t.voff * pitch;
  double* dest = B.b[t.dsel] + ( t.doff + row ) * pitch;
  const int k_lo = t.tri == 2 ? ( o0 & ~( KC - 1 ) ) : 0, k_hi = t.tri == 1 ? min( t.K, ( ( o0 + GV_ROWS - 1 ) | ( KC - 1 ) ) + 1 ) : t.K;
  for ( int e = tid; e < ( k_hi - k_lo ) * PCT; e += SW_THREADS ) { const int kk = e / PCT, q = e % PCT; vs[kk][q] = q < pcn ? __ldcg( v + ( int64_t )( k_lo + kk ) * pitch + q ) : 0.0; }
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( t.mode == 0 && live && g == 0 && q < pcn ) ? __ldcg( dest + q ) : 0.0;
  __syncthreads();
  const int nch = ( k_hi - k_lo + KC - 1 ) / KC;
  const int64_t ldm = t.ld;
  for ( int ch = g; ch < nch; ch += 8 ) {
    const int k0 = k_lo + ch * KC, kn = min( KC, k_hi - k0 );
    double acc[PCT];
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
    if ( live ) {
      const double* Mc = t.M + row + ( int64_t )k0 * ldm;
      const double ( *vc )[PCT] = vs + ch * KC;
      if ( kn == KC ) {
#pragma unroll 1
        for ( int h = 0; h < KC; h += 16 ) {
          double mv[16];
#pragma unroll
          for ( int i = 0; i < 16; i++ ) mv[i] = __ldg( Mc + ( h + i ) * ldm );
          pin8( mv[0], mv[1], mv[2], mv[3], mv[4], mv[5], mv[6], mv[7] ); pin8( mv[8], mv[9], mv[10], mv[11], mv[12], mv[13], mv[14], mv[15] );
#pragma unroll
          for ( int i = 0; i < 16; i++ ) {
#pragma unroll
            for ( int q = 0; q < PCT; q++ ) acc[q] = fma( mv[i], vc[h + i][q], acc[q] );
          }
        }
      } else {
        for ( int i = 0; i < kn; i++ ) {
          const double m = __ldg( Mc + i * ldm );
#pragma unroll
          for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, vc[i][q], acc[q] );
        }
      }
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) part[ch][r][q] = acc[q];
  }
  __syncthreads();
  if ( g == 0 && live ) {
    for ( int ch = 0; ch < nch; ch++ ) {
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) cr[q] -= part[ch][r][q];
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) dest[q] = t.mode == 0 ? cr[q] : -cr[q];
  }
}

/* the same tile by ONE warp (K <= SW_WARP_K): lane = output row, the chunks one after the other -- same chains, same order of subtraction.
 * A CTA then has eight such tiles in flight instead of one with most of its warps idle.  wsm: SW_WARP_K * PCT doubles of this warp. */
template <int PCT>
__device__ __forceinline__ void sw_gemv_warp( const SvTask& t, int o0, const SvBases& B, int p, int pitch, double* wsm ) {
  double ( *vs )[PCT] = reinterpret_cast<double ( * )[PCT]>( wsm );
  const int pcn = min( PCT, p );
  const int lane = threadIdx.x & 31;
  const int row = o0 + lane;
  const bool live = row < t.N;
  const double* v = B.b[t.vsel] + t.voff * pitch;
  double* dest = B.b[t.dsel] + ( t.doff + row ) * pitch;
  const int k_lo = t.tri == 2 ? ( o0 & ~( KC - 1 ) ) : 0, k_hi = t.tri == 1 ? min( t.K, ( ( o0 + GV_ROWS - 1 ) | ( KC - 1 ) ) + 1 ) : t.K;
  for ( int e = lane; e < ( k_hi - k_lo ) * PCT; e += 32 ) { const int kk = e / PCT, q = e % PCT; vs[kk][q] = q < pcn ? __ldcg( v + ( int64_t )( k_lo + kk ) * pitch + q ) : 0.0; }
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( t.mode == 0 && live && q < pcn ) ? __ldcg( dest + q ) : 0.0;
  __syncwarp();
  const int nch = ( k_hi - k_lo + KC - 1 ) / KC;
  const int64_t ldm = t.ld;
  if ( live ) {
    for ( int ch = 0; ch < nch; ch++ ) {
      const int k0 = k_lo + ch * KC, kn = min( KC, k_hi - k0 );
      double acc[PCT];
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
      const double* Mc = t.M + row + ( int64_t )k0 * ldm;
      const double ( *vc )[PCT] = vs + ch * KC;
      if ( kn == KC ) {
#pragma unroll 1
        for ( int h = 0; h < KC; h += 16 ) {
          double mv[16];
#pragma unroll
          for ( int i = 0; i < 16; i++ ) mv[i] = __ldg( Mc + ( h + i ) * ldm );
          pin8( mv[0], mv[1], mv[2], mv[3], mv[4], mv[5], mv[6], mv[7] ); pin8( mv[8], mv[9], mv[10], mv[11], mv[12], mv[13], mv[14], mv[15] );
#pragma unroll
          for ( int i = 0; i < 16; i++ ) {
#pragma unroll
            for ( int q = 0; q < PCT; q++ ) acc[q] = fma( mv[i], vc[h + i][q], acc[q] );
          }
        }
      } else {
        for ( int i = 0; i < kn; i++ ) {
          const double m = __ldg( Mc + i * ldm );
#pragma unroll
          for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, vc[i][q], acc[q] );
        }
      }
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) cr[q] -= acc[q];
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) dest[q] = t.mode == 0 ? cr[q] : -cr[q];
  }
}

/* dest -= M' v over 64 output columns, 8 per warp (the arithmetic of sv_gemvT_kernel: per chunk of 64 rows one chain per column,
 * chunk sums subtracted in ascending order).  vidx != NULL: row kk of v is row vidx[kk] of the base (x of the ancestors through the
 * row list of the supernode -- the backward sweep needs no separate gather pass). */
template <int PCT>
__device__ __forceinline__ void sw_gemvT( const SvTask& t, int o0, const int* __restrict__ vidx, const SvBases& B, int p, int pitch, double* dsm ) {
  double ( *xs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm );
  double ( *slab )[8][KC + 1] = reinterpret_cast<double ( * )[8][KC + 1]>( dsm + GT_XS * PCT );
  const int pcn = min( PCT, p );
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbase = o0 + warp * 8;
  const double* v = vidx ? B.b[t.vsel] : B.b[t.vsel] + t.voff * pitch;
  const int mycol = cbase + ( lane & 7 );
  double* dest = B.b[t.dsel] + ( t.doff + mycol ) * pitch;
  const bool owner = lane < 8 && mycol < t.N;
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( owner && q < pcn ) ? __ldcg( dest + q ) : 0.0;
  const int ncol = max( 0, min( 8, t.N - cbase ) );
  const int64_t ldm = t.ld;
  const double* Mw = t.M + ( int64_t )cbase * ldm + lane;
  double pre[16];
  auto fetch = [&]( int k0 ) {
    const bool r0 = k0 + lane < t.K, r1 = k0 + lane + 32 < t.K;
#pragma unroll
    for ( int j = 0; j < 8; j++ ) {
      pre[2 * j] = ( j < ncol && r0 ) ? __ldg( Mw + k0 + j * ldm ) : 0.0;
      pre[2 * j + 1] = ( j < ncol && r1 ) ? __ldg( Mw + k0 + 32 + j * ldm ) : 0.0;
    }
  };
  if ( t.K > 0 ) fetch( 0 );
  for ( int kb = 0; kb < t.K; kb += GT_XS ) {
    const int rows = min( GT_XS, t.K - kb );
    __syncthreads();
    if ( vidx ) { for ( int e = tid; e < rows * PCT; e += SW_THREADS ) { const int kk = e / PCT, q = e % PCT; xs[kk][q] = q < pcn ? __ldcg( v + ( int64_t )vidx[t.voff + kb + kk] * pitch + q ) : 0.0; } }
    else { for ( int e = tid; e < rows * PCT; e += SW_THREADS ) { const int kk = e / PCT, q = e % PCT; xs[kk][q] = q < pcn ? __ldcg( v + ( int64_t )( kb + kk ) * pitch + q ) : 0.0; } }
    __syncthreads();
    for ( int k0 = 0; k0 < rows; k0 += KC ) {
#pragma unroll
      for ( int j = 0; j < 8; j++ ) { slab[warp][j][lane] = pre[2 * j]; slab[warp][j][lane + 32] = pre[2 * j + 1]; }
      if ( kb + k0 + KC < t.K ) fetch( kb + k0 + KC );
      __syncwarp();
      if ( owner ) {
        const int kn = min( KC, rows - k0 );
        double acc[PCT];
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
        const double* sl = slab[warp][lane & 7];
        if ( kn == KC ) {
#pragma unroll 16
          for ( int i = 0; i < KC; i++ ) {
            const double m = sl[i];
#pragma unroll
            for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, xs[k0 + i][q], acc[q] );
          }
        } else {
          for ( int i = 0; i < kn; i++ ) {
            const double m = sl[i];
#pragma unroll
            for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, xs[k0 + i][q], acc[q] );
          }
        }
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) cr[q] -= acc[q];
      }
      __syncwarp();
    }
  }
  if ( owner ) {
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) dest[q] = cr[q];
  }
}

/* forward gather of one parent, destination rows [lo, hi): y[pivot rows] += w_child, w_parent[update rows] += w_child, children in
 * list order (the order of fwd_gather_kernel, so the sums have the same bits) */
__device__ __forceinline__ void sw_fwd_gather( const SnMeta& P, int lo, int hi, const SnMeta* __restrict__ sn, const int* __restrict__ child_list,
                                               const int* __restrict__ relidx, double* __restrict__ y, double* __restrict__ w, int pitch ) {
  int tqs = 0; while ( tqs < 2 && ( 1 << tqs ) < pitch ) tqs++;
  const int TQ = 1 << tqs, tq = threadIdx.x & ( TQ - 1 ), ti = threadIdx.x >> tqs, tr = SW_THREADS >> tqs;
  for ( int ci = P.child_begin; ci < P.child_end; ci++ ) {
    const SnMeta C = sn[child_list[ci]];
    const int uc = C.f - C.k;
    const int* rel = relidx + C.rowptr + C.k;
    int a = 0, b = uc;
    if ( lo > 0 || hi < P.f ) {
      { int x = 0, z = uc; while ( x < z ) { const int m = ( x + z ) >> 1; if ( rel[m] < lo ) x = m + 1; else z = m; } a = x; }
      { int x = a, z = uc; while ( x < z ) { const int m = ( x + z ) >> 1; if ( rel[m] < hi ) x = m + 1; else z = m; } b = x; }
    }
    for ( int i = a + ti; i < b; i += tr ) {
      const int d = rel[i];
      const double* src = w + ( C.wgl + i ) * pitch;
      double* dst = d < P.k ? y + ( int64_t )( P.c0 + d ) * pitch : w + ( P.wgl + ( d - P.k ) ) * pitch;
      for ( int q = tq; q < pitch; q += TQ ) dst[q] = __ldcg( dst + q ) + __ldcg( src + q );
    }
    __syncthreads();
  }
}

/* ---- small supernodes (k <= KC pivots, f <= 512 rows): one WARP takes a whole supernode, eight supernodes to a unit.  Nine tenths of the
 * supernodes live here with ~13 KB of factor each; what they cost is latency (a dozen dependent memory round trips), so the answer is
 * eight times as many of them in flight per CTA.  The arithmetic is that of sv_small_fwd_kernel / sv_small_bwd_kernel (single-chunk
 * chains in ascending k), only the work split changes.  wsm: this warp's 2 * KC * PCT doubles of shared memory. */
template <int PCT>
__device__ __forceinline__ void sw_small_fwd_warp( const SnMeta& S, const SnMeta* __restrict__ sn, const int* __restrict__ child_list, const int* __restrict__ relidx,
                                                   const double* __restrict__ L, const double* __restrict__ sinvF, double* __restrict__ Y, double* __restrict__ Z, double* __restrict__ W,
                                                   int p, int pitch, double* wsm ) {
  double ( *ys )[PCT] = reinterpret_cast<double ( * )[PCT]>( wsm );
  double ( *zs )[PCT] = reinterpret_cast<double ( * )[PCT]>( wsm + KC * PCT );
  const int pcn = min( PCT, p );
  const int k = S.k, u = S.f - k, lane = threadIdx.x & 31;
  const int64_t ldx = ( k + 1 ) & ~1;
  for ( int e = lane; e < k * PCT; e += 32 ) { const int c = e / PCT, q = e % PCT; ys[c][q] = q < pcn ? __ldcg( Y + ( int64_t )( S.c0 + c ) * pitch + q ) : 0.0; }
  __syncwarp();
  for ( int ci = S.child_begin; ci < S.child_end; ci++ ) {      /* gather: children in list order, as fwd_gather_kernel */
    const SnMeta C = sn[child_list[ci]];
    const int uc = C.f - C.k;
    const int* rel = relidx + C.rowptr + C.k;
    for ( int e = lane; e < uc * PCT; e += 32 ) {
      const int i = e / PCT, q = e % PCT;
      if ( q < pcn ) {
        const int d = rel[i];
        const double sv = __ldcg( W + ( C.wgl + i ) * pitch + q );
        if ( d < k ) ys[d][q] += sv; else { double* dst = W + ( S.wgl + ( d - k ) ) * pitch + q; *dst = __ldcg( dst ) + sv; }
      }
    }
    __syncwarp();
  }
  for ( int r = lane; r < k; r += 32 ) {                       /* z = X y */
    const double* Xr = sinvF + S.sinv + r;
    double acc[PCT];
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
    int c = 0;
    for ( ; c + 8 <= k; c += 8 ) {
      double mv[8];
#pragma unroll
      for ( int j = 0; j < 8; j++ ) mv[j] = __ldg( Xr + ( c + j ) * ldx );
#pragma unroll
      for ( int j = 0; j < 8; j++ ) {
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = fma( mv[j], ys[c + j][q], acc[q] );
      }
    }
    for ( ; c < k; c++ ) {
      const double m = __ldg( Xr + c * ldx );
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, ys[c][q], acc[q] );
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) { const double z = -( 0.0 - acc[q] ); zs[r][q] = z; if ( q < pcn ) Z[( int64_t )( S.c0 + r ) * pitch + q] = z; }
  }
  __syncwarp();
  const double* Lp = L + S.lptr + k;
  for ( int i = lane; i < u; i += 32 ) {                       /* w[update rows] -= L21 z */
    double acc[PCT];
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
    const double* Lr = Lp + i;
    int c = 0;
    for ( ; c + 8 <= k; c += 8 ) {
      double mv[8];
#pragma unroll
      for ( int j = 0; j < 8; j++ ) mv[j] = __ldg( Lr + ( int64_t )( c + j ) * S.ld );
#pragma unroll
      for ( int j = 0; j < 8; j++ ) {
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = fma( mv[j], zs[c + j][q], acc[q] );
      }
    }
    for ( ; c < k; c++ ) {
      const double m = __ldg( Lr + ( int64_t )c * S.ld );
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, zs[c][q], acc[q] );
    }
    double* wp = W + ( S.wgl + i ) * pitch;
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) wp[q] = __ldcg( wp + q ) - acc[q];
  }
}

/* backward: v = z - L21' x[ancestors] (chunks of 64 update rows, one chain per pivot column and chunk), x = X' v.  A lane owns pivot
 * columns lane and lane + 32 and walks down their contiguous rows. */
template <int PCT>
__device__ __forceinline__ void sw_small_bwd_warp( const SnMeta& S, const double* __restrict__ M, const double* __restrict__ sinvB, const int* __restrict__ rowidx,
                                                   double* __restrict__ Y, const double* __restrict__ Z, int p, int pitch, double* wsm ) {
  double ( *xs )[PCT] = reinterpret_cast<double ( * )[PCT]>( wsm );
  double ( *vs )[PCT] = reinterpret_cast<double ( * )[PCT]>( wsm + KC * PCT );
  const int pcn = min( PCT, p );
  const int k = S.k, u = S.f - k, lane = threadIdx.x & 31;
  const int64_t ldx = ( k + 1 ) & ~1;
  const int* ri = rowidx + S.rowptr + k;
  const double* Mp = M + S.lptr + k;
  double cr[2][PCT];
#pragma unroll
  for ( int h = 0; h < 2; h++ )
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) { const int c = lane + 32 * h; cr[h][q] = ( c < k && q < pcn ) ? __ldcg( Z + ( int64_t )( S.c0 + c ) * pitch + q ) : 0.0; }
  for ( int k0 = 0; k0 < u; k0 += KC ) {
    const int kn = min( KC, u - k0 );
    __syncwarp();
    for ( int e = lane; e < kn * PCT; e += 32 ) { const int i = e / PCT, q = e % PCT; xs[i][q] = q < pcn ? __ldcg( Y + ( int64_t )ri[k0 + i] * pitch + q ) : 0.0; }
    __syncwarp();
#pragma unroll
    for ( int h = 0; h < 2; h++ ) {
      const int c = lane + 32 * h;
      if ( c < k ) {
        const double* Mc = Mp + k0 + ( int64_t )c * S.ld;
        double acc[PCT];
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
        int i = 0;
        for ( ; i + 8 <= kn; i += 8 ) {
          double mv[8];
#pragma unroll
          for ( int j = 0; j < 8; j++ ) mv[j] = __ldg( Mc + i + j );
#pragma unroll
          for ( int j = 0; j < 8; j++ ) {
#pragma unroll
            for ( int q = 0; q < PCT; q++ ) acc[q] = fma( mv[j], xs[i + j][q], acc[q] );
          }
        }
        for ( ; i < kn; i++ ) {
          const double m = __ldg( Mc + i );
#pragma unroll
          for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, xs[i][q], acc[q] );
        }
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) cr[h][q] -= acc[q];
      }
    }
  }
  __syncwarp();
#pragma unroll
  for ( int h = 0; h < 2; h++ ) { const int c = lane + 32 * h; if ( c < k ) {
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) vs[c][q] = cr[h][q]; } }
  __syncwarp();
  for ( int r = lane; r < k; r += 32 ) {
    const double* Xr = sinvB + S.sinv + r;
    double acc[PCT];
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
    int c = 0;
    for ( ; c + 8 <= k; c += 8 ) {
      double mv[8];
#pragma unroll
      for ( int j = 0; j < 8; j++ ) mv[j] = __ldg( Xr + ( c + j ) * ldx );
#pragma unroll
      for ( int j = 0; j < 8; j++ ) {
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = fma( mv[j], vs[c + j][q], acc[q] );
      }
    }
    for ( ; c < k; c++ ) {
      const double m = __ldg( Xr + c * ldx );
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, vs[c][q], acc[q] );
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) Y[( int64_t )( S.c0 + r ) * pitch + q] = -( 0.0 - acc[q] );
  }
}

struct SweepArgs {
  const SvUnit* units; int nunits;
  const SvUnit* inner;             /* the units of the UK_BATCH units, one per warp */
  int* head;                       /* ticket counter of this sweep */
  int* cnt;                        /* dependency counters (zeroed before the forward sweep) */
  int* err;                        /* 0, or 1 + the unit that gave up waiting */
  const SvTask* tasks;
  const SnMeta* sn; const int* child_list; const int* relidx; const int* rowidx;
  const double* Lf; const double* Lb;      /* factor panels read by the forward / backward sweep */
  const double* sinvF; const double* sinvB;
  SvBases B;                       /* 0: y, 1: z, 2: w (one region per supernode) */
  int p, pitch;
  unsigned long long* trace;       /* diagnostics (B200_SWEEP_TRACE): per unit { globaltimer at the start, clock64 at start / inputs ready / done, SM id }; NULL = off */
};

/* spin (one warp) until cnt[w1 .. w1+n1) >= v1 and cnt[w2] >= v2; false = gave up (error flag set by somebody, or the spin bound) */
__device__ __forceinline__ bool sw_wait( const SweepArgs& a, int ui, int w1, int n1, int v1, int w2, int v2 ) {
  const int lane = threadIdx.x & 31;
  unsigned spins = 0;
  for ( ;; ) {
    bool mine = true;
    for ( int i = lane; i < n1; i += 32 ) if ( ld_acquire( a.cnt + w1 + i ) < v1 ) mine = false;
    if ( lane == 0 && w2 >= 0 && ld_acquire( a.cnt + w2 ) < v2 ) mine = false;
    if ( __all_sync( 0xffffffffu, mine ) ) break;
    if ( ( ++spins & 255u ) == 0 ) {
      int bad = lane == 0 ? ( ld_acquire( a.err ) != 0 || spins > SW_SPIN_LIMIT ) : 0;
      bad = __shfl_sync( 0xffffffffu, bad, 0 );
      if ( bad ) { if ( lane == 0 ) atomicCAS( a.err, 0, ui + 1 ); return false; }
    }
    __nanosleep( 64 );
  }
  __threadfence();
  return true;
}

template <int PCT>
__global__ void __launch_bounds__( SW_THREADS, PCT == 4 ? 3 : SW_MINB ) sv_sweep_kernel( const SweepArgs a ) {
  extern __shared__ double dsm[];
  __shared__ int4 s_desc[2][4];    /* unit descriptors: the current one and the next (fetched while the current one runs) */
  __shared__ int s_ticket[2];
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  /* thread 32 keeps the CTA supplied: tickets are taken two units ahead and descriptors one unit ahead, so that neither round trip is exposed.
   * A CTA still runs its units in ticket order, which is all the no-deadlock argument needs: the first unfinished unit of the list is either some
   * CTA's current unit with all its inputs ready, or held by a CTA whose earlier tickets are all finished. */
  int t_next = 0, t_next2 = 0;
  if ( tid == 32 ) {
    const int t0 = atomicAdd( a.head, 1 ); t_next = atomicAdd( a.head, 1 ); t_next2 = atomicAdd( a.head, 1 );
    s_ticket[0] = t0;
    if ( t0 < a.nunits ) { const int4* src = reinterpret_cast<const int4*>( a.units + t0 ); for ( int i = 0; i < 4; i++ ) s_desc[0][i] = src[i]; }
    s_abort = 0;
  }
  int buf = 0;
  for ( ;; ) {
    __syncthreads();               /* descriptor `buf` is in place; the previous unit's shared memory is free */
    const int ui = s_ticket[buf];
    if ( ui >= a.nunits || s_abort ) return;
    const SvUnit u = *reinterpret_cast<const SvUnit*>( &s_desc[buf][0] );
    unsigned long long* trc = a.trace ? a.trace + 5 * ( size_t )ui : nullptr;
    if ( trc && tid == 0 ) { unsigned long long g; unsigned smid; asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( g ) ); asm volatile( "mov.u32 %0, %%smid;" : "=r"( smid ) );
      trc[0] = g; trc[1] = ( unsigned long long )clock64(); trc[4] = smid; }
    int t_new = 0;
    if ( tid == 32 ) {             /* in flight during this unit: the descriptor of the next ticket (global -> shared, asynchronously) and a new ticket */
      if ( t_next < a.nunits ) { const int4* src = reinterpret_cast<const int4*>( a.units + t_next ); for ( int i = 0; i < 4; i++ ) cp_async16( &s_desc[buf ^ 1][i], src + i, 16 ); }
      cp_async_commit();
      t_new = atomicAdd( a.head, 1 );
    }
    if ( u.kind == UK_BATCH ) {
      /* up to eight latency-bound units, one per warp, each with its own wait and its own signals */
      if ( warp < u.o0 ) {
        const SvUnit iu = a.inner[u.task + warp];
        if ( !sw_wait( a, ui, iu.w1, iu.n1, iu.v1, iu.w2, iu.v2 ) ) { if ( lane == 0 ) s_abort = 1; }
        else {
          double* wsm = dsm + warp * ( SW_WARP_K * PCT );
          if ( iu.kind == UK_GEMV ) { const SvTask t = a.tasks[iu.task]; sw_gemv_warp<PCT>( t, iu.o0, a.B, a.p, a.pitch, wsm ); }
          else {
            const SnMeta S = a.sn[iu.sn];
            if ( iu.kind == UK_FWD_SMALL ) sw_small_fwd_warp<PCT>( S, a.sn, a.child_list, a.relidx, a.Lf, a.sinvF, a.B.b[0], a.B.b[1], a.B.b[2], a.p, a.pitch, wsm );
            else sw_small_bwd_warp<PCT>( S, a.Lb, a.sinvB, a.rowidx, a.B.b[0], a.B.b[1], a.p, a.pitch, wsm );
          }
          __syncwarp();
          if ( lane == 0 ) { __threadfence(); if ( iu.sig >= 0 ) atomicAdd( a.cnt + iu.sig, 1 ); if ( iu.up >= 0 ) atomicAdd( a.cnt + iu.up, 1 ); }
        }
      }
      if ( trc ) { __syncthreads(); if ( tid == 0 ) { trc[2] = trc[1]; trc[3] = ( unsigned long long )clock64(); } }
    } else {
      if ( warp == 0 ) { if ( !sw_wait( a, ui, u.w1, u.n1, u.v1, u.w2, u.v2 ) && lane == 0 ) s_abort = 1; }
      __syncthreads();
      if ( s_abort ) return;
      if ( trc && tid == 0 ) trc[2] = ( unsigned long long )clock64();
      switch ( u.kind ) {
        case UK_GEMV: { const SvTask t = a.tasks[u.task]; sw_gemv<PCT>( t, u.o0, a.B, a.p, a.pitch, dsm ); break; }
        case UK_GEMVT: {
          const SvTask t = a.tasks[u.task];
          const int* vidx = nullptr;
          if ( u.aux ) { const SnMeta S = a.sn[u.sn]; vidx = a.rowidx + S.rowptr + S.k; }
          sw_gemvT<PCT>( t, u.o0, vidx, a.B, a.p, a.pitch, dsm );
          break; }
        case UK_FWD_GATHER: { const SnMeta P = a.sn[u.sn]; sw_fwd_gather( P, u.o0, u.aux, a.sn, a.child_list, a.relidx, a.B.b[0], a.B.b[2], a.pitch ); break; }
      }
      __syncthreads();
      if ( tid == 0 ) {            /* publish: the CTA's writes, then the counters */
        __threadfence();
        if ( u.sig >= 0 ) atomicAdd( a.cnt + u.sig, 1 );
        if ( u.up >= 0 ) atomicAdd( a.cnt + u.up, 1 );
        if ( trc ) trc[3] = ( unsigned long long )clock64();
      }
    }
    if ( tid == 32 ) {
      s_ticket[buf ^ 1] = t_next;
      cp_async_wait<0>();
      t_next = t_next2; t_next2 = t_new;
    }
    buf ^= 1;
  }
}

}  // namespace b200
