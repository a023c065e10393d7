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

enum { UK_FWD_GATHER = 0, UK_FWD_SMALL, UK_GEMV, UK_GEMVT, UK_BWD_SMALL };

struct SvUnit {          /* 64 bytes */
  int kind;
  int task;              /* UK_GEMV / UK_GEMVT: index into the task list                                                          */
  int o0;                /* first output of the tile (UK_GEMV: 32 rows, UK_GEMVT: 64 columns); UK_FWD_GATHER: first destination row */
  int aux;               /* UK_FWD_GATHER: end of the destination rows; UK_GEMVT: 1 = v is x gathered through the row list of sn   */
  int sn;                /* supernode                                                                                             */
  int w1, n1, v1;        /* wait until cnt[w1 .. w1+n1) >= v1  (n1 = 0: nothing to wait for)                                      */
  int w2, v2;            /* ... and cnt[w2] >= v2              (w2 < 0: nothing)                                                  */
  int sig;               /* cnt[sig] += 1 when the unit is done (sig < 0: none)                                                   */
  int dn, dtot, up;      /* cnt[dn] += 1; the unit that makes it dtot adds 1 to cnt[up] (supernode complete; up < 0: none)        */
  int pad0, pad1;
};
static_assert( sizeof( SvUnit ) == 64, "SvUnit is four 16-byte words" );

constexpr int SW_THREADS = 256;
constexpr int SW_GT_COLS = 64;                        /* columns per gemvT unit: 8 warps x 8 columns */
constexpr unsigned SW_SPIN_LIMIT = 1u << 22;          /* polls before a unit gives up (a few seconds): a broken schedule must not hang the GPU */
template <int PCT> constexpr int sw_smem() {
  return ( ( SB_WIDE + SVG_MAXCH * GV_ROWS ) * PCT > GT_XS * PCT + 8 * 8 * ( KC + 1 ) ? ( SB_WIDE + SVG_MAXCH * GV_ROWS ) * PCT : GT_XS * PCT + 8 * 8 * ( KC + 1 ) ) * 8;
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
    if ( vidx ) { for ( int e = tid; e < rows * PCT; e += SW_THREADS ) { const int kk = e / PCT, q = e % PCT; xs[kk][q] = q < pcn ? __ldcg( v + ( int64_t )vidx[kb + kk] * pitch + q ) : 0.0; } }
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

/* a whole small supernode (k <= KC pivots, f <= 512 rows), forward: gather the children, z = X y, w[update rows] -= L21 z */
template <int PCT>
__device__ __forceinline__ void sw_small_fwd( const SnMeta& S, const SnMeta* __restrict__ sn, const int* __restrict__ child_list, const int* __restrict__ relidx,
                                              const double* __restrict__ L, const double* __restrict__ sinvF, double* __restrict__ Y, double* __restrict__ Z, double* __restrict__ W,
                                              int p, int pitch, double* dsm ) {
  double ( *ys )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm );
  double ( *zs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm + KC * PCT );
  const int pcn = min( PCT, p );
  const int k = S.k, u = S.f - k, tid = threadIdx.x;
  const int64_t ldx = ( k + 1 ) & ~1;
  for ( int e = tid; e < k * PCT; e += SW_THREADS ) { const int c = e / PCT, q = e % PCT; ys[c][q] = q < pcn ? __ldcg( Y + ( int64_t )( S.c0 + c ) * pitch + q ) : 0.0; }
  __syncthreads();
  for ( int ci = S.child_begin; ci < S.child_end; ci++ ) {
    const SnMeta C = sn[child_list[ci]];
    const int uc = C.f - C.k;
    const int* rel = relidx + C.rowptr + C.k;
    for ( int e = tid; e < uc * PCT; e += SW_THREADS ) {
      const int i = e / PCT, q = e % PCT;
      if ( q < pcn ) {
        const int d = rel[i];
        const double sv = __ldcg( W + ( C.wgl + i ) * pitch + q );
        if ( d < k ) ys[d][q] += sv; else { double* dst = W + ( S.wgl + ( d - k ) ) * pitch + q; *dst = __ldcg( dst ) + sv; }
      }
    }
    __syncthreads();
  }
  if ( tid < k ) {
    const double* Xr = sinvF + S.sinv + tid;
    double acc[PCT];
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
    for ( int c = 0; c < k; c++ ) {
      const double m = __ldg( Xr + c * ldx );
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, ys[c][q], acc[q] );
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) { const double z = -( 0.0 - acc[q] ); zs[tid][q] = z; if ( q < pcn ) Z[( int64_t )( S.c0 + tid ) * pitch + q] = z; }
  }
  __syncthreads();
  const double* Lp = L + S.lptr + k;
  for ( int i = tid; i < u; i += SW_THREADS ) {
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

/* a whole small supernode, backward: v = z - L21' x[ancestors] (chunks of 64 update rows), x = X' v */
template <int PCT>
__device__ __forceinline__ void sw_small_bwd( const SnMeta& S, const double* __restrict__ M, const double* __restrict__ sinvB, const int* __restrict__ rowidx,
                                              double* __restrict__ Y, const double* __restrict__ Z, int p, int pitch, double* dsm ) {
  double ( *slab )[KC + 1] = reinterpret_cast<double ( * )[KC + 1]>( dsm );                 /* [col][row of the chunk] */
  double ( *xs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm + KC * ( KC + 1 ) );
  double ( *vs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm + KC * ( KC + 1 ) + KC * PCT );
  const int pcn = min( PCT, p );
  const int k = S.k, u = S.f - k, tid = threadIdx.x;
  const int64_t ldx = ( k + 1 ) & ~1;
  const int* ri = rowidx + S.rowptr + k;
  const double* Mp = M + S.lptr + k;
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( tid < k && q < pcn ) ? __ldcg( Z + ( int64_t )( S.c0 + tid ) * pitch + q ) : 0.0;
  for ( int k0 = 0; k0 < u; k0 += KC ) {
    const int kn = min( KC, u - k0 );
    __syncthreads();
    for ( int e = tid; e < kn * PCT; e += SW_THREADS ) { const int i = e / PCT, q = e % PCT; xs[i][q] = q < pcn ? __ldcg( Y + ( int64_t )ri[k0 + i] * pitch + q ) : 0.0; }
    for ( int e = tid; e < k * KC; e += SW_THREADS ) { const int i = e & ( KC - 1 ), c = e >> 6; if ( i < kn ) slab[c][i] = __ldg( Mp + k0 + i + ( int64_t )c * S.ld ); }
    __syncthreads();
    if ( tid < k ) {
      double acc[PCT];
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
      for ( int i = 0; i < kn; i++ ) {
        const double m = slab[tid][i];
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, xs[i][q], acc[q] );
      }
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) cr[q] -= acc[q];
    }
  }
  __syncthreads();
  if ( tid < k ) {
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) vs[tid][q] = cr[q];
  }
  __syncthreads();
  if ( tid < k ) {
    const double* Xr = sinvB + S.sinv + tid;
    double acc[PCT];
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
    for ( int c = 0; c < k; c++ ) {
      const double m = __ldg( Xr + c * ldx );
#pragma unroll
      for ( int q = 0; q < PCT; q++ ) acc[q] = fma( m, vs[c][q], acc[q] );
    }
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) Y[( int64_t )( S.c0 + tid ) * pitch + q] = -( 0.0 - acc[q] );
  }
}

struct SweepArgs {
  const SvUnit* units; int nunits;
  int* head;                       /* ticket counter of this sweep */
  int* cnt;                        /* dependency counters (zeroed before the forward sweep) */
  int* err;                        /* 0, or 1 + the unit that gave up waiting */
  const SvTask* tasks;
  const SnMeta* sn; const int* child_list; const int* relidx; const int* rowidx;
  const double* Lf; const double* Lb;      /* factor panels read by the forward / backward sweep */
  const double* sinvF; const double* sinvB;
  SvBases B;                       /* 0: y, 1: z, 2: w (one region per supernode) */
  int p, pitch;
};

template <int PCT>
__global__ void __launch_bounds__( SW_THREADS, PCT == 4 ? 3 : 4 ) sv_sweep_kernel( const SweepArgs a ) {
  extern __shared__ double dsm[];
  __shared__ int s_unit, s_abort;
  const int tid = threadIdx.x;
  for ( ;; ) {
    __syncthreads();               /* the previous unit's shared memory is free, s_unit may be overwritten */
    if ( tid == 0 ) { s_unit = atomicAdd( a.head, 1 ); s_abort = 0; }
    __syncthreads();
    const int ui = s_unit;
    if ( ui >= a.nunits ) return;
    const SvUnit u = a.units[ui];
    /* ---- wait for the inputs of this unit (warp 0 polls; every counter only grows) */
    if ( tid < 32 ) {
      unsigned spins = 0; bool ok = false;
      while ( !ok ) {
        bool mine = true;
        for ( int i = tid; i < u.n1; i += 32 ) if ( ld_acquire( a.cnt + u.w1 + i ) < u.v1 ) mine = false;
        if ( tid == 0 && u.w2 >= 0 && ld_acquire( a.cnt + u.w2 ) < u.v2 ) mine = false;
        ok = __all_sync( 0xffffffffu, mine );
        if ( !ok ) {
          if ( ( ++spins & 255u ) == 0 ) {
            int bad = tid == 0 ? ( ld_acquire( a.err ) != 0 || spins > SW_SPIN_LIMIT ) : 0;
            bad = __shfl_sync( 0xffffffffu, bad, 0 );
            if ( bad ) { if ( tid == 0 ) { atomicCAS( a.err, 0, ui + 1 ); s_abort = 1; } break; }
          }
          __nanosleep( 64 );
        }
      }
      __threadfence();
    }
    __syncthreads();
    if ( s_abort ) return;
    switch ( u.kind ) {
      case UK_GEMV: { const SvTask t = a.tasks[u.task]; sw_gemv<PCT>( t, u.o0, a.B, a.p, a.pitch, dsm ); break; }
      case UK_GEMVT: {
        const SvTask t = a.tasks[u.task];
        const int* vidx = nullptr;
        if ( u.aux ) { const SnMeta S = a.sn[u.sn]; vidx = a.rowidx + S.rowptr + S.k; }
        sw_gemvT<PCT>( t, u.o0, vidx, a.B, a.p, a.pitch, dsm );
        break; }
      case UK_FWD_GATHER: { const SnMeta P = a.sn[u.sn]; sw_fwd_gather( P, u.o0, u.aux, a.sn, a.child_list, a.relidx, a.B.b[0], a.B.b[2], a.pitch ); break; }
      case UK_FWD_SMALL: { const SnMeta S = a.sn[u.sn]; sw_small_fwd<PCT>( S, a.sn, a.child_list, a.relidx, a.Lf, a.sinvF, a.B.b[0], a.B.b[1], a.B.b[2], a.p, a.pitch, dsm ); break; }
      case UK_BWD_SMALL: { const SnMeta S = a.sn[u.sn]; sw_small_bwd<PCT>( S, a.Lb, a.sinvB, a.rowidx, a.B.b[0], a.B.b[1], a.p, a.pitch, dsm ); break; }
    }
    __syncthreads();
    if ( tid == 0 ) {              /* publish: the CTA's writes, then the counters */
      __threadfence();
      if ( u.sig >= 0 ) atomicAdd( a.cnt + u.sig, 1 );
      const int before = atomicAdd( a.cnt + u.dn, 1 );
      if ( before == u.dtot - 1 && u.up >= 0 ) { __threadfence(); atomicAdd( a.cnt + u.up, 1 ); }
    }
  }
}

}  // namespace b200
