/* b200_kernels.cuh -- device kernels of the multifrontal factor / solve (sm_100a).
 *
 * The numeric work the reference hands to cholmod_factorize / cholmod_solve
 * (src/solver_cholmod.c:272,303), umfpack_di_numeric / _solve
 * (src/solver_umfpack.c:98,133) and dmumps_c JOB=2/3 (src/solver_mumps.c:211-213,
 * 274-276) is done here by hand-written kernels:
 *
 *   scatter_A          CSC values -> frontal panels through the analysis map   (HBM bound)
 *   extend_add         child contribution blocks -> parent front, by parent
 *                      column block so that no atomics are needed               (HBM bound)
 *   potrf_diag         <=64x64 diagonal block Cholesky + triangular inverse     (latency)
 *   getrf_diag         <=64x64 diagonal block LU, partial pivoting inside it   (latency)
 *   trsm_rows          panel rows x inverse diagonal block                     (FP64)
 *   gemm_nt_dmma_bulk  C -= A * B'  on FP64 tensor-core tiles
 *                      (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no
 *                      f64 kind), operands staged by cp.async.bulk + mbarriers (FP64 roof)
 *   solve kernels      level-set forward/backward sweeps over supernode blocks (HBM bound)
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

namespace b200 {

constexpr int NB = 64;        /* diagonal block / inner panel width */
constexpr int GETRF_SMEM = ( 3 * NB * ( NB + 1 ) + 2 * NB ) * 8;
constexpr int TRSM_SMEM = ( NB * NB + NB * 128 ) * 8;

/* ------------------------------------------------------------------ tasks */
struct GemmTask {
  double* C; const double* A; const double* B;
  int ldc, lda, ldb;
  int M, N, K;
  int lower;   /* 1: store only elements with r + diag >= c */
  int diag;    /* (first row of C in the front) - (first col of C in the front) */
  int mode;    /* 0: C -= A B' ;  1: C = A B' (panel rows times the inverse diagonal block; C may alias A when N, K <= 64) */
};
struct Tile { int task; unsigned short tm, tn; };

struct DiagTask {
  double* A; double* AU; int lda; int nb;   /* A: block in the L panel; AU: same block of the U' panel (LU) */
  double* inv;      /* nb x nb, ld nb: inverse of L (Cholesky) / of unit-L (LU)      */
  double* inv2;     /* LU: inverse of U, stored transposed                            */
  int* piv;         /* LU: pivot row (block-local) chosen for each column            */
  double* inv3;     /* LU: inv(unit-L) with the block's row interchanges folded in as a column permutation (full nb x nb): the rows of the U' panel
                       below the block are then one tensor-pipe product  U21' = A12' inv3'  like every other row update; NULL: not wanted */
  int col;          /* global column of the block (for error reports)                */
};

struct TrsmTask {
  double* B; int ldb; int m;     /* m rows starting at B, nb columns */
  const double* T; int nb;       /* X = B * T'   (T nb x nb, ld nb)  */
  const int* piv;                /* LU U-panel: permute the nb columns of B on load (NULL = none) */
  double* out;                   /* result rows (same ld); NULL = in place */
  int full;                      /* 1: T is a full matrix (LDL': interchanges folded in), 0: lower triangular */
};
struct RowTile { int task; int r0; };

struct AsmTask { int parent; int col0; int ncols; };

struct SnMeta {           /* per supernode, device copy */
  int64_t lptr, uptr, cbptr, rowptr, invptr;
  int c0, k, f, parent, depth, child_begin, child_end, ld;   /* ld: leading dimension of the panel (f rounded up to even) */
  int64_t wptr;           /* solve workspace row offset within its level arena (tensor-path sweeps, p >= 5) */
  int64_t wgl;            /* ... within the one-region-per-supernode workspace of the persistent sweeps (p <= 4) */
  int64_t sinv;           /* offset of its solve-block inverses: block t at sinv + t * sbw^2, leading dimension = width rounded up to even */
  int sbw;                /* solve block width: SB, or SB_WIDE for the big fronts */
};


/* ---------------------------------------------------------- small helpers */
__device__ __forceinline__ void cp_async8( void* smem, const void* g, int src_bytes ) {
  unsigned s = ( unsigned )__cvta_generic_to_shared( smem );
  asm volatile( "cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"( s ), "l"( g ), "r"( src_bytes ) );
}
__device__ __forceinline__ void cp_async16( void* smem, const void* g, int src_bytes ) {
  unsigned s = ( unsigned )__cvta_generic_to_shared( smem );
  asm volatile( "cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"( s ), "l"( g ), "r"( src_bytes ) );
}
__device__ __forceinline__ void cp_async_commit() { asm volatile( "cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile( "cp.async.wait_group %0;\n" ::"n"( N ) ); }

__device__ __forceinline__ void dmma884( double& d0, double& d1, double a, double b ) {
  asm volatile( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"( d0 ), "+d"( d1 ) : "d"( a ), "d"( b ) );
}

/* ------------------------------------------------------------- scatter A */
__global__ void scatter_A_kernel( const double* __restrict__ vals, const int64_t* __restrict__ amap, int64_t nnz,
                                  double* __restrict__ L, double* __restrict__ U ) {
  int64_t e = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = ( int64_t )gridDim.x * blockDim.x;
  for ( ; e < nnz; e += stride ) {
    int64_t d = amap[e];
    if ( d < 0 ) continue;
    double v = vals[e];
    if ( d & ( ( int64_t )1 << 62 ) ) atomicAdd( &U[d & ~( ( int64_t )1 << 62 )], v );
    else atomicAdd( &L[d], v );
  }
}

/* static-pivot threshold of the LU / LDL' diagonal-block kernels: sqrt(eps) * max|a_ij|, formed on the device so that
 * values already resident in HBM (b200_factor_device) get the same scaling as host values */
__global__ void maxabs_kernel( const double* __restrict__ vals, int64_t nnz, unsigned long long* __restrict__ out_bits ) {
  int64_t e = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = ( int64_t )gridDim.x * blockDim.x;
  double m = 0.0;
  for ( ; e < nnz; e += stride ) { const double a = fabs( vals[e] ); if ( a > m && a < CUDART_INF ) m = a; }
  for ( int o = 16; o; o >>= 1 ) m = fmax( m, __shfl_xor_sync( 0xffffffffu, m, o ) );
  if ( ( threadIdx.x & 31 ) == 0 && m > 0.0 ) atomicMax( out_bits, ( unsigned long long )__double_as_longlong( m ) );
}
__global__ void tiny_from_max_kernel( const unsigned long long* __restrict__ max_bits, double* __restrict__ tiny ) {
  const double m = __longlong_as_double( ( long long )*max_bits );
  *tiny = 1.4901161193847656e-8 * ( m > 0.0 ? m : 1.0 );
}

/* ------------------------------------------------------------ extend-add
 * One CTA per (parent, block of destination columns); inside it each warp owns
 * a fixed subset of those columns.  Children are visited in a fixed order and
 * each child touches a destination column at most once, so the sums are
 * deterministic without atomics or barriers.  Lanes run down a source column
 * (coalesced 8-byte reads); destination rows are monotone in the lane index. */
template <bool LU>
__global__ void __launch_bounds__( 256 ) extend_add_kernel( const AsmTask* __restrict__ tasks, const SnMeta* __restrict__ sn,
                                                            const int* __restrict__ child_list, const int* __restrict__ relidx,
                                                            double* __restrict__ L, double* __restrict__ U,
                                                            double* __restrict__ cb_parent, const double* __restrict__ cb_child ) {
  const AsmTask t = tasks[blockIdx.x];
  const SnMeta P = sn[t.parent];
  const int kp = P.k, fp = P.f, up = fp - kp;
  const int64_t ldp = P.ld, ldup = ( up + 1 ) & ~1;
  double* Lp = L + P.lptr;
  double* Up = LU ? U + P.uptr : nullptr;
  double* CBp = cb_parent + P.cbptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int lo = t.col0, hi = t.col0 + t.ncols;
  for ( int ci = P.child_begin; ci < P.child_end; ci++ ) {
    const SnMeta C = sn[child_list[ci]];
    const int uc = C.f - C.k;
    if ( uc == 0 ) continue;
    const int64_t lduc = ( uc + 1 ) & ~1;
    const int* rel = relidx + C.rowptr + C.k;
    const double* src = cb_child + C.cbptr;
    /* range [a,b) of child columns whose destination column is in [lo,hi): rel is increasing */
    int a = 0, b = uc;
    { int x = 0, y = uc; while ( x < y ) { int m = ( x + y ) >> 1; if ( rel[m] < lo ) x = m + 1; else y = m; } a = x; }
    { int x = a, y = uc; while ( x < y ) { int m = ( x + y ) >> 1; if ( rel[m] < hi ) x = m + 1; else y = m; } b = x; }
    for ( int j = a; j < b; j++ ) {
      const int dj = rel[j];
      /* a destination column belongs to exactly one warp, whichever child feeds it:
       * children can then be processed back to back without a barrier */
      if ( ( ( dj - lo ) & ( nwarps - 1 ) ) != warp ) continue;
      const double* s = src + ( int64_t )j * lduc;
      if ( !LU ) {
        /* lower triangle only: i >= j, hence rel[i] >= dj */
        double* d = dj < kp ? Lp + ( int64_t )dj * ldp : CBp + ( int64_t )( dj - kp ) * ldup - kp;
        int i = j + lane;
        for ( ; i + 224 < uc; i += 256 ) {     /* eight independent gather / add / scatter chains per lane in flight */
          double sv[8], dv[8]; int rv[8];
#pragma unroll
          for ( int q = 0; q < 8; q++ ) { sv[q] = s[i + 32 * q]; rv[q] = rel[i + 32 * q]; }
#pragma unroll
          for ( int q = 0; q < 8; q++ ) dv[q] = d[rv[q]];
#pragma unroll
          for ( int q = 0; q < 8; q++ ) d[rv[q]] = dv[q] + sv[q];
        }
        for ( ; i < uc; i += 32 ) d[rel[i]] += s[i];
      } else {
        /* full block: (di,dj) goes to the L panel (di>=dj, dj pivot), the U' panel
         * (di<dj, di pivot; stored transposed) or the parent's contribution block */
        auto where = [&]( int di ) -> double* {
          if ( di >= dj ) return dj < kp ? Lp + ( int64_t )dj * ldp + di : CBp + ( int64_t )( dj - kp ) * ldup + ( di - kp );
          return di < kp ? Up + ( int64_t )di * ldp + dj : CBp + ( int64_t )( dj - kp ) * ldup + ( di - kp );
        };
        int i = lane;
        for ( ; i + 224 < uc; i += 256 ) {     /* eight independent gather / add / scatter chains per lane in flight, as in the symmetric case */
          double sv[8], dv[8]; double* dp[8];
#pragma unroll
          for ( int q = 0; q < 8; q++ ) { sv[q] = s[i + 32 * q]; dp[q] = where( rel[i + 32 * q] ); }
#pragma unroll
          for ( int q = 0; q < 8; q++ ) dv[q] = *dp[q];
#pragma unroll
          for ( int q = 0; q < 8; q++ ) *dp[q] = dv[q] + sv[q];
        }
        for ( ; i < uc; i += 32 ) *where( rel[i] ) += s[i];
      }
    }
  }
}

/* ---------------------------------------------------- diagonal block: LL'
 * Cholesky of an nb x nb (nb<=64) block, then the inverse of the triangular factor (used by the
 * row update and the solves).  This kernel is the head of the dependency chain of every panel, so it
 * is built for latency: 128 threads, the lane pair (2r, 2r+1) keeps row r of the block in registers
 * (even / odd columns); one barrier per column -- every row publishes l(r,j) and its current
 * a(r,j+1), from which all threads form the next pivot themselves; the inverse is formed column by
 * column (a lane pair per column, partial sums joined by one shuffle) from the factor in shared memory. */
constexpr int POTRF_LPR = 4;                         /* lanes per row of the block: 4 lanes x 16 columns (2 x 32 measured 33.4 us: half the instructions per thread per step) */
constexpr int POTRF_THREADS = NB * POTRF_LPR;
__global__ void __launch_bounds__( POTRF_THREADS ) potrf_diag_kernel( const DiagTask* __restrict__ tasks, int* __restrict__ err ) {
  constexpr int LPR = POTRF_LPR, H = NB / LPR;
  __shared__ double Ls[NB][NB + 1];      /* the factor, row-major; column j is written by all rows in step j (it doubles as the broadcast of l(:,j)) */
  __shared__ double nxt[2][NB];          /* a(:,j+1) before the update by column j */
  const DiagTask t = tasks[blockIdx.x];
  const int nb = t.nb, r = threadIdx.x / LPR, h = threadIdx.x % LPR;
  double a[H];                           /* a[i] = A(r, LPR*i+h) */
#pragma unroll
  for ( int i = 0; i < H; i++ ) { const int c = LPR * i + h; a[i] = ( r < nb && c <= r ) ? t.A[r + ( int64_t )c * t.lda] : ( c == r ? 1.0 : 0.0 ); }   /* beyond nb: identity */
  if ( h == 0 ) nxt[0][r] = a[0];
  /* The inverse X = inv(L) is formed IN THE SAME LOOP, one row per step: row j of L is complete once column j has been published, and
   * x(j,c) = -( sum_{q=c}^{j-1} l(j,q) x(q,c) ) / l(j,j) only needs that row and the earlier rows of X, which the lanes (c, h) keep in
   * registers (x(q,c) with q = h mod LPR).  The two recurrences are independent chains of 64 dependent steps each; interleaved they hide in one
   * another's latency instead of running back to back (round 2: 42 us -> 33 us with 2 lanes per row; profiles/r1_potrf_bench.txt). */
  const int c = r;
  double x[H];                           /* x[i] = X(LPR*i+h, c) */
#pragma unroll
  for ( int i = 0; i < H; i++ ) x[i] = 0.0;
  __syncthreads();
  double dj = nxt[0][0];                 /* pivot of the current column, known to every thread */
  bool bad = false;
#pragma unroll
  for ( int j = 0; j < NB; j++ ) {
    if ( j < nb ) {                      /* uniform over the CTA: small blocks stop early */
      if ( !( dj > 0.0 ) ) { bad = true; dj = 1.0; }
      const double rs = rsqrt( dj );     /* one reciprocal square root per column; l = a * rs instead of a division */
      const double sj = dj * rs;
      if ( h == ( j % LPR ) ) {          /* the lane that holds column j of its row */
        const double l = ( r > j ) ? a[j / LPR] * rs : ( r == j ? sj : 0.0 );
        a[j / LPR] = l;
        Ls[r][j] = l;
      }
      if ( j + 1 < NB && h == ( ( j + 1 ) % LPR ) ) nxt[( j + 1 ) & 1][r] = a[( j + 1 ) / LPR];
      __syncthreads();
      if ( j + 1 < NB ) {
        const double l = Ls[r][j], ln = Ls[j + 1][j];
        dj = nxt[( j + 1 ) & 1][j + 1] - ln * ln;          /* next pivot */
#pragma unroll
        for ( int i = ( j + 1 ) / LPR; i < H; i++ ) { const int cc = LPR * i + h; if ( cc > j ) a[i] -= l * Ls[cc][j]; }     /* entries right of the diagonal (c > r) are scratch: updating them too saves the predicate */
      }
      {                                  /* row j of the inverse */
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for ( int qi = 0; LPR * qi < j; qi += 2 ) {      /* rows q = LPR*qi+h < j; x(q,c) = 0 for q < c, so no predicate on c */
          if ( LPR * qi + h < j ) s0 += Ls[j][LPR * qi + h] * x[qi];
          if ( LPR * ( qi + 1 ) + h < j ) s1 += Ls[j][LPR * ( qi + 1 ) + h] * x[qi + 1];
        }
        double sum = s0 + s1;
#pragma unroll
        for ( int o = 1; o < LPR; o <<= 1 ) sum += __shfl_xor_sync( 0xffffffffu, sum, o );
        if ( h == ( j % LPR ) ) x[j / LPR] = ( j == c ) ? rs : ( j > c ? -sum * rs : 0.0 );
      }
      if ( bad && threadIdx.x == 0 ) { atomicCAS( err, 0, t.col + j + 1 ); bad = false; }
    }
  }
  /* store L and the inverse */
#pragma unroll
  for ( int i = 0; i < H; i++ ) {
    const int cc = LPR * i + h;
    if ( r < nb && cc <= r ) t.A[r + ( int64_t )cc * t.lda] = a[i];
  }
  if ( c < nb ) {
#pragma unroll
    for ( int i = 0; i < H; i++ ) { const int row = LPR * i + h; if ( row < nb ) t.inv[row + c * nb] = x[i]; }
  }
}

/* The same operation with rolled loops and the block in shared memory: a few hundred instructions instead of
 * ~12 000, so nothing streams through the instruction cache.  Used for blocks narrower than NB (the many small
 * supernodes of the deep tree levels), where the unrolled kernel above would mostly skip over its own code.
 * One barrier per column: column j stays unscaled in place (it is never read again), every thread rescales what
 * it needs with rs = 1/sqrt(a(j,j)); a final pass turns the columns into L. */
constexpr int POTRF_S_SMEM = ( 2 * NB * ( NB + 1 ) + NB ) * 8;
__global__ void __launch_bounds__( 256 ) potrf_diag_small_kernel( const DiagTask* __restrict__ tasks, int* __restrict__ err ) {
  extern __shared__ double dsm[];
  double ( *a )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm );
  double ( *x )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + NB * ( NB + 1 ) );
  double* rsv = dsm + 2 * NB * ( NB + 1 );
  const DiagTask t = tasks[blockIdx.x];
  const int nb = t.nb, tid = threadIdx.x;
  for ( int e = tid; e < nb * nb; e += 256 ) { const int r = e % nb, c = e / nb; a[r][c] = ( r >= c ) ? t.A[r + ( int64_t )c * t.lda] : 0.0; x[r][c] = 0.0; }
  __syncthreads();
  const int r = tid & ( NB - 1 ), cg = tid >> 6;
  for ( int j = 0; j < nb; j++ ) {
    double d = a[j][j];
    if ( !( d > 0.0 ) ) { if ( tid == 0 ) atomicCAS( err, 0, t.col + j + 1 ); d = 1.0; }
    const double rs = rsqrt( d );
    if ( tid == 0 ) rsv[j] = rs;
    if ( r > j && r < nb ) {
      const double lr = a[r][j] * rs * rs;            /* l(r,j) / sqrt(d): the other factor a(c,j) is taken unscaled */
      for ( int c = j + 1 + cg; c <= r; c += 4 ) a[r][c] -= lr * a[c][j];
    }
    __syncthreads();
  }
  for ( int e = tid; e < nb * nb; e += 256 ) { const int rr = e % nb, c = e / nb; if ( rr >= c ) a[rr][c] = ( rr == c ) ? a[rr][c] * rsv[c] : a[rr][c] * rsv[c]; }   /* column c of L = a(:,c) * rs(c) */
  __syncthreads();
  {
    const int col = tid >> 2, part = tid & 3;
    if ( col < nb && part == 0 ) x[col][col] = rsv[col];
    __syncwarp();
    for ( int i = 1; i < nb; i++ ) {
      double s = 0.0;
      if ( col < nb && i > col ) for ( int q = col + part; q < i; q += 4 ) s += a[i][q] * x[q][col];
      s += __shfl_xor_sync( 0xffffffffu, s, 1 );
      s += __shfl_xor_sync( 0xffffffffu, s, 2 );
      if ( col < nb && i > col && part == 0 ) x[i][col] = -s * rsv[i];
      __syncwarp();
    }
  }
  __syncthreads();
  for ( int e = tid; e < nb * nb; e += 256 ) {
    const int rr = e % nb, c = e / nb;
    if ( rr >= c ) t.A[rr + ( int64_t )c * t.lda] = a[rr][c];
    t.inv[rr + c * nb] = x[rr][c];
  }
}

/* ---------------------------------------------------- small fronts, one CTA per front
 * The deep tree levels hold ~90 % of the supernodes and ~0.04 % of the flops (128^3: 119 330 fronts with k <= 32
 * pivots and f <= 160 rows).  Driving them through the four generic launches of a level (diagonal block, row update,
 * in-panel update, trailing update on 64x64 tensor tiles that are almost all padding) costs a few milliseconds per
 * level.  Here one CTA takes a whole front: the f x k panel goes to shared memory, is eliminated column by column
 * (right-looking over ALL rows, so the row update needs no inverse; one barrier per column), the top block is inverted
 * for the sweeps, and the contribution block receives  -L21 L21'  in 4x4 register blocks.  Cholesky only. */
constexpr int SF_KMAX = 32, SF_FMAX = 160, SF_LD = SF_KMAX + 1;
constexpr int SF_SMEM = ( ( SF_FMAX + 4 ) * SF_LD + SF_KMAX * SF_LD + SF_KMAX ) * 8;
__global__ void __launch_bounds__( 256 ) small_front_kernel( const int* __restrict__ sns, const SnMeta* __restrict__ sn, double* __restrict__ L, double* __restrict__ inv,
                                                             double* __restrict__ CB, int* __restrict__ err ) {
  extern __shared__ double dsm[];
  double ( *P )[SF_LD] = reinterpret_cast<double ( * )[SF_LD]>( dsm );                                    /* [f][k] */
  double ( *X )[SF_LD] = reinterpret_cast<double ( * )[SF_LD]>( dsm + ( SF_FMAX + 4 ) * SF_LD );          /* [k][k] */
  double* rsv = dsm + ( SF_FMAX + 4 ) * SF_LD + SF_KMAX * SF_LD;
  const SnMeta S = sn[sns[blockIdx.x]];
  const int k = S.k, f = S.f, u = f - k, tid = threadIdx.x;
  double* Lp = L + S.lptr;
  for ( int e = tid; e < f * k; e += 256 ) { const int r = e % f, c = e / f; P[r][c] = ( r >= c ) ? Lp[r + ( int64_t )c * S.ld] : 0.0; }
  for ( int e = tid; e < 4 * k; e += 256 ) P[f + e / k][e % k] = 0.0;          /* rows read by the last 4x4 blocks of the update */
  for ( int e = tid; e < k * k; e += 256 ) X[e / k][e % k] = 0.0;
  __syncthreads();
  for ( int j = 0; j < k; j++ ) {
    double d = P[j][j];
    if ( !( d > 0.0 ) ) { if ( tid == 0 ) atomicCAS( err, 0, S.c0 + j + 1 ); d = 1.0; }
    const double rs = rsqrt( d );
    if ( tid == 0 ) rsv[j] = rs;
    const double rs2 = rs * rs;
    const int nrow = f - 1 - j, ncol = k - 1 - j;
    for ( int e = tid; e < nrow * ncol; e += 256 ) {          /* column j stays unscaled in place: l(r,j) l(c,j) = a(r,j) a(c,j) / d */
      const int r = j + 1 + e % nrow, c = j + 1 + e / nrow;
      if ( r >= c ) P[r][c] -= P[r][j] * rs2 * P[c][j];
    }
    __syncthreads();
  }
  for ( int e = tid; e < f * k; e += 256 ) {
    const int r = e % f, c = e / f;
    if ( r >= c ) { const double l = P[r][c] * rsv[c]; P[r][c] = l; Lp[r + ( int64_t )c * S.ld] = l; }
  }
  __syncthreads();
  { /* inverse of the top block, four lanes per column */
    const int col = tid >> 2, part = tid & 3;
    if ( col < k && part == 0 ) X[col][col] = rsv[col];
    __syncwarp();
    for ( int i = 1; i < k; i++ ) {
      double sacc = 0.0;
      if ( col < k && i > col ) for ( int q = col + part; q < i; q += 4 ) sacc += P[i][q] * X[q][col];
      sacc += __shfl_xor_sync( 0xffffffffu, sacc, 1 );
      sacc += __shfl_xor_sync( 0xffffffffu, sacc, 2 );
      if ( col < k && i > col && part == 0 ) X[i][col] = -sacc * rsv[i];
      __syncwarp();
    }
  }
  __syncthreads();
  { double* ip = inv + S.invptr; for ( int e = tid; e < k * k; e += 256 ) { const int r = e % k, c = e / k; ip[r + c * k] = X[r][c]; } }
  if ( u == 0 ) return;
  /* contribution block (lower triangle): C(i,j) -= sum_c L21(i,c) L21(j,c) */
  double* cbp = CB + S.cbptr;
  const int64_t ldu = ( u + 1 ) & ~1;
  const int nb4 = ( u + 3 ) >> 2;
  for ( int b = tid; b < nb4 * nb4; b += 256 ) {
    const int bi = b % nb4, bj = b / nb4;
    if ( bi < bj ) continue;
    double acc[4][4];
#pragma unroll
    for ( int x = 0; x < 4; x++ )
#pragma unroll
      for ( int y = 0; y < 4; y++ ) acc[x][y] = 0.0;
    const int ri = k + 4 * bi, rj = k + 4 * bj;
    for ( int c = 0; c < k; c++ ) {
      double av[4], bv[4];
#pragma unroll
      for ( int x = 0; x < 4; x++ ) { av[x] = P[ri + x][c]; bv[x] = P[rj + x][c]; }
#pragma unroll
      for ( int x = 0; x < 4; x++ )
#pragma unroll
        for ( int y = 0; y < 4; y++ ) acc[x][y] += av[x] * bv[y];
    }
#pragma unroll
    for ( int y = 0; y < 4; y++ )
#pragma unroll
      for ( int x = 0; x < 4; x++ ) {
        const int i = 4 * bi + x, jj = 4 * bj + y;
        if ( i < u && jj < u && i >= jj ) cbp[i + ( int64_t )jj * ldu] -= acc[x][y];
      }
  }
}

/* ------------------------------------------------------ diagonal block: LU
 * LU of an nb x nb block with partial pivoting restricted to the block's own
 * rows.  piv[j] = block-local row swapped into position j.  Tiny pivots are
 * replaced by +-tiny (static perturbation) and counted.  Outputs: packed L\U
 * in place, inv = inverse of unit-lower L, inv2 = (inverse of U) transposed. */
__global__ void __launch_bounds__( 256 ) getrf_diag_kernel( const DiagTask* __restrict__ tasks, const double* __restrict__ tiny_p, int* __restrict__ nperturbed, int* __restrict__ err ) {
  const double tiny = *tiny_p;
  extern __shared__ double dsm[];
  double ( *a )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm );
  double ( *x )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + NB * ( NB + 1 ) );
  double ( *y )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + 2 * NB * ( NB + 1 ) );
  double* colv = dsm + 3 * NB * ( NB + 1 );
  double* rowv = colv + NB;
  __shared__ int prow;
  const DiagTask t = tasks[blockIdx.x];
  const int nb = t.nb, tid = threadIdx.x;
  for ( int e = tid; e < nb * nb; e += 256 ) { int r = e % nb, c = e / nb; a[r][c] = ( r >= c ) ? t.A[r + ( int64_t )c * t.lda] : t.AU[c + ( int64_t )r * t.lda]; x[r][c] = 0.0; y[r][c] = 0.0; }
  __syncthreads();
  for ( int j = 0; j < nb; j++ ) {
    if ( tid < 32 ) { /* arg-max of |a[r][j]|, r>=j, first maximal row wins */
      double best = -1.0; int bi = j;
      for ( int r = j + tid; r < nb; r += 32 ) { double v = fabs( a[r][j] ); if ( v > best ) { best = v; bi = r; } }
      for ( int o = 16; o; o >>= 1 ) {
        double ob = __shfl_xor_sync( 0xffffffffu, best, o ); int oi = __shfl_xor_sync( 0xffffffffu, bi, o );
        if ( ob > best || ( ob == best && oi < bi ) ) { best = ob; bi = oi; }
      }
      if ( tid == 0 ) { prow = bi; t.piv[j] = bi; }
    }
    __syncthreads();
    const int p = prow;
    if ( p != j && tid < nb ) { double tmp = a[j][tid]; a[j][tid] = a[p][tid]; a[p][tid] = tmp; }
    __syncthreads();
    double d = a[j][j];
    if ( d != d && tid == 0 ) atomicCAS( err, 0, -( t.col + j + 1 ) );      /* NaN pivot column: the factorization cannot be repaired by a perturbation */
    if ( !( fabs( d ) > tiny ) ) { d = ( d < 0.0 ) ? -tiny : tiny; if ( tid == 0 ) { a[j][j] = d; atomicAdd( nperturbed, 1 ); } }
    if ( tid < nb ) { colv[tid] = tid > j ? a[tid][j] / d : 0.0; rowv[tid] = tid > j ? a[j][tid] : 0.0; }
    __syncthreads();
    if ( tid < nb && tid > j ) a[tid][j] = colv[tid];
    const int r = tid & 63, cg = tid >> 6;
    if ( r < nb && r > j ) {
      const double lr = colv[r];
      for ( int c = j + 1 + cg; c < nb; c += 4 ) a[r][c] -= lr * rowv[c];
    }
    { /* row j of x = inverse of the unit lower L and column j of y = inverse of the upper U, in the same step: row j of L (its interchange is done),
       * column j of U and the pivot are final and no later step touches them, so this chain of 64 dependent steps runs alongside the
       * elimination instead of behind it (as in potrf_diag).  Four lanes per column `col`; each group only reads its own earlier results. */
      const int col = tid >> 2, part = tid & 3;
      double s = 0.0, u = 0.0;
      if ( col < nb && j > col ) for ( int q = col + part; q < j; q += 4 ) { s += a[j][q] * x[q][col]; u += a[q][j] * y[col][q]; }
      s += __shfl_xor_sync( 0xffffffffu, s, 1 ); s += __shfl_xor_sync( 0xffffffffu, s, 2 );
      u += __shfl_xor_sync( 0xffffffffu, u, 1 ); u += __shfl_xor_sync( 0xffffffffu, u, 2 );
      if ( col < nb && part == 0 ) { if ( j == col ) { x[col][col] = 1.0; y[col][col] = 1.0 / d; } else if ( j > col ) { x[j][col] = -s; y[col][j] = -u / d; } }
    }
    __syncthreads();
  }
  __shared__ int org[NB];            /* org[q]: the column of the untouched tile that the interchanges, applied in order, bring to position q */
  if ( tid == 0 ) {
    for ( int q = 0; q < nb; q++ ) org[q] = q;
    for ( int j = 0; j < nb; j++ ) { const int pj = t.piv[j]; if ( pj != j ) { const int sw = org[j]; org[j] = org[pj]; org[pj] = sw; } }
  }
  __syncthreads();
  for ( int e = tid; e < nb * nb; e += 256 ) {
    int r = e % nb, c = e / nb;
    if ( t.inv3 ) t.inv3[r + org[c] * nb] = x[r][c];               /* inv(L) P: column q of inv(L) goes where the interchanges took it from */
    if ( r > c ) t.A[r + ( int64_t )c * t.lda] = a[r][c];          /* unit-lower L, strict part   */
    else t.AU[c + ( int64_t )r * t.lda] = a[r][c];                 /* U (incl. diagonal), as U'   */
    if ( r == c ) t.A[r + ( int64_t )c * t.lda] = 1.0;
    t.inv[r + c * nb] = x[r][c];     /* inv(L), lower                                  */
    t.inv2[r + c * nb] = y[c][r];    /* inv(U)' : y holds inv(U) (upper), store transpose */
  }
}

/* ------------------------------------------------------ diagonal block: LDL'
 * Symmetric indefinite nb x nb block (nb <= 64, lower triangle in the L panel):  P A P' = L D L'  with Bunch-Kaufman partial
 * pivoting (1x1 and 2x2 pivots, alpha = (1+sqrt(17))/8), the interchanges restricted to the block's own rows -- the same rule the
 * LU path follows; a column whose candidates are all below `tiny` is statically perturbed and counted.
 * Outputs (the conventions of the LU path, so that the row update, the solve-block inverses and the sweeps are shared):
 *   inv  = T  = inv(L) P        -> rows below:  W21 = A21 T'   (W = L D, kept in the second panel; it plays the part of U')
 *   inv2 = T2 = inv(D) inv(L) P -> rows below:  L21 = A21 T2'
 * Both are full nb x nb matrices (the interchanges are folded in as a column permutation).  The panels' diagonal blocks receive
 * L (unit lower, pivot order) and D (block diagonal, lower part); nothing reads them again. */
constexpr int LDLT_SMEM = ( 3 * NB * ( NB + 1 ) + 2 * NB ) * 8 + 3 * NB * 4;
__global__ void __launch_bounds__( 256 ) ldlt_diag_kernel( const DiagTask* __restrict__ tasks, const double* __restrict__ tiny_p, int* __restrict__ nperturbed, int* __restrict__ err ) {
  extern __shared__ double dsm[];
  double ( *a )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm );                          /* full symmetric active matrix; columns < j hold L */
  double ( *x )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + NB * ( NB + 1 ) );          /* inv(L)                                         */
  double ( *y )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + 2 * NB * ( NB + 1 ) );      /* inv(D) inv(L)                                  */
  double* c1 = dsm + 3 * NB * ( NB + 1 );       /* column j   of the active matrix before elimination */
  double* c2 = c1 + NB;                         /* column j+1 (2x2 pivots)                            */
  int* sig = reinterpret_cast<int*>( c2 + NB ); /* position -> original row of the block              */
  int* two = sig + NB;                          /* two[j] = 1: (j, j+1) is a 2x2 pivot                */
  __shared__ int s_kp, s_step; __shared__ double s_d11, s_d21, s_d22;
  const DiagTask t = tasks[blockIdx.x];
  const int nb = t.nb, tid = threadIdx.x, lane = tid & 31;
  const double tiny = *tiny_p;
  const double alpha = 0.6403882032022076;
  for ( int e = tid; e < nb * nb; e += 256 ) { const int r = e % nb, c = e / nb; const double v = r >= c ? t.A[r + ( int64_t )c * t.lda] : t.A[c + ( int64_t )r * t.lda]; a[r][c] = v; x[r][c] = 0.0; y[r][c] = 0.0; }
  if ( tid < NB ) { sig[tid] = tid; two[tid] = 0; }
  __syncthreads();
  int j = 0;
  while ( j < nb ) {
    if ( tid < 32 ) {   /* ---- pivot choice (warp 0) */
      double colmax = 0.0; int imax = j;
      for ( int i = j + 1 + lane; i < nb; i += 32 ) { const double v = fabs( a[i][j] ); if ( v > colmax ) { colmax = v; imax = i; } }
      for ( int o = 16; o; o >>= 1 ) { const double ov = __shfl_xor_sync( 0xffffffffu, colmax, o ); const int oi = __shfl_xor_sync( 0xffffffffu, imax, o ); if ( ov > colmax || ( ov == colmax && ov > 0.0 && oi < imax ) ) { colmax = ov; imax = oi; } }
      const double absakk = fabs( a[j][j] );
      if ( lane == 0 && ( absakk != absakk || colmax != colmax ) ) atomicCAS( err, 0, -( t.col + j + 1 ) );      /* NaN in the pivot column */
      int kp = j, step = 1;
      if ( !( fmax( absakk, colmax ) > tiny ) ) {                  /* nothing usable in this column: static perturbation (also catches NaN) */
        if ( lane == 0 ) { a[j][j] = ( a[j][j] < 0.0 ) ? -tiny : tiny; atomicAdd( nperturbed, 1 ); }
      } else if ( absakk < alpha * colmax ) {
        double rowmax = 0.0;
        for ( int q = j + lane; q < nb; q += 32 ) if ( q != imax ) rowmax = fmax( rowmax, fabs( a[imax][q] ) );
        for ( int o = 16; o; o >>= 1 ) rowmax = fmax( rowmax, __shfl_xor_sync( 0xffffffffu, rowmax, o ) );
        if ( absakk * rowmax >= alpha * colmax * colmax ) { kp = j; }
        else if ( fabs( a[imax][imax] ) >= alpha * rowmax ) { kp = imax; }
        else { kp = imax; step = 2; }
      }
      if ( lane == 0 ) { s_kp = kp; s_step = step; }
    }
    __syncthreads();
    const int step = s_step, kk = j + step - 1, kp = s_kp;
    if ( kp != kk ) {   /* ---- symmetric interchange kk <-> kp: rows over all columns, then columns over all rows */
      if ( tid < nb ) { const double tmp = a[kk][tid]; a[kk][tid] = a[kp][tid]; a[kp][tid] = tmp; }
      __syncthreads();
      if ( tid < nb && tid >= j ) { const double tmp = a[tid][kk]; a[tid][kk] = a[tid][kp]; a[tid][kp] = tmp; }      /* columns of the L part (< j) belong to other pivots: rows only */
      if ( tid == 0 ) { const int s = sig[kk]; sig[kk] = sig[kp]; sig[kp] = s; }
      __syncthreads();
    }
    if ( step == 1 ) {
      if ( tid == 0 ) s_d11 = a[j][j];
      if ( tid < nb ) c1[tid] = tid > j ? a[tid][j] : 0.0;
      __syncthreads();
      const double d = s_d11, rd = 1.0 / d;
      for ( int e = tid; e < ( nb - j - 1 ) * ( nb - j - 1 ); e += 256 ) {
        const int r = j + 1 + e % ( nb - j - 1 ), c = j + 1 + e / ( nb - j - 1 );
        a[r][c] -= c1[r] * rd * c1[c];
      }
      if ( tid < nb && tid > j ) a[tid][j] = c1[tid] * rd;
      __syncthreads();
      j += 1;
    } else {
      if ( tid == 0 ) { s_d11 = a[j][j]; s_d21 = a[j + 1][j]; s_d22 = a[j + 1][j + 1]; two[j] = 1; }
      if ( tid < nb ) { c1[tid] = tid > j + 1 ? a[tid][j] : 0.0; c2[tid] = tid > j + 1 ? a[tid][j + 1] : 0.0; }
      __syncthreads();
      /* inverse of the 2x2 pivot, scaled as in LAPACK dsytf2 for accuracy */
      const double d21 = s_d21, e11 = s_d11 / d21, e22 = s_d22 / d21, tt = 1.0 / ( e11 * e22 - 1.0 ), sc = tt / d21;
      for ( int e = tid; e < ( nb - j - 2 ) * ( nb - j - 2 ); e += 256 ) {
        const int r = j + 2 + e % ( nb - j - 2 ), c = j + 2 + e / ( nb - j - 2 );
        const double l1 = sc * ( e22 * c1[r] - c2[r] ), l2 = sc * ( e11 * c2[r] - c1[r] );      /* row r of L, columns j and j+1 */
        a[r][c] -= l1 * c1[c] + l2 * c2[c];
      }
      if ( tid < nb && tid > j + 1 ) { a[tid][j] = sc * ( e22 * c1[tid] - c2[tid] ); a[tid][j + 1] = sc * ( e11 * c2[tid] - c1[tid] ); }
      __syncthreads();
      j += 2;
    }
  }
  /* x = inverse of the unit lower L (a 2x2 pivot has l(j+1,j) = 0: its entry a[j+1][j] holds d21) */
  {
    const int col = tid >> 2, part = tid & 3;
    if ( col < nb && part == 0 ) x[col][col] = 1.0;
    __syncwarp();
    for ( int i = 1; i < nb; i++ ) {
      double s = 0.0;
      if ( col < nb && i > col ) for ( int q = col + part; q < i; q += 4 ) { const double l = ( q + 1 == i && two[q] ) ? 0.0 : a[i][q]; s += l * x[q][col]; }
      s += __shfl_xor_sync( 0xffffffffu, s, 1 ); s += __shfl_xor_sync( 0xffffffffu, s, 2 );
      if ( col < nb && i > col && part == 0 ) x[i][col] = -s;
      __syncwarp();
    }
  }
  __syncthreads();
  /* y = inv(D) x, row by row (2x2 pivots couple two rows) */
  for ( int e = tid; e < nb * nb; e += 256 ) {
    const int r = e % nb, c = e / nb;
    double v;
    if ( two[r] ) { const double d21 = a[r + 1][r], e11 = a[r][r] / d21, e22 = a[r + 1][r + 1] / d21, sc = 1.0 / ( ( e11 * e22 - 1.0 ) * d21 ); v = sc * ( e22 * x[r][c] - x[r + 1][c] ); }
    else if ( r > 0 && two[r - 1] ) { const double d21 = a[r][r - 1], e11 = a[r - 1][r - 1] / d21, e22 = a[r][r] / d21, sc = 1.0 / ( ( e11 * e22 - 1.0 ) * d21 ); v = sc * ( e11 * x[r][c] - x[r - 1][c] ); }
    else v = x[r][c] / a[r][r];
    y[r][c] = v;
  }
  __syncthreads();
  for ( int e = tid; e < nb * nb; e += 256 ) {
    const int r = e % nb, c = e / nb;
    t.inv[r + sig[c] * nb] = x[r][c];       /* T  = inv(L) P : column c of inv(L) lands in column sig[c] */
    t.inv2[r + sig[c] * nb] = y[r][c];      /* T2 = inv(D) inv(L) P */
    if ( r >= c ) {
      const bool sub2 = r == c + 1 && two[c];
      t.A[r + ( int64_t )c * t.lda] = r == c ? 1.0 : ( sub2 ? 0.0 : a[r][c] );
      t.AU[r + ( int64_t )c * t.lda] = r == c ? a[r][r] : ( sub2 ? a[r][c] : 0.0 );
    }
  }
}

/* -------------------------------------------------------------- trsm rows
 * X = B * T' for a tile of 128 rows; T (nb x nb) is lower triangular, so column
 * c of X needs q <= c only.  B is overwritten.  4x8 register tile per thread. */
__global__ void __launch_bounds__( 256 ) trsm_rows_kernel( const TrsmTask* __restrict__ tasks, const RowTile* __restrict__ tiles ) {
  extern __shared__ double sm[];
  const RowTile rt = tiles[blockIdx.x];
  const TrsmTask t = tasks[rt.task];
  const int nb = t.nb;
  double* Ts = sm;                 /* [q][c]  stride NB   : T[c,q]  */
  double* Bs = sm + NB * NB;       /* [q][r]  stride 128            */
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = min( 128, t.m - rt.r0 );
  double* Bg = t.B + rt.r0;
  double* Og = ( t.out ? t.out : t.B ) + rt.r0;
  for ( int e = tid; e < NB * NB; e += 256 ) { int c = e % NB, q = e / NB; Ts[q * NB + c] = ( c < nb && q < nb ) ? t.T[c + q * nb] : 0.0; }
  for ( int e = tid; e < nb * 128; e += 256 ) {
    int r = e & 127, q = e >> 7;
    Bs[q * 128 + r] = ( r < rows ) ? Bg[r + ( int64_t )q * t.ldb] : 0.0;
  }
  __syncthreads();
  if ( t.piv ) { /* apply the block's row interchanges to the columns of this tile, in order */
    for ( int j = 0; j < nb; j++ ) {
      const int p = t.piv[j];
      if ( p != j && tid < 128 ) { double tmp = Bs[j * 128 + tid]; Bs[j * 128 + tid] = Bs[p * 128 + tid]; Bs[p * 128 + tid] = tmp; }
      __syncthreads();
    }
  }
  double acc[4][8];
#pragma unroll
  for ( int i = 0; i < 4; i++ )
#pragma unroll
    for ( int j = 0; j < 8; j++ ) acc[i][j] = 0.0;
  const int c0 = warp * 8;
  if ( c0 < nb ) {
    const int qend = t.full ? nb : min( nb, c0 + 8 );
    for ( int q = 0; q < qend; q++ ) {
      double bv[4], tv[8];
#pragma unroll
      for ( int i = 0; i < 4; i++ ) bv[i] = Bs[q * 128 + lane + 32 * i];
#pragma unroll
      for ( int j = 0; j < 8; j++ ) tv[j] = Ts[q * NB + c0 + j];
#pragma unroll
      for ( int i = 0; i < 4; i++ )
#pragma unroll
        for ( int j = 0; j < 8; j++ ) acc[i][j] += bv[i] * tv[j];
    }
#pragma unroll
    for ( int j = 0; j < 8; j++ ) {
      const int c = c0 + j;
      if ( c < nb ) {
#pragma unroll
        for ( int i = 0; i < 4; i++ ) { const int r = lane + 32 * i; if ( r < rows ) Og[r + ( int64_t )c * t.ldb] = acc[i][j]; }
      }
    }
  }
}

/* ------------------------------------------- DMMA GEMM NT, bulk-copy pipeline
 * C[M x N] -= A[M x K] * B[N x K]'   (all column-major; mode 1: C = A B').
 * CTA tile BM x BN, K step 16, STAGES-deep pipeline.  Shared tiles are stored k-major with a row stride of BM+4 doubles so that the
 * m8n8k4 fragment reads (lane>>2 -> m, lane&3 -> k) hit 32 distinct banks; each warp owns a WM x WN sub-tile = (WM/8)x(WN/8)
 * DMMA.8x8x4 accumulators in registers.  The operand tiles are brought in by the async proxy: one elected producer warp issues, per
 * stage, one cp.async.bulk (SASS UBLKCP) per k-column of A and of B -- each column of a panel tile is a contiguous run of <=66
 * doubles -- and the bytes are counted on an mbarrier.  The 8 consumer warps only ever wait on "full[stage]" and signal
 * "empty[stage]": there is no block-wide barrier in the K loop, so a warp that gets ahead keeps the FP64 tensor pipe busy instead of
 * waiting for the slowest one (ncu on a cp.async + __syncthreads version: 13 % of all warp samples sat at the per-stage barrier).
 * A tile whose first row sits at an odd address is fetched one element early (16-byte aligned bulk copies) and every fragment read
 * is shifted.  Requirements (true for every factor task): lda, ldb even; panel columns padded to their even ld.
 * Tried and measured slower (profiles/r1_gemm_bulk_variants.txt, r1_gemm_variants_*.txt): 128x128 and 128x64 tiles at small K, 6 stages,
 * a persistent-CTA variant with a static tile order (28.3 against 32.9 TFLOP/s at K = 512, 752 against 673 ms per 128^3 factorization). */
__device__ __forceinline__ void mbar_init( uint64_t* bar, int count ) {
  asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"( ( unsigned )__cvta_generic_to_shared( bar ) ), "r"( count ) );
}
__device__ __forceinline__ void mbar_expect_tx( uint64_t* bar, int bytes ) {
  asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"( ( unsigned )__cvta_generic_to_shared( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void mbar_arrive( uint64_t* bar ) {
  asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"( ( unsigned )__cvta_generic_to_shared( bar ) ) : "memory" );
}
__device__ __forceinline__ void mbar_wait( uint64_t* bar, int parity ) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"( ( unsigned )__cvta_generic_to_shared( bar ) ), "r"( parity ) : "memory" );
}
__device__ __forceinline__ void bulk_g2s( void* smem, const void* g, int bytes, uint64_t* bar ) {
  asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"( ( unsigned )__cvta_generic_to_shared( smem ) ),
                "l"( g ), "r"( bytes ), "r"( ( unsigned )__cvta_generic_to_shared( bar ) ) : "memory" );
}

template <int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__( ( BM / WM ) * ( BN / WN ) * 32 + 32 ) gemm_nt_dmma_bulk_kernel( const GemmTask* __restrict__ tasks, const Tile* __restrict__ tiles,
                                                                                               const double* __restrict__ zeros ) {
  constexpr int BK = 16;
  constexpr int NCW = ( BM / WM ) * ( BN / WN );      /* consumer warps */
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int MT = WM / 8, NTL = WN / 8;
  extern __shared__ double sm[];
  double* As = sm;
  double* Bs = sm + STAGES * BK * LDA;
  uint64_t* full = reinterpret_cast<uint64_t*>( sm + STAGES * BK * ( LDA + LDB ) );
  uint64_t* empty = full + STAGES;
  const Tile tl = tiles[blockIdx.x];
  const GemmTask t = tasks[tl.task];
  const int m0 = tl.tm * BM, n0 = tl.tn * BN;
  const int mrem = t.M - m0, nrem = t.N - n0;
  const double* Ag = t.A + m0;
  const double* Bg = t.B + n0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sa = ( int )( ( ( uintptr_t )Ag >> 3 ) & 1 ), sb = ( int )( ( ( uintptr_t )Bg >> 3 ) & 1 );
  const int K = t.K, nk = ( K + BK - 1 ) / BK;
  if ( tid == 0 ) {
    for ( int s = 0; s < STAGES; s++ ) { mbar_init( &full[s], 1 ); mbar_init( &empty[s], NCW ); }
    asm volatile( "fence.mbarrier_init.release.cluster;\n" ::: "memory" );
    asm volatile( "fence.proxy.async.shared::cta;\n" ::: "memory" );
  }
  __syncthreads();

  if ( warp == NCW ) {
    /* ---------------- producer warp: lane l issues the copy of k-column l%16 of A (l<16) or B (l>=16),
     * so a whole stage (32 columns) goes out with one warp instruction */
    static_assert( BK == 16, "one lane per k-column of A and of B" );
    const int ra = ( ( min( BM, mrem ) + sa + 1 ) & ~1 ) * 8, rb = ( ( min( BN, nrem ) + sb + 1 ) & ~1 ) * 8;   /* bytes per k-column */
    const int q = lane & 15; const bool isb = lane >= 16;
    const double* g = isb ? Bg - sb : Ag - sa;
    const int64_t ldg = isb ? t.ldb : t.lda;
    const int bytes = isb ? rb : ra;
    for ( int kt = 0; kt < nk; kt++ ) {
      const int slot = kt % STAGES, use = kt / STAGES;
      if ( lane == 0 ) {
        if ( use > 0 ) mbar_wait( &empty[slot], ( use - 1 ) & 1 );
        mbar_expect_tx( &full[slot], BK * ( ra + rb ) );
      }
      __syncwarp();
      double* dst = isb ? Bs + slot * BK * LDB + q * LDB : As + slot * BK * LDA + q * LDA;
      const int kq = kt * BK + q;
      /* K tail: feed zeros so that the partial k-step adds nothing */
      bulk_g2s( dst, kq < K ? ( const void* )( g + ( int64_t )kq * ldg ) : ( const void* )zeros, bytes, &full[slot] );
    }
    return;
  }

  /* ---------------- consumers */
  const int wm0 = ( warp % ( BM / WM ) ) * WM, wn0 = ( warp / ( BM / WM ) ) * WN;
  double acc[MT][NTL][2];
#pragma unroll
  for ( int i = 0; i < MT; i++ )
#pragma unroll
    for ( int j = 0; j < NTL; j++ ) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const int fr = lane >> 2, fk = lane & 3;
  for ( int kt = 0; kt < nk; kt++ ) {
    const int slot = kt % STAGES;
    mbar_wait( &full[slot], ( kt / STAGES ) & 1 );
    const double* as = As + slot * BK * LDA + wm0 + fr + sa;
    const double* bs = Bs + slot * BK * LDB + wn0 + fr + sb;
#pragma unroll
    for ( int kk = 0; kk < BK; kk += 4 ) {
      double af[MT], bf[NTL];
#pragma unroll
      for ( int i = 0; i < MT; i++ ) af[i] = as[( kk + fk ) * LDA + i * 8];
#pragma unroll
      for ( int j = 0; j < NTL; j++ ) bf[j] = bs[( kk + fk ) * LDB + j * 8];
#pragma unroll
      for ( int i = 0; i < MT; i++ )
#pragma unroll
        for ( int j = 0; j < NTL; j++ ) dmma884( acc[i][j][0], acc[i][j][1], af[i], bf[j] );
    }
    __syncwarp();
    if ( lane == 0 ) mbar_arrive( &empty[slot] );     /* this warp is done reading the slot */
  }
  double* Cg = t.C;
#pragma unroll
  for ( int j = 0; j < NTL; j++ ) {
#pragma unroll
    for ( int h = 0; h < 2; h++ ) {
      const int c = n0 + wn0 + j * 8 + 2 * fk + h;
      if ( c < t.N ) {
        double* cc = Cg + ( int64_t )c * t.ldc;
#pragma unroll
        for ( int i = 0; i < MT; i++ ) {
          const int r = m0 + wm0 + i * 8 + fr;
          if ( r < t.M && ( !t.lower || r + t.diag >= c ) ) { if ( t.mode ) cc[r] = acc[i][j][h]; else cc[r] -= acc[i][j][h]; }
        }
      }
    }
  }
}

/* -------------------------------------------------------------- solves
 * Level-set sweeps over supernodes.  A supernode's k pivot columns are cut into SOLVE BLOCKS of sbw columns (256, or SB_WIDE
 * for the big fronts at the top of the tree); for every block the factorization leaves behind the explicit inverse of its
 * triangular diagonal block (build_sinv_kernel), so a sweep is a sequence of dense products per block b:
 *     forward :  z_b = X_b y_b ;            y[rows below b] -= L[rows below b, b] z_b
 *     backward:  x_b = X_b' (z_b - ...) ;   z[cols before b] -= L[b, cols before]' x_b     (update rows first, x gathered from the ancestors)
 * Every product is an SvTask "dest (-)= M v".  Two implementations execute the same task lists:
 *     p <= 4 right-hand sides : FMA kernels below (sv_gemv_kernel: outputs = rows of M; sv_gemvT_kernel: outputs = columns of M) -- HBM bound
 *     p >= 5                  : the FP64 tensor pipe, gemm_solve_dmma_kernel, with the right-hand sides as the M dimension of a DMMA tile
 * Both follow ONE canonical arithmetic, so that a column's bits do not depend on how many columns travel with it:
 *     the reduction index k is cut into chunks of KC = 64; inside a chunk the sum is an FMA chain from 0 in ascending k
 *     (measured: DMMA.8x8x4 is bit-identical to that chain, scripts/probe/dmma_order.cu); the chunk sums p_0, p_1, ... are then
 *     subtracted one after the other:  dest = ((dest - p_0) - p_1) - ...   (mode 1: dest = -(((0 - p_0) - p_1) - ...)).
 * Vectors are stored with the right-hand sides contiguous: element (row, q) at row * pitch + q. */
constexpr int KC = 64;          /* canonical chunk of the reduction index */
constexpr int SB = 256;         /* solve block width of ordinary supernodes */
constexpr int SB_WIDE = 1024;   /* ... of supernodes with at least SB_WIDE_MIN pivot columns */
constexpr int SB_WIDE_MIN = 2048;

struct SvTask {
  const double* M; int ld;      /* slab of the factor (or of a block inverse), column-major                                        */
  int K, N;                     /* reduction length, number of outputs.  gemv: M is N x K, outputs = rows; gemvT: M is K x N        */
  int tri;                      /* gemv: 0 dense, 1 lower triangular (output r uses k <= r), 2 upper triangular (k >= r)            */
  int mode;                     /* 0: dest -= M v ; 1: dest = M v                                                                   */
  int vsel, dsel;               /* vector bases: 0 Y, 1 Z, 2 W[0], 3 W[1]                                                           */
  int64_t voff, doff;           /* first row of v (K rows) and of dest (N rows) inside their base                                   */
};
struct SvTile { int task; int o0; };
struct SvBases { double* b[4]; };

__global__ void permute_in_kernel( const double* __restrict__ b, const int* __restrict__ perm, int n, int p, int pitch, double* __restrict__ y ) {
  int64_t e = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = ( int64_t )n * pitch, stride = ( int64_t )gridDim.x * blockDim.x;
  for ( ; e < tot; e += stride ) { const int q = ( int )( e % pitch ); const int64_t k = e / pitch; y[e] = q < p ? b[perm[k] + ( int64_t )q * n] : 0.0; }
}
__global__ void permute_out_kernel( const double* __restrict__ y, const int* __restrict__ perm, int n, int p, int pitch, double* __restrict__ x ) {
  int64_t e = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = ( int64_t )n * p, stride = ( int64_t )gridDim.x * blockDim.x;
  for ( ; e < tot; e += stride ) { const int k = ( int )( e % n ); const int64_t q = e / n; x[perm[k] + q * n] = y[( int64_t )k * pitch + q]; }
}

/* forward: add the children's update vectors into the parent's pivots (y) and update rows (w).
 * One CTA per (parent, range of GATHER_ROWS destination rows): the rows of a child that land in the
 * range are a contiguous run (relidx is increasing); children are visited in a fixed order. */
constexpr int GATHER_ROWS = 2048;
struct GatherTask { int parent; int lo; int hi; };
__global__ void __launch_bounds__( 256 ) fwd_gather_kernel( const GatherTask* __restrict__ tasks, const SnMeta* __restrict__ sn, const int* __restrict__ child_list,
                                                            const int* __restrict__ relidx, double* __restrict__ y, int pitch, int tqs,
                                                            double* __restrict__ w_parent, const double* __restrict__ w_child ) {
  const GatherTask t = tasks[blockIdx.x];
  const SnMeta P = sn[t.parent];
  /* thread = (row slot, right-hand-side slot): 2^tqs consecutive threads walk the right-hand sides of one row (coalesced) */
  const int TQ = 1 << tqs, tq = threadIdx.x & ( TQ - 1 ), ti = threadIdx.x >> tqs, tr = blockDim.x >> tqs;
  for ( int ci = P.child_begin; ci < P.child_end; ci++ ) {
    const SnMeta C = sn[child_list[ci]];
    const int uc = C.f - C.k;
    const int* rel = relidx + C.rowptr + C.k;
    int a = 0, b = uc;
    if ( t.lo > 0 || t.hi < P.f ) {
      { int x = 0, z = uc; while ( x < z ) { const int m = ( x + z ) >> 1; if ( rel[m] < t.lo ) x = m + 1; else z = m; } a = x; }
      { int x = a, z = uc; while ( x < z ) { const int m = ( x + z ) >> 1; if ( rel[m] < t.hi ) x = m + 1; else z = m; } b = x; }
    }
    for ( int i = a + ti; i < b; i += tr ) {
      const int d = rel[i];
      const double* src = w_child + ( C.wptr + i ) * pitch;
      double* dst = d < P.k ? y + ( int64_t )( P.c0 + d ) * pitch : w_parent + ( P.wptr + ( d - P.k ) ) * pitch;
      for ( int q = tq; q < pitch; q += TQ ) dst[q] += src[q];
    }
    __syncthreads();
  }
}
/* backward: xg[update row i of s] = x[global row], for every supernode of a level (x of the ancestors is final) */
__global__ void __launch_bounds__( 256 ) bwd_gather_kernel( const GatherTask* __restrict__ tasks, const SnMeta* __restrict__ sn, const int* __restrict__ rowidx,
                                                            const double* __restrict__ x, int pitch, int tqs, double* __restrict__ xg ) {
  const GatherTask t = tasks[blockIdx.x];
  const SnMeta S = sn[t.parent];
  const int* ri = rowidx + S.rowptr + S.k;
  const int TQ = 1 << tqs, tq = threadIdx.x & ( TQ - 1 ), ti = threadIdx.x >> tqs, tr = blockDim.x >> tqs;
  for ( int i = t.lo + ti; i < t.hi; i += tr ) {
    const double* src = x + ( int64_t )ri[i] * pitch;
    double* dst = xg + ( S.wptr + i ) * pitch;
    for ( int q = tq; q < pitch; q += TQ ) dst[q] = src[q];
  }
}

/* dest (-)= M v, outputs = rows of M (M is N x K column-major, K <= SB_WIDE): GV_ROWS = 32 output rows per CTA, thread = (row, one of 8
 * chunk lanes).  Chunk lane g forms the chains of chunks g, g+8, ... -- 16 independent coalesced loads in flight per thread (a warp runs
 * down 32 rows of a column of M), then the 16 FMAs of the chain; the chunk sums are parked in shared memory and subtracted in chunk
 * order by lane 0. */
constexpr int GV_ROWS = 32;
constexpr int SVG_MAXCH = SB_WIDE / KC;
template <int PCT> constexpr int sv_gemv_smem() { return ( SB_WIDE + SVG_MAXCH * GV_ROWS ) * PCT * 8; }
/* keeps the 16 loads of a batch together: every value must have arrived before the first FMA of the chain issues */
__device__ __forceinline__ void pin8( double& a, double& b, double& c, double& d, double& e, double& f, double& g, double& h ) {
  asm volatile( "" : "+d"( a ), "+d"( b ), "+d"( c ), "+d"( d ), "+d"( e ), "+d"( f ), "+d"( g ), "+d"( h ) );
}
template <int PCT>
__global__ void __launch_bounds__( 256 ) sv_gemv_kernel( const SvTask* __restrict__ tasks, const SvTile* __restrict__ tiles, SvBases B, int p, int pitch ) {
  extern __shared__ double dsm[];
  double ( *vs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm );                                          /* [K][PCT]            */
  double ( *part )[GV_ROWS][PCT] = reinterpret_cast<double ( * )[GV_ROWS][PCT]>( dsm + SB_WIDE * PCT );       /* [chunk][row][PCT]   */
  const SvTile tl = tiles[blockIdx.x];
  const SvTask t = tasks[tl.task];
  const int pc0 = blockIdx.y * PCT, pcn = min( PCT, p - pc0 );
  const int tid = threadIdx.x, r = tid & ( GV_ROWS - 1 ), g = tid >> 5;
  const int row = tl.o0 + r;
  const bool live = row < t.N;
  const double* v = B.b[t.vsel] + t.voff * pitch + pc0;
  double* dest = B.b[t.dsel] + ( t.doff + row ) * pitch + pc0;
  /* reduction window of this tile in whole chunks: lower triangular -> chunks up to the one holding the tile's last row, upper -> from the one holding its first */
  const int k_lo = t.tri == 2 ? ( tl.o0 & ~( KC - 1 ) ) : 0, k_hi = t.tri == 1 ? min( t.K, ( ( tl.o0 + GV_ROWS - 1 ) | ( KC - 1 ) ) + 1 ) : t.K;
  for ( int e = tid; e < ( k_hi - k_lo ) * PCT; e += 256 ) { const int kk = e / PCT, q = e % PCT; vs[kk][q] = q < pcn ? v[( int64_t )( k_lo + kk ) * pitch + q] : 0.0; }
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( t.mode == 0 && live && g == 0 && q < pcn ) ? dest[q] : 0.0;
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

/* dest -= M' v, outputs = columns of M (M is K x N column-major, the reduction runs down the contiguous rows): 64 output columns per
 * CTA of 4 warps, 16 per warp.  Per chunk of 64 rows a warp brings its 64 x 16 slab in with coalesced loads (the next chunk is
 * already in flight in registers), parks it in shared memory, and lanes 0..15 run one canonical chain each -- the FMA work is a
 * fraction of a percent of the pipe, the kernel streams. */
constexpr int GT_COLS = 64;
constexpr int GT_XS = 512;      /* rows of v staged per pass */
template <int PCT> constexpr int sv_gemvT_smem() { return ( GT_XS * PCT + 4 * 16 * ( KC + 1 ) ) * 8; }
template <int PCT>
__global__ void __launch_bounds__( 128 ) sv_gemvT_kernel( const SvTask* __restrict__ tasks, const SvTile* __restrict__ tiles, SvBases B, int p, int pitch ) {
  extern __shared__ double dsm[];
  double ( *xs )[PCT] = reinterpret_cast<double ( * )[PCT]>( dsm );                                                   /* [GT_XS][PCT]      */
  double ( *slab )[16][KC + 1] = reinterpret_cast<double ( * )[16][KC + 1]>( dsm + GT_XS * PCT );                      /* [warp][col][row]  */
  const SvTile tl = tiles[blockIdx.x];
  const SvTask t = tasks[tl.task];
  const int pc0 = blockIdx.y * PCT, pcn = min( PCT, p - pc0 );
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbase = tl.o0 + warp * 16;
  const double* v = B.b[t.vsel] + t.voff * pitch + pc0;
  const int mycol = cbase + ( lane & 15 );
  double* dest = B.b[t.dsel] + ( t.doff + mycol ) * pitch + pc0;
  const bool owner = lane < 16 && mycol < t.N;
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( owner && q < pcn ) ? dest[q] : 0.0;
  const int ncol = max( 0, min( 16, t.N - cbase ) );
  const int64_t ldm = t.ld;
  const double* Mw = t.M + ( int64_t )cbase * ldm + lane;
  double pre[32];
  auto fetch = [&]( int k0 ) {
    const bool r0 = k0 + lane < t.K, r1 = k0 + lane + 32 < t.K;
#pragma unroll
    for ( int j = 0; j < 16; j++ ) {
      pre[2 * j] = ( j < ncol && r0 ) ? __ldg( Mw + k0 + j * ldm ) : 0.0;
      pre[2 * j + 1] = ( j < ncol && r1 ) ? __ldg( Mw + k0 + 32 + j * ldm ) : 0.0;
    }
  };
  if ( t.K > 0 ) fetch( 0 );
  for ( int kb = 0; kb < t.K; kb += GT_XS ) {
    const int rows = min( GT_XS, t.K - kb );
    __syncthreads();
    for ( int e = tid; e < rows * PCT; e += 128 ) { const int kk = e / PCT, q = e % PCT; xs[kk][q] = q < pcn ? v[( int64_t )( kb + kk ) * pitch + q] : 0.0; }
    __syncthreads();
    for ( int k0 = 0; k0 < rows; k0 += KC ) {
#pragma unroll
      for ( int j = 0; j < 16; j++ ) { slab[warp][j][lane] = pre[2 * j]; slab[warp][j][lane + 32] = pre[2 * j + 1]; }
      if ( kb + k0 + KC < t.K ) fetch( kb + k0 + KC );
      __syncwarp();
      if ( owner ) {
        const int kn = min( KC, rows - k0 );
        double acc[PCT];
#pragma unroll
        for ( int q = 0; q < PCT; q++ ) acc[q] = 0.0;
        const double* sl = slab[warp][lane & 15];
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

/* ---- supernodes with at most KC pivot columns (one canonical chunk, one solve block): the whole supernode in one CTA per sweep.
 * Nine tenths of the supernodes live here; through the generic task lists each of them would cost three launch-wide tiles.
 * The arithmetic is the canonical one (single-chunk chains in ascending k), so the bits equal those of the generic kernels.
 * forward : z = X y ;  w[update rows] -= L21 z
 * backward: v = z - L21' x[ancestors] (chunks of 64 update rows, chain per chunk) ;  x = X' v                                  */
constexpr int SVS_THREADS = 128;
template <int PCT>
__global__ void __launch_bounds__( SVS_THREADS ) sv_small_fwd_kernel( const int* __restrict__ sns, int pchunks, const SnMeta* __restrict__ sn, const double* __restrict__ L,
                                                                     const double* __restrict__ sinvF, const double* __restrict__ Y, double* __restrict__ Z, double* __restrict__ W, int p, int pitch ) {
  __shared__ double ys[KC][PCT];
  __shared__ double zs[KC][PCT];
  const SnMeta S = sn[sns[blockIdx.x / pchunks]];
  const int pc0 = ( blockIdx.x % pchunks ) * PCT, pcn = min( PCT, p - pc0 );
  const int k = S.k, u = S.f - k, tid = threadIdx.x;
  const int64_t ldx = ( k + 1 ) & ~1;
  for ( int e = tid; e < k * PCT; e += SVS_THREADS ) { const int c = e / PCT, q = e % PCT; ys[c][q] = q < pcn ? Y[( int64_t )( S.c0 + c ) * pitch + pc0 + q] : 0.0; }
  __syncthreads();
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
    for ( int q = 0; q < PCT; q++ ) { const double z = -( 0.0 - acc[q] ); zs[tid][q] = z; if ( q < pcn ) Z[( int64_t )( S.c0 + tid ) * pitch + pc0 + q] = z; }
  }
  __syncthreads();
  const double* Lp = L + S.lptr + k;
  for ( int i = tid; i < u; i += SVS_THREADS ) {
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
    double* wp = W + ( S.wptr + i ) * pitch + pc0;
#pragma unroll
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) wp[q] -= acc[q];
  }
}

template <int PCT>
__global__ void __launch_bounds__( SVS_THREADS ) sv_small_bwd_kernel( const int* __restrict__ sns, int pchunks, const SnMeta* __restrict__ sn, const double* __restrict__ M,
                                                                     const double* __restrict__ sinvB, const int* __restrict__ rowidx, double* __restrict__ Y, const double* __restrict__ Z, int p, int pitch ) {
  __shared__ double slab[KC][KC + 1];     /* [col][row of the chunk] */
  __shared__ double xs[KC][PCT];
  __shared__ double vs[KC][PCT];
  const SnMeta S = sn[sns[blockIdx.x / pchunks]];
  const int pc0 = ( blockIdx.x % pchunks ) * PCT, pcn = min( PCT, p - pc0 );
  const int k = S.k, u = S.f - k, tid = threadIdx.x;
  const int64_t ldx = ( k + 1 ) & ~1;
  const int* ri = rowidx + S.rowptr + k;
  const double* Mp = M + S.lptr + k;
  double cr[PCT];
#pragma unroll
  for ( int q = 0; q < PCT; q++ ) cr[q] = ( tid < k && q < pcn ) ? Z[( int64_t )( S.c0 + tid ) * pitch + pc0 + q] : 0.0;
  for ( int k0 = 0; k0 < u; k0 += KC ) {
    const int kn = min( KC, u - k0 );
    __syncthreads();
    for ( int e = tid; e < kn * PCT; e += SVS_THREADS ) { const int i = e / PCT, q = e % PCT; xs[i][q] = q < pcn ? Y[( int64_t )ri[k0 + i] * pitch + pc0 + q] : 0.0; }
    for ( int e = tid; e < k * KC; e += SVS_THREADS ) { const int i = e & ( KC - 1 ), c = e >> 6; if ( i < kn ) slab[c][i] = __ldg( Mp + k0 + i + ( int64_t )c * S.ld ); }
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
    for ( int q = 0; q < PCT; q++ ) if ( q < pcn ) Y[( int64_t )( S.c0 + tid ) * pitch + pc0 + q] = -( 0.0 - acc[q] );
  }
}

/* ---- the same task lists on the FP64 tensor pipe (p >= 5): C'[p x N] (-)= V'[p x K] * Mop[N x K]'  with the right-hand sides as
 * the M dimension of the tile.  V' (lda = pitch) and a gemv-type M (ldb = ld) are "NT" operands exactly as in the factorization
 * kernel; a gemvT-type M has the reduction index contiguous, so its tile is staged n-major by one bulk copy of 16 rows per column
 * (BT = true).  Every KC k-steps the accumulators are flushed into the running C registers: the canonical arithmetic above.
 * grid.y walks over the 64-wide slices of the right-hand sides. */
struct SvGemm { double* C; const double* A; const double* B; int ldc, lda, ldb; int M, N, K; int mode; };
__global__ void sv_expand_kernel( const SvTask* __restrict__ tasks, int64_t ntasks, SvBases Bs, int p, int pitch, SvGemm* __restrict__ out ) {
  const int64_t i = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= ntasks ) return;
  const SvTask t = tasks[i];
  SvGemm g;
  g.C = Bs.b[t.dsel] + t.doff * pitch; g.A = Bs.b[t.vsel] + t.voff * pitch; g.B = t.M;
  g.ldc = pitch; g.lda = pitch; g.ldb = t.ld; g.M = p; g.N = t.N; g.K = t.K; g.mode = t.mode;
  out[i] = g;
}

template <bool BT>
__global__ void __launch_bounds__( 288, 2 ) gemm_solve_dmma_kernel( const SvGemm* __restrict__ tasks, const SvTile* __restrict__ tiles, const double* __restrict__ zeros ) {
  constexpr int BM = 64, BN = 64, WM = 32, WN = 16, BK = BT ? 32 : 16, STAGES = BT ? 3 : 4;      /* BT: 64 row-run copies per stage -- 256-byte runs keep the copy issue rate below the tensor pipe's appetite */
  constexpr int NCW = 8, LDA = BM + 4, LDB = BT ? BK + 4 : BN + 4;
  constexpr int BSTAGE = BT ? BN * LDB : BK * LDB;
  constexpr int MT = WM / 8, NTL = WN / 8;
  extern __shared__ double sm[];
  double* As = sm;
  double* Bs = sm + STAGES * BK * LDA;
  uint64_t* full = reinterpret_cast<uint64_t*>( Bs + STAGES * BSTAGE );
  uint64_t* empty = full + STAGES;
  const SvTile tl = tiles[blockIdx.x];
  const SvGemm t = tasks[tl.task];
  const int m0 = blockIdx.y * BM, n0 = tl.o0;
  if ( m0 >= t.M ) return;
  const int mrem = t.M - m0, nrem = t.N - n0;
  const double* Ag = t.A + m0;
  const double* Bg = BT ? t.B + ( int64_t )n0 * t.ldb : t.B + n0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  /* 16-byte alignment of the bulk copies: NT operands are fetched one element early along m / n when they start at an odd
   * address (sa, sb); a BT operand that starts at an odd ROW is fetched one row early and one row longer (18 rows per column
   * and stage), and every fragment read is shifted by one (sbk) */
  const int sa = ( int )( ( ( uintptr_t )Ag >> 3 ) & 1 ), sb = BT ? 0 : ( int )( ( ( uintptr_t )Bg >> 3 ) & 1 ), sbk = BT ? ( int )( ( ( uintptr_t )Bg >> 3 ) & 1 ) : 0;
  const int K = t.K, nk = ( K + BK - 1 ) / BK;
  if ( tid == 0 ) {
    for ( int s = 0; s < STAGES; s++ ) { mbar_init( &full[s], 1 ); mbar_init( &empty[s], NCW ); }
    asm volatile( "fence.mbarrier_init.release.cluster;\n" ::: "memory" );
    asm volatile( "fence.proxy.async.shared::cta;\n" ::: "memory" );
  }
  __syncthreads();
  if ( warp == NCW ) {
    /* ---------------- producer warp */
    const int ra = ( ( min( BM, mrem ) + sa + 1 ) & ~1 ) * 8;
    if ( !BT ) {
      const int rb = ( ( min( BN, nrem ) + sb + 1 ) & ~1 ) * 8;
      const int q = lane & 15; const bool isb = lane >= 16;
      const double* g = isb ? Bg - sb : Ag - sa;
      const int64_t ldg = isb ? t.ldb : t.lda;
      const int bytes = isb ? rb : ra;
      for ( int kt = 0; kt < nk; kt++ ) {
        const int slot = kt % STAGES, use = kt / STAGES;
        if ( lane == 0 ) { if ( use > 0 ) mbar_wait( &empty[slot], ( use - 1 ) & 1 ); mbar_expect_tx( &full[slot], BK * ( ra + rb ) ); }
        __syncwarp();
        double* dst = isb ? Bs + slot * BSTAGE + q * LDB : As + slot * BK * LDA + q * LDA;
        const int kq = kt * BK + q;
        bulk_g2s( dst, kq < K ? ( const void* )( g + ( int64_t )kq * ldg ) : ( const void* )zeros, bytes, &full[slot] );
      }
    } else {
      /* A: lanes 0..15 one k-column each; B: every lane two columns of M, 16 contiguous rows (128 bytes) each */
      const int nval = min( BN, nrem ), cb = ( BK + 2 * sbk ) * 8;
      for ( int kt = 0; kt < nk; kt++ ) {
        const int slot = kt % STAGES, use = kt / STAGES;
        if ( lane == 0 ) { if ( use > 0 ) mbar_wait( &empty[slot], ( use - 1 ) & 1 ); mbar_expect_tx( &full[slot], BK * ra + nval * cb ); }
        __syncwarp();
        { const int kq = kt * BK + lane; bulk_g2s( As + slot * BK * LDA + lane * LDA, kq < K ? ( const void* )( Ag - sa + ( int64_t )kq * t.lda ) : ( const void* )zeros, ra, &full[slot] ); }      /* BK = 32: one k-column of A per lane */
#pragma unroll
        for ( int h = 0; h < 2; h++ ) { const int nn = lane + 32 * h; if ( nn < nval ) bulk_g2s( Bs + slot * BSTAGE + nn * LDB, Bg + ( int64_t )nn * t.ldb + kt * BK - sbk, cb, &full[slot] ); }
      }
    }
    return;
  }
  /* ---------------- consumers */
  const int wm0 = ( warp % ( BM / WM ) ) * WM, wn0 = ( warp / ( BM / WM ) ) * WN;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[MT][NTL][2], cr[MT][NTL][2];
#pragma unroll
  for ( int i = 0; i < MT; i++ )
#pragma unroll
    for ( int j = 0; j < NTL; j++ )
#pragma unroll
      for ( int h = 0; h < 2; h++ ) {
        acc[i][j][h] = 0.0;
        const int c = n0 + wn0 + j * 8 + 2 * fk + h, r = m0 + wm0 + i * 8 + fr;
        cr[i][j][h] = ( t.mode == 0 && c < t.N && r < t.M ) ? t.C[( int64_t )c * t.ldc + r] : 0.0;
      }
  for ( int kt = 0; kt < nk; kt++ ) {
    const int slot = kt % STAGES;
    mbar_wait( &full[slot], ( kt / STAGES ) & 1 );
    const double* as = As + slot * BK * LDA + wm0 + fr + sa;
    const double* bs = BT ? Bs + slot * BSTAGE + ( wn0 + fr ) * LDB + sbk : Bs + slot * BSTAGE + wn0 + fr + sb;
    const int kvalid = K - kt * BK;         /* BT: rows beyond K hold whatever follows the slab in memory -- never let them meet a product */
#pragma unroll
    for ( int kk = 0; kk < BK; kk += 4 ) {
      double af[MT], bf[NTL];
#pragma unroll
      for ( int i = 0; i < MT; i++ ) af[i] = as[( kk + fk ) * LDA + i * 8];
#pragma unroll
      for ( int j = 0; j < NTL; j++ ) bf[j] = BT ? ( kk + fk < kvalid ? bs[j * 8 * LDB + kk + fk] : 0.0 ) : bs[( kk + fk ) * LDB + j * 8];
#pragma unroll
      for ( int i = 0; i < MT; i++ )
#pragma unroll
        for ( int j = 0; j < NTL; j++ ) dmma884( acc[i][j][0], acc[i][j][1], af[i], bf[j] );
    }
    __syncwarp();
    if ( lane == 0 ) mbar_arrive( &empty[slot] );
    if ( ( ( kt + 1 ) & ( KC / BK - 1 ) ) == 0 || kt == nk - 1 ) {      /* end of a canonical chunk */
#pragma unroll
      for ( int i = 0; i < MT; i++ )
#pragma unroll
        for ( int j = 0; j < NTL; j++ )
#pragma unroll
          for ( int h = 0; h < 2; h++ ) { cr[i][j][h] -= acc[i][j][h]; acc[i][j][h] = 0.0; }
    }
  }
#pragma unroll
  for ( int j = 0; j < NTL; j++ ) {
#pragma unroll
    for ( int h = 0; h < 2; h++ ) {
      const int c = n0 + wn0 + j * 8 + 2 * fk + h;
      if ( c < t.N ) {
        double* cc = t.C + ( int64_t )c * t.ldc;
#pragma unroll
        for ( int i = 0; i < MT; i++ ) {
          const int r = m0 + wm0 + i * 8 + fr;
          if ( r < t.M ) cc[r] = t.mode == 0 ? cr[i][j][h] : -cr[i][j][h];
        }
      }
    }
  }
}

/* ---- explicit inverses of the SB-wide triangular diagonal blocks, from the panel and the 64x64 inverses the
 * factorization already produced.  One CTA per (solve block, 64-column block column jb):
 *   X_jj = T_j ;  X_ij = -T_i * sum_{q=jb}^{i-1} P_iq X_qj   (i > jb),   T_i = inv_i (* row interchanges of block i, LU)
 * X goes out untransposed (forward sweep) and/or transposed (backward sweep). */
struct SinvTask {
  const double* P; int ld;      /* panel (L or U') of the supernode                     */
  int j0, nb, jb;               /* solve block [j0, j0+nb), block column jb of it       */
  const double* inv64;          /* 64x64 inverses of the supernode (block i at +i*4096) */
  const int* piv;               /* LU: pivots of the supernode's columns (NULL: none)   */
  double* X; double* XT;        /* nb x nb outputs, leading dimension ldx (XT may be NULL) */
  int ldx;                      /* nb rounded up to even: 16-byte aligned columns for the bulk copies of the tensor-path sweeps */
};
constexpr int SINV_SMEM = 3 * NB * ( NB + 1 ) * 8 + NB * 4;

__global__ void __launch_bounds__( 256 ) build_sinv_kernel( const SinvTask* __restrict__ tasks ) {
  extern __shared__ double dsm[];
  double ( *As )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm );
  double ( *Bs )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + NB * ( NB + 1 ) );
  double ( *Ts )[NB + 1] = reinterpret_cast<double ( * )[NB + 1]>( dsm + 2 * NB * ( NB + 1 ) );
  int* sig = reinterpret_cast<int*>( dsm + 3 * NB * ( NB + 1 ) );
  const SinvTask t = tasks[blockIdx.x];
  const int tid = threadIdx.x, nb = t.nb, jb = t.jb;
  const int nsub = ( nb + NB - 1 ) / NB;
  const int wj = min( NB, nb - NB * jb );
  const int tr = ( tid & 15 ) * 4, tc = ( tid >> 4 ) * 4;
  for ( int i = jb; i < nsub; i++ ) {
    const int hi = min( NB, nb - NB * i );
    /* T_i (with the block's interchanges folded in as a column permutation) */
    __syncthreads();
    if ( tid < NB ) sig[tid] = tid;
    __syncthreads();
    if ( t.piv && tid == 0 ) { const int* pv = t.piv + t.j0 + NB * i; for ( int j = 0; j < hi; j++ ) { const int pj = pv[j]; if ( pj != j ) { const int s = sig[j]; sig[j] = sig[pj]; sig[pj] = s; } } }
    __syncthreads();
    { const double* iv = t.inv64 + ( int64_t )( t.j0 / NB + i ) * NB * NB;
      for ( int e = tid; e < NB * NB; e += 256 ) Ts[e % NB][e / NB] = 0.0;
      __syncthreads();
      for ( int e = tid; e < hi * hi; e += 256 ) { const int r = e % hi, c = e / hi; Ts[r][sig[c]] = iv[r + c * hi]; } }
    __syncthreads();
    double acc[4][4];
#pragma unroll
    for ( int a = 0; a < 4; a++ )
#pragma unroll
      for ( int c = 0; c < 4; c++ ) acc[a][c] = 0.0;
    if ( i == jb ) {
#pragma unroll
      for ( int a = 0; a < 4; a++ )
#pragma unroll
        for ( int c = 0; c < 4; c++ ) acc[a][c] = Ts[tr + a][tc + c];
    } else {
      for ( int q = jb; q < i; q++ ) {
        const int wq = min( NB, nb - NB * q );
        __syncthreads();
        for ( int e = tid; e < NB * NB; e += 256 ) {
          const int r = e % NB, c = e / NB;
          As[r][c] = ( r < hi && c < wq ) ? t.P[( t.j0 + NB * i + r ) + ( int64_t )( t.j0 + NB * q + c ) * t.ld] : 0.0;
          Bs[r][c] = ( r < wq && c < wj ) ? t.X[( NB * q + r ) + ( int64_t )( NB * jb + c ) * t.ldx] : 0.0;
        }
        __syncthreads();
        for ( int kk = 0; kk < NB; kk++ ) {
          double av[4], bv[4];
#pragma unroll
          for ( int a = 0; a < 4; a++ ) av[a] = As[tr + a][kk];
#pragma unroll
          for ( int c = 0; c < 4; c++ ) bv[c] = Bs[kk][tc + c];
#pragma unroll
          for ( int a = 0; a < 4; a++ )
#pragma unroll
            for ( int c = 0; c < 4; c++ ) acc[a][c] += av[a] * bv[c];
        }
      }
      __syncthreads();
#pragma unroll
      for ( int a = 0; a < 4; a++ )
#pragma unroll
        for ( int c = 0; c < 4; c++ ) { Bs[tr + a][tc + c] = acc[a][c]; acc[a][c] = 0.0; }
      __syncthreads();
      for ( int kk = 0; kk < NB; kk++ ) {
        double av[4], bv[4];
#pragma unroll
        for ( int a = 0; a < 4; a++ ) av[a] = Ts[tr + a][kk];
#pragma unroll
        for ( int c = 0; c < 4; c++ ) bv[c] = Bs[kk][tc + c];
#pragma unroll
        for ( int a = 0; a < 4; a++ )
#pragma unroll
          for ( int c = 0; c < 4; c++ ) acc[a][c] -= av[a] * bv[c];
      }
    }
#pragma unroll
    for ( int a = 0; a < 4; a++ )
#pragma unroll
      for ( int c = 0; c < 4; c++ ) {
        const int r = tr + a, cc = tc + c;
        if ( r < hi && cc < wj ) {
          const int gr = NB * i + r, gc = NB * jb + cc;
          t.X[gr + ( int64_t )gc * t.ldx] = acc[a][c];
          if ( t.XT ) t.XT[gc + ( int64_t )gr * t.ldx] = acc[a][c];
        }
      }
    /* blocks above the diagonal of this block column are zero */
    if ( i == jb ) {
      for ( int e = tid; e < NB * jb * wj; e += 256 ) { const int r = e % ( NB * jb ), c = e / ( NB * jb ); t.X[r + ( int64_t )( NB * jb + c ) * t.ldx] = 0.0; if ( t.XT ) t.XT[( NB * jb + c ) + ( int64_t )r * t.ldx] = 0.0; }
    }
  }
}

/* residual r = b - A x with the sum carried in double-double (error-free products through
 * fma, Knuth two-sum), so that one refinement step recovers x to working accuracy */
__global__ void residual_dd_kernel( const int64_t* __restrict__ rp, const int* __restrict__ col, const int* __restrict__ src,
                                    const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ b,
                                    int n, int p, double* __restrict__ r, unsigned long long* __restrict__ omega_bits ) {
  int64_t e = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = ( int64_t )n * p, stride = ( int64_t )gridDim.x * blockDim.x;
  double omega = 0.0;   /* componentwise backward error max_i |r_i| / (|A||x| + |b|)_i  (Oettli-Prager) */
  for ( ; e < tot; e += stride ) {
    const int i = ( int )( e % n ); const int64_t c = e / n;
    const double* xc = x + c * n;
    double hi = b[e], lo = 0.0, den = fabs( b[e] );
    for ( int64_t q = rp[i]; q < rp[i + 1]; q++ ) {
      const double a = vals[src[q]], xv = xc[col[q]];
      const double pr = -a * xv;
      const double pe = fma( -a, xv, -pr );        /* exact: -a*xv = pr + pe */
      const double s = hi + pr;
      const double bb = s - hi;
      const double se = ( hi - ( s - bb ) ) + ( pr - bb );
      hi = s; lo += se + pe;
      den += fabs( pr );
    }
    const double ri = hi + lo;
    r[e] = ri;
    /* NaN / Inf anywhere in x must surface as omega = +Inf (fmax would silently drop a NaN) */
    double w = den > 0.0 ? fabs( ri ) / den : ( ri != 0.0 ? CUDART_INF : 0.0 );
    if ( !( w == w ) ) w = CUDART_INF;
    omega = fmax( omega, w );
  }
  for ( int o = 16; o; o >>= 1 ) omega = fmax( omega, __shfl_xor_sync( 0xffffffffu, omega, o ) );
  if ( ( threadIdx.x & 31 ) == 0 && omega > 0.0 ) atomicMax( omega_bits, ( unsigned long long )__double_as_longlong( omega ) );   /* positive doubles order like integers */
}
__global__ void axpy_kernel( double* __restrict__ x, const double* __restrict__ d, int64_t tot ) {
  int64_t e = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = ( int64_t )gridDim.x * blockDim.x;
  for ( ; e < tot; e += stride ) x[e] += d[e];
}

/* checksum of the factor: out[c0 + j] = sum of squares of the stored entries of column j of the supernode's panel (rows >= j),
 * one CTA per supernode, one warp per column, lanes stride the rows and are joined by a fixed shuffle tree -- the same bits
 * whatever the launch geometry.  Cholesky: the sum over all columns equals trace(A). */
__global__ void __launch_bounds__( 256 ) panel_colnorm_kernel( const SnMeta* __restrict__ sn, int ns, const double* __restrict__ L, double* __restrict__ out ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for ( int s = blockIdx.x; s < ns; s += gridDim.x ) {
    const SnMeta S = sn[s];
    for ( int j = warp; j < S.k; j += 8 ) {
      const double* col = L + S.lptr + ( int64_t )j * S.ld;
      double acc = 0.0;
      for ( int i = j + lane; i < S.f; i += 32 ) acc += col[i] * col[i];
      for ( int o = 16; o; o >>= 1 ) acc += __shfl_xor_sync( 0xffffffffu, acc, o );
      if ( lane == 0 ) out[S.c0 + j] = acc;
    }
  }
}

/* LU: verification of threshold partial pivoting over ALL fully-summed rows of every front.  The diagonal-block kernel pivots inside its
 * 64 rows (|l| <= 1 there); for the fully-summed rows below the block, l(i,j) = a~(i,j) / u(j,j), so  |u(j,j)| >= u * |a~(i,j)|  (the
 * acceptance test of threshold pivoting, UMFPACK / MUMPS) holds for every candidate row exactly when  max |l(i,j)| <= 1/u.
 * One pass over the pivot blocks of the finished L panels: largest multiplier, and how many exceed the bound. */
__global__ void __launch_bounds__( 256 ) lu_growth_kernel( const SnMeta* __restrict__ sn, int ns, const double* __restrict__ L, double bound,
                                                           unsigned long long* __restrict__ max_bits, int* __restrict__ nviol ) {
  double mx = 0.0; int nv = 0;
  for ( int s = blockIdx.x; s < ns; s += gridDim.x ) {
    const SnMeta S = sn[s];
    if ( S.k <= NB ) continue;
    for ( int j = threadIdx.x >> 5; j < S.k; j += 8 ) {
      const double* col = L + S.lptr + ( int64_t )j * S.ld;
      for ( int i = ( j / NB + 1 ) * NB + ( threadIdx.x & 31 ); i < S.k; i += 32 ) { const double v = fabs( col[i] ); if ( !( v <= bound ) ) nv++; if ( v > mx || v != v ) mx = v != v ? CUDART_INF : v; }
    }
  }
  for ( int o = 16; o; o >>= 1 ) { mx = fmax( mx, __shfl_xor_sync( 0xffffffffu, mx, o ) ); nv += __shfl_xor_sync( 0xffffffffu, nv, o ); }
  if ( ( threadIdx.x & 31 ) == 0 ) { if ( mx > 0.0 ) atomicMax( max_bits, ( unsigned long long )__double_as_longlong( mx ) ); if ( nv ) atomicAdd( nviol, nv ); }
}

/* micro-benchmark: issue-rate roof of DMMA.8x8x4 (registers only, 16 independent accumulators per warp) */
__global__ void __launch_bounds__( 256 ) dmma_peak_kernel( double* out, int iters ) {
  double acc[16][2];
#pragma unroll
  for ( int i = 0; i < 16; i++ ) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for ( int it = 0; it < iters; it++ ) {
#pragma unroll
    for ( int i = 0; i < 16; i++ ) dmma884( acc[i][0], acc[i][1], a, b );
  }
  double s = 0.0;
#pragma unroll
  for ( int i = 0; i < 16; i++ ) s += acc[i][0] + acc[i][1];
  if ( s == 12345.678 ) out[0] = s;
}
template <int ILP>
__global__ void dmma_occ_kernel( double* out, int iters ) {
  double acc[ILP][2];
#pragma unroll
  for ( int i = 0; i < ILP; i++ ) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for ( int it = 0; it < iters; it++ ) {
#pragma unroll
    for ( int i = 0; i < ILP; i++ ) dmma884( acc[i][0], acc[i][1], a, b );
  }
  double s = 0.0;
#pragma unroll
  for ( int i = 0; i < ILP; i++ ) s += acc[i][0] + acc[i][1];
  if ( s == 12345.678 ) out[0] = s;
}
/* micro-benchmark: FP64 FMA pipe roof (DFMA, 16 independent chains per thread) */
__global__ void __launch_bounds__( 256 ) dfma_peak_kernel( double* out, int iters ) {
  double acc[16];
#pragma unroll
  for ( int i = 0; i < 16; i++ ) acc[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for ( int it = 0; it < iters; it++ ) {
#pragma unroll
    for ( int i = 0; i < 16; i++ ) acc[i] = fma( acc[i], a, b );
  }
  double s = 0.0;
#pragma unroll
  for ( int i = 0; i < 16; i++ ) s += acc[i];
  if ( s == 12345.678 ) out[0] = s;
}

/* micro-benchmark: plain streaming copy */
__global__ void stream_copy_kernel( const double2* __restrict__ a, double2* __restrict__ b, int64_t n2 ) {
  int64_t i = ( int64_t )blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = ( int64_t )gridDim.x * blockDim.x;
  for ( ; i < n2; i += stride ) b[i] = a[i];
}

} /* namespace b200 */
