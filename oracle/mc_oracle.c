/* mc_oracle.c -- CPU oracle for the factor/solve hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libmeagre_crowd_b200.so) never does.
 *
 * What it restates.  The reference has no numeric code of its own: the path runs inside
 * CHOLMOD (src/solver_cholmod.c:257,272,303), UMFPACK (src/solver_umfpack.c:73,98,133)
 * and MUMPS (src/solver_mumps.c:176-179,211-213,274-276), none vendored under
 * /root/reference and none installable here (get_dependencies:12-40 pins SuiteSparse 5.1.2,
 * MUMPS 4.9.2).  The oracle therefore restates the published algorithms those libraries
 * implement and is PINNED on what the reference's tests pin: x for the six golden answer
 * files tests/{unsym,sym,sym_posdef}-{default,rhs1}-ans.mm (testsuite.at:312-346).
 * Etree / column counts / supernodes / nnz(L) are "parity unpinned" in the reference (no
 * fixture records them); here they are pinned against a dense Cholesky of the permuted
 * matrix instead (tests/test_symbolic.py).
 *
 *   oracle_dense_solve        LU with partial pivoting, dense (what UMFPACK/MUMPS reduce to at n=5)
 *   oracle_etree              Liu's elimination tree
 *   oracle_chol_uplooking     sparse up-looking Cholesky (Davis, "Direct Methods", ch. 4):
 *                             L, column counts, nnz(L) for a GIVEN ordering
 *   oracle_chol_solve         x = P' L^-T L^-1 P b
 *   oracle_mf_cholesky        supernodal multifrontal Cholesky on the CPU with BLAS-3
 *                             (dpotrf/dtrsm/dsyrk, the kernels CHOLMOD's supernodal path calls),
 *                             OpenMP over the fronts of a tree level; used as the timed CPU
 *                             baseline and as a mid-size check of the GPU factor
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------ dense LU */
int oracle_dense_solve( int n, const double* A, const double* b, int p, double* x ) {
  double* M = malloc( sizeof( double ) * n * n );
  int* piv = malloc( sizeof( int ) * n );
  memcpy( M, A, sizeof( double ) * n * n );
  memcpy( x, b, sizeof( double ) * n * p );
  for ( int j = 0; j < n; j++ ) {
    int pr = j; double best = fabs( M[j + ( size_t )j * n] );
    for ( int i = j + 1; i < n; i++ ) if ( fabs( M[i + ( size_t )j * n] ) > best ) { best = fabs( M[i + ( size_t )j * n] ); pr = i; }
    if ( best == 0.0 ) { free( M ); free( piv ); return -1; }
    piv[j] = pr;
    if ( pr != j ) {
      for ( int c = 0; c < n; c++ ) { double t = M[j + ( size_t )c * n]; M[j + ( size_t )c * n] = M[pr + ( size_t )c * n]; M[pr + ( size_t )c * n] = t; }
      for ( int c = 0; c < p; c++ ) { double t = x[j + ( size_t )c * n]; x[j + ( size_t )c * n] = x[pr + ( size_t )c * n]; x[pr + ( size_t )c * n] = t; }
    }
    const double d = M[j + ( size_t )j * n];
    for ( int i = j + 1; i < n; i++ ) M[i + ( size_t )j * n] /= d;
    for ( int c = j + 1; c < n; c++ ) { const double u = M[j + ( size_t )c * n]; if ( u != 0.0 ) for ( int i = j + 1; i < n; i++ ) M[i + ( size_t )c * n] -= M[i + ( size_t )j * n] * u; }
    for ( int c = 0; c < p; c++ ) { const double u = x[j + ( size_t )c * n]; if ( u != 0.0 ) for ( int i = j + 1; i < n; i++ ) x[i + ( size_t )c * n] -= M[i + ( size_t )j * n] * u; }
  }
  for ( int c = 0; c < p; c++ )
    for ( int j = n - 1; j >= 0; j-- ) {
      double v = x[j + ( size_t )c * n] / M[j + ( size_t )j * n];
      x[j + ( size_t )c * n] = v;
      for ( int i = 0; i < j; i++ ) x[i + ( size_t )c * n] -= M[i + ( size_t )j * n] * v;
    }
  free( M ); free( piv );
  return 0;
}

/* ------------------------------------- dense symmetric indefinite solve: P A P' = L D L'
 * Bunch-Kaufman partial pivoting (1x1 / 2x2 pivots, alpha = (1+sqrt(17))/8), the published algorithm behind LAPACK dsytf2 and
 * the symmetric-indefinite mode of the reference's MUMPS backend (SYM=2, src/solver_mumps.c:44-51; the reference's tests run
 * tests/sym.mm on its general backends, testsuite.at:325-329).  A: n x n column-major, full symmetric; b, x: n x p column-major.
 * Pivots are searched over the WHOLE remaining matrix here; the device path restricts them to a diagonal block and perturbs,
 * so agreement of the solutions is what the tests check.  Returns 0, or j+1 if column j is exactly zero. */
int oracle_ldlt_solve( int n, const double* A, const double* b, int p, double* x ) {
  const double alpha = ( 1.0 + sqrt( 17.0 ) ) / 8.0;
  double* M = malloc( sizeof( double ) * ( size_t )n * n ); memcpy( M, A, sizeof( double ) * ( size_t )n * n );
  int* sig = malloc( sizeof( int ) * n ); int* two = calloc( n, sizeof( int ) );
  for ( int i = 0; i < n; i++ ) sig[i] = i;
#define AT( r, c ) M[( r ) + ( size_t )( c ) * n]
  int j = 0, ret = 0;
  while ( j < n ) {
    double colmax = 0; int imax = j;
    for ( int i = j + 1; i < n; i++ ) if ( fabs( AT( i, j ) ) > colmax ) { colmax = fabs( AT( i, j ) ); imax = i; }
    const double absakk = fabs( AT( j, j ) );
    int kp = j, step = 1;
    if ( absakk == 0.0 && colmax == 0.0 ) { ret = j + 1; break; }
    if ( absakk < alpha * colmax ) {
      double rowmax = 0;
      for ( int q = j; q < n; q++ ) if ( q != imax ) { const double v = fabs( q < imax ? AT( imax, q ) : AT( q, imax ) ); if ( v > rowmax ) rowmax = v; }
      if ( absakk * rowmax >= alpha * colmax * colmax ) kp = j;
      else if ( fabs( AT( imax, imax ) ) >= alpha * rowmax ) kp = imax;
      else { kp = imax; step = 2; }
    }
    const int kk = j + step - 1;
    if ( kp != kk ) { /* symmetric interchange on the lower triangle (and on the computed columns of L) */
      for ( int c = 0; c < kk; c++ ) { const double t = AT( kk, c ); AT( kk, c ) = AT( kp, c ); AT( kp, c ) = t; }
      for ( int r = kk + 1; r < kp; r++ ) { const double t = AT( r, kk ); AT( r, kk ) = AT( kp, r ); AT( kp, r ) = t; }
      for ( int r = kp + 1; r < n; r++ ) { const double t = AT( r, kk ); AT( r, kk ) = AT( r, kp ); AT( r, kp ) = t; }
      { const double t = AT( kk, kk ); AT( kk, kk ) = AT( kp, kp ); AT( kp, kp ) = t; }
      { const int t = sig[kk]; sig[kk] = sig[kp]; sig[kp] = t; }
    }
    if ( step == 1 ) {
      const double d = AT( j, j );
      for ( int c = j + 1; c < n; c++ ) { const double l = AT( c, j ) / d; for ( int r = c; r < n; r++ ) AT( r, c ) -= AT( r, j ) * l; }
      for ( int r = j + 1; r < n; r++ ) AT( r, j ) /= d;
      j += 1;
    } else {
      two[j] = 1;
      const double d21 = AT( j + 1, j ), e11 = AT( j, j ) / d21, e22 = AT( j + 1, j + 1 ) / d21, sc = 1.0 / ( ( e11 * e22 - 1.0 ) * d21 );
      for ( int c = j + 2; c < n; c++ ) {
        const double l1 = sc * ( e22 * AT( c, j ) - AT( c, j + 1 ) ), l2 = sc * ( e11 * AT( c, j + 1 ) - AT( c, j ) );
        for ( int r = c; r < n; r++ ) AT( r, c ) -= AT( r, j ) * l1 + AT( r, j + 1 ) * l2;
      }
      for ( int r = j + 2; r < n; r++ ) { const double a1 = AT( r, j ), a2 = AT( r, j + 1 ); AT( r, j ) = sc * ( e22 * a1 - a2 ); AT( r, j + 1 ) = sc * ( e11 * a2 - a1 ); }
      j += 2;
    }
  }
  if ( ret == 0 ) {
    double* y = malloc( sizeof( double ) * n );
    for ( int c = 0; c < p; c++ ) {
      for ( int i = 0; i < n; i++ ) y[i] = b[sig[i] + ( size_t )c * n];
      for ( int q = 0; q < n; q++ ) { const int w = two[q] ? 2 : 1; for ( int e = 0; e < w; e++ ) for ( int r = q + w; r < n; r++ ) y[r] -= AT( r, q + e ) * y[q + e]; q += w - 1; }
      for ( int q = 0; q < n; q++ ) {
        if ( two[q] ) { const double d21 = AT( q + 1, q ), e11 = AT( q, q ) / d21, e22 = AT( q + 1, q + 1 ) / d21, sc = 1.0 / ( ( e11 * e22 - 1.0 ) * d21 ); const double a = y[q], bb = y[q + 1]; y[q] = sc * ( e22 * a - bb ); y[q + 1] = sc * ( e11 * bb - a ); q++; }
        else y[q] /= AT( q, q );
      }
      for ( int q = n - 1; q >= 0; q-- ) { const int lo = ( q > 0 && two[q - 1] ) ? q - 1 : q; const int w = q - lo + 1; for ( int e = 0; e < w; e++ ) { double v = y[lo + e]; for ( int r = lo + w; r < n; r++ ) v -= AT( r, lo + e ) * y[r]; y[lo + e] = v; } q = lo; }
      for ( int i = 0; i < n; i++ ) x[sig[i] + ( size_t )c * n] = y[i];
    }
    free( y );
  }
#undef AT
  free( M ); free( sig ); free( two );
  return ret;
}

/* ------------------------------------------------------- etree (Liu 1990)
 * cp/ri: CSC of the symmetric matrix, only entries with row < col are read */
void oracle_etree( int n, const int* cp, const int* ri, int* parent ) {
  int* anc = malloc( sizeof( int ) * n );
  for ( int k = 0; k < n; k++ ) {
    parent[k] = -1; anc[k] = -1;
    for ( int p = cp[k]; p < cp[k + 1]; p++ ) {
      int i = ri[p];
      while ( i != -1 && i < k ) { int nx = anc[i]; anc[i] = k; if ( nx == -1 ) parent[i] = k; i = nx; }
    }
  }
  free( anc );
}

/* --------------------------------------- sparse up-looking Cholesky
 * Input: upper triangle (row<=col) of the ALREADY PERMUTED matrix in CSC (rows may be unsorted).
 * Output: parent[n], colcount[n]; if Lp/Li/Lx != NULL the factor itself (Lp has n+1 entries,
 * Li/Lx must hold sum(colcount) entries -- call once with NULL to size). Returns nnz(L) or
 * -(k+1) if pivot k is not positive. */
int64_t oracle_chol_uplooking( int n, const int* cp, const int* ri, const double* vx, int* parent, int* colcount,
                               int64_t* Lp, int* Li, double* Lx ) {
  oracle_etree( n, cp, ri, parent );
  int* flag = malloc( sizeof( int ) * n );
  int* stack = malloc( sizeof( int ) * n );
  int* pattern = malloc( sizeof( int ) * n );
  /* pass 1: column counts by walking every row subtree (exact, O(nnz(L))) */
  for ( int k = 0; k < n; k++ ) { colcount[k] = 1; flag[k] = -1; }
  for ( int k = 0; k < n; k++ ) {
    flag[k] = k;
    for ( int p = cp[k]; p < cp[k + 1]; p++ ) {
      int i = ri[p];
      if ( i >= k ) continue;
      for ( ; flag[i] != k; i = parent[i] ) { colcount[i]++; flag[i] = k; }
    }
  }
  int64_t nnz = 0;
  for ( int k = 0; k < n; k++ ) nnz += colcount[k];
  if ( !Lp || !Li || !Lx ) { free( flag ); free( stack ); free( pattern ); return nnz; }
  /* pass 2: numeric, row k of L at a time */
  int64_t* next = malloc( sizeof( int64_t ) * n );
  double* xw = calloc( n, sizeof( double ) );
  Lp[0] = 0;
  for ( int k = 0; k < n; k++ ) { Lp[k + 1] = Lp[k] + colcount[k]; next[k] = Lp[k]; flag[k] = -1; }
  int64_t ret = nnz;
  for ( int k = 0; k < n; k++ ) {
    /* nonzero pattern of row k = reach of the entries of column k of the upper triangle in the etree */
    int top = n;
    flag[k] = k;
    double d = 0.0;
    for ( int p = cp[k]; p < cp[k + 1]; p++ ) {
      int i = ri[p];
      if ( i > k ) continue;
      if ( i == k ) { d += vx[p]; continue; }
      xw[i] += vx[p];
      int len = 0;
      for ( ; flag[i] != k; i = parent[i] ) { stack[len++] = i; flag[i] = k; }
      while ( len > 0 ) pattern[--top] = stack[--len];
    }
    for ( ; top < n; top++ ) {
      const int i = pattern[top];
      const double lki = xw[i] / Lx[Lp[i]];
      xw[i] = 0.0;
      for ( int64_t q = Lp[i] + 1; q < next[i]; q++ ) xw[Li[q]] -= Lx[q] * lki;
      d -= lki * lki;
      const int64_t q = next[i]++;
      Li[q] = k; Lx[q] = lki;
    }
    if ( !( d > 0.0 ) ) { ret = -( int64_t )( k + 1 ); break; }
    const int64_t q = next[k]++;
    Li[q] = k; Lx[q] = sqrt( d );
  }
  free( next ); free( xw ); free( flag ); free( stack ); free( pattern );
  return ret;
}

/* x = L^-T L^-1 b for one right-hand side, in place (permuted numbering) */
void oracle_chol_solve( int n, const int64_t* Lp, const int* Li, const double* Lx, double* x ) {
  for ( int j = 0; j < n; j++ ) { x[j] /= Lx[Lp[j]]; for ( int64_t q = Lp[j] + 1; q < Lp[j + 1]; q++ ) x[Li[q]] -= Lx[q] * x[j]; }
  for ( int j = n - 1; j >= 0; j-- ) { for ( int64_t q = Lp[j] + 1; q < Lp[j + 1]; q++ ) x[j] -= Lx[q] * x[Li[q]]; x[j] /= Lx[Lp[j]]; }
}

/* -------------------------------------------- BLAS-3 multifrontal baseline */
typedef void ( *dpotrf_t )( const char*, const int*, double*, const int*, int* );
typedef void ( *dtrsm_t )( const char*, const char*, const char*, const char*, const int*, const int*, const double*, const double*, const int*, double*, const int* );
typedef void ( *dsyrk_t )( const char*, const char*, const int*, const int*, const double*, const double*, const int*, const double*, double*, const int* );
typedef void ( *setthr_t )( int );
static dpotrf_t f_dpotrf; static dtrsm_t f_dtrsm; static dsyrk_t f_dsyrk; static setthr_t f_setthr; static int blas_state = 0;

/* returns 1 if an optimised BLAS was found (path list separated by ':'), 0 -> built-in loops */
int oracle_load_blas( const char* paths ) {
  if ( blas_state ) return blas_state > 0;
  blas_state = -1;
  if ( !paths ) return 0;
  char* dup = strdup( paths );
  for ( char* tok = strtok( dup, ":" ); tok; tok = strtok( NULL, ":" ) ) {
    void* h = dlopen( tok, RTLD_NOW | RTLD_LOCAL );
    if ( !h ) continue;
    const char* pre[2] = { "scipy_", "" };
    for ( int k = 0; k < 2; k++ ) {
      char nm[64];
      snprintf( nm, 64, "%sdpotrf_", pre[k] ); f_dpotrf = ( dpotrf_t )dlsym( h, nm );
      snprintf( nm, 64, "%sdtrsm_", pre[k] ); f_dtrsm = ( dtrsm_t )dlsym( h, nm );
      snprintf( nm, 64, "%sdsyrk_", pre[k] ); f_dsyrk = ( dsyrk_t )dlsym( h, nm );
      snprintf( nm, 64, "%sopenblas_set_num_threads", pre[k] ); f_setthr = ( setthr_t )dlsym( h, nm );
      if ( f_dpotrf && f_dtrsm && f_dsyrk ) { blas_state = 1; free( dup ); return 1; }
    }
    dlclose( h );
  }
  free( dup );
  f_dpotrf = NULL; f_dtrsm = NULL; f_dsyrk = NULL;
  return 0;
}

/* dense front: F is f x f column-major, lower triangle valid; factor k pivots, update the rest */
static int front_factor( double* F, int f, int k ) {
  const int u = f - k;
  if ( blas_state > 0 ) {
    int info = 0; const double one = 1.0, mone = -1.0;
    f_dpotrf( "L", &k, F, &f, &info );
    if ( info ) return info;
    if ( u > 0 ) {
      f_dtrsm( "R", "L", "T", "N", &u, &k, &one, F, &f, F + k, &f );
      f_dsyrk( "L", "N", &u, &k, &mone, F + k, &f, &one, F + k + ( size_t )k * f, &f );
    }
    return 0;
  }
  for ( int j = 0; j < k; j++ ) {
    double d = F[j + ( size_t )j * f];
    if ( !( d > 0.0 ) ) return j + 1;
    d = sqrt( d ); F[j + ( size_t )j * f] = d;
    for ( int i = j + 1; i < f; i++ ) F[i + ( size_t )j * f] /= d;
    for ( int c = j + 1; c < f; c++ ) { const double l = F[c + ( size_t )j * f]; if ( l != 0.0 ) for ( int i = c; i < f; i++ ) F[i + ( size_t )c * f] -= F[i + ( size_t )j * f] * l; }
  }
  return 0;
}

/* Multifrontal Cholesky driven by a supernode partition + front row lists (the analysis output
 * of the backend under test, or any other valid one).  A: upper CSC of the ORIGINAL matrix;
 * iperm: old->new.  Lpanels receives the f_s x k_s panels at lptr[s] (same layout as the GPU path).
 * threads: OpenMP threads over the fronts of one level (BLAS calls are made single-threaded
 * below the top of the tree and multi-threaded for big fronts). Returns 0, or >0 if not SPD. */
int oracle_mf_cholesky( int n, const int* cp, const int* ri, const double* vx, const int* iperm,
                        int ns, const int* sn_start, const int* sn_parent, const int* sn_depth, int maxdepth,
                        const int64_t* rowptr, const int* rowidx, const int64_t* lptr, double* Lpanels, int threads ) {
  ( void )n;
  int err = 0;
  double** cb = calloc( ns, sizeof( double* ) );
  /* children lists */
  int* chead = malloc( sizeof( int ) * ns ); int* cnext = malloc( sizeof( int ) * ns );
  for ( int s = 0; s < ns; s++ ) chead[s] = -1;
  for ( int s = ns - 1; s >= 0; s-- ) if ( sn_parent[s] >= 0 ) { cnext[s] = chead[sn_parent[s]]; chead[sn_parent[s]] = s; }
  /* entries of A grouped by the supernode of their (permuted) column */
  int* col2sn = malloc( sizeof( int ) * n );
  for ( int s = 0; s < ns; s++ ) for ( int j = sn_start[s]; j < sn_start[s + 1]; j++ ) col2sn[j] = s;
  int64_t* acnt = calloc( ns + 1, sizeof( int64_t ) );
  const int64_t nnz = cp[n];
  for ( int j = 0; j < n; j++ ) for ( int p = cp[j]; p < cp[j + 1]; p++ ) { int i = ri[p]; if ( i > j ) continue; int a = iperm[i], b = iperm[j]; acnt[col2sn[a < b ? a : b] + 1]++; }
  for ( int s = 0; s < ns; s++ ) acnt[s + 1] += acnt[s];
  int* er = malloc( sizeof( int ) * ( nnz + 1 ) ); int* ec = malloc( sizeof( int ) * ( nnz + 1 ) ); double* ev = malloc( sizeof( double ) * ( nnz + 1 ) );
  { int64_t* w = malloc( sizeof( int64_t ) * ( ns + 1 ) ); memcpy( w, acnt, sizeof( int64_t ) * ( ns + 1 ) );
    for ( int j = 0; j < n; j++ ) for ( int p = cp[j]; p < cp[j + 1]; p++ ) { int i = ri[p]; if ( i > j ) continue; int a = iperm[i], b = iperm[j]; int c = a < b ? a : b, r = a < b ? b : a; int64_t q = w[col2sn[c]]++; er[q] = r; ec[q] = c; ev[q] = vx[p]; }
    free( w ); }
  /* level lists */
  int* lcnt = calloc( maxdepth + 2, sizeof( int ) );
  for ( int s = 0; s < ns; s++ ) lcnt[sn_depth[s] + 1]++;
  for ( int d = 0; d <= maxdepth; d++ ) lcnt[d + 1] += lcnt[d];
  int* lorder = malloc( sizeof( int ) * ns );
  { int* w = malloc( sizeof( int ) * ( maxdepth + 2 ) ); memcpy( w, lcnt, sizeof( int ) * ( maxdepth + 2 ) ); for ( int s = 0; s < ns; s++ ) lorder[w[sn_depth[s]]++] = s; free( w ); }
  if ( threads <= 0 ) {
#ifdef _OPENMP
    threads = omp_get_max_threads();
#else
    threads = 1;
#endif
  }
  for ( int d = maxdepth; d >= 0 && !err; d-- ) {
    const int cnt = lcnt[d + 1] - lcnt[d];
    /* many fronts: one thread each with a serial BLAS; few fronts: serial loop with a threaded BLAS */
    const int par = 2 * cnt >= threads;
    if ( f_setthr ) f_setthr( par ? 1 : threads );
    #pragma omp parallel for schedule(dynamic,1) num_threads(threads) if(par)
    for ( int t = 0; t < cnt; t++ ) {
      const int s = lorder[lcnt[d] + t];
      const int c0 = sn_start[s], k = sn_start[s + 1] - c0;
      const int f = ( int )( rowptr[s + 1] - rowptr[s] ), u = f - k;
      const int* rows = rowidx + rowptr[s];
      /* big fronts (serial loop over the level, threaded BLAS): zero-fill, extend-add and copy-out are spread over the
       * threads as well -- left serial they, not the BLAS-3 calls, bound the top of the tree on many cores */
      const int inner = par ? 1 : threads;
      double* F = malloc( ( size_t )f * f * sizeof( double ) );
      #pragma omp parallel for schedule(static) num_threads(inner) if(inner > 1 && f >= 512)
      for ( int j = 0; j < f; j++ ) memset( F + ( size_t )j * f, 0, sizeof( double ) * f );
      int* pos = malloc( sizeof( int ) * f );  /* binary search helper not needed: build a local map via sorted rows */
      /* assemble A */
      for ( int64_t q = acnt[s]; q < acnt[s + 1]; q++ ) {
        int r = er[q], c = ec[q], lr;
        if ( r < c0 + k ) lr = r - c0;
        else { int a = k, b = f - 1; lr = -1; while ( a <= b ) { int m = ( a + b ) >> 1; if ( rows[m] == r ) { lr = m; break; } else if ( rows[m] < r ) a = m + 1; else b = m - 1; } }
        F[lr + ( size_t )( c - c0 ) * f] += ev[q];
      }
      /* extend-add the children */
      for ( int c = chead[s]; c != -1; c = cnext[c] ) {
        const int kc = sn_start[c + 1] - sn_start[c], fc = ( int )( rowptr[c + 1] - rowptr[c] ), uc = fc - kc;
        if ( uc == 0 ) continue;
        const int* rc = rowidx + rowptr[c] + kc;
        int* rel = malloc( sizeof( int ) * uc );
        int a = 0;
        for ( int i = 0; i < uc; i++ ) { while ( rows[a] != rc[i] ) a++; rel[i] = a; }
        const double* C = cb[c];
        #pragma omp parallel for schedule(dynamic,16) num_threads(inner) if(inner > 1 && uc >= 512)
        for ( int j = 0; j < uc; j++ ) { double* dst = F + ( size_t )rel[j] * f; const double* src = C + ( size_t )j * uc; for ( int i = j; i < uc; i++ ) dst[rel[i]] += src[i]; }
        free( rel ); free( cb[c] ); cb[c] = NULL;
      }
      int e = front_factor( F, f, k );
      if ( e ) {
        #pragma omp critical
        err = c0 + e;
      }
      /* keep the panel, pass the contribution block up */
      double* P = Lpanels + lptr[s];
      const size_t ld = ( ( size_t )f + 1 ) & ~( size_t )1;   /* panels use an even leading dimension */
      #pragma omp parallel for schedule(static) num_threads(inner) if(inner > 1 && f >= 512)
      for ( int j = 0; j < k; j++ ) { memcpy( P + ( size_t )j * ld + j, F + ( size_t )j * f + j, sizeof( double ) * ( f - j ) ); for ( int i = 0; i < j; i++ ) P[( size_t )j * ld + i] = 0.0; }
      if ( u > 0 ) {
        double* C = malloc( sizeof( double ) * ( size_t )u * u );
        #pragma omp parallel for schedule(static) num_threads(inner) if(inner > 1 && u >= 512)
        for ( int j = 0; j < u; j++ ) memcpy( C + ( size_t )j * u, F + ( size_t )( k + j ) * f + k, sizeof( double ) * u );
        cb[s] = C;
      }
      free( pos ); free( F );
    }
  }
  for ( int s = 0; s < ns; s++ ) free( cb[s] );
  free( cb ); free( chead ); free( cnext ); free( col2sn ); free( acnt ); free( er ); free( ec ); free( ev ); free( lcnt ); free( lorder );
  return err;
}

/* triangular solves on the panels (one right-hand side, permuted numbering, in place) */
void oracle_mf_solve( int ns, const int* sn_start, const int64_t* rowptr, const int* rowidx, const int64_t* lptr, const double* Lpanels, double* x ) {
  for ( int s = 0; s < ns; s++ ) {
    const int c0 = sn_start[s], k = sn_start[s + 1] - c0, f = ( int )( rowptr[s + 1] - rowptr[s] );
    const int* rows = rowidx + rowptr[s]; const double* P = Lpanels + lptr[s]; const size_t ld = ( ( size_t )f + 1 ) & ~( size_t )1;
    for ( int j = 0; j < k; j++ ) { const double v = x[c0 + j] / P[j + ( size_t )j * ld]; x[c0 + j] = v; for ( int i = j + 1; i < f; i++ ) x[rows[i]] -= P[i + ( size_t )j * ld] * v; }
  }
  for ( int s = ns - 1; s >= 0; s-- ) {
    const int c0 = sn_start[s], k = sn_start[s + 1] - c0, f = ( int )( rowptr[s + 1] - rowptr[s] );
    const int* rows = rowidx + rowptr[s]; const double* P = Lpanels + lptr[s]; const size_t ld = ( ( size_t )f + 1 ) & ~( size_t )1;
    for ( int j = k - 1; j >= 0; j-- ) { double v = x[c0 + j]; for ( int i = j + 1; i < f; i++ ) v -= P[i + ( size_t )j * ld] * x[rows[i]]; x[c0 + j] = v / P[j + ( size_t )j * ld]; }
  }
}
