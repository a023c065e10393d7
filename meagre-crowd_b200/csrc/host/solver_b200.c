/* solver_b200.c -- the plugin wrapper: five callbacks behind solver_lookup[].
 *
 * Same contract as the reference's wrappers (src/solver_cholmod.c:33-342,
 * src/solver_umfpack.c:33-151):
 *   - init allocates s->specific; finalize releases it (src/solvers.c:451,461)
 *   - A arrives as CSC, base 0; symmetric input as SM_SYMMETRIC/UPPER_TRIANGULAR
 *     (read zero-copy exactly like _matrix2cholmodsparse, src/solver_cholmod.c:46-105:
 *     colptr = A->jj, rowidx = A->ii, values = A->dd), unsymmetric input as full CSC
 *   - b arrives DCOL m x p; x is returned as DCOL with a malloc'd dd that the
 *     caller frees (src/solver_umfpack.c:122-131, src/solvers.c:585-601)
 *   - errors: message on stderr + abort(), like the reference's assert()s
 *   - analyze/factorize/evaluate may be repeated on one state (-r N)
 * Tuning comes from the environment because the callbacks carry no options:
 *   B200_DEVICE (default 0), B200_ND_LEAF, B200_NO_RELAX, B200_REFINE (steps; default automatic), B200_VERBOSE,
 *   B200_SYM_KIND = cholesky | ldlt | lu  (symmetric input: default Cholesky, L D L' with Bunch-Kaufman pivots when a pivot is not positive).
 * There is no CPU fallback: without a usable GPU, init aborts.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "solver_b200.h"

typedef struct {
  b200_ctx_t* ctx;
  b200_symbolic_t* sym;
  int kind;               /* kind of the current analysis */
  int indefinite;         /* a symmetric input turned out not to be positive definite: later analyses of this state go straight to L D L' */
  /* full-matrix expansion used when a symmetric (upper) input goes through LU */
  unsigned int* fcolptr; unsigned int* frowidx; double* fvals; size_t* fsrc; size_t fnz;
  size_t n;
} b200_state_t;

static void die( const char* what ) {
  fprintf( stderr, "b200 solver: %s: %s\n", what, b200_last_error() );
  abort();
}

void solver_init_b200( solver_state_t* s ) {
  assert( s != NULL );
  b200_state_t* p = calloc( 1, sizeof( b200_state_t ) );
  assert( p != NULL );
  s->specific = p;
  if ( s->mpi_rank != 0 ) return;
  const char* dev = getenv( "B200_DEVICE" );
  p->ctx = b200_ctx_create( dev ? atoi( dev ) : 0 );
  if ( p->ctx == NULL ) die( "cannot create a CUDA context (no CPU fallback exists)" );
  const char* rf = getenv( "B200_REFINE" );
  if ( rf ) b200_set_refinement( p->ctx, atoi( rf ) );
}

static void free_expansion( b200_state_t* p ) {
  free( p->fcolptr ); free( p->frowidx ); free( p->fvals ); free( p->fsrc );
  p->fcolptr = p->frowidx = NULL; p->fvals = NULL; p->fsrc = NULL; p->fnz = 0;
}

/* upper-triangular CSC -> full CSC; fsrc[q] = index of the source entry */
static void expand_symmetric( b200_state_t* p, const matrix_t* A ) {
  free_expansion( p );
  const size_t n = A->n;
  unsigned int* cnt = calloc( n + 1, sizeof( unsigned int ) );
  for ( size_t j = 0; j < n; j++ ) for ( unsigned int q = A->jj[j]; q < A->jj[j + 1]; q++ ) { cnt[j + 1]++; if ( A->ii[q] != j ) cnt[A->ii[q] + 1]++; }
  for ( size_t j = 0; j < n; j++ ) cnt[j + 1] += cnt[j];
  p->fnz = cnt[n];
  p->fcolptr = malloc( sizeof( unsigned int ) * ( n + 1 ) ); memcpy( p->fcolptr, cnt, sizeof( unsigned int ) * ( n + 1 ) );
  p->frowidx = malloc( sizeof( unsigned int ) * ( p->fnz ? p->fnz : 1 ) );
  p->fsrc = malloc( sizeof( size_t ) * ( p->fnz ? p->fnz : 1 ) );
  p->fvals = malloc( sizeof( double ) * ( p->fnz ? p->fnz : 1 ) );
  for ( size_t j = 0; j < n; j++ ) for ( unsigned int q = A->jj[j]; q < A->jj[j + 1]; q++ ) {
      unsigned int i = A->ii[q];
      unsigned int w = cnt[j]++; p->frowidx[w] = i; p->fsrc[w] = q;
      if ( i != j ) { w = cnt[i]++; p->frowidx[w] = ( unsigned int )j; p->fsrc[w] = q; }
    }
  free( cnt );
}

static void analyze_kind( solver_state_t* s, matrix_t* A, int kind ) {
  b200_state_t* p = s->specific;
  b200_options_t o; b200_default_options( &o );
  const char* e;
  if ( ( e = getenv( "B200_ND_LEAF" ) ) ) o.nd_leaf = atoi( e );
  if ( getenv( "B200_NO_RELAX" ) ) o.relax = 0;
  if ( p->sym ) { b200_symbolic_free( p->sym ); p->sym = NULL; }   /* repeated analyze releases the previous symbolic */
  const unsigned int* cp = A->jj; const unsigned int* ri = A->ii;
  if ( kind == B200_KIND_LU && A->sym == SM_SYMMETRIC ) { expand_symmetric( p, A ); cp = p->fcolptr; ri = p->frowidx; }
  p->sym = b200_analyze( ( int )A->n, cp, ri, kind, &o );
  if ( p->sym == NULL ) { fprintf( stderr, "b200 solver: symbolic analysis failed\n" ); abort(); }
  p->kind = kind; p->n = A->n;
  if ( b200_upload_symbolic( p->ctx, p->sym ) != B200_OK ) die( "upload of the symbolic structure failed" );
  if ( s->verbosity >= 3 || getenv( "B200_VERBOSE" ) )
    fprintf( stderr, "b200: n=%d kind=%s nnz(L)=%lld (stored %lld) flops=%.4g supernodes=%d (fundamental %d) levels=%d maxfront=%d order=%.3fs symbolic=%.3fs\n",
             p->sym->n, kind == B200_KIND_LU ? "LU" : kind == B200_KIND_LDLT ? "LDLT" : "Cholesky", ( long long )p->sym->nnzL, ( long long )p->sym->stored_nnzL,
             p->sym->flops * ( kind == B200_KIND_LU ? 2 : 1 ), p->sym->ns, p->sym->nfund, p->sym->maxdepth + 1, p->sym->maxfront, p->sym->t_order, p->sym->t_symbolic );
}

void solver_analyze_b200( solver_state_t* s, matrix_t* A ) {
  assert( s != NULL );
  if ( s->mpi_rank != 0 ) return;
  assert( A != NULL );
  b200_state_t* p = s->specific;
  assert( p != NULL );
  assert( validate_matrix( A ) == 0 );
  assert( A->format == SM_CSC && A->base == FIRST_INDEX_ZERO && A->data_type == REAL_DOUBLE );
  assert( A->m == A->n );
  assert( A->sym == SM_UNSYMMETRIC || ( A->sym == SM_SYMMETRIC && A->location == UPPER_TRIANGULAR ) );
  int kind = A->sym == SM_SYMMETRIC ? ( p->indefinite ? B200_KIND_LDLT : B200_KIND_CHOLESKY ) : B200_KIND_LU;
  const char* sk = getenv( "B200_SYM_KIND" );
  if ( sk && A->sym == SM_SYMMETRIC ) kind = strcmp( sk, "ldlt" ) == 0 ? B200_KIND_LDLT : strcmp( sk, "lu" ) == 0 ? B200_KIND_LU : B200_KIND_CHOLESKY;
  analyze_kind( s, A, kind );
}

void solver_factorize_b200( solver_state_t* s, matrix_t* A ) {
  assert( s != NULL );
  if ( s->mpi_rank != 0 ) return;
  assert( A != NULL );
  b200_state_t* p = s->specific;
  assert( p != NULL && p->sym != NULL );
  assert( A->format == SM_CSC && A->base == FIRST_INDEX_ZERO );
  int r;
  if ( p->kind == B200_KIND_CHOLESKY ) {
    r = b200_factor_host( p->ctx, A->dd );
    if ( r == B200_ERR_NOT_SPD ) {
      /* symmetric but not positive definite (tests/sym.mm, which the reference runs on its general backends, testsuite.at:325-329):
       * the SAME analysis serves L D L' -- only the numeric schedule is rebuilt, with pivoting diagonal blocks and the second panel */
      if ( s->verbosity >= 3 || getenv( "B200_VERBOSE" ) ) fprintf( stderr, "b200: %s -- switching to LDL'\n", b200_last_error() );
      p->indefinite = 1;
      p->sym->kind = B200_KIND_LDLT; p->kind = B200_KIND_LDLT;
      if ( b200_upload_symbolic( p->ctx, p->sym ) != B200_OK ) die( "upload of the symbolic structure failed" );
    } else if ( r != B200_OK ) die( "factorization failed" );
    else return;
  }
  if ( p->kind == B200_KIND_LDLT ) {
    r = b200_factor_host( p->ctx, A->dd );
    if ( r != B200_OK ) die( "LDL' factorization failed" );
    return;
  }
  const double* vals = A->dd;
  if ( A->sym == SM_SYMMETRIC ) { /* LU of a symmetric input: values follow the expansion */
    const double* d = A->dd;
    for ( size_t q = 0; q < p->fnz; q++ ) p->fvals[q] = d[p->fsrc[q]];
    vals = p->fvals;
  }
  r = b200_factor_host( p->ctx, vals );
  if ( r != B200_OK ) die( "LU factorization failed" );
}

void solver_evaluate_b200( solver_state_t* s, matrix_t* b, matrix_t* x ) {
  assert( s != NULL );
  if ( s->mpi_rank != 0 ) return;
  assert( b != NULL && x != NULL && b != x );
  b200_state_t* p = s->specific;
  assert( p != NULL && p->sym != NULL );
  int ierr = convert_matrix( b, DCOL, FIRST_INDEX_ZERO );
  assert( ierr == 0 ); ( void )ierr;
  assert( b->data_type == REAL_DOUBLE );
  assert( b->m == p->n );
  const size_t np = b->n;
  double* xd = malloc( sizeof( double ) * p->n * ( np ? np : 1 ) );   /* ownership passes to the caller */
  assert( xd != NULL );
  if ( b200_solve_host( p->ctx, b->dd, ( int )np, xd ) != B200_OK ) die( "solve failed" );
  { /* statically perturbed pivots are only acceptable if refinement recovered a backward-stable answer: the reference
     * aborts on a singular factorization (assert( status == UMFPACK_OK ), src/solver_umfpack.c:98-101) */
    b200_stats_t st;
    if ( b200_get_stats( p->ctx, &st ) == B200_OK && st.npiv_perturbed > 0 ) {
      if ( st.backward_error > 1e-8 ) { fprintf( stderr, "b200 solver: %d perturbed pivots and componentwise backward error %.2e after refinement: matrix is numerically singular\n", st.npiv_perturbed, st.backward_error ); abort(); }
      if ( s->verbosity >= 1 ) fprintf( stderr, "b200 solver: warning: %d pivots were perturbed; backward error after %d refinement steps %.2e\n", st.npiv_perturbed, st.refine_steps, st.backward_error );
    }
  }
  clear_matrix( x );
  x->m = p->n; x->n = np; x->nz = p->n * np;
  x->format = DCOL; x->base = FIRST_INDEX_ZERO; x->sym = SM_UNSYMMETRIC; x->location = MC_STORE_BOTH; x->data_type = REAL_DOUBLE;
  x->dd = xd; x->ii = NULL; x->jj = NULL;
}

void solver_finalize_b200( solver_state_t* s ) {
  if ( s == NULL ) return;
  b200_state_t* p = s->specific;
  if ( p != NULL ) {
    if ( p->sym ) b200_symbolic_free( p->sym );
    if ( p->ctx ) b200_ctx_destroy( p->ctx );
    free_expansion( p );
  }
  free( p );
}
