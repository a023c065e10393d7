/* mc_solvers.c -- the solver plugin table and its dispatcher.
 *
 * Re-statement of the reference's src/solvers.c:
 *   _convert_matrix_A   :35-180   A is mutated IN PLACE to a (symmetry storage,
 *                                 format, base) the selected row advertises
 *   _convert_matrix_b   :183-318  same for the right-hand side
 *   lookup / listing    :326-379
 *   solver(), solver_solve, init/finalize, analyze/factorize/evaluate :420-611
 * and of the table in src/solver_lookup.h:123-339, which here has ONE row: the
 * B200 backend (the third-party rows are #ifdef HAVE_X in the reference and none
 * of those libraries exists in this image).  There is no MPI in the image, so the
 * single-process branch of the dispatcher (is_mpi == 0) is the only one.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mc_solver.h"
#include "solver_b200.h"

const struct solver_properties_t solver_lookup[] = {
  { "b200", "B200-multifrontal", "meagre-crowd_b200", "this repository", "0.1", "GPL",
    "https://github.com/boyle/meagre-crowd (backend for)",
    &solver_init_b200,
    &solver_analyze_b200,
    &solver_factorize_b200,
    &solver_evaluate_b200,
    &solver_finalize_b200,
    SOLVER_B200_CAPABILITIES,
    SOLVER_SINGLE_THREADED_ONLY, /* one process drives the GPUs; main never calls MPI_Init */
    "    * J. W. H. Liu, The role of elimination trees in sparse factorization, SIMAX 11 (1990).\n"
    "    * J. R. Gilbert, E. G. Ng, B. W. Peyton, An efficient algorithm to compute row and\n"
    "      column counts for sparse Cholesky factorization, SIMAX 15 (1994).\n"
    "    * I. S. Duff, J. K. Reid, The multifrontal solution of indefinite sparse symmetric\n"
    "      linear equations, ACM TOMS 9 (1983).\n"
    "    * A. George, Nested dissection of a regular finite element mesh, SINUM 10 (1973).\n" },
  { 0 }
};

static int convert_A_for( const int solver, matrix_t* A ) {
  const unsigned int c = solver_lookup[solver].capabilities;
  int ierr;
  switch ( A->sym ) {
    case SM_UNSYMMETRIC:
      assert( c & SOLVES_UNSYMMETRIC );
      break;
    case SM_SYMMETRIC:
      if ( c & SOLVES_SYMMETRIC ) {
        /* keep the current storage when the backend takes it; otherwise prefer the
         * cheaper triangle swap before expanding to both triangles */
        enum matrix_symmetric_storage_t want = A->location;
        const unsigned int has_up = c & SOLVES_SYMMETRIC_UPPER_TRIANGULAR, has_lo = c & SOLVES_SYMMETRIC_LOWER_TRIANGULAR, has_both = c & SOLVES_SYMMETRIC_BOTH;
        switch ( A->location ) {
          case MC_STORE_BOTH:    if ( !has_both ) want = has_up ? UPPER_TRIANGULAR : LOWER_TRIANGULAR; break;
          case UPPER_TRIANGULAR: if ( !has_up ) want = has_lo ? LOWER_TRIANGULAR : MC_STORE_BOTH; break;
          case LOWER_TRIANGULAR: if ( !has_lo ) want = has_up ? UPPER_TRIANGULAR : MC_STORE_BOTH; break;
        }
        if ( want != A->location ) { ierr = convert_matrix_symmetry( A, want ); assert( ierr == 0 ); }
      } else if ( c & SOLVES_UNSYMMETRIC ) {
        ierr = convert_matrix_symmetry( A, MC_STORE_BOTH ); assert( ierr == 0 );
        A->sym = SM_UNSYMMETRIC;
      } else assert( 0 );
      break;
    default:
      assert( 0 ); /* skew-symmetric / hermitian: not handled by the reference either (src/solvers.c:106-111) */
  }
  enum matrix_base_t base = A->base;
  if ( base == FIRST_INDEX_ZERO && !( c & SOLVES_BASE_ZERO ) ) base = FIRST_INDEX_ONE;
  else if ( base == FIRST_INDEX_ONE && !( c & SOLVES_BASE_ONE ) ) base = FIRST_INDEX_ZERO;
  enum matrix_format_t format = A->format;
  assert( format != INVALID );
  if ( format == DROW || format == DCOL ) format = SM_COO;
  if ( format == SM_COO && !( c & SOLVES_FORMAT_COO ) ) format = ( c & SOLVES_FORMAT_CSR ) ? SM_CSR : SM_CSC;
  else if ( format == SM_CSC && !( c & SOLVES_FORMAT_CSC ) ) format = ( c & SOLVES_FORMAT_COO ) ? SM_COO : SM_CSR;
  else if ( format == SM_CSR && !( c & SOLVES_FORMAT_CSR ) ) format = ( c & SOLVES_FORMAT_COO ) ? SM_COO : SM_CSC;
  ierr = convert_matrix( A, format, base );
  assert( ierr == 0 );
  ( void )ierr;
  return 0;
}

static int convert_b_for( const int solver, matrix_t* b ) {
  const unsigned int c = solver_lookup[solver].capabilities;
  int ierr;
  if ( b->sym != SM_UNSYMMETRIC ) { ierr = convert_matrix_symmetry( b, MC_STORE_BOTH ); assert( ierr == 0 ); }
  assert( c & ( SOLVES_BASE_ZERO | SOLVES_BASE_ONE ) );
  enum matrix_base_t base = b->base;
  if ( base == FIRST_INDEX_ZERO && !( c & SOLVES_BASE_ZERO ) ) base = FIRST_INDEX_ONE;
  else if ( base == FIRST_INDEX_ONE && !( c & SOLVES_BASE_ONE ) ) base = FIRST_INDEX_ZERO;
  assert( c & ( SOLVES_RHS_DROW | SOLVES_RHS_DCOL | SOLVES_RHS_COO | SOLVES_RHS_CSC | SOLVES_RHS_CSR ) );
  /* preference orders follow src/solvers.c:222-313 */
  static const struct { enum matrix_format_t f; unsigned int bit; } all[5] = {
    { DROW, SOLVES_RHS_DROW }, { DCOL, SOLVES_RHS_DCOL }, { SM_COO, SOLVES_RHS_COO }, { SM_CSC, SOLVES_RHS_CSC }, { SM_CSR, SOLVES_RHS_CSR } };
  static const int pref[6][4] = {
    { 0, 0, 0, 0 },
    /* DROW */ { 1, 2, 3, 4 }, /* DCOL */ { 0, 2, 3, 4 }, /* COO */ { 3, 4, 0, 1 }, /* CSC */ { 2, 4, 0, 1 }, /* CSR */ { 2, 3, 0, 1 } };
  enum matrix_format_t format = b->format;
  assert( format != INVALID );
  int self = format == DROW ? 0 : format == DCOL ? 1 : format == SM_COO ? 2 : format == SM_CSC ? 3 : 4;
  if ( !( c & all[self].bit ) ) {
    int k;
    for ( k = 0; k < 4; k++ ) if ( c & all[pref[format][k]].bit ) { format = all[pref[format][k]].f; break; }
    assert( k < 4 );
  }
  ierr = convert_matrix( b, format, base );
  assert( ierr == 0 );
  ( void )ierr;
  return 0;
}

static int valid_solver( const int solver ) { return solver >= 0 && solver_lookup[solver].shortname != NULL; }

int lookup_solver_by_shortname( const char* shortname ) {
  assert( shortname != NULL );
  for ( int i = 0; solver_lookup[i].shortname != NULL; i++ )
    if ( strncmp( solver_lookup[i].shortname, shortname, SOLVER_SHORTNAME_MAX_LEN ) == 0 ) return i;
  return -1;
}

const char* solver2str( const int solver ) { return valid_solver( solver ) ? solver_lookup[solver].name : "<invalid>"; }

void printf_solvers( const unsigned int verbosity ) {
  printf( "Available solvers:\n" );
  size_t w = 0;
  for ( int i = 0; solver_lookup[i].shortname != NULL; i++ ) { size_t l = strlen( solver_lookup[i].shortname ); if ( l > w ) w = l; }
  for ( int i = 0; solver_lookup[i].shortname != NULL; i++ ) {
    const struct solver_properties_t* p = &solver_lookup[i];
    printf( "  %-*s    %s %s (%s, %s)\n", ( int )w, p->shortname, p->name, p->version, p->author, p->license );
    if ( verbosity >= 1 ) printf( "%*s      %s, %s\n", ( int )w, "", p->organization, p->url );
    if ( verbosity >= 2 ) printf( "    References:\n%s\n\n", p->references );
  }
}

int solver_can_do( const int solver, matrix_t* A, matrix_t* b ) { ( void )solver; ( void )A; ( void )b; return 1; }
int solver_uses_mpi( const int solver ) { return ( solver_lookup[solver].multicore & SOLVER_CAN_USE_MPI ) != 0; }
int solver_requires_mpi( const int solver ) { return ( solver_lookup[solver].multicore & SOLVER_REQUIRES_MPI ) != 0; }
int solver_uses_omp( const int solver ) { return ( solver_lookup[solver].multicore & SOLVER_CAN_USE_OMP ) != 0; }
int solver_requires_omp( const int solver ) { return ( solver_lookup[solver].multicore & SOLVER_REQUIRES_OMP ) != 0; }
int select_solver( matrix_t* A, matrix_t* b ) { ( void )A; ( void )b; return 0; }

void solver( const int solver_id, const int verbosity, const int mpi_rank, matrix_t* A, matrix_t* b, matrix_t* x ) {
  solver_state_t* s = solver_init( solver_id, verbosity, mpi_rank, NULL );
  assert( s != NULL );
  solver_analyze( s, A );
  solver_factorize( s, A );
  solver_evaluate( s, b, x );
  solver_finalize( s );
}

void solver_solve( solver_state_t* s, matrix_t* A, matrix_t* b, matrix_t* x ) {
  solver_analyze( s, A );
  solver_factorize( s, A );
  solver_evaluate( s, b, x );
}

solver_state_t* solver_init( const int solver_id, const int verbosity, const int mpi_rank, perftimer_t* timer ) {
  solver_state_t* s = malloc( sizeof( solver_state_t ) );
  assert( s != NULL );
  s->solver = solver_id; s->verbosity = verbosity; s->mpi_rank = mpi_rank; s->timer = timer; s->specific = NULL;
  if ( valid_solver( solver_id ) && solver_lookup[solver_id].init != NULL ) solver_lookup[solver_id].init( s );
  return s;
}

void solver_finalize( solver_state_t* s ) {
  assert( s != NULL );
  const int id = s->solver;
  if ( valid_solver( id ) && solver_lookup[id].analyze != NULL ) solver_lookup[id].finalize( s ); /* sic: tests .analyze (src/solvers.c:461) */
  s->specific = NULL;
  free( s );
}

void solver_analyze( solver_state_t* s, matrix_t* A ) {
  assert( s != NULL );
  const int id = s->solver;
  if ( s->mpi_rank == 0 ) { assert( A != NULL ); convert_A_for( id, A ); }
  perftimer_inc( s->timer, "analyze", -1 );
  if ( valid_solver( id ) && solver_lookup[id].analyze != NULL ) solver_lookup[id].analyze( s, A );
}

void solver_factorize( solver_state_t* s, matrix_t* A ) {
  assert( s != NULL );
  const int id = s->solver;
  if ( s->mpi_rank == 0 ) { assert( A != NULL ); convert_A_for( id, A ); }
  perftimer_inc( s->timer, "factorize", -1 );
  if ( valid_solver( id ) && solver_lookup[id].factorize != NULL ) solver_lookup[id].factorize( s, A );
}

void solver_evaluate( solver_state_t* s, matrix_t* b, matrix_t* x ) {
  assert( s != NULL );
  const int id = s->solver;
  const unsigned int c = solver_lookup[id].capabilities;
  matrix_t bb = {0};
  int loops = 1;
  if ( s->mpi_rank == 0 ) {
    assert( b != NULL ); assert( x != NULL );
    if ( b->n != 1 && ( c & SOLVES_RHS_VECTOR_ONLY ) ) {
      /* one column at a time through a shallow copy whose dd is advanced (src/solvers.c:516-531) */
      assert( c & SOLVES_RHS_DCOL );
      convert_matrix( b, DCOL, FIRST_INDEX_ZERO );
      bb = *b; bb.n = 1; bb.nz = bb.m; loops = ( int )b->n;
    } else { convert_b_for( id, b ); bb = *b; loops = 1; }
  }
  perftimer_inc( s->timer, "evaluate", -1 );
  if ( valid_solver( id ) && solver_lookup[id].evaluate != NULL ) {
    if ( s->mpi_rank == 0 ) { clear_matrix( x ); x->format = DCOL; }
    for ( int i = 0; i < loops; i++ ) {
      matrix_t xx = {0};
      solver_lookup[id].evaluate( s, s->mpi_rank == 0 ? &bb : NULL, s->mpi_rank == 0 ? &xx : NULL );
      if ( s->mpi_rank != 0 ) continue;
      if ( i != loops - 1 ) { double* dd = bb.dd; bb.dd = dd + bb.m; }
      if ( x->dd == NULL ) { *x = xx; xx.dd = NULL; memset( &xx, 0, sizeof( xx ) ); }
      else {
        /* append the new column(s); the reference's memcpy scales the destination
         * offset by sizeof(double) twice (src/solvers.c:597) -- here the intended
         * column-major append is done */
        convert_matrix( &xx, DCOL, FIRST_INDEX_ZERO );
        x->m = xx.m;
        x->dd = realloc( x->dd, x->m * ( x->n + xx.n ) * sizeof( double ) );
        assert( x->dd != NULL );
        memcpy( ( double* )x->dd + x->m * x->n, xx.dd, x->m * xx.n * sizeof( double ) );
        x->n += xx.n; x->nz = x->m * x->n;
        clear_matrix( &xx );
      }
    }
  } else { clear_matrix( x ); x->format = INVALID; }
  perftimer_inc( s->timer, "done", -1 );
}

/* src/meagre-crowd.c:51-83 */
int results_match( matrix_t* expected_matrix, matrix_t* result_matrix, const double precision ) {
  assert( expected_matrix != NULL && result_matrix != NULL );
  assert( expected_matrix->format != INVALID && result_matrix->format != INVALID );
  assert( expected_matrix->data_type == REAL_DOUBLE && result_matrix->data_type == REAL_DOUBLE );
  assert( result_matrix->m == expected_matrix->m && result_matrix->n == expected_matrix->n );
  int ret = convert_matrix( expected_matrix, DCOL, FIRST_INDEX_ZERO ); assert( ret == 0 );
  ret = convert_matrix( result_matrix, DCOL, FIRST_INDEX_ZERO ); assert( ret == 0 );
  ( void )ret;
  const double* e = expected_matrix->dd; const double* r = result_matrix->dd;
  const size_t tot = result_matrix->m * result_matrix->n;
  for ( size_t k = 0; k < tot; k++ ) if ( r[k] < e[k] - precision || r[k] > e[k] + precision || r[k] != r[k] ) return 0;
  return 1;
}
