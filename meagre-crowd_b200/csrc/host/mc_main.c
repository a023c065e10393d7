/* mc_main.c -- drop-in `meagre-crowd` command line driver.
 *
 * Behavioural re-statement of the reference's src/meagre-crowd.c:96-475 and
 * src/args.c:32-205 for the single-process case (there is no MPI in this image
 * and the B200 row is SOLVER_SINGLE_THREADED_ONLY, so main never initialises MPI,
 * src/meagre-crowd.c:151-158): same options, same default right-hand side
 * b[i]=i (0-based, :219-238), same PASS/FAIL and exit codes (0/1/12/100), same
 * `-t` CSV line (:434-446) and `-v` summary (:315-320).
 */
#include <argp.h>
#include <assert.h>
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include "mc_solver.h"

#define MC_PACKAGE_STRING "meagre-crowd 0.4.6"

struct parse_args {
  char* input; char* output; char* rhs; char* expected;
  double expected_precision;
  unsigned int timing_enabled, verbosity, rep;
  int mpi_rank; int solver;
};

const char* argp_program_version = MC_PACKAGE_STRING;

static error_t parse_opt( int key, char* arg, struct argp_state* state ) {
  struct parse_args* args = state->input;
  switch ( key ) {
    case 'h': case '?': argp_state_help( state, state->out_stream, ARGP_HELP_STD_HELP ); break;
    case -1: argp_state_help( state, state->out_stream, ARGP_HELP_SHORT_USAGE | ARGP_HELP_EXIT_OK ); break;
    case 'V': printf( "%s\n", MC_PACKAGE_STRING ); exit( EXIT_SUCCESS ); break;
    case 't': args->timing_enabled++; break;
    case 'v': args->verbosity++; break;
    case -2: printf_solvers( args->verbosity ); exit( EXIT_SUCCESS ); break;
    case 's':
      args->solver = lookup_solver_by_shortname( arg );
      if ( args->solver < 0 ) { fprintf( stderr, "invalid solver (-s)\n" ); exit( EXIT_FAILURE ); }
      break;
    case 'r': { int i = atoi( arg ); args->rep = i < 0 ? 0 : i; } break;
    case 'i': { args->input = arg; FILE* f = fopen( arg, "r" ); if ( f == NULL ) { perror( "input error" ); exit( EXIT_FAILURE ); } fclose( f ); } break;
    case 'b': { args->rhs = arg; FILE* f = fopen( arg, "r" ); if ( f == NULL ) { perror( "right-hand-side error" ); exit( EXIT_FAILURE ); } fclose( f ); } break;
    case 'e': { args->expected = arg; FILE* f = fopen( arg, "r" ); if ( f == NULL ) { perror( "expected-output error" ); exit( EXIT_FAILURE ); } fclose( f ); } break;
    case 'p':
      if ( sscanf( arg, "%lf", &args->expected_precision ) != 1 ) { fprintf( stderr, "bad precision (non-floating point number)\n" ); exit( EXIT_FAILURE ); }
      if ( args->expected_precision < 0 ) { fprintf( stderr, "precision must be non-negative\n" ); exit( EXIT_FAILURE ); }
      break;
    case 'o':
      args->output = arg;
      if ( strncmp( arg, "-", 2 ) != 0 ) { FILE* f = fopen( arg, "r" ); if ( f != NULL ) { fclose( f ); perror( "output error: file exists" ); exit( EXIT_FAILURE ); } }
      break;
    default: return ARGP_ERR_UNKNOWN;
  }
  return 0;
}

static int parse_args( int argc, char** argv, struct parse_args* args ) {
  static char doc[] = "\nSolves Ax=b for x, where A is sparse.\nGiven an input matrix A and right-hand side b, find x.\n\n"
                      "Limitations: only 'Matrix Market' format (*.mm) is supported in this build.\n\nOptions:";
  static const struct argp_option opt[] = {
    { "help", 'h', 0, 0, "Give this help list" },
    { 0, '?', 0, OPTION_ALIAS },
    { "usage", -1, 0, 0, "Show usage information", -1 },
    { "version", 'V', 0, 0, "Show version information", -1 },
    { "input", 'i', "FILE", 0, "Input matrix from FILE (A)", 10 },
    { "right-hand-side", 'b', "FILE", 0, "RHS matrix from FILE (b)", 11 },
    { "expected-answer", 'e', "FILE", 0, "Expected matrix as a FILE (x)", 12 },
    { "precision", 'p', "<float>", 0, "Precision of comparison with expectation (a floating point number)", 12 },
    { "output", 'o', "FILE", 0, "Output matrix to FILE (x) ('-' is stdout)", 13 },
    { "verbose", 'v', 0, 0, "Increase verbosity", 20 },
    { "timing", 't', 0, 0, "Show/increase timing information", 21 },
    { "repeat", 'r', "N", 0, "Repeat calculations N times", 6 },
    { "solver", 's', "SOLVER", 0, "Select SOLVER", 5 },
    { "list-solvers", -2, 0, 0, "List available SOLVERs (-vv for more details)", 5 },
    { 0 }
  };
  const struct argp p = { opt, parse_opt, 0, doc };
  if ( argc == 1 ) {
    char* prog = strndup( argv[0], 100 );
    argp_help( &p, stderr, ARGP_HELP_SHORT_USAGE, basename( prog ) );
    free( prog );
    return EXIT_FAILURE;
  }
  argp_parse( &p, argc, argv, ARGP_NO_HELP, 0, args );
  return EXIT_SUCCESS;
}

int main( int argc, char** argv ) {
  int retval;
  struct parse_args* args = calloc( 1, sizeof( struct parse_args ) );
  assert( args != NULL );
  args->expected_precision = 5e-14;   /* src/meagre-crowd.c:104 */
  if ( ( retval = parse_args( argc, argv, args ) ) != EXIT_SUCCESS ) { free( args ); return retval; }
  perftimer_t* timer = perftimer_malloc();
  const int c_mpi = 0, c_omp = 0;      /* not launched by mpirun: "threads" column prints 0 (src/meagre-crowd.c:127-130) */

  matrix_t* b = malloc_matrix(); matrix_t* expected = malloc_matrix(); matrix_t* A = malloc_matrix(); matrix_t* rhs = malloc_matrix();
  if ( ( retval = load_matrix( args->input, A ) ) != 0 ) return 1;
  assert( validate_matrix( A ) == 0 );
  const unsigned int m = matrix_rows( A );
  if ( args->rhs != NULL ) {
    if ( ( retval = load_matrix( args->rhs, b ) ) != 0 ) return 1;
    assert( b->m == m );
  } else {
    *b = ( matrix_t ) { 0 };
    b->m = m; b->n = 1; b->nz = m; b->format = DCOL; b->data_type = REAL_DOUBLE;
    b->dd = malloc( m * sizeof( double ) );
    double* d = b->dd;
    for ( unsigned int i = 0; i < m; i++ ) d[i] = i;
  }
  assert( validate_matrix( b ) == 0 );
  if ( args->expected != NULL ) {
    if ( ( retval = load_matrix( args->expected, expected ) ) != 0 ) return 1;
    assert( validate_matrix( expected ) == 0 );
    assert( expected->m == m );
  }
  if ( args->verbosity >= 1 ) {
    int ierr = convert_matrix( A, SM_COO, FIRST_INDEX_ZERO ); assert( ierr == 0 ); ( void )ierr;
    const char* sym = A->sym == SM_UNSYMMETRIC ? "unsymmetric" : A->sym == SM_SYMMETRIC ? "symmetric" : A->sym == SM_SKEW_SYMMETRIC ? "skew symmetric" : "hermitian";
    /* src/meagre-crowd.c:277-295: the storage suffix only at -vv, never for MC_STORE_BOTH */
    const char* loc = ( args->verbosity < 2 || A->sym == SM_UNSYMMETRIC ) ? "" : A->location == UPPER_TRIANGULAR ? " (upper)" : A->location == LOWER_TRIANGULAR ? " (lower)" : "";
    printf( "Ax=b: A is %zux%zu, nz=%zu, %s%s, %s, b is %zux%zu, nz=%zu\nsolved with %s on %d core%s, %d thread%s\n",
            A->m, A->n, A->nz, sym, loc, "real", b->m, b->n, b->nz, solver2str( args->solver ), c_mpi, c_mpi == 1 ? "" : "s", c_omp, c_omp == 1 ? "" : "s" );
    if ( args->verbosity >= 2 ) { printf_matrix( "  A", A ); printf_matrix( "  b", b ); }
  }

  solver_state_t* state = solver_init( args->solver, args->verbosity, args->mpi_rank, timer );
  unsigned int r = 0;
  do {
    if ( r != 0 ) perftimer_restart( &timer );
    solver_solve( state, A, b, rhs );
    r++;
  } while ( r < args->rep );

  if ( args->output != NULL ) {
    if ( strncmp( args->output, "-", 2 ) == 0 ) printf_matrix( "  x", rhs );
    else if ( save_matrix( rhs, args->output ) != 0 ) retval = 12;
  }
  if ( expected->format != INVALID ) {
    perftimer_inc( timer, "test", -1 );
    if ( results_match( expected, rhs, args->expected_precision ) ) { retval = 0; printf( "PASS\n" ); }
    else { retval = 100; printf( "FAIL\n" ); }
  }
  solver_finalize( state );

  struct rusage usage;
  int rr = getrusage( RUSAGE_SELF, &usage );
  if ( rr == 0 && args->verbosity >= 2 )
    printf( "memory usage: maximum resident=%lgMB, page reclaims=%ld, page faults w/ IO=%ld\n", usage.ru_maxrss / 1e3, usage.ru_minflt, usage.ru_majflt );
  const double mem_sum = usage.ru_maxrss / 1e3;
  if ( args->timing_enabled == 1 ) {
    printf( "status, solver, threads, rows, max. memory (MB),  total memory (MB), " );
    perftimer_printf_csv_header( timer, 2 );
    printf( "%s, %s, %d, %zd, %lg, %lg, ", retval == 100 ? "FAIL" : "PASS", solver2str( args->solver ), c_mpi, A->m, usage.ru_maxrss / 1e3, mem_sum );
    perftimer_printf_csv_body( timer, 2 );
  } else if ( args->timing_enabled != 0 ) perftimer_printf( timer, args->timing_enabled - 2 );

  free_matrix( b ); free_matrix( expected ); free_matrix( A ); free_matrix( rhs );
  perftimer_free( timer );
  free( args );
  return retval;
}
