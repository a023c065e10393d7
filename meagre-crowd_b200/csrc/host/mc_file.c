/* mc_file.c -- MatrixMarket (.mm) input/output for the drop-in CLI.
 *
 * Behavioural re-statement of the reference's src/file.c:
 *   load_matrix / save_matrix and their stderr texts       src/file.c:640-694, 739-784
 *   readmm grammar (banner, comments, size line, data)     src/file.c:132-262, 425-630
 *   writer: x -> COO base 0 -> "%d %d %.13e" one-based      src/file.c:697-726 + the six
 *           golden files tests/{unsym,sym,sym_posdef}-{default,rhs1}-ans.mm
 * Symmetric files are stored as the lower triangle (src/file.c:247-248);
 * "general" files go through detect_matrix_symmetry (src/file.c:664-665).
 */
#include <stdio.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <ctype.h>
#include <errno.h>
#include <assert.h>
#include "mc_matrix.h"

#define MM_LINE_MAX 1024

/* tokenisers: strict like src/file.c:584-630 -- indices are digit strings,
 * values start with a digit or sign and contain only [0-9.eE+-] */
static int tok_uint( const char** s, long* out ) {
  const char* p = *s;
  while ( *p == ' ' || *p == '\t' ) p++;
  if ( !isdigit( ( unsigned char )*p ) ) return -1;
  long v = 0;
  while ( isdigit( ( unsigned char )*p ) ) { v = v * 10 + ( *p - '0' ); p++; }
  if ( *p != '\0' && !isspace( ( unsigned char )*p ) ) return -1;
  *out = v; *s = p;
  return 0;
}
static int tok_real( const char** s, double* out ) {
  const char* p = *s;
  while ( *p == ' ' || *p == '\t' ) p++;
  if ( !( isdigit( ( unsigned char )*p ) || *p == '+' || *p == '-' ) ) return -1;
  const char* q = p;
  while ( *q && ( isdigit( ( unsigned char )*q ) || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-' ) ) q++;
  if ( *q != '\0' && !isspace( ( unsigned char )*q ) ) return -1;
  char buf[128]; size_t l = ( size_t )( q - p );
  if ( l == 0 || l >= sizeof( buf ) ) return -1;
  memcpy( buf, p, l ); buf[l] = 0;
  char* end; errno = 0;
  double v = strtod( buf, &end );
  if ( *end != 0 || errno == ERANGE ) return -1;
  *out = v; *s = q;
  return 0;
}
static int rest_blank( const char* p ) { while ( *p ) { if ( !isspace( ( unsigned char )*p ) ) return 0; p++; } return 1; }

/* returns 0 on success; negative codes as readmm (src/file.c:96-128) */
static int readmm( const char* filename, matrix_t* A ) {
  FILE* f = fopen( filename, "r" );
  if ( f == NULL ) return -2;
  char line[MM_LINE_MAX + 2];
  if ( fgets( line, sizeof( line ), f ) == NULL ) { fclose( f ); return -11; }
  if ( strlen( line ) > MM_LINE_MAX ) { fclose( f ); return -13; }
  if ( strncasecmp( line, "%%MatrixMarket", 14 ) != 0 || !isspace( ( unsigned char )line[14] ) ) { fclose( f ); return -12; }
  char obj[64] = "", fmt[64] = "", dt[64] = "", sy[64] = "";
  if ( sscanf( line + 14, "%63s %63s %63s %63s", obj, fmt, dt, sy ) != 4 ) { fclose( f ); return -12; }
  if ( strcasecmp( obj, "matrix" ) != 0 ) { fclose( f ); return -3; }
  int coord;
  if ( strcasecmp( fmt, "coordinate" ) == 0 ) coord = 1; else if ( strcasecmp( fmt, "array" ) == 0 ) coord = 0; else { fclose( f ); return -4; }
  if ( strcasecmp( dt, "real" ) != 0 ) { fclose( f ); return -5; } /* complex/integer/pattern: not handled by this build */
  enum matrix_symmetry_t sym;
  if ( strcasecmp( sy, "general" ) == 0 ) sym = SM_UNSYMMETRIC;
  else if ( strcasecmp( sy, "symmetric" ) == 0 ) sym = SM_SYMMETRIC;
  else if ( strcasecmp( sy, "skew-symmetric" ) == 0 ) sym = SM_SKEW_SYMMETRIC;
  else if ( strcasecmp( sy, "hermitian" ) == 0 ) sym = SM_HERMITIAN;
  else { fclose( f ); return -12; }
  /* comments / blank lines are allowed only before the size line */
  long rows = 0, cols = 0, nz = 0;
  for ( ;; ) {
    if ( fgets( line, sizeof( line ), f ) == NULL ) { fclose( f ); return -11; }
    if ( strlen( line ) > MM_LINE_MAX ) { fclose( f ); return -13; }
    if ( line[0] == '%' || rest_blank( line ) ) continue;
    const char* p = line;
    if ( tok_uint( &p, &rows ) || tok_uint( &p, &cols ) ) { fclose( f ); return -12; }
    if ( coord ) { if ( tok_uint( &p, &nz ) ) { fclose( f ); return -12; } }
    else {
      if ( rows > 0 && cols > LONG_MAX / rows ) { fclose( f ); return -12; }     /* rows*cols must not wrap */
      nz = rows * cols;
    }
    if ( !rest_blank( p ) ) { fclose( f ); return -12; }
    if ( rows > ( long )UINT_MAX || cols > ( long )UINT_MAX || ( unsigned long )nz > SIZE_MAX / sizeof( double ) ) { fclose( f ); return -12; }
    break;
  }
  clear_matrix( A );
  A->m = ( size_t )rows; A->n = ( size_t )cols; A->nz = ( size_t )nz;
  A->base = FIRST_INDEX_ONE;
  A->location = LOWER_TRIANGULAR;
  A->data_type = REAL_DOUBLE;
  A->sym = sym;
  double* dd = malloc( sizeof( double ) * ( nz ? nz : 1 ) );
  if ( dd == NULL ) { fclose( f ); clear_matrix( A ); return -12; }
  int ret = 0;
  if ( coord ) {
    A->format = SM_COO;
    unsigned int* ii = malloc( sizeof( unsigned int ) * ( nz ? nz : 1 ) );
    unsigned int* jj = malloc( sizeof( unsigned int ) * ( nz ? nz : 1 ) );
    if ( ii == NULL || jj == NULL ) { free( ii ); free( jj ); free( dd ); fclose( f ); clear_matrix( A ); return -12; }
    for ( long k = 0; k < nz; k++ ) {
      if ( fgets( line, sizeof( line ), f ) == NULL ) { ret = -11; break; }
      if ( strlen( line ) > MM_LINE_MAX ) { ret = -13; break; }
      const char* p = line; long i, j; double v;
      if ( tok_uint( &p, &i ) || tok_uint( &p, &j ) || tok_real( &p, &v ) || !rest_blank( p ) ) { ret = -12; break; }
      if ( i < 1 || i > rows || j < 1 || j > cols ) { ret = -12; break; }     /* 1-based indices inside the declared size (the reference leaves this as a TODO) */
      ii[k] = ( unsigned int )i; jj[k] = ( unsigned int )j; dd[k] = v;
    }
    A->ii = ii; A->jj = jj;
  } else {
    A->format = DCOL;
    if ( sym != SM_UNSYMMETRIC ) ret = -5;
    for ( long k = 0; k < nz && ret == 0; k++ ) {
      if ( fgets( line, sizeof( line ), f ) == NULL ) { ret = -11; break; }
      const char* p = line; double v;
      if ( tok_real( &p, &v ) || !rest_blank( p ) ) { ret = -12; break; }
      dd[k] = v;
    }
    A->base = FIRST_INDEX_ZERO; A->location = MC_STORE_BOTH;
  }
  A->dd = dd;
  fclose( f );
  if ( A->sym == SM_UNSYMMETRIC ) A->location = MC_STORE_BOTH;
  if ( ret != 0 ) clear_matrix( A );
  else if ( nz == 0 ) { free( A->dd ); free( A->ii ); free( A->jj ); A->dd = NULL; A->ii = A->jj = NULL; }
  return ret;
}

enum { EXT_MM, EXT_HB, EXT_MAT };
/* texts match src/file.c:739-784 */
static int ext_of( const char* n, int* ext, int is_input ) {
  const size_t s = strnlen( n, 4096 );
  const char* io = is_input ? "input" : "output";
  const char* rw = is_input ? "reader" : "writer";
  if ( s > 3 && strcmp( n + s - 3, ".mm" ) == 0 ) { *ext = EXT_MM; return 0; }
  if ( s > 3 && strcmp( n + s - 3, ".hb" ) == 0 ) { fprintf( stderr, "%s error: Sorry Harwell-Boeing %s is broken\n", io, rw ); return 1; }
  if ( s > 3 && strcmp( n + s - 3, ".rb" ) == 0 ) { fprintf( stderr, "%s error: Sorry Rutherford-Boeing %s is broken\n", io, rw ); return 1; }
  if ( s > 4 && strcmp( n + s - 4, ".mat" ) == 0 ) { fprintf( stderr, "%s error: Matlab file %s was not enabled\n", io, rw ); return 1; }
  fprintf( stderr, "%s error: Unrecognized file extension\n", io );
  return 1;
}

int load_matrix( char* n, matrix_t* A ) {
  assert( A != NULL );
  if ( n == NULL ) { fprintf( stderr, "input error: No input specified (-i)\n" ); return 1; }
  clear_matrix( A );
  int ext, ret;
  if ( ( ret = ext_of( n, &ext, 1 ) ) != 0 ) return ret;
  ret = readmm( n, A );
  if ( ret != 0 ) { fprintf( stderr, "input error: Failed to load matrix\n" ); return ret; }
  if ( A->sym == SM_UNSYMMETRIC ) detect_matrix_symmetry( A );
  return 0;
}

int save_matrix( matrix_t* AA, char* n ) {
  if ( n == NULL ) { fprintf( stderr, "output error: No output specified (-o)\n" ); return 1; }
  int ext, ret;
  if ( ( ret = ext_of( n, &ext, 0 ) ) != 0 ) return ret;
  if ( ( ret = convert_matrix( AA, SM_COO, FIRST_INDEX_ZERO ) ) != 0 ) { fprintf( stderr, "output error: Failed to store matrix\n" ); return ret; }
  FILE* f = fopen( n, "w" );
  if ( f == NULL ) { fprintf( stderr, "output error: Failed to store matrix\n" ); return 1; }
  const double* d = AA->dd;
  fprintf( f, "%%%%MatrixMarket matrix coordinate real general\n" );
  fprintf( f, "%zu %zu %zu\n", AA->m, AA->n, AA->nz );
  for ( size_t p = 0; p < AA->nz; p++ ) fprintf( f, "%u %u %.13e\n", AA->ii[p] + 1, AA->jj[p] + 1, d[p] );
  fclose( f );
  return 0;
}
