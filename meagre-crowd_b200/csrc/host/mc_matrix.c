/* mc_matrix.c -- matrix_t lifecycle and conversions at the plugin boundary.
 *
 * Re-statement (not a copy) of the behaviour of the reference's src/matrix.c
 * that the dispatcher and driver rely on:
 *   malloc/free/clear/copy            src/matrix.c:117-175
 *   convert_matrix (format x base)    src/matrix.c:898-1076
 *   dense<->COO, dropping |v|<=1e-15  src/matrix.c:743-815
 *   convert_matrix_symmetry           src/matrix.c:1247-1306
 *   detect_matrix_symmetry            src/matrix.c:1309-1369   (tol 5e-15)
 *   validate_matrix                   src/matrix.c:1378-1428
 *   printf_matrix                     src/matrix.c:1430-1456
 * The reference delegates COO<->CSC/CSR to BeBOP (un-vendored) and detects
 * symmetry in O(nz^2); here every conversion is a counting sort, O(nz + n),
 * and symmetry detection sorts the entries, O(nz log nz), with identical results.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <assert.h>
#include <math.h>
#include "mc_matrix.h"

matrix_t* malloc_matrix( void ) { return calloc( 1, sizeof( matrix_t ) ); }

void clear_matrix( matrix_t* m ) {
  if ( m == NULL ) return;
  free( m->dd ); free( m->ii ); free( m->jj );
  memset( m, 0, sizeof( *m ) );
}

void free_matrix( matrix_t* m ) {
  if ( m == NULL ) return;
  clear_matrix( m );
  free( m );
}

static size_t ii_len( const matrix_t* m ) {
  switch ( m->format ) {
    case SM_COO: case SM_CSC: return m->nz;
    case SM_CSR: return m->m + 1;
    default: return 0;
  }
}
static size_t jj_len( const matrix_t* m ) {
  switch ( m->format ) {
    case SM_COO: case SM_CSR: return m->nz;
    case SM_CSC: return m->n + 1;
    default: return 0;
  }
}

matrix_t* copy_matrix( matrix_t* m ) {
  if ( m == NULL ) return NULL;
  matrix_t* c = malloc_matrix();
  if ( c == NULL ) return NULL;
  *c = *m; c->dd = NULL; c->ii = NULL; c->jj = NULL;
  assert( m->data_type == REAL_DOUBLE || m->dd == NULL );
  if ( m->dd ) { c->dd = malloc( sizeof( double ) * m->nz ); memcpy( c->dd, m->dd, sizeof( double ) * m->nz ); }
  if ( m->ii ) { size_t l = ii_len( m ); c->ii = malloc( sizeof( unsigned int ) * l ); memcpy( c->ii, m->ii, sizeof( unsigned int ) * l ); }
  if ( m->jj ) { size_t l = jj_len( m ); c->jj = malloc( sizeof( unsigned int ) * l ); memcpy( c->jj, m->jj, sizeof( unsigned int ) * l ); }
  return c;
}

/* ---- internal: everything funnels through COO, base 0 */
static int to_coo0( matrix_t* m ) {
  const size_t nz = m->nz;
  switch ( m->format ) {
    case SM_COO:
      if ( m->base == FIRST_INDEX_ONE ) for ( size_t p = 0; p < nz; p++ ) { m->ii[p]--; m->jj[p]--; }
      break;
    case SM_CSC: {
      unsigned int* jj = malloc( sizeof( unsigned int ) * ( nz ? nz : 1 ) );
      const unsigned int b = m->base;
      for ( size_t j = 0; j < m->n; j++ ) for ( unsigned int p = m->jj[j] - b; p < m->jj[j + 1] - b; p++ ) jj[p] = ( unsigned int )j;
      if ( b ) for ( size_t p = 0; p < nz; p++ ) m->ii[p]--;
      free( m->jj ); m->jj = jj;
      break;
    }
    case SM_CSR: {
      unsigned int* ii = malloc( sizeof( unsigned int ) * ( nz ? nz : 1 ) );
      const unsigned int b = m->base;
      for ( size_t i = 0; i < m->m; i++ ) for ( unsigned int p = m->ii[i] - b; p < m->ii[i + 1] - b; p++ ) ii[p] = ( unsigned int )i;
      if ( b ) for ( size_t p = 0; p < nz; p++ ) m->jj[p]--;
      free( m->ii ); m->ii = ii;
      break;
    }
    case DROW: case DCOL: {
      /* keep entries with |v| > 1e-15 (src/matrix.c:745,775), scanning row by row */
      const double* d = m->dd;
      size_t cnt = 0;
      for ( size_t p = 0; p < m->m * m->n; p++ ) if ( fabs( d[p] ) > 1e-15 ) cnt++;
      unsigned int* ii = malloc( sizeof( unsigned int ) * ( cnt ? cnt : 1 ) );
      unsigned int* jj = malloc( sizeof( unsigned int ) * ( cnt ? cnt : 1 ) );
      double* dd = malloc( sizeof( double ) * ( cnt ? cnt : 1 ) );
      size_t q = 0;
      for ( size_t i = 0; i < m->m; i++ ) for ( size_t j = 0; j < m->n; j++ ) {
          double v = ( m->format == DCOL ) ? d[j * m->m + i] : d[i * m->n + j];
          if ( fabs( v ) > 1e-15 ) { ii[q] = ( unsigned int )i; jj[q] = ( unsigned int )j; dd[q] = v; q++; }
        }
      free( m->dd );
      if ( cnt == 0 ) { free( ii ); free( jj ); free( dd ); ii = jj = NULL; dd = NULL; }
      m->dd = dd; m->ii = ii; m->jj = jj; m->nz = cnt;
      break;
    }
    default: return -1;
  }
  m->format = SM_COO; m->base = FIRST_INDEX_ZERO;
  return 0;
}

static int from_coo0( matrix_t* m, enum matrix_format_t f, enum matrix_base_t b ) {
  const size_t nz = m->nz;
  switch ( f ) {
    case SM_COO:
      if ( b == FIRST_INDEX_ONE ) for ( size_t p = 0; p < nz; p++ ) { m->ii[p]++; m->jj[p]++; }
      break;
    case SM_CSC: case SM_CSR: {
      /* stable counting sort on the compressed index */
      const int bycol = ( f == SM_CSC );
      const size_t dim = bycol ? m->n : m->m;
      unsigned int* key = bycol ? m->jj : m->ii;
      unsigned int* oth = bycol ? m->ii : m->jj;
      unsigned int* ptr = calloc( dim + 1, sizeof( unsigned int ) );
      for ( size_t p = 0; p < nz; p++ ) ptr[key[p] + 1]++;
      for ( size_t j = 0; j < dim; j++ ) ptr[j + 1] += ptr[j];
      unsigned int* w = malloc( sizeof( unsigned int ) * ( dim + 1 ) );
      memcpy( w, ptr, sizeof( unsigned int ) * ( dim + 1 ) );
      unsigned int* idx = malloc( sizeof( unsigned int ) * ( nz ? nz : 1 ) );
      double* dd = malloc( sizeof( double ) * ( nz ? nz : 1 ) );
      const double* d = m->dd;
      for ( size_t p = 0; p < nz; p++ ) { unsigned int q = w[key[p]]++; idx[q] = oth[p] + b; dd[q] = d[p]; }
      for ( size_t j = 0; j <= dim; j++ ) ptr[j] += b;
      free( w ); free( m->ii ); free( m->jj ); free( m->dd );
      if ( nz == 0 ) { free( idx ); free( dd ); idx = NULL; dd = NULL; free( ptr ); ptr = NULL; }
      m->dd = dd;
      if ( bycol ) { m->ii = idx; m->jj = ptr; } else { m->jj = idx; m->ii = ptr; }
      break;
    }
    case DROW: case DCOL: {
      double* dd = calloc( m->m * m->n ? m->m * m->n : 1, sizeof( double ) );
      const double* d = m->dd;
      for ( size_t p = 0; p < nz; p++ ) {
        size_t i = m->ii[p], j = m->jj[p];
        size_t q = ( f == DCOL ) ? j * m->m + i : i * m->n + j;
        dd[q] += d[p];
        if ( m->sym == SM_SYMMETRIC && m->location != MC_STORE_BOTH && i != j ) {
          size_t q2 = ( f == DCOL ) ? i * m->m + j : j * m->n + i;
          dd[q2] += d[p];
        }
      }
      if ( m->sym == SM_SYMMETRIC ) m->location = MC_STORE_BOTH;
      free( m->dd ); free( m->ii ); free( m->jj );
      m->dd = dd; m->ii = NULL; m->jj = NULL; m->nz = m->m * m->n;
      break;
    }
    default: return -1;
  }
  m->format = f; m->base = b;
  return 0;
}

int convert_matrix( matrix_t* m, enum matrix_format_t f, enum matrix_base_t b ) {
  if ( m == NULL || f == INVALID ) return -1;
  if ( m->format == INVALID ) return -1;
  if ( m->format == f && ( m->base == b || f == DROW || f == DCOL ) ) { m->base = b; return 0; }
  assert( m->data_type == REAL_DOUBLE );
  /* dense <-> dense: in-place transposition of the storage order */
  if ( ( m->format == DROW || m->format == DCOL ) && ( f == DROW || f == DCOL ) ) {
    if ( m->m != 1 && m->n != 1 ) {
      double* dd = malloc( sizeof( double ) * m->m * m->n );
      const double* d = m->dd;
      for ( size_t i = 0; i < m->m; i++ ) for ( size_t j = 0; j < m->n; j++ ) {
          if ( f == DCOL ) dd[j * m->m + i] = d[i * m->n + j]; else dd[i * m->n + j] = d[j * m->m + i];
        }
      free( m->dd ); m->dd = dd;
    } /* vectors: same memory either way (the dispatcher relies on this, src/solvers.c:572-576) */
    m->format = f; m->base = b;
    return 0;
  }
  int r = to_coo0( m );
  if ( r ) return r;
  return from_coo0( m, f, b );
}

int convert_matrix_symmetry( matrix_t* m, enum matrix_symmetric_storage_t loc ) {
  assert( m->sym == SM_SYMMETRIC );
  if ( m->location == loc ) return 0;
  const enum matrix_format_t old_format = m->format;
  const enum matrix_base_t old_base = m->base;
  if ( old_format == DROW || old_format == DCOL ) { m->location = loc; return 0; } /* dense holds everything */
  int r = to_coo0( m );
  if ( r ) return r;
  double* d = m->dd;
  if ( m->location == MC_STORE_BOTH ) {
    /* keep the lower triangle (row >= col), then mirror if the upper one is wanted */
    size_t q = 0;
    for ( size_t p = 0; p < m->nz; p++ ) if ( m->ii[p] >= m->jj[p] ) { m->ii[q] = m->ii[p]; m->jj[q] = m->jj[p]; d[q] = d[p]; q++; }
    m->nz = q;
    m->location = LOWER_TRIANGULAR;
  }
  if ( loc == MC_STORE_BOTH ) {
    size_t off = 0;
    for ( size_t p = 0; p < m->nz; p++ ) if ( m->ii[p] != m->jj[p] ) off++;
    const size_t nz2 = m->nz + off;
    m->ii = realloc( m->ii, sizeof( unsigned int ) * ( nz2 ? nz2 : 1 ) );
    m->jj = realloc( m->jj, sizeof( unsigned int ) * ( nz2 ? nz2 : 1 ) );
    m->dd = realloc( m->dd, sizeof( double ) * ( nz2 ? nz2 : 1 ) );
    d = m->dd;
    size_t q = m->nz;
    for ( size_t p = 0; p < m->nz; p++ ) if ( m->ii[p] != m->jj[p] ) { m->ii[q] = m->jj[p]; m->jj[q] = m->ii[p]; d[q] = d[p]; q++; }
    m->nz = nz2;
    m->location = MC_STORE_BOTH;
  } else if ( m->location != loc ) {
    unsigned int* t = m->ii; m->ii = m->jj; m->jj = t;
    m->location = loc;
  }
  return from_coo0( m, old_format, old_base );
}

typedef struct { unsigned int r, c; double v; } ent_t;
static int cmp_ent( const void* a, const void* b ) {
  const ent_t* x = a; const ent_t* y = b;
  if ( x->r != y->r ) return x->r < y->r ? -1 : 1;
  if ( x->c != y->c ) return x->c < y->c ? -1 : 1;
  return 0;
}

/* A general matrix is symmetric when every off-diagonal entry has a mirror whose
 * value agrees within 5e-15 (src/matrix.c:1321,1336-1359). */
int detect_matrix_symmetry( matrix_t* m ) {
  if ( m->sym != SM_UNSYMMETRIC ) return 0;
  if ( m->m != m->n ) return 0;
  if ( m->format == DROW || m->format == DCOL ) {
    const double* d = m->dd; const size_t n = m->n; int sym = 1;
    for ( size_t i = 0; i < n && sym; i++ ) for ( size_t j = i + 1; j < n; j++ ) if ( fabs( d[i * n + j] - d[j * n + i] ) >= 5e-15 ) { sym = 0; break; }
    if ( sym ) { m->sym = SM_SYMMETRIC; m->location = MC_STORE_BOTH; }
    return 0;
  }
  const enum matrix_format_t old_format = m->format;
  const enum matrix_base_t old_base = m->base;
  int r = to_coo0( m );
  if ( r ) return r;
  const size_t nz = m->nz;
  ent_t* E = malloc( sizeof( ent_t ) * ( nz ? nz : 1 ) );
  const double* d = m->dd;
  for ( size_t p = 0; p < nz; p++ ) { E[p].r = m->ii[p]; E[p].c = m->jj[p]; E[p].v = d[p]; }
  qsort( E, nz, sizeof( ent_t ), cmp_ent );
  int sym = 1;
  for ( size_t p = 0; p < nz && sym; p++ ) {
    if ( E[p].r == E[p].c ) continue;
    ent_t key = { E[p].c, E[p].r, 0 };
    ent_t* hit = bsearch( &key, E, nz, sizeof( ent_t ), cmp_ent );
    if ( hit == NULL || !( E[p].v < hit->v + 5e-15 && E[p].v > hit->v - 5e-15 ) ) sym = 0;
  }
  free( E );
  if ( sym ) { m->sym = SM_SYMMETRIC; m->location = MC_STORE_BOTH; }
  return from_coo0( m, old_format, old_base );
}

int validate_matrix( matrix_t* m ) {
  assert( m != NULL );
  if ( ( m->m == 0 || m->n == 0 ) && m->format != INVALID ) return -1;
  if ( ( m->format == DROW || m->format == DCOL ) && m->nz != m->m * m->n ) return -1;
  if ( m->format == INVALID ) {
    if ( m->m != 0 || m->n != 0 || m->nz != 0 ) return -1;
    if ( m->ii != NULL || m->jj != NULL || m->dd != NULL ) return -2;
    return 0;
  }
  if ( m->nz == 0 && ( m->ii != NULL || m->jj != NULL || m->dd != NULL ) ) return -2;
  if ( m->data_type == SM_PATTERN && m->dd != NULL ) return -2;
  if ( ( m->format == DROW || m->format == DCOL ) && ( m->ii != NULL || m->jj != NULL ) ) return -2;
  if ( m->nz > 0 ) {
    if ( m->format == SM_CSR && m->ii[m->m] != m->nz + m->base ) return -3;
    if ( m->format == SM_CSC && m->jj[m->n] != m->nz + m->base ) return -3;
    if ( m->format == SM_CSR && m->ii[0] != ( unsigned int )m->base ) return -4;
    if ( m->format == SM_CSC && m->jj[0] != ( unsigned int )m->base ) return -4;
  }
  return 0;
}

/* 0 when the two matrices hold the same entries (after normalising to sorted COO base 0) */
int cmp_matrix( matrix_t* a, matrix_t* b ) {
  if ( a == NULL || b == NULL ) return a == b ? 0 : -1;
  if ( a->m != b->m || a->n != b->n ) return -1;
  matrix_t* x = copy_matrix( a ); matrix_t* y = copy_matrix( b );
  int ret = 0;
  if ( x->sym == SM_SYMMETRIC && x->format != INVALID ) convert_matrix_symmetry( x, MC_STORE_BOTH );
  if ( y->sym == SM_SYMMETRIC && y->format != INVALID ) convert_matrix_symmetry( y, MC_STORE_BOTH );
  if ( x->format != INVALID ) convert_matrix( x, SM_COO, FIRST_INDEX_ZERO );
  if ( y->format != INVALID ) convert_matrix( y, SM_COO, FIRST_INDEX_ZERO );
  if ( x->nz != y->nz ) ret = -2;
  else {
    ent_t* E = malloc( sizeof( ent_t ) * ( x->nz ? x->nz : 1 ) ); ent_t* F = malloc( sizeof( ent_t ) * ( x->nz ? x->nz : 1 ) );
    for ( size_t p = 0; p < x->nz; p++ ) { E[p] = ( ent_t ) { x->ii[p], x->jj[p], ( ( double* )x->dd )[p] }; F[p] = ( ent_t ) { y->ii[p], y->jj[p], ( ( double* )y->dd )[p] }; }
    qsort( E, x->nz, sizeof( ent_t ), cmp_ent ); qsort( F, x->nz, sizeof( ent_t ), cmp_ent );
    for ( size_t p = 0; p < x->nz; p++ ) if ( E[p].r != F[p].r || E[p].c != F[p].c || E[p].v != F[p].v ) { ret = -3; break; }
    free( E ); free( F );
  }
  free_matrix( x ); free_matrix( y );
  return ret;
}

void printf_matrix( char const* const pre, matrix_t* m ) {
  assert( m != NULL );
  matrix_t* const c = copy_matrix( m );
  assert( c != NULL );
  if ( c->format != INVALID ) convert_matrix( c, SM_COO, FIRST_INDEX_ZERO );
  const double* v = c->dd;
  if ( c->m == 0 || c->n == 0 ) printf( "%s is empty\n", pre );
  else if ( c->nz == 0 ) printf( "%s is all-zero\n", pre );
  else if ( c->n == 1 ) for ( size_t i = 0; i < c->nz; i++ ) printf( "%s(%i)=%.2f\n", pre, c->ii[i], v[i] );
  else for ( size_t i = 0; i < c->nz; i++ ) printf( "%s(%i,%i)=%.2f\n", pre, c->ii[i], c->jj[i], v[i] );
  free_matrix( c );
}

unsigned int matrix_rows( const matrix_t* const A ) { return A == NULL ? 0 : ( unsigned int )A->m; }
