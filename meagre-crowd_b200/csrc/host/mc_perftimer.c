/* mc_perftimer.c -- wall-clock tic list that defines "factor s" / "solve s".
 *
 * Behavioural re-statement of the reference's src/perftimer.c: a singly linked
 * list of gettimeofday() tics with a nesting depth (:84-124), rounds chained
 * through ->old by perftimer_restart (:519-527), chart output (:201-280) and the
 * CSV header/body used by `meagre-crowd -t` (:294-411).  A tic's duration runs
 * to the next tic of the same or shallower depth; per-tic times are averaged
 * over the rounds that contain that tic.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mc_solver.h"

perftimer_t* perftimer_malloc( void ) { return calloc( 1, sizeof( perftimer_t ) ); }

void perftimer_free( perftimer_t* h ) {
  while ( h != NULL ) {
    perftimer_tic_t* p = h->head;
    while ( p != NULL ) { perftimer_tic_t* nx = p->next; free( p->desc ); free( p ); p = nx; }
    perftimer_t* old = h->old;
    free( h );
    h = old;
  }
}

int perftimer_inc( perftimer_t* h, char const* const s, const size_t n ) {
  if ( h == NULL ) return -1;
  perftimer_tic_t* t = malloc( sizeof( *t ) );
  if ( t == NULL ) return -2;
  gettimeofday( &t->now, NULL );
  t->depth = h->current_depth;
  t->next = NULL;
  size_t l = s ? strnlen( s, n ) : 0;
  t->desc = l ? strndup( s, l ) : NULL;
  if ( h->tail == NULL ) h->head = h->tail = t; else { h->tail->next = t; h->tail = t; }
  return 0;
}

static double tic_diff( perftimer_tic_t const* a, perftimer_tic_t const* b ) {
  if ( a == NULL || b == NULL ) return 0.0;
  return ( double )( b->now.tv_sec - a->now.tv_sec ) + 1e-6 * ( double )( b->now.tv_usec - a->now.tv_usec );
}

unsigned int perftimer_rounds( perftimer_t const* h ) {
  unsigned int c = 0;
  for ( ; h != NULL; h = h->old ) if ( h->head != NULL && h->head != h->tail ) c++;
  return c;
}

double perftimer_wall( perftimer_t const* const h ) {
  if ( h == NULL ) return 0.0;
  perftimer_t const* first = h;
  while ( first->old != NULL ) first = first->old;
  return tic_diff( first->head, h->tail );
}

double perftimer_wall_av( perftimer_t const* h ) {
  if ( h == NULL ) return 0.0;
  double runs = perftimer_rounds( h ), sum = 0.0;
  if ( runs == 0 ) runs = 1;
  for ( ; h != NULL; h = h->old ) sum += tic_diff( h->head, h->tail );
  return sum / runs;
}

double perftimer_delta( perftimer_t const* const h ) {
  if ( h == NULL || h->head == NULL || h->head == h->tail ) return 0.0;
  perftimer_tic_t const* p = h->head;
  while ( p->next->next != NULL ) p = p->next;
  return tic_diff( p, h->tail );
}

void perftimer_restart( perftimer_t** ph ) {
  if ( ph == NULL ) return;
  perftimer_t* nh = perftimer_malloc();
  if ( nh == NULL ) { perftimer_free( *ph ); *ph = NULL; return; }
  nh->old = *ph;
  *ph = nh;
}

void perftimer_adjust_depth( perftimer_t* const h, int d ) {
  if ( h == NULL ) return;
  if ( d < 0 && ( unsigned int )( -d ) > h->current_depth ) h->current_depth = 0;
  else h->current_depth += d;
}

/* per-link accumulated seconds and round counts; returns number of links */
static unsigned int link_times( perftimer_t const* h, double** t_out, unsigned int** r_out ) {
  unsigned int mmax = 0;
  for ( perftimer_t const* hs = h; hs != NULL; hs = hs->old ) {
    unsigned int c = 0;
    for ( perftimer_tic_t const* p = hs->head; p != NULL; p = p->next ) c++;
    if ( c > mmax ) mmax = c;
  }
  double* t = calloc( mmax + 1, sizeof( double ) );
  unsigned int* r = calloc( mmax + 1, sizeof( unsigned int ) );
  for ( perftimer_t const* hs = h; hs != NULL; hs = hs->old ) {
    perftimer_tic_t const* p = hs->head;
    for ( unsigned int i = 0; i < mmax && p != NULL && p->next != NULL; i++ ) {
      perftimer_tic_t const* stop = p->next;
      while ( stop->next != NULL && stop->depth > p->depth ) stop = stop->next;
      t[i] += tic_diff( p, stop ); r[i]++;
      p = p->next;
    }
  }
  *t_out = t; *r_out = r;
  return mmax;
}

void perftimer_printf( perftimer_t const* const h, const unsigned int d ) {
  if ( h == NULL || h->head == NULL ) return;
  double* t; unsigned int* r;
  link_times( h, &t, &r );
  int i = 0;
  for ( perftimer_tic_t const* p = h->head; p->next != NULL; p = p->next, i++ ) {
    if ( p->desc == NULL || p->depth > d ) continue;
    for ( unsigned int j = 0; j < p->depth; j++ ) printf( "  " );
    printf( "%s", p->desc );
    if ( t[i] > 0.0 ) {
      if ( r[i] > 1 ) printf( ":\t%0.3fms (av. %0.3fms in %d rounds)\n", t[i] * 1e3, t[i] / ( double )r[i] * 1e3, r[i] );
      else printf( ":\t%0.3fms\n", t[i] * 1e3 );
    } else printf( "\n" );
  }
  printf( "total:\t%0.3fms", perftimer_wall( h ) * 1e3 );
  unsigned int rn = perftimer_rounds( h );
  if ( rn > 1 ) printf( " (av. %0.3fms in %d rounds)", perftimer_wall_av( h ) * 1e3, rn );
  printf( "\n" );
  free( t ); free( r );
}

void perftimer_printf_csv_header( perftimer_t const* const h, const unsigned int d ) {
  if ( h == NULL || h->head == NULL ) return;
  for ( perftimer_tic_t const* p = h->head; p->next != NULL; p = p->next )
    if ( p->desc != NULL && p->depth <= d ) printf( "%s, ", p->desc );
  printf( "total (ms)" );
  if ( perftimer_rounds( h ) > 1 ) printf( ", rounds" );
  printf( "\n" );
}

void perftimer_printf_csv_body( perftimer_t const* const h, const unsigned int d ) {
  if ( h == NULL || h->head == NULL ) return;
  double* t; unsigned int* r;
  link_times( h, &t, &r );
  int i = 0;
  for ( perftimer_tic_t const* p = h->head; p->next != NULL; p = p->next, i++ )
    if ( p->desc != NULL && p->depth <= d ) printf( "%0.3f, ", t[i] / ( double )( r[i] ? r[i] : 1 ) * 1e3 );
  printf( "%0.3f", perftimer_wall_av( h ) * 1e3 );
  unsigned int rn = perftimer_rounds( h );
  if ( rn > 1 ) printf( ", %d", rn );
  printf( "\n" );
  free( t ); free( r );
}

double perftimer_tic_ms( perftimer_t const* h, char const* name ) {
  if ( h == NULL || h->head == NULL || name == NULL ) return -1.0;
  double* t; unsigned int* r;
  link_times( h, &t, &r );
  double out = -1.0; int i = 0;
  for ( perftimer_tic_t const* p = h->head; p->next != NULL; p = p->next, i++ )
    if ( p->desc != NULL && strcmp( p->desc, name ) == 0 ) { out = t[i] / ( double )( r[i] ? r[i] : 1 ) * 1e3; break; }
  free( t ); free( r );
  return out;
}
