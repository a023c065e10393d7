/* b200_plan.c -- how one factorization is spread over G GPUs (host C, pattern only).
 *
 * The reference scales across ranks only inside its third-party back ends (MUMPS:
 * src/solver_mumps.c:44-51 hands MPI_COMM_WORLD to dmumps_c; subtree-to-process mapping
 * and ScaLAPACK root fronts live in un-vendored code).  This file restates the textbook
 * scheme (Geist-Ng / Pothen-Sun proportional mapping; 1-D block-cyclic fronts) for the
 * b200 back end:
 *
 *   - every supernode gets a contiguous group of ranks [first, first+size) by proportional
 *     mapping of the subtree work: the root gets all ranks, children split their parent's
 *     group in proportion to their subtree flops, a group of one rank owns its whole subtree
 *     (no communication below that point);
 *   - a supernode whose group has more than one rank and whose front is large enough is
 *     DISTRIBUTED: its front columns (pivot columns and contribution-block columns alike,
 *     indexed 0..f-1) are dealt out in blocks of W columns, block b to rank
 *     first + b mod size, W = the outer panel width of its tree level;
 *   - the only data that ever crosses NVLink: factored panels of distributed fronts
 *     (owner -> its group), contribution-block columns at the boundary child -> distributed
 *     parent (owner of the child column -> owner of the destination column), and the finished
 *     factor panels at the end (each holder -> the ranks that lack them), so that every rank
 *     can run the triangular sweeps.
 *
 * Everything here is deterministic and depends on (symbolic, world) only, so every rank
 * derives the same plan without talking to the others.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "solver_b200.h"

/* Outer panel width (= K of the trailing update = column-block width of distributed fronts) from a level's largest
 * front.  On one GPU the factorization is bound by the DMMA update, which likes K = 512; with several GPUs the top
 * fronts are bound by the chain panel -> transfer -> update of the next panel's columns, whose in-panel and
 * next-panel parts grow with the SQUARE of the width; a cap can be set (b200_set_outer_width_cap, env B200_NBO_CAP),
 * but measured at 4 GPUs it does not pay (256: 273 ms, 512: 266 ms) -- per-block costs dominate the chain. */
static int g_width_cap = 0;
void b200_set_outer_width_cap( int cap ) { g_width_cap = cap; }
int b200_outer_width( int maxfront ) {
  int w = maxfront >= 6144 ? 512 : maxfront >= 1536 ? 256 : 128;
  if ( g_width_cap > 0 && w > g_width_cap ) w = g_width_cap;
  return w;
}

static int cmp_work_desc( const void* a, const void* b, void* ctx ) {
  const double* w = ( const double* )ctx;
  const int x = *( const int* )a, y = *( const int* )b;
  if ( w[x] > w[y] ) return -1;
  if ( w[x] < w[y] ) return 1;
  return x < y ? -1 : x > y;
}

/* split the group [first, first+P) among the children (sorted by work, heaviest first) */
static void map_children( b200_plan_t* pl, const int* ch, int nch, int first, int P ) {
  if ( nch == 0 ) return;
  double W = 0;
  for ( int i = 0; i < nch; i++ ) W += pl->work[ch[i]];
  double pos = 0.0;
  for ( int i = 0; i < nch; i++ ) {
    const int c = ch[i];
    if ( P == 1 || !( W > 0 ) ) { pl->grp_first[c] = first; pl->grp_size[c] = 1; continue; }
    const double share = P * pl->work[c] / W;
    int a, sz;
    if ( share >= 1.0 ) {
      a = ( int )floor( pos + 0.5 );
      int e = ( int )floor( pos + share + 0.5 );
      if ( a > P - 1 ) a = P - 1;
      if ( e > P ) e = P;
      if ( e <= a ) e = a + 1;
      sz = e - a;
    } else {
      a = ( int )floor( pos + 0.5 * share );
      if ( a > P - 1 ) a = P - 1;
      if ( a < 0 ) a = 0;
      sz = 1;
    }
    pl->grp_first[c] = first + a; pl->grp_size[c] = sz;
    pos += share;
  }
}

b200_plan_t* b200_plan_create( const b200_symbolic_t* S, int world, int min_dist_front ) {
  if ( !S || world < 1 ) return NULL;
  const int ns = S->ns;
  b200_plan_t* pl = calloc( 1, sizeof( *pl ) );
  pl->world = world; pl->ns = ns; pl->nlevels = S->maxdepth + 1;
  pl->grp_first = calloc( ns > 0 ? ns : 1, sizeof( int ) );
  pl->grp_size = calloc( ns > 0 ? ns : 1, sizeof( int ) );
  pl->dist = calloc( ns > 0 ? ns : 1, sizeof( int ) );
  pl->work = calloc( ns > 0 ? ns : 1, sizeof( double ) );
  pl->level_width = calloc( pl->nlevels, sizeof( int ) );
  /* outer panel width per level (from the level's largest front -- over ALL supernodes, so that every rank agrees) */
  {
    int* maxf = calloc( pl->nlevels, sizeof( int ) );
    for ( int s = 0; s < ns; s++ ) { const int f = ( int )( S->rowptr[s + 1] - S->rowptr[s] ); if ( f > maxf[S->sn_depth[s]] ) maxf[S->sn_depth[s]] = f; }
    for ( int d = 0; d < pl->nlevels; d++ ) pl->level_width[d] = b200_outer_width( maxf[d] );
    free( maxf );
  }
  /* subtree work: supernodes are numbered in postorder (children before parents) */
  for ( int s = 0; s < ns; s++ ) {
    const double k = S->sn_start[s + 1] - S->sn_start[s], f = ( double )( S->rowptr[s + 1] - S->rowptr[s] );
    pl->work[s] += k * f * f - k * k * f + k * k * k / 3.0;
    if ( S->sn_parent[s] >= 0 ) pl->work[S->sn_parent[s]] += pl->work[s];
  }
  /* children lists, heaviest first */
  int* cptr = calloc( ns + 2, sizeof( int ) );
  int* clist = calloc( ns > 0 ? ns : 1, sizeof( int ) );
  /* slot ns collects the roots */
  {
    int* cnt = calloc( ns + 1, sizeof( int ) );
    for ( int s = 0; s < ns; s++ ) cnt[S->sn_parent[s] >= 0 ? S->sn_parent[s] : ns]++;
    int acc = 0;
    for ( int s = 0; s <= ns; s++ ) { cptr[s] = acc; acc += cnt[s]; }
    cptr[ns + 1] = acc;
    memset( cnt, 0, sizeof( int ) * ( ns + 1 ) );
    for ( int s = 0; s < ns; s++ ) { const int p = S->sn_parent[s] >= 0 ? S->sn_parent[s] : ns; clist[cptr[p] + cnt[p]++] = s; }
    free( cnt );
  }
  for ( int s = 0; s <= ns; s++ ) qsort_r( clist + cptr[s], cptr[s + 1] - cptr[s], sizeof( int ), cmp_work_desc, pl->work );
  /* top-down: parents have larger indices than their children */
  map_children( pl, clist + cptr[ns], cptr[ns + 1] - cptr[ns], 0, world );
  for ( int s = ns - 1; s >= 0; s-- ) map_children( pl, clist + cptr[s], cptr[s + 1] - cptr[s], pl->grp_first[s], pl->grp_size[s] );
  for ( int s = 0; s < ns; s++ ) {
    const int f = ( int )( S->rowptr[s + 1] - S->rowptr[s] );
    pl->dist[s] = pl->grp_size[s] > 1 && f >= min_dist_front;
  }
  free( cptr ); free( clist );
  return pl;
}

void b200_plan_free( b200_plan_t* pl ) {
  if ( !pl ) return;
  free( pl->grp_first ); free( pl->grp_size ); free( pl->dist ); free( pl->work ); free( pl->level_width );
  free( pl );
}

/* rank that owns (computes) front-local column c of supernode s */
int b200_plan_col_owner( const b200_plan_t* pl, const b200_symbolic_t* S, int s, int c ) {
  if ( !pl->dist[s] ) return pl->grp_first[s];
  const int W = pl->level_width[S->sn_depth[s]];
  return pl->grp_first[s] + ( c / W ) % pl->grp_size[s];
}

/* does rank r hold / take part in supernode s ? */
int b200_plan_member( const b200_plan_t* pl, int s, int r ) {
  if ( !pl->dist[s] ) return r == pl->grp_first[s];
  return r >= pl->grp_first[s] && r < pl->grp_first[s] + pl->grp_size[s];
}

/* Contribution-block columns that must move before the extend-add of tree level `depth`:
 * column j of child c (front-local column k_c + j, computed by its column owner) is needed by the
 * owner of the parent's column relidx[j].  Consecutive columns with the same (src, dst) form one
 * transfer of whole stored columns (leading dimension u_c rounded up to even) inside the child's
 * contribution-block arena.  Returns the number of transfers; fills at most cap of them. */
int64_t b200_plan_cb_transfers( const b200_plan_t* pl, const b200_symbolic_t* S, int depth, b200_xfer_t* out, int64_t cap ) {
  int64_t nx = 0;
  for ( int c = 0; c < S->ns; c++ ) {
    const int p = S->sn_parent[c];
    if ( p < 0 || S->sn_depth[p] != depth ) continue;
    const int kc = S->sn_start[c + 1] - S->sn_start[c], fc = ( int )( S->rowptr[c + 1] - S->rowptr[c] ), uc = fc - kc;
    if ( uc == 0 ) continue;
    if ( !pl->dist[p] && !pl->dist[c] && pl->grp_first[p] == pl->grp_first[c] ) continue;   /* same single owner */
    const int64_t ldu = ( uc + 1 ) & ~1;
    const int* rel = S->relidx + S->rowptr[c] + kc;
    int j0 = 0;
    while ( j0 < uc ) {
      const int src = b200_plan_col_owner( pl, S, c, kc + j0 ), dst = b200_plan_col_owner( pl, S, p, rel[j0] );
      int j1 = j0 + 1;
      while ( j1 < uc && b200_plan_col_owner( pl, S, c, kc + j1 ) == src && b200_plan_col_owner( pl, S, p, rel[j1] ) == dst ) j1++;
      if ( src != dst ) {
        if ( nx < cap ) { b200_xfer_t x; x.src = src; x.dst = dst; x.sn = c; x.offset = S->cbptr[c] + ( int64_t )j0 * ldu; x.count = ( int64_t )( j1 - j0 ) * ldu; out[nx] = x; }
        nx++;
      }
      j0 = j1;
    }
  }
  return nx;
}
