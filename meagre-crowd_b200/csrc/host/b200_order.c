/* b200_order.c -- fill-reducing nested-dissection ordering (host C, from scratch).
 *
 * Replaces the ordering step that the reference delegates to un-vendored
 * libraries inside cholmod_analyze (src/solver_cholmod.c:257), umfpack_di_symbolic
 * (src/solver_umfpack.c:73) and MUMPS JOB=1 (src/solver_mumps.c:176-179).
 *
 * Method: recursive multilevel vertex-separator bisection on the graph of A+A':
 *   coarsen by heavy-edge matching -> grow initial bisections on the coarsest
 *   graph -> turn the edge cut into a vertex separator -> refine the separator
 *   with a two-sided node Fiduccia-Mattheyses pass at every uncoarsening level.
 * Leaves (<= leaf vertices) are ordered by exact minimum degree on the leaf's
 * own elimination graph. The two halves are independent OpenMP tasks.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "solver_b200.h"

typedef struct {
  int nv, ne;
  int* xadj;
  int* adj;
  int* vwgt;
  int* ewgt;
} graph_t;

/* Tuning knobs from the environment.  Read ONCE in b200_order_nd, before the OpenMP region starts, so that no task
 * ever sees a half-initialised value: the ordering must not depend on thread timing (ranks of a multi-GPU run derive
 * their plans from it independently). */
typedef struct { int fm_cap, tries, fm_big, fm_small, starts, verbose; double ub, pen; } nd_knobs_t;
static nd_knobs_t g_knobs = { 400, 12, 3, 5, 4, 0, 1.40, 2.0 };
static void read_knobs( void ) {
  nd_knobs_t k = { 400, 12, 3, 5, 4, 0, 1.40, 2.0 };
  const char* e;
  if ( ( e = getenv( "B200_FM_CAP" ) ) ) k.fm_cap = atoi( e );
  if ( ( e = getenv( "B200_ND_TRIES" ) ) ) k.tries = atoi( e );
  if ( ( e = getenv( "B200_FM_PASSES_BIG" ) ) ) k.fm_big = atoi( e );
  if ( ( e = getenv( "B200_FM_PASSES" ) ) ) k.fm_small = atoi( e );
  if ( ( e = getenv( "B200_ND_STARTS" ) ) ) { k.starts = atoi( e ); if ( k.starts < 1 ) k.starts = 1; if ( k.starts > 16 ) k.starts = 16; }
  if ( ( e = getenv( "B200_ND_UB" ) ) ) k.ub = atof( e );
  if ( ( e = getenv( "B200_ND_IMB_PENALTY" ) ) ) k.pen = atof( e );
  k.verbose = getenv( "B200_ND_VERBOSE" ) != NULL;
  g_knobs = k;
}

static void graph_free( graph_t* g ) {
  if ( !g ) return;
  free( g->xadj ); free( g->adj ); free( g->vwgt ); free( g->ewgt ); free( g );
}

static inline uint32_t rng_next( uint32_t* s ) {
  uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; *s = x ? x : 0x9e3779b9u; return *s;
}

/* ------------------------------------------------------------------ heap
 * indexed max-heap on integer keys (gain), one per destination side */
typedef struct { int n; int* node; int* key; int* pos; } heap_t;
static void heap_init( heap_t* h, int cap ) {
  h->n = 0; h->node = malloc( sizeof( int ) * ( cap + 1 ) ); h->key = malloc( sizeof( int ) * ( cap + 1 ) );
  h->pos = malloc( sizeof( int ) * ( cap + 1 ) );
  for ( int i = 0; i < cap; i++ ) h->pos[i] = -1;
}
static void heap_free( heap_t* h ) { free( h->node ); free( h->key ); free( h->pos ); }
static void heap_reset( heap_t* h ) { for ( int i = 0; i < h->n; i++ ) h->pos[h->node[i]] = -1; h->n = 0; }
static inline void heap_swap( heap_t* h, int a, int b ) {
  int t = h->node[a]; h->node[a] = h->node[b]; h->node[b] = t;
  h->pos[h->node[a]] = a; h->pos[h->node[b]] = b;
}
static inline int heap_less( heap_t* h, int a, int b ) { /* a has lower priority than b */
  int ka = h->key[h->node[a]], kb = h->key[h->node[b]];
  return ka < kb || ( ka == kb && h->node[a] > h->node[b] );
}
static void heap_up( heap_t* h, int i ) {
  while ( i > 0 ) { int p = ( i - 1 ) >> 1; if ( heap_less( h, p, i ) ) { heap_swap( h, p, i ); i = p; } else break; }
}
static void heap_down( heap_t* h, int i ) {
  for ( ;; ) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if ( l < h->n && heap_less( h, m, l ) ) m = l;
    if ( r < h->n && heap_less( h, m, r ) ) m = r;
    if ( m == i ) break;
    heap_swap( h, m, i ); i = m;
  }
}
static void heap_upsert( heap_t* h, int v, int key ) {
  if ( h->pos[v] < 0 ) { h->key[v] = key; h->node[h->n] = v; h->pos[v] = h->n; h->n++; heap_up( h, h->n - 1 ); }
  else { int old = h->key[v]; h->key[v] = key; if ( key > old ) heap_up( h, h->pos[v] ); else heap_down( h, h->pos[v] ); }
}
static void heap_remove( heap_t* h, int v ) {
  int i = h->pos[v]; if ( i < 0 ) return;
  h->n--;
  if ( i != h->n ) { heap_swap( h, i, h->n ); h->pos[v] = -1; heap_up( h, i ); heap_down( h, i ); }
  else h->pos[v] = -1;
}
static inline int heap_top( heap_t* h ) { return h->n ? h->node[0] : -1; }

/* ------------------------------------------------------------ coarsening */
static graph_t* coarsen( const graph_t* g, int* cmap, uint32_t* rng, int maxvwgt ) {
  const int nv = g->nv;
  int* match = malloc( sizeof( int ) * nv );
  int* order = malloc( sizeof( int ) * nv );
  int* cfirst = malloc( sizeof( int ) * nv );
  for ( int i = 0; i < nv; i++ ) { match[i] = -1; order[i] = i; }
  for ( int i = nv - 1; i > 0; i-- ) { int j = rng_next( rng ) % ( uint32_t )( i + 1 ); int t = order[i]; order[i] = order[j]; order[j] = t; }
  int cnv = 0;
  for ( int ii = 0; ii < nv; ii++ ) {
    int v = order[ii];
    if ( match[v] >= 0 ) continue;
    int best = -1, bestw = -1;
    if ( g->vwgt[v] < maxvwgt ) {
      for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) {
        int u = g->adj[e];
        if ( match[u] < 0 && g->ewgt[e] > bestw && g->vwgt[v] + g->vwgt[u] <= maxvwgt ) { best = u; bestw = g->ewgt[e]; }
      }
    }
    if ( best >= 0 ) { match[v] = best; match[best] = v; cmap[v] = cmap[best] = cnv; }
    else { match[v] = v; cmap[v] = cnv; }
    cfirst[cnv++] = v;
  }
  free( order );
  graph_t* c = calloc( 1, sizeof( graph_t ) );
  c->nv = cnv;
  c->xadj = malloc( sizeof( int ) * ( cnv + 1 ) );
  c->vwgt = malloc( sizeof( int ) * cnv );
  c->adj = malloc( sizeof( int ) * ( g->ne > 0 ? g->ne : 1 ) );
  c->ewgt = malloc( sizeof( int ) * ( g->ne > 0 ? g->ne : 1 ) );
  int* slot = malloc( sizeof( int ) * cnv );
  for ( int i = 0; i < cnv; i++ ) slot[i] = -1;
  int ne = 0;
  for ( int cv = 0; cv < cnv; cv++ ) {
    c->xadj[cv] = ne;
    const int start = ne;
    int mem[2]; int nm = 1; mem[0] = cfirst[cv];
    if ( match[mem[0]] != mem[0] ) { mem[1] = match[mem[0]]; nm = 2; }
    int w = 0;
    for ( int k = 0; k < nm; k++ ) {
      int v = mem[k]; w += g->vwgt[v];
      for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) {
        int cu = cmap[g->adj[e]];
        if ( cu == cv ) continue;
        int s = slot[cu];
        if ( s >= start ) c->ewgt[s] += g->ewgt[e];
        else { slot[cu] = ne; c->adj[ne] = cu; c->ewgt[ne] = g->ewgt[e]; ne++; }
      }
    }
    c->vwgt[cv] = w;
  }
  c->xadj[cnv] = ne; c->ne = ne;
  free( slot ); free( match ); free( cfirst );
  return c;
}

/* --------------------------------------------- node separator refinement
 * where[v] in {0,1,2}; 2 = separator. Moving separator vertex v to side s
 * pulls its neighbours on side 1-s into the separator. */
typedef struct { int v; int npull; } move_t;

static inline int sep_gain( const graph_t* g, const int* where, int v, int to ) {
  int other = 1 - to, ext = 0;
  for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) if ( where[g->adj[e]] == other ) ext += g->vwgt[g->adj[e]];
  return g->vwgt[v] - ext;
}

static void node_fm( const graph_t* g, int* where, int* pw, int maxpw, int npasses, heap_t* H, int* locked, int* pulled_buf, move_t* moves ) {
  const int nv = g->nv;
  for ( int pass = 0; pass < npasses; pass++ ) {
    heap_reset( &H[0] ); heap_reset( &H[1] );
    int nsepv = 0;
    for ( int v = 0; v < nv; v++ ) {
      locked[v] = 0;
      if ( where[v] == 2 ) { heap_upsert( &H[0], v, sep_gain( g, where, v, 0 ) ); heap_upsert( &H[1], v, sep_gain( g, where, v, 1 ) ); nsepv++; }
    }
    if ( nsepv == 0 ) return;
    const int fm_cap = g_knobs.fm_cap;
    const int limit = nsepv < 40 ? 40 + nsepv : ( nsepv / 4 + 60 > fm_cap ? fm_cap : nsepv / 4 + 60 );
    int nmoves = 0, npulled = 0, bestmoves = 0;
    int bestsep = pw[2], bestdiff = abs( pw[0] - pw[1] );
    int cursep = pw[2];
    int since = 0;
    const int initsep = pw[2];
    ( void )initsep;
    while ( since < limit ) {
      int t0 = heap_top( &H[0] ), t1 = heap_top( &H[1] );
      if ( t0 < 0 && t1 < 0 ) break;
      int to;
      if ( t0 < 0 ) to = 1; else if ( t1 < 0 ) to = 0;
      else {
        int g0 = H[0].key[t0], g1 = H[1].key[t1];
        /* prefer the move that keeps balance; otherwise the higher gain */
        int ok0 = pw[0] + g->vwgt[t0] <= maxpw, ok1 = pw[1] + g->vwgt[t1] <= maxpw;
        if ( ok0 && !ok1 ) to = 0; else if ( ok1 && !ok0 ) to = 1;
        else if ( g0 != g1 ) to = g0 > g1 ? 0 : 1;
        else to = pw[0] <= pw[1] ? 0 : 1;
      }
      int v = heap_top( &H[to] );
      if ( pw[to] + g->vwgt[v] > maxpw ) { /* cannot move there; drop from this heap */
        heap_remove( &H[to], v );
        continue;
      }
      const int other = 1 - to;
      heap_remove( &H[0], v ); heap_remove( &H[1], v );
      locked[v] = 1;
      where[v] = to; pw[2] -= g->vwgt[v]; pw[to] += g->vwgt[v];
      int np = 0;
      for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) {
        int u = g->adj[e];
        if ( where[u] == other ) {
          where[u] = 2; pw[other] -= g->vwgt[u]; pw[2] += g->vwgt[u];
          pulled_buf[npulled + np] = u; np++;
        }
      }
      /* refresh gains of affected separator vertices */
      for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) {
        int u = g->adj[e];
        if ( where[u] == 2 && !locked[u] ) {
          heap_upsert( &H[0], u, sep_gain( g, where, u, 0 ) ); heap_upsert( &H[1], u, sep_gain( g, where, u, 1 ) );
        }
      }
      for ( int k = 0; k < np; k++ ) {
        int u = pulled_buf[npulled + k];
        for ( int e = g->xadj[u]; e < g->xadj[u + 1]; e++ ) {
          int w = g->adj[e];
          if ( where[w] == 2 && !locked[w] && w != u ) {
            heap_upsert( &H[0], w, sep_gain( g, where, w, 0 ) ); heap_upsert( &H[1], w, sep_gain( g, where, w, 1 ) );
          }
        }
      }
      moves[nmoves].v = v; moves[nmoves].npull = np; nmoves++; npulled += np;
      cursep = pw[2];
      int diff = abs( pw[0] - pw[1] );
      if ( cursep < bestsep || ( cursep == bestsep && diff < bestdiff ) ) { bestsep = cursep; bestdiff = diff; bestmoves = nmoves; since = 0; }
      else since++;
      if ( nmoves >= nv ) break;
    }
    /* roll back to the best prefix */
    while ( nmoves > bestmoves ) {
      nmoves--;
      int v = moves[nmoves].v, np = moves[nmoves].npull;
      int to = where[v], other = 1 - to;
      for ( int k = 0; k < np; k++ ) { int u = pulled_buf[--npulled]; where[u] = other; pw[2] -= g->vwgt[u]; pw[other] += g->vwgt[u]; }
      where[v] = 2; pw[to] -= g->vwgt[v]; pw[2] += g->vwgt[v];
    }
    if ( bestmoves == 0 ) break;
  }
}

/* grow a bisection from a seed, convert to a vertex separator */
static void grow_bisection( const graph_t* g, int seed, int* where, int* pw, int* queue ) {
  const int nv = g->nv;
  int total = 0;
  for ( int v = 0; v < nv; v++ ) { where[v] = 1; total += g->vwgt[v]; }
  int half = total / 2, w0 = 0, qh = 0, qt = 0, scan = 0;
  queue[qt++] = seed; where[seed] = 0; w0 += g->vwgt[seed];
  while ( w0 < half ) {
    if ( qh == qt ) { /* disconnected: restart from an untouched vertex */
      while ( scan < nv && where[scan] == 0 ) scan++;
      if ( scan == nv ) break;
      queue[qt++] = scan; where[scan] = 0; w0 += g->vwgt[scan];
      continue;
    }
    int v = queue[qh++];
    for ( int e = g->xadj[v]; e < g->xadj[v + 1] && w0 < half; e++ ) {
      int u = g->adj[e];
      if ( where[u] == 1 ) { where[u] = 0; w0 += g->vwgt[u]; queue[qt++] = u; }
    }
  }
  /* separator = side-1 vertices adjacent to side 0, or the converse, whichever is lighter */
  int b0 = 0, b1 = 0;
  for ( int v = 0; v < nv; v++ ) {
    int s = where[v], isb = 0;
    for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) if ( where[g->adj[e]] != s ) { isb = 1; break; }
    if ( isb ) { if ( s == 0 ) b0 += g->vwgt[v]; else b1 += g->vwgt[v]; }
  }
  int from = b0 <= b1 ? 0 : 1;
  for ( int v = 0; v < nv; v++ ) {
    if ( where[v] != from ) continue;
    for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) if ( where[g->adj[e]] == 1 - from ) { where[v] = 3; break; }
  }
  pw[0] = pw[1] = pw[2] = 0;
  for ( int v = 0; v < nv; v++ ) { if ( where[v] == 3 ) where[v] = 2; pw[where[v]] += g->vwgt[v]; }
}

static inline long sep_score( const int* pw, int maxpw ) {
  long s = pw[2];
  int mx = pw[0] > pw[1] ? pw[0] : pw[1];
  if ( mx > maxpw ) s += 4L * ( mx - maxpw );
  if ( pw[0] == 0 || pw[1] == 0 ) s += ( long )( pw[0] + pw[1] + pw[2] );
  return s;
}

/* compute a vertex separator of g; where[] out */
static void multilevel_separator( graph_t* g0, int* where0, uint32_t* rng ) {
  enum { MAXLEV = 64 };
  graph_t* G[MAXLEV]; int* CM[MAXLEV];
  int nl = 0;
  G[0] = g0;
  long total = 0;
  for ( int v = 0; v < g0->nv; v++ ) total += g0->vwgt[v];
  const int coarsen_to = 100;
  int maxvwgt = ( int )( 1.5 * total / coarsen_to ); if ( maxvwgt < 2 ) maxvwgt = 2;
  while ( G[nl]->nv > coarsen_to && nl + 1 < MAXLEV ) {
    int* cmap = malloc( sizeof( int ) * G[nl]->nv );
    graph_t* c = coarsen( G[nl], cmap, rng, maxvwgt );
    if ( c->nv > 0.93 * G[nl]->nv ) { graph_free( c ); free( cmap ); break; }
    CM[nl] = cmap; G[nl + 1] = c; nl++;
  }
  const double ub_env = g_knobs.ub;     /* default 1.40: measured with 4 starts: 1.40 beats 1.20 / 1.30 / 1.50 on 128^3, 96^3, 2-D 1024^2 and 27-pt 64^3 (fill falls when a side may hold up to 70 %) */
  const double ub = ub_env;
  int maxpw = ( int )( ub * total / 2.0 ) + 1;
  /* initial separators on the coarsest graph */
  graph_t* gc = G[nl];
  int cnv = gc->nv;
  int* where = malloc( sizeof( int ) * cnv );
  int* best = malloc( sizeof( int ) * cnv );
  int* queue = malloc( sizeof( int ) * cnv );
  int* locked = malloc( sizeof( int ) * g0->nv );
  int* pulled = malloc( sizeof( int ) * ( g0->nv + 1 ) );
  move_t* moves = malloc( sizeof( move_t ) * ( g0->nv + 1 ) );
  heap_t H[2];
  heap_init( &H[0], g0->nv ); heap_init( &H[1], g0->nv );
  long bestscore = -1; int bestpw[3] = {0, 0, 0};
  const int tries_env = g_knobs.tries;
  int ntries = cnv < tries_env ? cnv : tries_env;
  for ( int t = 0; t < ntries; t++ ) {
    int seed = rng_next( rng ) % ( uint32_t )cnv;
    int pw[3];
    grow_bisection( gc, seed, where, pw, queue );
    node_fm( gc, where, pw, maxpw, 4, H, locked, pulled, moves );
    long sc = sep_score( pw, maxpw );
    if ( bestscore < 0 || sc < bestscore ) { bestscore = sc; memcpy( best, where, sizeof( int ) * cnv ); memcpy( bestpw, pw, sizeof( pw ) ); }
  }
  free( where ); free( queue );
  int* cur = best; int pw[3] = { bestpw[0], bestpw[1], bestpw[2] };
  /* uncoarsen + refine */
  for ( int l = nl - 1; l >= 0; l-- ) {
    graph_t* gf = G[l];
    int* wf = ( l == 0 ) ? where0 : malloc( sizeof( int ) * gf->nv );
    for ( int v = 0; v < gf->nv; v++ ) wf[v] = cur[CM[l][v]];
    free( cur ); cur = wf;
    /* projected separators are thick: vertices with no neighbour on one side can leave for free */
    const int fm_big = g_knobs.fm_big, fm_small = g_knobs.fm_small;
    node_fm( gf, cur, pw, maxpw, gf->nv > 200000 ? fm_big : fm_small, H, locked, pulled, moves );
    graph_free( G[l + 1] ); free( CM[l] );
  }
  if ( nl == 0 ) { memcpy( where0, cur, sizeof( int ) * g0->nv ); free( cur ); }
  heap_free( &H[0] ); heap_free( &H[1] );
  free( locked ); free( pulled ); free( moves );
}

/* ------------------------------------------------ leaf: exact minimum degree
 * on the leaf's own graph using adjacency sets as bit masks (nv <= 64*LW) */
static void leaf_mindeg( const graph_t* g, int* order ) {
  const int nv = g->nv;
  const int W = ( nv + 63 ) / 64;
  uint64_t* A = calloc( ( size_t )nv * W, sizeof( uint64_t ) );
  char* done = calloc( nv, 1 );
  for ( int v = 0; v < nv; v++ )
    for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) { int u = g->adj[e]; A[( size_t )v * W + ( u >> 6 )] |= 1ull << ( u & 63 ); }
  for ( int k = 0; k < nv; k++ ) {
    int bv = -1, bd = 1 << 30;
    for ( int v = 0; v < nv; v++ ) {
      if ( done[v] ) continue;
      int d = 0;
      for ( int w = 0; w < W; w++ ) d += __builtin_popcountll( A[( size_t )v * W + w] );
      if ( d < bd ) { bd = d; bv = v; }
    }
    order[k] = bv; done[bv] = 1;
    uint64_t* Av = &A[( size_t )bv * W];
    /* eliminate: neighbours become a clique, bv disappears */
    for ( int w = 0; w < W; w++ ) {
      uint64_t m = Av[w];
      while ( m ) {
        int b = __builtin_ctzll( m ); m &= m - 1;
        int u = w * 64 + b;
        uint64_t* Au = &A[( size_t )u * W];
        for ( int x = 0; x < W; x++ ) Au[x] |= Av[x];
        Au[u >> 6] &= ~( 1ull << ( u & 63 ) );
        Au[bv >> 6] &= ~( 1ull << ( bv & 63 ) );
      }
    }
  }
  free( A ); free( done );
}

/* induced subgraph on vertices with where[v]==side */
static graph_t* induced( const graph_t* g, const int* where, int side, int* newid, int** label_out, const int* label ) {
  int nv = 0;
  for ( int v = 0; v < g->nv; v++ ) if ( where[v] == side ) newid[v] = nv++;
  graph_t* s = calloc( 1, sizeof( graph_t ) );
  s->nv = nv; s->xadj = malloc( sizeof( int ) * ( nv + 1 ) ); s->vwgt = malloc( sizeof( int ) * ( nv > 0 ? nv : 1 ) );
  int* lab = malloc( sizeof( int ) * ( nv > 0 ? nv : 1 ) );
  int ne = 0;
  for ( int v = 0; v < g->nv; v++ ) if ( where[v] == side ) for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) if ( where[g->adj[e]] == side ) ne++;
  s->adj = malloc( sizeof( int ) * ( ne > 0 ? ne : 1 ) ); s->ewgt = malloc( sizeof( int ) * ( ne > 0 ? ne : 1 ) );
  ne = 0; int k = 0;
  for ( int v = 0; v < g->nv; v++ ) {
    if ( where[v] != side ) continue;
    s->xadj[k] = ne; s->vwgt[k] = 1; lab[k] = label[v]; k++;
    for ( int e = g->xadj[v]; e < g->xadj[v + 1]; e++ ) { int u = g->adj[e]; if ( where[u] == side ) { s->adj[ne] = newid[u]; s->ewgt[ne] = 1; ne++; } }
  }
  s->xadj[nv] = ne; s->ne = ne;
  *label_out = lab;
  return s;
}

/* g and label are consumed (freed) */
static void nd_rec( graph_t* g, int* label, int* perm, int first, int leaf, int depth ) {
  const int nv = g->nv;
  if ( nv == 0 ) { graph_free( g ); free( label ); return; }
  if ( nv <= leaf ) {
    int* ord = malloc( sizeof( int ) * nv );
    leaf_mindeg( g, ord );
    for ( int k = 0; k < nv; k++ ) perm[first + k] = label[ord[k]];
    free( ord ); graph_free( g ); free( label );
    return;
  }
  int* where = malloc( sizeof( int ) * nv );
  int n0 = 0, n1 = 0, n2 = 0;
  {
    /* Multi-start on the large subgraphs: the top separators decide the fill (their fronts cost ~|S|^3), and one
     * multilevel run is a lottery (128^3: 1.35e13 ... 1.80e13 flops depending on the balance bound alone).  The
     * top of the recursion has idle cores anyway (1, 2, 4 ... subgraphs), so R independent runs with different
     * seeds go out as tasks and the smallest separator (with a mild imbalance penalty) wins; ties go to the
     * lower run index, so the result does not depend on the number of threads. */
    const int starts_env = g_knobs.starts;
    int R = nv >= 1000000 ? 4 * starts_env : nv >= 200000 ? 2 * starts_env : nv >= 30000 ? starts_env : 1;    /* more where a front costs the most */
    if ( R > 16 ) R = 16;
    if ( starts_env == 1 ) R = 1;          /* B200_ND_STARTS=1: the single-start ordering of mid-round (reproducibility of earlier measurements) */
    int* cand[16]; double score[16];
    for ( int r = 0; r < R; r++ ) cand[r] = r == 0 ? where : malloc( sizeof( int ) * nv );
    for ( int r = 0; r < R; r++ ) {
      #pragma omp task default(shared) firstprivate(r) if ( R > 1 )
      {
        uint32_t rng = 0x12345u + 2654435761u * ( uint32_t )first + 40503u * ( uint32_t )nv + 0x9E3779B9u * ( uint32_t )r;
        multilevel_separator( g, cand[r], &rng );
        long c0 = 0, c1 = 0, c2 = 0;
        for ( int v = 0; v < nv; v++ ) { if ( cand[r][v] == 0 ) c0++; else if ( cand[r][v] == 1 ) c1++; else c2++; }
        const double imb = ( c0 + c1 ) > 0 ? ( double )( c0 > c1 ? c0 - c1 : c1 - c0 ) / ( double )( c0 + c1 ) : 1.0;
        const double pen = g_knobs.pen;
        score[r] = ( c0 == 0 || c1 == 0 ) ? 1e300 : ( double )c2 * ( 1.0 + pen * imb * imb );
      }
    }
    #pragma omp taskwait
    int best = 0;
    for ( int r = 1; r < R; r++ ) if ( score[r] < score[best] ) best = r;
    if ( best != 0 ) memcpy( where, cand[best], sizeof( int ) * nv );
    for ( int r = 1; r < R; r++ ) free( cand[r] );
  }
  for ( int v = 0; v < nv; v++ ) { if ( where[v] == 0 ) n0++; else if ( where[v] == 1 ) n1++; else n2++; }
  { const int verbose = g_knobs.verbose; if ( verbose && nv >= 20000 ) fprintf( stderr, "nd depth %d: nv=%d -> %d | %d | sep %d\n", depth, nv, n0, n1, n2 ); }
  if ( n0 == 0 || n1 == 0 ) {
    /* bisection failed (e.g. a clique): order what is left by degree-sorted natural order in chunks */
    if ( nv <= 4096 ) {
      int* ord = malloc( sizeof( int ) * nv ); leaf_mindeg( g, ord );
      for ( int k = 0; k < nv; k++ ) perm[first + k] = label[ord[k]];
      free( ord );
    } else for ( int k = 0; k < nv; k++ ) perm[first + k] = label[k];
    free( where ); graph_free( g ); free( label );
    return;
  }
  int* newid = malloc( sizeof( int ) * nv );
  int *lab0, *lab1;
  graph_t* g0 = induced( g, where, 0, newid, &lab0, label );
  graph_t* g1 = induced( g, where, 1, newid, &lab1, label );
  int k = first + n0 + n1;
  for ( int v = 0; v < nv; v++ ) if ( where[v] == 2 ) perm[k++] = label[v];
  free( newid ); free( where ); graph_free( g ); free( label );
  if ( nv > 20000 ) {
    #pragma omp task default(shared) firstprivate(g0, lab0, first, leaf, depth)
    nd_rec( g0, lab0, perm, first, leaf, depth + 1 );
    #pragma omp task default(shared) firstprivate(g1, lab1, first, n0, leaf, depth)
    nd_rec( g1, lab1, perm, first + n0, leaf, depth + 1 );
    #pragma omp taskwait
  } else {
    nd_rec( g0, lab0, perm, first, leaf, depth + 1 );
    nd_rec( g1, lab1, perm, first + n0, leaf, depth + 1 );
  }
}

/* xadj/adj: symmetric graph without self loops. perm[k] = vertex eliminated k-th. */
int b200_order_nd( int n, const int* xadj, const int* adj, int leaf, int threads, int* perm ) {
  if ( n < 0 || !perm ) return B200_ERR_ARG;
  if ( n == 0 ) return B200_OK;
  if ( leaf <= 0 ) leaf = 96;
  graph_t* g = calloc( 1, sizeof( graph_t ) );
  g->nv = n; g->ne = xadj[n];
  g->xadj = malloc( sizeof( int ) * ( n + 1 ) ); memcpy( g->xadj, xadj, sizeof( int ) * ( n + 1 ) );
  g->adj = malloc( sizeof( int ) * ( g->ne > 0 ? g->ne : 1 ) ); memcpy( g->adj, adj, sizeof( int ) * g->ne );
  g->vwgt = malloc( sizeof( int ) * n ); g->ewgt = malloc( sizeof( int ) * ( g->ne > 0 ? g->ne : 1 ) );
  for ( int i = 0; i < n; i++ ) g->vwgt[i] = 1;
  for ( int i = 0; i < g->ne; i++ ) g->ewgt[i] = 1;
  int* label = malloc( sizeof( int ) * n );
  for ( int i = 0; i < n; i++ ) label[i] = i;
  read_knobs();
#ifdef _OPENMP
  int nt = threads > 0 ? threads : omp_get_max_threads();
  #pragma omp parallel num_threads(nt)
  #pragma omp single
#else
  ( void )threads;
#endif
  nd_rec( g, label, perm, 0, leaf, 0 );
  return B200_OK;
}
