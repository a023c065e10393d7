/* b200_symbolic.c -- host-side symbolic analysis for the B200 multifrontal solver.
 *
 * This is the "analyze" phase of the plugin (src/solvers.c:471-481 ->
 * table.analyze): pattern only, values never read (src/solvers.h:95).
 * The reference performs it inside un-vendored cholmod_analyze
 * (src/solver_cholmod.c:257) / umfpack_di_symbolic (src/solver_umfpack.c:73);
 * here it is re-stated from the published algorithms:
 *   elimination tree        -- Liu 1990 (path compression with virtual ancestors)
 *   column counts           -- Gilbert, Ng, Peyton 1994 (skeleton / least common ancestor)
 *   fundamental supernodes  -- Liu, Ng, Peyton 1993
 *   relaxed amalgamation    -- Ashcraft & Grimes 1989, with CHOLMOD's documented
 *                              default thresholds nrelax {4,16,48}, zrelax {0.8,0.1,0.05}
 *   frontal row structures, relative indices, assembly map -- Duff & Reid multifrontal
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdint.h>
#include <sys/time.h>
#include "solver_b200.h"

static double now_s( void ) { struct timeval tv; gettimeofday( &tv, NULL ); return tv.tv_sec + 1e-6 * tv.tv_usec; }

void b200_default_options( b200_options_t* o ) {
  memset( o, 0, sizeof( *o ) );
  o->ordering = B200_ORDER_ND;
  o->nd_leaf = 96;
  o->relax = 1;
  o->nrelax[0] = 4; o->nrelax[1] = 16; o->nrelax[2] = 48;
  o->zrelax[0] = 0.8; o->zrelax[1] = 0.1; o->zrelax[2] = 0.05;
  o->threads = 0;
}

void b200_symbolic_free( b200_symbolic_t* S ) {
  if ( !S ) return;
  free( S->perm ); free( S->iperm ); free( S->parent ); free( S->colcount );
  free( S->sn_start ); free( S->sn_parent ); free( S->sn_depth );
  free( S->rowptr ); free( S->rowidx ); free( S->relidx ); free( S->lptr ); free( S->uptr );
  free( S->amap ); free( S->cbptr ); free( S->spmv_ptr ); free( S->spmv_col ); free( S->spmv_src );
  free( S );
}

/* Liu's algorithm on an upper-triangular CSC pattern (rows i<j in column j). */
int b200_etree( int n, const int* colptr, const int* rowidx, int* parent ) {
  int* anc = malloc( sizeof( int ) * ( n > 0 ? n : 1 ) );
  if ( !anc ) return B200_ERR_NOMEM;
  for ( int j = 0; j < n; j++ ) {
    parent[j] = -1; anc[j] = -1;
    for ( int p = colptr[j]; p < colptr[j + 1]; p++ ) {
      int i = rowidx[p];
      while ( i != -1 && i < j ) {
        int nx = anc[i];
        anc[i] = j;
        if ( nx == -1 ) parent[i] = j;
        i = nx;
      }
    }
  }
  free( anc );
  return B200_OK;
}

/* postorder of a forest; children visited in the order given by key (ascending, ties by index).
 * post[k] = node visited k-th. */
static void postorder( int n, const int* parent, const int* key, int* post ) {
  int* head = malloc( sizeof( int ) * ( n + 1 ) );
  int* next = malloc( sizeof( int ) * ( n + 1 ) );
  int* stack = malloc( sizeof( int ) * ( n + 1 ) );
  int* order = malloc( sizeof( int ) * ( n + 1 ) );
  for ( int i = 0; i <= n; i++ ) head[i] = -1;
  /* build child lists so that popping from head yields ascending key: insert in descending key order */
  if ( key ) {
    /* counting sort by key (keys are 1..n) */
    int* cnt = calloc( n + 2, sizeof( int ) );
    for ( int i = 0; i < n; i++ ) cnt[key[i] + 1 > n + 1 ? n + 1 : key[i] + 1]++;
    for ( int i = 1; i <= n + 1; i++ ) cnt[i] += cnt[i - 1];
    for ( int i = 0; i < n; i++ ) { int k = key[i] > n ? n : key[i]; order[cnt[k]++] = i; }
    free( cnt );
  } else for ( int i = 0; i < n; i++ ) order[i] = i;
  for ( int t = n - 1; t >= 0; t-- ) { /* descending key => lists end up ascending */
    int v = order[t];
    int p = parent[v] < 0 ? n : parent[v];
    next[v] = head[p]; head[p] = v;
  }
  int k = 0;
  for ( int r = head[n]; r != -1; r = next[r] ) {
    int top = 0; stack[0] = r;
    while ( top >= 0 ) {
      int v = stack[top];
      int c = head[v];
      if ( c == -1 ) { post[k++] = v; top--; }
      else { head[v] = next[c]; stack[++top] = c; }
    }
  }
  free( head ); free( next ); free( stack ); free( order );
}

/* Gilbert-Ng-Peyton column counts.  post[k] = k-th node of a postorder of the etree
 * (NULL: the matrix is already postordered).  lp/li: strictly-lower pattern by column. */
static void colcounts_post( int n, const int* parent, const int* post, const int* lp, const int* li, int* cc ) {
  int* first = malloc( sizeof( int ) * n );
  int* maxfirst = malloc( sizeof( int ) * n );
  int* prevleaf = malloc( sizeof( int ) * n );
  int* anc = malloc( sizeof( int ) * n );
  for ( int j = 0; j < n; j++ ) { first[j] = -1; maxfirst[j] = -1; prevleaf[j] = -1; anc[j] = j; }
  for ( int k = 0; k < n; k++ ) {
    int j = post ? post[k] : k;
    cc[j] = ( first[j] == -1 ) ? 1 : 0;   /* leaves of the etree start at 1 */
    for ( ; j != -1 && first[j] == -1; j = parent[j] ) first[j] = k;
  }
  for ( int k = 0; k < n; k++ ) {
    const int j = post ? post[k] : k;
    if ( parent[j] != -1 ) cc[parent[j]]--;            /* j is not a root */
    for ( int p = lp[j]; p < lp[j + 1]; p++ ) {
      int i = li[p];
      if ( i <= j || first[j] <= maxfirst[i] ) continue; /* j is not a leaf of the row subtree of i */
      maxfirst[i] = first[j];
      int jprev = prevleaf[i];
      prevleaf[i] = j;
      cc[j]++;
      if ( jprev != -1 ) {
        int q = jprev;
        while ( q != anc[q] ) q = anc[q];
        for ( int s = jprev; s != q; ) { int sp = anc[s]; anc[s] = q; s = sp; }
        cc[q]--;                                        /* q = least common ancestor(jprev, j) */
      }
    }
    if ( parent[j] != -1 ) anc[j] = parent[j];
  }
  for ( int k = 0; k < n; k++ ) { const int j = post ? post[k] : k; if ( parent[j] != -1 ) cc[parent[j]] += cc[j]; }
  free( first ); free( maxfirst ); free( prevleaf ); free( anc );
}
static void colcounts_postordered( int n, const int* parent, const int* lp, const int* li, int* cc ) { colcounts_post( n, parent, NULL, lp, li, cc ); }

static int cmp_int( const void* a, const void* b ) { int x = *( const int* )a, y = *( const int* )b; return ( x > y ) - ( x < y ); }

/* pattern of the permuted matrix, strictly lower part by column, sorted, de-duplicated.
 * For kind LU the pattern of A+A' is used. */
static int build_lower_pattern( int n, const unsigned int* cp, const unsigned int* ri, int kind, const int* iperm, int** lp_out, int** li_out ) {
  int* cnt = calloc( n + 1, sizeof( int ) );
  int64_t nz = cp[n];
  for ( int j = 0; j < n; j++ )
    for ( unsigned int p = cp[j]; p < cp[j + 1]; p++ ) {
      int i = ( int )ri[p];
      if ( i == j ) continue;
      if ( kind != B200_KIND_LU && i > j ) continue;
      int a = iperm[i], b = iperm[j];
      int c = a < b ? a : b;
      cnt[c + 1]++;
    }
  ( void )nz;
  for ( int j = 0; j < n; j++ ) cnt[j + 1] += cnt[j];
  int* lp = malloc( sizeof( int ) * ( n + 1 ) );
  memcpy( lp, cnt, sizeof( int ) * ( n + 1 ) );
  int* li = malloc( sizeof( int ) * ( lp[n] > 0 ? lp[n] : 1 ) );
  for ( int j = 0; j < n; j++ )
    for ( unsigned int p = cp[j]; p < cp[j + 1]; p++ ) {
      int i = ( int )ri[p];
      if ( i == j ) continue;
      if ( kind != B200_KIND_LU && i > j ) continue;
      int a = iperm[i], b = iperm[j];
      int c = a < b ? a : b, r = a < b ? b : a;
      li[cnt[c]++] = r;
    }
  free( cnt );
  /* sort + unique per column */
  int* nlp = malloc( sizeof( int ) * ( n + 1 ) );
  int w = 0;
  for ( int j = 0; j < n; j++ ) {
    int s = lp[j], e = lp[j + 1];
    qsort( li + s, e - s, sizeof( int ), cmp_int );
    nlp[j] = w;
    for ( int p = s; p < e; p++ ) if ( p == s || li[p] != li[p - 1] ) li[w++] = li[p];
  }
  nlp[n] = w;
  free( lp );
  *lp_out = nlp; *li_out = li;
  return 0;
}

/* transpose a strictly-lower CSC pattern into strictly-upper CSC (rows i<j in column j) */
static void transpose_pattern( int n, const int* lp, const int* li, int** up_out, int** ui_out ) {
  int* up = calloc( n + 1, sizeof( int ) );
  for ( int p = 0; p < lp[n]; p++ ) up[li[p] + 1]++;
  for ( int j = 0; j < n; j++ ) up[j + 1] += up[j];
  int* ui = malloc( sizeof( int ) * ( lp[n] > 0 ? lp[n] : 1 ) );
  int* w = malloc( sizeof( int ) * ( n + 1 ) );
  memcpy( w, up, sizeof( int ) * ( n + 1 ) );
  for ( int j = 0; j < n; j++ ) for ( int p = lp[j]; p < lp[j + 1]; p++ ) ui[w[li[p]]++] = j;
  free( w );
  *up_out = up; *ui_out = ui;
}

static inline double panel_nnz( double k, double f ) { return k * f - k * ( k - 1 ) / 2.0; }
static inline double panel_flops( double k, double f ) { /* sum_{j<k} (f-j)^2 */
  double a = f, b = f - k; /* sum_{m=b+1}^{a} m^2 */
  return ( a * ( a + 1 ) * ( 2 * a + 1 ) - b * ( b + 1 ) * ( 2 * b + 1 ) ) / 6.0;
}

b200_symbolic_t* b200_analyze( int n, const unsigned int* cp, const unsigned int* ri, int kind, const b200_options_t* opts_in ) {
  b200_options_t opts;
  if ( opts_in ) opts = *opts_in; else b200_default_options( &opts );
  if ( n <= 0 || !cp || ( kind != B200_KIND_CHOLESKY && kind != B200_KIND_LU && kind != B200_KIND_LDLT ) ) return NULL;
  if ( ( uint64_t )cp[n] > 0x7fffffffu ) return NULL;
  /* a malformed caller must end here, not in an out-of-bounds write on the host or the device */
  if ( cp[0] != 0 || ( cp[n] > 0 && !ri ) ) { fprintf( stderr, "b200_analyze: bad column pointers\n" ); return NULL; }
  for ( int j = 0; j < n; j++ ) if ( cp[j + 1] < cp[j] ) { fprintf( stderr, "b200_analyze: column pointers are not monotone (column %d)\n", j ); return NULL; }
  for ( unsigned int p = 0; p < cp[n]; p++ ) if ( ri[p] >= ( unsigned int )n ) { fprintf( stderr, "b200_analyze: row index %u out of range at entry %u\n", ri[p], p ); return NULL; }
  double t0 = now_s();
  b200_symbolic_t* S = calloc( 1, sizeof( *S ) );
  S->n = n; S->kind = kind; S->nnzA = cp[n];
  S->perm = malloc( sizeof( int ) * n ); S->iperm = malloc( sizeof( int ) * n );

  /* ---- 1. ordering on the graph of A+A' */
  if ( opts.ordering == B200_ORDER_GIVEN && opts.user_perm ) {
    memcpy( S->perm, opts.user_perm, sizeof( int ) * n );
  } else if ( opts.ordering == B200_ORDER_NATURAL ) {
    for ( int i = 0; i < n; i++ ) S->perm[i] = i;
  } else {
    int* id = malloc( sizeof( int ) * n );
    for ( int i = 0; i < n; i++ ) id[i] = i;
    int *lp, *li, *up, *ui;
    build_lower_pattern( n, cp, ri, kind, id, &lp, &li );
    transpose_pattern( n, lp, li, &up, &ui );
    int* xadj = malloc( sizeof( int ) * ( n + 1 ) );
    int* adj = malloc( sizeof( int ) * ( 2 * ( size_t )lp[n] + 1 ) );
    int e = 0;
    for ( int j = 0; j < n; j++ ) {
      xadj[j] = e;
      for ( int p = up[j]; p < up[j + 1]; p++ ) adj[e++] = ui[p];
      for ( int p = lp[j]; p < lp[j + 1]; p++ ) adj[e++] = li[p];
    }
    xadj[n] = e;
    free( lp ); free( li ); free( up ); free( ui ); free( id );
    b200_order_nd( n, xadj, adj, opts.nd_leaf, opts.threads, S->perm );
    free( xadj ); free( adj );
  }
  for ( int i = 0; i < n; i++ ) S->iperm[i] = -1;
  for ( int k = 0; k < n; k++ ) {
    int o = S->perm[k];
    if ( o < 0 || o >= n || S->iperm[o] != -1 ) { fprintf( stderr, "b200_analyze: ordering is not a permutation\n" ); b200_symbolic_free( S ); return NULL; }
    S->iperm[o] = k;
  }
  S->t_order = now_s() - t0;
  double t1 = now_s();

  /* ---- 2. etree, postorder (largest-front child last), relabel */
  int *lp, *li, *up, *ui;
  int* parent = malloc( sizeof( int ) * n );
  int* cc = malloc( sizeof( int ) * n );
  int* post = malloc( sizeof( int ) * n );
  for ( int round = 0; round < 2; round++ ) {
    build_lower_pattern( n, cp, ri, kind, S->iperm, &lp, &li );
    transpose_pattern( n, lp, li, &up, &ui );
    b200_etree( n, up, ui, parent );
    if ( round == 0 ) postorder( n, parent, NULL, post );
    else {
      /* second round: matrix is postordered already, so counts are valid; re-postorder by count */
      colcounts_postordered( n, parent, lp, li, cc );
      postorder( n, parent, cc, post );
    }
    /* compose: new position k holds old (current) column post[k] */
    int* np = malloc( sizeof( int ) * n );
    for ( int k = 0; k < n; k++ ) np[k] = S->perm[post[k]];
    memcpy( S->perm, np, sizeof( int ) * n ); free( np );
    for ( int k = 0; k < n; k++ ) S->iperm[S->perm[k]] = k;
    free( lp ); free( li ); free( up ); free( ui );
  }
  build_lower_pattern( n, cp, ri, kind, S->iperm, &lp, &li );
  transpose_pattern( n, lp, li, &up, &ui );
  b200_etree( n, up, ui, parent );
  colcounts_postordered( n, parent, lp, li, cc );
  S->parent = parent; S->colcount = cc;
  free( post );
  S->nnzL = 0; S->flops = 0;
  for ( int j = 0; j < n; j++ ) { S->nnzL += cc[j]; S->flops += ( double )cc[j] * cc[j]; }

  /* ---- 3. fundamental supernodes */
  int* nchild = calloc( n, sizeof( int ) );
  for ( int j = 0; j < n; j++ ) if ( parent[j] >= 0 ) nchild[parent[j]]++;
  int* fstart = malloc( sizeof( int ) * ( n + 1 ) );
  int nf = 0;
  for ( int j = 0; j < n; j++ ) {
    int cont = j > 0 && parent[j - 1] == j && cc[j - 1] == cc[j] + 1 && nchild[j] == 1;
    if ( !cont ) fstart[nf++] = j;
  }
  fstart[nf] = n;
  S->nfund = nf;
  free( nchild );

  /* ---- 4. relaxed amalgamation on the supernodal tree (Ashcraft-Grimes): a child may be absorbed by
   * its parent whichever position it has among the siblings; the columns are renumbered afterwards so
   * that every merged group is contiguous (any topological order of the etree gives the same fill).
   * A merge is accepted when (a) the merged front still fits one 64x64 device tile, or (b) the fraction
   * of explicit zeros stays under CHOLMOD's documented tiers (nrelax {4,16,48}, zrelax {0.8,0.1,0.05}). */
  int ns = 0;
  int* sn_start = malloc( sizeof( int ) * ( nf + 1 ) );
  if ( !opts.relax ) {
    for ( int s = 0; s < nf; s++ ) sn_start[ns++] = fstart[s];
    sn_start[ns] = n;
  } else {
    int* col2f = malloc( sizeof( int ) * n );
    for ( int s = 0; s < nf; s++ ) for ( int j = fstart[s]; j < fstart[s + 1]; j++ ) col2f[j] = s;
    int* fpar = malloc( sizeof( int ) * nf );
    for ( int s = 0; s < nf; s++ ) { int p = parent[fstart[s + 1] - 1]; fpar[s] = p < 0 ? -1 : col2f[p]; }
    double* gk = malloc( sizeof( double ) * nf ); double* gf = malloc( sizeof( double ) * nf ); double* gz = malloc( sizeof( double ) * nf );
    int* into = malloc( sizeof( int ) * nf );        /* into[c] = supernode that absorbed c, -1 if none */
    int* head = malloc( sizeof( int ) * nf ); int* tail = malloc( sizeof( int ) * nf ); int* nxt = malloc( sizeof( int ) * nf );
    for ( int s = 0; s < nf; s++ ) { gk[s] = fstart[s + 1] - fstart[s]; gf[s] = cc[fstart[s]]; gz[s] = 0; into[s] = -1; head[s] = tail[s] = -1; nxt[s] = -1; }
    for ( int s = 0; s < nf; s++ ) { int p = fpar[s]; if ( p >= 0 ) { if ( head[p] < 0 ) head[p] = s; else nxt[tail[p]] = s; tail[p] = s; } }
    int* kids = malloc( sizeof( int ) * nf ); int nk;
    for ( int p = 0; p < nf; p++ ) {
      nk = 0;
      for ( int c = head[p]; c != -1; c = nxt[c] ) kids[nk++] = c;
      if ( nk == 0 ) continue;
      /* largest front first: the chain child is tried before the small siblings (insertion sort, lists are short) */
      for ( int a = 1; a < nk; a++ ) { int v = kids[a], q = a - 1; while ( q >= 0 && gf[kids[q]] < gf[v] ) { kids[q + 1] = kids[q]; q--; } kids[q + 1] = v; }
      int nh = -1, nt = -1;   /* new child list of p */
      for ( int a = 0; a < nk; a++ ) {
        const int c = kids[a];
        const double kc = gk[c], fc = gf[c], zc = gz[c], kp = gk[p], fp = gf[p], zp = gz[p];
        const double k = kc + kp, f = kc + fp;
        const double tot = panel_nnz( k, f );
        const double z = tot - ( panel_nnz( kc, fc ) - zc ) - ( panel_nnz( kp, fp ) - zp );
        const double frac = z / tot;
        int ok;
        if ( f <= 64 ) ok = 1;
        else if ( k <= opts.nrelax[0] ) ok = 1;
        else if ( k <= opts.nrelax[1] ) ok = frac < opts.zrelax[0];
        else if ( k <= opts.nrelax[2] ) ok = frac < opts.zrelax[1];
        else ok = frac < opts.zrelax[2];
        if ( ok ) {
          into[c] = p; gk[p] = k; gf[p] = f; gz[p] = z;
          if ( head[c] >= 0 ) { if ( nh < 0 ) nh = head[c]; else nxt[nt] = head[c]; nt = tail[c]; }   /* adopt the grandchildren */
        } else { if ( nh < 0 ) nh = c; else nxt[nt] = c; nt = c; nxt[c] = -1; }
      }
      if ( nt >= 0 ) nxt[nt] = -1;
      head[p] = nh; tail[p] = nt;
    }
    free( kids );
    /* group root of every fundamental supernode */
    int* root = malloc( sizeof( int ) * nf );
    for ( int s = nf - 1; s >= 0; s-- ) root[s] = into[s] < 0 ? s : root[into[s]];
    int* gid = malloc( sizeof( int ) * nf ); int ng = 0;
    for ( int s = 0; s < nf; s++ ) if ( into[s] < 0 ) gid[s] = ng++;
    int* gpar = malloc( sizeof( int ) * ( ng > 0 ? ng : 1 ) ); int* gkey = malloc( sizeof( int ) * ( ng > 0 ? ng : 1 ) ); int* gpost = malloc( sizeof( int ) * ( ng > 0 ? ng : 1 ) );
    for ( int s = 0; s < nf; s++ ) if ( into[s] < 0 ) { gpar[gid[s]] = fpar[s] < 0 ? -1 : gid[root[fpar[s]]]; double f = gf[s]; gkey[gid[s]] = f > ng ? ng : ( int )f; }
    postorder( ng, gpar, gkey, gpost );
    /* members of each group in ascending (= topological) order */
    int* mcnt = calloc( ng + 1, sizeof( int ) );
    for ( int s = 0; s < nf; s++ ) mcnt[gid[root[s]] + 1]++;
    for ( int g = 0; g < ng; g++ ) mcnt[g + 1] += mcnt[g];
    int* mem = malloc( sizeof( int ) * nf );
    { int* w = malloc( sizeof( int ) * ( ng + 1 ) ); memcpy( w, mcnt, sizeof( int ) * ( ng + 1 ) ); for ( int s = 0; s < nf; s++ ) mem[w[gid[root[s]]]++] = s; free( w ); }
    /* new column order */
    int* np = malloc( sizeof( int ) * n ); int pos = 0;
    for ( int t = 0; t < ng; t++ ) {
      const int g = gpost[t];
      sn_start[ns++] = pos;
      for ( int q = mcnt[g]; q < mcnt[g + 1]; q++ ) { const int fs = mem[q]; for ( int j = fstart[fs]; j < fstart[fs + 1]; j++ ) np[pos++] = S->perm[j]; }
    }
    sn_start[ns] = n;
    memcpy( S->perm, np, sizeof( int ) * n ); free( np );
    for ( int k = 0; k < n; k++ ) S->iperm[S->perm[k]] = k;
    free( col2f ); free( fpar ); free( gk ); free( gf ); free( gz ); free( into ); free( head ); free( tail ); free( nxt );
    free( root ); free( gid ); free( gpar ); free( gkey ); free( gpost ); free( mcnt ); free( mem );
    /* the renumbering is a topological reordering: same fill, but the arrays must follow the new labels */
    free( lp ); free( li ); free( up ); free( ui );
    build_lower_pattern( n, cp, ri, kind, S->iperm, &lp, &li );
    transpose_pattern( n, lp, li, &up, &ui );
    b200_etree( n, up, ui, parent );
    /* the new numbering is topological but not a postorder (siblings' subtrees interleave): give GNP one */
    { int* post2 = malloc( sizeof( int ) * n ); postorder( n, parent, NULL, post2 ); colcounts_post( n, parent, post2, lp, li, cc ); free( post2 ); }
  }
  free( fstart );
  S->ns = ns; S->sn_start = realloc( sn_start, sizeof( int ) * ( ns + 1 ) );
  sn_start = S->sn_start;
  int* col2sn = malloc( sizeof( int ) * n );
  for ( int s = 0; s < ns; s++ ) for ( int j = sn_start[s]; j < sn_start[s + 1]; j++ ) col2sn[j] = s;
  S->sn_parent = malloc( sizeof( int ) * ns );
  for ( int s = 0; s < ns; s++ ) { int p = parent[sn_start[s + 1] - 1]; S->sn_parent[s] = p < 0 ? -1 : col2sn[p]; }

  /* ---- 5. frontal row structures */
  int* chead = malloc( sizeof( int ) * ns ); int* cnext = malloc( sizeof( int ) * ns );
  for ( int s = 0; s < ns; s++ ) chead[s] = -1;
  for ( int s = ns - 1; s >= 0; s-- ) { int p = S->sn_parent[s]; if ( p >= 0 ) { cnext[s] = chead[p]; chead[p] = s; } }
  S->rowptr = malloc( sizeof( int64_t ) * ( ns + 1 ) );
  int64_t cap = ( int64_t )n * 4 + 1024, used = 0;
  int* rows = malloc( sizeof( int ) * cap );
  int* mark = malloc( sizeof( int ) * n );
  for ( int i = 0; i < n; i++ ) mark[i] = -1;
  int maxfront = 0;
  for ( int s = 0; s < ns; s++ ) {
    int c0 = sn_start[s], c1 = sn_start[s + 1], k = c1 - c0;
    /* upper bound of this front: k + sum of candidates */
    int64_t need = k;
    for ( int j = c0; j < c1; j++ ) need += lp[j + 1] - lp[j];
    for ( int c = chead[s]; c != -1; c = cnext[c] ) need += S->rowptr[c + 1] - S->rowptr[c];
    if ( used + need > cap ) { while ( used + need > cap ) cap *= 2; rows = realloc( rows, sizeof( int ) * cap ); }
    S->rowptr[s] = used;
    int* r = rows + used; int f = 0;
    for ( int j = c0; j < c1; j++ ) { r[f++] = j; mark[j] = s; }
    for ( int j = c0; j < c1; j++ )
      for ( int p = lp[j]; p < lp[j + 1]; p++ ) { int i = li[p]; if ( i >= c1 && mark[i] != s ) { mark[i] = s; r[f++] = i; } }
    for ( int c = chead[s]; c != -1; c = cnext[c] ) {
      int kc = sn_start[c + 1] - sn_start[c];
      const int* rc = rows + S->rowptr[c];
      int fc = ( int )( S->rowptr[c + 1] - S->rowptr[c] );
      for ( int q = kc; q < fc; q++ ) { int i = rc[q]; if ( i >= c1 && mark[i] != s ) { mark[i] = s; r[f++] = i; } }
    }
    qsort( r + k, f - k, sizeof( int ), cmp_int );
    used += f; S->rowptr[s + 1] = used;
    if ( f > maxfront ) maxfront = f;
  }
  S->rowidx = realloc( rows, sizeof( int ) * ( used > 0 ? used : 1 ) );
  rows = S->rowidx;
  S->maxfront = maxfront;
  free( lp ); free( li ); free( up ); free( ui );

  /* ---- 6. depth, storage offsets, contribution-block arenas */
  S->sn_depth = malloc( sizeof( int ) * ns );
  int maxdepth = 0;
  for ( int s = ns - 1; s >= 0; s-- ) { int p = S->sn_parent[s]; S->sn_depth[s] = p < 0 ? 0 : S->sn_depth[p] + 1; if ( S->sn_depth[s] > maxdepth ) maxdepth = S->sn_depth[s]; }
  S->maxdepth = maxdepth;
  S->lptr = malloc( sizeof( int64_t ) * ( ns + 1 ) );
  S->uptr = malloc( sizeof( int64_t ) * ( ns + 1 ) );
  S->cbptr = malloc( sizeof( int64_t ) * ns );
  int64_t* dsum = calloc( maxdepth + 1, sizeof( int64_t ) );
  int64_t lo = 0, uo = 0; double sf = 0;
  for ( int s = 0; s < ns; s++ ) {
    int64_t k = sn_start[s + 1] - sn_start[s], f = S->rowptr[s + 1] - S->rowptr[s], u = f - k;
    const int64_t ld = ( f + 1 ) & ~( int64_t )1, ldu = ( u + 1 ) & ~( int64_t )1;   /* even leading dimensions: 16-byte aligned columns */
    S->lptr[s] = lo; lo += ld * k;
    S->uptr[s] = ( kind != B200_KIND_CHOLESKY ) ? S->lptr[s] : 0; ( void )uo;
    sf += panel_flops( ( double )k, ( double )f );
    int d = S->sn_depth[s];
    S->cbptr[s] = dsum[d]; dsum[d] += ldu * u;
  }
  S->lptr[ns] = lo; S->uptr[ns] = ( kind != B200_KIND_CHOLESKY ) ? lo : 0; S->stored_nnzL = lo; S->stored_flops = sf;
  S->cb_arena[0] = S->cb_arena[1] = 0;
  for ( int d = 0; d <= maxdepth; d++ ) if ( dsum[d] > S->cb_arena[d & 1] ) S->cb_arena[d & 1] = dsum[d];
  free( dsum );

  /* ---- 7. relative indices child -> parent front */
  S->relidx = malloc( sizeof( int ) * ( used > 0 ? used : 1 ) );
  for ( int64_t q = 0; q < used; q++ ) S->relidx[q] = -1;
  int* pos = mark; /* reuse: pos[row] = index in the current parent's row list */
  for ( int p = 0; p < ns; p++ ) {
    if ( chead[p] == -1 ) continue;
    const int* rp = rows + S->rowptr[p]; int fp = ( int )( S->rowptr[p + 1] - S->rowptr[p] );
    for ( int q = 0; q < fp; q++ ) pos[rp[q]] = q;
    for ( int c = chead[p]; c != -1; c = cnext[c] ) {
      int kc = sn_start[c + 1] - sn_start[c]; int fc = ( int )( S->rowptr[c + 1] - S->rowptr[c] );
      const int* rc = rows + S->rowptr[c];
      for ( int q = kc; q < fc; q++ ) S->relidx[S->rowptr[c] + q] = pos[rc[q]];
    }
  }
  free( chead ); free( cnext );

  /* ---- 8. assembly map of the input entries */
  S->amap = malloc( sizeof( int64_t ) * ( S->nnzA > 0 ? S->nnzA : 1 ) );
  for ( int j = 0; j < n; j++ ) {
    for ( unsigned int p = cp[j]; p < cp[j + 1]; p++ ) {
      int i = ( int )ri[p];
      if ( kind != B200_KIND_LU && i > j ) { S->amap[p] = -1; continue; }
      int pr = S->iperm[i], pc = S->iperm[j];   /* entry A(i,j) -> (pr,pc) of the permuted matrix */
      int r, c, upper = 0;
      if ( kind != B200_KIND_LU ) { r = pr > pc ? pr : pc; c = pr > pc ? pc : pr; }
      else if ( pr >= pc ) { r = pr; c = pc; }
      else { r = pc; c = pr; upper = 1; }        /* U(pr,pc): row pr belongs to supernode of pr */
      int s = col2sn[c];
      int c0 = sn_start[s], k = sn_start[s + 1] - c0;
      int64_t f = S->rowptr[s + 1] - S->rowptr[s];
      const int64_t ld = ( f + 1 ) & ~( int64_t )1;
      int lr;
      if ( r < c0 + k ) lr = r - c0;
      else {
        const int* rs = rows + S->rowptr[s];
        int a = k, b = ( int )f - 1; lr = -1;
        while ( a <= b ) { int m = ( a + b ) >> 1; if ( rs[m] == r ) { lr = m; break; } else if ( rs[m] < r ) a = m + 1; else b = m - 1; }
        if ( lr < 0 ) { fprintf( stderr, "b200_analyze: internal error, row %d missing from front %d\n", r, s ); b200_symbolic_free( S ); free( col2sn ); return NULL; }
      }
      /* the L panel and (LU) the U' panel share one layout: column = pivot, row = front position */
      S->amap[p] = S->lptr[s] + ( int64_t )( c - c0 ) * ld + lr;
      if ( upper ) S->amap[p] |= ( int64_t )1 << 62;
    }
  }
  free( col2sn ); free( mark );

  /* ---- 9. row-wise copy of the full pattern (original numbering) for the residual b - A x */
  {
    int64_t* rp = calloc( ( size_t )n + 1, sizeof( int64_t ) );
    for ( int j = 0; j < n; j++ )
      for ( unsigned int p = cp[j]; p < cp[j + 1]; p++ ) {
        int i = ( int )ri[p];
        if ( kind != B200_KIND_LU ) { if ( i > j ) continue; rp[i + 1]++; if ( i != j ) rp[j + 1]++; }
        else rp[i + 1]++;
      }
    for ( int i = 0; i < n; i++ ) rp[i + 1] += rp[i];
    int64_t* w = malloc( sizeof( int64_t ) * ( ( size_t )n + 1 ) ); memcpy( w, rp, sizeof( int64_t ) * ( ( size_t )n + 1 ) );
    S->spmv_col = malloc( sizeof( int ) * ( rp[n] > 0 ? rp[n] : 1 ) ); S->spmv_src = malloc( sizeof( int ) * ( rp[n] > 0 ? rp[n] : 1 ) );
    for ( int j = 0; j < n; j++ )
      for ( unsigned int p = cp[j]; p < cp[j + 1]; p++ ) {
        int i = ( int )ri[p];
        if ( kind != B200_KIND_LU && i > j ) continue;
        int64_t q = w[i]++; S->spmv_col[q] = j; S->spmv_src[q] = ( int )p;
        if ( kind != B200_KIND_LU && i != j ) { q = w[j]++; S->spmv_col[q] = i; S->spmv_src[q] = ( int )p; }
      }
    free( w );
    S->spmv_ptr = rp;
  }
  S->t_symbolic = now_s() - t1;
  return S;
}
