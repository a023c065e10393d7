"""Synthetic stencil systems of SURVEY.md 8(d) / BASELINE.json configs (no RNG, deterministic).

C2: 2-D 5-point Laplacian (diag 4, off -1), node id y*nx+x
C3: 3-D 7-point Laplacian (diag 6, off -1), node id (z*ny+y)*nx+x
C4: 3-D 27-point convection-diffusion: centre 26, 26 neighbours -1, plus first-order
    upwind convection with velocity (1, 1/2, 1/4)*Pe_h on the 6 face neighbours / diagonal
All return CSC base 0 as (n, colptr[uint32], rowidx[uint32], vals[float64]); symmetric
systems return the UPPER triangle (row <= col), which is what the dispatcher hands a
SOLVES_SYMMETRIC_UPPER_TRIANGULAR backend (reference src/solvers.c:42-92).
"""
import numpy as np


def _coo_to_csc(n, rows, cols, vals):
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr)
    return n, colptr.astype(np.uint32), rows.astype(np.uint32), vals.astype(np.float64)


def _grid_ids(shape):
    return np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)


def poisson2d_upper(nx, ny=None):
    ny = ny or nx
    ids = _grid_ids((ny, nx))
    n = ids.size
    r = [ids.ravel()]; c = [ids.ravel()]; v = [np.full(n, 4.0)]
    for a, b in ((ids[:, :-1], ids[:, 1:]), (ids[:-1, :], ids[1:, :])):
        r.append(a.ravel()); c.append(b.ravel()); v.append(np.full(a.size, -1.0))   # a < b: upper
    return _coo_to_csc(n, np.concatenate(r), np.concatenate(c), np.concatenate(v))


def poisson3d_upper(nx, ny=None, nz=None):
    ny = ny or nx; nz = nz or nx
    ids = _grid_ids((nz, ny, nx))
    n = ids.size
    r = [ids.ravel()]; c = [ids.ravel()]; v = [np.full(n, 6.0)]
    for a, b in ((ids[:, :, :-1], ids[:, :, 1:]), (ids[:, :-1, :], ids[:, 1:, :]), (ids[:-1, :, :], ids[1:, :, :])):
        r.append(a.ravel()); c.append(b.ravel()); v.append(np.full(a.size, -1.0))
    return _coo_to_csc(n, np.concatenate(r), np.concatenate(c), np.concatenate(v))


def convdiff27_full(nx, ny=None, nz=None, peh=0.5):
    ny = ny or nx; nz = nz or nx
    ids = _grid_ids((nz, ny, nx))
    n = ids.size
    vel = {(0, 0, 1): 1.0 * peh, (0, 1, 0): 0.5 * peh, (1, 0, 0): 0.25 * peh}   # (dz,dy,dx) -> velocity component
    diag = np.full(n, 26.0 + peh * (1.0 + 0.5 + 0.25))
    r = [ids.ravel()]; c = [ids.ravel()]; v = [diag]

    def sl(d, n_):
        return (slice(0, n_ - 1), slice(1, n_)) if d == 1 else (slice(1, n_), slice(0, n_ - 1)) if d == -1 else (slice(0, n_), slice(0, n_))

    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dz == dy == dx == 0:
                    continue
                sz, sy, sx = sl(dz, nz), sl(dy, ny), sl(dx, nx)
                me = ids[sz[0], sy[0], sx[0]].ravel()       # row i
                nb = ids[sz[1], sy[1], sx[1]].ravel()       # column j = neighbour in direction (dz,dy,dx)
                val = -1.0
                # upwind: flow in +direction, so the upstream neighbour (-direction) gets -velocity
                key = (-dz, -dy, -dx)
                if key in vel and abs(dz) + abs(dy) + abs(dx) == 1:
                    val -= vel[key]
                r.append(me); c.append(nb); v.append(np.full(me.size, val))
    return _coo_to_csc(n, np.concatenate(r), np.concatenate(c), np.concatenate(v))


def convdiff7_central(nx, ny=None, nz=None, peh=8.0):
    """3-D 7-point convection-diffusion with CENTRAL differences for the convection term: diag 6, neighbour in +direction
    -1 + v*peh/2, in -direction -1 - v*peh/2 (velocity v = 1, 1/2, 1/4).  For peh > 2 the matrix is far from diagonally dominant
    (off-diagonal row sum 6 + 1.75*peh against a diagonal of 6) and strongly unsymmetric: the LU pivoting test matrix."""
    ny = ny or nx; nz = nz or nx
    ids = _grid_ids((nz, ny, nx))
    n = ids.size
    r = [ids.ravel()]; c = [ids.ravel()]; v = [np.full(n, 6.0)]
    for axis, vel in ((2, 1.0), (1, 0.5), (0, 0.25)):
        lo = [slice(None)] * 3; hi = [slice(None)] * 3
        lo[axis] = slice(0, -1); hi[axis] = slice(1, None)
        a = ids[tuple(lo)].ravel(); b = ids[tuple(hi)].ravel()          # b is the +direction neighbour of a
        r.append(a); c.append(b); v.append(np.full(a.size, -1.0 + 0.5 * vel * peh))
        r.append(b); c.append(a); v.append(np.full(a.size, -1.0 - 0.5 * vel * peh))
    return _coo_to_csc(n, np.concatenate(r), np.concatenate(c), np.concatenate(v))


def random_weak_diagonal(n, density=0.01, diag=0.05, seed=0):
    """random sparse unsymmetric matrix (values N(0,1)) whose diagonal is diag * N(0,1): nothing favours the diagonal as a pivot"""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    A = (A + sp.diags(diag * rng.standard_normal(n))).tocsc()
    A.sum_duplicates(); A.sort_indices()
    return n, A.indptr.astype(np.uint32), A.indices.astype(np.uint32), A.data.astype(np.float64)


def upper_to_full(n, colptr, rowidx, vals):
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr.astype(np.int64)))
    rows = rowidx.astype(np.int64)
    off = rows != cols
    return _coo_to_csc(n, np.concatenate([rows, cols[off]]), np.concatenate([cols, rows[off]]), np.concatenate([vals, vals[off]]))


def to_scipy(n, colptr, rowidx, vals, symmetric_upper=False):
    import scipy.sparse as sp
    A = sp.csc_matrix((vals, rowidx.astype(np.int64), colptr.astype(np.int64)), shape=(n, n))
    if symmetric_upper:
        A = A + sp.triu(A, 1).T
    return A.tocsc()


def write_mm(path, n, colptr, rowidx, vals, symmetric_upper=False):
    """MatrixMarket writer honouring the reference reader's grammar:
    symmetric files store the LOWER triangle, no blank/comment lines inside the data."""
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr.astype(np.int64)))
    rows = rowidx.astype(np.int64)
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real %s\n" % ("symmetric" if symmetric_upper else "general"))
        f.write("%d %d %d\n" % (n, n, len(vals)))
        if symmetric_upper:
            rows, cols = cols, rows     # (i<=j) upper -> (j,i) lower
        for i, j, x in zip(rows, cols, vals):
            f.write("%d %d %.17g\n" % (i + 1, j + 1, x) if x != int(x) else "%d %d %d.0\n" % (i + 1, j + 1, int(x)))
