import numpy as np


def sym_dict(mc, S):
    """numpy copies of the arrays of a b200_symbolic_t"""
    s = S.contents
    ns, n = s.ns, s.n
    d = dict(ns=ns, n=n, maxdepth=s.maxdepth, nnzL=s.nnzL, flops=s.flops, nfund=s.nfund, kind=s.kind, nnzA=s.nnzA,
             stored_nnzL=s.stored_nnzL, maxfront=s.maxfront)
    for name, ln, dt in (("perm", n, np.int32), ("iperm", n, np.int32), ("parent", n, np.int32), ("colcount", n, np.int32),
                         ("sn_start", ns + 1, np.int32), ("sn_parent", ns, np.int32), ("sn_depth", ns, np.int32),
                         ("rowptr", ns + 1, np.int64), ("lptr", ns + 1, np.int64), ("cbptr", ns, np.int64)):
        d[name] = mc.sym_array(S, name, ln, dt)
    d["rowidx"] = mc.sym_array(S, "rowidx", int(d["rowptr"][-1]), np.int32)
    d["relidx"] = mc.sym_array(S, "relidx", int(d["rowptr"][-1]), np.int32)
    d["amap"] = mc.sym_array(S, "amap", int(s.nnzA), np.int64)
    return d


def read_mm_dense(path):
    rows = [l for l in open(path) if not l.startswith("%")]
    m, n, nz = map(int, rows[0].split())
    A = np.zeros((m, n))
    for l in rows[1:1 + nz]:
        i, j, v = l.split()
        A[int(i) - 1, int(j) - 1] += float(v)
    return A


def backward_error(A, x, b):
    """scaled backward error ||Ax-b|| / (||A|| ||x|| + ||b||) (north_star), infinity norms"""
    import scipy.sparse.linalg as sla
    r = A @ x - b
    return np.linalg.norm(r, np.inf) / (sla.norm(A, np.inf) * np.linalg.norm(x, np.inf) + np.linalg.norm(b, np.inf))
