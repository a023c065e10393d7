"""ctypes binding of oracle/libmc_oracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package (see the header of mc_oracle.c).
"""
import ctypes as C
import glob
import os
import subprocess
import numpy as np

DIR = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(DIR, "libmc_oracle.so")
_lib = None


def build():
    r = subprocess.run(["make", "-C", DIR], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        ip, dp, lp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.oracle_dense_solve.argtypes = [C.c_int, dp, dp, C.c_int, dp]
        L.oracle_ldlt_solve.argtypes = [C.c_int, dp, dp, C.c_int, dp]
        L.oracle_etree.argtypes = [C.c_int, ip, ip, ip]
        L.oracle_chol_uplooking.restype = C.c_int64
        L.oracle_chol_uplooking.argtypes = [C.c_int, ip, ip, dp, ip, ip, lp, ip, dp]
        L.oracle_chol_solve.argtypes = [C.c_int, lp, ip, dp, dp]
        L.oracle_load_blas.argtypes = [C.c_char_p]
        L.oracle_mf_cholesky.argtypes = [C.c_int, ip, ip, dp, ip, C.c_int, ip, ip, ip, C.c_int, lp, ip, lp, dp, C.c_int]
        L.oracle_mf_solve.argtypes = [C.c_int, ip, lp, ip, lp, dp, dp]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def dense_solve(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b2 = np.asfortranarray(np.asarray(b, dtype=np.float64).reshape(A.shape[0], -1))
    x = np.empty_like(b2, order="F")
    r = lib().oracle_dense_solve(A.shape[0], _p(A, C.c_double), _p(b2, C.c_double), b2.shape[1], _p(x, C.c_double))
    if r != 0:
        raise ZeroDivisionError("singular matrix")
    return x.reshape(np.shape(b)) if np.ndim(b) == 1 else x


def ldlt_solve(A, b):
    """dense symmetric indefinite solve by Bunch-Kaufman L D L' (A full symmetric)"""
    A = np.asfortranarray(A, dtype=np.float64)
    b2 = np.asfortranarray(np.asarray(b, dtype=np.float64).reshape(A.shape[0], -1))
    x = np.empty_like(b2, order="F")
    r = lib().oracle_ldlt_solve(A.shape[0], _p(A, C.c_double), _p(b2, C.c_double), b2.shape[1], _p(x, C.c_double))
    if r != 0:
        raise ZeroDivisionError(f"column {r - 1} is exactly zero")
    return x.reshape(np.shape(b)) if np.ndim(b) == 1 else x


def permuted_upper(n, colptr, rowidx, vals, perm):
    """upper triangle (CSC, int32) of P A P' for a symmetric A given by its upper triangle"""
    import scipy.sparse as sp
    A = sp.csc_matrix((vals, rowidx.astype(np.int64), colptr.astype(np.int64)), shape=(n, n))
    A = A + sp.triu(A, 1).T
    PA = sp.triu(A.tocsr()[perm][:, perm].tocsc(), 0).tocsc()
    PA.sort_indices()
    return PA.indptr.astype(np.int32), PA.indices.astype(np.int32), PA.data.astype(np.float64)


def chol_uplooking(n, cp, ri, vx, numeric=True):
    """returns dict(parent, colcount, nnz, Lp, Li, Lx) for the already permuted upper-CSC matrix"""
    L = lib()
    cp = np.ascontiguousarray(cp, np.int32); ri = np.ascontiguousarray(ri, np.int32); vx = np.ascontiguousarray(vx, np.float64)
    parent = np.empty(n, np.int32); cc = np.empty(n, np.int32)
    nnz = L.oracle_chol_uplooking(n, _p(cp, C.c_int), _p(ri, C.c_int), _p(vx, C.c_double), _p(parent, C.c_int), _p(cc, C.c_int), None, None, None)
    out = dict(parent=parent, colcount=cc, nnz=int(nnz))
    if numeric:
        Lp = np.empty(n + 1, np.int64); Li = np.empty(nnz, np.int32); Lx = np.empty(nnz, np.float64)
        r = L.oracle_chol_uplooking(n, _p(cp, C.c_int), _p(ri, C.c_int), _p(vx, C.c_double), _p(parent, C.c_int), _p(cc, C.c_int), _p(Lp, C.c_int64), _p(Li, C.c_int), _p(Lx, C.c_double))
        if r < 0:
            raise ValueError(f"not positive definite at pivot {-r - 1}")
        out.update(Lp=Lp, Li=Li, Lx=Lx)
    return out


def chol_solve(fac, b):
    x = np.array(b, dtype=np.float64, copy=True)
    lib().oracle_chol_solve(len(x), _p(fac["Lp"], C.c_int64), _p(fac["Li"], C.c_int), _p(fac["Lx"], C.c_double), _p(x, C.c_double))
    return x


def load_blas():
    """dlopen an optimised BLAS/LAPACK for the timed CPU baseline (numpy/scipy's bundled OpenBLAS)"""
    cands = []
    try:
        import scipy, numpy
        for base in (os.path.dirname(scipy.__file__) + ".libs", os.path.dirname(numpy.__file__) + ".libs"):
            cands += sorted(glob.glob(os.path.join(base, "libscipy_openblas*.so")) + glob.glob(os.path.join(base, "libopenblas*.so")))
    except Exception:
        pass
    cands = [c for c in cands if "64_" not in os.path.basename(c)] + []   # LP64 interface only
    return bool(lib().oracle_load_blas(":".join(cands).encode())) if cands else False


def mf_cholesky(n, colptr, rowidx, vals, sym, threads=0):
    """CPU multifrontal Cholesky on the analysis `sym` (a dict of numpy arrays); returns the panels"""
    L = lib()
    cp = np.ascontiguousarray(colptr, np.int32); ri = np.ascontiguousarray(rowidx, np.int32); vx = np.ascontiguousarray(vals, np.float64)
    panels = np.zeros(int(sym["lptr"][-1]), dtype=np.float64)
    r = L.oracle_mf_cholesky(n, _p(cp, C.c_int), _p(ri, C.c_int), _p(vx, C.c_double), _p(sym["iperm"], C.c_int), sym["ns"],
                             _p(sym["sn_start"], C.c_int), _p(sym["sn_parent"], C.c_int), _p(sym["sn_depth"], C.c_int), int(sym["maxdepth"]),
                             _p(sym["rowptr"], C.c_int64), _p(sym["rowidx"], C.c_int), _p(sym["lptr"], C.c_int64), _p(panels, C.c_double), threads)
    if r != 0:
        raise ValueError(f"not positive definite near column {r - 1}")
    return panels


def mf_solve(sym, panels, b):
    """x = A^-1 b on the CPU panels (original numbering in, original numbering out)"""
    perm = sym["perm"]
    y = np.ascontiguousarray(np.asarray(b, np.float64)[perm])
    lib().oracle_mf_solve(sym["ns"], _p(sym["sn_start"], C.c_int), _p(sym["rowptr"], C.c_int64), _p(sym["rowidx"], C.c_int), _p(sym["lptr"], C.c_int64), _p(panels, C.c_double), _p(y, C.c_double))
    x = np.empty_like(y); x[perm] = y
    return x
