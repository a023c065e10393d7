"""Pin the CPU oracle on every known-answer fixture the reference's tests hold for this path
(testsuite.at:312-346: unsym/sym/sym_posdef x default/rhs1, tolerances 5e-14, 5e-12, 5e-11)."""
import numpy as np
import pytest
import oracle

CASES = [("unsym", "default", 5e-14), ("unsym", "rhs1", 5e-14), ("sym", "default", 5e-14), ("sym", "rhs1", 5e-14),
         ("sym_posdef", "default", 5e-12), ("sym_posdef", "rhs1", 5e-11)]


def read_mm_dense(path):
    rows = [l for l in open(path) if not l.startswith("%")]
    m, n, nz = map(int, rows[0].split())
    A = np.zeros((m, n))
    for l in rows[1:1 + nz]:
        i, j, v = l.split()
        A[int(i) - 1, int(j) - 1] += float(v)
    return A


@pytest.mark.parametrize("mat,rhs,tol", CASES)
def test_oracle_dense_lu_reproduces_golden_answers(golden, mat, rhs, tol):
    A = read_mm_dense(f"{golden}/{mat}.mm")
    b = np.arange(5.0) if rhs == "default" else read_mm_dense(f"{golden}/rhs1.mm")[:, 0]   # b[i]=i, 0-based (src/meagre-crowd.c:233-234)
    e = read_mm_dense(f"{golden}/{mat}-{rhs}-ans.mm")[:, 0]
    x = oracle.dense_solve(A, b)
    assert np.all(np.abs(x - e) <= tol), np.abs(x - e).max()


def test_oracle_sparse_cholesky_reproduces_sym_posdef_answers(golden):
    A = read_mm_dense(f"{golden}/sym_posdef.mm")
    import scipy.sparse as sp
    U = sp.triu(sp.csc_matrix(A)).tocsc()
    fac = oracle.chol_uplooking(5, U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data)
    for rhs, tol in (("default", 5e-12), ("rhs1", 5e-11)):
        b = np.arange(5.0) if rhs == "default" else read_mm_dense(f"{golden}/rhs1.mm")[:, 0]
        e = read_mm_dense(f"{golden}/sym_posdef-{rhs}-ans.mm")[:, 0]
        assert np.all(np.abs(oracle.chol_solve(fac, b) - e) <= tol)
    Ld = np.zeros((5, 5))
    for j in range(5):
        for q in range(fac["Lp"][j], fac["Lp"][j + 1]):
            Ld[fac["Li"][q], j] = fac["Lx"][q]
    assert np.allclose(Ld, np.linalg.cholesky(A), atol=1e-13)


def test_oracle_rejects_indefinite(golden):
    A = read_mm_dense(f"{golden}/sym.mm")
    import scipy.sparse as sp
    U = sp.triu(sp.csc_matrix(A)).tocsc()
    with pytest.raises(ValueError):
        oracle.chol_uplooking(5, U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data)


def test_oracle_multifrontal_matches_uplooking(mc):
    """the BLAS-3 multifrontal baseline and the scalar up-looking oracle agree on a 3-D grid"""
    from meagre_crowd_b200 import stencils
    n, cp, ri, v = stencils.poisson3d_upper(9)
    S = mc.analyze(n, cp, ri)
    from helpers import sym_dict
    sym = sym_dict(mc, S)
    oracle.load_blas()
    panels = oracle.mf_cholesky(n, cp, ri, v, sym)
    pcp, pri, pvx = oracle.permuted_upper(n, cp, ri, v, sym["perm"])
    fac = oracle.chol_uplooking(n, pcp, pri, pvx)
    b = np.arange(n, dtype=float)
    x1 = oracle.mf_solve(sym, panels, b)
    y = oracle.chol_solve(fac, b[sym["perm"]]); x2 = np.empty(n); x2[sym["perm"]] = y
    assert np.abs(x1 - x2).max() <= 1e-12 * np.abs(x2).max()
    mc.lib().b200_symbolic_free(S)


@pytest.mark.parametrize("rhs", ["default", "rhs1"])
def test_oracle_ldlt_reproduces_sym_answers(golden, rhs):
    """tests/sym.mm is symmetric indefinite: the Bunch-Kaufman L D L' oracle must land on the reference's answers (5e-14)"""
    A = read_mm_dense(f"{golden}/sym.mm")
    assert np.array_equal(A, A.T) and np.linalg.eigvalsh(A).min() < 0 < np.linalg.eigvalsh(A).max()
    b = np.arange(5.0) if rhs == "default" else read_mm_dense(f"{golden}/rhs1.mm")[:, 0]
    e = read_mm_dense(f"{golden}/sym-{rhs}-ans.mm")[:, 0]
    assert np.all(np.abs(oracle.ldlt_solve(A, b) - e) <= 5e-14)


def test_oracle_ldlt_needs_2x2_pivots_and_agrees_with_lu():
    """zero diagonal (a saddle-point matrix): only 2x2 pivots factor it; answer against the dense LU oracle"""
    rng = np.random.default_rng(3)
    K = rng.standard_normal((12, 12)); K = K @ K.T + 12 * np.eye(12); B = rng.standard_normal((5, 12))
    A = np.block([[K, B.T], [B, np.zeros((5, 5))]])
    P = rng.permutation(17); A = A[np.ix_(P, P)]
    b = rng.standard_normal((17, 2))
    assert np.abs(oracle.ldlt_solve(A, b) - oracle.dense_solve(A, b)).max() <= 1e-12

