"""Host-side symbolic analysis vs the oracle: bit-exact etree, column counts, nnz(L) and
supernode partition GIVEN THE SAME ORDERING (north_star), plus structural invariants of the
front row lists, relative indices and the assembly map. CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp
import oracle
from helpers import sym_dict
from meagre_crowd_b200 import stencils

GRIDS = [("p2d-7x5", lambda: stencils.poisson2d_upper(7, 5)), ("p2d-24", lambda: stencils.poisson2d_upper(24)),
         ("p3d-6", lambda: stencils.poisson3d_upper(6)), ("p3d-10", lambda: stencils.poisson3d_upper(10)),
         ("p3d-9x7x5", lambda: stencils.poisson3d_upper(9, 7, 5))]


def dense_colcounts(Adense, perm):
    L = np.linalg.cholesky(Adense[np.ix_(perm, perm)])
    return (np.abs(L) > 0).sum(axis=0)


@pytest.mark.parametrize("name,gen", GRIDS)
@pytest.mark.parametrize("ordering", ["nd", "natural"])
def test_etree_colcounts_nnz_bit_exact(mc, name, gen, ordering):
    n, cp, ri, v = gen()
    S = mc.analyze(n, cp, ri, ordering=mc.ORDER_ND if ordering == "nd" else mc.ORDER_NATURAL)
    d = sym_dict(mc, S)
    perm = d["perm"]
    assert sorted(perm.tolist()) == list(range(n)) and np.array_equal(d["iperm"][perm], np.arange(n))
    pcp, pri, pvx = oracle.permuted_upper(n, cp, ri, v, perm)
    fac = oracle.chol_uplooking(n, pcp, pri, pvx, numeric=False)
    assert np.array_equal(d["parent"], fac["parent"])          # etree, bit exact
    assert np.array_equal(d["colcount"], fac["colcount"])      # column counts, bit exact
    assert d["nnzL"] == fac["nnz"] == int(d["colcount"].sum())
    assert d["flops"] == float((d["colcount"].astype(np.float64) ** 2).sum())
    # independent pin of the oracle itself: dense Cholesky of the permuted matrix -- the column counts, and the elimination tree
    # read off the factor by its definition  parent[j] = min { i > j : l_ij != 0 }  (no shared code with b200_etree / oracle_etree)
    A = stencils.to_scipy(n, cp, ri, v, True).toarray()
    assert np.array_equal(d["colcount"], dense_colcounts(A, perm))
    Ld = np.abs(np.linalg.cholesky(A[np.ix_(perm, perm)])) > 0
    dense_parent = np.array([next((i for i in range(j + 1, n) if Ld[i, j]), -1) for j in range(n)], dtype=np.int32)
    assert np.array_equal(d["parent"], dense_parent)
    # postordered: every parent has a larger index than its children
    assert np.all((d["parent"] > np.arange(n)) | (d["parent"] == -1))
    mc.lib().b200_symbolic_free(S)


@pytest.mark.parametrize("name,gen", GRIDS)
def test_fundamental_supernodes_match_definition(mc, name, gen):
    n, cp, ri, v = gen()
    S = mc.analyze(n, cp, ri, relax=False)
    d = sym_dict(mc, S)
    parent, cc = d["parent"], d["colcount"]
    nchild = np.bincount(parent[parent >= 0], minlength=n)
    starts = [0] + [j for j in range(1, n) if not (parent[j - 1] == j and cc[j - 1] == cc[j] + 1 and nchild[j] == 1)]
    assert d["ns"] == d["nfund"] == len(starts)
    assert np.array_equal(d["sn_start"], np.array(starts + [n], dtype=np.int32))
    # without relaxation the stored structure is the exact structure of L
    fsum = sum(int(d["rowptr"][s + 1] - d["rowptr"][s]) * int(d["sn_start"][s + 1] - d["sn_start"][s]) -
               (int(d["sn_start"][s + 1] - d["sn_start"][s]) * (int(d["sn_start"][s + 1] - d["sn_start"][s]) - 1)) // 2 for s in range(d["ns"]))
    assert fsum == d["nnzL"]
    mc.lib().b200_symbolic_free(S)


@pytest.mark.parametrize("name,gen", GRIDS)
@pytest.mark.parametrize("kind", [0, 1])
def test_front_structure_and_assembly_map(mc, name, gen, kind):
    n, cp, ri, v = gen()
    if kind == 1:
        n, cp, ri, v = stencils.upper_to_full(n, cp, ri, v)
        v = v + 0.01 * np.arange(len(v)) / len(v)     # unsymmetric values on a symmetric pattern
    S = mc.analyze(n, cp, ri, kind=kind)
    d = sym_dict(mc, S)
    ns = d["ns"]
    A = sp.csc_matrix((v, ri.astype(np.int64), cp.astype(np.int64)), shape=(n, n)).toarray()
    if kind == 0:
        A = A + np.triu(A, 1).T
    PA = A[np.ix_(d["perm"], d["perm"])]
    # structure of the exact factor of the permuted pattern (symmetric pattern => same for LU)
    Lpat = np.abs(np.linalg.cholesky(np.abs(PA != 0) * -1.0 + np.diag(np.full(n, 2.0 * n)))) > 0
    flatL = np.zeros(int(d["lptr"][-1])); flatU = np.zeros(int(d["lptr"][-1]))
    cols = np.repeat(np.arange(n), np.diff(cp.astype(np.int64)))
    for e in range(len(v)):
        t = d["amap"][e]
        if t < 0:
            assert kind == 0 and ri[e] > cols[e]
            continue
        if t & (1 << 62):
            flatU[t & ~(1 << 62)] += v[e]
        else:
            flatL[t] += v[e]
    seen = np.zeros(n, dtype=bool)
    for s in range(ns):
        c0, c1 = d["sn_start"][s], d["sn_start"][s + 1]
        rows = d["rowidx"][d["rowptr"][s]:d["rowptr"][s + 1]]
        k, f = c1 - c0, len(rows)
        assert np.array_equal(rows[:k], np.arange(c0, c1)) and np.all(np.diff(rows) > 0)
        seen[c0:c1] = True
        # every structural nonzero of L in these columns is inside the front (relaxation may add rows, never drop)
        need = np.nonzero(Lpat[:, c0:c1].any(axis=1))[0]
        assert set(need[need >= c0]).issubset(set(rows.tolist()))
        p = d["sn_parent"][s]
        if p >= 0:
            prow = d["rowidx"][d["rowptr"][p]:d["rowptr"][p + 1]]
            rel = d["relidx"][d["rowptr"][s] + k:d["rowptr"][s + 1]]
            assert np.array_equal(prow[rel], rows[k:])              # relative indices hit the same global rows
            assert d["sn_depth"][s] == d["sn_depth"][p] + 1
            assert p > s
        else:
            assert f == k and d["sn_depth"][s] == 0
        # the assembly map reproduces P A P' inside the panel
        panel = mc.panel_view(flatL, d["lptr"][s], f, k)
        ref = PA[np.ix_(rows, np.arange(c0, c1))].copy()
        if kind == 0:
            ref[:k, :] = np.tril(ref[:k, :])
            assert np.array_equal(panel, ref)
        else:
            refl = ref.copy(); refl[:k, :] = np.tril(refl[:k, :])
            assert np.array_equal(panel, refl)
            upanel = mc.panel_view(flatU, d["lptr"][s], f, k)
            refu = PA[np.ix_(np.arange(c0, c1), rows)].T.copy(); refu[:k, :] = np.tril(refu[:k, :], -1)
            assert np.array_equal(upanel, refu)
    assert seen.all()
    mc.lib().b200_symbolic_free(S)


def test_nested_dissection_reduces_fill(mc):
    n, cp, ri, v = stencils.poisson3d_upper(16)
    S0 = mc.analyze(n, cp, ri, ordering=mc.ORDER_NATURAL); S1 = mc.analyze(n, cp, ri)
    assert S1.contents.nnzL < 0.5 * S0.contents.nnzL and S1.contents.flops < 0.35 * S0.contents.flops
    # geometric ND on 16^3 (recursive plane bisection) needs about 3.9e5 entries; the graph ND must be in that class
    assert S1.contents.nnzL < 4.5e5
    mc.lib().b200_symbolic_free(S0); mc.lib().b200_symbolic_free(S1)


def test_given_ordering_and_unsorted_duplicate_input(mc):
    n, cp, ri, v = stencils.poisson2d_upper(9)
    rng = np.random.default_rng(0)
    perm = rng.permutation(n).astype(np.int32)
    S = mc.analyze(n, cp, ri, perm=perm)
    d = sym_dict(mc, S)
    # the analysis may re-postorder, but the result must be an equivalent elimination order: same nnz(L)
    pcp, pri, pvx = oracle.permuted_upper(n, cp, ri, v, perm)
    assert d["nnzL"] == oracle.chol_uplooking(n, pcp, pri, pvx, numeric=False)["nnz"]
    mc.lib().b200_symbolic_free(S)
    # rows inside a column in arbitrary order (BeBOP gives no guarantee; CHOLMOD is told sorted=0)
    cp64 = cp.astype(np.int64); ri2 = ri.copy()
    for j in range(n):
        seg = ri2[cp64[j]:cp64[j + 1]]; rng.shuffle(seg)
    S2 = mc.analyze(n, cp, ri2, perm=perm)
    assert S2.contents.nnzL == d["nnzL"]
    mc.lib().b200_symbolic_free(S2)


def test_order_nd_standalone_is_a_permutation_and_deterministic(mc):
    import ctypes as C
    n, cp, ri, v = stencils.poisson3d_upper(12)
    A = stencils.to_scipy(n, cp, ri, v, True); A.setdiag(0); A.eliminate_zeros(); A = A.tocsr()
    xadj = A.indptr.astype(np.int32); adj = A.indices.astype(np.int32)
    out = []
    for threads in (1, 4):
        p = np.empty(n, np.int32)
        r = mc.lib().b200_order_nd(n, xadj.ctypes.data_as(C.POINTER(C.c_int)), adj.ctypes.data_as(C.POINTER(C.c_int)), 64, threads, p.ctypes.data_as(C.POINTER(C.c_int)))
        assert r == 0 and sorted(p.tolist()) == list(range(n))
        out.append(p)
    assert np.array_equal(out[0], out[1])      # same ordering regardless of the thread count


@pytest.mark.parametrize("name,gen", [("convdiff27_9", lambda: stencils.convdiff27_full(9)), ("convdiff27_12x7x5", lambda: stencils.convdiff27_full(12, 7, 5))])
def test_lu_fill_equals_superlu_fill_on_the_same_ordering(mc, name, gen):
    """nnz(L) and nnz(U) for a given ordering against an INDEPENDENT sparse LU (north_star: "bit-exact ... nnz(L)/nnz(U) given the same
    ordering"): SuperLU (scipy) factors P A P' with its own column ordering switched off, no supernode relaxation and diagonal pivots
    (the matrices are diagonally dominant, so no row interchange changes the structure).  For a structurally symmetric matrix its L and U
    then hold exactly the fill of the symbolic Cholesky of the pattern -- the column counts of this repo's analysis."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    n, cp, ri, v = gen()
    S = mc.analyze(n, cp, ri, mc.KIND_LU, relax=False)
    d = sym_dict(mc, S)
    A = stencils.to_scipy(n, cp, ri, v, False).tocsc()
    P = d["perm"]
    Ap = A[P][:, P].tocsc()
    lu = sla.splu(Ap, permc_spec="NATURAL", diag_pivot_thresh=0.0, relax=1, panel_size=1, options=dict(SymmetricMode=False))
    assert np.array_equal(lu.perm_r, np.arange(n)) and np.array_equal(lu.perm_c, np.arange(n)), "SuperLU pivoted: the comparison needs diagonal pivots"
    Ls = lu.L.tocsc(); Us = lu.U.tocsc()
    Ls.eliminate_zeros(); Us.eliminate_zeros()
    cc = d["colcount"].astype(np.int64)                     # entries of column j of L, diagonal included
    assert int(d["nnzL"]) == int(cc.sum())
    # structural fill: SuperLU stores every structural nonzero it computed; numerical cancellation to exact zero does not occur on these matrices
    assert np.array_equal(np.diff(Ls.indptr), cc), "column counts of L differ from SuperLU's"
    assert np.array_equal(np.diff(Us.tocsr().indptr), cc), "row counts of U differ from SuperLU's (U' has the structure of L)"
    assert Ls.nnz == d["nnzL"] and Us.nnz == d["nnzL"]
    mc.lib().b200_symbolic_free(S)
