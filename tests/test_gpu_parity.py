"""Parity tests proper: the CUDA path, called through the C-ABI (plugin callbacks and shim),
against the CPU oracle and the reference's golden answer files. Run with -m gpu on a B200.

Tolerances (north_star): x within the reference tests' absolute tolerances on the fixtures
(5e-14 / 5e-12 / 5e-11, testsuite.at:312-346); scaled backward error
||Ax-b|| / (||A|| ||x|| + ||b||) <= 1e-12 for SPD systems and for the LU cases here;
L factors within 1e-12 relative of the oracle's (FP64, different summation order).
"""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import scipy.sparse as sp
import oracle
from helpers import sym_dict, read_mm_dense, backward_error
from meagre_crowd_b200 import stencils

pytestmark = pytest.mark.gpu

GOLDEN_CASES = [("unsym", "default", 5e-14), ("unsym", "rhs1", 5e-14), ("sym", "default", 5e-14), ("sym", "rhs1", 5e-14),
                ("sym_posdef", "default", 5e-12), ("sym_posdef", "rhs1", 5e-11)]


def plugin_solve(mc, A_path, b=None, b_path=None, rep=1):
    """drive the plugin exactly like main(): load -> solver_init -> solver_solve x rep -> finalize"""
    L = mc.lib()
    A = L.malloc_matrix(); assert L.load_matrix(A_path.encode(), A) == 0
    if b_path:
        bm = L.malloc_matrix(); assert L.load_matrix(b_path.encode(), bm) == 0
    else:
        bm = mc.make_matrix_dcol(np.arange(A.contents.m, dtype=float) if b is None else b)
    x = L.malloc_matrix()
    timer = L.perftimer_malloc()
    st = L.solver_init(L.lookup_solver_by_shortname(b"b200"), 0, 0, timer)
    for _ in range(rep):
        L.solver_solve(st, A, bm, x)
    assert L.validate_matrix(x) == 0
    out = mc.matrix_to_numpy(x)
    ticks = {k: L.perftimer_tic_ms(timer, k.encode()) for k in ("analyze", "factorize", "evaluate")}
    L.solver_finalize(st)
    L.free_matrix(A); L.free_matrix(bm); L.free_matrix(x); L.perftimer_free(timer)
    return out, ticks


@pytest.mark.parametrize("mat,rhs,tol", GOLDEN_CASES)
def test_known_answer_fixtures_through_plugin(mc, golden, mat, rhs, tol):
    x, ticks = plugin_solve(mc, f"{golden}/{mat}.mm", b_path=None if rhs == "default" else f"{golden}/rhs1.mm")
    e = read_mm_dense(f"{golden}/{mat}-{rhs}-ans.mm")
    assert x.shape == e.shape and np.all(np.abs(x - e) <= tol), np.abs(x - e).max()
    assert all(v >= 0 for v in ticks.values())       # the dispatcher's ticks were recorded


@pytest.mark.parametrize("mat,rhs,tol", GOLDEN_CASES)
def test_known_answer_fixtures_through_cli(mc, golden, mat, rhs, tol):
    """testsuite.at:312-346 verbatim: --expected=... [--right-hand-side=rhs1.mm] --solver=b200 => PASS"""
    args = [mc.CLI_PATH, f"--input={golden}/{mat}.mm", f"--expected={golden}/{mat}-{rhs}-ans.mm", "--solver=b200", f"--precision={tol}"]
    if rhs == "rhs1":
        args.append(f"--right-hand-side={golden}/rhs1.mm")
    r = subprocess.run(args, capture_output=True, text=True)
    assert (r.returncode, r.stdout) == (0, "PASS\n"), (r.stdout, r.stderr)


def test_cli_output_formats(mc, golden, tmp_path):
    """testsuite.at:187-233: silent run, `-o -` print format, -b, -p 0 => FAIL(100), wrong answer => FAIL, -t CSV, -v"""
    u = f"{golden}/unsym.mm"
    r = subprocess.run([mc.CLI_PATH, "-i", u], capture_output=True, text=True); assert (r.returncode, r.stdout) == (0, "")
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-o", "-"], capture_output=True, text=True)
    assert r.stdout == "  x(0)=1.00\n  x(1)=0.43\n  x(2)=0.25\n  x(3)=-0.09\n  x(4)=0.25\n"
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-b", f"{golden}/rhs1.mm", "-o", "-"], capture_output=True, text=True)
    assert r.stdout == "  x(0)=1.00\n  x(1)=-0.04\n  x(2)=0.25\n  x(3)=-0.36\n  x(4)=0.25\n"
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-e", f"{golden}/unsym-default-ans.mm", "-p", "0"], capture_output=True, text=True)
    assert (r.returncode, r.stdout) == (100, "FAIL\n")
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-e", f"{golden}/unsym-rhs1-ans.mm"], capture_output=True, text=True)
    assert (r.returncode, r.stdout) == (100, "FAIL\n")
    out = tmp_path / "x.mm"
    r = subprocess.run([mc.CLI_PATH, "-i", f"{golden}/sym_posdef.mm", "-o", str(out)], capture_output=True, text=True); assert r.returncode == 0
    got = read_mm_dense(str(out)); ref = read_mm_dense(f"{golden}/sym_posdef-default-ans.mm")
    assert open(out).readline() == "%%MatrixMarket matrix coordinate real general\n" and np.all(np.abs(got - ref) <= 5e-12)
    r = subprocess.run([mc.CLI_PATH, "-i", str(out), "-o", str(out)], capture_output=True, text=True); assert r.returncode == 1   # refuses to overwrite
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-t"], capture_output=True, text=True)
    head, body = r.stdout.splitlines()
    assert head == "status, solver, threads, rows, max. memory (MB),  total memory (MB), analyze, factorize, evaluate, total (ms)"
    assert body.startswith("PASS, B200-multifrontal, 0, 5, ") and len(body.split(", ")) == 10
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-t", "-e", f"{golden}/unsym-default-ans.mm"], capture_output=True, text=True)
    assert r.stdout.splitlines()[1].endswith("analyze, factorize, evaluate, done, total (ms)")     # the extra "test" tic (src/meagre-crowd.c:370)
    assert subprocess.run([mc.CLI_PATH, "-i", u, "-t", "-r", "3"], capture_output=True, text=True).returncode == 0
    r = subprocess.run([mc.CLI_PATH, "-i", u, "-v"], capture_output=True, text=True)
    assert r.stdout.startswith("Ax=b: A is 5x5, nz=11, unsymmetric, real, b is 5x1, nz=5\nsolved with B200-multifrontal on 0 cores, 0 threads\n")
    for flags in ("-vv", "-vvv", "-tt", "-ttt", "-tttt"):
        assert subprocess.run([mc.CLI_PATH, "-i", u, flags], capture_output=True, text=True).returncode == 0


def test_repeated_solve_and_multiple_rhs_through_plugin(mc, golden):
    """-r N calls analyze/factorize/evaluate N times on one state; a p-column b arrives in one call"""
    A = read_mm_dense(f"{golden}/sym_posdef.mm")
    B = np.stack([np.arange(5.0), np.array([4.0, 1, 2, 3, 4]), np.ones(5)], axis=1)
    x, _ = plugin_solve(mc, f"{golden}/sym_posdef.mm", b=B, rep=3)
    assert x.shape == (5, 3) and np.allclose(A @ x, B, atol=1e-11)


CHOL_GRIDS = [("p2d-1x1", lambda: stencils.poisson2d_upper(1, 1)), ("p2d-3x2", lambda: stencils.poisson2d_upper(3, 2)),
              ("p2d-33", lambda: stencils.poisson2d_upper(33)), ("p2d-130x7", lambda: stencils.poisson2d_upper(130, 7)),
              ("p3d-11", lambda: stencils.poisson3d_upper(11)), ("p3d-20x13x9", lambda: stencils.poisson3d_upper(20, 13, 9)),
              ("p3d-26", lambda: stencils.poisson3d_upper(26))]


@pytest.mark.parametrize("name,gen", CHOL_GRIDS)
@pytest.mark.parametrize("relax", [True, False])
def test_cholesky_factor_and_solution_vs_oracle(mc, name, gen, relax):
    n, cp, ri, v = gen()
    v = v * (1.0 + 0.3 * np.abs(np.sin(np.arange(len(v)))) * (ri.astype(np.int64) == np.repeat(np.arange(n), np.diff(cp.astype(np.int64)))))   # non-constant diagonal, still diagonally dominant
    s = mc.Solver()
    s.analyze(n, cp, ri, 0, relax=relax)
    assert s.factor(v) == 0, s.err()
    d = sym_dict(mc, s.S)
    pcp, pri, pvx = oracle.permuted_upper(n, cp, ri, v, d["perm"])
    fac = oracle.chol_uplooking(n, pcp, pri, pvx)
    # (1) the factor itself, column by column, against the oracle's L
    flat = s.read_factor(0)
    worst = 0.0
    for sn in range(d["ns"]):
        c0, c1 = d["sn_start"][sn], d["sn_start"][sn + 1]
        rows = d["rowidx"][d["rowptr"][sn]:d["rowptr"][sn + 1]]; f = len(rows)
        panel = mc.panel_view(flat, d["lptr"][sn], f, c1 - c0)
        for j in range(c0, c1):
            col = np.zeros(n); col[rows[j - c0:]] = panel[j - c0:, j - c0]
            ref = np.zeros(n); ref[fac["Li"][fac["Lp"][j]:fac["Lp"][j + 1]]] = fac["Lx"][fac["Lp"][j]:fac["Lp"][j + 1]]
            worst = max(worst, np.abs(col - ref).max() / np.abs(ref).max())
    assert worst <= 1e-12, worst
    # (2) solutions for 1 and 5 right-hand sides against the oracle's solve
    B = np.stack([np.arange(n, dtype=float), np.ones(n), np.cos(np.arange(n)), (np.arange(n) % 7) - 3.0, np.zeros(n)], axis=1)
    X = s.solve(B); x1 = s.solve(B[:, 0])
    assert np.array_equal(X[:, 0], x1)                      # same bits whatever the batch
    A = stencils.to_scipy(n, cp, ri, v, True)
    for k in range(B.shape[1]):
        y = oracle.chol_solve(fac, B[d["perm"], k]); xo = np.empty(n); xo[d["perm"]] = y
        scale = max(np.abs(xo).max(), 1e-300)
        assert np.abs(X[:, k] - xo).max() <= 1e-11 * scale
        if np.any(B[:, k]):
            assert backward_error(A, X[:, k], B[:, k]) <= 1e-12
    assert np.all(X[:, 4] == 0.0)
    s.close()


def test_factor_is_deterministic_and_refactorable(mc):
    """idempotence: factoring the same values twice gives identical bits; new values reuse the analysis"""
    n, cp, ri, v = stencils.poisson3d_upper(14)
    s = mc.Solver(); s.analyze(n, cp, ri, 0)
    assert s.factor(v) == 0; a = s.read_factor(0).copy()
    assert s.factor(v) == 0; b = s.read_factor(0)
    assert np.array_equal(a, b)
    assert s.factor(2.0 * v) == 0
    x = s.solve(np.ones(n)); A = stencils.to_scipy(n, cp, ri, 2.0 * v, True)
    assert backward_error(A, x, np.ones(n)) <= 1e-12
    s.close()


def test_solve_is_linear(mc):
    n, cp, ri, v = stencils.poisson2d_upper(50)
    s = mc.Solver(); s.analyze(n, cp, ri, 0); assert s.factor(v) == 0
    rng = np.random.default_rng(1); b1 = rng.standard_normal(n); b2 = rng.standard_normal(n)
    x1, x2, x3 = s.solve(b1), s.solve(b2), s.solve(2.0 * b1 - 3.0 * b2)
    assert np.abs(x3 - (2.0 * x1 - 3.0 * x2)).max() <= 1e-10 * np.abs(x3).max()
    s.close()


@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_sweeps_over_several_256_column_blocks_and_rhs_chunks(mc, kind):
    """root supernode wider than one 256-column solve block (look-ahead finish by chunk partials, both sweeps), 19 right-hand
    sides = three chunks of eight: every column must be bit-identical to the one-column solve and agree with SuperLU"""
    import scipy.sparse.linalg as sla
    if kind == "cholesky":
        n, cp, ri, v = stencils.poisson3d_upper(30); A = stencils.to_scipy(n, cp, ri, v, True); k = mc.KIND_CHOLESKY
    else:
        n, cp, ri, v = stencils.convdiff27_full(22); A = stencils.to_scipy(n, cp, ri, v, False); k = mc.KIND_LU
    s = mc.Solver(); S = s.analyze(n, cp, ri, k)
    st = mc.sym_array(s.S, "sn_start", S.ns + 1, np.int32)
    assert np.diff(st).max() > 2 * 256, "the case must have a supernode spanning several solve blocks"
    s.L.b200_set_refinement(s.ctx, 0)
    assert s.factor(v) == 0
    rng = np.random.default_rng(7); B = rng.standard_normal((n, 19))
    X = s.solve(B)
    for c in (0, 7, 8, 18):
        assert np.array_equal(X[:, c], s.solve(B[:, c].copy()))
    lu = sla.splu(A.tocsc())
    Xr = lu.solve(B)
    assert np.abs(X - Xr).max() <= 1e-10 * np.abs(Xr).max()
    for c in range(19):
        assert backward_error(A, X[:, c], B[:, c]) <= 1e-12
    s.close()


@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_wide_solve_blocks_same_bits_for_every_batch(mc, kind, monkeypatch):
    """the big fronts use 1024-column solve blocks (B200_SOLVE_WIDE_MIN forces them on a small system): FMA kernels (p <= 4) and the
    tensor path (p >= 5) follow one canonical arithmetic, so a column has the same bits in every batch"""
    import scipy.sparse.linalg as sla
    monkeypatch.setenv("B200_SOLVE_WIDE_MIN", "512")
    if kind == "cholesky":
        n, cp, ri, v = stencils.poisson3d_upper(34, 33, 35); A = stencils.to_scipy(n, cp, ri, v, True); k = mc.KIND_CHOLESKY
    else:
        n, cp, ri, v = stencils.convdiff27_full(36, 34, 35); A = stencils.to_scipy(n, cp, ri, v, False); k = mc.KIND_LU
    s = mc.Solver(); S = s.analyze(n, cp, ri, k)
    st = mc.sym_array(s.S, "sn_start", S.ns + 1, np.int32)
    assert np.diff(st).max() > 1024 + 64, "the case must have a supernode with a full wide block and a ragged second one"
    s.L.b200_set_refinement(s.ctx, 0)
    assert s.factor(v) == 0
    rng = np.random.default_rng(11); B = rng.standard_normal((n, 70))
    X = s.solve(B)                                           # tensor path, two 64-wide slices of right-hand sides
    for c in (0, 63, 64, 69):
        assert np.array_equal(X[:, c], s.solve(B[:, c].copy()))          # FMA path, p = 1
    assert np.array_equal(X[:, 3:6], s.solve(B[:, 3:6].copy()))          # FMA path, p = 3
    assert np.array_equal(X[:, 10:21], s.solve(B[:, 10:21].copy()))      # tensor path, odd p
    Xr = sla.splu(A.tocsc()).solve(B)
    assert np.abs(X - Xr).max() <= 1e-10 * np.abs(Xr).max()
    for c in range(0, 70, 7):
        assert backward_error(A, X[:, c], B[:, c]) <= 1e-12
    s.close()


def test_not_positive_definite_is_reported_not_hidden(mc, golden):
    Ad = read_mm_dense(f"{golden}/sym.mm")
    U = sp.triu(sp.csc_matrix(Ad)).tocsc(); U.sort_indices()
    s = mc.Solver(); s.analyze(5, U.indptr, U.indices, 0)
    assert s.factor(U.data) == mc.B200_ERR_NOT_SPD and "positive definite" in s.err()
    with pytest.raises(RuntimeError):
        s.solve(np.ones(5))                       # no solve after a failed factor
    s.close()


def shifted_laplacian(nx, ny, nz, shift):
    """3-D 7-point Laplacian minus shift*I: symmetric indefinite when the shift lies inside the spectrum (0, 12)"""
    n, cp, ri, v = stencils.poisson3d_upper(nx, ny, nz)
    diag = ri.astype(np.int64) == np.repeat(np.arange(n), np.diff(cp.astype(np.int64)))
    v = v.copy(); v[diag] -= shift
    return n, cp, ri, v


@pytest.mark.parametrize("rhs,tol", [("default", 5e-14), ("rhs1", 5e-14)])
def test_ldlt_reproduces_sym_fixture(mc, golden, rhs, tol):
    """tests/sym.mm (indefinite) through kind = LDLT of the shim: testsuite.at:325-329 at the reference's tolerance"""
    Ad = read_mm_dense(f"{golden}/sym.mm")
    U = sp.triu(sp.csc_matrix(Ad)).tocsc(); U.sort_indices()
    s = mc.Solver(); s.analyze(5, U.indptr, U.indices, mc.KIND_LDLT)
    assert s.factor(U.data) == 0, s.err()
    assert s.stats().npiv_perturbed == 0
    b = np.arange(5.0) if rhs == "default" else read_mm_dense(f"{golden}/rhs1.mm")[:, 0]
    x = s.solve(b)
    assert np.all(np.abs(x - read_mm_dense(f"{golden}/sym-{rhs}-ans.mm")[:, 0]) <= tol)
    s.close()


def test_plugin_routes_indefinite_symmetric_input_through_ldlt(mc, golden, capfd):
    """the plugin tries Cholesky, sees a non-positive pivot and switches the SAME analysis to L D L' (no LU re-analysis)"""
    os.environ["B200_VERBOSE"] = "1"
    try:
        x, _ = plugin_solve(mc, f"{golden}/sym.mm", rep=2)
    finally:
        del os.environ["B200_VERBOSE"]
    err = capfd.readouterr().err
    assert "switching to LDL'" in err and "kind=LDLT" in err and "kind=LU" not in err
    assert np.all(np.abs(x - read_mm_dense(f"{golden}/sym-default-ans.mm")) <= 5e-14)


@pytest.mark.parametrize("dims,shift", [((12, 12, 12), 3.3), ((30, 29, 31), 5.1), ((20, 7, 5), 11.0)])
def test_ldlt_indefinite_grids_vs_oracles(mc, dims, shift):
    """shifted Laplacians (eigenvalues of both signs): dense Bunch-Kaufman oracle where it finishes, SuperLU otherwise; several
    64-blocks per supernode, 1 / 3 / 9 right-hand sides (FMA and tensor sweeps) with identical bits per column"""
    import scipy.sparse.linalg as sla
    n, cp, ri, v = shifted_laplacian(*dims, shift)
    A = stencils.to_scipy(n, cp, ri, v, True)
    s = mc.Solver(); s.analyze(n, cp, ri, mc.KIND_LDLT)
    assert s.factor(v) == 0, s.err()
    rng = np.random.default_rng(5); B = rng.standard_normal((n, 9))
    X = s.solve(B)
    assert np.array_equal(X[:, 4], s.solve(B[:, 4].copy())) and np.array_equal(X[:, :3], s.solve(B[:, :3].copy()))
    for c in range(9):
        assert backward_error(A, X[:, c], B[:, c]) <= 1e-12
    Xr = oracle.ldlt_solve(A.toarray(), B) if n <= 2000 else sla.splu(A.tocsc()).solve(B)
    assert np.abs(X - Xr).max() <= 1e-8 * np.abs(Xr).max()
    s.close()


def test_ldlt_indefinite_at_64cubed(mc):
    """a 3-D indefinite system at a size no scalar oracle reaches: backward error <= 1e-12 (the north_star bound), no perturbed pivot,
    the factorization repeatable bit for bit, inertia-free checksum  sum_j ||L(:,j)||^2 finite"""
    n, cp, ri, v = shifted_laplacian(64, 64, 64, 2.7)
    A = stencils.to_scipy(n, cp, ri, v, True)
    s = mc.Solver(); s.analyze(n, cp, ri, mc.KIND_LDLT)
    assert s.factor(v) == 0, s.err()
    st = s.stats()
    b = np.arange(n, dtype=float)
    x = s.solve(b)
    assert backward_error(A, x, b) <= 1e-12
    assert s.stats().npiv_perturbed == 0
    cs = s.colnorms(); assert np.all(np.isfinite(cs))
    assert s.factor(v) == 0 and np.array_equal(cs, s.colnorms())
    print(f"LDLT 64^3: factor {st.factor_ms:.1f} ms, refine_steps {s.stats().refine_steps}")
    s.close()


def test_ldlt_on_a_positive_definite_matrix_equals_cholesky(mc):
    """on an SPD system no interchange happens: L D L' and L L' give the same solution to rounding"""
    n, cp, ri, v = stencils.poisson3d_upper(16)
    b = np.arange(n, dtype=float)
    xs = []
    for kind in (mc.KIND_CHOLESKY, mc.KIND_LDLT):
        s = mc.Solver(); s.analyze(n, cp, ri, kind); assert s.factor(v) == 0; xs.append(s.solve(b)); s.close()
    assert np.abs(xs[0] - xs[1]).max() <= 1e-12 * np.abs(xs[0]).max()


def test_ldlt_saddle_point_zero_diagonal_block(mc):
    """KKT matrix [[K, B'], [B, 0]]: zero diagonal entries need 2x2 pivots (or a perturbation that refinement repairs)"""
    rng = np.random.default_rng(2)
    n1, cp, ri, v = stencils.poisson2d_upper(20)
    K = stencils.to_scipy(n1, cp, ri, v, True)
    Bm = sp.random(60, n1, density=0.02, random_state=4, format="csr") + sp.eye(60, n1, format="csr")
    A = sp.bmat([[K, Bm.T], [Bm, None]], format="csc")
    n = A.shape[0]
    U = sp.triu(A, 0).tocsc(); U.sort_indices()
    s = mc.Solver(); s.analyze(n, U.indptr, U.indices, mc.KIND_LDLT)
    assert s.factor(U.data) == 0, s.err()
    b = rng.standard_normal(n)
    x = s.solve(b)
    assert backward_error(A, x, b) <= 1e-12
    xo = oracle.ldlt_solve(A.toarray(), b)
    assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()
    s.close()


LU_CASES = [("cd27-6", lambda: stencils.convdiff27_full(6)), ("cd27-12x9x7", lambda: stencils.convdiff27_full(12, 9, 7)),
            ("cd27-16", lambda: stencils.convdiff27_full(16)), ("p3d-full-10", lambda: stencils.upper_to_full(*stencils.poisson3d_upper(10)))]


@pytest.mark.parametrize("name,gen", LU_CASES)
def test_lu_vs_oracle_and_backward_error(mc, name, gen):
    n, cp, ri, v = gen()
    s = mc.Solver(); s.analyze(n, cp, ri, 1)
    assert s.factor(v) == 0, s.err()
    assert s.stats().npiv_perturbed == 0
    A = stencils.to_scipy(n, cp, ri, v)
    B = np.stack([np.arange(n, dtype=float), np.sin(np.arange(n))], axis=1)
    X = s.solve(B)
    for k in range(2):
        assert backward_error(A, X[:, k], B[:, k]) <= 1e-12
    if n <= 1200:
        Xo = oracle.dense_solve(A.toarray(), B)     # dense LU with partial pivoting (the UMFPACK/MUMPS class of answer)
        assert np.abs(X - Xo).max() <= 1e-10 * np.abs(Xo).max()
    import scipy.sparse.linalg as sla
    Xs = sla.splu(A.tocsc()).solve(B)               # independent third party (SuperLU)
    assert np.abs(X - Xs).max() <= 1e-9 * np.abs(Xs).max()
    s.close()


@pytest.mark.parametrize("dims,peh", [((24, 24, 24), 4.0), ((24, 24, 24), 64.0), ((40, 38, 36), 16.0)])
def test_lu_threshold_pivoting_holds_on_matrices_that_are_not_diagonally_dominant(mc, dims, peh):
    """central-difference convection at cell Peclet numbers >> 2: off-diagonal row sums several times the diagonal, strongly unsymmetric.
    The factorization must satisfy threshold partial pivoting (u = 0.1, UMFPACK's default, src/solver_umfpack.c:98) over ALL fully-summed
    rows of every front -- verified on the device from the finished multipliers -- with no perturbed pivot, and agree with SuperLU."""
    import scipy.sparse.linalg as sla
    n, cp, ri, v = stencils.convdiff7_central(*dims, peh=peh)
    A = stencils.to_scipy(n, cp, ri, v)
    off = np.asarray(np.abs(A - sp.diags(A.diagonal())).sum(axis=1)).ravel()
    assert np.median(off) > 1.2 * 6.0                                                   # interior rows are not diagonally dominant
    s = mc.Solver(); S = s.analyze(n, cp, ri, mc.KIND_LU)
    assert np.diff(mc.sym_array(s.S, "sn_start", S.ns + 1, np.int32)).max() > 256       # fronts with several diagonal blocks
    assert s.factor(v) == 0, s.err()
    st = s.stats()
    assert st.npiv_perturbed == 0 and st.npiv_threshold_violations == 0 and 1.0 <= st.lu_max_multiplier <= 1.0 / st.pivot_threshold
    B = np.stack([np.arange(n, dtype=float), np.cos(np.arange(n))], axis=1)
    X = s.solve(B)
    for c in range(2):
        assert backward_error(A, X[:, c], B[:, c]) <= 1e-12
    Xs = sla.splu(A.tocsc()).solve(B)
    assert np.abs(X - Xs).max() <= 1e-10 * np.abs(Xs).max()
    s.close()


def test_lu_weak_diagonal_is_flagged_never_silently_wrong(mc, tmp_path):
    """random sparse matrix with a weak diagonal: pivots from outside the 64-row diagonal blocks (or delayed pivots) would be needed.
    The backend must SAY so: multipliers beyond 1/u and perturbed pivots are counted, the shim reports the true backward error of the
    refined solution, and the plugin / CLI abort like the reference's assert( status == UMFPACK_OK ) instead of printing garbage."""
    n, cp, ri, v = stencils.random_weak_diagonal(1500, 0.004)
    A = stencils.to_scipy(n, cp, ri, v)
    s = mc.Solver(); s.analyze(n, cp, ri, mc.KIND_LU)
    assert s.factor(v) == 0
    st = s.stats()
    b = np.arange(n, dtype=float)
    try:
        x = s.solve(b)
        be = backward_error(A, x, b)
        st2 = s.stats()
        assert be <= 1e-10 or ((st.npiv_threshold_violations > 0 or st.npiv_perturbed > 0) and st2.backward_error > 1e-12)
    except RuntimeError as e:                                   # non-finite solution: B200_ERR_SINGULAR
        assert "not finite" in str(e)
    s.close()
    path = tmp_path / "weak.mm"
    stencils.write_mm(str(path), n, cp, ri, v)
    r = subprocess.run([mc.CLI_PATH, "-i", str(path), "-o", "-"], capture_output=True, text=True)
    if st.npiv_perturbed > 0:
        assert r.returncode != 0 and ("numerically singular" in r.stderr or "not finite" in r.stderr), (r.returncode, r.stderr[-300:])


def test_nan_in_the_matrix_fails_the_factorization(mc):
    n, cp, ri, v = stencils.convdiff27_full(6)
    v = v.copy(); v[len(v) // 2] = np.nan
    s = mc.Solver(); s.analyze(n, cp, ri, mc.KIND_LU)
    assert s.factor(v) == mc.B200_ERR_SINGULAR and "NaN" in s.err()
    s.close()


def test_lu_needs_pivoting_inside_the_front(mc, golden):
    """tests/unsym.mm has zero diagonal entries: only row interchanges make it factorable"""
    Ad = read_mm_dense(f"{golden}/unsym.mm")
    assert Ad[1, 1] == 0 and Ad[2, 2] == 0
    A = sp.csc_matrix(Ad); A.sort_indices()
    s = mc.Solver(); s.analyze(5, A.indptr, A.indices, 1)
    assert s.factor(A.data) == 0 and s.stats().npiv_perturbed == 0
    x = s.solve(np.arange(5.0))
    assert np.all(np.abs(x - read_mm_dense(f"{golden}/unsym-default-ans.mm")[:, 0]) <= 5e-14)
    s.close()


def test_headline_shape_properties_at_64cubed(mc):
    """size-independent properties at a size the scalar oracle cannot reach in seconds:
    backward error <= 1e-12, residual orthogonality free; agreement with the BLAS-3 CPU multifrontal oracle"""
    n, cp, ri, v = stencils.poisson3d_upper(64)
    s = mc.Solver(); S = s.analyze(n, cp, ri, 0)
    assert s.factor(v) == 0, s.err()
    b = np.arange(n, dtype=float)
    x = s.solve(b)
    A = stencils.to_scipy(n, cp, ri, v, True)
    assert backward_error(A, x, b) <= 1e-12
    d = sym_dict(mc, s.S)
    oracle.load_blas()
    panels = oracle.mf_cholesky(n, cp, ri, v, d)
    xo = oracle.mf_solve(d, panels, b)
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()
    # checksum of the factor: sum of squares of row i of L equals a_ii => trace identity
    flat = s.read_factor(0)
    tr = 0.0
    for sn in range(d["ns"]):
        k = d["sn_start"][sn + 1] - d["sn_start"][sn]; f = d["rowptr"][sn + 1] - d["rowptr"][sn]
        p = mc.panel_view(flat, d["lptr"][sn], f, k)
        tr += float((np.tril(p[:k]) ** 2).sum() + (p[k:] ** 2).sum())
    assert abs(tr - 6.0 * n) <= 1e-9 * 6.0 * n          # ||L||_F^2 = trace(A)
    # the factor itself, panel by panel, against the CPU multifrontal oracle (OpenBLAS potrf / trsm / syrk: another summation order)
    worst = 0.0
    for sn in range(d["ns"]):
        k = d["sn_start"][sn + 1] - d["sn_start"][sn]; f = d["rowptr"][sn + 1] - d["rowptr"][sn]
        g = mc.panel_view(flat, d["lptr"][sn], f, k); o = mc.panel_view(panels, d["lptr"][sn], f, k)
        g = np.vstack([np.tril(g[:k]), g[k:]]); o = np.vstack([np.tril(o[:k]), o[k:]])
        worst = max(worst, float(np.abs(g - o).max() / max(np.abs(o).max(), 1e-300)))
    assert worst <= 1e-11, worst
    s.close()


@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_persistent_sweeps_same_bits_as_level_set_and_tensor_sweeps(mc, kind, monkeypatch):
    """the sweeps exist three times: level-set launches of FMA kernels (p = 2..4), ONE persistent kernel per sweep whose units wait on
    device-side counters (p = 1 in production; csrc/cuda/b200_sweep.cuh) and the tensor path (p >= 5).  All three follow one canonical
    arithmetic: a column has the same bits whichever of them solves it.  B200_SWEEP_PMAX / B200_SOLVE_LEVELSET select the path."""
    if kind == "cholesky":
        n, cp, ri, v = stencils.poisson3d_upper(30); k = mc.KIND_CHOLESKY
    else:
        n, cp, ri, v = stencils.convdiff27_full(16); k = mc.KIND_LU
    rng = np.random.default_rng(11); B = rng.standard_normal((n, 8))
    out = {}
    for name, env in (("persistent", {"B200_SWEEP_PMAX": "4"}), ("levelset", {"B200_SOLVE_LEVELSET": "1"})):
        for key in ("B200_SWEEP_PMAX", "B200_SOLVE_LEVELSET"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        s = mc.Solver(); s.analyze(n, cp, ri, k)
        s.L.b200_set_refinement(s.ctx, 0)
        assert s.factor(v) == 0
        out[name] = [s.solve(np.asfortranarray(B[:, :p]) if p > 1 else B[:, 0].copy()).reshape(n, -1, order="F") for p in (1, 2, 3, 4)]
        if name == "persistent":
            out["tensor"] = s.solve(np.asfortranarray(B)).reshape(n, -1, order="F")
        s.close()
    for i, p in enumerate((1, 2, 3, 4)):
        assert np.array_equal(out["persistent"][i], out["levelset"][i]), f"p={p}: persistent and level-set sweeps differ"
        assert np.array_equal(out["persistent"][i], out["tensor"][:, :p]), f"p={p}: persistent sweeps and the tensor path differ"
