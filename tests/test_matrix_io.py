"""Host plumbing either side of the path: matrix_t conversions, symmetry detection, MatrixMarket
reader/writer and the CLI's argument errors -- behaviour of the reference's src/matrix.c, src/file.c,
src/args.c as its own tests exercise it (tests/unit-matrix.c, testsuite.at:155-309). CPU only."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from helpers import read_mm_dense


def load(mc, path):
    L = mc.lib()
    m = L.malloc_matrix()
    assert L.load_matrix(path.encode(), m) == 0
    assert L.validate_matrix(m) == 0
    return m


def test_load_detects_symmetry_like_the_reference(mc, golden):
    L = mc.lib()
    m = load(mc, f"{golden}/sym_posdef.mm")        # declared general, stored with both triangles
    assert (m.contents.sym, m.contents.location, m.contents.format, m.contents.base, m.contents.nz) == (mc.SM_SYMMETRIC, mc.MC_STORE_BOTH, mc.SM_COO, mc.FIRST_INDEX_ONE, 15)
    L.free_matrix(m)
    m = load(mc, f"{golden}/unsym.mm")
    assert m.contents.sym == mc.SM_UNSYMMETRIC and m.contents.nz == 11
    L.free_matrix(m)
    m = load(mc, f"{golden}/sym.mm")               # symmetric indefinite
    assert m.contents.sym == mc.SM_SYMMETRIC
    L.free_matrix(m)


def test_dispatcher_style_conversion_to_upper_csc(mc, golden):
    """what _convert_matrix_A (src/solvers.c:35-180) hands a CSC|BASE_ZERO|SYMMETRIC_UPPER backend"""
    L = mc.lib()
    m = load(mc, f"{golden}/sym_posdef.mm")
    assert L.convert_matrix_symmetry(m, mc.UPPER_TRIANGULAR) == 0
    assert L.convert_matrix(m, mc.SM_CSC, mc.FIRST_INDEX_ZERO) == 0
    mm = m.contents
    assert (mm.format, mm.base, mm.location, mm.nz) == (mc.SM_CSC, 0, mc.UPPER_TRIANGULAR, 10)
    jj = np.ctypeslib.as_array(mm.jj, shape=(6,)); ii = np.ctypeslib.as_array(mm.ii, shape=(10,))
    dd = np.frombuffer((C.c_double * 10).from_address(mm.dd), dtype=np.float64)
    assert jj[0] == 0 and jj[5] == 10
    A = np.zeros((5, 5))
    for j in range(5):
        for p in range(jj[j], jj[j + 1]):
            assert ii[p] <= j
            A[ii[p], j] = dd[p]
    ref = read_mm_dense(f"{golden}/sym_posdef.mm")
    assert np.array_equal(A, np.triu(ref))
    L.free_matrix(m)


@pytest.mark.parametrize("fmt", ["DROW", "DCOL", "SM_COO", "SM_CSC", "SM_CSR"])
@pytest.mark.parametrize("base", [0, 1])
def test_format_base_round_trips(mc, golden, fmt, base):
    """every format x base conversion round-trips (tests/unit-matrix.c:149-210)"""
    L = mc.lib()
    a = load(mc, f"{golden}/unsym.mm")
    b = L.copy_matrix(a)
    assert L.convert_matrix(b, getattr(mc, fmt), base) == 0 and L.validate_matrix(b) == 0
    for f2 in (mc.SM_CSR, mc.DCOL, mc.SM_CSC, mc.SM_COO):
        assert L.convert_matrix(b, f2, 1 - base) == 0 and L.validate_matrix(b) == 0
    assert L.convert_matrix(b, mc.SM_COO, mc.FIRST_INDEX_ONE) == 0
    assert L.cmp_matrix(a, b) == 0
    L.free_matrix(a); L.free_matrix(b)


def test_symmetry_storage_conversions(mc, golden):
    L = mc.lib()
    a = load(mc, f"{golden}/sym.mm")
    for loc in (mc.UPPER_TRIANGULAR, mc.LOWER_TRIANGULAR, mc.MC_STORE_BOTH, mc.LOWER_TRIANGULAR, mc.UPPER_TRIANGULAR, mc.MC_STORE_BOTH):
        b = L.copy_matrix(a)
        assert L.convert_matrix_symmetry(b, loc) == 0 and b.contents.location == loc
        assert b.contents.nz == (11 if loc == mc.MC_STORE_BOTH else 8)
        assert L.cmp_matrix(a, b) == 0
        L.free_matrix(b)
    L.free_matrix(a)


def test_writer_reproduces_golden_file_bytes(mc, golden, tmp_path):
    """load an answer file, save it again: identical bytes (format '%d %d %.13e', src/file.c:697-726 + BeBOP)"""
    L = mc.lib()
    for name in ("unsym-default-ans.mm", "sym_posdef-rhs1-ans.mm", "sym-default-ans.mm"):
        m = load(mc, f"{golden}/{name}")
        m.contents.sym = mc.SM_UNSYMMETRIC      # a column vector is never symmetric
        out = tmp_path / name
        assert L.save_matrix(m, str(out).encode()) == 0
        assert out.read_bytes() == open(f"{golden}/{name}", "rb").read()
        L.free_matrix(m)


def test_reader_grammar_errors(mc, tmp_path):
    L = mc.lib()
    cases = {"broken.mm": " Jibberish file\nnot even the right format\n",
             "pattern.mm": "%%MatrixMarket matrix coordinate pattern general\n2 2 1\n1 1\n",
             "short.mm": "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n",
             "dotfive.mm": "%%MatrixMarket matrix coordinate real general\n1 1 1\n1 1 .5\n",
             "blank_in_data.mm": "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n\n2 2 1.0\n"}
    for name, text in cases.items():
        p = tmp_path / name; p.write_text(text)
        m = L.malloc_matrix()
        assert L.load_matrix(str(p).encode(), m) != 0, name
        L.free_matrix(m)
    ok = tmp_path / "ok.mm"
    ok.write_text("%%MatrixMarket MATRIX Coordinate Real Symmetric\n% comment\n\n3 3 4\n1 1 2.0\n2 2 2e0\n3 3 +2.0\n3 1 -1\n")
    m = L.malloc_matrix()
    assert L.load_matrix(str(ok).encode(), m) == 0
    assert (m.contents.sym, m.contents.location, m.contents.nz) == (mc.SM_SYMMETRIC, mc.LOWER_TRIANGULAR, 4)
    L.free_matrix(m)
    arr = tmp_path / "arr.mm"
    arr.write_text("%%MatrixMarket matrix array real general\n2 2\n1.0\n2.0\n3.0\n4.0\n")
    m = L.malloc_matrix()
    assert L.load_matrix(str(arr).encode(), m) == 0 and m.contents.format == mc.DCOL
    assert np.array_equal(mc.matrix_to_numpy(m), np.array([[1.0, 3.0], [2.0, 4.0]]))
    L.free_matrix(m)


def run_cli(mc, *args, cwd=None):
    return subprocess.run([mc.CLI_PATH, *args], capture_output=True, text=True, cwd=cwd)


def test_cli_argument_errors_match_testsuite(mc, golden, tmp_path):
    """exit codes and stderr texts of testsuite.at:155-284 that need no solver run"""
    r = run_cli(mc); assert r.returncode == 1 and "Usage" in r.stderr
    for a in ("--help", "-h", "-?"):
        assert run_cli(mc, a).returncode == 0
    r = run_cli(mc, "--version"); assert (r.returncode, r.stdout) == (0, "meagre-crowd 0.4.6\n")
    r = run_cli(mc, "-V"); assert r.stdout == "meagre-crowd 0.4.6\n"
    for a in (["--list-solvers"], ["-v", "--list-solvers"], ["-vv", "--list-solvers"], ["--list-solvers", "-vvv"]):
        r = run_cli(mc, *a); assert r.returncode == 0 and "b200" in r.stdout
    assert run_cli(mc, "--input").returncode == 64 and run_cli(mc, "-i").returncode == 64
    r = run_cli(mc, "-i", "bad"); assert (r.returncode, r.stdout, r.stderr) == (1, "", "input error: No such file or directory\n")
    r = run_cli(mc, "-i", "-"); assert (r.returncode, r.stderr) == (1, "input error: No such file or directory\n")
    (tmp_path / "broken.mm").write_text(" Jibberish file\nnot even the right format\n")
    r = run_cli(mc, "-i", "broken.mm", cwd=tmp_path); assert (r.returncode, r.stdout, r.stderr) == (1, "", "input error: Failed to load matrix\n")
    (tmp_path / "test.rb").write_text("x"); (tmp_path / "test.hb").write_text("x")
    r = run_cli(mc, "-i", "test.rb", cwd=tmp_path); assert (r.returncode, r.stderr) == (1, "input error: Sorry Rutherford-Boeing reader is broken\n")
    r = run_cli(mc, "-i", "test.hb", cwd=tmp_path); assert (r.returncode, r.stderr) == (1, "input error: Sorry Harwell-Boeing reader is broken\n")
    u = f"{golden}/unsym.mm"
    r = run_cli(mc, "-i", u, "-p", "-0.1"); assert (r.returncode, r.stderr) == (1, "precision must be non-negative\n")
    r = run_cli(mc, "-i", u, "-p", "abcd"); assert (r.returncode, r.stderr) == (1, "bad precision (non-floating point number)\n")
    r = run_cli(mc, "-i", u, "-s", "nosuchsolver"); assert (r.returncode, r.stderr) == (1, "invalid solver (-s)\n")
    assert run_cli(mc, "-i", u, "-s").returncode == 64 and run_cli(mc, "-i", u, "-o").returncode == 64 and run_cli(mc, "-i", u, "-e").returncode == 64


def test_results_match_semantics(mc):
    """abs(x-e) <= precision element-wise; -p 0 needs bit equality (src/meagre-crowd.c:51-83)"""
    L = mc.lib()
    e = mc.make_matrix_dcol(np.array([1.0, 2.0, 3.0])); x = mc.make_matrix_dcol(np.array([1.0, 2.0 + 4e-14, 3.0]))
    assert L.results_match(e, x, 5e-14) == 1 and L.results_match(e, x, 3e-14) == 0 and L.results_match(e, x, 0.0) == 0
    assert L.results_match(e, e, 0.0) == 1
    L.free_matrix(e); L.free_matrix(x)


def test_perftimer_ticks(mc):
    import time
    L = mc.lib()
    t = L.perftimer_malloc()
    L.perftimer_inc.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    assert L.perftimer_inc(None, b"x", 10) == -1        # NULL timer is a no-op (src/perftimer.c:85-86)
    for name in (b"analyze", b"factorize", b"evaluate", b"done"):
        L.perftimer_inc(t, name, 100); time.sleep(0.002)
    assert L.perftimer_tic_ms(t, b"factorize") >= 1.5 and L.perftimer_tic_ms(t, b"nope") < 0
    assert L.perftimer_wall(t) >= 0.005
    L.perftimer_free(t)
