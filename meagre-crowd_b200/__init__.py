"""ctypes binding of libmeagre_crowd_b200.so -- the C-ABI declared in include/*.h.

This module is plumbing only: it loads the shared library, mirrors the C
structs (matrix_t of the reference's src/matrix.h:68-99, solver_state_t of
src/solvers.h:45-51, b200_symbolic_t of include/solver_b200.h) and exposes thin
helpers used by tests/ and bench.py.  There is no Python implementation of any
numeric step: if the library (or a CUDA device) is missing the calls fail loudly.

The directory name contains a hyphen (task layout), so import it with
`importlib` -- see `load_package()` in __graft_entry__.py -- or via
`tests/conftest.py`, which registers it as module `meagre_crowd_b200`.
"""
import ctypes as C
import os
import subprocess
import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmeagre_crowd_b200.so")
CLI_PATH = os.path.join(PKG_DIR, "meagre-crowd")

# ---- enums (include/mc_matrix.h)
FIRST_INDEX_ZERO, FIRST_INDEX_ONE = 0, 1
INVALID, DROW, DCOL, SM_COO, SM_CSC, SM_CSR = range(6)
SM_UNSYMMETRIC, SM_SYMMETRIC, SM_SKEW_SYMMETRIC, SM_HERMITIAN = range(4)
MC_STORE_BOTH, UPPER_TRIANGULAR, LOWER_TRIANGULAR = range(3)
REAL_DOUBLE = 0
B200_OK, B200_ERR_ARG, B200_ERR_NOMEM, B200_ERR_CUDA, B200_ERR_NOT_SPD, B200_ERR_SINGULAR, B200_ERR_STATE = 0, -1, -2, -3, -4, -5, -6
KIND_CHOLESKY, KIND_LU, KIND_LDLT = 0, 1, 2
ORDER_ND, ORDER_NATURAL, ORDER_GIVEN = 0, 1, 2


class matrix_t(C.Structure):
    _fields_ = [("m", C.c_size_t), ("n", C.c_size_t), ("nz", C.c_size_t),
                ("base", C.c_int), ("format", C.c_int), ("sym", C.c_int), ("location", C.c_int), ("data_type", C.c_int),
                ("dd", C.c_void_p), ("ii", C.POINTER(C.c_uint)), ("jj", C.POINTER(C.c_uint))]


class solver_state_t(C.Structure):
    _fields_ = [("solver", C.c_int), ("mpi_rank", C.c_int), ("verbosity", C.c_int), ("timer", C.c_void_p), ("specific", C.c_void_p)]


class b200_options_t(C.Structure):
    _fields_ = [("ordering", C.c_int), ("user_perm", C.POINTER(C.c_int)), ("nd_leaf", C.c_int), ("relax", C.c_int),
                ("nrelax", C.c_int * 3), ("zrelax", C.c_double * 3), ("threads", C.c_int)]


class b200_symbolic_t(C.Structure):
    _fields_ = [("n", C.c_int), ("kind", C.c_int), ("nnzA", C.c_int64),
                ("perm", C.POINTER(C.c_int)), ("iperm", C.POINTER(C.c_int)), ("parent", C.POINTER(C.c_int)), ("colcount", C.POINTER(C.c_int)),
                ("nnzL", C.c_int64), ("flops", C.c_double), ("nfund", C.c_int), ("ns", C.c_int),
                ("sn_start", C.POINTER(C.c_int)), ("sn_parent", C.POINTER(C.c_int)), ("sn_depth", C.POINTER(C.c_int)),
                ("rowptr", C.POINTER(C.c_int64)), ("rowidx", C.POINTER(C.c_int)), ("relidx", C.POINTER(C.c_int)),
                ("lptr", C.POINTER(C.c_int64)), ("uptr", C.POINTER(C.c_int64)), ("amap", C.POINTER(C.c_int64)),
                ("stored_nnzL", C.c_int64), ("stored_flops", C.c_double), ("maxfront", C.c_int), ("maxdepth", C.c_int),
                ("cb_arena", C.c_int64 * 2), ("cbptr", C.POINTER(C.c_int64)),
                ("spmv_ptr", C.POINTER(C.c_int64)), ("spmv_col", C.POINTER(C.c_int)), ("spmv_src", C.POINTER(C.c_int)), ("t_order", C.c_double), ("t_symbolic", C.c_double)]


class b200_stats_t(C.Structure):
    _fields_ = [("factor_ms", C.c_double), ("solve_ms", C.c_double), ("gemm_ms", C.c_double), ("gemm_flops", C.c_double), ("gemm_alg_flops", C.c_double),
                ("launches", C.c_int64), ("solve_launches", C.c_int64), ("gemm_launches", C.c_int64),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("device_bytes", C.c_int64), ("npiv_perturbed", C.c_int), ("refine_steps", C.c_int), ("backward_error", C.c_double), ("first_backward_error", C.c_double), ("comm_bytes", C.c_double),
                ("lu_max_multiplier", C.c_double), ("pivot_threshold", C.c_double), ("npiv_threshold_violations", C.c_int)]


class b200_plan_t(C.Structure):
    _fields_ = [("world", C.c_int), ("ns", C.c_int), ("nlevels", C.c_int), ("grp_first", C.POINTER(C.c_int)), ("grp_size", C.POINTER(C.c_int)),
                ("dist", C.POINTER(C.c_int)), ("work", C.POINTER(C.c_double)), ("level_width", C.POINTER(C.c_int))]


class b200_xfer_t(C.Structure):
    _fields_ = [("src", C.c_int), ("dst", C.c_int), ("sn", C.c_int), ("offset", C.c_int64), ("count", C.c_int64)]


class b200_sched_entry_t(C.Structure):
    _fields_ = [("kind", C.c_int), ("stream", C.c_int), ("wait_ev", C.c_int), ("record_ev", C.c_int), ("depth", C.c_int),
                ("first_msg", C.c_int64), ("nmsg", C.c_int64), ("count", C.c_int64), ("flops", C.c_double), ("kavg", C.c_double), ("bytes", C.c_double)]


class b200_access_t(C.Structure):
    _fields_ = [("launch", C.c_int), ("which", C.c_int), ("write", C.c_int), ("offset", C.c_int64), ("count", C.c_int64)]


class b200_msg_t(C.Structure):
    _fields_ = [("src", C.c_int), ("dst", C.c_int), ("which", C.c_int), ("offset", C.c_int64), ("count", C.c_int64)]


# every symbol include/*.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    # mc_matrix.h
    "malloc_matrix", "free_matrix", "clear_matrix", "copy_matrix", "cmp_matrix", "convert_matrix", "convert_matrix_symmetry",
    "detect_matrix_symmetry", "validate_matrix", "printf_matrix", "load_matrix", "save_matrix", "matrix_rows",
    # mc_solver.h
    "perftimer_malloc", "perftimer_free", "perftimer_inc", "perftimer_restart", "perftimer_adjust_depth", "perftimer_wall",
    "perftimer_wall_av", "perftimer_delta", "perftimer_rounds", "perftimer_printf", "perftimer_printf_csv_header",
    "perftimer_printf_csv_body", "perftimer_tic_ms", "solver_lookup", "lookup_solver_by_shortname", "solver2str", "printf_solvers",
    "solver_can_do", "solver_uses_mpi", "solver_requires_mpi", "solver_uses_omp", "solver_requires_omp", "select_solver", "solver",
    "solver_solve", "solver_init", "solver_finalize", "solver_analyze", "solver_factorize", "solver_evaluate", "results_match",
    # solver_b200.h
    "solver_init_b200", "solver_analyze_b200", "solver_factorize_b200", "solver_evaluate_b200", "solver_finalize_b200",
    "b200_default_options", "b200_analyze", "b200_symbolic_free", "b200_order_nd", "b200_etree",
    "b200_ctx_create", "b200_ctx_destroy", "b200_last_error", "b200_upload_symbolic", "b200_factor_host", "b200_factor_device",
    "b200_solve_host", "b200_solve_device", "b200_dev_alloc", "b200_dev_free", "b200_dev_upload", "b200_dev_download",
    "b200_host_alloc_pinned", "b200_host_free_pinned", "b200_dev_sync", "b200_flush_l2", "b200_get_stats", "b200_set_profile",
    "b200_get_profile", "b200_debug_read", "b200_set_refinement", "b200_bench_dgemm_tflops", "b200_bench_dmma_occupancy", "b200_bench_fp64_peak_tflops", "b200_bench_stream_gbs",
    "b200_outer_width", "b200_set_outer_width_cap", "b200_plan_create", "b200_plan_free", "b200_plan_col_owner", "b200_plan_member", "b200_plan_cb_transfers",
    "b200_mg_unique_id", "b200_mg_init", "b200_mg_rank", "b200_bench_potrf_us", "b200_bench_dgemm_bulk_variant", "b200_ctx_create_dry", "b200_debug_schedule", "b200_debug_messages", "b200_debug_kind_name", "b200_debug_accesses", "b200_factor_colnorms", "b200_debug_sweep_units", "b200_debug_sweep_tasks", "b200_debug_sweep_counters", "b200_debug_sweep_trace",
]

_lib = None


def build(verbose=False):
    """make -C meagre-crowd_b200 (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)"""
    r = subprocess.run(["make", "-C", PKG_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libmeagre_crowd_b200.so failed")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no Python/CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.malloc_matrix.restype = C.POINTER(matrix_t)
    L.free_matrix.argtypes = [C.POINTER(matrix_t)]
    L.clear_matrix.argtypes = [C.POINTER(matrix_t)]
    L.copy_matrix.restype = C.POINTER(matrix_t); L.copy_matrix.argtypes = [C.POINTER(matrix_t)]
    L.cmp_matrix.argtypes = [C.POINTER(matrix_t)] * 2
    L.convert_matrix.argtypes = [C.POINTER(matrix_t), i32, i32]
    L.convert_matrix_symmetry.argtypes = [C.POINTER(matrix_t), i32]
    L.detect_matrix_symmetry.argtypes = [C.POINTER(matrix_t)]
    L.validate_matrix.argtypes = [C.POINTER(matrix_t)]
    L.load_matrix.argtypes = [C.c_char_p, C.POINTER(matrix_t)]
    L.save_matrix.argtypes = [C.POINTER(matrix_t), C.c_char_p]
    L.results_match.argtypes = [C.POINTER(matrix_t), C.POINTER(matrix_t), dbl]
    L.perftimer_malloc.restype = vp
    L.perftimer_free.argtypes = [vp]
    L.perftimer_tic_ms.restype = dbl; L.perftimer_tic_ms.argtypes = [vp, C.c_char_p]
    L.perftimer_wall.restype = dbl; L.perftimer_wall.argtypes = [vp]
    L.lookup_solver_by_shortname.argtypes = [C.c_char_p]
    L.solver2str.restype = C.c_char_p; L.solver2str.argtypes = [i32]
    L.solver_init.restype = C.POINTER(solver_state_t); L.solver_init.argtypes = [i32, i32, i32, vp]
    L.solver_finalize.argtypes = [C.POINTER(solver_state_t)]
    L.solver_analyze.argtypes = [C.POINTER(solver_state_t), C.POINTER(matrix_t)]
    L.solver_factorize.argtypes = [C.POINTER(solver_state_t), C.POINTER(matrix_t)]
    L.solver_evaluate.argtypes = [C.POINTER(solver_state_t), C.POINTER(matrix_t), C.POINTER(matrix_t)]
    L.solver_solve.argtypes = [C.POINTER(solver_state_t)] + [C.POINTER(matrix_t)] * 3
    L.solver.argtypes = [i32, i32, i32] + [C.POINTER(matrix_t)] * 3
    L.b200_default_options.argtypes = [C.POINTER(b200_options_t)]
    L.b200_analyze.restype = C.POINTER(b200_symbolic_t)
    L.b200_analyze.argtypes = [i32, C.POINTER(C.c_uint), C.POINTER(C.c_uint), i32, C.POINTER(b200_options_t)]
    L.b200_symbolic_free.argtypes = [C.POINTER(b200_symbolic_t)]
    L.b200_order_nd.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), i32, i32, C.POINTER(i32)]
    L.b200_etree.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.b200_plan_create.restype = C.POINTER(b200_plan_t); L.b200_plan_create.argtypes = [C.POINTER(b200_symbolic_t), i32, i32]
    L.b200_plan_free.argtypes = [C.POINTER(b200_plan_t)]
    L.b200_plan_col_owner.argtypes = [C.POINTER(b200_plan_t), C.POINTER(b200_symbolic_t), i32, i32]
    L.b200_plan_member.argtypes = [C.POINTER(b200_plan_t), i32, i32]
    L.b200_plan_cb_transfers.restype = i64; L.b200_plan_cb_transfers.argtypes = [C.POINTER(b200_plan_t), C.POINTER(b200_symbolic_t), i32, C.POINTER(b200_xfer_t), i64]
    L.b200_mg_unique_id.argtypes = [C.c_char_p]
    L.b200_mg_init.argtypes = [vp, i32, i32, C.c_char_p]
    L.b200_mg_rank.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.b200_ctx_create_dry.restype = vp; L.b200_ctx_create_dry.argtypes = [i32, i32]
    L.b200_debug_schedule.restype = i64; L.b200_debug_schedule.argtypes = [vp, C.POINTER(b200_sched_entry_t), i64]
    L.b200_debug_messages.restype = i64; L.b200_debug_messages.argtypes = [vp, C.POINTER(b200_msg_t), i64]
    L.b200_debug_accesses.restype = i64; L.b200_debug_accesses.argtypes = [vp, C.POINTER(b200_access_t), i64]
    L.b200_debug_sweep_units.restype = i64; L.b200_debug_sweep_units.argtypes = [vp, C.c_int, vp, i64]
    L.b200_debug_sweep_tasks.restype = i64; L.b200_debug_sweep_tasks.argtypes = [vp, vp, i64]
    L.b200_debug_sweep_counters.restype = i64; L.b200_debug_sweep_counters.argtypes = [vp]
    L.b200_debug_sweep_trace.restype = i64; L.b200_debug_sweep_trace.argtypes = [vp, C.c_int, vp, vp, i64]
    L.b200_debug_kind_name.restype = C.c_char_p; L.b200_debug_kind_name.argtypes = [i32]
    L.b200_ctx_create.restype = vp; L.b200_ctx_create.argtypes = [i32]
    L.b200_ctx_destroy.argtypes = [vp]
    L.b200_last_error.restype = C.c_char_p
    L.b200_upload_symbolic.argtypes = [vp, C.POINTER(b200_symbolic_t)]
    L.b200_factor_host.argtypes = [vp, vp]; L.b200_factor_device.argtypes = [vp, vp]
    L.b200_solve_host.argtypes = [vp, vp, i32, vp]; L.b200_solve_device.argtypes = [vp, vp, i32, vp]
    L.b200_dev_alloc.restype = vp; L.b200_dev_alloc.argtypes = [vp, C.c_size_t]
    L.b200_dev_free.argtypes = [vp, vp]
    L.b200_dev_upload.argtypes = [vp, vp, vp, C.c_size_t]; L.b200_dev_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.b200_host_alloc_pinned.restype = vp; L.b200_host_alloc_pinned.argtypes = [C.c_size_t]
    L.b200_host_free_pinned.argtypes = [vp]
    L.b200_dev_sync.argtypes = [vp]; L.b200_flush_l2.argtypes = [vp]
    L.b200_get_stats.argtypes = [vp, C.POINTER(b200_stats_t)]
    L.b200_set_profile.argtypes = [vp, i32]
    L.b200_get_profile.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.b200_set_refinement.argtypes = [vp, i32]
    L.b200_debug_read.argtypes = [vp, i32, vp, i64, i64]
    L.b200_factor_colnorms.argtypes = [vp, vp]
    L.b200_bench_potrf_us.restype = dbl; L.b200_bench_potrf_us.argtypes = [vp, i32, i32, i32, i32]
    L.b200_bench_dgemm_bulk_variant.restype = dbl; L.b200_bench_dgemm_bulk_variant.argtypes = [vp, i32, i32, i32, i32, i32]
    L.b200_bench_dgemm_tflops.restype = dbl; L.b200_bench_dgemm_tflops.argtypes = [vp, i32, i32, i32, i32]
    L.b200_bench_dmma_occupancy.restype = dbl; L.b200_bench_dmma_occupancy.argtypes = [vp, i32, i32]
    L.b200_bench_fp64_peak_tflops.restype = dbl; L.b200_bench_fp64_peak_tflops.argtypes = [vp, i32, i32]
    L.b200_bench_stream_gbs.restype = dbl; L.b200_bench_stream_gbs.argtypes = [vp, C.c_size_t, i32]
    _lib = L
    return L


# ------------------------------------------------------------------ helpers
_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]


def _malloc_copy(arr):
    """copy a numpy array into a libc malloc block (matrix_t owns its arrays and free()s them)"""
    arr = np.ascontiguousarray(arr)
    p = _libc.malloc(max(arr.nbytes, 8))
    C.memmove(p, arr.ctypes.data, arr.nbytes)
    return p


def make_matrix_csc(n, colptr, rowidx, vals, sym=SM_UNSYMMETRIC, location=MC_STORE_BOTH):
    """matrix_t* in CSC base 0 owning malloc'd copies of the arrays (free with lib().free_matrix)"""
    L = lib()
    m = L.malloc_matrix()
    mm = m.contents
    mm.m = n; mm.n = n; mm.nz = len(vals)
    mm.base = FIRST_INDEX_ZERO; mm.format = SM_CSC; mm.sym = sym; mm.location = location; mm.data_type = REAL_DOUBLE
    mm.dd = _malloc_copy(np.asarray(vals, dtype=np.float64))
    mm.ii = C.cast(_malloc_copy(np.asarray(rowidx, dtype=np.uint32)), C.POINTER(C.c_uint))
    mm.jj = C.cast(_malloc_copy(np.asarray(colptr, dtype=np.uint32)), C.POINTER(C.c_uint))
    return m


def make_matrix_dcol(arr):
    """matrix_t* DCOL from a 1-D or 2-D (m x p) numpy array"""
    L = lib()
    a = np.asarray(arr, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    m = L.malloc_matrix()
    mm = m.contents
    mm.m, mm.n = a.shape; mm.nz = a.size
    mm.base = FIRST_INDEX_ZERO; mm.format = DCOL; mm.sym = SM_UNSYMMETRIC; mm.location = MC_STORE_BOTH; mm.data_type = REAL_DOUBLE
    mm.dd = _malloc_copy(np.asfortranarray(a).ravel(order="F"))
    return m


def matrix_to_numpy(m):
    """dense numpy copy (m x n) of a DCOL/DROW matrix_t"""
    mm = m.contents if hasattr(m, "contents") else m
    assert mm.format in (DCOL, DROW)
    buf = (C.c_double * (mm.m * mm.n)).from_address(mm.dd)
    a = np.frombuffer(buf, dtype=np.float64).copy()
    return a.reshape((mm.m, mm.n), order="F" if mm.format == DCOL else "C")


def analyze(n, colptr, rowidx, kind=KIND_CHOLESKY, ordering=ORDER_ND, perm=None, relax=True, nd_leaf=None, threads=0):
    """b200_analyze on numpy CSC arrays; returns POINTER(b200_symbolic_t) (free with lib().b200_symbolic_free)"""
    L = lib()
    o = b200_options_t()
    L.b200_default_options(C.byref(o))
    o.ordering = ordering
    o.relax = 1 if relax else 0
    o.threads = threads
    if nd_leaf:
        o.nd_leaf = nd_leaf
    keep = None
    if perm is not None:
        keep = np.ascontiguousarray(perm, dtype=np.int32)
        o.ordering = ORDER_GIVEN
        o.user_perm = keep.ctypes.data_as(C.POINTER(C.c_int))
    cp = np.ascontiguousarray(colptr, dtype=np.uint32)
    ri = np.ascontiguousarray(rowidx, dtype=np.uint32)
    S = L.b200_analyze(n, cp.ctypes.data_as(C.POINTER(C.c_uint)), ri.ctypes.data_as(C.POINTER(C.c_uint)), kind, C.byref(o))
    if not S:
        raise RuntimeError("b200_analyze failed")
    return S


def shared_ordering(n, colptr, rowidx, kind, rank, world, dist, threads=0):
    """one process per GPU: rank 0 computes the fill-reducing ordering with all its host threads and broadcasts it
    (torch.distributed), so that N ranks do not run N copies of the most expensive host step on N-times fewer cores each.
    Every rank (rank 0 included) then runs the rest of the analysis on the GIVEN permutation -- the same code path
    everywhere, hence bit-identical symbolic structures.  Returns the permutation (numpy int32)."""
    if world == 1:
        S = analyze(n, colptr, rowidx, kind, threads=threads)
        perm = sym_array(S, "perm", n, np.int32).copy()
        lib().b200_symbolic_free(S)
        return perm
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if rank == 0:
        S = analyze(n, colptr, rowidx, kind, threads=threads)
        t = torch.from_numpy(sym_array(S, "perm", n, np.int32).copy()).to(dev)
        lib().b200_symbolic_free(S)
    else:
        t = torch.empty(n, dtype=torch.int32, device=dev)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def panel_view(flat, offset, f, k):
    """f x k view of a factor panel: column-major, leading dimension f rounded up to even"""
    ld = (int(f) + 1) & ~1
    return flat[int(offset):int(offset) + ld * int(k)].reshape((ld, int(k)), order="F")[:int(f)]


def sym_array(S, name, length, dtype):
    p = getattr(S.contents, name)
    ct = {np.int32: C.c_int, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(length,)).copy()


class Solver:
    """analyze/factor/solve through the shim (b200_* entry points), numpy in / numpy out"""

    def __init__(self, device=0):
        self.L = lib()
        self.ctx = self.L.b200_ctx_create(device)
        if not self.ctx:
            raise RuntimeError("b200_ctx_create failed: " + self.L.b200_last_error().decode())
        self.S = None

    def err(self):
        return self.L.b200_last_error().decode()

    def join(self, rank, world, id256=None, dist=None):
        """multi-process mode: one rank per GPU.  Either pass the 256 bytes of b200_mg_unique_id, or a torch.distributed
        module with an initialised process group (rank 0 creates the id and broadcasts it)."""
        if world > 1 and id256 is None:
            import torch
            buf = C.create_string_buffer(256)
            if rank == 0 and self.L.b200_mg_unique_id(buf) != B200_OK:
                raise RuntimeError("b200_mg_unique_id: " + self.err())
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=0)
            id256 = bytes(t.cpu().tolist())
        r = self.L.b200_mg_init(self.ctx, rank, world, id256)
        if r != B200_OK:
            raise RuntimeError("b200_mg_init: " + self.err())

    def analyze(self, n, colptr, rowidx, kind=KIND_CHOLESKY, **kw):
        """kw: ordering / perm / relax / nd_leaf / threads of analyze(); perm = a given elimination order (e.g. shared_ordering)"""
        if self.S:
            self.L.b200_symbolic_free(self.S)
        self.S = analyze(n, colptr, rowidx, kind, **kw)
        r = self.L.b200_upload_symbolic(self.ctx, self.S)
        if r != B200_OK:
            raise RuntimeError("b200_upload_symbolic: " + self.err())
        self.n = n
        return self.S.contents

    def factor(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        return self.L.b200_factor_host(self.ctx, v.ctypes.data)

    def solve(self, b):
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        bb = np.asfortranarray(b.reshape(self.n, -1))
        x = np.empty_like(bb, order="F")
        r = self.L.b200_solve_host(self.ctx, bb.ctypes.data, bb.shape[1], x.ctypes.data)
        if r != B200_OK:
            raise RuntimeError("b200_solve_host: " + self.err())
        return x[:, 0].copy() if one else x

    def read_factor(self, which=0):
        """download the whole L (0) or U' (1) panel storage"""
        n = self.S.contents.stored_nnzL
        out = np.empty(n, dtype=np.float64)
        r = self.L.b200_debug_read(self.ctx, which, out.ctypes.data, 0, n)
        if r != B200_OK:
            raise RuntimeError("b200_debug_read: " + self.err())
        return out

    def dense_L(self, which=0):
        """assemble the dense n x n lower factor (of the permuted matrix) from the panels; small n only"""
        S = self.S.contents
        n, ns = S.n, S.ns
        st = sym_array(self.S, "sn_start", ns + 1, np.int32); rp = sym_array(self.S, "rowptr", ns + 1, np.int64)
        ri = sym_array(self.S, "rowidx", int(rp[ns]), np.int32); lp = sym_array(self.S, "lptr", ns + 1, np.int64)
        flat = self.read_factor(which)
        Ld = np.zeros((n, n))
        for s in range(ns):
            k = st[s + 1] - st[s]; f = rp[s + 1] - rp[s]
            rows = ri[rp[s]:rp[s + 1]]
            panel = panel_view(flat, lp[s], f, k)
            Ld[np.ix_(rows, np.arange(st[s], st[s + 1]))] = panel
        return np.tril(Ld)

    def colnorms(self):
        """per-column sums of squares of L (permuted numbering); deterministic checksum of the factor"""
        out = np.empty(self.n, dtype=np.float64)
        r = self.L.b200_factor_colnorms(self.ctx, out.ctypes.data)
        if r != B200_OK:
            raise RuntimeError("b200_factor_colnorms: " + self.err())
        return out

    def stats(self):
        st = b200_stats_t()
        self.L.b200_get_stats(self.ctx, C.byref(st))
        return st

    def close(self):
        if self.S:
            self.L.b200_symbolic_free(self.S); self.S = None
        if self.ctx:
            self.L.b200_ctx_destroy(self.ctx); self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
