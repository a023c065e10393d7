"""per-unit time stamps of the persistent sweeps of one solve: B200_SWEEP_TRACE=1 python scripts/gpu_sweep_trace.py grid p out.npz"""
import sys, os, numpy as np, ctypes as C
os.environ["B200_SWEEP_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
grid = int(sys.argv[1]); p = int(sys.argv[2]); out = sys.argv[3]
n, cp, ri, v = stencils.poisson3d_upper(grid)
s = m.Solver(); S = s.analyze(n, cp, ri, 0)
assert s.factor(v) == 0
s.L.b200_set_refinement(s.ctx, 0)
b = (np.arange(n * p, dtype=np.float64).reshape(n, p, order="F") % 1021) - 500.0
for _ in range(3): x = s.solve(b if p > 1 else b[:, 0])
print("solve_ms", s.stats().solve_ms)
res = {}
for d, nm in ((0, "f"), (1, "b")):
    nu = s.L.b200_debug_sweep_trace(s.ctx, d, None, None, 0)
    tr = np.zeros((nu, 5), dtype=np.uint64); un = np.zeros((nu, 16), dtype=np.int32)
    s.L.b200_debug_sweep_trace(s.ctx, d, tr.ctypes.data_as(C.c_void_p), un.ctypes.data_as(C.c_void_p), nu)
    res["trace_" + nm] = tr; res["units_" + nm] = un
ns = S.ns
res["sn_depth"] = m.sym_array(s.S, "sn_depth", ns, np.int32); res["sn_start"] = m.sym_array(s.S, "sn_start", ns + 1, np.int32); res["rowptr"] = m.sym_array(s.S, "rowptr", ns + 1, np.int64)
np.savez_compressed(out, **res)
s.close()
