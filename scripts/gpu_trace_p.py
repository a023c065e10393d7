"""per-launch trace of one solve with p right-hand sides: B200_TRACE_ALL=1 python scripts/gpu_trace_p.py grid p > log 2>&1"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n, cp, ri, v = stencils.poisson3d_upper(grid)
s = m.Solver(); S = s.analyze(n, cp, ri, 0)
assert s.factor(v) == 0
s.L.b200_set_refinement(s.ctx, 0)
b = np.asfortranarray(((np.arange(n)[:, None] + np.arange(p)[None, :]) % 1024).astype(np.float64))
x = s.solve(b); print("solve_ms", s.stats().solve_ms)
s.L.b200_set_profile(s.ctx, 1)
x = s.solve(b); st = s.stats()
print("profiled solve_ms", st.solve_ms)
s.close()
