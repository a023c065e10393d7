"""per-launch trace of one factor + one solve (profile mode serialises the launches): B200_TRACE_ALL=1 python scripts/gpu_trace.py grid > log 2>&1"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, cp, ri, v = stencils.poisson3d_upper(grid)
s = m.Solver(); S = s.analyze(n, cp, ri, 0)
assert s.factor(v) == 0
b = np.arange(n, dtype=np.float64)
x = s.solve(b)
s.L.b200_set_profile(s.ctx, 1)
assert s.factor(v) == 0
x = s.solve(b); st = s.stats()
print("factor_ms", st.factor_ms, "solve_ms", st.solve_ms)
s.close()
