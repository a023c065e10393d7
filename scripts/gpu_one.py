"""one factor + solve of poisson3d grid^3 (for ncu captures): python scripts/gpu_one.py grid [p]"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n, cp, ri, v = stencils.poisson3d_upper(grid)
s = m.Solver(); S = s.analyze(n, cp, ri, 0)
assert s.factor(v) == 0
b = np.arange(n * p, dtype=np.float64).reshape(n, p, order="F") % 1024 if p > 1 else np.arange(n, dtype=np.float64)
x = s.solve(b); st = s.stats()
print("factor_ms", st.factor_ms, "solve_ms", st.solve_ms)
s.close()
