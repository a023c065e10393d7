"""solve-only run for ncu: factor once (poisson3d grid^3), then `reps` solves with p right-hand sides, no refinement"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
refine = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n, cp, ri, v = stencils.poisson3d_upper(grid)
s = m.Solver(); S = s.analyze(n, cp, ri, 0)
s.L.b200_set_refinement(s.ctx, refine)
assert s.factor(v) == 0
b = (np.arange(n * p, dtype=np.float64).reshape(n, p, order="F") % 1024) if p > 1 else np.arange(n, dtype=np.float64)
for _ in range(reps):
    x = s.solve(b); st = s.stats()
    print(f"grid={grid} p={p} solve_ms={st.solve_ms:.3f} launches={st.solve_launches} refine={st.refine_steps} omega={st.backward_error:.3e} omega0={st.first_backward_error:.3e} "
          f"GB/s(min bytes)={(16.0 * S.nnzL + 32.0 * n * p) * (1 + st.refine_steps) / st.solve_ms / 1e6:.1f} stored GB/s={(16.0 * S.stored_nnzL) * (1 + st.refine_steps) / st.solve_ms / 1e6:.1f}", flush=True)
A = stencils.to_scipy(n, cp, ri, v, True)
x0 = x[:, 0] if p > 1 else x; b0 = b[:, 0] if p > 1 else b
print("berr", np.abs(A @ x0 - b0).max() / (12 * np.abs(x0).max() + np.abs(b0).max()))
s.close()
