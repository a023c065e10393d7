"""LU on matrices that are not diagonally dominant: perturbed pivots, largest multiplier over the fully-summed rows, threshold violations, backward error"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
import scipy.sparse.linalg as sla

def run(name, gen):
    n, cp, ri, v = gen
    A = stencils.to_scipy(n, cp, ri, v)
    s = m.Solver(); S = s.analyze(n, cp, ri, m.KIND_LU)
    r = s.factor(v); st = s.stats()
    if r != 0:
        print(name, "factor failed", r, s.err()); s.close(); return
    b = np.arange(n, dtype=float)
    try:
        x = s.solve(b); st2 = s.stats()
        be = np.abs(A @ x - b).max() / (sla.norm(A, np.inf) * np.abs(x).max() + np.abs(b).max())
        xs = sla.splu(A.tocsc()).solve(b)
        print(f"{name}: n={n} maxk={int(np.diff(m.sym_array(s.S, 'sn_start', S.ns + 1, np.int32)).max())} perturbed={st.npiv_perturbed} max|l|={st.lu_max_multiplier:.3g} violations={st.npiv_threshold_violations} refine={st2.refine_steps} berr={be:.2e} |x-x_superlu|/|x|={np.abs(x - xs).max() / np.abs(xs).max():.2e}", flush=True)
    except Exception as e:
        print(name, "solve failed:", e)
    s.close()

for peh in (0.5, 4.0, 16.0, 64.0):
    run(f"cd7-central-24 peh={peh}", stencils.convdiff7_central(24, peh=peh))
run("cd7-central-40 peh=16", stencils.convdiff7_central(40, peh=16.0))
for nn, dens in ((2000, 0.004), (4000, 0.002)):
    run(f"random-weak-diag n={nn}", stencils.random_weak_diagonal(nn, dens))
