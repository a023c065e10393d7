"""persistent sweeps (p <= 4) against the level-set FMA path (B200_SOLVE_LEVELSET=1) and the tensor path (p = 8): bit identity on small
grids, then timings at a large one: python scripts/gpu_sweep_check.py [big_grid]"""
import sys, os, time, numpy as np
os.environ["B200_SWEEP_PMAX"] = "4"      # the persistent path for every p <= 4 (production uses it for p = 1)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils


def run(gen, kind, p, levelset, refine=0, reps=1):
    if levelset: os.environ["B200_SOLVE_LEVELSET"] = "1"
    else: os.environ.pop("B200_SOLVE_LEVELSET", None)
    n, cp, ri, v = gen()
    s = m.Solver(); S = s.analyze(n, cp, ri, kind)
    assert s.factor(v) == 0, s.err()
    s.L.b200_set_refinement(s.ctx, refine)
    b = (np.arange(n * p, dtype=np.float64).reshape(n, p, order="F") % 1021) - 500.0
    best = 1e30
    for _ in range(reps):
        x = s.solve(b if p > 1 else b[:, 0]); best = min(best, s.stats().solve_ms)
    s.close()
    return np.asarray(x).reshape(n, -1, order="F"), best


ok = True
for name, gen, kind in (("p3d-12", lambda: stencils.poisson3d_upper(12), 0), ("p3d-24", lambda: stencils.poisson3d_upper(24), 0), ("p2d-300", lambda: stencils.poisson2d_upper(300), 0),
                        ("p3d-40", lambda: stencils.poisson3d_upper(40), 0), ("cd27-14-lu", lambda: stencils.convdiff27_full(14), 1)):
    x8, _ = run(gen, kind, 8, False)
    for p in (1, 2, 3, 4):
        xs, _ = run(gen, kind, p, False); xl, _ = run(gen, kind, p, True)
        same_l = np.array_equal(xs, xl); same_t = np.array_equal(xs, x8[:, :p])
        print(f"{name} p={p}: sweep == levelset {same_l}, sweep == tensor path columns {same_t}, max|diff| {np.abs(xs - xl).max():.2e}", flush=True)
        ok = ok and same_l and same_t
print("BIT-IDENTICAL" if ok else "MISMATCH", flush=True)
big = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for p in (1, 2, 4):
    for ls in (False, True):
        _, ms = run(lambda: stencils.poisson3d_upper(big), 0, p, ls, reps=5)
        print(f"grid {big} p={p} {'levelset' if ls else 'sweep   '} solve_ms {ms:.3f}", flush=True)
