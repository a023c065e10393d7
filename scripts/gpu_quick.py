"""quick GPU sanity run: factor+solve of small stencil systems, backward error, timings"""
import importlib.util, sys, time, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(ROOT, "meagre-crowd_b200", "__init__.py"), submodule_search_locations=[os.path.join(ROOT, "meagre-crowd_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["meagre_crowd_b200"] = m; spec.loader.exec_module(m)
from meagre_crowd_b200 import stencils

def berr(A, x, b):
    import scipy.sparse.linalg as sla
    r = A @ x - b
    return np.linalg.norm(r, np.inf) / (sla.norm(A, np.inf) * np.linalg.norm(x, np.inf) + np.linalg.norm(b, np.inf))

def run(name, gen, kind, sym, p=1, profile=False):
    n, cp, ri, v = gen
    A = stencils.to_scipy(n, cp, ri, v, symmetric_upper=sym)
    s = m.Solver()
    t = time.time(); S = s.analyze(n, cp, ri, kind); ta = time.time() - t
    if profile: s.L.b200_set_profile(s.ctx, 1)
    t = time.time(); r = s.factor(v); tf = time.time() - t
    if r != 0:
        print(name, "FACTOR FAILED", r, s.err()); return
    st = s.stats()
    b = np.arange(n * p, dtype=np.float64).reshape(n, p, order="F") % 1024 if p > 1 else np.arange(n, dtype=np.float64)
    t = time.time(); x = s.solve(b); ts = time.time() - t
    st2 = s.stats()
    be = max(berr(A, x[:, k] if p > 1 else x, b[:, k] if p > 1 else b) for k in range(p))
    fl = S.flops * (2 if kind else 1)
    print(f"{name}: n={n} nnzL={S.nnzL:.3g} flops={fl:.3g} ns={S.ns} analyze={ta:.2f}s factor_wall={tf*1e3:.1f}ms dev={st.factor_ms:.1f}ms ({fl/st.factor_ms/1e9:.2f} TF/s) launches={st.launches} solve_wall={ts*1e3:.1f}ms dev={st2.solve_ms:.2f}ms launches={st2.solve_launches} berr={be:.2e}", flush=True)
    if profile:
        x = s.solve(b)
        buf = (b"\0" * 4096); import ctypes as C
        cb = C.create_string_buffer(4096); s.L.b200_get_profile(s.ctx, cb, 4096); print(cb.value.decode())
    s.close()

if __name__ == "__main__":
    L = m.lib()
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run("p2d-8", stencils.poisson2d_upper(8), 0, True)
        run("p2d-40", stencils.poisson2d_upper(40), 0, True)
        run("p3d-12", stencils.poisson3d_upper(12), 0, True)
        run("p3d-24 p=3", stencils.poisson3d_upper(24), 0, True, p=3)
        run("p3d-32", stencils.poisson3d_upper(32), 0, True, profile=True)
        run("lu p3d-12 (full)", stencils.upper_to_full(*stencils.poisson3d_upper(12)), 1, False)
        run("lu cd27-10", stencils.convdiff27_full(10), 1, False)
        run("lu cd27-20", stencils.convdiff27_full(20), 1, False, p=2)
    elif which == "mid":
        run("p3d-64", stencils.poisson3d_upper(64), 0, True, profile=True)
        run("p2d-1024", stencils.poisson2d_upper(1024), 0, True, profile=True)
    elif which == "configs":   # BASELINE.json configs[1], [3] at full size, plus the multi-RHS sweep of configs[4]
        run("C2 p2d-1024", stencils.poisson2d_upper(1024), 0, True, profile=False)
        run("C2 p2d-1024", stencils.poisson2d_upper(1024), 0, True, profile=False)
        run("C4 lu cd27-96", stencils.convdiff27_full(96), 1, False, profile=False)
    elif which == "big":
        run("p3d-128", stencils.poisson3d_upper(128), 0, True, profile=False)
        run("p3d-128", stencils.poisson3d_upper(128), 0, True, profile=True)
    elif which == "peaks":
        s = m.Solver()
        for (mm, nn, kk) in [(8192, 8192, 256), (8192, 8192, 2048), (16384, 16384, 256), (4096, 4096, 64)]:
            print("dgemm", mm, nn, kk, "TF/s", s.L.b200_bench_dgemm_tflops(s.ctx, mm, nn, kk, 5), flush=True)
        print("stream GB/s", s.L.b200_bench_stream_gbs(s.ctx, 1 << 30, 5))
        s.close()
