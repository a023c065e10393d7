"""per-kind kernel times of one factorization of a BASELINE config (profile mode serialises the launches): python scripts/gpu_trace_cfg.py c2|c3|c4"""
import sys, os, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
n, cp, ri, v = {"c2": lambda: stencils.poisson2d_upper(1024), "c3": lambda: stencils.poisson3d_upper(128), "c4": lambda: stencils.convdiff27_full(96)}[cfg]()
kind = m.KIND_LU if cfg == "c4" else m.KIND_CHOLESKY
s = m.Solver(); S = s.analyze(n, cp, ri, kind)
assert s.factor(v) == 0
assert s.factor(v) == 0; print("factor_ms", s.stats().factor_ms)
s.L.b200_set_profile(s.ctx, 1); assert s.factor(v) == 0; s.L.b200_set_profile(s.ctx, 0)
buf = C.create_string_buffer(16384); s.L.b200_get_profile(s.ctx, buf, 16384)
print("\n".join(l for l in buf.value.decode().splitlines() if " launches=" in l and "launches=0 " not in l))
s.close()
