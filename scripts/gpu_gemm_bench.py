"""production DMMA update kernel, TFLOP/s at a few shapes"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package(); s = m.Solver()
print(" | ".join(f"{sh}: {s.L.b200_bench_dgemm_tflops(s.ctx, *sh, 4):.1f}" for sh in ((8192, 8192, 512), (8192, 8192, 256), (4096, 4096, 128), (16384, 512, 512), (20000, 64, 448), (8192, 8192, 64))), flush=True)
s.close()
