import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package(); s = m.Solver()
for nb, nt in ((64, 1), (64, 24), (64, 2000), (32, 1), (32, 2000), (12, 1), (12, 20000), (48, 1)):
    print(f"nb={nb:3d} tasks={nt:6d}  registers/unrolled {s.L.b200_bench_potrf_us(s.ctx, 0, nb, nt, 5):9.1f} us   smem/rolled {s.L.b200_bench_potrf_us(s.ctx, 1, nb, nt, 5):9.1f} us", flush=True)
s.close()
