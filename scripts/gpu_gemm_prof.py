"""the DMMA update kernel alone at a root-front shape (for ncu): C[8192x8192] -= A[8192x512] B[8192x512]'"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
s = m.Solver()
M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 8192, 512)
print("dgemm", M, N, K, "TF/s", s.L.b200_bench_dgemm_tflops(s.ctx, M, N, K, 3))
s.close()
