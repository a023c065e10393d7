"""tile-shape sweep of the bulk-copy DMMA kernel (TFLOP/s, best of 4)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package(); s = m.Solver()
names = ["64x64 w32x16 s4 (production)", "128x64 w32x32 s4", "64x128 w32x32 s4", "128x64 w64x16 s4", "64x64 w32x16 s6", "64x64 w16x32 s4", "128x64 w32x32 s3", "128x128 w64x32 s3", "64x64 w32x32 s4"]
for shape in ((8192, 8192, 512), (8192, 8192, 256), (4096, 4096, 128), (16384, 512, 512)):
    print(shape, " | ".join(f"{v}:{s.L.b200_bench_dgemm_bulk_variant(s.ctx, v, *shape, 4):.1f}" for v in range(len(names))), flush=True)
print("variants:", "; ".join(f"{i}={n}" for i, n in enumerate(names)))
s.close()
