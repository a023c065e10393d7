import importlib.util, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(ROOT, "meagre-crowd_b200", "__init__.py"), submodule_search_locations=[os.path.join(ROOT, "meagre-crowd_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["meagre_crowd_b200"] = m; spec.loader.exec_module(m)
s = m.Solver()
shapes = [(8192, 8192, 2048), (8192, 8192, 256), (16384, 16384, 256)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for (mm, nn, kk) in shapes:
    print("dgemm", mm, nn, kk, "TF/s %.2f" % s.L.b200_bench_dgemm_tflops(s.ctx, mm, nn, kk, 3), flush=True)
print("dmma roof %.2f dfma roof %.2f" % (s.L.b200_bench_fp64_peak_tflops(s.ctx, 0, 3), s.L.b200_bench_fp64_peak_tflops(s.ctx, 1, 3)))
s.close()
