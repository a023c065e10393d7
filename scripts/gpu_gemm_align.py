import importlib.util, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(ROOT, "meagre-crowd_b200", "__init__.py"), submodule_search_locations=[os.path.join(ROOT, "meagre-crowd_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["meagre_crowd_b200"] = m; spec.loader.exec_module(m)
s = m.Solver()
for shape in [(8192, 8192, 512), (8191, 8191, 512), (16384, 16384, 512), (16383, 16383, 512), (8192, 8192, 64), (8191, 8191, 64)]:
    print(shape, "v13: %.1f  v7: %.1f" % (s.L.b200_bench_dgemm_variant(s.ctx, 13, *shape, 3), s.L.b200_bench_dgemm_variant(s.ctx, 7, *shape, 3)), flush=True)
s.close()
