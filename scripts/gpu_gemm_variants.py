import importlib.util, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(ROOT, "meagre-crowd_b200", "__init__.py"), submodule_search_locations=[os.path.join(ROOT, "meagre-crowd_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["meagre_crowd_b200"] = m; spec.loader.exec_module(m)
s = m.Solver()
for shape in [(8192, 8192, 1024), (8192, 8192, 256), (2048, 2048, 256), (8192, 8192, 64)]:
    out = []
    for var in (7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18):
        out.append("%d:%.1f" % (var, s.L.b200_bench_dgemm_variant(s.ctx, var, *shape, 3)))
    print(shape, " ".join(out), flush=True)
s.close()
