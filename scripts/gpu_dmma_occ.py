import importlib.util, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(ROOT, "meagre-crowd_b200", "__init__.py"), submodule_search_locations=[os.path.join(ROOT, "meagre-crowd_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["meagre_crowd_b200"] = m; spec.loader.exec_module(m)
s = m.Solver()
print("TF/s by warps/SM (rows) x independent accumulators per warp (cols 1,2,4,8,16)")
for w in (4, 8, 12, 16, 24, 32, 64):
    print(w, " ".join("%.1f" % s.L.b200_bench_dmma_occupancy(s.ctx, w, ilp) for ilp in (1, 2, 4, 8, 16)), flush=True)
s.close()
