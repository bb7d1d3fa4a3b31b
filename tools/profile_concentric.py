import sys, os, time, cProfile, pstats, io
sys.path.insert(0, "."); sys.path.insert(0, "bin")
import importlib.util
spec = importlib.util.spec_from_file_location("cs", "bin/concentric_sampler_b200.py"); cs = importlib.util.module_from_spec(spec); spec.loader.exec_module(cs)
from skysampler_b200 import concentric
columns = {"cols_dc": ["C0", "C1", "C2"], "cols_wr": ["LOGR"], "cols_wcr": ["C0", "C1", "C2", "LOGR"]}
shells = [cs.make_shell(s, 100000, 300) for s in range(4)]
concentric.run_concentric([cs.make_shell(9, 20000, 300)], columns, nsamples=20000)   # warm up CUDA
pr = cProfile.Profile(); pr.enable(); t0 = time.time()
res = concentric.run_concentric(shells, columns, nsamples=1000000, bandwidth=0.1, philox=True)
dt = time.time() - t0; pr.disable()
print("4 shells wall %.2fs" % dt)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3500])
