"""Python-side overhead of Context.basefc (ctypes marshalling, result views): cProfile over repeated calls."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcltk_b200 import engine, workload
ctx = engine.get_context(0)
w = workload.make_basefc_workload(ctx, int(float(sys.argv[1])) if len(sys.argv) > 1 else 30000000, 10000, 60000, seed=7)
seg_mode = sys.argv[2] if len(sys.argv) > 2 else "tiny"
for _ in range(3):
    ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, 10000, w.params, segments=seg_mode)
pr = cProfile.Profile()
t = time.perf_counter()
pr.enable()
for _ in range(20):
    seg = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, 10000, w.params, segments=seg_mode)
    tm = ctx.timing()
pr.disable()
print("wall per call %.3f ms, in-library %.3f ms" % (1e3 * (time.perf_counter() - t) / 20, tm[12]))
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

# where the wall time of one call goes: the raw C call against the library's own clock and the Python around it
raw = ctx.lib.xg_basefc
acc = {"raw": 0.0}
class Timed(object):
    def __call__(self, *a):
        t0 = time.perf_counter()
        rc = raw(*a)
        acc["raw"] += time.perf_counter() - t0
        return rc
ctx.lib.xg_basefc = Timed()
tot = lib_ms = 0.0
for _ in range(20):
    t0 = time.perf_counter()
    seg = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, 10000, w.params, segments=seg_mode)
    tot += time.perf_counter() - t0
    lib_ms += ctx.timing()[12]
print("per call: Context.basefc %.3f ms, raw xg_basefc %.3f ms, library clock %.3f ms" % (1e3 * tot / 20, 1e3 * acc["raw"] / 20, lib_ms / 20))
