"""One basefc call on synthetic reads (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcltk_b200 import engine, workload
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000000
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
n_feat = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ctx = engine.get_context(0)
w = workload.make_basefc_workload(ctx, n_reads, n_cells, n_feat, seed=7)
for _ in range(reps):
    row, col, val, shape = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, n_cells, w.params)
    print(len(val), int(val.sum()), ctx.timing())
