import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from xcltk_b200 import engine, workload
from oracle import oracle
ctx = engine.get_context(0)
n = int(float(sys.argv[1])); cells = int(sys.argv[2])
w = workload.make_basefc_workload(ctx, n, cells, 60000, seed=3)
host = w.dreads.download()
exp = oracle.basefc(host, w.gid, w.beg, w.end, w.cell_keys, cells, oracle.params(workload.Conf()), os.cpu_count())
row, col, val, _ = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, cells, w.params)
print("ok" if all(np.array_equal(a, b) for a, b in zip((row, col, val), exp)) else "MISMATCH", len(val), flush=True)
