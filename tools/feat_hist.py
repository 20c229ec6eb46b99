"""Distribution of the finalize work over the features of the bench shape: words per feature ~ row non-zeros x (words / nnz)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xcltk_b200 import engine, workload
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300000000
ctx = engine.get_context(0)
w = workload.make_basefc_workload(ctx, n_reads, 10000, 60000, seed=7)
seg = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, 10000, w.params, segments=os.environ.get("TUNE_SEGMENTS", "tiny"))
cnt = np.asarray(seg.row_cnt, dtype=np.int64)
r, c, v = seg.to_sorted()
words = np.bincount(r, weights=v, minlength=len(cnt)).astype(np.int64)      # distinct (cell, UMI) per feature <= appended words
print("features %d, nnz %d, distinct triples %d" % (len(cnt), cnt.sum(), words.sum()))
edges = [0, 1, 64, 256, 819, 1638, 3276, 6553, 13107, 26214, 52428, 104857, 1 << 40]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (words >= lo) & (words < hi)
    print("words [%7d, %7d): %6d features, %5.1f %% of the words, %5.1f %% of nnz, mean nnz/feature %.0f" % (
        lo, min(hi, 1 << 30), m.sum(), 100.0 * words[m].sum() / max(1, words.sum()), 100.0 * cnt[m].sum() / max(1, cnt.sum()),
        cnt[m].mean() if m.any() else 0))
