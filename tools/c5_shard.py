"""Config C5 (atlas shape: 100k cells, 1 Mb bins, 2B reads over 8 GPUs): the shard ONE of the 8 GPUs counts.

    python tools/c5_shard.py [total reads] [world] [rank] [cells]

Parity first (the same generator at 3M reads, 100k cells, 1 Mb bins, against the CPU oracle), then the shard
`rank` of `world` of the library (contiguous genomic chunk, reads with halo, as bench.py --scaling strong cuts it):
basefc wall / device time.  Nothing is exchanged between the GPUs, so the 8-GPU time is the slowest shard's.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from xcltk_b200 import engine, workload  # noqa: E402

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2000000000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cells = int(sys.argv[4]) if len(sys.argv) > 4 else 100000
ctx = engine.get_context(0)

from oracle import oracle  # noqa: E402
w = workload.make_basefc_workload(ctx, 3000000, cells, 0, seed=5, bins_kb=1000)
host = w.dreads.download()
exp = oracle.basefc(host, w.gid, w.beg, w.end, w.cell_keys, cells, oracle.params(workload.Conf()), os.cpu_count() or 1)
row, col, val, _ = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, cells, w.params)
ok = all(np.array_equal(a, b) for a, b in zip((row, col, val), exp))
print("parity vs oracle (3M reads, %d cells, %d bins of 1 Mb): %s, nnz %d" % (cells, len(w.gid), "OK" if ok else "MISMATCH", len(val)), flush=True)
host.close()
w.dreads.close()

t = time.perf_counter()
w = workload.make_basefc_workload(ctx, total, cells, 0, seed=7, bins_kb=1000, part=(rank, world))
print("shard %d/%d of a %.2e-read library: %d bins, %d reads generated in %.1f s" % (
    rank, world, total, len(w.gid), w.n_reads, time.perf_counter() - t), flush=True)
best = None
for rep in range(4):
    t = time.perf_counter()
    seg = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, cells, w.params, segments=os.environ.get("TUNE_SEGMENTS", "tiny"))
    dt = 1e3 * (time.perf_counter() - t)
    tm = ctx.timing()
    line = (dt, tm[0], tm[3], tm[1], int(tm[5]), tm[6] / 1e9, seg.nnz, int(tm[14]), int(tm[15]))
    if best is None or dt < best[0]:
        best = line
print("basefc on the shard: call %.1f ms, device %.1f, epochs-span %.1f, count kernels %.1f ms, %d epochs, pool %.2f GB, nnz %d, "
      "segment / set features %d / %d -> %.2e reads/s on this GPU" % (best + (w.n_reads / (best[0] * 1e-3),)), flush=True)
