"""baf pileup + count on the bench shape: wall vs device time, result sizes, host profile."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcltk_b200 import engine, workload  # noqa: E402

ctx = engine.get_context(0)
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50000000
b = workload.make_baf_workload(ctx, n_reads, 5000, 200000, seed=11)


def step():
    t0 = time.perf_counter()
    totals, st = ctx.baf_pileup(b.dreads, b.snp_gid, b.snp_pos, b.cell_keys, b.n_cells, b.params)
    t1 = time.perf_counter()
    tp = ctx.timing()
    keep = (totals.sum(axis=1) >= 1).astype(np.uint8)
    t2 = time.perf_counter()
    ad, dp, oth = ctx.baf_count(st, b.reg_ptr, b.reg_snp, b.hap_of, keep, True)
    t3 = time.perf_counter()
    tc = ctx.timing()
    st.close()
    return (1e3 * (t1 - t0), tp[0], 1e3 * (t2 - t1), 1e3 * (t3 - t2), tc[0], len(ad[2]), len(dp[2]), len(oth[2]), tp[6], tc[6])


for k in range(4):
    print("pileup wall %.2f dev %.2f | keep %.2f | count wall %.2f dev %.2f | nnz AD %d DP %d OTH %d | pairs %d combos %d" % step())
pr = cProfile.Profile()
pr.enable()
for k in range(5):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
