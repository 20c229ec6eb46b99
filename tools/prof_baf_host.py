"""Where the wall time of the two baf calls goes (host phases reported by the library + Python around them)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xcltk_b200 import engine, workload
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50000000
ctx = engine.get_context(0)
b = workload.make_baf_workload(ctx, n, 5000, 200000, seed=8)
for rep in range(4):
    t0 = time.perf_counter()
    totals, st = ctx.baf_pileup(b.dreads, b.snp_gid, b.snp_pos, b.cell_keys, 5000, b.params, reuse_totals=True)
    t1 = time.perf_counter()
    tp = ctx.timing()
    keep = ((totals[:, 0] + totals[:, 1] + totals[:, 2] + totals[:, 3] + totals[:, 4]) >= 1).astype(np.uint8)
    t2 = time.perf_counter()
    ad, dp, oth = ctx.baf_count(st, b.reg_ptr, b.reg_snp, b.hap_of, keep, True)
    t3 = time.perf_counter()
    tc = ctx.timing()
    st.close()
    t4 = time.perf_counter()
    print("scan %.3f ms pairs %d | pileup wall %.2f ms (device %.2f; host phases %.2f / %.2f / %.2f)  keep %.2f  count wall %.2f ms (device %.2f; host %.2f / %.2f / %.2f)  close %.2f" % (
        tp[1], int(tp[6]), 1e3 * (t1 - t0), tp[0], tp[8], tp[9], tp[10], 1e3 * (t2 - t1), 1e3 * (t3 - t2), tc[0], tc[8], tc[9], tc[10], 1e3 * (t4 - t3)))
for rep in range(4):
    t0 = time.perf_counter()
    ad, dp, oth = ctx.baf_fc(b.dreads, b.snp_gid, b.snp_pos, b.cell_keys, 5000, b.params, b.snp_ref, b.snp_alt, 1, 0.0,
                             b.reg_ptr, b.reg_snp, b.hap_of, True)
    t1 = time.perf_counter()
    tm = ctx.timing()
    print("fused: wall %.2f ms (device %.2f, scan %.3f; host at pileup end %.2f / count queued %.2f / done %.2f) launches %d nnz %d" % (
        1e3 * (t1 - t0), tm[0], tm[1], tm[8], tm[9], tm[10], int(tm[2]), len(ad[2]) + len(dp[2]) + len(oth[2])))
