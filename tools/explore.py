"""Ad-hoc exploration on the GPU box: synthetic reads -> basefc / baf timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xcltk_b200 import lib, engine, workload

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 500
n_feat = int(sys.argv[3]) if len(sys.argv) > 3 else 33472
ctx = engine.get_context(0)
w = workload.make_basefc_workload(ctx, n_reads, n_cells, n_feat, seed=7)
for it in range(3):
    t = time.time()
    row, col, val, shape = ctx.basefc(w.dreads, w.gid, w.beg, w.end, w.cell_keys, n_cells, w.params)
    dt = time.time() - t
    tm = ctx.timing()
    print("basefc n=%d cells=%d feats=%d nnz=%d sum=%d wall=%.1fms kern=%.2fms count=%.2fms epochs=%d launches=%d pool=%.1fMB staging=%.1fM d2h=%.2fms" % (
        n_reads, n_cells, n_feat, len(val), int(val.sum()), dt * 1e3, tm[0], tm[1], tm[5], tm[2], tm[6] / 1e6, tm[7] / 1e6, tm[4]))
    print("   host ms: index=%.1f windows=%.1f plan=%.1f upload=%.1f call=%.1f span=%.2f" % (tm[8], tm[9], tm[10], tm[11], tm[12], tm[3]))

if len(sys.argv) > 4:
    nb = int(float(sys.argv[4]))
    wb = workload.make_baf_workload(ctx, nb, 5000, 200000, seed=7)
    for it in range(3):
        t = time.time()
        totals, st = ctx.baf_pileup(wb.dreads, wb.snp_gid, wb.snp_pos, wb.cell_keys, 5000, wb.params)
        t1 = ctx.timing()
        keep = (totals.sum(axis=1) >= 1).astype(np.uint8)
        ad, dp, oth = ctx.baf_count(st, wb.reg_ptr, wb.reg_snp, wb.hap_of, keep, True)
        t2 = ctx.timing()
        st.close()
        print("baf n=%d pairs=%d kept=%d AD=%d DP=%d OTH=%d wall=%.1fms pileup=%.2fms scan=%.2fms count=%.2fms launches=%d+%d" % (
            nb, t1[6], keep.sum(), len(ad[2]), len(dp[2]), len(oth[2]), (time.time() - t) * 1e3, t1[0], t1[1], t2[0], t1[2], t2[2]))
