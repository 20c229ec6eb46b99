"""SMART-seq-like input: many small BAMs (one per cell), query-name keys.  Device vs host decoder.
usage: prof_decode_multi.py [n_bams] [reads_per_bam]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcltk_b200 import engine, lib, synth  # noqa: E402

n_bams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
per = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200000
contigs = [("chr%d" % (i + 1), 100000000) for i in range(8)]
paths = []
t0 = time.time()
for b in range(n_bams):
    p = "/tmp/prof_multi_%d_%d.bam" % (per, b)
    if not os.path.exists(p):
        synth.write_fast_bam(p, per, contigs, n_cells=1, seed=100 + b, threads=os.cpu_count() or 1)
    paths.append(p)
print("%d BAMs, %.1f MB in %.1f s" % (n_bams, sum(os.path.getsize(p) for p in paths) / 1e6, time.time() - t0), flush=True)
ctx = engine.get_context(0)
maps = [np.arange(len(contigs), dtype=np.int32)] * n_bams
for umi in ("UB", None):
    for rep in range(2):
        ks = lib.KeySpace()
        t0 = time.time()
        dev, seen = ctx.decode_bams(paths, maps, None, umi, False, ks)
        dt = time.time() - t0
        t = ctx.timing()
        print("device umi=%s: %.3f s  %.1f Mreads/s | stream %.0f ms, interned %d (%d distinct) in %.0f ms, windows %d" % (
            umi, dt, seen / dt / 1e6, t[4], int(t[6]), int(t[11]), t[7], int(t[5])), flush=True)
        dev.close()
    ks = lib.KeySpace()
    t0 = time.time()
    host = lib.decode_bams(paths, maps, None, umi, False, ks, os.cpu_count() or 1)
    t1 = time.time()
    d = ctx.upload(host)
    print("host   umi=%s: decode %.3f s + upload %.3f s  %.1f Mreads/s" % (umi, t1 - t0, time.time() - t1, host.n / (time.time() - t0) / 1e6), flush=True)
    d.close()
    host.close()
