"""Decode throughput: host decoder (xg_decode_bams + upload) vs device decoder on a synthetic
coordinate-sorted 10x-style BAM.  usage: prof_decode.py [n_reads] [host_threads]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcltk_b200 import engine, lib, synth  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5000000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
path = "/tmp/prof_decode_%d.bam" % n
contigs = [("chr%d" % (i + 1), 100000000) for i in range(8)]
t0 = time.time()
if not os.path.exists(path):
    synth.write_fast_bam(path, n, contigs, n_cells=5000, seed=3, threads=min(threads, 32))
print("bam: %.1f MB written in %.1f s" % (os.path.getsize(path) / 1e6, time.time() - t0), flush=True)
ctx = engine.get_context(0)
maps = [np.arange(len(contigs), dtype=np.int32)]
umi = None if os.environ.get("PROF_UMI", "UB") == "None" else "UB"       # None: query-name keys, interned on the host
for want_seq in (False, True):
    for rep in range(2):
        t0 = time.time()
        dev, seen = ctx.decode_bams([path], maps, "CB", umi, want_seq, lib.KeySpace())
        dt = time.time() - t0
        t = ctx.timing()
        print("device want_seq=%d: %.3f s  %.1f Mreads/s | pread %.0f ms, read+h2d+inflate %.1f ms, walk %.1f ms, "
              "extract %.1f ms, alloc %.0f ms, trim %.0f ms, interned %d (%d distinct) in %.0f ms, call %.0f ms" % (want_seq, dt, seen / dt / 1e6, t[8], t[4], t[2], t[3], t[9], t[10], int(t[6]), int(t[11]), t[7], t[12]),
              flush=True)
        dev.close()
    if os.environ.get("PROF_SKIP_HOST"):
        continue
    t0 = time.time()
    ks = lib.KeySpace()
    host = lib.decode_bams([path], maps, "CB", umi, want_seq, ks, threads)
    t1 = time.time()
    d = ctx.upload(host)
    t2 = time.time()
    print("host   want_seq=%d (%d threads): decode %.3f s + upload %.3f s  %.1f Mreads/s" % (
        want_seq, threads, t1 - t0, t2 - t1, host.n / (t2 - t0) / 1e6), flush=True)
    d.close()
    host.close()
