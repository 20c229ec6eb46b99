"""Cold-start costs of one CLI-like run: context creation, first device decode, first basefc."""
import os
import sys
import time

t0 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from xcltk_b200 import engine, lib, synth  # noqa: E402
print("imports %.0f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
bam = "/tmp/first_call.bam"
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 6000000
contigs = [("chr%d" % c, 100000000) for c in range(1, 6)]
if not os.path.exists(bam):
    synth.write_fast_bam(bam, n, contigs, 1000, seed=5, threads=os.cpu_count() or 1)
t = time.perf_counter()
lib.load()
print("lib.load %.0f ms" % (1e3 * (time.perf_counter() - t)), flush=True)
t = time.perf_counter()
ctx = engine.get_context(0)
print("xg_create %.0f ms" % (1e3 * (time.perf_counter() - t)), flush=True)
maps = [np.arange(5, dtype=np.int32)]
for k in range(3):
    t = time.perf_counter()
    d, seen = ctx.decode_bams([bam], maps, "CB", "UB", False)
    print("decode %d: %.0f ms" % (k, 1e3 * (time.perf_counter() - t)), flush=True)
    d.close()
