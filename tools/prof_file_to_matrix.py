"""Where the time of the user's call goes: fc_wrapper(BAM, barcodes, features, out_dir) on a synthetic
10x-style BAM (cProfile, cumulative).  usage: prof_file_to_matrix.py [n_reads] [n_windows_per_contig]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcltk_b200 import synth  # noqa: E402
from xcltk_b200.rdr.fc.main import fc_wrapper  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 6000000
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
td = tempfile.mkdtemp()
bam = os.path.join(td, "s.bam")
t0 = time.time()
barcodes = synth.write_fast_bam(bam, n, [("chr%d" % c, 100000000) for c in range(1, 6)], 1000, seed=5,
                                threads=os.cpu_count() or 1)
print("bam %.1f MB in %.1f s" % (os.path.getsize(bam) / 1e6, time.time() - t0), flush=True)
bc_fn, ft_fn = os.path.join(td, "barcodes.tsv"), os.path.join(td, "features.tsv")
open(bc_fn, "w").write("".join(b + "\n" for b in barcodes))
with open(ft_fn, "w") as fp:
    step = 100000000 // nw
    for c in range(1, 6):
        for k in range(nw):
            fp.write("chr%d\t%d\t%d\tw%d_%d\n" % (c, k * step + 1, k * step + (step * 3 // 2 if k & 1 else step), c, k))
for k in range(2):
    t = time.perf_counter()
    fc_wrapper(bam, bc_fn, ft_fn, os.path.join(td, "warm%d" % k), ncores=os.cpu_count() or 1)
    print("call %d: %.1f ms" % (k, 1e3 * (time.perf_counter() - t)), flush=True)
for mode in ("0", "1") * int(os.environ.get("PROF_AB_ROUNDS", "2")):              # Matrix-Market text on the host / on the device
    os.environ["XCLTK_B200_DEVICE_MTX"] = mode
    t = time.perf_counter()
    fc_wrapper(bam, bc_fn, ft_fn, os.path.join(td, "ab" + mode), ncores=os.cpu_count() or 1)
    print("XCLTK_B200_DEVICE_MTX=%s: %.1f ms" % (mode, 1e3 * (time.perf_counter() - t)), flush=True)
os.environ.pop("XCLTK_B200_DEVICE_MTX")
if os.environ.get("PROF_AB_ONLY"):
    sys.exit(0)
pr = cProfile.Profile()
pr.enable()
fc_wrapper(bam, bc_fn, ft_fn, os.path.join(td, "prof"), ncores=os.cpu_count() or 1)
pr.disable()
print("mtx %.1f MB" % (os.path.getsize(os.path.join(td, "prof", "matrix.mtx")) / 1e6))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
