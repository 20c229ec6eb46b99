"""Ad-hoc robustness check of the device decoder: random byte flips in a small BAM must end in a clean
error, a decline or a valid batch -- never in a CUDA fault.  usage: fuzz_decode.py [n_trials]"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pathlib  # noqa: E402
import test_gpu_decode as t  # noqa: E402
from xcltk_b200 import engine, lib  # noqa: E402

n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = engine.get_context(0)
d = pathlib.Path("/tmp/fuzz_decode")
d.mkdir(exist_ok=True)
good = t.tenx_bam(d, 4000, 99, "good.bam", level=6)
raw = open(good, "rb").read()
maps = t.full_maps([good])
rng = random.Random(1)
outcomes = {}
for k in range(n_trials):
    b = bytearray(raw)
    for _ in range(rng.choice([1, 1, 2, 5])):
        pos = rng.randrange(200, len(b) - 40)
        b[pos] ^= 1 << rng.randrange(8)
    p = str(d / "bad.bam")
    open(p, "wb").write(bytes(b))
    try:
        res = ctx.decode_bams([p], maps, "CB", "UB", True, lib.KeySpace())
        if res is None:
            out = "declined"
        else:
            out = "decoded"
            res[0].close()
    except lib.XgError as e:
        out = "error %d" % e.code
        if e.code == -5:
            print("CUDA error at trial", k, e)
            raise
    outcomes[out] = outcomes.get(out, 0) + 1
print(outcomes)
# the context still works
res = ctx.decode_bams([good], maps, "CB", "UB", True, lib.KeySpace())
print("good file after the fuzz:", res[1], "records")
