"""Per-source-line instruction and stall shares of one kernel of an ncu report (build container, no GPU).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [mangled-name substring] [min share %]

ncu's CSV source page is per SASS instruction; the line table comes from the cubin inside the in-tree .so
(`nvdisasm -g`), matched by instruction offset.  The .so must be the build the report was taken from.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
mangled = sys.argv[3] if len(sys.argv) > 3 else None
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7

# NCU_LAUNCH_SKIP=n: the (n+1)-th matching launch of the report
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip",
                      os.environ.get("NCU_LAUNCH_SKIP", "0"), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
kname = rows[start - 1][1] if start else ""
hdr = rows[start]
ci = {h: i for i, h in enumerate(hdr)}
inst = []
for r in rows[start + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        break
    inst.append(r)
base = int(inst[0][0], 16)

# line table
so = os.path.join(ROOT, "xcltk_b200", "_lib", "libxcltk_b200.so")
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
line_of = {}
want = mangled or re.sub(r"[^A-Za-z0-9_]", "", kre.split("|")[0])
for f in os.listdir(td):
    if not f.endswith(".cubin"):
        continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, take = None, None, False
    for ln in sass.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1)
            take = want in cur_fn and not line_of
            if take:
                line_of = {}
                kfn = cur_fn
            continue
        if not take:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur_line
    if line_of:
        break

agg = collections.OrderedDict()
tot_i = tot_s = 0.0
stall_cols = [h for h in hdr if h.startswith("stall_")]
for r in inst:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg.setdefault(key, [0.0, 0.0, 0.0, collections.Counter()])
    n_i = float(r[ci["Instructions Executed"]] or 0)
    n_t = float(r[ci["Thread Instructions Executed"]] or 0)
    n_s = float(r[ci["# Samples"]] or 0)
    a[0] += n_i
    a[1] += n_t
    a[2] += n_s
    for h in stall_cols:
        v = float(r[ci[h]] or 0)
        if v:
            a[3][h[6:]] += v
    tot_i += n_i
    tot_s += n_s
src_cache = {}


def src(key):
    f, n = key
    p = os.path.join(ROOT, "xcltk_b200", "csrc", f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[p]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""


print("kernel: %s" % kname)
print("warp instructions %.0f, stall samples %.0f; lines with >= %.1f %% of either:" % (tot_i, tot_s, min_share))
print("%-18s %6s %6s %5s  %-34s %s" % ("line", "inst%", "samp%", "lanes", "top stalls", "source"))
for key, a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    si, ss = 100 * a[0] / tot_i, 100 * a[2] / max(1.0, tot_s)
    if si < min_share and ss < min_share:
        continue
    top = " ".join("%s:%.0f" % (k, 100 * v / max(1.0, a[2])) for k, v in a[3].most_common(3))
    print("%-18s %6.2f %6.2f %5.1f  %-34s %s" % ("%s:%d" % key, si, ss, a[1] / max(1.0, a[0]), top, src(key)))
