"""Turn ncu outputs into small text summaries for profiles/ (run in the build container).

  python tools/summarize_ncu.py launches <launches.csv> > profiles/rNN_launches.txt
  python tools/summarize_ncu.py report <prof.ncu-rep> [kernel name substring]  > profiles/rNN_<kernel>.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i]
    ci = {h: k for k, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[i + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[ci["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0, r[ci["Grid Size"]], r[ci["Block Size"]]])
        a[0] += 1
        a[1] += float(r[ci["Metric Value"]]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)")
    print("%-30s %6s %12s %7s   %s" % ("kernel", "n", "total ms", "share", "grid x block (last)"))
    for k, v in agg.items():
        print("%-30s %6d %12.3f %6.1f%%   %s x %s" % (k, v[0], v[1], 100 * v[1] / tot, v[2], v[3]))
    print("%-30s %6s %12.3f" % ("total", "", tot))


def report(path, pick=None):
    """pick: substring of the kernel name (metrics and source page of that kernel only)"""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for v in rows[2:]:
        if pick and pick not in v[hdr.index("Kernel Name")]:
            continue
        print("== kernel:", v[hdr.index("Kernel Name")][:100])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print("  %-82s %-14s %s" % (h, units[i], v[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    if pick:
        blocks = [b for b in blocks if pick in b["name"]]
    for b in blocks[:1]:
        hdr = b["rows"][0]
        data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
        ci = {h: i for i, h in enumerate(hdr)}
        tot_s = sum(int(r[ci["# Samples"]] or 0) for r in data) or 1
        tot_i = sum(int(r[ci["Instructions Executed"]] or 0) for r in data) or 1
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((h, sum(int(r[ci[h]] or 0) for r in data)) for h in stalls), key=lambda x: -x[1])
        print("== stall samples by reason (source page):", ", ".join("%s %.1f%%" % (h, 100.0 * n / tot_s) for h, n in agg[:8]))
        print("== SASS lines with > 1.2%% of the stall samples (of %d SASS instructions, %d warp-instructions executed)" % (len(data), tot_i))
        for k, r in enumerate(data):
            s = int(r[ci["# Samples"]] or 0)
            if s > tot_s * 0.012:
                top = max(stalls, key=lambda h: int(r[ci[h]] or 0))
                print("  %5d %-62s %5.1f%%  lanes %4.1f  %s" % (k, r[ci["Source"]][:62], 100.0 * s / tot_s,
                                                              float(r[ci["Avg. Threads Executed"]] or 0), top))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](*sys.argv[2:4])
