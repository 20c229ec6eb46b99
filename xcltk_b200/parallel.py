"""Multi-GPU sharding of the counting paths by contiguous genomic chunks.

The reference parallelises by cutting the feature list into `nproc` consecutive slices, one
forked worker each (xcltk/rdr/fc/main.py:191-235, xcltk/baf/fc/main.py:156-211), and merges the
per-worker shards by renumbering rows (merge_mtx, rdr/fc/utils.py:54-94).  Here the units are
the same -- features (basefc) / regions (baf) are independent, so there is no exchange step and
no collective -- but the cut is made along the genome and balanced by READS, not by feature
count: features are ordered by (contig, start) and split where the cumulative number of reads
starting before them crosses k/N of the total.  Every shard is counted on its own GPU against
the read batch (tiles that cannot touch the shard's features are skipped on the device), and
the per-shard rows go back to input order on the host.
"""

import threading

import numpy as np


def genomic_order(gid, beg):
    """Indices of the valid features sorted by (contig, start)."""
    gid = np.asarray(gid)
    beg = np.asarray(beg)
    valid = np.nonzero(gid >= 0)[0]
    return valid[np.lexsort((beg[valid], gid[valid]))]


def reads_before(gid, beg, runs, pos_of_run):
    """For every feature: (rank of its contig in the read stream, reads of that contig that
    start before the feature) -> a monotone load coordinate along the stream.

    runs: [(bam_idx, gid, rec_beg, rec_end)]; pos_of_run(r) -> sorted start positions of run r
    (callers pass the tile index `first_pos` with XG_TILE granularity or the exact array)."""
    gid = np.asarray(gid)
    beg = np.asarray(beg)
    load = np.zeros(len(gid), dtype=np.float64)
    per_gid = {}
    for r, (_b, g, rb, re_) in enumerate(runs):
        per_gid.setdefault(g, []).append(r)
    base = 0.0
    for g in sorted(per_gid):
        sel = np.nonzero(gid == g)[0]
        tot = 0.0
        acc = np.zeros(len(sel), dtype=np.float64)
        for r in per_gid[g]:
            pos, weight = pos_of_run(r)
            acc += np.searchsorted(pos, beg[sel], side="left") * weight
            tot += len(pos) * weight
        load[sel] = base + acc
        base += tot
    return load, base


def partition(gid, beg, load, total, n_shards):
    """Cut the genomically ordered features into <= n_shards contiguous groups of about equal
    read load.  Returns a list of index arrays (input indices); invalid features (gid < 0,
    never fetched) ride along with the first shard so that every row is owned exactly once."""
    order = genomic_order(gid, beg)
    n = len(np.asarray(gid))
    shards = []
    if len(order):
        cuts = np.searchsorted(load[order], [total * k / float(n_shards) for k in range(1, n_shards)], side="left")
        for part in np.split(order, cuts):
            shards.append(part)
    else:
        shards.append(order)
    while len(shards) < n_shards:
        shards.append(np.zeros(0, dtype=order.dtype))
    rest = np.setdiff1d(np.arange(n), order, assume_unique=True)
    shards[0] = np.concatenate([shards[0], rest]).astype(np.int64)
    return [np.sort(s.astype(np.int64)) for s in shards]


def merge_coo(parts, shard_rows, n_rows):
    """parts[k] = (row, col, val) with rows indexing shard_rows[k]; returns the union sorted by
    (global row, col) -- the host-side replacement of merge_mtx."""
    rows, cols, vals = [], [], []
    for (r, c, v), idx in zip(parts, shard_rows):
        if len(v):
            rows.append(np.asarray(idx)[np.asarray(r)])
            cols.append(np.asarray(c))
            vals.append(np.asarray(v))
    if not rows:
        z = np.zeros(0, dtype=np.int32)
        return z, z.copy(), z.copy()
    row = np.concatenate(rows).astype(np.int64)
    col = np.concatenate(cols).astype(np.int64)
    val = np.concatenate(vals)
    order = np.argsort(row * (int(col.max()) + 1) + col, kind="stable")
    return row[order].astype(np.int32), col[order].astype(np.int32), val[order].astype(np.int32)


def emulated():
    """$XCLTK_B200_EMULATE_SHARDS=1: all shards on device 0, one after the other -- the sharded path on a
    one-GPU box (tests)."""
    import os
    return os.environ.get("XCLTK_B200_EMULATE_SHARDS", "0") not in ("0", "", "no", "false")


def device_list(n):
    return (0,) * n if emulated() else tuple(range(n))


def run_on_devices(n, fn):
    """fn(k) for k in range(n), one host thread per device (ctypes calls release the GIL)."""
    if emulated():
        return [fn(k) for k in range(n)]
    out, err = [None] * n, [None] * n

    def work(k):
        try:
            out[k] = fn(k)
        except BaseException as e:        # re-raised in the caller's thread
            err[k] = e
    threads = [threading.Thread(target=work, args=(k,)) for k in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in err:
        if e is not None:
            for o in out:              # what the other devices produced must not outlive the failure
                _close_result(o)
            raise e
    return out


def _close_result(o):
    """close() whatever a per-device worker returned (a batch, a tuple holding one ...)."""
    if o is None:
        return
    if hasattr(o, "close"):
        try:
            o.close()
        except Exception:
            pass
    elif isinstance(o, (tuple, list)):
        for x in o:
            _close_result(x)


def n_devices(requested=None):
    """GPUs to shard over: `requested`, else $XCLTK_B200_GPUS, else 1."""
    import os
    n = requested if requested else int(os.environ.get("XCLTK_B200_GPUS", "1") or 1)
    return max(1, int(n))
