"""Loads tests/golden/*/reads.npz (decoded record arrays of a BAM that cannot travel)."""

import numpy as np

from xcltk_b200 import lib
from xcltk_b200.utils.sam import resolve_tid


def load_npz_reads(path, chroms, use_cell_keys=True):
    """Returns (ArrayReads, KeySpace, gid_of) with contigs renumbered to `chroms` the way
    sam_fetch resolves names (records on other contigs keep gid -1 and are never fetched)."""
    z = np.load(path, allow_pickle=False)
    ks = lib.KeySpace()
    cell_keys = np.array([ks.encode(str(s)) for s in z["cell_names"]] + [lib.XG_KEY_NONE], dtype=np.uint64)
    umi_keys = np.array([ks.encode(str(s)) for s in z["umi_names"]] + [lib.XG_KEY_NONE], dtype=np.uint64)   # -1: no tag
    keys = np.stack([cell_keys[z["cell_idx"]], umi_keys[z["umi_idx"]]], axis=1)
    index = {}
    for tid, name in enumerate(z["ref_names"]):
        index.setdefault(str(name), tid)
    gid_of, tid_to_gid = {}, {}
    for g, c in enumerate(chroms):
        tid = resolve_tid(index, c)
        gid_of[c] = g
        if tid >= 0:
            tid_to_gid[tid] = g
    runs = [(b, tid_to_gid.get(int(tid), -1), rb, re_) for b, tid, rb, re_ in z["runs"]]
    reads = lib.ArrayReads(z["pos_end"], z["fmq"], z["cig_off"], keys, z["cigar"], runs, z["seq_off"], z["seq"],
                           int(z["max_aln_len"]), int(z["max_span"]))
    return reads, ks, gid_of
