"""Worker of tests/test_parallel.py: one rank of a world_size-N gloo group.  Each rank counts its
genomic shard of the features of a golden case (CPU oracle as the counting backend -- the host
sharding / merge logic is what is under test), rank 0 gathers and merges the rows."""

import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_path):
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    import oracle_backend
    from oracle import oracle
    from util import resolve
    from xcltk_b200 import lib, parallel
    from xcltk_b200.rdr.fc.main import feature_arrays, load_region_from_txt

    r = resolve("c1_chr22_10x", "rdr_defaults")
    regs = load_region_from_txt(r["features"])
    with open(r["barcodes"]) as fp:
        barcodes = sorted(x.strip() for x in fp)
    chroms = list(dict.fromkeys(x.chrom for x in regs))
    host, ks, gid_of = oracle_backend.decode_host(r["sam"], chroms, "CB", "UB", False)
    gid, beg, end = feature_arrays(regs, gid_of)
    tile_pos = {}
    for rec_beg, n_rec, run, first_pos, max_end in host.tiles():
        tile_pos.setdefault(run, []).append(first_pos)
    load, total = parallel.reads_before(
        gid, beg, host.runs, lambda k: (np.asarray(tile_pos.get(k, []), dtype=np.int64), float(lib.XG_TILE)))
    shards = parallel.partition(gid, beg, load, total, world)
    sh = shards[rank]

    class Conf(object):
        min_mapq, min_len, min_include, incl_flag, excl_flag, no_orphan = 20, 30, 0.9, 0, 772, True
        use_barcodes = staticmethod(lambda: True)
        use_umi = staticmethod(lambda: True)
    keys = np.array([ks.encode(b) for b in barcodes], dtype=np.uint64)
    part = oracle.basefc(host, gid[sh], beg[sh], end[sh], keys, len(barcodes), oracle.params(Conf()), 1)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((part, sh), gathered, dst=0)
    # timing reduction of bench.py: max over ranks, sum over ranks
    import torch
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([float(len(sh))], dtype=torch.float64)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    if rank == 0:
        row, col, val = parallel.merge_coo([g[0] for g in gathered], [g[1] for g in gathered], len(regs))
        full = oracle.basefc(host, gid, beg, end, keys, len(barcodes), oracle.params(Conf()), 1)
        with open(out_path, "wb") as fp:
            pickle.dump({"merged": (row, col, val), "full": full, "max": float(t.item()), "sum": float(s.item()),
                         "n_feat": len(regs), "shard_sizes": [len(g[1]) for g in gathered]}, fp)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
