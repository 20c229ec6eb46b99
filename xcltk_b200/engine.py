"""Host-side plumbing shared by basefc and baf: BAM decode -> HBM, parameter packing, output.

Replaces the process pool of the reference (`multiprocessing.Pool` over feature chunks,
xcltk/rdr/fc/main.py:191-235, xcltk/baf/fc/main.py:156-211): one context per GPU, the decoded
read batch is uploaded once and every feature / SNP is evaluated against it on the device.
"""

import logging
import math
import os

import numpy as np

from . import lib
from .utils.sam import build_tid_maps

INT32_MAX = 2147483647


def min_mapq_int(min_mapq):
    """`read.mapq < conf.min_mapq` with a possibly float threshold (rdr/fc/main.py:126,
    core.py:47): for integer mapq, mapq < x  <=>  mapq < ceil(x)."""
    return int(math.ceil(min_mapq))


def include_threshold(min_include, max_aln_len):
    """Integer form of the include test (rdr/fc/core.py:160-165).

    Fraction mode iff 0 < min_include < 1: keep iff not (m / float(n) < min_include); the
    table gives, for every aligned length n, the smallest m that is kept, found with the
    reference's own float expression (m / float(n) is monotone in m, so bisection is exact).
    Otherwise length mode: keep iff not (m < min_include)  <=>  m >= ceil(min_include).
    Returns (table or None, min_len)."""
    if 0 < min_include < 1:
        tab = np.empty(max(1, max_aln_len) + 1, dtype=np.int32)
        tab[0] = INT32_MAX            # n == 0: the reference's frac is None (never reached with min_len >= 1)
        for n in range(1, len(tab)):
            lo, hi = 0, n + 1         # smallest m in [0, n+1) with not (m/float(n) < f); n+1 = none
            while lo < hi:
                mid = (lo + hi) // 2
                if mid / float(n) < min_include:
                    lo = mid + 1
                else:
                    hi = mid
            tab[n] = lo if lo <= n else INT32_MAX
        return tab, 0
    return None, int(math.ceil(min_include))


class ReadBatch(object):
    """Decoded reads of all BAMs resident on one GPU + what is needed to address them."""

    def __init__(self, ctx, dreads, keyspace, gid_of, stats):
        self.ctx, self.dreads, self.keyspace, self.gid_of, self.stats = ctx, dreads, keyspace, gid_of, stats

    def close(self):
        if self.dreads is not None:
            self.dreads.close()
            self.dreads = None
        host = getattr(self, "host", None)
        if host is not None:             # zero-copy batch: the pinned records were still in use
            host.close()
            self.host = None


_contexts = {}


def get_context(device=0):
    """One xg_ctx per device, created on first use.  Raises XgError without a GPU: the
    counting paths have no CPU fallback."""
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = lib.Context(device)
        _contexts[device] = ctx
    return ctx


def device_decode_enabled():
    """The device decoder (BGZF inflate + BAM parse on the GPU) is tried first unless
    $XCLTK_B200_DEVICE_DECODE is 0; files it declines go through the host decoder."""
    return os.environ.get("XCLTK_B200_DEVICE_DECODE", "1") not in ("0", "", "no", "false")


def _device_decode(ctx, sam_fn_list, tid_maps, cell_tag, umi_tag, want_seq, keyspace=None):
    """(dreads, stats) or None."""
    if not device_decode_enabled():
        return None
    res = ctx.decode_bams(sam_fn_list, tid_maps, cell_tag, umi_tag, want_seq, keyspace)
    if res is None:
        logging.getLogger(__name__).info("device decoder declined (%s); decoding on the host.",
                                         getattr(ctx, "decode_fallback_reason", "?"))
        return None
    dreads, seen = res
    i = dreads.info()
    return dreads, {"n_reads": i["n_reads"], "n_records_seen": seen, "max_aln_len": i["max_aln_len"],
                    "max_span": i["max_span"], "bytes": i["bytes"], "decoder": "device"}


def load_reads(sam_fn_list, chroms, cell_tag, umi_tag, want_seq, n_threads=0, device=0, mapped=False,
               host_only=False):
    """Bring every BAM's reads to `device` as one batch.

    The device decoder is tried first (the compressed files cross PCIe, the batch is built in
    HBM); when it declines -- records crossing BGZF blocks, keys that need the intern table --
    the BAMs are decoded on the host (multi-threaded) and uploaded.
    chroms: distinct (already 'chr'-stripped) contig names the features / SNPs use; reads on
    other contigs can never be fetched by the reference and are dropped at decode time.
    Host decoder only -- mapped: keep the records in pinned host memory and copy only pos/end
    (baf pileup); host_only: no upload at all, the caller streams the batch with
    Context.basefc_host."""
    ctx = get_context(device)
    ks = lib.KeySpace()
    bam_refs = [lib.bam_references(fn) for fn in sam_fn_list]
    gid_of, tid_maps = build_tid_maps(bam_refs, list(chroms))
    dev = _device_decode(ctx, sam_fn_list, tid_maps, cell_tag, umi_tag, want_seq, ks)
    if dev is not None:
        return ReadBatch(ctx, dev[0], ks, gid_of, dev[1])
    host = lib.decode_bams(sam_fn_list, tid_maps, cell_tag, umi_tag, want_seq, ks, n_threads)
    stats = {"n_reads": host.n, "n_records_seen": host.n_records_seen, "max_aln_len": host.max_aln_len,
             "max_span": host.max_span, "bytes": host.nbytes()}
    if mapped or host_only:
        batch = ReadBatch(ctx, None if host_only else ctx.map_reads(host), ks, gid_of, stats)
        batch.host = host
        return batch
    dreads = ctx.upload(host)
    host.close()
    return ReadBatch(ctx, dreads, ks, gid_of, stats)


class MultiBatch(object):
    """The decoded batch resident on several GPUs (one full copy each; the genomic sharding
    happens on the feature side, see parallel.py) + the tile starts used to balance shards."""

    def __init__(self, batches, runs, tile_pos):
        self.batches, self.runs, self._tile_pos = batches, runs, tile_pos
        b0 = batches[0]
        self.keyspace, self.gid_of, self.stats = b0.keyspace, b0.gid_of, b0.stats

    def pos_of_run(self, r):
        """(sorted start positions sampled once per tile, reads per sample) of run r."""
        return self._tile_pos[r], float(lib.XG_TILE)

    def close(self):
        for b in self.batches:
            b.close()


def _tile_pos(runs, tiles):
    tile_pos = {}
    for rec_beg, n_rec, run, first_pos, max_end in tiles:
        tile_pos.setdefault(run, []).append(first_pos)
    tile_pos = {r: np.asarray(v, dtype=np.int64) for r, v in tile_pos.items()}
    for r in range(len(runs)):
        tile_pos.setdefault(r, np.zeros(0, dtype=np.int64))
    return tile_pos


def load_reads_multi(sam_fn_list, chroms, cell_tag, umi_tag, want_seq, n_threads=0, devices=(0,)):
    """The batch on every device in `devices`: each GPU decodes the files itself (device
    decoder, one host thread per GPU), else they are decoded once on the host and uploaded."""
    from . import parallel
    ctxs = [get_context(d) for d in devices]
    ks = lib.KeySpace()
    bam_refs = [lib.bam_references(fn) for fn in sam_fn_list]
    gid_of, tid_maps = build_tid_maps(bam_refs, list(chroms))
    devs = parallel.run_on_devices(len(ctxs), lambda k: _device_decode(ctxs[k], sam_fn_list, tid_maps, cell_tag,
                                                                       umi_tag, want_seq, ks))
    if all(d is not None for d in devs):
        runs, tiles = devs[0][0].index()
        return MultiBatch([ReadBatch(c, d[0], ks, gid_of, d[1]) for c, d in zip(ctxs, devs)], runs, _tile_pos(runs, tiles))
    for d in devs:
        if d is not None:
            d[0].close()
    host = lib.decode_bams(sam_fn_list, tid_maps, cell_tag, umi_tag, want_seq, ks, n_threads)
    stats = {"n_reads": host.n, "n_records_seen": host.n_records_seen, "max_aln_len": host.max_aln_len,
             "max_span": host.max_span, "bytes": host.nbytes()}
    runs = list(host.runs)
    tile_pos = _tile_pos(runs, host.tiles())
    dreads = parallel.run_on_devices(len(ctxs), lambda k: ctxs[k].upload(host))
    host.close()
    return MultiBatch([ReadBatch(c, d, ks, gid_of, stats) for c, d in zip(ctxs, dreads)], runs, tile_pos)


def make_params(conf, max_aln_len, with_include):
    tab, incl_len = (None, 0)
    if with_include:
        tab, incl_len = include_threshold(conf.min_include, max_aln_len)
    return lib.ParamsBox(min_mapq_int(conf.min_mapq), conf.min_len, conf.incl_flag, conf.excl_flag,
                         conf.no_orphan, conf.use_barcodes(), conf.use_umi(), tab, incl_len)


def write_mtx(path, n_rows_in, row, col, val, emitted, n_cols, n_threads=0):
    """MatrixMarket text exactly as merge_mtx writes it (rdr/fc/utils.py:65-67,80-86): header,
    `%%`, `nrow\tncol\tnnz`, then 1-based `row\tcol\tval` lines with the rows renumbered over
    the emitted features.  (row, col, val): 0-based, sorted by (row, col); emitted: bool per
    input row.  Formatting is done by the library's multi-threaded writer."""
    row = np.asarray(row)
    counts = np.bincount(row, minlength=n_rows_in) if len(row) else np.zeros(n_rows_in, dtype=np.int64)
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    out_row = np.where(emitted, np.cumsum(emitted), 0).astype(np.int32)
    lib.write_mtx(path, row_ptr, out_row, int(np.count_nonzero(emitted)), n_cols, col, val, n_threads)


def n_decode_threads(nproc):
    n = int(nproc) if nproc else 1
    return max(1, min(n, os.cpu_count() or 1))
