"""Test-only backends: the host mirror (fc_wrapper / afc_wrapper) with the CUDA counting
call replaced by the CPU oracle.  Used to pin the oracle (and the host logic around it)
against the reference's golden outputs on machines without a GPU."""

import numpy as np

from oracle import oracle
from xcltk_b200 import lib
from xcltk_b200.utils.sam import build_tid_maps


def decode_host(sam_fn_list, chroms, cell_tag, umi_tag, want_seq, n_threads=2):
    if sam_fn_list[0].endswith(".npz"):
        from npz_reads import load_npz_reads
        return load_npz_reads(sam_fn_list[0], list(chroms))
    ks = lib.KeySpace()
    bam_refs = [lib.bam_references(fn) for fn in sam_fn_list]
    gid_of, tid_maps = build_tid_maps(bam_refs, list(chroms))
    host = lib.decode_bams(sam_fn_list, tid_maps, cell_tag, umi_tag, want_seq, ks, n_threads)
    return host, ks, gid_of


def feature_arrays(regs, gid_of):
    gid = np.array([gid_of.get(r.chrom, -1) for r in regs], dtype=np.int32)
    beg = np.array([r.start - 1 for r in regs], dtype=np.int64)
    end = np.array([r.end - 1 for r in regs], dtype=np.int64)
    bad = (beg < 0) | (end <= beg) | (end > 2147483647)
    gid[bad] = -1
    beg[bad] = 0
    end[bad] = 0
    return gid, beg.astype(np.int32), end.astype(np.int32)


def oracle_count_features(conf, batch=None):
    regs = conf.reg_list
    chroms = list(dict.fromkeys(r.chrom for r in regs))
    host, ks, gid_of = decode_host(conf.sam_fn_list, chroms, conf.cell_tag, conf.umi_tag, False)
    gid, beg, end = feature_arrays(regs, gid_of)
    keys = np.array([ks.encode(b) for b in conf.barcodes], dtype=np.uint64) if conf.use_barcodes() else None
    row, col, val = oracle.basefc(host, gid, beg, end, keys, len(conf.samples), oracle.params(conf), n_threads=2)
    conf.last_timing = [0.0] * 8
    conf.last_stats = {"n_reads": host.n}
    host.close()
    return row, col, val


def oracle_count_regions(conf, regs, batch=None):
    snps = conf.snp_set.snps
    chroms = list(dict.fromkeys(s.chrom for s in snps))
    host, ks, gid_of = decode_host(conf.sam_fn_list, chroms, conf.cell_tag, conf.umi_tag, True)
    gid = np.array([gid_of.get(s.chrom, -1) for s in snps], dtype=np.int32)
    pos0 = np.array([s.pos - 1 for s in snps], dtype=np.int64)
    gid[pos0 < 0] = -1
    pos0[pos0 < 0] = 0
    keys = np.array([ks.encode(b) for b in conf.barcodes], dtype=np.uint64) if conf.use_barcodes() else None
    reg_ptr = np.zeros(len(regs) + 1, dtype=np.int64)
    reg_snp = []
    for r, reg in enumerate(regs):
        if reg.snp_list:
            reg_snp.extend(s.index for s in reg.snp_list)
        reg_ptr[r + 1] = len(reg_snp)
    par = oracle.params(conf)          # baf has no include test: min_include defaults to 0
    ad, dp, oth = oracle.baf(host, gid, pos0.astype(np.int32), "".join(s.ref for s in snps),
                             "".join(s.alt for s in snps), [s.ref_idx for s in snps],
                             [s.alt_idx for s in snps], reg_ptr, np.array(reg_snp, dtype=np.int32), keys,
                             len(conf.samples), par, conf.min_count, conf.min_maf, conf.no_dup_hap, n_threads=2)
    host.close()
    shape = (len(regs), len(conf.samples))
    return tuple((m[0], m[1], m[2], shape) for m in (ad, dp, oth))
