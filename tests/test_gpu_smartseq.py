"""Config C4 (SMART-seq, well-based): many per-cell BAMs, no cell barcode, no UMI tag -- the column is the BAM's
place in the list (sample IDs) and reads are collapsed by query name (xcltk/rdr/fc/main.py:353-369,425-429;
rdr/fc/mcount.py:36-43; baf/fc/mcount.py:113-116).  The public entry points on the GPU against the same entry
points with the CPU oracle as counting backend, byte for byte."""

import logging
import os

import numpy as np
import pytest

import oracle_backend
from util import BAF_FILES, RDR_FILES, compare_dirs

pytestmark = pytest.mark.gpu
logging.disable(logging.CRITICAL)


def smartseq_bams(ctx, td, n_bams, reads_per_bam, chroms, want_seq=False, snps=None):
    """n_bams BAMs of `reads_per_bam` records each, generated in HBM and written with the molecule's UMI text as the
    query name (records of a molecule share their name the way mates do); no CB / UB tags in the files."""
    from xcltk_b200 import lib, workload
    names = [c for c in workload.HG38_CHROMS if c in chroms]
    contigs = [(c, workload.HG38_LEN[c]) for c in names]
    paths, w = [], None
    for b in range(n_bams):
        if want_seq:
            w = workload.make_baf_workload(ctx, reads_per_bam, 40, snps, seed=100 + b, chroms=chroms, snp_seed=5)
        else:
            w = workload.make_basefc_workload(ctx, reads_per_bam, 40, 33472, seed=100 + b, chroms=chroms)
        host = w.dreads.download()
        p = os.path.join(td, "cell%03d.bam" % b)
        lib.write_bam(p, host, contigs, None, None, None, level=1, n_threads=4, name_from_umi=True)
        paths.append(p)
        host.close()
        w.dreads.close()
    return paths, w


def test_basefc_many_cell_bams_by_query_name(gpu_ctx, tmp_path, monkeypatch):
    from xcltk_b200.rdr.fc import main as rdr_main
    n_bams = 48
    paths, w = smartseq_bams(gpu_ctx, str(tmp_path), n_bams, 15000, {"20", "21", "22"})
    lst, ids, feats = str(tmp_path / "bams.lst"), str(tmp_path / "ids.tsv"), str(tmp_path / "features.tsv")
    open(lst, "w").write("".join(p + "\n" for p in paths))
    open(ids, "w").write("".join("well_%03d\n" % b for b in range(n_bams)))
    open(feats, "w").write("".join("%s\t%d\t%d\t%s\n" % f for f in w.feats))
    kw = dict(sam_list_fn=lst, sample_id_fn=ids, cell_tag=None, umi_tag=None, ncores=4)
    got = str(tmp_path / "gpu")
    assert rdr_main.fc_wrapper(None, None, feats, got, **kw) == 0
    monkeypatch.setattr(rdr_main, "count_features", oracle_backend.oracle_count_features)
    exp = str(tmp_path / "cpu")
    assert rdr_main.fc_wrapper(None, None, feats, exp, **kw) == 0
    compare_dirs(exp, got, RDR_FILES)
    with open(os.path.join(got, "matrix.mtx")) as fp:
        dims = fp.read().split("\n", 3)[2].split("\t")
    assert int(dims[1]) == n_bams and int(dims[2]) > 10000
    assert open(os.path.join(got, "barcodes.tsv")).read().split("\n")[1] == "well_001"


def test_baf_many_cell_bams_by_query_name(gpu_ctx, tmp_path, monkeypatch):
    from xcltk_b200.baf.fc import main as baf_main
    n_bams = 16
    paths, w = smartseq_bams(gpu_ctx, str(tmp_path), n_bams, 40000, {"21", "22"}, want_seq=True, snps=1500)
    lst, feats, snps = str(tmp_path / "bams.lst"), str(tmp_path / "features.tsv"), str(tmp_path / "snps.tsv")
    open(lst, "w").write("".join(p + "\n" for p in paths))
    open(feats, "w").write("".join("%s\t%d\t%d\t%s\n" % f for f in w.feats))
    names = [c for c in ("21", "22")]
    with open(snps, "w") as fp:
        fp.write("chrom\tpos\tref\talt\tref_hap\talt_hap\n")
        for g, p, r, a, h in zip(w.snp_gid, w.snp_pos, w.snp_ref, w.snp_alt, w.snp_ref_hap):
            fp.write("%s\t%d\t%s\t%s\t%d\t%d\n" % (names[int(g)], int(p) + 1, "ACGT"[r], "ACGT"[a], int(h), 1 - int(h)))
    kw = dict(sam_list_fn=lst, cell_tag=None, umi_tag=None, ncores=4, output_all_reg=True,
              sample_ids=",".join("w%d" % b for b in range(n_bams)))
    got = str(tmp_path / "gpu")
    assert baf_main.afc_wrapper(None, None, feats, snps, got, **kw) == 0
    monkeypatch.setattr(baf_main, "count_regions", oracle_backend.oracle_count_regions)
    exp = str(tmp_path / "cpu")
    assert baf_main.afc_wrapper(None, None, feats, snps, exp, **kw) == 0
    compare_dirs(exp, got, BAF_FILES)
    with open(os.path.join(got, "xcltk.DP.mtx")) as fp:
        assert int(fp.read().split("\n", 3)[2].split("\t")[2]) > 500


def test_sharded_file_counting_equals_unsharded_at_scale(gpu_ctx, tmp_path, monkeypatch):
    """3M reads over chr19-22 with spliced reads (5 kb introns) and genes up to 1 Mb: the library split into 6 byte
    ranges gives the matrix of the unsplit file, and no shard decodes much more than its share."""
    from xcltk_b200 import engine, lib, workload
    from xcltk_b200.rdr.fc import main as rdr_main
    chroms = {"19", "20", "21", "22"}
    w = workload.make_basefc_workload(gpu_ctx, 3000000, 800, 33472, seed=77, chroms=chroms)
    host = w.dreads.download()
    names = [c for c in workload.HG38_CHROMS if c in chroms]
    bam = str(tmp_path / "lib.bam")
    lib.write_bam(bam, host, [("chr" + c, workload.HG38_LEN[c]) for c in names], None, "CB", "UB", level=1, n_threads=4)
    host.close()
    ks = lib.KeySpace()
    bc, ft = str(tmp_path / "bc.tsv"), str(tmp_path / "ft.tsv")
    open(bc, "w").write("".join(ks.decode(int(k)) + "\n" for k in w.cell_keys))
    open(ft, "w").write("".join("%s\t%d\t%d\t%s\n" % f for f in w.feats))
    w.dreads.close()
    one = str(tmp_path / "one")
    assert rdr_main.fc_wrapper(bam, bc, ft, one, ncores=4) == 0
    monkeypatch.setenv("XCLTK_B200_GPUS", "6")
    monkeypatch.setenv("XCLTK_B200_EMULATE_SHARDS", "1")
    seen = []
    real = engine.load_reads_sharded
    monkeypatch.setattr(engine, "load_reads_sharded", lambda *a, **kw: seen.append(real(*a, **kw)) or seen[-1])
    six = str(tmp_path / "six")
    assert rdr_main.fc_wrapper(bam, bc, ft, six, ncores=4) == 0
    compare_dirs(one, six, RDR_FILES)
    st = seen[0].stats
    sizes = [sum(hi - lo for lo, hi in per) for per in st["byte_ranges"]]
    assert max(sizes) < 0.35 * os.path.getsize(bam) and st["max_span"] <= st["halo_bp"]
    assert st["n_reads"] < 1.25 * 3000000
