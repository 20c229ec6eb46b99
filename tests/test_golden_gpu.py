"""GPU parity: the public Python entry points (fc_wrapper / afc_wrapper, which call the CUDA
path through the C-ABI) must reproduce, byte for byte, the files the unmodified reference
wrote for the same inputs (tests/golden/*/expected, made by oracle/make_golden.py)."""

import logging
import os

import pytest

from util import BAF_FILES, RDR_FILES, compare_dirs, golden_runs, read, resolve

pytestmark = pytest.mark.gpu
logging.disable(logging.CRITICAL)


def use_npz_reads(monkeypatch, r):
    """Cases whose BAM cannot travel feed the decoded record arrays to the same upload path."""
    if not r["reads_npz"]:
        return
    from npz_reads import load_npz_reads
    from xcltk_b200 import engine

    def load_reads(sam_fn_list, chroms, cell_tag, umi_tag, want_seq, n_threads=0, device=0, mapped=False,
                   host_only=False):
        ctx = engine.get_context(device)
        host, ks, gid_of = load_npz_reads(sam_fn_list[0], list(chroms))
        stats = {"n_reads": host.n, "n_records_seen": host.n, "max_aln_len": host.max_aln_len,
                 "max_span": host.max_span, "bytes": host.nbytes()}
        return engine.ReadBatch(ctx, ctx.upload(host), ks, gid_of, stats)
    monkeypatch.setattr(engine, "load_reads", load_reads)


@pytest.mark.parametrize("case,run", golden_runs("basefc"))
def test_basefc_matches_reference(case, run, tmp_path, gpu_ctx, monkeypatch):
    from xcltk_b200.rdr.fc.main import fc_wrapper
    r = resolve(case, run)
    use_npz_reads(monkeypatch, r)
    out = str(tmp_path / "out")
    ret = fc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], out, **r["kwargs"])
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, RDR_FILES)


@pytest.mark.parametrize("case,run", golden_runs("baf"))
def test_baf_matches_reference(case, run, tmp_path, gpu_ctx, monkeypatch):
    from xcltk_b200.baf.fc.main import afc_wrapper
    r = resolve(case, run)
    use_npz_reads(monkeypatch, r)
    out = str(tmp_path / "out")
    ret = afc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, BAF_FILES)


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("case,run", [("c1_chr22_10x", "rdr_defaults"), ("c1_chr22_10x", "rdr_umi_none"),
                                      ("d1_basefc_mini", "defaults"), ("d3_sample_mode", "rdr_ids"),
                                      ("c1_chr22_10x", "baf_all_reg_dup"), ("c1_chr22_10x", "baf_defaults"),
                                      ("d2_baf_mini", "defaults"), ("d3_sample_mode", "baf_p2p1")])
@pytest.mark.parametrize("reblock", [False, True])
def test_region_sharded_over_gpus_matches_reference(case, run, reblock, tmp_path, monkeypatch):
    """Product path sharded by genomic chunks over all GPUs of the box (>= 2), rows merged on
    the host: still byte-identical to the reference.  reblock: the BAMs in htslib's block layout,
    so that every GPU decodes them itself with the device decoder."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("XCLTK_B200_GPUS", str(min(n, 4)))
    r = resolve(case, run)
    if reblock:
        from xcltk_b200 import synth
        sams = []
        for k, p in enumerate(r["sam"]):
            q = str(tmp_path / ("%d_%s" % (k, os.path.basename(p))))
            synth.reblock_bam(p, q)
            sams.append(q)
        r["sam"] = sams
    out = str(tmp_path / "out")
    if r["kind"] == "basefc":
        from xcltk_b200.rdr.fc.main import fc_wrapper
        ret = fc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], out, **r["kwargs"])
        files = RDR_FILES
    else:
        from xcltk_b200.baf.fc.main import afc_wrapper
        ret = afc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
        files = BAF_FILES
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, files)


def _device_decoder_runs():
    """golden runs whose BAMs travel (the npz case has none)"""
    return [(c, r) for c, r in golden_runs() if not resolve(c, r)["reads_npz"]]


@pytest.mark.parametrize("case,run", _device_decoder_runs())
def test_goldens_through_the_device_decoder(case, run, tmp_path, gpu_ctx, monkeypatch):
    """The same golden runs with the BAMs laid out as htslib writes them (whole records per
    BGZF block), so that the device decoder takes them: BAM -> inflate + parse on the GPU ->
    counting kernels -> the reference's bytes.  Runs whose keys need the intern table (query-name
    UMIs, sample mode) have them interned by the host keyspace and stay on the device decoder."""
    from xcltk_b200 import engine, synth
    r = resolve(case, run)
    sams = []
    for k, p in enumerate(r["sam"]):
        q = str(tmp_path / ("%d_%s" % (k, os.path.basename(p))))
        synth.reblock_bam(p, q)
        sams.append(q)
    used = []
    real = engine._device_decode

    def spy(*a, **kw):
        res = real(*a, **kw)
        used.append(res is not None)
        return res
    monkeypatch.setattr(engine, "_device_decode", spy)
    out = str(tmp_path / "out")
    if r["kind"] == "basefc":
        from xcltk_b200.rdr.fc.main import fc_wrapper
        ret = fc_wrapper(",".join(sams), r["barcodes"], r["features"], out, **r["kwargs"])
        files = RDR_FILES
    else:
        from xcltk_b200.baf.fc.main import afc_wrapper
        ret = afc_wrapper(",".join(sams), r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
        files = BAF_FILES
    assert ret == int(read(r["expected"] + "/RETCODE"))
    if ret == 0:
        compare_dirs(r["expected"], out, files)
        if used:       # query-name UMIs and free-text tags go through the keyspace: nothing is declined
            assert all(used), "the device decoder declined a BAM in htslib layout"


@pytest.mark.parametrize("run", [r for c, r in golden_runs() if c == "bch869_smartseq"])
def test_bch869_bam_file_matches_reference(run, tmp_path, gpu_ctx, monkeypatch):
    """The reference's real BAM, as a file: device decoder (it is htslib-written) -> counting kernels -> the bytes
    the unmodified reference wrote for it."""
    from util import GOLD
    from xcltk_b200 import engine
    r = resolve("bch869_smartseq", run)
    bam = os.path.join(GOLD, "bch869_smartseq", "BCH869.output.bam")
    used = []
    real = engine._device_decode

    def spy(*a, **kw):
        res = real(*a, **kw)
        used.append(res is not None)
        return res
    monkeypatch.setattr(engine, "_device_decode", spy)
    out = str(tmp_path / "out")
    if r["kind"] == "basefc":
        from xcltk_b200.rdr.fc.main import fc_wrapper
        ret = fc_wrapper(bam, r["barcodes"], r["features"], out, **r["kwargs"])
        files = RDR_FILES
    else:
        from xcltk_b200.baf.fc.main import afc_wrapper
        ret = afc_wrapper(bam, r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
        files = BAF_FILES
    assert ret == int(read(r["expected"] + "/RETCODE")) == 0
    compare_dirs(r["expected"], out, files)
    assert used and all(used), "the device decoder declined the htslib-written BAM"


@pytest.mark.parametrize("n_shards", [2, 5])
@pytest.mark.parametrize("case,run", [("c1_chr22_10x", "rdr_defaults"), ("c1_chr22_10x", "rdr_umi_none"),
                                      ("c1_chr22_10x", "baf_all_reg_dup"), ("c1_chr22_10x", "baf_count3_maf0.1")])
def test_one_library_split_by_byte_ranges_matches_reference(case, run, n_shards, tmp_path, gpu_ctx, monkeypatch):
    """The multi-GPU product path on one GPU (XCLTK_B200_EMULATE_SHARDS: the shards run one after the other on
    device 0): the BAM is cut at BGZF block boundaries, every shard decodes only the blocks of its genomic chunk
    (+ halo) and counts the features / regions that start in it; the merged rows are the reference's bytes."""
    from xcltk_b200 import engine, synth
    monkeypatch.setenv("XCLTK_B200_GPUS", str(n_shards))
    monkeypatch.setenv("XCLTK_B200_EMULATE_SHARDS", "1")
    r = resolve(case, run)
    sams = []
    for k, p in enumerate(r["sam"]):
        q = str(tmp_path / ("%d_%s" % (k, os.path.basename(p))))
        synth.reblock_bam(p, q)                         # htslib's block layout (the goldens' BAMs are cut blindly)
        sams.append(q)
    seen = []
    real = engine.load_reads_sharded

    def spy(*a, **kw):
        b = real(*a, **kw)
        seen.append(b)
        return b
    monkeypatch.setattr(engine, "load_reads_sharded", spy)
    out = str(tmp_path / "out")
    if r["kind"] == "basefc":
        from xcltk_b200.rdr.fc.main import fc_wrapper
        ret = fc_wrapper(",".join(sams), r["barcodes"], r["features"], out, **r["kwargs"])
        files = RDR_FILES
    else:
        from xcltk_b200.baf.fc.main import afc_wrapper
        ret = afc_wrapper(",".join(sams), r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
        files = BAF_FILES
    assert ret == 0
    compare_dirs(r["expected"], out, files)
    assert seen and seen[0] is not None, "the library was not split"
    st = seen[0].stats
    whole = os.path.getsize(sams[0])
    decoded = sum(hi - lo for per in st["byte_ranges"] for lo, hi in per)
    assert decoded < 1.6 * whole          # every shard decodes its part (+ halo), not the whole file
    assert sum(len(s) for s in seen[0].shards) > 0
