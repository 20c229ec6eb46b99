"""GPU parity: the public Python entry points (fc_wrapper / afc_wrapper, which call the CUDA
path through the C-ABI) must reproduce, byte for byte, the files the unmodified reference
wrote for the same inputs (tests/golden/*/expected, made by oracle/make_golden.py)."""

import logging

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
def test_region_sharded_over_gpus_matches_reference(case, run, tmp_path, monkeypatch):
    """Product path sharded by genomic chunks over all GPUs of the box (>= 2), rows merged on
    the host: still byte-identical to the reference."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("XCLTK_B200_GPUS", str(min(n, 4)))
    r = resolve(case, run)
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
