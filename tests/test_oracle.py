"""Pins the CPU oracle (oracle/xg_oracle.c) and the host-side logic against the unmodified
reference: every golden run is reproduced byte for byte with the oracle as the counting
backend.  No GPU needed."""

import logging

import pytest

import oracle_backend
from util import BAF_FILES, RDR_FILES, compare_dirs, golden_runs, read, resolve

logging.disable(logging.CRITICAL)


@pytest.mark.parametrize("case,run", golden_runs("basefc"))
def test_oracle_basefc_matches_reference(case, run, tmp_path, monkeypatch):
    from xcltk_b200.rdr.fc import main as rdr_main
    monkeypatch.setattr(rdr_main, "count_features", oracle_backend.oracle_count_features)
    r = resolve(case, run)
    out = str(tmp_path / "out")
    ret = rdr_main.fc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], out, **r["kwargs"])
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, RDR_FILES)


@pytest.mark.parametrize("case,run", golden_runs("baf"))
def test_oracle_baf_matches_reference(case, run, tmp_path, monkeypatch):
    from xcltk_b200.baf.fc import main as baf_main
    monkeypatch.setattr(baf_main, "count_regions", oracle_backend.oracle_count_regions)
    r = resolve(case, run)
    out = str(tmp_path / "out")
    ret = baf_main.afc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, BAF_FILES)
