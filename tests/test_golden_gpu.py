"""GPU parity: the public Python entry points (fc_wrapper / afc_wrapper, which call the CUDA
path through the C-ABI) must reproduce, byte for byte, the files the unmodified reference
wrote for the same inputs (tests/golden/*/expected, made by oracle/make_golden.py)."""

import logging

import pytest

from util import BAF_FILES, RDR_FILES, compare_dirs, golden_runs, read, resolve

pytestmark = pytest.mark.gpu
logging.disable(logging.CRITICAL)


@pytest.mark.parametrize("case,run", golden_runs("basefc"))
def test_basefc_matches_reference(case, run, tmp_path, gpu_ctx):
    from xcltk_b200.rdr.fc.main import fc_wrapper
    r = resolve(case, run)
    out = str(tmp_path / "out")
    ret = fc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], out, **r["kwargs"])
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, RDR_FILES)


@pytest.mark.parametrize("case,run", golden_runs("baf"))
def test_baf_matches_reference(case, run, tmp_path, gpu_ctx):
    from xcltk_b200.baf.fc.main import afc_wrapper
    r = resolve(case, run)
    out = str(tmp_path / "out")
    ret = afc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], r["snps"], out, **r["kwargs"])
    assert ret == int(read(r["expected"] + "/RETCODE"))
    compare_dirs(r["expected"], out, BAF_FILES)
