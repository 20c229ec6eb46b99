"""Reader of a cellsnp-lite output directory (the input of the local-phasing pre-step of `xcltk baf`).

Replaces xcltk/utils/csp_io.py:16-63 (load_data) for the one thing afc_core needs from it
(xcltk/baf/fc/main.py:419-454, 112-149): cells, SNP coordinates and the AD / DP count matrices.  No AnnData:
the matrices stay sparse (cell x SNP, CSC) and are sliced per region.
"""

import gzip
import os

import numpy as np
import scipy.io
import scipy.sparse


class CellSNPData(object):
    """cells: names in file order; chrom / pos: one per SNP (VCF order); AD, DP: cell x SNP sparse counts."""

    def __init__(self, cells, chrom, pos, AD, DP):
        self.cells = np.asarray(cells, dtype=object)
        self.chrom = np.asarray(chrom, dtype=object)
        self.pos = np.asarray(pos, dtype=np.int64)
        self.AD = scipy.sparse.csc_matrix(AD)
        self.DP = scipy.sparse.csc_matrix(DP)
        assert self.AD.shape == self.DP.shape == (len(self.cells), len(self.pos))

    @property
    def shape(self):
        return self.AD.shape

    def subset_snps(self, idx):
        idx = np.asarray(idx)
        return CellSNPData(self.cells, self.chrom[idx], self.pos[idx], self.AD[:, idx], self.DP[:, idx])

    def subset_cells(self, mask):
        mask = np.asarray(mask, dtype=bool)
        return CellSNPData(self.cells[mask], self.chrom, self.pos, self.AD[mask, :], self.DP[mask, :])


def _load_vcf_sites(fn):
    op = gzip.open if fn.lower().endswith(".gz") else open
    chrom, pos = [], []
    seen_header = False
    with op(fn, "rt") as fp:
        for line in fp:
            if line.startswith("#"):
                seen_header = seen_header or line.startswith("#CHROM")
                continue
            if not line.strip():
                continue
            f = line.split("\t", 2)
            chrom.append(f[0])
            pos.append(int(f[1]))
    if not seen_header:
        raise IOError("no #CHROM line in '%s'" % fn)
    return chrom, pos


def load_data(data_dir, is_gzip=True):
    """cellSNP.base.vcf(.gz), cellSNP.samples.tsv, cellSNP.tag.{AD,DP}.mtx of `data_dir` (csp_io.py:16-63)."""
    chrom, pos = _load_vcf_sites(os.path.join(data_dir, "cellSNP.base.vcf" + (".gz" if is_gzip else "")))
    with open(os.path.join(data_dir, "cellSNP.samples.tsv")) as fp:
        cells = [x.rstrip("\n") for x in fp if x.rstrip("\n")]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)       # scipy's sparse-array migration notice
        ad = scipy.io.mmread(os.path.join(data_dir, "cellSNP.tag.AD.mtx"))        # SNP x cell on disk
        dp = scipy.io.mmread(os.path.join(data_dir, "cellSNP.tag.DP.mtx"))
    return CellSNPData(cells, chrom, pos, scipy.sparse.csc_matrix(ad.T), scipy.sparse.csc_matrix(dp.T))
