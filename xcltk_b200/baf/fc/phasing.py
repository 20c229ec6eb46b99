"""Region-wise local phasing before the counting (xcltk/baf/fc/phasing.py:13-78)."""

from logging import warning as warn

import numpy as np

from ..localphase import snp_local_phasing


def reg_local_phasing(reg, AD, DP):
    """AD, DP: cell x SNP count arrays of the cellsnp-lite SNPs inside `reg`, columns in the order of
    reg.snp_list.  SNPs nobody expresses leave the region's list; the SNPs the EM flips swap their haplotype
    indices (the SNP objects are shared by every region that holds them).  Returns (reg, flip) with flip = None
    when the phasing could not be done."""
    cell_idx = DP.sum(axis=1) > 0
    snp_idx = DP.sum(axis=0) > 0
    AD, DP = AD[np.ix_(cell_idx, snp_idx)], DP[np.ix_(cell_idx, snp_idx)]
    reg.snp_list = [s for keep, s in zip(snp_idx, reg.snp_list) if keep]
    BD = DP - AD
    flip = np.array([snp.ref_idx == 1 for snp in reg.snp_list])
    AD_ref_phased = AD * (1 - flip.T) + BD * flip.T          # ALT counts -> counts of haplotype 1
    flip = snp_local_phasing(AD_ref_phased, DP, np.array([s.pos for s in reg.snp_list]))
    if flip is None:
        warn("local phasing for region '%s' failed!" % reg.name)
        return reg, None
    flip = (1 - flip if np.mean(flip) > 0.5 else flip + 0).astype(int)     # as few changes as possible
    assert len(reg.snp_list) == len(flip)
    for snp, f in zip(reg.snp_list, flip):
        if f == 1:
            snp.ref_idx, snp.alt_idx = 1 - snp.ref_idx, 1 - snp.alt_idx
            snp.gt = {snp.ref: snp.ref_idx, snp.alt: snp.alt_idx}
    return reg, flip
