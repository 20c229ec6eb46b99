"""Contig-name resolution with the reference's 'chr'-prefix tolerance.

`sam_fetch` (xcltk/utils/sam.py:85-118) tries `sam.fetch(chrom, ...)` and, when that raises
(unknown contig), the name with the 'chr' prefix toggled.  Here the same rule maps a
(stripped) feature / SNP contig name to the tid of one BAM, once per BAM, instead of once
per fetch.
"""

BAM_FPAIRED = 1
BAM_FPROPER_PAIR = 2


def resolve_tid(ref_index, chrom):
    """ref_index: dict contig name -> tid (first occurrence). Returns tid or -1."""
    tid = ref_index.get(chrom)
    if tid is not None:
        return tid
    alt = chrom[3:] if chrom.startswith("chr") else "chr" + chrom
    tid = ref_index.get(alt)
    return -1 if tid is None else tid


def build_tid_maps(bam_refs, chroms):
    """bam_refs: per BAM list of (name, len); chroms: distinct contig names used by the
    features / SNPs.  Returns (gid_of: dict chrom -> gid, tid_maps: per BAM int list
    tid -> gid or -1).  Raises ValueError when two names resolve to one contig of a BAM."""
    gid_of = {c: i for i, c in enumerate(chroms)}
    tid_maps = []
    for refs in bam_refs:
        index = {}
        for tid, (name, _len) in enumerate(refs):
            index.setdefault(name, tid)
        m = [-1] * len(refs)
        for c, g in gid_of.items():
            tid = resolve_tid(index, c)
            if tid >= 0:
                if m[tid] >= 0 and m[tid] != g:
                    raise ValueError("contig names '%s' and '%s' resolve to the same BAM contig '%s'"
                                     % (chroms[m[tid]], c, refs[tid][0]))
                m[tid] = g
        tid_maps.append(m)
    return gid_of, tid_maps
