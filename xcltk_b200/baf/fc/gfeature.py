"""Phased SNPs and block regions (data carriers of xcltk/baf/fc/gfeature.py:8-72) plus a
sorted-array SNP set that answers the region join without an interval tree."""

import bisect

from ...utils.grange import Region, format_chrom


class SNP(Region):
    """chrom ('chr' stripped), 1-based pos, ref/alt base, haplotype index of each allele."""

    def __init__(self, chrom, pos, ref, alt, ref_idx, alt_idx):
        super().__init__(chrom, pos, pos + 1)
        self.pos = pos
        self.ref = ref
        self.alt = alt
        self.ref_idx = ref_idx
        self.alt_idx = alt_idx
        self.gt = {ref: ref_idx, alt: alt_idx}
        self.index = -1          # position in SNPSet.snps (device SNP id)

    def get_id(self):
        return "%s_%d" % (self.chrom, self.pos)

    def get_region_allele_index(self, base):
        return self.gt[base] if base in self.gt else -1


class BlockRegion(Region):
    def __init__(self, chrom, start, end, name=None, snp_list=None):
        super().__init__(chrom, start, end)
        self.name = name
        self.snp_list = snp_list


class SNPSet(object):
    """All loaded SNPs (duplicates kept, as RegionSet(is_uniq=False), utils/grange.py:104-138).
    `fetch` returns the SNPs with start <= pos < end on a contig, sorted by pos
    (== sorted(RegionSet.fetch(...), key=pos), baf/fc/main.py:91-98)."""

    def __init__(self):
        self.snps = []
        self._by_chrom = None

    def add(self, snp):
        snp.index = len(self.snps)
        self.snps.append(snp)
        self._by_chrom = None
        return 0

    def get_n(self):
        return len(self.snps)

    def _index(self):
        if self._by_chrom is None:
            d = {}
            for s in self.snps:
                d.setdefault(s.chrom, []).append(s)
            self._by_chrom = {}
            for c, lst in d.items():
                lst.sort(key=lambda s: s.pos)          # stable: file order among equal pos
                self._by_chrom[c] = (lst, [s.pos for s in lst])
        return self._by_chrom

    def fetch(self, chrom, start, end):
        idx = self._index().get(format_chrom(chrom))
        if idx is None or start >= end:
            return []
        lst, pos = idx
        return lst[bisect.bisect_left(pos, start):bisect.bisect_left(pos, end)]

    def destroy(self):
        self.snps = []
        self._by_chrom = None
