"""Genomic region helpers (semantics of xcltk/utils/grange.py:8-63,263-264)."""


def format_chrom(chrom):
    """Strip a leading 'chr' (case-insensitive test, as grange.py:263-264)."""
    return chrom[3:] if chrom.lower().startswith("chr") else chrom


class Region(object):
    """1-based start (inclusive) / end (exclusive); `chrom` has 'chr' stripped."""

    def __init__(self, chrom, start, end, rid=None):
        self.chrom = format_chrom(chrom)
        self.start = start
        self.end = end
        self._rid = rid
        self.len = self.end - self.start

    def get_id(self):
        if self._rid is None:
            self._rid = "%s_%d_%d" % (self.chrom, self.start, self.end)
        return self._rid

    def get_len(self):
        return self.len
