"""intervaltree-surface shim -- ORACLE / TEST INFRASTRUCTURE ONLY.

The reference uses intervaltree only through xcltk/utils/grange.py:120,135,164:
`tree[begin:end] = data`, `tree[begin:end]` -> set of Interval(begin, end, data),
`tree.clear()`.  Published semantics of intervaltree 3.x restated: half-open
intervals, null intervals (begin >= end) are rejected on insert, a query with
begin >= end returns the empty set, overlap iff iv.begin < end and iv.end > begin.
"""

from collections import namedtuple

Interval = namedtuple("Interval", ["begin", "end", "data"])


class IntervalTree(object):
    def __init__(self):
        self._ivs = []
        self._sorted = True

    def __setitem__(self, index, value):
        begin, end = index.start, index.stop
        if begin >= end:
            raise ValueError("IntervalTree: Null Interval objects not allowed in "
                             "IntervalTree: Interval(%r, %r, %r)" % (begin, end, value))
        self._ivs.append(Interval(begin, end, value))
        self._sorted = False

    def __getitem__(self, index):
        if not isinstance(index, slice):
            begin, end = index, index + 1
        else:
            begin, end = index.start, index.stop
        if not self._ivs or begin >= end:
            return set()
        if not self._sorted:
            self._ivs.sort(key=lambda iv: (iv.begin, iv.end))
            self._begins = [iv.begin for iv in self._ivs]
            self._maxlen = max(iv.end - iv.begin for iv in self._ivs)
            self._sorted = True
        import bisect
        lo = bisect.bisect_left(self._begins, begin - self._maxlen)
        hi = bisect.bisect_left(self._begins, end)
        out = []
        for k in range(lo, hi):
            iv = self._ivs[k]
            if iv.end > begin:
                out.append(iv)
        return _IvSet(out)

    def clear(self):
        self._ivs = []
        self._sorted = True

    def __len__(self):
        return len(self._ivs)


class _IvSet(list):
    """Unordered result container (Interval.data may be unhashable-by-value objects;
    the reference only iterates over it and takes len())."""
    pass
