"""pysam-surface shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never on the product path).

pysam/htslib are not installable in this image (no network), so the reference's
unmodified modules (imported from /root/reference) are run on top of this stand-in.
It restates, in pure Python, only the slice of the pysam API that the reference's
two counting paths touch (call sites: xcltk/rdr/fc/core.py:47-60,75,
xcltk/rdr/fc/mcount.py:38-40,120, xcltk/baf/fc/core.py:19-32,48,
xcltk/baf/fc/mcount.py:54-56,113-115,224, xcltk/utils/sam.py:21-27,106,114):

    AlignmentFile(fn, "r").fetch(contig, start, stop) / .close()
    AlignedSegment.{mapq, flag, has_tag, get_tag, positions, cigartuples,
                    query_sequence, query_name}

Semantics follow the SAM/BAM specification (SAMv1 section 4.2) and the htslib rules
listed in SURVEY.md Appendix A.3.  Parity with real pysam is UNPINNED (the reference
ships no tests); this shim and the C++ decoder in xcltk_b200/csrc are two independent
implementations that are cross-checked record by record in tests/.
"""

import bisect
import gzip
import struct

_SEQ_CODES = "=ACMGRSVTWYHKDBN"
_CONSUMES_REF = (0, 2, 3, 7, 8)       # M D N = X
_ALIGNED = (0, 7, 8)                  # M = X
BAM_FUNMAP = 4


class AlignedSegment(object):
    __slots__ = ("tid", "pos", "mapq", "flag", "query_name", "_cigar", "_l_seq",
                 "_seq_raw", "_aux_raw", "_tags", "_positions", "_endpos")

    def __init__(self, tid, pos, mapq, flag, name, cigar, l_seq, seq_raw, aux_raw):
        self.tid = tid
        self.pos = pos
        self.mapq = mapq
        self.flag = flag
        self.query_name = name
        self._cigar = cigar
        self._l_seq = l_seq
        self._seq_raw = seq_raw
        self._aux_raw = aux_raw
        self._tags = None
        self._positions = None
        self._endpos = None

    # htslib bam_endpos(): pos + rlen, rlen = sum of ref-consuming ops, 0 for
    # FUNMAP records, and a zero rlen counts as 1.
    @property
    def endpos(self):
        if self._endpos is None:
            rlen = 0
            if not (self.flag & BAM_FUNMAP):
                for w in self._cigar:
                    if (w & 15) in _CONSUMES_REF:
                        rlen += w >> 4
            if rlen == 0:
                rlen = 1
            self._endpos = self.pos + rlen
        return self._endpos

    @property
    def cigartuples(self):
        if not self._cigar:
            return None
        return [(w & 15, w >> 4) for w in self._cigar]

    # pysam get_reference_positions(): reference positions of M/=/X bases only.
    @property
    def positions(self):
        if self._positions is None:
            out = []
            p = self.pos
            for w in self._cigar:
                op, l = w & 15, w >> 4
                if op in _ALIGNED:
                    out.extend(range(p, p + l))
                    p += l
                elif op == 2 or op == 3:
                    p += l
            self._positions = out
        return self._positions

    @property
    def query_sequence(self):
        if self._l_seq == 0:
            return None
        raw = self._seq_raw
        chars = []
        for i in range(self._l_seq):
            b = raw[i >> 1]
            chars.append(_SEQ_CODES[(b >> 4) if (i & 1) == 0 else (b & 15)])
        return "".join(chars)

    def _parse_tags(self):
        tags = {}
        raw = self._aux_raw
        i, n = 0, len(raw)
        while i + 3 <= n:
            tag = raw[i:i + 2].decode("ascii", "replace")
            typ = chr(raw[i + 2])
            i += 3
            if typ == "A":
                val = chr(raw[i]); i += 1
            elif typ == "c":
                val = struct.unpack_from("<b", raw, i)[0]; i += 1
            elif typ == "C":
                val = raw[i]; i += 1
            elif typ == "s":
                val = struct.unpack_from("<h", raw, i)[0]; i += 2
            elif typ == "S":
                val = struct.unpack_from("<H", raw, i)[0]; i += 2
            elif typ == "i":
                val = struct.unpack_from("<i", raw, i)[0]; i += 4
            elif typ == "I":
                val = struct.unpack_from("<I", raw, i)[0]; i += 4
            elif typ == "f":
                val = struct.unpack_from("<f", raw, i)[0]; i += 4
            elif typ in "ZH":
                j = raw.index(b"\0", i)
                val = raw[i:j].decode("ascii", "replace"); i = j + 1
            elif typ == "B":
                sub = chr(raw[i]); cnt = struct.unpack_from("<I", raw, i + 1)[0]
                i += 5
                fmt, size = {"c": ("b", 1), "C": ("B", 1), "s": ("h", 2), "S": ("H", 2),
                             "i": ("i", 4), "I": ("I", 4), "f": ("f", 4)}[sub]
                val = list(struct.unpack_from("<%d%s" % (cnt, fmt), raw, i))
                i += cnt * size
            else:
                raise ValueError("unknown aux type '%s'" % typ)
            if tag not in tags:           # htslib bam_aux_get returns the first match
                tags[tag] = val
        self._tags = tags

    def has_tag(self, tag):
        if self._tags is None:
            self._parse_tags()
        return tag in self._tags

    def get_tag(self, tag):
        if self._tags is None:
            self._parse_tags()
        if tag not in self._tags:
            raise KeyError("tag '%s' not present" % tag)
        return self._tags[tag]


def read_bam(fn):
    """Return (header_text, [(name, length)], [AlignedSegment in file order])."""
    with gzip.open(fn, "rb") as fp:       # BGZF = concatenated gzip members
        data = fp.read()
    if data[:4] != b"BAM\1":
        raise ValueError("not a BAM file: %s" % fn)
    l_text = struct.unpack_from("<i", data, 4)[0]
    off = 8
    text = data[off:off + l_text].decode("ascii", "replace")
    off += l_text
    n_ref = struct.unpack_from("<i", data, off)[0]
    off += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, off)[0]
        off += 4
        name = data[off:off + l_name - 1].decode("ascii")
        off += l_name
        l_ref = struct.unpack_from("<i", data, off)[0]
        off += 4
        refs.append((name, l_ref))
    recs = []
    n = len(data)
    unpack_core = struct.Struct("<iiiBBHHHIiii").unpack_from
    while off + 4 <= n:
        (block_size, tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq,
         _ntid, _npos, _tlen) = unpack_core(data, off)
        p = off + 36
        name = data[p:p + l_name - 1].decode("ascii", "replace")
        p += l_name
        cigar = struct.unpack_from("<%dI" % n_cig, data, p) if n_cig else ()
        p += 4 * n_cig
        seq_raw = data[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2 + l_seq
        aux_raw = data[p:off + 4 + block_size]
        recs.append(AlignedSegment(tid, pos, mapq, flag, name, cigar, l_seq,
                                   seq_raw, aux_raw))
        off += 4 + block_size
    return text, refs, recs


class AlignmentFile(object):
    """Loads the whole BAM; `fetch` answers from per-contig position arrays.

    Real pysam needs a .bai/.csi index for fetch; the shim scans instead, so the
    result is what a valid index would give (all records with
    tid match, pos < stop and bam_endpos > start, in file order; A.3)."""

    def __init__(self, fn, mode="r", **kwargs):
        self.filename = fn
        self.text, refs, recs = read_bam(fn)
        self.references = tuple(r[0] for r in refs)
        self.lengths = tuple(r[1] for r in refs)
        self._tid = {}
        for i, name in enumerate(self.references):
            if name not in self._tid:
                self._tid[name] = i
        self._recs = {}
        for r in recs:
            if r.tid >= 0:
                self._recs.setdefault(r.tid, []).append(r)
        self._pos = {}
        self._maxspan = {}
        for tid, lst in self._recs.items():
            last = -1
            for r in lst:
                if r.pos < last:
                    raise ValueError("BAM is not coordinate sorted: %s" % fn)
                last = r.pos
            self._pos[tid] = [r.pos for r in lst]
            self._maxspan[tid] = max(r.endpos - r.pos for r in lst)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def get_tid(self, name):
        return self._tid.get(name, -1)

    def fetch(self, contig=None, start=None, stop=None, **kwargs):
        if contig is None:
            raise ValueError("shim supports region fetch only")
        if contig not in self._tid:
            raise ValueError("invalid contig `%s`" % contig)
        tid = self._tid[contig]
        if start is None:
            start = 0
        if stop is None:
            stop = 1 << 29
        if start < 0:
            raise ValueError("start out of range (%i)" % start)
        if stop < 0:
            raise ValueError("stop out of range (%i)" % stop)
        if start > stop:
            raise ValueError("invalid coordinates: start (%i) > stop (%i)" % (start, stop))
        return self._iter(tid, start, stop)

    def _iter(self, tid, start, stop):
        lst = self._recs.get(tid)
        if not lst:
            return
        pos = self._pos[tid]
        i = bisect.bisect_left(pos, start - self._maxspan[tid])
        j = bisect.bisect_left(pos, stop)
        for k in range(i, j):
            r = lst[k]
            if r.endpos > start:
                yield r


# Names the reference imports at module import time but never calls on the two
# counting paths (utils/zfile.py:53 write-bgzip only; baf/fixref.py; utils/vcf.py).
class BGZFile(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: BGZFile is not available")


class FastaFile(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: FastaFile is not available")


class VariantFile(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: VariantFile is not available")


def tabix_index(*a, **k):
    raise NotImplementedError("shim: tabix_index is not available")


__version__ = "0.0-shim"
