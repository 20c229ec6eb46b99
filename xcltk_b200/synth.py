"""Synthetic 10x-style / SMART-seq-style BAM generation (host, test + bench scale).

Nothing like this exists in the reference (it ships one real fixture BAM and no
tests); the distributions follow SURVEY.md section 8(d) C1/C2/C4.
Records are plain tuples so that the same list can be written as a BAM
(`write_bam`) and compared field-by-field with what the decoders return.

Record tuple: (name, flag, tid, pos0, mapq, cigar[(op,len)...], seq, tags[(tag,type,val)...])
"""

import random
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"
_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(payload, level=6):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = co.compress(payload) + co.flush()
    bsize = len(cdata) + 25                      # total block size - 1
    if bsize > 65535:
        raise ValueError("BGZF block too large")
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    tail = struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload))
    return head + cdata + tail


def reg2bin(beg, end):
    """SAMv1 section 5.3 bin for [beg, end)."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def pack_record(rec):
    name, flag, tid, pos, mapq, cigar, seq, tags = rec
    bname = name.encode("ascii") + b"\0"
    rlen = sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))
    end = pos + (rlen if rlen > 0 and not (flag & 4) else 1)
    l_seq = len(seq)
    packed_seq = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        code = _SEQ_CODE.get(c.upper(), 15)
        if i & 1:
            packed_seq[i >> 1] |= code
        else:
            packed_seq[i >> 1] |= code << 4
    aux = bytearray()
    for tag, typ, val in tags:
        aux += tag.encode("ascii") + typ.encode("ascii")
        if typ == "Z" or typ == "H":
            aux += val.encode("ascii") + b"\0"
        elif typ == "A":
            aux += val.encode("ascii")[:1]
        elif typ == "i":
            aux += struct.pack("<i", val)
        elif typ == "I":
            aux += struct.pack("<I", val)
        elif typ == "c":
            aux += struct.pack("<b", val)
        elif typ == "C":
            aux += struct.pack("<B", val)
        elif typ == "s":
            aux += struct.pack("<h", val)
        elif typ == "S":
            aux += struct.pack("<H", val)
        elif typ == "f":
            aux += struct.pack("<f", val)
        elif typ == "B":                 # val = (subtype, [values])
            sub, arr = val
            fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
            aux += sub.encode("ascii") + struct.pack("<I", len(arr))
            aux += struct.pack("<%d%s" % (len(arr), fmt), *arr)
        else:
            raise ValueError("unsupported aux type %r" % typ)
    body = struct.pack("<iiBBHHHIiii", tid, pos, len(bname), mapq,
                       reg2bin(pos, end) if tid >= 0 else 4680, len(cigar), flag, l_seq,
                       -1, -1, 0)
    body += bname
    body += b"".join(struct.pack("<I", (l << 4) | op) for op, l in cigar)
    body += bytes(packed_seq) + b"\xff" * l_seq + bytes(aux)
    return struct.pack("<i", len(body)) + body


BGZF_MAX_PAYLOAD = 0xff00      # htslib's BGZF_BLOCK_SIZE


def write_bam(path, refs, records, header_text=None, block=60000, level=6, align=True):
    """Write a coordinate-sorted BAM (no index; the decoders scan).

    align=True lays the blocks out as htslib does (bam_hdr_write ends with a flush; bam_write1
    flushes before a record that would not fit, so a record crosses a block boundary only when it
    is larger than a block).  align=False cuts the stream every `block` bytes regardless of the
    records -- legal BGZF/BAM that no htslib writer produces; the device decoder must refuse it."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(
            "@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in refs)
    text = header_text.encode("ascii")
    buf = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for n, l in refs:
        bn = n.encode("ascii") + b"\0"
        buf += struct.pack("<i", len(bn)) + bn + struct.pack("<i", l)
    with open(path, "wb") as fp:
        if align:
            def flush():
                while buf:
                    fp.write(_bgzf_block(bytes(buf[:BGZF_MAX_PAYLOAD]), level))
                    del buf[:BGZF_MAX_PAYLOAD]
            flush()
            for rec in records:
                r = pack_record(rec)
                if len(buf) + len(r) > BGZF_MAX_PAYLOAD:
                    flush()
                buf += r
            flush()
            fp.write(BGZF_EOF)
            return
        for rec in records:
            buf += pack_record(rec)
            while len(buf) >= block:
                fp.write(_bgzf_block(bytes(buf[:block]), level))
                del buf[:block]
        if buf:
            fp.write(_bgzf_block(bytes(buf), level))
        fp.write(BGZF_EOF)


def reblock_bam(src, dst, level=6):
    """Rewrite a BAM with htslib's block layout (header flushed, whole records per block);
    the uncompressed stream is unchanged."""
    import gzip
    with gzip.open(src, "rb") as fp:
        u = fp.read()
    o = 8 + struct.unpack_from("<i", u, 4)[0]
    n_ref = struct.unpack_from("<i", u, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 8 + struct.unpack_from("<i", u, o)[0]
    with open(dst, "wb") as fp:
        for i in range(0, o, BGZF_MAX_PAYLOAD):
            fp.write(_bgzf_block(u[i:min(o, i + BGZF_MAX_PAYLOAD)], level))
        beg = q = o
        while q < len(u):
            nxt = q + 4 + struct.unpack_from("<i", u, q)[0]
            if nxt - beg > BGZF_MAX_PAYLOAD and q > beg:
                fp.write(_bgzf_block(u[beg:q], level))
                beg = q
            q = nxt
        for i in range(beg, len(u), BGZF_MAX_PAYLOAD):       # an oversized record is cut, as htslib does
            fp.write(_bgzf_block(u[i:min(len(u), i + BGZF_MAX_PAYLOAD)], level))
        fp.write(BGZF_EOF)


def make_barcodes(rng, n, suffix="-1", length=16):
    out = set()
    while len(out) < n:
        out.add("".join(rng.choice("ACGT") for _ in range(length)) + suffix)
    return sorted(out)


def load_features(fn, chroms=None):
    """(chrom_stripped, start1, end1_incl, name) rows of a header-less TSV."""
    import gzip
    op = gzip.open if fn.lower().endswith((".gz", ".gzip")) else open
    out = []
    with op(fn, "rt") as fp:
        for line in fp:
            p = line.rstrip().split("\t")
            c = p[0][3:] if p[0].lower().startswith("chr") else p[0]
            if chroms is None or c in chroms:
                out.append((c, int(p[1]), int(p[2]), p[3]))
    return out


def synth_features(rng, contigs, n, mean_len=30000, nested_frac=0.3):
    """Random gene-like features incl. nested/overlapping ones; chrom names without 'chr'."""
    feats = []
    names = [c for c, _ in contigs]
    lens = dict(contigs)
    for i in range(n):
        c = rng.choice(names)
        if feats and rng.random() < nested_frac:
            pc, ps, pe, _ = rng.choice(feats)
            c = pc
            s = rng.randint(max(1, ps - 200), pe)
            e = min(lens[c], s + max(8, int(rng.expovariate(1.0 / (mean_len / 4)))))
        else:
            s = rng.randint(1, lens[c] - 10)
            e = min(lens[c], s + max(8, int(rng.expovariate(1.0 / mean_len))))
        feats.append((c, s, e, "g%d" % i))
    return feats


def _rand_cigar(rng, L, mix):
    r = rng.random()
    if r < mix[0]:
        return [(0, L)]
    if r < mix[0] + mix[1]:
        a = rng.randint(5, L - 5)
        return [(0, a), (3, rng.randint(80, 5000)), (0, L - a)]
    if r < mix[0] + mix[1] + mix[2]:
        s = rng.randint(1, L // 2)
        if rng.random() < 0.5:
            return [(4, s), (0, L - s)]
        return [(0, L - s), (4, s)]
    a = rng.randint(5, L - 10)
    return [(0, a), (2, 2), (0, L - a - 3), (1, 3)]


def gen_10x_records(seed, contigs, feats, n_reads, barcodes, read_len=91,
                    snps=None, chr_prefix="", umi_len=12, cigar_mix=(0.78, 0.15, 0.04),
                    reads_per_umi=3.0, paired=False, with_tags=True, name_prefix="r"):
    """10x-style reads placed inside feature spans.

    contigs: [(name_without_prefix, length)]; BAM contig name = chr_prefix + name.
    feats: (chrom, start1, end1, name) rows; snps: optional dict chrom -> sorted list of
    (pos1, ref, alt, ref_hap) so that molecules carry haplotype-consistent bases.
    Returns (refs, records sorted by (tid, pos)).
    """
    rng = random.Random(seed)
    tid_of = {c: i for i, (c, _) in enumerate(contigs)}
    clen = dict(contigs)
    feats = [f for f in feats if f[0] in tid_of]
    if not feats:
        raise ValueError("no features on the given contigs")
    w = [max(1, f[2] - f[1] + 1) for f in feats]
    import bisect
    snp_pos = {c: [s[0] for s in lst] for c, lst in (snps or {}).items()}
    recs = []
    other_cb = make_barcodes(rng, max(1, len(barcodes) // 10), suffix="-1")
    mol = 0
    f_cache = []
    while len(recs) < n_reads:
        mol += 1
        if not f_cache:
            f_cache = rng.choices(feats, weights=w, k=256)
        c, fs, fe, _ = f_cache.pop()
        anchor = rng.randint(max(1, fs - 50), min(clen[c] - 6000 - read_len, fe + 50))
        if anchor < 1:
            continue
        cb = rng.choice(barcodes) if rng.random() >= 0.05 else rng.choice(other_cb)
        umi = "".join(rng.choice("ACGT") for _ in range(umi_len))
        hap = rng.randrange(2)
        k = 1 + int(rng.expovariate(1.0 / max(1e-9, reads_per_umi - 1.0))) if reads_per_umi > 1 else 1
        for j in range(k):
            pos1 = max(1, anchor + rng.randint(-150, 150))
            cigar = _rand_cigar(rng, read_len, cigar_mix)
            r = rng.random()
            flag = 16 if rng.random() < 0.5 else 0
            if paired:
                flag = rng.choice((99, 147, 83, 163))
            if r < 0.04:
                flag |= 256
            elif r < 0.09:
                flag |= 1024
            elif r < 0.10:
                flag |= 2048
            mapq = 255 if rng.random() < 0.85 else rng.choice((0, 1, 3))
            seq = [rng.choice("ACGT") for _ in range(read_len)]
            if snps and c in snps:           # haplotype-consistent bases at covered SNPs
                p = pos1 - 1
                q = 0
                for op, l in cigar:
                    if op in (0, 7, 8):
                        lo = bisect.bisect_left(snp_pos[c], p + 1)
                        hi = bisect.bisect_left(snp_pos[c], p + l + 1)
                        for si in range(lo, hi):
                            spos, ref, alt, ref_hap = snps[c][si]
                            rr = rng.random()
                            if rr < 0.02:
                                b = rng.choice("ACGTN")
                            else:
                                h = hap if rr >= 0.03 else 1 - hap
                                b = ref if ref_hap == h else alt
                            seq[q + (spos - 1 - p)] = b
                        p += l
                        q += l
                    elif op in (2, 3):
                        p += l
                    elif op in (1, 4):
                        q += l
            tags = [("NH", "C", 1)]
            if with_tags:
                rr = rng.random()
                if rr >= 0.01:
                    tags.append(("CB", "Z", cb))
                if not 0.01 <= rr < 0.02:
                    tags.append(("UB", "Z", umi if rr >= 0.025 else ""))
            name = "%s%07d" % (name_prefix, len(recs)) if not paired else "%s%07d" % (name_prefix, mol * 4 + (j >> 1))
            recs.append((name, flag, tid_of[c], pos1 - 1, mapq, cigar, "".join(seq), tags))
            if len(recs) >= n_reads:
                break
    recs.sort(key=lambda r: (r[2], r[3]))
    refs = [(chr_prefix + c, l) for c, l in contigs]
    return refs, recs


def gen_snps(seed, feats, n, contigs=None):
    """Phased het SNPs inside feature spans: dict chrom -> sorted [(pos1, ref, alt, ref_hap)]."""
    rng = random.Random(seed)
    if contigs is not None:
        feats = [f for f in feats if f[0] in contigs]
    w = [max(1, f[2] - f[1] + 1) for f in feats]
    out = {}
    seen = set()
    while sum(len(v) for v in out.values()) < n:
        c, fs, fe, _ = rng.choices(feats, weights=w)[0]
        pos = rng.randint(max(1, fs), fe)
        if (c, pos) in seen:
            continue
        seen.add((c, pos))
        ref = rng.choice("ACGT")
        alt = rng.choice([b for b in "ACGT" if b != ref])
        out.setdefault(c, []).append((pos, ref, alt, rng.randrange(2)))
    for c in out:
        out[c].sort()
    return out


def write_snp_tsv(path, snps, chr_prefix=""):
    with open(path, "w") as fp:
        fp.write("chrom\tpos\tref\talt\tref_hap\talt_hap\n")
        for c in snps:
            for pos, ref, alt, rh in snps[c]:
                fp.write("%s%s\t%d\t%s\t%s\t%d\t%d\n" % (chr_prefix, c, pos, ref, alt, rh, 1 - rh))


def write_features(path, feats):
    with open(path, "w") as fp:
        for c, s, e, n in feats:
            fp.write("%s\t%d\t%d\t%s\n" % (c, s, e, n))


def write_lines(path, items):
    with open(path, "w") as fp:
        for x in items:
            fp.write("%s\n" % x)


def write_fast_bam(path, n_reads, contigs, n_cells=500, seed=7, read_len=91, level=4, threads=8):
    """Vectorised writer of a large coordinate-sorted 10x-style BAM (bench: decode throughput).

    Every record has the same layout (name of 10 chars, CIGAR `<read_len>M`, CB:Z 16 nt + "-1",
    UB:Z 12 nt, NH:C), so the whole uncompressed stream is built with numpy and compressed into
    BGZF blocks by a thread pool (zlib releases the GIL).  Returns the barcode list."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.RandomState(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    bc_codes = rng.randint(0, 4, size=(n_cells, 16))
    barcodes = ["".join("ACGT"[x] for x in row) + "-1" for row in bc_codes]
    l_name, l_seq = 11, read_len
    n_seq_b = (l_seq + 1) // 2
    aux_len = 4 + (3 + 19) + (3 + 13)
    body = 32 + l_name + 4 + n_seq_b + l_seq + aux_len
    rec_len = 4 + body
    tot_len = sum(l for _, l in contigs)
    per = [int(round(n_reads * l / float(tot_len))) for _, l in contigs]
    per[-1] += n_reads - sum(per)
    header_text = ("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs)).encode()
    head = bytearray(b"BAM\1" + struct.pack("<i", len(header_text)) + header_text + struct.pack("<i", len(contigs)))
    for nm, ln in contigs:
        b = nm.encode() + b"\0"
        head += struct.pack("<i", len(b)) + b + struct.pack("<i", ln)
    chunks = []
    serial = 0
    for tid, ((nm, ln), n) in enumerate(zip(contigs, per)):
        if n <= 0:
            continue
        rec = np.zeros((n, rec_len), dtype=np.uint8)
        pos = np.sort(rng.randint(0, max(1, ln - read_len - 1), size=n)).astype("<i4")

        def put(col, arr):
            rec[:, col:col + arr.shape[1]] = arr
        put(0, np.full(n, body, dtype="<i4").view(np.uint8).reshape(n, 4))
        put(4, np.full(n, tid, dtype="<i4").view(np.uint8).reshape(n, 4))
        put(8, pos.view(np.uint8).reshape(n, 4))
        rec[:, 12] = l_name
        rec[:, 13] = np.where(rng.rand(n) < 0.85, 255, 3)
        put(14, np.full(n, 4681, dtype="<u2").view(np.uint8).reshape(n, 2))
        put(16, np.full(n, 1, dtype="<u2").view(np.uint8).reshape(n, 2))
        flag = (np.where(rng.rand(n) < 0.5, 16, 0) | np.where(rng.rand(n) < 0.05, 1024, 0)).astype("<u2")
        put(18, flag.view(np.uint8).reshape(n, 2))
        put(20, np.full(n, l_seq, dtype="<u4").view(np.uint8).reshape(n, 4))
        put(24, np.full((n, 2), -1, dtype="<i4").view(np.uint8).reshape(n, 8))
        o = 36
        ids = np.arange(serial, serial + n)
        serial += n
        digits = (ids[:, None] // 10 ** np.arange(9, -1, -1)[None, :]) % 10
        put(o, (digits + 48).astype(np.uint8))
        o += l_name                                          # NUL already there
        put(o, np.full(n, (read_len << 4) | 0, dtype="<u4").view(np.uint8).reshape(n, 4))
        o += 4
        nib = 1 << rng.randint(0, 4, size=(n, 2 * n_seq_b)).astype(np.uint8)
        put(o, ((nib[:, 0::2] << 4) | nib[:, 1::2]).astype(np.uint8))
        o += n_seq_b
        rec[:, o:o + l_seq] = 0xFF
        o += l_seq
        rec[:, o:o + 4] = np.frombuffer(b"NHC\x01", dtype=np.uint8)
        o += 4
        rec[:, o:o + 3] = np.frombuffer(b"CBZ", dtype=np.uint8)
        cell = rng.randint(0, n_cells, size=n)
        put(o + 3, acgt[bc_codes[cell]])
        rec[:, o + 19:o + 21] = np.frombuffer(b"-1", dtype=np.uint8)
        o += 22
        rec[:, o:o + 3] = np.frombuffer(b"UBZ", dtype=np.uint8)
        put(o + 3, acgt[rng.randint(0, 4, size=(n, 12))])
        chunks.append(rec.tobytes())
    raw = b"".join(chunks)
    step = (BGZF_MAX_PAYLOAD // rec_len) * rec_len          # whole records per block, as htslib writes
    blocks = [bytes(head)] + [raw[i:i + step] for i in range(0, len(raw), step)]
    with ThreadPoolExecutor(max_workers=threads) as ex, open(path, "wb") as fp:
        for blk in ex.map(lambda b: _bgzf_block(b, level), blocks):
            fp.write(blk)
        fp.write(BGZF_EOF)
    return barcodes
