"""Device decoder (xcltk_b200/csrc/gpu_decode.cu: BGZF inflate + BAM parse on the GPU) against
the host decoder (decode.cpp, itself checked record by record against an independent Python BAM
reader in test_decode.py): every array of the batch must be identical.  The inflate kernel is
checked byte for byte against zlib."""

import gzip
import os
import random

import numpy as np
import pytest

from util import GOLD

pytestmark = pytest.mark.gpu


def host_decode(paths, maps, cell_tag, umi_tag, want_seq, with_ks=False):
    from xcltk_b200 import lib
    ks = lib.KeySpace()
    h = lib.decode_bams(paths, maps, cell_tag, umi_tag, want_seq, ks, 4)
    return (h, ks) if with_ks else h


def full_maps(paths):
    from xcltk_b200 import lib
    return [np.arange(len(lib.bam_references(p)), dtype=np.int32) for p in paths]


def assert_same_keys(got, exp, ks_got, ks_exp):
    """packed / special keys must be identical; interned ones (bit 63, opaque ids of two
    keyspaces) must name the same strings"""
    got, exp = got.reshape(-1), exp.reshape(-1)
    special = np.uint64(0xFFFFFFFFFFFFFFFE)
    ig = ((got >> np.uint64(63)) == 1) & (got < special)
    ie = ((exp >> np.uint64(63)) == 1) & (exp < special)
    assert np.array_equal(ig, ie)
    assert np.array_equal(got[~ig], exp[~ie])
    if ig.any():
        assert ks_got is not None and ks_exp is not None
        pairs = np.unique(np.stack([got[ig], exp[ie]], axis=1), axis=0)
        assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))     # a bijection
        for a, b in pairs[:: max(1, len(pairs) // 3000)]:
            assert ks_got.decode(int(a)) == ks_exp.decode(int(b))


def assert_same_batch(dev, seen, host, ks_dev=None, ks_host=None):
    got = dev.download()
    info = dev.info()
    assert info["n_reads"] == host.n == got.n
    assert seen == host.n_records_seen
    assert (info["max_aln_len"], info["max_span"]) == (host.max_aln_len, host.max_span)
    assert_same_keys(got.keys, host.keys, ks_dev, ks_host)
    for name in ("pos_end", "fmq", "cig_off", "cigar"):
        assert np.array_equal(getattr(got, name), getattr(host, name)), name
    assert got.has_seq == host.has_seq
    if host.has_seq:
        assert np.array_equal(got.seq_off, host.seq_off)
        assert np.array_equal(got.seq, host.seq)
    assert got.runs == host.runs
    assert got.tiles() == host.tiles()
    got.close()


def tenx_bam(tmp_path, n_reads, seed, name="t.bam", level=6, align=True, mix=None):
    from xcltk_b200 import synth
    rng = random.Random(seed)
    contigs = [("chr1", 2000000), ("chr2", 1200000), ("chrX", 600000), ("chrM", 16000)]
    feats = synth.synth_features(rng, [(c, l - 20000) for c, l in contigs[:3]], 60)    # reads stay on the contig
    barcodes = synth.make_barcodes(rng, 50)
    kw = {} if mix is None else {"cigar_mix": mix}
    refs, recs = synth.gen_10x_records(seed, contigs, feats, n_reads, barcodes, **kw)
    p = str(tmp_path / name)
    synth.write_bam(p, refs, recs, level=level, align=align)
    return p


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_kernel_matches_zlib(gpu_ctx, tmp_path, level):
    """stored (level 0), fixed-Huffman-heavy (tiny blocks) and dynamic blocks"""
    p = tenx_bam(tmp_path, 6000, 3 + level, level=level)
    with gzip.open(p, "rb") as fp:
        exp = fp.read()
    got = gpu_ctx.bgzf_inflate(p)
    assert got.tobytes() == exp


def test_inflate_kernel_tiny_and_incompressible_blocks(gpu_ctx, tmp_path):
    from xcltk_b200 import synth
    rng = random.Random(5)
    payloads = [b"", b"A", b"ACGT" * 3, bytes(rng.getrandbits(8) for _ in range(40000)), b"\0" * 65280,
                bytes(rng.choice(b"ACGT") for _ in range(65280)), (b"x" * 258 + b"y") * 200]
    p = str(tmp_path / "blocks.bgzf")
    with open(p, "wb") as fp:
        for k, pl in enumerate(payloads):
            fp.write(synth._bgzf_block(pl, level=(1, 6, 9)[k % 3]))
        fp.write(synth.BGZF_EOF)
    assert gpu_ctx.bgzf_inflate(p).tobytes() == b"".join(payloads)


@pytest.mark.parametrize("want_seq,cell_tag,umi_tag", [(True, "CB", "UB"), (False, "CB", "UB"), (True, "CB", "CB"),
                                                       (False, None, "UB")])
def test_device_decode_equals_host_decode(gpu_ctx, tmp_path, want_seq, cell_tag, umi_tag):
    p = tenx_bam(tmp_path, 30000, 11)
    maps = full_maps([p])
    host = host_decode([p], maps, cell_tag, umi_tag, want_seq)
    dev, seen = gpu_ctx.decode_bams([p], maps, cell_tag, umi_tag, want_seq)
    assert_same_batch(dev, seen, host)
    dev.close()


def test_device_decode_drops_contigs_and_joins_bams(gpu_ctx, tmp_path):
    """tid_map drops chr2 / chrM; two BAMs are concatenated in list order (fetch order, B4)"""
    p1 = tenx_bam(tmp_path, 20000, 12, "a.bam")
    p2 = tenx_bam(tmp_path, 9000, 13, "b.bam", level=1)
    maps = [np.array([0, -1, 1, -1], dtype=np.int32), np.array([0, -1, 1, -1], dtype=np.int32)]
    host = host_decode([p1, p2], maps, "CB", "UB", True)
    dev, seen = gpu_ctx.decode_bams([p1, p2], maps, "CB", "UB", True)
    assert host.n < host.n_records_seen
    assert [r[0] for r in host.runs] == [0, 0, 1, 1]
    assert_same_batch(dev, seen, host)
    dev.close()


def test_device_decode_handbuilt_cigars(gpu_ctx, tmp_path):
    from xcltk_b200 import synth
    M, I, D, N, S, H, P, EQ, X = range(9)
    recs = [
        ("a", 4, 0, 10, 0, [(M, 20)], "A" * 20, [("CB", "Z", "ACGT-1"), ("UB", "Z", "AC")]),
        ("b", 0, 0, 20, 30, [], "", [("UB", "Z", "")]),
        ("c", 0, 0, 30, 30, [(S, 3), (EQ, 5), (X, 2), (I, 4), (D, 6), (N, 100), (M, 7), (H, 9)],
         "ACGTNACGTNACGTNACGTNA", [("XX", "B", ("S", [1, 2, 3])), ("CB", "Z", "ACGT-1"), ("UB", "A", "T")]),
        ("d", 0, 0, 40, 30, [(M, 1), (I, 1)] * 150, "AC" * 150, [("CB", "i", 5), ("UB", "Z", "01234567")]),
        ("e", 0, 1, 5, 30, [(M, 10)], "ACGTACGTAC", [("CB", "Z", ""), ("xf", "i", 25), ("UB", "Z", "NNN-")]),
        ("f", 0, -1, -1, 0, [], "ACGT", []),
    ]
    p = str(tmp_path / "h.bam")
    synth.write_bam(p, [("chr1", 100000), ("chr2", 50000)], recs)
    maps = full_maps([p])
    host = host_decode([p], maps, "CB", "UB", True)
    dev, seen = gpu_ctx.decode_bams([p], maps, "CB", "UB", True)
    assert host.n == 5 and seen == 6
    assert_same_batch(dev, seen, host)
    dev.close()


def test_device_decode_refuses_what_needs_the_host(gpu_ctx, tmp_path):
    from xcltk_b200 import lib
    p = tenx_bam(tmp_path, 5000, 14, "u.bam", align=False)         # records straddle blocks
    maps = full_maps([p])
    assert gpu_ctx.decode_bams([p], maps, "CB", "UB", True) is None
    assert "block boundaries" in gpu_ctx.decode_fallback_reason
    q = tenx_bam(tmp_path, 5000, 14, "q.bam")
    assert gpu_ctx.decode_bams([q], maps, "CB", None, True) is None   # query-name keys need a keyspace
    assert "keyspace" in gpu_ctx.decode_fallback_reason
    # ... and the host decoder reads both
    assert host_decode([p], maps, "CB", "UB", True).n == host_decode([q], maps, "CB", None, True).n
    with pytest.raises(lib.XgError):
        gpu_ctx.decode_bams([str(tmp_path / "missing.bam")], maps, "CB", "UB", True)


def test_device_decode_rejects_unsorted_and_corrupt(gpu_ctx, tmp_path):
    from xcltk_b200 import lib, synth
    recs = [("r%d" % i, 0, 0, pos, 30, [(0, 50)], "A" * 50, [("CB", "Z", "ACGT"), ("UB", "Z", "AC")])
            for i, pos in enumerate([100, 300, 200])]
    p = str(tmp_path / "unsorted.bam")
    synth.write_bam(p, [("chr1", 100000)], recs)
    with pytest.raises(lib.XgError) as ei:
        gpu_ctx.decode_bams([p], full_maps([p]), "CB", "UB", True)
    assert ei.value.code == -3 and "sorted" in str(ei.value)
    good = tenx_bam(tmp_path, 3000, 15, "good.bam")
    raw = bytearray(open(good, "rb").read())
    raw[len(raw) // 2] ^= 0x5a                                     # flip bits inside a deflate stream
    bad = str(tmp_path / "bad.bam")
    open(bad, "wb").write(bytes(raw))
    # neither decoder checks the gzip CRC; a flipped bit shows as a bad stream, a bad record or
    # an order violation -- or (device only) as a layout the host decoder is asked to judge
    try:
        res = gpu_ctx.decode_bams([bad], full_maps([good]), "CB", "UB", True)
        assert res is None
        with pytest.raises(lib.XgError):
            host_decode([bad], full_maps([good]), "CB", "UB", True)
    except lib.XgError as e:
        assert e.code == -3


BCH869 = os.path.join(GOLD, "bch869_smartseq", "BCH869.output.bam")     # the reference's own fixture, htslib-written


def golden_bams():
    out = []
    for case in sorted(os.listdir(GOLD)):
        d = os.path.join(GOLD, case)
        out += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".bam")]
    return [p for p in out if p != BCH869]


def test_real_htslib_bam_goes_through_the_device_decoder_as_it_is(gpu_ctx):
    """BCH869.output.bam (preprocess/deprecated/merge_smartseq of the reference: samtools-merged SMART-seq reads,
    paired, secondaries, N / I / D / S operations, RG tags) was written by htslib: k_inflate / k_walk / k_extract take
    it without re-blocking, and give the host decoder's arrays -- and the arrays committed with the goldens."""
    from xcltk_b200 import lib
    maps = full_maps([BCH869])
    host, ks_h = host_decode([BCH869], maps, "RG", None, True, with_ks=True)
    assert host.n > 30000
    ks_d = lib.KeySpace()
    res = gpu_ctx.decode_bams([BCH869], maps, "RG", None, True, keyspace=ks_d)
    assert res is not None, getattr(gpu_ctx, "decode_fallback_reason", "")
    assert_same_batch(res[0], res[1], host, ks_d, ks_h)
    got = res[0].download()
    z = np.load(os.path.join(GOLD, "bch869_smartseq", "reads.npz"), allow_pickle=False)
    for name in ("pos_end", "fmq", "cigar", "seq_off", "seq"):
        assert np.array_equal(getattr(got, name), z[name]), name
    assert np.array_equal(got.cig_off, z["cig_off"][:-1])
    names = [str(x) for x in z["umi_names"]]
    for i in range(0, got.n, 97):                     # query names (the counting key without a UMI tag)
        assert ks_d.decode(int(got.keys[i, 1])) == names[int(z["umi_idx"][i])]
    got.close()
    res[0].close()
    host.close()


@pytest.mark.parametrize("path", golden_bams())
def test_device_decode_golden_bams(gpu_ctx, tmp_path, path):
    """The golden BAMs were cut into blocks without regard to records (no htslib writer does
    that): refused as they are, identical to the host decode once laid out as htslib would."""
    from xcltk_b200 import synth
    maps = full_maps([path])
    host = host_decode([path], maps, "CB", "UB", True)
    if host.n > 400:
        assert gpu_ctx.decode_bams([path], maps, "CB", "UB", True) is None
    p = str(tmp_path / "reblocked.bam")
    synth.reblock_bam(path, p)
    with gzip.open(path, "rb") as f1, gzip.open(p, "rb") as f2:
        assert f1.read() == f2.read()
    res = gpu_ctx.decode_bams([p], maps, "CB", "UB", True)
    if res is None:                                    # hand-built BAMs may carry free-text tags
        assert "keyspace" in gpu_ctx.decode_fallback_reason
        return
    assert_same_batch(res[0], res[1], host)
    res[0].close()


@pytest.mark.parametrize("window", [None, 300000, 140000])
def test_device_decode_small_staging_chunks_and_windows(gpu_ctx, tmp_path, monkeypatch, window):
    """file -> HBM in many chunks: blocks cut by chunk boundaries are carried over; with a small
    XG_DECODE_WINDOW the file goes through many windows (bounded device memory): the batch grows
    window by window, runs and sort order are stitched across them; two BAMs share the buffers"""
    monkeypatch.setenv("XG_STAGE_BYTES", str(132 << 10))
    if window:
        monkeypatch.setenv("XG_DECODE_WINDOW", str(window))
    p = tenx_bam(tmp_path, 40000, 16, level=1)
    q = tenx_bam(tmp_path, 15000, 17, "q.bam", level=6)
    assert os.path.getsize(p) > 8 * (132 << 10)
    maps = full_maps([p, q])
    maps[1][1] = -1
    host = host_decode([p, q], maps, "CB", "UB", True)
    dev, seen = gpu_ctx.decode_bams([p, q], maps, "CB", "UB", True)
    if window:
        assert gpu_ctx.timing()[5] >= 6          # windows
    assert_same_batch(dev, seen, host)
    dev.close()


def test_device_decode_large_uniform_bam(gpu_ctx, tmp_path):
    from xcltk_b200 import synth
    p = str(tmp_path / "big.bam")
    contigs = [("chr%d" % (i + 1), 3000000) for i in range(6)]
    synth.write_fast_bam(p, 1500000, contigs, n_cells=300, seed=3)
    maps = full_maps([p])
    host = host_decode([p], maps, "CB", "UB", True)
    dev, seen = gpu_ctx.decode_bams([p], maps, "CB", "UB", True)
    assert_same_batch(dev, seen, host)
    dev.close()


def _random_records(seed, n):
    """records with every aux type around the cell / UMI tags, CIGARs from none to > 255 ops,
    odd sequence lengths, missing tags, unplaced reads at the end"""
    rng = random.Random(seed)
    alpha = "ACGTN-0123456789"
    recs = []
    pos = {0: 0, 1: 0}
    for i in range(n):
        tid = 0 if i < n * 0.6 else 1
        pos[tid] += rng.randrange(0, 40)
        kind = rng.random()
        if kind < 0.05:
            cigar, l_seq = [], rng.choice([0, 5])
        elif kind < 0.10:
            cigar = [(0, 1), (1, 1)] * rng.randrange(130, 200)          # > 255 ops
            l_seq = sum(l for op, l in cigar if op in (0, 1, 4, 7, 8))
        else:
            ops = []
            for _ in range(rng.randrange(1, 7)):
                ops.append((rng.choice([0, 0, 0, 1, 2, 3, 4, 7, 8, 5, 6]), rng.randrange(1, 60)))
            cigar = ops
            l_seq = sum(l for op, l in cigar if op in (0, 1, 4, 7, 8))
        seq = "".join(rng.choice("ACGTN=MRY") for _ in range(l_seq))
        tags = []

        def noise():
            for _ in range(rng.randrange(0, 4)):
                t = rng.choice(["NH", "HI", "AS", "nM", "xf", "RG", "GX", "ZZ", "fl", "ba"])
                typ = rng.choice("AcCsSiIfZHB")
                if typ == "A":
                    v = rng.choice("ACGTXYZ!")
                elif typ in "cCsSiI":
                    lo, hi = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535),
                              "i": (-2 ** 31, 2 ** 31 - 1), "I": (0, 2 ** 32 - 1)}[typ]
                    v = rng.randint(lo, hi)
                elif typ == "f":
                    v = rng.random() * 100
                elif typ == "Z":
                    v = "".join(rng.choice("abc:/ ACGT0123") for _ in range(rng.randrange(0, 30)))
                elif typ == "H":
                    v = "".join(rng.choice("0123456789ABCDEF") for _ in range(2 * rng.randrange(0, 8)))
                else:
                    sub = rng.choice("cCsSiIf")
                    cnt = rng.randrange(0, 9)
                    if sub == "f":
                        v = (sub, [rng.random() for _ in range(cnt)])
                    else:
                        lo, hi = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535),
                                  "i": (-2 ** 31, 2 ** 31 - 1), "I": (0, 2 ** 32 - 1)}[sub]
                        v = (sub, [rng.randint(lo, hi) for _ in range(cnt)])
                tags.append((t, typ, v))
        noise()
        r = rng.random()
        if r < 0.85:
            tags.append(("CB", "Z", "".join(rng.choice(alpha[:6]) for _ in range(rng.randrange(0, 17))) +
                         rng.choice(["", "-1", "-2"])))
        elif r < 0.90:
            tags.append(("CB", "A", rng.choice("ACGTN-7")))
        elif r < 0.93:
            tags.append(("CB", "i", rng.randrange(0, 100)))               # not a string: never a listed barcode
        noise()
        r = rng.random()
        if r < 0.9:
            tags.append(("UB", "Z", "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 9)))))
        elif r < 0.95:
            tags.append(("UB", "A", rng.choice("ACGT5")))
        if rng.random() < 0.1:
            tags.append(("CB", "Z", "TTTT"))                              # a second CB: the first one wins
        noise()
        flag = rng.choice([0, 16, 4, 99, 147, 256, 1024, 2048])
        name = "q%d" % i + "x" * rng.randrange(0, 20)
        recs.append((name, flag, tid, pos[tid], rng.choice([0, 3, 20, 255]), cigar, seq, tags))
    for i in range(5):
        recs.append(("u%d" % i, 4, -1, -1, 0, [], "ACGT", [("CB", "Z", "ACGT")]))
    return recs


@pytest.mark.parametrize("seed,want_seq", [(1, True), (2, False), (3, True)])
def test_device_decode_random_records_all_aux_types(gpu_ctx, tmp_path, seed, want_seq):
    from xcltk_b200 import synth
    recs = _random_records(seed, 6000)
    p = str(tmp_path / "r.bam")
    synth.write_bam(p, [("chr1", 10000000), ("chr2", 10000000)], recs, level=(1, 6, 9)[seed % 3])
    maps = full_maps([p])
    host = host_decode([p], maps, "CB", "UB", want_seq)
    res = gpu_ctx.decode_bams([p], maps, "CB", "UB", want_seq)
    assert res is not None, gpu_ctx.decode_fallback_reason
    assert_same_batch(res[0], res[1], host)
    res[0].close()


def test_device_decode_empty_and_header_heavy_bams(gpu_ctx, tmp_path):
    """no records at all; only unplaced records; a header of 6000 contigs spanning several blocks
    with the records on the last contigs"""
    from xcltk_b200 import synth
    tag = [("CB", "Z", "ACGT-1"), ("UB", "Z", "ACGTAC")]
    cases = {
        "empty": ([("chr1", 1000)], []),
        "unplaced": ([("chr1", 1000)], [("u%d" % i, 4, -1, -1, 0, [], "ACGT", tag) for i in range(10)]),
        "bighdr": ([("contig_with_a_long_name_%06d" % i, 100000 + i) for i in range(6000)],
                   [("r%d" % i, 0, 5990 + i // 100, 10 * (i % 100), 30, [(0, 40)], "ACGT" * 10, tag) for i in range(900)]),
    }
    for name, (refs, recs) in cases.items():
        p = str(tmp_path / (name + ".bam"))
        synth.write_bam(p, refs, recs)
        maps = full_maps([p])
        if name == "bighdr":
            maps[0][5995] = -1                       # one contig in the middle of the data is not wanted
            assert os.path.getsize(p) > 0 and len(refs) * 40 > synth.BGZF_MAX_PAYLOAD
        host = host_decode([p], maps, "CB", "UB", True)
        res = gpu_ctx.decode_bams([p], maps, "CB", "UB", True)
        assert res is not None, (name, gpu_ctx.decode_fallback_reason)
        assert_same_batch(res[0], res[1], host)
        res[0].close()


def _odd_key_records(seed, n):
    """values that do not pack into 63 bits: free-text and long barcodes, integer UMI tags
    (zero is falsy -> EMPTY), one-character tags outside ACGTN-, and repeats of all of them"""
    rng = random.Random(seed)
    cbs = ["cell_%d" % i for i in range(40)] + ["ACGT" * 6 + "-1", "acgtacgt", "ACGT", "", "N-1"]
    recs, pos = [], 0
    for i in range(n):
        pos += rng.randrange(0, 30)
        tags = []
        r = rng.random()
        if r < 0.8:
            tags.append(("CB", "Z", rng.choice(cbs)))
        elif r < 0.9:
            tags.append(("CB", "A", rng.choice("QxA!")))
        r = rng.random()
        if r < 0.4:
            tags.append(("UB", "Z", rng.choice(["umi%d" % rng.randrange(200), "ACGTACGTAC", "ACGT" * 8, ""])))
        elif r < 0.8:
            typ = rng.choice("cCsSiI")
            lo, hi = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535),
                      "i": (-2 ** 31, 2 ** 31 - 1), "I": (0, 2 ** 32 - 1)}[typ]
            tags.append(("UB", typ, rng.choice([0, 1, -1, 7, lo, hi, rng.randint(lo, hi)]) if lo < 0 else
                         rng.choice([0, 1, 7, hi, rng.randint(lo, hi)])))
        elif r < 0.9:
            tags.append(("UB", "A", rng.choice("zZ#A")))
        recs.append(("read:%d:%s" % (i // 2, "x" * rng.randrange(0, 12)), rng.choice([0, 16, 99, 147]), 0, pos, 30,
                     [(0, 50)], "ACGTN" * 10, tags))
    return recs


@pytest.mark.parametrize("cell_tag,umi_tag", [("CB", "UB"), ("CB", None), (None, None), (None, "UB")])
def test_device_decode_interns_odd_keys_through_the_keyspace(gpu_ctx, tmp_path, monkeypatch, cell_tag, umi_tag):
    """free-text barcodes, integer UMI tags, query names (--UMItag None): gathered on the device,
    interned by the host keyspace, patched into the batch -- same strings as the host decoder,
    also across windows and two BAMs"""
    from xcltk_b200 import lib, synth
    monkeypatch.setenv("XG_STAGE_BYTES", str(132 << 10))
    monkeypatch.setenv("XG_DECODE_WINDOW", str(200000))
    paths = []
    for k, n in enumerate((9000, 4000)):
        p = str(tmp_path / ("odd%d.bam" % k))
        synth.write_bam(p, [("chr1", 10000000)], _odd_key_records(20 + k, n), level=1)
        paths.append(p)
    maps = full_maps(paths)
    host, ks_host = host_decode(paths, maps, cell_tag, umi_tag, True, with_ks=True)
    ks = lib.KeySpace()
    res = gpu_ctx.decode_bams(paths, maps, cell_tag, umi_tag, True, ks)
    assert res is not None, gpu_ctx.decode_fallback_reason
    assert ks.n_interned() == ks_host.n_interned() > (0 if (cell_tag, umi_tag) == (None, "UB") else 40)
    t = gpu_ctx.timing()
    assert t[6] > 0 and 0 < t[11] <= t[6]                # values handed over / distinct strings per window
    if (cell_tag, umi_tag) == ("CB", "UB"):
        assert t[11] < t[6] / 2                          # 45 barcodes over thousands of reads: interned once per window
    assert_same_batch(res[0], res[1], host, ks, ks_host)
    res[0].close()


def test_device_decode_leaves_float_umis_to_the_host(gpu_ctx, tmp_path):
    from xcltk_b200 import lib, synth
    recs = [("r%d" % i, 0, 0, 10 * i, 30, [(0, 50)], "A" * 50, [("CB", "Z", "ACGT"), ("UB", "f", 1.5 + i)]) for i in range(50)]
    p = str(tmp_path / "f.bam")
    synth.write_bam(p, [("chr1", 100000)], recs)
    maps = full_maps([p])
    assert gpu_ctx.decode_bams([p], maps, "CB", "UB", True, lib.KeySpace()) is None
    assert "float" in gpu_ctx.decode_fallback_reason
    assert host_decode([p], maps, "CB", "UB", True).n == 50


def test_device_decode_survives_random_corruption(gpu_ctx, tmp_path):
    """bit flips anywhere in the file end in a format error (both decoders verify the gzip CRC32
    of every block, as htslib does), a decline or -- flips in bytes nobody reads -- a valid
    stream; never in a CUDA fault: the context decodes the intact file right afterwards."""
    from xcltk_b200 import lib
    good = tenx_bam(tmp_path, 3000, 41, "good.bam")
    raw = open(good, "rb").read()
    maps = full_maps([good])
    rng = random.Random(7)
    seen_kinds = set()
    for k in range(60):
        b = bytearray(raw)
        for _ in range(rng.choice([1, 1, 2, 6])):
            b[rng.randrange(100, len(b) - 30)] ^= 1 << rng.randrange(8)
        p = str(tmp_path / "bad.bam")
        with open(p, "wb") as fp:
            fp.write(bytes(b))
        try:
            res = gpu_ctx.decode_bams([p], maps, "CB", "UB", True, lib.KeySpace())
            if res is not None:
                res[0].close()
            seen_kinds.add("decoded" if res is not None else "declined")
        except lib.XgError as e:
            assert e.code in (-3, -2), str(e)            # XG_E_FORMAT / XG_E_IO, not XG_E_CUDA
            seen_kinds.add("error")
    assert "error" in seen_kinds
    host = host_decode([good], maps, "CB", "UB", True)
    dev, seen = gpu_ctx.decode_bams([good], maps, "CB", "UB", True)
    assert_same_batch(dev, seen, host)
    dev.close()


@pytest.mark.parametrize("want_seq", [False, True])
def test_device_generated_batch_survives_the_file_round_trip(gpu_ctx, tmp_path, want_seq):
    """records generated in HBM -> xg_write_bam -> device decoder: the batch comes back array for array (the file
    legs of bench.py rest on this: the BAM holds exactly the records whose matrix is known)."""
    from xcltk_b200 import lib, workload
    if want_seq:
        w = workload.make_baf_workload(gpu_ctx, 300000, 200, 3000, seed=41, chroms={"20", "21", "22"})
        names = ["20", "21", "22"]
    else:
        w = workload.make_basefc_workload(gpu_ctx, 400000, 300, 33472, seed=40, chroms={"21", "22"})
        names = ["21", "22"]
    host = w.dreads.download()
    contigs = [("chr" + c, workload.HG38_LEN[c]) for c in names]
    p = str(tmp_path / "rt.bam")
    lib.write_bam(p, host, contigs, None, "CB", "UB", level=1, n_threads=4)
    maps = [np.arange(len(contigs), dtype=np.int32)]
    res = gpu_ctx.decode_bams([p], maps, "CB", "UB", want_seq)
    assert res is not None, getattr(gpu_ctx, "decode_fallback_reason", "")
    if not want_seq:
        assert_same_batch(res[0], res[1], host)
    else:
        # the generator fills whole sequence words; a BAM record holds ceil(91 / 2) bytes of them: compare the bases
        got = res[0].download()
        for name in ("pos_end", "fmq", "cig_off", "cigar", "keys", "seq_off"):
            assert np.array_equal(getattr(got, name), getattr(host, name)), name
        a = got.seq.view(np.uint8).reshape(got.n, -1)
        b = host.seq.view(np.uint8).reshape(host.n, -1)
        assert a.shape == b.shape and np.array_equal(a[:, :45], b[:, :45]) and np.array_equal(a[:, 45] >> 4, b[:, 45] >> 4)
        assert got.runs == host.runs and got.tiles() == host.tiles()
        got.close()
    res[0].close()
    host.close()
    w.dreads.close()
