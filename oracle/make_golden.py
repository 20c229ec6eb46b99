#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference on the pysam shim.

ORACLE / TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs
/root/reference); the GPU box sees just the committed fixtures.  Usage:

    python oracle/make_golden.py [case ...]

Each case directory holds its inputs (small BAMs written by xcltk_b200.synth, barcode /
feature / SNP files), `case.json` (the runs: wrapper kwargs) and
`expected/<run>/` = the reference's output files, byte for byte
(rdr: features.tsv barcodes.tsv matrix.mtx -- xcltk/rdr/fc/main.py:378-383;
 baf: xcltk.region.tsv xcltk.samples.tsv xcltk.{AD,DP,OTH}.mtx -- xcltk/baf/fc/main.py:373-379).
"""

import json
import os
import random
import shutil
import sys
import time
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
sys.path.insert(0, ROOT)

from xcltk_b200 import synth  # noqa: E402


def _reference_modules():
    for p in (os.path.join(HERE, "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging
    logging.disable(logging.CRITICAL)
    from xcltk.rdr.fc.main import fc_wrapper
    from xcltk.baf.fc.main import afc_wrapper
    return fc_wrapper, afc_wrapper


def resolve_run(case_dir, case, run):
    """Merge case-level defaults with per-run overrides; make paths absolute."""
    g = dict(case.get("defaults", {}))
    g.update({k: v for k, v in run.items() if k in ("sam", "barcodes", "features", "snps", "kind")})
    kw = {}
    for k, v in run.get("kwargs", {}).items():
        if isinstance(v, str) and v.startswith("@"):
            v = os.path.join(case_dir, v[1:])
        kw[k] = v
    ab = lambda x: (x if os.path.isabs(x) else os.path.join(case_dir, x)) if x else None
    return dict(kind=g["kind"], sam=[ab(x) for x in g["sam"]], barcodes=ab(g.get("barcodes")),
                features=ab(g["features"]), snps=ab(g.get("snps")), kwargs=kw)


def run_reference(case_dir, case, run, out_dir):
    """Run one `run` entry of case.json with the reference; returns its return code."""
    fc_wrapper, afc_wrapper = _reference_modules()
    r = resolve_run(case_dir, case, run)
    devnull = open(os.devnull, "w")
    old = sys.stderr
    sys.stderr = devnull
    try:
        if r["kind"] == "basefc":
            ret = fc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], out_dir, **r["kwargs"])
        else:
            ret = afc_wrapper(",".join(r["sam"]), r["barcodes"], r["features"], r["snps"], out_dir,
                              **r["kwargs"])
    finally:
        sys.stderr = old
        devnull.close()
    return ret


def finish_case(case_dir, case):
    with open(os.path.join(case_dir, "case.json"), "w") as fp:
        json.dump(case, fp, indent=1, sort_keys=True)
        fp.write("\n")
    exp = os.path.join(case_dir, "expected")
    shutil.rmtree(exp, ignore_errors=True)
    os.makedirs(exp)
    for run in case["runs"]:
        out = os.path.join(exp, run["name"])
        t_ref = time.perf_counter()
        ret = run_reference(case_dir, case, run, out)
        print("  reference wall time for %s: %.1f s" % (run["name"], time.perf_counter() - t_ref))
        with open(os.path.join(out, "RETCODE"), "w") as fp:
            fp.write("%d\n" % ret)
        left = sorted(f for f in os.listdir(out) if "pickle" in f or f[-1].isdigit())
        assert not left, left
        print("  %-28s ret=%d  %s" % (run["name"], ret, summarize(out)))


def summarize(out):
    s = []
    for f in sorted(os.listdir(out)):
        if f.endswith(".mtx"):
            with open(os.path.join(out, f)) as fp:
                lines = fp.read().splitlines()
            s.append("%s[%s]" % (f.replace("xcltk.", ""), lines[2].replace("\t", ",")))
    return " ".join(s)


def fresh(name):
    d = os.path.join(GOLD, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


def seq_with(L, overrides):
    s = ["A"] * L
    for i, b in overrides.items():
        s[i] = b
    return "".join(s)


# ------------------------------------------------------------------ mini cases
def case_d1():
    d = fresh("d1_basefc_mini")
    M, N, S = 0, 3, 4

    def r(name, pos, flag, mapq, cigar, cb, ub, L=40):
        tags = []
        if cb is not None:
            tags.append(("CB", "Z", cb))
        if ub is not None:
            tags.append(("UB", "Z", ub))
        return (name, flag, 0, pos, mapq, cigar, "A" * L, tags)

    recs = [
        r("r1", 100, 0, 255, [(M, 40)], "BBB", "U1"),
        r("r2", 100, 0, 255, [(M, 40)], "BBB", "U1"),
        r("r3", 105, 0, 255, [(M, 40)], "AAA", "U1"),
        r("r6", 120, 0, 10, [(M, 40)], "AAA", "U4"),
        r("r7", 120, 256, 255, [(M, 40)], "AAA", "U5"),
        r("r8", 120, 0, 255, [(M, 40)], "ZZZ", "U6"),
        r("r9", 120, 0, 255, [(M, 40)], "AAA", None),
        r("r10", 120, 0, 255, [(S, 15), (M, 25)], "AAA", "U7"),
        r("r11", 120, 1, 255, [(M, 40)], "AAA", "U8"),
        r("r12", 120, 2048, 255, [(M, 40)], "AAA", "U9"),
        r("r13", 120, 0, 255, [(M, 40)], "AAA", ""),
        r("r4", 170, 0, 255, [(M, 40)], "AAA", "U2"),
        r("r5", 180, 0, 255, [(M, 20), (N, 500), (M, 20)], "AAA", "U3"),
    ]
    recs.sort(key=lambda x: x[3])
    synth.write_bam(os.path.join(d, "a.bam"), [("chr1", 100000)], recs)
    synth.write_lines(os.path.join(d, "barcodes.tsv"), ["BBB", "AAA"])
    with open(os.path.join(d, "features.tsv"), "w") as fp:
        fp.write("chr1\t175\t400\tg2\n1\t101\t200\tg1\nchr1\t600\t800\tg3\n"
                 "chr1\t0\t50\tg0\nchrQ\t1\t100\tgq\n")
    case = {"defaults": {"kind": "basefc", "sam": ["a.bam"], "barcodes": "barcodes.tsv",
                         "features": "features.tsv"}, "runs": [
                {"name": "defaults", "kwargs": {}},
                {"name": "min_include_20", "kwargs": {"min_include": 20}},
                {"name": "umi_none", "kwargs": {"umi_tag": "None"}},
                {"name": "min_include_0", "kwargs": {"min_include": 0}},
                {"name": "min_include_1.0", "kwargs": {"min_include": 1.0}},
                {"name": "frac_0.5_mapq0_len1", "kwargs": {"min_include": 0.5, "min_mapq": 0, "min_len": 1}},
                {"name": "no_all_reg", "kwargs": {"output_all_reg": False}},
                {"name": "incl_flag_16", "kwargs": {"incl_flag": 16}},
                {"name": "count_orphan", "kwargs": {"no_orphan": False}},
                {"name": "ncores3", "kwargs": {"ncores": 3}},
                {"name": "mapq_float", "kwargs": {"min_mapq": 9.5}},
            ]}
    finish_case(d, case)


def case_d2():
    d = fresh("d2_baf_mini")
    M, N = 0, 3

    def r(name, pos, cigar, cb, ub, ov, L=40):
        L = sum(l for op, l in cigar if op in (0, 1, 4, 7, 8))
        return (name, 0, 0, pos, 255, cigar, seq_with(L, ov), [("CB", "Z", cb), ("UB", "Z", ub)])

    recs = [
        r("a1", 100, [(M, 10), (N, 30), (M, 30)], "AAA", "U1", {}),
        r("a2", 105, [(M, 40)], "AAA", "U1", {14: "T"}),
        r("b1", 110, [(M, 40)], "AAA", "U2", {9: "C", 39: "G"}),
        r("c1", 112, [(M, 40)], "AAA", "U3", {7: "T"}),
        r("d1", 115, [(M, 40)], "BBB", "U4", {4: "G", 34: "C"}),
        r("e1", 118, [(M, 30)], "BBB", "U5", {1: "N"}),
        r("f1", 145, [(M, 40)], "BBB", "U6", {4: "G"}),
    ]
    synth.write_bam(os.path.join(d, "a.bam"), [("1", 100000)], recs)
    synth.write_lines(os.path.join(d, "barcodes.tsv"), ["AAA", "BBB"])
    with open(os.path.join(d, "features.tsv"), "w") as fp:
        fp.write("1\t101\t400\tg1\n1\t500\t600\tg_nosnp\n1\t130\t160\tg_nested\n")
    with open(os.path.join(d, "snps.tsv"), "w") as fp:
        fp.write("chrom\tpos\tref\talt\tref_hap\talt_hap\n"
                 "1\t120\tC\tT\t0\t1\n1\t150\tG\tA\t1\t0\n1\t550\tA\tC\t0\t1\n")
    with open(os.path.join(d, "snps.vcf"), "w") as fp:
        fp.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n"
                 "chr1\t120\t.\tc\tT\t.\tPASS\t.\tGT:AD\t0|1:3\n"
                 "1\t150\t.\tG\tA\t.\tPASS\t.\tAD:GT\t3:1/0\n"
                 "1\t550\t.\tA\tC\t.\tPASS\t.\tGT\t0|1\n"
                 "1\t151\t.\tA\tCT\t.\tPASS\t.\tGT\t0|1\n"
                 "1\t152\t.\tA\tC\t.\tPASS\t.\tGT\t1|1\n"
                 "1\t153\t.\tA\tC\t.\tPASS\t.\tDP\t4\n")
    case = {"defaults": {"kind": "baf", "sam": ["a.bam"], "barcodes": "barcodes.tsv",
                         "features": "features.tsv", "snps": "snps.tsv"}, "runs": [
                {"name": "defaults", "kwargs": {}},
                {"name": "dup_hap_all_reg", "kwargs": {"no_dup_hap": False, "output_all_reg": True}},
                {"name": "min_count_6", "kwargs": {"min_count": 6}},
                {"name": "min_count_5", "kwargs": {"min_count": 5, "output_all_reg": True}},
                {"name": "min_maf_0.3", "kwargs": {"min_maf": 0.3, "output_all_reg": True}},
                {"name": "vcf", "snps": "snps.vcf", "kwargs": {"output_all_reg": True}},
                {"name": "umi_none", "kwargs": {"umi_tag": "None", "output_all_reg": True}},
                {"name": "ncores2", "kwargs": {"ncores": 2, "output_all_reg": True}},
            ]}
    finish_case(d, case)


def case_d3():
    d = fresh("d3_sample_mode")
    M = 0

    def r(name, pos, flag, ov, tags=()):
        return (name, flag, 0, pos, 255, [(M, 40)], seq_with(40, ov), list(tags))

    w1 = [r("q1", 100, 99, {19: "C"}), r("q2", 100, 355, {}), r("q1", 110, 147, {9: "T"})]
    w2 = [r("q1", 100, 83, {19: "T"}), r("q9", 100, 1123, {})]
    tg = [("CB", "Z", "AAA"), ("UB", "Z", "U1")]
    p1 = [r("x1", 105, 0, {14: "T"}, tg)]
    p2 = [r("x2", 100, 0, {19: "C"}, tg)]
    for nm, recs in (("w1", w1), ("w2", w2), ("p1", p1), ("p2", p2)):
        synth.write_bam(os.path.join(d, nm + ".bam"), [("1", 100000)], recs)
    synth.write_lines(os.path.join(d, "barcodes.tsv"), ["AAA"])
    synth.write_lines(os.path.join(d, "sample_ids.tsv"), ["cellB", "cellA"])
    with open(os.path.join(d, "features.tsv"), "w") as fp:
        fp.write("1\t101\t200\tg1\n")
    with open(os.path.join(d, "snps.tsv"), "w") as fp:
        fp.write("chrom\tpos\tref\talt\tref_hap\talt_hap\n1\t120\tC\tT\t0\t1\n")
    nb = {"cell_tag": None, "umi_tag": None}
    case = {"defaults": {"sam": ["w1.bam", "w2.bam"], "barcodes": None, "features": "features.tsv",
                         "snps": "snps.tsv"}, "runs": [
        {"name": "rdr_ids", "kind": "basefc", "kwargs": dict(nb, sample_ids="cellB,cellA")},
        {"name": "rdr_default_ids", "kind": "basefc", "kwargs": dict(nb)},
        {"name": "rdr_ids_file", "kind": "basefc", "kwargs": dict(nb, sample_id_fn="@sample_ids.tsv")},
        {"name": "rdr_pooled_p1p2", "kind": "basefc", "sam": ["p1.bam", "p2.bam"],
         "barcodes": "barcodes.tsv", "kwargs": {}},
        {"name": "baf_ids", "kind": "baf", "kwargs": dict(nb, sample_ids="cellB,cellA")},
        {"name": "baf_p1p2", "kind": "baf", "sam": ["p1.bam", "p2.bam"], "barcodes": "barcodes.tsv",
         "kwargs": {}},
        {"name": "baf_p2p1", "kind": "baf", "sam": ["p2.bam", "p1.bam"], "barcodes": "barcodes.tsv",
         "kwargs": {}},
    ]}
    finish_case(d, case)


# ------------------------------------------------------------------ synthetic 10x cases
def case_chr22(name="c1_chr22_10x", n_reads=40000, n_bc=60, seed=7):
    d = fresh(name)
    rng = random.Random(seed)
    feats_all = synth.load_features(os.path.join(
        REF, "data/anno/annotate_genes_hg38_update_20230126.txt"))
    f22 = [f for f in feats_all if f[0] == "22"]
    # features on other contigs exercise the "unknown contig -> empty row" path (R4)
    keep = f22 + [f for f in feats_all if f[0] in ("21", "X")][:200]
    rng.shuffle(keep)
    keep = keep[:700]
    bcs = synth.make_barcodes(rng, n_bc)
    snps = synth.gen_snps(seed + 1, [f for f in keep if f[0] == "22"], 3000)
    refs, recs = synth.gen_10x_records(seed, [("22", 50818468)], keep, n_reads, bcs,
                                       chr_prefix="chr", snps=snps)
    synth.write_bam(os.path.join(d, "a.bam"), refs, recs)
    shuffled = list(bcs)
    rng.shuffle(shuffled)
    synth.write_lines(os.path.join(d, "barcodes.tsv"), shuffled)
    with open(os.path.join(d, "features.tsv"), "w") as fp:
        for i, (c, s, e, n) in enumerate(keep):
            fp.write("%s%s\t%d\t%d\t%s\n" % ("chr" if i % 3 == 0 else "", c, s, e, n))
    synth.write_snp_tsv(os.path.join(d, "snps.tsv"), snps, chr_prefix="")
    case = {"defaults": {"sam": ["a.bam"], "barcodes": "barcodes.tsv", "features": "features.tsv",
                         "snps": "snps.tsv"}, "runs": [
        {"name": "rdr_defaults", "kind": "basefc", "kwargs": {"ncores": 4}},
        {"name": "rdr_frac_0.5", "kind": "basefc", "kwargs": {"min_include": 0.5, "ncores": 4}},
        {"name": "rdr_len_30_mapq0", "kind": "basefc",
         "kwargs": {"min_include": 30, "min_mapq": 0, "ncores": 4}},
        {"name": "rdr_umi_none", "kind": "basefc", "kwargs": {"umi_tag": "None", "ncores": 4}},
        {"name": "baf_defaults", "kind": "baf", "kwargs": {"ncores": 4}},
        {"name": "baf_all_reg_dup", "kind": "baf",
         "kwargs": {"output_all_reg": True, "no_dup_hap": False, "ncores": 4}},
        {"name": "baf_count3_maf0.1", "kind": "baf",
         "kwargs": {"min_count": 3, "min_maf": 0.1, "ncores": 4, "output_all_reg": True}},
        {"name": "baf_umi_none", "kind": "baf", "kwargs": {"umi_tag": "None", "ncores": 4}},
    ]}
    finish_case(d, case)


CASES = {"d1": case_d1, "d2": case_d2, "d3": case_d3, "chr22": case_chr22}



# ------------------------------------------------------------------ the reference's only real fixture
def case_bch869():
    """BCH869 SMART-seq BAM (preprocess/deprecated/merge_smartseq, 32 764 reads, RG = cell).
    The BAM itself stays in /root/reference; the fixture holds the decoded record arrays
    (xcltk_b200 decoder, cross-checked against oracle/shim in tests/test_decode.py) + the
    reference's outputs on the real BAM."""
    import gzip
    import numpy as np
    from xcltk_b200 import lib
    d = fresh("bch869_smartseq")
    src = os.path.join(REF, "preprocess/deprecated/merge_smartseq")
    bam = os.path.join(src, "BCH869.output.bam")
    shutil.copyfile(bam, os.path.join(d, "BCH869.output.bam"))      # travels: the GPU tests decode it on the device
    # inputs the reference reads
    with open(os.path.join(src, "BCH869.output.492.RG.barcodes.tsv")) as fp:
        barcodes = [x.strip() for x in fp if x.strip()]
    synth.write_lines(os.path.join(d, "barcodes.tsv"), barcodes)
    with open(os.path.join(REF, "data/anno/annotate_genes_hg19_update_20230126.txt")) as fp, \
            gzip.GzipFile(os.path.join(d, "features.tsv.gz"), "wb", mtime=0) as out:
        for line in fp:
            out.write(("\t".join(line.rstrip("\n").split("\t")[:4]) + "\n").encode())
    # decoded records (identity contig map: gid = tid)
    ks = lib.KeySpace()
    refs = lib.bam_references(bam)
    hr = lib.decode_bams([bam], [np.arange(len(refs), dtype=np.int32)], "RG", None, True, ks, 4)
    cells, umis = {}, {}
    cell_idx = np.full(hr.n, -1, dtype=np.int32)
    umi_idx = np.zeros(hr.n, dtype=np.int32)
    for i in range(hr.n):
        ck, uk = int(hr.keys[i, 0]), int(hr.keys[i, 1])
        if ck != lib.XG_KEY_NONE:
            cell_idx[i] = cells.setdefault(ks.decode(ck), len(cells))
        umi_idx[i] = umis.setdefault(ks.decode(uk), len(umis))
    # SNPs at covered positions (REF = an observed base) so that the pileup has something to count
    rng = random.Random(869)
    nt16 = "=ACMGRSVTWYHKDBN"
    snp_rows, seen = [], set()
    while len(snp_rows) < 2547:
        i = rng.randrange(hr.n)
        f = int(hr.fmq[i])
        if (f >> 24) != 0:
            continue                      # simple reads only: query index = ref offset
        pos, end = (int(x) for x in hr.pos_end[i])
        q = rng.randrange(end - pos)
        tid = [r for r in hr.runs if r[2] <= i < r[3]][0][1]
        if (tid, pos + q) in seen:
            continue
        raw = hr.seq[int(hr.seq_off[i]):int(hr.seq_off[i]) + 16].tobytes()
        b = raw[q >> 1]
        base = nt16[(b & 15) if (q & 1) else (b >> 4)]
        if base not in "ACGT":
            continue
        seen.add((tid, pos + q))
        alt = rng.choice([x for x in "ACGT" if x != base])
        a1 = rng.randrange(2)
        snp_rows.append((refs[tid][0], pos + q + 1, base, alt, a1, 1 - a1))
    with open(os.path.join(d, "snps.tsv"), "w") as fp:
        fp.write("chrom\tpos\tref\talt\tref_hap\talt_hap\n")
        for r in snp_rows:
            fp.write("%s\t%d\t%s\t%s\t%d\t%d\n" % r)
    np.savez_compressed(
        os.path.join(d, "reads.npz"), pos_end=hr.pos_end, fmq=hr.fmq, cig_off=np.append(hr.cig_off, len(hr.cigar)),
        cigar=hr.cigar, seq_off=hr.seq_off, seq=hr.seq, runs=np.array(hr.runs, dtype=np.int64),
        cell_idx=cell_idx, umi_idx=umi_idx, cell_names=np.array(list(cells)), umi_names=np.array(list(umis)),
        ref_names=np.array([r[0] for r in refs]), max_aln_len=hr.max_aln_len, max_span=hr.max_span)
    hr.close()
    # `sam` holds the real BAM for the reference run; the tests feed reads.npz instead
    case = {"defaults": {"sam": [bam], "barcodes": "barcodes.tsv", "features": "features.tsv.gz",
                         "snps": "snps.tsv", "reads_npz": "reads.npz"}, "runs": [
        {"name": "rdr_rg_barcodes", "kind": "basefc", "kwargs": {"cell_tag": "RG", "umi_tag": "None", "ncores": 8}},
        {"name": "rdr_rg_frac0.5_len20", "kind": "basefc",
         "kwargs": {"cell_tag": "RG", "umi_tag": "None", "ncores": 8, "min_include": 0.5, "min_len": 20, "min_mapq": 0}},
        {"name": "rdr_sample_id", "kind": "basefc", "barcodes": None,
         "kwargs": {"cell_tag": None, "umi_tag": None, "sample_ids": "BCH869", "ncores": 8}},
        {"name": "baf_rg_all_reg", "kind": "baf",
         "kwargs": {"cell_tag": "RG", "umi_tag": None, "ncores": 8, "output_all_reg": True}},
        {"name": "baf_rg_count2", "kind": "baf",
         "kwargs": {"cell_tag": "RG", "umi_tag": None, "ncores": 8, "min_count": 2, "no_dup_hap": False}},
    ]}
    finish_case(d, case)


CASES["bch869"] = case_bch869


# ------------------------------------------------------------------ config C1 at its full shape
def case_c1_full(n_reads=1000000, n_bc=500, seed=7):
    """BASELINE.json configs[0] / SURVEY.md 8(d) C1 as named: ~1M reads on chr22 (BAM contig `chr22`, features
    without the prefix: the sam_fetch fallback), 500 barcodes, ALL 33 472 rows of the hg38 gene table.  The BAM
    (tens of MB) does not travel; the fixture holds the decoded record arrays -- what the counting kernels see --
    and the files the unmodified reference wrote from the BAM itself."""
    import numpy as np
    from xcltk_b200 import lib
    d = fresh("c1_full_chr22")
    rng = random.Random(seed)
    feats_all = synth.load_features(os.path.join(REF, "data/anno/annotate_genes_hg38_update_20230126.txt"))
    assert len(feats_all) == 33472
    bcs = synth.make_barcodes(rng, n_bc)
    refs, recs = synth.gen_10x_records(seed, [("22", 50818468)], [f for f in feats_all if f[0] == "22"], n_reads, bcs,
                                       chr_prefix="chr")
    tmp = tempfile.mkdtemp()
    bam = os.path.join(tmp, "c1.bam")
    synth.write_bam(bam, refs, recs)
    shuffled = list(bcs)
    rng.shuffle(shuffled)
    synth.write_lines(os.path.join(d, "barcodes.tsv"), shuffled)
    import gzip
    with gzip.GzipFile(os.path.join(d, "features.tsv.gz"), "wb", mtime=0) as out:
        for c, s, e, n in feats_all:
            out.write(("%s\t%d\t%d\t%s\n" % (c, s, e, n)).encode())
    ks = lib.KeySpace()
    hr = lib.decode_bams([bam], [np.arange(1, dtype=np.int32)], "CB", "UB", False, ks, 8)
    cells, umis = {}, {}
    cell_idx = np.full(hr.n, -1, dtype=np.int32)
    umi_idx = np.full(hr.n, -1, dtype=np.int32)
    special = {lib.XG_KEY_NONE: "\x00none", 0: ""}
    for i in range(hr.n):
        ck, uk = int(hr.keys[i, 0]), int(hr.keys[i, 1])
        if ck != lib.XG_KEY_NONE:
            cell_idx[i] = cells.setdefault(ks.decode(ck), len(cells))
        if uk != lib.XG_KEY_NONE:
            umi_idx[i] = umis.setdefault(ks.decode(uk) if uk else "", len(umis))
    np.savez_compressed(
        os.path.join(d, "reads.npz"), pos_end=hr.pos_end, fmq=hr.fmq, cig_off=np.append(hr.cig_off, len(hr.cigar)),
        cigar=hr.cigar, seq_off=np.zeros(0, np.uint32), seq=np.zeros(0, np.uint32),
        runs=np.array(hr.runs, dtype=np.int64), cell_idx=cell_idx, umi_idx=umi_idx,
        cell_names=np.array(list(cells)), umi_names=np.array(list(umis)), ref_names=np.array([r[0] for r in refs]),
        max_aln_len=hr.max_aln_len, max_span=hr.max_span)
    print("  %d records decoded, %d cells, %d UMIs" % (hr.n, len(cells), len(umis)))
    hr.close()
    case = {"defaults": {"sam": [bam], "barcodes": "barcodes.tsv", "features": "features.tsv.gz",
                         "reads_npz": "reads.npz"}, "runs": [
        {"name": "rdr_defaults", "kind": "basefc", "kwargs": {"ncores": 8}},
    ]}
    finish_case(d, case)
    # the BAM is gone after this: keep case.json free of the temporary path
    case["defaults"]["sam"] = ["c1.bam (regenerate with oracle/make_golden.py c1_full)"]
    with open(os.path.join(d, "case.json"), "w") as fp:
        json.dump(case, fp, indent=1, sort_keys=True)
        fp.write("\n")
    shutil.rmtree(tmp, ignore_errors=True)


CASES["c1_full"] = case_c1_full


# ------------------------------------------------------------------ afc with the local-phasing pre-step
def case_local_phasing(seed=23):
    """afc_wrapper as `xcltk baf` calls it (baf/pipeline.py:341-360): cellsnp_dir given, so long regions
    (>= 50 kb, >= 2 SNPs spanning >= 50 kb) have their SNPs re-phased by the EM before counting
    (baf/fc/main.py:112-149).  BAM, barcodes, features and SNPs are those of c1_chr22_10x; the cellsnp-lite
    directory is synthetic: two clones of cells with allelic ratios 0.15 / 0.85 in every region and a third of
    the SNPs phased the wrong way round, so that the EM has flips to find."""
    import gzip
    import numpy as np
    import scipy.io
    import scipy.sparse
    d = fresh("c2_local_phasing")
    src = os.path.join(GOLD, "c1_chr22_10x")
    with open(os.path.join(src, "barcodes.tsv")) as fp:
        cells = [x.strip() for x in fp if x.strip()]
    snps = []
    with open(os.path.join(src, "snps.tsv")) as fp:
        next(fp)
        for line in fp:
            c, pos, ref, alt, rh, ah = line.rstrip("\n").split("\t")
            snps.append((c, int(pos), ref, alt, int(rh)))
    rng = np.random.RandomState(seed)
    n_snp, n_cell = len(snps), len(cells)
    clone = rng.randint(0, 3, size=n_cell)                       # 0 / 1: the two CNA clones, 2: balanced cells
    theta = np.where(clone == 0, 0.15, np.where(clone == 1, 0.85, 0.5))
    wrong = rng.rand(n_snp) < 0.33                               # phased the wrong way round by "Eagle2"
    DP = rng.poisson(1.2, size=(n_snp, n_cell)) * (rng.rand(n_snp, n_cell) < 0.6)
    # ALT count: haplotype-1 allele is ALT iff ref_hap == 0; theta = fraction of haplotype 1
    hap1_is_alt = np.array([s[4] == 0 for s in snps]) ^ wrong
    p = np.where(hap1_is_alt[:, None], theta[None, :], 1 - theta[None, :])
    AD = rng.binomial(DP, p)
    cs = os.path.join(d, "cellsnp")
    os.makedirs(cs)
    order = sorted(cells)                                        # cellsnp-lite writes its own (sorted) cell order
    perm = [cells.index(c) for c in order]
    synth.write_lines(os.path.join(cs, "cellSNP.samples.tsv"), order)
    with gzip.GzipFile(os.path.join(cs, "cellSNP.base.vcf.gz"), "wb", mtime=0) as fp:
        fp.write(b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for c, pos, ref, alt, _ in snps:
            fp.write(("%s\t%d\t.\t%s\t%s\t.\tPASS\tAD=1;DP=2;OTH=0\n" % (c, pos, ref, alt)).encode())
    for name, m in (("AD", AD), ("DP", DP), ("OTH", np.zeros_like(DP))):
        scipy.io.mmwrite(os.path.join(cs, "cellSNP.tag.%s.mtx" % name), scipy.sparse.csr_matrix(m[:, perm]),
                         field="integer")
    synth.write_lines(os.path.join(d, "ref_cells.tsv"), [c for c, k in zip(cells, clone) if k == 2][:8])
    rel = lambda f: os.path.join("..", "c1_chr22_10x", f)
    case = {"defaults": {"kind": "baf", "sam": [rel("a.bam")], "barcodes": rel("barcodes.tsv"),
                         "features": rel("features.tsv"), "snps": rel("snps.tsv")}, "runs": [
        {"name": "pipeline_call", "kwargs": {"cellsnp_dir": "@cellsnp", "output_all_reg": True, "ncores": 4}},
        {"name": "ref_cells", "kwargs": {"cellsnp_dir": "@cellsnp", "ref_cell_fn": "@ref_cells.tsv", "ncores": 4}},
        {"name": "no_phasing", "kwargs": {"output_all_reg": True, "ncores": 4}},
    ]}
    finish_case(d, case)


CASES["local_phasing"] = case_local_phasing


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    todo = sys.argv[1:] or list(CASES)
    for c in todo:
        print("case", c)
        CASES[c]()
