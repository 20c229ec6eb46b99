"""Helpers shared by the tests: golden-case enumeration and byte comparison."""

import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

RDR_FILES = ["features.tsv", "barcodes.tsv", "matrix.mtx"]
BAF_FILES = ["xcltk.region.tsv", "xcltk.samples.tsv", "xcltk.AD.mtx", "xcltk.DP.mtx", "xcltk.OTH.mtx"]


def golden_runs(kind=None):
    """[(case_name, run_name)] of every committed golden run (optionally one kind)."""
    out = []
    for case in sorted(os.listdir(GOLD)):
        cj = os.path.join(GOLD, case, "case.json")
        if not os.path.isfile(cj):
            continue
        with open(cj) as fp:
            spec = json.load(fp)
        for run in spec["runs"]:
            k = run.get("kind", spec.get("defaults", {}).get("kind"))
            if kind is None or k == kind:
                out.append((case, run["name"]))
    return out


def resolve(case, run_name):
    case_dir = os.path.join(GOLD, case)
    with open(os.path.join(case_dir, "case.json")) as fp:
        spec = json.load(fp)
    run = [r for r in spec["runs"] if r["name"] == run_name][0]
    g = dict(spec.get("defaults", {}))
    g.update({k: v for k, v in run.items() if k in ("sam", "barcodes", "features", "snps", "kind")})
    kw = {}
    for k, v in run.get("kwargs", {}).items():
        if isinstance(v, str) and v.startswith("@"):
            v = os.path.join(case_dir, v[1:])
        kw[k] = v

    def ab(x):
        return os.path.join(case_dir, x) if x else None
    npz = ab(spec.get("defaults", {}).get("reads_npz"))
    sam = [npz] if npz else [ab(x) for x in g["sam"]]      # the BAM of an npz case cannot travel
    return dict(kind=g["kind"], sam=sam, barcodes=ab(g.get("barcodes")),
                features=ab(g["features"]), snps=ab(g.get("snps")), kwargs=kw, reads_npz=npz,
                expected=os.path.join(case_dir, "expected", run_name))


def read(path):
    with open(path, "rb") as fp:
        return fp.read()


def compare_dirs(expected, got, files):
    for f in files:
        e, g = read(os.path.join(expected, f)), read(os.path.join(got, f))
        assert e == g, "%s differs from the reference output\n--- expected\n%s\n--- got\n%s" % (
            f, e[:600].decode(), g[:600].decode())
    extra = sorted(set(os.listdir(got)) - set(files))
    assert not extra, "unexpected files left in the output directory: %s" % extra
