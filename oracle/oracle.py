"""ctypes front end of oracle/xg_oracle.c -- ORACLE / TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by xcltk_b200/.  Inputs are the decoded record arrays (xcltk_b200.lib.HostReads).
"""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libxg_oracle.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u64p = C.POINTER(C.c_uint64)
c_i8p = C.POINTER(C.c_int8)


class OrcCoo(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("cap", C.c_int64), ("row", c_i32p), ("col", c_i32p), ("val", c_i32p)]


class OrcParams(C.Structure):
    _fields_ = [("min_mapq", C.c_double), ("min_len", C.c_int32), ("min_include", C.c_double),
                ("min_include_is_int", C.c_int32), ("incl_flag", C.c_uint32), ("excl_flag", C.c_uint32),
                ("no_orphan", C.c_int32), ("use_cell_tag", C.c_int32), ("need_umi_tag", C.c_int32)]


class OrcSnps(C.Structure):
    _fields_ = [("n", C.c_int32), ("gid", c_i32p), ("pos", c_i32p), ("ref", C.c_char_p), ("alt", C.c_char_p),
                ("ref_idx", c_i8p), ("alt_idx", c_i8p)]


def build(force=False):
    src = os.path.join(HERE, "xg_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_basefc.restype = C.c_int
        _lib.orc_baf.restype = C.c_int
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _coo(lib, m):
    n = int(m.nnz)
    out = tuple(np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int32)
                for p in (m.row, m.col, m.val))
    lib.orc_coo_free(C.byref(m))
    return out


def params(conf):
    """conf: any object with the reference's Config fields (min_mapq, min_len, ...)."""
    mi = getattr(conf, "min_include", 0)
    return OrcParams(float(conf.min_mapq), int(conf.min_len), float(mi), int(isinstance(mi, int)),
                     int(conf.incl_flag), int(conf.excl_flag), int(bool(conf.no_orphan)),
                     int(conf.use_barcodes()), int(conf.use_umi()))


def basefc(host_reads, gid, beg, end, cell_keys, n_samples, par, n_threads=1):
    """Returns (row, col, val) sorted by (row, col), 0-based."""
    lib = load()
    gid, beg, end = (np.ascontiguousarray(a, dtype=np.int32) for a in (gid, beg, end))
    keys = np.ascontiguousarray(cell_keys if cell_keys is not None else [], dtype=np.uint64)
    out = OrcCoo()
    rc = lib.orc_basefc(host_reads.ptr, C.c_int32(len(gid)), _ptr(gid, c_i32p), _ptr(beg, c_i32p),
                        _ptr(end, c_i32p), C.c_int32(len(keys)), _ptr(keys, c_u64p), C.c_int32(n_samples),
                        C.byref(par), C.c_int32(n_threads), C.byref(out))
    if rc != 0:
        raise MemoryError("orc_basefc failed")
    return _coo(lib, out)


def baf(host_reads, snp_gid, snp_pos0, snp_ref, snp_alt, ref_idx, alt_idx, reg_ptr, reg_snp, cell_keys,
        n_samples, par, min_count, min_maf, no_dup_hap, n_threads=1):
    """snp_ref / snp_alt: str of base letters, one per SNP.  Returns (AD, DP, OTH) triples."""
    lib = load()
    g, p = (np.ascontiguousarray(a, dtype=np.int32) for a in (snp_gid, snp_pos0))
    ri, ai = (np.ascontiguousarray(a, dtype=np.int8) for a in (ref_idx, alt_idx))
    rp = np.ascontiguousarray(reg_ptr, dtype=np.int64)
    rs = np.ascontiguousarray(reg_snp, dtype=np.int32)
    keys = np.ascontiguousarray(cell_keys if cell_keys is not None else [], dtype=np.uint64)
    rb, ab = snp_ref.encode("ascii"), snp_alt.encode("ascii")
    snps = OrcSnps(len(g), _ptr(g, c_i32p), _ptr(p, c_i32p), rb, ab, _ptr(ri, c_i8p), _ptr(ai, c_i8p))
    ad, dp, oth = OrcCoo(), OrcCoo(), OrcCoo()
    rc = lib.orc_baf(host_reads.ptr, C.byref(snps), C.c_int32(len(rp) - 1), _ptr(rp, c_i64p),
                     _ptr(rs, c_i32p), C.c_int32(len(keys)), _ptr(keys, c_u64p), C.c_int32(n_samples),
                     C.byref(par), C.c_double(float(min_count)), C.c_double(float(min_maf)),
                     C.c_int32(int(bool(no_dup_hap))), C.c_int32(n_threads),
                     C.byref(ad), C.byref(dp), C.byref(oth))
    if rc != 0:
        raise MemoryError("orc_baf failed")
    return _coo(lib, ad), _coo(lib, dp), _coo(lib, oth)
