"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/xcltk_b200.h declares; device entry points fail loudly (no CPU fallback)."""

import ctypes
import os
import re

import pytest

from util import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "xcltk_b200.h")) as fp:
        text = fp.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(xg_lib):
    from xcltk_b200 import lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(xg_lib, n), "library does not export %s" % n
        assert n in lib.SYMBOLS, "%s has no ctypes prototype in xcltk_b200/lib.py" % n
    for n in lib.SYMBOLS:
        assert n in names, "%s is bound but not declared in the header" % n


def test_library_is_in_tree(xg_lib):
    from xcltk_b200 import lib
    assert os.path.realpath(lib.lib_path()).startswith(os.path.realpath(ROOT))
    assert b"xcltk_b200" in xg_lib.xg_version()


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_device_entry_points_fail_loudly_without_gpu():
    from xcltk_b200 import lib
    with pytest.raises(lib.XgError) as ei:
        lib.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under xcltk_b200/ may reference it."""
    bad = []
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "xcltk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, f), errors="replace") as fp:
                    src = fp.read()
                if re.search(r"^\s*(from|import)\s+oracle\b|xg_oracle|oracle/_build", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
