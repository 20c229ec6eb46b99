"""The device DEFLATE decoder's logic (xcltk_b200/csrc/inflate.cuh) compiled for the host with
a one-lane group (tests/inflate_host_harness.cpp defines the warp intrinsics as identities) and
checked block by block against zlib.  The multi-lane paths are covered on the GPU
(test_gpu_decode.py)."""

import os
import random
import subprocess

import pytest

from util import GOLD, ROOT

from xcltk_b200 import synth


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("inflate") / "harness")
    src = os.path.join(ROOT, "tests", "inflate_host_harness.cpp")
    subprocess.check_call(["g++", "-O1", "-o", exe, src, "-lz"])
    return exe


def run(exe, path):
    out = subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout
    n_blocks, n_bad = (int(x) for x in out.strip().splitlines()[-1].replace(" blocks,", "").replace(" bad", "").split())
    return n_blocks, n_bad, out


def test_golden_bams_inflate_like_zlib(harness):
    n = 0
    for case in sorted(os.listdir(GOLD)):
        d = os.path.join(GOLD, case)
        for f in sorted(os.listdir(d)):
            if f.endswith(".bam"):
                n_blocks, n_bad, out = run(harness, os.path.join(d, f))
                assert n_blocks > 0 and n_bad == 0, out
                n += n_blocks
    assert n > 100


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_block_kinds(harness, tmp_path, level):
    """stored, fixed-code and dynamic-code blocks; runs, random bytes, empty payloads"""
    rng = random.Random(level)
    payloads = [b"", b"A", b"ACGT" * 3, bytes(rng.getrandbits(8) for _ in range(40000)), b"\0" * 65280,
                bytes(rng.choice(b"ACGT") for _ in range(65280)), (b"x" * 258 + b"y") * 200,
                bytes(rng.choice(b"ACGTN\xff\xff\xff\xff") for _ in range(50000)),
                b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 400) for _ in range(200))[:65280]]
    p = str(tmp_path / "blocks.bgzf")
    with open(p, "wb") as fp:
        for pl in payloads:
            fp.write(synth._bgzf_block(pl, level=level))
        fp.write(synth.BGZF_EOF)
    n_blocks, n_bad, out = run(harness, p)
    assert n_blocks == len(payloads) + 1 and n_bad == 0, out
