"""GPU parity of the database merge (reference merge_db.cpp:278-820): kwg_merge_slices against a numpy restatement of the
reference's bit-by-bit move, and merge_database_files / `kwage_tools merge_db` against the UNMODIFIED reference merge_db
(oracle/_ref/merge_db, compiled from /root/reference by oracle/Makefile) on the same files, byte for byte."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from kwage_b200 import capi, hostapi as H
from kwage_b200.host import build as hbuild
from oracle import oracle_py as O
import util

pytestmark = pytest.mark.gpu

REF_MERGE = os.path.join(O.REF_DIR, "merge_db")


def setup_module(_):
    hbuild.build()


def merge_np(src1, n1, src2, n2, n_dst1):
    """merge_db.cpp:533-566 with unpackbits: dst1 = src1 columns ++ first (n_dst1 - n1) of src2, dst2 = the rest of src2"""
    b1 = np.unpackbits(src1, axis=1, bitorder="little")[:, :n1]
    b2 = np.unpackbits(src2, axis=1, bitorder="little")[:, :n2]
    take = n_dst1 - n1
    d1 = np.packbits(np.concatenate([b1, b2[:, :take]], axis=1), axis=1, bitorder="little")
    d2 = np.packbits(b2[:, take:], axis=1, bitorder="little") if take < n2 else None
    return d1, d2


@pytest.mark.parametrize("n1,n2,n_dst1,n_slices", [(13, 10, 23, 1000), (8, 8, 16, 64), (1, 1, 2, 5), (1500, 700, 2048, 300), (7, 30, 20, 257),
                                                    (31, 33, 31, 100), (64, 64, 100, 33), (5, 2043, 2048, 50), (2047, 2047, 2048, 17)])
def test_merge_slices_matches_bitwise_move(n1, n2, n_dst1, n_slices):
    rng = np.random.default_rng(n1 * 131 + n2)
    p1, p2 = (n1 + 7) // 8, (n2 + 7) // 8
    # (padding bits of a valid file are zero; garbage there must not leak into the destinations either)
    src1 = rng.integers(0, 256, size=(n_slices, p1), dtype=np.uint8)
    src2 = rng.integers(0, 256, size=(n_slices, p2), dtype=np.uint8)
    exp1, exp2 = merge_np(src1, n1, src2, n2, n_dst1)
    dst1 = np.full_like(exp1, 0xAA)
    dst2 = np.full_like(exp2, 0xAA) if exp2 is not None else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    capi.check(capi.lib().kwg_merge_slices(0, vp(src1), n1, vp(src2), n2, n_slices, n_dst1, vp(dst1), vp(dst2)))
    assert np.array_equal(dst1, exp1)
    if exp2 is not None:
        assert np.array_equal(dst2, exp2)


def test_merge_slices_rejects_bad_arguments():
    a = np.zeros((4, 2), np.uint8)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    L = capi.lib()
    assert L.kwg_merge_slices(0, vp(a), 10, vp(a), 10, 4, 9, vp(a), None) == capi.KWG_ERR_INVALID_ARG      # n_dst1 < n1
    assert L.kwg_merge_slices(0, vp(a), 10, vp(a), 10, 4, 21, vp(a), None) == capi.KWG_ERR_INVALID_ARG     # n_dst1 > n1 + n2
    assert L.kwg_merge_slices(0, vp(a), 10, vp(a), 10, 4, 15, vp(a), None) == capi.KWG_ERR_INVALID_ARG     # remainder without dst2


def make_db(path, first_filter, n, L, k=31, h=3):
    d = os.path.dirname(path)
    files = []
    for j in range(n):
        acc = util.fixture_accession(first_filter + j)
        bits = O.gen_filter_bits(4000 + L, first_filter + j, (1 << L) // 8)
        f = os.path.join(d, "%s.bloom" % acc)
        assert H.write_bloom_file(f, acc, k, L, h, bits)
        files.append(f)
    assert H.build_db(path, k, L, h, files)
    for f in files:
        os.remove(f)


@pytest.mark.skipif(not os.path.exists(REF_MERGE), reason="the compiled reference did not travel to this box")
@pytest.mark.parametrize("n_large,n_small", [(13, 10), (1500, 700), (2040, 9)])
def test_merge_database_files_equal_the_reference(n_large, n_small, tmp_path):
    L = 18
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir()
    ref.mkdir()
    make_db(str(ours / "large.db"), 0, n_large, L)
    make_db(str(ours / "small.db"), n_large, n_small, L)
    for f in ("large.db", "small.db"):
        shutil.copy(str(ours / f), str(ref / f))
    # the reference's main sorts by filter count and merges the smaller file into the larger (merge_db.cpp:225-246)
    r = subprocess.run([REF_MERGE, str(ref / "large.db"), str(ref / "small.db")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    left = H.merge_db(str(ours / "large.db"), str(ours / "small.db"))
    assert sorted(os.listdir(str(ours))) == sorted(os.listdir(str(ref)))
    for f in os.listdir(str(ref)):
        a, b = np.fromfile(str(ours / f), np.uint8), np.fromfile(str(ref / f), np.uint8)
        assert len(a) == len(b) and np.array_equal(a, b), f
    total = n_large + n_small
    assert left == (total if total < 2048 else (total - 2048))
    assert os.path.exists(str(ours / "small.db")) == (total > 2048)


@pytest.mark.skipif(not os.path.exists(REF_MERGE), reason="the compiled reference did not travel to this box")
def test_merge_db_cli_groups_and_merges_like_the_reference(tmp_path):
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir()
    ref.mkdir()
    make_db(str(ours / "a.db"), 0, 5, 18)
    make_db(str(ours / "b.db"), 5, 40, 18)
    make_db(str(ours / "c.db"), 45, 12, 18)
    make_db(str(ours / "d.db"), 0, 3, 19)          # another parameter group: left alone
    for f in os.listdir(str(ours)):
        shutil.copy(str(ours / f), str(ref / f))
    names = ["a.db", "b.db", "c.db", "d.db"]
    r = subprocess.run([REF_MERGE] + [str(ref / n) for n in names], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o = subprocess.run([H.TOOLS_BIN, "merge_db"] + [str(ours / n) for n in names], capture_output=True, text=True)
    assert o.returncode == 0, o.stderr
    assert sorted(os.listdir(str(ours))) == sorted(os.listdir(str(ref)))
    for f in os.listdir(str(ref)):
        assert np.array_equal(np.fromfile(str(ours / f), np.uint8), np.fromfile(str(ref / f), np.uint8)), f


def test_merge_database_files_refuses_what_the_reference_refuses(tmp_path):
    make_db(str(tmp_path / "a.db"), 0, 5, 18)
    make_db(str(tmp_path / "b.db"), 5, 6, 19)
    with pytest.raises(RuntimeError, match="Incompatible"):
        H.merge_db(str(tmp_path / "a.db"), str(tmp_path / "b.db"))
    make_db(str(tmp_path / "c.db"), 5, 6, 18)
    raw = bytearray(open(str(tmp_path / "c.db"), "rb").read())
    raw[44 + 1000] ^= 1                                   # a flipped slice bit: crc32 mismatch (merge_db.cpp:612-618)
    open(str(tmp_path / "c.db"), "wb").write(bytes(raw))
    with pytest.raises(RuntimeError, match="CRC32"):
        H.merge_db(str(tmp_path / "a.db"), str(tmp_path / "c.db"))
