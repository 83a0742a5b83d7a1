"""CPU: the DEVICE arithmetic (kwage_b200/csrc/bitops.cuh compiled as host C++ with intrinsic shims,
tests/host_emul/emul.cpp) against the oracle.  Catches kernel-math bugs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_py as O
import synth_cases as S

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def E():
    e = C.CDLL(os.path.join(HERE, "host_emul", "libkwage_emul.so"))
    e.emu_scan.restype = C.c_uint64
    e.emu_synth_rnd.restype = C.c_uint64
    e.emu_synth_rnd.argtypes = [C.c_uint64] * 3
    return e


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def scan(E, bases, offsets, k, nh):
    n = len(bases)
    words = np.zeros(n + 1, np.uint64)
    pos = np.zeros(n + 1, np.uint64)
    hs = np.zeros((n + 1) * nh, np.uint32)
    cnt = E.emu_scan(p(bases), C.c_uint64(n), p(offsets), C.c_uint64(len(offsets) - 1), C.c_uint32(k), C.c_uint32(nh),
                     p(words), p(pos), p(hs))
    return words[:cnt], pos[:cnt], hs[: cnt * nh].reshape(cnt, nh)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 21, 25, 30, 31, 32])
def test_tile_scan_matches_oracle(E, k):
    nh = 1 + (k % 8)
    n_reads = 150
    flat = S.mutate(O.gen_reads(100 + k, 0, n_reads, 160), 100 + k, n_rate=41, lower_rate=5)
    bases, offsets = S.ragged(flat, 100 + k, n_reads, 0, 160)     # crosses several 4096-base tiles
    w, ps, hs = scan(E, bases, offsets, k, nh)
    ow, op = [], []
    for r in range(n_reads):
        a, b = int(offsets[r]), int(offsets[r + 1])
        ww, ll = O.canonical_kmers(bases[a:b], k)
        ow.append(ww)
        op.append(ll + np.uint64(a))
    ow, op = np.concatenate(ow), np.concatenate(op)
    assert np.array_equal(w, ow) and np.array_equal(ps, op)
    sel = np.arange(0, len(ow), max(1, len(ow) // 300))
    exp = np.array([[O.murmur3_word(int(ow[i]), k, s) for s in range(nh)] for i in sel], dtype=np.uint32).reshape(-1, nh)
    assert np.array_equal(hs[sel], exp)


@pytest.mark.parametrize("k", [1, 5, 16, 31, 32, 33, 34, 40, 47, 48, 49, 62, 63])
def test_wide_tile_scan_matches_wide_oracle(E, k):
    """the 128-bit helpers (window_sense_wide, canonical_wide, murmur3_multi_wide, window_ok_wide) as kmer_scan_wide_kernel drives
    them; parity beyond k = 32 is unpinned (the reference stops there), the yardstick is oracle kwo_raw_insert_wide"""
    nh, L = 1 + (k % 8), 16
    n_reads = 150
    flat = S.mutate(O.gen_reads(300 + k, 0, n_reads, 200), 300 + k, n_rate=97, lower_rate=5)
    bases, offsets = S.ragged(flat, 300 + k, n_reads, 0, 200)
    exp, n = O.raw_insert_wide(bases, offsets, k, nh, L)
    bits = np.zeros((1 << L) // 8, np.uint8)
    E.emu_scan_wide_insert.restype = C.c_uint64
    cnt = E.emu_scan_wide_insert(p(bases), C.c_uint64(len(bases)), p(offsets), C.c_uint64(n_reads), C.c_uint32(k), C.c_uint32(nh),
                                 C.c_uint32(L), p(bits))
    assert cnt == n and n > 0
    assert np.array_equal(bits, exp)


@pytest.mark.parametrize("k", [1, 2, 3, 7, 8, 15, 16, 17, 21, 30, 31, 32])
def test_window_ok_word_equals_window_ok(E, k):
    """the 32-positions-at-once window test (sliding OR by doubling) against the per-position one, on sparse and dense bitmaps"""
    rng = np.random.default_rng(k)
    for density in (0.0, 0.002, 0.02, 0.2, 1.0):
        bad = np.packbits(rng.random(64 * 32) < density, bitorder="little").view(np.uint32).copy()
        start = np.packbits(rng.random(64 * 32) < density * 0.5 + 0.005, bitorder="little").view(np.uint32).copy()
        assert E.emu_window_ok_word_check(p(bad), p(start), C.c_uint32(len(bad)), C.c_uint32(k)) == 0


def test_hash_of_reference_layout_word(E):
    for seq, k in S.HASH_KAT_INPUTS:
        words, _ = O.canonical_kmers(seq, k)
        for w in words[:5]:
            out = np.zeros(5, np.uint32)
            E.emu_hash_word(C.c_uint64(int(w)), C.c_uint32(k), p(out))
            assert list(out) == [O.murmur3_word(w, k, s) for s in range(5)]


def test_transpose32(E):
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = rng.integers(0, 2 ** 32, 32, dtype=np.uint64).astype(np.uint32)
        out = np.zeros(32, np.uint32)
        E.emu_transpose32(p(a), p(out))
        bits = (a[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1      # bits[i][b]
        exp = (bits.T.astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint32)
        assert np.array_equal(out, exp)


@pytest.mark.parametrize("n,nsub", [(0, 8), (1, 8), (15, 8), (16, 8), (17, 8), (970, 8), (4970, 8), (8192, 8), (3000, 64),
                                    (20000, 256), (32768, 32), (32768, 256)])
def test_bitsliced_counters(E, n, nsub):
    rng = np.random.default_rng(n + nsub)
    vecs = rng.integers(0, 2 ** 32, (max(n, 1), 4), dtype=np.uint64).astype(np.uint32)
    if n > 5000:
        vecs |= rng.integers(0, 2 ** 32, (n, 4), dtype=np.uint64).astype(np.uint32)   # dense: large counts
    counts = np.zeros(128, np.uint32)
    E.emu_count128(p(vecs), C.c_uint32(n), C.c_uint32(nsub), p(counts))
    exp = np.zeros(128, np.uint64)
    if n:
        b = (vecs[:n, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1
        exp = b.reshape(n, 128).sum(axis=0)
    assert np.array_equal(counts.astype(np.uint64), exp.astype(np.uint64))


@pytest.mark.parametrize("nsub,n_lanes", [(8, 2), (16, 1), (32, 1)])
def test_search_early_exit_never_loses_a_hit(E, nsub, n_lanes):
    """search_count_kernel<NH, true>: a chunk stops being read once the sum of the substreams' bounds (exact maximum of
    the partial counts + k-mers not looked at yet, possibly stale) falls below the count a hit needs.  Replayed on the
    host with the substreams taking turns in random order: whenever somebody stops, every column's TRUE count is below
    `need`; when a column reaches `need`, nobody stops and the counts are the full ones; and chunks without a hit do stop."""
    rng = np.random.default_rng(nsub * 7 + n_lanes)
    cols = 128 * n_lanes
    stopped_some = 0
    for trial in range(60):
        n = int(rng.choice([40, 200, 970, 970, 2000, 5000]))
        need = max(1, int(n * float(rng.choice([0.2, 0.5, 0.5, 0.9, 1.0]))))
        dens = float(rng.choice([0.02, 0.125, 0.3]))
        m = (rng.random((n, cols)) < dens)
        plant = int(rng.integers(0, 4))                     # 0: no hit; else a column just below / at / above `need`
        if plant:
            c = int(rng.integers(0, cols))
            k_m = min(n, max(0, need + (plant - 2)))
            m[:, c] = False
            rows = rng.choice(n, k_m, replace=False) if trial % 2 else np.arange(n - k_m, n)
            m[rows, c] = True
        true = m.sum(axis=0)
        vecs = np.packbits(m.reshape(n, cols // 32, 32), axis=2, bitorder="little").view(np.uint32).reshape(n, n_lanes, 4)
        order = rng.integers(0, nsub, 4 * ((n + 16 * nsub - 1) // (16 * nsub)) * nsub).astype(np.uint32)
        if trial % 3 == 0:
            order = np.sort(order)                          # one substream races ahead of the others
        counts = np.zeros(cols, np.uint32)
        n_stop = E.emu_count_exit(p(np.ascontiguousarray(vecs)), C.c_uint32(n), C.c_uint32(n_lanes), C.c_uint32(nsub), C.c_uint32(need),
                                  p(order), C.c_uint32(len(order)), p(counts))
        if true.max() >= need:
            assert n_stop == 0 and np.array_equal(counts, true), (trial, n, need)
        if n_stop:
            assert true.max() < need and counts.max() < need and np.all(counts <= true), (trial, n, need)
            stopped_some += 1
        else:
            assert np.array_equal(counts, true)
    assert stopped_some >= 5


@pytest.mark.parametrize("nsub", [8, 16, 32, 64, 256])
def test_bitsliced_counters_cannot_wrap(E, nsub):
    # a filter that holds EVERY k-mer of a long query used to wrap a substream's 10-plane counter (1024 -> 0)
    cap = E.emu_seg_cap(C.c_uint32(nsub))
    assert cap <= 32768 and cap // nsub <= 1023
    for n in (cap - 1, cap, cap + 1, 2 * cap + 1, 40000):
        vecs = np.full((n, 4), 0xFFFFFFFF, np.uint32)
        vecs[:, 3] &= np.uint32(0x7FFFFFFF)          # one column that never matches
        counts = np.zeros(128, np.uint32)
        E.emu_count128(p(vecs), C.c_uint32(n), C.c_uint32(nsub), p(counts))
        assert list(counts[:127]) == [n] * 127 and counts[127] == 0


def test_synth_generator_agrees(E):
    for seed, stream, ctr in [(0, 0, 0), (1, 2, 3), (12345, 999999, 1 << 40), (2 ** 64 - 1, 2 ** 63, 7)]:
        assert E.emu_synth_rnd(seed, stream, ctr) == O.lib().kwo_rnd(seed, stream, ctr) == int(S.rnd(seed, stream, ctr))


# ---------------------------------------------------------------------------------------------- crc32
@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("n_bytes", [4, 8, 60, 64, 68, 252, 256, 260, 8188, 8192, 8196, 16384, 16388, 3 * 16384 + 12,
                                     1 << 20, (1 << 20) + 8, 1024 * 8192 + 4, 1025 * 16384 + 64])
def test_crc32_tiling_tables_and_combine_equal_zlib(E, n_bytes, fast):
    # the device algorithm of crc32.cu replayed on the host with the tables the kernels use (crc_tables.h)
    import zlib
    E.emu_crc32.restype = C.c_uint32
    E.emu_crc32.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int]
    rng = np.random.default_rng(n_bytes + fast)
    a = rng.integers(0, 256, n_bytes, dtype=np.uint8)
    for crc_in in (0, 0xDEADBEEF):
        assert E.emu_crc32(p(a), n_bytes, crc_in, fast) == zlib.crc32(a.tobytes(), crc_in)
    z = np.zeros(n_bytes, np.uint8)
    assert E.emu_crc32(p(z), n_bytes, 0, fast) == zlib.crc32(z.tobytes())


def test_crc32_shift_tables_are_the_matrices(E):
    rng = np.random.default_rng(1)
    for v in [0, 1, 0x80000000, 0xFFFFFFFF] + [int(x) for x in rng.integers(0, 1 << 32, 200, dtype=np.uint64)]:
        assert E.emu_crc_shift_tables_agree(C.c_uint32(v)) == 1
