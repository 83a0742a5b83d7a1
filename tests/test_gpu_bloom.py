"""GPU parity: Bloom construction through the C ABI vs the oracle and the reference's golden digests."""
import numpy as np
import pytest

from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S
import util
from conftest import load_golden

pytestmark = pytest.mark.gpu


def ragged_case(seed, n_reads=400, max_len=170, n_rate=37, lower_rate=5):
    flat = S.mutate(O.gen_reads(seed, 0, n_reads, max_len), seed, n_rate=n_rate, lower_rate=lower_rate)
    return S.ragged(flat, seed, n_reads, 0, max_len)


@pytest.mark.parametrize("k,nh,L", [(31, 3, 20), (31, 1, 12), (21, 5, 18), (32, 4, 22), (1, 2, 8), (4, 8, 10), (15, 7, 16),
                                    (25, 3, 26), (31, 3, 29), (17, 2, 32), (31, 3, 30), (21, 5, 31)])
def test_raw_insert_bit_exact(k, nh, L):
    bases, offsets = ragged_case(1000 + k * 7 + nh)
    exp, n = O.raw_insert(bases, offsets, k, nh, L)
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_reads(bases, offsets)
        assert b.num_valid() == n
        got = b.finalize()
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("k,nh,L", [(33, 3, 20), (40, 2, 14), (47, 5, 22), (48, 8, 18), (62, 1, 12), (63, 3, 29), (63, 4, 30)])
def test_raw_insert_wide_k_matches_the_wide_restatement(k, nh, L):
    """k in 33..63 (raw mode only; BASELINE.json configs[4]).  PARITY UNPINNED: the reference stops at k = 32 (word.h:10);
    the yardstick is oracle kwo_raw_insert_wide, which equals the pinned narrow restatement for k <= 32 and an independent
    pure-Python statement above (tests/test_oracle_wide.py).  ASCII and packed input, reads across several tiles."""
    flat = S.mutate(O.gen_reads(2000 + k, 0, 600, 260), 2000 + k, n_rate=211, lower_rate=5)
    bases, offsets = S.ragged(flat, 2000 + k, 600, 0, 260)
    exp, n = O.raw_insert_wide(bases, offsets, k, nh, L)
    assert n > 0
    packed, mask = capi.pack_2na(bases)
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_reads(bases, offsets)
        assert b.num_valid() == n
        assert np.array_equal(b.finalize(), exp)
        b.reset()
        b.add_packed(packed, mask, offsets)
        assert b.num_valid() == n
        assert np.array_equal(b.finalize(), exp)


def test_wide_k_is_refused_outside_raw_mode():
    with pytest.raises(capi.KwageError):
        capi.BloomBuilder(33, min_kmer_count=1, log2_count_len=20, log2_max_len=24)
    with pytest.raises(capi.KwageError):
        capi.BloomBuilder(64, raw_num_hash=3, raw_log2_len=20)


def test_raw_insert_accumulates_over_calls_and_reset():
    k, nh, L = 31, 3, 22
    b1, o1 = S.uniform_reads(7, 0, 3000, 150)
    b2, o2 = ragged_case(8, n_reads=900)
    exp, n1 = O.raw_insert(b1, o1, k, nh, L)
    exp, n2 = O.raw_insert(b2, o2, k, nh, L, bits=exp)
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_reads(b1, o1)
        b.add_reads(b2, o2)
        assert b.num_valid() == n1 + n2
        assert np.array_equal(b.finalize(), exp)
        b.reset()
        assert b.num_valid() == 0
        b.add_reads(b2, o2)
        exp2, _ = O.raw_insert(b2, o2, k, nh, L)
        assert np.array_equal(b.finalize(), exp2)


def test_raw_insert_edge_inputs():
    k, nh, L = 31, 3, 16
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_reads(np.zeros(0, np.uint8), np.zeros(1, np.uint64))                 # no reads
        b.add_reads(np.frombuffer(b"ACGT", np.uint8), np.array([0, 0, 4, 4], np.uint64))   # empty + too-short reads
        assert b.num_valid() == 0 and not b.finalize().any()
        seq = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACG", np.uint8)          # exactly one k-mer
        b.add_reads(seq, np.array([0, 31], np.uint64))
        exp, n = O.raw_insert(seq, np.array([0, 31], np.uint64), k, nh, L)
        assert n == 1 and b.num_valid() == 1 and np.array_equal(b.finalize(), exp)
        # offsets that do not start at zero
        b.reset()
        two = np.concatenate([seq, seq[::-1].copy()])
        b.add_reads(two, np.array([31, 62], np.uint64))
        exp, _ = O.raw_insert(two, np.array([31, 62], np.uint64), k, nh, L)
        assert np.array_equal(b.finalize(), exp)


def run_counting(case, bases, offsets, split=None):
    lc = O.counting_log2_len(case["num_bp"])
    with capi.BloomBuilder(case["k"], min_kmer_count=case["min_count"], log2_count_len=lc, log2_max_len=case["lmax"]) as b:
        if split is None:
            b.add_reads(bases, offsets)
        else:   # the same stream delivered in several calls (the reference adds one fragment at a time)
            n = len(offsets) - 1
            cuts = [0] + [n * i // split for i in range(1, split)] + [n]
            for a, z in zip(cuts[:-1], cuts[1:]):
                b.add_reads(bases, offsets[a: z + 1])
        n_valid = b.num_valid()
        param = O.optimal_bloom_param(n_valid, case["p"], case["lmin"], case["lmax"])
        bits = b.finalize(*param) if param else None
    return lc, n_valid, param, bits


@pytest.mark.parametrize("name", ["uniform_k31", "ragged_k21", "k32", "k15_dups", "small_count_filter", "cfg1_mt64",
                                  "min_count_2", "min_count_5"])
def test_counting_mode_matches_reference_golden(name):
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    lc, n_valid, param, bits = run_counting(case, bases, offsets)
    assert lc == g["log2_count_len"]
    assert n_valid == g["num_kmer"]
    assert list(param) == [g["log2_len"], g["num_hash"]]
    assert O.crc32(bits) == g["bits_crc32"]
    assert util.sha256(bits) == g["bits_sha256"]


@pytest.mark.parametrize("name,split", [("uniform_k31", None), ("ragged_k21", 7), ("small_count_filter", 3), ("k15_dups", None), ("cfg1_mt64", None)])
def test_counting_mode_two_level_radix_path_stays_exact(name, split, monkeypatch):
    """min_kmer_count 1 takes the one-level first-touch path (bloom_first.cuh) up to log2_count_len 30 and the two-level
    radix partition (bloom_count.cuh) above; KWG_COUNT_TWO_LEVEL (read when the handle is created) forces the latter so
    that it stays covered at the sizes the goldens have"""
    monkeypatch.setenv("KWG_COUNT_TWO_LEVEL", "1")
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    _, n_valid, param, bits = run_counting(case, bases, offsets, split=split)
    assert n_valid == g["num_kmer"] and util.sha256(bits) == g["bits_sha256"]


@pytest.mark.parametrize("name,split", [("ragged_k21", 7), ("small_count_filter", 3), ("k15_dups", 40), ("min_count_2", 5),
                                        ("min_count_5", 3)])
def test_counting_mode_is_stream_order_exact_across_calls(name, split):
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    _, n_valid, param, bits = run_counting(case, bases, offsets, split=split)
    assert n_valid == g["num_kmer"] and util.sha256(bits) == g["bits_sha256"]


def test_counting_mode_finalize_any_parameters_equals_fold():
    # finalize(L, h) must equal the reference's fold for ANY (L <= Lmax, h <= 5), not just the optimum
    case = dict(S.MAKE_BLOOM_CASES["ragged_k21"])
    bases, offsets = S.make_bloom_reads(case)
    lc = O.counting_log2_len(case["num_bp"])
    ob = O.Builder(case["k"], 1, lc, case["lmax"])
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(case["k"], min_kmer_count=1, log2_count_len=lc, log2_max_len=case["lmax"]) as b:
        b.add_reads(bases, offsets)
        assert b.num_valid() == ob.num_valid()
        for L, h in [(18, 1), (19, 5), (22, 3), (26, 2)]:
            assert np.array_equal(b.finalize(L, h), ob.finalize(L, h)), (L, h)
    ob.close()


@pytest.mark.parametrize("no_fold", [False, True])
@pytest.mark.parametrize("name,split", [("small_count_filter", 3), ("k15_dups", None), ("ragged_k21", 2)])
def test_finalize_seed_pairs_from_the_touched_bitmap(name, split, no_fold, monkeypatch):
    """min_kmer_count 1, first-touch path: finalize takes the seed pairs (0,1) and (2,3) out of the counting tables' touched
    bitmap when L <= log2_count_len (fold_touched_kernel) and only a last odd seed out of the word list.  Every num_hash,
    L below / at / above the counting-filter length, accessions added in several calls (the bitmap accumulates), and
    the path that sets every bit from the word list (KWG_NO_FOLD, read when the handle is created)."""
    if no_fold:
        monkeypatch.setenv("KWG_NO_FOLD", "1")
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    lc = O.counting_log2_len(case["num_bp"])
    lmax = max(case["lmax"], lc + 1)
    ob = O.Builder(case["k"], 1, lc, lmax)
    ob.add_reads(bases, offsets)
    n = len(offsets) - 1
    cuts = [0, n] if not split else [0] + [n * (i + 1) // split for i in range(split)]
    with capi.BloomBuilder(case["k"], min_kmer_count=1, log2_count_len=lc, log2_max_len=lmax) as b:
        for a, z in zip(cuts[:-1], cuts[1:]):
            if z > a:
                b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == ob.num_valid()
        for L in (5, 6, 11, lc - 1, lc, lc + 1):
            for h in (1, 2, 3, 4, 5):
                assert np.array_equal(b.finalize(L, h), ob.finalize(L, h)), (L, h)
        # an accession without k-mers after a reset: an all-zero filter, not a stale bitmap
        b.reset()
        b.add_reads(np.frombuffer(b"ACGTNNACGT", dtype=np.uint8), np.array([0, 10], dtype=np.uint64))
        assert b.num_valid() == 0 and not b.finalize(12, 4).any()
    ob.close()


@pytest.mark.parametrize("c,name", [(1, "ragged_k21"), (1, "small_count_filter"), (3, "k15_dups")])
@pytest.mark.parametrize("two_level", [False, True])
def test_checkpoint_and_rollback_restore_the_counting_state(c, name, two_level, monkeypatch):
    """kwg_bloom_checkpoint / kwg_bloom_rollback (what the host needs to stop at the fragment that crosses max_num_kmer,
    make_bloom.cpp:208-214): after a rollback the handle behaves as if the batches since the checkpoint had never been
    added -- counts and filters equal the oracle's for every continuation, on all three construction paths."""
    if two_level:
        monkeypatch.setenv("KWG_COUNT_TWO_LEVEL", "1")
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    lc = O.counting_log2_len(case["num_bp"])
    n = len(offsets) - 1
    a, z = n // 3, 2 * n // 3

    def oracle(parts):
        ob = O.Builder(case["k"], c, lc, case["lmax"])
        for p0, p1 in parts:
            ob.add_reads(bases, offsets[p0: p1 + 1])
        out = (ob.num_valid(), ob.finalize(20, 3), ob.finalize(19, 4))
        ob.close()
        return out

    with capi.BloomBuilder(case["k"], min_kmer_count=c, log2_count_len=lc, log2_max_len=case["lmax"]) as b:
        with pytest.raises(capi.KwageError):
            b.rollback()                                     # nothing to go back to
        b.checkpoint()                                       # before the first batch
        b.add_reads(bases, offsets[0: a + 1])
        b.rollback()
        assert b.num_valid() == 0 and not b.finalize(12, 2).any()
        b.add_reads(bases, offsets[0: a + 1])
        b.checkpoint()
        b.add_reads(bases, offsets[a: z + 1])
        n_all = b.num_valid()
        for cont in ([(a, a + 7)], [(z, n)], [(a, z), (z, n)]):
            b.rollback()
            for p0, p1 in cont:
                b.add_reads(bases, offsets[p0: p1 + 1])
            exp = oracle([(0, a)] + cont)
            assert b.num_valid() == exp[0], cont
            assert np.array_equal(b.finalize(20, 3), exp[1]) and np.array_equal(b.finalize(19, 4), exp[2]), cont
        assert n_all == oracle([(0, a), (a, z)])[0]


def test_filters_beyond_one_l2_window_are_filled_window_by_window():
    # filters of more than 2^29 bits are filled one 64 MiB window per pass (raw scan and counting-mode finalize alike)
    case = dict(S.MAKE_BLOOM_CASES["uniform_k31"])
    bases, offsets = S.make_bloom_reads(case)
    lc = O.counting_log2_len(case["num_bp"])
    ob = O.Builder(case["k"], 1, lc, 32)
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(case["k"], min_kmer_count=1, log2_count_len=lc, log2_max_len=32) as b:
        b.add_reads(bases, offsets)
        for L, h in [(30, 3), (32, 5)]:
            got = b.finalize(L, h)
            assert np.array_equal(got, ob.finalize(L, h)), (L, h)
            del got
    ob.close()


def skewed_case(seed, n_reads=6000, poly=1500):
    """Random ragged reads + a block of identical poly-A reads + repeats of the first reads: heavy
    duplication puts hundreds of thousands of records into the same few partition buckets."""
    bases, offsets = ragged_case(seed, n_reads=n_reads)
    poly_a = np.full(poly * 150, ord("A"), np.uint8)
    rep = bases[: int(offsets[n_reads // 4])]
    flat = np.concatenate([bases, poly_a, rep, poly_a[: 150 * 7]])
    offs = np.concatenate([offsets,
                           offsets[-1] + np.arange(1, poly + 1, dtype=np.uint64) * np.uint64(150),
                           offsets[-1] + np.uint64(poly * 150) + offsets[1: n_reads // 4 + 1],
                           offsets[-1] + np.uint64(poly * 150) + offsets[n_reads // 4] + np.arange(1, 8, dtype=np.uint64) * np.uint64(150)])
    return flat, offs.astype(np.uint64)


@pytest.mark.parametrize("lc,split", [(18, 1), (23, 2), (24, 1), (25, 3), (27, 1), (30, 2), (32, 1)])
def test_counting_mode_every_partition_geometry(lc, split):
    # log2_count_len decides the partition geometry (single level up to 23, two levels above); the
    # result must equal the oracle's sequential counting filters for every one of them
    k, lmax = 31, 24
    bases, offsets = skewed_case(50 + lc)
    ob = O.Builder(k, 1, lc, lmax)
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=lmax) as b:
        n = len(offsets) - 1
        cuts = [0] + [n * i // split for i in range(1, split)] + [n]
        for a, z in zip(cuts[:-1], cuts[1:]):
            b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == ob.num_valid()
        for L, h in [(22, 3), (24, 5)]:
            assert np.array_equal(b.finalize(L, h), ob.finalize(L, h)), (L, h)
    ob.close()


@pytest.mark.parametrize("k,lc,n_reads", [(32, 22, 5000), (31, 23, 20000), (31, 21, 20000), (31, 20, 30000)])
def test_counting_mode_many_tiles_per_bucket(k, lc, n_reads):
    # single-level geometries with more tiles than one staging window holds: the resolve kernel gathers a
    # bucket's runs in several rounds (and, beyond 1024 tiles, several run-table passes)
    bases, offsets = S.uniform_reads(13, 0, n_reads, 150)
    ob = O.Builder(k, 1, lc, 24)
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=24) as b:
        b.add_reads(bases, offsets)
        assert b.num_valid() == ob.num_valid()
        assert np.array_equal(b.finalize(24, 3), ob.finalize(24, 3))
    ob.close()


@pytest.mark.parametrize("n_reads", [2400, 2500, 2600, 2700])
def test_staged_records_end_in_an_overhang(n_reads):
    # lc = 21: 128 buckets, one run of ~64 records per (bucket, tile) -- about half of them "long" (read directly), half
    # staged.  Around 184 tiles the staged half of a bucket just exceeds one staging window; when the run that overhangs the
    # window is the last one, the second round has nothing to stage and must resolve nothing (it used to resolve the first
    # window again: 576 k-mers lost on this input).
    case = dict(kind="coverage", seed=884924425, genome=200000, n_reads=4000, read_len=150, num_bp=-1)
    bases, offsets = S.make_bloom_reads(case)
    for c in (1, 2):
        ob = O.Builder(31, c, 21, 24)
        ob.add_reads(bases, offsets[: n_reads + 1])
        with capi.BloomBuilder(31, min_kmer_count=c, log2_count_len=21, log2_max_len=24) as b:
            b.add_reads(bases, offsets[: n_reads + 1])
            assert b.num_valid() == ob.num_valid()
            assert np.array_equal(b.finalize(24, 3), ob.finalize(24, 3))
        ob.close()


def test_counting_mode_small_filter_heavy_shadowing():
    # many more k-mers than counting slots: most first occurrences are shadowed by earlier k-mers, so
    # the stream-order rule (minimum position per slot, displacement of later occurrences) decides almost
    # every k-mer
    k, lc, lmax = 21, 18, 22
    bases, offsets = S.uniform_reads(77, 0, 20000, 100)
    ob = O.Builder(k, 1, lc, lmax)
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=lmax) as b:
        b.add_reads(bases, offsets[:7001])
        b.add_reads(bases, offsets[7000:])
        assert b.num_valid() == ob.num_valid()
        assert b.num_valid() < 20000 * 80 // 2
        assert np.array_equal(b.finalize(22, 2), ob.finalize(22, 2))
    ob.close()


def test_counting_mode_reset_and_invalid_statuses():
    case = dict(S.MAKE_BLOOM_CASES["no_kmers"])
    bases, offsets = S.make_bloom_reads(case)
    lc = O.counting_log2_len(case["num_bp"])
    with capi.BloomBuilder(case["k"], min_kmer_count=1, log2_count_len=lc, log2_max_len=case["lmax"]) as b:
        b.add_reads(bases, offsets)
        assert b.num_valid() == 0            # -> optimal_bloom_param throws -> STATUS_BLOOM_INVALID on host
        c2 = dict(S.MAKE_BLOOM_CASES["uniform_k31"])
        b2, o2 = S.make_bloom_reads(c2)
        b.add_reads(b2, o2)
        n_a = b.num_valid()
        b.reset()
        b.add_reads(b2, o2)
        assert b.num_valid() == n_a


# ------------------------------------------------------------------------------ min_kmer_count > 1
# k = 31 k-mers whose murmur3 values for seeds 0 and 1 agree in their low 18 bits: with log2_count_len = 18 both
# hashes of the first counting filter fall on one slot, which the reference increments twice per visit
# (make_bloom.cpp:553-554,586-592)
DUP_KMERS = [b"CTGATCATTGGGTAATACATTGAAGGCCCAT", b"CCTTGATATGCCACGCAAGCTCGTCTTCCAA"]


def test_dup_kmers_really_collide():
    for s in DUP_KMERS:
        w, _ = O.canonical_kmers(s, 31)
        assert (O.murmur3_word(w[0], 31, 0) ^ O.murmur3_word(w[0], 31, 1)) & 0x3FFFF == 0


def reads_from_list(reads):
    lens = np.array([len(r) for r in reads], np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    return np.concatenate([np.frombuffer(bytes(r), np.uint8) for r in reads]), offsets


def abundance_case(seed, genome=60000, n_reads=9000, read_len=100, dup_copies=(11, 6)):
    """Coverage ~15x cut by non-ACGT bytes (k-mer abundances spread around 10, on both sides of every threshold
    tested), plus the slot-doubling k-mers as stand-alone reads scattered through the stream."""
    case = dict(kind="coverage", seed=seed, genome=genome, n_reads=n_reads, read_len=read_len, num_bp=-1)
    bases, offsets = S.make_bloom_reads(case)
    bases = S.mutate(bases, seed, n_rate=211, lower_rate=9)
    reads = [bases[int(offsets[i]): int(offsets[i + 1])] for i in range(n_reads)]
    for kmer, copies in zip(DUP_KMERS, dup_copies):
        for j in range(copies):
            at = int(S.rnd(seed, 0xD0 + copies, np.array([j], np.uint64))[0] % np.uint64(len(reads)))
            reads.insert(at, np.frombuffer(kmer, np.uint8))
    return reads_from_list(reads)


def check_against_oracle(bases, offsets, k, c, lc, lmax, split=1, params=((22, 3), (24, 5))):
    ob = O.Builder(k, c, lc, lmax)
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(k, min_kmer_count=c, log2_count_len=lc, log2_max_len=lmax) as b:
        n = len(offsets) - 1
        cuts = [0] + [n * i // split for i in range(1, split)] + [n]
        for a, z in zip(cuts[:-1], cuts[1:]):
            b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == ob.num_valid()
        for L, h in params:
            assert np.array_equal(b.finalize(L, h), ob.finalize(L, h)), (L, h)
        n_valid = b.num_valid()
    ob.close()
    return n_valid


@pytest.mark.parametrize("c,lc,split", [(2, 18, 1), (2, 24, 3), (3, 18, 4), (3, 30, 1), (5, 18, 1), (5, 20, 2), (5, 23, 1), (5, 25, 5),
                                        (5, 30, 2), (8, 18, 3), (8, 27, 1), (14, 18, 2), (15, 18, 1), (15, 24, 2), (15, 32, 1)])
def test_min_kmer_count_levels_match_sequential_counters(c, lc, split):
    # conservative-update counters resolved level by level must equal the reference's sequential loop for every
    # threshold, partition geometry and batching; lc = 18 makes slot collisions (and weight-2 records) common
    bases, offsets = abundance_case(100 + c)
    n_valid = check_against_oracle(bases, offsets, 31, c, lc, 24, split)
    assert n_valid > 1000


@pytest.mark.parametrize("c", [2, 5, 9])
def test_min_kmer_count_on_skewed_input(c):
    # poly-A block + repeated reads: runs of >64 records (read straight from HBM) and multi-round buckets
    bases, offsets = skewed_case(300 + c)
    check_against_oracle(bases, offsets, 31, c, 22, 24, split=2)


@pytest.mark.parametrize("c,copies", [(2, 1), (2, 2), (2, 3), (3, 2), (3, 3), (3, 4), (6, 5), (6, 6), (14, 13), (14, 14), (15, 14)])
def test_weight2_records_follow_the_double_increment(c, copies):
    # one k-mer whose two first-filter hashes share a slot: the slot goes 0 -> 2 -> 2 -> 4 ... while the second
    # filter's slots go up by one, so the k-mer turns valid at a visit that depends on the double increments
    reads = [np.frombuffer(DUP_KMERS[0], np.uint8)] * copies + [np.frombuffer(DUP_KMERS[1], np.uint8)] * (copies - 1)
    bases, offsets = reads_from_list(reads)
    check_against_oracle(bases, offsets, 31, c, 18, 24, params=((20, 3),))
    check_against_oracle(bases, offsets, 31, c, 18, 24, split=max(1, copies // 2), params=((20, 5),))


def test_counter_wrap_is_reported_not_mimicked():
    # min_kmer_count 15: the 15th visit of the slot-doubling k-mer takes its 4-bit counter from 14 to 16 = 0 in the
    # reference (bloom.h 4-bit field).  That order-dependent wrap is the one case the device refuses, loudly.
    bases, offsets = reads_from_list([np.frombuffer(DUP_KMERS[0], np.uint8)] * 15)
    with capi.BloomBuilder(31, min_kmer_count=15, log2_count_len=18, log2_max_len=24) as b:
        b.add_reads(bases, offsets)
        with pytest.raises(capi.KwageError) as e:
            b.num_valid()
        assert e.value.code == capi.KWG_ERR_UNSUPPORTED and "wrapped" in str(e.value)
    # the same stream with a different counting-filter length has no shared slot and no wrap
    check_against_oracle(bases, offsets, 31, 15, 19, 24, params=((20, 3),))


def test_min_kmer_count_reset_between_accessions():
    bases, offsets = abundance_case(7, n_reads=3000)
    ob = O.Builder(31, 4, 20, 24)
    ob.add_reads(bases, offsets)
    with capi.BloomBuilder(31, min_kmer_count=4, log2_count_len=20, log2_max_len=24) as b:
        b.add_reads(bases, offsets[:1501])
        b.reset()
        b.add_reads(bases, offsets)
        assert b.num_valid() == ob.num_valid()
        assert np.array_equal(b.finalize(24, 3), ob.finalize(24, 3))
    ob.close()


def test_unsupported_and_bad_arguments_raise():
    with pytest.raises(capi.KwageError) as e:
        capi.BloomBuilder(31, min_kmer_count=16, log2_count_len=20, log2_max_len=24)
    assert e.value.code == capi.KWG_ERR_INVALID_ARG
    with pytest.raises(capi.KwageError):
        capi.BloomBuilder(64, raw_num_hash=3, raw_log2_len=20)               # raw mode goes up to 63 (an extension), no further
    with pytest.raises(capi.KwageError):
        capi.BloomBuilder(33, min_kmer_count=1, log2_count_len=20, log2_max_len=24)      # counting mode: the reference's limit
    with capi.BloomBuilder(31, raw_num_hash=3, raw_log2_len=20) as b:
        with pytest.raises(capi.KwageError):
            b.add_reads(np.zeros(10, np.uint8), np.array([5, 2], np.uint64))      # decreasing offsets
