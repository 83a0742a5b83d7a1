"""GPU parity of the packed (NCBI 2na + not-a-base mask) entry points kwg_bloom_add_packed[_dev] and of the chunked host
feed of raw mode: the same reads as ASCII through the oracle (reference make_bloom.cpp:506-621, word.h:73-104) must give
the same filters and counts."""
import numpy as np
import pytest

from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S

pytestmark = pytest.mark.gpu


def ragged_case(seed, n_reads=400, max_len=170, n_rate=37, lower_rate=5):
    flat = S.mutate(O.gen_reads(seed, 0, n_reads, max_len), seed, n_rate=n_rate, lower_rate=lower_rate)
    return S.ragged(flat, seed, n_reads, 0, max_len)


@pytest.mark.parametrize("k,nh,L", [(31, 3, 20), (21, 5, 18), (32, 4, 22), (1, 2, 8), (4, 8, 10), (15, 7, 16)])
def test_packed_raw_insert_bit_exact(k, nh, L):
    bases, offsets = ragged_case(3000 + k * 7 + nh)        # N's and lower case inside
    packed, mask = capi.pack_2na(bases)
    assert mask is not None
    exp, n = O.raw_insert(bases, offsets, k, nh, L)
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_packed(packed, mask, offsets)
        assert b.num_valid() == n
        assert np.array_equal(b.finalize(), exp)


def test_packed_without_mask_and_unaligned_call_starts():
    """no mask (every base is ACGT); calls whose first read starts at any base offset: the bases between the previous
    16-base boundary and the first read of a call ride along and must not count (k = 1 .. 32, incl. windows that fit
    inside that lead-in)"""
    bases, offsets = S.ragged(O.gen_reads(99, 0, 300, 90), 99, 300, 0, 90)
    packed, mask = capi.pack_2na(bases)
    assert mask is None
    n = len(offsets) - 1
    for k, nh, L in ((31, 3, 18), (3, 2, 10), (1, 1, 6), (16, 4, 14)):
        exp, n_exp = O.raw_insert(bases, offsets, k, nh, L)
        with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
            cuts = [0, 1, 2, 77, 78, 200, n]
            for a, z in zip(cuts[:-1], cuts[1:]):
                b.add_packed(packed, None, offsets[a: z + 1])
            assert b.num_valid() == n_exp
            assert np.array_equal(b.finalize(), exp)
        assert any(int(offsets[c]) % 16 for c in cuts[1:-1])


@pytest.mark.parametrize("c", [1, 3])
def test_packed_counting_matches_oracle(c):
    flat = O.gen_reads(515, 0, 700, 120)
    if c > 1:                                   # recurring k-mers so that the threshold lets some through
        flat = np.concatenate([flat] + [flat[: 200 * 120]] * (c + 1))
    flat = S.mutate(flat, 515, n_rate=301, lower_rate=7)
    n_reads = len(flat) // 120
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(120)
    packed, mask = capi.pack_2na(flat)
    lc = O.counting_log2_len(int(offsets[-1]))
    ob = O.Builder(31, c, lc, 24)
    ob.add_reads(flat, offsets)
    with capi.BloomBuilder(31, min_kmer_count=c, log2_count_len=lc, log2_max_len=24) as b:
        half = n_reads // 2 + 1
        b.add_packed(packed, mask, offsets[: half + 1])
        b.add_packed(packed, mask, offsets[half:])
        assert b.num_valid() == ob.num_valid()
        assert np.array_equal(b.finalize(22, 3), ob.finalize(22, 3))
    ob.close()


def test_chunked_host_feed_raw_and_packed_large_batch():
    """batches of more than 16 Mi bases travel piece by piece behind the scan (raw mode: new this round; packed: pieces of
    the 2na stream): same filter as the oracle's, ASCII and packed"""
    n_reads, rl, k, nh, L = 150_000, 150, 31, 3, 27
    bases = S.mutate(O.gen_reads(777, 0, n_reads, rl), 777, n_rate=5003, lower_rate=11)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(rl)
    assert len(bases) > (16 << 20)
    exp, n = O.raw_insert(bases, offsets, k, nh, L)
    packed, mask = capi.pack_2na(bases)
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_reads(bases, offsets)
        assert b.num_valid() == n
        assert np.array_equal(b.finalize(), exp)
        b.reset()
        b.add_packed(packed, mask, offsets)
        assert b.num_valid() == n
        assert np.array_equal(b.finalize(), exp)
    # counting mode through the packed feed: the ASCII path (itself checked against the oracle) is the yardstick here
    lc = O.counting_log2_len(int(offsets[-1]))
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=30) as b:
        b.add_reads(bases, offsets)
        n_a = b.num_valid()
        bits_a = b.finalize(26, 3)
        b.reset()
        b.add_packed(packed, mask, offsets)
        assert b.num_valid() == n_a
        assert np.array_equal(b.finalize(26, 3), bits_a)


def test_packed_dev_entry_point():
    import torch
    bases, offsets = ragged_case(4242, n_reads=900)
    packed, mask = capi.pack_2na(bases)
    exp, n = O.raw_insert(bases, offsets, 31, 3, 20)
    d_p = torch.from_numpy(np.concatenate([packed, np.zeros(16, np.uint8)])).cuda()
    d_m = torch.from_numpy(np.concatenate([mask, np.zeros(16, np.uint8)])).cuda()
    d_o = torch.from_numpy(offsets.view(np.int64)).cuda()
    with capi.BloomBuilder(31, raw_num_hash=3, raw_log2_len=20) as b:
        b.add_packed_dev(d_p.data_ptr(), d_m.data_ptr(), d_o.data_ptr(), len(offsets) - 1, int(offsets[-1]))
        assert b.num_valid() == n
        assert np.array_equal(b.finalize(), exp)
