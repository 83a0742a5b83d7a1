"""GPU properties at the benchmark's full accession size (BASELINE.json configs[1]: 1e6 x 150 bp reads, 1.2e8 k-mer
occurrences, lc = 30 -> 65,536 final buckets, two-level partition, chunked host feed).  The oracle needs minutes at
this size, so these tests use properties that hold for any input:
  * the result does not depend on how the stream is cut into add_reads calls (stream-order exactness);
  * adding reads that were already added changes nothing (every one of their touches finds its slot taken);
  * every bit of the counting-mode filter is a bit of the raw-mode filter (same L, h), and the k-mers the counting
    filters dropped are exactly the difference between the number of distinct k-mers and num_valid;
  * a smaller accession embedded at the start of the stream gives the oracle's exact num_valid for that prefix."""
import numpy as np
import pytest

from kwage_b200 import capi, hostapi as H
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

K, N_READS, READ_LEN, LMAX = 31, 1_000_000, 150, 32


@pytest.fixture(scope="module")
def accession():
    bases = O.gen_reads(20260101, 0, N_READS, READ_LEN)
    offsets = np.arange(N_READS + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    return bases, offsets


def test_full_size_accession_is_split_and_repeat_invariant(accession):
    bases, offsets = accession
    lc = H.counting_filter_log2_len(int(offsets[-1]))
    assert lc == 30
    with capi.BloomBuilder(K, min_kmer_count=1, log2_count_len=lc, log2_max_len=LMAX) as b:
        b.add_reads(bases, offsets)
        n_one = b.num_valid()
        L, h = H.optimal_bloom_param(K, n_one, 0.25, 18, LMAX)
        assert (L, h) == (29, 3)
        bits_one = b.finalize(L, h)
        # the same reads again: nothing may change
        b.add_reads(bases, offsets[: N_READS // 3 + 1])
        assert b.num_valid() == n_one
        assert np.array_equal(b.finalize(L, h), bits_one)
        # a fresh accession fed in five uneven calls
        b.reset()
        cuts = [0, 1, 99_999, 400_000, 400_001, N_READS]
        for a, z in zip(cuts[:-1], cuts[1:]):
            b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == n_one
        assert np.array_equal(b.finalize(L, h), bits_one)
    # uniform random 150-mers: practically every 31-mer is distinct, the counting filters shadow ~3e-4 of them
    n_kmers = N_READS * (READ_LEN - K + 1)
    assert 0.9990 * n_kmers < n_one < n_kmers

    with capi.BloomBuilder(K, raw_num_hash=h, raw_log2_len=L) as r:
        r.add_reads(bases, offsets)
        assert r.num_valid() == n_kmers
        raw = r.finalize()
    assert not np.any(bits_one & ~raw), "counting-mode filter has a bit the raw filter lacks"
    missing = int(np.unpackbits(raw & ~bits_one).sum())
    assert 0 < missing <= h * (n_kmers - n_one)


def test_full_geometry_prefix_matches_oracle(accession):
    # the full-size geometry (lc = 30) on a prefix the oracle can do in seconds
    bases, offsets = accession
    n = 60_000
    ob = O.Builder(K, 1, 30, 26)
    ob.add_reads(bases, offsets[: n + 1])
    with capi.BloomBuilder(K, min_kmer_count=1, log2_count_len=30, log2_max_len=26) as b:
        b.add_reads(bases, offsets[: n // 2 + 1])
        b.add_reads(bases, offsets[n // 2: n + 1])
        assert b.num_valid() == ob.num_valid()
        assert np.array_equal(b.finalize(26, 4), ob.finalize(26, 4))
    ob.close()


# ------------------------------------------------------------------------------ min_kmer_count 5 (the reference default)
@pytest.fixture(scope="module")
def covered_accession():
    """1e6 reads sampled from a 5 Mb random genome: ~30x coverage, so most k-mers pass the threshold of 5."""
    genome = O.gen_reads(20260202, 0, 1, 5_000_000)
    rng = np.random.default_rng(20260203)
    starts = rng.integers(0, len(genome) - READ_LEN, N_READS)
    bases = np.empty(N_READS * READ_LEN, np.uint8)
    step = 100_000
    for a in range(0, N_READS, step):
        idx = starts[a: a + step, None] + np.arange(READ_LEN)[None, :]
        bases[a * READ_LEN: (a + step) * READ_LEN] = genome[idx].reshape(-1)
    offsets = np.arange(N_READS + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    return bases, offsets


def test_full_size_min_count_5_is_split_invariant_and_matches_oracle_prefix(covered_accession):
    bases, offsets = covered_accession
    with capi.BloomBuilder(K, min_kmer_count=5, log2_count_len=30, log2_max_len=LMAX) as b:
        b.add_reads(bases, offsets)
        n_one = b.num_valid()
        # every 31-mer of the genome seen at least 5 times turns valid exactly once: a little under 5e6 of them
        assert 4_500_000 < n_one < 5_000_000
        L, h = H.optimal_bloom_param(K, n_one, 0.25, 18, LMAX)
        bits_one, crc = b.finalize_crc(L, h)
        import zlib
        assert crc == zlib.crc32(bits_one.tobytes())
        # the counters of the first pass persist: the same reads again add nothing new (all counters are >= 5 or the
        # k-mer stays below the threshold only if it still has fewer than 5 occurrences -- so count again from scratch)
        b.reset()
        cuts = [0, 3, 250_000, 250_001, 777_777, N_READS]
        for a, z in zip(cuts[:-1], cuts[1:]):
            b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == n_one
        assert np.array_equal(b.finalize(L, h), bits_one)
    # the full-size geometry on a prefix the oracle does in seconds, thresholds 2 and 5
    n = 80_000
    for c in (2, 5):
        ob = O.Builder(K, c, 30, 26)
        ob.add_reads(bases, offsets[: n + 1])
        with capi.BloomBuilder(K, min_kmer_count=c, log2_count_len=30, log2_max_len=26) as b:
            b.add_reads(bases, offsets[: n // 3 + 1])
            b.add_reads(bases, offsets[n // 3: n + 1])
            assert b.num_valid() == ob.num_valid()
            assert np.array_equal(b.finalize(24, 3), ob.finalize(24, 3))
        ob.close()


# ------------------------------------------------------------------------------ more than 2^28 start positions in one call
@pytest.mark.parametrize("c", [1, 3])
def test_device_batch_longer_than_one_sub_batch(c):
    """kwg_bloom_add_reads_dev cuts a batch into sub-batches of 2^28 start positions (a touch record carries a 28-bit
    position); reads straddle the cut and the counting-filter state has to carry over.  The same 3.3e8 bases delivered in
    one call and in three calls cut elsewhere must give the same filter."""
    import torch
    n_reads, glen = 2_200_000, 40_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(4711)
    genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (glen,), generator=g, device="cuda")]
    starts = torch.randint(0, glen - READ_LEN, (n_reads,), generator=g, device="cuda")
    bases = torch.empty(n_reads * READ_LEN + 16, dtype=torch.uint8, device="cuda")
    for a in range(0, n_reads, 100_000):
        z = min(n_reads, a + 100_000)
        idx = starts[a:z, None] + torch.arange(READ_LEN, device="cuda")[None, :]
        bases[a * READ_LEN: z * READ_LEN] = genome[idx].reshape(-1)
    del genome, idx
    offsets = torch.arange(n_reads + 1, dtype=torch.int64, device="cuda") * READ_LEN
    n_bases = n_reads * READ_LEN
    assert n_bases > (1 << 28) and (1 << 28) % READ_LEN != 0          # the cut falls inside a read
    lc = H.counting_filter_log2_len(n_bases)
    torch.cuda.synchronize()
    with capi.BloomBuilder(K, min_kmer_count=c, log2_count_len=lc, log2_max_len=LMAX) as b:
        b.add_reads_dev(bases.data_ptr(), offsets.data_ptr(), n_reads, n_bases)
        n_one = b.num_valid()
        L, h = H.optimal_bloom_param(K, n_one, 0.25, 18, LMAX)
        bits_one = b.finalize(L, h)
        b.reset()
        cuts = [0, 700_000, 1_500_008, n_reads]          # (device pointers must stay 16-byte aligned: multiples of 8 reads)
        for a, z in zip(cuts[:-1], cuts[1:]):
            off = (offsets[a: z + 1] - offsets[a]).contiguous()
            assert (a * READ_LEN) % 16 == 0
            b.add_reads_dev(bases.data_ptr() + a * READ_LEN, off.data_ptr(), z - a, (z - a) * READ_LEN)
        b.sync()
        assert b.num_valid() == n_one
        assert np.array_equal(b.finalize(L, h), bits_one)
    # ~8x coverage: with c = 1 every distinct 31-mer of the covered genome counts once, with c = 3 those seen 3 times
    assert (3.0e7 < n_one < 4.1e7) if c == 1 else (1.5e7 < n_one < 4.0e7)
