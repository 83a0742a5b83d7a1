"""min_kmer_count 1: the filter bits of seeds (2t, 2t+1) are the fold of counting table t's "touched" bitmap.

kwg_bloom_finalize takes seeds 0..(num_hash & ~1)-1 out of the touched bitmap of the counting tables
(bloom_build.cu::fold_touched_kernel) instead of hashing every valid k-mer again.  This test checks the identity
behind it on the CPU, against the oracle's sequential restatement of make_bloom.cpp:506-621 and its fold (337-354):
heavy shadowing (a 2^18-slot table under 4e5 touches), duplicated reads, every num_hash, L <= lc.
"""
import numpy as np
import pytest

from oracle import oracle_py as O
import synth_cases as S


def touched_and_valid(bases, offsets, k, lc):
    """-> (touched bitmaps of the two tables as bool arrays, list of (h0..h4) of the valid occurrences)"""
    m = (1 << lc) - 1
    t = [np.zeros(1 << lc, dtype=bool), np.zeros(1 << lc, dtype=bool)]
    valid = []
    for r in range(len(offsets) - 1):
        words, _ = O.canonical_kmers(bytes(bases[int(offsets[r]):int(offsets[r + 1])]), k)
        for w in words:
            h = [O.murmur3_word(int(w), k, s) for s in range(5)]
            slots = [(0, h[0] & m), (0, h[1] & m), (1, h[2] & m), (1, h[3] & m)]
            if not all(t[a][b] for a, b in slots):          # some counter is zero: valid, and all four are non-zero afterwards
                valid.append(h)
            for a, b in slots:
                t[a][b] = True
    return t, valid


def fold(table, L):
    return table.reshape(-1, 1 << L).any(axis=0)


@pytest.mark.parametrize("seed,n_reads,dups", [(5, 700, 0), (6, 1500, 300), (7, 2500, 2500)])
def test_filter_bits_of_a_seed_pair_are_the_fold_of_the_touched_bitmap(seed, n_reads, dups):
    k, lc, read_len = 31, 18, 70
    bases, offsets = S.uniform_reads(seed, 0, n_reads, read_len)
    if dups:
        bases = np.concatenate([bases, bases[: dups * read_len]])
        offsets = np.arange(n_reads + dups + 1, dtype=np.uint64) * np.uint64(read_len)
    t, valid = touched_and_valid(bases, offsets, k, lc)
    b = O.Builder(k, 1, lc, 24)
    b.add_reads(bases, offsets)
    assert b.num_valid() == len(valid)
    n_occ = sum(max(0, int(offsets[i + 1] - offsets[i]) - k + 1) for i in range(len(offsets) - 1))
    assert len(valid) < n_occ                                # the case has occurrences that are not valid
    for L in (10, 15, 18):
        for nh in (1, 2, 3, 4, 5):
            ref = np.unpackbits(b.finalize(L, nh), bitorder="little").astype(bool)
            n_fold = nh & ~1
            mine = np.zeros(1 << L, dtype=bool)
            for tbl in range(n_fold // 2):
                mine |= fold(t[tbl], L)
            for h in valid:
                for s in range(n_fold, nh):
                    mine[h[s] & ((1 << L) - 1)] = True
            assert np.array_equal(mine, ref), (L, nh)
    b.close()
