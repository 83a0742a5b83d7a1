"""CPU: the level-by-level formulation of the reference's conservative-update counting filters (the algorithm
bloom_count.cuh implements for min_kmer_count > 1) as a small Python model, checked against

  * a direct transcription of the reference's sequential loop (make_bloom.cpp:546-601) on abstract slot tuples, with tiny
    tables so that collisions, both-hashes-on-one-slot ("double") records and batch cuts are everywhere, and
  * the C oracle on real reads (slots from the real murmur3 values).

The model mirrors the kernels: one record per (occurrence, slot) with weight 2 when both hashes of a table meet, records in
ARBITRARY order, per level an atomicMin per slot over the eligible records, wins per occurrence, eligible at the next level
<=> no win, valid <=> a win at level c-1, persistent 4-bit counters between batches, a wrap of the 4-bit field reported."""
import random

import numpy as np
import pytest

from oracle import oracle_py as O
import synth_cases as S


def sequential(occ, nslots, c):
    first, second, valid = [0] * nslots, [0] * nslots, []
    for t, (a0, a1, b0, b1) in enumerate(occ):
        f0, f1, s0, s1 = first[a0], first[a1], second[b0], second[b1]
        m = min(f0, f1, s0, s1)
        if m < c:
            if m == c - 1:
                valid.append(t)
            if f0 == m:
                first[a0] = (first[a0] + 1) & 15
            if f1 == m:
                first[a1] = (first[a1] + 1) & 15
            if s0 == m:
                second[b0] = (second[b0] + 1) & 15
            if s1 == m:
                second[b1] = (second[b1] + 1) & 15
    return valid, first, second


def by_levels(occ, nslots, c, batches, rng):
    INF = 1 << 40
    cnt = [0] * (2 * nslots)                       # slot = index * 2 + table
    valid, wrapped = [], False
    n = len(occ)
    size = max(1, (n + batches - 1) // batches)
    for b0 in range(0, n, size):
        sub = occ[b0: b0 + size]
        recs = []
        for t, (a0, a1, c0, c1) in enumerate(sub):
            recs += [(a0 * 2, t, 2)] if a0 == a1 else [(a0 * 2, t, 1), (a1 * 2, t, 1)]
            recs += [(c0 * 2 + 1, t, 2)] if c0 == c1 else [(c0 * 2 + 1, t, 1), (c1 * 2 + 1, t, 1)]
        rng.shuffle(recs)                            # the kernels see the records in no particular order
        elig = [True] * len(sub)
        for v in range(c):
            tile, wins = {}, [0] * len(sub)
            for s, t, w in recs:
                if not elig[t] or cnt[s] > v:
                    continue
                val = ((t + 1) << 1) | (0 if w == 2 else 1)
                old = tile.get(s, INF)
                if val < old:
                    tile[s] = val
                    wins[t] += w
                    if old != INF:
                        wins[(old >> 1) - 1] -= 2 - (old & 1)      # displaced holder: its win is taken back
            for s, val in tile.items():
                nv = v + 2 - (val & 1)
                wrapped = wrapped or nv > 15
                cnt[s] = nv & 15
            if v == c - 1:
                valid += [b0 + t for t in range(len(sub)) if elig[t] and wins[t] > 0]
            else:
                elig = [elig[t] and wins[t] == 0 for t in range(len(sub))]
    return sorted(valid), cnt, wrapped


@pytest.mark.parametrize("seed", range(6))
def test_level_recursion_equals_the_sequential_counters(seed):
    rng = random.Random(seed)
    checked = 0
    for _ in range(400):
        nslots = rng.choice([2, 3, 5, 8, 16, 64])
        c = rng.choice([1, 2, 3, 5, 14, 15])
        n = rng.choice([5, 20, 100, 300])
        kmers = [tuple(rng.randrange(nslots) for _ in range(4)) for _ in range(rng.choice([3, 10, 50, 1000]))]
        occ = [rng.choice(kmers) for _ in range(n)]
        v1, first, second = sequential(occ, nslots, c)
        v2, cnt, wrapped = by_levels(occ, nslots, c, rng.choice([1, 2, 3]), rng)
        if wrapped:
            assert c == 15                            # the one case the device refuses instead of reproducing
            continue
        assert v1 == v2
        assert all(cnt[2 * i] == first[i] and cnt[2 * i + 1] == second[i] for i in range(nslots))
        checked += 1
    assert checked > 300


@pytest.mark.parametrize("c", [2, 3, 5])
def test_level_model_on_real_kmers_equals_the_oracle(c):
    # slots from the real murmur3 values of real canonical k-mers; the oracle is the C restatement of the reference.
    # (The oracle's tables are at least 2^18 slots; the model is run with the same 18-bit masks.)
    k, lc = 31, 18
    case = dict(kind="coverage", seed=40 + c, genome=3000, n_reads=300, read_len=80, num_bp=-1)
    bases, offsets = S.make_bloom_reads(case)
    ob = O.Builder(k, c, lc, 20)
    ob.add_reads(bases, offsets)
    expect = ob.num_valid()
    ob.close()
    mask = (1 << lc) - 1
    occ = []
    for r in range(len(offsets) - 1):
        words, _ = O.canonical_kmers(bases[int(offsets[r]): int(offsets[r + 1])], k)
        for w in words:
            h = [O.murmur3_word(int(w), k, s) & mask for s in range(4)]
            occ.append((h[0], h[1], h[2], h[3]))
    v, _, wrapped = by_levels(occ, 1 << lc, c, 2, random.Random(1))
    assert not wrapped and len(v) == expect
    assert len(sequential(occ, 1 << lc, c)[0]) == expect
