"""Randomised soak of the counting construction against the oracle (not part of the test suite): random thresholds,
counting-filter lengths, batch splits and input mixes, for a fixed wall-clock budget.  compute-sanitizer is closed on
this pool, so repeated randomised parity is what stands in for racecheck."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
n = 0
wraps = 0
while time.time() - t0 < budget:
    c = int(rng.choice([1, 1, 2, 3, 5, 5, 8, 15]))
    lc = int(rng.choice([18, 19, 21, 23, 24, 26, 28, 30]))
    k = int(rng.choice([31, 31, 32, 21, 15]))
    genome = int(rng.choice([3000, 20000, 200000]))
    n_reads = int(rng.choice([500, 4000, 20000]))
    read_len = int(rng.choice([60, 100, 150]))
    seed = int(rng.integers(1, 1 << 30))
    case = dict(kind="coverage", seed=seed, genome=genome, n_reads=n_reads, read_len=read_len, num_bp=-1)
    bases, offsets = S.make_bloom_reads(case)
    mut = dup = None
    if rng.random() < 0.5:
        mut = (int(rng.choice([0, 97, 500])), int(rng.choice([0, 5])))
        bases = S.mutate(bases, seed, n_rate=mut[0], lower_rate=mut[1])
    if rng.random() < 0.3:      # heavy duplication: a block of identical reads
        dup = int(rng.integers(50, 2000))
        rep = np.tile(bases[: read_len], dup)
        bases = np.concatenate([bases, rep])
        offsets = np.arange(len(bases) // read_len + 1, dtype=np.uint64) * np.uint64(read_len)
    split = int(rng.choice([1, 1, 2, 5]))
    ob = O.Builder(k, c, lc, 24)
    ob.add_reads(bases, offsets)
    try:
        with capi.BloomBuilder(k, min_kmer_count=c, log2_count_len=lc, log2_max_len=24) as b:
            nr = len(offsets) - 1
            cuts = [0] + sorted(int(x) for x in rng.integers(0, nr + 1, split - 1)) + [nr]
            for a, z in zip(cuts[:-1], cuts[1:]):
                b.add_reads(bases, offsets[a: z + 1])
            nv = b.num_valid()
            assert nv == ob.num_valid(), ("num_valid", c, lc, k, genome, n_reads, read_len, seed, split, cuts, mut, dup, nv, ob.num_valid())
            L = int(rng.choice([20, 22, 24]))
            h = int(rng.integers(1, 6))
            bits, crc = b.finalize_crc(L, h)
            assert np.array_equal(bits, ob.finalize(L, h)), ("bits", c, lc, k, genome, n_reads, read_len, seed, split, cuts, mut, dup)
            assert crc == O.crc32(bits)
    except capi.KwageError as e:
        # the one refused case: min_kmer_count 15 and a double increment from 14 (reported, never mimicked)
        assert c == 15 and "wrapped" in str(e), (c, lc, str(e))
        wraps += 1
    ob.close()
    n += 1
print("stress ok: %d random cases in %.0f s (%d refused for a counter wrap)" % (n, time.time() - t0, wraps))
