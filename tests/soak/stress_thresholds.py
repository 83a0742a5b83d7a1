"""Targeted parity around the internal thresholds of the counting construction: run-table passes (1024 runs per bucket),
staging windows (9150 records), the long-run limit (64 records), one partition tile (2048 positions) more or less."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); n = 0
read_len = 128                       # 16 reads = one partition tile of 2048 positions
while time.time() - t0 < budget:
    c = int(rng.choice([1, 1, 2, 4]))
    lc = int(rng.choice([18, 19, 20, 21, 22, 23]))
    buckets = 1 << (lc + 1 - 15)
    kind = int(rng.integers(0, 3))
    if kind == 0:      # run-table passes: 1024 / 2048 tiles, +- a few
        tiles = int(rng.choice([1024, 2048])) + int(rng.integers(-3, 4))
    elif kind == 1:    # staged records of a bucket around one / two staging windows: 8192 * tiles / buckets ~ 9150 * m
        m = int(rng.choice([1, 2, 3]))
        tiles = max(2, int(9150 * m * buckets / 8192 * float(rng.choice([0.5, 1.0])) ) + int(rng.integers(-4, 5)))   # (half the runs are long when they average 64)
    else:              # runs around the long-run limit: 8192 / buckets ~ 64 <=> 128 buckets (lc = 21); any size
        lc, buckets = 21, 128
        tiles = int(rng.integers(20, 700))
    tiles = min(tiles, 2100)
    n_reads = tiles * 16 - int(rng.integers(0, 16))
    cov = float(rng.choice([1.2, 4.0, 12.0]))
    seed = int(rng.integers(1, 1 << 30))
    case = dict(kind="coverage", seed=seed, genome=max(1000, int(n_reads * read_len / cov)), n_reads=n_reads, read_len=read_len, num_bp=-1)
    bases, offsets = S.make_bloom_reads(case)
    split = int(rng.choice([1, 1, 2]))
    ob = O.Builder(31, c, lc, 24); ob.add_reads(bases, offsets)
    with capi.BloomBuilder(31, min_kmer_count=c, log2_count_len=lc, log2_max_len=24) as b:
        cuts = [0] + sorted(int(x) for x in rng.integers(0, n_reads + 1, split - 1)) + [n_reads]
        for a, z in zip(cuts[:-1], cuts[1:]):
            b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == ob.num_valid(), ("num_valid", c, lc, n_reads, cov, seed, cuts, b.num_valid(), ob.num_valid())
        assert np.array_equal(b.finalize(24, 2), ob.finalize(24, 2)), ("bits", c, lc, n_reads, cov, seed, cuts)
    ob.close(); n += 1
print("stress ok: %d threshold cases in %.0f s" % (n, time.time() - t0))
