"""Medium-size randomised parity (1e6 .. 1.5e7 bases per case): single-level geometries with thousands of runs per bucket
(several run-table passes and staging rounds), two-level geometries with real chunk counts, thresholds 1..5."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); n = 0
while time.time() - t0 < budget:
    c = int(rng.choice([1, 1, 2, 3, 5, 9]))
    lc = int(rng.choice([19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 30]))
    n_reads = int(rng.choice([8000, 30000, 60000, 100000, 250000]))
    read_len = int(rng.choice([100, 150]))
    cov = float(rng.choice([1.5, 6.0, 20.0]))
    seed = int(rng.integers(1, 1 << 30))
    case = dict(kind="coverage", seed=seed, genome=max(1000, int(n_reads * read_len / cov)), n_reads=n_reads, read_len=read_len, num_bp=-1)
    bases, offsets = S.make_bloom_reads(case)
    split = int(rng.choice([1, 2, 3]))
    ob = O.Builder(31, c, lc, 26); ob.add_reads(bases, offsets)
    with capi.BloomBuilder(31, min_kmer_count=c, log2_count_len=lc, log2_max_len=26) as b:
        cuts = [0] + sorted(int(x) for x in rng.integers(0, n_reads + 1, split - 1)) + [n_reads]
        for a, z in zip(cuts[:-1], cuts[1:]):
            b.add_reads(bases, offsets[a: z + 1])
        assert b.num_valid() == ob.num_valid(), ("num_valid", c, lc, n_reads, read_len, cov, seed, cuts, b.num_valid(), ob.num_valid())
        assert np.array_equal(b.finalize(26, 3), ob.finalize(26, 3)), ("bits", c, lc, n_reads, read_len, cov, seed, cuts)
    ob.close(); n += 1
    print("ok", c, lc, n_reads, read_len, cov, cuts, flush=True)
print("stress ok: %d medium cases in %.0f s" % (n, time.time() - t0))
