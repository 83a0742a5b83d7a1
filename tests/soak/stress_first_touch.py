"""Randomised parity of the first-touch construction (bloom_first.cuh: min_kmer_count 1, counting filters up to 2^30 slots)
against the oracle: random counting-filter lengths (1 .. 2048 buckets), coverage, low-complexity blocks (poly-A reads,
repeated reads: thousands of records of one round on a few slots), N's and lower case, ragged reads, random cuts into
calls, ASCII or 2-bit packed input (kwg_bloom_add_packed, incl. calls that start mid-byte of the 2na stream)."""
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
    lc = int(rng.choice([18, 19, 20, 21, 22, 24, 26, 28, 29, 30]))
    k = int(rng.choice([31, 31, 31, 21, 32, 15]))
    n_reads = int(rng.choice([300, 3000, 20000, 60000, 150000]))
    max_len = int(rng.choice([60, 100, 150, 250]))
    cov = float(rng.choice([0.0, 0.0, 3.0, 25.0]))
    seed = int(rng.integers(1, 1 << 30))
    if cov > 0:
        case = dict(kind="coverage", seed=seed, genome=max(1000, int(n_reads * max_len / cov)), n_reads=n_reads, read_len=max_len, num_bp=-1)
        bases, offsets = S.make_bloom_reads(case)
    else:
        bases, offsets = S.uniform_reads(seed, 0, n_reads, max_len)
    bases = bases.copy()
    kind = int(rng.integers(0, 4))
    if kind == 1:                                    # a block of poly-A reads and a block of copies of the first reads
        a = int(rng.integers(0, max(1, n_reads // 2))) * max_len
        z = min(len(bases), a + int(rng.integers(1, 4000)) * max_len)
        bases[a:z] = ord("A")
        m = min(len(bases) // 3, int(rng.integers(1, 3000)) * max_len)
        bases[len(bases) - m:] = bases[:m]
    elif kind == 2:                                  # N's, IUPAC codes, lower case
        bases = S.mutate(bases, seed, n_rate=int(rng.choice([17, 301, 5003])), lower_rate=int(rng.choice([3, 11])))
    elif kind == 3:                                  # ragged reads (empty ones too)
        bases, offsets = S.ragged(bases, seed, n_reads, 0, max_len)
    n_r = len(offsets) - 1
    split = int(rng.choice([1, 1, 2, 4]))
    packed_in = bool(rng.integers(0, 2))
    ob = O.Builder(k, 1, lc, 26); ob.add_reads(bases, offsets)
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=lc, log2_max_len=26) as b:
        cuts = [0] + sorted(int(x) for x in rng.integers(0, n_r + 1, split - 1)) + [n_r]
        if packed_in:
            p2, mask = capi.pack_2na(bases)
        for a, z in zip(cuts[:-1], cuts[1:]):
            if packed_in:
                b.add_packed(p2, mask, offsets[a: z + 1])
            else:
                b.add_reads(bases, offsets[a: z + 1])
        tag = (lc, k, n_reads, max_len, cov, kind, seed, cuts, packed_in)
        assert b.num_valid() == ob.num_valid(), ("num_valid",) + tag + (b.num_valid(), ob.num_valid())
        L, h = int(rng.choice([18, 22, 26])), int(rng.choice([1, 3, 5]))
        assert np.array_equal(b.finalize(L, h), ob.finalize(L, h)), ("bits",) + tag
    ob.close(); n += 1
    print("ok", *tag, flush=True)
print("stress ok: %d first-touch cases in %.0f s" % (n, time.time() - t0))
