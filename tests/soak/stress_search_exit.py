"""Randomised parity of the thresholded search (kwg_search: search_count_kernel<NH, true>, which stops reading the rows of a
column chunk once no column of it can reach the threshold) against the oracle's restatement of kwage.cpp:340-541: random slab
widths (8 / 16 / 32 lanes per row, chunks with a single active lane, slabs too narrow for the early exit), hash counts, k,
background densities, thresholds incl. 1.0, queries from a few k-mers to more than one segment, and columns planted with a
fraction of a query's k-mers just below / at / above the threshold, early or late in the k-mer order."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from kwage_b200 import capi
from oracle import oracle_py as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
n = 0
n_hits = 0
while time.time() - t0 < budget:
    k = int(rng.choice([15, 21, 31, 31, 32]))
    h = int(rng.integers(1, 6))
    L = int(rng.choice([10, 12, 13]))
    nf = int(rng.choice([520, 600, 1000, 1100, 2048, 4096, 4100, 4200, 8192, 9000]))
    row = (nf + 7) // 8
    dens = int(rng.integers(0, 3))                       # background: ~50 %, ~25 %, ~6 % of the bits
    slices = rng.integers(0, 256, ((1 << L), row), dtype=np.uint8)
    for _ in range(dens):
        slices &= rng.integers(0, 256, ((1 << L), row), dtype=np.uint8)
    t = float(rng.choice([1.0, 0.9, 0.75, 0.5, 0.5, 0.33, 0.2, 0.05]))
    queries = []
    for i in range(int(rng.integers(2, 7))):
        ln = int(rng.choice([k + 3, 100, 400, 1000, 1000, 2500, 9500 if nf <= 4200 else 1000]))
        queries.append(bytes(O.gen_reads(int(rng.integers(1, 1 << 20)), 0, 1, ln)).decode())
    mask = (1 << L) - 1
    for s in queries:
        words = O.query_kmers(s, k)
        nk = len(words)
        if nk == 0:
            continue
        for j in range(int(rng.integers(0, 5))):
            c = int(rng.integers(0, nf))
            f = min(1.0, max(0.0, t + float(rng.choice([-0.06, -0.01, 0.0, 0.002, 0.05, 1.0]))))
            m = min(nk, int(np.ceil(f * nk)))
            mode = int(rng.integers(0, 3))
            pick = range(nk - m, nk) if mode == 0 else range(m) if mode == 1 else rng.choice(nk, m, replace=False)
            for i in pick:
                for sd in range(h):
                    r = O.murmur3_word(int(words[i]), k, sd) & mask
                    slices[r, c >> 3] |= np.uint8(1 << (c & 7))
    if nf % 8:
        slices[:, -1] &= (1 << (nf % 8)) - 1
    with capi.Database.load(slices, k, h, L, nf) as db:
        hits, nkq = db.search(queries, t)
    got = {(int(x["query"]), int(x["filter"]), int(x["num_match"])) for x in hits}
    exp = set()
    for qi, s in enumerate(queries):
        hf, hm, nq = O.search_matches(slices, nf, L, h, k, s, t)
        assert nkq[qi] == nq
        for f, m in zip(hf, hm):
            exp.add((qi, int(f), int(m)))
    assert got == exp, ("search hits", k, h, L, nf, dens, t, [len(q) for q in queries], len(got), len(exp), sorted(got ^ exp)[:5])
    n += 1
    n_hits += len(exp)
print("stress ok: %d random thresholded searches (%d hits) in %.0f s" % (n, n_hits, time.time() - t0))
