"""Randomised soak of transposition (+ device crc32), column slabs and search against the oracle / zlib."""
import sys, time, os, zlib
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from kwage_b200 import capi
from oracle import oracle_py as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
n = 0
while time.time() - t0 < budget:
    # ---- transpose (+ running checksums) of n filters x bits
    nf = int(rng.choice([1, 3, 8, 31, 32, 64, 257, 300, 1024, 2048]))
    bits = int(rng.choice([8, 64, 1000 * 8, 4096, 1 << 15, 1 << 18]))
    filters = [rng.integers(0, 256, bits // 8, dtype=np.uint8) for _ in range(nf)]
    exp = O.transpose(filters, bits)
    got = capi.transpose(filters, bits)
    assert np.array_equal(got, exp), ("transpose", nf, bits)
    if bits % 32 == 0:
        want_dest = nf % 32 == 0
        s, fc, dc = capi.transpose_crc(filters, bits, rng.integers(0, 1 << 32, nf, dtype=np.uint64).astype(np.uint32) * 0, 5 if want_dest else None)
        assert np.array_equal(s, exp) and list(fc) == [zlib.crc32(f.tobytes()) for f in filters], ("transpose_crc", nf, bits)
        if want_dest:
            assert dc == zlib.crc32(exp.tobytes(), 5)
    # ---- search: random slab geometry, several files side by side, random queries incl. Ns and short ones
    k = int(rng.choice([15, 21, 31, 32]))
    h = int(rng.integers(1, 6))
    L = int(rng.choice([10, 12, 14]))
    widths = [int(x) for x in rng.choice([1, 5, 8, 13, 64, 100, 2048], size=int(rng.integers(1, 4)))]
    total = sum(widths)
    fl = [O.gen_filter_bits(int(rng.integers(1, 1 << 20)), j, (1 << L) // 8) for j in range(total)]
    slices = O.transpose(fl, 1 << L)
    queries = []
    for i in range(int(rng.integers(1, 6))):
        q = bytearray(O.gen_reads(int(rng.integers(1, 1 << 20)), 0, 1, int(rng.choice([5, k, k + 1, 100, 700, 3000]))))
        if rng.random() < 0.3 and len(q) > 10:
            q[len(q) // 2] = ord("N")
        queries.append(bytes(q).decode())
    t = float(rng.choice([1.0, 0.5, 0.1, 0.01]))
    with capi.Database.alloc(k, h, L, total) as slab:
        col = 0
        for w in widths:
            slab.upload_columns(col, w, 0, O.transpose(fl[col: col + w], 1 << L))
            col += w
        counts, nk = slab.search_counts(queries)
        hits, _ = slab.search(queries, t)
    hit_set = {(int(x["query"]), int(x["filter"]), int(x["num_match"])) for x in hits}
    exp_hits = set()
    for qi, q in enumerate(queries):
        e, nq = O.search_counts(slices, total, L, h, k, q)
        assert nk[qi] == nq and np.array_equal(counts[qi], e), ("search_counts", k, h, L, widths, len(q))
        hf, hm, _ = O.search_matches(slices, total, L, h, k, q, t)
        for f, m in zip(hf, hm):
            exp_hits.add((qi, int(f), int(m)))
    assert hit_set == exp_hits, ("search hits", k, h, L, widths, t, len(hit_set), len(exp_hits))
    n += 1
print("stress ok: %d random rounds (transpose + crc + slab search) in %.0f s" % (n, time.time() - t0))
