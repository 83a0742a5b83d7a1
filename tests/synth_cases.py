"""Deterministic synthetic inputs shared by tests/golden/make_golden.py (which runs the reference on
them) and the tests (which run the oracle / the CUDA path on them).  Everything derives from the
counter-based generator kwo_rnd (oracle/kwage_oracle.c), re-stated here in numpy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle_py as O  # noqa: E402

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rnd(seed, stream, ctr):
    """numpy restatement of kwo_rnd(seed, stream, ctr); stream/ctr may be arrays."""
    with np.errstate(over="ignore"):
        s = _mix64(np.uint64(seed) ^ _mix64(np.asarray(stream, dtype=np.uint64)))
        return _mix64(s + np.asarray(ctr, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))


def uniform_reads(seed, first_read, n_reads, read_len):
    bases = O.gen_reads(seed, first_read, n_reads, read_len)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return bases, offsets


def mutate(bases, seed, n_rate=0, lower_rate=0):
    """Every ~n_rate-th base becomes 'N' (or another non-ACGT byte), every ~lower_rate-th lower case."""
    b = bases.copy()
    idx = np.arange(len(b), dtype=np.uint64)
    r = rnd(seed, 0xBAD, idx)
    if n_rate:
        hit = (r % np.uint64(n_rate)) == 0
        junk = np.frombuffer(b"NnRY-.*", dtype=np.uint8)
        b[hit] = junk[((r[hit] >> np.uint64(20)) % np.uint64(len(junk))).astype(np.int64)]
    if lower_rate:
        hit = ((r >> np.uint64(32)) % np.uint64(lower_rate)) == 0
        b[hit] = b[hit] | 0x20
    return b


def ragged(bases, seed, n_reads, min_len, max_len):
    """Cut the flat base array into reads of pseudo-random lengths in [min_len, max_len] (0 allowed)."""
    lens = (min_len + (rnd(seed, 0x1E9, np.arange(n_reads, dtype=np.uint64)) % np.uint64(max_len - min_len + 1))).astype(np.uint64)
    offsets = np.concatenate([[np.uint64(0)], np.cumsum(lens, dtype=np.uint64)])
    assert offsets[-1] <= len(bases)
    return bases[: int(offsets[-1])].copy(), offsets


def write_reads_file(path, bases, offsets):
    with open(path, "wb") as f:
        for r in range(len(offsets) - 1):
            f.write(bytes(bases[int(offsets[r]): int(offsets[r + 1])]))
            f.write(b"\n")


# ---------------------------------------------------------------------------------------- KATs
HASH_KAT_INPUTS = [
    ("ACGTACGTACGTACGTACGTACGTACGTACG", 21),
    ("ACGTACGTACGTACGTACGTACGTACGTACG", 31),
    ("T" * 31, 31),
    ("GATTACA" * 5, 31),
    ("GATTACA" * 5, 32),
    ("ACGTN" + "ACGT" * 8 + "AC", 31),
    ("acgtACGTnACGTTGCAtgcaGGCCAATTggccaattACGTACGAT", 11),
    ("GATTACAGATTACACATTAGGATTACA", 1),
    ("GATTACAGATTACACATTAGGATTACA", 2),
    ("GATTACAGATTACACATTAGGATTACA", 3),
    ("GATTACAGATTACACATTAGGATTACA", 4),
    ("GATTACAGATTACACATTAGGATTACA", 5),
    ("GATTACAGATTACACATTAGGATTACA", 8),
    ("TTGACCAGTTAGCCATAGGACCATTAGGACCAGATTTAGACCAGGGATTTACCCAGATAGAGACCCATTTG", 16),
    ("TTGACCAGTTAGCCATAGGACCATTAGGACCAGATTTAGACCAGGGATTTACCCAGATAGAGACCCATTTG", 25),
    ("TTGACCAGTTAGCCATAGGACCATTAGGACCAGATTTAGACCAGGGATTTACCCAGATAGAGACCCATTTG", 30),
    ("TTGACCAGTTAGCCATAGGACCATTAGGACCAGATTTAGACCAGGGATTTACCCAGATAGAGACCCATTTG", 32),
    ("A" * 40, 32),
    ("C" * 35 + "G" * 35, 32),
]

PARAM_KAT_INPUTS = [
    (31, 1000000, 0.25, 18, 32), (31, 12000000, 0.25, 18, 32), (31, 120000000, 0.25, 18, 32),
    (31, 1000000000, 0.25, 18, 32), (31, 3000000000, 0.25, 18, 32), (31, 11998171, 0.25, 18, 32),
    (31, 1, 0.25, 18, 32), (31, 100, 0.01, 5, 10), (31, 449534, 0.25, 18, 24), (31, 5000, 0.001, 10, 20),
    (21, 262144, 0.25, 18, 18), (21, 200000, 0.5, 18, 18), (31, 90000, 0.25, 18, 18), (31, 100000, 0.25, 18, 18),
    (31, 2147483648, 0.25, 18, 32), (31, 1490000000, 0.25, 18, 32), (31, 1500000000, 0.25, 18, 32),
    (31, 119958110, 0.25, 18, 32), (31, 250000, 0.05, 18, 32), (31, 7777777, 0.1, 20, 30),
]

MAXKMER_KAT_INPUTS = [(0.25, 18, 32), (0.25, 18, 24), (0.01, 18, 32), (0.5, 10, 20), (0.05, 18, 30)]

# ---------------------------------------------------------------------------------------- construction
# num_bp feeds the counting-filter size (reference make_bloom.cpp:104-129); -1 = sum of read lengths
MAKE_BLOOM_CASES = {
    "uniform_k31": dict(kind="uniform", seed=11, n_reads=3000, read_len=150, k=31, min_count=1, p=0.25, lmin=18, lmax=24, num_bp=-1),
    "ragged_k21": dict(kind="ragged", seed=12, n_reads=20000, min_len=0, max_len=120, n_rate=53, lower_rate=7, k=21,
                       min_count=1, p=0.25, lmin=18, lmax=26, num_bp=-1),
    "k32": dict(kind="uniform", seed=13, n_reads=5000, read_len=150, k=32, min_count=1, p=0.25, lmin=18, lmax=25, num_bp=-1),
    "k15_dups": dict(kind="coverage", seed=14, genome=20000, n_reads=4000, read_len=100, k=15, min_count=1, p=0.25,
                     lmin=18, lmax=24, num_bp=-1),
    "small_count_filter": dict(kind="uniform", seed=15, n_reads=40000, read_len=150, k=31, min_count=1, p=0.25, lmin=18, lmax=26,
                               num_bp=3000),   # metadata under-reports: Lc clamps to 18 -> many slot collisions
    "min_count_2": dict(kind="coverage", seed=16, genome=30000, n_reads=6000, read_len=100, k=31, min_count=2, p=0.25,
                        lmin=18, lmax=24, num_bp=-1),
    "min_count_5": dict(kind="coverage", seed=17, genome=20000, n_reads=10000, read_len=100, k=31, min_count=5, p=0.25,
                        lmin=18, lmax=24, num_bp=-1),
    "invalid_too_many": dict(kind="uniform", seed=18, n_reads=2000, read_len=150, k=31, min_count=1, p=0.25, lmin=18, lmax=18, num_bp=-1),
    "no_kmers": dict(kind="uniform", seed=19, n_reads=50, read_len=20, k=31, min_count=1, p=0.25, lmin=18, lmax=24, num_bp=-1),
    "cfg1_mt64": dict(kind="mt64", seed=12345, n_reads=100000, read_len=150, k=31, min_count=1, p=0.25, lmin=18, lmax=32, num_bp=-1),
}


def make_bloom_reads(case):
    kind = case["kind"]
    if kind == "uniform":
        bases, offsets = uniform_reads(case["seed"], 0, case["n_reads"], case["read_len"])
    elif kind == "mt64":
        bases = O.gen_reads_mt64(case["seed"], case["n_reads"], case["read_len"])
        offsets = np.arange(case["n_reads"] + 1, dtype=np.uint64) * np.uint64(case["read_len"])
    elif kind == "ragged":
        flat = O.gen_reads(case["seed"], 0, case["n_reads"], case["max_len"])
        flat = mutate(flat, case["seed"], case.get("n_rate", 0), case.get("lower_rate", 0))
        bases, offsets = ragged(flat, case["seed"], case["n_reads"], case["min_len"], case["max_len"])
    elif kind == "coverage":
        genome = O.gen_reads(case["seed"], 0, 1, case["genome"])
        starts = (rnd(case["seed"], 0xC0FE, np.arange(case["n_reads"], dtype=np.uint64)) %
                  np.uint64(case["genome"] - case["read_len"])).astype(np.int64)
        idx = starts[:, None] + np.arange(case["read_len"])[None, :]
        bases = genome[idx].reshape(-1).copy()
        offsets = np.arange(case["n_reads"] + 1, dtype=np.uint64) * np.uint64(case["read_len"])
    else:
        raise ValueError(kind)
    if case["num_bp"] == -1:
        case["num_bp"] = int(offsets[-1])
    return bases, offsets


# ---------------------------------------------------------------------------------------- transposition
BUILD_DB_CASES = {
    "n257_L18": dict(n=257, L=18, k=31, h=3, seed=5),      # the reference's own db_debug.cpp shape (257 = non multiple of 8)
    "n8_L16": dict(n=8, L=16, k=31, h=3, seed=6),
    "n300_L16": dict(n=300, L=16, k=25, h=4, seed=7),
    "n1_L20": dict(n=1, L=20, k=31, h=1, seed=8),
    "n2048_L14": dict(n=2048, L=14, k=31, h=3, seed=9),    # MAX_NUM_FILTER_CHUNK columns (options.h:137)
}


def build_db_filters(case):
    nbytes = (1 << case["L"]) // 8
    return [O.gen_filter_bits(case["seed"], j, nbytes) for j in range(case["n"])]


# ---------------------------------------------------------------------------------------- search
SEARCH_CASES = {
    "accessions": dict(kind="accessions", n=12, seed=21, n_reads=2000, read_len=150, k=31, lmin=18, lmax=24,
                       thresholds=[1.0, 0.5, 0.2, 0.01]),
    "random_n257": dict(kind="random", n=257, L=18, k=31, h=3, seed=5, thresholds=[1.0, 0.05, 0.01, 0.001]),
}


def search_accession_reads(case, j):
    return uniform_reads(case["seed"] + 1000 * (j + 1), 0, case["n_reads"], case["read_len"])


def search_queries(case):
    """-> list of (name, sequence str)"""
    qs = []
    rnd_seq = lambda s, n: bytes(O.gen_reads(case["seed"] + 77, s, 1, n)).decode()  # noqa: E731
    if case["kind"] == "accessions":
        b3, o3 = search_accession_reads(case, 3)
        b5, o5 = search_accession_reads(case, 5)
        r3 = bytes(b3[: 150 * 4]).decode()                      # four reads of accession 3, concatenated
        r5 = bytes(b5[150 * 10: 150 * 12]).decode()
        qs.append(("reads_of_3", r3))
        qs.append(("mix_5_random", r5[:200] + rnd_seq(1, 400)))
        qs.append(("random_1kb", rnd_seq(2, 1000)))
        qs.append(("short", "ACGTACGTACGT"))
        qs.append(("with_N", r3[:100] + "N" + r3[101:260] + "nn" + r5[:90]))
        qs.append(("lower", r5[:300].lower()))
        qs.append(("dup_kmers", r3[:80] * 3))
        qs.append(("exact_k", r3[:31]))
    else:
        for i in range(6):
            qs.append(("rand%d" % i, rnd_seq(10 + i, 200 + 150 * i)))
        qs.append(("short", "ACGT"))
        qs.append(("polyA", "A" * 100))
    return qs
