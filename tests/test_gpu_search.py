"""GPU parity: bit-sliced search through the C ABI vs the oracle and the reference's golden hit lists."""
import numpy as np
import pytest

from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S
import util
from conftest import load_golden

pytestmark = pytest.mark.gpu


def hits_by_query(hits):
    out = {}
    for h in hits:
        out.setdefault(int(h["query"]), []).append((int(h["filter"]), int(h["num_match"])))
    return out


@pytest.mark.parametrize("name", list(S.SEARCH_CASES))
def test_search_matches_reference_golden(name):
    g = load_golden("search")[name]
    case = S.SEARCH_CASES[name]
    dbd = util.search_case_db(name)
    queries = S.search_queries(case)
    seqs = [s for _, s in queries]
    with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"]) as db:
        counts, nk = db.search_counts(seqs)
        for qi, s in enumerate(seqs):
            exp, n = O.search_counts(dbd["slices"], dbd["n"], dbd["L"], dbd["h"], dbd["k"], s)
            assert nk[qi] == n
            assert np.array_equal(counts[qi], exp), (name, qi)
        for t in case["thresholds"]:
            hits, nk2 = db.search(seqs, t)
            assert np.array_equal(nk2, nk)
            assert util.golden_rows_from_hits(queries, hits_by_query(hits), nk) == g["results"][repr(t)], (name, t)
            hits3, _ = db.search_ptrs(seqs, t)
            assert np.array_equal(hits3, hits)
            # ordered by (query, filter)
            key = hits["query"].astype(np.int64) * (1 << 32) + hits["filter"]
            assert np.all(np.diff(key) > 0)


@pytest.mark.parametrize("n_filters,L,h,k", [(1, 10, 1, 31), (8, 12, 3, 21), (100, 14, 2, 31), (129, 12, 5, 15), (1000, 12, 3, 31),
                                             (2048, 12, 3, 32), (4100, 10, 4, 31), (9000, 8, 8, 25)])
def test_search_counts_shapes(n_filters, L, h, k):
    rng = np.random.default_rng(n_filters + L)
    row = (n_filters + 7) // 8
    slices = rng.integers(0, 256, ((1 << L), row), dtype=np.uint8) | rng.integers(0, 256, ((1 << L), row), dtype=np.uint8)
    if n_filters % 8:
        slices[:, -1] &= (1 << (n_filters % 8)) - 1            # padding bits are zero in a .db (build_db.cpp:267)
    seqs = [bytes(O.gen_reads(5, i, 1, ln)).decode() for i, ln in enumerate([0, 10, k, k + 1, 100, 333, 1000, 2500])]
    seqs.append("ACGTN" * 50)
    seqs.append("A" * 500)
    with capi.Database.load(slices, k, h, L, n_filters) as db:
        counts, nk = db.search_counts(seqs)
        for qi, s in enumerate(seqs):
            exp, n = O.search_counts(slices, n_filters, L, h, k, s)
            assert nk[qi] == n
            assert np.array_equal(counts[qi], exp), qi
        for t in (1.0, 0.7, 0.3):
            hits, _ = db.search(seqs, t)
            hb = hits_by_query(hits)
            for qi, s in enumerate(seqs):
                hf, hm, n = O.search_matches(slices, n_filters, L, h, k, s, t)
                assert hb.get(qi, []) == [(int(f), int(m)) for f, m in zip(hf, hm)], (qi, t)


def test_search_long_query_many_kmers():
    # more unique k-mers than one counter segment (32768) exercises the segment loop
    n_filters, L, h, k = 200, 16, 2, 31
    rng = np.random.default_rng(9)
    row = (n_filters + 7) // 8
    slices = rng.integers(0, 256, ((1 << L), row), dtype=np.uint8) | rng.integers(0, 256, ((1 << L), row), dtype=np.uint8) \
        | rng.integers(0, 256, ((1 << L), row), dtype=np.uint8)
    seq = bytes(O.gen_reads(6, 0, 1, 80000)).decode()
    with capi.Database.load(slices, k, h, L, n_filters) as db:
        counts, nk = db.search_counts([seq, seq[:100]])
    exp, n = O.search_counts(slices, n_filters, L, h, k, seq)
    assert nk[0] == n and n > 70000 and np.array_equal(counts[0], exp)


@pytest.mark.parametrize("n_filters", [600, 2048, 4100])
def test_search_filter_holding_every_kmer_of_a_long_query(n_filters):
    # with more than 512 filters a substream got exactly 1024 k-mers per full segment and its 10-plane
    # counter wrapped to 0 for a filter that matches all of them.  Planted columns: one all-ones filter per 128-column
    # group boundary, the rest sparse; the oracle counts are the reference's (kwage.cpp:404-483).
    L, h, k = 14, 2, 31
    rng = np.random.default_rng(n_filters)
    row = (n_filters + 7) // 8
    slices = rng.integers(0, 256, ((1 << L), row), dtype=np.uint8) & rng.integers(0, 256, ((1 << L), row), dtype=np.uint8)
    full = [0, 127, 128, 511, 512, n_filters - 1]
    for f in full:
        slices[:, f >> 3] |= np.uint8(1 << (f & 7))
    seq = bytes(O.gen_reads(11, 0, 1, 45000)).decode()
    with capi.Database.load(slices, k, h, L, n_filters) as db:
        counts, nk = db.search_counts([seq])
        hits, _ = db.search([seq], 1.0)
    exp, n = O.search_counts(slices, n_filters, L, h, k, seq)
    assert nk[0] == n and n > 40000
    assert np.array_equal(counts[0], exp)
    assert all(int(counts[0][f]) == n for f in full)
    assert sorted(int(x) for x in hits["filter"]) == sorted(set(full))


def test_search_column_slabs_equal_whole():
    # multi-GPU layout: every device holds a column slab; the concatenation of slab counts is the answer
    dbd = util.search_case_db("random_n257")
    seqs = [s for _, s in S.search_queries(S.SEARCH_CASES["random_n257"])]
    with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"]) as db:
        whole, nk = db.search_counts(seqs)
    parts = []
    for a, z in [(0, 64), (64, 200), (200, 257)]:
        with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"], col_begin=a, col_end=z) as db:
            c, nk2 = db.search_counts(seqs)
            assert np.array_equal(nk2, nk) and c.shape[1] == z - a
            parts.append(c)
    assert np.array_equal(np.concatenate(parts, axis=1), whole)


def test_search_bad_arguments():
    dbd = util.search_case_db("random_n257")
    with capi.Database.load(dbd["slices"], dbd["k"], dbd["h"], dbd["L"], dbd["n"]) as db:
        for t in (0.0, -0.5, 1.5):
            with pytest.raises(capi.KwageError):
                db.search(["ACGT" * 20], t)
        hits, nk = db.search([], 0.5)
        assert len(hits) == 0
    with pytest.raises(capi.KwageError):
        capi.Database.load(dbd["slices"], 33, 3, dbd["L"], dbd["n"])
    with pytest.raises(capi.KwageError):
        capi.Database.load(dbd["slices"], 31, 3, dbd["L"], dbd["n"], col_begin=4, col_end=64)


# ------------------------------------------------------------------------------ several files as one column slab
@pytest.mark.parametrize("widths", [[13, 64, 21], [8, 2048, 5, 3], [2048, 2048], [1, 1, 30, 33, 7]])
def test_files_side_by_side_in_one_slab_search_like_one_database(widths):
    # kwg_db_upload_columns lays the slice regions of several files (any filter counts -> any bit offsets) side by side;
    # counts and hits must equal those of a database built from all the filters at once
    k, h, L = 25, 3, 12
    total = sum(widths)
    filters = [O.gen_filter_bits(4242, j, (1 << L) // 8) for j in range(total)]
    all_slices = O.transpose(filters, 1 << L)
    queries = [bytes(O.gen_reads(5 + i, 0, 1, 200 + 37 * i)).decode() for i in range(6)]
    with capi.Database.load(all_slices, k, h, L, total) as one:
        exp_counts, exp_nk = one.search_counts(queries)
        exp_hits, _ = one.search(queries, 0.3)
    with capi.Database.alloc(k, h, L, total) as slab:
        col = 0
        for w in widths:
            part = O.transpose(filters[col: col + w], 1 << L)          # what this file's slice region holds
            half = (1 << L) // 2 + 3
            slab.upload_columns(col, w, 0, part[:half])                 # in two row pieces, like the streaming loader
            slab.upload_columns(col, w, half, part[half:])
            col += w
        got_counts, got_nk = slab.search_counts(queries)
        got_hits, _ = slab.search(queries, 0.3)
    assert np.array_equal(got_nk, exp_nk) and np.array_equal(got_counts, exp_counts)
    assert np.array_equal(got_hits, exp_hits)
    with capi.Database.alloc(k, h, L, total) as slab:
        with pytest.raises(capi.KwageError):
            slab.upload_columns(total - 2, 5, 0, np.zeros((4, 1), np.uint8))    # column range outside the slab


@pytest.mark.parametrize("n_filters", [300, 600, 1100, 4100, 8200])
def test_search_stops_reading_a_chunk_that_cannot_reach_the_threshold(n_filters, monkeypatch):
    """kwg_search reads the rows of a (query, 4096-column chunk) only as long as some column of the chunk can still reach
    the threshold (search_count_kernel<NH, true>; the reference's early exit, kwage.cpp:397,459-482).  The hit lists must
    be the reference's whatever stops where: columns planted with 30 .. 100 % of a query's k-mers around every threshold,
    over a random background, chunks with and without a hit, slabs of 8 / 16 / 32 lanes per row and one (300 filters) that
    is too narrow for the early exit; KWG_SEARCH_NO_EXIT (every row is read) gives the same lists."""
    k, h, L = 31, 3, 13
    rng = np.random.default_rng(n_filters)
    row = (n_filters + 7) // 8
    slices = rng.integers(0, 256, ((1 << L), row), dtype=np.uint8) & rng.integers(0, 256, ((1 << L), row), dtype=np.uint8)
    seqs = [bytes(O.gen_reads(77, i, 1, ln)).decode() for i, ln in enumerate([1000, 1000, 400, 2200, 95, 1000, 31, 20])]
    fracs = [0.3, 0.45, 0.5, 0.55, 0.75, 0.97, 1.0]
    planted = {}
    for qi, s in enumerate(seqs):
        words = O.query_kmers(s, k)
        n = len(words)
        if n == 0:
            continue
        for j, f in enumerate(fracs):
            c = int(rng.integers(0, n_filters))
            m = min(n, int(np.ceil(f * n)))
            # the LAST m k-mers of the query (the reference's worst case: no match early, all matches late) or a random subset
            pick = range(n - m, n) if (qi + j) % 2 == 0 else rng.choice(n, m, replace=False)
            for i in pick:
                for sd in range(h):
                    r = O.murmur3_word(int(words[i]), k, sd) & ((1 << L) - 1)
                    slices[r, c >> 3] |= np.uint8(1 << (c & 7))
            planted[(qi, c)] = f
    if n_filters % 8:
        slices[:, -1] &= (1 << (n_filters % 8)) - 1
    results = {}
    for no_exit in (False, True):
        if no_exit:
            monkeypatch.setenv("KWG_SEARCH_NO_EXIT", "1")
        with capi.Database.load(slices, k, h, L, n_filters) as db:
            for t in (0.3, 0.5, 0.75, 1.0):
                hits, nk = db.search(seqs, t)
                got = hits_by_query(hits)
                if not no_exit:
                    for qi, s in enumerate(seqs):
                        hf, hm, _ = O.search_matches(slices, n_filters, L, h, k, s, t)
                        exp = sorted((int(f), int(c)) for f, c in zip(hf, hm))
                        assert sorted(got.get(qi, [])) == exp, (n_filters, t, qi)
                    assert any(got.values()), (n_filters, t)
                results.setdefault(t, []).append(got)
    for t, (a, b) in results.items():
        assert a == b, t
