"""CPU: the oracle (C restatement) against the golden vectors generated from the compiled
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import oracle_py as O
import synth_cases as S
import util
from conftest import load_golden


def test_hash_kats_match_reference():
    kats = load_golden("hash_kats")
    assert len(kats) == len(S.HASH_KAT_INPUTS)
    n_rows = 0
    for kat in kats:
        words, loc5 = O.canonical_kmers(kat["seq"], kat["k"])
        assert len(words) == len(kat["rows"])
        for w, l5, row in zip(words, loc5, kat["rows"]):
            assert int(l5) == row["loc5"]
            assert int(w) == int(row["canon"], 16)
            assert int(w) == min(int(row["sense"], 16), int(row["anti"], 16))
            for s, hx in enumerate(row["hashes"]):
                assert O.murmur3_word(w, kat["k"], s) == int(hx, 16)
                # identical to textbook murmur3_x86_32 of the canonical ASCII k-mer
                assert O.murmur3_bytes(row["text"], s) == int(hx, 16)
            n_rows += 1
    assert n_rows > 300


def test_survey_pinned_kats():
    # SURVEY.md section 8c table (captured from the compiled reference)
    w, _ = O.canonical_kmers("ACGTACGTACGTACGTACGTACGTACGTACG", 31)
    assert int(w[0]) == 0x06c6c6c6c6c6c6c6
    assert [O.murmur3_word(w[0], 31, s) for s in range(5)] == [0x7f7b12a3, 0x16d21cd3, 0x1d86c8ab, 0x14dc07ff, 0xdd02e49a]
    w, _ = O.canonical_kmers("T" * 31, 31)
    assert int(w[0]) == 0 and O.murmur3_word(0, 31, 0) == 0x30e9726e
    w, l5 = O.canonical_kmers("ACGTN" + "ACGT" * 8 + "AC", 31)
    assert int(l5[0]) == 5 and int(w[0]) == 0x06c6c6c6c6c6c6c6


def test_param_kats_match_reference():
    g = load_golden("param_kats")
    for e in g["optimal_bloom_param"]:
        r = O.optimal_bloom_param(e["n"], e["p"], e["lmin"], e["lmax"])
        assert (None if r is None else list(r)) == e["result"], e
    for e in g["approximate_max_kmers"]:
        assert O.approximate_max_kmers(e["p"], e["lmin"], e["lmax"]) == e["result"], e


@pytest.mark.parametrize("name", list(S.MAKE_BLOOM_CASES))
def test_make_bloom_matches_reference(name):
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    r = O.make_bloom(bases, offsets, case["k"], case["min_count"], case["p"], case["lmin"], case["lmax"], case["num_bp"])
    assert r["num_kmer"] == g["num_kmer"]
    assert r["log2_count_len"] == g["log2_count_len"]
    if g["status"] == 14:
        assert r["status"] == "success"
        assert (r["log2_len"], r["num_hash"]) == (g["log2_len"], g["num_hash"])
        assert O.crc32(r["bits"]) == g["bits_crc32"]
        assert util.sha256(r["bits"]) == g["bits_sha256"]
    else:
        assert g["status"] == 16 and r["status"] == "invalid"


@pytest.mark.parametrize("name", list(S.BUILD_DB_CASES))
def test_transpose_matches_reference(name):
    g = load_golden("build_db")[name]
    case = S.BUILD_DB_CASES[name]
    filters = S.build_db_filters(case)
    slices = O.transpose(filters, 1 << case["L"])
    assert util.sha256(slices) == g["slices_sha256"]
    assert O.crc32(slices.reshape(-1)) == g["slices_crc32"]


def test_transpose_is_chunkable():
    # build_db.cpp:259 works chunk by chunk: transposing pieces and stacking them is the same thing
    case = S.BUILD_DB_CASES["n300_L16"]
    filters = S.build_db_filters(case)
    whole = O.transpose(filters, 1 << case["L"])
    half = (1 << case["L"]) // 2
    a = O.transpose([f[: half // 8] for f in filters], half)
    b = O.transpose([f[half // 8:] for f in filters], half)
    assert np.array_equal(np.concatenate([a, b]), whole)


@pytest.mark.parametrize("name", list(S.SEARCH_CASES))
def test_search_matches_reference(name):
    g = load_golden("search")[name]
    case = S.SEARCH_CASES[name]
    db = util.search_case_db(name)
    assert (db["L"], db["h"]) == (g["L"], g["h"])
    queries = S.search_queries(case)
    for t in case["thresholds"]:
        hits, nk = {}, []
        for qi, (_, seq) in enumerate(queries):
            hf, hm, n = O.search_matches(db["slices"], db["n"], db["L"], db["h"], db["k"], seq, t)
            nk.append(n)
            hits[qi] = list(zip(hf, hm))
            # the early exits never change the match set: full counts give the same answer
            counts, n2 = O.search_counts(db["slices"], db["n"], db["L"], db["h"], db["k"], seq)
            assert n2 == n
            if n:
                thr = int(np.float32(t) * np.float32(n))
                exp = [(f, (n if t == 1.0 else int(c))) for f, c in enumerate(counts) if (c == n if t == 1.0 else c >= thr)]
                assert exp == [(int(f), int(m)) for f, m in zip(hf, hm)]
        assert util.golden_rows_from_hits(queries, hits, nk) == g["results"][repr(t)], (name, t)
