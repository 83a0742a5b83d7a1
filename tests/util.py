"""Helpers shared by the CPU and GPU tests (TEST INFRASTRUCTURE: uses the oracle)."""
import hashlib

import numpy as np

from oracle import oracle_py as O
import synth_cases as S


def sha256(b):
    return hashlib.sha256(np.ascontiguousarray(b).tobytes()).hexdigest()


def fixture_accession(j):
    return "SRR%07d" % (1000000 + j)


_db_cache = {}


def search_case_db(name):
    """Builds the search fixture's database with the ORACLE pipeline (make_bloom restatement +
    transpose restatement).  Returns dict(filters, slices, L, h, k, n)."""
    if name in _db_cache:
        return _db_cache[name]
    case = S.SEARCH_CASES[name]
    if case["kind"] == "random":
        L, h = case["L"], case["h"]
        filters = [O.gen_filter_bits(case["seed"], j, (1 << L) // 8) for j in range(case["n"])]
    else:
        filters = []
        L = h = None
        for j in range(case["n"]):
            bases, offsets = S.search_accession_reads(case, j)
            r = O.make_bloom(bases, offsets, case["k"], 1, 0.25, case["lmin"], case["lmax"], int(offsets[-1]))
            assert r["status"] == "success"
            if L is None:
                L, h = r["log2_len"], r["num_hash"]
            assert (L, h) == (r["log2_len"], r["num_hash"])
            filters.append(r["bits"])
    slices = O.transpose(filters, 1 << L)
    out = dict(filters=filters, slices=slices, L=L, h=h, k=case["k"], n=case["n"])
    _db_cache[name] = out
    return out


def golden_rows_from_hits(queries, hits_by_query, nk):
    """-> sorted rows [query name, n_kmers, num_match, accession] like make_golden.parse_csv"""
    rows = []
    for qi, (qn, _) in enumerate(queries):
        for f, m in hits_by_query.get(qi, []):
            rows.append([qn, int(nk[qi]), int(m), fixture_accession(int(f))])
    return sorted(rows)
