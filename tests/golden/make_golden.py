#!/usr/bin/env python
"""Generates tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile) on deterministic synthetic inputs.

Run here (where /root/reference exists):   python tests/golden/make_golden.py
The inputs are re-created inside the tests from the same seeds (tests/synth_cases.py); only the
reference's outputs (digests, parameters, hit lists) are stored.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_py as O  # noqa: E402
import synth_cases as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha256(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def gen_hash_kats():
    out = []
    for seq, k in S.HASH_KAT_INPUTS:
        rows = []
        for line in O.ref_driver("hash", k, 8, seq).splitlines():
            f = line.split()
            rows.append(dict(loc5=int(f[0]), sense=f[1], anti=f[2], canon=f[3], hashes=f[4:12], text=f[12]))
        out.append(dict(seq=seq, k=k, rows=rows))
    return out


def gen_param_kats():
    opt = []
    for (k, n, p, lmin, lmax) in S.PARAM_KAT_INPUTS:
        r = O.ref_driver("optparam", k, n, p, lmin, lmax).strip()
        opt.append(dict(k=k, n=n, p=p, lmin=lmin, lmax=lmax, result=None if r == "throw" else [int(x) for x in r.split()]))
    mx = []
    for (p, lmin, lmax) in S.MAXKMER_KAT_INPUTS:
        mx.append(dict(p=p, lmin=lmin, lmax=lmax, result=int(O.ref_driver("maxkmers", p, lmin, lmax).strip())))
    return dict(optimal_bloom_param=opt, approximate_max_kmers=mx)


def read_bloom_bits(path, log2_len):
    data = open(path, "rb").read()
    nbytes = (1 << log2_len) // 8
    return data[:-nbytes], data[-nbytes:]


def gen_make_bloom(tmp):
    out = {}
    for name, case in S.MAKE_BLOOM_CASES.items():
        bases, offsets = S.make_bloom_reads(case)
        d = os.path.join(tmp, name)
        os.makedirs(d, exist_ok=True)
        acc = "SRR%06d" % (1 + list(S.MAKE_BLOOM_CASES).index(name))
        S.write_reads_file(os.path.join(d, acc + ".reads"), bases, offsets)
        r = O.ref_make_bloom(acc, d, d, case["k"], case["min_count"], case["p"], case["lmin"], case["lmax"], case["num_bp"])
        entry = dict(status=r["status"], num_kmer=r["num_kmer"], num_bp=r["num_bp"], log2_len=r["log_2_filter_len"],
                     num_hash=r["num_hash"], log2_count_len=r["log_2_counting_filter_len"], accession=acc)
        if r["status"] == 14:
            path = os.path.join(d, acc + ".bloom")
            header, bits = read_bloom_bits(path, r["log_2_filter_len"])
            entry.update(file_size=os.path.getsize(path), header_sha256=sha256(header), header_hex=header.hex(),
                         bits_sha256=sha256(bits), bits_crc32=O.crc32(bits), file_sha256=sha256(open(path, "rb").read()))
        out[name] = entry
        print("make_bloom", name, {k: v for k, v in entry.items() if k != "header_hex"})
    return out


def gen_build_db(tmp):
    out = {}
    for name, case in S.BUILD_DB_CASES.items():
        d = os.path.join(tmp, "db_" + name)
        os.makedirs(d, exist_ok=True)
        files = O.ref_driver("gen_blooms", d, case["n"], case["L"], case["k"], case["h"], case["seed"]).split()
        lst = os.path.join(d, "list.txt")
        open(lst, "w").write("\n".join(files) + "\n")
        db = os.path.join(d, "out.db")
        O.ref_driver("build_db", db, case["k"], case["L"], case["h"], lst)
        data = open(db, "rb").read()
        row = (case["n"] + 7) // 8
        nslice = (1 << case["L"]) * row
        entry = dict(file_size=len(data), file_sha256=sha256(data), header_hex=data[:44].hex(),
                     slices_sha256=sha256(data[44:44 + nslice]), slices_crc32=O.crc32(data[44:44 + nslice]),
                     tail_sha256=sha256(data[44 + nslice:]),
                     bloom0_sha256=sha256(open(files[0], "rb").read()))
        out[name] = entry
        print("build_db", name, {k: v for k, v in entry.items()})
    return out


def parse_csv(text):
    rows = []
    for line in text.splitlines():
        if not line or line.startswith("query,"):
            continue
        f = line.split(",")
        rows.append([f[0].strip('"'), int(f[1]), int(f[2]), f[4].strip('"')])
    return sorted(rows)


def gen_search(tmp):
    out = {}
    for name, case in S.SEARCH_CASES.items():
        d = os.path.join(tmp, "search_" + name)
        os.makedirs(d, exist_ok=True)
        if case["kind"] == "random":
            files = O.ref_driver("gen_blooms", d, case["n"], case["L"], case["k"], case["h"], case["seed"]).split()
            L, h = case["L"], case["h"]
        else:
            files = []
            L = h = None
            for j in range(case["n"]):
                acc = "SRR%07d" % (1000000 + j)
                bases, offsets = S.search_accession_reads(case, j)
                S.write_reads_file(os.path.join(d, acc + ".reads"), bases, offsets)
                r = O.ref_make_bloom(acc, d, d, case["k"], 1, 0.25, case["lmin"], case["lmax"], int(offsets[-1]))
                assert r["status"] == 14, r
                if L is None:
                    L, h = r["log_2_filter_len"], r["num_hash"]
                assert (L, h) == (r["log_2_filter_len"], r["num_hash"]), "accessions must share Bloom parameters"
                files.append(os.path.join(d, acc + ".bloom"))
        lst = os.path.join(d, "list.txt")
        open(lst, "w").write("\n".join(files) + "\n")
        db = os.path.join(d, "test.db")
        O.ref_driver("build_db", db, case["k"], L, h, lst)
        queries = S.search_queries(case)
        fa = os.path.join(d, "q.fa")
        with open(fa, "w") as f:
            for qn, qs in queries:
                f.write(">%s\n%s\n" % (qn, qs))
        res = {}
        for t in case["thresholds"]:
            r = O.ref_kwage(["-d", db, "-i", fa, "-t", repr(t), "--o.csv"], omp_threads=1)
            assert r.returncode == 0, r.stderr
            res[repr(t)] = parse_csv(r.stdout)
        data = open(db, "rb").read()
        out[name] = dict(L=L, h=h, db_sha256=sha256(data), db_size=len(data), results=res)
        print("search", name, "L", L, "h", h, {t: len(v) for t, v in res.items()})
    return out


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    with tempfile.TemporaryDirectory() as tmp:
        json.dump(gen_hash_kats(), open(os.path.join(OUT, "hash_kats.json"), "w"), indent=0)
        json.dump(gen_param_kats(), open(os.path.join(OUT, "param_kats.json"), "w"), indent=0)
        json.dump(gen_make_bloom(tmp), open(os.path.join(OUT, "make_bloom.json"), "w"), indent=1)
        json.dump(gen_build_db(tmp), open(os.path.join(OUT, "build_db.json"), "w"), indent=1)
        json.dump(gen_search(tmp), open(os.path.join(OUT, "search.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
