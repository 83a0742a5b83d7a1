"""GPU: the C++ host layer end to end -- make_bloom_filter / build_db / kwage CLI -- must produce the
reference's files byte for byte and the reference's search output (golden vectors came from the
unmodified reference; where oracle/_ref travelled to this box it is also run side by side)."""
import os
import subprocess

import numpy as np
import pytest

from kwage_b200 import hostapi as H
from kwage_b200.host import build as hbuild
from oracle import oracle_py as O
import synth_cases as S
import util
from conftest import load_golden

pytestmark = pytest.mark.gpu


def setup_module(_):
    hbuild.build()


def sha_file(path):
    return util.sha256(np.fromfile(path, dtype=np.uint8))


@pytest.mark.parametrize("name", ["uniform_k31", "ragged_k21", "k32", "k15_dups", "small_count_filter", "invalid_too_many", "no_kmers",
                                  "min_count_2", "min_count_5"])
def test_make_bloom_filter_writes_the_reference_file(name, tmp_path):
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    reads = str(tmp_path / (g["accession"] + ".reads"))
    S.write_reads_file(reads, bases, offsets)
    r = H.make_bloom_file(g["accession"], reads, case["num_bp"], str(tmp_path), k=case["k"], min_kmer_count=case["min_count"],
                          p=case["p"], min_log2=case["lmin"], max_log2=case["lmax"])
    assert r["status"] == g["status"], r
    assert r["log2_count_len"] == g["log2_count_len"]
    if g["status"] == 14:
        assert (r["num_kmer"], r["log2_len"], r["num_hash"]) == (g["num_kmer"], g["log2_len"], g["num_hash"])
        out = str(tmp_path / (g["accession"] + ".bloom"))
        assert os.path.getsize(out) == g["file_size"]
        assert sha_file(out) == g["file_sha256"]          # header, crc32, FilterInfo and bits all identical
    else:
        assert not os.path.exists(str(tmp_path / (g["accession"] + ".bloom")))
        if g["status"] == 16 and g["num_kmer"]:
            # STATUS_BLOOM_INVALID: the reference stops at the fragment that takes num_kmer beyond max_num_kmer
            # (make_bloom.cpp:208-214); the progress record at the abort is the compiled reference's
            assert (r["num_kmer"], r["num_bp"]) == (g["num_kmer"], g["num_bp"]), r
            assert (r["curr_read"], r["curr_fragment"]) == (g["num_bp"] // case["read_len"] - 1, 1)


@pytest.mark.parametrize("name,batch", [("ragged_k21", 4096), ("ragged_k21", 100), ("min_count_2", 33000), ("uniform_k31", 64)])
def test_make_bloom_filter_pipeline_batch_boundaries(name, batch, tmp_path, monkeypatch):
    """the parser thread packs batch n + 1 while the device works on batch n; with tiny batches (KWAGE_BATCH_BASES) every
    boundary case is hit: batches that end mid-byte of the 2na stream, fragments longer than a batch that go over in
    pieces overlapping by k - 1 bases, empty fragments at the cuts -- the file must still be the reference's"""
    monkeypatch.setenv("KWAGE_BATCH_BASES", str(batch))
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    reads = str(tmp_path / (g["accession"] + ".reads"))
    S.write_reads_file(reads, bases, offsets)
    r = H.make_bloom_file(g["accession"], reads, case["num_bp"], str(tmp_path), k=case["k"], min_kmer_count=case["min_count"],
                          p=case["p"], min_log2=case["lmin"], max_log2=case["lmax"])
    assert r["status"] == g["status"] == 14, r
    assert (r["num_kmer"], r["log2_len"], r["num_hash"]) == (g["num_kmer"], g["log2_len"], g["num_hash"])
    assert sha_file(str(tmp_path / (g["accession"] + ".bloom"))) == g["file_sha256"]


def test_make_bloom_filter_fastq_gz_input(tmp_path):
    """FASTQ.gz -> .bloom through the parser thread (the reference's SequenceIterator role, parse_sequence.cpp:72-262):
    same file as from the plain reads"""
    import gzip
    name = "uniform_k31"
    g = load_golden("make_bloom")[name]
    case = dict(S.MAKE_BLOOM_CASES[name])
    bases, offsets = S.make_bloom_reads(case)
    fq = str(tmp_path / (g["accession"] + ".fastq.gz"))
    with gzip.open(fq, "wb", compresslevel=1) as f:
        for r in range(len(offsets) - 1):
            seq = bytes(bases[int(offsets[r]): int(offsets[r + 1])])
            f.write(b"@r%d\n" % r + seq + b"\n+\n" + b"I" * len(seq) + b"\n")
    r = H.make_bloom_file(g["accession"], fq, case["num_bp"], str(tmp_path), k=case["k"], min_kmer_count=case["min_count"],
                          p=case["p"], min_log2=case["lmin"], max_log2=case["lmax"])
    assert r["status"] == 14, r
    assert sha_file(str(tmp_path / (g["accession"] + ".bloom"))) == g["file_sha256"]


def test_make_bloom_filter_default_min_count_5_no_cpu_path(tmp_path):
    # the reference's default --min-kmer-count is 5 (options.h); the device path must be the one that ran
    from kwage_b200 import capi
    case = dict(S.MAKE_BLOOM_CASES["min_count_5"])
    bases, offsets = S.make_bloom_reads(case)
    reads = str(tmp_path / "SRR000007.reads")
    S.write_reads_file(reads, bases, offsets)
    before = capi.launch_count()
    r = H.make_bloom_file("SRR000007", reads, case["num_bp"], str(tmp_path), k=31, min_kmer_count=5, max_log2=24)
    assert r["status"] == H.STATUS_BLOOM_SUCCESS, r
    assert r["num_kmer"] == load_golden("make_bloom")["min_count_5"]["num_kmer"]
    assert capi.launch_count() > before


@pytest.mark.parametrize("name", list(S.BUILD_DB_CASES))
def test_build_db_writes_the_reference_file(name, tmp_path):
    g = load_golden("build_db")[name]
    case = S.BUILD_DB_CASES[name]
    files = []
    for j, bits in enumerate(S.build_db_filters(case)):
        path = str(tmp_path / (util.fixture_accession(j) + ".bloom"))
        assert H.write_bloom_file(path, util.fixture_accession(j), case["k"], case["L"], case["h"], bits)
        files.append(path)
    assert sha_file(files[0]) == g["bloom0_sha256"]        # the .bloom writer is byte-exact too
    db = str(tmp_path / "out.db")
    assert H.build_db(db, case["k"], case["L"], case["h"], files)
    data = np.fromfile(db, dtype=np.uint8)
    assert len(data) == g["file_size"]
    assert bytes(data[:44]).hex() == g["header_hex"]       # incl. crc32 of the slices and info_start
    assert util.sha256(data) == g["file_sha256"]


def test_build_db_rejects_bad_inputs(tmp_path):
    case = S.BUILD_DB_CASES["n8_L16"]
    files = []
    for j, bits in enumerate(S.build_db_filters(case)):
        path = str(tmp_path / (util.fixture_accession(j) + ".bloom"))
        H.write_bloom_file(path, util.fixture_accession(j), case["k"], case["L"], case["h"], bits)
        files.append(path)
    db = str(tmp_path / "bad.db")
    assert not H.build_db(db, case["k"], case["L"], case["h"] + 1, files)           # inconsistent parameters
    assert not H.build_db(db, case["k"], case["L"], case["h"], files + [str(tmp_path / "missing.bloom")])
    assert not H.build_db(db, case["k"], case["L"], case["h"], [])                  # empty inventory
    raw = bytearray(open(files[3], "rb").read())
    raw[-5] ^= 0x10                                                                  # flip one filter bit -> crc mismatch
    open(files[3], "wb").write(bytes(raw))
    assert not H.build_db(db, case["k"], case["L"], case["h"], files)
    raw[0] = 0x00                                                                    # "in progress" guard byte
    open(files[3], "wb").write(bytes(raw))
    assert not H.build_db(db, case["k"], case["L"], case["h"], files)


def parse_csv(text):
    rows = []
    for line in text.splitlines():
        if not line or line.startswith("query,"):
            continue
        f = line.split(",")
        rows.append([f[0].strip('"'), int(f[1]), int(f[2]), f[4].strip('"')])
    return sorted(rows)


@pytest.mark.parametrize("name", list(S.SEARCH_CASES))
def test_kwage_cli_matches_reference_output(name, tmp_path):
    g = load_golden("search")[name]
    case = S.SEARCH_CASES[name]
    dbd = util.search_case_db(name)
    files = []
    for j, bits in enumerate(dbd["filters"]):
        path = str(tmp_path / (util.fixture_accession(j) + ".bloom"))
        H.write_bloom_file(path, util.fixture_accession(j), dbd["k"], dbd["L"], dbd["h"], bits)
        files.append(path)
    db = str(tmp_path / "test.db")
    assert H.build_db(db, dbd["k"], dbd["L"], dbd["h"], files)
    assert sha_file(db) == g["db_sha256"]                  # same database file as the reference built
    fa = str(tmp_path / "q.fa")
    with open(fa, "w") as f:
        for qn, qs in S.search_queries(case):
            f.write(">%s\n%s\n" % (qn, qs))
    for t in case["thresholds"]:
        r = subprocess.run([H.KWAGE_BIN, "-d", str(tmp_path), "-i", fa, "-t", repr(t), "--o.csv"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert parse_csv(r.stdout) == g["results"][repr(t)], (name, t)
        if O.have_ref():                                   # the real reference binary on the same inputs
            ref = O.ref_kwage(["-d", db, "-i", fa, "-t", repr(t), "--o.csv"], omp_threads=1)
            assert ref.returncode == 0
            assert sorted(ref.stdout.splitlines()) == sorted(r.stdout.splitlines())
    # command-line sequences and JSON output
    q = S.search_queries(case)[0][1]
    r = subprocess.run([H.KWAGE_BIN, "-d", db, "-t", "0.2", "--o.json", q], capture_output=True, text=True)
    assert r.returncode == 0
    if name == "accessions":
        assert '"query": "command line seq 0"' in r.stdout and '"run": "SRR1000003"' in r.stdout
    if O.have_ref():
        ref = O.ref_kwage(["-d", db, "-t", "0.2", "--o.json", q], omp_threads=1)
        strip = lambda s: sorted(x.strip().rstrip(",") for x in s.splitlines())   # ties are unordered in both
        assert strip(ref.stdout) == strip(r.stdout)


def test_kwage_cli_several_files_share_one_slab(tmp_path):
    # a directory of .db files with equal Bloom parameters (uneven filter counts: any bit offset) is searched as ONE column
    # slab by default and file by file with --max-slab-gib 0; both must print the same matches, and so must the reference
    k, L, h = 31, 14, 3
    widths = [13, 64, 21]
    dbs, j = [], 0
    dbdir = tmp_path / "dbs"
    dbdir.mkdir()
    for fi, w in enumerate(widths):
        files = []
        for _ in range(w):
            path = str(tmp_path / (util.fixture_accession(j) + ".bloom"))
            bases, offsets = S.uniform_reads(900 + j, 0, 40, 120)
            bits, _ = O.raw_insert(bases, offsets, k, h, L)
            H.write_bloom_file(path, util.fixture_accession(j), k, L, h, bits)
            files.append(path)
            j += 1
        db = str(dbdir / ("part%d.db" % fi))
        assert H.build_db(db, k, L, h, files)
        dbs.append(db)
    fa = str(tmp_path / "q.fa")
    with open(fa, "w") as f:
        for acc in (0, 12, 13, 76, 77, 97):                    # first / last filters of every file
            bases, _ = S.uniform_reads(900 + acc, 0, 2, 120)
            f.write(">from_%d\n%s\n" % (acc, bytes(bases).decode()))
        f.write(">random\n%s\n" % bytes(O.gen_reads(1, 0, 1, 500)).decode())
    outs = []
    # one slab; file by file; file by file shared out over two host threads on the same device
    from kwage_b200 import capi
    variants = [[], ["--max-slab-gib", "0"], ["--max-slab-gib", "0", "--device", "0,0"]]
    if capi.device_count() >= 2:
        # two devices: one round of two slabs through kwg_search_gather (NCCL), the third slab on the plain path
        variants.append(["--max-slab-gib", "0", "--device", "0,1"])
    for extra in variants:
        r = subprocess.run([H.KWAGE_BIN, "-d", str(dbdir), "-i", fa, "-t", "0.5", "--o.csv"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(parse_csv(r.stdout))
    assert all(o == outs[0] for o in outs) and len(outs[0]) >= 6
    found = {(q, acc) for q, _, _, acc in outs[0]}
    for acc in (0, 12, 13, 76, 77, 97):
        assert ("from_%d" % acc, util.fixture_accession(acc)) in found
    if O.have_ref():
        ref = O.ref_kwage(["-d", str(dbdir), "-i", fa, "-t", "0.5", "--o.csv"], omp_threads=1)
        assert ref.returncode == 0 and parse_csv(ref.stdout) == outs[0]


def test_kwage_cli_errors():
    r = subprocess.run([H.KWAGE_BIN, "-i", "nothing.fa"], capture_output=True, text=True)
    assert r.returncode != 0 and "database" in r.stderr
    r = subprocess.run([H.KWAGE_BIN, "-d", "/nonexistent.db", "ACGT"], capture_output=True, text=True)
    assert r.returncode != 0
