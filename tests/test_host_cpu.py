"""CPU: host-layer logic (C++ kwage_b200/host) against the reference's golden vectors."""
import numpy as np
import pytest

from kwage_b200 import hostapi as H
from kwage_b200.host import build as hbuild
import synth_cases as S
from conftest import load_golden


def setup_module(_):
    hbuild.build()


def test_optimal_bloom_param_matches_reference():
    g = load_golden("param_kats")
    for e in g["optimal_bloom_param"]:
        r = H.optimal_bloom_param(e["k"], e["n"], e["p"], e["lmin"], e["lmax"])
        assert (None if r is None else list(r)) == e["result"], e
    for e in g["approximate_max_kmers"]:
        assert H.approximate_max_kmers(e["p"], e["lmin"], e["lmax"]) == e["result"]


def test_counting_filter_length_matches_reference():
    g = load_golden("make_bloom")
    for name, e in g.items():
        case = dict(S.MAKE_BLOOM_CASES[name])
        S.make_bloom_reads(case)                       # resolves num_bp == -1 to the true base count
        assert H.counting_filter_log2_len(case["num_bp"]) == e["log2_count_len"], name
    assert H.counting_filter_log2_len(0) == 32          # no metadata (make_bloom.cpp:106)
    assert H.counting_filter_log2_len(1) == 18
    assert H.counting_filter_log2_len(10 ** 12) == 32


def test_accession_packing():
    # the packed values appear in the reference's files: bytes 21..28 of every golden .bloom header
    g = load_golden("make_bloom")
    for e in g.values():
        if "header_hex" in e:
            packed = int.from_bytes(bytes.fromhex(e["header_hex"])[21:29], "little")
            assert H.str_to_accession(e["accession"]) == packed
            assert H.accession_to_str(packed) == e["accession"]
    for s in ["SRR1", "ERR0000000001", "DRR000347", "srr12345"]:
        a = H.str_to_accession(s)
        assert a != 0 and H.accession_to_str(a) == s.upper()
    assert H.str_to_accession("SR12") == 0 and H.str_to_accession("SRRX") == 0      # reference throws


def test_host_packer_matches_the_numpy_packer():
    """stages.cpp::pack_2na (what the parser thread of make_bloom_filter runs) against capi.pack_2na on fragments of
    every length and alignment, with N, IUPAC codes and lower case inside"""
    from kwage_b200 import capi
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNnRY-*", dtype=np.uint8)
    for trial in range(30):
        lens = rng.integers(0, 41, size=rng.integers(1, 30))
        if trial % 3 == 0:
            frs = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)] for n in lens]       # no bad bases at all
        else:
            frs = [alphabet[rng.integers(0, len(alphabet), size=n)] for n in lens]
        flat = np.concatenate(frs) if len(frs) else np.zeros(0, np.uint8)
        packed, mask, any_bad = H.pack_2na([bytes(f) for f in frs])
        exp_p, exp_m = capi.pack_2na(flat)
        n = len(flat)
        assert np.array_equal(packed, exp_p[: (n + 3) // 4])
        if exp_m is None:
            assert not any_bad and not mask.any()
        else:
            assert any_bad and np.array_equal(mask, exp_m[: (n + 7) // 8])


def _fnv(frags):
    h = 1469598103934665603
    for f in frags:
        for c in f:
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        h = ((h ^ 0xFF) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _digest(frags):
    return (len(frags), sum(len(f) for f in frags), max([len(f) for f in frags] or [0]), _fnv(frags))


def _rand_frags(rng, n, max_len):
    alphabet = np.frombuffer(b"ACGTNacgtn", dtype=np.uint8)
    return [bytes(alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, max_len + 1)))]) for _ in range(n)]


@pytest.mark.parametrize("fmt", ["fastq", "fastq.gz", "fasta", "fasta.gz", "fa_crlf", "fastq_no_final_newline", "fasta_long.gz"])
def test_read_streaming_equals_a_plain_parse(fmt, tmp_path):
    """FASTA / FASTQ (.gz) streaming of the host layer (stages.cpp::GzSequenceReads: an inflate thread hands 4 MiB blocks to
    the line cutter; the SequenceIterator role, parse_sequence.cpp:72-262): the fragments are the ones a plain Python parse
    of the same file yields -- lines that straddle blocks, multi-line FASTA records, CRLF, a last line without a newline,
    empty reads, files of several blocks."""
    import gzip
    rng = np.random.default_rng(len(fmt))
    big = fmt == "fasta_long.gz"
    frags = _rand_frags(rng, 40 if big else 3000, 600000 if big else 300)
    eol = b"\r\n" if fmt == "fa_crlf" else b"\n"
    out = bytearray()
    if fmt.startswith("fastq"):
        for i, f in enumerate(frags):
            out += b"@r%d" % i + eol + f + eol + b"+" + eol + b"I" * len(f) + eol
        if fmt == "fastq_no_final_newline":
            out = out[:-1]
        exp = frags
    else:
        for i, f in enumerate(frags):
            out += b">s%d some text" % i + eol
            w = int(rng.choice([60, 70, 1000]))
            for a in range(0, len(f), w):
                out += f[a:a + w] + eol
        exp = frags
    ext = ".fastq" if fmt.startswith("fastq") else ".fasta"
    path = str(tmp_path / ("x" + ext + (".gz" if fmt.endswith(".gz") else "")))
    if fmt.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(bytes(out))
    else:
        with open(path, "wb") as f:
            f.write(bytes(out))
    assert len(out) > (9 << 20) or not big           # the long case spans several 4 MiB blocks
    assert H.parse_digest(path) == _digest(exp)


def test_read_streaming_of_a_reads_file_and_of_an_empty_file(tmp_path):
    rng = np.random.default_rng(3)
    frags = _rand_frags(rng, 500, 200)
    p1 = str(tmp_path / "a.reads")
    with open(p1, "wb") as f:
        f.write(b"".join(x + b"\n" for x in frags))
    assert H.parse_digest(p1) == _digest(frags)
    p2 = str(tmp_path / "e.fastq")
    open(p2, "wb").close()
    assert H.parse_digest(p2) == (0, 0, 0, _fnv([]))
