"""CPU: host-layer logic (C++ kwage_b200/host) against the reference's golden vectors."""
import numpy as np

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
