"""GPU parity: kwg_transpose through the C ABI vs the oracle and the reference's golden digests."""
import numpy as np
import pytest

from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S
import util
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(S.BUILD_DB_CASES))
def test_transpose_matches_reference_golden(name):
    g = load_golden("build_db")[name]
    case = S.BUILD_DB_CASES[name]
    filters = S.build_db_filters(case)
    got = capi.transpose(filters, 1 << case["L"])
    assert got.shape == (1 << case["L"], (case["n"] + 7) // 8)
    assert util.sha256(got) == g["slices_sha256"]
    assert O.crc32(got.reshape(-1)) == g["slices_crc32"]


@pytest.mark.parametrize("n,bits", [(1, 8), (3, 24), (7, 1000), (8, 1024), (9, 4104), (31, 32), (33, 40), (257, 4096 + 8),
                                    (255, 2048), (256, 3072), (511, 1032), (1000, 520), (2049, 264)])
def test_transpose_ragged_shapes(n, bits):
    rng = np.random.default_rng(n * 131 + bits)
    filters = [rng.integers(0, 256, bits // 8, dtype=np.uint8) for _ in range(n)]
    exp = O.transpose(filters, bits)
    got = capi.transpose(filters, bits)
    assert np.array_equal(got, exp)


def test_transpose_chunked_like_build_db():
    # build_db.cpp:259 walks the filter in chunks of max_buffer_slice; each chunk is one kwg_transpose call
    case = S.BUILD_DB_CASES["n300_L16"]
    filters = S.build_db_filters(case)
    whole = O.transpose(filters, 1 << case["L"])
    chunk = 1 << 13
    parts = [capi.transpose([f[o // 8: (o + chunk) // 8] for f in filters], chunk) for o in range(0, 1 << case["L"], chunk)]
    assert np.array_equal(np.concatenate(parts), whole)


def test_transpose_round_trip_is_identity():
    # size-independent property: transposing the slices back gives the filters (bit-matrix involution)
    n, bits = 4096, 1 << 15
    filters = [O.gen_filter_bits(77, j, bits // 8) for j in range(n)]
    slices = capi.transpose(filters, bits)                     # (bits, n/8)
    back = capi.transpose([slices[r] for r in range(bits)], n)  # (n, bits/8)
    assert np.array_equal(back, np.stack(filters))


def test_transpose_bad_arguments():
    with pytest.raises(capi.KwageError):
        capi.transpose([np.zeros(2, np.uint8)], 12)            # not a multiple of 8
    with pytest.raises(capi.KwageError):
        capi.transpose([], 64)


def test_transpose_from_several_threads_and_after_releasing_the_caches():
    """kwg_transpose keeps its streams and staging buffers in a per-device cache behind a per-device lock: callers on
    one device take turns (and get the right answer each), kwg_release_caches() frees the buffers and the next call
    simply allocates again"""
    import threading
    rng = np.random.default_rng(5)
    cases = []
    for t in range(4):
        n, bits = int(rng.integers(1, 300)), 8 * int(rng.integers(1, 600))
        filters = [rng.integers(0, 256, bits // 8, dtype=np.uint8) for _ in range(n)]
        cases.append((filters, bits, O.transpose(filters, bits)))
    out = [None] * len(cases)

    def work(i):
        for _ in range(3):
            out[i] = capi.transpose(cases[i][0], cases[i][1])
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i, (_, _, exp) in enumerate(cases):
        assert np.array_equal(out[i], exp)
    capi.lib().kwg_release_caches()
    assert np.array_equal(capi.transpose(cases[0][0], cases[0][1]), cases[0][2])
