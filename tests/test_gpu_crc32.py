"""GPU parity: the device crc32 (crc32.cu) against zlib's -- the function the reference calls for the filter bits of a
.bloom file (bloom.cpp:328-336) and for the source filters / slice region of a .db file (build_db.cpp:281-282,307)."""
import zlib

import numpy as np
import pytest
import torch

from kwage_b200 import capi
from oracle import oracle_py as O
import synth_cases as S

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n_bytes", [4, 8, 60, 64, 68, 16380, 16384, 16388, 65536, 1 << 20, (1 << 20) + 4, 3 * 16384 + 12,
                                     1024 * 16384, 1025 * 16384 + 8, 64 << 20])
def test_flat_buffer_equals_zlib(n_bytes):
    rng = np.random.default_rng(n_bytes)
    a = rng.integers(0, 256, n_bytes, dtype=np.uint8)
    d = dev(a)
    assert capi.crc32_dev(d.data_ptr(), 1, n_bytes, n_bytes) == zlib.crc32(a.tobytes())
    # the oracle's crc32 (the function the golden files were pinned with) agrees
    if n_bytes <= 1 << 20:
        assert O.crc32(a) == zlib.crc32(a.tobytes())


def test_running_value_chains_like_crc32_z():
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, 5 * 16384 + 24, dtype=np.uint8)
    d = dev(a)
    cuts = [0, 4, 16384, 16388, 40000, len(a)]
    crc = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        crc = capi.crc32_dev(d.data_ptr() + lo, 1, hi - lo, hi - lo, crc)
        assert crc == zlib.crc32(a[:hi].tobytes())
    # zeros and all-ones (the register must survive long runs of either)
    z = torch.zeros(1 << 22, dtype=torch.uint8, device="cuda")
    assert capi.crc32_dev(z.data_ptr(), 1, 1 << 22, 1 << 22) == zlib.crc32(bytes(1 << 22))
    z.fill_(255)
    assert capi.crc32_dev(z.data_ptr(), 1, 1 << 22, 1 << 22, 0x12345678) == zlib.crc32(b"\xff" * (1 << 22), 0x12345678)


@pytest.mark.parametrize("n_rows,row_bytes,pitch", [(1000, 4, 16), (4096, 36, 48), (257, 256, 256), (70000, 12, 16), (3, 100000, 100016)])
def test_pitched_rows_equal_zlib_of_the_packed_rows(n_rows, row_bytes, pitch):
    rng = np.random.default_rng(n_rows)
    a = rng.integers(0, 256, (n_rows, pitch), dtype=np.uint8)
    d = dev(a)
    exp = zlib.crc32(np.ascontiguousarray(a[:, :row_bytes]).tobytes(), 99)
    assert capi.crc32_dev(d.data_ptr(), n_rows, row_bytes, pitch, 99) == exp


def test_bad_shapes_are_refused():
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for args in [(1, 6, 6), (2, 4, 6), (0, 4, 4)]:
        with pytest.raises(capi.KwageError):
            capi.crc32_dev(d.data_ptr(), *args)
    with pytest.raises(capi.KwageError):
        capi.crc32_dev(d.data_ptr() + 2, 1, 8, 8)


@pytest.mark.parametrize("k,nh,L", [(31, 3, 20), (21, 5, 26), (31, 2, 5)])
def test_finalize_crc_raw_and_counting(k, nh, L):
    bases, offsets = S.uniform_reads(5, 0, 3000, 150)
    with capi.BloomBuilder(k, raw_num_hash=nh, raw_log2_len=L) as b:
        b.add_reads(bases, offsets)
        bits, crc = b.finalize_crc()
        assert np.array_equal(bits, b.finalize()) and crc == zlib.crc32(bits.tobytes())
    with capi.BloomBuilder(k, min_kmer_count=1, log2_count_len=20, log2_max_len=26) as b:
        b.add_reads(bases, offsets)
        bits, crc = b.finalize_crc(L, min(nh, 5))
        assert np.array_equal(bits, b.finalize(L, min(nh, 5))) and crc == zlib.crc32(bits.tobytes())


@pytest.mark.parametrize("n,bits", [(32, 1024), (2048, 4096), (96, 32 * 1000), (64, 1 << 20)])
def test_transpose_crc_advances_both_running_values(n, bits):
    rng = np.random.default_rng(n + bits)
    filters = [rng.integers(0, 256, bits // 8 * 2, dtype=np.uint8) for _ in range(n)]
    first = [f[: bits // 8] for f in filters]
    second = [f[bits // 8:] for f in filters]
    s1, fc, dc = capi.transpose_crc(first, bits, np.zeros(n, np.uint32), 0)
    assert np.array_equal(s1, capi.transpose(first, bits))
    assert list(fc) == [zlib.crc32(f.tobytes()) for f in first] and dc == zlib.crc32(s1.tobytes())
    s2, fc2, dc2 = capi.transpose_crc(second, bits, fc, dc)
    assert list(fc2) == [zlib.crc32(f.tobytes()) for f in filters]
    assert dc2 == zlib.crc32(s2.tobytes(), zlib.crc32(s1.tobytes()))
    # filter values alone work for any column count; the slice value needs whole 32-bit words per row
    odd = first[: n - 3] if n > 3 else first
    s3, fc3, _ = capi.transpose_crc(odd, bits, np.zeros(len(odd), np.uint32), None)
    assert np.array_equal(s3, capi.transpose(odd, bits)) and list(fc3) == [zlib.crc32(f.tobytes()) for f in odd]
    with pytest.raises(capi.KwageError):
        capi.transpose_crc(odd, bits, None, 0)
