"""CPU: the k <= 63 restatement (oracle kwo_raw_insert_wide, 128-bit words) -- PARITY UNPINNED beyond the reference's
k <= 32 (word.h:10).  What can be pinned is pinned: for every k <= 32 it must give exactly what the narrow restatement
(itself checked against the compiled reference's golden vectors) gives, and for k > 32 it is checked against an
independent pure-Python statement of the same rules on small inputs."""
import numpy as np
import pytest

from oracle import oracle_py as O
import synth_cases as S


def ragged_case(seed, n_reads=120, max_len=140):
    flat = S.mutate(O.gen_reads(seed, 0, n_reads, max_len), seed, n_rate=41, lower_rate=5)
    return S.ragged(flat, seed, n_reads, 0, max_len)


@pytest.mark.parametrize("k,nh,L", [(1, 2, 8), (4, 8, 10), (15, 7, 16), (21, 5, 18), (31, 3, 20), (32, 4, 22)])
def test_wide_restatement_equals_narrow_up_to_k32(k, nh, L):
    bases, offsets = ragged_case(500 + k)
    a, na = O.raw_insert(bases, offsets, k, nh, L)
    b, nb = O.raw_insert_wide(bases, offsets, k, nh, L)
    assert na == nb and np.array_equal(a, b)


def murmur3_py(data, seed):
    c1, c2, M = 0xcc9e2d51, 0x1b873593, 0xFFFFFFFF
    rotl = lambda x, r: ((x << r) | (x >> (32 - r))) & M
    h = seed
    nb = len(data) // 4
    for i in range(nb):
        k1 = int.from_bytes(data[4 * i: 4 * i + 4], "little")
        k1 = (k1 * c1) & M; k1 = rotl(k1, 15); k1 = (k1 * c2) & M
        h ^= k1; h = rotl(h, 13); h = (h * 5 + 0xe6546b64) & M
    tail = data[4 * nb:]
    if tail:
        k1 = int.from_bytes(tail, "little")
        k1 = (k1 * c1) & M; k1 = rotl(k1, 15); k1 = (k1 * c2) & M
        h ^= k1
    h ^= len(data)
    h ^= h >> 16; h = (h * 0x85ebca6b) & M; h ^= h >> 13; h = (h * 0xc2b2ae35) & M; h ^= h >> 16
    return h


def raw_insert_py(bases, offsets, k, nh, L):
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    bits = np.zeros((1 << L) // 8, dtype=np.uint8)
    n = 0
    for r in range(len(offsets) - 1):
        read = bytes(bases[int(offsets[r]): int(offsets[r + 1])]).decode("latin1")
        for p in range(len(read) - k + 1):
            w = read[p: p + k].upper()
            if any(ch not in "ACGT" for ch in w):
                continue
            rc = "".join(comp[ch] for ch in reversed(w))
            canon = min(w, rc)                      # A < C < G < T in ASCII as in the 2-bit code: same order as the integers
            for h in range(nh):
                b = murmur3_py(canon.encode(), h) & ((1 << L) - 1)
                bits[b >> 3] |= 1 << (b & 7)
            n += 1
    return bits, n


@pytest.mark.parametrize("k,nh,L", [(33, 3, 14), (40, 2, 12), (47, 5, 16), (63, 1, 10), (31, 3, 14), (32, 2, 12)])
def test_wide_restatement_equals_pure_python(k, nh, L):
    bases, offsets = ragged_case(900 + k, n_reads=40, max_len=150)
    a, na = O.raw_insert_wide(bases, offsets, k, nh, L)
    b, nb = raw_insert_py(bases, offsets, k, nh, L)
    assert na == nb and na > 0 and np.array_equal(a, b)
