"""ctypes binding of oracle/libkwage_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
It also wraps the compiled UNMODIFIED reference under oracle/_ref/ (ref_driver, kwage).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkwage_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_lib = None


def build():
    """Compile the C restatement and, when /root/reference is present, the reference itself."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    # the reference's host code with the build_db shim of INTEGRATION.md, linked against the product library (when built)
    if os.path.exists(os.path.join(HERE, "..", "kwage_b200", "lib", "libkwage_cuda.so")):
        subprocess.run(["make", "-s", "-C", HERE, "shim"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u64, u32, p = C.c_uint64, C.c_uint32, C.c_void_p
        L.kwo_rnd.restype = u64
        L.kwo_rnd.argtypes = [u64, u64, u64]
        L.kwo_gen_reads.argtypes = [u64, u64, u64, u32, p]
        L.kwo_gen_filter_bits.argtypes = [u64, u64, u64, p]
        L.kwo_gen_reads_mt64.argtypes = [u64, u64, u32, p]
        L.kwo_canonical_kmers.restype = u64
        L.kwo_canonical_kmers.argtypes = [p, u64, u32, p, p]
        L.kwo_murmur3_word.restype = u32
        L.kwo_murmur3_word.argtypes = [u64, u32, u32]
        L.kwo_murmur3_bytes.restype = u32
        L.kwo_murmur3_bytes.argtypes = [p, u32, u32]
        L.kwo_crc32.restype = u32
        L.kwo_crc32.argtypes = [p, u64]
        L.kwo_raw_insert.restype = u64
        L.kwo_raw_insert.argtypes = [p, p, u64, u32, u32, u32, p]
        L.kwo_raw_insert_wide.restype = u64
        L.kwo_raw_insert_wide.argtypes = [p, p, u64, u32, u32, u32, p]
        L.kwo_optimal_bloom_param.restype = C.c_int
        L.kwo_optimal_bloom_param.argtypes = [u64, C.c_float, u32, u32, C.POINTER(u32), C.POINTER(u32)]
        L.kwo_approximate_max_kmers.restype = u64
        L.kwo_approximate_max_kmers.argtypes = [C.c_float, u32, u32]
        L.kwo_counting_log2_len.restype = u32
        L.kwo_counting_log2_len.argtypes = [u64]
        L.kwo_builder_create.restype = p
        L.kwo_builder_create.argtypes = [u32, u32, u32, u32]
        L.kwo_builder_destroy.argtypes = [p]
        L.kwo_builder_add_reads.argtypes = [p, p, p, u64]
        L.kwo_builder_add_reads_limit.restype = u64
        L.kwo_builder_add_reads_limit.argtypes = [p, p, p, u64, u64]
        L.kwo_builder_num_valid.restype = u64
        L.kwo_builder_num_valid.argtypes = [p]
        L.kwo_builder_finalize.argtypes = [p, u32, u32, p]
        L.kwo_transpose.argtypes = [p, u32, u64, p]
        L.kwo_query_kmers.restype = u64
        L.kwo_query_kmers.argtypes = [p, u64, u32, p]
        L.kwo_search_counts.restype = u64
        L.kwo_search_counts.argtypes = [p, u32, u32, u32, u32, p, u64, p]
        L.kwo_search_matches.restype = u64
        L.kwo_search_matches.argtypes = [p, u32, u32, u32, u32, p, u64, C.c_float, p, p, C.POINTER(u32)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _bytes_arr(seq):
    if isinstance(seq, str):
        seq = seq.encode()
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


# ---------------------------------------------------------------- synthetic data
def gen_reads(seed, first_read, n_reads, read_len):
    out = np.empty(n_reads * read_len, dtype=np.uint8)
    lib().kwo_gen_reads(seed, first_read, n_reads, read_len, _ptr(out))
    return out


def gen_reads_mt64(seed, n_reads, read_len):
    out = np.empty(n_reads * read_len, dtype=np.uint8)
    lib().kwo_gen_reads_mt64(seed, n_reads, read_len, _ptr(out))
    return out


def gen_filter_bits(seed, filter_index, n_bytes):
    assert n_bytes % 8 == 0
    out = np.empty(n_bytes // 8, dtype=np.uint64)
    lib().kwo_gen_filter_bits(seed, filter_index, n_bytes // 8, _ptr(out))
    return out.view(np.uint8)


# ---------------------------------------------------------------- word / hash
def canonical_kmers(seq, k):
    s = _bytes_arr(seq)
    words = np.empty(max(len(s), 1), dtype=np.uint64)
    loc5 = np.empty(max(len(s), 1), dtype=np.uint64)
    n = lib().kwo_canonical_kmers(_ptr(s), len(s), k, _ptr(words), _ptr(loc5))
    return words[:n].copy(), loc5[:n].copy()


def murmur3_word(word, k, seed):
    return lib().kwo_murmur3_word(int(word), k, seed)


def murmur3_bytes(data, seed):
    d = _bytes_arr(data)
    return lib().kwo_murmur3_bytes(_ptr(d), len(d), seed)


def crc32(buf):
    b = _bytes_arr(buf)
    return lib().kwo_crc32(_ptr(b), len(b))


# ---------------------------------------------------------------- construction
def raw_insert(bases, offsets, k, num_hash, log2_len, bits=None):
    bases = _bytes_arr(bases)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if bits is None:
        bits = np.zeros(max((1 << log2_len) // 8, 1), dtype=np.uint8)
    n = lib().kwo_raw_insert(_ptr(bases), _ptr(offsets), len(offsets) - 1, k, num_hash, log2_len, _ptr(bits))
    return bits, n


def raw_insert_wide(bases, offsets, k, num_hash, log2_len, bits=None):
    """k up to 63 on a 128-bit word: PARITY UNPINNED (the reference stops at k = 32, word.h:10); equals raw_insert for k <= 32"""
    bases = _bytes_arr(bases)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if bits is None:
        bits = np.zeros(max((1 << log2_len) // 8, 1), dtype=np.uint8)
    n = lib().kwo_raw_insert_wide(_ptr(bases), _ptr(offsets), len(offsets) - 1, k, num_hash, log2_len, _ptr(bits))
    return bits, n


def optimal_bloom_param(num_kmer, p, min_log2, max_log2):
    """Returns (log2_len, num_hash) or None where the reference throws."""
    L, h = C.c_uint32(0), C.c_uint32(0)
    rc = lib().kwo_optimal_bloom_param(num_kmer, p, min_log2, max_log2, C.byref(L), C.byref(h))
    return None if rc != 0 else (L.value, h.value)


def approximate_max_kmers(p, min_log2, max_log2):
    return lib().kwo_approximate_max_kmers(p, min_log2, max_log2)


def counting_log2_len(num_bp):
    return lib().kwo_counting_log2_len(num_bp)


class Builder:
    """Sequential restatement of the reference's counting-filter construction."""

    def __init__(self, k, min_count, log2_count_len, log2_max_len):
        self.h = lib().kwo_builder_create(k, min_count, log2_count_len, log2_max_len)
        if not self.h:
            raise MemoryError("kwo_builder_create")

    def add_reads(self, bases, offsets):
        bases = _bytes_arr(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lib().kwo_builder_add_reads(self.h, _ptr(bases), _ptr(offsets), len(offsets) - 1)

    def add_reads_limit(self, bases, offsets, max_num_kmer):
        """Stops like the reference as soon as num_kmer > max_num_kmer; returns reads consumed."""
        bases = _bytes_arr(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        return lib().kwo_builder_add_reads_limit(self.h, _ptr(bases), _ptr(offsets), len(offsets) - 1, max_num_kmer)

    def num_valid(self):
        return lib().kwo_builder_num_valid(self.h)

    def finalize(self, log2_len, num_hash):
        out = np.empty(max((1 << log2_len) // 8, 1), dtype=np.uint8)
        lib().kwo_builder_finalize(self.h, log2_len, num_hash, _ptr(out))
        return out

    def close(self):
        if self.h:
            lib().kwo_builder_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def make_bloom(bases, offsets, k, min_count, p, min_log2, max_log2, num_bp):
    """Restatement of make_bloom_filter()'s numeric result (reference make_bloom.cpp:76-448).
    Returns dict(status, num_kmer, log2_len, num_hash, log2_count_len, bits)."""
    lc = counting_log2_len(num_bp)
    max_kmer = approximate_max_kmers(p, min_log2, max_log2)
    b = Builder(k, min_count, lc, max_log2)
    try:
        b.add_reads_limit(bases, offsets, max_kmer)
        n = b.num_valid()
        res = dict(status="invalid", num_kmer=n, log2_count_len=lc, log2_len=0, num_hash=0, bits=None)
        if n > max_kmer:
            return res
        param = optimal_bloom_param(n, p, min_log2, max_log2)
        if param is None:
            return res
        res.update(status="success", log2_len=param[0], num_hash=param[1], bits=b.finalize(*param))
        return res
    finally:
        b.close()


# ---------------------------------------------------------------- transpose / search
def transpose(filters, chunk_bits):
    """filters: list of uint8 arrays with >= chunk_bits/8 bytes each -> (chunk_bits, ceil(n/8)) uint8."""
    n = len(filters)
    keep = [np.ascontiguousarray(f, dtype=np.uint8) for f in filters]
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in keep])
    row = (n + 7) // 8
    dest = np.empty(chunk_bits * row, dtype=np.uint8)
    lib().kwo_transpose(ptrs, n, chunk_bits, _ptr(dest))
    return dest.reshape(chunk_bits, row)


def query_kmers(query, k):
    q = _bytes_arr(query)
    out = np.empty(max(len(q), 1), dtype=np.uint64)
    n = lib().kwo_query_kmers(_ptr(q), len(q), k, _ptr(out))
    return out[:n].copy()


def search_counts(slices, n_filters, log2_len, num_hash, k, query):
    q = _bytes_arr(query)
    s = np.ascontiguousarray(slices, dtype=np.uint8)
    counts = np.zeros(n_filters, dtype=np.uint32)
    n = lib().kwo_search_counts(_ptr(s), n_filters, log2_len, num_hash, k, _ptr(q), len(q), _ptr(counts))
    return counts, n


def search_matches(slices, n_filters, log2_len, num_hash, k, query, threshold):
    q = _bytes_arr(query)
    s = np.ascontiguousarray(slices, dtype=np.uint8)
    hf = np.zeros(n_filters, dtype=np.uint32)
    hm = np.zeros(n_filters, dtype=np.uint32)
    nq = C.c_uint32(0)
    nh = lib().kwo_search_matches(_ptr(s), n_filters, log2_len, num_hash, k, _ptr(q), len(q),
                                  C.c_float(threshold), _ptr(hf), _ptr(hm), C.byref(nq))
    return hf[:nh].copy(), hm[:nh].copy(), nq.value


# ---------------------------------------------------------------- compiled reference (oracle/_ref)
def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "ref_driver")) and os.path.exists(os.path.join(REF_DIR, "kwage"))


def ref_driver(*args, env=None, check=True):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run([os.path.join(REF_DIR, "ref_driver")] + [str(a) for a in args],
                       capture_output=True, text=True, env=e)
    if check and r.returncode != 0:
        raise RuntimeError("ref_driver %s failed: %s" % (args, r.stderr))
    return r.stdout


def ref_make_bloom(acc, reads_dir, bloom_dir, k, min_count, p, min_log2, max_log2, num_bp):
    out = ref_driver("make_bloom", acc, bloom_dir, k, min_count, p, min_log2, max_log2, num_bp,
                     env={"KWAGE_READS_DIR": reads_dir})
    return json.loads(out)


def ref_kwage(args, omp_threads=None):
    e = dict(os.environ)
    if omp_threads:
        e["OMP_NUM_THREADS"] = str(omp_threads)
    return subprocess.run([os.path.join(REF_DIR, "kwage")] + [str(a) for a in args],
                          capture_output=True, text=True, env=e)
