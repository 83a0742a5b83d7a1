"""ctypes binding of libkwage_cuda.so (include/kwage_cuda.h).

This module is a thin binding, not an implementation: every call goes to the CUDA library and
raises KwageError when the library reports a failure.  There is no CPU fallback -- if the shared
library has not been built, importing the binding's `lib()` raises.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libkwage_cuda.so")

KWG_OK = 0
KWG_ERR_INVALID_ARG = -1
KWG_ERR_CUDA = -2
KWG_ERR_NO_MEMORY = -3
KWG_ERR_UNSUPPORTED = -4
KWG_ERR_STATE = -5

# every symbol include/kwage_cuda.h declares
EXPORTS = [
    "kwg_last_error", "kwg_version", "kwg_device_count", "kwg_launch_count",
    "kwg_bloom_create", "kwg_bloom_create_raw", "kwg_bloom_add_reads", "kwg_bloom_add_reads_dev",
    "kwg_bloom_add_packed", "kwg_bloom_add_packed_dev",
    "kwg_bloom_num_valid", "kwg_bloom_finalize", "kwg_bloom_finalize_dev", "kwg_bloom_finalize_crc", "kwg_bloom_reset",
    "kwg_bloom_checkpoint", "kwg_bloom_rollback",
    "kwg_bloom_sync", "kwg_bloom_destroy", "kwg_bloom_stream",
    "kwg_transpose", "kwg_transpose_dev", "kwg_transpose_crc", "kwg_crc32_dev", "kwg_host_alloc", "kwg_host_free",
    "kwg_merge_slices", "kwg_release_caches", "kwg_db_upload_rows_async", "kwg_db_upload_columns_async",
    "kwg_db_load", "kwg_db_alloc", "kwg_db_upload_rows", "kwg_db_upload_columns", "kwg_db_attach_dev", "kwg_db_unload", "kwg_search", "kwg_search_ptrs",
    "kwg_search_counts", "kwg_search_counts_dev", "kwg_db_sync", "kwg_db_stream", "kwg_free_hits",
    "kwg_db_set_count_budget", "kwg_search_hits_dev",
    "kwg_comm_get_unique_id", "kwg_comm_create", "kwg_comm_create_all", "kwg_comm_destroy", "kwg_search_gather", "kwg_merge_hits",
    "kwg_synth_reads_dev", "kwg_synth_filter_bits_dev", "kwg_synth_plant_dev",
    "kwg_bloom_set_timing", "kwg_bloom_get_timing", "kwg_db_set_timing", "kwg_db_get_timing",
]
T_SCAN_A, T_SCAN_B, T_INSERT, T_AUX, T_SEARCH, T_HITS, T_REGROUP, T_RESOLVE, T_COUNT = 0, 1, 2, 3, 4, 5, 6, 7, 8


class KwageError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libkwage_cuda error %d: %s" % (code, msg))
        self.code = code


class Hit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("filter", C.c_uint32), ("num_match", C.c_uint32)]


HIT_DTYPE = np.dtype([("query", np.uint32), ("filter", np.uint32), ("num_match", np.uint32)])

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libkwage_cuda.so is not built (run `python -m kwage_b200.build`); "
                          "there is no CPU fallback for the KWAGE hot path")
    L = C.CDLL(LIB_PATH)
    u64, u32, vp, i32, f32 = C.c_uint64, C.c_uint32, C.c_void_p, C.c_int, C.c_float
    pvp = C.POINTER(C.c_void_p)
    L.kwg_last_error.restype = C.c_char_p
    L.kwg_version.restype = C.c_char_p
    L.kwg_device_count.argtypes = [C.POINTER(i32)]
    L.kwg_launch_count.restype = u64
    L.kwg_bloom_create.argtypes = [pvp, i32, u32, u32, u32, u32]
    L.kwg_bloom_create_raw.argtypes = [pvp, i32, u32, u32, u32]
    L.kwg_bloom_add_reads.argtypes = [vp, vp, vp, u64]
    L.kwg_bloom_add_reads_dev.argtypes = [vp, vp, vp, u64, u64]
    L.kwg_bloom_add_packed.argtypes = [vp, vp, vp, vp, u64]
    L.kwg_bloom_add_packed_dev.argtypes = [vp, vp, vp, vp, u64, u64]
    L.kwg_bloom_num_valid.argtypes = [vp, C.POINTER(u64)]
    L.kwg_bloom_finalize.argtypes = [vp, u32, u32, vp]
    L.kwg_bloom_finalize_dev.argtypes = [vp, u32, u32, vp]
    L.kwg_bloom_reset.argtypes = [vp]
    L.kwg_bloom_checkpoint.argtypes = [vp]
    L.kwg_bloom_rollback.argtypes = [vp]
    L.kwg_bloom_sync.argtypes = [vp]
    L.kwg_bloom_destroy.argtypes = [vp]
    L.kwg_bloom_destroy.restype = None
    L.kwg_bloom_stream.argtypes = [vp, pvp]
    L.kwg_transpose.argtypes = [i32, vp, u32, u64, vp]
    L.kwg_transpose_dev.argtypes = [i32, vp, u64, u32, u64, vp, u64, vp]
    L.kwg_transpose_crc.argtypes = [i32, vp, u32, u64, vp, vp, vp]
    L.kwg_crc32_dev.argtypes = [i32, vp, u64, u64, u64, u32, C.POINTER(u32), vp]
    L.kwg_host_alloc.argtypes = [u64]
    L.kwg_host_alloc.restype = vp
    L.kwg_host_free.argtypes = [vp]
    L.kwg_host_free.restype = None
    L.kwg_bloom_finalize_crc.argtypes = [vp, u32, u32, vp, C.POINTER(u32)]
    L.kwg_db_load.argtypes = [pvp, i32, vp, u32, u32, u32, u32, u32, u32]
    L.kwg_db_alloc.argtypes = [pvp, i32, u32, u32, u32, u32, u32, u32]
    L.kwg_db_upload_rows.argtypes = [vp, u64, u64, vp]
    L.kwg_db_upload_columns.argtypes = [vp, u32, u32, u64, u64, vp]
    L.kwg_db_upload_rows_async.argtypes = [vp, u64, u64, vp]
    L.kwg_db_upload_columns_async.argtypes = [vp, u32, u32, u64, u64, vp]
    L.kwg_release_caches.restype = None
    L.kwg_merge_slices.argtypes = [i32, vp, u32, vp, u32, u64, u32, vp, vp]
    L.kwg_db_attach_dev.argtypes = [pvp, i32, vp, u64, u32, u32, u32, u32]
    L.kwg_db_unload.argtypes = [vp]
    L.kwg_db_unload.restype = None
    L.kwg_search.argtypes = [vp, vp, vp, u32, f32, vp, C.POINTER(C.POINTER(Hit)), C.POINTER(u64)]
    L.kwg_search_ptrs.argtypes = [vp, vp, vp, u32, f32, vp, C.POINTER(C.POINTER(Hit)), C.POINTER(u64)]
    L.kwg_search_counts.argtypes = [vp, vp, vp, u32, vp, vp]
    L.kwg_search_counts_dev.argtypes = [vp, vp, vp, u32, u64, vp, vp, u64]
    L.kwg_db_sync.argtypes = [vp]
    L.kwg_db_set_count_budget.argtypes = [vp, u64]
    L.kwg_search_hits_dev.argtypes = [vp, vp, vp, u32, f32, vp, u32, pvp, C.POINTER(u64)]
    L.kwg_comm_get_unique_id.argtypes = [vp]
    L.kwg_comm_create.argtypes = [pvp, i32, i32, i32, vp]
    L.kwg_comm_create_all.argtypes = [pvp, i32, vp]
    L.kwg_comm_destroy.argtypes = [vp]
    L.kwg_comm_destroy.restype = None
    L.kwg_search_gather.argtypes = [vp, vp, i32, vp, vp, u32, f32, u32, vp, C.POINTER(C.POINTER(Hit)), C.POINTER(u64)]
    L.kwg_merge_hits.argtypes = [vp, vp, u32, u32, vp]
    L.kwg_merge_hits.restype = None
    L.kwg_db_stream.argtypes = [vp, pvp]
    L.kwg_free_hits.argtypes = [C.POINTER(Hit)]
    L.kwg_free_hits.restype = None
    L.kwg_bloom_set_timing.argtypes = [vp, i32]
    L.kwg_bloom_get_timing.argtypes = [vp, vp, vp]
    L.kwg_db_set_timing.argtypes = [vp, i32]
    L.kwg_db_get_timing.argtypes = [vp, vp, vp]
    L.kwg_synth_reads_dev.argtypes = [i32, u64, u64, u64, u32, vp, vp, vp]
    L.kwg_synth_filter_bits_dev.argtypes = [i32, u64, u64, u32, u64, u64, vp, vp]
    L.kwg_synth_plant_dev.argtypes = [i32, vp, u64, u32, u32, u32, u32, vp, vp, u32, u32, u32, u32, vp]
    _lib = L
    return L


def check(rc):
    if rc != KWG_OK:
        raise KwageError(rc, lib().kwg_last_error().decode(errors="replace"))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_bases(x):
    if isinstance(x, str):
        x = x.encode()
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    a = np.asarray(x)
    if a.dtype != np.uint8:
        a = a.astype(np.uint8)
    return np.ascontiguousarray(a)


def device_count():
    n = C.c_int(0)
    rc = lib().kwg_device_count(C.byref(n))
    return n.value if rc == KWG_OK else 0


def launch_count():
    return int(lib().kwg_launch_count())


def flatten(seqs):
    """list of str/bytes/uint8 arrays -> (uint8 bases, uint64 offsets)"""
    arrs = [_as_bases(s) for s in seqs]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([len(a) for a in arrs])
    bases = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(bases, dtype=np.uint8), offsets


def pack_2na(bases):
    """ASCII bases -> (NCBI 2na bytes, not-a-base mask or None): the input format of kwg_bloom_add_packed.  Test and
    bench helper (numpy); a host that feeds the library packs on its parser threads (kwage_b200/host/stages.cpp)."""
    b = _as_bases(bases)
    n = len(b)
    u = b & 0xDF
    code = np.zeros(n, dtype=np.uint8)
    code[u == ord("C")] = 1
    code[u == ord("G")] = 2
    code[u == ord("T")] = 3
    bad = ~((u == ord("A")) | (u == ord("C")) | (u == ord("G")) | (u == ord("T")))
    pad = (-n) % 4
    c4 = np.concatenate([code, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    packed = ((c4[:, 0] << 6) | (c4[:, 1] << 4) | (c4[:, 2] << 2) | c4[:, 3]).astype(np.uint8)
    mask = None
    if bad.any():
        mask = np.packbits(bad, bitorder="little")
        if len(mask) % 2:
            mask = np.concatenate([mask, np.zeros(1, dtype=np.uint8)])
    return np.ascontiguousarray(packed), mask


class BloomBuilder:
    """kwg_bloom_t: streaming construction of one Bloom filter (one accession)."""

    def __init__(self, kmer_len, *, device=0, min_kmer_count=1, log2_count_len=None, log2_max_len=32,
                 raw_num_hash=None, raw_log2_len=None):
        self.h = C.c_void_p()
        self.raw = raw_num_hash is not None
        if self.raw:
            check(lib().kwg_bloom_create_raw(C.byref(self.h), device, kmer_len, raw_num_hash, raw_log2_len))
            self.num_hash, self.log2_len = raw_num_hash, raw_log2_len
        else:
            check(lib().kwg_bloom_create(C.byref(self.h), device, kmer_len, min_kmer_count, log2_count_len, log2_max_len))

    def add_reads(self, bases, offsets):
        bases = _as_bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        check(lib().kwg_bloom_add_reads(self.h, _np_ptr(bases), _np_ptr(offsets), len(offsets) - 1))

    def add_packed(self, packed, bad_mask, offsets):
        """2na bytes + optional not-a-base mask (pack_2na) + offsets in bases"""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if bad_mask is not None:
            bad_mask = np.ascontiguousarray(bad_mask, dtype=np.uint8)
        check(lib().kwg_bloom_add_packed(self.h, _np_ptr(packed), _np_ptr(bad_mask) if bad_mask is not None else None,
                                         _np_ptr(offsets), len(offsets) - 1))

    def add_packed_ptr(self, packed_ptr, bad_ptr, offsets_ptr, n_reads):
        check(lib().kwg_bloom_add_packed(self.h, C.c_void_p(packed_ptr), C.c_void_p(bad_ptr) if bad_ptr else None, C.c_void_p(offsets_ptr), n_reads))

    def add_packed_dev(self, d_packed_ptr, d_bad_ptr, d_offsets_ptr, n_reads, n_bases):
        check(lib().kwg_bloom_add_packed_dev(self.h, C.c_void_p(d_packed_ptr), C.c_void_p(d_bad_ptr) if d_bad_ptr else None,
                                             C.c_void_p(d_offsets_ptr), n_reads, n_bases))

    def add_reads_ptr(self, bases_ptr, offsets_ptr, n_reads):
        """Host pointers (e.g. pinned torch tensors): no numpy conversion."""
        check(lib().kwg_bloom_add_reads(self.h, C.c_void_p(bases_ptr), C.c_void_p(offsets_ptr), n_reads))

    def add_reads_dev(self, d_bases_ptr, d_offsets_ptr, n_reads, n_bases):
        check(lib().kwg_bloom_add_reads_dev(self.h, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_reads, n_bases))

    def num_valid(self):
        n = C.c_uint64(0)
        check(lib().kwg_bloom_num_valid(self.h, C.byref(n)))
        return n.value

    def finalize(self, log2_len=None, num_hash=None, out=None):
        if self.raw:
            log2_len, num_hash = self.log2_len, self.num_hash
        if out is None:
            out = np.empty((1 << log2_len) // 8, dtype=np.uint8)
        check(lib().kwg_bloom_finalize(self.h, log2_len, num_hash, _np_ptr(out)))
        return out

    def finalize_crc(self, log2_len=None, num_hash=None):
        """-> (bits, zlib crc32 of the bits computed on the device)"""
        if self.raw:
            log2_len, num_hash = self.log2_len, self.num_hash
        out = np.empty((1 << log2_len) // 8, dtype=np.uint8)
        crc = C.c_uint32(0)
        check(lib().kwg_bloom_finalize_crc(self.h, log2_len, num_hash, _np_ptr(out), C.byref(crc)))
        return out, crc.value

    def finalize_crc_ptr(self, log2_len, num_hash, out_ptr):
        crc = C.c_uint32(0)
        check(lib().kwg_bloom_finalize_crc(self.h, log2_len, num_hash, C.c_void_p(out_ptr), C.byref(crc)))
        return crc.value

    def finalize_ptr(self, log2_len, num_hash, out_ptr):
        check(lib().kwg_bloom_finalize(self.h, log2_len, num_hash, C.c_void_p(out_ptr)))

    def finalize_dev(self, log2_len, num_hash, d_out_ptr):
        check(lib().kwg_bloom_finalize_dev(self.h, log2_len, num_hash, C.c_void_p(d_out_ptr)))

    def reset(self):
        check(lib().kwg_bloom_reset(self.h))

    def checkpoint(self):
        check(lib().kwg_bloom_checkpoint(self.h))

    def rollback(self):
        check(lib().kwg_bloom_rollback(self.h))

    def set_timing(self, enable=True):
        check(lib().kwg_bloom_set_timing(self.h, int(enable)))

    def get_timing(self):
        """-> (ms[T_COUNT], launches[T_COUNT]) accumulated since the last call; synchronises."""
        ms = np.zeros(T_COUNT, dtype=np.float64)
        n = np.zeros(T_COUNT, dtype=np.uint64)
        check(lib().kwg_bloom_get_timing(self.h, _np_ptr(ms), _np_ptr(n)))
        return ms, n

    def sync(self):
        check(lib().kwg_bloom_sync(self.h))

    def stream(self):
        s = C.c_void_p()
        check(lib().kwg_bloom_stream(self.h, C.byref(s)))
        return s.value or 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().kwg_bloom_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def transpose(filters, chunk_bits, *, device=0):
    """filters: list of uint8 arrays (>= chunk_bits/8 bytes each) -> (chunk_bits, ceil(n/8)) uint8
    through kwg_transpose (host buffers in, host buffer out)."""
    n = len(filters)
    keep = [np.ascontiguousarray(f, dtype=np.uint8) for f in filters]
    ptrs = (C.c_void_p * max(n, 1))(*[f.ctypes.data for f in keep])
    row = (n + 7) // 8
    dest = np.empty(chunk_bits * row, dtype=np.uint8)
    check(lib().kwg_transpose(device, ptrs, n, chunk_bits, _np_ptr(dest)))
    return dest.reshape(chunk_bits, row)


def transpose_crc(filters, chunk_bits, filter_crc=None, dest_crc=None, *, device=0):
    """kwg_transpose_crc: -> (slices, running per-filter crc32 array or None, running slice crc32 or None)"""
    n = len(filters)
    keep = [np.ascontiguousarray(f, dtype=np.uint8) for f in filters]
    ptrs = (C.c_void_p * max(n, 1))(*[f.ctypes.data for f in keep])
    row = (n + 7) // 8
    dest = np.empty(chunk_bits * row, dtype=np.uint8)
    fc = None if filter_crc is None else np.ascontiguousarray(filter_crc, dtype=np.uint32).copy()
    dc = None if dest_crc is None else C.c_uint32(dest_crc)
    check(lib().kwg_transpose_crc(device, ptrs, n, chunk_bits, _np_ptr(dest), None if fc is None else _np_ptr(fc),
                                  None if dc is None else C.byref(dc)))
    return dest.reshape(chunk_bits, row), fc, (None if dc is None else dc.value)


def crc32_dev(d_ptr, n_rows, row_bytes, row_pitch, crc_in=0, *, device=0, stream=0):
    out = C.c_uint32(0)
    check(lib().kwg_crc32_dev(device, C.c_void_p(d_ptr), n_rows, row_bytes, row_pitch, crc_in, C.byref(out), C.c_void_p(stream)))
    return out.value


def transpose_dev(d_filters_ptr, filter_pitch, n_filters, chunk_bits, d_dest_ptr, dest_pitch, *, device=0, stream=0):
    check(lib().kwg_transpose_dev(device, C.c_void_p(d_filters_ptr), filter_pitch, n_filters, chunk_bits,
                                  C.c_void_p(d_dest_ptr), dest_pitch, C.c_void_p(stream)))


class Database:
    """kwg_db_t: slice region of one database (or one column slab of it) resident in HBM."""

    def __init__(self, h, n_filters, col_begin=0):
        self.h = h
        self.n_filters = n_filters
        self.col_begin = col_begin

    @classmethod
    def load(cls, slices, kmer_len, num_hash, log2_len, n_filters_total, *, device=0, col_begin=0, col_end=None):
        s = np.ascontiguousarray(slices, dtype=np.uint8)
        if col_end is None:
            col_end = n_filters_total
        h = C.c_void_p()
        check(lib().kwg_db_load(C.byref(h), device, _np_ptr(s), kmer_len, num_hash, log2_len, n_filters_total, col_begin, col_end))
        return cls(h, col_end - col_begin, col_begin)

    @classmethod
    def alloc(cls, kmer_len, num_hash, log2_len, n_filters_total, *, device=0, col_begin=0, col_end=None):
        if col_end is None:
            col_end = n_filters_total
        h = C.c_void_p()
        check(lib().kwg_db_alloc(C.byref(h), device, kmer_len, num_hash, log2_len, n_filters_total, col_begin, col_end))
        return cls(h, col_end - col_begin, col_begin)

    @classmethod
    def alloc(cls, kmer_len, num_hash, log2_len, n_filters, *, device=0):
        h = C.c_void_p()
        check(lib().kwg_db_alloc(C.byref(h), device, kmer_len, num_hash, log2_len, n_filters, 0, n_filters))
        return cls(h, n_filters)

    def upload_columns(self, col_begin, n_cols, row_begin, rows):
        """rows: (n_rows, ceil(n_cols/8)) uint8 = a file's slices -> columns [col_begin, col_begin + n_cols) of the slab"""
        r = np.ascontiguousarray(rows, dtype=np.uint8)
        assert r.shape[1] == (n_cols + 7) // 8
        check(lib().kwg_db_upload_columns(self.h, col_begin, n_cols, row_begin, r.shape[0], _np_ptr(r)))

    def upload_rows(self, row_begin, rows):
        r = np.ascontiguousarray(rows, dtype=np.uint8)
        check(lib().kwg_db_upload_rows(self.h, row_begin, r.shape[0], _np_ptr(r)))

    @classmethod
    def attach_dev(cls, d_slices_ptr, row_pitch, kmer_len, num_hash, log2_len, n_filters, *, device=0):
        h = C.c_void_p()
        check(lib().kwg_db_attach_dev(C.byref(h), device, C.c_void_p(d_slices_ptr), row_pitch, kmer_len, num_hash, log2_len, n_filters))
        return cls(h, n_filters)

    def search(self, queries, threshold):
        """-> (hits structured array ordered by (query, filter), n_query_kmers)"""
        bases, offsets = flatten(queries)
        return self.search_flat(bases, offsets, threshold)

    def search_flat(self, bases, offsets, threshold):
        bases = _as_bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nq = len(offsets) - 1
        nk = np.zeros(max(nq, 1), dtype=np.uint32)
        hp = C.POINTER(Hit)()
        nh = C.c_uint64(0)
        check(lib().kwg_search(self.h, _np_ptr(bases), _np_ptr(offsets), nq, C.c_float(threshold), _np_ptr(nk), C.byref(hp), C.byref(nh)))
        try:
            if nh.value:
                buf = C.string_at(hp, nh.value * C.sizeof(Hit))
                hits = np.frombuffer(buf, dtype=HIT_DTYPE).copy()
            else:
                hits = np.zeros(0, dtype=HIT_DTYPE)
        finally:
            if nh.value:
                lib().kwg_free_hits(hp)
        return hits, nk[:nq]

    def set_count_budget(self, n_bytes):
        check(lib().kwg_db_set_count_budget(self.h, int(n_bytes)))

    def search_hits_dev(self, bases, offsets, threshold, filter0=0):
        """kwg_search_hits_dev: -> (device pointer of the hit list, number of hits, n_query_kmers)"""
        bases = _as_bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nq = len(offsets) - 1
        nk = np.zeros(max(nq, 1), dtype=np.uint32)
        dp = C.c_void_p()
        nh = C.c_uint64(0)
        check(lib().kwg_search_hits_dev(self.h, _np_ptr(bases), _np_ptr(offsets), nq, C.c_float(threshold), _np_ptr(nk), int(filter0),
                                        C.byref(dp), C.byref(nh)))
        return dp.value or 0, nh.value, nk[:nq]

    def search_gather(self, comm, bases, offsets, threshold, filter0, root=0):
        """kwg_search_gather (collective over `comm`): merged hits with global filter indices on the root, an empty
        array elsewhere; plus n_query_kmers."""
        bases = _as_bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nq = len(offsets) - 1
        nk = np.zeros(max(nq, 1), dtype=np.uint32)
        hp = C.POINTER(Hit)()
        nh = C.c_uint64(0)
        check(lib().kwg_search_gather(self.h, comm.h, int(root), _np_ptr(bases), _np_ptr(offsets), nq, C.c_float(threshold), int(filter0),
                                      _np_ptr(nk), C.byref(hp), C.byref(nh)))
        try:
            hits = np.frombuffer(C.string_at(hp, nh.value * C.sizeof(Hit)), dtype=HIT_DTYPE).copy() if nh.value else np.zeros(0, dtype=HIT_DTYPE)
        finally:
            if nh.value:
                lib().kwg_free_hits(hp)
        return hits, nk[:nq]

    def search_gather_ptr(self, comm, bases_ptr, offsets_ptr, n_queries, threshold, filter0, nk_ptr, root=0):
        """Same with raw host pointers (benchmarks); returns the number of hits delivered to this rank."""
        hp = C.POINTER(Hit)()
        nh = C.c_uint64(0)
        check(lib().kwg_search_gather(self.h, comm.h, int(root), bases_ptr, offsets_ptr, int(n_queries), C.c_float(threshold), int(filter0),
                                      nk_ptr, C.byref(hp), C.byref(nh)))
        n = nh.value
        if n:
            lib().kwg_free_hits(hp)
        return n

    def search_ptrs(self, queries, threshold):
        """Same through kwg_search_ptrs (one pointer per query)."""
        arrs = [_as_bases(q) for q in queries]
        nq = len(arrs)
        ptrs = (C.c_void_p * max(nq, 1))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], dtype=np.uint64)
        nk = np.zeros(max(nq, 1), dtype=np.uint32)
        hp = C.POINTER(Hit)()
        nh = C.c_uint64(0)
        check(lib().kwg_search_ptrs(self.h, ptrs, _np_ptr(lens), nq, C.c_float(threshold), _np_ptr(nk), C.byref(hp), C.byref(nh)))
        hits = np.zeros(0, dtype=HIT_DTYPE)
        if nh.value:
            hits = np.frombuffer(C.string_at(hp, nh.value * C.sizeof(Hit)), dtype=HIT_DTYPE).copy()
            lib().kwg_free_hits(hp)
        return hits, nk[:nq]

    def search_counts(self, queries):
        bases, offsets = flatten(queries)
        return self.search_counts_flat(bases, offsets)

    def search_counts_flat(self, bases, offsets):
        bases = _as_bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nq = len(offsets) - 1
        nk = np.zeros(max(nq, 1), dtype=np.uint32)
        counts = np.zeros((max(nq, 1), self.n_filters), dtype=np.uint32)
        check(lib().kwg_search_counts(self.h, _np_ptr(bases), _np_ptr(offsets), nq, _np_ptr(nk), _np_ptr(counts)))
        return counts[:nq], nk[:nq]

    def search_counts_dev(self, d_bases_ptr, d_offsets_ptr, n_queries, n_bases, d_nk_ptr, d_counts_ptr, count_pitch):
        check(lib().kwg_search_counts_dev(self.h, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_queries, n_bases,
                                          C.c_void_p(d_nk_ptr), C.c_void_p(d_counts_ptr), count_pitch))

    def sync(self):
        check(lib().kwg_db_sync(self.h))

    def set_timing(self, enable=True):
        check(lib().kwg_db_set_timing(self.h, int(enable)))

    def get_timing(self):
        ms = np.zeros(T_COUNT, dtype=np.float64)
        n = np.zeros(T_COUNT, dtype=np.uint64)
        check(lib().kwg_db_get_timing(self.h, _np_ptr(ms), _np_ptr(n)))
        return ms, n

    def stream(self):
        s = C.c_void_p()
        check(lib().kwg_db_stream(self.h, C.byref(s)))
        return s.value or 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().kwg_db_unload(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """kwg_comm_t: the NCCL communicator behind kwg_search_gather (one per device)."""

    def __init__(self, h, rank, n_ranks):
        self.h, self.rank, self.n_ranks = h, rank, n_ranks

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib().kwg_comm_get_unique_id(buf))
        return bytes(buf)

    @classmethod
    def create(cls, device, n_ranks, rank, uid):
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(lib().kwg_comm_create(C.byref(h), device, n_ranks, rank, buf))
        return cls(h, rank, n_ranks)

    @classmethod
    def create_all(cls, devices):
        n = len(devices)
        hs = (C.c_void_p * n)()
        devs = (C.c_int * n)(*devices)
        check(lib().kwg_comm_create_all(hs, n, devs))
        return [cls(C.c_void_p(hs[i]), i, n) for i in range(n)]

    def close(self):
        if self.h:
            lib().kwg_comm_destroy(self.h)
            self.h = None


def merge_hits(lists, n_queries):
    """kwg_merge_hits: per-slab hit arrays (each ordered by (query, filter), slabs in column order, global filter
    indices) -> one array ordered by (query, filter)."""
    lens = np.array([len(x) for x in lists], dtype=np.uint64)
    flat = np.concatenate([np.asarray(x, dtype=HIT_DTYPE) for x in lists]) if len(lists) else np.zeros(0, dtype=HIT_DTYPE)
    out = np.zeros(len(flat), dtype=HIT_DTYPE)
    if len(flat):
        lib().kwg_merge_hits(_np_ptr(flat), _np_ptr(lens), len(lists), int(n_queries), _np_ptr(out))
    return out


def synth_reads_dev(seed, first_read, n_reads, read_len, d_bases_ptr, d_offsets_ptr=0, *, device=0, stream=0):
    check(lib().kwg_synth_reads_dev(device, seed, first_read, n_reads, read_len, C.c_void_p(d_bases_ptr),
                                    C.c_void_p(d_offsets_ptr), C.c_void_p(stream)))


def synth_plant_dev(d_slab_ptr, row_pitch, k, num_hash, log2_len, column, d_bases_ptr, d_offsets_ptr, query_first, query_stride,
                    n_planted, plant_len, *, device=0, stream=0):
    check(lib().kwg_synth_plant_dev(device, C.c_void_p(d_slab_ptr), row_pitch, k, num_hash, log2_len, column, C.c_void_p(d_bases_ptr),
                                    C.c_void_p(d_offsets_ptr), query_first, query_stride, n_planted, plant_len, C.c_void_p(stream)))


def synth_filter_bits_dev(seed, first_filter, n_filters, filter_bytes, filter_pitch, d_filters_ptr, *, device=0, stream=0):
    check(lib().kwg_synth_filter_bits_dev(device, seed, first_filter, n_filters, filter_bytes, filter_pitch,
                                          C.c_void_p(d_filters_ptr), C.c_void_p(stream)))
