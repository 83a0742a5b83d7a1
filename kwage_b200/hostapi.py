"""ctypes binding of libkwage_host.so: the C++ host layer (kwage_b200/host/) that mirrors the
reference's make_bloom_filter / build_db / parameter search on top of libkwage_cuda.so."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libkwage_host.so")
BIN_DIR = os.path.join(PKG, "bin")
KWAGE_BIN = os.path.join(BIN_DIR, "kwage")
TOOLS_BIN = os.path.join(BIN_DIR, "kwage_tools")

STATUS_BLOOM_SUCCESS, STATUS_BLOOM_FAIL, STATUS_BLOOM_INVALID = 14, 15, 16

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libkwage_host.so is not built (run `python -m kwage_b200.host.build`)")
        L = C.CDLL(LIB_PATH)
        u64, u32, f32, i32, cp = C.c_uint64, C.c_uint32, C.c_float, C.c_int, C.c_char_p
        L.kwh_optimal_bloom_param.argtypes = [u32, u64, f32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
        L.kwh_approximate_max_kmers.restype = u64
        L.kwh_approximate_max_kmers.argtypes = [f32, u32, u32]
        L.kwh_counting_filter_log2_len.restype = u32
        L.kwh_counting_filter_log2_len.argtypes = [u64]
        L.kwh_str_to_accession.restype = u64
        L.kwh_str_to_accession.argtypes = [cp]
        L.kwh_accession_to_str.argtypes = [u64, cp, C.c_size_t]
        L.kwh_make_bloom_file.argtypes = [cp, cp, u64, cp, u32, u32, f32, u32, u32, i32, C.POINTER(u64), C.POINTER(u32),
                                          C.POINTER(u32), C.POINTER(u32), cp, C.c_size_t, C.POINTER(u64)]
        L.kwh_write_bloom_file.argtypes = [cp, cp, u32, u32, u32, C.c_void_p]
        L.kwh_build_db.argtypes = [cp, u32, u32, u32, cp, i32]
        L.kwh_merge_db.argtypes = [cp, cp, u64, i32, cp, C.c_size_t]
        L.kwh_merge_db.restype = C.c_long
        L.kwh_db_load_seconds.argtypes = [cp, i32, C.POINTER(u64)]
        L.kwh_db_load_seconds.restype = C.c_double
        L.kwh_pack_2na.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_void_p, u64]
        _lib = L
    return _lib


def optimal_bloom_param(kmer_len, num_kmer, p, min_log2, max_log2):
    """-> (log2_len, num_hash) or None where the reference throws (bloom.cpp:16,66)."""
    L, h = C.c_uint32(0), C.c_uint32(0)
    rc = lib().kwh_optimal_bloom_param(kmer_len, num_kmer, p, min_log2, max_log2, C.byref(L), C.byref(h))
    return None if rc != 0 else (L.value, h.value)


def approximate_max_kmers(p, min_log2, max_log2):
    return lib().kwh_approximate_max_kmers(p, min_log2, max_log2)


def counting_filter_log2_len(num_bp):
    return lib().kwh_counting_filter_log2_len(num_bp)


def str_to_accession(s):
    return lib().kwh_str_to_accession(s.encode())


def accession_to_str(a):
    buf = C.create_string_buffer(32)
    lib().kwh_accession_to_str(a, buf, 32)
    return buf.value.decode()


def parse_digest(path):
    """(fragments, bases, longest fragment, FNV-1a hash) of a reads / FASTA / FASTQ(.gz) file as the host layer streams it"""
    out = (C.c_uint64 * 4)()
    lib().kwh_parse_digest.argtypes = [C.c_char_p, C.POINTER(C.c_uint64)]
    if lib().kwh_parse_digest(path.encode(), out):
        raise RuntimeError("kwh_parse_digest: cannot open " + path)
    return tuple(int(x) for x in out)


def make_bloom_file(accession, reads_path, num_bp, bloom_dir, *, k=31, min_kmer_count=1, p=0.25, min_log2=18, max_log2=32, device=0):
    n, L, h, lc = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    err = C.create_string_buffer(512)
    prog = (C.c_uint64 * 3)()
    status = lib().kwh_make_bloom_file(accession.encode(), reads_path.encode(), num_bp, bloom_dir.encode(), k, min_kmer_count, p,
                                       min_log2, max_log2, device, C.byref(n), C.byref(L), C.byref(h), C.byref(lc), err, 512, prog)
    return dict(status=status, num_kmer=n.value, log2_len=L.value, num_hash=h.value, log2_count_len=lc.value, error=err.value.decode(),
                num_bp=prog[0], curr_read=prog[1], curr_fragment=prog[2])


def write_bloom_file(path, accession, k, log2_len, num_hash, bits):
    import numpy as np
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    return bool(lib().kwh_write_bloom_file(path.encode(), accession.encode(), k, log2_len, num_hash, b.ctypes.data_as(C.c_void_p)))


def build_db(filename, k, log2_len, num_hash, bloom_files, *, device=0):
    return bool(lib().kwh_build_db(filename.encode(), k, log2_len, num_hash, "\n".join(bloom_files).encode(), device))


def pack_2na(fragments):
    """The host layer's packer (stages.cpp::pack_2na) over a list of fragments appended one after the other:
    -> (2na bytes, not-a-base mask, any_bad)."""
    import numpy as np
    frs = [np.frombuffer(f if isinstance(f, (bytes, bytearray)) else bytes(f), dtype=np.uint8) for f in fragments]
    total = sum(len(f) for f in frs)
    packed = np.zeros(total // 4 + 8, dtype=np.uint8)
    mask = np.zeros(total // 8 + 8, dtype=np.uint8)
    cur, any_bad = 0, False
    for f in frs:
        if len(f) and cur % 4 == 0:
            packed[cur // 4] = 0
        f = np.ascontiguousarray(f)
        any_bad |= bool(lib().kwh_pack_2na(packed.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), cur,
                                           f.ctypes.data_as(C.c_void_p), len(f)))
        cur += len(f)
    return packed[: (total + 3) // 4], mask[: (total + 7) // 8], any_bad


def merge_db(file_1, file_2, max_num_filters=0, *, device=0):
    """merge_database_files() (reference merge_db.cpp:278): -> filters in the file that can still take more; raises on error"""
    err = C.create_string_buffer(512)
    r = lib().kwh_merge_db(file_1.encode(), file_2.encode(), max_num_filters, device, err, 512)
    if r < 0:
        raise RuntimeError(err.value.decode() or "merge_database_files failed")
    return r


def db_load_seconds(paths, *, device=0):
    """SubjectDatabase over one or several .db files: (wall-clock seconds of the load, bytes of slices)"""
    n = C.c_uint64(0)
    sec = lib().kwh_db_load_seconds("\n".join(paths).encode(), device, C.byref(n))
    if sec < 0:
        raise RuntimeError("SubjectDatabase load failed")
    return sec, n.value
