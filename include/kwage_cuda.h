/*
 * kwage_cuda.h -- C ABI of libkwage_cuda.so, the B200 (sm_100a) implementation of KWAGE's k-mer
 * Bloom-filter hot path.  This is the drop-in boundary: plain pointers and sizes, no C++ or torch
 * types.  The reference has no FFI/plugin interface; each group of entry points
 * below replaces the body of one reference function, cited as file:line into the reference tree.
 *
 * Conventions
 *   - every function returns 0 (KWG_OK) or a negative kwg_status; kwg_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread.  Nothing
 *     ever throws across this boundary; the reference-side shim turns non-zero into
 *     `throw __FILE__ ":...";` so the reference's existing catch blocks keep working
 *     (make_bloom.cpp:454-501, build_db.cpp:435-453, kwage.cpp:180-187).
 *   - the library owns all device memory; the caller owns every host buffer.  Host input buffers
 *     may be reused as soon as the call returns.
 *   - one handle = one CUDA stream: handles may be used from different host threads concurrently
 *     (kwage.cpp:76-87 enters search() from an OpenMP region); a single handle must not be used by
 *     two threads at once.  The handle-less calls (kwg_transpose*, kwg_merge_slices) may be called
 *     from any thread; calls for one device take turns, calls for different devices run side by side.
 *   - there is NO CPU fallback: with no usable CUDA device every call fails with KWG_ERR_CUDA.
 *   - bit order everywhere is the reference's BitVector order (bloom.h:131-163): bit i of a
 *     vector lives in byte i/8 at bit position i%8.
 *   - functions suffixed _dev take DEVICE pointers (already resident in HBM) and run
 *     asynchronously on the handle's stream unless stated otherwise.
 */
#ifndef KWAGE_CUDA_H
#define KWAGE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KWG_MAX_KMER_LEN 32   /* reference word.h:10 */
#define KWG_MAX_KMER_LEN_RAW 63   /* raw mode only: an extension beyond the reference (BASELINE.json configs[4]); parity unpinned above 32 */
#define KWG_MAX_NUM_HASH 8    /* reference hash.cpp:7 (construction uses <= 5: bloom.h:21) */
#define KWG_COUNT_NUM_HASH 5  /* hashes evaluated per k-mer in counting mode: bloom.h:21 */

typedef enum {
	KWG_OK = 0,
	KWG_ERR_INVALID_ARG = -1,   /* argument outside the reference's limits */
	KWG_ERR_CUDA = -2,          /* CUDA runtime error / no device */
	KWG_ERR_NO_MEMORY = -3,     /* device or pinned-host allocation failed */
	KWG_ERR_UNSUPPORTED = -4,   /* valid for the reference but not implemented on the device yet */
	KWG_ERR_STATE = -5          /* call sequence error (e.g. add_reads after finalize) */
} kwg_status;

const char* kwg_last_error(void);
const char* kwg_version(void);
int kwg_device_count(int* count);
/* Number of kernels this library has launched from the calling process so far (monotonic). */
uint64_t kwg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Bloom construction.  Replaces, inside make_bloom_filter() (make_bloom.cpp:76): the table
 * allocation 151-166, every count_words() call 201-203/239-241/281-283 (body 506-621), the
 * num_kmer bookkeeping that feeds the max_num_kmer checks 208/246/288, and the fold 337-354.
 * NGS iteration, number_of_bases(), optimal_bloom_param(), crc32 and file writing stay on host.
 *
 * Counting mode (kwg_bloom_create) reproduces the reference's pair of 4-bit counting Bloom
 * filters with conservative update (make_bloom.cpp:63-69,546-601) followed by the 5-vector
 * fold, bit-exactly, for every min_kmer_count in [1,15] (the reference's default is 5), including
 * the double increment when both hashes of one table meet on a slot.  One case is refused instead
 * of reproduced: with min_kmer_count == 15 such a double increment can take a 4-bit counter from
 * 14 to 16 = 0 in the reference; the library detects it and kwg_bloom_num_valid / kwg_bloom_finalize
 * return KWG_ERR_UNSUPPORTED.
 * Device memory per handle (besides ~64 bytes per start position of the largest batch): min_kmer_count 1 keeps one bit per
 * counting-filter slot (2^log2_count_len / 4 bytes); min_kmer_count > 1 keeps the reference's 4-bit counters
 * (2^log2_count_len bytes) plus a dense copy of the current batch's touch records for the later counter levels
 * (72 KiB per 2^15 slots: 4.5 GiB at log2_count_len 30, 18 GiB at 32; skipped if it cannot be allocated).
 *   log2_count_len : log2 of the counting-filter length, [18,32] (make_bloom.cpp:104-129)
 *   log2_max_len   : opt.max_log_2_filter_len, <= 32 (make_bloom.cpp:137-140)
 *
 * Raw mode (kwg_bloom_create_raw) sets bit (murmur3(kmer, seed=h) & (2^log2_len - 1)) for
 * h < num_hash for every valid canonical k-mer: the ground-truth construction of the reference's
 * own rig (bloom_test.cpp:268-275).  Raw mode also takes kmer_len 33..63 (128-bit words, the same rules); the reference
 * has no such k-mers (word.h:10), so there is nothing to be bit-exact with: checked against oracle/kwo_raw_insert_wide.
 * ------------------------------------------------------------------------------------------ */
typedef struct kwg_bloom kwg_bloom_t;

int kwg_bloom_create(kwg_bloom_t** out, int device, uint32_t kmer_len, uint32_t min_kmer_count,
	uint32_t log2_count_len, uint32_t log2_max_len);
int kwg_bloom_create_raw(kwg_bloom_t** out, int device, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len);

/* Reads/fragments as ASCII, concatenated; read r is bases[offsets[r] .. offsets[r+1]).  n_reads+1
 * offsets.  Stream order = array order (it matters in counting mode).  Any byte other than
 * ACGTacgt breaks k-mers exactly like word.h:98-100. */
int kwg_bloom_add_reads(kwg_bloom_t* b, const char* bases, const uint64_t* offsets, uint64_t n_reads);
/* Same, inputs already in HBM.  offsets[0] must be 0 and offsets[n_reads] == n_bases. */
int kwg_bloom_add_reads_dev(kwg_bloom_t* b, const char* d_bases, const uint64_t* d_offsets,
	uint64_t n_reads, uint64_t n_bases);

/* The same reads 2-bit packed, as the SRA stores them (NCBI 2na): four bases per byte, the first in bits 7..6, A=0 C=1
 * G=2 T=3 (the reference's own code order, word.h:19); base i of the call is in byte i/4.  bad_mask (may be NULL: every
 * base is one of ACGT) has bit i%8 of byte i/8 set where base i is anything else (N, IUPAC codes: word.h:98-100 breaks
 * the k-mer there); 2-byte aligned.  offsets are in bases, as above: reads need not start on byte boundaries.  A
 * quarter of the PCIe traffic of the ASCII call (plus 1/8 byte per base when there is a mask) for hosts that hold packed
 * reads or pack them on their parser threads (make_bloom.cpp:194-300 is the loop this feeds). */
int kwg_bloom_add_packed(kwg_bloom_t* b, const uint8_t* packed, const uint8_t* bad_mask, const uint64_t* offsets, uint64_t n_reads);
int kwg_bloom_add_packed_dev(kwg_bloom_t* b, const uint8_t* d_packed, const uint8_t* d_bad_mask, const uint64_t* d_offsets,
	uint64_t n_reads, uint64_t n_bases);

/* Counting mode: number of "valid" k-mers so far == BloomProgress::num_kmer (make_bloom.cpp:563).
 * Raw mode: number of k-mer occurrences inserted.  Synchronises the handle's stream. */
int kwg_bloom_num_valid(kwg_bloom_t* b, uint64_t* n);

/* Counting mode: build the final filter for the parameters the host chose with
 * optimal_bloom_param(); identical to folding valid_bits[h < num_hash] (make_bloom.cpp:337-354).
 * (min_kmer_count 1, log2_len <= log2_count_len: the bits of the seed pairs (0,1) and (2,3) are taken from the
 * counting tables themselves -- a slot is non-zero iff a valid k-mer put one of its two table hashes there,
 * make_bloom.cpp:546-601 -- and only a last odd seed is set k-mer by k-mer; KWG_NO_FOLD=1 in the environment when the
 * handle is created: every seed is.)
 * Raw mode: log2_len/num_hash must equal the creation values.  out_bits: 2^log2_len/8 bytes.
 * May be called more than once (e.g. with different parameters). */
int kwg_bloom_finalize(kwg_bloom_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* out_bits);
int kwg_bloom_finalize_dev(kwg_bloom_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* d_out_bits);
/* kwg_bloom_finalize that also returns BitVector::crc32 of the filter bits (bloom.cpp:328-336: zlib crc32 seeded with 0),
 * computed on the device while the bits travel to the host; replaces filter.update_crc32() (make_bloom.cpp:395). */
int kwg_bloom_finalize_crc(kwg_bloom_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* out_bits, uint32_t* crc32);

/* Forget everything added so far; keeps the allocations for the next accession. */
int kwg_bloom_reset(kwg_bloom_t* b);

/* Counting mode.  The reference tests `max_num_kmer < num_kmer` after EVERY fragment and stops there
 * (make_bloom.cpp:208-214, 246-252, 288-294); a caller that adds fragments in batches keeps that behaviour with
 * these two: kwg_bloom_checkpoint before a batch that may cross the limit (num_kmer so far + k-mer positions of the
 * batch > max_num_kmer), and, if kwg_bloom_num_valid then exceeds the limit, kwg_bloom_rollback followed by shorter
 * prefixes of the batch until the first fragment that crosses it is known: num_kmer, num_bp and the fragment counters
 * of the progress record are then the reference's.  The checkpoint is the counting-filter state (the two tables:
 * 2^(lc+1) bits for min_kmer_count 1, 2^lc bytes otherwise) plus the counters; one checkpoint per handle, a later
 * one replaces it, kwg_bloom_rollback may be called any number of times. */
int kwg_bloom_checkpoint(kwg_bloom_t* b);
int kwg_bloom_rollback(kwg_bloom_t* b);
int kwg_bloom_sync(kwg_bloom_t* b);
void kwg_bloom_destroy(kwg_bloom_t* b);

/* ------------------------------------------------------------------------------------------
 * Transposition.  Replaces the body of build_db()'s chunk loop, build_db.cpp:267-304 (memset of
 * dest + the bit-by-bit scatter 288-303).  File validation, per-filter crc32_z (281-282), the
 * output crc32_z (307) and all file I/O stay on host.
 *   filter_chunks : n_filters host pointers, each to chunk_bits/8 bytes of one filter
 *   chunk_bits    : number of slices in this chunk, a multiple of 8 (build_db.cpp:238-243)
 *   dest          : chunk_bits * ceil(n_filters/8) bytes, FULLY overwritten, padding bits zero
 * ------------------------------------------------------------------------------------------ */
int kwg_transpose(int device, const uint8_t* const* filter_chunks, uint32_t n_filters,
	uint64_t chunk_bits, uint8_t* dest);
/* Device-resident variant: filter j starts at d_filters + j*filter_pitch (pitch % 16 == 0), slice
 * k is written at d_dest + k*dest_pitch (dest_pitch % 16 == 0, >= ceil(n_filters/8)); bytes of a
 * slice beyond ceil(n_filters/8) up to dest_pitch are zero-filled.  chunk_bits % 32 == 0.
 * Runs on `stream` (a cudaStream_t, may be NULL) and does not synchronise. */
int kwg_transpose_dev(int device, const uint8_t* d_filters, uint64_t filter_pitch, uint32_t n_filters,
	uint64_t chunk_bits, uint8_t* d_dest, uint64_t dest_pitch, void* stream);
/* kwg_transpose that also advances the two kinds of running zlib crc32 values build_db keeps, on the device:
 *   filter_crc[j] : crc32_z(filter_crc[j], chunk of filter j)   -- build_db.cpp:281-282 (checked against the .bloom header, 321-333)
 *   *dest_crc     : crc32_z(*dest_crc, dest, chunk_bits * ceil(n_filters/8))   -- build_db.cpp:307 (DBFileHeader::crc32)
 * Either pointer may be NULL.  Needs chunk_bits % 32 == 0 and, for dest_crc, n_filters % 32 == 0 (messages made of
 * 32-bit words); otherwise KWG_ERR_INVALID_ARG and the caller keeps the host crc32. */
int kwg_transpose_crc(int device, const uint8_t* const* filter_chunks, uint32_t n_filters,
	uint64_t chunk_bits, uint8_t* dest, uint32_t* filter_crc, uint32_t* dest_crc);

/* ------------------------------------------------------------------------------------------
 * Column concatenation of database files: the body of merge_database_files()'s chunk loop, merge_db.cpp:489-584 (the
 * copy of source 1's slice and the bit-by-bit move of source 2's behind it, 533-566).  File handling, the running
 * crc32_z values (519-520, 576, 586) and the FilterInfo sections stay on host.
 *   src1, src2 : n_slices slices of ceil(n1/8) resp. ceil(n2/8) bytes (the files' slice layout, kwage.h:30-72)
 *   n_dst1     : filters of destination 1 = n1 + the leading (n_dst1 - n1) columns of source 2; the other columns of
 *                source 2 (has_remainder, merge_db.cpp:330) go to dst2, which may be NULL when there are none
 *   dst1, dst2 : n_slices slices of ceil(n_dst1/8) resp. ceil((n1 + n2 - n_dst1)/8) bytes, fully overwritten
 * ------------------------------------------------------------------------------------------ */
int kwg_merge_slices(int device, const uint8_t* src1, uint32_t n1, const uint8_t* src2, uint32_t n2, uint64_t n_slices,
	uint32_t n_dst1, uint8_t* dst1, uint8_t* dst2);

/* kwg_transpose[_crc] keep their streams and staging buffers (up to ~0.5 GB per device) between calls; this releases them. */
void kwg_release_caches(void);

/* Page-locked host memory for the staging buffers of the calls above (filter chunks, slices, read batches): the
 * library's copies then run at PCIe rate and overlap with its kernels.  Optional: any host pointer works. */
void* kwg_host_alloc(uint64_t bytes);
void kwg_host_free(void* p);

/* zlib crc32(crc_in, message) of a message resident in HBM: n_rows rows of row_bytes bytes, row r at d_data + r*row_pitch
 * (row_pitch == row_bytes or n_rows == 1: a flat buffer).  row_bytes, row_pitch and d_data multiples of 4.
 * Runs on `stream`, synchronises it, writes the host word *crc_out. */
int kwg_crc32_dev(int device, const uint8_t* d_data, uint64_t n_rows, uint64_t row_bytes, uint64_t row_pitch,
	uint32_t crc_in, uint32_t* crc_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Search.  Replaces, inside search() (kwage.cpp:340): k-mer extraction 352-366, the per-k-mer
 * slice loops 404-483 and the match decision 489-503.  Option parsing, FASTA iteration,
 * FilterInfo lookup 505-515, MatchResult sorting and CSV/JSON output stay on host.
 *
 * A kwg_db_t is the slice region of one database file (or one column slab of it) resident in
 * HBM: 2^log2_len slices of ceil(n_filters/8) bytes (kwage.h:30-72, build_db.cpp:259-314).
 * ------------------------------------------------------------------------------------------ */
typedef struct kwg_db kwg_db_t;

typedef struct {
	uint32_t query;      /* index into the queries of the call */
	uint32_t filter;     /* column index within this kwg_db_t (add the slab's first column) */
	uint32_t num_match;  /* MatchResult::num_kmers_found (kwage.cpp:519-520) */
} kwg_hit_t;

/* slices: host pointer to the file's slice region, 2^log2_len rows of ceil(n_filters_total/8)
 * bytes.  Only columns [col_begin, col_end) are kept on this device (col_begin % 8 == 0); pass
 * 0, n_filters_total for the whole file. */
int kwg_db_load(kwg_db_t** out, int device, const uint8_t* slices, uint32_t kmer_len, uint32_t num_hash,
	uint32_t log2_len, uint32_t n_filters_total, uint32_t col_begin, uint32_t col_end);
/* Streaming variant for slice regions larger than host memory: allocate the (zeroed) slab for
 * columns [col_begin, col_end) of a file with n_filters_total columns, then upload row ranges in
 * any order.  rows: n_rows consecutive slices of the FILE layout (ceil(n_filters_total/8) bytes
 * each) starting at slice row_begin. */
int kwg_db_alloc(kwg_db_t** out, int device, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len,
	uint32_t n_filters_total, uint32_t col_begin, uint32_t col_end);
int kwg_db_upload_rows(kwg_db_t* db, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows);
/* The same without waiting for the copy: `rows` must stay untouched until kwg_db_sync(db) (or the next synchronous call on
 * this handle) returns.  With two page-locked buffers (kwg_host_alloc) the host reads piece n + 1 of the file while piece n
 * travels: what SubjectDatabase does for the slice regions the reference seeks through on every query (kwage.cpp:404-483). */
int kwg_db_upload_rows_async(kwg_db_t* db, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows);
/* Several database files as one column slab (the reference keeps <= 2048 filters per file: 256-byte rows; wide rows gather
 * at three times the HBM rate).  After kwg_db_alloc(n_filters_total = sum of the files' filters, 0, n_filters_total):
 * rows [row_begin, row_begin + n_rows) of a file with n_cols filters -- n_rows * ceil(n_cols/8) host bytes, exactly the
 * file's slice region -- go to the columns [col_begin, col_begin + n_cols) of the slab; any bit offset.  Each column range
 * may be written once (bits are OR-ed into the zero-initialised slab).  Hits then carry slab column indices. */
int kwg_db_upload_columns(kwg_db_t* db, uint32_t col_begin, uint32_t n_cols, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows);
int kwg_db_upload_columns_async(kwg_db_t* db, uint32_t col_begin, uint32_t n_cols, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows);

/* Use slices that are already in HBM (e.g. written by kwg_transpose_dev); row k at
 * d_slices + k*row_pitch, row_pitch % 16 == 0, bits >= n_filters in a row must be zero.
 * The memory is borrowed, not owned. */
int kwg_db_attach_dev(kwg_db_t** out, int device, const uint8_t* d_slices, uint64_t row_pitch,
	uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len, uint32_t n_filters);
void kwg_db_unload(kwg_db_t* db);

/* queries: concatenated ASCII, query q is bases[offsets[q] .. offsets[q+1]).
 * n_query_kmers[q] (may be NULL) receives the number of unique canonical k-mers of query q
 * (kwage.cpp:368).  threshold in (0,1] with the reference's float arithmetic (kwage.cpp:349,388):
 * 1.0f -> every k-mer must match; otherwise num_match >= (unsigned)(threshold * n).
 * *hits is allocated by the library (release with kwg_free_hits), ordered by (query, filter).
 * Like the reference (kwage.cpp:397,459-482: "even the best matching Bloom filter does not have enough matches")
 * the thresholded calls -- kwg_search, kwg_search_ptrs, kwg_search_hits_dev, kwg_search_gather -- stop reading the
 * slices of a query once no filter can reach the threshold any more (decided per chunk of 4096 filter columns); the
 * hit list is the same either way and every hit carries its full num_match.  kwg_search_counts* always read every
 * slice.  KWG_SEARCH_NO_EXIT=1 in the environment when the handle is created: the thresholded calls do too. */
int kwg_search(kwg_db_t* db, const char* bases, const uint64_t* offsets, uint32_t n_queries, float threshold,
	uint32_t* n_query_kmers, kwg_hit_t** hits, uint64_t* n_hits);
/* Same with one pointer per query (the shape search() is called with, kwage.cpp:119,137). */
int kwg_search_ptrs(kwg_db_t* db, const char* const* queries, const uint64_t* query_len, uint32_t n_queries,
	float threshold, uint32_t* n_query_kmers, kwg_hit_t** hits, uint64_t* n_hits);
/* Raw per-filter counts (parity tests, multi-GPU gather): counts[q*n_filters + f]. */
int kwg_search_counts(kwg_db_t* db, const char* bases, const uint64_t* offsets, uint32_t n_queries,
	uint32_t* n_query_kmers, uint32_t* counts);
/* Device-resident variant: d_counts is n_queries * count_pitch uint32 (count_pitch >= n_filters,
 * multiple of 4).  d_n_query_kmers: n_queries uint32.  d_bases: 4-byte aligned, read in whole 32-bit words (up to the
 * word that holds base n_bases - 1).  Asynchronous on the db's stream. */
int kwg_search_counts_dev(kwg_db_t* db, const char* d_bases, const uint64_t* d_offsets, uint32_t n_queries,
	uint64_t n_bases, uint32_t* d_n_query_kmers, uint32_t* d_counts, uint64_t count_pitch);
int kwg_db_sync(kwg_db_t* db);
void kwg_free_hits(kwg_hit_t* hits);
/* kwg_search works through the queries in batches so that its per-(query, filter) counts stay below this many bytes
 * (default 1 GiB); the reference streams one query at a time (kwage.cpp:116-148). */
int kwg_db_set_count_budget(kwg_db_t* db, uint64_t bytes);
/* kwg_search that leaves the hit list in HBM (owned by the handle, valid until its next search): *d_hits is a DEVICE pointer.
 * filter0 is added to every filter index (the slab's first column in the whole database). */
int kwg_search_hits_dev(kwg_db_t* db, const char* bases, const uint64_t* offsets, uint32_t n_queries, float threshold,
	uint32_t* n_query_kmers, uint32_t filter0, const kwg_hit_t** d_hits, uint64_t* n_hits);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU search: every device holds a column slab of the database (the reference's unit is the database file,
 * one OpenMP thread each: kwage.cpp:76-87) and searches all the queries; the per-slab hit lists are gathered on the
 * root with one exchange over NCCL and merged into kwg_search's (query, filter) order with global filter indices --
 * what the reference's critical section kwage.cpp:154-177 does with its thread-local maps.
 * One kwg_comm_t per device; either one per process (kwg_comm_create with an id made by rank 0 and handed to the other
 * ranks by the host's own means: MPI, torch.distributed, a file) or all of them in one process (kwg_comm_create_all,
 * then one host thread per device).  NCCL is loaded at run time (libnccl.so.2); without it these calls fail with
 * KWG_ERR_CUDA and nothing else in the library is affected.
 * ------------------------------------------------------------------------------------------ */
typedef struct kwg_comm kwg_comm_t;
#define KWG_COMM_ID_BYTES 128
int kwg_comm_get_unique_id(uint8_t* id /* KWG_COMM_ID_BYTES */);
int kwg_comm_create(kwg_comm_t** out, int device, int n_ranks, int rank, const uint8_t* id);
int kwg_comm_create_all(kwg_comm_t** out /* n */, int n, const int* devices);
void kwg_comm_destroy(kwg_comm_t* c);
/* Collective: every rank calls it with the same queries and threshold and its own slab (filter0 = first column of the
 * slab).  On the root *hits / *n_hits receive the merged list (release with kwg_free_hits); elsewhere *n_hits = 0. */
int kwg_search_gather(kwg_db_t* db, kwg_comm_t* comm, int root, const char* bases, const uint64_t* offsets, uint32_t n_queries,
	float threshold, uint32_t filter0, uint32_t* n_query_kmers, kwg_hit_t** hits, uint64_t* n_hits);
/* The merge step on its own (host only): n_lists lists laid end to end, each ordered by (query, filter), list r holding
 * the columns before those of list r + 1; out receives them ordered by (query, filter). */
void kwg_merge_hits(const kwg_hit_t* lists, const uint64_t* list_len, uint32_t n_lists, uint32_t n_queries, kwg_hit_t* out);

/* ------------------------------------------------------------------------------------------
 * Device-side synthetic inputs for benchmarks (same generators as oracle/kwage_oracle.c).
 * ------------------------------------------------------------------------------------------ */
int kwg_synth_reads_dev(int device, uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t read_len,
	char* d_bases, uint64_t* d_offsets /* n_reads+1, may be NULL */, void* stream);
int kwg_synth_filter_bits_dev(int device, uint64_t seed, uint64_t first_filter, uint32_t n_filters,
	uint64_t filter_bytes, uint64_t filter_pitch, uint8_t* d_filters, void* stream);
/* Known positives for the search benchmark: sets bit `column` of the rows of every k-mer of the first plant_len
 * bases of queries query_first, query_first + query_stride, ... (n_planted of them) in a device slice slab. */
int kwg_synth_plant_dev(int device, uint8_t* d_slab, uint64_t row_pitch, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len,
	uint32_t column, const char* d_bases, const uint64_t* d_offsets, uint32_t query_first, uint32_t query_stride, uint32_t n_planted,
	uint32_t plant_len, void* stream);

/* Per-kernel device timing (CUDA events on the handle's stream around every launch).  Enable, run,
 * then get: ms[i] / launches[i] accumulate since the last get.  Kernel ids: */
#define KWG_T_SCAN_A 0      /* bloom: partition_scan_kernel (counting) or the raw-insert scan */
#define KWG_T_SCAN_B 1      /* bloom: kmer_scan_kernel pass B (valid-word list) */
#define KWG_T_INSERT 2      /* bloom: insert_words_kernel (finalize) */
#define KWG_T_AUX 3         /* bloom: mark_read_starts / flatten;  db: query_kmers_kernel */
#define KWG_T_SEARCH 4      /* db: search_count_kernel */
#define KWG_T_HITS 5        /* db: hits_kernel */
#define KWG_T_REGROUP 6     /* bloom: group_count + group_prefix + regroup_kernel (level-2 partition) */
#define KWG_T_RESOLVE 7     /* bloom: resolve_kernel (first-touch resolution in shared memory) */
#define KWG_T_COUNT 8
int kwg_bloom_set_timing(kwg_bloom_t* b, int enable);
int kwg_bloom_get_timing(kwg_bloom_t* b, double* ms /* KWG_T_COUNT */, uint64_t* launches /* KWG_T_COUNT */);
int kwg_db_set_timing(kwg_db_t* db, int enable);
int kwg_db_get_timing(kwg_db_t* db, double* ms, uint64_t* launches);

/* Elapsed-time helpers on a handle's stream, so that callers that only see the C ABI can time
 * device work with CUDA events on the stream the kernels actually run on. */
int kwg_bloom_stream(kwg_bloom_t* b, void** stream);
int kwg_db_stream(kwg_db_t* db, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* KWAGE_CUDA_H */
