// kwage_host.h -- C++ host layer above the C ABI (include/kwage_cuda.h).
//
// It mirrors, name for name, the reference's interface for the three stages of the hot path so
// that a user of the reference finds the same entry points, argument meaning, status codes, error
// behaviour and on-disk formats -- with the compute delegated to libkwage_cuda.so:
//
//   make_bloom_filter()  <-> reference make_bloom.cpp:76   (maestro.h:129-131)
//   build_db()           <-> reference build_db.cpp:24     (maestro.h:123)
//   search()             <-> reference kwage.cpp:340       (kwage.cpp:26-31)
//   BloomParam, optimal_bloom_param, approximate_max_kmers <-> bloom.h:546-621, bloom.cpp:10-121
//   FilterInfo, DBFileHeader, .bloom/.db writers/readers   <-> bloom.h:474-537, kwage.h:30-72,
//                                                              binary_io.cpp:182-265
//
// Error convention = the reference's: helpers throw `const char*`; stage functions catch and map
// to status bytes / bool.  There is no CPU implementation of the compute here: if the CUDA library
// reports an error the stage fails (STATUS_BLOOM_FAIL / false / rethrow).
#ifndef KWAGE_B200_HOST_H
#define KWAGE_B200_HOST_H

#include <cstdint>
#include <deque>
#include <fstream>
#include <iostream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/kwage_cuda.h"

namespace kwage {

// ---- status bytes (reference maestro.h:27-31)
enum : unsigned char {
	STATUS_BLOOM_SUCCESS = 14,
	STATUS_BLOOM_FAIL = 15,
	STATUS_BLOOM_INVALID = 16,
	STATUS_DATABASE_SUCCESS = 17,
	STATUS_DATABASE_FAIL = 18
};

// ---- hash function ids (reference hash.h:8-11)
enum { MURMUR_HASH_32 = 0, UNKNOWN_HASH = 1 };
typedef int HashFunction;

#define KWAGE_MIN_NUM_HASH 1             // bloom.h:20
#define KWAGE_MAX_NUM_HASH 5             // bloom.h:21
#define KWAGE_BLOOM_MAGIC_IN_PROGRESS 0x00   // bloom.h:25
#define KWAGE_BLOOM_MAGIC_COMPLETE 0xFF      // bloom.h:28
#define KWAGE_MAGIC_NUMBER 0x20191025u       // kwage.h:22
#define KWAGE_DBFILE_VERSION 2u              // kwage.h:23
#define KWAGE_MAX_COUNT 15u                  // make_bloom.cpp:61

// ---- accessions (reference sra_accession.cpp:27-96)
typedef uint64_t SraAccession;
const SraAccession INVALID_ACCESSION = 0;
SraAccession str_to_accession(const std::string& s);
std::string accession_to_str(const SraAccession& a);

struct Date {
	uint32_t day, month, year;
	Date() : day(0), month(0), year(0) {}
	bool is_valid() const { return year != 0 && month != 0 && day != 0; }
};

// ---- Bloom parameters (reference bloom.h:546-621)
struct BloomParam {
	uint32_t kmer_len;
	uint32_t log_2_filter_len;
	uint32_t num_hash;
	HashFunction hash_func;
	BloomParam() : kmer_len(0), log_2_filter_len(0), num_hash(0), hash_func(0) {}
	bool operator==(const BloomParam& r) const
	{
		return kmer_len == r.kmer_len && log_2_filter_len == r.log_2_filter_len && num_hash == r.num_hash && hash_func == r.hash_func;
	}
	bool operator!=(const BloomParam& r) const { return !(*this == r); }
	size_t filter_len() const { return size_t(1) << log_2_filter_len; }
};

// throws "…: No kmers found" / "…: Unable to satisfy Bloom filter probability bound" (bloom.cpp:16,66)
BloomParam optimal_bloom_param(const uint32_t& kmer_len, const size_t& num_kmer, const float& p, const HashFunction& func,
	const uint32_t& min_log_2_filter_len, const uint32_t& max_log_2_filter_len);
size_t approximate_max_kmers(const float& p, const HashFunction& func, const uint32_t& min_log_2_filter_len,
	const uint32_t& max_log_2_filter_len);
// counting-filter length from the metadata base count (make_bloom.cpp:104-129); 0 bases -> 32
uint32_t counting_filter_log2_len(uint64_t num_bp);

// ---- SRA metadata carried with every filter (reference bloom.h:474-537); serialised field by field
struct FilterInfo {
	SraAccession run_accession, experiment_accession;
	std::string experiment_title, experiment_design_description, experiment_library_name, experiment_library_strategy,
		experiment_library_source, experiment_library_selection, experiment_instrument_model;
	SraAccession sample_accession;
	std::string sample_taxa;
	std::unordered_map<std::string, std::string> sample_attributes;
	SraAccession study_accession;
	std::string study_title, study_abstract;
	uint64_t number_of_spots, number_of_bases;
	Date date_received;
	FilterInfo() : run_accession(0), experiment_accession(0), sample_accession(0), study_accession(0), number_of_spots(0), number_of_bases(0) {}
	std::string csv_string() const;                              // bloom.cpp:123-126
	std::string json_string(const std::string& prefix) const;    // bloom.cpp:128-326
};
void binary_write(std::ostream& out, const FilterInfo& info);
void binary_read(std::istream& in, FilterInfo& info);
void binary_write(std::ostream& out, const BloomParam& p);
void binary_read(std::istream& in, BloomParam& p);

// ---- database file header (reference kwage.h:30-72), 44 bytes on disk
struct DBFileHeader {
	uint32_t magic, version, crc32, kmer_len, num_hash, log_2_filter_len, num_filter;
	HashFunction hash_func;
	uint32_t compression;
	uint64_t info_start;
	DBFileHeader() : magic(KWAGE_MAGIC_NUMBER), version(KWAGE_DBFILE_VERSION), crc32(0), kmer_len(0), num_hash(0), log_2_filter_len(0),
		num_filter(0), hash_func(0), compression(0), info_start(0) {}
	size_t filter_len() const { return size_t(1) << log_2_filter_len; }
};
void binary_write(std::ostream& out, const DBFileHeader& h);
void binary_read(std::istream& in, DBFileHeader& h);

// .bloom writer/reader (reference binary_io.cpp:182-237): magic, param, crc32(bits), info, bits
void write_bloom_file(std::ostream& out, const BloomParam& param, const FilterInfo& info, const uint8_t* bits);
void write_bloom_file(std::ostream& out, const BloomParam& param, const FilterInfo& info, const uint8_t* bits, uint32_t crc);
struct BloomFileHeader { BloomParam param; uint32_t crc32; FilterInfo info; std::streampos bits_start; };
void read_bloom_header(std::istream& in, BloomFileHeader& h);   // leaves the stream at the first bit byte
uint32_t crc32_bytes(uint32_t crc, const uint8_t* p, size_t n); // zlib crc32_z (bloom.cpp:328-336)

// ---- options (the fields of the reference's MaestroOptions / SearchOptions that the stages read)
struct MaestroOptions {
	float false_positive_probability;   // options.h:147  default 0.25
	unsigned int min_kmer_count;        // default 5 (options.h:152)
	unsigned int kmer_len;              // default 31
	unsigned int min_log_2_filter_len;  // default 18
	unsigned int max_log_2_filter_len;  // default 32
	HashFunction hash_func;
	bool verbose;
	int device;                         // new: CUDA device of this worker rank (local_rank % n_gpu)
	MaestroOptions() : false_positive_probability(0.25f), min_kmer_count(5), kmer_len(31), min_log_2_filter_len(18),
		max_log_2_filter_len(32), hash_func(MURMUR_HASH_32), verbose(false), device(0) {}
};

struct SearchOptions {
	enum { OUTPUT_CSV, OUTPUT_JSON };
	float threshold;                    // default 1.0 (options.h:153)
	int output_format;                  // default JSON (options.h:154)
	int device;
	SearchOptions() : threshold(1.0f), output_format(OUTPUT_JSON), device(0) {}
};

// ---- progress record (reference maestro.h:67-102)
struct BloomProgress {
	size_t num_primary_align, curr_primary_align, num_unaligned_read, curr_unaligned_read, num_read, curr_read, curr_fragment,
		num_kmer, num_bp, log_2_counting_filter_len;
	std::string error;
	bool valid_read_collection;
	BloomProgress() : num_primary_align(0), curr_primary_align(0), num_unaligned_read(0), curr_unaligned_read(0), num_read(0),
		curr_read(0), curr_fragment(0), num_kmer(0), num_bp(0), log_2_counting_filter_len(0), valid_read_collection(false) {}
};

// ---- read streaming: stands where the NGS ReadCollection stands in the reference
// (make_bloom.cpp:170-300).  One call = one fragment, in stream order.
class ReadSource {
public:
	virtual ~ReadSource() {}
	virtual bool next_fragment(std::string& bases) = 0;
	virtual uint64_t read_count() const { return 0; }
};
// "<dir>/<accession>.reads" (one read per line), .fa/.fasta/.fna/.fastq (optionally .gz): the
// reference's SequenceIterator role (parse_sequence.cpp:72-262).  Throws const char* when missing.
ReadSource* open_read_collection(const std::string& accession_or_path);
// SRA metadata hook (sra_meta.cpp:17): defaults to the KWAGE_NUM_BASES environment variable or the
// size of the reads file; return 0 for "unknown" (-> 4 GiB-slot counting filter like the reference).
uint64_t number_of_bases(const std::string& accession);
void set_number_of_bases_hook(uint64_t (*hook)(const std::string&));

// 2-bit packing of a fragment into a batch for kwg_bloom_add_packed (NCBI 2na + not-a-base mask); see stages.cpp
bool pack_2na(uint8_t* packed, uint8_t* mask, uint64_t cursor, const char* bases, size_t n);

// ---- the three stages
unsigned char make_bloom_filter(const SraAccession& acc, const FilterInfo& info, BloomParam& param, BloomProgress& progress,
	const std::string& bloom_dir, const MaestroOptions& opt, bool force_unaligned = false);
unsigned char make_bloom_filter(ReadSource& reads, uint64_t num_bp, const SraAccession& acc, const FilterInfo& info, BloomParam& param,
	BloomProgress& progress, const std::string& bloom_dir, const MaestroOptions& opt);

bool build_db(const std::string& filename, const BloomParam& param, const std::deque<std::string>& bloom_files);
void set_build_db_device(int device);

// merge_db (reference merge_db.cpp:278-820): append file_2's filters to file_1's columns, overflow becomes the new file_2
std::pair<size_t, std::string> merge_database_files(const std::string& file_1, const std::string& file_2, const size_t& max_num_filters, int device = 0);
size_t max_filters_per_database_file(uint32_t log_2_filter_len);   // merge_db.cpp:88-100

struct MatchResult {                     // reference output.h:9-33
	unsigned int num_kmers_found, num_query_kmer;
	FilterInfo subject_info;
	MatchResult() : num_kmers_found(0), num_query_kmer(0) {}
	MatchResult(unsigned int found, unsigned int n, const FilterInfo& info) : num_kmers_found(found), num_query_kmer(n), subject_info(info) {}
	bool operator<(const MatchResult& r) const { return num_kmers_found > r.num_kmers_found; }   // descending
};

// One database file resident in HBM: the GPU-side counterpart of the reference's open ifstream.
class SubjectDatabase {
public:
	explicit SubjectDatabase(const std::string& filename, int device = 0);
	// Several database files with the same (kmer_len, num_hash, log_2_filter_len, hash_func) side by side in ONE column
	// slab: the reference keeps <= 2048 filters per file (options.h:137-138), i.e. 256-byte slices, and wide slices are
	// gathered three times faster.  Filter indices of the slab run through the files in the order given; header() is the
	// first file's with num_filter = the sum.  Throws if the files do not agree.
	explicit SubjectDatabase(const std::vector<std::string>& filenames, int device = 0);
	~SubjectDatabase();
	// can these files share a slab?  (reads the two headers)
	static bool compatible(const std::string& file_a, const std::string& file_b);
	// bytes of HBM the slices of this file need as part of a slab
	static uint64_t slab_bytes(const std::string& filename);
	const DBFileHeader& header() const { return hdr; }
	// Batched form of search(): all queries against this file in one pass over the device.
	// Returns true if any query matched.  results[query_id] gets one MatchResult per matching filter.
	bool search(std::unordered_map<size_t, std::deque<MatchResult> >& results, const std::vector<std::string>& queries,
		const std::vector<size_t>& query_ids, const SearchOptions& opt);
	// Multi-device form: every device holds one slab; all of them search the same queries and the hit lists are gathered
	// on `root` with one NCCL exchange inside the library (kwg_search_gather), the stand-in for the reference's critical
	// section over thread-local maps (kwage.cpp:154-177).  Collective: one host thread per device calls it.  filter0 =
	// first global filter index of this slab.  On the root, hits (global filter indices) and n_kmers are filled.
	void search_gather(kwg_comm_t* comm, int root, uint32_t filter0, const std::vector<std::string>& queries, float threshold,
		std::vector<kwg_hit_t>& hits, std::vector<uint32_t>& n_kmers);
	FilterInfo filter_info(uint32_t filter);   // kwage.cpp:505-515
private:
	SubjectDatabase(const SubjectDatabase&);
	SubjectDatabase& operator=(const SubjectDatabase&);
	void open_files(const std::vector<std::string>& filenames, int device);
	struct Part { std::ifstream* fin; DBFileHeader hdr; uint32_t col_begin; uint64_t slices_start; };
	std::vector<Part> parts;     // one per file, in column order
	DBFileHeader hdr;
	kwg_db_t* db;
};

// Same contract as the reference's search() (kwage.cpp:340): one query against one database.
bool search(std::unordered_map<size_t, std::deque<MatchResult> >& results, SubjectDatabase& subject, const std::string& query,
	const size_t& query_id, const SearchOptions& opt);

// output (reference output.h:35-112)
void write_csv_header(std::ostream& out);
void write_csv(std::ostream& out, const std::string& query, const std::deque<MatchResult>& r);
void write_json_header(std::ostream& out, bool multiple);
void write_json(std::ostream& out, const std::string& query, bool multiple, bool first, const float& threshold, const std::deque<MatchResult>& r);
void write_json_footer(std::ostream& out, bool multiple);

} // namespace kwage
#endif
