// extern "C" view of the host layer so that tests and bench.py (Python, ctypes) can drive the same
// C++ entry points a C++ caller uses.  Thin: no logic of its own.
#include <algorithm>
#include <chrono>
#include <cstring>

#include "kwage_host.h"

using namespace kwage;

extern "C" {

int kwh_optimal_bloom_param(uint32_t kmer_len, uint64_t num_kmer, float p, uint32_t min_log2, uint32_t max_log2,
	uint32_t* log2_len, uint32_t* num_hash)
{
	try {
		const BloomParam r = optimal_bloom_param(kmer_len, num_kmer, p, MURMUR_HASH_32, min_log2, max_log2);
		*log2_len = r.log_2_filter_len;
		*num_hash = r.num_hash;
		return 0;
	}
	catch (...) { return -1; }
}

uint64_t kwh_approximate_max_kmers(float p, uint32_t min_log2, uint32_t max_log2)
{
	return approximate_max_kmers(p, MURMUR_HASH_32, min_log2, max_log2);
}

uint32_t kwh_counting_filter_log2_len(uint64_t num_bp) { return counting_filter_log2_len(num_bp); }

uint64_t kwh_str_to_accession(const char* s)
{
	try { return str_to_accession(s); } catch (...) { return 0; }
}

void kwh_accession_to_str(uint64_t acc, char* out, size_t cap)
{
	const std::string s = accession_to_str(acc);
	std::strncpy(out, s.c_str(), cap - 1);
	out[cap - 1] = 0;
}

int kwh_pack_2na(uint8_t* packed, uint8_t* mask, uint64_t cursor, const char* bases, uint64_t n)
{
	return pack_2na(packed, mask, cursor, bases, (size_t)n) ? 1 : 0;
}

// the read streaming alone (open_read_collection + next_fragment): fragments, bases, longest fragment and an FNV-1a hash of
// the fragments with a separator after each -- what the parser tests compare with a plain Python parse of the same file
int kwh_parse_digest(const char* path, uint64_t* out /* [4] */)
{
	try {
		ReadSource* src = open_read_collection(path);
		std::string frag;
		uint64_t n = 0, bases = 0, longest = 0, h = 1469598103934665603ull;
		while (src->next_fragment(frag)) {
			++n; bases += frag.size(); longest = std::max<uint64_t>(longest, frag.size());
			for (size_t i = 0; i < frag.size(); ++i) { h ^= (unsigned char)frag[i]; h *= 1099511628211ull; }
			h ^= 0xFFu; h *= 1099511628211ull;
		}
		delete src;
		out[0] = n; out[1] = bases; out[2] = longest; out[3] = h;
		return 0;
	}
	catch (const char*) { return 1; }
}

// make_bloom_filter() on a reads file; results through plain out-parameters
int kwh_make_bloom_file(const char* accession, const char* reads_path, uint64_t num_bp, const char* bloom_dir, uint32_t kmer_len,
	uint32_t min_kmer_count, float p, uint32_t min_log2, uint32_t max_log2, int device,
	uint64_t* num_kmer, uint32_t* log2_len, uint32_t* num_hash, uint32_t* log2_count_len, char* error, size_t error_cap,
	uint64_t* progress_out /* NULL or [3]: num_bp, curr_read, curr_fragment of the progress record */)
{
	MaestroOptions opt;
	opt.kmer_len = kmer_len; opt.min_kmer_count = min_kmer_count; opt.false_positive_probability = p;
	opt.min_log_2_filter_len = min_log2; opt.max_log_2_filter_len = max_log2; opt.device = device;
	FilterInfo info;
	BloomParam param;
	BloomProgress progress;
	unsigned char status = STATUS_BLOOM_FAIL;
	try {
		info.run_accession = str_to_accession(accession);
		ReadSource* src = open_read_collection(reads_path);
		status = make_bloom_filter(*src, num_bp, info.run_accession, info, param, progress, bloom_dir, opt);
		delete src;
	}
	catch (const char* e) { progress.error = e; }
	*num_kmer = progress.num_kmer; *log2_len = param.log_2_filter_len; *num_hash = param.num_hash;
	*log2_count_len = (uint32_t)progress.log_2_counting_filter_len;
	if (progress_out) { progress_out[0] = progress.num_bp; progress_out[1] = progress.curr_read; progress_out[2] = progress.curr_fragment; }
	if (error && error_cap) { std::strncpy(error, progress.error.c_str(), error_cap - 1); error[error_cap - 1] = 0; }
	return status;
}

// write a .bloom file for raw filter bits (metadata: run accession only)
int kwh_write_bloom_file(const char* path, const char* accession, uint32_t kmer_len, uint32_t log2_len, uint32_t num_hash, const uint8_t* bits)
{
	try {
		BloomParam param;
		param.kmer_len = kmer_len; param.log_2_filter_len = log2_len; param.num_hash = num_hash; param.hash_func = MURMUR_HASH_32;
		FilterInfo info;
		info.run_accession = str_to_accession(accession);
		std::ofstream fout(path, std::ios::binary);
		if (!fout) return 0;
		write_bloom_file(fout, param, info, bits);
		return fout ? 1 : 0;
	}
	catch (...) { return 0; }
}

// build_db(); bloom file paths separated by '\n'
int kwh_build_db(const char* filename, uint32_t kmer_len, uint32_t log2_len, uint32_t num_hash, const char* bloom_paths, int device)
{
	BloomParam param;
	param.kmer_len = kmer_len; param.log_2_filter_len = log2_len; param.num_hash = num_hash; param.hash_func = MURMUR_HASH_32;
	std::deque<std::string> files;
	const char* p = bloom_paths;
	while (*p) {
		const char* e = std::strchr(p, '\n');
		const std::string s = e ? std::string(p, e) : std::string(p);
		if (!s.empty()) files.push_back(s);
		if (!e) break;
		p = e + 1;
	}
	set_build_db_device(device);
	return build_db(filename, param, files) ? 1 : 0;
}

// SubjectDatabase(path): wall-clock seconds of loading the file's slice region into HBM (bench: db_load); < 0 on error
double kwh_db_load_seconds(const char* paths, int device, uint64_t* slab_bytes)
{
	try {
		std::vector<std::string> files;
		const char* p = paths;
		while (*p) {
			const char* e = std::strchr(p, '\n');
			const std::string s = e ? std::string(p, e) : std::string(p);
			if (!s.empty()) files.push_back(s);
			if (!e) break;
			p = e + 1;
		}
		uint64_t total = 0;
		for (size_t i = 0; i < files.size(); ++i) total += SubjectDatabase::slab_bytes(files[i]);
		const auto t0 = std::chrono::steady_clock::now();
		SubjectDatabase db(files, device);
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (slab_bytes) *slab_bytes = total;
		return sec;
	}
	catch (const char* e) { std::cerr << e << std::endl; return -1.0; }
	catch (...) { return -1.0; }
}

// merge_database_files(); returns the remaining-capacity filter count (>= 0) or -1 on error
long kwh_merge_db(const char* file_1, const char* file_2, uint64_t max_num_filters, int device, char* error, size_t error_cap)
{
	try {
		const size_t max_f = max_num_filters ? (size_t)max_num_filters : 0;
		size_t m = max_f;
		if (!m) {
			std::ifstream f(file_1, std::ios::binary);
			DBFileHeader h;
			binary_read(f, h);
			m = max_filters_per_database_file(h.log_2_filter_len);
		}
		return (long)merge_database_files(file_1, file_2, m, device).first;
	}
	catch (const char* e) {
		if (error && error_cap) { std::strncpy(error, e, error_cap - 1); error[error_cap - 1] = 0; }
		return -1;
	}
	catch (...) { return -1; }
}

} // extern "C"
