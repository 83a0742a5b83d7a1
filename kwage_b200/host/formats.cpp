// On-disk formats and parameter search of the host layer.  Byte-exact with the reference's files
// (SURVEY.md appendix A): tests/test_host_files.py compares against files written by the
// unmodified reference (oracle/_ref).
#include "kwage_host.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <sstream>
#include <zlib.h>

namespace kwage {

// ---------------------------------------------------------------- accessions
// 64-bit packing, reference sra_accession.cpp:27-96: low 4 bits = digits-1, the rest is the value
// of "3 letters base 26, then the digits base 10" read left to right.
SraAccession str_to_accession(const std::string& s)
{
	size_t letters = 0, digits = 0;
	uint64_t value = 0;
	for (size_t i = 0; i < s.size(); ++i) {
		const int c = std::toupper((unsigned char)s[i]);
		if (c >= 'A' && c <= 'Z') { ++letters; value = value * 26 + (uint64_t)(c - 'A'); }
		else if (s[i] >= '0' && s[i] <= '9') { ++digits; value = value * 10 + (uint64_t)(s[i] - '0'); }
	}
	if (letters != 3 || digits == 0 || digits > 10) throw __FILE__ ":str_to_accession: Unable to parse accession string";
	const SraAccession packed = (digits - 1) | (value << 4);
	if (packed == INVALID_ACCESSION) throw __FILE__ ":str_to_accession: Mapped input string to INVALID_ACCESSION";
	return packed;
}

std::string accession_to_str(const SraAccession& a)
{
	const size_t digits = (a & 0xF) + 1;
	uint64_t value = (a >> 4) & 0x0FFFFFFFFFFFFFFFull;
	std::string out;
	for (size_t i = 0; i < digits; ++i) { out.push_back(char('0' + value % 10)); value /= 10; }
	for (size_t i = 0; i < 3; ++i) { out.push_back(char('A' + value % 26)); value /= 26; }
	std::reverse(out.begin(), out.end());
	return out;
}

// ---------------------------------------------------------------- primitive I/O (binary_io.h:28-54, binary_io.cpp:13-53)
template <class T> static void put(std::ostream& out, const T& v)
{
	out.write(reinterpret_cast<const char*>(&v), sizeof(T));
	if (!out) throw __FILE__ ":binary_write<>: Unable to write simple";
}
template <class T> static void get(std::istream& in, T& v)
{
	in.read(reinterpret_cast<char*>(&v), sizeof(T));
	if (!in) throw __FILE__ ":binary_read<>: Unable to read simple";
}
static void put_str(std::ostream& out, const std::string& s)   // NUL terminated, no length prefix
{
	out.write(s.c_str(), (std::streamsize)s.size() + 1);
	if (!out) throw __FILE__ ":binary_write<string>: Unable to write string";
}
static void get_str(std::istream& in, std::string& s)
{
	s.clear();
	for (;;) {
		char c = 0;
		in.read(&c, 1);
		if (!in) throw __FILE__ ":binary_read<string>: Unable to read string";
		if (c == '\0') break;
		s.push_back(c);
	}
}

void binary_write(std::ostream& out, const BloomParam& p)
{
	put(out, p.kmer_len); put(out, p.log_2_filter_len); put(out, p.num_hash); put(out, p.hash_func);
}
void binary_read(std::istream& in, BloomParam& p)
{
	get(in, p.kmer_len); get(in, p.log_2_filter_len); get(in, p.num_hash); get(in, p.hash_func);
}

void binary_write(std::ostream& out, const FilterInfo& f)       // field order: bloom.h:478-496
{
	put(out, f.run_accession); put(out, f.experiment_accession);
	put_str(out, f.experiment_title); put_str(out, f.experiment_design_description); put_str(out, f.experiment_library_name);
	put_str(out, f.experiment_library_strategy); put_str(out, f.experiment_library_source);
	put_str(out, f.experiment_library_selection); put_str(out, f.experiment_instrument_model);
	put(out, f.sample_accession); put_str(out, f.sample_taxa);
	put(out, (uint64_t)f.sample_attributes.size());
	for (std::unordered_map<std::string, std::string>::const_iterator i = f.sample_attributes.begin(); i != f.sample_attributes.end(); ++i) {
		put_str(out, i->first); put_str(out, i->second);
	}
	put(out, f.study_accession); put_str(out, f.study_title); put_str(out, f.study_abstract);
	put(out, f.number_of_spots); put(out, f.number_of_bases);
	put(out, f.date_received.day); put(out, f.date_received.month); put(out, f.date_received.year);
}
void binary_read(std::istream& in, FilterInfo& f)
{
	get(in, f.run_accession); get(in, f.experiment_accession);
	get_str(in, f.experiment_title); get_str(in, f.experiment_design_description); get_str(in, f.experiment_library_name);
	get_str(in, f.experiment_library_strategy); get_str(in, f.experiment_library_source);
	get_str(in, f.experiment_library_selection); get_str(in, f.experiment_instrument_model);
	get(in, f.sample_accession); get_str(in, f.sample_taxa);
	uint64_t n = 0;
	get(in, n);
	f.sample_attributes.clear();
	for (uint64_t i = 0; i < n; ++i) {      // same insertion sequence as the reference -> same iteration order on rewrite
		std::pair<std::string, std::string> kv;
		get_str(in, kv.first); get_str(in, kv.second);
		f.sample_attributes.insert(kv);
	}
	get(in, f.study_accession); get_str(in, f.study_title); get_str(in, f.study_abstract);
	get(in, f.number_of_spots); get(in, f.number_of_bases);
	get(in, f.date_received.day); get(in, f.date_received.month); get(in, f.date_received.year);
}

void binary_write(std::ostream& out, const DBFileHeader& h)     // kwage.h:34-44, no padding: 44 bytes
{
	put(out, h.magic); put(out, h.version); put(out, h.crc32); put(out, h.kmer_len); put(out, h.num_hash);
	put(out, h.log_2_filter_len); put(out, h.num_filter); put(out, h.hash_func); put(out, h.compression); put(out, h.info_start);
}
void binary_read(std::istream& in, DBFileHeader& h)
{
	get(in, h.magic); get(in, h.version); get(in, h.crc32); get(in, h.kmer_len); get(in, h.num_hash);
	get(in, h.log_2_filter_len); get(in, h.num_filter); get(in, h.hash_func); get(in, h.compression); get(in, h.info_start);
}

uint32_t crc32_bytes(uint32_t crc, const uint8_t* p, size_t n)
{
	return (uint32_t)::crc32_z(crc, p, n);
}

void write_bloom_file(std::ostream& out, const BloomParam& param, const FilterInfo& info, const uint8_t* bits)
{
	const size_t nbytes = param.filter_len() / 8 + ((param.filter_len() % 8) ? 1 : 0);
	write_bloom_file(out, param, info, bits, crc32_bytes((uint32_t)::crc32_z(0L, Z_NULL, 0), bits, nbytes));
}

// crc: BitVector::crc32 of the bits (bloom.cpp:328-336), e.g. from kwg_bloom_finalize_crc
void write_bloom_file(std::ostream& out, const BloomParam& param, const FilterInfo& info, const uint8_t* bits, uint32_t crc)
{
	const std::streampos begin = out.tellp();
	put(out, (unsigned char)KWAGE_BLOOM_MAGIC_IN_PROGRESS);
	binary_write(out, param);
	const size_t nbytes = param.filter_len() / 8 + ((param.filter_len() % 8) ? 1 : 0);
	put(out, crc);
	binary_write(out, info);
	out.write(reinterpret_cast<const char*>(bits), (std::streamsize)nbytes);
	const std::streampos end = out.tellp();
	out.seekp(begin);                       // the record is complete: flip the guard byte
	put(out, (unsigned char)KWAGE_BLOOM_MAGIC_COMPLETE);
	out.seekp(end);
	if (!out) throw __FILE__ ":binary_write<BloomFilter>: Unable to write BloomFilter";
}

void read_bloom_header(std::istream& in, BloomFileHeader& h)
{
	unsigned char magic = 0;
	get(in, magic);
	if (magic != KWAGE_BLOOM_MAGIC_COMPLETE) throw __FILE__ ":binary_read<BloomFilter>: Filter record is not complete!";
	binary_read(in, h.param);
	get(in, h.crc32);
	binary_read(in, h.info);
	h.bits_start = in.tellg();
}

// ---------------------------------------------------------------- FilterInfo text output
std::string FilterInfo::csv_string() const { return accession_to_str(run_accession); }

std::string FilterInfo::json_string(const std::string& prefix) const
{
	std::stringstream out;
	bool any = false;
	struct Emit {
		std::stringstream& out; const std::string& prefix; bool& any;
		void operator()(const char* key, const std::string& value) const
		{
			if (any) out << ",\n";
			out << prefix << '"' << key << "\": \"" << value << '"';
			any = true;
		}
	} emit = {out, prefix, any};
	if (run_accession != INVALID_ACCESSION) emit("run", accession_to_str(run_accession));
	if (date_received.is_valid()) {
		std::stringstream d;
		d << date_received.year << '-' << date_received.month << '-' << date_received.day;
		emit("date received", d.str());
	}
	if (experiment_accession != INVALID_ACCESSION) emit("experiment", accession_to_str(experiment_accession));
	if (!experiment_title.empty()) emit("experiment title", experiment_title);
	if (!experiment_design_description.empty()) emit("experiment design", experiment_design_description);
	if (!experiment_library_name.empty()) emit("experiment library name", experiment_library_name);
	if (!experiment_library_strategy.empty()) emit("experiment library strategy", experiment_library_strategy);
	if (!experiment_library_source.empty()) emit("experiment library source", experiment_library_source);
	if (!experiment_library_selection.empty()) emit("experiment library selection", experiment_library_selection);
	if (!experiment_instrument_model.empty()) emit("experiment instrument model", experiment_instrument_model);
	if (sample_accession != INVALID_ACCESSION) emit("sample", accession_to_str(sample_accession));
	if (!sample_taxa.empty()) emit("sample taxa", sample_taxa);
	if (!sample_attributes.empty()) {
		if (any) out << ",\n";
		out << prefix << "\"sample attributes\": [\n";
		bool first = true;
		for (std::unordered_map<std::string, std::string>::const_iterator i = sample_attributes.begin(); i != sample_attributes.end(); ++i) {
			if (!first) out << ",\n";
			out << prefix << "\t{\n" << prefix << "\t\t\"tag\": \"" << i->first << "\",\n" << prefix << "\t\t\"value\": \"" << i->second
			    << "\"\n" << prefix << "\t}";
			first = false;
		}
		out << '\n' << prefix << ']';
		any = true;
	}
	if (study_accession != INVALID_ACCESSION) emit("study", accession_to_str(study_accession));
	if (!study_title.empty()) emit("study title", study_title);
	if (!study_abstract.empty()) emit("study abstract", study_abstract);
	return out.str();
}

// ---------------------------------------------------------------- parameter search (host math, same libm as the reference)
BloomParam optimal_bloom_param(const uint32_t& kmer_len, const size_t& num_kmer, const float& p_max, const HashFunction& func,
	const uint32_t& min_log2, const uint32_t& max_log2)
{
	if (num_kmer == 0) throw __FILE__ ":optimal_bloom_param: No kmers found";
	BloomParam best;
	best.hash_func = func;
	best.kmer_len = kmer_len;
	bool found = false;
	// smallest length first; inside a length the hash count with the lowest false-positive rate
	for (best.log_2_filter_len = min_log2; best.log_2_filter_len <= max_log2; ++best.log_2_filter_len) {
		float best_p = 10.0f;
		const uint64_t len = 1ULL << best.log_2_filter_len;
		for (uint32_t h = KWAGE_MIN_NUM_HASH; h <= KWAGE_MAX_NUM_HASH; ++h) {
			const double p = std::pow(1.0 - std::pow(1.0 - 1.0 / len, num_kmer * h), h);
			if (p <= p_max && p < best_p) {
				best_p = p;
				best.num_hash = h;
				found = true;
			}
		}
		if (found) return best;
	}
	throw __FILE__ ":optimal_bloom_param: Unable to satisfy Bloom filter probability bound";
}

size_t approximate_max_kmers(const float& p_max, const HashFunction&, const uint32_t& min_log2, const uint32_t& max_log2)
{
	for (size_t lk = 1; lk < 8 * sizeof(size_t); ++lk) {
		const size_t num_kmer = size_t(1) << lk;
		bool found = false;
		for (size_t L = min_log2; L <= max_log2 && !found; ++L) {
			const float best_p = 10.0f;
			const uint64_t len = 1ULL << L;
			for (uint32_t h = KWAGE_MIN_NUM_HASH; h <= KWAGE_MAX_NUM_HASH && !found; ++h) {
				const double p = std::pow(1.0 - std::pow(1.0 - 1.0 / len, num_kmer * h), h);
				if (p <= p_max && p < best_p) found = true;
			}
		}
		if (!found) return num_kmer;   // smallest power of two that cannot be stored
	}
	return 0xFFFFFFFFFFFFFFFFull;
}

uint32_t counting_filter_log2_len(uint64_t num_bp)
{
	const uint64_t lo = 18, hi = 32;                    // make_bloom.cpp:21-22
	if (num_bp == 0) return (uint32_t)hi;               // no metadata: worst case (make_bloom.cpp:106)
	// two counting filters, two hashes each, false-positive target 1e-2 (make_bloom.cpp:25,116)
	const double len = 1.0 / (1.0 - std::pow(1.0 - std::pow(1.0e-2, 1.0 / 4.0), 1.0 / (2 * num_bp)));
	uint64_t L = (uint64_t)std::ceil(std::log(len) / std::log(2.0));
	if (L > hi) L = hi;
	if (L < lo) L = lo;
	return (uint32_t)L;
}

} // namespace kwage
