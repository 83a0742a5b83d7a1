// The three stage functions of the hot path on top of the C ABI.  Host code keeps what the
// reference keeps on the host (streaming, parameter choice, CRCs, files, metadata); everything
// the reference computes in its inner loops is a libkwage_cuda call.
#include "kwage_host.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <sys/stat.h>
#include <zlib.h>

namespace kwage {

static void cuda_check(int rc, const char* where)
{
	if (rc == KWG_OK) return;
	// keep the reference's convention: a C string that names the site; the detail goes to stderr
	std::cerr << where << ": " << kwg_last_error() << std::endl;
	throw where;
}

// ================================================================ read streaming
namespace {

class LineReads : public ReadSource {       // "<accession>.reads": one read per line
public:
	explicit LineReads(const std::string& path) : fin(path.c_str()) { if (!fin) throw __FILE__ ":open_read_collection: Unable to open reads file"; }
	bool next_fragment(std::string& bases) { return (bool)std::getline(fin, bases); }
private:
	std::ifstream fin;
};

class GzSequenceReads : public ReadSource {  // FASTA / FASTQ, optionally gzip (zlib reads plain files too)
public:
	explicit GzSequenceReads(const std::string& path) : gz(gzopen(path.c_str(), "rb")), peeked(false), fastq(false), first(true)
	{
		if (!gz) throw __FILE__ ":open_read_collection: Unable to open sequence file";
		gzbuffer(gz, 1 << 20);
	}
	~GzSequenceReads() { if (gz) gzclose(gz); }
	bool next_fragment(std::string& bases)
	{
		std::string line;
		bases.clear();
		if (!next_line(line)) return false;
		if (first) { fastq = !line.empty() && line[0] == '@'; first = false; }
		if (fastq) {                          // @id / bases / + / qualities
			if (!next_line(bases)) return false;
			std::string skip;
			next_line(skip); next_line(skip);
			return true;
		}
		// FASTA: header line, then sequence lines until the next '>'
		while (next_line(line)) {
			if (!line.empty() && line[0] == '>') { pending = line; peeked = true; break; }
			bases += line;
		}
		return true;
	}
private:
	bool next_line(std::string& line)
	{
		if (peeked) { line = pending; peeked = false; return true; }
		line.clear();
		char buf[1 << 16];
		bool any = false;
		while (gzgets(gz, buf, sizeof(buf))) {
			any = true;
			size_t n = std::strlen(buf);
			const bool eol = n && buf[n - 1] == '\n';
			while (n && (buf[n - 1] == '\n' || buf[n - 1] == '\r')) --n;
			line.append(buf, n);
			if (eol) break;
		}
		return any;
	}
	gzFile gz;
	std::string pending;
	bool peeked, fastq, first;
};

bool ends_with(const std::string& s, const char* suffix)
{
	const size_t n = std::strlen(suffix);
	return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

bool file_exists(const std::string& p)
{
	struct stat st;
	return ::stat(p.c_str(), &st) == 0;
}

uint64_t (*g_num_bases_hook)(const std::string&) = NULL;
int g_build_db_device = 0;

} // namespace

ReadSource* open_read_collection(const std::string& name)
{
	static const char* seq_ext[] = {".fa", ".fasta", ".fna", ".fastq", ".fq", ".fa.gz", ".fasta.gz", ".fna.gz", ".fastq.gz", ".fq.gz"};
	for (size_t i = 0; i < sizeof(seq_ext) / sizeof(seq_ext[0]); ++i)
		if (ends_with(name, seq_ext[i]) && file_exists(name)) return new GzSequenceReads(name);
	if (ends_with(name, ".reads") && file_exists(name)) return new LineReads(name);
	const char* dir = std::getenv("KWAGE_READS_DIR");
	const std::string base = std::string(dir ? dir : ".") + "/" + name;
	if (file_exists(base + ".reads")) return new LineReads(base + ".reads");
	for (size_t i = 0; i < sizeof(seq_ext) / sizeof(seq_ext[0]); ++i)
		if (file_exists(base + seq_ext[i])) return new GzSequenceReads(base + seq_ext[i]);
	throw __FILE__ ":open_read_collection: No read collection for accession";
}

void set_number_of_bases_hook(uint64_t (*hook)(const std::string&)) { g_num_bases_hook = hook; }

uint64_t number_of_bases(const std::string& accession)
{
	if (g_num_bases_hook) return g_num_bases_hook(accession);
	if (const char* e = std::getenv("KWAGE_NUM_BASES")) return std::strtoull(e, NULL, 10);
	return 0;     // unknown -> largest counting filter, as in the reference when metadata is missing
}

void set_build_db_device(int device) { g_build_db_device = device; }

// ================================================================ construction
// Mirrors make_bloom_filter() (reference make_bloom.cpp:76-504).  Fragments are batched so that one
// libkwage_cuda call carries ~64 MiB; stream order is preserved (it matters for the counting filter).
unsigned char make_bloom_filter(ReadSource& reads, uint64_t num_bp, const SraAccession& acc, const FilterInfo& info, BloomParam& param,
	BloomProgress& progress, const std::string& bloom_dir, const MaestroOptions& opt)
{
	kwg_bloom_t* builder = NULL;
	try {
		if (opt.min_kmer_count > KWAGE_MAX_COUNT)
			throw __FILE__ ":make_bloom_filter: min_kmer_count is too large. See the comments in make_bloom.cpp for parameter settings.";

		progress.log_2_counting_filter_len = counting_filter_log2_len(num_bp);
		const size_t max_num_kmer = approximate_max_kmers(opt.false_positive_probability, opt.hash_func,
			opt.min_log_2_filter_len, opt.max_log_2_filter_len);

		cuda_check(kwg_bloom_create(&builder, opt.device, opt.kmer_len, opt.min_kmer_count,
			(uint32_t)progress.log_2_counting_filter_len, opt.max_log_2_filter_len), __FILE__ ":make_bloom_filter: kwg_bloom_create failed");
		progress.valid_read_collection = true;

		const size_t batch_bytes = size_t(64) << 20;
		std::string flat, fragment;
		std::vector<uint64_t> offsets(1, 0);
		flat.reserve(batch_bytes + (1 << 20));
		bool more = true;
		while (more) {
			more = reads.next_fragment(fragment);
			if (more) {
				progress.num_bp += fragment.size();
				flat += fragment;
				offsets.push_back(flat.size());
				++progress.curr_read;
				++progress.num_read;
			}
			if ((!more && offsets.size() > 1) || flat.size() >= batch_bytes) {
				cuda_check(kwg_bloom_add_reads(builder, flat.data(), offsets.data(), offsets.size() - 1),
					__FILE__ ":make_bloom_filter: kwg_bloom_add_reads failed");
				uint64_t n = 0;
				cuda_check(kwg_bloom_num_valid(builder, &n), __FILE__ ":make_bloom_filter: kwg_bloom_num_valid failed");
				progress.num_kmer = n;
				flat.clear();
				offsets.assign(1, 0);
				// the reference tests this after every fragment (make_bloom.cpp:208,246,288); per batch
				// the outcome (STATUS_BLOOM_INVALID) is the same, only num_kmer at the abort differs
				if (max_num_kmer < progress.num_kmer) {
					kwg_bloom_destroy(builder);
					return STATUS_BLOOM_INVALID;
				}
			}
		}

		try {
			param = optimal_bloom_param(opt.kmer_len, progress.num_kmer, opt.false_positive_probability, opt.hash_func,
				opt.min_log_2_filter_len, opt.max_log_2_filter_len);
		}
		catch (...) {
			kwg_bloom_destroy(builder);
			return STATUS_BLOOM_INVALID;
		}

		std::vector<uint8_t> bits(param.filter_len() / 8 + ((param.filter_len() % 8) ? 1 : 0));
		// the bits and their crc32 (filter.update_crc32(), make_bloom.cpp:395) both come from the device
		uint32_t bits_crc = 0;
		cuda_check(kwg_bloom_finalize_crc(builder, param.log_2_filter_len, param.num_hash, bits.data(), &bits_crc),
			__FILE__ ":make_bloom_filter: kwg_bloom_finalize failed");
		kwg_bloom_destroy(builder);
		builder = NULL;

		const std::string output_file = bloom_dir + "/" + accession_to_str(acc) + ".bloom";
		std::ofstream fout(output_file.c_str(), std::ios::binary);
		if (!fout) throw __FILE__ ":main: Unable to open Bloom filter file for writing";
		write_bloom_file(fout, param, info, bits.data(), bits_crc);
		fout.close();
	}
	catch (const char* error) {
		if (builder) kwg_bloom_destroy(builder);
		progress.error = error;
		return STATUS_BLOOM_FAIL;
	}
	catch (const std::exception& error) {
		if (builder) kwg_bloom_destroy(builder);
		progress.error = error.what();
		return STATUS_BLOOM_FAIL;
	}
	catch (...) {
		if (builder) kwg_bloom_destroy(builder);
		return STATUS_BLOOM_FAIL;
	}
	return STATUS_BLOOM_SUCCESS;
}

unsigned char make_bloom_filter(const SraAccession& acc, const FilterInfo& info, BloomParam& param, BloomProgress& progress,
	const std::string& bloom_dir, const MaestroOptions& opt, bool /*force_unaligned*/)
{
	ReadSource* src = NULL;
	try {
		const std::string accession = accession_to_str(acc);
		const uint64_t num_bp = number_of_bases(accession);
		src = open_read_collection(accession);
		const unsigned char status = make_bloom_filter(*src, num_bp, acc, info, param, progress, bloom_dir, opt);
		delete src;
		return status;
	}
	catch (const char* error) {
		delete src;
		progress.error = error;
		return STATUS_BLOOM_FAIL;
	}
	catch (...) {
		delete src;
		return STATUS_BLOOM_FAIL;
	}
}

// page-locked when the CUDA library can provide it, plain memory otherwise (the calls accept any host pointer)
class PinnedBuffer {
public:
	explicit PinnedBuffer(size_t n) : p(NULL), pinned(true)
	{
		p = static_cast<uint8_t*>(kwg_host_alloc(n ? n : 1));
		if (!p) { pinned = false; p = static_cast<uint8_t*>(malloc(n ? n : 1)); }
		if (!p) throw __FILE__ ":build_db: Unable to allocate a staging buffer";
	}
	~PinnedBuffer() { if (pinned) kwg_host_free(p); else free(p); }
	uint8_t* data() { return p; }
private:
	PinnedBuffer(const PinnedBuffer&);
	PinnedBuffer& operator=(const PinnedBuffer&);
	uint8_t* p;
	bool pinned;
};

// ================================================================ transposition
// Mirrors build_db() (reference build_db.cpp:24-456): same validation, same chunking of the slice
// axis (4,194,304 slices per chunk, build_db.cpp:243), same CRC bookkeeping, same file layout.
bool build_db(const std::string& filename, const BloomParam& param, const std::deque<std::string>& bloom_files)
{
	std::vector<std::ifstream*> fin;
	try {
		const size_t num_filter = bloom_files.size();
		if (num_filter == 0) throw __FILE__ ":build_db: Empty Bloom filter inventory file";

		fin.assign(num_filter, NULL);
		std::vector<BloomFileHeader> heads(num_filter);
		for (size_t i = 0; i < num_filter; ++i) {
			fin[i] = new std::ifstream(bloom_files[i].c_str(), std::ios::binary);
			if (!*fin[i]) throw __FILE__ ":build_db: Unable to open Bloom filter file";
		}
		for (size_t i = 0; i < num_filter; ++i) {
			try { read_bloom_header(*fin[i], heads[i]); }
			catch (...) { throw __FILE__ ":build_db: Incomplete Bloom filter"; }
			if (param != heads[i].param) throw __FILE__ ":build_db: Inconsistent Bloom parameters";
		}

		DBFileHeader header;
		header.crc32 = 0;
		header.kmer_len = param.kmer_len;
		header.num_hash = param.num_hash;
		header.log_2_filter_len = param.log_2_filter_len;
		header.num_filter = (uint32_t)num_filter;
		header.hash_func = param.hash_func;
		header.compression = 0;   // NO_COMPRESSION (kwage.h:16-20, build_db.cpp:197-199)

		std::ofstream fout(filename.c_str(), std::ios::binary);
		if (!fout) throw __FILE__ ":build_db: Unable to open output file for writing";
		binary_write(fout, header);

		const size_t filter_len = param.filter_len();
		const size_t max_buffer_slice = size_t(524288) * 8;
		const size_t bytes_per_slice = num_filter / 8 + ((num_filter % 8) ? 1 : 0);
		// staging buffers in page-locked memory: the transposition's copies run at PCIe rate and overlap its kernels
		PinnedBuffer src(num_filter * (std::min(max_buffer_slice, filter_len) / 8 + 1));
		PinnedBuffer dest(std::min(max_buffer_slice, filter_len) * bytes_per_slice);
		std::vector<const uint8_t*> chunk_ptr(num_filter);
		std::vector<uint32_t> running_crc(num_filter, 0);

		for (size_t i = 0; i < filter_len; i += max_buffer_slice) {
			const size_t num_buffer_slice = std::min(max_buffer_slice, filter_len - i);
			const size_t chunk_bytes = num_buffer_slice / 8 + ((num_buffer_slice % 8) ? 1 : 0);
			// the running crc32 values (per source filter, build_db.cpp:281-282; of the slices, 307) are advanced on the
			// device next to the transposition when the messages are made of whole 32-bit words, else by zlib here
			const bool dev_filter_crc = num_buffer_slice % 32 == 0;
			const bool dev_dest_crc = dev_filter_crc && num_filter % 32 == 0;
			for (size_t j = 0; j < num_filter; ++j) {
				uint8_t* p = src.data() + j * chunk_bytes;
				fin[j]->read(reinterpret_cast<char*>(p), (std::streamsize)chunk_bytes);
				if (!*fin[j]) throw __FILE__ ":build_db: Error reading filter bytes";
				if (!dev_filter_crc) running_crc[j] = crc32_bytes(running_crc[j], p, chunk_bytes);
				chunk_ptr[j] = p;
			}
			// the bitwise transposition at the heart of the bit-sliced approach (build_db.cpp:288-303)
			const size_t curr_dest_len = num_buffer_slice * bytes_per_slice;
			if (dev_filter_crc) {
				cuda_check(kwg_transpose_crc(g_build_db_device, chunk_ptr.data(), (uint32_t)num_filter, num_buffer_slice, dest.data(),
					running_crc.data(), dev_dest_crc ? &header.crc32 : NULL), __FILE__ ":build_db: kwg_transpose failed");
			} else {
				cuda_check(kwg_transpose(g_build_db_device, chunk_ptr.data(), (uint32_t)num_filter, num_buffer_slice, dest.data()),
					__FILE__ ":build_db: kwg_transpose failed");
			}
			if (!dev_dest_crc) header.crc32 = crc32_bytes(header.crc32, dest.data(), curr_dest_len);
			fout.write(reinterpret_cast<const char*>(dest.data()), (std::streamsize)curr_dest_len);
			if (!fout) throw __FILE__ ":build_db: Unable to write transpose buffer to disk";
		}
		for (size_t i = 0; i < num_filter; ++i) { delete fin[i]; fin[i] = NULL; }

		for (size_t i = 0; i < num_filter; ++i)
			if (heads[i].crc32 != running_crc[i]) throw __FILE__ ":build_db: One or more invalid Bloom filter CRC32 values";

		// metadata: a table of absolute offsets, then the FilterInfo records (build_db.cpp:371-429)
		std::vector<uint64_t> info_loc(num_filter, 0);
		header.info_start = (uint64_t)fout.tellp();
		fout.write(reinterpret_cast<const char*>(info_loc.data()), (std::streamsize)(num_filter * sizeof(uint64_t)));
		for (size_t i = 0; i < num_filter; ++i) {
			info_loc[i] = (uint64_t)fout.tellp();
			binary_write(fout, heads[i].info);
		}
		fout.seekp((std::streamoff)header.info_start);
		fout.write(reinterpret_cast<const char*>(info_loc.data()), (std::streamsize)(num_filter * sizeof(uint64_t)));
		fout.seekp(0);
		binary_write(fout, header);
		if (!fout) throw __FILE__ ":build_db: Error writing database file header (final)";
		fout.close();
	}
	catch (...) {
		for (size_t i = 0; i < fin.size(); ++i) delete fin[i];
		return false;
	}
	return true;
}

// ================================================================ search
static void read_db_header(std::ifstream& fin, DBFileHeader& h)
{
	if (!fin) throw __FILE__ ":main: I/O error";
	binary_read(fin, h);
	if (!fin) throw __FILE__ ":main: Unable to read header";
	if (h.magic != KWAGE_MAGIC_NUMBER) throw __FILE__ ":main: Not a KWAGE database file";
}

SubjectDatabase::SubjectDatabase(const std::string& filename, int device) : db(NULL)
{
	open_files(std::vector<std::string>(1, filename), device);
}

SubjectDatabase::SubjectDatabase(const std::vector<std::string>& filenames, int device) : db(NULL)
{
	if (filenames.empty()) throw __FILE__ ":main: No database file";
	open_files(filenames, device);
}

bool SubjectDatabase::compatible(const std::string& file_a, const std::string& file_b)
{
	try {
		std::ifstream fa(file_a.c_str(), std::ios::binary), fb(file_b.c_str(), std::ios::binary);
		DBFileHeader a, b;
		read_db_header(fa, a);
		read_db_header(fb, b);
		return a.kmer_len == b.kmer_len && a.num_hash == b.num_hash && a.log_2_filter_len == b.log_2_filter_len &&
		       a.hash_func == b.hash_func && a.compression == b.compression;
	}
	catch (...) { return false; }
}

uint64_t SubjectDatabase::slab_bytes(const std::string& filename)
{
	std::ifstream f(filename.c_str(), std::ios::binary);
	DBFileHeader h;
	read_db_header(f, h);
	return (uint64_t(1) << h.log_2_filter_len) * (h.num_filter / 8 + 1);
}

void SubjectDatabase::open_files(const std::vector<std::string>& filenames, int device)
{
	try {
		uint64_t total = 0;
		for (size_t f = 0; f < filenames.size(); ++f) {
			Part p;
			p.fin = new std::ifstream(filenames[f].c_str(), std::ios::binary);
			parts.push_back(p);
			Part& q = parts.back();
			read_db_header(*q.fin, q.hdr);
			q.col_begin = (uint32_t)total;
			total += q.hdr.num_filter;
			const DBFileHeader& h0 = parts[0].hdr;
			if (q.hdr.kmer_len != h0.kmer_len || q.hdr.num_hash != h0.num_hash || q.hdr.log_2_filter_len != h0.log_2_filter_len ||
			    q.hdr.hash_func != h0.hash_func)
				throw __FILE__ ":main: Database files with different Bloom parameters cannot share a slab";
		}
		if (total == 0 || total > 0xFFFFFFFFull) throw __FILE__ ":main: Bad number of filters";
		hdr = parts[0].hdr;
		hdr.num_filter = (uint32_t)total;
		cuda_check(kwg_db_alloc(&db, device, hdr.kmer_len, hdr.num_hash, hdr.log_2_filter_len, hdr.num_filter, 0, hdr.num_filter),
			__FILE__ ":search: kwg_db_alloc failed");
		// stream every file's slice region into its columns in bounded pieces (a region can exceed host memory)
		const uint64_t n_rows = 1ULL << hdr.log_2_filter_len;
		for (size_t f = 0; f < parts.size(); ++f) {
			Part& q = parts[f];
			const size_t slice_size = q.hdr.num_filter / 8 + ((q.hdr.num_filter % 8) ? 1 : 0);
			const uint64_t piece_rows = std::max<uint64_t>(1, (uint64_t(256) << 20) / std::max<size_t>(slice_size, 1));
			std::vector<uint8_t> buf((size_t)(std::min(piece_rows, n_rows) * slice_size));
			for (uint64_t r = 0; r < n_rows; r += piece_rows) {
				const uint64_t rows = std::min(piece_rows, n_rows - r);
				q.fin->read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(rows * slice_size));
				if (!*q.fin) throw __FILE__ ":search: Error reading slice from file (1)";
				const int rc = (parts.size() == 1) ? kwg_db_upload_rows(db, r, rows, buf.data())
				                                   : kwg_db_upload_columns(db, q.col_begin, q.hdr.num_filter, r, rows, buf.data());
				cuda_check(rc, __FILE__ ":search: kwg_db_upload failed");
			}
		}
	}
	catch (...) {
		if (db) { kwg_db_unload(db); db = NULL; }
		for (size_t f = 0; f < parts.size(); ++f) delete parts[f].fin;
		parts.clear();
		throw;
	}
}

SubjectDatabase::~SubjectDatabase()
{
	if (db) kwg_db_unload(db);
	for (size_t f = 0; f < parts.size(); ++f) delete parts[f].fin;
}

FilterInfo SubjectDatabase::filter_info(uint32_t filter)
{
	size_t f = 0;
	while (f + 1 < parts.size() && filter >= parts[f + 1].col_begin) ++f;      // the file that owns this column
	Part& q = parts[f];
	const uint32_t local = filter - q.col_begin;
	q.fin->clear();
	q.fin->seekg((std::streamoff)(q.hdr.info_start + (uint64_t)local * sizeof(uint64_t)));
	uint64_t loc = 0;
	q.fin->read(reinterpret_cast<char*>(&loc), sizeof(loc));
	q.fin->seekg((std::streamoff)loc);
	FilterInfo info;
	binary_read(*q.fin, info);
	return info;
}

bool SubjectDatabase::search(std::unordered_map<size_t, std::deque<MatchResult> >& results, const std::vector<std::string>& queries,
	const std::vector<size_t>& query_ids, const SearchOptions& opt)
{
	if (queries.empty()) return false;
	if (queries.size() > 0xFFFFFFFFull) throw __FILE__ ":search: more than 2^32 queries in one call (split the query set)";
	std::vector<const char*> ptrs(queries.size());
	std::vector<uint64_t> lens(queries.size());
	for (size_t i = 0; i < queries.size(); ++i) { ptrs[i] = queries[i].data(); lens[i] = queries[i].size(); }
	std::vector<uint32_t> n_kmers(queries.size(), 0);
	kwg_hit_t* hits = NULL;
	uint64_t n_hits = 0;
	cuda_check(kwg_search_ptrs(db, ptrs.data(), lens.data(), (uint32_t)queries.size(), opt.threshold, n_kmers.data(), &hits, &n_hits),
		__FILE__ ":search: kwg_search failed");
	std::unordered_map<uint32_t, FilterInfo> info_cache;
	for (uint64_t i = 0; i < n_hits; ++i) {
		const kwg_hit_t& h = hits[i];
		std::unordered_map<uint32_t, FilterInfo>::iterator it = info_cache.find(h.filter);
		if (it == info_cache.end()) it = info_cache.insert(std::make_pair(h.filter, filter_info(h.filter))).first;
		results[query_ids[h.query]].push_back(MatchResult(h.num_match, n_kmers[h.query], it->second));
	}
	kwg_free_hits(hits);
	return n_hits > 0;
}

bool search(std::unordered_map<size_t, std::deque<MatchResult> >& results, SubjectDatabase& subject, const std::string& query,
	const size_t& query_id, const SearchOptions& opt)
{
	return subject.search(results, std::vector<std::string>(1, query), std::vector<size_t>(1, query_id), opt);
}

// ================================================================ output (reference output.h:35-112)
void write_csv_header(std::ostream& out) { out << "query,num_kmers,num_kmers_found,percent_kmers_found,sample_metadata\n"; }

void write_csv(std::ostream& out, const std::string& query, const std::deque<MatchResult>& r)
{
	for (std::deque<MatchResult>::const_iterator i = r.begin(); i != r.end(); ++i) {
		const float norm = i->num_query_kmer ? 1.0f / i->num_query_kmer : 0.0f;
		out << '"' << query << "\"," << i->num_query_kmer << ',' << i->num_kmers_found << ',' << (100.0f * i->num_kmers_found) * norm
		    << ",\"" << i->subject_info.csv_string() << '"' << std::endl;
	}
}

void write_json_header(std::ostream& out, bool multiple) { if (multiple) out << '['; }

void write_json(std::ostream& out, const std::string& query, bool multiple, bool first, const float& threshold, const std::deque<MatchResult>& r)
{
	const std::string prefix = multiple ? "\t" : "";
	out << ((multiple && !first) ? "," : "") << '\n' << prefix << "{\n" << prefix << "\t\"query\": \"" << query << "\",\n" << prefix
	    << "\t\"threshold\": " << std::showpoint << std::setprecision(1) << std::fixed << threshold << ",\n" << prefix << "\t\"results\": [";
	for (std::deque<MatchResult>::const_iterator i = r.begin(); i != r.end(); ++i) {
		const float norm = i->num_query_kmer ? 1.0f / i->num_query_kmer : 0.0f;
		out << ((i != r.begin()) ? "," : "") << "\n" << prefix << "\t\t{\n" << prefix << "\t\t\t\"percent_kmers_found\": "
		    << (100.0 * i->num_kmers_found) * norm << ",\n" << prefix << "\t\t\t\"num_kmers\": " << i->num_query_kmer << ",\n" << prefix
		    << "\t\t\t\"num_kmers_found\": " << i->num_kmers_found << ",\n" << prefix << "\t\t\t\"sample_metadata\": {\n"
		    << i->subject_info.json_string(prefix + "\t\t\t\t") << "\n" << prefix << "\t\t\t}\n" << prefix << "\t\t}";
	}
	if (!r.empty()) out << "\n" << prefix << '\t';
	out << "]\n" << prefix << "}";
}

void write_json_footer(std::ostream& out, bool multiple) { if (multiple) out << "\n]\n"; }

} // namespace kwage
