// The three stage functions of the hot path on top of the C ABI.  Host code keeps what the
// reference keeps on the host (streaming, parameter choice, CRCs, files, metadata); everything
// the reference computes in its inner loops is a libkwage_cuda call.
#include "kwage_host.h"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <mutex>
#include <sstream>
#include <fcntl.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
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

// FASTA / FASTQ, optionally gzip (zlib reads plain files too): the SequenceIterator role (parse_sequence.cpp:72-262).
// zlib inflate is what bounds FASTQ.gz -> .bloom, so it gets a thread of its own: the file is inflated block by block
// into a small ring while the caller (BatchPipeline's parser thread) cuts the previous block into lines and packs them.
class GzSequenceReads : public ReadSource {
public:
	explicit GzSequenceReads(const std::string& path) : gz(gzopen(path.c_str(), "rb")), fill(0), drain(0), stop(false), cur(NULL), pos(0),
		peeked(false), fastq(false), first(true)
	{
		if (!gz) throw __FILE__ ":open_read_collection: Unable to open sequence file";
		gzbuffer(gz, 1 << 20);
		for (size_t i = 0; i < N_BLOCKS; ++i) blocks[i].data.resize(BLOCK_BYTES);
		inflater = std::thread(&GzSequenceReads::inflate_loop, this);
	}
	~GzSequenceReads()
	{
		{ std::lock_guard<std::mutex> l(mu); stop = true; }
		cv.notify_all();
		if (inflater.joinable()) inflater.join();
		if (gz) gzclose(gz);
	}
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
	static const size_t BLOCK_BYTES = size_t(4) << 20, N_BLOCKS = 3;
	struct Block { std::vector<char> data; size_t len; bool last; Block() : len(0), last(false) {} };

	void inflate_loop()
	{
		for (;;) {
			Block* b;
			{
				std::unique_lock<std::mutex> l(mu);
				cv.wait(l, [&] { return stop || fill - drain < N_BLOCKS; });
				if (stop) return;
				b = &blocks[fill % N_BLOCKS];
			}
			const int n = gzread(gz, b->data.data(), (unsigned)BLOCK_BYTES);
			b->len = n > 0 ? (size_t)n : 0;
			b->last = n < (int)BLOCK_BYTES;          // short read: end of file (or an error: the stream ends here, like gzgets)
			{
				std::lock_guard<std::mutex> l(mu);
				++fill;
			}
			cv.notify_all();
			if (b->last) return;
		}
	}
	// the next inflated block, or false at the end of the file
	bool next_block()
	{
		if (cur) {
			const bool was_last = cur->last;
			{ std::lock_guard<std::mutex> l(mu); ++drain; }
			cv.notify_all();
			cur = NULL;
			if (was_last) { at_end = true; return false; }
		}
		if (at_end) return false;
		std::unique_lock<std::mutex> l(mu);
		cv.wait(l, [&] { return fill > drain; });
		cur = &blocks[drain % N_BLOCKS];
		pos = 0;
		return true;
	}
	// one line without its end-of-line characters; false when nothing is left (a last line without '\n' still counts)
	bool next_line(std::string& line)
	{
		if (peeked) { line = pending; peeked = false; return true; }
		line.clear();
		bool any = false;
		for (;;) {
			if (!cur || pos >= cur->len) {
				if (!next_block()) break;
				if (cur->len == 0) continue;
			}
			any = true;
			const char* p = cur->data.data() + pos;
			const size_t left = cur->len - pos;
			const char* nl = (const char*)std::memchr(p, '\n', left);
			if (nl) {
				line.append(p, (size_t)(nl - p));
				pos += (size_t)(nl - p) + 1;
				break;
			}
			line.append(p, left);
			pos = cur->len;
		}
		while (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
		return any;
	}
	gzFile gz;
	Block blocks[N_BLOCKS];
	size_t fill, drain;                      // blocks inflated / handed back (under mu)
	bool stop;
	std::mutex mu;
	std::condition_variable cv;
	std::thread inflater;
	Block* cur;
	size_t pos;
	bool at_end = false;
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

// page-locked when the CUDA library can provide it, plain memory otherwise (the calls accept any host pointer)
class PinnedBuffer {
public:
	explicit PinnedBuffer(size_t n) : p(NULL), pinned(true)
	{
		p = static_cast<uint8_t*>(kwg_host_alloc(n ? n : 1));
		if (!p) { pinned = false; p = static_cast<uint8_t*>(malloc(n ? n : 1)); }
		if (!p) throw __FILE__ ":PinnedBuffer: Unable to allocate a staging buffer";
	}
	~PinnedBuffer() { if (pinned) kwg_host_free(p); else free(p); }
	uint8_t* data() { return p; }
private:
	PinnedBuffer(const PinnedBuffer&);
	PinnedBuffer& operator=(const PinnedBuffer&);
	uint8_t* p;
	bool pinned;
};

// ================================================================ read ingestion
// 2-bit packing on the parser thread: the bases of a fragment are appended to the batch at base position `cursor` in
// the format of kwg_bloom_add_packed (NCBI 2na: four bases per byte, the first in bits 7..6; A=0 C=1 G=2 T=3, lower case
// folded, word.h:19,80-101) and every other byte sets its bit in the not-a-base mask (LSB first), which breaks the
// k-mers there exactly like word.h:98-100.  `mask` must be zero where nothing has been written yet; returns true if the
// fragment held such a byte.
namespace {
struct PackTable {
	uint8_t code[256];
	PackTable()
	{
		std::memset(code, 0x80, sizeof(code));
		code[(unsigned char)'A'] = code[(unsigned char)'a'] = 0; code[(unsigned char)'C'] = code[(unsigned char)'c'] = 1;
		code[(unsigned char)'G'] = code[(unsigned char)'g'] = 2; code[(unsigned char)'T'] = code[(unsigned char)'t'] = 3;
	}
};
const PackTable g_pack;
}

bool pack_2na(uint8_t* packed, uint8_t* mask, uint64_t cursor, const char* bases, size_t n)
{
	const unsigned char* s = reinterpret_cast<const unsigned char*>(bases);
	unsigned any = 0;
	size_t i = 0;
	uint64_t pos = cursor;
	// up to the next byte boundary of the 2na stream (the byte already holds the previous fragment's last bases)
	for (; i < n && (pos & 3u); ++i, ++pos) {
		const unsigned c = g_pack.code[s[i]];
		if (c & 0x80u) { mask[pos >> 3] |= (uint8_t)(1u << (pos & 7u)); any = 1; }
		packed[pos >> 2] |= (uint8_t)((c & 3u) << (6 - 2 * (pos & 3u)));
	}
	for (; i + 4 <= n; i += 4, pos += 4) {
		const unsigned c0 = g_pack.code[s[i]], c1 = g_pack.code[s[i + 1]], c2 = g_pack.code[s[i + 2]], c3 = g_pack.code[s[i + 3]];
		packed[pos >> 2] = (uint8_t)(((c0 & 3u) << 6) | ((c1 & 3u) << 4) | ((c2 & 3u) << 2) | (c3 & 3u));
		if ((c0 | c1 | c2 | c3) & 0x80u) {
			const unsigned m = (c0 >> 7) | ((c1 >> 7) << 1) | ((c2 >> 7) << 2) | ((c3 >> 7) << 3);
			mask[pos >> 3] |= (uint8_t)(m << (pos & 7u));          // (pos % 4 == 0: the four flags stay inside the byte)
			any = 1;
		}
	}
	if (i < n) {
		packed[pos >> 2] = 0;
		for (; i < n; ++i, ++pos) {
			const unsigned c = g_pack.code[s[i]];
			if (c & 0x80u) { mask[pos >> 3] |= (uint8_t)(1u << (pos & 7u)); any = 1; }
			packed[pos >> 2] |= (uint8_t)((c & 3u) << (6 - 2 * (pos & 3u)));
		}
	}
	return any != 0;
}

// The reference's fragment loop (make_bloom.cpp:194-300) reads one fragment, counts its k-mers, reads the next.  Here a
// parser thread reads and packs batch n + 1 while the device works on batch n: a ring of page-locked batches, the
// parser fills, the caller's thread feeds them to kwg_bloom_add_packed in stream order.
namespace {

struct ReadBatch {
	PinnedBuffer packed, mask;
	std::vector<uint64_t> offsets;
	uint64_t n_bases, n_fragments;           // (a fragment longer than a batch arrives as several reads)
	bool any_bad, last;
	std::string error;                       // what the parser threw, rethrown on the feeding thread
	explicit ReadBatch(size_t cap_bases) : packed(cap_bases / 4 + 64), mask(cap_bases / 8 + 64), n_bases(0), n_fragments(0), any_bad(false), last(false) {}
};

class BatchPipeline {
public:
	BatchPipeline(ReadSource& src, size_t batch_bases, size_t n_batches, size_t kmer_len) : overlap(kmer_len ? kmer_len - 1 : 0), reads(src), cap(batch_bases),
		max_batches(n_batches), stop(false)
	{
		// (page-locking memory costs ~0.5 ms per MiB: batches are created when the parser first needs them, so a small
		// accession pins one batch and not the whole ring)
		worker = std::thread(&BatchPipeline::run, this);
	}
	~BatchPipeline()
	{
		{ std::lock_guard<std::mutex> l(mu); stop = true; }
		cv.notify_all();
		if (worker.joinable()) worker.join();
		for (size_t i = 0; i < pool.size(); ++i) delete pool[i];
	}
	ReadBatch* next()                        // blocks until the parser has a batch; the last one has last == true
	{
		std::unique_lock<std::mutex> l(mu);
		cv.wait(l, [this] { return failed || !full_q.empty(); });
		if (full_q.empty()) throw __FILE__ ":make_bloom_filter: Unable to allocate a read batch";
		ReadBatch* b = full_q.front();
		full_q.pop_front();
		return b;
	}
	void release(ReadBatch* b)
	{
		{ std::lock_guard<std::mutex> l(mu); free_q.push_back(b); }
		cv.notify_all();
	}
private:
	void run()
	{
		std::string fragment;
		bool pending = false;                    // `fragment` (from pending_off on) still has to go out
		size_t pending_off = 0;
		bool more = true;
		while (more || pending) {
			ReadBatch* b = NULL;
			{
				std::unique_lock<std::mutex> l(mu);
				if (free_q.empty() && pool.size() < max_batches) {
					l.unlock();
					ReadBatch* fresh = NULL;
					try { fresh = new ReadBatch(cap + (1u << 16)); } catch (...) { fresh = NULL; }
					l.lock();
					if (fresh) { pool.push_back(fresh); free_q.push_back(fresh); }
					else if (pool.empty()) { stop = true; failed = true; l.unlock(); cv.notify_all(); return; }
				}
				cv.wait(l, [this] { return stop || !free_q.empty(); });
				if (stop) return;
				b = free_q.front();
				free_q.pop_front();
			}
			b->offsets.assign(1, 0);
			b->n_bases = 0; b->n_fragments = 0; b->any_bad = false; b->last = false; b->error.clear();
			std::memset(b->mask.data(), 0, cap / 8 + 64);
			try {
				while (b->n_bases < cap) {
					if (!pending) {
						more = reads.next_fragment(fragment);
						if (!more) break;
						pending = true;
						pending_off = 0;
					}
					const size_t remaining = fragment.size() - pending_off, room = cap - b->n_bases;
					size_t take = remaining;
					if (remaining > room) {
						if (b->n_bases) break;           // goes out with the next batch
						// a fragment longer than a whole batch: pieces that overlap by k - 1 bases hold every one of its
						// k-mers exactly once and in order (a window lies in the first piece that contains it whole)
						take = cap;
					}
					append(b, fragment.data() + pending_off, take);
					if (take == remaining) { pending = false; ++b->n_fragments; }
					else pending_off += take - (overlap < take ? overlap : 0);
				}
			}
			catch (const char* e) { b->error = e; more = false; pending = false; }
			catch (const std::exception& e) { b->error = e.what(); more = false; pending = false; }
			catch (...) { b->error = "unknown error while reading"; more = false; pending = false; }
			b->last = !more && !pending;
			{ std::lock_guard<std::mutex> l(mu); full_q.push_back(b); }
			cv.notify_all();
		}
	}
	void append(ReadBatch* b, const char* bases, size_t n)
	{
		if ((b->n_bases & 3u) == 0) b->packed.data()[b->n_bases >> 2] = 0;
		if (n) b->any_bad |= pack_2na(b->packed.data(), b->mask.data(), b->n_bases, bases, n);
		b->n_bases += n;
		b->offsets.push_back(b->n_bases);
	}
	size_t overlap;                              // kmer_len - 1
	ReadSource& reads;
	size_t cap, max_batches;
	bool failed = false;
	std::vector<ReadBatch*> pool;
	std::deque<ReadBatch*> free_q, full_q;
	std::mutex mu;
	std::condition_variable cv;
	std::thread worker;
	bool stop;
};

} // namespace

// bases per batch: 64 Mi (16 MiB packed); KWAGE_BATCH_BASES overrides it (tests drive the batch boundaries with it)
static size_t batch_bases()
{
	if (const char* e = std::getenv("KWAGE_BATCH_BASES")) {
		const size_t v = (size_t)std::strtoull(e, NULL, 10);
		if (v >= 64) return v;
	}
	return size_t(64) << 20;
}

// ================================================================ construction
// Mirrors make_bloom_filter() (reference make_bloom.cpp:76-504).  Fragments are batched so that one libkwage_cuda call
// carries ~64 Mi bases (16 MiB packed); stream order is preserved (it matters for the counting filter).
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

		{
			// (an accession known to be small pins a batch of its own size, not 64 Mi bases)
			const size_t bb = num_bp ? std::min<size_t>(batch_bases(), std::max<size_t>(size_t(4) << 20, (size_t)num_bp + (1u << 16))) : batch_bases();
			BatchPipeline pipe(reads, bb, 3, opt.kmer_len);
			bool last = false;
			while (!last) {
				ReadBatch* rb = pipe.next();
				if (!rb->error.empty()) {
					progress.error = rb->error;
					pipe.release(rb);
					kwg_bloom_destroy(builder);
					return STATUS_BLOOM_FAIL;
				}
				last = rb->last;
				const size_t n_reads = rb->offsets.size() - 1;
				progress.num_bp += rb->n_bases;
				progress.curr_read += rb->n_fragments;
				progress.num_read += rb->n_fragments;
				if (n_reads) {
					// The reference tests the limit after every fragment (make_bloom.cpp:208,246,288).  A batch can only
					// cross it if the k-mers so far plus the bases of the batch exceed it: then the counting state is
					// saved first, and if the limit is crossed the batch is replayed in prefixes until the fragment
					// that crossed it is known -- num_kmer, num_bp and the fragment counters at the abort are the reference's.
					const size_t bp_before = progress.num_bp - rb->n_bases, reads_before = progress.curr_read - rb->n_fragments;
					const bool at_risk = progress.num_kmer + rb->n_bases > max_num_kmer;
					if (at_risk) cuda_check(kwg_bloom_checkpoint(builder), __FILE__ ":make_bloom_filter: kwg_bloom_checkpoint failed");
					const uint8_t* mask = rb->any_bad ? rb->mask.data() : NULL;
					int rc = kwg_bloom_add_packed(builder, rb->packed.data(), mask, rb->offsets.data(), n_reads);
					uint64_t n = 0;
					if (rc == KWG_OK) rc = kwg_bloom_num_valid(builder, &n);
					if (rc == KWG_OK && at_risk && max_num_kmer < n) {
						// smallest prefix of the batch that crosses the limit: hi reads do, lo reads do not
						size_t lo = 0, hi = n_reads;
						uint64_t n_hi = n;
						while (rc == KWG_OK && hi - lo > 1) {
							const size_t mid = lo + (hi - lo) / 2;
							uint64_t n_mid = 0;
							rc = kwg_bloom_rollback(builder);
							if (rc == KWG_OK) rc = kwg_bloom_add_packed(builder, rb->packed.data(), mask, rb->offsets.data(), mid);
							if (rc == KWG_OK) rc = kwg_bloom_num_valid(builder, &n_mid);
							if (max_num_kmer < n_mid) { hi = mid; n_hi = n_mid; } else lo = mid;
						}
						if (rc == KWG_OK) {
							progress.num_kmer = n_hi;
							progress.num_bp = bp_before + (rb->offsets[hi] - rb->offsets[0]);
							progress.curr_read = reads_before + (hi - 1);      // (the read that crossed the limit is not counted as done)
							progress.curr_fragment = 1;
							pipe.release(rb);
							kwg_bloom_destroy(builder);
							return STATUS_BLOOM_INVALID;
						}
					}
					pipe.release(rb);                 // (the calls have copied the batch: the parser may refill it)
					cuda_check(rc, __FILE__ ":make_bloom_filter: kwg_bloom_add_packed failed");
					progress.num_kmer = n;
					if (max_num_kmer < progress.num_kmer) {       // (not reached: a batch that can cross the limit is at risk)
						kwg_bloom_destroy(builder);
						return STATUS_BLOOM_INVALID;
					}
				} else {
					pipe.release(rb);
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

// ================================================================ merging database files
// Mirrors merge_database_files() (reference merge_db.cpp:278-820): file 2's filters are appended to file 1's columns; if
// file 1 fills up (max_num_filters) the rest of file 2's columns become the new file 2.  Same checks, same temp files
// and renames, same header / crc32 / FilterInfo layout; the per-slice bit moving (merge_db.cpp:533-566) is kwg_merge_slices.
// Returns {filters in the file that can still take more, its name} like the reference ({0, ""} if none).
std::pair<size_t, std::string> merge_database_files(const std::string& file_1, const std::string& file_2, const size_t& max_num_filters, int device)
{
	const std::string tmp_suffix = ".tmp";
	std::pair<size_t, std::string> ret(size_t(0), "");
	std::ifstream fin_1(file_1.c_str(), std::ios::binary);
	if (!fin_1) throw __FILE__ ":merge_database_files: Unable to open database file 1";
	std::ifstream fin_2(file_2.c_str(), std::ios::binary);
	if (!fin_2) throw __FILE__ ":merge_database_files: Unable to open database file 2";
	DBFileHeader src_1, src_2;
	binary_read(fin_1, src_1);
	if (!fin_1) throw __FILE__ ":merge_database_files: Error reading header 1";
	binary_read(fin_2, src_2);
	if (!fin_2) throw __FILE__ ":merge_database_files: Error reading header 2";
	if (src_1.log_2_filter_len != src_2.log_2_filter_len || src_1.num_hash != src_2.num_hash || src_1.kmer_len != src_2.kmer_len ||
	    src_1.hash_func != src_2.hash_func)
		throw __FILE__ ":merge_database_files: Incompatible database files";
	if (src_1.compression != 0 || src_2.compression != 0)
		throw __FILE__ ":merge_database_files: Compressed database files are not currently supported";
	if (src_1.num_filter >= max_num_filters) throw __FILE__ ":merge_database_files: Database file 1 has more than expected filters";
	if (src_2.num_filter >= max_num_filters) throw __FILE__ ":merge_database_files: Database file 2 has more than expected filters";
	if (src_1.num_filter == 0 || src_2.num_filter == 0) throw __FILE__ ":merge_database_files: A database file without filters";

	const bool has_remainder = (size_t(src_1.num_filter) + src_2.num_filter) > max_num_filters;
	const std::string dst_file_1 = file_1 + tmp_suffix;
	const std::string dst_file_2 = has_remainder ? file_2 + tmp_suffix : "";
	if (file_exists(dst_file_1)) throw __FILE__ ":merge_database_files: Temp database file 1 already exists";
	if (has_remainder && file_exists(dst_file_2)) throw __FILE__ ":merge_database_files: Temp database file 2 already exists";
	std::ofstream fout_1(dst_file_1.c_str(), std::ios::binary), fout_2;
	if (!fout_1) throw __FILE__ ":merge_database_files: Unable to open new_file_1";
	if (has_remainder) {
		fout_2.open(dst_file_2.c_str(), std::ios::binary);
		if (!fout_2) throw __FILE__ ":merge_database_files: Unable to open new_file_2";
	}

	DBFileHeader dst_1 = src_1, dst_2 = src_2;
	dst_1.crc32 = dst_2.crc32 = 0;
	dst_1.info_start = dst_2.info_start = 0;
	if (has_remainder) {
		dst_1.num_filter = (uint32_t)max_num_filters;
		dst_2.num_filter = (uint32_t)((size_t(src_1.num_filter) + src_2.num_filter) - max_num_filters);
		ret = std::make_pair(size_t(dst_2.num_filter), file_2);
	}
	else {
		dst_1.num_filter = src_1.num_filter + src_2.num_filter;
		dst_2.num_filter = 0;
		if (dst_1.num_filter < max_num_filters) ret = std::make_pair(size_t(dst_1.num_filter), file_1);
	}
	binary_write(fout_1, dst_1);
	if (!fout_1) throw __FILE__ ":merge_database_files: Error writing database file 1 header placeholder";
	if (has_remainder) {
		binary_write(fout_2, dst_2);
		if (!fout_2) throw __FILE__ ":merge_database_files: Error writing database file 1 header placeholder";
	}

	const size_t num_bitslice = src_1.filter_len();
	const size_t bps_src_1 = src_1.num_filter / 8 + ((src_1.num_filter % 8) ? 1 : 0), bps_src_2 = src_2.num_filter / 8 + ((src_2.num_filter % 8) ? 1 : 0);
	const size_t bps_dst_1 = dst_1.num_filter / 8 + ((dst_1.num_filter % 8) ? 1 : 0), bps_dst_2 = dst_2.num_filter / 8 + ((dst_2.num_filter % 8) ? 1 : 0);
	// chunks of slices sized for ~32 MiB per buffer (the reference moves 1024 slices at a time)
	const size_t per_chunk = std::max<size_t>(1024, (size_t(32) << 20) / std::max(bps_dst_1, size_t(1)));
	{
		PinnedBuffer b_src_1(per_chunk * bps_src_1), b_src_2(per_chunk * bps_src_2), b_dst_1(per_chunk * bps_dst_1), b_dst_2(per_chunk * bps_dst_2 + 1);
		uint32_t crc_src_1 = 0, crc_src_2 = 0;
		for (size_t i = 0; i < num_bitslice; i += per_chunk) {
			const size_t n = std::min(per_chunk, num_bitslice - i);
			fin_1.read(reinterpret_cast<char*>(b_src_1.data()), (std::streamsize)(n * bps_src_1));
			if (!fin_1) throw __FILE__ ":merge_database_files: Error reading bitslices from source file 1";
			fin_2.read(reinterpret_cast<char*>(b_src_2.data()), (std::streamsize)(n * bps_src_2));
			if (!fin_2) throw __FILE__ ":merge_database_files: Error reading bitslices from source file 2";
			crc_src_1 = crc32_bytes(crc_src_1, b_src_1.data(), n * bps_src_1);
			crc_src_2 = crc32_bytes(crc_src_2, b_src_2.data(), n * bps_src_2);
			cuda_check(kwg_merge_slices(device, b_src_1.data(), src_1.num_filter, b_src_2.data(), src_2.num_filter, n, dst_1.num_filter,
				b_dst_1.data(), has_remainder ? b_dst_2.data() : NULL), __FILE__ ":merge_database_files: kwg_merge_slices failed");
			fout_1.write(reinterpret_cast<const char*>(b_dst_1.data()), (std::streamsize)(n * bps_dst_1));
			if (!fout_1) throw __FILE__ ":merge_database_files: Error writing bitslices to destination file 1";
			dst_1.crc32 = crc32_bytes(dst_1.crc32, b_dst_1.data(), n * bps_dst_1);
			if (has_remainder) {
				fout_2.write(reinterpret_cast<const char*>(b_dst_2.data()), (std::streamsize)(n * bps_dst_2));
				if (!fout_2) throw __FILE__ ":merge_database_files: Error writing bitslices to destination file 2";
				dst_2.crc32 = crc32_bytes(dst_2.crc32, b_dst_2.data(), n * bps_dst_2);
			}
		}
		if (crc_src_1 != src_1.crc32) throw __FILE__ ":merge_database_files: Invalid CRC32 value for source database file 1";
		if (crc_src_2 != src_2.crc32) throw __FILE__ ":merge_database_files: Invalid CRC32 value for source database file 2";
	}

	// metadata: location tables (placeholders first), then the FilterInfo records in column order
	dst_1.info_start = (uint64_t)fout_1.tellp();
	if (has_remainder) dst_2.info_start = (uint64_t)fout_2.tellp();
	std::vector<uint64_t> loc_1(dst_1.num_filter, 0), loc_2(dst_2.num_filter, 0);
	fout_1.write(reinterpret_cast<const char*>(loc_1.data()), (std::streamsize)(loc_1.size() * sizeof(uint64_t)));
	if (!fout_1) throw __FILE__ ":merge_database_files: Error writing dummy metadata location buffer to destination databaes file 1";
	if (has_remainder) {
		fout_2.write(reinterpret_cast<const char*>(loc_2.data()), (std::streamsize)(loc_2.size() * sizeof(uint64_t)));
		if (!fout_2) throw __FILE__ ":merge_database_files: Error writing dummy metadata location buffer to destination databaes file 2";
	}
	fin_1.seekg((std::streamoff)((uint64_t)fin_1.tellg() + sizeof(uint64_t) * src_1.num_filter));
	fin_2.seekg((std::streamoff)((uint64_t)fin_2.tellg() + sizeof(uint64_t) * src_2.num_filter));
	const size_t take = dst_1.num_filter - src_1.num_filter;
	for (size_t i = 0; i < src_1.num_filter; ++i) {
		FilterInfo info;
		binary_read(fin_1, info);
		loc_1[i] = (uint64_t)fout_1.tellp();
		binary_write(fout_1, info);
		if (!fout_1) throw __FILE__ ":merge_database_files: Error writing Bloom filter info to destination file 1";
	}
	for (size_t i = 0; i < take; ++i) {
		FilterInfo info;
		binary_read(fin_2, info);
		loc_1[src_1.num_filter + i] = (uint64_t)fout_1.tellp();
		binary_write(fout_1, info);
		if (!fout_1) throw __FILE__ ":merge_database_files: Error writing Bloom filter info to destination file 1 (b)";
	}
	for (size_t i = 0; i < dst_2.num_filter && has_remainder; ++i) {
		FilterInfo info;
		binary_read(fin_2, info);
		loc_2[i] = (uint64_t)fout_2.tellp();
		binary_write(fout_2, info);
		if (!fout_2) throw __FILE__ ":merge_database_files: Error writing Bloom filter info to destination file 2";
	}
	fout_1.seekp(0);
	binary_write(fout_1, dst_1);
	if (!fout_1) throw __FILE__ ":merge_database_files: Error writing database file 1 header (final)";
	fout_1.seekp((std::streamoff)dst_1.info_start);
	fout_1.write(reinterpret_cast<const char*>(loc_1.data()), (std::streamsize)(loc_1.size() * sizeof(uint64_t)));
	if (!fout_1) throw __FILE__ ":merge_database_files: Error writing metadata location buffer to destination file 1";
	fout_1.close();
	if (has_remainder) {
		fout_2.seekp(0);
		binary_write(fout_2, dst_2);
		if (!fout_2) throw __FILE__ ":merge_database_files: Error writing database file 2 header (final)";
		fout_2.seekp((std::streamoff)dst_2.info_start);
		fout_2.write(reinterpret_cast<const char*>(loc_2.data()), (std::streamsize)(loc_2.size() * sizeof(uint64_t)));
		if (!fout_2) throw __FILE__ ":merge_database_files: Error writing metadata location buffer to destination file 2";
		fout_2.close();
	}
	fin_1.close();
	fin_2.close();
	if (std::rename(dst_file_1.c_str(), file_1.c_str()) != 0) throw __FILE__ ":merge_database_files: Error renaming database file 1";
	if (has_remainder) {
		if (std::rename(dst_file_2.c_str(), file_2.c_str()) != 0) throw __FILE__ ":merge_database_files: Error renaming database file 2";
	}
	else if (::unlink(file_2.c_str()) != 0) throw __FILE__ ":merge_database_files: Error removing database file 2";
	return ret;
}

// filters per database file for a filter length (merge_db.cpp:88-100, options.h:137-138): 2048, fewer when that would pass 64 GiB
size_t max_filters_per_database_file(uint32_t log_2_filter_len)
{
	const uint64_t num_bloom = (64ull * 8ull * (1ull << 30)) >> log_2_filter_len;
	return (size_t)std::min<uint64_t>(2048, num_bloom);
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

// n bytes of a file from `offset` into dst, by up to four threads (pread on one descriptor)
static void read_slices(const std::string& path, uint64_t offset, uint8_t* dst, size_t n)
{
	const int fd = ::open(path.c_str(), O_RDONLY);
	if (fd < 0) throw __FILE__ ":search: Error reading slice from file (open)";
	const size_t n_thr = n >= (size_t(8) << 20) ? 4 : 1;
	const size_t per = (n + n_thr - 1) / n_thr;
	std::vector<std::thread> thr;
	std::vector<int> ok(n_thr, 1);
	for (size_t t = 0; t < n_thr; ++t) {
		const size_t a = t * per, z = std::min(n, a + per);
		if (a >= z) continue;
		thr.push_back(std::thread([fd, offset, dst, a, z, t, &ok]() {
			size_t done = a;
			while (done < z) {
				const ssize_t got = ::pread(fd, dst + done, z - done, (off_t)(offset + done));
				if (got <= 0) { ok[t] = 0; return; }
				done += (size_t)got;
			}
		}));
	}
	for (size_t t = 0; t < thr.size(); ++t) thr[t].join();
	::close(fd);
	for (size_t t = 0; t < n_thr; ++t) if (!ok[t]) throw __FILE__ ":search: Error reading slice from file (1)";
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
			q.slices_start = (uint64_t)q.fin->tellg();
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
		// Stream every file's slice region into its columns: two page-locked buffers alternate, piece n + 1 is read from
		// the file (by several threads: one thread copies out of the page cache at a fraction of the PCIe rate) while
		// piece n travels to the device.  A region may exceed host memory.
		const uint64_t n_rows = 1ULL << hdr.log_2_filter_len;
		// (page-locking costs ~0.5 ms per MiB: small files get small buffers)
		uint64_t largest = 0;
		for (size_t f = 0; f < parts.size(); ++f)
			largest = std::max<uint64_t>(largest, n_rows * (parts[f].hdr.num_filter / 8 + ((parts[f].hdr.num_filter % 8) ? 1 : 0)));
		const size_t piece_bytes = (size_t)std::min<uint64_t>(uint64_t(64) << 20, std::max<uint64_t>(largest, 4096));
		PinnedBuffer buf0(piece_bytes), buf1(piece_bytes);
		uint8_t* bufs[2] = {buf0.data(), buf1.data()};
		struct Piece { size_t part; uint64_t row, rows; };
		std::vector<Piece> pieces;
		for (size_t f = 0; f < parts.size(); ++f) {
			const size_t slice_size = parts[f].hdr.num_filter / 8 + ((parts[f].hdr.num_filter % 8) ? 1 : 0);
			const uint64_t piece_rows = std::max<uint64_t>(1, piece_bytes / std::max<size_t>(slice_size, 1));
			if (slice_size > piece_bytes) throw __FILE__ ":search: A slice is larger than the load buffer";
			for (uint64_t r = 0; r < n_rows; r += piece_rows) pieces.push_back(Piece{f, r, std::min(piece_rows, n_rows - r)});
		}
		for (size_t i = 0; i < pieces.size(); ++i) {
			const Piece& pc = pieces[i];
			Part& q = parts[pc.part];
			const size_t slice_size = q.hdr.num_filter / 8 + ((q.hdr.num_filter % 8) ? 1 : 0);
			uint8_t* buf = bufs[i & 1];
			read_slices(filenames[pc.part], q.slices_start + pc.row * slice_size, buf, (size_t)(pc.rows * slice_size));
			// (piece i - 1 travelled while piece i was read; its buffer is the next one to be filled)
			if (i) cuda_check(kwg_db_sync(db), __FILE__ ":search: kwg_db_sync failed");
			const int rc = (parts.size() == 1) ? kwg_db_upload_rows_async(db, pc.row, pc.rows, buf)
			                                   : kwg_db_upload_columns_async(db, q.col_begin, q.hdr.num_filter, pc.row, pc.rows, buf);
			cuda_check(rc, __FILE__ ":search: kwg_db_upload failed");
		}
		cuda_check(kwg_db_sync(db), __FILE__ ":search: kwg_db_sync failed");
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

void SubjectDatabase::search_gather(kwg_comm_t* comm, int root, uint32_t filter0, const std::vector<std::string>& queries, float threshold,
	std::vector<kwg_hit_t>& hits_out, std::vector<uint32_t>& n_kmers)
{
	if (queries.size() > 0xFFFFFFFFull) throw __FILE__ ":search: more than 2^32 queries in one call (split the query set)";
	std::string flat;
	std::vector<uint64_t> offsets(1, 0);
	for (size_t i = 0; i < queries.size(); ++i) { flat += queries[i]; offsets.push_back(flat.size()); }
	if (flat.empty()) flat.push_back('N');
	n_kmers.assign(queries.size(), 0);
	kwg_hit_t* hits = NULL;
	uint64_t n_hits = 0;
	cuda_check(kwg_search_gather(db, comm, root, flat.data(), offsets.data(), (uint32_t)queries.size(), threshold, filter0,
		n_kmers.data(), &hits, &n_hits), __FILE__ ":search: kwg_search_gather failed");
	hits_out.assign(hits, hits + n_hits);
	kwg_free_hits(hits);
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
