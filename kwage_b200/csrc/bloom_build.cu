// Bloom construction on sm_100a: read tiles -> 2-bit canonical k-mers -> murmur3 multi-hash ->
//   raw mode      : red.or into the (L2-resident where it fits) filter
//   counting mode : exact, order-free form of the reference's two 4-bit counting Bloom filters
//                   for min_kmer_count == 1 (reference make_bloom.cpp:506-621):
//        a slot of the counting filters is non-zero at the time occurrence t is processed iff an
//        earlier occurrence touched it (counters only ever grow from 0 and never wrap for c == 1),
//        so t is "valid" iff it is the FIRST toucher (minimum stream position) of at least one of
//        its four slots {first[h0], first[h1], second[h2], second[h3]}.
//        pass A: old = atomicMin(T[slot], stream position)          (HBM sector RMW)
//                old > position  -> this occurrence is (so far) the first toucher: its "won" bit is set;
//                if old was another occurrence of this batch, that one has just been displaced and
//                is flagged for a re-check (execution order is not stream order)
//        pass B: no random access: valid <=> won bit, except flagged occurrences, which re-read their
//                four slots (T[slot] == own position).  Valid canonical words are appended to a
//                compact list in HBM (instead of the reference's 5 x 2^Lmax valid_bits vectors)
//        finalize: for every listed word set bit (hash_h & (2^L-1)), h < num_hash -- which is what
//                the reference's fold of valid_bits[h] computes (make_bloom.cpp:337-354).
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace kwg {

constexpr int SCAN_THREADS = 256;
constexpr int TILE_BASES = 4096;                 // k-mer start positions per tile
constexpr int TILE_HALO = 32;                    // >= KWG_MAX_KMER_LEN - 1, keeps loads 16-byte granular
constexpr int TILE_LOAD = TILE_BASES + TILE_HALO;
constexpr int TILE_VEC = TILE_LOAD / 16;         // 16-base groups per tile
constexpr uint32_t T_EMPTY = 0xFFFFFFFFu;
constexpr uint64_t LIST_CHUNK_LOG2 = 22;         // valid-word list grows in 32 MiB chunks
constexpr uint64_t LIST_CHUNK = 1ull << LIST_CHUNK_LOG2;
constexpr uint64_t MAX_BATCH_BASES = 1ull << 28; // host batches are cut at read boundaries near this

enum ScanMode { MODE_RAW = 0, MODE_PASS_A = 1, MODE_PASS_B = 2 };

struct ScanParams {
	const char* bases;           // device, 16-byte aligned
	uint64_t n_bases;
	const uint32_t* start_mask;  // bit p set <=> a read starts at base p
	uint32_t k;
	// raw
	uint32_t* filter;
	uint32_t filter_mask;
	// counting
	uint32_t* T;                 // [2][2^lc]
	uint32_t count_mask;
	uint64_t count_len;
	uint32_t epoch_base;         // stream position of base 0 of this batch within the epoch
	uint64_t* const* list_chunks;
	unsigned long long* counter; // valid k-mers (counting) or inserted occurrences (raw)
	uint32_t* won;               // bit p: occurrence at batch position p was the first toucher of a slot when it ran
	uint32_t* recheck;           // bit p: occurrence p was displaced afterwards; pass B must re-read its slots
};

template <int MODE, int NH>
__global__ void __launch_bounds__(SCAN_THREADS)
kmer_scan_kernel(const ScanParams P)
{
	__shared__ uint32_t s_codes[TILE_VEC + 2];
	__shared__ uint32_t s_bad[TILE_LOAD / 32 + 2];
	__shared__ uint32_t s_start[TILE_LOAD / 32 + 2];
	__shared__ uint64_t s_words[MODE == MODE_PASS_B ? TILE_BASES : 1];   // pass B: valid words of the tile
	__shared__ uint32_t s_count;
	__shared__ unsigned long long s_base;

	const uint32_t tid = threadIdx.x;
	if (tid == 0) s_count = 0;
	const uint32_t k = P.k;
	const uint64_t t0 = (uint64_t)blockIdx.x * TILE_BASES;

	// ---- stage 1: 128-bit coalesced loads of the tile (+halo), encode, park in shared memory
	for (uint32_t v = tid; v < TILE_VEC; v += SCAN_THREADS) {
		const uint64_t g = t0 + (uint64_t)v * 16;
		uint32_t codes = 0, bad16 = 0xFFFFu;
		if (g + 16 <= P.n_bases) {
			encode16(ld_nc_v4(P.bases + g), codes, bad16);
		} else if (g < P.n_bases) {
			// ragged end of the batch: bytes past n_bases behave like separators
			uint32_t w[4] = {0, 0, 0, 0};
			for (uint32_t j = 0; j < 16; ++j) {
				const uint32_t b = (g + j < P.n_bases) ? (uint8_t)P.bases[g + j] : (uint32_t)'N';
				w[j >> 2] |= b << (8 * (j & 3));
			}
			encode16(make_uint4(w[0], w[1], w[2], w[3]), codes, bad16);
		}
		s_codes[v] = codes;
		reinterpret_cast<uint16_t*>(s_bad)[v] = (uint16_t)bad16;
	}
	for (uint32_t v = tid; v < TILE_LOAD / 32 + 1; v += SCAN_THREADS) {
		const uint64_t w = (t0 >> 5) + v;
		s_start[v] = (w * 32 < P.n_bases) ? P.start_mask[w] : 0u;
	}
	if (tid == 0) {
		s_codes[TILE_VEC] = 0; s_codes[TILE_VEC + 1] = 0;
		s_bad[TILE_LOAD / 32] = 0xFFFFFFFFu; s_bad[TILE_LOAD / 32 + 1] = 0xFFFFFFFFu;
		s_start[TILE_LOAD / 32 + 1] = 0;
	}
	__syncthreads();

	unsigned long long raw_local = 0;

	// ---- stage 2: one k-mer start position per thread per iteration
#pragma unroll 1
	for (uint32_t it = 0; it < TILE_BASES / SCAN_THREADS; ++it) {
		const uint32_t p = it * SCAN_THREADS + tid;
		// a read that starts strictly inside the window breaks it (fragments are independent,
		// reference make_bloom.cpp:277-283 calls count_words once per fragment)
		const bool ok = window_ok(s_bad, s_start, p, k);

		Canon c;
		c.word = 0; c.low = 0;
		if (ok) c = canonical(window_sense(s_codes, p, k), k);

		if (MODE == MODE_RAW) {
			if (ok) {
				uint32_t h[NH];
				murmur3_multi<NH>(c.low, k, h);
#pragma unroll
				for (int s = 0; s < NH; ++s) {
					const uint32_t bit = h[s] & P.filter_mask;
					atomicOr(P.filter + (bit >> 5), 1u << (bit & 31));
				}
				++raw_local;
			}
		} else {
			const uint32_t pos = P.epoch_base + (uint32_t)(t0 + p);
			uint32_t* T1 = P.T;
			uint32_t* T2 = P.T + P.count_len;
			const uint32_t lane = tid & 31;
			const uint64_t word = (t0 + p) >> 5;            // a warp owns one 32-position word of the bitmaps
			if (MODE == MODE_PASS_A) {
				bool won = false;
				if (ok) {
					uint32_t h[4];
					murmur3_multi<4>(c.low, k, h);
					uint32_t old[4];
					old[0] = atomicMin(T1 + (h[0] & P.count_mask), pos);
					old[1] = atomicMin(T1 + (h[1] & P.count_mask), pos);
					old[2] = atomicMin(T2 + (h[2] & P.count_mask), pos);
					old[3] = atomicMin(T2 + (h[3] & P.count_mask), pos);
#pragma unroll
					for (int s = 0; s < 4; ++s) {
						if (old[s] > pos) {
							won = true;
							if (old[s] != T_EMPTY) {            // displaced a later occurrence of this batch
								const uint32_t q = old[s] - P.epoch_base;
								atomicOr(P.recheck + (q >> 5), 1u << (q & 31));
							}
						}
					}
				}
				const uint32_t m = __ballot_sync(0xFFFFFFFFu, won);
				if (lane == 0) P.won[word] = m;
			} else {
				bool valid = false;
				if (ok) {
					const bool flagged = (P.recheck[word] >> lane) & 1u;
					valid = (P.won[word] >> lane) & 1u;
					if (flagged) {
						uint32_t h[4];
						murmur3_multi<4>(c.low, k, h);
						valid = (T1[h[0] & P.count_mask] == pos) | (T1[h[1] & P.count_mask] == pos) |
						        (T2[h[2] & P.count_mask] == pos) | (T2[h[3] & P.count_mask] == pos);
					}
				}
				// block-aggregated append: valid words are compacted in shared memory first, so the
				// whole tile costs ONE atomicAdd on the global list cursor and coalesced stores
				const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
				if (m) {
					const uint32_t leader = __ffs(m) - 1;
					uint32_t base = 0;
					if (lane == leader) base = atomicAdd(&s_count, (uint32_t)__popc(m));
					base = __shfl_sync(0xFFFFFFFFu, base, leader);
					if (valid) s_words[base + __popc(m & ((1u << lane) - 1u))] = c.word;
				}
			}
		}
	}

	if (MODE == MODE_PASS_B) {
		__syncthreads();
		const uint32_t n = s_count;
		if (n) {
			if (tid == 0) s_base = atomicAdd(P.counter, (unsigned long long)n);
			__syncthreads();
			const unsigned long long base = s_base;
			for (uint32_t i = tid; i < n; i += SCAN_THREADS) {
				const uint64_t at = base + i;
				P.list_chunks[at >> LIST_CHUNK_LOG2][at & (LIST_CHUNK - 1)] = s_words[i];
			}
		}
	}

	if (MODE == MODE_RAW) {
		// occurrences inserted (BloomProgress-style bookkeeping for the raw rig)
		for (int o = 16; o > 0; o >>= 1) raw_local += __shfl_down_sync(0xFFFFFFFFu, raw_local, o);
		if ((tid & 31) == 0 && raw_local) atomicAdd(P.counter, raw_local);
	}
}

// bit p of start_mask <=> some read starts at base p of the batch
__global__ void mark_read_starts_kernel(const uint64_t* __restrict__ offsets, uint64_t n_reads, uint64_t off0,
	uint64_t n_bases, uint32_t* __restrict__ start_mask)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_reads) return;
	const uint64_t p = offsets[i] - off0;
	if (p < n_bases) atomicOr(start_mask + (p >> 5), 1u << (p & 31));
}

// epoch roll-over: every touched slot becomes "touched before anything in the new epoch"
__global__ void flatten_epoch_kernel(uint32_t* __restrict__ T, uint64_t n)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		if (T[i] != T_EMPTY) T[i] = 0u;
}

// finalize (counting mode): set bit (hash_h & mask) for every listed word
template <int NH>
__global__ void __launch_bounds__(256)
insert_words_kernel(uint64_t* const* __restrict__ chunks, uint64_t n_words, uint32_t k, uint32_t* __restrict__ filter,
	uint32_t filter_mask)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
		const uint64_t w = chunks[i >> LIST_CHUNK_LOG2][i & (LIST_CHUNK - 1)];
		const uint64_t low = reverse_groups(w, k);
		uint32_t h[NH];
		murmur3_multi<NH>(low, k, h);
#pragma unroll
		for (int s = 0; s < NH; ++s) {
			const uint32_t bit = h[s] & filter_mask;
			atomicOr(filter + (bit >> 5), 1u << (bit & 31));
		}
	}
}

} // namespace kwg

using namespace kwg;

struct kwg_bloom {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool raw = false;
	uint32_t k = 0, min_count = 0, lc = 0, lmax = 0, raw_nh = 0, raw_L = 0;
	// counting mode
	uint32_t* T = nullptr;
	uint64_t epoch_pos = 1;              // next free stream position in the current epoch (0 = "earlier epoch")
	std::vector<uint64_t*> chunks;
	uint64_t** d_chunk_table = nullptr;
	size_t table_cap = 0;
	// both modes
	unsigned long long* d_counter = nullptr;
	unsigned long long* h_counter = nullptr;   // pinned
	uint32_t* d_filter = nullptr;              // raw: the filter; counting: finalize scratch
	size_t filter_cap = 0;
	// staging for host inputs
	char* d_bases = nullptr;
	size_t bases_cap = 0;
	uint64_t* d_offsets = nullptr;
	size_t offsets_cap = 0;
	uint32_t* d_start = nullptr;
	size_t start_cap = 0;
	uint32_t* d_won = nullptr;
	size_t won_cap = 0;
	uint32_t* d_recheck = nullptr;
	size_t recheck_cap = 0;
	KernelTimers timers;
};

static int grow(void** p, size_t* cap, size_t need)
{
	if (need <= *cap) return KWG_OK;
	if (*p) KWG_CUDA(cudaFree(*p));
	*p = nullptr; *cap = 0;
	const size_t want = round_up(need + need / 8, 256);
	KWG_CUDA(cudaMalloc(p, want));
	*cap = want;
	return KWG_OK;
}

static int read_counter(kwg_bloom* b, uint64_t* out)
{
	KWG_CUDA(cudaMemcpyAsync(b->h_counter, b->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	*out = *b->h_counter;
	return KWG_OK;
}

static int ensure_list_capacity(kwg_bloom* b, uint64_t words)
{
	const size_t need_chunks = (size_t)ceil_div(words, LIST_CHUNK);
	if (need_chunks <= b->chunks.size()) return KWG_OK;
	while (b->chunks.size() < need_chunks) {
		uint64_t* c = nullptr;
		KWG_CUDA(cudaMalloc(&c, LIST_CHUNK * sizeof(uint64_t)));
		b->chunks.push_back(c);
	}
	if (b->chunks.size() > b->table_cap) {
		// kernels that still use the old table must finish before it is released
		KWG_CUDA(cudaStreamSynchronize(b->stream));
		if (b->d_chunk_table) KWG_CUDA(cudaFree(b->d_chunk_table));
		b->d_chunk_table = nullptr;
		b->table_cap = std::max<size_t>(64, b->chunks.size() * 2);
		KWG_CUDA(cudaMalloc(&b->d_chunk_table, b->table_cap * sizeof(uint64_t*)));
	}
	KWG_CUDA(cudaMemcpyAsync(b->d_chunk_table, b->chunks.data(), b->chunks.size() * sizeof(uint64_t*),
		cudaMemcpyHostToDevice, b->stream));
	// the host vector may reallocate later: make sure the copy has consumed it
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	return KWG_OK;
}

template <int MODE>
static int launch_scan(kwg_bloom* b, const ScanParams& P)
{
	const uint64_t tiles = ceil_div(P.n_bases, TILE_BASES);
	if (tiles == 0) return KWG_OK;
	if (tiles > 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "batch too large");
	const dim3 grid((unsigned)tiles), block(SCAN_THREADS);
	b->timers.begin(MODE == MODE_PASS_B ? KWG_T_SCAN_B : KWG_T_SCAN_A, b->stream);
	if (MODE == MODE_RAW) {
		switch (b->raw_nh) {
#define KWG_CASE(N) case N: kmer_scan_kernel<MODE_RAW, N><<<grid, block, 0, b->stream>>>(P); break;
			KWG_CASE(1) KWG_CASE(2) KWG_CASE(3) KWG_CASE(4) KWG_CASE(5) KWG_CASE(6) KWG_CASE(7) KWG_CASE(8)
#undef KWG_CASE
			default: return fail(KWG_ERR_INVALID_ARG, "num_hash out of range");
		}
	} else {
		kmer_scan_kernel<MODE, 4><<<grid, block, 0, b->stream>>>(P);
	}
	b->timers.end(b->stream);
	KWG_LAUNCHED();
	return KWG_OK;
}

// All inputs on the device: d_bases (16-byte aligned), d_offsets[n_reads+1] with offsets relative
// to off0.  n_bases < 2^32 - 2.
static int add_batch_dev(kwg_bloom* b, const char* d_bases, const uint64_t* d_offsets, uint64_t n_reads,
	uint64_t off0, uint64_t n_bases)
{
	if (n_bases == 0 || n_reads == 0) return KWG_OK;
	if (n_bases >= 0xFFFFFFF0ull) return fail(KWG_ERR_INVALID_ARG, "a device batch must hold fewer than 2^32-16 bases");
	if ((reinterpret_cast<uintptr_t>(d_bases) & 15u) != 0) return fail(KWG_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");

	const size_t start_words = (size_t)(n_bases / 32 + 2);
	int rc = grow((void**)&b->d_start, &b->start_cap, start_words * sizeof(uint32_t));
	if (rc) return rc;
	KWG_CUDA(cudaMemsetAsync(b->d_start, 0, start_words * sizeof(uint32_t), b->stream));
	b->timers.begin(KWG_T_AUX, b->stream);
	mark_read_starts_kernel<<<(unsigned)ceil_div(n_reads, 256), 256, 0, b->stream>>>(d_offsets, n_reads, off0, n_bases, b->d_start);
	b->timers.end(b->stream);
	KWG_LAUNCHED();

	ScanParams P{};
	P.bases = d_bases;
	P.n_bases = n_bases;
	P.start_mask = b->d_start;
	P.k = b->k;
	P.counter = b->d_counter;

	if (b->raw) {
		P.filter = b->d_filter;
		P.filter_mask = (b->raw_L >= 32) ? 0xFFFFFFFFu : ((1u << b->raw_L) - 1u);
		return launch_scan<MODE_RAW>(b, P);
	}

	// counting mode
	if (b->epoch_pos + n_bases >= (uint64_t)T_EMPTY) {
		flatten_epoch_kernel<<<sm_count(b->device) * 8, 256, 0, b->stream>>>(b->T, 2ull << b->lc);
		KWG_LAUNCHED();
		b->epoch_pos = 1;
	}
	uint64_t n_valid = 0;
	rc = read_counter(b, &n_valid);
	if (rc) return rc;
	rc = ensure_list_capacity(b, n_valid + n_bases);
	if (rc) return rc;

	const size_t bitmap_words = (size_t)(ceil_div(n_bases, TILE_BASES) * (TILE_BASES / 32));
	if ((rc = grow((void**)&b->d_won, &b->won_cap, bitmap_words * sizeof(uint32_t)))) return rc;
	if ((rc = grow((void**)&b->d_recheck, &b->recheck_cap, bitmap_words * sizeof(uint32_t)))) return rc;
	KWG_CUDA(cudaMemsetAsync(b->d_recheck, 0, bitmap_words * sizeof(uint32_t), b->stream));
	P.won = b->d_won;
	P.recheck = b->d_recheck;
	P.T = b->T;
	P.count_len = 1ull << b->lc;
	P.count_mask = (b->lc >= 32) ? 0xFFFFFFFFu : ((1u << b->lc) - 1u);
	P.epoch_base = (uint32_t)b->epoch_pos;
	P.list_chunks = b->d_chunk_table;
	rc = launch_scan<MODE_PASS_A>(b, P);
	if (rc) return rc;
	rc = launch_scan<MODE_PASS_B>(b, P);
	if (rc) return rc;
	b->epoch_pos += n_bases;
	return KWG_OK;
}

static int bloom_alloc_common(kwg_bloom* b)
{
	if (const char* e = getenv("KWG_L2_FETCH")) {    // experiment knob: L2 fetch granularity hint (32/64/128)
		cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
		size_t v = 0;
		cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
		fprintf(stderr, "[kwg] L2 fetch granularity = %zu\n", v);
	}
	KWG_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	KWG_CUDA(cudaMalloc(&b->d_counter, sizeof(unsigned long long)));
	KWG_CUDA(cudaMallocHost(&b->h_counter, sizeof(unsigned long long)));
	KWG_CUDA(cudaMemsetAsync(b->d_counter, 0, sizeof(unsigned long long), b->stream));
	return KWG_OK;
}

extern "C" {

void kwg_bloom_destroy(kwg_bloom_t* b)
{
	if (!b) return;
	cudaSetDevice(b->device);
	if (b->stream) cudaStreamSynchronize(b->stream);
	cudaFree(b->T);
	for (uint64_t* c : b->chunks) cudaFree(c);
	cudaFree(b->d_chunk_table);
	cudaFree(b->d_counter);
	if (b->h_counter) cudaFreeHost(b->h_counter);
	cudaFree(b->d_filter);
	cudaFree(b->d_bases);
	cudaFree(b->d_offsets);
	cudaFree(b->d_start);
	cudaFree(b->d_won);
	cudaFree(b->d_recheck);
	if (b->stream) cudaStreamDestroy(b->stream);
	delete b;
}

int kwg_bloom_create(kwg_bloom_t** out, int device, uint32_t kmer_len, uint32_t min_kmer_count,
	uint32_t log2_count_len, uint32_t log2_max_len)
{
	if (!out) return fail(KWG_ERR_INVALID_ARG, "out is NULL");
	*out = nullptr;
	if (kmer_len < 1 || kmer_len > KWG_MAX_KMER_LEN) return fail(KWG_ERR_INVALID_ARG, "kmer_len must be in [1,32] (reference word.h:10)");
	if (min_kmer_count < 1 || min_kmer_count > 15) return fail(KWG_ERR_INVALID_ARG, "min_kmer_count must be in [1,15] (reference make_bloom.cpp:61,90-92)");
	if (log2_count_len < 18 || log2_count_len > 32) return fail(KWG_ERR_INVALID_ARG, "log2_count_len must be in [18,32] (reference make_bloom.cpp:21-22)");
	if (log2_max_len > 32) return fail(KWG_ERR_INVALID_ARG, "log2_max_len must be <= 32 (32-bit hash)");
	if (min_kmer_count != 1)
		return fail(KWG_ERR_UNSUPPORTED, "counting mode with min_kmer_count > 1 is order dependent (conservative-update counters, "
			"reference make_bloom.cpp:546-601) and is not implemented on the device");
	int rc = select_device(device);
	if (rc) return rc;
	kwg_bloom* b = new kwg_bloom();
	b->device = device; b->raw = false; b->k = kmer_len; b->min_count = min_kmer_count;
	b->lc = log2_count_len; b->lmax = log2_max_len;
	rc = bloom_alloc_common(b);
	if (rc == KWG_OK) {
		const size_t bytes = (size_t)(2ull << b->lc) * sizeof(uint32_t);
		cudaError_t e = cudaMalloc(&b->T, bytes);
		if (e != cudaSuccess) rc = fail(KWG_ERR_NO_MEMORY, std::string("first-touch table: ") + cudaGetErrorString(e));
		else if (cudaMemsetAsync(b->T, 0xFF, bytes, b->stream) != cudaSuccess) rc = fail(KWG_ERR_CUDA, "memset of first-touch table failed");
	}
	if (rc) { kwg_bloom_destroy(b); return rc; }
	*out = b;
	return KWG_OK;
}

int kwg_bloom_create_raw(kwg_bloom_t** out, int device, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len)
{
	if (!out) return fail(KWG_ERR_INVALID_ARG, "out is NULL");
	*out = nullptr;
	if (kmer_len < 1 || kmer_len > KWG_MAX_KMER_LEN) return fail(KWG_ERR_INVALID_ARG, "kmer_len must be in [1,32] (reference word.h:10)");
	if (num_hash < 1 || num_hash > KWG_MAX_NUM_HASH) return fail(KWG_ERR_INVALID_ARG, "num_hash must be in [1,8] (reference hash.cpp:7,243-245)");
	if (log2_len < 5 || log2_len > 32) return fail(KWG_ERR_INVALID_ARG, "log2_len must be in [5,32]");
	int rc = select_device(device);
	if (rc) return rc;
	kwg_bloom* b = new kwg_bloom();
	b->device = device; b->raw = true; b->k = kmer_len; b->raw_nh = num_hash; b->raw_L = log2_len;
	rc = bloom_alloc_common(b);
	if (rc == KWG_OK) {
		const size_t bytes = (size_t)1 << (log2_len - 3);
		rc = grow((void**)&b->d_filter, &b->filter_cap, bytes);
		if (rc == KWG_OK && cudaMemsetAsync(b->d_filter, 0, bytes, b->stream) != cudaSuccess) rc = fail(KWG_ERR_CUDA, "memset of filter failed");
	}
	if (rc) { kwg_bloom_destroy(b); return rc; }
	*out = b;
	return KWG_OK;
}

int kwg_bloom_reset(kwg_bloom_t* b)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	int rc = select_device(b->device);
	if (rc) return rc;
	KWG_CUDA(cudaMemsetAsync(b->d_counter, 0, sizeof(unsigned long long), b->stream));
	if (b->raw) {
		KWG_CUDA(cudaMemsetAsync(b->d_filter, 0, (size_t)1 << (b->raw_L - 3), b->stream));
	} else {
		KWG_CUDA(cudaMemsetAsync(b->T, 0xFF, (size_t)(2ull << b->lc) * sizeof(uint32_t), b->stream));
		b->epoch_pos = 1;
	}
	return KWG_OK;
}

int kwg_bloom_add_reads_dev(kwg_bloom_t* b, const char* d_bases, const uint64_t* d_offsets, uint64_t n_reads, uint64_t n_bases)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (n_reads == 0 || n_bases == 0) return KWG_OK;
	if (!d_bases || !d_offsets) return fail(KWG_ERR_INVALID_ARG, "NULL input");
	int rc = select_device(b->device);
	if (rc) return rc;
	return add_batch_dev(b, d_bases, d_offsets, n_reads, 0, n_bases);
}

int kwg_bloom_add_reads(kwg_bloom_t* b, const char* bases, const uint64_t* offsets, uint64_t n_reads)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (n_reads == 0) return KWG_OK;
	if (!bases || !offsets) return fail(KWG_ERR_INVALID_ARG, "NULL input");
	int rc = select_device(b->device);
	if (rc) return rc;
	for (uint64_t r = 0; r < n_reads; ++r)
		if (offsets[r + 1] < offsets[r]) return fail(KWG_ERR_INVALID_ARG, "offsets must be non-decreasing");

	uint64_t r0 = 0;
	while (r0 < n_reads) {
		// cut a batch of whole reads of at most MAX_BATCH_BASES bases (at least one read)
		uint64_t r1 = r0 + 1;
		while (r1 < n_reads && offsets[r1 + 1] - offsets[r0] <= MAX_BATCH_BASES) ++r1;
		const uint64_t off0 = offsets[r0];
		const uint64_t nb = offsets[r1] - off0;
		if (nb >= 0xFFFFFFF0ull) return fail(KWG_ERR_INVALID_ARG, "a single read of 2^32 bases or more is not supported");
		if (nb > 0) {
			rc = grow((void**)&b->d_bases, &b->bases_cap, round_up(nb, 16) + 16);
			if (rc) return rc;
			rc = grow((void**)&b->d_offsets, &b->offsets_cap, (r1 - r0 + 1) * sizeof(uint64_t));
			if (rc) return rc;
			KWG_CUDA(cudaMemcpyAsync(b->d_bases, bases + off0, nb, cudaMemcpyHostToDevice, b->stream));
			KWG_CUDA(cudaMemcpyAsync(b->d_offsets, offsets + r0, (r1 - r0 + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
			rc = add_batch_dev(b, b->d_bases, b->d_offsets, r1 - r0, off0, nb);
			if (rc) return rc;
		}
		r0 = r1;
	}
	// host buffers may be reused by the caller as soon as we return
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	return KWG_OK;
}

int kwg_bloom_num_valid(kwg_bloom_t* b, uint64_t* n)
{
	if (!b || !n) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(b->device);
	if (rc) return rc;
	return read_counter(b, n);
}

int kwg_bloom_finalize_dev(kwg_bloom_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* d_out_bits)
{
	if (!b || !d_out_bits) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(b->device);
	if (rc) return rc;
	if (b->raw) {
		if (log2_len != b->raw_L || num_hash != b->raw_nh) return fail(KWG_ERR_INVALID_ARG, "raw mode: parameters differ from creation");
		KWG_CUDA(cudaMemcpyAsync(d_out_bits, b->d_filter, (size_t)1 << (log2_len - 3), cudaMemcpyDeviceToDevice, b->stream));
		return KWG_OK;
	}
	if (num_hash < 1 || num_hash > KWG_COUNT_NUM_HASH) return fail(KWG_ERR_INVALID_ARG, "num_hash must be in [1,5] (reference bloom.h:20-21)");
	if (log2_len < 5 || log2_len > b->lmax) return fail(KWG_ERR_INVALID_ARG, "log2_len must be in [5, log2_max_len]");
	if ((reinterpret_cast<uintptr_t>(d_out_bits) & 3u) != 0) return fail(KWG_ERR_INVALID_ARG, "d_out_bits must be 4-byte aligned");
	uint64_t n_valid = 0;
	rc = read_counter(b, &n_valid);
	if (rc) return rc;
	const size_t bytes = (size_t)1 << (log2_len - 3);
	KWG_CUDA(cudaMemsetAsync(d_out_bits, 0, bytes, b->stream));
	if (n_valid) {
		const uint32_t mask = (log2_len >= 32) ? 0xFFFFFFFFu : ((1u << log2_len) - 1u);
		const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(n_valid, 256), (uint64_t)sm_count(b->device) * 16);
		uint32_t* f = reinterpret_cast<uint32_t*>(d_out_bits);
		b->timers.begin(KWG_T_INSERT, b->stream);
		switch (num_hash) {
#define KWG_CASE(N) case N: insert_words_kernel<N><<<grid, 256, 0, b->stream>>>(b->d_chunk_table, n_valid, b->k, f, mask); break;
			KWG_CASE(1) KWG_CASE(2) KWG_CASE(3) KWG_CASE(4) KWG_CASE(5)
#undef KWG_CASE
		}
		b->timers.end(b->stream);
		KWG_LAUNCHED();
	}
	return KWG_OK;
}

int kwg_bloom_finalize(kwg_bloom_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* out_bits)
{
	if (!b || !out_bits) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(b->device);
	if (rc) return rc;
	if (log2_len < 5 || log2_len > 32) return fail(KWG_ERR_INVALID_ARG, "log2_len must be in [5,32]");
	const size_t bytes = (size_t)1 << (log2_len - 3);
	if (b->raw) {
		if (log2_len != b->raw_L || num_hash != b->raw_nh) return fail(KWG_ERR_INVALID_ARG, "raw mode: parameters differ from creation");
		KWG_CUDA(cudaMemcpyAsync(out_bits, b->d_filter, bytes, cudaMemcpyDeviceToHost, b->stream));
		KWG_CUDA(cudaStreamSynchronize(b->stream));
		return KWG_OK;
	}
	rc = grow((void**)&b->d_filter, &b->filter_cap, bytes);
	if (rc) return rc;
	rc = kwg_bloom_finalize_dev(b, log2_len, num_hash, reinterpret_cast<uint8_t*>(b->d_filter));
	if (rc) return rc;
	KWG_CUDA(cudaMemcpyAsync(out_bits, b->d_filter, bytes, cudaMemcpyDeviceToHost, b->stream));
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	return KWG_OK;
}

int kwg_bloom_sync(kwg_bloom_t* b)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	int rc = select_device(b->device);
	if (rc) return rc;
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	return KWG_OK;
}

int kwg_bloom_set_timing(kwg_bloom_t* b, int enable)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	b->timers.enabled = enable != 0;
	return KWG_OK;
}

int kwg_bloom_get_timing(kwg_bloom_t* b, double* ms, uint64_t* launches)
{
	if (!b || !ms || !launches) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(b->device);
	if (rc) return rc;
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	for (int i = 0; i < KWG_T_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
	b->timers.collect(ms, launches, KWG_T_COUNT);
	return KWG_OK;
}

int kwg_bloom_stream(kwg_bloom_t* b, void** stream)
{
	if (!b || !stream) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*stream = (void*)b->stream;
	return KWG_OK;
}

} // extern "C"
