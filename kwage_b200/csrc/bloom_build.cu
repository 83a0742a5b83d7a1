// Bloom construction on sm_100a: read tiles -> 2-bit canonical k-mers -> murmur3 multi-hash ->
//   raw mode      : red.or into the (L2-resident where it fits) filter
//   counting mode : exact, order-free form of the reference's two 4-bit counting Bloom filters
//                   (reference make_bloom.cpp:506-621).  min_kmer_count == 1:
//        a slot of the counting filters is non-zero at the time occurrence t is processed iff an
//        earlier occurrence touched it (counters only ever grow from 0 and never wrap for c == 1),
//        so t is "valid" iff it is the FIRST toucher (minimum stream position) of at least one of
//        its four slots {first[h0], first[h1], second[h2], second[h3]}.
//        The first toucher of every slot is found by radix-partitioning (slot, position) records and
//        resolving each 2^15-slot bucket in shared memory (bloom_count.cuh); what persists between
//        batches of one accession is a 1-bit-per-slot "touched" bitmap.
//        pass B: valid <=> fewer than 4 of the occurrence's touches lost.  Valid canonical words are
//                appended to a compact list in HBM (instead of the reference's 5 x 2^Lmax valid_bits vectors)
//        finalize: for every listed word set bit (hash_h & (2^L-1)), h < num_hash -- which is what
//                the reference's fold of valid_bits[h] computes (make_bloom.cpp:337-354).
//        min_kmer_count = c > 1: the conservative-update counters are resolved level by level, one
//        resolve launch per counter value 0 .. c-1 over the same sorted records (bloom_count.cuh,
//        resolve_kernel<true>); what persists between batches is the 4-bit counter of every slot, i.e.
//        the reference's own two tables.  Exact, including the double increment when both hashes of a
//        table meet; a 4-bit wrap (only possible for c == 15) is detected and reported as an error.
#include "common.cuh"
#include "bloom_count.cuh"
#include "bloom_first.cuh"

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
constexpr uint64_t LIST_CHUNK_LOG2 = 22;         // valid-word list grows in 32 MiB chunks
constexpr uint64_t LIST_CHUNK = 1ull << LIST_CHUNK_LOG2;
constexpr uint64_t MAX_BATCH_BASES = 1ull << 28; // host batches are cut at read boundaries near this
// A filter of up to 2^29 bits (64 MiB) stays in the 126 MB L2 and takes ~2e11 red.or per second; a larger one lives in
// HBM, where every bit set is a DRAM sector read-modify-write (2.6e10 /s, profiles/r1e_sweep.md).  Larger filters are
// therefore filled one 64 MiB window at a time: every pass re-derives the hashes (cheap next to the atomics it saves)
// and only sets the bits that fall into its window.
constexpr uint32_t WINDOW_LOG2 = 29;

enum ScanMode { MODE_RAW = 0, MODE_PASS_B = 2 };

struct ScanParams {
	BaseSource src;              // device, 16-byte aligned
	const uint32_t* start_mask;  // bit p set <=> a read starts at base p
	uint32_t k;
	// raw
	uint32_t* filter;
	uint32_t filter_mask;
	uint32_t win_id, n_win;      // filters beyond WINDOW_LOG2 bits: this launch only sets the bits of window win_id
	// counting
	uint64_t pos0;               // absolute base index of the first start position of this sub-batch
	uint64_t n_pos;              // start positions in this sub-batch
	uint64_t* const* list_chunks;
	unsigned long long* counter; // valid k-mers (counting) or inserted occurrences (raw)
	const uint32_t* loss;        // 4-bit counters, 8 positions per word: touches of the occurrence that lost
	const uint32_t* elig;        // min_kmer_count > 1: bit per position, occurrence eligible at the last level (NULL: all)
	uint32_t wins;               // `loss` holds wins of the last level (min_kmer_count > 1) instead of losses
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
	const uint64_t rel0 = (uint64_t)blockIdx.x * TILE_BASES;   // position inside the sub-batch
	const uint64_t t0 = P.pos0 + rel0;                          // absolute base index

	// ---- stage 1: 128-bit coalesced loads of the tile (+halo), encode, park in shared memory
	for (uint32_t v = tid; v < TILE_VEC; v += SCAN_THREADS) {
		const uint64_t g = t0 + (uint64_t)v * 16;
		uint32_t codes, bad16;
		load_group16(P.src, g, codes, bad16);         // (the ragged end of the batch reads as separators)
		s_codes[v] = codes;
		reinterpret_cast<uint16_t*>(s_bad)[v] = (uint16_t)bad16;
	}
	for (uint32_t v = tid; v < TILE_LOAD / 32 + 1; v += SCAN_THREADS) {
		const uint64_t w = (t0 >> 5) + v;
		s_start[v] = (w * 32 < P.src.n_bases) ? P.start_mask[w] : 0u;
	}
	if (tid == 0) {
		s_codes[TILE_VEC] = 0; s_codes[TILE_VEC + 1] = 0;
		s_bad[TILE_LOAD / 32] = 0xFFFFFFFFu; s_bad[TILE_LOAD / 32 + 1] = 0xFFFFFFFFu;
		s_start[TILE_LOAD / 32 + 1] = 0;
	}
	__syncthreads();

	unsigned long long raw_local = 0;
	const uint64_t pol_keep = l2_policy_evict_last();     // raw mode: the filter asks to stay in the L2

	// ---- stage 2: one k-mer start position per thread per iteration
#pragma unroll 1
	for (uint32_t it = 0; it < TILE_BASES / SCAN_THREADS; ++it) {
		const uint32_t p = it * SCAN_THREADS + tid;
		// a read that starts strictly inside the window breaks it (fragments are independent,
		// reference make_bloom.cpp:277-283 calls count_words once per fragment)
		const bool ok = (rel0 + p < P.n_pos) && window_ok(s_bad, s_start, p, k);

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
					if (P.n_win == 1 || (bit >> WINDOW_LOG2) == P.win_id) atomicOr(P.filter + (bit >> 5), 1u << (bit & 31));
				}
				if (P.win_id == 0) ++raw_local;
			}
		} else {
			const uint32_t lane = tid & 31;
			bool valid = false;
			if (ok) {
				const uint64_t rel = rel0 + p;
				const uint32_t a = (P.loss[rel >> 3] >> ((rel & 7u) << 2)) & 0xFu;
				// min count 1: fewer than 4 touches lost; min count c > 1: eligible at level c-1 and a win there
				valid = P.wins ? (a != 0u && (!P.elig || ((P.elig[rel >> 5] >> (rel & 31u)) & 1u))) : (a < 4u);
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

// Raw mode for k in 33..63 (BASELINE.json configs[4] sweeps k up to 63; the reference stops at 32, word.h:10, so parity is
// UNPINNED there: the yardstick is oracle kwo_raw_insert_wide).  Same structure as kmer_scan_kernel<MODE_RAW>, 128-bit
// words, a 64-base halo.
constexpr int WT_HALO = 64;
constexpr int WT_LOAD = TILE_BASES + WT_HALO;
constexpr int WT_VEC = WT_LOAD / 16;

template <int NH>
__global__ void __launch_bounds__(SCAN_THREADS)
kmer_scan_wide_kernel(const ScanParams P)
{
	__shared__ uint32_t s_codes[WT_VEC + 4];
	__shared__ uint32_t s_bad[WT_LOAD / 32 + 3];
	__shared__ uint32_t s_start[WT_LOAD / 32 + 3];

	const uint32_t tid = threadIdx.x;
	const uint32_t k = P.k;
	const uint64_t rel0 = (uint64_t)blockIdx.x * TILE_BASES;
	const uint64_t t0 = P.pos0 + rel0;
	for (uint32_t v = tid; v < (uint32_t)WT_VEC; v += SCAN_THREADS) {
		uint32_t codes, bad16;
		load_group16(P.src, t0 + (uint64_t)v * 16, codes, bad16);
		s_codes[v] = codes;
		reinterpret_cast<uint16_t*>(s_bad)[v] = (uint16_t)bad16;
	}
	for (uint32_t v = tid; v < (uint32_t)(WT_LOAD / 32 + 1); v += SCAN_THREADS) {
		const uint64_t w = (t0 >> 5) + v;
		s_start[v] = (w * 32 < P.src.n_bases) ? P.start_mask[w] : 0u;
	}
	if (tid < 4) s_codes[WT_VEC + tid] = 0;
	if (tid < 3) s_bad[WT_LOAD / 32 + tid] = 0xFFFFFFFFu;
	if (tid < 2) s_start[WT_LOAD / 32 + 1 + tid] = 0;
	__syncthreads();

	unsigned long long raw_local = 0;
	const uint64_t pol_keep = l2_policy_evict_last();
#pragma unroll 1
	for (uint32_t it = 0; it < TILE_BASES / SCAN_THREADS; ++it) {
		const uint32_t p = it * SCAN_THREADS + tid;
		if ((rel0 + p < P.n_pos) && window_ok_wide(s_bad, s_start, p, k)) {
			const CanonWide c = canonical_wide(window_sense_wide(s_codes, p, k), k);
			uint32_t h[NH];
			murmur3_multi_wide<NH>(c.low, k, h);
#pragma unroll
			for (int s = 0; s < NH; ++s) {
				const uint32_t bit = h[s] & P.filter_mask;
				if (P.n_win == 1 || (bit >> WINDOW_LOG2) == P.win_id) red_or_hint(P.filter + (bit >> 5), 1u << (bit & 31), pol_keep);
			}
			if (P.win_id == 0) ++raw_local;
		}
	}
	for (int o = 16; o > 0; o >>= 1) raw_local += __shfl_down_sync(0xFFFFFFFFu, raw_local, o);
	if ((tid & 31) == 0 && raw_local) atomicAdd(P.counter, raw_local);
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

// finalize (counting mode): set bit (hash_h & mask), h = seed0 .. seed0 + NH - 1, for every listed word
template <int NH>
__global__ void __launch_bounds__(256)
insert_words_kernel(uint64_t* const* __restrict__ chunks, uint64_t n_words, uint32_t k, uint32_t* __restrict__ filter,
	uint32_t filter_mask, uint32_t win_id, uint32_t n_win, const uint32_t* __restrict__ invalid, uint32_t seed0)
{
	// the word list streams through once (evict first) and must not push the filter out of the L2, where the red.or
	// of a filter of up to 2^29 bits are served (evict last)
	const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
		// first-touch path: the list holds every occurrence; the few that read four non-zero counters are skipped
		if (invalid && ((invalid[i >> 5] >> (i & 31u)) & 1u)) continue;
		const uint64_t w = ld_nc_u64_hint(&chunks[i >> LIST_CHUNK_LOG2][i & (LIST_CHUNK - 1)], pol_stream);
		const uint64_t low = reverse_groups(w, k);
		uint32_t h[NH];
		murmur3_multi<NH>(low, k, h, seed0);
#pragma unroll
		for (int s = 0; s < NH; ++s) {
			const uint32_t bit = h[s] & filter_mask;
			if (n_win == 1 || (bit >> WINDOW_LOG2) == win_id) red_or_hint(filter + (bit >> 5), 1u << (bit & 31), pol_keep);
		}
	}
}

// finalize, first-touch path, num_hash 3 with seeds (0,1) folded out of the touched bitmap: seed 2's hash of every entry
// was kept by ft_hash_kernel behind the words of the list chunk (it is one of the four counting hashes), so the last
// seed costs one 4-byte load and one red.or per valid occurrence.  Four entries per thread.
__global__ void __launch_bounds__(256)
insert_hash_kernel(uint64_t* const* __restrict__ chunks, uint64_t n_words, uint32_t* __restrict__ filter,
	uint32_t filter_mask, uint32_t win_id, uint32_t n_win, const uint32_t* __restrict__ invalid)
{
	const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
	for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n_words; i += stride) {
		const uint32_t* h2 = reinterpret_cast<const uint32_t*>(chunks[i >> LIST_CHUNK_LOG2] + LIST_CHUNK);
		const uint4 v = ld_nc_v4_hint(h2 + (i & (LIST_CHUNK - 1)), pol_stream);      // (the chunk is allocated whole: entries beyond n_words are readable)
		const uint32_t inv = (invalid[i >> 5] >> (i & 31u)) & 0xFu;
		const uint32_t h[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			if (i + q < n_words && !((inv >> q) & 1u)) {
				const uint32_t bit = h[q] & filter_mask;
				if (n_win == 1 || (bit >> WINDOW_LOG2) == win_id) red_or_hint(filter + (bit >> 5), 1u << (bit & 31), pol_keep);
			}
		}
	}
}

// kwg_bloom_rollback: the boundary word of the invalid bitmap keeps the marks of the entries below the checkpoint
__global__ void mask_word_kernel(uint32_t* w, uint32_t keep) { *w &= keep; }

// finalize, min_kmer_count == 1: the filter bits of a PAIR of seeds are the fold of a counting table's touched bitmap.
//
// Seeds 2t and 2t + 1 index counting table t (make_bloom.cpp:546-551), and the same hash values masked to the filter
// length index the filter (make_bloom.cpp:573-577).  With min_kmer_count 1 a slot of table t is non-zero iff some
// occurrence touched it; the FIRST occurrence to touch a slot read a zero counter there, so it was valid and both of
// its table-t hashes are in the filter; an occurrence that is not valid found all its slots touched and adds nothing
// to the bitmap.  Hence { h_s & (2^L - 1) : valid occurrences, s in {2t, 2t+1} } = { slot & (2^L - 1) : touched slots
// of table t } whenever L <= lc and both seeds of the table are in use: 2^lc bits are read instead of two random
// read-modify-writes per k-mer occurrence.  A last odd seed (num_hash 3 or 5) is still set occurrence by occurrence
// (insert_hash_kernel / insert_words_kernel).  Filter word w = OR of the words w + j * 2^(L-5) of every table, j < 2^(lc-L).
// Checked against the oracle on the CPU (tests/test_fold_touched_cpu.py) and on the GPU over every (L, num_hash).
__device__ __forceinline__ uint32_t vec_or(uint32_t a, uint32_t b) { return a | b; }
__device__ __forceinline__ uint4 vec_or(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }

template <typename V>
__global__ void __launch_bounds__(256)
fold_touched_kernel(const V* __restrict__ touched, uint32_t n_tables, uint64_t table_vecs, uint64_t filter_vecs, V* __restrict__ filter)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= filter_vecs) return;
	V acc = touched[i];
	for (uint32_t t = 0; t < n_tables; ++t)
		for (uint64_t j = (t == 0) ? filter_vecs : 0; j < table_vecs; j += filter_vecs) acc = vec_or(acc, touched[t * table_vecs + j + i]);
	filter[i] = acc;
}

} // namespace kwg

using namespace kwg;

struct kwg_bloom {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool raw = false;
	uint32_t k = 0, min_count = 0, lc = 0, lmax = 0, raw_nh = 0, raw_L = 0;
	// counting mode
	CountGeom geom{};
	uint32_t* d_touched = nullptr;       // min_count == 1: 2^(lc+1) bits, slot touched by an earlier batch of this accession
	uint16_t* d_cnt = nullptr;           // min_count > 1: 2^(lc+1) 4-bit counters (the reference's two tables)
	uint32_t* d_elig = nullptr;  size_t elig_cap = 0;   // min_count > 1: eligibility bitmap of the current sub-batch
	uint64_t* d_dense = nullptr;         // min_count > 1: dense per-bucket copies of the records for the levels >= 1
	uint32_t* d_dense_len = nullptr;     // [n_buckets] extents, [1] number of buckets without a dense copy, [n_buckets] their list
	bool touched_dirty = false;          // false: nothing added since create/reset (the bitmap need not be read)
	bool no_fold = false;                // KWG_NO_FOLD=1: finalize sets every bit from the word list (A/B runs, tests of that path)
	std::vector<uint64_t*> chunks;
	uint64_t** d_chunk_table = nullptr;
	size_t table_cap = 0;
	// per-batch scratch of the partition pipeline (bloom_count.cuh), grown on demand
	uint64_t* d_rec1 = nullptr;  size_t rec1_cap = 0;
	uint64_t* d_rec2 = nullptr;  size_t rec2_cap = 0;
	uint16_t* d_offs1 = nullptr; size_t offs1_cap = 0;
	uint16_t* d_offs2 = nullptr; size_t offs2_cap = 0;
	uint32_t* d_cnt1 = nullptr;  size_t cnt1_cap = 0;
	uint64_t* d_base2 = nullptr; size_t base2_cap = 0;
	uint32_t* d_cbase = nullptr; size_t cbase_cap = 0;
	uint64_t* d_chunk_rec = nullptr; size_t chunk_rec_cap = 0;
	uint4* d_chunk_meta = nullptr; size_t chunk_meta_cap = 0;
	uint32_t* d_cfirst = nullptr; size_t cfirst_cap = 0;
	unsigned long long* d_tot_rec = nullptr; size_t tot_rec_cap = 0;
	uint32_t* d_tot_chk = nullptr; size_t tot_chk_cap = 0;
	// min_count == 1, lc <= 30: page chains of the one-level partition (bloom_first.cuh)
	bool use_ft = false;
	uint8_t* d_ft_pool = nullptr;  size_t ft_pool_cap = 0;
	uint32_t* d_ft_log = nullptr;  size_t ft_log_cap = 0;
	uint32_t* d_ft_plist = nullptr; size_t ft_plist_cap = 0;
	uint2* d_ft_info = nullptr;    size_t ft_info_cap = 0;
	uint32_t* d_ft_tiles = nullptr; size_t ft_tiles_cap = 0;
	uint4* d_ft_hm = nullptr;      size_t ft_hm_cap = 0;
	uint32_t* d_ft_meta = nullptr; size_t ft_meta_cap = 0;
	uint2* d_ft_carry = nullptr;   size_t ft_carry_cap = 0;
	uint32_t* d_inv = nullptr;     size_t inv_cap = 0;     // bit i: list entry i is an occurrence that read four non-zero counters
	uint64_t n_list_host = 0;            // entries of the word list (first-touch path: every k-mer occurrence, valid or not)
	uint64_t n_valid_host = 0;           // last value read from d_counter (saves a device round trip in finalize)
	bool n_valid_known = false;
	// host feed: bases stream in on a copy stream while the partition scan already runs on what has arrived
	cudaStream_t copy_stream = nullptr;
	std::vector<cudaEvent_t> feed_events;
	uint32_t* d_loss = nullptr;  size_t loss_cap = 0;
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
	uint16_t* d_bad = nullptr;    // packed input: not-a-base mask of the batch
	size_t bad_cap = 0;
	// kwg_bloom_checkpoint / kwg_bloom_rollback: the counting state before a batch that may take num_kmer beyond the limit
	uint8_t* d_ckpt = nullptr; size_t ckpt_cap = 0;
	struct { bool valid = false, touched_dirty = false; unsigned long long counter[3] = {0, 0, 0}; } ck;
	uint32_t* d_crc_ws = nullptr; size_t crc_ws_cap = 0;   // kwg_bloom_finalize_crc
	uint32_t* h_crc = nullptr;                             // pinned
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

// d_counter[0] = valid k-mers so far, d_counter[1] != 0: a 4-bit counter wrapped (min_kmer_count == 15 only),
// d_counter[2] = entries of the word list (first-touch path; elsewhere the list holds exactly the valid k-mers)
static int read_counter(kwg_bloom* b, uint64_t* out)
{
	if (b->n_valid_known) { *out = b->n_valid_host; return KWG_OK; }
	KWG_CUDA(cudaMemcpyAsync(b->h_counter, b->d_counter, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	if (b->h_counter[1])
		return fail(KWG_ERR_UNSUPPORTED, "a 4-bit counter of the counting filter wrapped from 15 to 0 (min_kmer_count 15 with both hashes of "
			"a table on one slot, reference make_bloom.cpp:586-592); this order-dependent case is not reproduced on the device");
	*out = *b->h_counter;
	b->n_valid_host = *out;
	b->n_list_host = b->use_ft ? b->h_counter[2] : *out;
	b->n_valid_known = true;
	return KWG_OK;
}

static int ensure_list_capacity(kwg_bloom* b, uint64_t words)
{
	const size_t need_chunks = (size_t)ceil_div(words, LIST_CHUNK);
	if (need_chunks <= b->chunks.size()) return KWG_OK;
	while (b->chunks.size() < need_chunks) {
		uint64_t* c = nullptr;
		// first-touch path: the chunk's words are followed by seed 2's hash of every entry (insert_hash_kernel)
		KWG_CUDA(cudaMalloc(&c, LIST_CHUNK * (sizeof(uint64_t) + (b->use_ft ? sizeof(uint32_t) : 0))));
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
	const uint64_t tiles = ceil_div(P.n_pos, TILE_BASES);
	if (tiles == 0) return KWG_OK;
	if (tiles > 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "batch too large");
	const dim3 grid((unsigned)tiles), block(SCAN_THREADS);
	b->timers.begin(MODE == MODE_PASS_B ? KWG_T_SCAN_B : KWG_T_SCAN_A, b->stream);
	if (MODE == MODE_RAW && b->k > KWG_MAX_KMER_LEN) {
		switch (b->raw_nh) {
#define KWG_CASE(N) case N: kmer_scan_wide_kernel<N><<<grid, block, 0, b->stream>>>(P); break;
			KWG_CASE(1) KWG_CASE(2) KWG_CASE(3) KWG_CASE(4) KWG_CASE(5) KWG_CASE(6) KWG_CASE(7) KWG_CASE(8)
#undef KWG_CASE
			default: return fail(KWG_ERR_INVALID_ARG, "num_hash out of range");
		}
	} else if (MODE == MODE_RAW) {
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

static size_t partition_smem_bytes()
{
	return (size_t)PT_REC * 8 + (size_t)(4 * PT_POS + PT_VEC + 2 + 2 * (PT_LOAD / 32 + 2) + PT_POS / 32 + MAX_FAN + 1 + MAX_FAN + PT_THREADS / 32) * 4;
}
static size_t regroup_smem_bytes()
{
	return (size_t)2 * STAGE_REC * 8 + (size_t)(CHUNK_REC + 2) * 8 + 32 + 2 * sizeof(RegroupUnit) +
	       (size_t)(2 * MAX_FAN + 2 + (MAX_FAN + 1) + MAX_FAN + 16 + 4) * 4;
}
static size_t resolve_smem_bytes(bool levels = false)
{
	return (size_t)FINAL_SLOTS * 4 + (size_t)STAGE_REC * 8 + 16 + (size_t)((levels ? FINAL_SLOTS / 8 : FINAL_SLOTS / 32) + 2 * RS_THREADS + 32 + 2) * 4 +
	       (size_t)RS_THREADS * 2;
}

static int count_kernels_init()
{
	KWG_CUDA(cudaFuncSetAttribute(partition_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)partition_smem_bytes()));
	KWG_CUDA(cudaFuncSetAttribute(regroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)regroup_smem_bytes()));
	KWG_CUDA(cudaFuncSetAttribute(resolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolve_smem_bytes()));
	KWG_CUDA(cudaFuncSetAttribute(resolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolve_smem_bytes(true)));
	KWG_CUDA(cudaFuncSetAttribute(resolve_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolve_smem_bytes(true)));
	KWG_CUDA(cudaFuncSetAttribute(ft_append_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_append_smem_bytes()));
	KWG_CUDA(cudaFuncSetAttribute(ft_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_resolve_smem_bytes()));
	return KWG_OK;
}

constexpr size_t FEED_CHUNK = 16u << 20;          // bases per piece of the host feed, at least
constexpr size_t FEED_MAX_PIECES = 4;             // ... and at most this many pieces: every piece costs a copy, two event calls and
                                                  // its launches, and the callers that care about throughput keep several
                                                  // accessions in flight (whose copies and kernels overlap anyway)
static inline size_t feed_chunk(uint64_t n_bases)
{
	return std::max<size_t>(FEED_CHUNK, (size_t)round_up(ceil_div(n_bases, FEED_MAX_PIECES), 1u << 20));
}

// Bases that are still on the host when the first kernel is launched: they are copied piece by piece on the copy
// stream and the scan is launched piece by piece behind them, so that the H2D copy hides behind the kernels.
struct HostFeed {
	const char* h_bases = nullptr;     // what goes to BaseSource::bases (ASCII: one byte per base, packed: four bases per byte)
	const uint16_t* h_bad = nullptr;   // packed: the not-a-base mask (NULL: none)
	uint32_t lead = 0;                 // packed: the first `lead` (< 16) bases belong to the batch before and must not count
};
struct FeedPiece { uint32_t tile0, n_tiles; size_t off, len; };   // tiles of the scan that may run once bases [off, off + len) have arrived

// pieces of about `chunk` bases; a tile of `pos` start positions reads `load` bases from its first position
static std::vector<FeedPiece> feed_plan(uint64_t n_bases, size_t chunk, uint32_t pos, uint32_t load, uint32_t n_tiles)
{
	std::vector<FeedPiece> pieces;
	const size_t n_chunks = (size_t)ceil_div(n_bases, chunk);
	uint32_t t_done = 0;
	for (size_t c = 0; c < n_chunks; ++c) {
		const size_t off = c * chunk, len = std::min<size_t>(chunk, (size_t)n_bases - off);
		const uint32_t t_end = (c + 1 == n_chunks) ? n_tiles
			: (uint32_t)std::min<uint64_t>(n_tiles, (off + len >= (size_t)load) ? (off + len - load) / pos + 1 : 0);
		pieces.push_back(FeedPiece{t_done, t_end > t_done ? t_end - t_done : 0u, off, len});
		t_done = std::max(t_done, t_end);
	}
	return pieces;
}

// packed input: the not-a-base mask of the batch (the caller's, or zeros) with the lead-in blanked
static int stage_mask(const BaseSource& S, const HostFeed& F, cudaStream_t st)
{
	if (!S.packed || !S.bad_mask) return KWG_OK;
	uint16_t* d = const_cast<uint16_t*>(S.bad_mask);
	const size_t bytes = (size_t)ceil_div(S.n_bases, 8);
	if (F.h_bad) KWG_CUDA(cudaMemcpyAsync(d, F.h_bad, bytes, cudaMemcpyHostToDevice, st));
	else KWG_CUDA(cudaMemsetAsync(d, 0, round_up(bytes, 2), st));
	if (F.lead) {
		// (two bytes from pageable memory: staged by the runtime before the call returns)
		const uint16_t w0 = (uint16_t)((F.h_bad ? F.h_bad[0] : 0u) | ((1u << F.lead) - 1u));
		KWG_CUDA(cudaMemcpyAsync(d, &w0, sizeof(w0), cudaMemcpyHostToDevice, st));
	}
	return KWG_OK;
}

static int feed_begin(kwg_bloom* b, size_t n_pieces, const BaseSource& S, const HostFeed& F)
{
	if (!b->copy_stream) KWG_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
	while (b->feed_events.size() < n_pieces + 1) {
		cudaEvent_t e;
		KWG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		b->feed_events.push_back(e);
	}
	// the staging buffer may still be read by work queued earlier on the compute stream
	KWG_CUDA(cudaEventRecord(b->feed_events[n_pieces], b->stream));
	KWG_CUDA(cudaStreamWaitEvent(b->copy_stream, b->feed_events[n_pieces], 0));
	return stage_mask(S, F, b->copy_stream);
}

// piece l travels; the compute stream waits for it
static int feed_piece(kwg_bloom* b, size_t l, const FeedPiece& L, const BaseSource& S, const HostFeed& F)
{
	size_t o = L.off, n = L.len;
	if (S.packed) { n = (size_t)ceil_div(o + n, 4) - o / 4; o /= 4; }      // (piece boundaries are multiples of 4 bases)
	KWG_CUDA(cudaMemcpyAsync(const_cast<char*>(S.bases) + o, F.h_bases + o, n, cudaMemcpyHostToDevice, b->copy_stream));
	KWG_CUDA(cudaEventRecord(b->feed_events[l], b->copy_stream));
	KWG_CUDA(cudaStreamWaitEvent(b->stream, b->feed_events[l], 0));
	return KWG_OK;
}

// min_kmer_count == 1, at most FT_MAX_BUCKETS buckets of 2^20 slots (lc <= 30): count / scan / hash by tile, append into
// page chains, resolve bucket by bucket against a 1-bit-per-slot tile, mark the invalid occurrences (bloom_first.cuh).
// The canonical words go straight into the accession's list (entries list_base + ordinal), there is no pass B.
static int count_first_touch(kwg_bloom* b, const ScanParams& S, uint64_t pos0, uint64_t n_pos, const HostFeed* h_feed, uint64_t list_base)
{
	const CountGeom& G = b->geom;
	const uint32_t slot_bits = G.lc + 1;
	const uint32_t n_buckets = slot_bits > (uint32_t)FT_BUCKET_LOG2 ? 1u << (slot_bits - FT_BUCKET_LOG2) : 1u;
	const uint32_t bucket_words = (slot_bits > (uint32_t)FT_BUCKET_LOG2 ? 1u << FT_BUCKET_LOG2 : 1u << slot_bits) / 32;
	const uint32_t sms = (uint32_t)sm_count(b->device);
	const uint32_t n_tiles = (uint32_t)ceil_div(n_pos, HT_POS);
	int rc;

	// The count / scan / hash passes run piece by piece behind the host feed (one piece when the bases are resident);
	// the append is ONE launch over all the ordinals of the sub-batch: every append block opens a chain per bucket, and
	// the resolver pays for every chain (a table entry, a padded last unit), so fewer and longer chains are cheaper
	// than an append that starts before the last piece has arrived.
	std::vector<FeedPiece> pieces;
	if (!h_feed) pieces.push_back(FeedPiece{0, n_tiles, 0, 0});
	else pieces = feed_plan(S.src.n_bases, feed_chunk(S.src.n_bases), HT_POS, HT_LOAD, n_tiles);
	const uint64_t ords = (uint64_t)n_tiles * HT_POS;                          // upper bound of the ordinals
	const uint32_t append_grid = (uint32_t)std::min<uint64_t>(sms, ceil_div(ords, FT_SUB));
	const uint32_t n_chains = append_grid;
	const uint64_t max_per_cta = std::max<uint64_t>(FT_SUB, round_up(ceil_div(ords, append_grid), FT_SUB));   // ordinals per append block, upper bound
	if (n_chains > (uint32_t)FT_MAX_CHAINS) return fail(KWG_ERR_CUDA, "internal: too many partition chains");
	const uint32_t max_chains = (uint32_t)round_up(std::max(n_chains, 1u), 32);

	// page size: about four pages per chain, 4 .. 32 units; pool per chain: exact worst case
	const uint64_t chain_units = max_per_cta * 4 / FT_UNIT / n_buckets;
	uint32_t pu_log2 = 2;
	while (pu_log2 < 5 && (8ull << pu_log2) <= chain_units) ++pu_log2;
	// a chain's pages are counted in 16 bits: even if one bucket took every record of a block's stretch
	while (pu_log2 < 5 && ceil_div(max_per_cta * 4, (uint64_t)FT_UNIT << pu_log2) + 1 > 0xFFFFull) ++pu_log2;
	if (ceil_div(max_per_cta * 4, (uint64_t)FT_UNIT << pu_log2) + 1 > 0xFFFFull) return fail(KWG_ERR_CUDA, "internal: partition stretch too long");
	const uint64_t ppc64 = ceil_div(max_per_cta * 4, (uint64_t)FT_UNIT << pu_log2) + 2ull * n_buckets + 2;
	if (ppc64 * max_chains >= 0xFFFFFFFFull || ppc64 >= (1ull << FT_SEQ_BITS)) return fail(KWG_ERR_CUDA, "internal: page pool too large");
	const uint32_t ppc = (uint32_t)ppc64;
	const size_t page_bytes = (size_t)FT_UNIT_BYTES << pu_log2;
	const size_t meta_words = 2;
	if ((rc = grow((void**)&b->d_ft_pool, &b->ft_pool_cap, (size_t)n_chains * ppc * page_bytes))) return rc;
	if ((rc = grow((void**)&b->d_ft_log, &b->ft_log_cap, (size_t)n_chains * ppc * sizeof(uint32_t)))) return rc;
	if ((rc = grow((void**)&b->d_ft_plist, &b->ft_plist_cap, (size_t)n_chains * ppc * sizeof(uint32_t)))) return rc;
	if ((rc = grow((void**)&b->d_ft_info, &b->ft_info_cap, (size_t)n_buckets * max_chains * sizeof(uint2)))) return rc;
	if ((rc = grow((void**)&b->d_ft_tiles, &b->ft_tiles_cap, (size_t)n_tiles * sizeof(uint32_t)))) return rc;
	if ((rc = grow((void**)&b->d_ft_hm, &b->ft_hm_cap, (size_t)n_tiles * HT_POS * sizeof(uint4)))) return rc;
	if ((rc = grow((void**)&b->d_ft_meta, &b->ft_meta_cap, meta_words * sizeof(uint32_t)))) return rc;
	if ((rc = grow((void**)&b->d_ft_carry, &b->ft_carry_cap, (size_t)sms * 2 * FR_CARRY * sizeof(uint2)))) return rc;
	KWG_CUDA(cudaMemsetAsync(b->d_ft_meta, 0, meta_words * sizeof(uint32_t), b->stream));

	FtTileParams T{};
	T.src = S.src; T.start_mask = S.start_mask; T.k = S.k;
	T.pos0 = pos0; T.n_pos = n_pos;
	T.tile_cnt = b->d_ft_tiles;
	T.count_mask = G.count_mask;
	T.hm = b->d_ft_hm;
	T.list_chunks = b->d_chunk_table;
	T.list_base = list_base;
	T.loss = b->d_loss;

	FtAppendParams A{};
	A.hm = b->d_ft_hm; A.meta = b->d_ft_meta;
	A.lc = G.lc; A.n_buckets = n_buckets;
	A.max_chains = max_chains; A.pu_log2 = pu_log2; A.ppc = ppc;
	A.pool = b->d_ft_pool; A.page_log = b->d_ft_log; A.plist = b->d_ft_plist; A.info = b->d_ft_info;

	if (h_feed && (rc = feed_begin(b, pieces.size(), S.src, *h_feed))) return rc;
	for (size_t l = 0; l < pieces.size(); ++l) {
		const FeedPiece& L = pieces[l];
		if (h_feed && (rc = feed_piece(b, l, L, S.src, *h_feed))) return rc;
		if (L.n_tiles == 0) continue;
		T.tile0 = L.tile0;
		b->timers.begin(KWG_T_SCAN_A, b->stream);
		ft_count_kernel<<<L.n_tiles, HT_THREADS, 0, b->stream>>>(T);
		KWG_LAUNCHED();
		ft_scan_kernel<<<1, 1024, 0, b->stream>>>(b->d_ft_tiles, L.tile0, L.n_tiles, b->d_ft_meta);
		KWG_LAUNCHED();
		ft_hash_kernel<<<L.n_tiles, HT_THREADS, 0, b->stream>>>(T);
		b->timers.end(b->stream);
		KWG_LAUNCHED();
	}
	b->timers.begin(KWG_T_REGROUP, b->stream);
	ft_append_kernel<<<append_grid, FT_THREADS, ft_append_smem_bytes(), b->stream>>>(A);
	b->timers.end(b->stream);
	KWG_LAUNCHED();

	FtResolveParams K3{};
	K3.pool = b->d_ft_pool; K3.plist = b->d_ft_plist; K3.info = b->d_ft_info;
	K3.n_chains = n_chains; K3.max_chains = max_chains; K3.n_buckets = n_buckets; K3.pu_log2 = pu_log2;
	K3.bucket_words = bucket_words;
	K3.touched = b->d_touched;
	K3.have_prior = b->touched_dirty ? 1u : 0u;
	K3.loss = b->d_loss;
	K3.carry_scratch = b->d_ft_carry;
	b->timers.begin(KWG_T_RESOLVE, b->stream);
	ft_resolve_kernel<<<std::min(n_buckets, sms), FR_THREADS, ft_resolve_smem_bytes(), b->stream>>>(K3);
	b->timers.end(b->stream);
	KWG_LAUNCHED();
	b->timers.begin(KWG_T_SCAN_B, b->stream);
	ft_finish_kernel<<<(unsigned)ceil_div(ceil_div(n_pos, 8), 256), 256, 0, b->stream>>>(b->d_loss, b->d_ft_meta, list_base, b->d_inv, b->d_counter);
	b->timers.end(b->stream);
	KWG_LAUNCHED();
	return KWG_OK;
}

// One sub-batch of the counting construction: start positions [pos0, pos0 + n_pos) of the batch in S.
// partition (K1) -> [regroup (K2)] -> resolve (K3) -> pass B (valid-word list).
// h_feed != NULL: the bases of this (single) sub-batch are still on the host at h_feed; they are copied in
// FEED_CHUNK pieces on the copy stream and the partition scan is launched piece by piece behind them.
static int count_sub_batch(kwg_bloom* b, const ScanParams& S, uint64_t pos0, uint64_t n_pos, const HostFeed* h_feed)
{
	const CountGeom& G = b->geom;
	const uint32_t F1 = 1u << G.f1_log2, F2 = 1u << G.f2_log2;
	const uint64_t n_tiles = ceil_div(n_pos, PT_POS);
	const uint32_t ntp = (uint32_t)round_up(n_tiles, 64);
	const bool two_level = G.f2_log2 != 0;
	int rc;

	uint64_t n_valid = 0;
	rc = read_counter(b, &n_valid);
	if (rc) return rc;
	rc = ensure_list_capacity(b, b->n_list_host + n_pos);
	if (rc) return rc;

	const size_t loss_words = (size_t)(round_up(n_pos, 8192) / 8);
	if ((rc = grow((void**)&b->d_loss, &b->loss_cap, loss_words * sizeof(uint32_t)))) return rc;
	KWG_CUDA(cudaMemsetAsync(b->d_loss, 0, loss_words * sizeof(uint32_t), b->stream));
	if (b->use_ft) {
		// the invalid bitmap grows with the list (kept contiguous: it is 1/64 of the list)
		const uint64_t list_base = b->n_list_host;
		const size_t inv_need = (size_t)round_up(ceil_div(list_base + n_pos, 8), 256);
		if (inv_need > b->inv_cap) {
			uint32_t* n = nullptr;
			const size_t want = round_up(inv_need + inv_need / 2, 256);
			KWG_CUDA(cudaMalloc(&n, want));
			KWG_CUDA(cudaMemsetAsync(n, 0, want, b->stream));
			if (b->d_inv) {
				KWG_CUDA(cudaMemcpyAsync(n, b->d_inv, b->inv_cap, cudaMemcpyDeviceToDevice, b->stream));
				KWG_CUDA(cudaStreamSynchronize(b->stream));
				KWG_CUDA(cudaFree(b->d_inv));
			}
			b->d_inv = n; b->inv_cap = want;
		}
		if ((rc = count_first_touch(b, S, pos0, n_pos, h_feed, list_base))) return rc;
		b->touched_dirty = true;
		b->n_valid_known = false;
		return KWG_OK;
	}
	if ((rc = grow((void**)&b->d_rec1, &b->rec1_cap, (size_t)n_tiles * PT_REC * sizeof(uint64_t)))) return rc;
	if ((rc = grow((void**)&b->d_offs1, &b->offs1_cap, (size_t)(F1 + 1) * ntp * sizeof(uint16_t)))) return rc;
	if (b->min_count > 1 && (rc = grow((void**)&b->d_elig, &b->elig_cap, loss_words / 4 * sizeof(uint32_t)))) return rc;

	PartParams K1{};
	K1.src = S.src; K1.start_mask = S.start_mask; K1.k = S.k;
	K1.pos0 = pos0; K1.n_pos = n_pos;
	K1.count_mask = G.count_mask;
	K1.f1_log2 = G.f1_log2;
	K1.shift1 = FINAL_LOG2 + G.f2_log2;
	K1.table_shift = G.f1_log2 - 1;
	K1.ntp = ntp;
	K1.rec1 = b->d_rec1;
	K1.offs1 = b->d_offs1;
	K1.tile0 = 0;
	b->timers.begin(KWG_T_SCAN_A, b->stream);
	if (!h_feed) {
		partition_scan_kernel<<<(unsigned)n_tiles, PT_THREADS, partition_smem_bytes(), b->stream>>>(K1);
		KWG_LAUNCHED();
	} else {
		const std::vector<FeedPiece> pieces = feed_plan(S.src.n_bases, feed_chunk(S.src.n_bases), PT_POS, PT_LOAD, (uint32_t)n_tiles);
		if ((rc = feed_begin(b, pieces.size(), S.src, *h_feed))) return rc;
		for (size_t l = 0; l < pieces.size(); ++l) {
			if ((rc = feed_piece(b, l, pieces[l], S.src, *h_feed))) return rc;
			if (pieces[l].n_tiles) {
				K1.tile0 = pieces[l].tile0;
				partition_scan_kernel<<<pieces[l].n_tiles, PT_THREADS, partition_smem_bytes(), b->stream>>>(K1);
				KWG_LAUNCHED();
			}
		}
	}
	b->timers.end(b->stream);

	ResolveParams K3{};
	K3.n_buckets = 1u << G.nb_log2;
	K3.touched = b->d_touched;
	K3.have_prior = b->touched_dirty ? 1u : 0u;
	K3.loss = b->d_loss;
	if (!two_level) {
		K3.rec = b->d_rec1;
		K3.offs = b->d_offs1;
		K3.cfirst = nullptr;
		K3.single_nci = (uint32_t)n_tiles;
		K3.chunk_rec = nullptr;
		K3.chunk_stride = PT_REC;
		K3.row_pitch = ntp;
		K3.f2_log2 = G.nb_log2;
	} else {
		const uint32_t G1 = F1;                                   // tiles per group: ~CHUNK_REC records per (bucket, group)
		const uint32_t NG = (uint32_t)ceil_div(n_tiles, G1);
		const size_t n_pairs = (size_t)F1 * NG;
		const size_t max_chunks = n_pairs + (size_t)n_tiles * PT_REC / CHUNK_REC + 1;
		if ((rc = grow((void**)&b->d_cnt1, &b->cnt1_cap, n_pairs * sizeof(uint32_t)))) return rc;
		if ((rc = grow((void**)&b->d_base2, &b->base2_cap, n_pairs * sizeof(uint64_t)))) return rc;
		if ((rc = grow((void**)&b->d_cbase, &b->cbase_cap, n_pairs * sizeof(uint32_t)))) return rc;
		if ((rc = grow((void**)&b->d_cfirst, &b->cfirst_cap, (size_t)(F1 + 1) * sizeof(uint32_t)))) return rc;
		if ((rc = grow((void**)&b->d_tot_rec, &b->tot_rec_cap, (size_t)F1 * sizeof(unsigned long long)))) return rc;
		if ((rc = grow((void**)&b->d_tot_chk, &b->tot_chk_cap, (size_t)F1 * sizeof(uint32_t)))) return rc;
		if ((rc = grow((void**)&b->d_chunk_rec, &b->chunk_rec_cap, max_chunks * sizeof(uint64_t)))) return rc;
		if ((rc = grow((void**)&b->d_chunk_meta, &b->chunk_meta_cap, max_chunks * sizeof(uint4)))) return rc;
		if ((rc = grow((void**)&b->d_offs2, &b->offs2_cap, max_chunks * (F2 + 1) * sizeof(uint16_t)))) return rc;
		if ((rc = grow((void**)&b->d_rec2, &b->rec2_cap, ((size_t)n_tiles * PT_REC + n_pairs + 2) * sizeof(uint64_t)))) return rc;

		b->timers.begin(KWG_T_REGROUP, b->stream);
		group_count_kernel<<<(unsigned)ceil_div(n_pairs, 8), 256, 0, b->stream>>>(b->d_offs1, ntp, (uint32_t)n_tiles, F1, G1, NG, b->d_cnt1);
		KWG_LAUNCHED();
		group_scan_kernel<<<F1, GS_THREADS, 0, b->stream>>>(b->d_cnt1, NG, b->d_base2, b->d_cbase, b->d_tot_rec, b->d_tot_chk);
		KWG_LAUNCHED();
		group_finish_kernel<<<F1, GS_THREADS, 0, b->stream>>>(b->d_cnt1, NG, F1, b->d_tot_rec, b->d_tot_chk, b->d_base2, b->d_cbase,
			b->d_chunk_rec, b->d_chunk_meta, b->d_cfirst);
		KWG_LAUNCHED();
		RegroupParams K2{};
		K2.rec1 = b->d_rec1; K2.offs1 = b->d_offs1; K2.ntp = ntp; K2.n_tiles = (uint32_t)n_tiles;
		K2.F1 = F1; K2.G1 = G1; K2.f2_log2 = G.f2_log2;
		K2.cfirst = b->d_cfirst;
		K2.chunk_rec = b->d_chunk_rec; K2.chunk_meta = b->d_chunk_meta;
		K2.rec2 = b->d_rec2; K2.offs2 = b->d_offs2;
		const unsigned ggrid = (unsigned)std::min<uint64_t>(max_chunks, (uint64_t)sm_count(b->device));
		regroup_kernel<<<ggrid, RG_BLOCK, regroup_smem_bytes(), b->stream>>>(K2);
		b->timers.end(b->stream);
		KWG_LAUNCHED();

		K3.rec = b->d_rec2;
		K3.offs = b->d_offs2;
		K3.cfirst = b->d_cfirst;
		K3.single_nci = 0;
		K3.chunk_rec = b->d_chunk_rec;
		K3.chunk_stride = 0;
		K3.row_pitch = 0;
		K3.f2_log2 = G.f2_log2;
	}
	const unsigned rgrid = (unsigned)std::min<uint64_t>(K3.n_buckets, (uint64_t)sm_count(b->device));
	const uint32_t* d_elig_final = nullptr;
	b->timers.begin(KWG_T_RESOLVE, b->stream);
	if (b->min_count == 1) {
		resolve_kernel<false><<<rgrid, RS_THREADS, resolve_smem_bytes(), b->stream>>>(K3);
		KWG_LAUNCHED();
	} else {
		// one launch per counter level; between two levels the eligibility bitmap is narrowed
		const uint64_t elig_words = n_tiles * PT_POS / 32;
		K3.cnt = b->d_cnt;
		K3.wrap_flag = reinterpret_cast<uint32_t*>(b->d_counter + 1);
		K3.dense = b->d_dense;
		K3.dense_len = b->d_dense_len;
		K3.n_not_dense = b->d_dense_len ? b->d_dense_len + K3.n_buckets : nullptr;
		K3.nd_list = b->d_dense_len ? b->d_dense_len + K3.n_buckets + 1 : nullptr;
		if (K3.dense) KWG_CUDA(cudaMemsetAsync(K3.n_not_dense, 0, sizeof(uint32_t), b->stream));
		for (uint32_t level = 0; level < b->min_count; ++level) {
			if (level) {
				elig_update_kernel<<<(unsigned)ceil_div(elig_words, 256), 256, 0, b->stream>>>(b->d_elig, b->d_loss, elig_words, level == 1);
				KWG_LAUNCHED();
			}
			K3.level = level;
			K3.elig = level ? b->d_elig : nullptr;
			if (level && K3.dense) {
				resolve_dense_kernel<<<rgrid, RS_THREADS, resolve_smem_bytes(true), b->stream>>>(K3);
				KWG_LAUNCHED();
			}
			resolve_kernel<true><<<rgrid, RS_THREADS, resolve_smem_bytes(true), b->stream>>>(K3);     // (level >= 1: only the buckets without a dense copy)
			KWG_LAUNCHED();
		}
		d_elig_final = b->min_count > 1 ? b->d_elig : nullptr;
	}
	b->timers.end(b->stream);

	b->touched_dirty = true;
	b->n_valid_known = false;

	ScanParams P = S;
	P.pos0 = pos0;
	P.n_pos = n_pos;
	P.loss = b->d_loss;
	P.elig = d_elig_final;
	P.wins = b->min_count > 1 ? 1u : 0u;
	P.list_chunks = b->d_chunk_table;
	return launch_scan<MODE_PASS_B>(b, P);
}

// All inputs on the device (or on their way: `feed`): src.bases 16-byte aligned, d_offsets[n_reads+1] with offsets
// relative to off0.  n_bases < 2^32 - 16.
static int add_batch_dev(kwg_bloom* b, const BaseSource& src, const uint64_t* d_offsets, uint64_t n_reads,
	uint64_t off0, const HostFeed* feed = nullptr)
{
	const uint64_t n_bases = src.n_bases;
	if (n_bases == 0 || n_reads == 0) return KWG_OK;
	if (n_bases >= 0xFFFFFFF0ull) return fail(KWG_ERR_INVALID_ARG, "a device batch must hold fewer than 2^32-16 bases");
	if (b->raw) b->n_valid_known = false;      // (counting mode invalidates after it has read the counter, below)
	if ((reinterpret_cast<uintptr_t>(src.bases) & 15u) != 0) return fail(KWG_ERR_INVALID_ARG, "d_bases must be 16-byte aligned");
	if ((reinterpret_cast<uintptr_t>(src.bad_mask) & 1u) != 0) return fail(KWG_ERR_INVALID_ARG, "the not-a-base mask must be 2-byte aligned");

	const size_t start_words = (size_t)(n_bases / 32 + 2);
	int rc = grow((void**)&b->d_start, &b->start_cap, start_words * sizeof(uint32_t));
	if (rc) return rc;
	KWG_CUDA(cudaMemsetAsync(b->d_start, 0, start_words * sizeof(uint32_t), b->stream));
	b->timers.begin(KWG_T_AUX, b->stream);
	mark_read_starts_kernel<<<(unsigned)ceil_div(n_reads, 256), 256, 0, b->stream>>>(d_offsets, n_reads, off0, n_bases, b->d_start);
	b->timers.end(b->stream);
	KWG_LAUNCHED();

	ScanParams P{};
	P.src = src;
	P.start_mask = b->d_start;
	P.k = b->k;
	P.counter = b->d_counter;
	P.pos0 = 0;
	P.n_pos = n_bases;

	if (b->raw) {
		P.filter = b->d_filter;
		P.filter_mask = (b->raw_L >= 32) ? 0xFFFFFFFFu : ((1u << b->raw_L) - 1u);
		P.n_win = (b->raw_L > WINDOW_LOG2) ? 1u << (b->raw_L - WINDOW_LOG2) : 1u;
		const uint32_t n_tiles = (uint32_t)ceil_div(n_bases, TILE_BASES);
		// the first pass over the bases runs piece by piece behind the host feed; further window passes find them in HBM
		std::vector<FeedPiece> pieces;
		if (feed) {
			pieces = feed_plan(n_bases, feed_chunk(n_bases), TILE_BASES, b->k > KWG_MAX_KMER_LEN ? WT_LOAD : TILE_LOAD, n_tiles);
			if ((rc = feed_begin(b, pieces.size(), src, *feed))) return rc;
		} else {
			pieces.push_back(FeedPiece{0, n_tiles, 0, (size_t)n_bases});
		}
		for (P.win_id = 0; P.win_id < P.n_win; ++P.win_id) {
			for (size_t l = 0; l < pieces.size(); ++l) {
				if (feed && P.win_id == 0 && (rc = feed_piece(b, l, pieces[l], src, *feed))) return rc;
				if (!pieces[l].n_tiles) continue;
				P.pos0 = (uint64_t)pieces[l].tile0 * TILE_BASES;
				P.n_pos = std::min<uint64_t>((uint64_t)pieces[l].n_tiles * TILE_BASES, n_bases - P.pos0);
				if ((rc = launch_scan<MODE_RAW>(b, P))) return rc;
			}
		}
		return KWG_OK;
	}

	// counting mode: sub-batches of just under 2^28 start positions (a record carries a 28-bit position)
	if (feed && n_bases > FT_MAX_POS) {
		// (several sub-batches: the bases go over in one piece first)
		if ((rc = feed_begin(b, 1, src, *feed))) return rc;
		if ((rc = feed_piece(b, 0, FeedPiece{0, 0, 0, (size_t)n_bases}, src, *feed))) return rc;
		feed = nullptr;
	}
	for (uint64_t pos0 = 0; pos0 < n_bases; pos0 += FT_MAX_POS) {
		const uint64_t n_pos = std::min<uint64_t>(FT_MAX_POS, n_bases - pos0);
		rc = count_sub_batch(b, P, pos0, n_pos, feed);
		if (rc) return rc;
	}
	return KWG_OK;
}

static int bloom_alloc_common(kwg_bloom* b)
{
	KWG_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	KWG_CUDA(cudaMalloc(&b->d_counter, 3 * sizeof(unsigned long long)));
	KWG_CUDA(cudaMallocHost(&b->h_counter, 3 * sizeof(unsigned long long)));
	KWG_CUDA(cudaMemsetAsync(b->d_counter, 0, 3 * sizeof(unsigned long long), b->stream));
	return KWG_OK;
}

extern "C" {

void kwg_bloom_destroy(kwg_bloom_t* b)
{
	if (!b) return;
	cudaSetDevice(b->device);
	if (b->stream) cudaStreamSynchronize(b->stream);
	cudaFree(b->d_touched); cudaFree(b->d_cnt); cudaFree(b->d_elig); cudaFree(b->d_dense); cudaFree(b->d_dense_len);
	cudaFree(b->d_rec1); cudaFree(b->d_rec2); cudaFree(b->d_offs1); cudaFree(b->d_offs2);
	cudaFree(b->d_cnt1); cudaFree(b->d_base2); cudaFree(b->d_cbase); cudaFree(b->d_chunk_rec); cudaFree(b->d_chunk_meta);
	cudaFree(b->d_cfirst); cudaFree(b->d_loss); cudaFree(b->d_tot_rec); cudaFree(b->d_tot_chk);
	cudaFree(b->d_ft_pool); cudaFree(b->d_ft_log); cudaFree(b->d_ft_plist); cudaFree(b->d_ft_info);
	cudaFree(b->d_ft_tiles); cudaFree(b->d_ft_hm); cudaFree(b->d_ft_meta); cudaFree(b->d_ft_carry); cudaFree(b->d_inv);
	for (cudaEvent_t e : b->feed_events) cudaEventDestroy(e);
	if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
	for (uint64_t* c : b->chunks) cudaFree(c);
	cudaFree(b->d_chunk_table);
	cudaFree(b->d_counter);
	if (b->h_counter) cudaFreeHost(b->h_counter);
	cudaFree(b->d_filter);
	cudaFree(b->d_crc_ws);
	cudaFree(b->d_ckpt);
	if (b->h_crc) cudaFreeHost(b->h_crc);
	cudaFree(b->d_bases);
	cudaFree(b->d_offsets);
	cudaFree(b->d_start);
	cudaFree(b->d_bad);
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
	int rc = select_device(device);
	if (rc) return rc;
	kwg_bloom* b = new kwg_bloom();
	b->device = device; b->raw = false; b->k = kmer_len; b->min_count = min_kmer_count;
	b->lc = log2_count_len; b->lmax = log2_max_len;
	rc = bloom_alloc_common(b);
	if (rc == KWG_OK) {
		b->geom = count_geometry(b->lc);
		// the one-level partition covers up to 2048 buckets of 2^20 slots; KWG_COUNT_TWO_LEVEL=1 keeps the first design (A/B runs)
		b->use_ft = min_kmer_count == 1 && b->lc + 1 <= (uint32_t)FT_BUCKET_LOG2 + 11 && !getenv("KWG_COUNT_TWO_LEVEL");
		b->no_fold = getenv("KWG_NO_FOLD") != nullptr;
		rc = count_kernels_init();
		if (rc == KWG_OK) {
			// two tables of 2^lc slots: one bit each (min count 1: the first batch after create/reset writes every word
			// without reading it) or the reference's 4-bit counters (cleared here and by reset)
			const size_t bytes = (size_t)1 << (b->lc + 1 - (min_kmer_count == 1 ? 3 : 1));
			cudaError_t e = (min_kmer_count == 1) ? cudaMalloc(&b->d_touched, bytes) : cudaMalloc(&b->d_cnt, bytes);
			if (e != cudaSuccess) rc = fail(KWG_ERR_NO_MEMORY, std::string("counting-filter state: ") + cudaGetErrorString(e));
			else if (b->d_cnt && cudaMemsetAsync(b->d_cnt, 0, bytes, b->stream) != cudaSuccess) rc = fail(KWG_ERR_CUDA, "memset of the counting filters failed");
			if (rc == KWG_OK && min_kmer_count > 1 && !getenv("KWG_NO_DENSE")) {
				// one staging window per final bucket (4.5 GiB at lc = 30); optional: without it every level gathers again
				const size_t nb = (size_t)1 << b->geom.nb_log2;
				if (cudaMalloc(&b->d_dense, nb * STAGE_REC * sizeof(uint64_t)) != cudaSuccess ||
				    cudaMalloc(&b->d_dense_len, (2 * nb + 1) * sizeof(uint32_t)) != cudaSuccess) {
					cudaGetLastError();
					cudaFree(b->d_dense); cudaFree(b->d_dense_len);
					b->d_dense = nullptr; b->d_dense_len = nullptr;
				}
			}
		}
	}
	if (rc) { kwg_bloom_destroy(b); return rc; }
	*out = b;
	return KWG_OK;
}

int kwg_bloom_create_raw(kwg_bloom_t** out, int device, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len)
{
	if (!out) return fail(KWG_ERR_INVALID_ARG, "out is NULL");
	*out = nullptr;
	if (kmer_len < 1 || kmer_len > KWG_MAX_KMER_LEN_RAW)
		return fail(KWG_ERR_INVALID_ARG, "kmer_len must be in [1,63] in raw mode (the reference stops at 32, word.h:10; 33..63 is an extension)");
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
	KWG_CUDA(cudaMemsetAsync(b->d_counter, 0, 3 * sizeof(unsigned long long), b->stream));
	b->n_valid_host = 0;
	b->n_list_host = 0;
	b->n_valid_known = true;
	if (b->d_inv) KWG_CUDA(cudaMemsetAsync(b->d_inv, 0, b->inv_cap, b->stream));
	if (b->raw) {
		KWG_CUDA(cudaMemsetAsync(b->d_filter, 0, (size_t)1 << (b->raw_L - 3), b->stream));
	} else {
		b->touched_dirty = false;            // the next batch rewrites every word of the touched bitmap
		if (b->d_cnt) KWG_CUDA(cudaMemsetAsync(b->d_cnt, 0, (size_t)1 << b->lc, b->stream));
	}
	return KWG_OK;
}

// bytes of the state that an added batch changes for good: the touched bitmap (min_kmer_count 1) or the 4-bit counters
static size_t counting_state_bytes(const kwg_bloom* b)
{
	return b->min_count == 1 ? (size_t)1 << (b->lc + 1 - 3) : (size_t)1 << b->lc;
}

int kwg_bloom_checkpoint(kwg_bloom_t* b)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (b->raw) return fail(KWG_ERR_UNSUPPORTED, "raw mode has no k-mer limit to stop at (reference rig bloom_test.cpp)");
	int rc = select_device(b->device);
	if (rc) return rc;
	const size_t bytes = counting_state_bytes(b);
	b->ck.valid = false;
	if ((rc = grow((void**)&b->d_ckpt, &b->ckpt_cap, bytes))) return rc;
	// (min_kmer_count 1 before the first batch: the bitmap holds nothing yet and need not be kept)
	if (b->min_count > 1 || b->touched_dirty)
		KWG_CUDA(cudaMemcpyAsync(b->d_ckpt, b->min_count == 1 ? (const void*)b->d_touched : (const void*)b->d_cnt, bytes, cudaMemcpyDeviceToDevice, b->stream));
	KWG_CUDA(cudaMemcpyAsync(b->h_counter, b->d_counter, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	for (int i = 0; i < 3; ++i) b->ck.counter[i] = b->h_counter[i];
	b->ck.touched_dirty = b->touched_dirty;
	b->ck.valid = true;
	return KWG_OK;
}

int kwg_bloom_rollback(kwg_bloom_t* b)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (b->raw || !b->ck.valid) return fail(KWG_ERR_INVALID_ARG, "no checkpoint to go back to");
	int rc = select_device(b->device);
	if (rc) return rc;
	const size_t bytes = counting_state_bytes(b);
	if (b->min_count > 1 || b->ck.touched_dirty)
		KWG_CUDA(cudaMemcpyAsync(b->min_count == 1 ? (void*)b->d_touched : (void*)b->d_cnt, b->d_ckpt, bytes, cudaMemcpyDeviceToDevice, b->stream));
	b->touched_dirty = b->ck.touched_dirty;
	for (int i = 0; i < 3; ++i) b->h_counter[i] = b->ck.counter[i];
	KWG_CUDA(cudaMemcpyAsync(b->d_counter, b->h_counter, 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice, b->stream));
	b->n_valid_host = b->ck.counter[0];
	b->n_list_host = b->use_ft ? b->ck.counter[2] : b->ck.counter[0];
	b->n_valid_known = true;
	if (b->d_inv) {
		// first-touch path: the occurrences listed since the checkpoint are forgotten, and so are their invalid marks
		const uint64_t w0 = b->n_list_host >> 5;
		const uint32_t keep = (1u << (b->n_list_host & 31u)) - 1u;
		if ((w0 + 1) * 4 <= b->inv_cap) {
			mask_word_kernel<<<1, 1, 0, b->stream>>>(b->d_inv + w0, keep);
			KWG_LAUNCHED();
			KWG_CUDA(cudaMemsetAsync(b->d_inv + w0 + 1, 0, b->inv_cap - (w0 + 1) * 4, b->stream));
		}
	}
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	return KWG_OK;
}

int kwg_bloom_add_reads_dev(kwg_bloom_t* b, const char* d_bases, const uint64_t* d_offsets, uint64_t n_reads, uint64_t n_bases)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (n_reads == 0 || n_bases == 0) return KWG_OK;
	if (!d_bases || !d_offsets) return fail(KWG_ERR_INVALID_ARG, "NULL input");
	int rc = select_device(b->device);
	if (rc) return rc;
	return add_batch_dev(b, BaseSource{d_bases, nullptr, n_bases, 0u}, d_offsets, n_reads, 0);
}

int kwg_bloom_add_packed_dev(kwg_bloom_t* b, const uint8_t* d_packed, const uint8_t* d_bad_mask, const uint64_t* d_offsets,
	uint64_t n_reads, uint64_t n_bases)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (n_reads == 0 || n_bases == 0) return KWG_OK;
	if (!d_packed || !d_offsets) return fail(KWG_ERR_INVALID_ARG, "NULL input");
	int rc = select_device(b->device);
	if (rc) return rc;
	return add_batch_dev(b, BaseSource{reinterpret_cast<const char*>(d_packed), reinterpret_cast<const uint16_t*>(d_bad_mask), n_bases, 1u},
		d_offsets, n_reads, 0);
}

// Host batches: whole reads, at most MAX_BATCH_BASES bases each (at least one read).  ASCII (`packed` false: `bases` holds one
// byte per base) or 2na (`bases` holds four bases per byte from base 0 of the call, `bad` the optional mask).
static int add_reads_host(kwg_bloom* b, const char* bases, const uint16_t* bad, bool packed, const uint64_t* offsets, uint64_t n_reads)
{
	int rc;
	{
		uint64_t wrong = 0;                    // branch-free so that the compiler vectorises the scan
		for (uint64_t r = 0; r < n_reads; ++r) wrong |= (uint64_t)(offsets[r + 1] < offsets[r]);
		if (wrong) return fail(KWG_ERR_INVALID_ARG, "offsets must be non-decreasing");
	}
	uint64_t r0 = 0;
	while (r0 < n_reads) {
		uint64_t r1;
		if (offsets[n_reads] - offsets[r0] <= MAX_BATCH_BASES) {
			r1 = n_reads;
		} else {
			r1 = (uint64_t)(std::upper_bound(offsets + r0 + 1, offsets + n_reads + 1, offsets[r0] + MAX_BATCH_BASES) - offsets) - 1;
			if (r1 <= r0) r1 = r0 + 1;
		}
		// a packed batch starts on a 16-bit word of the mask (hence on a byte of the 2na stream): the bases between that
		// boundary and the first read of the batch ride along and are blanked through the mask
		const uint64_t first = offsets[r0];
		const uint64_t off0 = packed ? (first & ~(uint64_t)15) : first;
		const uint32_t lead = (uint32_t)(first - off0);
		const uint64_t nb = offsets[r1] - off0;
		if (nb >= 0xFFFFFFF0ull) return fail(KWG_ERR_INVALID_ARG, "a single read of 2^32 bases or more is not supported");
		if (offsets[r1] > first) {
			const size_t base_bytes = packed ? (size_t)ceil_div(nb, 4) : (size_t)nb;
			const bool mask = packed && (bad || lead);
			rc = grow((void**)&b->d_bases, &b->bases_cap, round_up(base_bytes, 16) + 16);
			if (rc) return rc;
			rc = grow((void**)&b->d_offsets, &b->offsets_cap, (r1 - r0 + 1) * sizeof(uint64_t));
			if (rc) return rc;
			if (mask && (rc = grow((void**)&b->d_bad, &b->bad_cap, round_up(ceil_div(nb, 8), 16) + 16))) return rc;
			KWG_CUDA(cudaMemcpyAsync(b->d_offsets, offsets + r0, (r1 - r0 + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
			const BaseSource src{b->d_bases, mask ? b->d_bad : nullptr, nb, packed ? 1u : 0u};
			const HostFeed feed{bases + (packed ? off0 / 4 : off0), (packed && bad) ? bad + off0 / 16 : nullptr, lead};
			// batches of more than one piece are fed piece by piece behind the scan; small ones go over in one copy
			const bool chunked = nb > FEED_CHUNK;
			if (!chunked) {
				KWG_CUDA(cudaMemcpyAsync(b->d_bases, feed.h_bases, base_bytes, cudaMemcpyHostToDevice, b->stream));
				if ((rc = stage_mask(src, feed, b->stream))) return rc;
			}
			rc = add_batch_dev(b, src, b->d_offsets, r1 - r0, off0, chunked ? &feed : nullptr);
			if (rc) return rc;
		}
		r0 = r1;
	}
	// host buffers may be reused by the caller as soon as we return
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	return KWG_OK;
}

int kwg_bloom_add_reads(kwg_bloom_t* b, const char* bases, const uint64_t* offsets, uint64_t n_reads)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (n_reads == 0) return KWG_OK;
	if (!bases || !offsets) return fail(KWG_ERR_INVALID_ARG, "NULL input");
	int rc = select_device(b->device);
	if (rc) return rc;
	return add_reads_host(b, bases, nullptr, false, offsets, n_reads);
}

int kwg_bloom_add_packed(kwg_bloom_t* b, const uint8_t* packed, const uint8_t* bad_mask, const uint64_t* offsets, uint64_t n_reads)
{
	if (!b) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (n_reads == 0) return KWG_OK;
	if (!packed || !offsets) return fail(KWG_ERR_INVALID_ARG, "NULL input");
	if ((reinterpret_cast<uintptr_t>(bad_mask) & 1u) != 0) return fail(KWG_ERR_INVALID_ARG, "bad_mask must be 2-byte aligned");
	int rc = select_device(b->device);
	if (rc) return rc;
	return add_reads_host(b, reinterpret_cast<const char*>(packed), reinterpret_cast<const uint16_t*>(bad_mask), true, offsets, n_reads);
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
	const uint64_t n_list = b->n_list_host;
	uint32_t* f = reinterpret_cast<uint32_t*>(d_out_bits);
	// seeds [0, n_fold) come out of the touched bitmap (fold_touched_kernel), the rest out of the word list
	uint32_t n_fold = 0;
	if (b->use_ft && b->touched_dirty && n_list && log2_len <= b->lc && !b->no_fold) n_fold = num_hash & ~1u;
	b->timers.begin(KWG_T_INSERT, b->stream);
	if (n_fold) {
		const uint64_t table_words = (uint64_t)1 << (b->lc - 5), filter_words = (uint64_t)1 << (log2_len - 5);
		if (filter_words >= 4 && (reinterpret_cast<uintptr_t>(d_out_bits) & 15u) == 0) {
			const uint64_t fv = filter_words / 4;
			fold_touched_kernel<uint4><<<(unsigned)ceil_div(fv, 256), 256, 0, b->stream>>>(reinterpret_cast<const uint4*>(b->d_touched),
				n_fold / 2, table_words / 4, fv, reinterpret_cast<uint4*>(f));
		} else {
			fold_touched_kernel<uint32_t><<<(unsigned)ceil_div(filter_words, 256), 256, 0, b->stream>>>(b->d_touched,
				n_fold / 2, table_words, filter_words, f);
		}
		KWG_LAUNCHED();
	} else {
		KWG_CUDA(cudaMemsetAsync(d_out_bits, 0, bytes, b->stream));
	}
	if (n_list && n_fold < num_hash) {
		const uint32_t mask = (log2_len >= 32) ? 0xFFFFFFFFu : ((1u << log2_len) - 1u);
		const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(n_list, 256), (uint64_t)sm_count(b->device) * 16);
#define KWG_CASE(N) case N: insert_words_kernel<N><<<grid, 256, 0, b->stream>>>(b->d_chunk_table, n_list, b->k, f, mask, w, n_win, b->use_ft ? b->d_inv : nullptr, n_fold); break;
		const uint32_t n_win = (log2_len > WINDOW_LOG2) ? 1u << (log2_len - WINDOW_LOG2) : 1u;
		const bool kept_hash = b->use_ft && n_fold == 2 && num_hash == 3;      // seed 2 of every entry lies behind the words
		for (uint32_t w = 0; w < n_win; ++w) {
			if (kept_hash) {
				const unsigned hgrid = (unsigned)std::min<uint64_t>(ceil_div(n_list, 1024), (uint64_t)sm_count(b->device) * 16);
				insert_hash_kernel<<<hgrid, 256, 0, b->stream>>>(b->d_chunk_table, n_list, f, mask, w, n_win, b->d_inv);
			} else switch (num_hash - n_fold) {
				KWG_CASE(1) KWG_CASE(2) KWG_CASE(3) KWG_CASE(4) KWG_CASE(5)
			}
			KWG_LAUNCHED();
		}
#undef KWG_CASE
	}
	b->timers.end(b->stream);
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

int kwg_bloom_finalize_crc(kwg_bloom_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* out_bits, uint32_t* crc32)
{
	if (!b || !out_bits || !crc32) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(b->device);
	if (rc) return rc;
	if (log2_len < 5 || log2_len > 32) return fail(KWG_ERR_INVALID_ARG, "log2_len must be in [5,32]");
	const size_t bytes = (size_t)1 << (log2_len - 3);
	const uint8_t* d_bits = nullptr;
	if (b->raw) {
		if (log2_len != b->raw_L || num_hash != b->raw_nh) return fail(KWG_ERR_INVALID_ARG, "raw mode: parameters differ from creation");
		d_bits = reinterpret_cast<const uint8_t*>(b->d_filter);
	} else {
		rc = grow((void**)&b->d_filter, &b->filter_cap, bytes);
		if (rc) return rc;
		rc = kwg_bloom_finalize_dev(b, log2_len, num_hash, reinterpret_cast<uint8_t*>(b->d_filter));
		if (rc) return rc;
		d_bits = reinterpret_cast<const uint8_t*>(b->d_filter);
	}
	const size_t ws_words = crc32_workspace_words(1, bytes) + 1;
	rc = grow((void**)&b->d_crc_ws, &b->crc_ws_cap, ws_words * sizeof(uint32_t));
	if (rc) return rc;
	if (!b->h_crc) KWG_CUDA(cudaMallocHost(&b->h_crc, sizeof(uint32_t)));
	// the checksum kernels read the filter while it is still warm in L2; the 4-byte result rides behind the bits
	b->timers.begin(KWG_T_AUX, b->stream);
	rc = crc32_launch(b->device, d_bits, 1, 0, 1, bytes, bytes, nullptr, b->d_crc_ws + ws_words - 1, b->d_crc_ws, b->stream);
	b->timers.end(b->stream);
	if (rc) return rc;
	KWG_CUDA(cudaMemcpyAsync(out_bits, d_bits, bytes, cudaMemcpyDeviceToHost, b->stream));
	KWG_CUDA(cudaMemcpyAsync(b->h_crc, b->d_crc_ws + ws_words - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
	KWG_CUDA(cudaStreamSynchronize(b->stream));
	*crc32 = *b->h_crc;
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
