// Bit-sliced search (reference kwage.cpp:340-541) on sm_100a over an HBM-resident slice slab.
//
//   query_kmers_kernel : ASCII query -> canonical k-mers -> de-duplicated with a hash set (the
//                        reference sorts + std::unique, kwage.cpp:362-366; only the SET matters for
//                        counts) -> murmur3 row indices for every unique k-mer (kwage.cpp:411-412)
//   search_count_kernel: THE hot kernel.  For every unique query k-mer gather its num_hash slices
//                        (rows) with 128-bit loads, AND them, and accumulate the per-filter match
//                        bits in bit-sliced carry-save counters (Harley-Seal, 16 k-mers per block)
//                        instead of the reference's per-bit increment_count (bloom.h:291-330).
//                        One thread owns 128 filter columns; the counters are expanded to the
//                        reference's uint32 counts once per (query, column chunk).
//   hit kernels        : threshold with the reference's float arithmetic (kwage.cpp:349,388,
//                        489-503) and compact (query, filter, num_match) in (query, filter) order.
#include "common.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <mutex>

#include <dlfcn.h>
#include <nccl.h>      // types only: the library is loaded with dlopen

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace kwg {

constexpr uint64_t SET_EMPTY = ~0ull;      // never a canonical word: T^32 is not canonical (A^32 is)
constexpr int QK_THREADS = 128;

// ------------------------------------------------------------------------------------------
// query -> unique canonical k-mers -> row indices
// ------------------------------------------------------------------------------------------
constexpr int QK_TILE = 2048;                       // k-mer start positions per staged tile
constexpr int QK_LOAD = QK_TILE + 32;               // bytes staged per tile (halo >= k - 1)

// Queries of up to QK_SMEM_SLOTS / 2 bases (the BIGSI use case: genes) keep their hash set in shared memory; longer
// ones use their region of the table in HBM, which this kernel clears for them (and only for them).
constexpr int QK_SMEM_SLOTS = 4096;

__global__ void __launch_bounds__(256)
query_table_init_kernel(const uint64_t* __restrict__ offsets, uint64_t* __restrict__ table)
{
	const uint32_t q = blockIdx.x;
	const uint64_t o0 = offsets[q] - offsets[0], len = offsets[q + 1] - offsets[q];
	if (2 * len <= (uint64_t)QK_SMEM_SLOTS) return;
	uint64_t* tab = table + 2 * o0;
	for (uint64_t i = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; i < 2 * len; i += (uint64_t)gridDim.y * blockDim.x) tab[i] = SET_EMPTY;
}

template <int NH>
__global__ void __launch_bounds__(QK_THREADS)
query_kmers_kernel(const char* __restrict__ bases, uint64_t n_bases_total, const uint64_t* __restrict__ offsets, uint32_t k, uint32_t filter_mask,
	uint64_t* __restrict__ table,       // 2 slots per base: query q owns table[2*o0 .. 2*o1)
	uint64_t* __restrict__ kmers,       // unique words of query q at kmers[o0 ..)
	uint32_t* __restrict__ rows,        // row indices at rows[(o0 + i) * NH + h]
	uint32_t* __restrict__ n_kmers)     // per query, zero-initialised
{
	// same front end as the construction kernels: the tile is staged once, 16 bases at a time become 32 bits
	// of 2-bit codes + 16 "not ACGT" flags (encode16), and every start position extracts its window from
	// shared memory (no per-position byte loop)
	__shared__ uint32_t s_w[QK_LOAD / 4 + 8];          // the aligned 32-bit words that cover the tile
	__shared__ uint32_t s_codes[QK_LOAD / 16 + 2];
	__shared__ uint32_t s_bad[QK_LOAD / 32 + 2];
	__shared__ uint32_t s_nostart[QK_LOAD / 32 + 2];
	__shared__ unsigned long long s_set[QK_SMEM_SLOTS];

	const uint32_t q = blockIdx.x, tid = threadIdx.x;
	const uint64_t o0 = offsets[q] - offsets[0], o1 = offsets[q + 1] - offsets[0];
	const uint64_t len = o1 - o0;
	if (len < k) return;
	const uint64_t n_pos = len - k + 1;
	const uint64_t tsize = 2 * len;
	// a short query is one block's work (its set lives in this block's shared memory): the other blocks of its row leave
	const bool in_smem = tsize <= (uint64_t)QK_SMEM_SLOTS;
	if (in_smem && blockIdx.y != 0) return;
	unsigned long long* tab = in_smem ? s_set : reinterpret_cast<unsigned long long*>(table + 2 * o0);
	const uint32_t* gw = reinterpret_cast<const uint32_t*>(bases);
	const uint64_t n_words_total = (n_bases_total + 3) >> 2;
	for (uint32_t v = tid; v < (uint32_t)(QK_LOAD / 32 + 2); v += QK_THREADS) s_nostart[v] = 0u;
	if (in_smem) for (uint32_t v = tid; v < (uint32_t)tsize; v += QK_THREADS) s_set[v] = SET_EMPTY;
	const uint64_t tile_first = in_smem ? 0 : (uint64_t)blockIdx.y * QK_TILE, tile_step = in_smem ? QK_TILE : (uint64_t)gridDim.y * QK_TILE;

	for (uint64_t tile0 = tile_first; tile0 < n_pos; tile0 += tile_step) {
		__syncthreads();                                    // previous tile's readers are done
		const uint64_t avail = len - tile0;
		// A query starts at any byte: the tile is fetched as whole aligned words (coalesced, all in flight together -- a
		// byte-wise copy is a chain of dependent round trips) and re-aligned with funnel shifts on the way to encode16.
		const uint64_t g0 = o0 + tile0, w0 = g0 >> 2;
		const uint32_t sh = 8u * (uint32_t)(g0 & 3u);
		for (uint32_t i = tid; i < (uint32_t)(QK_LOAD / 4 + 1); i += QK_THREADS) s_w[i] = (w0 + i < n_words_total) ? ld_nc_u32(gw + w0 + i) : 0u;
		__syncthreads();
		for (uint32_t v = tid; v < (uint32_t)(QK_LOAD / 16); v += QK_THREADS) {
			uint32_t codes, bad16;
			const uint32_t* w = s_w + 4 * v;
			encode16(make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh),
				__funnelshift_r(w[3], w[4], sh)), codes, bad16);
			// bases at and beyond the end of the query read as separators
			const uint64_t first = 16ull * v;
			if (first >= avail) bad16 = 0xFFFFu;
			else if (first + 16 > avail) bad16 |= (0xFFFFu << (uint32_t)(avail - first)) & 0xFFFFu;
			s_codes[v] = codes;
			reinterpret_cast<uint16_t*>(s_bad)[v] = (uint16_t)bad16;
		}
		if (tid == 0) {
			s_codes[QK_LOAD / 16] = 0; s_codes[QK_LOAD / 16 + 1] = 0;
			s_bad[QK_LOAD / 32] = 0xFFFFFFFFu; s_bad[QK_LOAD / 32 + 1] = 0xFFFFFFFFu;
		}
		__syncthreads();
		// One start position per thread and iteration; the trip count is the same for every lane, so that the warp can
		// vote below.  The probing loop only decides WHO inserted a new k-mer: lanes leave it after different numbers of
		// probes, and anything heavy inside it (the hashes) would run once per probe with a few lanes each.  New k-mers of
		// a warp take their list places with one counter update, then every inserting lane hashes its k-mer once.
		const uint64_t left = n_pos - tile0;
		const uint32_t n_here = (uint32_t)(left < (uint64_t)QK_TILE ? left : (uint64_t)QK_TILE);
		const uint32_t lane = tid & 31u;
		for (uint32_t p0 = 0; p0 < n_here; p0 += QK_THREADS) {
			const uint32_t p = p0 + tid;
			bool inserted = false;
			Canon c;
			c.word = 0; c.low = 0;
			if (p < n_here && window_ok(s_bad, s_nostart, p, k)) {
				c = canonical(window_sense(s_codes, p, k), k);
				uint64_t slot = __umul64hi(mix64(c.word), tsize);      // (a 64-bit remainder costs more than the rest of the loop)
				for (;;) {
					const unsigned long long old = atomicCAS(tab + slot, (unsigned long long)SET_EMPTY, (unsigned long long)c.word);
					if (old == SET_EMPTY) { inserted = true; break; }
					if (old == c.word) break;
					slot = (slot + 1 == tsize) ? 0 : slot + 1;
				}
			}
			const uint32_t m = __ballot_sync(0xFFFFFFFFu, inserted);
			if (m) {
				const uint32_t leader = (uint32_t)__ffs(m) - 1u;
				uint32_t base = 0;
				if (lane == leader) base = atomicAdd(n_kmers + q, (uint32_t)__popc(m));
				base = __shfl_sync(0xFFFFFFFFu, base, leader);
				if (inserted) {
					const uint32_t i = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
					kmers[o0 + i] = c.word;
					uint32_t h[NH];
					murmur3_multi<NH>(c.low, k, h);
#pragma unroll
					for (int t = 0; t < NH; ++t) rows[(o0 + i) * NH + t] = h[t] & filter_mask;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// gather + AND + bit-sliced count
// ------------------------------------------------------------------------------------------
constexpr int SC_WARPS = 8;
constexpr int SC_THREADS = SC_WARPS * 32;
constexpr int SC_LOW = 4;                  // ones, twos, fours, eights
constexpr int SC_UP = 6;                   // 16s .. 512s  -> < 1024 k-mers per substream per segment (SC_SUB_CAP, bitops.cuh)
constexpr int SC_PLANES = SC_LOW + SC_UP;
constexpr int SC_TOT = 16;                 // planes of the merged per-segment total (< 65536)

struct SearchParams {
	const uint8_t* slab;          // rows of row_pitch bytes
	uint64_t row_pitch;
	uint32_t vec_per_row;         // row_pitch / 16
	uint32_t lanes_per_row;       // power of two <= 32
	uint32_t n_chunks;            // ceil(vec_per_row / lanes_per_row)
	const uint64_t* offsets;      // per query, base offsets (device)
	const uint32_t* rows;         // [(o0 + i) * NH + h]
	const uint32_t* n_kmers;      // per query
	uint32_t* counts;             // [q * count_pitch + f]
	uint64_t count_pitch;         // multiple of 4
	float threshold;              // EXIT variant: the match threshold of the call (kwage.cpp:349,388)
};

// EXIT: the caller only wants the filters that reach the threshold (kwg_search*), so a block may stop reading rows as
// soon as no column of its chunk can reach it any more -- the reference's own early exit (kwage.cpp:397,459-482: "even
// the best matching Bloom filter does not have enough matches"), taken per 4096-column chunk instead of per file.
// Every substream publishes an upper bound of what it can still contribute to ANY column: the exact maximum of its
// partial counts over its columns (a descent through the bit planes) plus the k-mers it has not looked at yet.  The
// bounds only fall, so a stale one is still a bound and no barrier is needed; when their sum is below the count a hit
// needs, the warp leaves the loop.  The counts of such a chunk are partial -- and below the threshold, which is all
// hits_kernel asks.  A chunk that holds a hit never stops: its bound is at least the hit's final count.
template <int NH, bool EXIT>
__global__ void __launch_bounds__(SC_THREADS, 4)
search_count_kernel(const SearchParams P)
{
	extern __shared__ uint32_t sm_planes[];   // [substream][plane][lanes_per_row * 4], EXIT: + [substream] bounds

	const uint32_t q = blockIdx.x / P.n_chunks;
	const uint32_t chunk = blockIdx.x % P.n_chunks;
	const uint32_t n = P.n_kmers[q];
	const uint64_t o0 = P.offsets[q] - P.offsets[0];

	const uint32_t lpr = P.lanes_per_row;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t groups = 32u / lpr;
	const uint32_t nsub = SC_WARPS * groups;
	const uint32_t sub = warp * groups + lane / lpr;
	const uint32_t cl = lane % lpr;
	const uint32_t col4 = chunk * lpr + cl;                 // uint4 index within a row
	const bool active = col4 < P.vec_per_row;
	const uint8_t* col_ptr = P.slab + (uint64_t)col4 * 16;
	const uint32_t words_per_chunk = lpr * 4;               // 32-bit column words handled by this block
	const uint32_t seg_cap = search_seg_cap(nsub);

	// EXIT: count a hit needs (hits_kernel::is_hit); single-segment queries, at most 32 substreams (one bound per lane)
	uint32_t need = 0;
	bool can_exit = false;
	volatile uint32_t* s_ub = sm_planes + (uint64_t)nsub * SC_PLANES * words_per_chunk;
	if (EXIT) {
		need = (P.threshold == 1.0f) ? n : (uint32_t)__fmul_rn(P.threshold, (float)n);
		can_exit = n <= seg_cap && need > 0 && nsub <= 32;
	}

	for (uint32_t seg0 = 0; seg0 == 0 || seg0 < n; seg0 += seg_cap) {
		const uint32_t seg_n = (n > seg0) ? min(seg_cap, n - seg0) : 0u;
		const uint32_t* rows = P.rows + (o0 + seg0) * NH;

		uint4 pl[SC_PLANES];
#pragma unroll
		for (int i = 0; i < SC_PLANES; ++i) pl[i] = make_uint4(0, 0, 0, 0);

		const uint32_t n_blk = (seg_n + 16 * nsub - 1) / (16 * nsub);
		const uint32_t my_total = search_sub_total(seg_n, sub, nsub);                       // k-mers of my substream
		if (EXIT && can_exit) {
			if (cl == 0) s_ub[sub] = my_total;
			__syncthreads();
		}
		// the first block of 16 after which a bound can fall below `need`: fewer than `need` k-mers are left
		const uint32_t blk_first = (EXIT && can_exit) ? search_exit_first_blk(seg_n, need, nsub) : 0xFFFFFFFFu;
		if (active || (EXIT && can_exit)) {
#pragma unroll 1
			for (uint32_t blk = 0; blk < n_blk; ++blk) {
				uint4 fA = make_uint4(0, 0, 0, 0), eA = make_uint4(0, 0, 0, 0);
#pragma unroll
				for (int quad = 0; quad < 4; ++quad) {
					uint4 v[4];
#pragma unroll
					for (int u = 0; u < 4; ++u) {
						const uint32_t i = (blk * 16 + quad * 4 + u) * nsub + sub;
						v[u] = make_uint4(0, 0, 0, 0);
						if (i < seg_n && (!EXIT || active)) {
							const uint32_t* r = rows + (uint64_t)i * NH;
							uint4 acc = ld_nc_v4(col_ptr + (uint64_t)__ldg(r) * P.row_pitch);
#pragma unroll
							for (int h = 1; h < NH; ++h) acc = and4(acc, ld_nc_v4(col_ptr + (uint64_t)__ldg(r + h) * P.row_pitch));
							v[u] = acc;
						}
					}
					uint4 tA, tB, f;
					csa(pl[0], tA, v[0], v[1]);
					csa(pl[0], tB, v[2], v[3]);
					csa(pl[1], f, tA, tB);
					if (quad == 0 || quad == 2) fA = f;
					else {
						uint4 e;
						csa(pl[2], e, fA, f);
						if (quad == 1) eA = e;
						else {
							uint4 c16;
							csa(pl[3], c16, eA, e);
							// ripple the "sixteens" carry into the upper planes
#pragma unroll
							for (int up = SC_LOW; up < SC_PLANES; ++up) {
								const uint4 t = and4(pl[up], c16);
								pl[up].x ^= c16.x; pl[up].y ^= c16.y; pl[up].z ^= c16.z; pl[up].w ^= c16.w;
								c16 = t;
							}
						}
					}
				}
				if (EXIT && search_exit_check_at(blk, blk_first, n_blk)) {
					// (every lane of the warp is here: with can_exit the loop is not predicated on `active`)
					const uint32_t gmask = (lpr == 32) ? 0xFFFFFFFFu : (((1u << lpr) - 1u) << (lane & ~(lpr - 1u)));
					uint4 cand = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
					uint32_t mx = 0;
#pragma unroll
					for (int p = SC_PLANES - 1; p >= 0; --p) {
						const uint4 t = and4(cand, pl[p]);
						const bool some = (__ballot_sync(0xFFFFFFFFu, (t.x | t.y | t.z | t.w) != 0u) & gmask) != 0u;
						cand.x = some ? t.x : cand.x; cand.y = some ? t.y : cand.y; cand.z = some ? t.z : cand.z; cand.w = some ? t.w : cand.w;
						mx |= some ? 1u << p : 0u;
					}
					if (cl == 0) s_ub[sub] = mx + search_sub_left(my_total, blk);
					__syncwarp();
					uint32_t left = (lane < nsub) ? s_ub[lane] : 0u;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) left += __shfl_xor_sync(0xFFFFFFFFu, left, o);
					if (left < need) break;
				}
			}
		}

		// ---- merge the substreams bit-sliced, expand to uint32 counts, write once per segment
		__syncthreads();   // previous segment's readers are done with sm_planes
		if (active) {
#pragma unroll
			for (int i = 0; i < SC_PLANES; ++i) {
				uint32_t* dst = sm_planes + ((uint64_t)(sub * SC_PLANES + i) * words_per_chunk) + cl * 4;
				*reinterpret_cast<uint4*>(dst) = pl[i];
			}
		}
		__syncthreads();

		for (uint32_t w = threadIdx.x; w < words_per_chunk; w += SC_THREADS) {
			const uint32_t vec = chunk * lpr + (w >> 2);
			if (vec >= P.vec_per_row) continue;
			uint32_t tot[SC_TOT];
#pragma unroll
			for (int i = 0; i < SC_TOT; ++i) tot[i] = 0;
			for (uint32_t s = 0; s < nsub; ++s) {
				uint32_t x[SC_PLANES];
#pragma unroll
				for (int i = 0; i < SC_PLANES; ++i) x[i] = sm_planes[(uint64_t)(s * SC_PLANES + i) * words_per_chunk + w];
				bitsliced_add<SC_PLANES>(tot, x);
			}
			const uint64_t col0 = (uint64_t)vec * 128 + (w & 3) * 32;
			uint32_t* out = P.counts + (uint64_t)q * P.count_pitch + col0;
#pragma unroll
			for (int nb = 0; nb < 8; ++nb) {
				if (col0 + 4 * nb + 4 <= P.count_pitch) {
					uint4 c = expand_counts4(tot, nb);
					uint4* o4 = reinterpret_cast<uint4*>(out + 4 * nb);
					if (seg0 != 0) {
						const uint4 prev = *o4;
						c.x += prev.x; c.y += prev.y; c.z += prev.z; c.w += prev.w;
					}
					*o4 = c;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// threshold + compaction
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_hit(uint32_t cnt, uint32_t n, float threshold, bool complete, uint32_t thr)
{
	return complete ? (cnt == n) : (cnt >= thr);
}

// one warp per query; pass 0 counts, pass 1 writes at hit_base[q] in filter order
template <int PASS>
__global__ void __launch_bounds__(256)
hits_kernel(const uint32_t* __restrict__ counts, uint64_t count_pitch, uint32_t n_filters, const uint32_t* __restrict__ n_kmers,
	uint32_t n_queries, float threshold, uint32_t* __restrict__ hit_count, const uint64_t* __restrict__ hit_base,
	kwg_hit_t* __restrict__ hits, uint32_t query0, uint32_t filter0)
{
	const uint32_t q = (blockIdx.x * 256 + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	if (q >= n_queries) return;
	const uint32_t n = n_kmers[q];
	if (n == 0) {                                       // kwage.cpp:370-372: query shorter than k
		if (PASS == 0 && lane == 0) hit_count[q] = 0;
		return;
	}
	const bool complete = (threshold == 1.0f);          // kwage.cpp:349
	const uint32_t thr = (uint32_t)__fmul_rn(threshold, (float)n);   // kwage.cpp:388
	const uint32_t* row = counts + (uint64_t)q * count_pitch;
	uint64_t at = (PASS == 1) ? hit_base[q] : 0;
	uint32_t total = 0;
	for (uint32_t f0 = 0; f0 < n_filters; f0 += 32) {
		const uint32_t f = f0 + lane;
		const uint32_t cnt = (f < n_filters) ? row[f] : 0u;
		const bool hit = (f < n_filters) && is_hit(cnt, n, threshold, complete, thr);
		const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
		if (PASS == 1 && hit) {
			kwg_hit_t h;
			h.query = query0 + q; h.filter = filter0 + f; h.num_match = complete ? n : cnt;   // kwage.cpp:519-520
			hits[at + __popc(m & ((1u << lane) - 1u))] = h;
		}
		at += __popc(m);
		total += __popc(m);
	}
	if (PASS == 0 && lane == 0) hit_count[q] = total;
}

// exclusive prefix of the per-query hit counts (one block): hit_base[q] = *total + sum of the counts before q; *total advances
__global__ void __launch_bounds__(1024)
hit_scan_kernel(const uint32_t* __restrict__ hit_count, uint32_t n_queries, uint64_t* __restrict__ hit_base, unsigned long long* __restrict__ total)
{
	__shared__ unsigned long long s_w[32];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned long long carry = *total;
	for (uint32_t q0 = 0; q0 < n_queries; q0 += 1024) {
		const uint32_t q = q0 + tid;
		const unsigned long long c = (q < n_queries) ? hit_count[q] : 0ull;
		unsigned long long inc = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
			if (lane >= (uint32_t)o) inc += y;
		}
		__syncthreads();
		if (lane == 31) s_w[warp] = inc;
		__syncthreads();
		unsigned long long base = 0, tot = 0;
		for (int w = 0; w < 32; ++w) { if ((uint32_t)w < warp) base += s_w[w]; tot += s_w[w]; }
		if (q < n_queries) hit_base[q] = carry + base + inc - c;
		carry += tot;
	}
	__syncthreads();
	if (tid == 0) *total = carry;
}

} // namespace kwg

using namespace kwg;

struct kwg_db {
	int device = 0;
	cudaStream_t stream = nullptr;
	uint8_t* slab = nullptr;
	bool owns_slab = false;
	uint64_t row_pitch = 0;
	uint32_t k = 0, num_hash = 0, log2_len = 0, n_filters = 0;
	uint32_t n_filters_total = 0, col_begin = 0;   // file geometry (kwg_db_alloc / kwg_db_load)
	bool no_exit = false;             // KWG_SEARCH_NO_EXIT=1 (read when the handle is created): kwg_search* read every row (A/B runs)
	// per-call scratch, grown on demand
	char* d_bases = nullptr;          size_t bases_cap = 0;
	uint64_t* d_offsets = nullptr;    size_t offsets_cap = 0;
	uint64_t* d_table = nullptr;      size_t table_cap = 0;
	uint64_t* d_kmers = nullptr;      size_t kmers_cap = 0;
	uint32_t* d_rows = nullptr;       size_t rows_cap = 0;
	uint32_t* d_nk = nullptr;         size_t nk_cap = 0;
	uint32_t* d_counts = nullptr;     size_t counts_cap = 0;
	uint32_t* d_hit_count = nullptr;  size_t hit_count_cap = 0;
	uint64_t* d_hit_base = nullptr;   size_t hit_base_cap = 0;
	kwg_hit_t* d_hits = nullptr;      size_t hits_cap = 0;
	unsigned long long* d_hit_total = nullptr;          // running length of d_hits
	unsigned long long* h_hit_total = nullptr;          // pinned
	uint64_t count_budget = 1ull << 30;                 // bytes of per-(query, filter) counts held at a time (kwg_search batches its queries)
	uint8_t* d_stage[2] = {nullptr, nullptr};           // kwg_db_upload_columns: pieces of a file on their way into the slab
	size_t stage_cap[2] = {0, 0};
	int stage_next = 0;
	KernelTimers timers;
};

static int grow_db(void** p, size_t* cap, size_t need)
{
	if (need <= *cap) return KWG_OK;
	if (*p) KWG_CUDA(cudaFree(*p));
	*p = nullptr; *cap = 0;
	const size_t want = round_up(need + need / 8, 256);
	KWG_CUDA(cudaMalloc(p, want));
	*cap = want;
	return KWG_OK;
}

static int db_common_init(kwg_db* db, int device, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len, uint32_t n_filters)
{
	if (kmer_len < 1 || kmer_len > KWG_MAX_KMER_LEN) return fail(KWG_ERR_INVALID_ARG, "kmer_len must be in [1,32] (reference word.h:10)");
	if (num_hash < 1 || num_hash > KWG_MAX_NUM_HASH) return fail(KWG_ERR_INVALID_ARG, "num_hash must be in [1,8]");
	if (log2_len > 32) return fail(KWG_ERR_INVALID_ARG, "log2_len must be <= 32 (32-bit hash)");
	if (n_filters == 0) return fail(KWG_ERR_INVALID_ARG, "n_filters must be > 0");
	int rc = select_device(device);
	if (rc) return rc;
	db->device = device; db->k = kmer_len; db->num_hash = num_hash; db->log2_len = log2_len; db->n_filters = n_filters;
	db->no_exit = getenv("KWG_SEARCH_NO_EXIT") != nullptr;
	KWG_CUDA(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
	return KWG_OK;
}

// Device-side pipeline shared by every search entry point.  d_bases/d_offsets on the device.
// exit_threshold > 0: only the filters that reach this threshold matter to the caller (search_count_kernel<NH, true>)
static int search_counts_device(kwg_db* db, const char* d_bases, const uint64_t* d_offsets, uint32_t n_queries,
	uint64_t n_bases, uint32_t max_query_len, uint32_t* d_nk, uint32_t* d_counts, uint64_t count_pitch, float exit_threshold = 0.0f)
{
	int rc;
	if ((rc = grow_db((void**)&db->d_table, &db->table_cap, std::max<uint64_t>(n_bases, 1) * 2 * sizeof(uint64_t)))) return rc;
	if ((rc = grow_db((void**)&db->d_kmers, &db->kmers_cap, std::max<uint64_t>(n_bases, 1) * sizeof(uint64_t)))) return rc;
	if ((rc = grow_db((void**)&db->d_rows, &db->rows_cap, std::max<uint64_t>(n_bases, 1) * db->num_hash * sizeof(uint32_t)))) return rc;
	KWG_CUDA(cudaMemsetAsync(d_nk, 0, (size_t)n_queries * sizeof(uint32_t), db->stream));

	const uint32_t filter_mask = (db->log2_len >= 32) ? 0xFFFFFFFFu : ((1u << db->log2_len) - 1u);
	// blockIdx.y strides over the tiles of a query: a performance knob only (any query length is covered by the
	// kernel's loop).  The device entry point does not know the longest query: assume at most 4x the average.
	const uint64_t len_bound = std::min<uint64_t>(max_query_len, 4 * ceil_div(std::max<uint64_t>(n_bases, 1), std::max<uint32_t>(n_queries, 1)));
	const uint32_t parts = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ceil_div(len_bound, QK_TILE), 1), 256);
	const dim3 qgrid(n_queries, parts);
	db->timers.begin(KWG_T_AUX, db->stream);
	query_table_init_kernel<<<qgrid, 256, 0, db->stream>>>(d_offsets, db->d_table);
	KWG_LAUNCHED();
	switch (db->num_hash) {
#define KWG_CASE(N) case N: query_kmers_kernel<N><<<qgrid, QK_THREADS, 0, db->stream>>>(d_bases, n_bases, d_offsets, db->k, filter_mask, \
		db->d_table, db->d_kmers, db->d_rows, d_nk); break;
		KWG_CASE(1) KWG_CASE(2) KWG_CASE(3) KWG_CASE(4) KWG_CASE(5) KWG_CASE(6) KWG_CASE(7) KWG_CASE(8)
#undef KWG_CASE
	}
	db->timers.end(db->stream);
	KWG_LAUNCHED();

	SearchParams P{};
	P.slab = db->slab;
	P.row_pitch = db->row_pitch;
	P.vec_per_row = (uint32_t)(db->row_pitch / 16);
	uint32_t lpr = 1;
	while (lpr < 32 && lpr < P.vec_per_row) lpr <<= 1;
	P.lanes_per_row = lpr;
	P.n_chunks = (uint32_t)ceil_div(P.vec_per_row, lpr);
	P.offsets = d_offsets;
	P.rows = db->d_rows;
	P.n_kmers = d_nk;
	P.counts = d_counts;
	P.count_pitch = count_pitch;
	const uint64_t blocks = (uint64_t)n_queries * P.n_chunks;
	if (blocks > 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "too many (query, column chunk) pairs for one launch");
	const uint32_t nsub = SC_WARPS * (32 / lpr);
	const bool early = exit_threshold > 0.0f && !db->no_exit;
	P.threshold = exit_threshold;
	const size_t smem = (size_t)nsub * SC_PLANES * lpr * 4 * sizeof(uint32_t) + (early ? nsub * sizeof(uint32_t) : 0);
	// (the limit is raised when a launch needs more than any launch on this device before it, not on every call -- every
	// runtime call counts when eight processes drive eight devices through one host -- and it is never lowered: another
	// handle on the device may still need the larger value)
	static std::mutex attr_mu;
	static size_t attr_set[64][KWG_MAX_NUM_HASH + 1][2];
	bool raise;
	{
		std::lock_guard<std::mutex> lock(attr_mu);
		size_t& cur = attr_set[db->device & 63][db->num_hash][early ? 1 : 0];
		raise = smem > cur;
		if (raise) cur = smem;
	}
	db->timers.begin(KWG_T_SEARCH, db->stream);
	switch (db->num_hash) {
#define KWG_CASE(N) case N: \
		if (early) { \
			if (raise) KWG_CUDA(cudaFuncSetAttribute(search_count_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
			search_count_kernel<N, true><<<(unsigned)blocks, SC_THREADS, smem, db->stream>>>(P); \
		} else { \
			if (raise) KWG_CUDA(cudaFuncSetAttribute(search_count_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
			search_count_kernel<N, false><<<(unsigned)blocks, SC_THREADS, smem, db->stream>>>(P); \
		} break;
		KWG_CASE(1) KWG_CASE(2) KWG_CASE(3) KWG_CASE(4) KWG_CASE(5) KWG_CASE(6) KWG_CASE(7) KWG_CASE(8)
#undef KWG_CASE
	}
	db->timers.end(db->stream);
	KWG_LAUNCHED();
	return KWG_OK;
}

// Host inputs -> device scratch; returns n_bases and the longest query.
static int stage_queries(kwg_db* db, const char* bases, const uint64_t* offsets, uint32_t n_queries, uint64_t* n_bases_out,
	uint32_t* max_len_out)
{
	const uint64_t off0 = offsets[0];
	uint64_t max_len = 0;
	for (uint32_t q = 0; q < n_queries; ++q) {
		if (offsets[q + 1] < offsets[q]) return fail(KWG_ERR_INVALID_ARG, "offsets must be non-decreasing");
		max_len = std::max(max_len, offsets[q + 1] - offsets[q]);
	}
	if (max_len >= 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "a single query of 2^31 bases or more is not supported");
	const uint64_t n_bases = offsets[n_queries] - off0;
	int rc;
	if ((rc = grow_db((void**)&db->d_bases, &db->bases_cap, n_bases + 16))) return rc;
	if ((rc = grow_db((void**)&db->d_offsets, &db->offsets_cap, ((size_t)n_queries + 1) * sizeof(uint64_t)))) return rc;
	if ((rc = grow_db((void**)&db->d_nk, &db->nk_cap, std::max<size_t>(n_queries, 1) * sizeof(uint32_t)))) return rc;
	if (n_bases) KWG_CUDA(cudaMemcpyAsync(db->d_bases, bases + off0, n_bases, cudaMemcpyHostToDevice, db->stream));
	KWG_CUDA(cudaMemcpyAsync(db->d_offsets, offsets, ((size_t)n_queries + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, db->stream));
	*n_bases_out = n_bases;
	*max_len_out = (uint32_t)max_len;
	return KWG_OK;
}

extern "C" {

void kwg_db_unload(kwg_db_t* db)
{
	if (!db) return;
	cudaSetDevice(db->device);
	if (db->stream) cudaStreamSynchronize(db->stream);
	if (db->owns_slab) cudaFree(db->slab);
	cudaFree(db->d_stage[0]); cudaFree(db->d_stage[1]);
	cudaFree(db->d_bases); cudaFree(db->d_offsets); cudaFree(db->d_table); cudaFree(db->d_kmers); cudaFree(db->d_rows);
	cudaFree(db->d_nk); cudaFree(db->d_counts); cudaFree(db->d_hit_count); cudaFree(db->d_hit_base); cudaFree(db->d_hits);
	cudaFree(db->d_hit_total);
	if (db->h_hit_total) cudaFreeHost(db->h_hit_total);
	if (db->stream) cudaStreamDestroy(db->stream);
	delete db;
}

int kwg_db_alloc(kwg_db_t** out, int device, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len,
	uint32_t n_filters_total, uint32_t col_begin, uint32_t col_end)
{
	if (!out) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*out = nullptr;
	if (col_begin >= col_end || col_end > n_filters_total) return fail(KWG_ERR_INVALID_ARG, "bad column range");
	if (col_begin % 8) return fail(KWG_ERR_INVALID_ARG, "col_begin must be a multiple of 8");
	kwg_db* db = new kwg_db();
	int rc = db_common_init(db, device, kmer_len, num_hash, log2_len, col_end - col_begin);
	if (rc) { kwg_db_unload(db); return rc; }
	const uint64_t n_rows = 1ull << log2_len;
	db->n_filters_total = n_filters_total;
	db->col_begin = col_begin;
	db->row_pitch = round_up(ceil_div(col_end - col_begin, 8), 16);
	db->owns_slab = true;
	cudaError_t e = cudaMalloc(&db->slab, (size_t)(n_rows * db->row_pitch));
	if (e != cudaSuccess) {
		rc = fail(KWG_ERR_NO_MEMORY, std::string("slice slab: ") + cudaGetErrorString(e));
		kwg_db_unload(db);
		return rc;
	}
	// padding bytes of every row must read as zero (build_db.cpp:267 zeroes them in the file too)
	e = cudaMemsetAsync(db->slab, 0, (size_t)(n_rows * db->row_pitch), db->stream);
	if (e != cudaSuccess) {
		rc = fail(KWG_ERR_CUDA, std::string("slab clear: ") + cudaGetErrorString(e));
		kwg_db_unload(db);
		return rc;
	}
	*out = db;
	return KWG_OK;
}

int kwg_db_upload_rows_async(kwg_db_t* db, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows)
{
	if (!db || !rows) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (!db->owns_slab || db->n_filters_total == 0) return fail(KWG_ERR_STATE, "handle was not created by kwg_db_alloc/kwg_db_load");
	if (row_begin + n_rows > (1ull << db->log2_len)) return fail(KWG_ERR_INVALID_ARG, "row range outside the filter");
	if (n_rows == 0) return KWG_OK;
	int rc = select_device(db->device);
	if (rc) return rc;
	const uint64_t src_pitch = ceil_div(db->n_filters_total, 8);
	const uint64_t width = ceil_div(db->n_filters, 8);
	if (src_pitch == width && width == db->row_pitch) {
		// the file's rows are the slab's rows: one flat copy at the full PCIe rate
		KWG_CUDA(cudaMemcpyAsync(db->slab + row_begin * db->row_pitch, rows, (size_t)(n_rows * width), cudaMemcpyHostToDevice, db->stream));
	} else {
		KWG_CUDA(cudaMemcpy2DAsync(db->slab + row_begin * db->row_pitch, db->row_pitch, rows + db->col_begin / 8, src_pitch, width, n_rows,
			cudaMemcpyHostToDevice, db->stream));
	}
	return KWG_OK;
}

int kwg_db_upload_rows(kwg_db_t* db, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows)
{
	int rc = kwg_db_upload_rows_async(db, row_begin, n_rows, rows);
	if (rc) return rc;
	if (n_rows) KWG_CUDA(cudaStreamSynchronize(db->stream));   // the caller may reuse `rows`
	return KWG_OK;
}

// ---- several database files as ONE column slab
// The reference keeps at most 2048 filters per .db file (options.h:137-138): rows of 256 bytes.  A gather of 256-byte rows
// reaches a third of the HBM rate a gather of >= 1 KiB rows does (profiles/r1e_sweep.md), so files with the same (k, hashes,
// length) are laid side by side in one slab: file f owns the columns [col_begin_f, col_begin_f + n_f) of every row.
__global__ void __launch_bounds__(256)
place_columns_kernel(const uint8_t* __restrict__ src, uint64_t src_pitch, uint32_t n_cols, uint32_t col_begin, uint64_t n_rows,
	uint32_t* __restrict__ slab, uint64_t slab_pitch_words, uint32_t w0, uint32_t n_words)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rows * n_words) return;
	const uint64_t row = i / n_words;
	const uint32_t w = w0 + (uint32_t)(i % n_words);
	const uint32_t lo = max(32u * w, col_begin), hi = min(32u * w + 32u, col_begin + n_cols);      // destination bits of this word
	const uint32_t s = lo - col_begin, n = hi - lo;                                                // source bits [s, s + n)
	const uint8_t* p = src + row * src_pitch;
	const uint32_t b0 = s >> 3, b1 = (s + n - 1) >> 3;
	uint64_t v = 0;
	for (uint32_t b = b0; b <= b1; ++b) v |= (uint64_t)p[b] << (8 * (b - b0));                    // at most 5 bytes
	uint32_t bits = (uint32_t)(v >> (s & 7u));
	if (n < 32) bits &= (1u << n) - 1u;
	bits <<= (lo & 31u);
	if (bits) atomicOr(slab + row * slab_pitch_words + w, bits);
}

int kwg_db_upload_columns_async(kwg_db_t* db, uint32_t col_begin, uint32_t n_cols, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows)
{
	if (!db || !rows) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (!db->owns_slab || db->n_filters_total == 0) return fail(KWG_ERR_STATE, "handle was not created by kwg_db_alloc");
	if (n_cols == 0 || (uint64_t)col_begin + n_cols > db->n_filters) return fail(KWG_ERR_INVALID_ARG, "column range outside the slab");
	if (row_begin + n_rows > (1ull << db->log2_len)) return fail(KWG_ERR_INVALID_ARG, "row range outside the filter");
	if (n_rows == 0) return KWG_OK;
	int rc = select_device(db->device);
	if (rc) return rc;
	const uint64_t src_pitch = ceil_div(n_cols, 8);
	// the piece crosses PCIe as ONE flat copy (a pitched host-to-device copy of 256-byte rows runs far below the link
	// rate), then it is laid into its columns at HBM speed: a pitched device copy for whole bytes, else bits OR-ed into
	// place (the slab starts out all zero).  Two staging buffers alternate so that a piece can travel while the one
	// before it is still being placed.
	const size_t bytes = (size_t)(n_rows * src_pitch);
	const int sb = db->stage_next;
	db->stage_next ^= 1;
	rc = grow_db((void**)&db->d_stage[sb], &db->stage_cap[sb], bytes);
	if (rc) return rc;
	KWG_CUDA(cudaMemcpyAsync(db->d_stage[sb], rows, bytes, cudaMemcpyHostToDevice, db->stream));
	if (col_begin % 8 == 0 && n_cols % 8 == 0) {
		KWG_CUDA(cudaMemcpy2DAsync(db->slab + row_begin * db->row_pitch + col_begin / 8, db->row_pitch, db->d_stage[sb], src_pitch, src_pitch, n_rows,
			cudaMemcpyDeviceToDevice, db->stream));
	} else {
		const uint32_t w0 = col_begin / 32, n_words = (col_begin + n_cols - 1) / 32 - w0 + 1;
		const uint64_t total = n_rows * n_words;
		if (ceil_div(total, 256) > 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "piece too large: upload fewer rows per call");
		place_columns_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, db->stream>>>(db->d_stage[sb], src_pitch, n_cols,
			col_begin, n_rows, reinterpret_cast<uint32_t*>(db->slab + row_begin * db->row_pitch), db->row_pitch / 4, w0, n_words);
		KWG_LAUNCHED();
	}
	return KWG_OK;
}

int kwg_db_upload_columns(kwg_db_t* db, uint32_t col_begin, uint32_t n_cols, uint64_t row_begin, uint64_t n_rows, const uint8_t* rows)
{
	int rc = kwg_db_upload_columns_async(db, col_begin, n_cols, row_begin, n_rows, rows);
	if (rc) return rc;
	if (n_rows) KWG_CUDA(cudaStreamSynchronize(db->stream));   // the caller may reuse `rows`
	return KWG_OK;
}

int kwg_db_load(kwg_db_t** out, int device, const uint8_t* slices, uint32_t kmer_len, uint32_t num_hash,
	uint32_t log2_len, uint32_t n_filters_total, uint32_t col_begin, uint32_t col_end)
{
	if (!out || !slices) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = kwg_db_alloc(out, device, kmer_len, num_hash, log2_len, n_filters_total, col_begin, col_end);
	if (rc) return rc;
	rc = kwg_db_upload_rows(*out, 0, 1ull << log2_len, slices);
	if (rc) { kwg_db_unload(*out); *out = nullptr; }
	return rc;
}

int kwg_db_attach_dev(kwg_db_t** out, int device, const uint8_t* d_slices, uint64_t row_pitch,
	uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len, uint32_t n_filters)
{
	if (!out || !d_slices) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*out = nullptr;
	if (row_pitch % 16 || row_pitch < ceil_div(n_filters, 8)) return fail(KWG_ERR_INVALID_ARG, "row_pitch must be a multiple of 16 and cover a slice");
	if (reinterpret_cast<uintptr_t>(d_slices) & 15u) return fail(KWG_ERR_INVALID_ARG, "d_slices must be 16-byte aligned");
	kwg_db* db = new kwg_db();
	int rc = db_common_init(db, device, kmer_len, num_hash, log2_len, n_filters);
	if (rc) { kwg_db_unload(db); return rc; }
	db->slab = const_cast<uint8_t*>(d_slices);
	db->owns_slab = false;
	db->row_pitch = row_pitch;
	*out = db;
	return KWG_OK;
}

int kwg_search_counts_dev(kwg_db_t* db, const char* d_bases, const uint64_t* d_offsets, uint32_t n_queries,
	uint64_t n_bases, uint32_t* d_n_query_kmers, uint32_t* d_counts, uint64_t count_pitch)
{
	if (!db || !d_bases || !d_offsets || !d_n_query_kmers || !d_counts) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (count_pitch % 4 || count_pitch < db->n_filters) return fail(KWG_ERR_INVALID_ARG, "count_pitch must be a multiple of 4 and >= n_filters");
	if (reinterpret_cast<uintptr_t>(d_counts) & 15u) return fail(KWG_ERR_INVALID_ARG, "d_counts must be 16-byte aligned");
	if (reinterpret_cast<uintptr_t>(d_bases) & 3u) return fail(KWG_ERR_INVALID_ARG, "d_bases must be 4-byte aligned");
	if (n_queries == 0) return KWG_OK;
	int rc = select_device(db->device);
	if (rc) return rc;
	// without the host copy of the offsets the longest query is unknown: assume the worst case
	return search_counts_device(db, d_bases, d_offsets, n_queries, n_bases, (uint32_t)std::min<uint64_t>(n_bases, 0x7FFFFFFFull),
		d_n_query_kmers, d_counts, count_pitch);
}

int kwg_search_counts(kwg_db_t* db, const char* bases, const uint64_t* offsets, uint32_t n_queries,
	uint32_t* n_query_kmers, uint32_t* counts)
{
	if (!db || !bases || !offsets || !counts) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (n_queries == 0) return KWG_OK;
	int rc = select_device(db->device);
	if (rc) return rc;
	uint64_t n_bases = 0;
	uint32_t max_len = 0;
	if ((rc = stage_queries(db, bases, offsets, n_queries, &n_bases, &max_len))) return rc;
	const uint64_t pitch = round_up(db->n_filters, 4);
	if ((rc = grow_db((void**)&db->d_counts, &db->counts_cap, (size_t)n_queries * pitch * sizeof(uint32_t)))) return rc;
	if ((rc = search_counts_device(db, db->d_bases, db->d_offsets, n_queries, n_bases, max_len, db->d_nk, db->d_counts, pitch))) return rc;
	KWG_CUDA(cudaMemcpy2DAsync(counts, (size_t)db->n_filters * sizeof(uint32_t), db->d_counts, pitch * sizeof(uint32_t),
		(size_t)db->n_filters * sizeof(uint32_t), n_queries, cudaMemcpyDeviceToHost, db->stream));
	if (n_query_kmers)
		KWG_CUDA(cudaMemcpyAsync(n_query_kmers, db->d_nk, (size_t)n_queries * sizeof(uint32_t), cudaMemcpyDeviceToHost, db->stream));
	KWG_CUDA(cudaStreamSynchronize(db->stream));
	return KWG_OK;
}

// Queries [q0, q1) that fit the handle's budget for the per-(query, filter) counts and the 2^31-base staging limit.
static uint32_t next_query_batch(const kwg_db* db, const uint64_t* offsets, uint32_t q0, uint32_t n_queries, uint64_t pitch)
{
	const uint64_t max_q = std::max<uint64_t>(1, db->count_budget / (pitch * sizeof(uint32_t)));
	uint32_t q1 = (uint32_t)std::min<uint64_t>(n_queries, (uint64_t)q0 + max_q);
	while (q1 > q0 + 1 && offsets[q1] - offsets[q0] > (1ull << 30)) q1 = q0 + (q1 - q0) / 2;
	return q1;
}

// Shared by kwg_search and kwg_search_hits_dev: the hits of all the queries, compacted on the device in (query, filter)
// order into db->d_hits; queries go through in batches so that counts and scratch stay within the handle's budget
// (the reference streams one query at a time, kwage.cpp:116-148); filter indices are offset by filter0.
static int search_hits_device(kwg_db* db, const char* bases, const uint64_t* offsets, uint32_t n_queries, float threshold,
	uint32_t* n_query_kmers, uint32_t filter0, uint64_t* total_out)
{
	int rc;
	const uint64_t pitch = round_up(db->n_filters, 4);
	if (!db->d_hit_total) {
		KWG_CUDA(cudaMalloc(&db->d_hit_total, sizeof(unsigned long long)));
		KWG_CUDA(cudaMallocHost(&db->h_hit_total, sizeof(unsigned long long)));
	}
	KWG_CUDA(cudaMemsetAsync(db->d_hit_total, 0, sizeof(unsigned long long), db->stream));
	uint64_t total = 0;
	for (uint32_t q0 = 0; q0 < n_queries;) {
		const uint32_t q1 = next_query_batch(db, offsets, q0, n_queries, pitch);
		const uint32_t nq = q1 - q0;
		uint64_t n_bases = 0;
		uint32_t max_len = 0;
		if ((rc = stage_queries(db, bases, offsets + q0, nq, &n_bases, &max_len))) return rc;
		if ((rc = grow_db((void**)&db->d_counts, &db->counts_cap, (size_t)nq * pitch * sizeof(uint32_t)))) return rc;
		if ((rc = grow_db((void**)&db->d_hit_count, &db->hit_count_cap, (size_t)nq * sizeof(uint32_t)))) return rc;
		if ((rc = grow_db((void**)&db->d_hit_base, &db->hit_base_cap, (size_t)nq * sizeof(uint64_t)))) return rc;
		if ((rc = search_counts_device(db, db->d_bases, db->d_offsets, nq, n_bases, max_len, db->d_nk, db->d_counts, pitch, threshold))) return rc;

		const unsigned hgrid = (unsigned)ceil_div((uint64_t)nq * 32, 256);
		db->timers.begin(KWG_T_HITS, db->stream);
		hits_kernel<0><<<hgrid, 256, 0, db->stream>>>(db->d_counts, pitch, db->n_filters, db->d_nk, nq, threshold,
			db->d_hit_count, nullptr, nullptr, 0, 0);
		KWG_LAUNCHED();
		hit_scan_kernel<<<1, 1024, 0, db->stream>>>(db->d_hit_count, nq, db->d_hit_base, db->d_hit_total);
		db->timers.end(db->stream);
		KWG_LAUNCHED();
		KWG_CUDA(cudaMemcpyAsync(db->h_hit_total, db->d_hit_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, db->stream));
		if (n_query_kmers)
			KWG_CUDA(cudaMemcpyAsync(n_query_kmers + q0, db->d_nk, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, db->stream));
		KWG_CUDA(cudaStreamSynchronize(db->stream));
		const uint64_t new_total = *db->h_hit_total;
		if (new_total > total) {
			if (new_total * sizeof(kwg_hit_t) > db->hits_cap) {
				// grow, keeping the hits of the batches before
				kwg_hit_t* n = nullptr;
				const size_t want = round_up(new_total * sizeof(kwg_hit_t) * 3 / 2, 256);
				KWG_CUDA(cudaMalloc(&n, want));
				if (total) KWG_CUDA(cudaMemcpyAsync(n, db->d_hits, total * sizeof(kwg_hit_t), cudaMemcpyDeviceToDevice, db->stream));
				KWG_CUDA(cudaStreamSynchronize(db->stream));
				if (db->d_hits) KWG_CUDA(cudaFree(db->d_hits));
				db->d_hits = n; db->hits_cap = want;
			}
			db->timers.begin(KWG_T_HITS, db->stream);
			hits_kernel<1><<<hgrid, 256, 0, db->stream>>>(db->d_counts, pitch, db->n_filters, db->d_nk, nq, threshold,
				db->d_hit_count, db->d_hit_base, db->d_hits, q0, filter0);
			db->timers.end(db->stream);
			KWG_LAUNCHED();
		}
		total = new_total;
		q0 = q1;
	}
	*total_out = total;
	return KWG_OK;
}

int kwg_search(kwg_db_t* db, const char* bases, const uint64_t* offsets, uint32_t n_queries, float threshold,
	uint32_t* n_query_kmers, kwg_hit_t** hits, uint64_t* n_hits)
{
	if (!db || !bases || !offsets || !hits || !n_hits) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*hits = nullptr; *n_hits = 0;
	if (!(threshold > 0.0f && threshold <= 1.0f)) return fail(KWG_ERR_INVALID_ARG, "threshold must be in (0,1] (reference options.cpp:186-191)");
	if (n_queries == 0) return KWG_OK;
	int rc = select_device(db->device);
	if (rc) return rc;
	uint64_t total = 0;
	if ((rc = search_hits_device(db, bases, offsets, n_queries, threshold, n_query_kmers, 0, &total))) return rc;
	if (total == 0) return KWG_OK;
	kwg_hit_t* h = (kwg_hit_t*)std::malloc((size_t)total * sizeof(kwg_hit_t));
	if (!h) return fail(KWG_ERR_NO_MEMORY, "host allocation of the hit list failed");
	cudaError_t e = cudaMemcpyAsync(h, db->d_hits, (size_t)total * sizeof(kwg_hit_t), cudaMemcpyDeviceToHost, db->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(db->stream);
	if (e != cudaSuccess) {
		std::free(h);
		return fail(KWG_ERR_CUDA, std::string("hit list copy: ") + cudaGetErrorString(e));
	}
	*hits = h;
	*n_hits = total;
	return KWG_OK;
}

int kwg_search_hits_dev(kwg_db_t* db, const char* bases, const uint64_t* offsets, uint32_t n_queries, float threshold,
	uint32_t* n_query_kmers, uint32_t filter0, const kwg_hit_t** d_hits, uint64_t* n_hits)
{
	if (!db || !bases || !offsets || !d_hits || !n_hits) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*d_hits = nullptr; *n_hits = 0;
	if (!(threshold > 0.0f && threshold <= 1.0f)) return fail(KWG_ERR_INVALID_ARG, "threshold must be in (0,1] (reference options.cpp:186-191)");
	if (n_queries == 0) return KWG_OK;
	int rc = select_device(db->device);
	if (rc) return rc;
	uint64_t total = 0;
	if ((rc = search_hits_device(db, bases, offsets, n_queries, threshold, n_query_kmers, filter0, &total))) return rc;
	*d_hits = db->d_hits;
	*n_hits = total;
	return KWG_OK;
}

int kwg_db_set_count_budget(kwg_db_t* db, uint64_t bytes)
{
	if (!db) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	if (bytes < 4096) return fail(KWG_ERR_INVALID_ARG, "budget too small");
	db->count_budget = bytes;
	return KWG_OK;
}

int kwg_search_ptrs(kwg_db_t* db, const char* const* queries, const uint64_t* query_len, uint32_t n_queries,
	float threshold, uint32_t* n_query_kmers, kwg_hit_t** hits, uint64_t* n_hits)
{
	if (!queries || !query_len) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	std::vector<uint64_t> offsets((size_t)n_queries + 1, 0);
	for (uint32_t q = 0; q < n_queries; ++q) offsets[q + 1] = offsets[q] + query_len[q];
	std::vector<char> flat((size_t)offsets[n_queries] + 1);
	for (uint32_t q = 0; q < n_queries; ++q)
		if (query_len[q]) std::memcpy(flat.data() + offsets[q], queries[q], (size_t)query_len[q]);
	return kwg_search(db, flat.data(), offsets.data(), n_queries, threshold, n_query_kmers, hits, n_hits);
}

// ------------------------------------------------------------------------------------------ multi-GPU gather
// The one exchange of the path: every device holds a column slab, searches all the queries, and the
// compacted hit lists -- a few KB -- travel to the root over NCCL (NVLink): one all-gather of the list lengths, one
// grouped send/recv of the lists themselves straight from HBM, one counting sort by query on the root's host.
// NCCL is loaded at run time (dlopen of libnccl.so.2: the one the process already has, e.g. PyTorch's, else the
// system's), so the library has no link-time dependency on it and single-GPU users never touch it.
namespace {
struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	std::string error;
};

NcclApi& nccl_api()
{
	static NcclApi api = [] {
		NcclApi a;
		a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!a.lib) { a.error = std::string("libnccl.so.2 could not be loaded: ") + dlerror(); return a; }
#define KWG_SYM(field, name) \
		*reinterpret_cast<void**>(&a.field) = dlsym(a.lib, name); \
		if (!a.field && a.error.empty()) a.error = std::string("libnccl.so.2 lacks ") + name;
		KWG_SYM(GetUniqueId, "ncclGetUniqueId") KWG_SYM(CommInitRank, "ncclCommInitRank") KWG_SYM(CommInitAll, "ncclCommInitAll")
		KWG_SYM(CommDestroy, "ncclCommDestroy") KWG_SYM(AllGather, "ncclAllGather") KWG_SYM(Send, "ncclSend") KWG_SYM(Recv, "ncclRecv")
		KWG_SYM(GroupStart, "ncclGroupStart") KWG_SYM(GroupEnd, "ncclGroupEnd") KWG_SYM(GetErrorString, "ncclGetErrorString")
#undef KWG_SYM
		return a;
	}();
	return api;
}
} // namespace

#define KWG_NCCL(expr)                                                                                   \
	do {                                                                                                 \
		ncclResult_t _r = (expr);                                                                        \
		if (_r != ncclSuccess) return fail(KWG_ERR_CUDA, std::string(#expr) + ": " + nccl_api().GetErrorString(_r)); \
	} while (0)

struct kwg_comm {
	ncclComm_t comm = nullptr;
	int device = 0, rank = 0, n_ranks = 1;
	unsigned long long* d_totals = nullptr;       // [n_ranks + 1]: every rank's list length; [n_ranks] = this rank's
	unsigned long long* h_totals = nullptr;       // pinned
	kwg_hit_t* d_gathered = nullptr; size_t gathered_cap = 0;
};

static int comm_finish(kwg_comm* c)
{
	int rc = select_device(c->device);
	if (rc) return rc;
	KWG_CUDA(cudaMalloc(&c->d_totals, (size_t)(c->n_ranks + 1) * sizeof(unsigned long long)));
	KWG_CUDA(cudaMallocHost(&c->h_totals, (size_t)(c->n_ranks + 1) * sizeof(unsigned long long)));
	return KWG_OK;
}

void kwg_merge_hits(const kwg_hit_t* lists, const uint64_t* list_len, uint32_t n_lists, uint32_t n_queries, kwg_hit_t* out)
{
	// counting sort by query, stable over (list, position): lists are ordered by (query, filter) and list r holds the
	// columns before those of list r + 1, so the result is ordered by (query, filter) -- kwg_search's order
	std::vector<uint64_t> at((size_t)n_queries + 1, 0);
	uint64_t total = 0;
	for (uint32_t r = 0; r < n_lists; ++r) total += list_len[r];
	for (uint64_t i = 0; i < total; ++i) ++at[(size_t)lists[i].query + 1];
	for (uint32_t q = 0; q < n_queries; ++q) at[q + 1] += at[q];
	for (uint64_t i = 0; i < total; ++i) out[at[lists[i].query]++] = lists[i];
}

int kwg_comm_get_unique_id(uint8_t* id)
{
	if (!id) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	NcclApi& N = nccl_api();
	if (!N.error.empty()) return fail(KWG_ERR_CUDA, N.error);
	ncclUniqueId u;
	KWG_NCCL(N.GetUniqueId(&u));
	static_assert(sizeof(u) == KWG_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
	std::memcpy(id, &u, sizeof(u));
	return KWG_OK;
}

int kwg_comm_create(kwg_comm_t** out, int device, int n_ranks, int rank, const uint8_t* id)
{
	if (!out || !id) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*out = nullptr;
	if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(KWG_ERR_INVALID_ARG, "bad rank");
	NcclApi& N = nccl_api();
	if (!N.error.empty()) return fail(KWG_ERR_CUDA, N.error);
	int rc = select_device(device);
	if (rc) return rc;
	kwg_comm* c = new kwg_comm();
	c->device = device; c->rank = rank; c->n_ranks = n_ranks;
	ncclUniqueId u;
	std::memcpy(&u, id, sizeof(u));
	ncclResult_t r = N.CommInitRank(&c->comm, n_ranks, u, rank);
	if (r != ncclSuccess) { delete c; return fail(KWG_ERR_CUDA, std::string("ncclCommInitRank: ") + N.GetErrorString(r)); }
	rc = comm_finish(c);
	if (rc) { kwg_comm_destroy(c); return rc; }
	*out = c;
	return KWG_OK;
}

int kwg_comm_create_all(kwg_comm_t** out, int n, const int* devices)
{
	if (!out || !devices || n < 1) return fail(KWG_ERR_INVALID_ARG, "bad argument");
	NcclApi& N = nccl_api();
	if (!N.error.empty()) return fail(KWG_ERR_CUDA, N.error);
	std::vector<ncclComm_t> comms((size_t)n);
	KWG_NCCL(N.CommInitAll(comms.data(), n, devices));
	for (int i = 0; i < n; ++i) {
		kwg_comm* c = new kwg_comm();
		c->comm = comms[(size_t)i]; c->device = devices[i]; c->rank = i; c->n_ranks = n;
		out[i] = c;
		int rc = comm_finish(c);
		if (rc) return rc;
	}
	return KWG_OK;
}

void kwg_comm_destroy(kwg_comm_t* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->comm) nccl_api().CommDestroy(c->comm);
	cudaFree(c->d_totals);
	if (c->h_totals) cudaFreeHost(c->h_totals);
	cudaFree(c->d_gathered);
	delete c;
}

int kwg_search_gather(kwg_db_t* db, kwg_comm_t* c, int root, const char* bases, const uint64_t* offsets, uint32_t n_queries,
	float threshold, uint32_t filter0, uint32_t* n_query_kmers, kwg_hit_t** hits, uint64_t* n_hits)
{
	if (!db || !c || !bases || !offsets || !hits || !n_hits) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*hits = nullptr; *n_hits = 0;
	if (root < 0 || root >= c->n_ranks) return fail(KWG_ERR_INVALID_ARG, "bad root");
	if (c->device != db->device) return fail(KWG_ERR_INVALID_ARG, "communicator and database are on different devices");
	if (!(threshold > 0.0f && threshold <= 1.0f)) return fail(KWG_ERR_INVALID_ARG, "threshold must be in (0,1] (reference options.cpp:186-191)");
	NcclApi& N = nccl_api();
	int rc = select_device(db->device);
	if (rc) return rc;
	// KWG_GATHER_TRACE=1: host wall clock of the phases of this call on stderr (scaling studies; read once per process)
	static const bool trace = getenv("KWG_GATHER_TRACE") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
	uint64_t mine = 0;
	if (n_queries && (rc = search_hits_device(db, bases, offsets, n_queries, threshold, n_query_kmers, filter0, &mine))) return rc;
	const double ms_search = since(t0);
	const auto t1 = std::chrono::steady_clock::now();
	const int W = c->n_ranks;
	// list lengths of all ranks
	c->h_totals[W] = mine;
	KWG_CUDA(cudaMemcpyAsync(c->d_totals + W, c->h_totals + W, sizeof(unsigned long long), cudaMemcpyHostToDevice, db->stream));
	KWG_NCCL(N.AllGather(c->d_totals + W, c->d_totals, 1, ncclUint64, c->comm, db->stream));
	KWG_CUDA(cudaMemcpyAsync(c->h_totals, c->d_totals, (size_t)W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, db->stream));
	KWG_CUDA(cudaStreamSynchronize(db->stream));
	uint64_t total = 0;
	std::vector<uint64_t> len((size_t)W), off((size_t)W);
	for (int r = 0; r < W; ++r) { len[(size_t)r] = c->h_totals[r]; off[(size_t)r] = total; total += len[(size_t)r]; }
	if (c->rank == root && total * sizeof(kwg_hit_t) > c->gathered_cap) {
		if (c->d_gathered) KWG_CUDA(cudaFree(c->d_gathered));
		c->d_gathered = nullptr; c->gathered_cap = 0;
		const size_t want = round_up(total * sizeof(kwg_hit_t) * 3 / 2, 256);
		KWG_CUDA(cudaMalloc(&c->d_gathered, want));
		c->gathered_cap = want;
	}
	const double ms_sizes = since(t1);
	const auto t2 = std::chrono::steady_clock::now();
	// the lists themselves, HBM to HBM (3 words per hit)
	KWG_NCCL(N.GroupStart());
	if (c->rank == root) {
		for (int r = 0; r < W; ++r)
			if (r != root && len[(size_t)r]) KWG_NCCL(N.Recv(c->d_gathered + off[(size_t)r], len[(size_t)r] * 3, ncclUint32, r, c->comm, db->stream));
	} else if (mine) {
		KWG_NCCL(N.Send(db->d_hits, mine * 3, ncclUint32, root, c->comm, db->stream));
	}
	KWG_NCCL(N.GroupEnd());
	if (c->rank != root || total == 0) {
		KWG_CUDA(cudaStreamSynchronize(db->stream));      // the list may be overwritten by the next call
		if (trace) fprintf(stderr, "[kwg gather] rank %d: search %.3f ms, sizes %.3f ms, lists %.3f ms\n", c->rank, ms_search, ms_sizes, since(t2));
		return KWG_OK;
	}
	if (mine) KWG_CUDA(cudaMemcpyAsync(c->d_gathered + off[(size_t)root], db->d_hits, mine * sizeof(kwg_hit_t), cudaMemcpyDeviceToDevice, db->stream));
	std::vector<kwg_hit_t> flat((size_t)total);
	KWG_CUDA(cudaMemcpyAsync(flat.data(), c->d_gathered, total * sizeof(kwg_hit_t), cudaMemcpyDeviceToHost, db->stream));
	KWG_CUDA(cudaStreamSynchronize(db->stream));
	kwg_hit_t* h = (kwg_hit_t*)std::malloc((size_t)total * sizeof(kwg_hit_t));
	if (!h) return fail(KWG_ERR_NO_MEMORY, "host allocation of the hit list failed");
	const double ms_lists = since(t2);
	const auto t3 = std::chrono::steady_clock::now();
	kwg_merge_hits(flat.data(), len.data(), (uint32_t)W, n_queries, h);
	*hits = h;
	*n_hits = total;
	if (trace) fprintf(stderr, "[kwg gather] root %d: search %.3f ms, sizes %.3f ms, lists %.3f ms, merge %.3f ms (%llu hits)\n", c->rank, ms_search, ms_sizes,
		ms_lists, since(t3), (unsigned long long)total);
	return KWG_OK;
}

void kwg_free_hits(kwg_hit_t* hits) { std::free(hits); }

int kwg_db_sync(kwg_db_t* db)
{
	if (!db) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	int rc = select_device(db->device);
	if (rc) return rc;
	KWG_CUDA(cudaStreamSynchronize(db->stream));
	return KWG_OK;
}

int kwg_db_set_timing(kwg_db_t* db, int enable)
{
	if (!db) return fail(KWG_ERR_INVALID_ARG, "handle is NULL");
	db->timers.enabled = enable != 0;
	return KWG_OK;
}

int kwg_db_get_timing(kwg_db_t* db, double* ms, uint64_t* launches)
{
	if (!db || !ms || !launches) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(db->device);
	if (rc) return rc;
	KWG_CUDA(cudaStreamSynchronize(db->stream));
	for (int i = 0; i < KWG_T_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
	db->timers.collect(ms, launches, KWG_T_COUNT);
	return KWG_OK;
}

int kwg_db_stream(kwg_db_t* db, void** stream)
{
	if (!db || !stream) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	*stream = (void*)db->stream;
	return KWG_OK;
}

} // extern "C"
