// Counting-mode construction for min_kmer_count == 1 (reference make_bloom.cpp:506-621), second design:
// ONE partition level and a 1-bit-per-slot resolution in stream order.
//
// With min_kmer_count == 1 an occurrence is valid iff it is the FIRST toucher (earliest in the stream) of at
// least one of its four counting-filter slots (bloom_build.cu has the derivation).  The first design
// (bloom_count.cuh) kept "smallest position seen" per slot -- 32 bits per slot, hence 2^15-slot buckets, 65536
// of them for a 1e6-read accession, a two-level partition and 8-byte records moved four times (128 B/k-mer).
// Here the records of a bucket reach the resolver IN STREAM ORDER, so "was this slot touched before me" is
// one bit: a bucket is 2^20 slots (a 128 KiB shared-memory bitmap), 2048 of them at lc = 30, one partition
// level, and a record is 6 bytes (20-bit slot, 28-bit ordinal) written once and read once.
//
// Occurrences are numbered densely: the ORDINAL of a k-mer occurrence is its rank among the valid windows of
// the sub-batch (stream order).  Ordinal o of a batch lands at entry list_base + o of the accession's word
// list, its losses in nibble o of the loss array: no compaction pass afterwards (pass B of the first design).
//
//   ft_count_kernel    valid windows per tile of 2048 positions (encode + window test only)
//   ft_scan_kernel     exclusive prefix over the tiles: first ordinal of every tile
//   ft_hash_kernel     per tile: canonical k-mer and the 4 counting hashes of every valid window, written
//                      densely by ordinal (16 B) + the canonical word into the accession's list (8 B)
//   ft_append_kernel   one persistent block per SM walks a contiguous range of ordinals, 1024 per round, one per
//                      thread: each touch is appended to the 16-record ring of its bucket in shared memory
//                      (software write combining); the append that completes a unit of 8 records puts the
//                      bucket on the round's flush list, and the list is flushed one bucket per thread
//                      (32 B of low words + 16 B of high halves) into the block's own page chains: the chain
//                      (block c, bucket b) holds the bucket's records of that stretch of the stream, ordered by
//                      round (ordinal >> 10).  Pages come from the block's own pool (no global atomics, exact
//                      worst-case size); an epilogue sorts the page log into per-bucket page lists.
//   ft_resolve_kernel  one bucket at a time per persistent block: bitmap of the bucket from the persistent
//                      touched-bitmap, then the chains in stream order, a window of 8192 new records at a time,
//                      one unit per thread, kept in registers.  A window ends where the last round of its last
//                      chain begins (records of one round are in no particular order; the rest of that round is
//                      carried into the next window), so two records of one slot either meet in one window -- the
//                      smaller ordinal wins, found through short lists -- or the earlier window holds the
//                      earlier one.  Losers add 1 to the 4-bit loss counter of their occurrence.
//   ft_finish_kernel   occurrences with 4 losses (every counter they read was non-zero) are marked in the accession's
//                      invalid bitmap (what finalize skips) and left out of the valid count; the list length advances.
//
// Exact for any input: rings that fill up make the round repeat, a bucket may receive every record of a round
// (the carry holds 4096), a window in which thousands of records meet on a few slots
// (poly-A reads) is settled by min-reduction rounds instead of the lists.
#pragma once
#include "bloom_count.cuh"

namespace kwg {

constexpr int FT_THREADS = 1024;
constexpr int FT_SUB_LOG2 = 10;
constexpr int FT_SUB = 1 << FT_SUB_LOG2;          // ordinals per append round (one per thread): the resolver's ordering unit
constexpr int FT_PER = FT_SUB / FT_THREADS;       // ordinals per thread and round
constexpr int FT_REC = 4 * FT_PER;                // touches per thread and round
constexpr int FT_BUCKET_LOG2 = 20;                // slots per bucket
constexpr int FT_MAX_BUCKETS = 2048;
constexpr int FT_RING = 16;                       // records staged per bucket
constexpr int FT_UNIT = 8;                        // records per flush unit
constexpr int FT_UNIT_BYTES = 48;                 // 8 x u32 (slot | ordinal low 12 << 20) + 8 x u16 (ordinal >> 12)
constexpr int FT_RING_BYTES = 96;                 // 16 x u32 + 16 x u16
constexpr uint32_t FT_NULL_LO = 0xFFFFFFFFu;      // padding record: ordinal 2^28 - 1 (never a real one)
constexpr uint32_t FT_NULL_HI = 0xFFFFu;
constexpr uint32_t FT_NULL_ORD = 0xFFFFFFFu;
constexpr uint32_t FT_SEQ_BITS = 20;              // page log entry: bucket << 20 | page number within the chain
constexpr uint64_t FT_MAX_POS = (1ull << 28) - 8192;        // positions per sub-batch
constexpr int FT_MAX_CHAINS = 1024;               // chains (append blocks over all launches) per sub-batch
constexpr uint32_t FT_TWIN = 0xFFFFFFFFu;         // hash slot of a touch that fell on its table-mate's slot
constexpr uint32_t FT_LIST_LOG2 = 22;             // entries per chunk of the accession's word list (LIST_CHUNK_LOG2)

// tiles of the count / hash kernels
constexpr int HT_THREADS = 256;
constexpr int HT_POS = 2048;
constexpr int HT_LOAD = HT_POS + 32;
constexpr int HT_VEC = HT_LOAD / 16;
constexpr int HT_IT = HT_POS / HT_THREADS;

struct FtTileParams {
	BaseSource src;              // device, 16-byte aligned (whole batch)
	const uint32_t* start_mask;
	uint32_t k;
	uint64_t pos0;               // absolute base index of the sub-batch's first start position (multiple of 16)
	uint64_t n_pos;              // start positions in the sub-batch
	uint32_t tile0;              // first tile of this launch
	uint32_t* tile_cnt;          // [tiles of the sub-batch]: valid windows of the tile, then (scan) ordinal of its first one
	// hash kernel only
	uint32_t count_mask;
	uint4* hm;                   // [ordinal] the four counting hashes (masked), FT_TWIN where a touch doubles its table-mate
	uint64_t* const* list_chunks;
	uint64_t list_base;          // list entry of ordinal 0 of this sub-batch
	uint32_t* loss;
};

// stage the tile: 2-bit codes (hash kernel), bad-base and read-start bitmaps
template <bool CODES>
__device__ __forceinline__ void ft_stage_tile(const FtTileParams& P, uint64_t t0, uint32_t* s_codes, uint32_t* s_bad, uint32_t* s_start)
{
	const uint32_t tid = threadIdx.x;
	for (uint32_t v = tid; v < (uint32_t)HT_VEC; v += HT_THREADS) {
		const uint64_t g = t0 + (uint64_t)v * 16;
		uint32_t codes, bad16;
		load_group16(P.src, g, codes, bad16);
		if (CODES) s_codes[v] = codes;
		reinterpret_cast<uint16_t*>(s_bad)[v] = (uint16_t)bad16;
	}
	for (uint32_t v = tid; v < (uint32_t)(HT_LOAD / 32 + 1); v += HT_THREADS) {
		const uint64_t w = (t0 >> 5) + v;
		s_start[v] = (w * 32 < P.src.n_bases) ? P.start_mask[w] : 0u;
	}
	if (tid == 0) {
		if (CODES) { s_codes[HT_VEC] = 0; s_codes[HT_VEC + 1] = 0; }
		s_bad[HT_LOAD / 32] = 0xFFFFFFFFu; s_bad[HT_LOAD / 32 + 1] = 0xFFFFFFFFu;
		s_start[HT_LOAD / 32 + 1] = 0;
	}
}

// valid windows of 32 consecutive start positions of the tile, as a bit mask (positions at and beyond n_pos excluded)
__device__ __forceinline__ uint32_t ft_ok_word(const FtTileParams& P, const uint32_t* s_bad, const uint32_t* s_start, uint64_t rel0, uint32_t v)
{
	const uint64_t first = rel0 + 32ull * v;
	if (first >= P.n_pos) return 0u;
	uint32_t ok = window_ok_word(s_bad, s_start, v, P.k);
	if (first + 32 > P.n_pos) ok &= (1u << (uint32_t)(P.n_pos - first)) - 1u;
	return ok;
}

__global__ void __launch_bounds__(HT_THREADS)
ft_count_kernel(const FtTileParams P)
{
	__shared__ uint32_t s_bad[HT_LOAD / 32 + 2], s_start[HT_LOAD / 32 + 2], s_n[HT_THREADS / 32];
	const uint32_t tid = threadIdx.x;
	const uint64_t tile = (uint64_t)blockIdx.x + P.tile0;
	const uint64_t rel0 = tile * HT_POS;
	ft_stage_tile<false>(P, P.pos0 + rel0, nullptr, s_bad, s_start);
	__syncthreads();
	// 32 start positions per thread at once (window_ok_word): the first 64 threads cover the tile
	uint32_t n = (tid < (uint32_t)(HT_POS / 32)) ? (uint32_t)__popc(ft_ok_word(P, s_bad, s_start, rel0, tid)) : 0u;
	for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xFFFFFFFFu, n, o);
	if ((tid & 31) == 0) s_n[tid >> 5] = n;
	__syncthreads();
	if (tid == 0) {
		uint32_t t = 0;
		for (int w = 0; w < HT_THREADS / 32; ++w) t += s_n[w];
		P.tile_cnt[tile] = t;
	}
}

// meta[0] = ordinals of the sub-batch so far (the scan is launched once per piece of the host feed, in stream order)
__global__ void __launch_bounds__(1024)
ft_scan_kernel(uint32_t* __restrict__ tile_cnt, uint32_t tile0, uint32_t n_tiles, uint32_t* __restrict__ meta)
{
	__shared__ uint32_t s_warp[32];
	const uint32_t tid = threadIdx.x;
	uint32_t carry = meta[0];
	for (uint32_t i0 = 0; i0 < n_tiles; i0 += 1024) {
		const uint32_t i = i0 + tid;
		const uint32_t c = (i < n_tiles) ? tile_cnt[tile0 + i] : 0u;
		uint32_t total;
		const uint32_t ex = block_exclusive_scan_1024(c, total, s_warp);
		if (i < n_tiles) tile_cnt[tile0 + i] = carry + ex;
		carry += total;
	}
	__syncthreads();
	if (tid == 0) meta[0] = carry;
}

__global__ void __launch_bounds__(HT_THREADS)
ft_hash_kernel(const FtTileParams P)
{
	__shared__ uint32_t s_codes[HT_VEC + 2], s_bad[HT_LOAD / 32 + 2], s_start[HT_LOAD / 32 + 2];
	__shared__ uint32_t s_pref[HT_POS / 32 + 1];          // valid windows before every 32-position group of the tile
	__shared__ uint32_t s_okw[HT_POS / 32];               // which windows of the group are k-mers
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t k = P.k;
	const uint64_t tile = (uint64_t)blockIdx.x + P.tile0;
	const uint64_t rel0 = tile * HT_POS;
	ft_stage_tile<true>(P, P.pos0 + rel0, s_codes, s_bad, s_start);
	__syncthreads();
	// which windows are k-mers: one word of 32 start positions per thread (window_ok_word); group g = positions 32 g ..
	if (tid < (uint32_t)(HT_POS / 32)) {
		const uint32_t ok = ft_ok_word(P, s_bad, s_start, rel0, tid);
		s_okw[tid] = ok;
		s_pref[tid] = __popc(ok);
	}
	__syncthreads();
	if (warp == 0) {
		// exclusive prefix over the 64 groups (two per lane)
		const uint32_t a = s_pref[2 * lane], b = s_pref[2 * lane + 1];
		uint32_t x = a + b;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
			if (lane >= (uint32_t)o) x += y;
		}
		const uint32_t ex = x - (a + b);
		s_pref[2 * lane] = ex; s_pref[2 * lane + 1] = ex + a;
	}
	__syncthreads();
	const uint32_t ord_tile = P.tile_cnt[tile];
#pragma unroll 1
	for (uint32_t it = 0; it < (uint32_t)HT_IT; ++it) {
		const uint32_t m = s_okw[it * (HT_THREADS / 32) + warp];       // (position it * 256 + tid lies in group it * 8 + warp, bit lane)
		const bool ok = (m >> lane) & 1u;
		if (ok) {
			const uint32_t p = it * HT_THREADS + tid;
			const uint32_t ord = ord_tile + s_pref[it * (HT_THREADS / 32) + warp] + __popc(m & ((1u << lane) - 1u));
			const Canon c = canonical(window_sense(s_codes, p, k), k);
			uint32_t h[4];
			murmur3_multi<4>(c.low, k, h);
			uint4 v = make_uint4(h[0] & P.count_mask, h[1] & P.count_mask, h[2] & P.count_mask, h[3] & P.count_mask);
			// both hashes of one table on one slot (reference: the counter is read once and incremented twice,
			// make_bloom.cpp:553-554,586-592): one record, and the twin touch shares its fate -- it is charged as a
			// loss right away (if the record wins, the occurrence is valid with or without it)
			if (v.y == v.x) { v.y = FT_TWIN; atomicAdd(&P.loss[ord >> 3], 1u << ((ord & 7u) << 2)); }
			if (v.w == v.z) { v.w = FT_TWIN; atomicAdd(&P.loss[ord >> 3], 1u << ((ord & 7u) << 2)); }
			P.hm[ord] = v;
			const uint64_t at = P.list_base + ord;
			uint64_t* chunk = P.list_chunks[at >> FT_LIST_LOG2];
			chunk[at & ((1ull << FT_LIST_LOG2) - 1)] = c.word;
			// seed 2's hash, whole: with three hashes it is the one filter bit that does not come out of the touched bitmap
			reinterpret_cast<uint32_t*>(chunk + (1ull << FT_LIST_LOG2))[at & ((1ull << FT_LIST_LOG2) - 1)] = h[2];
		}
	}
}

// ------------------------------------------------------------------------------------------ append
struct FtAppendParams {
	const uint4* hm;             // [ordinal]
	const uint32_t* meta;        // ft_scan_kernel: [0] = ordinals of the sub-batch
	uint32_t lc;
	uint32_t n_buckets;          // power of two <= FT_MAX_BUCKETS
	uint32_t max_chains;         // pitch of an info row
	uint32_t pu_log2;            // units per page, log2
	uint32_t ppc;                // pages per chain pool
	uint8_t* pool;               // [chain][ppc] pages of (48 << pu_log2) bytes
	uint32_t* page_log;          // [chain][ppc] bucket << 20 | page number within (chain, bucket), in allocation order
	uint32_t* plist;             // [chain][ppc] page ids sorted by bucket, each bucket's pages in order
	uint2* info;                 // [bucket][max_chains]: x = plist index of the chain's first page, y = units in the chain
};

static inline size_t ft_append_smem_bytes()
{
	return (size_t)FT_MAX_BUCKETS * FT_RING_BYTES + (size_t)FT_MAX_BUCKETS * (4 + 4 + 2 + 2 + 2) + 64;
}

constexpr uint32_t FT_DIFF_MASK = 0x7FFFFu;      // (records appended - 8 * units flushed) is taken modulo 2^19

// Everything of bucket b that makes whole units goes to the bucket's page chain.  One thread per bucket and round.
__device__ __forceinline__ void ft_flush_bucket(const FtAppendParams& P, uint8_t* s_ring, const uint32_t* s_head, uint32_t* s_page,
	uint16_t* s_tailu, uint16_t* s_npg, uint32_t* s_next_page, uint32_t chain, uint32_t b, bool pad_rest)
{
	const uint32_t pu = 1u << P.pu_log2;
	const uint32_t head = s_head[b];
	uint32_t tu = s_tailu[b];
	uint32_t present = (head - (tu << 3)) & FT_DIFF_MASK;
	uint8_t* rg = s_ring + (size_t)b * FT_RING_BYTES;
	if (pad_rest && (present & 7u)) {
		// end of the block's range: the last, partial unit is padded with null records
		for (uint32_t i = head; i & 7u; ++i) {
			reinterpret_cast<uint32_t*>(rg)[i & 15u] = FT_NULL_LO;
			reinterpret_cast<uint16_t*>(rg + 64)[i & 15u] = (uint16_t)FT_NULL_HI;
		}
		present = (present + 7u) & ~7u;
	}
	if (present < (uint32_t)FT_UNIT) return;
	uint32_t ps = s_page[b];
	uint32_t fill = ps & 63u, page = ps >> 6;
	do {
		if (fill == pu) {
			page = atomicAdd(s_next_page, 1u);
			const uint32_t seq = s_npg[b];                   // (the host sizes the pages so that a chain has fewer than 2^16 of them)
			s_npg[b] = (uint16_t)(seq + 1u);
			P.page_log[(uint64_t)chain * P.ppc + page] = (b << FT_SEQ_BITS) | seq;
			fill = 0;
		}
		const uint32_t m = tu & 1u;
		const uint4 l0 = *reinterpret_cast<const uint4*>(rg + 32 * m);
		const uint4 l1 = *reinterpret_cast<const uint4*>(rg + 32 * m + 16);
		const uint4 h0 = *reinterpret_cast<const uint4*>(rg + 64 + 16 * m);
		uint4* dst = reinterpret_cast<uint4*>(P.pool + (((uint64_t)chain * P.ppc + page) * pu + fill) * FT_UNIT_BYTES);
		dst[0] = l0; dst[1] = l1; dst[2] = h0;
		++fill; ++tu;
		present -= FT_UNIT;
	} while (present >= (uint32_t)FT_UNIT);
	s_page[b] = (page << 6) | fill;
	s_tailu[b] = (uint16_t)tu;
}

__global__ void __launch_bounds__(FT_THREADS, 1)
ft_append_kernel(const FtAppendParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint8_t* s_ring = smem_raw;                                                        // [FT_MAX_BUCKETS][96]
	uint32_t* s_head = reinterpret_cast<uint32_t*>(smem_raw + (size_t)FT_MAX_BUCKETS * FT_RING_BYTES);   // records appended (free running)
	uint32_t* s_page = s_head + FT_MAX_BUCKETS;                                        // page << 6 | units in it
	uint16_t* s_tailu = reinterpret_cast<uint16_t*>(s_page + FT_MAX_BUCKETS);          // units flushed (mod 2^16)
	uint16_t* s_npg = s_tailu + FT_MAX_BUCKETS;                                        // pages of the bucket so far
	uint16_t* s_items = s_npg + FT_MAX_BUCKETS;                                        // buckets with a whole unit this round
	uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_items + FT_MAX_BUCKETS);          // [0] next page of the pool, [1..2] item counts by round parity

	const uint32_t tid = threadIdx.x;
	const uint32_t chain = blockIdx.x;
	const uint32_t pu = 1u << P.pu_log2;
	const uint32_t nb = P.n_buckets;

	for (uint32_t b = tid; b < (uint32_t)FT_MAX_BUCKETS; b += FT_THREADS) {
		s_head[b] = 0; s_page[b] = pu; s_tailu[b] = 0; s_npg[b] = 0;
	}
	if (tid < 3) s_misc[tid] = 0;
	__syncthreads();

	// my stretch of the sub-batch's ordinals; rounds are aligned to multiples of 1024 ordinals (the resolver's ordering unit)
	const uint32_t n_ok = P.meta[0], ord0 = 0;
	const uint32_t per = (((n_ok + gridDim.x - 1) / gridDim.x) + FT_SUB - 1) & ~(uint32_t)(FT_SUB - 1);
	const uint32_t o_begin = ord0 + min(n_ok, blockIdx.x * per);
	const uint32_t o_end = ord0 + min(n_ok, (blockIdx.x + 1) * per);
	uint32_t round = 0;

	const uint32_t lane = tid & 31u;
	uint4 v_next[FT_PER];
#pragma unroll
	for (int i = 0; i < FT_PER; ++i) {
		const uint32_t o = (o_begin & ~(uint32_t)(FT_SUB - 1)) + i * FT_THREADS + tid;
		v_next[i] = make_uint4(FT_TWIN, FT_TWIN, FT_TWIN, FT_TWIN);
		if (o >= o_begin && o < o_end) v_next[i] = ld_nc_v4(P.hm + o);
	}
#pragma unroll 1
	for (uint32_t a = o_begin & ~(uint32_t)(FT_SUB - 1); a < o_end; a += FT_SUB) {
		// a touch is kept as its hash (FT_TWIN: nothing to append); bucket and record word are derived when needed
		uint32_t hm[FT_REC], pend = 0;
#pragma unroll
		for (int i = 0; i < FT_PER; ++i) {
			const uint32_t o = a + i * FT_THREADS + tid;
			const uint4 v = v_next[i];
			hm[4 * i] = v.x; hm[4 * i + 1] = v.y; hm[4 * i + 2] = v.z; hm[4 * i + 3] = v.w;
			// the next round's hashes travel during this one
			v_next[i] = make_uint4(FT_TWIN, FT_TWIN, FT_TWIN, FT_TWIN);
			if (o + FT_SUB < o_end) v_next[i] = ld_nc_v4(P.hm + o + FT_SUB);
		}
#pragma unroll
		for (int j = 0; j < FT_REC; ++j) pend |= (hm[j] != FT_TWIN) ? 1u << j : 0u;
#define FT_SLOT(j) ((((uint32_t)(((j) >> 1) & 1)) << P.lc) | hm[j])
#define FT_BKT(j) ((FT_SLOT(j) >> FT_BUCKET_LOG2) & (nb - 1u))
		// append what fits into the rings; the append that completes a bucket's first whole unit of the round puts the
		// bucket on the flush list; the list is flushed one bucket per thread; again if a ring was full
		while (true) {
			uint32_t* items_n = &s_misc[1 + (round & 1u)];
			// (the ring counters first, then the stores: independent shared-memory operations in flight together; tried: a
			// warp-uniform straight-line path for the common case "every lane has four records and no ring is full" --
			// 2.74 vs 2.69 ms, the per-record predication is not what this kernel waits for)
			uint32_t old[FT_REC], tl[FT_REC];
#pragma unroll
			for (int j = 0; j < FT_REC; ++j) {
				old[j] = 0; tl[j] = 0;
				if ((pend >> j) & 1u) { const uint32_t b = FT_BKT(j); old[j] = atomicAdd(&s_head[b], 1u); tl[j] = s_tailu[b]; }
			}
			uint32_t pushm = 0;
#pragma unroll
			for (int j = 0; j < FT_REC; ++j) {
				if ((pend >> j) & 1u) {
					const uint32_t slot = FT_SLOT(j), b = (slot >> FT_BUCKET_LOG2) & (nb - 1u);
					const uint32_t d = (old[j] - (tl[j] << 3)) & FT_DIFF_MASK;
					if (d < (uint32_t)FT_RING) {
						const uint32_t o = a + (j >> 2) * FT_THREADS + tid;
						uint8_t* r = s_ring + (size_t)b * FT_RING_BYTES;
						reinterpret_cast<uint32_t*>(r)[old[j] & 15u] = (slot & ((1u << FT_BUCKET_LOG2) - 1u)) | ((o & 0xFFFu) << FT_BUCKET_LOG2);
						reinterpret_cast<uint16_t*>(r + 64)[old[j] & 15u] = (uint16_t)(o >> 12);
						pend &= ~(1u << j);
						if (d == (uint32_t)FT_UNIT - 1u) pushm |= 1u << j;
					} else {
						atomicSub(&s_head[b], 1u);
					}
				}
			}
			// flush list: one counter update per warp
			{
				const uint32_t mine = __popc(pushm);
				uint32_t inc = mine;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
					if (lane >= (uint32_t)o) inc += y;
				}
				const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
				if (total) {
					uint32_t base = 0;
					if (lane == 31) base = atomicAdd(items_n, total);
					base = __shfl_sync(0xFFFFFFFFu, base, 31) + inc - mine;
					for (uint32_t m = pushm; m; m &= m - 1u) {
						const int j = __ffs(m) - 1;
						uint32_t b = 0;
#pragma unroll
						for (int jj = 0; jj < FT_REC; ++jj) if (jj == j) b = FT_BKT(jj);
						s_items[base++] = (uint16_t)b;
					}
				}
			}
			__syncthreads();
			const uint32_t n_items = *items_n;
			if (tid == 0) s_misc[1 + ((round + 1u) & 1u)] = 0;
			// (tried: every other thread takes a flush so that all 32 warps share them -- slower, 2.87 vs 2.71 ms: the flushes
			// are issue work, not latency, and half-empty warps double it)
			for (uint32_t i = tid; i < n_items; i += FT_THREADS)
				ft_flush_bucket(P, s_ring, s_head, s_page, s_tailu, s_npg, &s_misc[0], chain, s_items[i], false);
			++round;
			if (!__syncthreads_or(pend != 0u)) break;
		}
#undef FT_SLOT
#undef FT_BKT
	}

	// ---- end of the range: pad and flush what is left in the rings
	__syncthreads();
	for (uint32_t b = tid; b < nb; b += FT_THREADS)
		ft_flush_bucket(P, s_ring, s_head, s_page, s_tailu, s_npg, &s_misc[0], chain, b, true);
	__syncthreads();

	// ---- page lists: prefix sum of the pages per bucket, then every logged page finds its place
	uint32_t* s_first = reinterpret_cast<uint32_t*>(s_ring);            // the rings are free now
	uint32_t* s_warp = s_first + FT_MAX_BUCKETS;
	{
		const uint32_t c0 = s_npg[2 * tid], c1 = s_npg[2 * tid + 1];
		uint32_t total;
		const uint32_t ex = block_exclusive_scan_1024(c0 + c1, total, s_warp);
		s_first[2 * tid] = ex;
		s_first[2 * tid + 1] = ex + c0;
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const uint32_t b = 2 * tid + q;
			if (b < nb) {
				const uint32_t n = q ? c1 : c0;
				const uint32_t units = n ? (n - 1u) * pu + (s_page[b] & 63u) : 0u;
				P.info[(uint64_t)b * P.max_chains + chain] = make_uint2((uint32_t)((uint64_t)chain * P.ppc) + (q ? ex + c0 : ex), units);
			}
		}
	}
	__syncthreads();
	const uint32_t n_alloc = s_misc[0];
	const uint64_t cbase = (uint64_t)chain * P.ppc;
	for (uint32_t i = tid; i < n_alloc; i += FT_THREADS) {
		const uint32_t e = __ldcg(&P.page_log[cbase + i]);
		P.plist[cbase + s_first[e >> FT_SEQ_BITS] + (e & ((1u << FT_SEQ_BITS) - 1u))] = (uint32_t)(cbase + i);
	}
}

// ------------------------------------------------------------------------------------------ resolver
constexpr int FR_THREADS = 1024;
constexpr int FR_WIN_UNITS = 1024;                       // units per window: one per thread, kept in registers
constexpr int FR_CARRY = 4 * FT_SUB;                     // one round can send at most this many records to one bucket
constexpr int FR_CARRY_SMEM = 1024;                      // ... of which this many wait in shared memory, the rest in the block's scratch
constexpr int FR_OWN = FT_UNIT;                          // records of the thread's own unit
constexpr int FR_SLOTS = FR_OWN + 1;                     // + one carried record
constexpr int FR_CONF = 512;                             // late-contender list / claimer list
constexpr int FR_FILT_WORDS = 512;                       // 16384-bit filter of the slots on the late list
constexpr int FR_PAGES = 3072;                           // page ids of the bucket kept in shared memory
constexpr uint32_t FR_NO_TILE = 0xFFFFFFFFu;

struct FtResolveParams {
	const uint8_t* pool;
	const uint32_t* plist;
	const uint2* info;
	uint32_t n_chains, max_chains, n_buckets, pu_log2;
	uint32_t bucket_words;       // words of the touched bitmap per bucket (2^15, fewer when the filters are smaller than a bucket)
	uint32_t* touched;
	uint32_t have_prior;         // 0: first batch after create/reset, the bitmap is known to be all zero
	uint32_t* loss;
	uint2* carry_scratch;        // [block][2][FR_CARRY] carried records beyond FR_CARRY_SMEM (x = low word, y = high half)
};

static inline size_t ft_resolve_smem_bytes()
{
	return (size_t)(1u << FT_BUCKET_LOG2) / 8 + (size_t)2 * FR_CARRY_SMEM * 6 +
	       (size_t)(3 * (FT_MAX_CHAINS + 1) + 1 + FR_PAGES + 2 * 4 * FR_CONF + 2 * FR_FILT_WORDS + 32 + 16) * 4;
}

// record q of a thread: high half (two to a register) and ordinal
#define FR_HI(rh, q) (((q) & 1) ? ((rh)[(q) >> 1] >> 16) : ((rh)[(q) >> 1] & 0xFFFFu))
#define FR_POS(rl, rh, q) ((FR_HI(rh, q) << 12) | ((rl)[q] >> FT_BUCKET_LOG2))

// (a reduction without return value: nothing waits for it; the occurrences that collected four losses are counted and
// marked by ft_finish_kernel)
__device__ __forceinline__ void ft_charge_loss(const FtResolveParams& P, uint32_t ord)
{
	atomicAdd(&P.loss[ord >> 3], 1u << ((ord & 7u) << 2));
}

// Per-window state lives in two copies selected by the parity of the window, so that a window's last phase (settling
// its conflicts) may overlap the first phase of the next one: three block barriers per window.
//   P1  own unit (registers) + one carried record: tail records -> next carry; touched before -> loss; else candidate
//   B1
//   P2  candidates claim their bit; a record that finds it set now is a late contender (list + filter)
//   B2
//   P3a claimers whose slot is in the filter go on the claimer list; next window's tail tile is published; resets
//   B3
//   P3b one thread per listed record: smallest ordinal of a slot wins, the others are charged
// A carry of more than 1024 records (a bucket that took most of a round: low-complexity reads) gets a window of its own
// that takes no new units.
__global__ void __launch_bounds__(FR_THREADS, 1)
ft_resolve_kernel(const FtResolveParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem_raw);                             // 2^15 words
	uint32_t* s_clo = s_bm + (1u << (FT_BUCKET_LOG2 - 5));                              // [2][FR_CARRY_SMEM] carry: low words
	uint16_t* s_chi = reinterpret_cast<uint16_t*>(s_clo + 2 * FR_CARRY_SMEM);           // [2][FR_CARRY_SMEM] carry: high halves
	uint32_t* s_ubase = reinterpret_cast<uint32_t*>(s_chi + 2 * FR_CARRY_SMEM);         // FT_MAX_CHAINS + 1: first unit of every chain
	uint32_t* s_pl0 = s_ubase + FT_MAX_CHAINS + 1;                                      // FT_MAX_CHAINS + 1: plist index of its first page
	uint32_t* s_pbase = s_pl0 + FT_MAX_CHAINS + 1;                                      // FT_MAX_CHAINS + 1: first page of every chain in s_pages
	uint32_t* s_pages = s_pbase + FT_MAX_CHAINS + 1;                                    // FR_PAGES page ids of the bucket, chain after chain
	uint32_t* s_list = s_pages + FR_PAGES + 1;                                          // [2][4][FR_CONF]: late slot, late ord, claimer slot, claimer ord (8-byte aligned)
	uint32_t* s_filt = s_list + 2 * 4 * FR_CONF;                                        // [2][FR_FILT_WORDS]
	uint32_t* s_warp = s_filt + 2 * FR_FILT_WORDS;                                      // 32
	uint32_t* s_misc = s_warp + 32;                                                     // per parity: [0] late, [1] claimers listed, [2] carry, [3] tail tile; [8] round to skip, [9] round carried

	const uint32_t tid = threadIdx.x;
	const uint32_t pu = 1u << P.pu_log2;
	const uint32_t bw = P.bucket_words;
	uint2* g_carry = P.carry_scratch + (size_t)blockIdx.x * 2 * FR_CARRY;

	for (uint32_t b = blockIdx.x; b < P.n_buckets; b += gridDim.x) {
		__syncthreads();                                   // the previous bucket has left shared memory
		// ---- bitmap of the bucket
		uint32_t* g_bm = P.touched + (uint64_t)b * bw;
		if (P.have_prior) {
			for (uint32_t i = tid; i < bw / 4; i += FR_THREADS) reinterpret_cast<uint4*>(s_bm)[i] = reinterpret_cast<const uint4*>(g_bm)[i];
		} else {
			for (uint32_t i = tid; i < bw / 4; i += FR_THREADS) reinterpret_cast<uint4*>(s_bm)[i] = make_uint4(0u, 0u, 0u, 0u);
		}
		for (uint32_t i = tid; i < (uint32_t)(2 * FR_FILT_WORDS); i += FR_THREADS) s_filt[i] = 0;
		if (tid < 10) s_misc[tid] = (tid == 3 || tid == 7 || tid >= 8) ? FR_NO_TILE : 0u;
		// ---- chain table: where every chain's units (and pages) start in the bucket's flat numbering
		{
			uint2 inf = make_uint2(0u, 0u);
			if (tid < P.n_chains) inf = P.info[(uint64_t)b * P.max_chains + tid];
			uint32_t total, ptotal;
			const uint32_t ex = block_exclusive_scan_1024(inf.y, total, s_warp);
			const uint32_t np = (inf.y + pu - 1u) >> P.pu_log2;
			const uint32_t pex = block_exclusive_scan_1024(np, ptotal, s_warp);
			s_ubase[tid] = ex;
			s_pl0[tid] = inf.x;
			s_pbase[tid] = pex;
			if (tid == FR_THREADS - 1) { s_ubase[FT_MAX_CHAINS] = total; s_pl0[FT_MAX_CHAINS] = 0; s_pbase[FT_MAX_CHAINS] = ptotal; }
		}
		__syncthreads();
		const uint32_t n_units = s_ubase[FT_MAX_CHAINS];
		// page ids of the bucket into shared memory (a warp per chain), as many as fit
		for (uint32_t c = tid >> 5; c < P.n_chains; c += FR_THREADS / 32) {
			const uint32_t p0 = s_pbase[c], p1 = s_pbase[c + 1], src = s_pl0[c];
			for (uint32_t i = p0 + (tid & 31u); i < p1 && i < (uint32_t)FR_PAGES; i += 32) s_pages[i] = __ldg(&P.plist[src + (i - p0)]);
		}
		__syncthreads();

		// one unit per thread: flat unit u -> chain (a few steps forward from the last one, else binary search) -> page -> 48 bytes
		uint4 f0, f1, f2;
		uint32_t c_hint = 0;
		auto fetch = [&](uint32_t u0) {
			const uint32_t u = u0 + tid;
			f0 = make_uint4(FT_NULL_LO, FT_NULL_LO, FT_NULL_LO, FT_NULL_LO); f1 = f0; f2 = f0;
			if (u < n_units) {
				uint32_t c = c_hint;
				int steps = 0;
				while (s_ubase[c + 1] <= u && steps < 8) { ++c; ++steps; }
				if (s_ubase[c + 1] <= u) {
					uint32_t lo = c, hi = FT_MAX_CHAINS;           // largest c with s_ubase[c] <= u
					while (hi - lo > 1u) {
						const uint32_t mid = (lo + hi) >> 1;
						if (s_ubase[mid] <= u) lo = mid; else hi = mid;
					}
					c = lo;
				}
				c_hint = c;
				const uint32_t j = u - s_ubase[c];
				const uint32_t pi = s_pbase[c] + (j >> P.pu_log2);
				const uint32_t page = (pi < (uint32_t)FR_PAGES) ? s_pages[pi] : __ldg(&P.plist[s_pl0[c] + (j >> P.pu_log2)]);
				const uint4* src = reinterpret_cast<const uint4*>(P.pool + ((uint64_t)page * pu + (j & (pu - 1u))) * FT_UNIT_BYTES);
				f0 = ld_nc_v4(src); f1 = ld_nc_v4(src + 1); f2 = ld_nc_v4(src + 2);
			}
		};
		// tile the window [u0, u0 + n_new) ends in: the round of its last record, unless the chain (or the bucket) ends there
		auto publish_tail = [&](uint32_t u0, uint32_t par) {
			const uint32_t n_new = min((uint32_t)FR_WIN_UNITS, n_units - u0);
			if (tid == n_new - 1u) {
				const uint32_t lo7 = f1.w, hi7 = f2.w >> 16;
				const bool null7 = (lo7 == FT_NULL_LO && hi7 == FT_NULL_HI);
				const bool last_window = u0 + n_new >= n_units;
				s_misc[4 * par + 3] = (last_window || null7) ? FR_NO_TILE : (((hi7 << 12) | (lo7 >> FT_BUCKET_LOG2)) >> FT_SUB_LOG2);
			}
		};

		if (n_units) {
			fetch(0);
			publish_tail(0, 0);
		}
		__syncthreads();
		uint32_t w = 0;
		uint32_t u0 = 0;
		while (u0 < n_units) {
			const uint32_t par = w & 1u, nxt = par ^ 1u;
			uint32_t* mi = s_misc + 4 * par;
			uint32_t* mn = s_misc + 4 * nxt;
			uint32_t n_carry = mi[2];
			if (n_carry > (uint32_t)FR_CARRY) __trap();        // cannot happen: a round is 1024 ordinals of 4 touches
			// A carry of more than 1024 records: this window takes no new units.  The records of one round are in no order, so
			// the WHOLE round must be settled together: what the units in flight hold of it joins the carry first (and the
			// next window skips those records), then the eight slots of the unit that is not taken hold the carried round.
			const bool carry_only = n_carry > (uint32_t)FR_CARRY_SMEM;
			const uint32_t skip_tile = carry_only ? FR_NO_TILE : s_misc[8];
			if (carry_only) {
				const uint32_t x_tile = s_misc[9];
				const uint32_t lo8[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
				const uint32_t hi4[4] = {f2.x, f2.y, f2.z, f2.w};
#pragma unroll
				for (int q = 0; q < FR_OWN; ++q) {
					const uint32_t lo = lo8[q], hi = (q & 1) ? (hi4[q >> 1] >> 16) : (hi4[q >> 1] & 0xFFFFu);
					if (((((hi << 12) | (lo >> FT_BUCKET_LOG2))) >> FT_SUB_LOG2) == x_tile) {
						const uint32_t at = atomicAdd(&mi[2], 1u);
						if (at < (uint32_t)FR_CARRY_SMEM) { s_clo[par * FR_CARRY_SMEM + at] = lo; s_chi[par * FR_CARRY_SMEM + at] = (uint16_t)hi; }
						else if (at < (uint32_t)FR_CARRY) __stcg(&g_carry[par * FR_CARRY + at], make_uint2(lo, hi));
					}
				}
				if (tid == 0) s_misc[8] = x_tile;
				__syncthreads();
				n_carry = mi[2];
				if (n_carry > (uint32_t)FR_CARRY) __trap();
			}
			const uint32_t tail_tile = carry_only ? FR_NO_TILE : mi[3];
			const bool more = !carry_only && u0 + FR_WIN_UNITS < n_units;

			// ---- P1 (a record is kept as its low word and, two to a register, its high half)
			uint32_t rl[FR_SLOTS], rh[(FR_SLOTS + 1) / 2];
			if (!carry_only) {
				rl[0] = f0.x; rl[1] = f0.y; rl[2] = f0.z; rl[3] = f0.w; rl[4] = f1.x; rl[5] = f1.y; rl[6] = f1.z; rl[7] = f1.w;
				rh[0] = f2.x; rh[1] = f2.y; rh[2] = f2.z; rh[3] = f2.w;
			} else {
				// the whole carry in this one window (its records are in no order, so they must be settled together): the
				// eight slots of the unit that is not taken hold up to 4 x 1024 carried records
#pragma unroll
				for (int q = 0; q < FR_OWN; ++q) rl[q] = FT_NULL_LO;
#pragma unroll
				for (int q = 0; q < FR_OWN / 2; ++q) rh[q] = 0xFFFFFFFFu;
#pragma unroll
				for (int q = 0; q < FR_CARRY / FR_THREADS; ++q) {
					const uint32_t i = tid + q * FR_THREADS;
					uint32_t lo = FT_NULL_LO, hi = FT_NULL_HI;
					if (i < n_carry) {
						if (i < (uint32_t)FR_CARRY_SMEM) { lo = s_clo[par * FR_CARRY_SMEM + i]; hi = s_chi[par * FR_CARRY_SMEM + i]; }
						else { const uint2 r = __ldcg(&g_carry[par * FR_CARRY + i]); lo = r.x; hi = r.y; }
					}
					rl[q] = lo;
					rh[q >> 1] = (q & 1) ? ((rh[q >> 1] & 0xFFFFu) | (hi << 16)) : ((rh[q >> 1] & 0xFFFF0000u) | hi);
				}
				if (tid == 0) mn[3] = mi[3];
			}
			rl[8] = FT_NULL_LO; rh[4] = FT_NULL_HI;
			if (!carry_only && tid < n_carry) { rl[8] = s_clo[par * FR_CARRY_SMEM + tid]; rh[4] = s_chi[par * FR_CARRY_SMEM + tid]; }
			// the next window's units start travelling now: they have the whole window to arrive
			if (more) fetch(u0 + FR_WIN_UNITS);

			uint32_t cand = 0;           // bit q: record q takes part and its slot was untouched before this window
#pragma unroll
			for (int q0 = 0; q0 < FR_SLOTS; q0 += 5) {
				uint32_t bmw[5];             // (several bitmap words first, then the decisions: loads in flight together)
#pragma unroll
				for (int q = q0; q < q0 + 5 && q < FR_SLOTS; ++q) bmw[q - q0] = s_bm[(rl[q] & ((1u << FT_BUCKET_LOG2) - 1u)) >> 5];
#pragma unroll
				for (int q = q0; q < q0 + 5 && q < FR_SLOTS; ++q) {
					const uint32_t lo = rl[q], hi = FR_HI(rh, q);
					const uint32_t ord = (hi << 12) | (lo >> FT_BUCKET_LOG2);
					if ((ord >> FT_SUB_LOG2) == skip_tile) continue;      // settled with its round by the window before
					if ((ord >> FT_SUB_LOG2) == tail_tile) {
						// the rest of the window's last round waits for the next window (in any order: it has none)
						const uint32_t at = atomicAdd(&mn[2], 1u);
						if (at < (uint32_t)FR_CARRY_SMEM) { s_clo[nxt * FR_CARRY_SMEM + at] = lo; s_chi[nxt * FR_CARRY_SMEM + at] = (uint16_t)hi; }
						else if (at < (uint32_t)FR_CARRY) __stcg(&g_carry[nxt * FR_CARRY + at], make_uint2(lo, hi));
					} else if (ord != FT_NULL_ORD) {
						if ((bmw[q - q0] >> (lo & 31u)) & 1u) ft_charge_loss(P, ord);
						else cand |= 1u << q;
					}
				}
			}
			__syncthreads();                                                            // B1

			// ---- P2: claim the slots; whoever finds the bit set now met another record of this window
			uint32_t late = 0;
			uint32_t* l_slot = s_list + par * 4 * FR_CONF;
			uint32_t* l_pos = l_slot + FR_CONF;
			uint32_t* c_slot = l_pos + FR_CONF;
			uint32_t* c_pos = c_slot + FR_CONF;
			uint32_t* filt = s_filt + par * FR_FILT_WORDS;
			// (several claims first, then their outcomes: independent shared-memory atomics in flight together)
#pragma unroll
			for (int q0 = 0; q0 < FR_SLOTS; q0 += 5) {
				uint32_t seen[5];
#pragma unroll
				for (int q = q0; q < q0 + 5 && q < FR_SLOTS; ++q) {
					const uint32_t slot = rl[q] & ((1u << FT_BUCKET_LOG2) - 1u);
					seen[q - q0] = 0;
					if ((cand >> q) & 1u) seen[q - q0] = atomicOr(&s_bm[slot >> 5], 1u << (slot & 31u));
				}
#pragma unroll
				for (int q = q0; q < q0 + 5 && q < FR_SLOTS; ++q) {
					const uint32_t slot = rl[q] & ((1u << FT_BUCKET_LOG2) - 1u), bit = 1u << (slot & 31u);
					if (((cand >> q) & 1u) && (seen[q - q0] & bit)) {
						late |= 1u << q;
						const uint32_t c = atomicAdd(&mi[0], 1u);
						if (c < (uint32_t)FR_CONF) { l_slot[c] = slot; l_pos[c] = FR_POS(rl, rh, q); }
						atomicOr(&filt[(slot >> 5) & (FR_FILT_WORDS - 1)], bit);
					}
				}
			}
			__syncthreads();                                                            // B2
			const uint32_t n_late = mi[0];
			// ---- P3a
			if (n_late && n_late <= (uint32_t)FR_CONF) {
				const uint32_t claim = cand & ~late;
				uint32_t fw[FR_SLOTS];
#pragma unroll
				for (int q = 0; q < FR_SLOTS; ++q) fw[q] = filt[((rl[q] & ((1u << FT_BUCKET_LOG2) - 1u)) >> 5) & (FR_FILT_WORDS - 1)];
#pragma unroll
				for (int q = 0; q < FR_SLOTS; ++q) {
					if (((claim >> q) & 1u) && ((fw[q] >> (rl[q] & 31u)) & 1u)) {
						const uint32_t c = atomicAdd(&mi[1], 1u);
						if (c < (uint32_t)FR_CONF) { c_slot[c] = rl[q] & ((1u << FT_BUCKET_LOG2) - 1u); c_pos[c] = FR_POS(rl, rh, q); }
					}
				}
			}
			// housekeeping for the next window: its lists, filter and tail tile; the carry buffer this window read is free again
			if (tid == 0) { mn[0] = 0; mn[1] = 0; mi[2] = 0; if (!carry_only) { s_misc[8] = FR_NO_TILE; s_misc[9] = tail_tile; } }
			if (tid < (uint32_t)FR_FILT_WORDS) s_filt[nxt * FR_FILT_WORDS + tid] = 0;
			if (more) publish_tail(u0 + FR_WIN_UNITS, nxt);
			__syncthreads();                                                            // B3
			const uint32_t n_claim = mi[1];
			if (n_late && n_late <= (uint32_t)FR_CONF && n_claim <= (uint32_t)FR_CONF) {
				// ---- P3b: a warp per listed record i, its lanes over the listed records c: i loses to an earlier c on its slot
				// (a claimer can only lose to a late contender: there is one claimer per slot)
				const uint32_t n_all = n_late + n_claim;
				for (uint32_t i = tid >> 5; i < n_all; i += FR_THREADS / 32) {
					const uint32_t si = (i < n_late) ? l_slot[i] : c_slot[i - n_late], pi = (i < n_late) ? l_pos[i] : c_pos[i - n_late];
					bool lose = false;
					for (uint32_t c = tid & 31u; c < n_late; c += 32) lose = lose || (l_slot[c] == si && l_pos[c] < pi);
					if (i < n_late) for (uint32_t c = tid & 31u; c < n_claim; c += 32) lose = lose || (c_slot[c] == si && c_pos[c] < pi);
					if (__any_sync(0xFFFFFFFFu, lose) && (tid & 31u) == 0) ft_charge_loss(P, pi);
				}
			} else if (n_late) {
				// ---- thousands of records on a few slots (low-complexity reads): min-reduction rounds over all the
				// records that found their slot untouched; table entry = slot << 28 | ordinal, the smallest wins the
				// entry, every record of that slot is settled, the others try again under another hash
				unsigned long long* tbl = reinterpret_cast<unsigned long long*>(l_slot);        // 4 * FR_CONF words = 1024 entries, 512 used
				uint32_t open = cand;
				for (uint32_t round = 0; ; ++round) {
					__syncthreads();
					if (tid < (uint32_t)FR_CONF) tbl[tid] = ~0ull;
					__syncthreads();
#pragma unroll
					for (int q = 0; q < FR_SLOTS; ++q) {
						if ((open >> q) & 1u) {
							const uint32_t slot = rl[q] & ((1u << FT_BUCKET_LOG2) - 1u);
							const uint32_t hsh = ((slot * 0x9E3779B1u) >> ((round % 20u) + 3u)) & (FR_CONF - 1);
							atomicMin(&tbl[hsh], ((unsigned long long)slot << 28) | FR_POS(rl, rh, q));
						}
					}
					__syncthreads();
#pragma unroll
					for (int q = 0; q < FR_SLOTS; ++q) {
						if ((open >> q) & 1u) {
							const uint32_t slot = rl[q] & ((1u << FT_BUCKET_LOG2) - 1u), pos = FR_POS(rl, rh, q);
							const uint32_t hsh = ((slot * 0x9E3779B1u) >> ((round % 20u) + 3u)) & (FR_CONF - 1);
							const unsigned long long e = tbl[hsh];
							if ((uint32_t)(e >> 28) == slot) {
								if ((uint32_t)(e & 0xFFFFFFFu) != pos) ft_charge_loss(P, pos);
								open &= ~(1u << q);
							}
						}
					}
					if (!__syncthreads_or(open != 0u)) break;
				}
			}
			++w;
			if (!carry_only) u0 += FR_WIN_UNITS;
		}
		__syncthreads();
		if (s_misc[4 * (w & 1u) + 2]) __trap();                // the last window has no tail
		// ---- the bitmap now holds the earlier batches' bits plus every slot first touched here
		if (n_units || !P.have_prior)
			for (uint32_t i = tid; i < bw / 4; i += FR_THREADS) reinterpret_cast<uint4*>(g_bm)[i] = reinterpret_cast<const uint4*>(s_bm)[i];
	}
}

// After the resolver: occurrences that lost all four touches are marked in the accession's invalid bitmap (bit list_base + o)
// and taken out of the valid count; the list length and the valid count advance by the ordinals of the sub-batch.
__global__ void __launch_bounds__(256)
ft_finish_kernel(const uint32_t* __restrict__ loss, const uint32_t* __restrict__ meta, uint64_t list_base, uint32_t* __restrict__ invalid,
	unsigned long long* __restrict__ counter)
{
	const uint32_t n_ord = meta[0];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) { atomicAdd(&counter[0], (unsigned long long)n_ord); atomicAdd(&counter[2], (unsigned long long)n_ord); }
	if ((uint64_t)i * 8 >= n_ord) return;
	uint32_t w = loss[i];
	w &= 0x44444444u;                                        // nibble == 4 <=> bit 2 set (a nibble never exceeds 4)
	if (w) atomicAdd(&counter[0], 0ull - (unsigned long long)__popc(w));
	while (w) {
		const uint32_t j = (uint32_t)(__ffs(w) - 1) >> 2;
		w &= w - 1u;
		const uint64_t at = list_base + (uint64_t)i * 8 + j;
		atomicOr(&invalid[at >> 5], 1u << (at & 31u));
	}
}

} // namespace kwg
