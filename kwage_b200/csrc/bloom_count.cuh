// Counting-mode construction (reference make_bloom.cpp:506-621, min_kmer_count == 1) as a radix
// partition of "touch" records followed by first-touch resolution in shared memory.
//
// Why this shape (measured on B200, profiles/ubench): a random atomic into an HBM-resident table
// runs at ~2.0e10 /s (DRAM sector read-modify-write), into an L2-resident table at ~1.3e11 /s, into
// shared memory at ~1.8e12 /s.  Every k-mer occurrence touches 4 pseudo-random slots of the two
// counting filters (2^(lc+1) slots: 2^31 for a 1e6-read accession), so the table can never be
// cache resident.  Instead each touch becomes an 8-byte record (slot, stream position) and the
// records are brought to the slots:
//
//   K1 partition_scan_kernel   tile of 2048 k-mer start positions: encode, canonical k-mer, 4 hashes,
//                              counting sort of the <= 8192 records by level-1 bucket in shared memory,
//                              coalesced copy-out of the sorted tile + one row of run offsets.
//                              No global atomics: every tile owns a fixed 64 KiB output window.
//   K2 regroup_kernel          (only when there are more than 512 final buckets) block (i, g) gathers
//                              the runs of level-1 bucket i from a group of tiles, sorts them by
//                              level-2 bucket in shared memory and writes one dense chunk at an exact
//                              (prefix-summed) position.
//   K3 resolve_kernel          one final bucket = 2^15 slots = a 128 KiB shared-memory tile of
//                              "smallest stream position that touched this slot": tile initialised from
//                              the persistent touched-bitmap, records streamed through atomicMin,
//                              every touch that is not (or stops being) the first toucher adds 1 to the
//                              4-bit loss counter of its occurrence; the bitmap is written back.
//   pass B (kmer_scan_kernel)  occurrence is valid <=> loss counter < 4  (it is the first toucher of at
//                              least one of its four slots, i.e. the reference read a zero counter).
//
// All run/chunk positions are exact (prefix sums of counts): memory use is deterministic and heavy
// duplication (poly-G reads, adapters) only makes one bucket longer, never overflows anything.
#pragma once
#include "common.cuh"

namespace kwg {

constexpr int PT_THREADS = 512;                   // partition_scan_kernel block size
constexpr int RG_THREADS = 512;                   // regroup_kernel block size
constexpr int PT_POS = 2048;                      // k-mer start positions per partition tile
constexpr int PT_REC = 4 * PT_POS;                // record slots per tile
constexpr int PT_LOAD = PT_POS + 32;              // bases staged per tile (halo >= k-1, 16-byte granular)
constexpr int PT_VEC = PT_LOAD / 16;
constexpr int FINAL_LOG2 = 15;                    // slots per final bucket (shared-memory tile of u32)
constexpr int FINAL_SLOTS = 1 << FINAL_LOG2;
constexpr int MAX_FAN_LOG2 = 9;                   // at most 512 buckets per partition level
constexpr int MAX_FAN = 1 << MAX_FAN_LOG2;
constexpr int CHUNK_REC = 8192;                   // records per level-2 chunk
constexpr int RS_THREADS = 1024;                  // resolve kernel block size
constexpr int RS_U = 8;                           // runs in flight per warp in the resolve kernel
constexpr uint32_t SLOT_EMPTY = 0xFFFFFFFFu;
constexpr uint64_t MAX_COUNT_POS = 1ull << 28;    // positions per counting sub-batch (record: 28-bit position)

struct CountGeom {
	uint32_t lc;             // log2 counting-filter length (per table)
	uint32_t count_mask;
	uint32_t nb_log2;        // log2(number of final buckets) = lc + 1 - FINAL_LOG2
	uint32_t f1_log2;        // level-1 fan-out
	uint32_t f2_log2;        // level-2 fan-out (0: single level)
};

static inline CountGeom count_geometry(uint32_t lc)
{
	CountGeom g{};
	g.lc = lc;
	g.count_mask = (lc >= 32) ? 0xFFFFFFFFu : ((1u << lc) - 1u);
	g.nb_log2 = lc + 1 - FINAL_LOG2;
	if (g.nb_log2 <= (uint32_t)MAX_FAN_LOG2) { g.f1_log2 = g.nb_log2; g.f2_log2 = 0; }
	else { g.f1_log2 = (g.nb_log2 + 1) / 2; g.f2_log2 = g.nb_log2 - g.f1_log2; }
	return g;
}

struct PartParams {
	const char* bases;           // device, 16-byte aligned (whole batch)
	uint64_t n_bases;
	const uint32_t* start_mask;
	uint32_t k;
	uint64_t pos0;               // absolute base index of the first start position of this sub-batch (multiple of 16)
	uint64_t n_pos;              // start positions in this sub-batch (<= 2^28)
	uint32_t count_mask;
	uint32_t f1_log2;
	uint32_t shift1;             // FINAL_LOG2 + f2_log2: record keeps the slot bits below this
	uint32_t table_shift;        // f1_log2 - 1: where the table index (0: first, 1: second) lands in the level-1 bucket
	uint32_t ntp;                // row pitch of offs1 (tiles, padded)
	uint64_t* rec1;              // [n_tiles][PT_REC]
	uint16_t* offs1;             // [(F1+1)][ntp]: start of bucket b inside the sorted tile; row F1 = record count
};

// exclusive prefix sum of n <= MAX_FAN counters in shared memory (in place) by a block of THREADS
// threads (256: two counters per thread, 512: one); returns the total.  s_warp: THREADS/32 words of scratch.
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan_512(uint32_t* v, uint32_t n, uint32_t* s_warp)
{
	constexpr int ITEMS = MAX_FAN / THREADS;
	static_assert(ITEMS == 1 || ITEMS == 2, "block size must be 256 or 512");
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t a = (ITEMS * tid < n) ? v[ITEMS * tid] : 0u;
	const uint32_t b = (ITEMS == 2 && 2 * tid + 1 < n) ? v[2 * tid + 1] : 0u;
	uint32_t x = a + b;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
		if (lane >= (uint32_t)o) x += y;
	}
	if (lane == 31) s_warp[warp] = x;
	__syncthreads();
	uint32_t base = 0, total = 0;
#pragma unroll
	for (int w = 0; w < THREADS / 32; ++w) {
		const uint32_t s = s_warp[w];
		if ((uint32_t)w < warp) base += s;
		total += s;
	}
	const uint32_t excl = base + x - (a + b);
	if (ITEMS * tid < n) v[ITEMS * tid] = excl;
	if (ITEMS == 2 && 2 * tid + 1 < n) v[2 * tid + 1] = excl + a;
	__syncthreads();
	return total;
}

// ------------------------------------------------------------------------------------------ K1
__global__ void __launch_bounds__(PT_THREADS, 2)
partition_scan_kernel(const PartParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint64_t* s_sorted = reinterpret_cast<uint64_t*>(smem_raw);                        // PT_REC records
	uint32_t* s_hm = reinterpret_cast<uint32_t*>(smem_raw + PT_REC * 8);               // [4][PT_POS]
	uint32_t* s_codes = s_hm + 4 * PT_POS;                                             // PT_VEC + 2
	uint32_t* s_bad = s_codes + PT_VEC + 2;                                            // PT_LOAD/32 + 2
	uint32_t* s_start = s_bad + PT_LOAD / 32 + 2;                                      // PT_LOAD/32 + 2
	uint32_t* s_ok = s_start + PT_LOAD / 32 + 2;                                       // PT_POS/32
	uint32_t* s_hist = s_ok + PT_POS / 32;                                             // MAX_FAN + 1
	uint32_t* s_cursor = s_hist + MAX_FAN + 1;                                         // MAX_FAN
	uint32_t* s_warp = s_cursor + MAX_FAN;                                             // 16

	const uint32_t tid = threadIdx.x;
	const uint32_t k = P.k;
	const uint64_t tile = blockIdx.x;
	const uint64_t rel0 = tile * PT_POS;               // sub-batch relative position of the tile
	const uint64_t t0 = P.pos0 + rel0;                 // absolute base index
	const uint32_t F1 = 1u << P.f1_log2;

	for (uint32_t v = tid; v < (uint32_t)PT_VEC; v += PT_THREADS) {
		const uint64_t g = t0 + (uint64_t)v * 16;
		uint32_t codes = 0, bad16 = 0xFFFFu;
		if (g + 16 <= P.n_bases) {
			encode16(ld_nc_v4(P.bases + g), codes, bad16);
		} else if (g < P.n_bases) {
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
	for (uint32_t v = tid; v < (uint32_t)(PT_LOAD / 32 + 1); v += PT_THREADS) {
		const uint64_t w = (t0 >> 5) + v;
		s_start[v] = (w * 32 < P.n_bases) ? P.start_mask[w] : 0u;
	}
	for (uint32_t v = tid; v <= F1; v += PT_THREADS) s_hist[v] = 0;
	if (tid == 0) {
		s_codes[PT_VEC] = 0; s_codes[PT_VEC + 1] = 0;
		s_bad[PT_LOAD / 32] = 0xFFFFFFFFu; s_bad[PT_LOAD / 32 + 1] = 0xFFFFFFFFu;
		s_start[PT_LOAD / 32 + 1] = 0;
	}
	__syncthreads();

	// ---- sweep 1: hashes -> parked slot indices + level-1 histogram
#pragma unroll 1
	for (uint32_t it = 0; it < (uint32_t)(PT_POS / PT_THREADS); ++it) {
		const uint32_t p = it * PT_THREADS + tid;
		const bool ok = (rel0 + p < P.n_pos) && window_ok(s_bad, s_start, p, k);
		if (ok) {
			const Canon c = canonical(window_sense(s_codes, p, k), k);
			uint32_t h[4];
			murmur3_multi<4>(c.low, k, h);
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t hm = h[j] & P.count_mask;
				s_hm[j * PT_POS + p] = hm;
				const uint32_t b1 = (hm >> P.shift1) | ((uint32_t)(j >> 1) << P.table_shift);
				atomicAdd(&s_hist[b1], 1u);
			}
		}
		const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
		if ((tid & 31) == 0) s_ok[p >> 5] = m;
	}
	__syncthreads();

	const uint32_t total = block_exclusive_scan_512<PT_THREADS>(s_hist, F1, s_warp);
	for (uint32_t b = tid; b < F1; b += PT_THREADS) {
		const uint32_t s = s_hist[b];
		s_cursor[b] = s;
		P.offs1[(uint64_t)b * P.ntp + tile] = (uint16_t)s;
	}
	if (tid == 0) P.offs1[(uint64_t)F1 * P.ntp + tile] = (uint16_t)total;
	__syncthreads();

	// ---- sweep 2: counting-sort scatter into the staging tile
	const uint32_t low_mask = (1u << P.shift1) - 1u;
#pragma unroll 1
	for (uint32_t it = 0; it < (uint32_t)(PT_POS / PT_THREADS); ++it) {
		const uint32_t p = it * PT_THREADS + tid;
		if ((s_ok[p >> 5] >> (p & 31)) & 1u) {
			const uint32_t pos = (uint32_t)(rel0 + p);
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t hm = s_hm[j * PT_POS + p];
				const uint32_t b1 = (hm >> P.shift1) | ((uint32_t)(j >> 1) << P.table_shift);
				const uint32_t idx = atomicAdd(&s_cursor[b1], 1u);
				s_sorted[idx] = ((uint64_t)(hm & low_mask) << 32) | pos;
			}
		}
	}
	__syncthreads();

	// ---- coalesced copy-out (16 bytes per thread per step; the tile window is 64 KiB aligned)
	uint4* dst = reinterpret_cast<uint4*>(P.rec1 + tile * PT_REC);
	const uint4* src = reinterpret_cast<const uint4*>(s_sorted);
	for (uint32_t i = tid; i < (total + 1) / 2; i += PT_THREADS) st_na_v4(dst + i, src[i]);
}

// ------------------------------------------------------------------------------------------ level-2 bookkeeping
// cnt1[i*NG + g] = records of level-1 bucket i inside tile group g  (one warp per pair)
__global__ void __launch_bounds__(256)
group_count_kernel(const uint16_t* __restrict__ offs1, uint32_t ntp, uint32_t n_tiles, uint32_t F1, uint32_t G1, uint32_t NG,
	uint32_t* __restrict__ cnt1)
{
	const uint32_t idx = blockIdx.x * 8 + (threadIdx.x >> 5);
	const uint32_t lane = threadIdx.x & 31;
	if (idx >= F1 * NG) return;
	const uint32_t i = idx / NG, g = idx % NG;
	const uint32_t t0 = g * G1, t1 = min(t0 + G1, n_tiles);
	uint32_t sum = 0;
	for (uint32_t t = t0 + lane; t < t1; t += 32)
		sum += (uint32_t)offs1[(uint64_t)(i + 1) * ntp + t] - (uint32_t)offs1[(uint64_t)i * ntp + t];
	for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, o);
	if (lane == 0) cnt1[idx] = sum;
}

// Exclusive prefix sums over the F1*NG (bucket-major) group counts: record base of every group, id of
// its first chunk, record base of every chunk, first chunk of every level-1 bucket.  One block.
__global__ void __launch_bounds__(1024)
group_prefix_kernel(const uint32_t* __restrict__ cnt1, uint32_t n, uint32_t NG, uint32_t F1,
	uint64_t* __restrict__ base2, uint32_t* __restrict__ cbase, uint64_t* __restrict__ chunk_rec, uint32_t* __restrict__ cfirst)
{
	__shared__ unsigned long long s_rec[32];
	__shared__ uint32_t s_chk[32];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t per = (n + 1023) / 1024;
	const uint32_t a = min(tid * per, n), b = min(a + per, n);
	unsigned long long rsum = 0;
	uint32_t csum = 0;
	for (uint32_t x = a; x < b; ++x) { const uint32_t c = cnt1[x]; rsum += c; csum += (c + CHUNK_REC - 1) / CHUNK_REC; }
	unsigned long long rinc = rsum;
	uint32_t cinc = csum;
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned long long ry = __shfl_up_sync(0xFFFFFFFFu, rinc, o);
		const uint32_t cy = __shfl_up_sync(0xFFFFFFFFu, cinc, o);
		if (lane >= (uint32_t)o) { rinc += ry; cinc += cy; }
	}
	if (lane == 31) { s_rec[warp] = rinc; s_chk[warp] = cinc; }
	__syncthreads();
	unsigned long long rbase = 0;
	uint32_t cb = 0, ctotal = 0;
	for (uint32_t w = 0; w < 32; ++w) {
		if (w < warp) { rbase += s_rec[w]; cb += s_chk[w]; }
		ctotal += s_chk[w];
	}
	unsigned long long r = rbase + rinc - rsum;
	uint32_t c = cb + cinc - csum;
	for (uint32_t x = a; x < b; ++x) {
		const uint32_t cnt = cnt1[x];
		base2[x] = r;
		cbase[x] = c;
		if (x % NG == 0) cfirst[x / NG] = c;
		const uint32_t nc = (cnt + CHUNK_REC - 1) / CHUNK_REC;
		for (uint32_t q = 0; q < nc; ++q) chunk_rec[c + q] = r + (unsigned long long)q * CHUNK_REC;
		r += cnt; c += nc;
	}
	if (tid == 0) cfirst[F1] = ctotal;
}

// ------------------------------------------------------------------------------------------ K2
struct RegroupParams {
	const uint64_t* rec1;
	const uint16_t* offs1;
	uint32_t ntp, n_tiles;
	uint32_t F1, G1, NG;
	uint32_t f2_log2;
	const uint32_t* cnt1;
	const uint64_t* base2;
	const uint32_t* cbase;
	const uint32_t* cfirst;
	uint64_t* rec2;
	uint16_t* offs2;             // bucket i: rows at cfirst[i]*(F2+1); entry (j, local chunk c) at + j*nci + c
};

// One block regroups the runs of level-1 bucket i coming from a group of G1 <= 512 tiles.  Every warp
// owns RG_R runs per round and keeps their first 32 records in registers: all RG_R loads (256 bytes
// each) are issued back to back, so one DRAM round trip covers the block's whole input, and the
// scatter pass re-uses the registers instead of reading the runs again.  Runs longer than a warp
// (~9 % at the usual 26 records per run) and the first round of 512-tile groups are re-read (L2).
constexpr int RG_R = 16;                                     // runs per warp per round
constexpr int RG_RPR = (RG_THREADS / 32) * RG_R;             // runs per block per round (256)
static_assert(2 * RG_RPR >= MAX_FAN, "two rounds must cover the largest tile group");

__global__ void __launch_bounds__(RG_THREADS, 2)
regroup_kernel(const RegroupParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint64_t* s_sorted = reinterpret_cast<uint64_t*>(smem_raw);             // CHUNK_REC records
	uint32_t* s_vs = reinterpret_cast<uint32_t*>(smem_raw + CHUNK_REC * 8); // MAX_FAN + 1: virtual start of each run
	uint32_t* s_roff = s_vs + MAX_FAN + 1;                                  // MAX_FAN: record offset of the run's part in this chunk
	uint32_t* s_hist = s_roff + MAX_FAN;                                    // MAX_FAN + 1
	uint32_t* s_cursor = s_hist + MAX_FAN + 1;                              // MAX_FAN
	uint32_t* s_warp = s_cursor + MAX_FAN;                                  // RG_THREADS / 32
	uint32_t* s_nlong = s_warp + RG_THREADS / 32;                           // 1: runs with more than 32 records in this chunk
	uint16_t* s_rs = reinterpret_cast<uint16_t*>(s_nlong + 1);              // MAX_FAN: run start inside its tile
	uint16_t* s_rn = s_rs + MAX_FAN;                                        // MAX_FAN: records of the run in this chunk
	uint16_t* s_long = s_rn + MAX_FAN;                                      // MAX_FAN: the long runs

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr uint32_t NW = RG_THREADS / 32;
	const uint32_t i = blockIdx.x % P.F1, g = blockIdx.x / P.F1;
	const uint32_t pair = i * P.NG + g;
	const uint32_t cnt = P.cnt1[pair];
	if (cnt == 0) return;
	const uint32_t t0 = g * P.G1;
	const uint32_t nt = min(P.G1, P.n_tiles - t0);
	const uint32_t F2 = 1u << P.f2_log2;
	const uint64_t* __restrict__ base = P.rec1 + (uint64_t)t0 * PT_REC;

	for (uint32_t r = tid; r < (uint32_t)MAX_FAN; r += RG_THREADS) {
		uint32_t s = 0, len = 0;
		if (r < nt) {
			s = P.offs1[(uint64_t)i * P.ntp + t0 + r];
			len = (uint32_t)P.offs1[(uint64_t)(i + 1) * P.ntp + t0 + r] - s;
		}
		s_rs[r] = (uint16_t)s;
		s_vs[r] = len;
	}
	__syncthreads();
	block_exclusive_scan_512<RG_THREADS>(s_vs, MAX_FAN, s_warp);
	if (tid == 0) s_vs[MAX_FAN] = cnt;
	__syncthreads();

	const uint32_t nchunk = (cnt + CHUNK_REC - 1) / CHUNK_REC;
	const uint32_t c0 = P.cfirst[i];
	const uint32_t nci = P.cfirst[i + 1] - c0;
	const uint32_t cl0 = P.cbase[pair] - c0;
	uint16_t* rows = P.offs2 + (uint64_t)c0 * (F2 + 1);
	const uint64_t out0 = P.base2[pair];
	const uint32_t n_rounds = (nt + RG_RPR - 1) / RG_RPR;       // 1 or 2

	for (uint32_t c = 0; c < nchunk; ++c) {
		const uint32_t cv0 = c * CHUNK_REC, cv1 = min(cnt, cv0 + CHUNK_REC);
		for (uint32_t v = tid; v <= F2; v += RG_THREADS) s_hist[v] = 0;
		if (tid == 0) *s_nlong = 0;
		__syncthreads();
		for (uint32_t r = tid; r < (uint32_t)MAX_FAN; r += RG_THREADS) {
			const uint32_t vs = s_vs[r], ve = s_vs[r + 1];
			const uint32_t lo = max(vs, cv0), hi = min(ve, cv1);
			const uint32_t n = (lo < hi) ? hi - lo : 0u;
			s_rn[r] = (uint16_t)n;
			s_roff[r] = r * PT_REC + s_rs[r] + (lo - vs);
			if (n > 32) s_long[atomicAdd(s_nlong, 1u)] = (uint16_t)r;
		}
		__syncthreads();
		const uint32_t nlong = *s_nlong;

		// pass A: level-2 histogram; the records of the last round stay in registers
		uint64_t rec[RG_R];
		for (uint32_t round = 0; round < n_rounds; ++round) {
			const uint32_t rb = round * RG_RPR + warp;
#pragma unroll
			for (int u = 0; u < RG_R; ++u) {
				const uint32_t r = rb + u * NW;
				rec[u] = (lane < s_rn[r]) ? base[s_roff[r] + lane] : 0ull;
			}
#pragma unroll
			for (int u = 0; u < RG_R; ++u)
				if (lane < s_rn[rb + u * NW]) atomicAdd(&s_hist[(uint32_t)(rec[u] >> 32) >> FINAL_LOG2], 1u);
		}
		for (uint32_t q = warp; q < nlong; q += NW) {
			const uint32_t r = s_long[q], n = s_rn[r];
			const uint64_t* run = base + s_roff[r];
			for (uint32_t x = 32 + lane; x < n; x += 32) atomicAdd(&s_hist[(uint32_t)(run[x] >> 32) >> FINAL_LOG2], 1u);
		}
		__syncthreads();
		block_exclusive_scan_512<RG_THREADS>(s_hist, F2, s_warp);
		for (uint32_t j = tid; j < F2; j += RG_THREADS) {
			const uint32_t s = s_hist[j];
			s_cursor[j] = s;
			rows[(uint64_t)j * nci + cl0 + c] = (uint16_t)s;
		}
		if (tid == 0) rows[(uint64_t)F2 * nci + cl0 + c] = (uint16_t)(cv1 - cv0);
		__syncthreads();

		// pass B: scatter into the staging chunk
		for (uint32_t round = 0; round + 1 < n_rounds; ++round) {          // earlier rounds: re-read (L2)
			const uint32_t rb = round * RG_RPR + warp;
			uint64_t tmp[RG_R];
#pragma unroll
			for (int u = 0; u < RG_R; ++u) {
				const uint32_t r = rb + u * NW;
				tmp[u] = (lane < s_rn[r]) ? base[s_roff[r] + lane] : 0ull;
			}
#pragma unroll
			for (int u = 0; u < RG_R; ++u)
				if (lane < s_rn[rb + u * NW]) s_sorted[atomicAdd(&s_cursor[(uint32_t)(tmp[u] >> 32) >> FINAL_LOG2], 1u)] = tmp[u];
		}
		{
			const uint32_t rb = (n_rounds - 1) * RG_RPR + warp;
#pragma unroll
			for (int u = 0; u < RG_R; ++u)
				if (lane < s_rn[rb + u * NW]) s_sorted[atomicAdd(&s_cursor[(uint32_t)(rec[u] >> 32) >> FINAL_LOG2], 1u)] = rec[u];
		}
		for (uint32_t q = warp; q < nlong; q += NW) {
			const uint32_t r = s_long[q], n = s_rn[r];
			const uint64_t* run = base + s_roff[r];
			for (uint32_t x = 32 + lane; x < n; x += 32) {
				const uint64_t v = run[x];
				s_sorted[atomicAdd(&s_cursor[(uint32_t)(v >> 32) >> FINAL_LOG2], 1u)] = v;
			}
		}
		__syncthreads();
		uint4* dst = reinterpret_cast<uint4*>(P.rec2 + out0 + cv0);
		if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
			const uint4* src = reinterpret_cast<const uint4*>(s_sorted);
			const uint32_t n2 = (cv1 - cv0) / 2;
			for (uint32_t x = tid; x < n2; x += RG_THREADS) st_na_v4(dst + x, src[x]);
			if (tid == 0 && ((cv1 - cv0) & 1u)) P.rec2[out0 + cv1 - 1] = s_sorted[cv1 - cv0 - 1];
		} else {
			uint64_t* d8 = P.rec2 + out0 + cv0;
			for (uint32_t x = tid; x < cv1 - cv0; x += RG_THREADS) d8[x] = s_sorted[x];
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------ K3
struct ResolveParams {
	const uint64_t* rec;         // rec1 (single level) or rec2
	const uint16_t* offs;        // run offsets, rows of bucket group i at cfirst[i]*(F2+1), entry (j, c) at + j*nci + c
	const uint32_t* cfirst;      // [F1 + 1] first chunk of every level-1 bucket (NULL for a single level)
	uint32_t single_nci;         // single level: number of tiles (= runs per bucket)
	const uint64_t* chunk_rec;   // record base of every chunk (ignored when chunk_stride != 0)
	uint64_t chunk_stride;       // single level: chunk c starts at c * PT_REC
	uint64_t row_pitch;          // single level: pitch of an offsets row (ntp); 0: use nci
	uint32_t f2_log2;            // sub-buckets per group (single level: all final buckets are one group)
	uint32_t n_buckets;          // final buckets
	uint32_t* touched;           // persistent bitmap, FINAL_SLOTS bits per final bucket
	uint32_t have_prior;         // 0: first batch after create/reset, the bitmap is known to be all zero
	uint32_t* loss;              // 4-bit loss counters, 8 positions per word
};

__device__ __forceinline__ uint64_t ld_nc_u64(const uint64_t* p)
{
	uint64_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
	return r;
}

__device__ __forceinline__ void resolve_record(uint32_t* s_tile, uint32_t* s_bm, uint32_t* loss, uint64_t rec)
{
	const uint32_t slot = (uint32_t)(rec >> 32) & (FINAL_SLOTS - 1);
	const uint32_t pos = (uint32_t)rec;
	const uint32_t v = pos + 1u;                       // 0 = touched by an earlier batch
	const uint32_t old = atomicMin(&s_tile[slot], v);
	if (old <= v) {
		// an earlier occurrence (or an earlier batch, or the same occurrence through its other hash
		// of this table) holds the slot: this touch read a non-zero counter
		atomicAdd(&loss[pos >> 3], 1u << ((pos & 7u) << 2));
	} else if (old != SLOT_EMPTY) {
		// a later occurrence had been processed first and has just been displaced
		const uint32_t q = old - 1u;
		atomicAdd(&loss[q >> 3], 1u << ((q & 7u) << 2));
	} else {
		atomicOr(&s_bm[slot >> 5], 1u << (slot & 31u));      // first touch of the slot in this accession
	}
}

// what a thread needs of a bucket before it can start: its word of the touched bitmap and run tid
struct BucketPrefetch { uint32_t bmw, len, off; };    // off: record index (< 2^30)

__device__ __forceinline__ BucketPrefetch prefetch_bucket(const ResolveParams& P, uint32_t b)
{
	const uint32_t tid = threadIdx.x;
	const uint32_t F2 = 1u << P.f2_log2;
	const uint32_t i = b >> P.f2_log2, j = b & (F2 - 1);
	const uint32_t c0 = P.cfirst ? P.cfirst[i] : 0u;
	const uint32_t nci = P.cfirst ? P.cfirst[i + 1] - c0 : P.single_nci;
	const uint64_t pitch = P.row_pitch ? P.row_pitch : (uint64_t)nci;
	const uint16_t* row_s = P.offs + (uint64_t)c0 * (F2 + 1) + (uint64_t)j * pitch;
	BucketPrefetch r;
	r.bmw = P.have_prior ? P.touched[(uint64_t)b * (FINAL_SLOTS / 32) + tid] : 0u;
	r.len = 0; r.off = 0;
	if (tid < nci) {
		const uint32_t s = row_s[tid], e = row_s[pitch + tid];
		const uint64_t base = P.chunk_stride ? (uint64_t)(c0 + tid) * P.chunk_stride : P.chunk_rec[c0 + tid];
		r.off = (uint32_t)(base + s);
		r.len = e - s;
	}
	return r;
}

__global__ void __launch_bounds__(RS_THREADS, 1)
resolve_kernel(const ResolveParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem_raw);                           // FINAL_SLOTS
	uint32_t* s_bm = s_tile + FINAL_SLOTS;                                              // FINAL_SLOTS / 32
	uint32_t* s_off = s_bm + FINAL_SLOTS / 32;                                          // RS_THREADS: record index of every run
	uint32_t* s_len = s_off + RS_THREADS;                                               // RS_THREADS
	uint32_t* s_nlong = s_len + RS_THREADS;                                             // 1: runs with more than 32 records
	uint16_t* s_long = reinterpret_cast<uint16_t*>(s_nlong + 1);                        // RS_THREADS: their indices

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t F2 = 1u << P.f2_log2;

	uint32_t b = blockIdx.x;
	BucketPrefetch pf{0u, 0u, 0u};
	if (b < P.n_buckets) pf = prefetch_bucket(P, b);

	for (; b < P.n_buckets; b += gridDim.x) {
		const uint32_t i = b >> P.f2_log2, j = b & (F2 - 1);
		const uint32_t c0 = P.cfirst ? P.cfirst[i] : 0u;
		const uint32_t nci = P.cfirst ? P.cfirst[i + 1] - c0 : P.single_nci;
		const uint64_t pitch = P.row_pitch ? P.row_pitch : (uint64_t)nci;
		const uint16_t* row_s = P.offs + (uint64_t)c0 * (F2 + 1) + (uint64_t)j * pitch;

		s_bm[tid] = pf.bmw;
		s_off[tid] = pf.off;
		s_len[tid] = pf.len;
		if (tid == 0) *s_nlong = 0;
		__syncthreads();
		if (pf.len > 32) s_long[atomicAdd(s_nlong, 1u)] = (uint16_t)tid;
		// tables of the next bucket travel while this one is resolved, and its runs are pulled into L2
		if (b + gridDim.x < P.n_buckets) {
			pf = prefetch_bucket(P, b + gridDim.x);
			if (pf.len) {
				const char* p0 = reinterpret_cast<const char*>(P.rec + pf.off);
				const char* p1 = p0 + (size_t)min(pf.len, 64u) * 8 - 1;
				for (const char* q = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)127); q <= p1; q += 128)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
			}
		}

		// tile <- touched bitmap of this bucket
		if (P.have_prior) {
#pragma unroll
			for (uint32_t q = 0; q < (uint32_t)(FINAL_SLOTS / (4 * RS_THREADS)); ++q) {
				const uint32_t s4 = (q * RS_THREADS + tid) * 4;
				const uint32_t bits = s_bm[s4 >> 5] >> (s4 & 31);
				uint4 v;
				v.x = (bits & 1u) ? 0u : SLOT_EMPTY;
				v.y = (bits & 2u) ? 0u : SLOT_EMPTY;
				v.z = (bits & 4u) ? 0u : SLOT_EMPTY;
				v.w = (bits & 8u) ? 0u : SLOT_EMPTY;
				reinterpret_cast<uint4*>(s_tile)[q * RS_THREADS + tid] = v;
			}
		} else {
			const uint4 e4 = make_uint4(SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY);
#pragma unroll
			for (uint32_t q = 0; q < (uint32_t)(FINAL_SLOTS / (4 * RS_THREADS)); ++q)
				reinterpret_cast<uint4*>(s_tile)[q * RS_THREADS + tid] = e4;
		}
		__syncthreads();

		for (uint32_t cb = 0; cb < nci; cb += RS_THREADS) {
			const uint32_t nrt = min((uint32_t)RS_THREADS, nci - cb);
			if (cb) {          // run tables beyond the prefetched first RS_THREADS runs
				uint32_t off = 0, len = 0;
				if (tid < nrt) {
					const uint32_t c = cb + tid;
					const uint32_t s = row_s[c], e = row_s[pitch + c];
					const uint64_t base = P.chunk_stride ? (uint64_t)(c0 + c) * P.chunk_stride : P.chunk_rec[c0 + c];
					off = (uint32_t)(base + s);
					len = e - s;
				}
				s_off[tid] = off;
				s_len[tid] = len;
				if (tid == 0) *s_nlong = 0;
				__syncthreads();
				if (len > 32) s_long[atomicAdd(s_nlong, 1u)] = (uint16_t)tid;
				__syncthreads();
			}
			// one warp per run, the first 32 records of RS_U runs in flight (s_len is zero beyond nrt)
			for (uint32_t r0 = warp; r0 < nrt; r0 += RS_U * (RS_THREADS / 32)) {
				uint64_t rec[RS_U];
				uint32_t len[RS_U];
#pragma unroll
				for (int u = 0; u < RS_U; ++u) {
					const uint32_t r = (r0 + u * (RS_THREADS / 32)) & (RS_THREADS - 1);
					len[u] = (r >= r0) ? s_len[r] : 0u;
					rec[u] = (lane < len[u]) ? ld_nc_u64(P.rec + s_off[r] + lane) : 0ull;
				}
#pragma unroll
				for (int u = 0; u < RS_U; ++u)
					if (lane < len[u]) resolve_record(s_tile, s_bm, P.loss, rec[u]);
			}
			const uint32_t nlong = *s_nlong;
			for (uint32_t q = warp; q < nlong; q += RS_THREADS / 32) {
				const uint32_t r = s_long[q], len = s_len[r];
				const uint64_t* run = P.rec + s_off[r];
				for (uint32_t x = 32 + lane; x < len; x += 32) resolve_record(s_tile, s_bm, P.loss, ld_nc_u64(run + x));
			}
			__syncthreads();
		}

		// the bitmap now holds the earlier batches' bits plus every slot first touched here
		P.touched[(uint64_t)b * (FINAL_SLOTS / 32) + tid] = s_bm[tid];
		__syncthreads();
	}
}

} // namespace kwg
