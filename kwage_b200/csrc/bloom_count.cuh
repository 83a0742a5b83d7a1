// Counting-mode construction (reference make_bloom.cpp:506-621) as a radix partition of "touch" records
// followed by first-touch resolution in shared memory: once for min_kmer_count == 1, once per counter
// level for min_kmer_count > 1 (resolve_kernel<true>, resolve_dense_kernel, elig_update_kernel below).
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
//   K2 regroup_kernel          (only when there are more than 512 final buckets) persistent, warp-specialised
//                              blocks: a producer group gathers the ~256 runs of one chunk = (level-1 bucket i,
//                              group of tiles) into a staging buffer with bulk async copies + mbarrier, consumer
//                              warps counting-sort it by level-2 bucket in shared memory and bulk-store one
//                              dense chunk at an exact (prefix-summed) position.
//   K3 resolve_kernel          one final bucket = 2^15 slots = a 128 KiB shared-memory tile of
//                              "smallest stream position that touched this slot": tile initialised from
//                              the persistent touched-bitmap, the bucket's runs staged by bulk async copies and
//                              streamed through atomicMin, every touch that is not (or stops being) the first
//                              toucher adds 1 to the 4-bit loss counter of its occurrence; the bitmap is
//                              written back.
//   pass B (kmer_scan_kernel)  occurrence is valid <=> loss counter < 4  (it is the first toucher of at
//                              least one of its four slots, i.e. the reference read a zero counter).
//   min_kmer_count = c > 1     K1/K2 as above, then K3 once per counter level v = 0 .. c-1 over the occurrences that
//                              won nothing at the levels below (the derivation is at resolve_kernel); level 0 leaves a
//                              dense copy of every bucket for the later levels, what persists between batches is the
//                              4-bit counter of every slot, valid <=> eligible at level c-1 and a win there.
//
// All run/chunk positions are exact (prefix sums of counts): memory use is deterministic and heavy
// duplication (poly-G reads, adapters) only makes one bucket longer, never overflows anything.
#pragma once
#include "common.cuh"

namespace kwg {

constexpr int PT_THREADS = 512;                   // partition_scan_kernel block size
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
constexpr uint32_t SLOT_EMPTY = 0xFFFFFFFFu;
constexpr uint64_t MAX_COUNT_POS = 1ull << 28;    // positions per counting sub-batch (record: 28-bit position)
constexpr uint32_t REC_POS_MASK = (1u << 28) - 1u;
constexpr uint32_t REC_DOUBLE = 1u << 28;         // record flag: the occurrence touches this slot through both hashes of the table

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
	BaseSource src;              // device, 16-byte aligned (whole batch)
	const uint32_t* start_mask;
	uint32_t k;
	uint64_t pos0;               // absolute base index of the first start position of this sub-batch (multiple of 16)
	uint64_t n_pos;              // start positions in this sub-batch (<= 2^28)
	uint32_t count_mask;
	uint32_t f1_log2;
	uint32_t shift1;             // FINAL_LOG2 + f2_log2: record keeps the slot bits below this
	uint32_t table_shift;        // f1_log2 - 1: where the table index (0: first, 1: second) lands in the level-1 bucket
	uint32_t ntp;                // row pitch of offs1 (tiles, padded)
	uint32_t tile0;              // first tile of this launch (the scan may be launched in pieces while the bases stream in)
	uint64_t* rec1;              // [n_tiles][PT_REC]
	uint16_t* offs1;             // [(F1+1)][ntp]: start of bucket b inside the sorted tile; row F1 = record count
};

// exclusive prefix sum of n <= MAX_FAN counters in shared memory (in place) by a block of THREADS
// threads (256: two counters per thread, 512/1024: one); returns the total.  s_warp: THREADS/32 words of scratch.
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan_512(uint32_t* v, uint32_t n, uint32_t* s_warp)
{
	constexpr int ITEMS = (MAX_FAN + THREADS - 1) / THREADS;
	static_assert(ITEMS == 1 || ITEMS == 2, "block size must be 256, 512 or 1024");
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
	const uint64_t tile = (uint64_t)blockIdx.x + P.tile0;
	const uint64_t rel0 = tile * PT_POS;               // sub-batch relative position of the tile
	const uint64_t t0 = P.pos0 + rel0;                 // absolute base index
	const uint32_t F1 = 1u << P.f1_log2;

	for (uint32_t v = tid; v < (uint32_t)PT_VEC; v += PT_THREADS) {
		const uint64_t g = t0 + (uint64_t)v * 16;
		uint32_t codes, bad16;
		load_group16(P.src, g, codes, bad16);
		s_codes[v] = codes;
		reinterpret_cast<uint16_t*>(s_bad)[v] = (uint16_t)bad16;
	}
	for (uint32_t v = tid; v < (uint32_t)(PT_LOAD / 32 + 1); v += PT_THREADS) {
		const uint64_t w = (t0 >> 5) + v;
		s_start[v] = (w * 32 < P.src.n_bases) ? P.start_mask[w] : 0u;
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
				// both hashes of one table on the same slot (reference: the counter is read once and incremented
				// twice, make_bloom.cpp:553-554,586-592): ONE record of weight 2, carried by the even hash
				if ((j & 1) && hm == (h[j - 1] & P.count_mask)) continue;
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
			for (int j = 0; j < 4; j += 2) {
				const uint32_t hm0 = s_hm[j * PT_POS + p], hm1 = s_hm[(j + 1) * PT_POS + p];
				const uint32_t dbl = (hm0 == hm1) ? REC_DOUBLE : 0u;
				const uint32_t b0 = (hm0 >> P.shift1) | ((uint32_t)(j >> 1) << P.table_shift);
				s_sorted[atomicAdd(&s_cursor[b0], 1u)] = ((uint64_t)(hm0 & low_mask) << 32) | pos | dbl;
				if (!dbl) {
					const uint32_t b1 = (hm1 >> P.shift1) | ((uint32_t)(j >> 1) << P.table_shift);
					s_sorted[atomicAdd(&s_cursor[b1], 1u)] = ((uint64_t)(hm1 & low_mask) << 32) | pos;
				}
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

// Exclusive prefix sums over the F1*NG (bucket-major) group counts, in two small kernels of F1 blocks:
//   group_scan_kernel   block i: prefix over the groups of bucket i (records rounded up to even so that every
//                       chunk starts 16-byte aligned for the bulk stores; chunks = ceil(records / CHUNK_REC))
//                       -> local prefixes in base2/cbase, bucket totals in tot_rec/tot_chk
//   group_finish_kernel block i: adds the totals of the buckets before it, writes the record base and the
//                       description of every chunk and the first chunk of every level-1 bucket
constexpr int GS_THREADS = 256;

__global__ void __launch_bounds__(GS_THREADS)
group_scan_kernel(const uint32_t* __restrict__ cnt1, uint32_t NG, uint64_t* __restrict__ base2, uint32_t* __restrict__ cbase,
	unsigned long long* __restrict__ tot_rec, uint32_t* __restrict__ tot_chk)
{
	__shared__ uint32_t s_r[GS_THREADS / 32], s_c[GS_THREADS / 32];
	const uint32_t i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned long long rcarry = 0;
	uint32_t ccarry = 0;
	for (uint32_t g0 = 0; g0 < NG; g0 += GS_THREADS) {
		const uint32_t g = g0 + tid;
		const uint32_t c = (g < NG) ? cnt1[(uint64_t)i * NG + g] : 0u;
		const uint32_t r = (c + 1u) & ~1u, k = (c + CHUNK_REC - 1) / CHUNK_REC;
		uint32_t rinc = r, cinc = k;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t ry = __shfl_up_sync(0xFFFFFFFFu, rinc, o), cy = __shfl_up_sync(0xFFFFFFFFu, cinc, o);
			if (lane >= (uint32_t)o) { rinc += ry; cinc += cy; }
		}
		__syncthreads();
		if (lane == 31) { s_r[warp] = rinc; s_c[warp] = cinc; }
		__syncthreads();
		uint32_t rb = 0, cb = 0, rt = 0, ctot = 0;
#pragma unroll
		for (int w = 0; w < GS_THREADS / 32; ++w) {
			if ((uint32_t)w < warp) { rb += s_r[w]; cb += s_c[w]; }
			rt += s_r[w]; ctot += s_c[w];
		}
		if (g < NG) {
			base2[(uint64_t)i * NG + g] = rcarry + rb + rinc - r;
			cbase[(uint64_t)i * NG + g] = ccarry + cb + cinc - k;
		}
		rcarry += rt; ccarry += ctot;
	}
	if (tid == 0) { tot_rec[i] = rcarry; tot_chk[i] = ccarry; }
}

__global__ void __launch_bounds__(GS_THREADS)
group_finish_kernel(const uint32_t* __restrict__ cnt1, uint32_t NG, uint32_t F1, const unsigned long long* __restrict__ tot_rec,
	const uint32_t* __restrict__ tot_chk, uint64_t* __restrict__ base2, uint32_t* __restrict__ cbase, uint64_t* __restrict__ chunk_rec,
	uint4* __restrict__ chunk_meta, uint32_t* __restrict__ cfirst)
{
	__shared__ unsigned long long s_r[GS_THREADS / 32];
	__shared__ uint32_t s_c[GS_THREADS / 32];
	const uint32_t i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned long long r = 0;
	uint32_t c = 0;
	for (uint32_t x = tid; x < i; x += GS_THREADS) { r += tot_rec[x]; c += tot_chk[x]; }
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xFFFFFFFFu, r, o); c += __shfl_xor_sync(0xFFFFFFFFu, c, o); }
	if (lane == 0) { s_r[warp] = r; s_c[warp] = c; }
	__syncthreads();
	unsigned long long roff = 0;
	uint32_t coff = 0;
#pragma unroll
	for (int w = 0; w < GS_THREADS / 32; ++w) { roff += s_r[w]; coff += s_c[w]; }
	if (tid == 0) {
		cfirst[i] = coff;
		if (i + 1 == F1) cfirst[F1] = coff + tot_chk[i];
	}
	for (uint32_t g = tid; g < NG; g += GS_THREADS) {
		const uint64_t x = (uint64_t)i * NG + g;
		const uint32_t cnt = cnt1[x];
		const unsigned long long rb = roff + base2[x];
		const uint32_t cb = coff + cbase[x];
		base2[x] = rb;
		cbase[x] = cb;
		const uint32_t nc = (cnt + CHUNK_REC - 1) / CHUNK_REC;
		for (uint32_t q = 0; q < nc; ++q) {
			chunk_rec[cb + q] = rb + (unsigned long long)q * CHUNK_REC;
			chunk_meta[cb + q] = make_uint4((i << 16) | g, cnt, q, 0u);     // (bucket i, group g), records of the pair, chunk within the pair
		}
	}
}

// ------------------------------------------------------------------------------------------ async-copy plumbing
// The runs a block gathers are short (~26 records = 208 bytes) and scattered.  Loading them with LDG costs
// ~40 instructions per run and leaves every warp waiting on its own few loads; instead one thread per run
// issues one bulk asynchronous copy (cp.async.bulk, the TMA engine's linear mode) into shared memory: all
// the runs of a block are in flight at once, no registers are held, and completion is one mbarrier wait.
// Bulk copies need 16-byte aligned addresses and sizes; runs are 8-byte aligned, so a copy is widened by
// up to one record on either side and the owner thread overwrites those strangers with REC_NULL.
constexpr int STAGE_REC = CHUNK_REC + 2 * MAX_FAN;           // staging capacity in records (72 KiB)
constexpr uint32_t REC_NULL_HI = 0xFFFFFFFFu;                // slot field of a padding record (real ones use < 2^24)
constexpr unsigned long long REC_NULL = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug must trap, not hang the device.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
	const uint32_t addr = smem_u32(bar);
	for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
		uint32_t done;
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(addr), "r"(parity) : "memory");
		if (done) return;
		__nanosleep(spin < 8 ? 32 : 256);       // waiting warps must not crowd the issue slots of the working ones
	}
	__trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
	             :: "l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy accesses to shared memory before this fence are ordered before later async-proxy accesses
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// exclusive prefix sum of one value per thread over a block of 1024 threads; returns the total
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t x, uint32_t& total, uint32_t* s_warp)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = x;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
		if (lane >= (uint32_t)o) inc += y;
	}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	uint32_t w = (lane < 32) ? s_warp[lane] : 0u;          // every warp scans the 32 warp totals
	uint32_t winc = w;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, winc, o);
		if (lane >= (uint32_t)o) winc += y;
	}
	total = __shfl_sync(0xFFFFFFFFu, winc, 31);
	const uint32_t wbase = __shfl_sync(0xFFFFFFFFu, winc - w, warp);
	__syncthreads();
	return wbase + inc - x;
}

// ------------------------------------------------------------------------------------------ K2
struct RegroupParams {
	const uint64_t* rec1;
	const uint16_t* offs1;
	uint32_t ntp, n_tiles;
	uint32_t F1, G1;
	uint32_t f2_log2;
	const uint32_t* cfirst;      // [F1 + 1]; cfirst[F1] = number of chunks = work units of this kernel
	const uint64_t* chunk_rec;
	const uint4* chunk_meta;     // x: i << 16 | g, y: records of the (i, g) pair, z: chunk index within the pair
	uint64_t* rec2;
	uint16_t* offs2;             // bucket i: rows at cfirst[i]*(F2+1); entry (j, local chunk c) at + j*nci + c
};

struct RegroupUnit {             // what the consumers need to know about a staged chunk
	uint32_t count;              // real records of the chunk
	uint32_t nci, cl;            // chunks of the level-1 bucket, index of this chunk among them
	uint32_t pad;
	unsigned long long rows;     // offs2 element index of the bucket's rows
	unsigned long long out;      // rec2 record index of the chunk
};

constexpr int RG_PRODUCERS = 128;                            // producer threads: each owns 4 consecutive runs
constexpr int RG_CONSUMERS = 512;                            // consumer threads: one record per lane
constexpr int RG_BLOCK = RG_PRODUCERS + RG_CONSUMERS;
static_assert(RG_PRODUCERS * 4 == MAX_FAN, "a producer thread owns four runs");

__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 2, %0;" :: "n"(RG_PRODUCERS) : "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(RG_CONSUMERS) : "memory"); }

// exclusive prefix of one value per producer thread; *total receives the sum.  s_pw: RG_PRODUCERS/32 words.
__device__ __forceinline__ uint32_t producer_exclusive_scan(uint32_t x, uint32_t* total, uint32_t* s_pw, uint32_t pt)
{
	const uint32_t lane = pt & 31, warp = pt >> 5;
	uint32_t inc = x;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
		if (lane >= (uint32_t)o) inc += y;
	}
	producer_sync();                              // previous use of s_pw is over
	if (lane == 31) s_pw[warp] = inc;
	producer_sync();
	uint32_t base = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < RG_PRODUCERS / 32; ++w) {
		const uint32_t v = s_pw[w];
		if ((uint32_t)w < warp) base += v;
		tot += v;
	}
	*total = tot;
	return base + inc - x;
}

// exclusive prefix sum of n <= RG_CONSUMERS counters in shared memory (in place) by the consumer threads
__device__ __forceinline__ void consumer_exclusive_scan(uint32_t* v, uint32_t n, uint32_t* s_warp, uint32_t ct)
{
	const uint32_t lane = ct & 31, warp = ct >> 5;
	const uint32_t a = (ct < n) ? v[ct] : 0u;
	uint32_t x = a;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
		if (lane >= (uint32_t)o) x += y;
	}
	if (lane == 31) s_warp[warp] = x;
	consumer_sync();
	uint32_t base = 0;
#pragma unroll
	for (int w = 0; w < RG_CONSUMERS / 32; ++w)
		if ((uint32_t)w < warp) base += s_warp[w];
	if (ct < n) v[ct] = base + x - a;
	consumer_sync();
}

// Persistent, warp-specialised blocks; one work unit = one chunk (<= CHUNK_REC records of level-1 bucket i
// coming from one group of G1 <= 512 tiles).
//   producer warps: each thread owns four consecutive runs of the unit (its slice of the run table is one
//                   64-bit load per row, prefetched one unit ahead), the group prefix-sums the 16-byte aligned
//                   spans into a staging layout and every thread issues its bulk async copies; completion is
//                   tracked by the buffer's "full" mbarrier (complete_tx)
//   consumer warps: wait for "full", blank the stranger records the 16-byte widening dragged in, histogram
//                   the level-2 buckets, prefix-sum, scatter into the sorted chunk, bulk-store it, and hand
//                   the staging buffer back through the "empty" mbarrier
// Two staging buffers: the runs of the next unit travel while this one is sorted.
__global__ void __launch_bounds__(RG_BLOCK, 1)
regroup_kernel(const RegroupParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint64_t* s_stage = reinterpret_cast<uint64_t*>(smem_raw);                          // [2][STAGE_REC]
	uint64_t* s_sorted = s_stage + 2 * STAGE_REC;                                       // CHUNK_REC (+2 spare)
	unsigned long long* s_full = reinterpret_cast<unsigned long long*>(s_sorted + CHUNK_REC + 2);  // [2]
	unsigned long long* s_empty = s_full + 2;                                           // [2]
	RegroupUnit* s_unit = reinterpret_cast<RegroupUnit*>(s_empty + 2);                  // [2]
	uint32_t* s_own = reinterpret_cast<uint32_t*>(s_unit + 2);                          // [2][MAX_FAN]: so | n << 14 | lead << 28
	uint32_t* s_staged = s_own + 2 * MAX_FAN;                                           // [2] records staged (incl. padding)
	uint32_t* s_hist = s_staged + 2;                                                    // MAX_FAN + 1
	uint32_t* s_cursor = s_hist + MAX_FAN + 1;                                          // MAX_FAN
	uint32_t* s_warp = s_cursor + MAX_FAN;                                              // 16 (consumer scans)
	uint32_t* s_pw = s_warp + 16;                                                       // 4 (producer scans)

	const uint32_t tid = threadIdx.x, lane = tid & 31;
	const uint32_t F2 = 1u << P.f2_log2;
	const uint32_t n_units = P.cfirst[P.F1];

	if (tid == 0) {
		mbar_init(&s_full[0], RG_PRODUCERS); mbar_init(&s_full[1], RG_PRODUCERS);
		mbar_init(&s_empty[0], RG_CONSUMERS / 32); mbar_init(&s_empty[1], RG_CONSUMERS / 32);
		mbar_init_fence();
	}
	__syncthreads();

	if (tid < (uint32_t)RG_PRODUCERS) {
		// ================================================================= producer group
		const uint32_t pt = tid;
		// run table slice of a unit: runs 4pt .. 4pt+3, starts and ends as 4 x u16 each
		uint2 ts = make_uint2(0, 0), te = make_uint2(0, 0);
		uint4 meta = make_uint4(0, 0, 0, 0);
		auto load_table = [&](uint32_t u) {
			meta = P.chunk_meta[u];
			const uint32_t i = meta.x >> 16, t0 = (meta.x & 0xFFFFu) * P.G1;
			const uint32_t nt = min(P.G1, P.n_tiles - t0);
			ts = make_uint2(0, 0); te = make_uint2(0, 0);
			if (4 * pt < nt) {             // rows are 8-byte aligned (ntp % 64 == 0, t0 % 4 == 0); slack past nt is never used
				const uint16_t* r0 = P.offs1 + (uint64_t)i * P.ntp + t0 + 4 * pt;
				ts = *reinterpret_cast<const uint2*>(r0);
				te = *reinterpret_cast<const uint2*>(r0 + P.ntp);
			}
		};
		uint32_t u = blockIdx.x;
		if (u < n_units) load_table(u);
		for (uint32_t it = 0; u < n_units; u += gridDim.x, ++it) {
			const uint32_t buf = it & 1u;
			const uint32_t i = meta.x >> 16, t0 = (meta.x & 0xFFFFu) * P.G1;
			const uint32_t nt = min(P.G1, P.n_tiles - t0);
			const uint32_t cnt = meta.y, clocal = meta.z;
			const uint32_t cv0 = clocal * CHUNK_REC, cv1 = min(cnt, cv0 + CHUNK_REC);
			uint32_t s0[4], n[4];
			s0[0] = ts.x & 0xFFFFu; s0[1] = ts.x >> 16; s0[2] = ts.y & 0xFFFFu; s0[3] = ts.y >> 16;
			n[0] = (te.x & 0xFFFFu) - s0[0]; n[1] = (te.x >> 16) - s0[1]; n[2] = (te.y & 0xFFFFu) - s0[2]; n[3] = (te.y >> 16) - s0[3];
#pragma unroll
			for (int q = 0; q < 4; ++q) if (4 * pt + q >= nt) { s0[q] = 0; n[q] = 0; }
			// unit header (one thread; its loads overlap the layout computation)
			unsigned long long out = 0;
			uint32_t c0 = 0, c1 = 0;
			if (pt == 0) { out = P.chunk_rec[u]; c0 = P.cfirst[i]; c1 = P.cfirst[i + 1]; }
			if (u + gridDim.x < n_units) load_table(u + gridDim.x);     // next unit's table travels meanwhile

			if (cnt > CHUNK_REC) {
				// a skewed group spans several chunks: clip the runs to the virtual range [cv0, cv1) of this one
				uint32_t tot;
				uint32_t vs = producer_exclusive_scan(n[0] + n[1] + n[2] + n[3], &tot, s_pw, pt);
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					const uint32_t lo = max(vs, cv0), hi = min(vs + n[q], cv1);
					vs += n[q];
					s0[q] += lo - (vs - n[q]);
					n[q] = (lo < hi) ? hi - lo : 0u;
				}
			}
			uint32_t span[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) span[q] = n[q] ? (((s0[q] & 1u) + n[q] + 1u) & ~1u) : 0u;
			const uint32_t mine = span[0] + span[1] + span[2] + span[3];
			uint32_t staged;
			uint32_t so = producer_exclusive_scan(mine, &staged, s_pw, pt);
			// the consumers must have left this buffer (they used it two units ago)
			if (it >= 2) mbar_wait(&s_empty[buf], ((it >> 1) - 1u) & 1u);
			uint4 own;
			uint32_t sos[4];
			sos[0] = so; sos[1] = sos[0] + span[0]; sos[2] = sos[1] + span[1]; sos[3] = sos[2] + span[2];
			own.x = sos[0] | (n[0] << 14) | ((s0[0] & 1u) << 28);
			own.y = sos[1] | (n[1] << 14) | ((s0[1] & 1u) << 28);
			own.z = sos[2] | (n[2] << 14) | ((s0[2] & 1u) << 28);
			own.w = sos[3] | (n[3] << 14) | ((s0[3] & 1u) << 28);
			reinterpret_cast<uint4*>(s_own + buf * MAX_FAN)[pt] = own;
			if (pt == 0) {
				RegroupUnit U;
				U.count = cv1 - cv0;
				U.nci = c1 - c0; U.cl = u - c0; U.pad = 0;
				U.rows = (unsigned long long)c0 * (F2 + 1);
				U.out = out;
				s_unit[buf] = U;
				s_staged[buf] = staged;
			}
			// arrive (releases the stores above) and announce this thread's bytes, then start its copies
			if (mine) mbar_arrive_expect_tx(&s_full[buf], mine * 8u); else mbar_arrive(&s_full[buf]);
			uint64_t* stage = s_stage + (size_t)buf * STAGE_REC;
			const uint64_t* src = P.rec1 + (uint64_t)(t0 + 4 * pt) * PT_REC;
#pragma unroll
			for (int q = 0; q < 4; ++q)
				if (n[q]) bulk_g2s(stage + sos[q], src + (uint64_t)q * PT_REC + (s0[q] - (s0[q] & 1u)), span[q] * 8u, &s_full[buf]);
		}
	} else {
		// ================================================================= consumer warps
		const uint32_t ct = tid - RG_PRODUCERS;
		uint32_t it = 0;
		for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
			const uint32_t buf = it & 1u;
			uint64_t* stage = s_stage + (size_t)buf * STAGE_REC;
			mbar_wait(&s_full[buf], (it >> 1) & 1u);
			{
				const uint32_t own = s_own[buf * MAX_FAN + ct];
				const uint32_t so = own & 0x3FFFu, n = (own >> 14) & 0x3FFFu, lead = own >> 28;
				if (n) {
					if (lead) stage[so] = REC_NULL;
					if ((lead + n) & 1u) stage[so + lead + n] = REC_NULL;
				}
			}
			for (uint32_t v = ct; v <= F2; v += RG_CONSUMERS) s_hist[v] = 0;
			consumer_sync();
			const RegroupUnit U = s_unit[buf];
			const uint32_t staged = s_staged[buf];

			// pass A: level-2 histogram, one record per lane
			for (uint32_t e = ct; e < staged; e += RG_CONSUMERS) {
				const uint32_t hi = (uint32_t)(stage[e] >> 32);
				if (hi != REC_NULL_HI) atomicAdd(&s_hist[hi >> FINAL_LOG2], 1u);
			}
			consumer_sync();
			consumer_exclusive_scan(s_hist, F2, s_warp, ct);
			uint16_t* rows = P.offs2 + U.rows;
			for (uint32_t j = ct; j < F2; j += RG_CONSUMERS) {
				const uint32_t s = s_hist[j];
				s_cursor[j] = s;
				rows[(uint64_t)j * U.nci + U.cl] = (uint16_t)s;
			}
			if (ct == 0) {
				rows[(uint64_t)F2 * U.nci + U.cl] = (uint16_t)U.count;
				bulk_wait_read();             // the previous chunk has left s_sorted
			}
			consumer_sync();
			// pass B: counting-sort scatter
			for (uint32_t e = ct; e < staged; e += RG_CONSUMERS) {
				const uint64_t rec = stage[e];
				const uint32_t hi = (uint32_t)(rec >> 32);
				if (hi != REC_NULL_HI) s_sorted[atomicAdd(&s_cursor[hi >> FINAL_LOG2], 1u)] = rec;
			}
			fence_async_smem();               // staging reads/stores and s_sorted stores before the async proxy touches either
			consumer_sync();
			if (lane == 0) mbar_arrive(&s_empty[buf]);
			if (ct == 0) bulk_s2g(P.rec2 + U.out, s_sorted, ((U.count + 1u) & ~1u) * 8u);
		}
		if (ct == 0) bulk_wait_all();
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
	uint32_t* loss;              // 4-bit counters, 8 positions per word: losses (resolve_kernel<false>) or wins (<true>)
	// min_kmer_count > 1 (resolve_kernel<true>): one launch per counter level
	uint16_t* cnt;               // persistent 4-bit counters of the two counting filters: bucket-major, slot s of a bucket = nibble s
	const uint32_t* elig;        // bit per position: occurrence read counters >= level on all four slots (NULL: level 0)
	uint32_t level;              // this launch decides which slots go from `level` to level + 1 (+ 2 for a weight-2 winner)
	uint32_t* wrap_flag;         // set when a counter would pass 15 (the reference wraps to 0, bloom.h 4-bit field)
	// dense copies of the buckets for the launches of level >= 1 (resolve_dense_kernel): level 0 stores the staged
	// records of every bucket that fits one staging window at dense + b * STAGE_REC and notes the extent
	uint64_t* dense;             // NULL: not used
	uint32_t* dense_len;         // [n_buckets] records (even), 0: the bucket has to be gathered again
	uint32_t* n_not_dense;       // buckets with records that have no dense copy
	uint32_t* nd_list;           // ... and which they are
};

constexpr uint32_t RS_LONG = 64;                             // runs longer than this are read directly, not staged
constexpr uint32_t RS_ROUND = STAGE_REC - (RS_LONG + 2);     // staging window per round (one run may overhang)

__device__ __forceinline__ uint64_t ld_nc_u64(const uint64_t* p)
{
	uint64_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
	return r;
}

// One touch record against the tile.  Tile entry = ((position + 1) << 1) | (single ? 1 : 0) of the earliest
// (eligible) occurrence seen so far, 0 = the slot already holds a larger count than this launch looks at,
// SLOT_EMPTY = untouched.  A weight-2 record stands for both hashes of one table on the same slot.
//
// LEVELS == false (min_kmer_count == 1): every occurrence takes part, and `acct` counts LOSSES per occurrence
//     (touches that found the slot taken; valid <=> fewer than 4): on low-coverage input almost every touch wins,
//     so counting the losers keeps the global atomics rare.  First touches are noted in the bitmap.
// LEVELS == true : `acct` counts WINS per occurrence (eligible at the next level <=> no win; valid <=> a win at the
//     last level): min_kmer_count > 1 is what one uses on high-coverage input, where almost every touch loses.  A
//     record that cannot lower the tile entry has no effect at all, so it is dropped after ONE shared-memory read
//     (a stale value is safe: entries only decrease) and only the few would-be winners look up their eligibility.
// Returns true iff the record became the holder of its slot (LEVELS only): only such records can be winners in the end,
// which is what lets the settle pass look at a handful of records per thread instead of all of them.
template <bool LEVELS>
__device__ __forceinline__ bool resolve_record(uint32_t* s_tile, uint32_t* s_bm, uint32_t* acct, const uint32_t* elig, uint32_t level, uint64_t rec,
	bool have_word = false, uint32_t word = 0u)        // have_word: the caller already fetched elig[pos >> 5]
{
	const uint32_t slot = (uint32_t)(rec >> 32) & (FINAL_SLOTS - 1);
	const uint32_t lo = (uint32_t)rec;
	const uint32_t pos = lo & REC_POS_MASK;
	const uint32_t dbl = (lo >> 28) & 1u;
	const uint32_t v = ((pos + 1u) << 1) | (dbl ^ 1u);
	if (LEVELS) {
		// s_bm holds the bucket's 4-bit counters here (8 slots per word)
		if (*reinterpret_cast<volatile uint32_t*>(&s_tile[slot]) < v) return false;
		if (((s_bm[slot >> 3] >> ((slot & 7u) << 2)) & 15u) > level) return false;      // the slot is already above this level
		if (have_word) { if (!((word >> (pos & 31u)) & 1u)) return false; }
		else if (elig && !((__ldg(&elig[pos >> 5]) >> (pos & 31u)) & 1u)) return false;
		const uint32_t old = atomicMin(&s_tile[slot], v);
		if (old > v) {
			atomicAdd(&acct[pos >> 3], (1u + dbl) << ((pos & 7u) << 2));
			if (old != SLOT_EMPTY) {
				// a later occurrence had been processed first and has just been displaced: take its win back
				// (the two updates of a nibble may land in either order; the sum of the word is what counts)
				const uint32_t q = (old >> 1) - 1u;
				atomicSub(&acct[q >> 3], (2u - (old & 1u)) << ((q & 7u) << 2));
			}
			return true;
		}
		return false;
	}
	const uint32_t old = atomicMin(&s_tile[slot], v);
	if (old <= v) {
		// an earlier occurrence (or an earlier batch) holds the slot: this touch read a non-zero counter
		atomicAdd(&acct[pos >> 3], (1u + dbl) << ((pos & 7u) << 2));
	} else if (old != SLOT_EMPTY) {
		// a later occurrence had been processed first and has just been displaced
		const uint32_t q = (old >> 1) - 1u;
		atomicAdd(&acct[q >> 3], (2u - (old & 1u)) << ((q & 7u) << 2));
	} else {
		atomicOr(&s_bm[slot >> 5], 1u << (slot & 31u));      // first touch of the slot in this accession
	}
	return false;
}

// LEVELS, after all records of the bucket went through resolve_record: is this record the winner of its slot?  Then
// the slot's counter goes from level to level + weight and the tile entry is handed back clean.
__device__ __forceinline__ bool settle_winner(uint32_t* s_tile, uint32_t* s_cn, uint32_t level, uint32_t* wrap_flag, uint64_t rec)
{
	const uint32_t slot = (uint32_t)(rec >> 32) & (FINAL_SLOTS - 1);
	const uint32_t lo = (uint32_t)rec;
	const uint32_t v = (((lo & REC_POS_MASK) + 1u) << 1) | (((lo >> 28) & 1u) ^ 1u);
	if (s_tile[slot] != v) return false;
	const uint32_t w = 2u - (v & 1u);
	if (level + w > 15u) *wrap_flag = 1u;           // the reference's 4-bit field would wrap to 0: reported, not mimicked
	else atomicAdd(&s_cn[slot >> 3], w << ((slot & 7u) << 2));
	s_tile[slot] = SLOT_EMPTY;
	return true;
}

// what a thread needs of a bucket before it can start: its word of the touched bitmap and run tid
struct BucketPrefetch { uint32_t bmw, len, off; };    // off: record index (< 2^30)

template <bool LEVELS>
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
	r.bmw = (!LEVELS && P.have_prior) ? P.touched[(uint64_t)b * (FINAL_SLOTS / 32) + tid] : 0u;
	r.len = 0; r.off = 0;
	if (tid < nci) {
		const uint32_t s = row_s[tid], e = row_s[pitch + tid];
		const uint64_t base = P.chunk_stride ? (uint64_t)(c0 + tid) * P.chunk_stride : P.chunk_rec[c0 + tid];
		r.off = (uint32_t)(base + s);
		r.len = e - s;
	}
	return r;
}

// One final bucket at a time per (persistent) block: 2^15 slots as a 128 KiB tile of u32 "smallest
// position that touched the slot".  The bucket's runs (one per chunk, ~26 records) are gathered into the
// staging buffer by bulk async copies while the tile is initialised, then resolved one record per lane.
//
// LEVELS (min_kmer_count > 1): the conservative-update counters of the reference (make_bloom.cpp:546-601) are
// resolved level by level.  Let T_v(s) be the stream position whose increment takes slot s to a count >= v.  An
// occurrence t reads min >= v iff t > T_v(s) for all its slots ("eligible at level v"), and
//     T_{v+1}(s) = min { t touching s : t eligible at level v },
// because the earliest such t still reads exactly v on s, v as its minimum, and therefore increments s.  One launch
// per level v = 0 .. c-1: the tile starts at 0 where the persistent counter is already above v (earlier batch, or
// a weight-2 winner of level v-1) and at SLOT_EMPTY elsewhere, eligible touches go through atomicMin, the losers'
// wins are summed per occurrence (eligible at v+1 <=> it won nothing), winners raise the counter.  An occurrence is
// valid (reads min == c-1) iff it is eligible at level c-1 and wins at least one slot there.
template <bool LEVELS>
__global__ void __launch_bounds__(RS_THREADS, 1)
resolve_kernel(const ResolveParams P)
{
	// LEVELS: the bitmap region holds the bucket's 4-bit counters instead (16 KiB)
	constexpr uint32_t BM_WORDS = LEVELS ? FINAL_SLOTS / 8 : FINAL_SLOTS / 32;
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem_raw);                           // FINAL_SLOTS
	uint64_t* s_stage = reinterpret_cast<uint64_t*>(s_tile + FINAL_SLOTS);              // STAGE_REC
	unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_stage + STAGE_REC);        // 1 (+1 pad)
	uint32_t* s_bm = reinterpret_cast<uint32_t*>(s_bar + 2);                            // BM_WORDS (16-byte aligned)
	uint32_t* s_off = s_bm + BM_WORDS;                                                  // RS_THREADS: record index of every run
	uint32_t* s_len = s_off + RS_THREADS;                                               // RS_THREADS
	uint32_t* s_warp = s_len + RS_THREADS;                                              // 32
	uint32_t* s_misc = s_warp + 32;                                                     // [0] long runs, [1] staged extent of the round
	uint16_t* s_long = reinterpret_cast<uint16_t*>(s_misc + 2);                         // RS_THREADS: indices of the long runs

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t F2 = 1u << P.f2_log2;

	if (tid == 0) { mbar_init(&s_bar[0], RS_THREADS); mbar_init_fence(); }
	if (LEVELS) {
		// the tile is handed back clean by every bucket (settle_winner / the sweep below): filled once
		const uint4 e4 = make_uint4(SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY);
#pragma unroll
		for (uint32_t q = 0; q < (uint32_t)(FINAL_SLOTS / (4 * RS_THREADS)); ++q)
			reinterpret_cast<uint4*>(s_tile)[q * RS_THREADS + tid] = e4;
	}
	__syncthreads();
	uint32_t parity = 0;

	// work list: all buckets, or (levels >= 1 with dense copies) the few buckets that have none
	const bool listed = LEVELS && P.level > 0 && P.dense;
	const uint32_t n_work = listed ? *P.n_not_dense : P.n_buckets;
	uint32_t it = blockIdx.x;
	uint32_t b = 0;
	BucketPrefetch pf{0u, 0u, 0u};
	if (it < n_work) { b = listed ? P.nd_list[it] : it; pf = prefetch_bucket<LEVELS>(P, b); }

	for (; it < n_work; it += gridDim.x) {
		const uint32_t i = b >> P.f2_log2, j = b & (F2 - 1);
		const uint32_t c0 = P.cfirst ? P.cfirst[i] : 0u;
		const uint32_t nci = P.cfirst ? P.cfirst[i + 1] - c0 : P.single_nci;
		const uint64_t pitch = P.row_pitch ? P.row_pitch : (uint64_t)nci;
		const uint16_t* row_s = P.offs + (uint64_t)c0 * (F2 + 1) + (uint64_t)j * pitch;
		const uint32_t b_cur = b;

		uint32_t my_off = pf.off, my_len = pf.len;
		if (!LEVELS) s_bm[tid] = pf.bmw;
		// tables of the next bucket travel while this one is resolved, and its runs are pulled into L2
		if (it + gridDim.x < n_work) {
			b = listed ? P.nd_list[it + gridDim.x] : it + gridDim.x;
			pf = prefetch_bucket<LEVELS>(P, b);
			if (pf.len) {
				const char* p0 = reinterpret_cast<const char*>(P.rec + pf.off);
				const char* p1 = p0 + (size_t)min(pf.len, RS_LONG) * 8 - 1;
				for (const char* q = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)127); q <= p1; q += 128)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
			}
		}

		bool tile_ready = false;
		bool stage_whole = false, has_records = false;       // LEVELS: every record of the bucket is in the staging buffer
		uint32_t held = 0;        // LEVELS: bit q <=> this thread's q-th record of the (last) round became its slot's holder
		uint32_t whole_extent = 0;
		for (uint32_t cb = 0; cb < nci; cb += RS_THREADS) {
			const uint32_t nrt = min((uint32_t)RS_THREADS, nci - cb);
			if (cb) {          // run tables beyond the prefetched first RS_THREADS runs
				my_off = 0; my_len = 0;
				if (tid < nrt) {
					const uint32_t c = cb + tid;
					const uint32_t s = row_s[c], e = row_s[pitch + c];
					const uint64_t base = P.chunk_stride ? (uint64_t)(c0 + c) * P.chunk_stride : P.chunk_rec[c0 + c];
					my_off = (uint32_t)(base + s);
					my_len = e - s;
				}
			}
			if (tid == 0) s_misc[0] = 0;
			// staging layout of the short runs: prefix sum of their 16-byte aligned spans
			const bool is_long = my_len > RS_LONG;
			const uint32_t lead = my_off & 1u;
			const uint32_t span = (my_len && !is_long) ? ((lead + my_len + 1u) & ~1u) : 0u;
			uint32_t staged;
			const uint32_t so = block_exclusive_scan_1024(span, staged, s_warp);      // (two block barriers inside)
			if (is_long) {
				const uint32_t q = atomicAdd(&s_misc[0], 1u);
				s_long[q] = (uint16_t)tid; s_off[q] = my_off; s_len[q] = my_len;
			}

			for (uint32_t r0 = 0; r0 < staged || r0 == 0; r0 += RS_ROUND) {
				if (r0) {
					// A run that overhung the previous window may have carried the staged records to their end: then no
					// run starts in this window, nobody below sets the extent, and the PREVIOUS window would be resolved a
					// second time (every record of the bucket one loss too many).  Found by the randomised soak
					// (tests/soak/stress_levels.py); regression: tests/test_gpu_bloom.py::test_staged_records_end_in_an_overhang.
					if (tid == 0) s_misc[1] = 0;
					__syncthreads();
				}
				const bool mine = span && so >= r0 && so < r0 + RS_ROUND;
				if (mine && (so + span >= r0 + RS_ROUND || so + span == staged)) s_misc[1] = so + span - r0;
				if (staged == 0 && tid == 0) s_misc[1] = 0;
				// the last run of the previous round may overhang the window: the first run of this one then
				// starts a few records in, and what lies before it must not be resolved a second time
				if (r0 && tid < RS_LONG + 2) s_stage[tid] = REC_NULL;
				if (LEVELS && tid == 0) bulk_wait_read();     // bulk stores of the previous bucket have left shared memory
				fence_async_smem();
				__syncthreads();
				// LEVELS: the bucket's counters travel with the first round
				const uint32_t cn_bytes = (LEVELS && !tile_ready && tid == 0) ? (uint32_t)(BM_WORDS * 4) : 0u;
				if (mine || cn_bytes) {
					mbar_arrive_expect_tx(&s_bar[0], (mine ? span * 8u : 0u) + cn_bytes);
					if (mine) bulk_g2s(s_stage + (so - r0), P.rec + (my_off - lead), span * 8u, &s_bar[0]);
					if (cn_bytes) bulk_g2s(s_bm, P.cnt + (uint64_t)b_cur * (FINAL_SLOTS / 4), cn_bytes, &s_bar[0]);
				} else {
					mbar_arrive(&s_bar[0]);
				}
				if (!LEVELS && !tile_ready) {
					// tile <- touched bitmap of this bucket, while the copies are in flight
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
				}
				tile_ready = true;
				mbar_wait(&s_bar[0], parity);
				parity ^= 1u;
				if (mine) {
					if (lead) s_stage[so - r0] = REC_NULL;
					if ((lead + my_len) & 1u) s_stage[so - r0 + lead + my_len] = REC_NULL;
				}
				__syncthreads();
				const uint32_t extent = s_misc[1];
				held = 0;
				uint32_t q = 0;
				for (uint32_t e = tid; e < extent; e += RS_THREADS, ++q) {
					const uint64_t rec = s_stage[e];
					if ((uint32_t)(rec >> 32) != REC_NULL_HI) {
						if (resolve_record<LEVELS>(s_tile, s_bm, P.loss, P.elig, P.level, rec)) held |= 1u << q;
					}
				}
				__syncthreads();
			}
			// long runs (heavy duplication): one warp per run straight from global memory
			const uint32_t nlong = s_misc[0];
			for (uint32_t q = warp; q < nlong; q += RS_THREADS / 32) {
				const uint64_t* run = P.rec + s_off[q];
				const uint32_t len = s_len[q];
				for (uint32_t x = lane; x < len; x += 32) resolve_record<LEVELS>(s_tile, s_bm, P.loss, P.elig, P.level, ld_nc_u64(run + x));
			}
			__syncthreads();
			if (LEVELS) {
				has_records = has_records || staged != 0u || nlong != 0u;
				stage_whole = nci <= (uint32_t)RS_THREADS && staged <= RS_ROUND && nlong == 0u && staged != 0u;
				whole_extent = s_misc[1];
			}
		}

		if (LEVELS) {
			if (tile_ready) {
				if (stage_whole) {
					// the usual case: the winners are found by one more pass over the staged records; they are
					// blanked (a winner is not eligible at any higher level) before the dense copy leaves
					// (only a record that held its slot at some point can be the winner: a few per thread, not all)
					for (uint32_t m = held; m; m &= m - 1u) {
						const uint32_t e = tid + (uint32_t)(__ffs(m) - 1) * RS_THREADS;
						if (settle_winner(s_tile, s_bm, P.level, P.wrap_flag, s_stage[e])) s_stage[e] = REC_NULL;
					}
				} else if (has_records) {
					// several rounds / long runs: sweep the whole tile
#pragma unroll 1
					for (uint32_t q = 0; q < (uint32_t)(FINAL_SLOTS / RS_THREADS); ++q) {
						const uint32_t slot = q * RS_THREADS + tid;
						const uint32_t t = s_tile[slot];
						if (t != SLOT_EMPTY) {
							const uint32_t w = 2u - (t & 1u);
							if (P.level + w > 15u) *P.wrap_flag = 1u;
							else atomicAdd(&s_bm[slot >> 3], w << ((slot & 7u) << 2));
							s_tile[slot] = SLOT_EMPTY;
						}
					}
				}
				fence_async_smem();
				__syncthreads();
				if (tid == 0) {
					if (has_records) bulk_s2g(P.cnt + (uint64_t)b_cur * (FINAL_SLOTS / 4), s_bm, BM_WORDS * 4u);
					if (P.level == 0 && P.dense && stage_whole) bulk_s2g(P.dense + (uint64_t)b_cur * STAGE_REC, s_stage, whole_extent * 8u);
				}
			}
			if (P.level == 0 && P.dense && tid == 0) {
				const bool dense_ok = tile_ready && stage_whole;
				P.dense_len[b_cur] = dense_ok ? whole_extent : 0u;
				if (!dense_ok && has_records) P.nd_list[atomicAdd(P.n_not_dense, 1u)] = b_cur;
			}
		} else {
			// the bitmap now holds the earlier batches' bits plus every slot first touched here
			P.touched[(uint64_t)b_cur * (FINAL_SLOTS / 32) + tid] = s_bm[tid];
			__syncthreads();
		}
	}
	if (LEVELS && tid == 0) bulk_wait_all();
}

// Levels >= 1 of min_kmer_count > 1 for the buckets that have a dense copy (the usual case: every bucket whose
// records fit one staging window): two bulk copies per bucket (records, counters) instead of ~256 small gathers, the
// next bucket's copy is pulled into L2 meanwhile, nothing sweeps the whole tile, and the winners of this level are
// blanked in the dense copy so that later levels drop them without an eligibility look-up.
__global__ void __launch_bounds__(RS_THREADS, 1)
resolve_dense_kernel(const ResolveParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem_raw);                           // FINAL_SLOTS
	uint64_t* s_stage = reinterpret_cast<uint64_t*>(s_tile + FINAL_SLOTS);              // STAGE_REC
	unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_stage + STAGE_REC);
	uint32_t* s_cn = reinterpret_cast<uint32_t*>(s_bar + 2);                            // FINAL_SLOTS / 8

	const uint32_t tid = threadIdx.x;
	if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_init_fence(); }
	{
		const uint4 e4 = make_uint4(SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY);
#pragma unroll
		for (uint32_t q = 0; q < (uint32_t)(FINAL_SLOTS / (4 * RS_THREADS)); ++q)
			reinterpret_cast<uint4*>(s_tile)[q * RS_THREADS + tid] = e4;
	}
	__syncthreads();
	uint32_t parity = 0;

	uint32_t b = blockIdx.x;
	uint32_t dlen_n = (b < P.n_buckets) ? P.dense_len[b] : 0u;
	for (; b < P.n_buckets; b += gridDim.x) {
		const uint32_t dlen = dlen_n;
		uint64_t* seg = P.dense + (uint64_t)b * STAGE_REC;
		uint16_t* cn = P.cnt + (uint64_t)b * (FINAL_SLOTS / 4);
		if (dlen && tid == 0) {
			// (the previous bucket's accesses to shared memory ended at its last barrier)
			bulk_wait_read();
			fence_async_smem();
			// the first record of every thread arrives on its own barrier: the eligibility words of those records (the
			// tile is empty, so every one of them is a would-be winner) are fetched while the rest is still on its way
			const uint32_t n0 = min(dlen, (uint32_t)RS_THREADS);
			mbar_arrive_expect_tx(&s_bar[0], n0 * 8u);
			bulk_g2s(s_stage, seg, n0 * 8u, &s_bar[0]);
			mbar_arrive_expect_tx(&s_bar[1], (dlen - n0) * 8u + (uint32_t)(FINAL_SLOTS / 2));
			if (dlen > n0) bulk_g2s(s_stage + n0, seg + n0, (dlen - n0) * 8u, &s_bar[1]);
			bulk_g2s(s_cn, cn, (uint32_t)(FINAL_SLOTS / 2), &s_bar[1]);
		}
		if (b + gridDim.x < P.n_buckets) {
			const uint32_t bn = b + gridDim.x;
			dlen_n = P.dense_len[bn];
			if (tid * 16u < dlen_n) asm volatile("prefetch.global.L2 [%0];" :: "l"(P.dense + (uint64_t)bn * STAGE_REC + tid * 16u));
			if (dlen_n && tid < (uint32_t)(FINAL_SLOTS / 2 / 128))
				asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(P.cnt + (uint64_t)bn * (FINAL_SLOTS / 4)) + tid * 128u));
		}
		if (dlen == 0u) continue;          // block-uniform: gathered by resolve_kernel<true> (or empty)

		mbar_wait(&s_bar[0], parity);
		const uint64_t rec0 = (tid < dlen) ? s_stage[tid] : REC_NULL;
		uint32_t word0 = 0;
		if ((uint32_t)(rec0 >> 32) != REC_NULL_HI) word0 = __ldg(&P.elig[((uint32_t)rec0 & REC_POS_MASK) >> 5]);
		mbar_wait(&s_bar[1], parity);
		parity ^= 1u;
		uint32_t held = 0, q = 1;       // bit q <=> this thread's q-th record became its slot's holder (dlen <= 9 * RS_THREADS)
		if ((uint32_t)(rec0 >> 32) != REC_NULL_HI && resolve_record<true>(s_tile, s_cn, P.loss, P.elig, P.level, rec0, true, word0)) held |= 1u;
		for (uint32_t e = tid + RS_THREADS; e < dlen; e += RS_THREADS, ++q) {
			const uint64_t rec = s_stage[e];
			if ((uint32_t)(rec >> 32) != REC_NULL_HI && resolve_record<true>(s_tile, s_cn, P.loss, P.elig, P.level, rec)) held |= 1u << q;
		}
		__syncthreads();
		// settle: only a record that held its slot at some point can be the winner -- a few per thread instead of a
		// second pass over every record (that pass was 3.4k of the 10.2k cycles per bucket)
		bool any = false;
		for (uint32_t m = held; m; m &= m - 1u) {
			const uint32_t e = tid + (uint32_t)(__ffs(m) - 1) * RS_THREADS;
			if (settle_winner(s_tile, s_cn, P.level, P.wrap_flag, s_stage[e])) { seg[e] = REC_NULL; any = true; }
		}
		fence_async_smem();
		if (__syncthreads_or(any) && tid == 0) bulk_s2g(cn, s_cn, (uint32_t)(FINAL_SLOTS / 2));
	}
	if (tid == 0) bulk_wait_all();
}

// Between two levels: eligible at the next level <=> eligible at this one and no slot won (win nibble == 0).
// One thread per 32 positions; the win counters are cleared for the next launch.
__global__ void __launch_bounds__(256)
elig_update_kernel(uint32_t* __restrict__ elig, uint32_t* __restrict__ wins, uint64_t n_words, uint32_t first)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_words) return;
	const uint4 l = reinterpret_cast<const uint4*>(wins)[i];
	const uint32_t w[4] = {l.x, l.y, l.z, l.w};
	uint32_t bits = 0;
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		uint32_t m = w[q] | (w[q] >> 1);
		m = (m | (m >> 2)) & 0x11111111u;                // bit 4j <=> nibble j != 0
		m = (m | (m >> 3)) & 0x03030303u;                // 2 bits per byte
		m = (m | (m >> 6)) & 0x000F000Fu;                // 4 bits per half
		m = (m | (m >> 12)) & 0xFFu;
		bits |= (m ^ 0xFFu) << (8 * q);
	}
	elig[i] = first ? bits : (elig[i] & bits);
	reinterpret_cast<uint4*>(wins)[i] = make_uint4(0u, 0u, 0u, 0u);
}

} // namespace kwg
