// CRC-32 (ISO-HDLC, the zlib crc32 the reference calls: bloom.cpp:328-336 for the filter bits of a .bloom file,
// build_db.cpp:145,281-282,307 for the source filters and the slice region of a .db file) on sm_100a.
//
// A CRC is linear over GF(2), so it parallelises: the message is cut into 16 KiB tiles counted from its END (a tile that
// sticks out before byte 0 is zero-prefixed, which leaves a zero register untouched), every thread of a tile takes the
// raw register (no initial value, no final inversion) of 64 contiguous bytes with four 256-entry tables in shared
// memory, and registers are merged pairwise with
//        raw(A || B) = x^(8|B|) * raw(A)  xor  raw(B)        (mod the CRC polynomial)
// where the multiplication by x^(8 * 64 * 2^l) is a 32x32 bit matrix prepared on the host by repeated squaring
// (the construction zlib's crc32_combine uses).  The caller's running value enters as ~crc xor-ed into the first
// four message bytes.  A message may be a 2-D region (rows of row_bytes at row_pitch: the padded slice rows of the
// transposed slab) and many equal-length messages are done in one launch (the 2048 source filters of a build_db chunk).
//
// Only lengths that are multiples of 4 bytes at 4-byte aligned addresses are taken; the host layer keeps zlib for the rest.
#include "common.cuh"
#include "crc_tables.h"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace kwg {

constexpr int CRC_THREADS = 256;
constexpr int CRC_WPT = 16;                                   // 32-bit words per thread
constexpr int CRC_TILE_WORDS = CRC_THREADS * CRC_WPT;         // 16 KiB
constexpr int CRC_TILE_LOG2 = 8;                              // tile = 64 bytes * 2^8
constexpr int CRC_FAN = 1024;                                 // registers merged per block of the combine kernel

struct CrcParams {
	const uint8_t* base;
	uint64_t n_seg, seg_stride;       // message s starts at base + s * seg_stride
	uint64_t n_rows, row_words, row_pitch;
	uint64_t n_tiles;                 // per message
	const uint32_t* crc_in;           // [n_seg] running values (NULL: 0)
	uint32_t* tile_crc;               // [n_seg][n_tiles], tile 0 = the END of the message
	const uint32_t* tables;           // [4][256] slice-by-4 tables, then [CRC_LEVELS][32] shift matrices
};

__device__ __forceinline__ uint32_t gf2_times(const uint32_t* mat, uint32_t vec)
{
	uint32_t r = 0;
#pragma unroll
	for (int i = 0; i < 32; ++i) r ^= mat[i] & (0u - ((vec >> i) & 1u));
	return r;
}

__global__ void __launch_bounds__(CRC_THREADS)
crc_tile_kernel(const CrcParams P)
{
	__shared__ uint32_t s_tab[4 * 256];
	__shared__ uint32_t s_mat[CRC_TILE_LOG2 * 32];
	__shared__ uint32_t s_w[CRC_TILE_WORDS + CRC_TILE_WORDS / 16];      // one pad word per 16: thread-contiguous reads are conflict free
	__shared__ uint32_t s_c[CRC_THREADS];

	const uint32_t tid = threadIdx.x;
	const uint64_t s = blockIdx.x / P.n_tiles, t = blockIdx.x % P.n_tiles;
	for (uint32_t i = tid; i < 4 * 256; i += CRC_THREADS) s_tab[i] = P.tables[i];
	s_mat[tid] = P.tables[4 * 256 + tid];                                 // CRC_TILE_LOG2 * 32 == CRC_THREADS

	const uint64_t W = P.n_rows * P.row_words;
	const long long start = (long long)W - (long long)(t + 1) * CRC_TILE_WORDS;
	const uint8_t* seg = P.base + s * P.seg_stride;
	const bool flat = P.n_rows == 1 || P.row_pitch == P.row_words * 4;
	const uint32_t init = ~(P.crc_in ? P.crc_in[s] : 0u);
#pragma unroll 4
	for (uint32_t j = 0; j < (uint32_t)CRC_WPT; ++j) {
		const uint32_t idx = j * CRC_THREADS + tid;
		const long long w = start + idx;
		uint32_t v = 0;
		if (w >= 0) {
			const uint8_t* p = flat ? seg + (uint64_t)w * 4 : seg + ((uint64_t)w / P.row_words) * P.row_pitch + ((uint64_t)w % P.row_words) * 4;
			v = ld_nc_u32(p);
			if (w == 0) v ^= init;
		}
		s_w[idx + (idx >> 4)] = v;
	}
	__syncthreads();

	uint32_t c = 0;
	const uint32_t* mine = s_w + tid * (CRC_WPT + 1);
#pragma unroll
	for (int j = 0; j < CRC_WPT; ++j) {
		const uint32_t x = c ^ mine[j];
		c = s_tab[768 + (x & 255u)] ^ s_tab[512 + ((x >> 8) & 255u)] ^ s_tab[256 + ((x >> 16) & 255u)] ^ s_tab[x >> 24];
	}
	s_c[tid] = c;
	__syncthreads();
	// thread t holds the register of bytes [64t, 64t + 64): merge neighbours, the earlier half is shifted past the later one
#pragma unroll
	for (int l = 0; l < CRC_TILE_LOG2; ++l) {
		const uint32_t stride = 1u << l;
		if ((tid & (2 * stride - 1)) == 0) s_c[tid] = gf2_times(s_mat + l * 32, s_c[tid]) ^ s_c[tid + stride];
		__syncthreads();
	}
	if (tid == 0) P.tile_crc[s * P.n_tiles + t] = s_c[0];
}

// ---------------------------------------------------------------- the fast path: flat, 32-byte aligned messages
// Random bytes index 32 shared-memory banks, so the 1 KiB tables of crc_tile_kernel cost ~3.5 bank conflicts per look-up
// (measured: 0.15 of the HBM peak).  Here every table entry is replicated once per lane (entry e of table k for lane l at
// ((k * 256 + e) << 5) + l: bank == lane, never a conflict; 128 KiB, filled once per persistent block), a thread owns
// 256 contiguous bytes fetched with eight 256-bit loads (one whole sector per lane per load, no staging) and runs four
// independent 64-byte chains (the look-up chain of one register is pure latency), and a tile is what ONE WARP covers
// (8 KiB, merged with shuffles): warps never meet at a barrier, so the loads of one overlap the look-ups of the others.
constexpr int CRC2_THREADS = 512;
constexpr int CRC2_BYTES_PT = 256;
constexpr int CRC2_TILE_BYTES = 32 * CRC2_BYTES_PT;             // one WARP per tile: 8 KiB = 64 * 2^7
constexpr int CRC2_TILE_LOG2 = 7;
constexpr size_t CRC2_SMEM = ((size_t)4 * 256 * 32 + (size_t)CRC2_SHIFT_LEVELS * 4 * 256) * 4;

__device__ __forceinline__ void ld_nc_v8(const void* p, uint32_t (&r)[8])
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

// x^(8 * 64 * 2^l) * v: four look-ups in the level's byte-indexed table instead of 32 masked xors
__device__ __forceinline__ uint32_t crc_shift(const uint32_t* s_shift, int l, uint32_t v)
{
	const uint32_t* t = s_shift + l * 1024;
	return t[v & 255u] ^ t[256u + ((v >> 8) & 255u)] ^ t[512u + ((v >> 16) & 255u)] ^ t[768u + (v >> 24)];
}

__global__ void __launch_bounds__(CRC2_THREADS, 1)
crc_tile256_kernel(const CrcParams P)
{
	extern __shared__ __align__(16) uint32_t smem_crc[];
	uint32_t* s_tab = smem_crc;                               // [4][256][32]
	uint32_t* s_shift = s_tab + 4 * 256 * 32;                 // [CRC2_SHIFT_LEVELS][4][256]

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (uint32_t i = tid; i < 4 * 256 * 32; i += CRC2_THREADS) s_tab[i] = P.tables[i >> 5];
	for (uint32_t i = tid; i < (uint32_t)CRC2_SHIFT_LEVELS * 1024; i += CRC2_THREADS) s_shift[i] = P.tables[CRC_SHIFT_OFFSET + i];
	__syncthreads();
	const uint32_t* tab = s_tab + lane;

	const uint64_t n_bytes = P.n_rows * P.row_words * 4;
	const uint64_t n_units = P.n_seg * P.n_tiles;
	const uint64_t stride = (uint64_t)gridDim.x * (CRC2_THREADS / 32);
	for (uint64_t u = (uint64_t)blockIdx.x * (CRC2_THREADS / 32) + warp; u < n_units; u += stride) {
		const uint64_t s = u / P.n_tiles, t = u % P.n_tiles;
		const uint8_t* seg = P.base + s * P.seg_stride;
		const long long start = (long long)n_bytes - (long long)(t + 1) * CRC2_TILE_BYTES + (long long)lane * CRC2_BYTES_PT;
		// all eight loads are issued before anything looks at their data (a use in between would serialise them)
		const uint32_t init = (start <= 0 && start > -(long long)CRC2_BYTES_PT) ? ~(P.crc_in ? P.crc_in[s] : 0u) : 0u;
		uint32_t w[8][8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const long long off = start + j * 32;
			if (off >= 0) ld_nc_v8(seg + off, w[j]);
			else {
#pragma unroll
				for (int q = 0; q < 8; ++q) w[j][q] = 0u;
			}
		}
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (start + j * 32 == 0) w[j][0] ^= init;
		uint32_t c[4] = {0u, 0u, 0u, 0u};
#pragma unroll
		for (int i = 0; i < 16; ++i) {
#pragma unroll
			for (int ch = 0; ch < 4; ++ch) {
				const uint32_t x = c[ch] ^ w[ch * 2 + (i >> 3)][i & 7];
				c[ch] = tab[(768u + (x & 255u)) << 5] ^ tab[(512u + ((x >> 8) & 255u)) << 5] ^ tab[(256u + ((x >> 16) & 255u)) << 5] ^ tab[(x >> 24) << 5];
			}
		}
		const uint32_t c01 = crc_shift(s_shift, 0, c[0]) ^ c[1], c23 = crc_shift(s_shift, 0, c[2]) ^ c[3];
		uint32_t r = crc_shift(s_shift, 1, c01) ^ c23;
		// lanes: 256 B -> 8 KiB (levels 2 .. 6); the earlier half is shifted past the later one
#pragma unroll
		for (int l = 0; l < 5; ++l) {
			const uint32_t other = __shfl_down_sync(0xFFFFFFFFu, r, 1u << l);
			if ((lane & ((2u << l) - 1u)) == 0) r = crc_shift(s_shift, 2 + l, r) ^ other;
		}
		if (lane == 0) P.tile_crc[u] = r;
	}
}

// in: [n_seg][n] registers, element 0 = the END of the message, every element standing for 64 * 2^level0 bytes;
// out: [n_seg][ceil(n / CRC_FAN)], every element standing for 64 * 2^(level0 + 10) bytes.  finalize: n_out == 1, write ~register.
__global__ void __launch_bounds__(CRC_FAN)
crc_combine_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ out, uint64_t n_out, int level0,
	const uint32_t* __restrict__ mats, int finalize)
{
	__shared__ uint32_t s_c[CRC_FAN];
	__shared__ uint32_t s_mat[10 * 32];
	const uint32_t tid = threadIdx.x;
	const uint64_t s = blockIdx.x / n_out, bx = blockIdx.x % n_out;
	if (tid < 320) s_mat[tid] = mats[(size_t)level0 * 32 + tid];
	const uint64_t i = bx * CRC_FAN + tid;
	s_c[tid] = (i < n) ? in[s * n + i] : 0u;
	__syncthreads();
#pragma unroll
	for (int l = 0; l < 10; ++l) {
		const uint32_t stride = 1u << l;
		// element tid + stride lies EARLIER in the message: it is shifted past the bytes the elements [tid, tid + stride) stand for
		if ((tid & (2 * stride - 1)) == 0) s_c[tid] ^= gf2_times(s_mat + l * 32, s_c[tid + stride]);
		__syncthreads();
	}
	if (tid == 0) out[s * n_out + bx] = finalize ? ~s_c[0] : s_c[0];
}

// ---------------------------------------------------------------- host: tables
static const uint32_t* crc_tables_dev(int device)
{
	static std::mutex mu;
	static std::vector<uint32_t*> per_device;
	std::lock_guard<std::mutex> lock(mu);
	if ((int)per_device.size() <= device) per_device.resize(device + 1, nullptr);
	if (per_device[device]) return per_device[device];

	std::vector<uint32_t> h;
	crc32_build_tables(h);          // crc_tables.h (also checked on the CPU: tests/test_host_emul.py)
	uint32_t* d = nullptr;
	if (cudaMalloc(&d, h.size() * sizeof(uint32_t)) != cudaSuccess) return nullptr;
	if (cudaMemcpy(d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
	per_device[device] = d;
	return d;
}

uint64_t crc32_tiles(uint64_t message_bytes) { return ceil_div(std::max<uint64_t>(message_bytes / 4, 1), CRC_TILE_WORDS); }   // (upper bound for both tile sizes)

// words of workspace for crc32_launch
size_t crc32_workspace_words(uint64_t n_seg, uint64_t message_bytes)
{
	const uint64_t nt = ceil_div(std::max<uint64_t>(message_bytes, 1), CRC2_TILE_BYTES);      // the smaller of the two tile sizes
	return (size_t)(n_seg * (nt + ceil_div(nt, CRC_FAN) + 2));
}

// d_crc_in (may be NULL) / d_crc_out: [n_seg] on the device; may alias.  Everything is queued on `stream`.
int crc32_launch(int device, const uint8_t* d_base, uint64_t n_seg, uint64_t seg_stride, uint64_t n_rows, uint64_t row_bytes,
	uint64_t row_pitch, const uint32_t* d_crc_in, uint32_t* d_crc_out, uint32_t* d_ws, cudaStream_t stream)
{
	if (n_seg == 0) return KWG_OK;
	if (n_rows == 0 || row_bytes == 0 || row_bytes % 4 || row_pitch % 4 || seg_stride % 4 || (reinterpret_cast<uintptr_t>(d_base) & 3u))
		return fail(KWG_ERR_INVALID_ARG, "device crc32 takes non-empty messages made of 4-byte aligned 32-bit words");
	const uint32_t* tab = crc_tables_dev(device);
	if (!tab) return fail(KWG_ERR_CUDA, "crc32 tables could not be uploaded");
	CrcParams P{};
	P.base = d_base; P.n_seg = n_seg; P.seg_stride = seg_stride;
	P.n_rows = n_rows; P.row_words = row_bytes / 4; P.row_pitch = row_pitch;
	P.crc_in = d_crc_in;
	P.tile_crc = d_ws;
	P.tables = tab;
	const uint64_t n_bytes = n_rows * row_bytes;
	const bool flat = n_rows == 1 || row_pitch == row_bytes;
	const bool fast = flat && n_bytes % 32 == 0 && seg_stride % 32 == 0 && (reinterpret_cast<uintptr_t>(d_base) & 31u) == 0 &&
	                  n_seg * n_bytes >= (uint64_t)(128 << 10);
	int level;
	if (fast) {
		static std::once_flag once[64];
		cudaError_t attr = cudaSuccess;
		std::call_once(once[device & 63], [&]() { attr = cudaFuncSetAttribute(crc_tile256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CRC2_SMEM); });
		if (attr != cudaSuccess) return fail(KWG_ERR_CUDA, std::string("crc32: ") + cudaGetErrorString(attr));
		P.n_tiles = ceil_div(n_bytes, CRC2_TILE_BYTES);
		const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(n_seg * P.n_tiles, CRC2_THREADS / 32), (uint64_t)sm_count(device));
		crc_tile256_kernel<<<grid, CRC2_THREADS, CRC2_SMEM, stream>>>(P);
		level = CRC2_TILE_LOG2;
	} else {
		P.n_tiles = crc32_tiles(n_bytes);
		if (n_seg * P.n_tiles > 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "crc32: too many tiles for one launch");
		crc_tile_kernel<<<(unsigned)(n_seg * P.n_tiles), CRC_THREADS, 0, stream>>>(P);
		level = CRC_TILE_LOG2;
	}
	KWG_LAUNCHED();
	uint32_t* in = d_ws;
	uint32_t* other = d_ws + n_seg * P.n_tiles;
	uint64_t n = P.n_tiles;
	for (;;) {
		const uint64_t n_out = ceil_div(n, CRC_FAN);
		const bool last = n_out == 1;
		uint32_t* out = last ? d_crc_out : other;
		crc_combine_kernel<<<(unsigned)(n_seg * n_out), CRC_FAN, 0, stream>>>(in, n, out, n_out, level, tab + 4 * 256, last ? 1 : 0);
		KWG_LAUNCHED();
		if (last) break;
		// (the next level's output fits where this level's input was: n_out <= n / 1024 + 1)
		uint32_t* t = in; in = other; other = t;
		n = n_out;
		level += 10;
		if (level + 10 > CRC_LEVELS + 10) return fail(KWG_ERR_INVALID_ARG, "crc32: message too long");
	}
	return KWG_OK;
}

} // namespace kwg

using namespace kwg;

extern "C" {

int kwg_crc32_dev(int device, const uint8_t* d_data, uint64_t n_rows, uint64_t row_bytes, uint64_t row_pitch, uint32_t crc_in,
	uint32_t* crc_out, void* stream)
{
	if (!d_data || !crc_out) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	int rc = select_device(device);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	// grow-only workspace per device; the call synchronises its stream before returning, so holding the lock for the
	// whole call is what keeps concurrent callers on one device apart
	struct Ws { uint32_t* p = nullptr; size_t words = 0; uint32_t* h = nullptr; };
	static std::mutex mu;
	static std::vector<Ws> cache;
	std::lock_guard<std::mutex> lock(mu);
	if ((int)cache.size() <= device) cache.resize(device + 1);
	Ws& W = cache[device];
	const size_t words = crc32_workspace_words(1, n_rows * row_bytes) + 2;
	if (words > W.words) {
		if (W.p) KWG_CUDA(cudaFree(W.p));
		W.p = nullptr; W.words = 0;
		KWG_CUDA(cudaMalloc(&W.p, (words + words / 4) * sizeof(uint32_t)));
		W.words = words + words / 4;
	}
	if (!W.h) KWG_CUDA(cudaMallocHost(&W.h, 2 * sizeof(uint32_t)));
	uint32_t* d_io = W.p + W.words - 2;
	W.h[0] = crc_in;
	KWG_CUDA(cudaMemcpyAsync(d_io, W.h, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	rc = crc32_launch(device, d_data, 1, 0, n_rows, row_bytes, row_pitch, d_io, d_io + 1, W.p, st);
	if (rc) return rc;
	KWG_CUDA(cudaMemcpyAsync(W.h + 1, d_io + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	KWG_CUDA(cudaStreamSynchronize(st));
	*crc_out = W.h[1];
	return KWG_OK;
}

} // extern "C"
