// Bloom filters -> bit-sliced database rows (reference build_db.cpp:259-304) on sm_100a.
//
// DB[k][j] = Bloom_j[k]: slice (row) k holds bit j of filter j at byte j/8, bit j%8.
// Each thread owns a 32 filters x 32 bits tile in registers: 32 coalesced 4-byte loads (a warp
// reads 128 contiguous bytes of one filter per instruction), a 5-stage masked-swap bit-matrix
// transpose, then the block (8 warps = 256 filters) exchanges the 32-bit pieces through a
// swizzled shared-memory tile so that every global store instruction writes whole 32-byte
// sectors of four rows.  Pure HBM streaming: 1 bit read + 1 bit written per bit moved.
#include "common.cuh"

#include <mutex>
#include <new>

#include <algorithm>
#include <vector>

namespace kwg {

constexpr int TR_WARPS = 8;                    // filter groups (of 32 filters) per block
constexpr int TR_THREADS = TR_WARPS * 32;
constexpr int TR_WORDS = 32;                   // 32-bit words of every filter per block (1024 slices)

__global__ void __launch_bounds__(TR_THREADS)
transpose_kernel(const uint8_t* __restrict__ filters, uint64_t filter_pitch, uint32_t n_filters, uint64_t n_words,
	uint8_t* __restrict__ dest, uint64_t dest_pitch, uint32_t ny)
{
	__shared__ __align__(16) uint32_t tile[32 * 32 * TR_WARPS];       // [b][lw][group ^ swizzle], 32 KiB

	// 1-D grid, column-group index fastest: blocks that run together write the same rows side by
	// side (whole 128-byte lines reach DRAM) while each still reads whole lines of its filters.
	const uint32_t by = blockIdx.x % ny;
	const uint64_t bx = blockIdx.x / ny;
	const uint32_t g = threadIdx.x >> 5, lw = threadIdx.x & 31;
	const uint64_t word = bx * TR_WORDS + lw;
	const uint32_t group = by * TR_WARPS + g;
	const uint32_t f0 = group * 32;

	uint32_t a[32];
	if (word < n_words && f0 < n_filters) {
		const uint8_t* src = filters + (uint64_t)f0 * filter_pitch + word * 4;
		if (f0 + 32 <= n_filters) {
#pragma unroll
			for (int i = 0; i < 32; ++i) a[i] = ld_nc_u32(src + (uint64_t)i * filter_pitch);
		} else {
#pragma unroll
			for (int i = 0; i < 32; ++i) a[i] = (f0 + i < n_filters) ? ld_nc_u32(src + (uint64_t)i * filter_pitch) : 0u;
		}
		transpose32(a);
	} else {
#pragma unroll
		for (int i = 0; i < 32; ++i) a[i] = 0u;
	}

	// a[b] = 32 column bits of row (word*32 + b).  Park at [b][lw][g ^ (lw>>2 & 7)]: conflict free.
	const uint32_t sw = (lw >> 2) & 7;
#pragma unroll
	for (int b = 0; b < 32; ++b) tile[(b * 32 + lw) * TR_WARPS + (g ^ sw)] = a[b];
	__syncthreads();

	// Copy-out with 16-byte accesses: a thread moves half a row of the block (4 groups = 128 filters), two neighbouring
	// lanes write the 32 bytes of one row, a warp covers 16 rows.  Thread = (half, q = lw & 3, b); iteration `it` is
	// lw >> 2, so the swizzle of the row (g ^ it) is known at compile time: the 16-byte piece holding the wanted half
	// is piece (half ^ (it >> 2)), and its four words are in the order j ^ (it & 3).  A quarter warp reads four
	// consecutive rows of the tile: conflict free.  8 LDS.128 + 8 STG.128 per thread instead of 32 + 32 narrow ones.
	const uint32_t hl = threadIdx.x & 1;                 // which 16 bytes of the row's 32
	const uint32_t q = (threadIdx.x >> 1) & 3;           // low two bits of lw
	const uint32_t bq = threadIdx.x >> 3;                // b: 0..31
	const uint64_t cbyte = (uint64_t)by * TR_WARPS * 4 + hl * 16;
	if (cbyte < dest_pitch) {
		uint8_t* out0 = dest + ((bx * TR_WORDS + q) * 32 + bq) * dest_pitch + cbyte;
		const uint32_t* t0 = tile + (bq * 32 + q) * TR_WARPS;
#pragma unroll
		for (int it = 0; it < 8; ++it) {
			const uint4 v = *reinterpret_cast<const uint4*>(t0 + (it * 4) * TR_WARPS + ((hl ^ (uint32_t)(it >> 2)) << 2));
			const uint32_t w[4] = {v.x, v.y, v.z, v.w};
			const uint4 o = make_uint4(w[0 ^ (it & 3)], w[1 ^ (it & 3)], w[2 ^ (it & 3)], w[3 ^ (it & 3)]);
			if (bx * TR_WORDS + it * 4 + q < n_words)
				st_na_v4(out0 + ((uint64_t)(it * 4) * 32) * dest_pitch, o);
		}
	}
}

int transpose_launch(const uint8_t* d_filters, uint64_t filter_pitch, uint32_t n_filters, uint64_t chunk_bits,
	uint8_t* d_dest, uint64_t dest_pitch, cudaStream_t stream)
{
	const uint64_t n_words = chunk_bits / 32;
	if (n_words == 0 || n_filters == 0) return KWG_OK;
	// every 4-byte column word up to dest_pitch is written (zero beyond n_filters)
	const uint64_t ny = ceil_div(dest_pitch / 4, TR_WARPS);         // blocks across a row
	const uint64_t blocks = ceil_div(n_words, TR_WORDS) * ny;
	if (blocks > 0x7FFFFFFFull) return fail(KWG_ERR_INVALID_ARG, "transpose chunk too large for one launch");
	transpose_kernel<<<(unsigned)blocks, TR_THREADS, 0, stream>>>(d_filters, filter_pitch, n_filters, n_words, d_dest, dest_pitch, (uint32_t)ny);
	KWG_LAUNCHED();
	return KWG_OK;
}

} // namespace kwg

using namespace kwg;

extern "C" {

int kwg_transpose_dev(int device, const uint8_t* d_filters, uint64_t filter_pitch, uint32_t n_filters,
	uint64_t chunk_bits, uint8_t* d_dest, uint64_t dest_pitch, void* stream)
{
	if (!d_filters || !d_dest) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (n_filters == 0) return fail(KWG_ERR_INVALID_ARG, "n_filters must be > 0 (reference build_db.cpp:30-32)");
	if (chunk_bits % 32) return fail(KWG_ERR_INVALID_ARG, "chunk_bits must be a multiple of 32 on the device path");
	if (filter_pitch % 16 || dest_pitch % 16) return fail(KWG_ERR_INVALID_ARG, "pitches must be multiples of 16 bytes");
	if (filter_pitch < chunk_bits / 8) return fail(KWG_ERR_INVALID_ARG, "filter_pitch smaller than a filter chunk");
	if (dest_pitch < ceil_div(n_filters, 8)) return fail(KWG_ERR_INVALID_ARG, "dest_pitch smaller than a slice");
	if ((reinterpret_cast<uintptr_t>(d_filters) | reinterpret_cast<uintptr_t>(d_dest)) & 15u)
		return fail(KWG_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
	int rc = select_device(device);
	if (rc) return rc;
	return transpose_launch(d_filters, filter_pitch, n_filters, chunk_bits, d_dest, dest_pitch, (cudaStream_t)stream);
}

struct TransposeCache {
	std::mutex mu;                     // one caller at a time per DEVICE (callers on other devices run beside it)
	cudaStream_t stream[2] = {nullptr, nullptr};
	cudaEvent_t crc_done[2] = {nullptr, nullptr};
	uint8_t* d_in[2] = {nullptr, nullptr};
	uint8_t* d_out[2] = {nullptr, nullptr};
	uint32_t* d_ws[2] = {nullptr, nullptr};
	uint32_t* d_crc = nullptr;
	size_t in_cap[2] = {0, 0}, out_cap[2] = {0, 0}, ws_cap[2] = {0, 0}, crc_cap = 0;
};

static std::mutex g_tr_mu;                           // guards the table below, never held while work runs
static std::vector<TransposeCache*> g_tr_caches;

static TransposeCache* transpose_cache(int device)
{
	std::lock_guard<std::mutex> lock(g_tr_mu);
	if ((int)g_tr_caches.size() <= device) g_tr_caches.resize(device + 1, nullptr);
	if (!g_tr_caches[device]) g_tr_caches[device] = new (std::nothrow) TransposeCache();
	return g_tr_caches[device];
}

static int transpose_host(int device, const uint8_t* const* filter_chunks, uint32_t n_filters, uint64_t chunk_bits, uint8_t* dest,
	uint32_t* filter_crc, uint32_t* dest_crc)
{
	if (!filter_chunks || !dest) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (n_filters == 0) return fail(KWG_ERR_INVALID_ARG, "n_filters must be > 0 (reference build_db.cpp:30-32)");
	if (chunk_bits % 8) return fail(KWG_ERR_INVALID_ARG, "chunk_bits must be a multiple of 8 (reference build_db.cpp:238-243)");
	if (chunk_bits == 0) return KWG_OK;
	if ((filter_crc || dest_crc) && chunk_bits % 32) return fail(KWG_ERR_INVALID_ARG, "device crc32 needs chunk_bits % 32 == 0");
	if (dest_crc && n_filters % 32) return fail(KWG_ERR_INVALID_ARG, "device crc32 of the slices needs n_filters % 32 == 0");
	int rc = select_device(device);
	if (rc) return rc;

	const uint64_t row_bytes = ceil_div(n_filters, 8);
	const uint64_t dest_pitch = round_up(row_bytes, 16);
	const uint64_t chunk_bytes = chunk_bits / 8;

	// The slice axis is worked through in pieces of ~128 MiB on TWO lanes (stream + staging buffers each): while one
	// piece is transposed and copied back, the next one is already coming in -- the two copy directions and the
	// kernel overlap.  The running crc32 values chain from piece to piece through an event.
	const uint64_t budget = 128ull << 20;
	uint64_t piece_bits = std::max<uint64_t>(1024, (budget / std::max<uint64_t>(n_filters / 8, 16)) & ~1023ull);
	piece_bits = std::min(piece_bits, round_up(chunk_bits, 32));
	const uint64_t piece_pitch = round_up(piece_bits / 8, 16);
	const int n_lanes = (chunk_bits > piece_bits) ? 2 : 1;
	for (uint32_t j = 0; j < n_filters; ++j)
		if (!filter_chunks[j]) return fail(KWG_ERR_INVALID_ARG, "NULL filter chunk");
	size_t src_stride = 0;                           // != 0: filter_chunks[j] == filter_chunks[0] + j * src_stride
	if (n_filters > 1 && filter_chunks[1] > filter_chunks[0] && (size_t)(filter_chunks[1] - filter_chunks[0]) >= chunk_bytes) {
		src_stride = (size_t)(filter_chunks[1] - filter_chunks[0]);
		for (uint32_t j = 2; j < n_filters && src_stride; ++j)
			if (filter_chunks[j] != filter_chunks[0] + (size_t)j * src_stride) src_stride = 0;
	}
	const bool want_crc = filter_crc || dest_crc;

	// Streams and staging buffers live in a per-device cache that only grows (kwg_release_caches frees it): build_db
	// calls this once per chunk, and allocating / releasing gigabytes per call costs anything from 10 to 400 ms.  The
	// device's lock is held for the whole call (it synchronises before returning), which keeps concurrent callers on ONE
	// device apart; callers on different devices do not meet.
	TransposeCache* cache = transpose_cache(device);
	if (!cache) return fail(KWG_ERR_NO_MEMORY, "transpose: out of host memory");
	std::lock_guard<std::mutex> lock(cache->mu);
	TransposeCache& C = *cache;
	cudaStream_t* stream = C.stream;
	cudaEvent_t* crc_done = C.crc_done;
	uint8_t** d_in = C.d_in;
	uint8_t** d_out = C.d_out;
	uint32_t** d_ws = C.d_ws;
	auto cleanup = [&]() {            // after an error: drain whatever was queued
		for (int l = 0; l < 2; ++l) if (stream[l]) cudaStreamSynchronize(stream[l]);
	};
#define KWG_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); \
	return fail(_e == cudaErrorMemoryAllocation ? KWG_ERR_NO_MEMORY : KWG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
#define KWG_GROW(ptr, cap, need) do { if ((need) > (cap)) { if (ptr) KWG_TRY(cudaFree(ptr)); (ptr) = nullptr; (cap) = 0; \
	KWG_TRY(cudaMalloc(&(ptr), (need))); (cap) = (need); } } while (0)
	for (int l = 0; l < n_lanes; ++l) {
		if (!stream[l]) KWG_TRY(cudaStreamCreateWithFlags(&stream[l], cudaStreamNonBlocking));
		if (!crc_done[l]) KWG_TRY(cudaEventCreateWithFlags(&crc_done[l], cudaEventDisableTiming));
		KWG_GROW(d_in[l], C.in_cap[l], (size_t)n_filters * piece_pitch);
		KWG_GROW(d_out[l], C.out_cap[l], (size_t)piece_bits * dest_pitch);
		if (want_crc) {
			const size_t ws = std::max(crc32_workspace_words(n_filters, piece_bits / 8), crc32_workspace_words(1, piece_bits * row_bytes));
			KWG_GROW(d_ws[l], C.ws_cap[l], ws * sizeof(uint32_t));
		}
	}
	uint32_t*& d_crc = C.d_crc;                      // [n_filters] running filter values, then the slice value
	if (want_crc) {
		KWG_GROW(d_crc, C.crc_cap, ((size_t)n_filters + 1) * sizeof(uint32_t));
		if (filter_crc) KWG_TRY(cudaMemcpyAsync(d_crc, filter_crc, (size_t)n_filters * sizeof(uint32_t), cudaMemcpyHostToDevice, stream[0]));
		if (dest_crc) KWG_TRY(cudaMemcpyAsync(d_crc + n_filters, dest_crc, sizeof(uint32_t), cudaMemcpyHostToDevice, stream[0]));
	}

	uint64_t piece = 0;
	for (uint64_t bit0 = 0; bit0 < chunk_bits; bit0 += piece_bits, ++piece) {
		const int l = (int)(piece % n_lanes);
		cudaStream_t st = stream[l];
		const uint64_t bits = std::min(piece_bits, chunk_bits - bit0);
		const uint64_t bytes = bits / 8;
		const uint64_t bits32 = round_up(bits, 32);
		// (work queued on this lane two pieces ago has to be over before its buffers are filled again: stream order)
		if (bits32 != bits) KWG_TRY(cudaMemsetAsync(d_in[l], 0, (size_t)n_filters * piece_pitch, st));
		if (src_stride) {
			// the caller's chunks are equally spaced (one staging buffer, as in build_db): ONE pitched copy per piece --
			// a copy call costs ~20 us whatever its size, 2048 of them per piece would cost more than the transfer
			KWG_TRY(cudaMemcpy2DAsync(d_in[l], piece_pitch, filter_chunks[0] + bit0 / 8, src_stride, bytes, n_filters, cudaMemcpyHostToDevice, st));
		} else {
			for (uint32_t j = 0; j < n_filters; ++j)
				KWG_TRY(cudaMemcpyAsync(d_in[l] + (uint64_t)j * piece_pitch, filter_chunks[j] + bit0 / 8, bytes, cudaMemcpyHostToDevice, st));
		}
		// the running checksums of this piece continue those of the previous piece (other lane)
		if (want_crc && piece > 0 && n_lanes == 2) KWG_TRY(cudaStreamWaitEvent(st, crc_done[1 - l], 0));
		if (filter_crc) {      // every filter's piece is one message of `bytes` bytes at d_in + j * piece_pitch
			rc = crc32_launch(device, d_in[l], n_filters, piece_pitch, 1, bytes, bytes, d_crc, d_crc, d_ws[l], st);
			if (rc) { cleanup(); return rc; }
		}
		rc = transpose_launch(d_in[l], piece_pitch, n_filters, bits32, d_out[l], dest_pitch, st);
		if (rc) { cleanup(); return rc; }
		if (dest_crc) {        // the slices of the piece: `bits` rows of row_bytes at dest_pitch, one running message
			rc = crc32_launch(device, d_out[l], 1, 0, bits, row_bytes, dest_pitch, d_crc + n_filters, d_crc + n_filters, d_ws[l], st);
			if (rc) { cleanup(); return rc; }
		}
		if (want_crc) KWG_TRY(cudaEventRecord(crc_done[l], st));
		KWG_TRY(cudaMemcpy2DAsync(dest + bit0 * row_bytes, row_bytes, d_out[l], dest_pitch, row_bytes, bits, cudaMemcpyDeviceToHost, st));
	}
	(void)chunk_bytes;
	const int last = (int)((piece + n_lanes - 1) % n_lanes);
	if (filter_crc) KWG_TRY(cudaMemcpyAsync(filter_crc, d_crc, (size_t)n_filters * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream[last]));
	if (dest_crc) KWG_TRY(cudaMemcpyAsync(dest_crc, d_crc + n_filters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream[last]));
	for (int l = 0; l < n_lanes; ++l) KWG_TRY(cudaStreamSynchronize(stream[l]));
#undef KWG_GROW
#undef KWG_TRY
	return KWG_OK;
}

void kwg_release_caches(void)
{
	std::vector<TransposeCache*> all;
	{
		std::lock_guard<std::mutex> lock(g_tr_mu);
		all = g_tr_caches;
	}
	for (size_t d = 0; d < all.size(); ++d) {
		TransposeCache* C = all[d];
		if (!C) continue;
		std::lock_guard<std::mutex> lock(C->mu);
		if (cudaSetDevice((int)d) != cudaSuccess) continue;
		for (int l = 0; l < 2; ++l) {
			if (C->stream[l]) cudaStreamSynchronize(C->stream[l]);
			cudaFree(C->d_in[l]); cudaFree(C->d_out[l]); cudaFree(C->d_ws[l]);
			C->d_in[l] = C->d_out[l] = nullptr; C->d_ws[l] = nullptr;
			C->in_cap[l] = C->out_cap[l] = C->ws_cap[l] = 0;
		}
		cudaFree(C->d_crc);
		C->d_crc = nullptr; C->crc_cap = 0;
	}
}

int kwg_transpose(int device, const uint8_t* const* filter_chunks, uint32_t n_filters, uint64_t chunk_bits, uint8_t* dest)
{
	return transpose_host(device, filter_chunks, n_filters, chunk_bits, dest, nullptr, nullptr);
}

int kwg_transpose_crc(int device, const uint8_t* const* filter_chunks, uint32_t n_filters, uint64_t chunk_bits, uint8_t* dest,
	uint32_t* filter_crc, uint32_t* dest_crc)
{
	return transpose_host(device, filter_chunks, n_filters, chunk_bits, dest, filter_crc, dest_crc);
}

} // extern "C"
