// Bloom filters -> bit-sliced database rows (reference build_db.cpp:259-304) on sm_100a.
//
// DB[k][j] = Bloom_j[k]: slice (row) k holds bit j of filter j at byte j/8, bit j%8.
// Each thread owns a 32 filters x 32 bits tile in registers: 32 coalesced 4-byte loads (a warp
// reads 128 contiguous bytes of one filter per instruction), a 5-stage masked-swap bit-matrix
// transpose, then the block (8 warps = 256 filters) exchanges the 32-bit pieces through a
// swizzled shared-memory tile so that every global store instruction writes whole 32-byte
// sectors of four rows.  Pure HBM streaming: 1 bit read + 1 bit written per bit moved.
#include "common.cuh"

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
	__shared__ uint32_t tile[32 * 32 * TR_WARPS];       // [b][lw][group ^ swizzle], 32 KiB

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

	// 8 consecutive lanes write the 32 bytes (8 groups) of one row; a warp covers 4 rows.  Fully unrolled: with
	// R = it*8 + (tid>>5), everything that depends on `it` is a compile-time constant (lwh = it>>2,
	// b = (it&3)*8 + tid>>5), so an iteration is one LDS, one address add and one predicated store.
	const uint32_t wd = threadIdx.x & 7;                 // group within the block
	const uint32_t q = (threadIdx.x >> 3) & 3;           // low two bits of lw
	const uint32_t r0 = threadIdx.x >> 5;                // 0..7
	const uint64_t cbyte = (uint64_t)by * TR_WARPS * 4 + wd * 4;
	if (cbyte < dest_pitch) {
		uint8_t* out0 = dest + ((bx * TR_WORDS + q) * 32 + r0) * dest_pitch + cbyte;
		const uint32_t* t0 = tile + (r0 * 32 + q) * TR_WARPS;
#pragma unroll
		for (int it = 0; it < 32; ++it) {
			const int lwh = it >> 2, bb = (it & 3) * 8;      // rlw = lwh*4 + q, b = bb + r0
			const uint32_t v = t0[(bb * 32 + lwh * 4) * TR_WARPS + (wd ^ lwh)];
			if (bx * TR_WORDS + lwh * 4 + q < n_words)
				st_na_u32(out0 + ((uint64_t)(lwh * 4) * 32 + bb) * dest_pitch, v);
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

	// Work through the slice axis in pieces that keep the staging buffers bounded (<= ~1 GiB each).
	const uint64_t budget = 1ull << 30;
	uint64_t piece_bits = std::max<uint64_t>(1024, (budget / std::max<uint64_t>(n_filters / 8, 16)) & ~1023ull);
	piece_bits = std::min(piece_bits, round_up(chunk_bits, 32));
	const uint64_t piece_pitch = round_up(piece_bits / 8, 16);

	cudaStream_t stream = nullptr;
	uint8_t *d_in = nullptr, *d_out = nullptr;
	uint32_t *d_crc = nullptr, *d_ws = nullptr;      // d_crc: [n_filters] running filter values, then the slice value
	auto cleanup = [&]() {
		if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
		cudaFree(d_in); cudaFree(d_out); cudaFree(d_crc); cudaFree(d_ws);
	};
#define KWG_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); \
	return fail(_e == cudaErrorMemoryAllocation ? KWG_ERR_NO_MEMORY : KWG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
	KWG_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
	KWG_TRY(cudaMalloc(&d_in, (size_t)n_filters * piece_pitch));
	KWG_TRY(cudaMalloc(&d_out, (size_t)piece_bits * dest_pitch));
	if (filter_crc || dest_crc) {
		const size_t ws = std::max(crc32_workspace_words(n_filters, piece_bits / 8), crc32_workspace_words(1, piece_bits * row_bytes));
		KWG_TRY(cudaMalloc(&d_ws, ws * sizeof(uint32_t)));
		KWG_TRY(cudaMalloc(&d_crc, ((size_t)n_filters + 1) * sizeof(uint32_t)));
		if (filter_crc) KWG_TRY(cudaMemcpyAsync(d_crc, filter_crc, (size_t)n_filters * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
		if (dest_crc) KWG_TRY(cudaMemcpyAsync(d_crc + n_filters, dest_crc, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
	}

	for (uint64_t bit0 = 0; bit0 < chunk_bits; bit0 += piece_bits) {
		const uint64_t bits = std::min(piece_bits, chunk_bits - bit0);
		const uint64_t bytes = bits / 8;
		const uint64_t bits32 = round_up(bits, 32);
		if (bits32 != bits) KWG_TRY(cudaMemsetAsync(d_in, 0, (size_t)n_filters * piece_pitch, stream));
		// one strided copy per filter: host pointers are unrelated
		for (uint32_t j = 0; j < n_filters; ++j) {
			if (!filter_chunks[j]) { cleanup(); return fail(KWG_ERR_INVALID_ARG, "NULL filter chunk"); }
			KWG_TRY(cudaMemcpyAsync(d_in + (uint64_t)j * piece_pitch, filter_chunks[j] + bit0 / 8, bytes, cudaMemcpyHostToDevice, stream));
		}
		if (filter_crc) {      // every filter's piece is one message of `bytes` bytes at d_in + j * piece_pitch
			rc = crc32_launch(device, d_in, n_filters, piece_pitch, 1, bytes, bytes, d_crc, d_crc, d_ws, stream);
			if (rc) { cleanup(); return rc; }
		}
		rc = transpose_launch(d_in, piece_pitch, n_filters, bits32, d_out, dest_pitch, stream);
		if (rc) { cleanup(); return rc; }
		if (dest_crc) {        // the slices of the piece: `bits` rows of row_bytes at dest_pitch, one running message
			rc = crc32_launch(device, d_out, 1, 0, bits, row_bytes, dest_pitch, d_crc + n_filters, d_crc + n_filters, d_ws, stream);
			if (rc) { cleanup(); return rc; }
		}
		KWG_TRY(cudaMemcpy2DAsync(dest + bit0 * row_bytes, row_bytes, d_out, dest_pitch, row_bytes, bits, cudaMemcpyDeviceToHost, stream));
		KWG_TRY(cudaStreamSynchronize(stream));
	}
	(void)chunk_bytes;
	if (filter_crc) KWG_TRY(cudaMemcpyAsync(filter_crc, d_crc, (size_t)n_filters * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	if (dest_crc) KWG_TRY(cudaMemcpyAsync(dest_crc, d_crc + n_filters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	if (filter_crc || dest_crc) KWG_TRY(cudaStreamSynchronize(stream));
#undef KWG_TRY
	cleanup();
	return KWG_OK;
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
