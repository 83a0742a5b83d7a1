// Column concatenation of bit-slices: the body of merge_database_files()'s chunk loop (reference merge_db.cpp:489-584),
// which copies source 1's slice, then moves the bits of source 2 one by one (BitVector::get_bit / set_bit) behind them and,
// when the first destination is full, into a second one.  Here a thread assembles one destination byte from at most two
// bytes of each source.
#include "common.cuh"

namespace kwg {

// bits [start, start + 8) of a slice of nbits bits (LSB first, bloom.h:131-163); positions outside [0, nbits) read as zero
__device__ __forceinline__ uint32_t slice_bits8(const uint8_t* __restrict__ row, int64_t nbits, int64_t start)
{
	if (start >= nbits || start <= -8) return 0u;
	const int64_t s = start > 0 ? start : 0;
	const int64_t e = (start + 8 < nbits) ? start + 8 : nbits;
	const uint64_t b0 = (uint64_t)s >> 3;
	uint32_t v = row[b0];
	if ((uint64_t)((e - 1) >> 3) > b0) v |= (uint32_t)row[b0 + 1] << 8;
	v >>= (uint32_t)(s & 7);
	v &= (1u << (uint32_t)(e - s)) - 1u;
	return v << (uint32_t)(s - start);
}

struct MergeParams {
	const uint8_t* src1; uint64_t p1; uint32_t n1;
	const uint8_t* src2; uint64_t p2; uint32_t n2;
	uint64_t n_slices;
	uint32_t take;             // columns of source 2 that follow source 1 in destination 1; the rest go to destination 2
	uint8_t* dst1; uint64_t pd1;
	uint8_t* dst2; uint64_t pd2;
};

__global__ void __launch_bounds__(256)
merge_slices_kernel(const MergeParams P)
{
	const uint64_t per = P.pd1 + P.pd2;
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_slices * per) return;
	const uint64_t r = i / per, b = i % per;
	const uint8_t* s1 = P.src1 + r * P.p1;
	const uint8_t* s2 = P.src2 + r * P.p2;
	if (b < P.pd1) {
		const int64_t bit = (int64_t)b * 8;
		P.dst1[r * P.pd1 + b] = (uint8_t)(slice_bits8(s1, P.n1, bit) | slice_bits8(s2, P.take, bit - (int64_t)P.n1));
	} else {
		const uint64_t b2 = b - P.pd1;
		P.dst2[r * P.pd2 + b2] = (uint8_t)slice_bits8(s2, P.n2, (int64_t)P.take + (int64_t)b2 * 8);
	}
}

} // namespace kwg

using namespace kwg;

extern "C" int kwg_merge_slices(int device, const uint8_t* src1, uint32_t n1, const uint8_t* src2, uint32_t n2, uint64_t n_slices,
	uint32_t n_dst1, uint8_t* dst1, uint8_t* dst2)
{
	if (!src1 || !src2 || !dst1) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (n1 == 0 || n2 == 0) return fail(KWG_ERR_INVALID_ARG, "a source without filters");
	if (n_dst1 < n1 || (uint64_t)n_dst1 > (uint64_t)n1 + n2) return fail(KWG_ERR_INVALID_ARG, "n_dst1 must lie in [n1, n1 + n2]");
	const uint32_t take = n_dst1 - n1, n_dst2 = n2 - take;
	if (n_dst2 && !dst2) return fail(KWG_ERR_INVALID_ARG, "dst2 is NULL although source 2 does not fit into destination 1");
	if (n_slices == 0) return KWG_OK;
	int rc = select_device(device);
	if (rc) return rc;
	const uint64_t p1 = ceil_div(n1, 8), p2 = ceil_div(n2, 8), pd1 = ceil_div(n_dst1, 8), pd2 = ceil_div(n_dst2, 8);
	cudaStream_t st = nullptr;
	uint8_t* d = nullptr;
	KWG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	// pieces of at most ~256 MiB of device staging; sources and destinations of a piece share one allocation
	const uint64_t per_slice = p1 + p2 + pd1 + pd2 + 4;
	const uint64_t piece = std::max<uint64_t>(1, std::min<uint64_t>(n_slices, (256ull << 20) / per_slice));
	auto up16 = [](uint64_t v) { return (v + 15) & ~(uint64_t)15; };
	const uint64_t o1 = 0, o2 = o1 + up16(piece * p1 + 1), o3 = o2 + up16(piece * p2 + 1), o4 = o3 + up16(piece * pd1);
	const uint64_t total = o4 + up16(piece * pd2 + 1);
	cudaError_t e = cudaMalloc(&d, (size_t)total);
	if (e != cudaSuccess) { cudaStreamDestroy(st); return fail(KWG_ERR_NO_MEMORY, std::string("kwg_merge_slices: ") + cudaGetErrorString(e)); }
	rc = KWG_OK;
	for (uint64_t r0 = 0; r0 < n_slices && rc == KWG_OK; r0 += piece) {
		const uint64_t n = std::min(piece, n_slices - r0);
		MergeParams P{};
		P.src1 = d + o1; P.p1 = p1; P.n1 = n1;
		P.src2 = d + o2; P.p2 = p2; P.n2 = n2;
		P.n_slices = n; P.take = take;
		P.dst1 = d + o3; P.pd1 = pd1;
		P.dst2 = d + o4; P.pd2 = pd2;
		const uint64_t threads = n * (pd1 + pd2);
		if (cudaMemcpyAsync(d + o1, src1 + r0 * p1, (size_t)(n * p1), cudaMemcpyHostToDevice, st) != cudaSuccess ||
		    cudaMemcpyAsync(d + o2, src2 + r0 * p2, (size_t)(n * p2), cudaMemcpyHostToDevice, st) != cudaSuccess) {
			rc = fail(KWG_ERR_CUDA, "kwg_merge_slices: copy to the device failed");
			break;
		}
		merge_slices_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, st>>>(P);
		g_launches.fetch_add(1);
		if (cudaGetLastError() != cudaSuccess ||
		    cudaMemcpyAsync(dst1 + r0 * pd1, d + o3, (size_t)(n * pd1), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
		    (pd2 && cudaMemcpyAsync(dst2 + r0 * pd2, d + o4, (size_t)(n * pd2), cudaMemcpyDeviceToHost, st) != cudaSuccess) ||
		    cudaStreamSynchronize(st) != cudaSuccess) {
			rc = fail(KWG_ERR_CUDA, std::string("kwg_merge_slices: ") + cudaGetErrorString(cudaGetLastError()));
		}
	}
	cudaFree(d);
	cudaStreamDestroy(st);
	return rc;
}
