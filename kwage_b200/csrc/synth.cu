// Device-side synthetic inputs for benchmarks: the same counter-based generators as
// oracle/kwage_oracle.c (kwo_gen_reads, kwo_gen_filter_bits), so that HBM-resident benchmark
// inputs can be re-created bit-for-bit on the CPU for parity checks.
#include "common.cuh"

namespace kwg {

__global__ void synth_reads_kernel(uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t read_len, uint32_t groups,
	char* __restrict__ bases, uint64_t* __restrict__ offsets)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (offsets && t <= n_reads) offsets[t] = t * read_len;
	if (t >= n_reads * groups) return;
	const uint64_t r = t / groups;
	const uint32_t g = (uint32_t)(t % groups);
	uint64_t w = synth_rnd(seed, first_read + r, g);
	char* out = bases + r * read_len + (uint64_t)g * 32;
	const uint32_t n = min(32u, read_len - g * 32);
	for (uint32_t j = 0; j < n; ++j, w >>= 2) out[j] = "ACGT"[w & 3];
}

__global__ void synth_filter_bits_kernel(uint64_t seed, uint64_t first_filter, uint32_t n_filters, uint64_t words_per_filter,
	uint64_t filter_pitch, uint8_t* __restrict__ filters)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t total = (uint64_t)n_filters * words_per_filter;
	for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
		const uint64_t j = t / words_per_filter, w = t % words_per_filter;
		const uint64_t v = synth_rnd(seed, first_filter + j, 2 * w) & synth_rnd(seed, first_filter + j, 2 * w + 1);
		*reinterpret_cast<uint64_t*>(filters + j * filter_pitch + w * 8) = v;
	}
}

} // namespace kwg

using namespace kwg;

extern "C" {

int kwg_synth_reads_dev(int device, uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t read_len,
	char* d_bases, uint64_t* d_offsets, void* stream)
{
	if (!d_bases) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (n_reads == 0 || read_len == 0) return KWG_OK;
	int rc = select_device(device);
	if (rc) return rc;
	const uint32_t groups = (read_len + 31) / 32;
	const uint64_t threads = std::max<uint64_t>(n_reads * groups, n_reads + 1);
	synth_reads_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(seed, first_read, n_reads, read_len, groups, d_bases, d_offsets);
	KWG_LAUNCHED();
	return KWG_OK;
}

int kwg_synth_filter_bits_dev(int device, uint64_t seed, uint64_t first_filter, uint32_t n_filters,
	uint64_t filter_bytes, uint64_t filter_pitch, uint8_t* d_filters, void* stream)
{
	if (!d_filters) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (filter_bytes % 8 || filter_pitch % 8 || filter_pitch < filter_bytes) return fail(KWG_ERR_INVALID_ARG, "filter_bytes/pitch must be multiples of 8");
	if (n_filters == 0 || filter_bytes == 0) return KWG_OK;
	int rc = select_device(device);
	if (rc) return rc;
	synth_filter_bits_kernel<<<sm_count(device) * 16, 256, 0, (cudaStream_t)stream>>>(seed, first_filter, n_filters, filter_bytes / 8, filter_pitch, d_filters);
	KWG_LAUNCHED();
	return KWG_OK;
}

} // extern "C"
