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

// Plants the k-mers of the first plant_len bases of every query_stride-th query into one column of a slice
// slab (sets bit `column` of row murmur3(kmer, h) & mask for h < NH): the benchmark's known positives.
__global__ void plant_kmers_kernel(uint8_t* __restrict__ slab, uint64_t row_pitch, uint32_t k, uint32_t num_hash, uint32_t mask,
	uint32_t column, const char* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t query_first, uint32_t query_stride,
	uint32_t plant_len)
{
	const uint32_t q = query_first + blockIdx.x * query_stride;
	const uint64_t o0 = offsets[q], o1 = offsets[q + 1];
	const uint64_t len = min((uint64_t)plant_len, o1 - o0);
	if (len < k) return;
	const char* s = bases + o0;
	for (uint64_t p = threadIdx.x; p + k <= len; p += blockDim.x) {
		uint64_t sense = 0;
		bool ok = true;
		for (uint32_t j = 0; j < k; ++j) {
			const uint32_t ch = (uint8_t)s[p + j], u = ch & 0xDFu;
			const uint32_t x = (ch >> 1) & 3u;
			ok = ok && (u == 'A' || u == 'C' || u == 'G' || u == 'T');
			sense = (sense << 2) | ((x ^ (x >> 1)) & 3u);
		}
		if (!ok) continue;
		const Canon c = canonical(sense, k);
		uint32_t h[KWG_MAX_NUM_HASH];
		murmur3_multi<KWG_MAX_NUM_HASH>(c.low, k, h);
		for (uint32_t t = 0; t < num_hash; ++t) {
			uint8_t* byte = slab + (uint64_t)(h[t] & mask) * row_pitch + column / 8;
			uint32_t* word = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(byte) & ~(uintptr_t)3);
			atomicOr(word, (1u << (column & 7u)) << (8 * (uint32_t)(reinterpret_cast<uintptr_t>(byte) & 3)));
		}
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

int kwg_synth_plant_dev(int device, uint8_t* d_slab, uint64_t row_pitch, uint32_t kmer_len, uint32_t num_hash, uint32_t log2_len,
	uint32_t column, const char* d_bases, const uint64_t* d_offsets, uint32_t query_first, uint32_t query_stride, uint32_t n_planted,
	uint32_t plant_len, void* stream)
{
	if (!d_slab || !d_bases || !d_offsets) return fail(KWG_ERR_INVALID_ARG, "NULL argument");
	if (kmer_len < 1 || kmer_len > KWG_MAX_KMER_LEN || num_hash < 1 || num_hash > KWG_MAX_NUM_HASH || log2_len > 32 || row_pitch % 4)
		return fail(KWG_ERR_INVALID_ARG, "bad planting parameters");
	if (n_planted == 0) return KWG_OK;
	int rc = select_device(device);
	if (rc) return rc;
	const uint32_t mask = (log2_len >= 32) ? 0xFFFFFFFFu : ((1u << log2_len) - 1u);
	plant_kmers_kernel<<<n_planted, 128, 0, (cudaStream_t)stream>>>(d_slab, row_pitch, kmer_len, num_hash, mask, column, d_bases, d_offsets,
		query_first, query_stride, plant_len);
	KWG_LAUNCHED();
	return KWG_OK;
}

} // extern "C"
