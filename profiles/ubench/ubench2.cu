// Design microbenchmark 2 (not part of the product): gather of short runs at a fixed stride.
// Each warp reads `piece` bytes (8 B per lane per step) from address  base + r * stride + off(block),
// r = 0..255 -- the access pattern of regroup_kernel's gather (tile windows 64 KiB apart).
// Question: does the exact power-of-two stride cost bandwidth (DRAM channel / L2 slice camping)?
// Also: random red.add into a 73 MB array while a 3.8 GB stream passes through L2, with and without
// an L2 persisting window for the array.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) gather(const uint64_t* base, uint64_t stride_rec, uint32_t n_windows, uint32_t units, uint32_t recs, unsigned long long* sink)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long acc = 0;
	for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
		const uint32_t g = u / 256, i = u % 256;                  // group of 256 windows, bucket i
		const uint64_t* p0 = base + (uint64_t)g * 256 * stride_rec + (uint64_t)i * recs + (i * 7u) % 5u;
		for (uint32_t r0 = warp; r0 < 256; r0 += 8 * 8) {
			uint64_t v[8];
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				const uint32_t r = r0 + q * 8;
				v[q] = (r < 256 && lane < recs) ? p0[(uint64_t)r * stride_rec + lane] : 0;
			}
#pragma unroll
			for (int q = 0; q < 8; ++q) acc += v[q];
		}
	}
	if (acc == 0x1234567ull) *sink = acc;
}

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// stream `n` 16-byte words (reads) and do one random red.add into `tab` per `every` words
__global__ void __launch_bounds__(256) stream_and_red(const uint4* src, uint64_t n, uint32_t* tab, uint32_t tab_words, uint32_t every, unsigned long long* sink)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		uint4 v;
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
		acc += v.x + v.w;
		if (i % every == 0) atomicAdd(tab + (mix((uint32_t)i * 2654435761u) % tab_words), 1u << (v.x & 28));
	}
	if (acc == 0x1234567ull) *sink = acc;
}

template <typename F> static float time_ms(F f, int reps = 3)
{
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	f(); CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
	return best;
}

int main()
{
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	printf("device %s, %d SMs, L2 %d MB, persisting L2 max %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20);
	const uint32_t n_windows = 73216;                      // 286 groups of 256
	const uint64_t max_stride = 8192 + 512;
	uint64_t* buf; unsigned long long* sink;
	CK(cudaMalloc(&buf, (size_t)n_windows * max_stride * 8 + (1 << 20))); CK(cudaMalloc(&sink, 8));
	CK(cudaMemset(buf, 1, (size_t)n_windows * max_stride * 8));
	const uint32_t units = n_windows;                      // (g, i) pairs: 286 * 256
	for (uint32_t recs : {26u, 32u}) {
		for (uint64_t stride : {8192ull, 8192ull + 2, 8192ull + 16, 8192ull + 32, 8192ull + 64, 8192ull + 130, 8192ull + 256, 8192ull + 512}) {
			const float t = time_ms([&] { gather<<<prop.multiProcessorCount * 8, 256>>>(buf, stride, n_windows, units, recs, sink); });
			printf("gather %u recs/run, window stride %llu records (%llu B): %.3f ms  %.0f GB/s\n", recs, (unsigned long long)stride,
			       (unsigned long long)stride * 8, t, (double)units * 256 * recs * 8 / t / 1e6);
		}
	}
	// loss-counter pattern
	const uint64_t n16 = (3840ull << 20) / 16;
	uint4* src; uint32_t* tab;
	const uint32_t tab_words = 73u << 18;                 // 73 MB
	CK(cudaMalloc(&src, n16 * 16)); CK(cudaMalloc(&tab, (size_t)tab_words * 4));
	CK(cudaMemset(src, 3, n16 * 16)); CK(cudaMemset(tab, 0, (size_t)tab_words * 4));
	cudaStream_t st; CK(cudaStreamCreate(&st));
	for (int persist = 0; persist < 2; ++persist) {
		if (persist) {
			CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize));
			cudaStreamAttrValue av = {};
			av.accessPolicyWindow.base_ptr = tab;
			av.accessPolicyWindow.num_bytes = (size_t)tab_words * 4;
			av.accessPolicyWindow.hitRatio = 1.0f;
			av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
		}
		for (uint32_t every : {1000000000u, 5u}) {        // none, one red per 80 bytes streamed (= 1 per 10 records)
			cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
			float best = 1e30f;
			for (int r = 0; r < 4; ++r) {
				CK(cudaEventRecord(a, st));
				stream_and_red<<<prop.multiProcessorCount * 8, 256, 0, st>>>(src, n16, tab, tab_words, every, sink);
				CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b));
				float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (r && ms < best) best = ms;
			}
			printf("stream 3.84 GB %s red.add every %u words into 73 MB, persisting window %s: %.3f ms (%.0f GB/s stream, %.1f G red/s)\n",
			       every > 1000 ? "without" : "with", every, persist ? "ON" : "off", best, n16 * 16.0 / best / 1e6, every > 1000 ? 0.0 : n16 / (double)every / best / 1e6);
		}
	}
	return 0;
}
