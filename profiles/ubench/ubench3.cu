// Design microbenchmark 3 (not part of the product): random gathers of contiguous pieces from a 64 GiB slab,
// the access pattern of search_count_kernel (one row segment per warp-level load).  What does HBM deliver as a
// function of the piece size, and does it help if two warps of a block fetch the two halves of a 1 KiB row?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }

// each warp-level load fetches 512 B (32 lanes x 16 B); a "piece" of `piece` bytes is fetched by piece/512
// warps of the same block at the same loop step (piece >= 512) or a 512 B load covers 512/piece pieces (< 512)
template <int U>
__global__ void __launch_bounds__(256) gather(const uint8_t* slab, uint64_t n_rows, uint32_t row_bytes, uint32_t piece, uint32_t iters, uint32_t* sink)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t wpp = piece >= 512 ? piece / 512 : 1;             // warps per piece
	const uint32_t group = warp / wpp, part = warp % wpp;
	const uint32_t lanes_per_piece = piece >= 512 ? 32 : piece / 16;
	uint32_t acc = 0;
	uint32_t seed = (blockIdx.x * 8 + group) * 0x9E3779B9u + 17u;
	for (uint32_t it = 0; it < iters; ++it) {
		uint4 v[U];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			seed = mix(seed + u + it * 131u);
			const uint32_t sub = lane / lanes_per_piece;                // several small pieces per warp load
			const uint64_t row = (uint64_t)mix(seed + sub * 7919u) % n_rows;
			const uint8_t* p = slab + row * row_bytes + (uint64_t)part * 512 + (uint64_t)(lane % lanes_per_piece) * 16;
			v[u] = ld_nc_v4(p);
		}
#pragma unroll
		for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].w;
	}
	if (acc == 0x12345u) *sink = acc;
}

int main()
{
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	const uint64_t bytes = 64ull << 30;
	uint8_t* slab; uint32_t* sink;
	CK(cudaMalloc(&slab, bytes)); CK(cudaMalloc(&sink, 4));
	CK(cudaMemset(slab, 1, bytes));
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	const uint32_t row_bytes = 2048;
	const uint64_t n_rows = bytes / row_bytes;
	for (uint32_t blocks_per_sm : {2u, 4u, 8u}) {
		for (uint32_t piece : {128u, 256u, 512u, 1024u, 2048u}) {
			const uint32_t iters = 256;
			const int grid = prop.multiProcessorCount * blocks_per_sm;
			float best = 1e30f;
			for (int r = 0; r < 3; ++r) {
				CK(cudaEventRecord(a));
				gather<12><<<grid, 256>>>(slab, n_rows, row_bytes, piece, iters, sink);
				CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
				float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
			}
			const double total = (double)grid * 8 * iters * 12 * 512;
			printf("%u blocks/SM x 8 warps x 12 loads in flight, pieces of %4u B: %.3f ms  %.0f GB/s\n", blocks_per_sm, piece, best, total / best / 1e6);
		}
	}
	return 0;
}
