// Design microbenchmarks (not part of the product): throughput of the primitives the bucketed
// first-touch construction leans on.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}

// MODE 0: atomicMin with return, 1: red.min, 2: ld, 3: red.or bit, 4: atomicAdd nibble (10% of ops)
template <int MODE>
__global__ void __launch_bounds__(512) global_ops(uint32_t* tab, uint32_t mask, uint32_t per_thread, uint32_t* sink)
{
	const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t acc = 0;
	uint32_t s = gid * 0x9E3779B9u + 12345u;
#pragma unroll 4
	for (uint32_t i = 0; i < per_thread; ++i) {
		s = mix(s + i);
		const uint32_t a = s & mask;
		const uint32_t v = s >> 4;
		if (MODE == 0) acc += atomicMin(tab + a, v);
		else if (MODE == 1) atomicMin(tab + a, v);
		else if (MODE == 2) acc += __ldcg(tab + a);
		else if (MODE == 3) atomicOr(tab + a, 1u << (v & 31));
		else if (MODE == 4) { if ((v & 15) == 0) atomicAdd(tab + a, 1u << (v & 28)); }
	}
	if (acc == 0x12345678u) *sink = acc;
}

// shared-memory atomics: MODE 0 atomicAdd with return on `nbins` counters, 1 without return,
// 2: match_any-based ranking (warp-private histogram, leader updates), 3: ballot-based (9 ballots)
template <int MODE>
__global__ void __launch_bounds__(256) smem_ops(uint32_t nbins_mask, uint32_t per_thread, uint32_t* sink)
{
	__shared__ uint32_t hist[8 * 512];
	for (uint32_t i = threadIdx.x; i < 8 * 512; i += blockDim.x) hist[i] = 0;
	__syncthreads();
	const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t acc = 0;
	uint32_t s = gid * 0x9E3779B9u + 777u;
#pragma unroll 4
	for (uint32_t i = 0; i < per_thread; ++i) {
		s = mix(s + i);
		const uint32_t d = s & nbins_mask;
		if (MODE == 0) acc += atomicAdd(hist + d, 1u);
		else if (MODE == 1) atomicAdd(hist + d, 1u);
		else if (MODE == 2) {
			const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
			const uint32_t leader = 31 - __clz(peers);
			uint32_t pre = 0;
			uint32_t* hp = hist + warp * 512 + d;
			if (lane == leader) { pre = *hp; *hp = pre + __popc(peers); }
			pre = __shfl_sync(0xFFFFFFFFu, pre, leader);
			acc += pre + __popc(peers & ((1u << lane) - 1u));
		} else {
			uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
			for (int b = 0; b < 9; ++b) {
				const uint32_t m = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
				peers &= ((d >> b) & 1u) ? m : ~m;
			}
			const uint32_t leader = 31 - __clz(peers);
			uint32_t pre = 0;
			uint32_t* hp = hist + warp * 512 + d;
			if (lane == leader) { pre = *hp; *hp = pre + __popc(peers); }
			pre = __shfl_sync(0xFFFFFFFFu, pre, leader);
			acc += pre + __popc(peers & ((1u << lane) - 1u));
		}
	}
	__syncthreads();
	if (acc == 0x12345678u) *sink = acc + hist[threadIdx.x];
}

template <typename F>
static float time_ms(F f, int reps = 3)
{
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	f();
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < reps; ++r) {
		CK(cudaEventRecord(a));
		f();
		CK(cudaEventRecord(b));
		CK(cudaEventSynchronize(b));
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		if (ms < best) best = ms;
	}
	return best;
}

int main()
{
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	printf("device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
	uint32_t* tab; uint32_t* sink;
	const size_t max_bytes = 1ull << 33;
	CK(cudaMalloc(&tab, max_bytes)); CK(cudaMalloc(&sink, 4));
	CK(cudaMemset(tab, 0xFF, max_bytes));
	const int blocks = prop.multiProcessorCount * 4, threads = 512;
	const uint32_t per_thread = 1024;
	const double ops = (double)blocks * threads * per_thread;
	const char* names[5] = {"atomicMin+ret", "red.min", "ld.cg", "red.or", "red.add 1/16"};
	for (int lg = 22; lg <= 33; ++lg) {
		if (lg > 27 && lg != 28 && lg != 30 && lg != 33) continue;
		const uint32_t mask = (uint32_t)((1ull << (lg - 2)) - 1);
		float t[5];
		t[0] = time_ms([&] { global_ops<0><<<blocks, threads>>>(tab, mask, per_thread, sink); });
		t[1] = time_ms([&] { global_ops<1><<<blocks, threads>>>(tab, mask, per_thread, sink); });
		t[2] = time_ms([&] { global_ops<2><<<blocks, threads>>>(tab, mask, per_thread, sink); });
		t[3] = time_ms([&] { global_ops<3><<<blocks, threads>>>(tab, mask, per_thread, sink); });
		t[4] = time_ms([&] { global_ops<4><<<blocks, threads>>>(tab, mask, per_thread, sink); });
		printf("region 2^%d B (%6.0f MB):", lg, (double)(1ull << lg) / 1048576.0);
		for (int m = 0; m < 5; ++m) printf("  %s %.3g G/s", names[m], ops / t[m] / 1e6 * (m == 4 ? 1.0 / 16 : 1.0));
		printf("\n");
	}
	const int sblocks = prop.multiProcessorCount * 8;
	const double sops = (double)sblocks * 256 * 4096;
	const char* snames[4] = {"ATOMS+ret", "ATOMS noret", "match_any rank", "9-ballot rank"};
	for (int nb = 8; nb <= 9; ++nb) {
		float t[4];
		t[0] = time_ms([&] { smem_ops<0><<<sblocks, 256>>>((1u << nb) - 1, 4096, sink); });
		t[1] = time_ms([&] { smem_ops<1><<<sblocks, 256>>>((1u << nb) - 1, 4096, sink); });
		t[2] = time_ms([&] { smem_ops<2><<<sblocks, 256>>>((1u << nb) - 1, 4096, sink); });
		t[3] = time_ms([&] { smem_ops<3><<<sblocks, 256>>>((1u << nb) - 1, 4096, sink); });
		printf("smem %d bins:", 1 << nb);
		for (int m = 0; m < 4; ++m) printf("  %s %.3g G/s", snames[m], sops / t[m] / 1e6);
		printf("\n");
	}
	return 0;
}
