// Shared-memory atomic throughput vs table size and operation (design microbenchmark for bloom_first.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench4 ubench4.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// MODE 0 atomicOr+ret  1 atomicOr noret  2 atomicAdd+ret  3 atomicAdd noret  4 ld  5 ld + st (non-atomic)  6 atomicOr+ret only 1 lane in 8 active
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t words_mask, uint32_t per_thread, uint32_t* sink)
{
	extern __shared__ uint32_t tab[];
	for (uint32_t i = threadIdx.x; i <= words_mask; i += blockDim.x) tab[i] = 0;
	__syncthreads();
	uint32_t acc = 0, s = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B9u + 99u;
#pragma unroll 4
	for (uint32_t i = 0; i < per_thread; ++i) {
		s = mix(s + i);
		const uint32_t a = s & words_mask, bit = 1u << (s >> 27);
		if (MODE == 0) acc += atomicOr(tab + a, bit);
		else if (MODE == 1) atomicOr(tab + a, bit);
		else if (MODE == 2) acc += atomicAdd(tab + a, 1u);
		else if (MODE == 3) atomicAdd(tab + a, 1u);
		else if (MODE == 4) acc += tab[a];
		else if (MODE == 5) { const uint32_t v = tab[a]; tab[a] = v | bit; acc += v; }
		else if (MODE == 6) { if ((s & 0x700) == 0) acc += atomicOr(tab + a, bit); }
	}
	if (acc == 0x12345678u) *sink = acc;
}

template <typename F> static float time_ms(F f)
{
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	f(); CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
	return best;
}

int main()
{
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	uint32_t* sink; CK(cudaMalloc(&sink, 4));
	const int blocks = prop.multiProcessorCount; const uint32_t per_thread = 2048;
	const double ops = (double)blocks * 1024 * per_thread;
	const char* names[7] = {"or+ret", "or", "add+ret", "add", "ld", "ld+st", "or+ret 1/8 lanes"};
#define RUN(M) { CK(cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
	const float t = time_ms([&] { k<M><<<blocks, 1024, bytes>>>(mask, per_thread, sink); }); \
	printf("  %s %.3g G/s (%.2f /clk/SM @1.9GHz)", names[M], ops / t / 1e6 * (M == 6 ? 0.125 : 1.0), ops / t / 1e6 * (M == 6 ? 0.125 : 1.0) / blocks / 1.9); }
	for (int lg = 9; lg <= 15; lg += 2) {
		const uint32_t mask = (1u << lg) - 1; const size_t bytes = (size_t)4 << lg;
		printf("table %6zu B:", bytes);
		RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
		printf("\n");
	}
	return 0;
}
