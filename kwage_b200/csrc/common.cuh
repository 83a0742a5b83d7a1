// Shared device helpers and host-side error plumbing for libkwage_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/kwage_cuda.h"
#include "bitops.cuh"

namespace kwg {

// ---------------------------------------------------------------- host: errors and launch count
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define KWG_CUDA(expr)                                                                              \
	do {                                                                                            \
		cudaError_t _e = (expr);                                                                    \
		if (_e != cudaSuccess) {                                                                    \
			return ::kwg::fail(_e == cudaErrorMemoryAllocation ? KWG_ERR_NO_MEMORY : KWG_ERR_CUDA,  \
				std::string(#expr) + ": " + cudaGetErrorString(_e));                                \
		}                                                                                           \
	} while (0)

#define KWG_LAUNCHED()                                   \
	do {                                                 \
		::kwg::g_launches.fetch_add(1);                  \
		KWG_CUDA(cudaGetLastError());                    \
	} while (0)

static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static inline uint64_t round_up(uint64_t a, uint64_t b) { return ceil_div(a, b) * b; }

// Optional per-handle kernel timing: CUDA events recorded on the handle's own stream around each
// kernel launch, summed per kernel id when the caller asks (kwg_*_get_timing).
struct KernelTimers {
	struct Span { int id; cudaEvent_t a, b; };
	bool enabled = false;
	std::vector<Span> spans;
	void begin(int id, cudaStream_t s)
	{
		if (!enabled) return;
		Span sp{id, nullptr, nullptr};
		if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
		cudaEventRecord(sp.a, s);
		spans.push_back(sp);
	}
	void end(cudaStream_t s)
	{
		if (!enabled || spans.empty()) return;
		cudaEventRecord(spans.back().b, s);
	}
	// caller has synchronised the stream; adds elapsed ms into ms[id] and the launch count into n[id]
	void collect(double* ms, uint64_t* n, int n_ids)
	{
		for (Span& sp : spans) {
			float t = 0.f;
			if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess && sp.id >= 0 && sp.id < n_ids) { ms[sp.id] += t; n[sp.id] += 1; }
			cudaEventDestroy(sp.a);
			cudaEventDestroy(sp.b);
		}
		spans.clear();
	}
};

int select_device(int device);   // cudaSetDevice with validation
int sm_count(int device);

// crc32.cu: zlib crc32 of n_seg equal-length messages resident in HBM, queued on `stream` (no synchronisation).
// Message s = n_rows rows of row_bytes at d_base + s * seg_stride + r * row_pitch; d_crc_in (NULL: 0) / d_crc_out: [n_seg].
size_t crc32_workspace_words(uint64_t n_seg, uint64_t message_bytes);
int crc32_launch(int device, const uint8_t* d_base, uint64_t n_seg, uint64_t seg_stride, uint64_t n_rows, uint64_t row_bytes,
	uint64_t row_pitch, const uint32_t* d_crc_in, uint32_t* d_crc_out, uint32_t* d_ws, cudaStream_t stream);

// ---------------------------------------------------------------- device: memory access helpers
__device__ __forceinline__ uint4 ld_nc_v4(const void* p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ uint32_t ld_nc_u32(const void* p)
{
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}
// L2 eviction policies (createpolicy + .L2::cache_hint): a one-pass stream must not push the working set
// (a filter that is being filled with red.or) out of the L2, and the working set asks to stay.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t ld_nc_u64_hint(const void* p, uint64_t pol)
{
	uint64_t r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
	return r;
}
__device__ __forceinline__ uint4 ld_nc_v4_hint(const void* p, uint64_t pol)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
	return r;
}
__device__ __forceinline__ void red_or_hint(uint32_t* p, uint32_t v, uint64_t pol)
{
	asm volatile("red.global.or.b32.L2::cache_hint [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}

// ---------------------------------------------------------------- device: the bases of a batch
// ASCII (one byte per base, what NGS hands the reference: make_bloom.cpp:201-203) or packed: NCBI 2na, four bases per
// byte, first base in bits 7..6, A=0 C=1 G=2 T=3 (the reference's own code order, word.h:19) plus an optional
// one-bit-per-base mask (LSB first) of the positions that do not hold ACGT.
struct BaseSource {
	const char* bases;          // 16-byte aligned
	const uint16_t* bad_mask;   // packed only; NULL: every base is one of ACGT
	uint64_t n_bases;
	uint32_t packed;
};

// 16 bases starting at base g (a multiple of 16): 2-bit codes, first base in the top bits, and one "not a base" flag
// per base.  Positions at and beyond n_bases read as separators.
__device__ __forceinline__ void load_group16(const BaseSource& S, uint64_t g, uint32_t& codes, uint32_t& bad16)
{
	codes = 0; bad16 = 0xFFFFu;
	if (g >= S.n_bases) return;
	if (S.packed) {
		codes = __byte_perm(ld_nc_u32(S.bases + (g >> 2)), 0u, 0x0123);
		bad16 = S.bad_mask ? (uint32_t)__ldg(S.bad_mask + (g >> 4)) : 0u;
		if (g + 16 > S.n_bases) bad16 |= (0xFFFFu << (uint32_t)(S.n_bases - g)) & 0xFFFFu;
	} else if (g + 16 <= S.n_bases) {
		encode16(ld_nc_v4(S.bases + g), codes, bad16);
	} else {
		uint32_t w[4] = {0, 0, 0, 0};
		for (uint32_t j = 0; j < 16; ++j) {
			const uint32_t c = (g + j < S.n_bases) ? (uint8_t)S.bases[g + j] : (uint32_t)'N';
			w[j >> 2] |= c << (8 * (j & 3));
		}
		encode16(make_uint4(w[0], w[1], w[2], w[3]), codes, bad16);
	}
}

__device__ __forceinline__ void st_na_v4(void* p, uint4 v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
	             :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_na_u32(void* p, uint32_t v)
{
	asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

} // namespace kwg
