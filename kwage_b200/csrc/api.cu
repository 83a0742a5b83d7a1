// Error plumbing, device selection and misc entry points of libkwage_cuda.so.
#include "common.cuh"

#include <mutex>
#include <vector>

namespace kwg {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }

int fail(int code, const std::string& msg)
{
	t_last_error = msg;
	return code;
}

int select_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return fail(KWG_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
	if (device < 0 || device >= n) return fail(KWG_ERR_INVALID_ARG, "device index out of range");
	KWG_CUDA(cudaSetDevice(device));
	return KWG_OK;
}

int sm_count(int device)
{
	static std::mutex mu;
	static std::vector<int> cache;
	std::lock_guard<std::mutex> lock(mu);
	if ((int)cache.size() <= device) cache.resize(device + 1, 0);
	if (cache[device] == 0) {
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
		cache[device] = v;
	}
	return cache[device];
}

} // namespace kwg

extern "C" {

const char* kwg_last_error(void) { return kwg::t_last_error.c_str(); }

const char* kwg_version(void) { return "kwage-b200 0.1 (sm_100a)"; }

int kwg_device_count(int* count)
{
	if (!count) return kwg::fail(KWG_ERR_INVALID_ARG, "count is NULL");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		*count = 0;
		return kwg::fail(KWG_ERR_CUDA, cudaGetErrorString(e));
	}
	*count = n;
	return KWG_OK;
}

uint64_t kwg_launch_count(void) { return kwg::g_launches.load(); }

void* kwg_host_alloc(uint64_t bytes)
{
	void* p = nullptr;
	if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
		cudaGetLastError();
		kwg::set_error("kwg_host_alloc: cudaMallocHost failed");
		return nullptr;
	}
	return p;
}

void kwg_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

} // extern "C"
