// Internal helpers: error reporting, CUDA checks, launch accounting, device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <string>
#include "../../include/adypt_b200.h"
#include "guard.h"

#ifndef ADYPT_NO_FMAD
#error "build with -fmad=false -DADYPT_NO_FMAD: the FP policy forbids implicit multiply-add contraction"
#endif

namespace adypt {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
extern std::atomic<uint64_t> g_launches;

#define ADYPT_CUDA(expr)                                                                                  \
	do {                                                                                                  \
		cudaError_t _e = (expr);                                                                          \
		if (_e != cudaSuccess)                                                                            \
			return ::adypt::fail(_e == cudaErrorMemoryAllocation ? ADYPT_ENOMEM : ADYPT_ECUDA,            \
			                     std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
	} while (0)

#define ADYPT_TRY(expr)            \
	do {                           \
		int _rc = (expr);          \
		if (_rc != ADYPT_OK) return _rc; \
	} while (0)

// a growable device buffer (never shrinks)
struct DeviceBuffer {
	void *ptr = nullptr;
	size_t cap = 0;
	int reserve(size_t bytes)
	{
		if (bytes <= cap) return ADYPT_OK;
		if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
		size_t want = bytes + bytes / 8;
		ADYPT_CUDA(cudaMalloc(&ptr, want));
		cap = want;
		return ADYPT_OK;
	}
	void release()
	{
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
	}
	template <class T> T *as() const { return (T *)ptr; }
};

struct DeviceGuard {
	int prev = -1;
	bool ok = false;
	explicit DeviceGuard(int dev)
	{
		if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
		ok = cudaSetDevice(dev) == cudaSuccess;
		// A failed call made earlier by anyone in this process (ours, torch's, ...) leaves its code in the runtime's
		// "last error" slot; the cudaGetLastError() after our next kernel launch would report it as ours. Start clean.
		(void)cudaGetLastError();
	}
	~DeviceGuard()
	{
		if (prev >= 0) cudaSetDevice(prev);
	}
};

} // namespace adypt
