// Shared host-side helpers for the C-ABI library: error string, CUDA error mapping.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

namespace ady {

enum : int {
    ADY_OK = 0,
    ADY_ERR_INVALID = -1,      // bad argument / unsupported configuration
    ADY_ERR_CUDA = -2,         // CUDA runtime error (launch, attribute, allocation at init)
    ADY_ERR_UNSUPPORTED = -3,  // geometry not compiled into this library
};

char* last_error_buf();
// kernels launched by this library in this process (adyolo_launch_count; bench.py's gpu_launches)
std::atomic<long long>& launch_counter();
int set_error(int code, const char* fmt, ...);

#define ADY_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::ady::set_error(::ady::ADY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,       \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define ADY_LAUNCH_CHECK(name)                                                                 \
    do {                                                                                       \
        ::ady::launch_counter().fetch_add(1, std::memory_order_relaxed);                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return ::ady::set_error(::ady::ADY_ERR_CUDA, "launch of %s failed: %s", name,      \
                                    cudaGetErrorString(_e));                                   \
    } while (0)

}  // namespace ady
