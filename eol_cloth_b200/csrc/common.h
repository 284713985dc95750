// common.h — context, error plumbing and small RAII helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/eolc.h"

namespace eolc {

void set_error(const char *fmt, ...);

#define EOLC_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            eolc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(err__)); \
            return EOLC_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

#define EOLC_REQUIRE(cond, msg)                        \
    do {                                               \
        if (!(cond)) {                                 \
            eolc::set_error("%s (%s)", msg, #cond);    \
            return EOLC_ERR_ARG;                       \
        }                                              \
    } while (0)

// device buffer that frees itself
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;   // a failed allocation leaves an empty buffer: a retry allocates again
        return e;
    }
    cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMallocHost((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
};

}  // namespace eolc

struct eolc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // second copy engine for the big device-to-host copies of the host entry points
    cudaEvent_t copy_event = nullptr;
};
