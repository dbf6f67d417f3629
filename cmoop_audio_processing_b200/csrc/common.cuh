// Shared host-side plumbing for libcmoop_b200: error reporting, launch counter,
// library-owned device scratch for the *_host entry points.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cmoop_b200.h"

namespace cmoop {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// host<->device traffic of every entry point (bench.py e2e "h2d_bytes_per_step" / "d2h_bytes_per_step"): all copies of the
// library go through these two wrappers
void count_copy(size_t bytes, cudaMemcpyKind kind);
inline cudaError_t copy_async(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    count_copy(bytes, kind);
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}
// cudaMemcpy from PAGEABLE host memory returns once the source has been staged -- "the DMA to final destination may not
// have completed" (CUDA runtime API, memcpy semantics).  Work on the legacy stream is ordered behind it, but this library
// launches on a NON-BLOCKING stream that does not synchronise with the legacy stream: a kernel launched right after such a
// copy read half-written task records (flaky illegal addresses, found with tools/calibrate_cost.py).  Host-to-device
// copies therefore wait for the legacy stream before returning.
inline cudaError_t copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    count_copy(bytes, kind);
    cudaError_t e = cudaMemcpy(dst, src, bytes, kind);
    if (e == cudaSuccess && kind == cudaMemcpyHostToDevice) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}

// Grow-only device/pinned scratch owned by the library (one per slot).
int bound_device();                 // device this process is bound to (-1 before the first device entry point)
void* device_scratch(int slot, size_t bytes);
void* pinned_scratch(int slot, size_t bytes);
cudaStream_t internal_stream();
bool ensure_device();   // false (and error set) when no CUDA device is usable

#define CMOOP_CUDA_OK(expr)                                                                 \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cmoop::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return CMOOP_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

#define CMOOP_REQUIRE(cond, ...)                \
    do {                                        \
        if (!(cond)) {                          \
            cmoop::set_error(__VA_ARGS__);      \
            return CMOOP_ERR_INVALID;           \
        }                                       \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace cmoop
