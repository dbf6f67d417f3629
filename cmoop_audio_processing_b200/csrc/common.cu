#include "common.cuh"

#include <stdarg.h>

#include <atomic>

namespace cmoop {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::atomic<uint64_t> g_h2d{0}, g_d2h{0};
void count_copy(size_t bytes, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyHostToDevice) g_h2d.fetch_add((uint64_t)bytes, std::memory_order_relaxed);
    else if (kind == cudaMemcpyDeviceToHost) g_d2h.fetch_add((uint64_t)bytes, std::memory_order_relaxed);
}

static const int kSlots = 16;
static void* g_dev[kSlots];
static size_t g_dev_bytes[kSlots];
static void* g_pin[kSlots];
static size_t g_pin_bytes[kSlots];
static cudaStream_t g_stream = nullptr;

bool ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no usable CUDA device (%s); libcmoop_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return false;
    }
    return true;
}

cudaStream_t internal_stream() {
    if (!g_stream) {
        if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) g_stream = nullptr;
    }
    return g_stream;
}

void* device_scratch(int slot, size_t bytes) {
    if (slot < 0 || slot >= kSlots) return nullptr;
    if (bytes <= g_dev_bytes[slot]) return g_dev[slot];
    if (g_dev[slot]) {
        cudaDeviceSynchronize();
        cudaFree(g_dev[slot]);
        g_dev[slot] = nullptr;
        g_dev_bytes[slot] = 0;
    }
    size_t want = align_up(bytes + bytes / 4, 256);
    if (cudaMalloc(&g_dev[slot], want) != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed for scratch slot %d", want, slot);
        g_dev[slot] = nullptr;
        return nullptr;
    }
    g_dev_bytes[slot] = want;
    return g_dev[slot];
}

void* pinned_scratch(int slot, size_t bytes) {
    if (slot < 0 || slot >= kSlots) return nullptr;
    if (bytes <= g_pin_bytes[slot]) return g_pin[slot];
    if (g_pin[slot]) {
        cudaDeviceSynchronize();
        cudaFreeHost(g_pin[slot]);
        g_pin[slot] = nullptr;
        g_pin_bytes[slot] = 0;
    }
    size_t want = align_up(bytes + bytes / 4, 256);
    if (cudaMallocHost(&g_pin[slot], want) != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed for scratch slot %d", want, slot);
        g_pin[slot] = nullptr;
        return nullptr;
    }
    g_pin_bytes[slot] = want;
    return g_pin[slot];
}

}  // namespace cmoop

extern "C" {

int cmoop_abi_version(void) { return 1; }

const char* cmoop_last_error(void) { return cmoop::g_err; }

int cmoop_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int cmoop_set_device(int device) {
    CMOOP_CUDA_OK(cudaSetDevice(device));
    return CMOOP_OK;
}

uint64_t cmoop_launch_count(void) { return cmoop::g_launches.load(); }

void cmoop_copy_bytes(uint64_t* h2d, uint64_t* d2h) {
    if (h2d) *h2d = cmoop::g_h2d.load();
    if (d2h) *d2h = cmoop::g_d2h.load();
}

}  // extern "C"
