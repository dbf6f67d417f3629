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

// The library keeps process-global device state (scratch slots, the internal stream, the CNN activation arena, the lanes'
// task-list blobs): ONE device per process, as in the one-process-per-GPU launch (torchrun).  The first entry point that needs
// the device binds the current one; running on another device afterwards is refused instead of silently reusing pointers
// that belong to the first.
static int g_bound_device = -1;

bool ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no usable CUDA device (%s); libcmoop_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return false;
    }
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;                                    // reported by the first real CUDA call
    }
    if (g_bound_device < 0) g_bound_device = cur;
    if (cur != g_bound_device) {
        set_error("libcmoop_b200 is bound to device %d in this process (scratch, arena and streams live there) but the current "
                  "device is %d: run one process per GPU", g_bound_device, cur);
        return false;
    }
    return true;
}
int bound_device() { return g_bound_device; }

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
    if (cmoop::bound_device() >= 0 && device != cmoop::bound_device()) {
        cmoop::set_error("cmoop_set_device(%d): this process already runs on device %d (one process per GPU)", device,
                         cmoop::bound_device());
        return CMOOP_ERR_UNSUPPORTED;
    }
    CMOOP_CUDA_OK(cudaSetDevice(device));
    return CMOOP_OK;
}

uint64_t cmoop_launch_count(void) { return cmoop::g_launches.load(); }

void cmoop_copy_bytes(uint64_t* h2d, uint64_t* d2h) {
    if (h2d) *h2d = cmoop::g_h2d.load();
    if (d2h) *d2h = cmoop::g_d2h.load();
}

}  // extern "C"
