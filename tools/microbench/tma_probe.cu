// Probe of the tiled-TMA behaviour the convolution kernels rely on (run once, results quoted in DESIGN.md):
//  (1) 4-D tensor map {C, W, H, N} over a dense NHWC bf16 tensor, box {64, Wp, 1, 1} at coordinates (c0, -p, h - p, n):
//      out-of-bounds elements (w < 0, w >= W, h outside [0, H), c >= C) are zero-filled -> the box IS a padded image row;
//  (2) SWIZZLE_128B: is the 16-byte-chunk XOR a function of the ABSOLUTE shared-memory address (bits 7..9) or of the row
//      index inside the box?  A destination that is 128-byte but not 1024-byte aligned tells them apart.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const CUtensorMap* __restrict__ tmap, int c0, int w0, int h0, int n0, int dst_row_off, int rows,
                      uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    for (int i = threadIdx.x; i < 64 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0xDEADDEADu;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)rows * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(base + dst_row_off * 128)), "l"(tmap), "r"(c0), "r"(w0), "r"(h0), "r"(n0), "r"(smem_u32(&bar))
            : "memory");
    }
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        if (++spins > (1u << 24)) __trap();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(base)[i];
}

int main() {
    const int N = 3, H = 5, W = 6, C = 32, p = 1, Wp = W + 2 * p;
    std::vector<__nv_bfloat16> h((size_t)N * H * W * C);
    // value encodes (n, h, w, c): exactly representable small integers: 1 + c + 32*(w + 8*(h + 8*n)) would overflow bf16
    // precision, so use two probes: value = c + 1 for the channel/chunk order, and value = 1 + w + 8*h + 64*n for position
    for (int mode = 0; mode < 2; ++mode) {
        for (int n = 0; n < N; ++n)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    for (int c = 0; c < C; ++c)
                        h[(((size_t)n * H + y) * W + x) * C + c] = __float2bfloat16(mode == 0 ? (float)(c + 1) : (float)(1 + x + 8 * y + 64 * n));
        __nv_bfloat16* d;
        cudaMalloc(&d, h.size() * 2);
        cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMap tm;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)Wp, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        cuInit(0);
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("mode %d: encode -> %d\n", mode, (int)r);
        CUtensorMap* dtm;
        cudaMalloc(&dtm, sizeof(tm));
        cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
        uint16_t* dout;
        cudaMalloc(&dout, 64 * 64 * 2);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
        for (int off : {0, 3}) {
            for (int hrow : {0, -1, 2}) {
                probe<<<1, 128, 16 * 1024>>>(dtm, 0, -p, hrow, 1, off, Wp, dout);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<uint16_t> o(64 * 64);
                cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
                printf("mode %d dst_row_off %d h %d (n = 1): per smem row [first element of each 16-byte chunk]\n", mode, off, hrow);
                for (int row = 0; row < off + Wp + 1; ++row) {
                    printf("  row %2d:", row);
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t bits = (uint32_t)o[row * 64 + ch * 8] << 16;
                        float f;
                        memcpy(&f, &bits, 4);
                        if (o[row * 64 + ch * 8] == 0xDEAD) printf("   ----"); else printf(" %6.0f", f);
                    }
                    printf("\n");
                }
            }
        }
        cudaFree(d); cudaFree(dtm); cudaFree(dout);
    }
    return 0;
}
