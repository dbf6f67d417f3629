// Micro-benchmark that sizes the weight-gradient redesign: issue-to-completion cost of tcgen05.mma (kind::f16, bf16,
// M = 128, cta_group::1) per operand layout -- K-major vs MN-major SWIZZLE_128B for A and B, N = 64 / 128 / 256, one or
// two CTAs per SM, aligned and unaligned MN-major start rows.  Operand values do not matter (zeros); the descriptors are
// the ones conv_tc.cu / conv_tc2.cu build.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_layouts umma_layouts.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// a_mn / b_mn: operand is MN-major; row_shift: extra start offset in 128-byte rows applied to MN-major A (tap shift)
__global__ void __launch_bounds__(64) bench(int N, int a_mn, int b_mn, int row_shift, int stages, int iters, float* out,
                                            int lbo_rows, int n_acc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = 16384 + 2048, b_bytes = (N <= 128 ? 16384 : 32768);
    for (int i = threadIdx.x; i < stages * (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3F803F80u ^ (uint32_t)(i * 2654435761u & 0x007F007Fu);   // finite bf16 pairs
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 32) {
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (a_mn) idesc |= 1u << 15;
        if (b_mn) idesc |= 1u << 16;
        // lean issue loop: descriptors of stage 0 are built once; a stage / k step only adds to the 14-bit address field
        const uint32_t a0 = smem_u32(base), b0 = a0 + a_bytes;
        const uint64_t ad0 = a_mn ? desc_sw128(a0 + row_shift * 128, lbo_rows * 128) : desc_sw128(a0 + row_shift * 128, 16);
        const uint64_t bd0 = b_mn ? desc_sw128(b0, 8192) : desc_sw128(b0, 16);
        const uint64_t a_k = (uint64_t)((a_mn ? 2048 : 32) >> 4), b_k = (uint64_t)((b_mn ? 2048 : 32) >> 4);
        const uint64_t st_step = (uint64_t)((a_bytes + b_bytes) >> 4);
        const uint32_t acc_step = (uint32_t)N, acc_end = (uint32_t)(N * n_acc);
        const long long t0 = clock64();
        uint64_t ad = ad0, bd = bd0;
        int s = 0;
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                umma(tmem + acc, ad + a_k * k4, bd + b_k * k4, idesc, it != 0);
                acc += acc_step;
                if (acc >= acc_end) acc = 0;
            }
            ad += st_step; bd += st_step;
            if (++s == stages) { s = 0; ad = ad0; bd = bd0; }
        }
        commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = (float)(t1 - t0) / (float)(iters * 4);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

int main() {
    float* d;
    cudaMalloc(&d, 1024 * sizeof(float));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    printf("%-5s %-4s %-4s %-6s %-8s %-5s %-5s %-10s %-8s\n", "N", "Amn", "Bmn", "shift", "lboRows", "nacc", "ctas", "clk/UMMA", "mathclk");
    struct Cfg { int N, a_mn, b_mn, shift, lbo_rows, n_acc, ctas; };
    std::vector<Cfg> cfgs;
    for (int ctas : {1, 2})
        for (int N : {32, 64, 128, 256})
            for (int n_acc : {1, 4}) {
                if (n_acc * N > 256) continue;
                cfgs.push_back({N, 0, 0, 0, 64, n_acc, ctas});          // K-major both
                cfgs.push_back({N, 0, 0, 3, 64, n_acc, ctas});          // K-major, A start shifted by 3 rows (conv_tc2 taps)
                cfgs.push_back({N, 1, 1, 0, 64, n_acc, ctas});          // MN-major both, aligned, LBO 8192 (wgrad_tc)
                cfgs.push_back({N, 1, 1, 3, 64, n_acc, ctas});          // MN-major, unaligned start row
                cfgs.push_back({N, 1, 1, 0, 1, n_acc, ctas});           // MN-major, LBO = one row (adjacent taps, wgrad_tc2)
                cfgs.push_back({N, 1, 1, 3, 1, n_acc, ctas});
                cfgs.push_back({N, 1, 1, 0, 22, n_acc, ctas});          // LBO = Wp rows, not a multiple of 8
                cfgs.push_back({N, 1, 1, 3, 24, n_acc, ctas});          // LBO = 24 rows (multiple of 8), unaligned start
                cfgs.push_back({N, 1, 0, 3, 24, n_acc, ctas});          // only A MN-major
            }
    for (const Cfg& c : cfgs) {
        const int stages = c.ctas == 1 ? 3 : 2;
        const size_t smem = 1024 + (size_t)stages * (16384 + 2048 + (c.N <= 128 ? 16384 : 32768));
        bench<<<148 * c.ctas, 64, smem>>>(c.N, c.a_mn, c.b_mn, c.shift, stages, 64, d, c.lbo_rows, c.n_acc);          // warm-up
        bench<<<148 * c.ctas, 64, smem>>>(c.N, c.a_mn, c.b_mn, c.shift, stages, 512, d, c.lbo_rows, c.n_acc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> h(148 * c.ctas);
        cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
        double s = 0;
        for (float v : h) s += v;
        printf("%-5d %-4d %-4d %-6d %-8d %-5d %-5d %-10.1f %-8d\n", c.N, c.a_mn, c.b_mn, c.shift, c.lbo_rows, c.n_acc, c.ctas,
               s / h.size(), c.N / 2);
    }
    return 0;
}
