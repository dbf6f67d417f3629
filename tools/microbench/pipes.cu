// Micro-benchmarks that size the MFCC kernel design: packed-fp32 (FFMA2/FADD2) issue rate, and whether
// SHFL and LDS/STS share one data pipe.  nvcc -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#define ULL unsigned long long
__device__ __forceinline__ ULL fma2(ULL a, ULL b, ULL c) { ULL d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ ULL add2(ULL a, ULL b) { ULL d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int ITER = 2048;
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int n) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float f[8]; ULL u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { f[j] = threadIdx.x + j; u[j] = ((ULL)__float_as_uint(f[j]) << 32) | __float_as_uint(f[j] * 0.5f); }
    float acc = 0.f; ULL cu = u[0];
    const float* base = sm + (threadIdx.x & 1023);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) f[j] = fma1(f[j], 1.0001f, 0.5f);                     // FFMA scalar
            if (MODE == 1) u[j] = fma2(u[j], cu, cu);                             // FFMA2
            if (MODE == 2) f[j] = add1(f[j], 0.5f);                              // FADD
            if (MODE == 3) u[j] = add2(u[j], cu);                                 // FADD2
            if (MODE == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(base + ((j * 32 + it) & 1023)))); acc += v; }   // LDS.32
            if (MODE == 5) { f[j] = __shfl_xor_sync(0xffffffffu, f[j], 1 + (j & 7)); }                // SHFL
            if (MODE == 6) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(base + ((j * 32 + it) & 1023)))); acc += v;
                             f[j] = __shfl_xor_sync(0xffffffffu, f[j], 1 + (j & 7)); }                // LDS + SHFL
            if (MODE == 7) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + 2 * ((threadIdx.x + j * 32 + it) & 1023)))); acc += v.x + v.y; }   // LDS.64
            if (MODE == 8) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sm + 4 * ((threadIdx.x + j * 32 + it) & 1023)))); acc += v.x + v.y + v.z + v.w; }   // LDS.128
            if (MODE == 9) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sm + 4 * ((j * 32 + it) & 1023)))); acc += v.x + v.y + v.z + v.w; }   // LDS.128 broadcast
            if (MODE == 10) { u[j] = fma2(u[j], cu, cu); float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(base + ((j * 32 + it) & 1023)))); acc += v; } // FFMA2 + LDS + FADD
            if (MODE == 11) { u[j] = fma2(u[j], cu, cu); f[j] = fma1(f[j], 1.0001f, 0.5f); }        // FFMA2 + FFMA
            if (MODE == 12) { asm volatile("st.shared.f32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(base + ((j * 32 + it) & 1023))), "f"(f[j])); } // STS
            if (MODE == 13) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(sm + 2 * ((threadIdx.x + j * 32 + it) & 1023))), "f"(f[j]), "f"(f[j])); } // STS.64
        }
    }
    float s = acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j] + __uint_as_float((unsigned)u[j]) + __uint_as_float((unsigned)(u[j] >> 32));
    if (s == 123.456f) out[0] = s;
}
template <int MODE>
void run(const char* name, float* d, double per_iter_ops) {
    int sm = 148;
    for (int warps : {4, 8, 16}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<MODE><<<sm, warps * 32>>>(d, 64);
        cudaEventRecord(a);
        k<MODE><<<sm, warps * 32>>>(d, ITER);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        double cyc = ms * 1e-3 * 1.965e9;
        double inst = (double)ITER * 8 * warps * per_iter_ops;
        printf("%-28s warps/SM=%2d  %.3f ms  warp-instr/clk/SM = %.3f  (clk/instr/SM %.3f)\n", name, warps, ms, inst / cyc, cyc / inst);
    }
}
int main() {
    float* d; cudaMalloc(&d, 1024);
    run<0>("FFMA", d, 1); run<1>("FFMA2", d, 1); run<2>("FADD", d, 1); run<3>("FADD2", d, 1);
    run<4>("LDS.32", d, 1); run<5>("SHFL", d, 1); run<6>("LDS.32+SHFL (pairs)", d, 1);
    run<7>("LDS.64", d, 1); run<8>("LDS.128", d, 1); run<9>("LDS.128 bcast", d, 1);
    run<10>("FFMA2+LDS+FADD (triples)", d, 1); run<11>("FFMA2+FFMA (pairs)", d, 1);
    run<12>("STS.32", d, 1); run<13>("STS.64", d, 1);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
