// Cost model of the shared-memory / shuffle data path on sm_100: loops of LDS.64, LDS.128, STS.64, STS.128, SHFL and
// mixes of them (32 warps per SM, independent operations, distinct addresses), in cycles per warp-instruction per SM.
// Answers: do shuffles share the pipe with LDS/STS (yes: times add), what does a 64-bit store cost against a load.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shfl_lds shfl_lds.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int NL, int NL4, int NS, int NS4, int NSH>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters)
{
    extern __shared__ float4 dyn[];
    float2* buf = reinterpret_cast<float2*>(dyn);          // [16][1024] float2 = 128 KB
    const int t = threadIdx.x;
    float2 acc = make_float2(t, t * 0.5f);
    float4 acc4 = make_float4(t, 1.f, 2.f, 3.f);
    for (int r = 0; r < 16; ++r) buf[r * 1024 + t] = acc;
    __syncthreads();
    float s = t;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NL; ++j) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa(&buf[j * 1024 + t]))); acc.x += v.x; acc.y += v.y; }
#pragma unroll
        for (int j = 0; j < NL4; ++j) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa(&dyn[(j & 7) * 1024 + t]))); acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w; }
#pragma unroll
        for (int j = 0; j < NS; ++j) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(sa(&buf[j * 1024 + t])), "f"(acc.x), "f"(acc.y) : "memory"); }
#pragma unroll
        for (int j = 0; j < NS4; ++j) { asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(sa(&dyn[(j & 7) * 1024 + t])), "f"(acc4.x), "f"(acc4.y), "f"(acc4.z), "f"(acc4.w) : "memory"); }
#pragma unroll
        for (int j = 0; j < NSH; ++j) { s += __shfl_sync(0xffffffffu, acc.x + j, (t + 1 + j) & 31); }
    }
    out[blockIdx.x * blockDim.x + t] = acc.x + acc.y + s + acc4.x + acc4.y + acc4.z + acc4.w;
}
template <int NL, int NL4, int NS, int NS4, int NSH>
void run(const char* name, float* d, int sms)
{
    const int iters = 2000, threads = 1024;
    const size_t smem = 16 * 1024 * sizeof(float2);
    cudaFuncSetAttribute(k<NL, NL4, NS, NS4, NSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NL, NL4, NS, NS4, NSH><<<sms, threads, smem>>>(d, 10);
    cudaEventRecord(a);
    k<NL, NL4, NS, NS4, NSH><<<sms, threads, smem>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3 / iters;
    const int n = NL + NL4 + NS + NS4 + NSH;
    printf("%-34s %7.1f cycles/iter/SM   %.2f cycles per warp-instruction  (%s)\n", name, cycles, cycles / (n * 32.0), cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, sms * 1024 * sizeof(float));
    run<16, 0, 0, 0, 0>("16 LDS.64", d, sms);
    run<0, 16, 0, 0, 0>("16 LDS.128", d, sms);
    run<0, 0, 16, 0, 0>("16 STS.64", d, sms);
    run<0, 0, 0, 16, 0>("16 STS.128", d, sms);
    run<0, 0, 0, 0, 16>("16 SHFL", d, sms);
    run<16, 0, 0, 0, 16>("16 LDS.64 + 16 SHFL", d, sms);
    run<16, 0, 16, 0, 0>("16 LDS.64 + 16 STS.64", d, sms);
    run<8, 0, 8, 0, 16>("8 LDS.64 + 8 STS.64 + 16 SHFL", d, sms);
    return 0;
}
