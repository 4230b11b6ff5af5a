// Reproducer: CUDA 12.9 ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (see oc_core.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cubin fuse2.cu && cuobjdump -sass fuse2.cubin | grep -E "FFMA|FMUL|FADD"
#include <cuda_runtime.h>
__device__ __forceinline__ float2 mul2_via_fma(float2 a, float2 b) { return __ffma2_rn(a, b, make_float2(-0.0f, -0.0f)); }
__global__ void k(const float2* in, float2* out)
{
    int t = threadIdx.x;
    float2 a = in[t], b = in[t + 32], c = in[t + 64], d = in[t + 96];
    // G: scalar products feeding a packed add
    float2 g = __fadd2_rn(make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), c);
    // A: product via fma(a,b,-0) then packed add
    float2 h = __fadd2_rn(mul2_via_fma(a, d), c);
    // B: packed mul, then add expressed as fma(m, 1, c)
    float2 i = __ffma2_rn(__fmul2_rn(b, d), make_float2(1.0f, 1.0f), c);
    // sum of two packed products
    float2 j = __fadd2_rn(__fmul2_rn(a, c), __fmul2_rn(b, c));
    out[t] = make_float2(g.x + h.x + i.x + j.x, g.y + h.y + i.y + j.y);
}
