// Do the packed FP32x2 instructions (FADD2/FMUL2/FFMA2) keep subnormals like the scalar ones?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out)
{
    float a = 1e-20f, b = 3e-20f;           // product 3e-40: subnormal
    float d1 = 1.5e-39f, d2 = 2.25e-39f;    // subnormal inputs
    float2 m = __fmul2_rn(make_float2(a, a), make_float2(b, b));
    float ms = __fmul_rn(a, b);
    float2 s = __fadd2_rn(make_float2(d1, d1), make_float2(d2, d2));
    float ss = __fadd_rn(d1, d2);
    float2 f = __ffma2_rn(make_float2(a, a), make_float2(b, b), make_float2(d1, d1));
    float fs = __fmaf_rn(a, b, d1);
    float2 g = __fmul2_rn(make_float2(d1, 2.0f), make_float2(2.0f, d2));   // subnormal operands
    float gs = __fmul_rn(d1, 2.0f);
    out[0] = __float_as_uint(m.x); out[1] = __float_as_uint(ms);
    out[2] = __float_as_uint(s.x); out[3] = __float_as_uint(ss);
    out[4] = __float_as_uint(f.y); out[5] = __float_as_uint(fs);
    out[6] = __float_as_uint(g.x); out[7] = __float_as_uint(gs);
    out[8] = __float_as_uint(g.y); out[9] = __float_as_uint(__fmul_rn(2.0f, d2));
}
int main()
{
    unsigned* d; unsigned h[10];
    cudaMalloc(&d, sizeof(h)); k<<<1, 1>>>(d); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* n[5] = { "mul -> subnormal result", "add of subnormals", "fma subnormal", "mul subnormal operand (lo)", "mul subnormal operand (hi)" };
    for (int i = 0; i < 5; ++i) printf("%-28s packed %08x scalar %08x %s\n", n[i], h[2 * i], h[2 * i + 1], h[2 * i] == h[2 * i + 1] ? "same" : "DIFFERENT");
    return 0;
}
