// Microbenchmark (development tool): issue rate of scalar FP32 vs packed FP32x2 (sm_100
// FFMA2/FADD2/FMUL2), mixed with ALU work, and MUFU rate.  Prints warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float2 b0 = make_float2(a0, a1), b1 = make_float2(a2, a3), b2 = make_float2(a4, a5), b3 = make_float2(a6, a7);
    float2 b4 = b0, b5 = b1, b6 = b2, b7 = b3;
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {          // 8 scalar FFMA
            a0 = __fmaf_rn(a0, 1.0001f, 0.5f); a1 = __fmaf_rn(a1, 1.0001f, 0.5f); a2 = __fmaf_rn(a2, 1.0001f, 0.5f); a3 = __fmaf_rn(a3, 1.0001f, 0.5f);
            a4 = __fmaf_rn(a4, 1.0001f, 0.5f); a5 = __fmaf_rn(a5, 1.0001f, 0.5f); a6 = __fmaf_rn(a6, 1.0001f, 0.5f); a7 = __fmaf_rn(a7, 1.0001f, 0.5f);
        } else if (MODE == 1) {   // 8 packed FFMA2 (16 flop-lanes)
            b0 = __ffma2_rn(b0, m, c); b1 = __ffma2_rn(b1, m, c); b2 = __ffma2_rn(b2, m, c); b3 = __ffma2_rn(b3, m, c);
            b4 = __ffma2_rn(b4, m, c); b5 = __ffma2_rn(b5, m, c); b6 = __ffma2_rn(b6, m, c); b7 = __ffma2_rn(b7, m, c);
        } else if (MODE == 2) {   // 8 packed FADD2
            b0 = __fadd2_rn(b0, c); b1 = __fadd2_rn(b1, c); b2 = __fadd2_rn(b2, c); b3 = __fadd2_rn(b3, c);
            b4 = __fadd2_rn(b4, c); b5 = __fadd2_rn(b5, c); b6 = __fadd2_rn(b6, c); b7 = __fadd2_rn(b7, c);
        } else if (MODE == 3) {   // 8 packed FMUL2
            b0 = __fmul2_rn(b0, m); b1 = __fmul2_rn(b1, m); b2 = __fmul2_rn(b2, m); b3 = __fmul2_rn(b3, m);
            b4 = __fmul2_rn(b4, m); b5 = __fmul2_rn(b5, m); b6 = __fmul2_rn(b6, m); b7 = __fmul2_rn(b7, m);
        } else if (MODE == 4) {   // 8 scalar FADD (non-fused rn)
            a0 = __fadd_rn(a0, 0.5f); a1 = __fadd_rn(a1, 0.5f); a2 = __fadd_rn(a2, 0.5f); a3 = __fadd_rn(a3, 0.5f);
            a4 = __fadd_rn(a4, 0.5f); a5 = __fadd_rn(a5, 0.5f); a6 = __fadd_rn(a6, 0.5f); a7 = __fadd_rn(a7, 0.5f);
        } else if (MODE == 5) {   // 4 FFMA2 + 4 integer ALU ops (co-issue FMA pipe / ALU pipe)
            b0 = __ffma2_rn(b0, m, c); b1 = __ffma2_rn(b1, m, c); b2 = __ffma2_rn(b2, m, c); b3 = __ffma2_rn(b3, m, c);
            i0 = (i0 ^ 0x5bd1e995) + it; i1 = (i1 ^ 0x1b873593) + it; i2 = (i2 ^ 0x5bd1e995) + it; i3 = (i3 ^ 0x1b873593) + it;
        } else if (MODE == 6) {   // 4 scalar FFMA + 4 integer ALU ops
            a0 = __fmaf_rn(a0, 1.0001f, 0.5f); a1 = __fmaf_rn(a1, 1.0001f, 0.5f); a2 = __fmaf_rn(a2, 1.0001f, 0.5f); a3 = __fmaf_rn(a3, 1.0001f, 0.5f);
            i0 = (i0 ^ 0x5bd1e995) + it; i1 = (i1 ^ 0x1b873593) + it; i2 = (i2 ^ 0x5bd1e995) + it; i3 = (i3 ^ 0x1b873593) + it;
        } else if (MODE == 7) {   // 4 MUFU.RSQ + 4 FFMA
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a3));
            a4 = __fmaf_rn(a4, 1.0001f, 0.5f); a5 = __fmaf_rn(a5, 1.0001f, 0.5f); a6 = __fmaf_rn(a6, 1.0001f, 0.5f); a7 = __fmaf_rn(a7, 1.0001f, 0.5f);
        } else if (MODE == 8) {   // 8 FSETP+FSEL pairs (ALU pipe)
            a0 = a0 > 1.0f ? a1 : a0; a1 = a1 > 2.0f ? a2 : a1; a2 = a2 > 3.0f ? a3 : a2; a3 = a3 > 4.0f ? a4 : a3;
            a4 = a4 > 1.0f ? a5 : a4; a5 = a5 > 2.0f ? a6 : a5; a6 = a6 > 3.0f ? a7 : a6; a7 = a7 > 4.0f ? a0 : a7;
        }
    }
    long long t1 = clock64();
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + b0.x + b0.y + b1.x + b1.y + b2.x + b2.y + b3.x + b3.y + b4.x + b5.y + b6.x + b7.y + (float)(i0 + i1 + i2 + i3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int warps)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    k<MODE><<<148, warps * 32>>>(out, cyc, 1.0f);
    k<MODE><<<148, warps * 32>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double wi = (double)ITERS * instr_per_iter * warps;
    printf("%-34s warps/SM=%2d  %.2f warp-instr/clk/SM  (%.0f cycles)\n", name, warps, wi / avg, avg);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<0>("8x FFMA scalar", 8, w);
        run<1>("8x FFMA2 packed", 8, w);
        run<2>("8x FADD2 packed", 8, w);
        run<3>("8x FMUL2 packed", 8, w);
        run<4>("8x FADD scalar", 8, w);
        run<5>("4x FFMA2 + 8 int ALU", 12, w);
        run<6>("4x FFMA + 8 int ALU", 12, w);
        run<7>("4x MUFU.RSQ + 4x FFMA", 8, w);
        run<8>("8x FSETP+FSEL", 16, w);
    }
    return 0;
}
