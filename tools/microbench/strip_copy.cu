// Is the marching kernels' MEMORY ACCESS PATTERN what bounds them?  Same pattern, no arithmetic: a CTA owns a strip of
// 128 columns and a segment of rows and walks down the rows; per row and thread it requests A[row][col] and B[row][col]
// (float4 each) with cp.async into shared memory, waits, and stores one float4 to C[row][col] - 48 bytes per element,
// the algorithmic traffic of a cloth step.  Reports elements/s and TB/s for the cloth sizes of the bench.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strip_copy strip_copy.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void cp16(void* s, const void* g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
template <int DEPTH>
__global__ void __launch_bounds__(128) k(const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, int U, int V, int rs, int nstrips)
{
    __shared__ float4 st[DEPTH + 1][2][128];
    const int strip = blockIdx.x % nstrips, seg = blockIdx.x / nstrips;
    const int col = strip * 128 + threadIdx.x;
    const int r0 = seg * rs, r1 = min(V, r0 + rs);
    if (col >= U) return;
    for (int r = r0 - DEPTH; r < r1; ++r) {
        const int lr = r + DEPTH, z = (lr - r0 + DEPTH + 1) % (DEPTH + 1);
        if (lr < r1) { cp16(&st[z][0][threadIdx.x], A + (size_t)lr * U + col); cp16(&st[z][1][threadIdx.x], B + (size_t)lr * U + col); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" :: "n"(DEPTH) : "memory");
        if (r >= r0) {
            const int zz = (r - r0 + DEPTH + 1) % (DEPTH + 1);
            const float4 a = st[zz][0][threadIdx.x], b = st[zz][1][threadIdx.x];
            C[(size_t)r * U + col] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w);
        }
    }
}
template <int DEPTH>
void run(int n, int rs, int reps)
{
    const size_t e = (size_t)n * n;
    float4 *buf[4];
    for (int i = 0; i < 4; ++i) { cudaMalloc(&buf[i], e * sizeof(float4)); cudaMemset(buf[i], 0, e * sizeof(float4)); }
    const int nstrips = (n + 127) / 128, nseg = (n + rs - 1) / rs;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) k<DEPTH><<<nstrips * nseg, 128>>>(buf[w % 3], buf[(w + 1) % 3], buf[(w + 2) % 3], n, n, rs, nstrips);
    cudaEventRecord(a);
    for (int w = 0; w < reps; ++w) k<DEPTH><<<nstrips * nseg, 128>>>(buf[w % 3], buf[(w + 1) % 3], buf[(w + 2) % 3], n, n, rs, nstrips);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double el = (double)e * reps / (ms * 1e-3);
    printf("%5d^2  rows/tile %4d  tiles %5d  depth %d: %7.1f G elements/s = %5.2f TB/s (48 B each)  %s\n", n, rs, nstrips * nseg, DEPTH, el / 1e9, el * 48 / 1e12,
           cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 4; ++i) cudaFree(buf[i]);
}
int main()
{
    run<0>(2048, 52, 200); run<1>(2048, 52, 200); run<3>(2048, 52, 200);
    run<0>(2048, 26, 200); run<3>(2048, 26, 200); run<3>(2048, 13, 200);
    run<0>(8192, 745, 20); run<1>(8192, 745, 20); run<3>(8192, 745, 20);
    run<0>(8192, 128, 20); run<3>(8192, 128, 20); run<3>(8192, 32, 20);
    return 0;
}
