// oc_normals.cuh — render hand-off: per-vertex normals of the cloth mesh, as the reference's lit demo
// computes them in UpdateNormals ("L:" = OpenCloth_ExplicitEuler_TextureMapped_Lit/.../main.cpp:684-707) over its
// triangle list (L:313-327: two triangles per cell, the diagonal alternating with the cell's parity).
//
// Each triangle adds cross(p2 - p1, p3 - p1) / 3.0f to its three vertices (L:693-700), then every normal is
// normalised (L:703-706).  Here one thread gathers the (up to six) triangles of its vertex in the order the list visits
// them, so the sums round exactly like the reference's scatter loop.  One deliberate difference: the reference never
// clears the accumulators between frames (L:698-700 add onto last frame's unit normal), which blends every frame with
// history; this is the history-free result, i.e. what the reference computes on its FIRST call.
#pragma once
#include "oc_core.cuh"

template <class M>
OC_HD void oc_tri_add(const float4* __restrict__ X, long long base, int U, int a1, int b1, int a2, int b2, int a3, int b3, f3& n)
{
    // triangle (p1, p2, p3) given as (row, column) pairs
    const float4 q1 = X[base + (long long)a1 * U + b1], q2 = X[base + (long long)a2 * U + b2], q3 = X[base + (long long)a3 * U + b3];
    const f3 e = make_f3(M::sub(q2.x, q1.x), M::sub(q2.y, q1.y), M::sub(q2.z, q1.z));      // p2 - p1
    const f3 f = make_f3(M::sub(q3.x, q1.x), M::sub(q3.y, q1.y), M::sub(q3.z, q1.z));      // p3 - p1
    // glm::cross (func_geometric.inl:181-193)
    const f3 c = make_f3(M::sub(M::mul(e.y, f.z), M::mul(f.y, e.z)), M::sub(M::mul(e.z, f.x), M::mul(f.z, e.x)), M::sub(M::mul(e.x, f.y), M::mul(f.x, e.y)));
    n.x = M::add(n.x, M::div(c.x, 3.0f)); n.y = M::add(n.y, M::div(c.y, 3.0f)); n.z = M::add(n.z, M::div(c.z, 3.0f));   // L:698-700
}

// normal of vertex (column i, row j) of cloth b
template <class M>
OC_HD f3 oc_vertex_normal(const OcConst& c, const float4* __restrict__ X, int b, int i, int j)
{
    const int U = c.U, V = c.V;
    const long long base = oc_index(c, b, 0, 0);
    f3 n = make_f3(0.0f, 0.0f, 0.0f);
    // the cells that contain the vertex, in the row-major order of the index list; cell (a, bb) has corners
    // i0 = (a, bb), i1 = (a, bb+1), i2 = (a+1, bb), i3 = (a+1, bb+1)                                   L:316-319
    for (int k = 0; k < 4; ++k) {
        const int a = j - 1 + (k >> 1), bb = i - 1 + (k & 1);
        if (a < 0 || a >= V - 1 || bb < 0 || bb >= U - 1) continue;
        const int me = (j - a) * 2 + (i - bb);                  // which corner of the cell the vertex is: 0..3
        if ((bb + a) % 2) {
            // (i0, i2, i1) then (i1, i2, i3)                                                         L:321-322
            if (me != 3) oc_tri_add<M>(X, base, U, a, bb, a + 1, bb, a, bb + 1, n);
            if (me != 0) oc_tri_add<M>(X, base, U, a, bb + 1, a + 1, bb, a + 1, bb + 1, n);
        } else {
            // (i0, i2, i3) then (i0, i3, i1)                                                         L:324-325
            if (me != 1) oc_tri_add<M>(X, base, U, a, bb, a + 1, bb, a + 1, bb + 1, n);
            if (me != 2) oc_tri_add<M>(X, base, U, a, bb, a + 1, bb + 1, a, bb + 1, n);
        }
    }
    // glm::normalize: n * (1 / sqrt(dot(n, n)))                                                       L:705
    const float inv = M::rcp(M::sqrt(M::dot(n, n)));
    return make_f3(M::mul(n.x, inv), M::mul(n.y, inv), M::mul(n.z, inv));
}

#ifdef __CUDACC__
// out: stride 3 or 4 floats per vertex (w = 0), cloth-major then row-major like the positions
__global__ void __launch_bounds__(128)
oc_k_normals(OcConst c, const float4* __restrict__ X, float* __restrict__ out, int stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, b = blockIdx.z;
    if (i >= c.U) return;
    const f3 n = oc_vertex_normal<MathExact>(c, X, b, i, j);
    float* o = out + (((long long)b * c.V + j) * c.U + i) * stride;
    o[0] = n.x; o[1] = n.y; o[2] = n.z; if (stride == 4) o[3] = 0.0f;
}
#endif
