// oc_gather.cuh — kernel 1: one thread per particle, 12-neighbour gather straight from global
// memory (L1/L2 provide the neighbour reuse).  Simple, bit-exact, used for tiny grids, as the
// in-library cross-check of the marching kernel, and as the "first correct CUDA path".
//
// Accumulation order = the order in which the reference's spring list (V:286-320) touches
// particle (i,j) in ComputeForces' second loop (V:462-483); derivation in SURVEY.md 8(a) A6
// and DESIGN.md.
#pragma once
#include "oc_core.cuh"

//
// The reference's two explicit-integrator siblings of the Verlet demo run through the same gather (SURVEY.md 8(f)3;
// "E:" = OpenCloth_ExplicitEuler/OpenCloth_ExplicitEuler/main.cpp, "S:" = OpenCloth_SemiImplicit/.../main.cpp): their
// state is X and V (buffer B holds V instead of X(t-1)), the spring force is the same formula on the stored velocities
// (E:434-466 / S:402-436), and a step writes two buffers (new X, new V).  kXV selects that form at compile time.
template <class M, bool kXV>
OC_HD void oc_nbr(const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                  int b, int ni, int nj, f3 xm, f3 vm, float rest, float nks, float kd, f3& F)
{
    long long n = oc_index(c, b, ni, nj);
    float4 a = A[n];
    float4 q = B[n];
    f3 xn = make_f3(a.x, a.y, a.z);
    f3 vn = kXV ? make_f3(q.x, q.y, q.z) : M::velocity(oc_delta<M>(a, q), c);
    f3 f = oc_spring<M>(xm, vm, xn, vn, rest, nks, kd);
    F.x = M::add(F.x, f.x); F.y = M::add(F.y, f.y); F.z = M::add(F.z, f.z);
}

// new X(t+1) of particle (i,j) of cloth b, as float4 with the collider flag in w
// (kXV: *vout receives the new V, w = 1)
template <class M, bool kXV = false>
OC_HD float4 oc_gather_particle(const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                                int b, int i, int j, float4* vout = nullptr)
{
    const int U = c.U, V = c.V;
    long long me = oc_index(c, b, i, j);
    float4 a = A[me];
    float4 q = B[me];
    f3 xm = make_f3(a.x, a.y, a.z);
    f3 d  = kXV ? make_f3(0.0f, 0.0f, 0.0f) : oc_delta<M>(a, q);
    f3 vm = kXV ? make_f3(q.x, q.y, q.z) : M::velocity(d, c);
    bool pinned = oc_pinned(c, b, i, j);
    f3 F = oc_base_force<M>(c, vm, pinned);
    if (!pinned) {
        // structural horizontal (V:288-291)
        if (i - 1 >= 0) oc_nbr<M, kXV>(c, A, B, b, i - 1, j, xm, vm, c.rh1[i - 1], c.nks_struct, c.kd_struct, F);
        if (i + 1 <  U) oc_nbr<M, kXV>(c, A, B, b, i + 1, j, xm, vm, c.rh1[i],     c.nks_struct, c.kd_struct, F);
        // structural vertical (V:294-297)
        if (j - 1 >= 0) oc_nbr<M, kXV>(c, A, B, b, i, j - 1, xm, vm, c.rv1[j - 1], c.nks_struct, c.kd_struct, F);
        if (j + 1 <  V) oc_nbr<M, kXV>(c, A, B, b, i, j + 1, xm, vm, c.rv1[j],     c.nks_struct, c.kd_struct, F);
        // shear (V:301-305)
        if (i - 1 >= 0 && j - 1 >= 0) oc_nbr<M, kXV>(c, A, B, b, i - 1, j - 1, xm, vm, M::sqrt(M::add(c.dx2[i - 1], c.dz2[j - 1])), c.nks_shear, c.kd_shear, F);
        if (i + 1 <  U && j - 1 >= 0) oc_nbr<M, kXV>(c, A, B, b, i + 1, j - 1, xm, vm, M::sqrt(M::add(c.dx2[i],     c.dz2[j - 1])), c.nks_shear, c.kd_shear, F);
        if (i - 1 >= 0 && j + 1 <  V) oc_nbr<M, kXV>(c, A, B, b, i - 1, j + 1, xm, vm, M::sqrt(M::add(c.dx2[i - 1], c.dz2[j])),     c.nks_shear, c.kd_shear, F);
        if (i + 1 <  U && j + 1 <  V) oc_nbr<M, kXV>(c, A, B, b, i + 1, j + 1, xm, vm, M::sqrt(M::add(c.dx2[i],     c.dz2[j])),     c.nks_shear, c.kd_shear, F);
        // bend horizontal, last spring of the row twice (V:309-314)
        if (i - 2 >= 0) oc_nbr<M, kXV>(c, A, B, b, i - 2, j, xm, vm, c.rh2[i - 2], c.nks_bend, c.kd_bend, F);
        if (i + 2 <  U) oc_nbr<M, kXV>(c, A, B, b, i + 2, j, xm, vm, c.rh2[i],     c.nks_bend, c.kd_bend, F);
        if (i == U - 3) oc_nbr<M, kXV>(c, A, B, b, i + 2, j, xm, vm, c.rh2[i],     c.nks_bend, c.kd_bend, F);
        if (i == U - 1) oc_nbr<M, kXV>(c, A, B, b, i - 2, j, xm, vm, c.rh2[i - 2], c.nks_bend, c.kd_bend, F);
        // bend vertical, last spring of the column twice (V:315-320)
        if (j - 2 >= 0) oc_nbr<M, kXV>(c, A, B, b, i, j - 2, xm, vm, c.rv2[j - 2], c.nks_bend, c.kd_bend, F);
        if (j + 2 <  V) oc_nbr<M, kXV>(c, A, B, b, i, j + 2, xm, vm, c.rv2[j],     c.nks_bend, c.kd_bend, F);
        if (j == V - 3) oc_nbr<M, kXV>(c, A, B, b, i, j + 2, xm, vm, c.rv2[j],     c.nks_bend, c.kd_bend, F);
        if (j == V - 1) oc_nbr<M, kXV>(c, A, B, b, i, j - 2, xm, vm, c.rv2[j - 2], c.nks_bend, c.kd_bend, F);
    }
    bool hit;
    if (kXV) {
        // IntegrateEuler (E:469-482) / IntegrateSemiImplicit (S:464-477), then EllipsoidCollision, which zeroes V (E:599)
        f3 vn = make_f3(M::add(vm.x, M::mul(F.x, c.dtm)), M::add(vm.y, M::mul(F.y, c.dtm)), M::add(vm.z, M::mul(F.z, c.dtm)));   // E:475
        const f3 vx = c.integ == 1 ? vm : vn;                                                     // E:476 oldV / S:470 new V
        f3 n = make_f3(M::add(xm.x, M::mul(c.dt, vx.x)), M::add(xm.y, M::mul(c.dt, vx.y)), M::add(xm.z, M::mul(c.dt, vx.z)));
        if (n.y < 0.0f) n.y = 0.0f;                                                               // E:478-480
        n = oc_collide<M>(c, n, &hit);
        if (hit) vn = make_f3(0.0f, 0.0f, 0.0f);
        *vout = make_float4(vn.x, vn.y, vn.z, oc_u2f(OC_W_PLAIN));
        return make_float4(n.x, n.y, n.z, oc_u2f(OC_W_PLAIN));
    }
    f3 n = oc_integrate_collide<M>(c, xm, d, F, &hit);
    return make_float4(n.x, n.y, n.z, oc_u2f(hit ? OC_W_HIT : OC_W_PLAIN));
}

#ifdef __CUDACC__
// grid: x = ceil(U/blockDim.x), y = rows to compute, z = cloth
template <class M>
__global__ void __launch_bounds__(128)
oc_k_gather(OcConst c, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, int row_a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = row_a + blockIdx.y;
    int b = blockIdx.z;
    if (i >= c.U) return;
    C[oc_index(c, b, i, j)] = oc_gather_particle<M>(c, A, B, b, i, j);
}
// state X, V (explicit / semi-implicit Euler): A = X, B = V -> C = new X, D = new V
template <class M>
__global__ void __launch_bounds__(128)
oc_k_gather_xv(OcConst c, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C, float4* __restrict__ D, int row_a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = row_a + blockIdx.y;
    int b = blockIdx.z;
    if (i >= c.U) return;
    float4 v;
    const long long o = oc_index(c, b, i, j);
    C[o] = oc_gather_particle<M, true>(c, A, B, b, i, j, &v);
    D[o] = v;
}
#endif
