// oc_resident.cuh — kernel 4: small cloths RESIDENT in shared memory.
//
// The reference's own configuration is a 21 x 21 cloth (V:59-62): 441 particles, 5 KB of state.  Stepping it with one
// launch per substep is pure launch latency, and one thread per particle is a serial chain of 12 springs.  Here one
// CTA owns one cloth (blockIdx.x = cloth of a batch), loads X(t), X(t-1) once, takes ALL the substeps of the oc_step
// call without touching HBM, and writes X(t+n), X(t+n-1) back.  A substep is two phases with a barrier each:
//   S  the FORWARD springs of the cloth, (+1,0) (+2,0) (0,+1) (0,+2) (+1,+1) (-1,+1) of every particle, two per work
//      item (3 N items over the CTA's threads, packed FP32x2 like oc_k_march): f = springForce(p, partner) into
//      F[type][p].  Every spring is evaluated once, by its upper/left end, with the branch-free exact sequences
//      (oc_spring2; IEEE intrinsics on the rare bad operand).
//   P  every particle is one work item: base force, the <= 14 spring terms in the reference's order (own forces added,
//      partners' forces subtracted: f(b,a) == -f(a,b) bitwise), IntegrateVerlet, EllipsoidCollision; then the new
//      velocity and X - X_last once per particle.
// The serial chain per substep is one spring + one accumulation instead of twelve springs.  Arithmetic and order
// are those of the other kernels: bit-identical to the reference in exact mode.
// Limits: whole cloths (no row band) of at most OC_RESIDENT_MAX_PARTICLES particles.
#pragma once
#include "oc_core.cuh"

#define OC_RESIDENT_MAX_PARTICLES 1536
#define OC_RESIDENT_THREADS 1024
#define OC_RESIDENT_MAX_STEPS 4096          /* substeps per launch (bounds the run time of one launch) */

// shared memory: 32 floats per particle, structure of arrays (plain offsets: no pointer tables in local memory)
struct OcResidentSmem {
    float* base; int N;
    OC_HD float* X(int k)  const { return base + (size_t)(0 + k) * N; }       // X(t)
    OC_HD float* Xp(int k) const { return base + (size_t)(3 + k) * N; }       // X(t-1)
    OC_HD float* V(int k)  const { return base + (size_t)(6 + k) * N; }       // velocity
    OC_HD float* D(int k)  const { return base + (size_t)(9 + k) * N; }       // X(t) - X_last(t)
    OC_HD float* F(int t, int k) const { return base + (size_t)(12 + 3 * t + k) * N; }   // forward spring forces by type
    OC_HD float* W()  const { return base + (size_t)30 * N; }                 // w of X(t) (collider flag)
    OC_HD float* Wp() const { return base + (size_t)31 * N; }                 // w of X(t-1)
    static OC_HD size_t bytes(int N) { return (size_t)32 * N * sizeof(float); }
};

template <class M, class Ctx>
OC_HD void oc_resident_body(Ctx& ctx, const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                            float4* __restrict__ dst, float4* __restrict__ dst_prev, int n_steps)
{
    const int U = c.U, V = c.V, N = U * V;
    const int tid = ctx.tid(), T = ctx.nthreads();
    OcResidentSmem s;
    s.base = reinterpret_cast<float*>(ctx.smem()); s.N = N;
    const long long base = (long long)ctx.bx() * c.cloth_stride;
    const float ydt = oc_rcp_bf(c.dt);

    // derived per-particle state from a position pair: X - X_last (V:530 flag honoured) and the velocity
    auto derive = [&](int p, f3 d) {
        bool bad = false;
        f3 v = oc_velocity_bf<M>(d, c, ydt, bad);
        if (M::kExact && bad) v = M::velocity(d, c);
        s.D(0)[p] = d.x; s.D(1)[p] = d.y; s.D(2)[p] = d.z;
        s.V(0)[p] = v.x; s.V(1)[p] = v.y; s.V(2)[p] = v.z;
    };
    for (int p = tid; p < N; p += T) {
        const float4 a = A[base + p], q = B[base + p];
        s.X(0)[p] = a.x; s.X(1)[p] = a.y; s.X(2)[p] = a.z; s.W()[p] = a.w;
        s.Xp(0)[p] = q.x; s.Xp(1)[p] = q.y; s.Xp(2)[p] = q.z; s.Wp()[p] = q.w;
        derive(p, oc_delta<M>(a, q));
    }
    ctx.sync();

    for (int step = 0; step < n_steps; ++step) {
        // ---- S: forward springs, two per work item (packed FP32x2): (+1,0)&(+2,0), (0,+1)&(0,+2), (+1,+1)&(-1,+1) ---
        for (int w = tid; w < 3 * N; w += T) {
            const int pr = w / N, p = w - pr * N;
            const int j = p / U, i = p - j * U;
            int qi0 = i, qj0 = j, qi1 = i, qj1 = j;
            float2 rest, nks, kd;
            bool badr = false;
            if (pr == 0) {                                                       // V:288-291, V:309-314
                qi0 = i + 1; qi1 = i + 2;
                rest = make_float2(c.rh1[i], c.rh2[i]);
                nks = make_float2(c.nks_struct, c.nks_bend); kd = make_float2(c.kd_struct, c.kd_bend);
            } else if (pr == 1) {                                                // V:294-297, V:315-320
                qj0 = j + 1; qj1 = j + 2;
                rest = make_float2(c.rv1[j], c.rv2[j]);
                nks = make_float2(c.nks_struct, c.nks_bend); kd = make_float2(c.kd_struct, c.kd_bend);
            } else {                                                             // V:301-305
                qi0 = i + 1; qj0 = j + 1; qi1 = i - 1; qj1 = j + 1;
                rest = oc_sqrt2<M>(make_float2(M::add(c.dx2[i], c.dz2[j]), M::add(c.dx2[i > 0 ? i - 1 : 0], c.dz2[j])), badr);
                nks = p_bc(c.nks_shear); kd = p_bc(c.kd_shear);
            }
            const bool e0 = qi0 < U && qj0 < V, e1 = qi1 >= 0 && qi1 < U && qj1 < V;
            if (!e0 && !e1) continue;                                            // no such springs: their slots are never read
            const f3 xp = make_f3(s.X(0)[p], s.X(1)[p], s.X(2)[p]), vp = make_f3(s.V(0)[p], s.V(1)[p], s.V(2)[p]);
            // a missing partner is replaced by a ghost at unit distance moving with the particle (result not stored)
            const int q0 = e0 ? qj0 * U + qi0 : p, q1 = e1 ? qj1 * U + qi1 : p;
            const float g0 = e0 ? 0.0f : 1.0f, g1 = e1 ? 0.0f : 1.0f;
            OcPair3 qx, qv;
            qx.x = make_float2(s.X(0)[q0] + g0, s.X(0)[q1] + g1); qx.y = make_float2(s.X(1)[q0], s.X(1)[q1]); qx.z = make_float2(s.X(2)[q0], s.X(2)[q1]);
            qv.x = make_float2(s.V(0)[q0], s.V(0)[q1]); qv.y = make_float2(s.V(1)[q0], s.V(1)[q1]); qv.z = make_float2(s.V(2)[q0], s.V(2)[q1]);
            bool bad = false;
            OcPair3 f = oc_spring2<M>(xp, vp, qx, qv, M::kExact ? rest : p_mul(rest, nks), nks, kd, c.one, bad);
            if (M::kExact && (bad | badr)) {                                     // rare: IEEE intrinsics
                if (pr == 2) rest = make_float2(M::sqrt(M::add(c.dx2[i], c.dz2[j])), M::sqrt(M::add(c.dx2[i > 0 ? i - 1 : 0], c.dz2[j])));
                const f3 f0 = oc_spring<M>(xp, vp, make_f3(qx.x.x, qx.y.x, qx.z.x), make_f3(qv.x.x, qv.y.x, qv.z.x), rest.x, nks.x, kd.x);
                const f3 f1 = oc_spring<M>(xp, vp, make_f3(qx.x.y, qx.y.y, qx.z.y), make_f3(qv.x.y, qv.y.y, qv.z.y), rest.y, nks.y, kd.y);
                f.x = make_float2(f0.x, f1.x); f.y = make_float2(f0.y, f1.y); f.z = make_float2(f0.z, f1.z);
            }
            if (e0) { s.F(2 * pr, 0)[p] = f.x.x; s.F(2 * pr, 1)[p] = f.y.x; s.F(2 * pr, 2)[p] = f.z.x; }
            if (e1) { s.F(2 * pr + 1, 0)[p] = f.x.y; s.F(2 * pr + 1, 1)[p] = f.y.y; s.F(2 * pr + 1, 2)[p] = f.z.y; }
        }
        ctx.sync();
        // ---- P: accumulate in the reference's order, integrate, collide -------------------------------------
        for (int p = tid; p < N; p += T) {
            const int j = p / U, i = p - j * U;
            const f3 xm = make_f3(s.X(0)[p], s.X(1)[p], s.X(2)[p]);
            const f3 vm = make_f3(s.V(0)[p], s.V(1)[p], s.V(2)[p]);
            const f3 d  = make_f3(s.D(0)[p], s.D(1)[p], s.D(2)[p]);
            const bool pinned = oc_pinned(c, ctx.bx(), i, j);
            f3 F = oc_base_force<M>(c, vm, pinned);
#define OC_ADD(t, q) { F.x = M::add(F.x, s.F(t, 0)[q]); F.y = M::add(F.y, s.F(t, 1)[q]); F.z = M::add(F.z, s.F(t, 2)[q]); }
#define OC_SUB(t, q) { F.x = M::sub(F.x, s.F(t, 0)[q]); F.y = M::sub(F.y, s.F(t, 1)[q]); F.z = M::sub(F.z, s.F(t, 2)[q]); }
            if (!pinned) {
                if (i - 1 >= 0) OC_SUB(0, p - 1)                                   //  1 (i-1, j)
                if (i + 1 <  U) OC_ADD(0, p)                                       //  2 (i+1, j)
                if (j - 1 >= 0) OC_SUB(2, p - U)                                   //  3 (i, j-1)
                if (j + 1 <  V) OC_ADD(2, p)                                       //  4 (i, j+1)
                if (i - 1 >= 0 && j - 1 >= 0) OC_SUB(4, p - U - 1)                 //  5 (i-1, j-1)
                if (i + 1 <  U && j - 1 >= 0) OC_SUB(5, p - U + 1)                 //  6 (i+1, j-1)
                if (i - 1 >= 0 && j + 1 <  V) OC_ADD(5, p)                         //  7 (i-1, j+1)
                if (i + 1 <  U && j + 1 <  V) OC_ADD(4, p)                         //  8 (i+1, j+1)
                if (i - 2 >= 0) OC_SUB(1, p - 2)                                   //  9 (i-2, j)
                if (i + 2 <  U) OC_ADD(1, p)                                       // 10 (i+2, j)
                if (i == U - 3) OC_ADD(1, p)                                       // 11 last bend spring of the row twice (V:313)
                if (i == U - 1 && i - 2 >= 0) OC_SUB(1, p - 2)
                if (j - 2 >= 0) OC_SUB(3, p - 2 * U)                               // 12 (i, j-2)
                if (j + 2 <  V) OC_ADD(3, p)                                       // 13 (i, j+2)
                if (j == V - 3) OC_ADD(3, p)                                       // 14 last bend spring of the column twice (V:319)
                if (j == V - 1 && j - 2 >= 0) OC_SUB(3, p - 2 * U)
            }
#undef OC_ADD
#undef OC_SUB
            bool hit;
            const f3 n = oc_integrate_collide<M>(c, xm, d, F, &hit);
            // rotate: X(t) becomes X(t-1); the particle's own entries only, nobody else reads them in this phase
            s.Xp(0)[p] = xm.x; s.Xp(1)[p] = xm.y; s.Xp(2)[p] = xm.z; s.Wp()[p] = s.W()[p];
            s.X(0)[p] = n.x; s.X(1)[p] = n.y; s.X(2)[p] = n.z; s.W()[p] = oc_u2f(hit ? OC_W_HIT : OC_W_PLAIN);
            derive(p, hit ? make_f3(0.0f, 0.0f, 0.0f) : make_f3(M::sub(n.x, xm.x), M::sub(n.y, xm.y), M::sub(n.z, xm.z)));
        }
        ctx.sync();
    }
    for (int p = tid; p < N; p += T) {
        dst[base + p] = make_float4(s.X(0)[p], s.X(1)[p], s.X(2)[p], s.W()[p]);
        if (n_steps > 1) dst_prev[base + p] = make_float4(s.Xp(0)[p], s.Xp(1)[p], s.Xp(2)[p], s.Wp()[p]);
    }
}

#ifdef __CUDACC__
struct OcResidentCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ int bx() const { return blockIdx.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ unsigned char* smem() const { extern __shared__ __align__(16) unsigned char oc_dyn_smem[]; return oc_dyn_smem; }
};
template <class M>
__global__ void __launch_bounds__(OC_RESIDENT_THREADS)
oc_k_resident(const __grid_constant__ OcConst c, const float4* __restrict__ A, const float4* __restrict__ B,
              float4* __restrict__ dst, float4* __restrict__ dst_prev, int n_steps)
{
    OcResidentCtx ctx;
    oc_resident_body<M, OcResidentCtx>(ctx, c, A, B, dst, dst_prev, n_steps);
}
#endif
