// oc_provot.cuh — ApplyProvotDynamicInverse, the over-stretch correction pass of the reference.
//
//   "V:" OpenCloth_Verlet/OpenCloth_Verlet/main.cpp:486-508   — present but disabled in StepPhysics (V:561);
//   "E:" OpenCloth_ExplicitEuler/.../main.cpp:554-577, "S:" OpenCloth_SemiImplicit/.../main.cpp:437-462 — enabled
//        (E:639 / S:531), after EllipsoidCollision.
// For every spring of the list, in list order: if it is longer than its rest length, both ends are moved by half the
// excess along the spring (a pinned end is not moved; V:498-505).  Bit-exact here in both forms:
//
//   * Euler demos: the correction is ADDED TO V while X stays fixed during the pass, so every spring's correction is
//     independent of the others and only the order of the additions into one particle matters: a per-particle gather in
//     the order the spring list touches the particle (the order of the force gather, oc_gather.cuh) reproduces the
//     sequential loop bit for bit, fully parallel (oc_k_provot_v).
//   * Verlet demo: the correction MOVES X IN PLACE, so every spring sees the positions its predecessors left — a
//     Gauss-Seidel sweep whose result depends on the list order.  The list order is reproduced, not approximated:
//     the springs of one class only conflict along one direction, so
//       - structural and bend springs along a row (V:288-291, V:309-314): rows are independent, one thread walks a row;
//       - structural and bend springs along a column (V:294-297, V:315-320): one thread walks a column (coalesced);
//       - shear springs (V:301-305, two per cell, cells row-major): cell (r, c) conflicts with its eight neighbours
//         only, and the list puts (r, c-1) and (r-1, c+1) — and through them every earlier conflicting cell — before
//         it.  Thread r takes row r of cells and runs two cells behind thread r-1: a skewed wavefront, cell (r, c) at
//         time c + 2r, one barrier per time step (oc_k_provot_shear).
//     The pass is inherently sequential along those directions (U + 2V barrier-separated steps for the shear class):
//     it is an optional pass, off by default like in the reference, and costs far more than the step it follows.
#pragma once
#include "oc_core.cuh"

// correction of one spring with first end p, second end q: normalize(p - q) * ((|p - q| - rest) / 2), or 0 if not stretched
template <class M>
OC_HD f3 oc_provot_delta(f3 p, f3 q, float rest, bool* on)
{
    const f3 d = make_f3(M::sub(p.x, q.x), M::sub(p.y, q.y), M::sub(p.z, q.z));      // V:491
    const float sqr = M::dot(d, d);
    float dist = M::sqrt(sqr);                                                        // V:492 glm::length
    *on = dist > rest;                                                                // V:493
    if (!*on) return make_f3(0.0f, 0.0f, 0.0f);
    const float inv = M::rcp(dist);                                                   // V:496 glm::normalize: 1.0f / sqrt(dot)
    dist = M::sub(dist, rest);                                                        // V:494
    dist = M::div(dist, 2.0f);                                                        // V:495
    return make_f3(M::mul(M::mul(d.x, inv), dist), M::mul(M::mul(d.y, inv), dist), M::mul(M::mul(d.z, inv), dist));   // V:496-497
}

// ---- Euler demos: V of particle (i,j) after the pass (X fixed) ------------------------------------------------
//   me_is_p1: whether this particle is the spring's first end in the list.  It is corrected unless it is pinned and
//   (it is the first end, or the first end is not pinned): E:567-574.
template <class M>
OC_HD void oc_provot_v_term(const OcConst& c, const float4* __restrict__ X, int b, int i, int j, int ni, int nj,
                            bool me_is_p1, float rest, f3 xm, bool pin_me, f3& v)
{
    const bool pin_q = oc_pinned(c, b, ni, nj);
    const bool upd = me_is_p1 ? !pin_me : (pin_q || !pin_me);
    if (!upd) return;
    const float4 a = X[oc_index(c, b, ni, nj)];
    bool on;
    const f3 d = oc_provot_delta<M>(xm, make_f3(a.x, a.y, a.z), rest, &on);
    if (on) { v.x = M::sub(v.x, d.x); v.y = M::sub(v.y, d.y); v.z = M::sub(v.z, d.z); }   // V[p1] -= deltaP == V[p2] += -deltaP
    (void)i; (void)j;
}
template <class M>
OC_HD float4 oc_provot_v_particle(const OcConst& c, const float4* __restrict__ X, const float4* Vb, int b, int i, int j)
{
    const int U = c.U, V = c.V;
    const long long me = oc_index(c, b, i, j);
    const float4 a = X[me];
    float4 vv = Vb[me];
    const f3 xm = make_f3(a.x, a.y, a.z);
    f3 v = make_f3(vv.x, vv.y, vv.z);
    const bool pin = oc_pinned(c, b, i, j);
    // the order in which the spring list touches the particle (as in oc_gather_particle); first ends: the left / upper
    // particle of a structural or bend spring, the upper-left of a "\" shear spring, the LOWER-left of a "/" one (V:304)
    if (i - 1 >= 0) oc_provot_v_term<M>(c, X, b, i, j, i - 1, j, false, c.rh1[i - 1], xm, pin, v);
    if (i + 1 <  U) oc_provot_v_term<M>(c, X, b, i, j, i + 1, j, true,  c.rh1[i],     xm, pin, v);
    if (j - 1 >= 0) oc_provot_v_term<M>(c, X, b, i, j, i, j - 1, false, c.rv1[j - 1], xm, pin, v);
    if (j + 1 <  V) oc_provot_v_term<M>(c, X, b, i, j, i, j + 1, true,  c.rv1[j],     xm, pin, v);
    if (i - 1 >= 0 && j - 1 >= 0) oc_provot_v_term<M>(c, X, b, i, j, i - 1, j - 1, false, M::sqrt(M::add(c.dx2[i - 1], c.dz2[j - 1])), xm, pin, v);
    if (i + 1 <  U && j - 1 >= 0) oc_provot_v_term<M>(c, X, b, i, j, i + 1, j - 1, true,  M::sqrt(M::add(c.dx2[i],     c.dz2[j - 1])), xm, pin, v);
    if (i - 1 >= 0 && j + 1 <  V) oc_provot_v_term<M>(c, X, b, i, j, i - 1, j + 1, false, M::sqrt(M::add(c.dx2[i - 1], c.dz2[j])),     xm, pin, v);
    if (i + 1 <  U && j + 1 <  V) oc_provot_v_term<M>(c, X, b, i, j, i + 1, j + 1, true,  M::sqrt(M::add(c.dx2[i],     c.dz2[j])),     xm, pin, v);
    if (i - 2 >= 0) oc_provot_v_term<M>(c, X, b, i, j, i - 2, j, false, c.rh2[i - 2], xm, pin, v);
    if (i + 2 <  U) oc_provot_v_term<M>(c, X, b, i, j, i + 2, j, true,  c.rh2[i],     xm, pin, v);
    if (i == U - 3) oc_provot_v_term<M>(c, X, b, i, j, i + 2, j, true,  c.rh2[i],     xm, pin, v);
    if (i == U - 1) oc_provot_v_term<M>(c, X, b, i, j, i - 2, j, false, c.rh2[i - 2], xm, pin, v);
    if (j - 2 >= 0) oc_provot_v_term<M>(c, X, b, i, j, i, j - 2, false, c.rv2[j - 2], xm, pin, v);
    if (j + 2 <  V) oc_provot_v_term<M>(c, X, b, i, j, i, j + 2, true,  c.rv2[j],     xm, pin, v);
    if (j == V - 3) oc_provot_v_term<M>(c, X, b, i, j, i, j + 2, true,  c.rv2[j],     xm, pin, v);
    if (j == V - 1) oc_provot_v_term<M>(c, X, b, i, j, i, j - 2, false, c.rv2[j - 2], xm, pin, v);
    vv.x = v.x; vv.y = v.y; vv.z = v.z;
    return vv;
}

// ---- Verlet demo: one spring of the in-place sweep (V:489-506) on positions held by the caller ------------------
template <class M>
OC_HD void oc_provot_x_spring(f3& p1, f3& p2, bool pin1, bool pin2, float rest)
{
    bool on;
    const f3 d = oc_provot_delta<M>(p1, p2, rest, &on);
    if (!on) return;
    if (pin1)      { p2.x = M::add(p2.x, d.x); p2.y = M::add(p2.y, d.y); p2.z = M::add(p2.z, d.z); }                 // V:498-499
    else if (pin2) { p1.x = M::sub(p1.x, d.x); p1.y = M::sub(p1.y, d.y); p1.z = M::sub(p1.z, d.z); }                 // V:500-501
    else {                                                                                                       // V:503-504
        p1.x = M::sub(p1.x, d.x); p1.y = M::sub(p1.y, d.y); p1.z = M::sub(p1.z, d.z);
        p2.x = M::add(p2.x, d.x); p2.y = M::add(p2.y, d.y); p2.z = M::add(p2.z, d.z);
    }
}

OC_HD f3 oc_ld3(const float4* p) { const float4 a = *p; return make_f3(a.x, a.y, a.z); }
OC_HD void oc_st3(float4* p, f3 v) { p->x = v.x; p->y = v.y; p->z = v.z; }      // w (collider flag / 1) untouched

// Before the in-place sweep of the Verlet form: particles the collider moved in this step carry X_last = X as a flag in
// w (oc_core.cuh).  The sweep is about to move X, so X_last is made explicit: X_last buffer <- X, flag cleared.
OC_HD void oc_provot_materialize(float4* X, float4* XL, long long t)
{
    float4 a = X[t];
    if (oc_hit(a.w)) { a.w = oc_u2f(OC_W_PLAIN); XL[t] = a; X[t] = a; }
}

// springs along row j of cloth b: reach 1 (structural, V:288-291) or reach 2 (bend, V:309-314, the row's last spring
// twice).  One thread walks the row; the particles it is working on live in registers.
template <class M, int REACH>
OC_HD void oc_provot_row(const OcConst& c, float4* X, int b, int j)
{
    const int U = c.U;
    float4* row = X + oc_index(c, b, 0, j);
    const float* rest = REACH == 1 ? c.rh1 : c.rh2;
    if (REACH == 1) {
        f3 p = oc_ld3(row);
        for (int i = 0; i + 1 < U; ++i) {
            f3 q = oc_ld3(row + i + 1);
            oc_provot_x_spring<M>(p, q, oc_pinned(c, b, i, j), oc_pinned(c, b, i + 1, j), rest[i]);
            oc_st3(row + i, p);
            p = q;
        }
        oc_st3(row + U - 1, p);
    } else {
        f3 p0 = oc_ld3(row), p1 = oc_ld3(row + 1);
        for (int i = 0; i + 2 < U; ++i) {
            f3 q = oc_ld3(row + i + 2);
            oc_provot_x_spring<M>(p0, q, oc_pinned(c, b, i, j), oc_pinned(c, b, i + 2, j), rest[i]);
            if (i == U - 3) oc_provot_x_spring<M>(p0, q, oc_pinned(c, b, i, j), oc_pinned(c, b, i + 2, j), rest[i]);       // V:313
            oc_st3(row + i, p0);
            p0 = p1; p1 = q;
        }
        oc_st3(row + U - 2, p0); oc_st3(row + U - 1, p1);
    }
}
// springs along column i: reach 1 (V:294-297) or 2 (V:315-320).  One thread per column: coalesced across the warp.
template <class M, int REACH>
OC_HD void oc_provot_col(const OcConst& c, float4* X, int b, int i)
{
    const int U = c.U, V = c.V;
    float4* col = X + oc_index(c, b, i, 0);
    const float* rest = REACH == 1 ? c.rv1 : c.rv2;
    if (REACH == 1) {
        f3 p = oc_ld3(col);
        for (int j = 0; j + 1 < V; ++j) {
            f3 q = oc_ld3(col + (long long)(j + 1) * U);
            oc_provot_x_spring<M>(p, q, oc_pinned(c, b, i, j), oc_pinned(c, b, i, j + 1), rest[j]);
            oc_st3(col + (long long)j * U, p);
            p = q;
        }
        oc_st3(col + (long long)(V - 1) * U, p);
    } else {
        f3 p0 = oc_ld3(col), p1 = oc_ld3(col + U);
        for (int j = 0; j + 2 < V; ++j) {
            f3 q = oc_ld3(col + (long long)(j + 2) * U);
            oc_provot_x_spring<M>(p0, q, oc_pinned(c, b, i, j), oc_pinned(c, b, i, j + 2), rest[j]);
            if (j == V - 3) oc_provot_x_spring<M>(p0, q, oc_pinned(c, b, i, j), oc_pinned(c, b, i, j + 2), rest[j]);       // V:319
            oc_st3(col + (long long)j * U, p0);
            p0 = p1; p1 = q;
        }
        oc_st3(col + (long long)(V - 2) * U, p0); oc_st3(col + (long long)(V - 1) * U, p1);
    }
}
// shear springs of cloth ctx.bx(): thread r of a block of T cell rows takes cell row r and lags two cells behind thread
// r-1; blocks of T cell rows one after the other.  Every time step ends with a barrier, which also makes the positions
// written in global memory by the other threads of the CTA visible.  (Template over the execution context like the
// step kernels: tests/emu runs the same body on the CPU under several thread schedules.)
template <class M, class Ctx>
OC_HD void oc_provot_shear_body(Ctx& ctx, const OcConst& c, float4* X)
{
    const int b = ctx.bx(), U = c.U, V = c.V;
    const int T = ctx.nthreads(), l = ctx.tid();
    for (int base = 0; base < V - 1; base += T) {
        const int r = base + l;
        const bool live = r < V - 1;
        const int rows_here = (V - 1 - base) < T ? (V - 1 - base) : T;
        const int steps = (U - 1) + 2 * (rows_here - 1);
        float4* up = X + oc_index(c, b, 0, live ? r : 0);
        float4* dn = up + U;
        const float dz2 = live ? c.dz2[r] : 0.0f;
        for (int t = 0; t < steps; ++t) {
            const int col = t - 2 * l;
            if (live && col >= 0 && col < U - 1) {
                const float rest = M::sqrt(M::add(c.dx2[col], dz2));
                f3 a = oc_ld3(up + col), bq = oc_ld3(up + col + 1), cq = oc_ld3(dn + col), d = oc_ld3(dn + col + 1);
                // (col, r) -> (col+1, r+1), then (col, r+1) -> (col+1, r)            V:303-304
                oc_provot_x_spring<M>(a, d, oc_pinned(c, b, col, r), oc_pinned(c, b, col + 1, r + 1), rest);
                oc_provot_x_spring<M>(cq, bq, oc_pinned(c, b, col, r + 1), oc_pinned(c, b, col + 1, r), rest);
                oc_st3(up + col, a); oc_st3(up + col + 1, bq); oc_st3(dn + col, cq); oc_st3(dn + col + 1, d);
            }
            ctx.sync();
        }
    }
}

#ifdef __CUDACC__
template <class M>
__global__ void __launch_bounds__(128)
oc_k_provot_v(OcConst c, const float4* __restrict__ X, const float4* Vin, float4* Vout)      // Vin may be Vout: a thread touches its own V only
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, b = blockIdx.z;
    if (i >= c.U) return;
    Vout[oc_index(c, b, i, j)] = oc_provot_v_particle<M>(c, X, Vin, b, i, j);
}
__global__ void oc_k_provot_materialize(OcConst c, float4* __restrict__ X, float4* __restrict__ XL)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < c.cloth_stride * c.batch) oc_provot_materialize(X, XL, t);
}
template <class M, int REACH>
__global__ void __launch_bounds__(128)
oc_k_provot_rows(OcConst c, float4* X)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < c.V) oc_provot_row<M, REACH>(c, X, blockIdx.y, j);
}
template <class M, int REACH>
__global__ void __launch_bounds__(128)
oc_k_provot_cols(OcConst c, float4* X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.U) oc_provot_col<M, REACH>(c, X, blockIdx.y, i);
}
struct OcProvotCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ int bx() const { return blockIdx.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
template <class M>
__global__ void __launch_bounds__(1024)
oc_k_provot_shear(OcConst c, float4* X)      // one CTA per cloth
{
    OcProvotCtx ctx;
    oc_provot_shear_body<M, OcProvotCtx>(ctx, c, X);
}
#endif
