// oc_march2.cuh — kernel 3: the marching stencil kernel with TWO COLUMNS PER THREAD (one substep per launch).
//
// Same algorithm, data flow and arithmetic as oc_march.cuh (read that header first); what changes is the
// mapping of work to threads, chosen to cut the instruction-issue and shared-memory load that bound the
// one-column kernel (profiles/r1_march_*):
//   * A thread owns two adjacent columns a = 2i, b = 2i+1 of the window.  The packed FP32x2 pairs are
//     now (particle a, particle b) for ONE spring type, so the own-column window, the (+2,0) partners and
//     every received (+2,0) force are naturally aligned 64-bit shared-memory accesses.
//   * Three of the twelve springs of a thread are internal (a-b, a-b', b-a'): their partner force never
//     touches shared memory.  Per particle the kernel issues ~20 shared loads instead of 55 and ~10
//     stores instead of 19.
//   * The force accumulation, the base force, the integration and the collider test run as packed
//     operations over the two particles.
// The accumulation ORDER per particle is unchanged (the reference's spring-list order), and so is every
// rounding: exact mode stays bit-identical.
//
// Kernel: S = 1 (no temporal blocking; k > 1 is served by oc_k_march).  WC window columns per CTA,
// WC/2 threads.  The steady loop (interior rows) has no predicates: in CTAs at a cloth edge the springs of
// columns that do not exist are multiplied by zero instead.  Edge ROWS (and the pipeline fill of a tile) use a
// generic, per-half predicated path.  A launch is a 1-D grid of tiles (OcSeg2); consecutive launches are
// chained tile by tile instead of by a grid-wide barrier (OcDep2).
#pragma once
#include "oc_core.cuh"
#include "oc_march.cuh"


template <int WC>
struct OcSmem2 {
    float X[6][OC_RING][WC + 4];        // x, y, z, vx, vy, vz     slot = row & 3, index = window column + 2
    float Dd[3][OC_RING][WC + 4];       // X - X_last
    float FH2[3][2][WC + 4];            // f(+2,0) of every column, slot = row & 1
    float FH1[3][2][WC / 2 + 2];        // f(+1,0) of each thread's b column, index = thread + 1
    float FDb[3][OC_RING][WC / 2 + 2];  // f(+1,+1) of the b column
    float FAa[3][OC_RING][WC / 2 + 2];  // f(-1,+1) of the a column
    float4 stage[4][WC / 2];            // landing zone of the asynchronous row loads: A[a], B[a], A[b], B[b] per thread
};

// Linked row bands (multi-GPU): the cloth is cut into row bands, one handle per GPU, and the bands are
// LINKED through peer memory (CUDA IPC across processes, plain peer access inside one).  There is no halo exchange
// step: the tiles at a band boundary store the two rows the neighbour's stencil reaches (bend springs, reach 2:
// V:311, V:317) straight into the neighbour's halo rows as they produce them (NVLink stores), and the per-tile
// dependency flags span the GPUs: after its last store a boundary tile releases one word per strip in the NEIGHBOUR's
// memory (system scope); the neighbour's boundary tiles of the next step poll their local copy.  Steps on different
// GPUs are thereby ordered tile by tile, without the host and without a grid-wide or machine-wide barrier.
// The values below are needed on the generic path and after the steady loop only.  Read as plain kernel parameters
// ptxas loads them into uniform registers at kernel entry, where they stay live across the steady loop and push the
// loop's own constants out of the uniform register file (measured: 33 extra LDCU per iteration, -5 %).  A pointer the
// compiler cannot see through makes them ordinary loads at the point of use.
template <class T> OC_HD const T* oc_opaque(const T* p)
{
#ifdef __CUDA_ARCH__
    asm volatile("" : "+l"(p));
#endif
    return p;
}
struct OcPeer2 {
    float4*         c[2];          // [0] upper, [1] lower neighbour: its destination buffer of this step, biased so that
                                   // element row * U + column is that particle's slot; nullptr = cloth edge / not linked
    unsigned*       flags_out[2];  // words [strip] in the neighbour's memory that this band's boundary tiles release
    const unsigned* flags_in[2];   // local words [strip] released by the neighbour's boundary tiles
    unsigned        epoch;         // linked steps taken, this one included (the same number on every band)
    int             ra, rb;        // rows of this launch (= the band's owned rows)
    int             nstrips;
    int             rev;           // this band launches its segments bottom to top.  Neighbouring bands alternate: a band's boundary
                                   // tiles then are the FIRST of its step on one side and the LAST on the other, and so are its
                                   // neighbour's on the facing side, which leaves every cross-GPU dependency about a whole step of
                                   // slack (with one direction everywhere the lower band's first tiles wait for the upper band's last
                                   // ones, and the bands drift apart by up to 1.8 steps per boundary before the reverse dependency holds them)
};

// ---- spring pair with a pair-valued first end (particles a and b of the thread) ----------------------
template <class M>
OC_HD OcPair3 oc_spring2v(const OcPair3& px, const OcPair3& pv, const OcPair3& qx, const OcPair3& qv,
                          float2 rest, float2 nks, float2 kd, float one, OcRange& rg)
{
    OcPair3 dp, dv, f;
    dp.x = p_sub(px.x, qx.x); dp.y = p_sub(px.y, qx.y); dp.z = p_sub(px.z, qx.z);                     // V:471
    dv.x = p_sub(pv.x, qv.x); dv.y = p_sub(pv.y, qv.y); dv.z = p_sub(pv.z, qv.z);                     // V:472
    if (M::kExact) {
        const float2 sqr  = p_sump<M>(p_mul(dp.z, dp.z), p_sump<M>(p_mul(dp.y, dp.y), p_mul(dp.x, dp.x), one), one);
        const float2 dist = oc_sqrt2<M>(sqr, rg);                                                    // V:473
#ifdef __CUDA_ARCH__
        const float2 y0  = p_rcp(dist);
        const float2 inv = p_fma(y0, p_fma(y0, p_neg(dist), p_bc(1.0f)), y0);
        const float2 a   = p_sump<M>(p_mul(dv.z, dp.z), p_sump<M>(p_mul(dv.y, dp.y), p_mul(dv.x, dp.x), one), one);
        rg.num(a.x); rg.num(a.y);
        const float2 q0  = p_mul(a, inv);
        const float2 q   = p_fma(inv, p_fma(q0, p_neg(dist), a), q0);
#else
        const float2 inv = make_float2(1.0f / dist.x, 1.0f / dist.y);
        const float2 a   = p_sump<M>(p_mul(dv.z, dp.z), p_sump<M>(p_mul(dv.y, dp.y), p_mul(dv.x, dp.x), one), one);
        const float2 q   = make_float2(a.x / dist.x, a.y / dist.y);
#endif
        const float2 left  = p_mul(nks, p_sub(dist, rest));                                          // V:475
        const float2 right = p_mul(kd, q);                                                           // V:476
        const float2 s = p_sump<M>(right, left, one);
        f.x = p_mul(s, p_mul(dp.x, inv)); f.y = p_mul(s, p_mul(dp.y, inv)); f.z = p_mul(s, p_mul(dp.z, inv));   // V:477
    } else {
        const float2 sqr  = p_fma(dp.z, dp.z, p_fma(dp.y, dp.y, p_mul(dp.x, dp.x)));
        const float2 rinv = p_rsq(sqr);
        const float2 dist = p_mul(sqr, rinv);
        const float2 left = p_fma(nks, dist, p_neg(rest));                      // rest pre-multiplied by nks
        const float2 dot  = p_fma(dv.z, dp.z, p_fma(dv.y, dp.y, p_mul(dv.x, dp.x)));
        const float2 s    = p_mul(p_fma(p_mul(kd, dot), rinv, left), rinv);
        f.x = p_mul(s, dp.x); f.y = p_mul(s, dp.y); f.z = p_mul(s, dp.z);
    }
    return f;
}

// Cold paths.  They are kept OUT OF LINE (noinline, scalar arguments and results in registers, operands re-read from
// shared memory, constants through a pointer to the __grid_constant__ kernel parameter): the steady loop of exact
// mode is bound by instruction issue and instruction-cache reach, and inlined the three of them were 516 of its
// 1183 instructions.
#ifdef __CUDA_ARCH__
#define OC_COLD __device__ __noinline__
#else
#define OC_COLD inline
#endif
// One spring pair of the thread redone with the IEEE intrinsics.
// kind: 0 (+1,0)  1 (+2,0)  2 (0,+1)  3 (0,+2)  4 (+1,+1)  5 (-1,+1); for 4 and 5 rest_* are the SQUARED rest lengths.
template <class M, int WC>
OC_COLD OcPair3 oc_march2_redo(const OcConst* c, const OcSmem2<WC>* s, int kind, int sl, int pa, float rest_a, float rest_b)
{
    const int s1 = (sl + 1) & (OC_RING - 1), s2 = (sl + 2) & (OC_RING - 1);
#define OC_LDX(slot, col) make_f3(s->X[0][slot][col], s->X[1][slot][col], s->X[2][slot][col])
#define OC_LDV(slot, col) make_f3(s->X[3][slot][col], s->X[4][slot][col], s->X[5][slot][col])
    // partner of a / of b: (slot, column)
    int sa = sl, ca = pa, sb = sl, cb = pa;
    float nks = c->nks_struct, kd = c->kd_struct;
    switch (kind) {
    case 0: sa = sl; ca = pa + 1; sb = sl; cb = pa + 2; break;
    case 1: sa = sl; ca = pa + 2; sb = sl; cb = pa + 3; nks = c->nks_bend; kd = c->kd_bend; break;
    case 2: sa = s1; ca = pa;     sb = s1; cb = pa + 1; break;
    case 3: sa = s2; ca = pa;     sb = s2; cb = pa + 1; nks = c->nks_bend; kd = c->kd_bend; break;
    case 4: sa = s1; ca = pa + 1; sb = s1; cb = pa + 2; nks = c->nks_shear; kd = c->kd_shear; break;
    default: sa = s1; ca = pa - 1; sb = s1; cb = pa;    nks = c->nks_shear; kd = c->kd_shear; break;
    }
    if (kind >= 4) { rest_a = M::sqrt(rest_a); rest_b = M::sqrt(rest_b); }
    const f3 fa = oc_spring<M>(OC_LDX(sl, pa),     OC_LDV(sl, pa),     OC_LDX(sa, ca), OC_LDV(sa, ca), rest_a, nks, kd);
    const f3 fb = oc_spring<M>(OC_LDX(sl, pa + 1), OC_LDV(sl, pa + 1), OC_LDX(sb, cb), OC_LDV(sb, cb), rest_b, nks, kd);
#undef OC_LDX
#undef OC_LDV
    OcPair3 f;
    f.x = make_float2(fa.x, fb.x); f.y = make_float2(fa.y, fb.y); f.z = make_float2(fa.z, fb.z);
    return f;
}
// (X - X_last) / dt of both particles with the IEEE division (an operand left the range of the branch-free form)
template <class M>
OC_COLD OcPair3 oc_march2_vel_slow(OcPair3 d, float dt)
{
    OcPair3 v;
    v.x = make_float2(M::div(d.x.x, dt), M::div(d.x.y, dt));
    v.y = make_float2(M::div(d.y.x, dt), M::div(d.y.y, dt));
    v.z = make_float2(M::div(d.z.x, dt), M::div(d.z.y, dt));
    return v;
}
// EllipsoidCollision of one particle that is inside the collider (V:514-530): p0 = X_0 - center, sq = dot(p0, p0),
// n = the integrated position; returns the projected position.
template <class M>
OC_COLD f3 oc_march2_collide(const OcConst* c, f3 d0, float sq, f3 n)
{
    const float distance = M::sqrt(sq);
    const float sc = M::sub(c->radius, distance);                                    // V:515
    if (M::kExact) d0 = make_f3(M::div(M::mul(sc, d0.x), distance), M::div(M::mul(sc, d0.y), distance), M::div(M::mul(sc, d0.z), distance));
    else { const float q = M::div(sc, distance); d0 = make_f3(q * d0.x, q * d0.y, q * d0.z); }
    const float ddx = M::dot(d0, make_f3(c->tinv[0][0], c->tinv[0][1], c->tinv[0][2]));     // V:520-528
    const float ddy = M::dot(d0, make_f3(c->tinv[1][0], c->tinv[1][1], c->tinv[1][2]));
    const float ddz = M::dot(d0, make_f3(c->tinv[2][0], c->tinv[2][1], c->tinv[2][2]));
    return make_f3(M::add(n.x, ddx), M::add(n.y, ddy), M::add(n.z, ddz));
}

// F (+|-)= g for both particles, or per half under predicates
// (g is a packed product s * n: p_sump / p_subp keep ptxas from contracting it into the accumulation)
template <class M, bool kAll> OC_HD void oc_acc2(OcPair3& F, const OcPair3& g, bool pa, bool pb, bool sub, float one)
{
    if (kAll) {
        if (sub) { F.x = p_subp<M>(F.x, g.x, one); F.y = p_subp<M>(F.y, g.y, one); F.z = p_subp<M>(F.z, g.z, one); }
        else     { F.x = p_sump<M>(g.x, F.x, one); F.y = p_sump<M>(g.y, F.y, one); F.z = p_sump<M>(g.z, F.z, one); }
    } else {
        if (pa) {
            if (sub) { F.x.x = M::sub(F.x.x, g.x.x); F.y.x = M::sub(F.y.x, g.y.x); F.z.x = M::sub(F.z.x, g.z.x); }
            else     { F.x.x = M::add(F.x.x, g.x.x); F.y.x = M::add(F.y.x, g.y.x); F.z.x = M::add(F.z.x, g.z.x); }
        }
        if (pb) {
            if (sub) { F.x.y = M::sub(F.x.y, g.x.y); F.y.y = M::sub(F.y.y, g.y.y); F.z.y = M::sub(F.z.y, g.z.y); }
            else     { F.x.y = M::add(F.x.y, g.x.y); F.y.y = M::add(F.y.y, g.y.y); F.z.y = M::add(F.z.y, g.z.y); }
        }
    }
}

template <class M, int WC, class Ctx>
struct OcMarch2 {
    typedef OcSmem2<WC> Smem;
    static constexpr int T = WC / 2;
    Ctx& ctx;
    const OcConst& c;
    const float4* __restrict__ A; const float4* __restrict__ B;
    float4* __restrict__ C;
    Smem* sm;
    int i, pa, ga, U, V;
    int lo, hi, plo, in_lo, in_hi, first, row0;
    bool oka, okb, sta, stb;                 // column exists / column is stored by this CTA
    float2 rh1, rh2, dx2ab, dx2ma;           // (rh1[ga], rh1[gb]) ... (dx2[ga-1], dx2[ga])
    float ydt;
    float rv1_n, rv2_n, dz2_n;
    long long goff;                          // element offset of (cloth, ga, row 0)
    // carried from iteration to iteration as 64-bit pairs (see oc_q2):
    OcPair3q me_x, me_v, w1_x, w1_v;         // own columns, rows row and row+1
    OcPair3q k1_q, k2a_q, k2b_q;             // carried (0,+1) of row-1, (0,+2) of row-1 and row-2
    f3 kDa, kAb;                             // carried internal shear forces of row-1: f(a->b'), f(b->a')
    const OcPeer2* peer;                     // linked row bands (kernel-parameter space; read on the generic path only)

    OC_HD OcMarch2(Ctx& ctx_, const OcConst& c_) : ctx(ctx_), c(c_) {}

    OC_HD OcPV2 ld_own(int slot) const
    {
        OcPV2 r;
        r.x.x = *reinterpret_cast<const float2*>(&sm->X[0][slot][pa]); r.x.y = *reinterpret_cast<const float2*>(&sm->X[1][slot][pa]);
        r.x.z = *reinterpret_cast<const float2*>(&sm->X[2][slot][pa]); r.v.x = *reinterpret_cast<const float2*>(&sm->X[3][slot][pa]);
        r.v.y = *reinterpret_cast<const float2*>(&sm->X[4][slot][pa]); r.v.z = *reinterpret_cast<const float2*>(&sm->X[5][slot][pa]);
        return r;
    }

    // One loaded row (both columns) -> position, velocity (X - X_last)/dt and X - X_last, once, into ring slot sl
    OC_HD void publish(int sl, const float4 laa, const float4 lqa, const float4 lab, const float4 lqb)
    {
        Smem& s = *sm;
        OcPair3 d;
        d.x = make_float2(M::sub(laa.x, lqa.x), M::sub(lab.x, lqb.x));
        d.y = make_float2(M::sub(laa.y, lqa.y), M::sub(lab.y, lqb.y));
        d.z = make_float2(M::sub(laa.z, lqa.z), M::sub(lab.z, lqb.z));
        if (oc_hit(laa.w)) { d.x.x = 0.0f; d.y.x = 0.0f; d.z.x = 0.0f; }       // X_last == X (V:530)
        if (oc_hit(lab.w)) { d.x.y = 0.0f; d.y.y = 0.0f; d.z.y = 0.0f; }
        OcPair3 v;
#ifdef __CUDA_ARCH__
        if (M::kExact) {
            OcRangeStrict rv; rv.init();
            rv.add(d.x.x); rv.add(d.x.y); rv.add(d.y.x); rv.add(d.y.y); rv.add(d.z.x); rv.add(d.z.y);
            const bool badv = (c.dt_bf == 0) | rv.bad(OC_VEL_LO_BITS, OC_VEL_HI_BITS);
            const float2 y = p_bc(ydt), nd = p_bc(-c.dt);
            float2 q0 = p_mul(d.x, y); v.x = p_fma(y, p_fma(q0, nd, d.x), q0);
            q0 = p_mul(d.y, y);        v.y = p_fma(y, p_fma(q0, nd, d.y), q0);
            q0 = p_mul(d.z, y);        v.z = p_fma(y, p_fma(q0, nd, d.z), q0);
            if (__builtin_expect(badv, 0)) {
                if (c.dbg & 4) atomicAdd(c.dbg_cnt + 2, 1ull);
                v = oc_march2_vel_slow<M>(d, c.dt);
            }
        } else
#endif
        {
            if (M::kExact) {
                v.x = make_float2(d.x.x / c.dt, d.x.y / c.dt); v.y = make_float2(d.y.x / c.dt, d.y.y / c.dt); v.z = make_float2(d.z.x / c.dt, d.z.y / c.dt);
            } else {
                v.x = p_mul(d.x, p_bc(c.inv_dt)); v.y = p_mul(d.y, p_bc(c.inv_dt)); v.z = p_mul(d.z, p_bc(c.inv_dt));
            }
        }
        *reinterpret_cast<float2*>(&s.X[0][sl][pa]) = make_float2(laa.x, lab.x);          // lrow = row + 4: same slot
        *reinterpret_cast<float2*>(&s.X[1][sl][pa]) = make_float2(laa.y, lab.y);
        *reinterpret_cast<float2*>(&s.X[2][sl][pa]) = make_float2(laa.z, lab.z);
        *reinterpret_cast<float2*>(&s.X[3][sl][pa]) = v.x;
        *reinterpret_cast<float2*>(&s.X[4][sl][pa]) = v.y;
        *reinterpret_cast<float2*>(&s.X[5][sl][pa]) = v.z;
        *reinterpret_cast<float2*>(&s.Dd[0][sl][pa]) = d.x;
        *reinterpret_cast<float2*>(&s.Dd[1][sl][pa]) = d.y;
        *reinterpret_cast<float2*>(&s.Dd[2][sl][pa]) = d.z;
    }

    template <bool kSteady, bool kInterior>
    OC_HD void iter(int it)
    {
        Smem& s = *sm;
        const int row = row0 + it;
        const int lrow = first + it;
        OcPV2 me, w1;
        me.x = p_unpack3(me_x); me.v = p_unpack3(me_v); w1.x = p_unpack3(w1_x); w1.v = p_unpack3(w1_v);
        const OcPair3 k1 = p_unpack3(k1_q), k2b = p_unpack3(k2b_q);
        // ---- asynchronous global loads of row lrow (both columns) into the thread's landing zone ------
        // (columns of the window outside the cloth get a benign far-away particle at rest)
        const bool doL = kSteady || (lrow >= in_lo && lrow < in_hi);
        if (doL) {
            const long long o = goff + (long long)lrow * U;
            if (kInterior || oka) { oc_cp_async16(&s.stage[0][i], A + o); oc_cp_async16(&s.stage[1][i], B + o); }
            else s.stage[0][i] = s.stage[1][i] = make_float4(1.0e3f + 8.0f * (float)pa, 1.0e3f, 1.0e3f + 8.0f * (float)(lrow & 63), oc_u2f(OC_W_PLAIN));
            if (kInterior || okb) { oc_cp_async16(&s.stage[2][i], A + o + 1); oc_cp_async16(&s.stage[3][i], B + o + 1); }
            else s.stage[2][i] = s.stage[3][i] = make_float4(1.0e3f + 8.0f * (float)(pa + 1), 1.0e3f, 1.0e3f + 8.0f * (float)(lrow & 63), oc_u2f(OC_W_PLAIN));
            oc_cp_async_commit();
        }
        const float rv1_j = rv1_n, rv2_j = rv2_n, dz2_j = dz2_n;
        {
            int r = row + 1;
            if (!kSteady) r = r < 0 ? 0 : (r >= V ? V - 1 : r);
            rv1_n = OC_LDG(c.rv1 + r); rv2_n = OC_LDG(c.rv2 + r); dz2_n = OC_LDG(c.dz2 + r);
        }
        const int sl = row & (OC_RING - 1);
        const int s1 = (sl + 1) & (OC_RING - 1), s2 = (sl + 2) & (OC_RING - 1), s3 = (sl + 3) & (OC_RING - 1);
        const int h = sl & 1;

        // ---- P phase ---------------------------------------------------------------------------------
        const bool doP = kSteady || (row >= plo && row < hi);
        OcPair3 gH1, gH2, gV1, gV2, gD, gA, dme;
        OcPV2 w2;
        if (doP) {
            if (!kSteady && row == plo) { me = ld_own(sl); w1 = ld_own(s1); }
            w2 = ld_own(s2);
            dme.x = *reinterpret_cast<const float2*>(&s.Dd[0][sl][pa]);
            dme.y = *reinterpret_cast<const float2*>(&s.Dd[1][sl][pa]);
            dme.z = *reinterpret_cast<const float2*>(&s.Dd[2][sl][pa]);
            // partners of the next thread's columns, this row: (a_n, b_n)
            OcPV2 n0;
            n0.x.x = *reinterpret_cast<const float2*>(&s.X[0][sl][pa + 2]); n0.x.y = *reinterpret_cast<const float2*>(&s.X[1][sl][pa + 2]);
            n0.x.z = *reinterpret_cast<const float2*>(&s.X[2][sl][pa + 2]); n0.v.x = *reinterpret_cast<const float2*>(&s.X[3][sl][pa + 2]);
            n0.v.y = *reinterpret_cast<const float2*>(&s.X[4][sl][pa + 2]); n0.v.z = *reinterpret_cast<const float2*>(&s.X[5][sl][pa + 2]);
            // shifted pairs: (b, a_n) this row; (b', a_n') and (b_prev', a') next row
            OcPV2 qH1, qD, qA;
            qH1.x.x = make_float2(me.x.x.y, n0.x.x.x); qH1.x.y = make_float2(me.x.y.y, n0.x.y.x); qH1.x.z = make_float2(me.x.z.y, n0.x.z.x);
            qH1.v.x = make_float2(me.v.x.y, n0.v.x.x); qH1.v.y = make_float2(me.v.y.y, n0.v.y.x); qH1.v.z = make_float2(me.v.z.y, n0.v.z.x);
            qD.x.x = make_float2(w1.x.x.y, s.X[0][s1][pa + 2]); qD.x.y = make_float2(w1.x.y.y, s.X[1][s1][pa + 2]); qD.x.z = make_float2(w1.x.z.y, s.X[2][s1][pa + 2]);
            qD.v.x = make_float2(w1.v.x.y, s.X[3][s1][pa + 2]); qD.v.y = make_float2(w1.v.y.y, s.X[4][s1][pa + 2]); qD.v.z = make_float2(w1.v.z.y, s.X[5][s1][pa + 2]);
            qA.x.x = make_float2(s.X[0][s1][pa - 1], w1.x.x.x); qA.x.y = make_float2(s.X[1][s1][pa - 1], w1.x.y.x); qA.x.z = make_float2(s.X[2][s1][pa - 1], w1.x.z.x);
            qA.v.x = make_float2(s.X[3][s1][pa - 1], w1.v.x.x); qA.v.y = make_float2(s.X[4][s1][pa - 1], w1.v.y.x); qA.v.z = make_float2(s.X[5][s1][pa - 1], w1.v.z.x);

            // exact mode: the operand ranges of the branch-free sqrt / division sequences, accumulated over the
            // six spring pairs and the two shear rest lengths (OcRange); one test for the whole iteration
            OcRange rg; rg.init();
            float2 rD = oc_sqrt2<M>(p_add(dx2ab, p_bc(dz2_j)), rg);         // cells (ga, row), (gb, row)
            float2 rA = oc_sqrt2<M>(p_add(dx2ma, p_bc(dz2_j)), rg);         // cells (ga-1, row), (ga, row)
            float2 rH1 = rh1, rH2 = rh2, rV1 = p_bc(rv1_j), rV2 = p_bc(rv2_j);
            const float2 nS = p_bc(c.nks_struct), kS = p_bc(c.kd_struct), nB = p_bc(c.nks_bend), kB = p_bc(c.kd_bend);
            const float2 nSh = p_bc(c.nks_shear), kSh = p_bc(c.kd_shear);
            if (!M::kExact) { rH1 = p_mul(rH1, nS); rH2 = p_mul(rH2, nB); rV1 = p_mul(rV1, nS); rV2 = p_mul(rV2, nB); rD = p_mul(rD, nSh); rA = p_mul(rA, nSh); }
            gH1 = oc_spring2v<M>(me.x, me.v, qH1.x, qH1.v, rH1, nS, kS, c.one, rg);
            gH2 = oc_spring2v<M>(me.x, me.v, n0.x,  n0.v,  rH2, nB, kB, c.one, rg);
            gV1 = oc_spring2v<M>(me.x, me.v, w1.x,  w1.v,  rV1, nS, kS, c.one, rg);
            gV2 = oc_spring2v<M>(me.x, me.v, w2.x,  w2.v,  rV2, nB, kB, c.one, rg);
            gD  = oc_spring2v<M>(me.x, me.v, qD.x,  qD.v,  rD,  nSh, kSh, c.one, rg);
            gA  = oc_spring2v<M>(me.x, me.v, qA.x,  qA.v,  rA,  nSh, kSh, c.one, rg);
            if (__builtin_expect(M::kExact && rg.bad(), 0)) {
                // rare: an operand left the exact range of the branch-free sequences -> all six pairs again with the
                // IEEE intrinsics (cold, out of line; operands re-read from shared memory)
#ifdef __CUDA_ARCH__
                if (c.dbg & 4) {                                            // development counters (OC_DEBUG=4)
                    atomicAdd(c.dbg_cnt, 1ull);
                    if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 1, 1ull);
                }
#endif
                gH1 = oc_march2_redo<M, WC>(&c, sm, 0, sl, pa, rh1.x, rh1.y);
                gH2 = oc_march2_redo<M, WC>(&c, sm, 1, sl, pa, rh2.x, rh2.y);
                gV1 = oc_march2_redo<M, WC>(&c, sm, 2, sl, pa, rv1_j, rv1_j);
                gV2 = oc_march2_redo<M, WC>(&c, sm, 3, sl, pa, rv2_j, rv2_j);
                gD  = oc_march2_redo<M, WC>(&c, sm, 4, sl, pa, M::add(dx2ab.x, dz2_j), M::add(dx2ab.y, dz2_j));
                gA  = oc_march2_redo<M, WC>(&c, sm, 5, sl, pa, M::add(dx2ma.x, dz2_j), M::add(dx2ma.y, dz2_j));
            }
            if (!kInterior || !kSteady) {
                // Window columns at a cloth edge: a spring to (or from) a column that does not exist is multiplied
                // by 0, every other one by 1 (exact).  The zero force then flows through the unpredicated packed
                // accumulation of the steady loop unchanged: F + (+-0) == F because an accumulator is never -0
                // (oc_core.cuh).  The ghost ends are finite far-away particles, so 0 * f is a zero, not a NaN.
                const int gb_ = ga + 1;
                const float2 mH1 = make_float2(oka && okb ? 1.0f : 0.0f, okb && gb_ + 1 < U ? 1.0f : 0.0f);   // a-b, b-a_next (also (+1,+1))
                const float2 mH2 = make_float2(oka && ga + 2 < U ? 1.0f : 0.0f, okb && gb_ + 2 < U ? 1.0f : 0.0f);
                const float2 mA  = make_float2(oka && ga - 1 >= 0 ? 1.0f : 0.0f, oka && okb ? 1.0f : 0.0f);    // a-b_prev', b-a'
                gH1.x = p_mul(gH1.x, mH1); gH1.y = p_mul(gH1.y, mH1); gH1.z = p_mul(gH1.z, mH1);
                gH2.x = p_mul(gH2.x, mH2); gH2.y = p_mul(gH2.y, mH2); gH2.z = p_mul(gH2.z, mH2);
                gD.x  = p_mul(gD.x,  mH1); gD.y  = p_mul(gD.y,  mH1); gD.z  = p_mul(gD.z,  mH1);
                gA.x  = p_mul(gA.x,  mA);  gA.y  = p_mul(gA.y,  mA);  gA.z  = p_mul(gA.z,  mA);
            }
            // publish the forces whose partner lives in another thread
            s.FH1[0][h][i + 1] = gH1.x.y; s.FH1[1][h][i + 1] = gH1.y.y; s.FH1[2][h][i + 1] = gH1.z.y;
            *reinterpret_cast<float2*>(&s.FH2[0][h][pa]) = gH2.x; *reinterpret_cast<float2*>(&s.FH2[1][h][pa]) = gH2.y; *reinterpret_cast<float2*>(&s.FH2[2][h][pa]) = gH2.z;
            s.FDb[0][sl][i + 1] = gD.x.y; s.FDb[1][sl][i + 1] = gD.y.y; s.FDb[2][sl][i + 1] = gD.z.y;
            s.FAa[0][sl][i + 1] = gA.x.x; s.FAa[1][sl][i + 1] = gA.y.x; s.FAa[2][sl][i + 1] = gA.z.x;
        }

        ctx.sync();

        // ---- G phase ---------------------------------------------------------------------------------
        const bool doG = kSteady || (row >= lo && row < hi);
        if (doG) {
            constexpr bool kAll = kSteady;           // no predicates: edge columns are handled by the zero masks above
            const int gb = ga + 1;
            const bool pin_a = !kSteady && oc_pinned(c, ctx.bz(), ga, row), pin_b = !kSteady && oc_pinned(c, ctx.bz(), gb, row);
            const bool ea = !pin_a, eb = !pin_b;                                  // springs act on the particle
            // existence of the horizontal neighbours of a and of b
            const bool al1 = kAll || ga - 1 >= 0, al2 = kAll || ga - 2 >= 0, ar1 = kAll || gb < U, ar2 = kAll || ga + 2 < U;
            const bool bl1 = kAll || ga >= 0,     bl2 = kAll || ga - 1 >= 0, br1 = kAll || gb + 1 < U, br2 = kAll || gb + 2 < U;
            const bool up1 = kSteady || row - 1 >= 0, up2 = kSteady || row - 2 >= 0, dn1 = kSteady || row + 1 < V, dn2 = kSteady || row + 2 < V;
            // F = 0 + gravity*mass (unless pinned) + DEFAULT_DAMPING*V     V:451-459
            OcPair3 F;
            F.x = make_float2(pin_a ? 0.0f : c.f0[0], pin_b ? 0.0f : c.f0[0]);
            F.y = make_float2(pin_a ? 0.0f : c.f0[1], pin_b ? 0.0f : c.f0[1]);
            F.z = make_float2(pin_a ? 0.0f : c.f0[2], pin_b ? 0.0f : c.f0[2]);
            F.x = p_sump<M>(p_mul(p_bc(c.damping), me.v.x), F.x, c.one);
            F.y = p_sump<M>(p_mul(p_bc(c.damping), me.v.y), F.y, c.one);
            F.z = p_sump<M>(p_mul(p_bc(c.damping), me.v.z), F.z, c.one);
            // 1  (i-1, j): a <- b of the previous thread (shared), b <- a (own pair, first half)
            {
                const bool p = ea && al1, q = eb && bl1;
                if (kAll || p) { F.x.x = M::sub(F.x.x, s.FH1[0][h][i]); F.y.x = M::sub(F.y.x, s.FH1[1][h][i]); F.z.x = M::sub(F.z.x, s.FH1[2][h][i]); }
                if (kAll || q) { F.x.y = M::sub(F.x.y, gH1.x.x); F.y.y = M::sub(F.y.y, gH1.y.x); F.z.y = M::sub(F.z.y, gH1.z.x); }
            }
            oc_acc2<M, kAll>(F, gH1, ea && ar1, eb && br1, false, c.one);                                  // 2  (i+1, j)
            oc_acc2<M, kAll>(F, k1,  ea && up1, eb && up1, true, c.one);                                   // 3  (i, j-1)
            oc_acc2<M, kAll>(F, gV1, ea && dn1, eb && dn1, false, c.one);                                  // 4  (i, j+1)
            // 5  (i-1, j-1): a <- b_prev at row-1 (shared), b <- a at row-1 (carried)
            {
                const bool p = ea && al1 && up1, q = eb && bl1 && up1;
                if (kAll || p) { F.x.x = M::sub(F.x.x, s.FDb[0][s3][i]); F.y.x = M::sub(F.y.x, s.FDb[1][s3][i]); F.z.x = M::sub(F.z.x, s.FDb[2][s3][i]); }
                if (kAll || q) { F.x.y = M::sub(F.x.y, kDa.x); F.y.y = M::sub(F.y.y, kDa.y); F.z.y = M::sub(F.z.y, kDa.z); }
            }
            // 6  (i+1, j-1): a <- b at row-1 (carried), b <- a_next at row-1 (shared)
            {
                const bool p = ea && ar1 && up1, q = eb && br1 && up1;
                if (kAll || p) { F.x.x = M::sub(F.x.x, kAb.x); F.y.x = M::sub(F.y.x, kAb.y); F.z.x = M::sub(F.z.x, kAb.z); }
                if (kAll || q) { F.x.y = M::sub(F.x.y, s.FAa[0][s3][i + 2]); F.y.y = M::sub(F.y.y, s.FAa[1][s3][i + 2]); F.z.y = M::sub(F.z.y, s.FAa[2][s3][i + 2]); }
            }
            oc_acc2<M, kAll>(F, gA, ea && al1 && dn1, eb && bl1 && dn1, false, c.one);                     // 7  (i-1, j+1)
            oc_acc2<M, kAll>(F, gD, ea && ar1 && dn1, eb && br1 && dn1, false, c.one);                     // 8  (i+1, j+1)
            OcPair3 r2;                                                                             // 9  (i-2, j)
            r2.x = *reinterpret_cast<const float2*>(&s.FH2[0][h][pa - 2]);
            r2.y = *reinterpret_cast<const float2*>(&s.FH2[1][h][pa - 2]);
            r2.z = *reinterpret_cast<const float2*>(&s.FH2[2][h][pa - 2]);
            oc_acc2<M, kAll>(F, r2,  ea && al2, eb && bl2, true, c.one);
            oc_acc2<M, kAll>(F, gH2, ea && ar2, eb && br2, false, c.one);                                  // 10 (i+2, j)
            if (!kAll) {                                                                            // 11 duplicated last bend spring of the row (V:313)
                oc_acc2<M, false>(F, gH2, ea && ga == U - 3, eb && gb == U - 3, false, c.one);
                oc_acc2<M, false>(F, r2,  ea && ga == U - 1, eb && gb == U - 1, true, c.one);
            } else if (!kInterior) {                                                                // same, as 0/1 multipliers
                const float2 dA = make_float2(ga == U - 3 ? 1.0f : 0.0f, gb == U - 3 ? 1.0f : 0.0f);
                const float2 dB = make_float2(ga == U - 1 ? 1.0f : 0.0f, gb == U - 1 ? 1.0f : 0.0f);
                F.x = p_sump<M>(p_mul(gH2.x, dA), F.x, c.one); F.y = p_sump<M>(p_mul(gH2.y, dA), F.y, c.one); F.z = p_sump<M>(p_mul(gH2.z, dA), F.z, c.one);
                F.x = p_subp<M>(F.x, p_mul(r2.x, dB), c.one);  F.y = p_subp<M>(F.y, p_mul(r2.y, dB), c.one);  F.z = p_subp<M>(F.z, p_mul(r2.z, dB), c.one);
            }
            oc_acc2<M, kAll>(F, k2b, ea && up2, eb && up2, true, c.one);                                   // 12 (i, j-2)
            oc_acc2<M, kAll>(F, gV2, ea && dn2, eb && dn2, false, c.one);                                  // 13 (i, j+2)
            if (!kSteady) {                                                                         // 14 duplicated last bend spring of the column (V:319)
                oc_acc2<M, false>(F, gV2, ea && row == V - 3, eb && row == V - 3, false, c.one);
                oc_acc2<M, false>(F, k2b, ea && row == V - 1, eb && row == V - 1, true, c.one);
            }
            // ---- IntegrateVerlet (V:428-444) + EllipsoidCollision (V:509-533), both particles ----------
            OcPair3 n;
            n.x = p_sump<M>(p_mul(p_bc(c.dt2m), F.x), p_add(me.x.x, dme.x), c.one);
            n.y = p_sump<M>(p_mul(p_bc(c.dt2m), F.y), p_add(me.x.y, dme.y), c.one);
            n.z = p_sump<M>(p_mul(p_bc(c.dt2m), F.z), p_add(me.x.z, dme.z), c.one);
            if (n.y.x < 0.0f) n.y.x = 0.0f;
            if (n.y.y < 0.0f) n.y.y = 0.0f;
            // A particle outside the collider's bounding sphere (OcConst::bs_*, conservative) cannot be inside the
            // ellipsoid: the transform of V:511-513 is skipped for it (most of the cloth, most of the time).
            bool hit_a = false, hit_b = false;
            const float2 ex = p_sub(n.x, p_bc(c.bs_c[0])), ey = p_sub(n.y, p_bc(c.bs_c[1])), ez = p_sub(n.z, p_bc(c.bs_c[2]));
            const float2 e2 = p_fma(ez, ez, p_fma(ey, ey, p_mul(ex, ex)));
            if ((e2.x <= c.bs_r2) | (e2.y <= c.bs_r2)) {
                OcPair3 p0;         // X_0 = inverse_ellipsoid * vec4(X,1) - center, rows x, y, z for (a, b)
                p0.x = p_sub(p_add(p_sump<M>(p_mul(p_bc(c.im[0][2]), n.z), p_sump<M>(p_mul(p_bc(c.im[0][1]), n.y), p_mul(p_bc(c.im[0][0]), n.x), c.one), c.one), p_bc(c.im[0][3])), p_bc(c.center[0]));
                p0.y = p_sub(p_add(p_sump<M>(p_mul(p_bc(c.im[1][2]), n.z), p_sump<M>(p_mul(p_bc(c.im[1][1]), n.y), p_mul(p_bc(c.im[1][0]), n.x), c.one), c.one), p_bc(c.im[1][3])), p_bc(c.center[1]));
                p0.z = p_sub(p_add(p_sump<M>(p_mul(p_bc(c.im[2][2]), n.z), p_sump<M>(p_mul(p_bc(c.im[2][1]), n.y), p_mul(p_bc(c.im[2][0]), n.x), c.one), c.one), p_bc(c.im[2][3])), p_bc(c.center[2]));
                const float2 sq = p_sump<M>(p_mul(p0.z, p0.z), p_sump<M>(p_mul(p0.y, p0.y), p_mul(p0.x, p0.x), c.one), c.one);
                hit_a = sq.x < 1.0f; hit_b = sq.y < 1.0f;                                           // V:513-514 (see oc_core.cuh)
#ifdef __CUDA_ARCH__
                if ((c.dbg & 4) && (hit_a | hit_b)) { atomicAdd(c.dbg_cnt + 3, 1ull); if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 3, 1ull << 32); }
#endif
                if (__builtin_expect(hit_a | hit_b, 0)) {
                    // EllipsoidCollision (V:514-530) of both particles at once, branch-free, with the same exact sqrt and
                    // division sequences as the springs; the result is taken per half where that particle is inside.
                    // (Contact zones are compact: whole CTAs spend every iteration here, so this path must be short.)
                    OcPair3 nn;
                    bool slow = false;
#ifdef __CUDA_ARCH__
                    if (M::kExact) {
                        OcRange rc; rc.init();
                        OcRangeStrict rn; rn.init();
                        const float2 distance = oc_sqrt2<M>(sq, rc);
                        const float2 sc = p_sub(p_bc(c.radius), distance);                                   // V:515
                        const float2 y0 = p_rcp(distance);
                        const float2 inv = p_fma(y0, p_fma(y0, p_neg(distance), p_bc(1.0f)), y0);            // 1/distance, correctly rounded
                        const float2 ax = p_mul(sc, p0.x), ay = p_mul(sc, p0.y), az = p_mul(sc, p0.z);
                        rn.add(ax.x); rn.add(ax.y); rn.add(ay.x); rn.add(ay.y); rn.add(az.x); rn.add(az.y);
                        float2 q0 = p_mul(ax, inv); const float2 dx = p_fma(inv, p_fma(q0, p_neg(distance), ax), q0);   // (sc*x0)/distance
                        q0 = p_mul(ay, inv);        const float2 dy = p_fma(inv, p_fma(q0, p_neg(distance), ay), q0);
                        q0 = p_mul(az, inv);        const float2 dz = p_fma(inv, p_fma(q0, p_neg(distance), az), q0);
                        // dot(d, transformInv row) = (dx*t0 + dy*t1) + dz*t2                                 V:520-528
                        nn.x = p_add(n.x, p_sump<M>(p_mul(dz, p_bc(c.tinv[0][2])), p_sump<M>(p_mul(dy, p_bc(c.tinv[0][1])), p_mul(dx, p_bc(c.tinv[0][0])), c.one), c.one));
                        nn.y = p_add(n.y, p_sump<M>(p_mul(dz, p_bc(c.tinv[1][2])), p_sump<M>(p_mul(dy, p_bc(c.tinv[1][1])), p_mul(dx, p_bc(c.tinv[1][0])), c.one), c.one));
                        nn.z = p_add(n.z, p_sump<M>(p_mul(dz, p_bc(c.tinv[2][2])), p_sump<M>(p_mul(dy, p_bc(c.tinv[2][1])), p_mul(dx, p_bc(c.tinv[2][0])), c.one), c.one));
                        // (a half that is not inside is tested too: if it trips the test the scalar path below still only
                        // touches the halves that are inside)
                        slow = rc.bad() | rn.bad(OC_NUM_LO_BITS, OC_NUM_HI_BITS);
                    } else
#endif
                    if (!M::kExact) {
                        const float2 rinv = p_rsq(sq);
                        const float2 q = p_mul(p_sub(p_bc(c.radius), p_mul(sq, rinv)), rinv);                // (radius - distance) / distance
                        const float2 dx = p_mul(q, p0.x), dy = p_mul(q, p0.y), dz = p_mul(q, p0.z);
                        nn.x = p_add(n.x, p_fma(dz, p_bc(c.tinv[0][2]), p_fma(dy, p_bc(c.tinv[0][1]), p_mul(dx, p_bc(c.tinv[0][0])))));
                        nn.y = p_add(n.y, p_fma(dz, p_bc(c.tinv[1][2]), p_fma(dy, p_bc(c.tinv[1][1]), p_mul(dx, p_bc(c.tinv[1][0])))));
                        nn.z = p_add(n.z, p_fma(dz, p_bc(c.tinv[2][2]), p_fma(dy, p_bc(c.tinv[2][1]), p_mul(dx, p_bc(c.tinv[2][0])))));
                    } else {
                        slow = true;                                      // host (emulator), exact mode: the scalar reference form
                    }
                    if (__builtin_expect(slow, 0)) {
                        if (hit_a) { const f3 r = oc_march2_collide<M>(&c, make_f3(p0.x.x, p0.y.x, p0.z.x), sq.x, make_f3(n.x.x, n.y.x, n.z.x)); nn.x.x = r.x; nn.y.x = r.y; nn.z.x = r.z; }
                        if (hit_b) { const f3 r = oc_march2_collide<M>(&c, make_f3(p0.x.y, p0.y.y, p0.z.y), sq.y, make_f3(n.x.y, n.y.y, n.z.y)); nn.x.y = r.x; nn.y.y = r.y; nn.z.y = r.z; }
                    }
                    if (hit_a) { n.x.x = nn.x.x; n.y.x = nn.y.x; n.z.x = nn.z.x; }
                    if (hit_b) { n.x.y = nn.x.y; n.y.y = nn.y.y; n.z.y = nn.z.y; }
                }
            }
            const long long o = goff + (long long)row * U;
            const float4 out_a = make_float4(n.x.x, n.y.x, n.z.x, oc_u2f(hit_a ? OC_W_HIT : OC_W_PLAIN));
            const float4 out_b = make_float4(n.x.y, n.y.y, n.z.y, oc_u2f(hit_b ? OC_W_HIT : OC_W_PLAIN));
            if (sta) C[o]     = out_a;
            if (stb) C[o + 1] = out_b;
            if (!kSteady) {
                // linked row bands: the first / last two rows of the band also go into the neighbour's halo (OcPeer2);
                // the steady range of a boundary tile excludes them, so the steady loop knows nothing of this
                const OcPeer2* pp = oc_opaque(peer);      // read at the point of use (see oc_opaque)
                float4* pc = nullptr;
                if (pp->c[0] && row < pp->ra + 2) pc = pp->c[0];
                if (pp->c[1] && row >= pp->rb - 2) pc = pp->c[1];
                if (pc) {
                    const long long po = (long long)row * U + ga;
                    if (sta) pc[po]     = out_a;
                    if (stb) pc[po + 1] = out_b;
                }
            }
        }
        if (doP) {
            k2b_q = k2a_q; k2a_q = p_pack3(gV2); k1_q = p_pack3(gV1);
            kDa = make_f3(gD.x.x, gD.y.x, gD.z.x);
            kAb = make_f3(gA.x.y, gA.y.y, gA.z.y);
            me_x = p_pack3(w1.x); me_v = p_pack3(w1.v); w1_x = p_pack3(w2.x); w1_v = p_pack3(w2.v);
        }

        // ---- publish the loaded row ------------------------------------------------------------------
        if (doL) {
            oc_cp_async_wait();
            publish(sl, s.stage[0][i], s.stage[1][i], s.stage[2][i], s.stage[3][i]);      // lrow = row + 4: same slot
        }
    }
};

#ifdef __CUDACC__
// development aid (OC_DEBUG=8): per-CTA time stamps  [0] entry  [1] set-up done  [2] lead-in done  [3] steady loop
// done  [4] exit, and [5] = SM id, at dbg_cnt[OC_DBG_TL_BASE + 8 * linear CTA index + k]
__device__ __forceinline__ void oc_timeline_mark(const OcConst& c, int k)
{
    const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (lin >= OC_DBG_TL_CTAS) return;
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    c.dbg_cnt[OC_DBG_TL_BASE + 8 * lin + k] = t;
    if (k == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); c.dbg_cnt[OC_DBG_TL_BASE + 8 * lin + 5] = sm; }
}
#endif

// Row segmentation of a launch: tile t of the 1-D grid -> (strip, segment) -> rows.  Uniform segments (rs_e = rs,
// n_extra = 0) are always valid.  In a SINGLE-WAVE launch the kernel ends when its slowest CTA does, and the CTAs
// of the first and last strip (masked edge path, half-empty last window) are the slow ones (time line in
// profiles/): the planner then gives the edge strips shorter segments (rs_e < rs) and a few more of them, so that
// all CTAs finish together.  The result does not depend on the segmentation: each row of a strip is computed by
// exactly one tile, from the same inputs.
struct OcSeg2 {
    int rs, rs_e;       // rows per segment: interior strips, edge strips
    int nstrips;
    int nseg_all;       // segments every strip has: tiles t < nstrips * nseg_all are (t % nstrips, t / nstrips)
    int n_extra;        // further tiles, alternately of the first and the last strip (the edge strips' additional segments)
    int rev;            // launch order of the segments: 0 top to bottom, 1 bottom to top (linked row bands alternate, see OcPeer2::rev);
                        // flag words are indexed by position (oc_seg2_index), whatever the launch order
};
OC_HD int oc_seg2_tiles(const OcSeg2& g) { return g.nstrips * g.nseg_all + g.n_extra; }
OC_HD void oc_seg2_tile(const OcSeg2& g, int t, int& bx, int& by)
{
    const int body = g.nstrips * g.nseg_all;
    if (t < body) { bx = t % g.nstrips; by = t / g.nstrips; if (g.rev) by = g.nseg_all - 1 - by; }
    else { const int e = t - body; bx = (e & 1) ? g.nstrips - 1 : 0; by = g.nseg_all + (e >> 1); }
}
OC_HD void oc_seg2_rows(const OcSeg2& g, int bx, int by, int ra, int rb, int& r0, int& r1)
{
    const int h = (bx == 0 || bx == g.nstrips - 1) ? g.rs_e : g.rs;
    r0 = ra + by * h;
    r1 = r0 + h;
    if (r0 > rb) r0 = rb;
    if (r1 > rb) r1 = rb;
}
// Completes g (rs, rs_e, nstrips given) so that every strip is covered.
inline void oc_seg2_finish(OcSeg2& g, int rows)
{
    const int ni = (rows + g.rs - 1) / g.rs, ne = (rows + g.rs_e - 1) / g.rs_e;
    if (g.nstrips <= 2 || ne <= ni) { g.nseg_all = g.nstrips <= 2 ? ne : ni; g.n_extra = 0; if (ne < ni) g.rs_e = g.rs; return; }
    g.nseg_all = ni; g.n_extra = 2 * (ne - ni);
}

// index of tile (strip, segment): its flag word, and with rev == 0 its place in the 1-D grid (inverse of oc_seg2_tile)
OC_HD int oc_seg2_index(const OcSeg2& g, int bx, int by)
{
    return by < g.nseg_all ? bx + g.nstrips * by : g.nstrips * g.nseg_all + 2 * (by - g.nseg_all) + (bx == 0 ? 0 : 1);
}

// What a launch has to wait for before it touches the state.  Consecutive steps are consecutive kernels on one
// stream, launched with programmatic stream serialization (the CTAs of step e+1 are placed while step e drains):
//   mode 0: griddepcontrol.wait, i.e. the whole previous grid;
//   mode 1: only the tiles of the previous launch that wrote what this tile reads: rows [r0-2, r1+2) of its own and
//           the two neighbouring strips.  Every tile publishes flags[its index] = epoch (release, after its last
//           store); a tile of the next launch polls the <= 12 flags it depends on (acquire).  This removes the
//           barrier between steps: a tile starts as soon as its neighbourhood is done, so the slow tail of one
//           launch overlaps the head of the next.  Older data (X(t-1), and the buffer being overwritten, last read
//           two steps ago) is covered transitively: the tiles waited for have themselves waited for theirs.
// The host only chains launches whose predecessor on the stream is the previous oc_k_march2 launch of the same handle
// (anything else that writes the state resets the chain to mode 0).
struct OcDep2 {
    unsigned* flags;          // one word per tile; nullptr: do not publish
    unsigned  epoch;          // this launch
    int       mode;
    int       pra, prb;       // previous launch: row range and segmentation
    OcSeg2    pseg;
    OcPeer2   peer;           // linked row bands: the neighbours' halos and flag words (all null otherwise)
};

// Flag word of the k-th (k = 0..3) segment of strip xs of the PREVIOUS launch that overlaps the rows [r0-2, r1+2) a
// tile reads; -1 if there is none.  (The host only uses mode 1 when at most four segments can overlap.)
OC_HD int oc_dep2_index(const OcDep2& d, int xs, int k, int r0, int r1)
{
    const int lo = r0 - 2 < d.pra ? d.pra : r0 - 2, hi = r1 + 2 > d.prb ? d.prb : r1 + 2;
    if (xs < 0 || xs >= d.pseg.nstrips || lo >= hi) return -1;
    const int h = (xs == 0 || xs == d.pseg.nstrips - 1) ? d.pseg.rs_e : d.pseg.rs;
    const int y = (lo - d.pra) / h + k, y_hi = (hi - 1 - d.pra) / h;
    return y <= y_hi ? oc_seg2_index(d.pseg, xs, y) : -1;
}
// The host's condition for mode 1.
//  (1) the tallest tile of this launch overlaps at most four segments of the previous one (twelve flags polled);
//  (2) tiles of at least 32 rows (performance only: short tiles lose more to the polling than they gain);
//  (3) same tile numbering, and the tile with a given index covers (nearly) the same rows in both launches, so that
//      it waits for its namesake.  This is what makes one flag word per tile index safe: a word is also written by
//      the launches after the one a poller waits for, and flags[i] >= e-1 must imply that tile i of launch e-1 is
//      done.  By (3) tile i of launch e only finishes after tile i of launch e-1 (it read its rows), and so on.
// Whole cloths have identical tilings; within a group of substeps of a row band the range shrinks by two rows per
// side and the segment height by at most one row, which (3) admits; anything else falls back to the grid-wide wait.
OC_HD bool oc_dep2_chainable(const OcSeg2& seg, int ra, int rb, const OcSeg2& pseg, int pra, int prb, bool ignore_height = false)
{
    const int h_max = seg.rs > seg.rs_e ? seg.rs : seg.rs_e, h_min = seg.rs < seg.rs_e ? seg.rs : seg.rs_e;
    const int p_min = pseg.rs < pseg.rs_e ? pseg.rs : pseg.rs_e;
    if (pseg.nstrips != seg.nstrips || h_max + 4 > 2 * p_min) return false;
    if (!ignore_height && h_min < 32) return false;
    if (seg.nseg_all != pseg.nseg_all || seg.n_extra != pseg.n_extra) return false;
    for (int cls = 0; cls < 2; ++cls) {                         // interior strips, edge strips
        const int h = cls ? seg.rs_e : seg.rs, ph = cls ? pseg.rs_e : pseg.rs;
        const int ymax = seg.nseg_all + (cls ? seg.n_extra / 2 : 0) - 1;
        const int lim = h < ph ? h : ph;
        const int d0 = ra - pra, d1 = (ra + ymax * h) - (pra + ymax * ph);
        if ((d0 < 0 ? -d0 : d0) + 2 >= lim || (d1 < 0 ? -d1 : d1) + 2 >= lim) return false;
    }
    (void)rb; (void)prb;
    return true;
}

template <class M, int WC, class Ctx>
OC_HD bool oc_march2_body(Ctx& ctx, const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                          float4* __restrict__ C, int ra, int rb, OcSeg2 seg, int x_halo, const OcDep2& dep)
{
    OcMarch2<M, WC, Ctx> m(ctx, c);
    m.A = A; m.B = B; m.C = C;
    m.sm = reinterpret_cast<OcSmem2<WC>*>(ctx.smem());
    const int i = ctx.tid();
    const int U = c.U, V = c.V;
    const int W_out = WC - 2 * x_halo;
    const int cx0 = ctx.bx() * W_out - x_halo;
    const int ga = cx0 + 2 * i, gb = ga + 1;
    int r0, r1;
    oc_seg2_rows(seg, ctx.bx(), ctx.by(), ra, rb, r0, r1);
    if (r0 >= r1) return ctx.wait_deps(dep, c, r0, r0);           // CTA-uniform: no rows left for this tile (it still orders itself after its predecessors)
    m.i = i; m.pa = 2 * i + 2; m.ga = ga; m.U = U; m.V = V;
    int lo = r0, hi = r1;
    int plo = lo - 2; if (plo < 0) plo = 0;
    int in_lo = plo, in_hi = hi + 2; if (in_hi > V) in_hi = V;
    const int first = lo - 2;
    const int n_it = r1 - first + OC_MARCH_LAG;
    const int row0 = first - OC_MARCH_LAG;
    m.lo = lo; m.hi = hi; m.plo = plo; m.in_lo = in_lo; m.in_hi = in_hi; m.first = first; m.row0 = row0;
    m.peer = &dep.peer;
    m.oka = ga >= 0 && ga < U; m.okb = gb >= 0 && gb < U;
    m.sta = m.oka && 2 * i >= x_halo && 2 * i < WC - x_halo;
    m.stb = m.okb && 2 * i + 1 >= x_halo && 2 * i + 1 < WC - x_halo;
    auto clampc = [&](int g) { return g < 0 ? 0 : (g >= U ? U - 1 : g); };
    m.rh1 = make_float2(OC_LDG(c.rh1 + clampc(ga)), OC_LDG(c.rh1 + clampc(gb)));
    m.rh2 = make_float2(OC_LDG(c.rh2 + clampc(ga)), OC_LDG(c.rh2 + clampc(gb)));
    m.dx2ab = make_float2(OC_LDG(c.dx2 + clampc(ga)), OC_LDG(c.dx2 + clampc(gb)));
    m.dx2ma = make_float2(OC_LDG(c.dx2 + clampc(ga - 1)), OC_LDG(c.dx2 + clampc(ga)));
    m.ydt = oc_rcp_bf(c.dt);
    m.goff = (long long)ctx.bz() * c.cloth_stride - (long long)c.row_lo * U + ga;
    {
        int r = row0 < 0 ? 0 : (row0 >= V ? V - 1 : row0);
        m.rv1_n = OC_LDG(c.rv1 + r); m.rv2_n = OC_LDG(c.rv2 + r); m.dz2_n = OC_LDG(c.dz2 + r);
    }
    const float2 z2 = make_float2(0.f, 0.f);
    m.me_x.x = m.me_x.y = m.me_x.z = p_pack(z2);
    m.me_v = m.w1_x = m.w1_v = m.k1_q = m.k2a_q = m.k2b_q = m.me_x;
    m.kDa = m.kAb = make_f3(0.f, 0.f, 0.f);

    // benign content for the pad columns / pad thread slots (never written by a particle)
    {
        OcSmem2<WC>& s = *m.sm;
        constexpr int T = WC / 2;
        for (int e = i; e < 6 * OC_RING * 4; e += T) {
            const int comp = e / (OC_RING * 4), slot = (e / 4) % OC_RING, pc = e % 4;
            const int col = pc < 2 ? pc : WC + pc;
            s.X[comp][slot][col] = comp < 3 ? 1.0e3f + 8.0f * (float)col : 0.0f;
            if (comp < 3) {
                s.Dd[comp][slot][col] = 0.0f;
                if (slot < 2) s.FH2[comp][slot][col] = 0.0f;
                if (pc == 0 || pc == 3) {
                    const int ti = pc == 0 ? 0 : T + 1;
                    s.FDb[comp][slot][ti] = 0.0f; s.FAa[comp][slot][ti] = 0.0f;
                    if (slot < 2) s.FH1[comp][slot][ti] = 0.0f;
                }
            }
        }
    }

    // steady range: interior rows, all activities on
    int st_lo = lo > plo + 1 ? lo : plo + 1; if (st_lo < 2) st_lo = 2;
    int st_hi = hi < V - 3 ? hi : V - 3;
    if (st_hi > in_hi - OC_MARCH_LAG) st_hi = in_hi - OC_MARCH_LAG;
    // linked row bands: the two rows pushed into a neighbour's halo are taken on the generic path
    if (dep.peer.c[0] && st_lo < dep.peer.ra + 2) st_lo = dep.peer.ra + 2;
    if (dep.peer.c[1] && st_hi > dep.peer.rb - 2) st_hi = dep.peer.rb - 2;
    if (st_lo < st_hi && !oc_rows_unpinned(c, ctx.bz(), st_lo, st_hi)) st_hi = st_lo;      // custom pins (oc_set_pins) in these rows: generic path only
    int it_lo = st_lo - row0, it_hi = st_hi - row0;
    if (it_lo < 0) it_lo = 0;
    if (it_hi > n_it) it_hi = n_it;
    if (it_hi <= it_lo) it_lo = it_hi = n_it;
    const bool interior = cx0 >= 2 && cx0 + WC + 2 <= U;          // CTA-uniform

#ifdef __CUDA_ARCH__
    if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 1);       // development: CTA timeline (OC_DEBUG=8)
#endif
    if (!ctx.wait_deps(dep, c, r0, r1)) return false;
    int it = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const int end = phase == 0 ? it_lo : n_it;
        for (; it < end; ++it) m.template iter<false, false>(it);
#ifdef __CUDA_ARCH__
        if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, phase == 0 ? 2 : 4);
#endif
        if (phase == 0) {
            if (interior) {
                // fast mode: two copies of the body, which removes most of the register moves that rotate the
                // loop-carried rows (measured +2.6 %; three copies are slower again).  Exact mode is not unrolled: its
                // 2 x 11 KB of hot code, with the cold fallback blocks interleaved, misses in the instruction cache
                // and loses 10 %.
                if (!M::kExact)
                    for (; it + 1 < it_hi; it += 2) { m.template iter<true, true>(it); m.template iter<true, true>(it + 1); }
                for (; it < it_hi; ++it) m.template iter<true, true>(it);
            } else {
                for (; it < it_hi; ++it) m.template iter<true, false>(it);
            }
#ifdef __CUDA_ARCH__
            if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 3);
#endif
        }
    }
    return true;
}

#ifdef __CUDACC__
// four CTAs of WC/2 threads per SM: 230-250 registers per thread (every trade of registers for a fifth CTA lost)
#define OC_M2_BOUNDS __launch_bounds__(WC / 2, 4)
// Poll a dependency flag until it reaches `want` (acquire; kSys: the word is written by another GPU).  The polls go to
// L2, with exponential back-off.  A wait that does not end (2 s for a flag of this GPU, 30 s for a neighbour GPU's:
// its process may lag) cannot happen in a correct chain of launches.  It is fatal for the step: the thread records why in
// the handle's error word (host-mapped; oc_sync / oc_download / oc_step turn it into OC_ERR_CUDA) and in a device-side
// poison counter that makes every other waiting tile give up at once, and returns false — the CTA then leaves without
// computing from stale rows and without publishing.  (No __trap(): a noreturn path in front of the steady loop changes
// ptxas's uniform-register allocation inside it — 33 more instructions per iteration, -5 %.)
template <bool kSys>
__device__ __forceinline__ bool oc_flag_wait(const OcConst& c, const unsigned* p, unsigned want)
{
    unsigned v, ns = 100, polls = 0;
    unsigned long long t0 = 0;
    for (;;) {
        if (kSys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        else      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if ((int)(v - want) >= 0) return true;
        __nanosleep(ns);
        if (ns < 1600) { ns += ns; continue; }
        if ((++polls & 63u) != 0u) continue;
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        const unsigned long long limit = (c.dbg & 32) ? 50000000ull : (kSys ? 30000000000ull : 2000000000ull);
        const bool poisoned = (*(volatile unsigned long long*)(c.dbg_cnt + 2) >> 40) != 0ull;
        if (poisoned || t - t0 > limit) {
            atomicAdd(c.dbg_cnt + 2, 1ull << 40);
            if (c.err) { *(volatile unsigned*)c.err = kSys ? 2u : 1u; __threadfence_system(); }
            return false;
        }
    }
}
struct OcDevCtx2 {          // grid = (tiles, 1, batch): the tile -> (strip, segment) map is OcSeg2's
    int x, y;
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int bx() const { return x; }
    __device__ __forceinline__ int by() const { return y; }
    __device__ __forceinline__ int bz() const { return blockIdx.z; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ unsigned char* smem() const { extern __shared__ __align__(16) unsigned char oc_dyn_smem[]; return oc_dyn_smem; }
    // Programmatic dependent launch: everything before this point (set-up, table loads, shared-memory pads) may
    // overlap the previous launch; the state buffers are only touched after it.  See OcDep2.
    // false: a dependency never arrived (oc_flag_wait); the caller leaves the kernel
    __device__ __forceinline__ bool wait_deps(const OcDep2& d, const OcConst& c, int r0, int r1) const
    {
        const int t = threadIdx.x;
        bool ok = true;
        if (d.mode == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        else if (t < 12) {
            const int idx = oc_dep2_index(d, x - 1 + t / 4, t % 4, r0, r1);
            if (idx >= 0)
                ok = oc_flag_wait<false>(c, d.flags + (size_t)blockIdx.z * oc_seg2_tiles(d.pseg) + idx, d.epoch - 1u + ((c.dbg & 32) ? 1000u : 0u));   // same cloth
        }
        // linked row bands: a tile that reads halo rows waits for the neighbour's boundary tiles of the previous step in
        // its own and the two adjacent strips (they wrote those rows, and they were the last readers of the neighbour's
        // halo rows this tile is about to overwrite)
        if (t >= 16 && t < 22) {
            const int side = (t - 16) / 3, xs = x - 1 + (t - 16) % 3;
            const bool reads_halo = side == 0 ? r0 < d.peer.ra + 2 : r1 > d.peer.rb - 2;
            if (d.peer.flags_in[side] && reads_halo && xs >= 0 && xs < d.peer.nstrips)
                ok = oc_flag_wait<true>(c, d.peer.flags_in[side] + xs, d.peer.epoch - 1u);
        }
        return __syncthreads_and(ok) != 0;
    }
    // after the tile's last store
    __device__ __forceinline__ void publish(const OcDep2& d, const OcSeg2& seg, int r0, int r1) const
    {
        if (!d.flags) return;                  // (a linked band always has its local flags)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(d.flags + (size_t)blockIdx.z * gridDim.x + oc_seg2_index(seg, x, y)), "r"(d.epoch) : "memory");
            const OcPeer2* pp = oc_opaque(&d.peer);
            const bool up = pp->flags_out[0] && r0 < pp->ra + 2 && r1 > r0, dn = pp->flags_out[1] && r1 > pp->rb - 2 && r1 > r0;
            if (up | dn) {
                __threadfence_system();
                if (up) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pp->flags_out[0] + x), "r"(pp->epoch) : "memory");
                if (dn) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pp->flags_out[1] + x), "r"(pp->epoch) : "memory");
            }
        }
    }
};
template <class M, int WC>
__global__ void OC_M2_BOUNDS
oc_k_march2(const __grid_constant__ OcConst c, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C,
            int ra, int rb, OcSeg2 seg, int x_halo, const __grid_constant__ OcDep2 dep)
{
    asm volatile("griddepcontrol.launch_dependents;");      // the next step may be placed as soon as CTA slots free up
    if ((c.dbg & 8) && threadIdx.x == 0) oc_timeline_mark(c, 0);
    OcDevCtx2 ctx;
    oc_seg2_tile(seg, blockIdx.x, ctx.x, ctx.y);
    if (!oc_march2_body<M, WC, OcDevCtx2>(ctx, c, A, B, C, ra, rb, seg, x_halo, dep)) return;      // a dependency timed out: nothing published
    int r0, r1;
    oc_seg2_rows(seg, ctx.x, ctx.y, ra, rb, r0, r1);
    ctx.publish(dep, seg, r0, r1);
}
#endif

// ---- host side (oc_march.cu) -------------------------------------------------------------------
int  oc_march2_configure(int device);
int  oc_march2_plan(const OcConst& c, bool exact, bool chained, bool linked, int ra, int rb, int sm_count, int occ_hint, OcMarchPlan* plan, OcSeg2* seg);
int  oc_march2_nstrips(int nx);
// host side of OcDep2: the chain of launches of one handle
struct OcChain2 {
    unsigned* flags; int cap;      // device flag words
    unsigned  epoch;
    bool      valid;               // the previous stream operation of the handle was the launch described below
    int       kind;                // ... of 0 = oc_k_march2, 1 = oc_k_twin, 2 = oc_k_stream (a chain never crosses kernels)
    int       pra, prb;
    OcSeg2    pseg;
};
#ifdef __CUDACC__
cudaError_t oc_march2_launch(const OcConst& c, bool exact, int ra, int rb, int sm_count,
                             const float4* A, const float4* B, float4* C, cudaStream_t stream, int* n_launches, OcChain2* chain,
                             const OcPeer2* peer = nullptr);
#endif
