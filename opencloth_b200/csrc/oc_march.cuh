// oc_march.cuh — kernel 2: the fused marching stencil kernel (the hot path).
//
// One launch advances the cloth by S substeps (temporal blocking).  It fuses, per substep, the
// reference's ComputeForces (V:448-484), IntegrateVerlet (V:428-444), EllipsoidCollision (V:509-533)
// and the pin mask (V:455, V:479-482) — "V:" = /root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp.
//
// Shape of the computation
//   * A CTA owns a column strip of TW columns and a segment of RS rows and MARCHES down the rows,
//     one row per iteration, one thread per column.  Neighbour data never comes from global memory:
//     each row is loaded once (coalesced float4), turned into (X, V, X-X_last) once, and published
//     into a 4-row shared-memory ring.
//   * Every spring is evaluated ONCE, by its upper/left end ("forward" springs: +1,0  +2,0  0,+1
//     0,+2  +1,+1  -1,+1), and the exactly negated force is handed to the partner: through shared
//     memory for the four cross-column springs, through registers for the two vertical ones (the
//     partner is the same thread one or two iterations later).  f(b,a) == -f(a,b) bit for bit, so
//     this halves the div/sqrt work without changing a single rounding.
//   * The gather phase adds the twelve (+ duplicated edge bend) spring forces in exactly the order
//     in which the reference's spring list touches the particle, then
//     integrates, clamps to the floor, collides and writes X(t+1) with the collider flag in w.
//   * Temporal blocking: S warp groups ("stages") of TW threads each form a systolic pipeline.
//     Stage s computes substep s+1 and publishes its rows into the ring of stage s+1, which runs
//     4 rows behind.  Only stage 0 reads global memory and only the last stage (and, for S >= 2,
//     the one before it, for X(t+S-1)) writes it: HBM traffic per substep drops to 64/S bytes.
//   * One __syncthreads per row iteration.  Hazard analysis (iteration it = P-phase, barrier,
//     G-phase):  rows read in P(it) were published in G(it-2) or earlier;  a ring slot (row & 3)
//     is rewritten 4 iterations after it was written, 1 iteration after its last read.
//
// The body is a template over an execution context so that tests/emu can run the identical code on
// the CPU with pthread barriers (kernel-logic check for the CPU-only test tier).
#pragma once
#include "oc_core.cuh"

#define OC_MARCH_MAX_STAGES 8
#define OC_MARCH_LAG 4          // rows between consecutive stages (must be a multiple of OC_RING)
#define OC_RING 4               // ring depth (power of two)

// Shared memory of one stage.  Structure-of-arrays rows (2 pad columns either side) in rings:
// a 32-bit load per component lets the two partners of a spring PAIR land in adjacent registers,
// ready for the packed FP32x2 instructions, and consecutive lanes hit consecutive banks.
template <int TW>
struct OcStageSmem {
    float  X[6][OC_RING][TW + 4];   // x, y, z, vx, vy, vz of the stage's input rows      slot = row & 3
    float4 D[OC_RING][TW + 4];      // X - X_last (w unused), read by the owning column only
    float  FH[6][2][TW + 4];        // f(+1,0).xyz, f(+2,0).xyz of row (row & 1): force ON the publishing
    float  FD[6][OC_RING][TW + 4];  // f(+1,+1).xyz, f(-1,+1).xyz of row (row & 3)   particle; the partner SUBTRACTS it
    float4 stage[2][TW];            // landing zone of stage 0's asynchronous row loads: A[col], B[col] per thread
};

#ifdef __CUDA_ARCH__
#define OC_LDG(p) __ldg(p)
// Optimisation fence on a float4 held in registers: whatever consumes it is scheduled after this
// point, and all four registers stay allocated until here.  Used to keep the consumers of a global
// load behind the barrier (its latency is then covered by the spring phase) and to stop ptxas from
// recycling a register of the in-flight 128-bit load (a write-after-write wait of a full DRAM latency).
#define OC_KEEP4(v) asm volatile("" : "+f"((v).x), "+f"((v).y), "+f"((v).z), "+f"((v).w))
#else
#define OC_LDG(p) (*(p))
#define OC_KEEP4(v) ((void)0)
#endif

// Asynchronous 16-byte copy global -> shared (LDGSTS): the row of the next-but-three iteration is requested
// at the top of an iteration without occupying registers and without giving ptxas the chance to sink the
// load next to its use; the requesting thread waits for its own copies just before it consumes them.
#ifdef __CUDA_ARCH__
OC_HD void oc_cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
OC_HD void oc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
OC_HD void oc_cp_async_wait()   { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N> OC_HD void oc_cp_async_wait_n() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }      // all but the newest N groups
#else
OC_HD void oc_cp_async16(void* smem_dst, const void* gsrc) { *reinterpret_cast<float4*>(smem_dst) = *reinterpret_cast<const float4*>(gsrc); }
OC_HD void oc_cp_async_commit() {}
OC_HD void oc_cp_async_wait() {}
template <int N> OC_HD void oc_cp_async_wait_n() {}
#endif

// position (x) and velocity (v) of two particles as pairs
struct OcPV2 { OcPair3 x, v; };
template <int TW>
OC_HD OcPV2 oc_ld_pv2(const OcStageSmem<TW>& sm, int slotA, int colA, int slotB, int colB)
{
    OcPV2 r;
    r.x.x = make_float2(sm.X[0][slotA][colA], sm.X[0][slotB][colB]);
    r.x.y = make_float2(sm.X[1][slotA][colA], sm.X[1][slotB][colB]);
    r.x.z = make_float2(sm.X[2][slotA][colA], sm.X[2][slotB][colB]);
    r.v.x = make_float2(sm.X[3][slotA][colA], sm.X[3][slotB][colB]);
    r.v.y = make_float2(sm.X[4][slotA][colA], sm.X[4][slotB][colB]);
    r.v.z = make_float2(sm.X[5][slotA][colA], sm.X[5][slotB][colB]);
    return r;
}
template <int TW>
OC_HD void oc_st_pvd(OcStageSmem<TW>& sm, int slot, int col, f3 x, f3 v, f3 d, float pad)
{
    sm.X[0][slot][col] = x.x; sm.X[1][slot][col] = x.y; sm.X[2][slot][col] = x.z;
    sm.X[3][slot][col] = v.x; sm.X[4][slot][col] = v.y; sm.X[5][slot][col] = v.z;
    sm.D[slot][col] = make_float4(d.x, d.y, d.z, pad);
}

// Force accumulator as (xy pair, z).  a - b is a + (-b) exactly, so the partner of a spring
// subtracts the force its publisher computed for itself (f(b,a) == -f(a,b), oc_core.cuh).
struct OcF { float2 xy; float z; };
template <class M> OC_HD void oc_add(OcF& F, float x, float y, float z, bool on)
{
    if (on) { F.xy.x = M::add(F.xy.x, x); F.xy.y = M::add(F.xy.y, y); F.z = M::add(F.z, z); }
}
template <class M> OC_HD void oc_sub(OcF& F, float x, float y, float z, bool on)
{
    if (on) { F.xy.x = M::sub(F.xy.x, x); F.xy.y = M::sub(F.xy.y, y); F.z = M::sub(F.z, z); }
}
// received force: (x, y) arrive in adjacent registers straight from shared memory -> one packed op
template <class M> OC_HD void oc_sub_pk(OcF& F, float2 xy, float z, bool on)
{
    if (on) { F.xy = p_sub(F.xy, xy); F.z = M::sub(F.z, z); }
}

// Cold path of exact mode: an operand of this lane left the range of the branch-free sequences
// (or a neighbour does not exist and the lane holds garbage).  Re-reads the inputs from shared memory
// and evaluates the six springs with the IEEE intrinsics.  Not inlined: keeps the hot loop small and
// its registers free.
template <class M, int TW>
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
void oc_march_redo(const OcStageSmem<TW>& in, int sl, int s1, int s2, int ci, const OcConst& c,
                   float rh1_i, float rh2_i, float dx2_i, float dx2_m, float rv1_j, float rv2_j, float dz2_j,
                   OcPair3& gH, OcPair3& gV, OcPair3& gS)
{
#define OC_LDX(slot, col) make_f3(in.X[0][slot][col], in.X[1][slot][col], in.X[2][slot][col])
#define OC_LDV(slot, col) make_f3(in.X[3][slot][col], in.X[4][slot][col], in.X[5][slot][col])
    const f3 mx = OC_LDX(sl, ci), mv = OC_LDV(sl, ci);
    const float sD = M::sqrt(M::add(dx2_i, dz2_j)), sA = M::sqrt(M::add(dx2_m, dz2_j));
    const f3 h1 = oc_spring<M>(mx, mv, OC_LDX(sl, ci + 1), OC_LDV(sl, ci + 1), rh1_i, c.nks_struct, c.kd_struct);
    const f3 h2 = oc_spring<M>(mx, mv, OC_LDX(sl, ci + 2), OC_LDV(sl, ci + 2), rh2_i, c.nks_bend,   c.kd_bend);
    const f3 v1 = oc_spring<M>(mx, mv, OC_LDX(s1, ci),     OC_LDV(s1, ci),     rv1_j, c.nks_struct, c.kd_struct);
    const f3 v2 = oc_spring<M>(mx, mv, OC_LDX(s2, ci),     OC_LDV(s2, ci),     rv2_j, c.nks_bend,   c.kd_bend);
    const f3 dd = oc_spring<M>(mx, mv, OC_LDX(s1, ci + 1), OC_LDV(s1, ci + 1), sD,    c.nks_shear,  c.kd_shear);
    const f3 da = oc_spring<M>(mx, mv, OC_LDX(s1, ci - 1), OC_LDV(s1, ci - 1), sA,    c.nks_shear,  c.kd_shear);
#undef OC_LDX
#undef OC_LDV
    gH.x = make_float2(h1.x, h2.x); gH.y = make_float2(h1.y, h2.y); gH.z = make_float2(h1.z, h2.z);
    gV.x = make_float2(v1.x, v2.x); gV.y = make_float2(v1.y, v2.y); gV.z = make_float2(v1.z, v2.z);
    gS.x = make_float2(dd.x, da.x); gS.y = make_float2(dd.y, da.y); gS.z = make_float2(dd.z, da.z);
}

// All per-thread state of the marching loop.  iter<kSteady, kSlot>() is one row iteration; the
// steady variant is used when this stage's row is interior (rows row-2 .. row+2 exist, no duplicated
// vertical bend spring, not the pinned row) and all three activities (load, spring phase, gather
// phase) are on: it carries no row predicates and no dead-path initialisations.
template <class M, int S, int TW, class Ctx>
struct OcMarch {
    typedef OcStageSmem<TW> Smem;
    Ctx& ctx;
    const OcConst& c;
    const float4* __restrict__ A; const float4* __restrict__ B;
    float4* __restrict__ C; float4* __restrict__ Dst;
    Smem* rings;
    int s, ci, gi, U, V, r0, r1;
    int lo_s, hi_s, plo_s, in_lo, in_hi, first, row0;
    bool col_ok, col_store;
    bool has_l1, has_l2, has_r1, has_r2, dup_r, dup_l;
    float rh1_i, rh2_i, dx2_i, dx2_m, ydt;
    float rv1_n, rv2_n, dz2_n;            // row constants of the NEXT iteration's row
    long long goff;                       // element offset of (cloth, gi, row 0) in a position buffer
    f3 k1, k2a, k2b;                      // vertical forces published by rows row-1 (0,+1), row-1 and row-2 (0,+2)

    OC_HD OcMarch(Ctx& ctx_, const OcConst& c_) : ctx(ctx_), c(c_) {}

    // kSlot = row & 3 when it is known at compile time (the steady loop is unrolled by 4 so that every
    // shared-memory address is the thread's base plus an immediate), -1 otherwise
    template <bool kSteady, int kSlot>
    OC_HD void iter(int it)
    {
        Smem& in = rings[s];
        const int row = row0 + it;                                // row this stage works on
        const int lrow = first + it;                              // row stage 0 loads
        // ---- stage 0: request row lrow asynchronously (LDGSTS into the thread's landing zone) ---------
        // (columns of the window that lie outside the cloth publish a benign far-away particle at rest, so
        // that the lanes next to them stay inside the operand range of the branch-free sequences)
        const bool doL = (s == 0) && (kSteady || (lrow >= in_lo && lrow < in_hi));
        if (doL) {
            Smem& st = rings[0];
            const int li = ci - 2;
            if (col_ok) {
                const long long o = goff + (long long)lrow * U;
                oc_cp_async16(&st.stage[0][li], A + o); oc_cp_async16(&st.stage[1][li], B + o);
            } else {
                // distinct per column and row: the springs between two such particles must not be degenerate
                st.stage[0][li] = st.stage[1][li] = make_float4(1.0e3f + 8.0f * (float)ci, 1.0e3f, 1.0e3f + 8.0f * (float)(lrow & 63), oc_u2f(OC_W_PLAIN));
            }
            oc_cp_async_commit();
        }
        const float rv1_j = rv1_n, rv2_j = rv2_n, dz2_j = dz2_n;
        {
            int r = row + 1;
            if (!kSteady) r = r < 0 ? 0 : (r >= V ? V - 1 : r);
            rv1_n = OC_LDG(c.rv1 + r); rv2_n = OC_LDG(c.rv2 + r); dz2_n = OC_LDG(c.dz2 + r);
        }
        const int sl = kSlot >= 0 ? kSlot : (row & (OC_RING - 1));           // slot of row
        const int s1 = (sl + 1) & (OC_RING - 1);                             // row + 1
        const int s2 = (sl + 2) & (OC_RING - 1);                             // row + 2
        const int s3 = (sl + 3) & (OC_RING - 1);                             // row - 1

        // ---- P phase: forward springs of row `row`, as three pairs ----------------------------------
        const bool doP = kSteady || (row >= plo_s && row < hi_s);
        OcPair3 gH, gV, gS;          // (f(+1,0), f(+2,0)), (f(0,+1), f(0,+2)), (f(+1,+1), f(-1,+1))
        f3 mx, mv, dme;
        if (doP) {
            mx = make_f3(in.X[0][sl][ci], in.X[1][sl][ci], in.X[2][sl][ci]);
            mv = make_f3(in.X[3][sl][ci], in.X[4][sl][ci], in.X[5][sl][ci]);
            const float4 d4 = in.D[sl][ci];
            dme = make_f3(d4.x, d4.y, d4.z);
            const OcPV2 nH = oc_ld_pv2<TW>(in, sl, ci + 1, sl, ci + 2);      // (i+1, j), (i+2, j)
            const OcPV2 nV = oc_ld_pv2<TW>(in, s1, ci,     s2, ci);          // (i, j+1), (i, j+2)
            const OcPV2 nS = oc_ld_pv2<TW>(in, s1, ci + 1, s1, ci - 1);      // (i+1, j+1), (i-1, j+1)
            bool bad = false;
            float2 rS = oc_sqrt2<M>(p_add(make_float2(dx2_i, dx2_m), p_bc(dz2_j)), bad);     // shear rest lengths
            float2 rH = make_float2(rh1_i, rh2_i), rV = make_float2(rv1_j, rv2_j);
            const float2 nksHV = make_float2(c.nks_struct, c.nks_bend), kdHV = make_float2(c.kd_struct, c.kd_bend);
            const float2 nksS = p_bc(c.nks_shear), kdS = p_bc(c.kd_shear);
            if (!M::kExact) { rH = p_mul(rH, nksHV); rV = p_mul(rV, nksHV); rS = p_mul(rS, nksS); }
            unsigned cls = 0;
            gH = oc_spring2<M>(mx, mv, nH.x, nH.v, rH, nksHV, kdHV, c.one, bad, &cls);
            gV = oc_spring2<M>(mx, mv, nV.x, nV.v, rV, nksHV, kdHV, c.one, bad, &cls);
            gS = oc_spring2<M>(mx, mv, nS.x, nS.v, rS, nksS,  kdS,  c.one, bad, &cls);
#if defined(OC_CLASSIFY) && defined(__CUDA_ARCH__)
            if (M::kExact && (c.dbg & 4) && col_store && row >= lo_s && row < hi_s) {
                if (cls & 1u) atomicAdd(c.dbg_cnt + 3, 1ull);                       // -0 numerator
                if (cls & 2u) atomicAdd(c.dbg_cnt + 3, 1ull << 20);                // tiny numerator
                if (cls & 4u) atomicAdd(c.dbg_cnt + 3, 1ull << 40);                // huge numerator
                if (cls & 8u) atomicAdd(c.dbg_cnt + 2, 1ull << 32);                // squared length
            }
#endif
            if (c.dbg & 1) bad = true;
            if (c.dbg & 2) bad = false;
#ifdef __CUDA_ARCH__
            if (M::kExact && bad && (c.dbg & 4)) {
                atomicAdd(c.dbg_cnt, 1ull);
                if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 1, 1ull);
            }
#endif
            if (M::kExact && bad) oc_march_redo<M, TW>(in, sl, s1, s2, ci, c, rh1_i, rh2_i, dx2_i, dx2_m, rv1_j, rv2_j, dz2_j, gH, gV, gS);
            in.FH[0][sl & 1][ci] = gH.x.x; in.FH[1][sl & 1][ci] = gH.y.x; in.FH[2][sl & 1][ci] = gH.z.x;
            in.FH[3][sl & 1][ci] = gH.x.y; in.FH[4][sl & 1][ci] = gH.y.y; in.FH[5][sl & 1][ci] = gH.z.y;
            in.FD[0][sl][ci] = gS.x.x; in.FD[1][sl][ci] = gS.y.x; in.FD[2][sl][ci] = gS.z.x;
            in.FD[3][sl][ci] = gS.x.y; in.FD[4][sl][ci] = gS.y.y; in.FD[5][sl][ci] = gS.z.y;
        }

        ctx.sync();

        // ---- G phase: gather in the reference's order, integrate, collide, hand on ------------------
        const bool doG = kSteady || (row >= lo_s && row < hi_s);
        if (doG) {
            const bool pinned = !kSteady && oc_pinned(c, ctx.bz(), gi, row);
            // F = 0 + gravity*mass (unless pinned) + DEFAULT_DAMPING*V     V:451-459
            OcF F;
            F.xy = pinned ? p_bc(0.0f) : make_float2(c.f0[0], c.f0[1]);
            F.z  = pinned ? 0.0f : c.f0[2];
            F.xy = p_sump<M>(p_mul(p_bc(c.damping), make_float2(mv.x, mv.y)), F.xy, c.one);
            F.z  = M::add(F.z, M::mul(c.damping, mv.z));
            if (!pinned) {
                const bool up1 = kSteady || row - 1 >= 0, up2 = kSteady || row - 2 >= 0;
                const bool dn1 = kSteady || row + 1 < V,  dn2 = kSteady || row + 2 < V;
                const int h = sl & 1;
                oc_sub_pk<M>(F, make_float2(in.FH[0][h][ci - 1], in.FH[1][h][ci - 1]), in.FH[2][h][ci - 1], has_l1);   // 1  (i-1, j)   structural
                oc_add<M>(F, gH.x.x, gH.y.x, gH.z.x, has_r1);                                                        // 2  (i+1, j)
                oc_sub<M>(F, k1.x, k1.y, k1.z, up1);                                                                 // 3  (i, j-1)
                oc_add<M>(F, gV.x.x, gV.y.x, gV.z.x, dn1);                                                           // 4  (i, j+1)
                oc_sub_pk<M>(F, make_float2(in.FD[0][s3][ci - 1], in.FD[1][s3][ci - 1]), in.FD[2][s3][ci - 1], has_l1 && up1);   // 5  (i-1, j-1) shear
                oc_sub_pk<M>(F, make_float2(in.FD[3][s3][ci + 1], in.FD[4][s3][ci + 1]), in.FD[5][s3][ci + 1], has_r1 && up1);   // 6  (i+1, j-1)
                oc_add<M>(F, gS.x.y, gS.y.y, gS.z.y, has_l1 && dn1);                                                 // 7  (i-1, j+1)
                oc_add<M>(F, gS.x.x, gS.y.x, gS.z.x, has_r1 && dn1);                                                 // 8  (i+1, j+1)
                const float2 h2xy = make_float2(in.FH[3][h][ci - 2], in.FH[4][h][ci - 2]);
                const float h2z = in.FH[5][h][ci - 2];
                oc_sub_pk<M>(F, h2xy, h2z, has_l2);                                                                  // 9  (i-2, j)   bend
                oc_add<M>(F, gH.x.y, gH.y.y, gH.z.y, has_r2);                                                        // 10 (i+2, j)
                oc_add<M>(F, gH.x.y, gH.y.y, gH.z.y, dup_r);                                                         // 11 duplicate of the row's last bend spring (V:313)
                oc_sub_pk<M>(F, h2xy, h2z, dup_l);
                oc_sub<M>(F, k2b.x, k2b.y, k2b.z, up2);                                                              // 12 (i, j-2)
                oc_add<M>(F, gV.x.y, gV.y.y, gV.z.y, dn2);                                                           // 13 (i, j+2)
                if (!kSteady) {
                    oc_add<M>(F, gV.x.y, gV.y.y, gV.z.y, row == V - 3);                                              // 14 duplicate of the column's last bend spring (V:319)
                    oc_sub<M>(F, k2b.x, k2b.y, k2b.z, row == V - 1);
                }
            }
            bool hit;
            float2 nxy; float nz;
            oc_integrate_collide2<M>(c, make_float2(mx.x, mx.y), mx.z, make_float2(dme.x, dme.y), dme.z, F.xy, F.z, nxy, nz, &hit);
            const float4 out = make_float4(nxy.x, nxy.y, nz, oc_u2f(hit ? OC_W_HIT : OC_W_PLAIN));
            if (s == S - 1) {
                if (col_store) C[goff + (long long)row * U] = out;                // X(t+S)
            } else {
                // new X_last is the old X (V:438) unless the collider moved the particle (V:530)
                float2 dnxy = p_sub(nxy, make_float2(mx.x, mx.y));
                float dnz = M::sub(nz, mx.z);
                if (hit) { dnxy = p_bc(0.0f); dnz = 0.0f; }
                bool badv = false;
                float2 vxy; float vz;
                oc_velocity2<M>(dnxy, dnz, c, ydt, badv, vxy, vz);
                if (M::kExact && badv) { const f3 v = M::velocity(make_f3(dnxy.x, dnxy.y, dnz), c); vxy = make_float2(v.x, v.y); vz = v.z; }
                oc_st_pvd<TW>(rings[s + 1], sl, ci, make_f3(nxy.x, nxy.y, nz), make_f3(vxy.x, vxy.y, vz), make_f3(dnxy.x, dnxy.y, dnz), 0.0f);
                if (s == S - 2 && col_store && row >= r0 && row < r1) Dst[goff + (long long)row * U] = out;   // X(t+S-1)
            }
        }
        if (doP) {
            k2b = k2a;
            k2a = make_f3(gV.x.y, gV.y.y, gV.z.y);
            k1  = make_f3(gV.x.x, gV.y.x, gV.z.x);
        }

        // ---- stage 0: publish the loaded row into its own ring --------------------------------------
        if (doL) {
            oc_cp_async_wait();
            const float4 la = rings[0].stage[0][ci - 2], lq = rings[0].stage[1][ci - 2];
            float2 dxy = p_sub(make_float2(la.x, la.y), make_float2(lq.x, lq.y));
            float dz = M::sub(la.z, lq.z);
            if (oc_hit(la.w)) { dxy = p_bc(0.0f); dz = 0.0f; }                   // X_last == X (V:530)
            bool badv = false;
            float2 vxy; float vz;
            oc_velocity2<M>(dxy, dz, c, ydt, badv, vxy, vz);
#ifdef __CUDA_ARCH__
            if (M::kExact && badv && (c.dbg & 4)) atomicAdd(c.dbg_cnt + 2, 1ull);
#endif
            if (M::kExact && badv) { const f3 v = M::velocity(make_f3(dxy.x, dxy.y, dz), c); vxy = make_float2(v.x, v.y); vz = v.z; }
            oc_st_pvd<TW>(rings[0], sl, ci, make_f3(la.x, la.y, la.z), make_f3(vxy.x, vxy.y, vz), make_f3(dxy.x, dxy.y, dz), lq.w);   // lrow = row + 4: same slot
        }
    }
};

// Ctx: tid(), bx(), by(), bz(), sync(), smem()  (see OcDevCtx below and tests/emu/oc_emu.cu)
template <class M, int S, int TW, class Ctx>
OC_HD void oc_march_body(Ctx& ctx, const OcConst& c,
                         const float4* __restrict__ A, const float4* __restrict__ B,
                         float4* __restrict__ C, float4* __restrict__ Dst,
                         int ra, int rb, int RS, int x_halo)
{
    OcMarch<M, S, TW, Ctx> m(ctx, c);
    m.A = A; m.B = B; m.C = C; m.Dst = Dst;
    m.rings = reinterpret_cast<OcStageSmem<TW>*>(ctx.smem());
    const int tid = ctx.tid();
    const int s = tid / TW;              // stage (substep s+1 of this launch)
    const int i = tid - s * TW;          // column lane
    const int U = c.U, V = c.V;
    // x_halo = 2*S columns either side are recomputed by the neighbouring strips; 0 when one strip
    // spans the whole cloth width (both strip edges are cloth edges: nothing to recompute)
    const int W_out = TW - 2 * x_halo;
    const int gi = ctx.bx() * W_out - x_halo + i;      // global column
    const int r0 = ra + ctx.by() * RS;
    const int r1 = (r0 + RS < rb) ? r0 + RS : rb;
    m.s = s; m.ci = i + 2; m.gi = gi; m.U = U; m.V = V; m.r0 = r0; m.r1 = r1;

    // rows this stage produces, rows it must run the spring phase on, rows it needs as input
    int lo_s = r0 - 2 * (S - 1 - s); if (lo_s < 0) lo_s = 0;
    int hi_s = r1 + 2 * (S - 1 - s); if (hi_s > V) hi_s = V;
    int plo_s = lo_s - 2; if (plo_s < 0) plo_s = 0;
    int lo_0 = r0 - 2 * (S - 1); if (lo_0 < 0) lo_0 = 0;
    int hi_0 = r1 + 2 * (S - 1); if (hi_0 > V) hi_0 = V;
    int in_lo = lo_0 - 2; if (in_lo < 0) in_lo = 0;
    int in_hi = hi_0 + 2; if (in_hi > V) in_hi = V;
    const int first = lo_0 - 2;
    const int n_it = r1 - first + OC_MARCH_LAG * S;
    const int row0 = first - OC_MARCH_LAG * (s + 1);
    m.lo_s = lo_s; m.hi_s = hi_s; m.plo_s = plo_s; m.in_lo = in_lo; m.in_hi = in_hi; m.first = first; m.row0 = row0;

    m.col_ok = gi >= 0 && gi < U;
    m.col_store = m.col_ok && i >= x_halo && i < TW - x_halo;      // columns valid after S substeps
    const int gic = gi < 0 ? 0 : (gi >= U ? U - 1 : gi);
    const int gim = gic > 0 ? gic - 1 : 0;
    m.rh1_i = OC_LDG(c.rh1 + gic); m.rh2_i = OC_LDG(c.rh2 + gic);
    m.dx2_i = OC_LDG(c.dx2 + gic); m.dx2_m = OC_LDG(c.dx2 + gim);
    m.has_l1 = gi - 1 >= 0; m.has_l2 = gi - 2 >= 0; m.has_r1 = gi + 1 < U; m.has_r2 = gi + 2 < U;
    m.dup_r = gi == U - 3; m.dup_l = gi == U - 1;
    m.ydt = oc_rcp_bf(c.dt);             // reciprocal of dt for the branch-free velocity division
    m.goff = (long long)ctx.bz() * c.cloth_stride - (long long)c.row_lo * U + gic;
    {
        int r = row0 < 0 ? 0 : (row0 >= V ? V - 1 : row0);
        m.rv1_n = OC_LDG(c.rv1 + r); m.rv2_n = OC_LDG(c.rv2 + r); m.dz2_n = OC_LDG(c.dz2 + r);
    }
    m.k1 = m.k2a = m.k2b = make_f3(0.f, 0.f, 0.f);

    // Pad columns (2 either side of the window) are never written by a particle: give them the same
    // benign content once.  The first spring phase comes at least OC_MARCH_LAG barriers later.
    {
        OcStageSmem<TW>& sm = m.rings[s];
        for (int e = i; e < 6 * OC_RING * 4; e += TW) {
            const int comp = e / (OC_RING * 4), slot = (e / 4) % OC_RING, pc = e % 4;
            const int col = pc < 2 ? pc : TW + pc;
            sm.X[comp][slot][col] = comp < 3 ? 1.0e3f : 0.0f;
            sm.FD[comp][slot][col] = 0.0f;
            if (slot < 2) sm.FH[comp][slot][col] = 0.0f;
        }
    }

    // Steady iterations of this stage: its row is interior and inside the rows it produces, the
    // springs of the two rows above have been evaluated, and (stage 0) the row to load exists.
    //   row >= max(lo_s, plo_s + 1, 2)   row < min(hi_s, V - 3)   lrow = row + 4(s+1) < in_hi
    int st_lo = lo_s > plo_s + 1 ? lo_s : plo_s + 1; if (st_lo < 2) st_lo = 2;
    int st_hi = hi_s < V - 3 ? hi_s : V - 3;
    if (s == 0 && st_hi > in_hi - OC_MARCH_LAG) st_hi = in_hi - OC_MARCH_LAG;
    if (st_lo < st_hi && !oc_rows_unpinned(c, ctx.bz(), st_lo, st_hi)) st_hi = st_lo;      // custom pins in these rows: generic path only
    int it_lo = st_lo - row0, it_hi = st_hi - row0;            // steady for it in [it_lo, it_hi)
    if (it_lo < 0) it_lo = 0;
    if (it_hi > n_it) it_hi = n_it;
    if (it_hi <= it_lo) it_lo = it_hi = n_it;                  // no steady range

    // the steady loop is unrolled by 4 and starts on a row with row & 3 == 0
    while (it_lo < it_hi && ((row0 + it_lo) & (OC_RING - 1)) != 0) ++it_lo;
    it_hi = it_lo + ((it_hi - it_lo) & ~(OC_RING - 1));
    if (it_hi <= it_lo) it_lo = it_hi = n_it;

    // generic iterations up to the steady range, the steady loop, generic iterations to the end
    int it = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const int end = phase == 0 ? it_lo : n_it;
        for (; it < end; ++it) m.template iter<false, -1>(it);
        if (phase == 0)
            for (; it < it_hi; it += 4) {
                m.template iter<true, 0>(it);
                m.template iter<true, 1>(it + 1);
                m.template iter<true, 2>(it + 2);
                m.template iter<true, 3>(it + 3);
            }
    }
}

#ifdef __CUDACC__
struct OcDevCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int bx() const { return blockIdx.x; }
    __device__ __forceinline__ int by() const { return blockIdx.y; }
    __device__ __forceinline__ int bz() const { return blockIdx.z; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ unsigned char* smem() const { extern __shared__ __align__(16) unsigned char oc_dyn_smem[]; return oc_dyn_smem; }
};

// resident CTAs per SM the register allocation is capped for: 4 x 128 threads, 2 x 256, 1 x 512
#define OC_MARCH_MIN_CTAS(threads) ((threads) <= 32 ? 16 : ((threads) <= 64 ? 8 : ((threads) <= 128 ? OC_CTAS128 : ((threads) <= 256 ? 2 : 1))))
#ifndef OC_CTAS128
#define OC_CTAS128 4
#endif

template <class M, int S, int TW>
__global__ void __launch_bounds__(S * TW, OC_MARCH_MIN_CTAS(S * TW))
oc_k_march(OcConst c, const float4* __restrict__ A, const float4* __restrict__ B,
           float4* __restrict__ C, float4* __restrict__ Dst, int ra, int rb, int RS, int x_halo)
{
    OcDevCtx ctx;
    oc_march_body<M, S, TW, OcDevCtx>(ctx, c, A, B, C, Dst, ra, rb, RS, x_halo);
}

#endif

// ---- host side (oc_march.cu) -------------------------------------------------------------------
int  oc_march_configure(int device);                  // opt in to large dynamic shared memory; 0 or cudaError_t
cudaError_t oc_march_launch(const OcConst& c, bool exact, int S, int ra, int rb, int sm_count,
                            const float4* A, const float4* B, float4* C, float4* Dst,
                            cudaStream_t stream, int* n_launches);
// geometry chosen for a launch (also used by the emulator and by bench.py's report)
struct OcMarchPlan { int TW, S, x_halo, W_out, nstrips, RS, nseg, threads; size_t smem; };
int  oc_march_plan(const OcConst& c, int S, int ra, int rb, int sm_count, int occ_hint, OcMarchPlan* plan);
