// oc_twin.cuh — kernel 5: the marching stencil kernel over TWIN TILES (one substep per launch).
//
// Same algorithm, data flow, arithmetic and accumulation order as oc_march.cuh / oc_march2.cuh (read those first:
// forward springs evaluated once in a P phase, one barrier, gather + IntegrateVerlet + EllipsoidCollision in the
// reference's order in a G phase, rows loaded with cp.async four iterations ahead into a shared-memory ring).  What
// changes is what the two halves of a packed FP32x2 operation are:
//   * oc_k_march  packs two SPRINGS of one particle,
//   * oc_k_march2 packs two ADJACENT COLUMNS of one row: every spring that crosses the column pair needs a "shifted"
//     operand pair that has to be assembled with register moves, a third of the accumulation is scalar, and the
//     profile shows it: 13 % MOV, 8 % scalar FADD, 38 % packed FP in the fast kernel (profiles/r1d_*),
//   * this kernel packs the SAME column and the SAME iteration of TWO INDEPENDENT TILES (the "twins": two row
//     segments of one column strip, or the same tile of two cloths of a batch).  The two halves never interact, so
//     every operand of every spring, every accumulation and the whole integration / collider step is a naturally
//     aligned 64-bit pair: shared memory holds float2 (tile 0, tile 1) per column, partners are plain LDS.64 at
//     column +-1 / +-2, received forces are LDS.64, and there is not a single half-swap or scalar lane operation in
//     the steady loop.  One thread = one column, WC threads per CTA.
// Per thread-iteration (two particle updates) the steady loop issues ~40 % fewer instructions than oc_k_march2.
//
// The partner (column+1, row+1) of the (+1,+1) shear spring is the (+1,0) partner of the next iteration and is
// carried in registers.  Edge columns are handled as in oc_k_march2 (0/1 multipliers on springs that do not exist,
// steady loop without predicates); edge ROWS, the pipeline fill of a tile, pinned rows and the two rows a linked
// band pushes to its neighbour take a generic path with per-HALF row predicates.
//
// Tiles, dependencies between consecutive launches and linked row bands are OcSeg2 / OcDep2 / OcPeer2 of
// oc_march2.cuh, with one flag word per TILE as there; a CTA waits for, and publishes, the flags of both its tiles.
#pragma once
#include "oc_core.cuh"
#include "oc_march.cuh"
#include "oc_march2.cuh"

// kExact: the ring holds X and V = (X - X_last)/dt (the reference's separately rounded velocity) and, beside it, X - X_last.
// Fast mode never forms V: deltaV = (d1 - d2)/dt and DEFAULT_DAMPING*V are evaluated from d = X - X_last with 1/dt folded
// into Kd and the damping constant, so its ring holds X and d, and there is no Dd array.
template <int WC, bool kExact>
struct OcSmemT {
    float2 X[6][OC_RING][WC + 4];       // x, y, z, then vx, vy, vz (exact) or dx, dy, dz (fast) of (tile 0, tile 1)   slot = iteration & 3, index = window column + 2
    float2 Dd[kExact ? 3 : 1][kExact ? OC_RING : 1][kExact ? WC + 4 : 1];      // X - X_last (exact mode only)
    float2 FH1[3][2][WC + 4];           // f(+1,0) of every column, slot = iteration & 1
    float2 FH2[3][2][WC + 4];           // f(+2,0)
    float2 FD[3][OC_RING][WC + 4];      // f(+1,+1)
    float2 FA[3][OC_RING][WC + 4];      // f(-1,+1)
    float4 stage[4][WC];                // landing zone of the asynchronous row loads: A, B of tile 0, A, B of tile 1
};

// how a CTA finds its two tiles
struct OcTwinMap {
    int pair_cloths;      // 0: tiles (strip, 2k) and (strip, 2k+1) of one cloth;  1: tile (strip, k) of cloths 2z and 2z+1
};

// One spring pair of the thread (the same spring of both tiles) redone with the IEEE intrinsics; cold, out of line.
// kind: 0 (+1,0)  1 (+2,0)  2 (0,+1)  3 (0,+2)  4 (+1,+1)  5 (-1,+1); for 4 and 5 rest_* are the SQUARED rest lengths.
template <class M, class S>
OC_COLD OcPair3 oc_twin_redo(const OcConst* c, const S* s, int kind, int sl, int ci, float rest_0, float rest_1)
{
    const int s1 = (sl + 1) & (OC_RING - 1), s2 = (sl + 2) & (OC_RING - 1);
    int sp = sl, cp = ci;
    float nks = c->nks_struct, kd = c->kd_struct;
    switch (kind) {
    case 0: cp = ci + 1; break;
    case 1: cp = ci + 2; nks = c->nks_bend; kd = c->kd_bend; break;
    case 2: sp = s1; break;
    case 3: sp = s2; nks = c->nks_bend; kd = c->kd_bend; break;
    case 4: sp = s1; cp = ci + 1; nks = c->nks_shear; kd = c->kd_shear; break;
    default: sp = s1; cp = ci - 1; nks = c->nks_shear; kd = c->kd_shear; break;
    }
    if (kind >= 4) { rest_0 = M::sqrt(rest_0); rest_1 = M::sqrt(rest_1); }
#define OC_T0(comp, slot, col) s->X[comp][slot][col].x
#define OC_T1(comp, slot, col) s->X[comp][slot][col].y
    const f3 f0 = oc_spring<M>(make_f3(OC_T0(0, sl, ci), OC_T0(1, sl, ci), OC_T0(2, sl, ci)), make_f3(OC_T0(3, sl, ci), OC_T0(4, sl, ci), OC_T0(5, sl, ci)),
                               make_f3(OC_T0(0, sp, cp), OC_T0(1, sp, cp), OC_T0(2, sp, cp)), make_f3(OC_T0(3, sp, cp), OC_T0(4, sp, cp), OC_T0(5, sp, cp)), rest_0, nks, kd);
    const f3 f1 = oc_spring<M>(make_f3(OC_T1(0, sl, ci), OC_T1(1, sl, ci), OC_T1(2, sl, ci)), make_f3(OC_T1(3, sl, ci), OC_T1(4, sl, ci), OC_T1(5, sl, ci)),
                               make_f3(OC_T1(0, sp, cp), OC_T1(1, sp, cp), OC_T1(2, sp, cp)), make_f3(OC_T1(3, sp, cp), OC_T1(4, sp, cp), OC_T1(5, sp, cp)), rest_1, nks, kd);
#undef OC_T0
#undef OC_T1
    OcPair3 f;
    f.x = make_float2(f0.x, f1.x); f.y = make_float2(f0.y, f1.y); f.z = make_float2(f0.z, f1.z);
    return f;
}

// The same spring of both tiles: p1 = (px, pv), p2 = (qx, qv), all pairs (tile 0, tile 1).  Exact mode is oc_spring2v (the
// reference's operations, branch-free IEEE sequences).  Fast mode: pv / qv are X - X_last (kd = Kd / dt) and
//   s = -Ks (dist - rest)/dist + Kd (dv.dp)/dist^2 = nks + rinv (kd dot rinv - nks rest),   rinv = rsqrt(dp.dp),
// 17 packed operations and two MUFU per spring pair.  rest is pre-multiplied by nks.
template <class M>
OC_HD OcPair3 oc_spring_twin(const OcPair3& px, const OcPair3& pv, const OcPair3& qx, const OcPair3& qv,
                             float2 rest, float2 nks, float2 kd, float one, OcRange& rg)
{
    if (M::kExact) return oc_spring2v<M>(px, pv, qx, qv, rest, nks, kd, one, rg);
    OcPair3 dp, dv, f;
    dp.x = p_sub(px.x, qx.x); dp.y = p_sub(px.y, qx.y); dp.z = p_sub(px.z, qx.z);
    dv.x = p_sub(pv.x, qv.x); dv.y = p_sub(pv.y, qv.y); dv.z = p_sub(pv.z, qv.z);
    const float2 sqr  = p_fma(dp.z, dp.z, p_fma(dp.y, dp.y, p_mul(dp.x, dp.x)));
    const float2 rinv = p_rsq(sqr);
    const float2 dot  = p_fma(dv.z, dp.z, p_fma(dv.y, dp.y, p_mul(dv.x, dp.x)));
    const float2 u    = p_fma(p_mul(kd, dot), rinv, p_neg(rest));
    const float2 s    = p_fma(u, rinv, nks);
    f.x = p_mul(s, dp.x); f.y = p_mul(s, dp.y); f.z = p_mul(s, dp.z);
    return f;
}

// 1: the own column (rows row, row+1) and the column+1 partner are carried in registers from iteration to iteration;
// 0: re-read from the ring every iteration (18 LDS.64 more, 36 registers and their rotation less)
#ifndef OC_TWIN_CARRY
#define OC_TWIN_CARRY 1
#endif
// 1: fast mode adds a particle's own spring forces, the carried vertical ones, gravity and damping as they are produced in
// the P phase (any order is as good as another there); only the sum crosses the barrier.  Exact mode keeps the reference's
// order, so its six own forces stay live until their turn in the G phase.
#ifndef OC_TWIN_LEAN
#define OC_TWIN_LEAN 1
#endif

template <class M, int WC, class Ctx>
struct OcTwin {
    typedef OcSmemT<WC, M::kExact> Smem;
    Ctx& ctx;
    const OcConst& c;
    const float4* __restrict__ A; const float4* __restrict__ B;
    float4* __restrict__ C;
    Smem* sm;
    int i, ci, gi, U, V;
    int bz0, bz1;
    int lo0, hi0, plo0, in_lo0, in_hi0;      // rows of tile 0: produced [lo, hi), spring phase from plo, loaded [in_lo, in_hi)
    int lo1, hi1, plo1, in_lo1, in_hi1;      // the same for tile 1
    int row0, dRow;                          // row of tile 0 at iteration 0; tile 1 is dRow rows further down
    bool ok, st;                             // column exists / column is stored by this CTA
    float rh1, rh2, dx2i, dx2m, ydt;
    float mR1, mR2, mL1;                     // 1 if the spring to column +1 / +2 / -1 exists, else 0
    float kdt_struct, kdt_shear, kdt_bend, damp_dt;      // fast mode: Kd / dt, DEFAULT_DAMPING / dt
    float2 rv1_n, rv2_n, dz2_n;              // row constants of the NEXT iteration's rows (tile 0, tile 1)
    long long goff0, dOff;                   // element offset of (cloth of tile 0, gi, row 0); tile 1 minus tile 0 at the same iteration
    // carried from iteration to iteration as 64-bit pairs (see oc_q2):
    OcPair3q me_x, me_v, w1_x, w1_v;         // own column, rows row and row+1
    OcPair3q n1_x, n1_v;                     // column +1, row row  (the (+1,+1) partner of the previous iteration)
    OcPair3q k1_q, k2a_q, k2b_q;             // carried (0,+1) of row-1, (0,+2) of row-1 and row-2
    const OcPeer2* peer;

    OC_HD OcTwin(Ctx& ctx_, const OcConst& c_) : ctx(ctx_), c(c_) {}

    OC_HD OcPV2 ld(int slot, int col) const
    {
        OcPV2 r;
        r.x.x = sm->X[0][slot][col]; r.x.y = sm->X[1][slot][col]; r.x.z = sm->X[2][slot][col];
        r.v.x = sm->X[3][slot][col]; r.v.y = sm->X[4][slot][col]; r.v.z = sm->X[5][slot][col];
        return r;
    }
    OC_HD float2 row_pair(const float* t, int r0_, int r1_) const { return make_float2(OC_LDG(t + r0_), OC_LDG(t + r1_)); }
    static OC_HD float4 benign(int ci_, int lrow) { return make_float4(1.0e3f + 8.0f * (float)ci_, 1.0e3f, 1.0e3f + 8.0f * (float)(lrow & 63), oc_u2f(OC_W_PLAIN)); }

    // One loaded row (both tiles) -> position, velocity (X - X_last)/dt and X - X_last, once, into ring slot sl
    OC_HD void publish(int sl, const float4 la0, const float4 lq0, const float4 la1, const float4 lq1)
    {
        Smem& s = *sm;
        OcPair3 d;
        d.x = make_float2(M::sub(la0.x, lq0.x), M::sub(la1.x, lq1.x));
        d.y = make_float2(M::sub(la0.y, lq0.y), M::sub(la1.y, lq1.y));
        d.z = make_float2(M::sub(la0.z, lq0.z), M::sub(la1.z, lq1.z));
        if (oc_hit(la0.w)) { d.x.x = 0.0f; d.y.x = 0.0f; d.z.x = 0.0f; }       // X_last == X (V:530)
        if (oc_hit(la1.w)) { d.x.y = 0.0f; d.y.y = 0.0f; d.z.y = 0.0f; }
        s.X[0][sl][ci] = make_float2(la0.x, la1.x);
        s.X[1][sl][ci] = make_float2(la0.y, la1.y);
        s.X[2][sl][ci] = make_float2(la0.z, la1.z);
        if (!M::kExact) { s.X[3][sl][ci] = d.x; s.X[4][sl][ci] = d.y; s.X[5][sl][ci] = d.z; return; }
        OcPair3 v;
#ifdef __CUDA_ARCH__
        {
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
        }
#else
        v.x = make_float2(d.x.x / c.dt, d.x.y / c.dt); v.y = make_float2(d.y.x / c.dt, d.y.y / c.dt); v.z = make_float2(d.z.x / c.dt, d.z.y / c.dt);
#endif
        s.X[3][sl][ci] = v.x;
        s.X[4][sl][ci] = v.y;
        s.X[5][sl][ci] = v.z;
        constexpr int kD = M::kExact ? 1 : 0;      // (indices that exist in both layouts; this point is reached in exact mode only)
        s.Dd[0 * kD][sl * kD][ci * kD] = d.x;
        s.Dd[1 * kD][sl * kD][ci * kD] = d.y;
        s.Dd[2 * kD][sl * kD][ci * kD] = d.z;
    }

    // the carried registers of the steady loop, from the ring (at the entry of the steady loop)
    OC_HD void load_carried(int it)
    {
        const int sl = it & (OC_RING - 1), s1 = (sl + 1) & (OC_RING - 1);
        const OcPV2 me = ld(sl, ci), w1 = ld(s1, ci), n1 = ld(sl, ci + 1);
        me_x = p_pack3(me.x); me_v = p_pack3(me.v); w1_x = p_pack3(w1.x); w1_v = p_pack3(w1.v); n1_x = p_pack3(n1.x); n1_v = p_pack3(n1.v);
    }

    // kSlot = iteration & 3 when it is known at compile time (unrolled steady loop), -1 otherwise
    template <bool kSteady, bool kInterior, int kSlot>
    OC_HD void iter(int it)
    {
        Smem& s = *sm;
        const int row_0 = row0 + it, row_1 = row_0 + dRow;
        const int lrow_0 = row_0 + OC_MARCH_LAG, lrow_1 = row_1 + OC_MARCH_LAG;
        OcPV2 me, w1, n1;
        me.x = p_unpack3(me_x); me.v = p_unpack3(me_v); w1.x = p_unpack3(w1_x); w1.v = p_unpack3(w1_v);
        n1.x = p_unpack3(n1_x); n1.v = p_unpack3(n1_v);
        const OcPair3 k1 = p_unpack3(k1_q), k2b = p_unpack3(k2b_q);
        // ---- asynchronous global loads of rows lrow_0 / lrow_1 into the thread's landing zone -----------------
        // (columns outside the cloth, and on the generic path rows that are not loaded, get a benign far-away
        // particle at rest: every ring row then holds finite, non-degenerate data)
        const bool doL0 = (kSteady || (lrow_0 >= in_lo0 && lrow_0 < in_hi0)) && (kInterior || ok);
        const bool doL1 = (kSteady || (lrow_1 >= in_lo1 && lrow_1 < in_hi1)) && (kInterior || ok);
        {
            const long long o = goff0 + (long long)lrow_0 * U;
            if (doL0) { oc_cp_async16(&s.stage[0][i], A + o); oc_cp_async16(&s.stage[1][i], B + o); }
            else s.stage[0][i] = s.stage[1][i] = benign(ci, lrow_0);
            if (doL1) { oc_cp_async16(&s.stage[2][i], A + o + dOff); oc_cp_async16(&s.stage[3][i], B + o + dOff); }
            else s.stage[2][i] = s.stage[3][i] = benign(ci, lrow_1);
            oc_cp_async_commit();
        }
        const float2 rv1_j = rv1_n, rv2_j = rv2_n, dz2_j = dz2_n;
        {
            int ra_ = row_0 + 1, rb_ = row_1 + 1;
            if (!kSteady) { ra_ = ra_ < 0 ? 0 : (ra_ >= V ? V - 1 : ra_); rb_ = rb_ < 0 ? 0 : (rb_ >= V ? V - 1 : rb_); }
            rv1_n = row_pair(c.rv1, ra_, rb_); rv2_n = row_pair(c.rv2, ra_, rb_); dz2_n = row_pair(c.dz2, ra_, rb_);
        }
        const int sl = kSlot >= 0 ? kSlot : (it & (OC_RING - 1));
        const int s1 = (sl + 1) & (OC_RING - 1), s2 = (sl + 2) & (OC_RING - 1), s3 = (sl + 3) & (OC_RING - 1);
        const int h = sl & 1;

        // ---- P phase ---------------------------------------------------------------------------------
        const bool doP0 = kSteady || (row_0 >= plo0 && row_0 < hi0);
        const bool doP1 = kSteady || (row_1 >= plo1 && row_1 < hi1);
        const bool doP = doP0 | doP1;
        OcPair3 gH1, gH2, gV1, gV2, gD, gA, dme;
        OcPair3 Fo, xn;                          // kLean: the thread's own part of F, and X + (X - X_last)
        OcPV2 w2, nD;
        constexpr bool kLean = kSteady && !M::kExact && OC_TWIN_LEAN;
        if (doP) {
            if (!kSteady || !OC_TWIN_CARRY) { me = ld(sl, ci); w1 = ld(s1, ci); n1 = ld(sl, ci + 1); }
            w2 = ld(s2, ci);
            if (M::kExact) { constexpr int kD = M::kExact ? 1 : 0; dme.x = s.Dd[0 * kD][sl * kD][ci * kD]; dme.y = s.Dd[1 * kD][sl * kD][ci * kD]; dme.z = s.Dd[2 * kD][sl * kD][ci * kD]; }
            else dme = me.v;                                                 // fast mode: the ring holds d itself
            const OcPV2 n2 = ld(sl, ci + 2);
            nD = ld(s1, ci + 1);
            const OcPV2 nA = ld(s1, ci - 1);
            // exact mode: the operand ranges of the branch-free sqrt / division sequences, accumulated over the six
            // spring pairs and the two shear rest lengths (OcRange); one test for the whole iteration
            OcRange rg; rg.init();
            float2 rD = oc_sqrt2<M>(p_add(p_bc(dx2i), dz2_j), rg);          // cell (gi, row)
            float2 rA = oc_sqrt2<M>(p_add(p_bc(dx2m), dz2_j), rg);          // cell (gi-1, row)
            float2 rH1 = p_bc(rh1), rH2 = p_bc(rh2), rV1 = rv1_j, rV2 = rv2_j;
            const float2 nS = p_bc(c.nks_struct), nB = p_bc(c.nks_bend), nSh = p_bc(c.nks_shear);
            // fast mode: Kd / dt, because the "velocities" of the ring are X - X_last
            const float2 kS = p_bc(M::kExact ? c.kd_struct : kdt_struct), kB = p_bc(M::kExact ? c.kd_bend : kdt_bend), kSh = p_bc(M::kExact ? c.kd_shear : kdt_shear);
            if (!M::kExact) { rH1 = p_mul(rH1, nS); rH2 = p_mul(rH2, nB); rV1 = p_mul(rV1, nS); rV2 = p_mul(rV2, nB); rD = p_mul(rD, nSh); rA = p_mul(rA, nSh); }
            gH1 = oc_spring_twin<M>(me.x, me.v, n1.x, n1.v, rH1, nS, kS, c.one, rg);
            gH2 = oc_spring_twin<M>(me.x, me.v, n2.x, n2.v, rH2, nB, kB, c.one, rg);
            gV1 = oc_spring_twin<M>(me.x, me.v, w1.x, w1.v, rV1, nS, kS, c.one, rg);
            gV2 = oc_spring_twin<M>(me.x, me.v, w2.x, w2.v, rV2, nB, kB, c.one, rg);
            gD  = oc_spring_twin<M>(me.x, me.v, nD.x, nD.v, rD,  nSh, kSh, c.one, rg);
            gA  = oc_spring_twin<M>(me.x, me.v, nA.x, nA.v, rA,  nSh, kSh, c.one, rg);
            if (__builtin_expect(M::kExact && rg.bad(), 0)) {
                // rare: an operand left the exact range of the branch-free sequences -> all six pairs again with the
                // IEEE intrinsics (cold, out of line; operands re-read from shared memory)
#ifdef __CUDA_ARCH__
                if (c.dbg & 4) {                                            // development counters (OC_DEBUG=4)
                    atomicAdd(c.dbg_cnt, 1ull);
                    if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 1, 1ull);
                }
#endif
                gH1 = oc_twin_redo<M, Smem>(&c, sm, 0, sl, ci, rh1, rh1);
                gH2 = oc_twin_redo<M, Smem>(&c, sm, 1, sl, ci, rh2, rh2);
                gV1 = oc_twin_redo<M, Smem>(&c, sm, 2, sl, ci, rv1_j.x, rv1_j.y);
                gV2 = oc_twin_redo<M, Smem>(&c, sm, 3, sl, ci, rv2_j.x, rv2_j.y);
                gD  = oc_twin_redo<M, Smem>(&c, sm, 4, sl, ci, M::add(dx2i, dz2_j.x), M::add(dx2i, dz2_j.y));
                gA  = oc_twin_redo<M, Smem>(&c, sm, 5, sl, ci, M::add(dx2m, dz2_j.x), M::add(dx2m, dz2_j.y));
            }
            if (!kInterior || !kSteady) {
                // Window columns at a cloth edge: a spring to (or from) a column that does not exist is multiplied by
                // 0, every other one by 1 (exact); see oc_march2.cuh.  The ghost ends are finite far-away particles.
                const float2 m1 = p_bc(mR1), m2 = p_bc(mR2), ml = p_bc(mL1);
                gH1.x = p_mul(gH1.x, m1); gH1.y = p_mul(gH1.y, m1); gH1.z = p_mul(gH1.z, m1);
                gH2.x = p_mul(gH2.x, m2); gH2.y = p_mul(gH2.y, m2); gH2.z = p_mul(gH2.z, m2);
                gD.x  = p_mul(gD.x,  m1); gD.y  = p_mul(gD.y,  m1); gD.z  = p_mul(gD.z,  m1);
                gA.x  = p_mul(gA.x,  ml); gA.y  = p_mul(gA.y,  ml); gA.z  = p_mul(gA.z,  ml);
            }
            if (kLean) {
                // fast mode, steady rows: everything of F that this thread already has (see OC_TWIN_LEAN)
                const float2 damp = p_bc(damp_dt);
                Fo.x = p_fma(damp, me.v.x, p_bc(c.f0[0])); Fo.y = p_fma(damp, me.v.y, p_bc(c.f0[1])); Fo.z = p_fma(damp, me.v.z, p_bc(c.f0[2]));
                Fo.x = p_add(Fo.x, p_sub(gH1.x, k1.x));  Fo.y = p_add(Fo.y, p_sub(gH1.y, k1.y));  Fo.z = p_add(Fo.z, p_sub(gH1.z, k1.z));
                Fo.x = p_add(Fo.x, p_sub(gV1.x, k2b.x)); Fo.y = p_add(Fo.y, p_sub(gV1.y, k2b.y)); Fo.z = p_add(Fo.z, p_sub(gV1.z, k2b.z));
                Fo.x = p_add(Fo.x, p_add(gA.x, gD.x));   Fo.y = p_add(Fo.y, p_add(gA.y, gD.y));   Fo.z = p_add(Fo.z, p_add(gA.z, gD.z));
                Fo.x = p_add(Fo.x, p_add(gH2.x, gV2.x)); Fo.y = p_add(Fo.y, p_add(gH2.y, gV2.y)); Fo.z = p_add(Fo.z, p_add(gH2.z, gV2.z));
                if (!kInterior) {                                                                   // duplicated last bend spring of the row (V:313)
                    const float2 dA = p_bc(gi == U - 3 ? 1.0f : 0.0f);
                    Fo.x = p_fma(gH2.x, dA, Fo.x); Fo.y = p_fma(gH2.y, dA, Fo.y); Fo.z = p_fma(gH2.z, dA, Fo.z);
                }
                xn.x = p_add(me.x.x, dme.x); xn.y = p_add(me.x.y, dme.y); xn.z = p_add(me.x.z, dme.z);
            }
            // publish the forces whose partner lives in another thread
            s.FH1[0][h][ci] = gH1.x; s.FH1[1][h][ci] = gH1.y; s.FH1[2][h][ci] = gH1.z;
            s.FH2[0][h][ci] = gH2.x; s.FH2[1][h][ci] = gH2.y; s.FH2[2][h][ci] = gH2.z;
            s.FD[0][sl][ci] = gD.x;  s.FD[1][sl][ci] = gD.y;  s.FD[2][sl][ci] = gD.z;
            s.FA[0][sl][ci] = gA.x;  s.FA[1][sl][ci] = gA.y;  s.FA[2][sl][ci] = gA.z;
        }

        ctx.sync();

        // ---- G phase ---------------------------------------------------------------------------------
        const bool doG0 = kSteady || (row_0 >= lo0 && row_0 < hi0);
        const bool doG1 = kSteady || (row_1 >= lo1 && row_1 < hi1);
        if (doG0 | doG1) {
            constexpr bool kAll = kSteady;           // no predicates: edge columns are handled by the zero masks above
            const bool pin_0 = !kSteady && oc_pinned(c, bz0, gi, row_0), pin_1 = !kSteady && oc_pinned(c, bz1, gi, row_1);
            const bool e0 = !pin_0, e1 = !pin_1;                                  // springs act on the particle
            const bool l1 = kAll || gi - 1 >= 0, l2 = kAll || gi - 2 >= 0, r1 = kAll || gi + 1 < U, r2 = kAll || gi + 2 < U;
            const bool u1_0 = kSteady || row_0 - 1 >= 0, u2_0 = kSteady || row_0 - 2 >= 0, d1_0 = kSteady || row_0 + 1 < V, d2_0 = kSteady || row_0 + 2 < V;
            const bool u1_1 = kSteady || row_1 - 1 >= 0, u2_1 = kSteady || row_1 - 2 >= 0, d1_1 = kSteady || row_1 + 1 < V, d2_1 = kSteady || row_1 + 2 < V;
            OcPair3 F, n;
            if (kLean) {
                // the four forces received from other threads (their publishers' forces on themselves: subtract)
                OcPair3 a, b;
                a.x = s.FH1[0][h][ci - 1]; a.y = s.FH1[1][h][ci - 1]; a.z = s.FH1[2][h][ci - 1];
                b.x = s.FD[0][s3][ci - 1]; b.y = s.FD[1][s3][ci - 1]; b.z = s.FD[2][s3][ci - 1];
                a.x = p_add(a.x, b.x); a.y = p_add(a.y, b.y); a.z = p_add(a.z, b.z);
                b.x = s.FA[0][s3][ci + 1]; b.y = s.FA[1][s3][ci + 1]; b.z = s.FA[2][s3][ci + 1];
                OcPair3 t2;
                t2.x = s.FH2[0][h][ci - 2]; t2.y = s.FH2[1][h][ci - 2]; t2.z = s.FH2[2][h][ci - 2];
                b.x = p_add(b.x, t2.x); b.y = p_add(b.y, t2.y); b.z = p_add(b.z, t2.z);
                if (!kInterior) {                                                                   // duplicated last bend spring of the row (V:313)
                    const float2 dB = p_bc(gi == U - 1 ? 1.0f : 0.0f);
                    b.x = p_fma(t2.x, dB, b.x); b.y = p_fma(t2.y, dB, b.y); b.z = p_fma(t2.z, dB, b.z);
                }
                F.x = p_sub(Fo.x, p_add(a.x, b.x)); F.y = p_sub(Fo.y, p_add(a.y, b.y)); F.z = p_sub(Fo.z, p_add(a.z, b.z));
                n.x = p_fma(p_bc(c.dt2m), F.x, xn.x); n.y = p_fma(p_bc(c.dt2m), F.y, xn.y); n.z = p_fma(p_bc(c.dt2m), F.z, xn.z);
            } else {
            // F = 0 + gravity*mass (unless pinned) + DEFAULT_DAMPING*V     V:451-459
            F.x = make_float2(pin_0 ? 0.0f : c.f0[0], pin_1 ? 0.0f : c.f0[0]);
            F.y = make_float2(pin_0 ? 0.0f : c.f0[1], pin_1 ? 0.0f : c.f0[1]);
            F.z = make_float2(pin_0 ? 0.0f : c.f0[2], pin_1 ? 0.0f : c.f0[2]);
            const float2 damp = p_bc(M::kExact ? c.damping : damp_dt);            // fast mode: DEFAULT_DAMPING / dt, times X - X_last
            F.x = p_sump<M>(p_mul(damp, me.v.x), F.x, c.one);
            F.y = p_sump<M>(p_mul(damp, me.v.y), F.y, c.one);
            F.z = p_sump<M>(p_mul(damp, me.v.z), F.z, c.one);
            OcPair3 t;
            t.x = s.FH1[0][h][ci - 1]; t.y = s.FH1[1][h][ci - 1]; t.z = s.FH1[2][h][ci - 1];
            oc_acc2<M, kAll>(F, t,   e0 && l1, e1 && l1, true, c.one);                                     // 1  (i-1, j)   structural
            oc_acc2<M, kAll>(F, gH1, e0 && r1, e1 && r1, false, c.one);                                    // 2  (i+1, j)
            oc_acc2<M, kAll>(F, k1,  e0 && u1_0, e1 && u1_1, true, c.one);                                 // 3  (i, j-1)
            oc_acc2<M, kAll>(F, gV1, e0 && d1_0, e1 && d1_1, false, c.one);                                // 4  (i, j+1)
            t.x = s.FD[0][s3][ci - 1]; t.y = s.FD[1][s3][ci - 1]; t.z = s.FD[2][s3][ci - 1];
            oc_acc2<M, kAll>(F, t,   e0 && l1 && u1_0, e1 && l1 && u1_1, true, c.one);                     // 5  (i-1, j-1) shear
            t.x = s.FA[0][s3][ci + 1]; t.y = s.FA[1][s3][ci + 1]; t.z = s.FA[2][s3][ci + 1];
            oc_acc2<M, kAll>(F, t,   e0 && r1 && u1_0, e1 && r1 && u1_1, true, c.one);                     // 6  (i+1, j-1)
            oc_acc2<M, kAll>(F, gA,  e0 && l1 && d1_0, e1 && l1 && d1_1, false, c.one);                    // 7  (i-1, j+1)
            oc_acc2<M, kAll>(F, gD,  e0 && r1 && d1_0, e1 && r1 && d1_1, false, c.one);                    // 8  (i+1, j+1)
            t.x = s.FH2[0][h][ci - 2]; t.y = s.FH2[1][h][ci - 2]; t.z = s.FH2[2][h][ci - 2];
            oc_acc2<M, kAll>(F, t,   e0 && l2, e1 && l2, true, c.one);                                     // 9  (i-2, j)   bend
            oc_acc2<M, kAll>(F, gH2, e0 && r2, e1 && r2, false, c.one);                                    // 10 (i+2, j)
            if (!kAll) {                                                                            // 11 duplicated last bend spring of the row (V:313)
                oc_acc2<M, false>(F, gH2, e0 && gi == U - 3, e1 && gi == U - 3, false, c.one);
                oc_acc2<M, false>(F, t,   e0 && gi == U - 1, e1 && gi == U - 1, true, c.one);
            } else if (!kInterior) {                                                                // same, as 0/1 multipliers
                const float2 dA = p_bc(gi == U - 3 ? 1.0f : 0.0f), dB = p_bc(gi == U - 1 ? 1.0f : 0.0f);
                F.x = p_sump<M>(p_mul(gH2.x, dA), F.x, c.one); F.y = p_sump<M>(p_mul(gH2.y, dA), F.y, c.one); F.z = p_sump<M>(p_mul(gH2.z, dA), F.z, c.one);
                F.x = p_subp<M>(F.x, p_mul(t.x, dB), c.one);   F.y = p_subp<M>(F.y, p_mul(t.y, dB), c.one);   F.z = p_subp<M>(F.z, p_mul(t.z, dB), c.one);
            }
            oc_acc2<M, kAll>(F, k2b, e0 && u2_0, e1 && u2_1, true, c.one);                                 // 12 (i, j-2)
            oc_acc2<M, kAll>(F, gV2, e0 && d2_0, e1 && d2_1, false, c.one);                                // 13 (i, j+2)
            if (!kSteady) {                                                                         // 14 duplicated last bend spring of the column (V:319)
                oc_acc2<M, false>(F, gV2, e0 && row_0 == V - 3, e1 && row_1 == V - 3, false, c.one);
                oc_acc2<M, false>(F, k2b, e0 && row_0 == V - 1, e1 && row_1 == V - 1, true, c.one);
            }
            // ---- IntegrateVerlet (V:428-444) + EllipsoidCollision (V:509-533), both tiles ----------------
            n.x = p_sump<M>(p_mul(p_bc(c.dt2m), F.x), p_add(me.x.x, dme.x), c.one);
            n.y = p_sump<M>(p_mul(p_bc(c.dt2m), F.y), p_add(me.x.y, dme.y), c.one);
            n.z = p_sump<M>(p_mul(p_bc(c.dt2m), F.z), p_add(me.x.z, dme.z), c.one);
            }
            if (n.y.x < 0.0f) n.y.x = 0.0f;
            if (n.y.y < 0.0f) n.y.y = 0.0f;
            // A particle outside the collider's bounding sphere (OcConst::bs_*, conservative) cannot be inside the
            // ellipsoid: the transform of V:511-513 is skipped for it (most of the cloth, most of the time).
            bool hit_0 = false, hit_1 = false;
            const float2 ex = p_sub(n.x, p_bc(c.bs_c[0])), ey = p_sub(n.y, p_bc(c.bs_c[1])), ez = p_sub(n.z, p_bc(c.bs_c[2]));
            const float2 e2 = p_fma(ez, ez, p_fma(ey, ey, p_mul(ex, ex)));
            if ((e2.x <= c.bs_r2) | (e2.y <= c.bs_r2)) {
                OcPair3 p0;         // X_0 = inverse_ellipsoid * vec4(X,1) - center, rows x, y, z for (tile 0, tile 1)
                p0.x = p_sub(p_add(p_sump<M>(p_mul(p_bc(c.im[0][2]), n.z), p_sump<M>(p_mul(p_bc(c.im[0][1]), n.y), p_mul(p_bc(c.im[0][0]), n.x), c.one), c.one), p_bc(c.im[0][3])), p_bc(c.center[0]));
                p0.y = p_sub(p_add(p_sump<M>(p_mul(p_bc(c.im[1][2]), n.z), p_sump<M>(p_mul(p_bc(c.im[1][1]), n.y), p_mul(p_bc(c.im[1][0]), n.x), c.one), c.one), p_bc(c.im[1][3])), p_bc(c.center[1]));
                p0.z = p_sub(p_add(p_sump<M>(p_mul(p_bc(c.im[2][2]), n.z), p_sump<M>(p_mul(p_bc(c.im[2][1]), n.y), p_mul(p_bc(c.im[2][0]), n.x), c.one), c.one), p_bc(c.im[2][3])), p_bc(c.center[2]));
                const float2 sq = p_sump<M>(p_mul(p0.z, p0.z), p_sump<M>(p_mul(p0.y, p0.y), p_mul(p0.x, p0.x), c.one), c.one);
                hit_0 = sq.x < 1.0f; hit_1 = sq.y < 1.0f;                                           // V:513-514 (see oc_core.cuh)
#ifdef __CUDA_ARCH__
                if ((c.dbg & 4) && (hit_0 | hit_1)) { atomicAdd(c.dbg_cnt + 3, 1ull); if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 3, 1ull << 32); }
#endif
                if (__builtin_expect(hit_0 | hit_1, 0)) {
                    // EllipsoidCollision (V:514-530) of both particles at once, branch-free, with the same exact sqrt and
                    // division sequences as the springs; the result is taken per half where that particle is inside.
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
                        if (hit_0) { const f3 r = oc_march2_collide<M>(&c, make_f3(p0.x.x, p0.y.x, p0.z.x), sq.x, make_f3(n.x.x, n.y.x, n.z.x)); nn.x.x = r.x; nn.y.x = r.y; nn.z.x = r.z; }
                        if (hit_1) { const f3 r = oc_march2_collide<M>(&c, make_f3(p0.x.y, p0.y.y, p0.z.y), sq.y, make_f3(n.x.y, n.y.y, n.z.y)); nn.x.y = r.x; nn.y.y = r.y; nn.z.y = r.z; }
                    }
                    if (hit_0) { n.x.x = nn.x.x; n.y.x = nn.y.x; n.z.x = nn.z.x; }
                    if (hit_1) { n.x.y = nn.x.y; n.y.y = nn.y.y; n.z.y = nn.z.y; }
                }
            }
            const long long o = goff0 + (long long)row_0 * U;
            const float4 out_0 = make_float4(n.x.x, n.y.x, n.z.x, oc_u2f(hit_0 ? OC_W_HIT : OC_W_PLAIN));
            const float4 out_1 = make_float4(n.x.y, n.y.y, n.z.y, oc_u2f(hit_1 ? OC_W_HIT : OC_W_PLAIN));
            if (st && doG0) C[o] = out_0;
            if (st && doG1) C[o + dOff] = out_1;
            if (!kSteady) {
                // linked row bands: the first / last two rows of the band also go into the neighbour's halo (OcPeer2);
                // the steady range of a boundary tile excludes them, so the steady loop knows nothing of this
                const OcPeer2* pp = oc_opaque(peer);      // read at the point of use (see oc_opaque)
                if (st && (pp->c[0] || pp->c[1])) {
                    if (doG0) {
                        float4* pc = nullptr;
                        if (pp->c[0] && row_0 < pp->ra + 2) pc = pp->c[0];
                        if (pp->c[1] && row_0 >= pp->rb - 2) pc = pp->c[1];
                        if (pc) pc[(long long)row_0 * U + gi] = out_0;
                    }
                    if (doG1) {
                        float4* pc = nullptr;
                        if (pp->c[0] && row_1 < pp->ra + 2) pc = pp->c[0];
                        if (pp->c[1] && row_1 >= pp->rb - 2) pc = pp->c[1];
                        if (pc) pc[(long long)row_1 * U + gi] = out_1;
                    }
                }
            }
        }
        if (doP) {
            k2b_q = k2a_q; k2a_q = p_pack3(gV2); k1_q = p_pack3(gV1);
            if (OC_TWIN_CARRY) {
                me_x = p_pack3(w1.x); me_v = p_pack3(w1.v); w1_x = p_pack3(w2.x); w1_v = p_pack3(w2.v);
                n1_x = p_pack3(nD.x); n1_v = p_pack3(nD.v);
            }
        }

        // ---- publish the loaded rows (lrow = row + 4: same slot) ---------------------------------------
        // (generic path: every iteration, so that rows that are not loaded hold the benign particle)
        oc_cp_async_wait();
        publish(sl, s.stage[0][i], s.stage[1][i], s.stage[2][i], s.stage[3][i]);
    }
};

// Rows of the two tiles of CTA (strip bx, index k, z) and the cloths they belong to.
OC_HD void oc_twin_tiles(const OcSeg2& seg, const OcTwinMap& map, int bx, int k, int z, int ra, int rb,
                         int by[2], int bz[2], int r0[2], int r1[2])
{
    if (map.pair_cloths) { if (seg.rev) k = seg.nseg_all - 1 - k; by[0] = by[1] = k; bz[0] = 2 * z; bz[1] = 2 * z + 1; }
    else { if (seg.rev) k = seg.nseg_all / 2 - 1 - k; by[0] = 2 * k; by[1] = 2 * k + 1; bz[0] = bz[1] = z; }      // adjacent segments: consecutive launches then finish, and depend on each other, in the same order
    oc_seg2_rows(seg, bx, by[0], ra, rb, r0[0], r1[0]);
    oc_seg2_rows(seg, bx, by[1], ra, rb, r0[1], r1[1]);
}

#ifndef OC_TWIN_UNROLL_FAST
#define OC_TWIN_UNROLL_FAST 2
#endif
#ifndef OC_TWIN_UNROLL_EXACT
#define OC_TWIN_UNROLL_EXACT 1
#endif

template <class M, int WC, class Ctx>
OC_HD bool oc_twin_body(Ctx& ctx, const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                        float4* __restrict__ C, int ra, int rb, OcSeg2 seg, int x_halo, OcTwinMap map, const OcDep2& dep)
{
    OcTwin<M, WC, Ctx> m(ctx, c);
    m.A = A; m.B = B; m.C = C;
    m.sm = reinterpret_cast<OcSmemT<WC, M::kExact>*>(ctx.smem());
    const int i = ctx.tid();
    const int U = c.U, V = c.V;
    const int W_out = WC - 2 * x_halo;
    const int cx0 = ctx.bx() * W_out - x_halo;
    const int gi = cx0 + i;
    int by[2], bz[2], r0[2], r1[2];
    oc_twin_tiles(seg, map, ctx.bx(), ctx.by(), ctx.bz(), ra, rb, by, bz, r0, r1);
    if (r0[0] >= r1[0] && r0[1] >= r1[1]) return ctx.wait_deps_twin(dep, c, seg, by, bz, r0, r1);      // CTA-uniform: no rows at all
    m.i = i; m.ci = i + 2; m.gi = gi; m.U = U; m.V = V; m.bz0 = bz[0]; m.bz1 = bz[1];
    // rows of one tile: produced [lo, hi), spring phase from plo, loaded [plo, in_hi); an empty tile loads, evaluates and
    // stores nothing
    auto tile_rows = [&](int t0, int t1, int& lo, int& hi, int& plo, int& in_lo, int& in_hi) {
        lo = t0; hi = t1;
        plo = lo - 2; if (plo < 0) plo = 0;
        in_hi = hi + 2; if (in_hi > V) in_hi = V;
        if (t0 >= t1) { plo = lo; in_hi = lo; }
        in_lo = plo;
    };
    tile_rows(r0[0], r1[0], m.lo0, m.hi0, m.plo0, m.in_lo0, m.in_hi0);
    tile_rows(r0[1], r1[1], m.lo1, m.hi1, m.plo1, m.in_lo1, m.in_hi1);
    const int n_rows = (r1[0] - r0[0]) > (r1[1] - r0[1]) ? (r1[0] - r0[0]) : (r1[1] - r0[1]);
    const bool empty0 = r0[0] >= r1[0], empty1 = r0[1] >= r1[1];
    // both tiles start at iteration 0: row(it) = r0 - 2 - LAG + it
    const int f0 = empty0 ? r0[1] : r0[0], f1 = empty1 ? r0[0] : r0[1];      // (an empty tile follows the other one's rows, all predicates off)
    m.row0 = f0 - 2 - OC_MARCH_LAG;
    m.dRow = f1 - f0;
    const int n_it = n_rows + 2 + OC_MARCH_LAG;
    m.peer = &dep.peer;
    m.ok = gi >= 0 && gi < U;
    m.st = m.ok && i >= x_halo && i < WC - x_halo;
    auto clampc = [&](int g) { return g < 0 ? 0 : (g >= U ? U - 1 : g); };
    m.rh1 = OC_LDG(c.rh1 + clampc(gi)); m.rh2 = OC_LDG(c.rh2 + clampc(gi));
    m.dx2i = OC_LDG(c.dx2 + clampc(gi)); m.dx2m = OC_LDG(c.dx2 + clampc(gi - 1));
    m.mR1 = (m.ok && gi + 1 < U) ? 1.0f : 0.0f; m.mR2 = (m.ok && gi + 2 < U) ? 1.0f : 0.0f; m.mL1 = (m.ok && gi - 1 >= 0) ? 1.0f : 0.0f;
    m.ydt = oc_rcp_bf(c.dt);
    m.kdt_struct = c.kd_struct * c.inv_dt; m.kdt_shear = c.kd_shear * c.inv_dt; m.kdt_bend = c.kd_bend * c.inv_dt; m.damp_dt = c.damping * c.inv_dt;
    m.goff0 = (long long)bz[0] * c.cloth_stride - (long long)c.row_lo * U + gi;
    m.dOff = (long long)(bz[1] - bz[0]) * c.cloth_stride + (long long)m.dRow * U;
    {
        auto clampr = [&](int r) { return r < 0 ? 0 : (r >= V ? V - 1 : r); };
        const int ra_ = clampr(m.row0), rb_ = clampr(m.row0 + m.dRow);
        m.rv1_n = m.row_pair(c.rv1, ra_, rb_); m.rv2_n = m.row_pair(c.rv2, ra_, rb_); m.dz2_n = m.row_pair(c.dz2, ra_, rb_);
    }
    const float2 z2 = make_float2(0.f, 0.f);
    m.me_x.x = m.me_x.y = m.me_x.z = p_pack(z2);
    m.me_v = m.w1_x = m.w1_v = m.n1_x = m.n1_v = m.k1_q = m.k2a_q = m.k2b_q = m.me_x;

    // benign content for the pad columns (never written by a particle) and, for the rings, for every slot: the first
    // generic iterations read ring rows that have not been published yet
    {
        OcSmemT<WC, M::kExact>& s = *m.sm;
        for (int e = i; e < OC_RING * (WC + 4); e += WC) {
            const int slot = e / (WC + 4), col = e % (WC + 4);
            const float p = 1.0e3f + 8.0f * (float)col;
            for (int comp = 0; comp < 3; ++comp) {
                s.X[comp][slot][col] = make_float2(p, p);
                s.X[comp + 3][slot][col] = z2;
                if (M::kExact) { constexpr int kD = M::kExact ? 1 : 0; s.Dd[comp * kD][slot * kD][col * kD] = z2; }
                s.FD[comp][slot][col] = z2; s.FA[comp][slot][col] = z2;
                if (slot < 2) { s.FH1[comp][slot][col] = z2; s.FH2[comp][slot][col] = z2; }
            }
        }
    }

    // steady range: interior rows of BOTH tiles, all activities on
    int it_lo = 0, it_hi = n_it;
    auto steady_rows = [&](int bzh, int lo, int hi, int plo, int in_hi, int r00) {
        int st_lo = lo > plo + 1 ? lo : plo + 1; if (st_lo < 2) st_lo = 2;
        int st_hi = hi < V - 3 ? hi : V - 3;
        if (st_hi > in_hi - OC_MARCH_LAG) st_hi = in_hi - OC_MARCH_LAG;
        // linked row bands: the two rows pushed into a neighbour's halo are taken on the generic path
        if (dep.peer.c[0] && st_lo < dep.peer.ra + 2) st_lo = dep.peer.ra + 2;
        if (dep.peer.c[1] && st_hi > dep.peer.rb - 2) st_hi = dep.peer.rb - 2;
        if (st_lo < st_hi && !oc_rows_unpinned(c, bzh, st_lo, st_hi)) st_hi = st_lo;      // custom pins (oc_set_pins) in these rows: generic path only
        if (st_lo - r00 > it_lo) it_lo = st_lo - r00;
        if (st_hi - r00 < it_hi) it_hi = st_hi - r00;
    };
    steady_rows(bz[0], m.lo0, m.hi0, m.plo0, m.in_hi0, m.row0);
    steady_rows(bz[1], m.lo1, m.hi1, m.plo1, m.in_hi1, m.row0 + m.dRow);
    if (empty0 || empty1) it_lo = it_hi = n_it;                 // an empty tile: generic path only (the planner avoids it)
    if (it_lo < 0) it_lo = 0;
    if (it_hi > n_it) it_hi = n_it;
    if (it_hi <= it_lo) it_lo = it_hi = n_it;
    const bool interior = cx0 >= 2 && cx0 + WC + 2 <= U;          // CTA-uniform

#ifdef __CUDA_ARCH__
    if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 1);       // development: CTA timeline (OC_DEBUG=8)
#endif
    if (!ctx.wait_deps_twin(dep, c, seg, by, bz, r0, r1)) return false;
    // the published ring rows must be visible before the first generic iteration reads them
    ctx.sync();
    int it = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const int end = phase == 0 ? it_lo : n_it;
        for (; it < end; ++it) m.template iter<false, false, -1>(it);
#ifdef __CUDA_ARCH__
        if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, phase == 0 ? 2 : 4);
#endif
        if (phase == 0 && it < it_hi) {
            m.load_carried(it);
            constexpr int kUn = M::kExact ? OC_TWIN_UNROLL_EXACT : OC_TWIN_UNROLL_FAST;
            if (interior) {
                if (kUn == 4) {
                    for (; it < it_hi && (it & 3) != 0; ++it) m.template iter<true, true, -1>(it);
                    for (; it + 3 < it_hi; it += 4) {
                        m.template iter<true, true, 0>(it); m.template iter<true, true, 1>(it + 1);
                        m.template iter<true, true, 2>(it + 2); m.template iter<true, true, 3>(it + 3);
                    }
                } else if (kUn == 2) {
                    for (; it + 1 < it_hi; it += 2) { m.template iter<true, true, -1>(it); m.template iter<true, true, -1>(it + 1); }
                } else if (kUn == 3) {
                    for (; it + 2 < it_hi; it += 3) { m.template iter<true, true, -1>(it); m.template iter<true, true, -1>(it + 1); m.template iter<true, true, -1>(it + 2); }
                } else if (kUn == 6) {
                    for (; it + 5 < it_hi; it += 6) {
                        m.template iter<true, true, -1>(it);     m.template iter<true, true, -1>(it + 1); m.template iter<true, true, -1>(it + 2);
                        m.template iter<true, true, -1>(it + 3); m.template iter<true, true, -1>(it + 4); m.template iter<true, true, -1>(it + 5);
                    }
                }
                for (; it < it_hi; ++it) m.template iter<true, true, -1>(it);
            } else {
                for (; it < it_hi; ++it) m.template iter<true, false, -1>(it);
            }
#ifdef __CUDA_ARCH__
            if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 3);
#endif
        }
    }
    return true;
}

#ifdef __CUDACC__
struct OcDevCtxT {          // grid = (strips * tile pairs per strip, 1, batch or batch / 2)
    int x, y;
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int bx() const { return x; }
    __device__ __forceinline__ int by() const { return y; }
    __device__ __forceinline__ int bz() const { return blockIdx.z; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ unsigned char* smem() const { extern __shared__ __align__(16) unsigned char oc_dyn_smem[]; return oc_dyn_smem; }
    // split-phase CTA barrier (mbarrier in shared memory): arrive now, wait for everybody's arrival later (oc_stream.cuh)
    __device__ __forceinline__ void bar_init(unsigned long long* b, int count) const
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void bar_arrive(unsigned long long* b) const
    {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
    }
    __device__ __forceinline__ void bar_wait(unsigned long long* b, unsigned parity) const
    {
        asm volatile("{\n\t.reg .pred p;\n\tOC_BAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra OC_BAR_DONE;\n\tbra OC_BAR_WAIT;\n\tOC_BAR_DONE:\n\t}"
                     :: "r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
    }
    // like OcDevCtx2::wait_deps, for both tiles: warp 0 polls for tile 0, warp 1 for tile 1
    __device__ __forceinline__ bool wait_deps_twin(const OcDep2& d, const OcConst& c, const OcSeg2& seg, const int by[2], const int bz[2],
                                                   const int r0[2], const int r1[2]) const
    {
        (void)seg; (void)by;
        const int t = threadIdx.x & 31, hh = threadIdx.x >> 5;
        bool ok = true;
        if (d.mode == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        else if (hh < 2 && t < 12) {
            const int idx = oc_dep2_index(d, x - 1 + t / 4, t % 4, r0[hh], r1[hh]);
            if (idx >= 0)
                ok = oc_flag_wait<false>(c, d.flags + (size_t)bz[hh] * oc_seg2_tiles(d.pseg) + idx, d.epoch - 1u + ((c.dbg & 32) ? 1000u : 0u));
        }
        if (hh < 2 && t >= 16 && t < 22) {
            const int side = (t - 16) / 3, xs = x - 1 + (t - 16) % 3;
            const bool reads_halo = r1[hh] > r0[hh] && (side == 0 ? r0[hh] < d.peer.ra + 2 : r1[hh] > d.peer.rb - 2);
            if (d.peer.flags_in[side] && reads_halo && xs >= 0 && xs < d.peer.nstrips)
                ok = oc_flag_wait<true>(c, d.peer.flags_in[side] + xs, d.peer.epoch - 1u);
        }
        return __syncthreads_and(ok) != 0;
    }
    // after the CTA's last store: the flags of both tiles
    __device__ __forceinline__ void publish(const OcDep2& d, const OcSeg2& seg, const int by[2], const int bz[2], const int r0[2], const int r1[2]) const
    {
        if (!d.flags) return;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const int tiles = oc_seg2_tiles(seg);
            for (int hh = 0; hh < 2; ++hh)
                asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(d.flags + (size_t)bz[hh] * tiles + oc_seg2_index(seg, x, by[hh])), "r"(d.epoch) : "memory");
            const OcPeer2* pp = oc_opaque(&d.peer);
            bool up = false, dn = false;
            for (int hh = 0; hh < 2; ++hh) {
                up |= pp->flags_out[0] && r1[hh] > r0[hh] && r0[hh] < pp->ra + 2;
                dn |= pp->flags_out[1] && r1[hh] > r0[hh] && r1[hh] > pp->rb - 2;
            }
            if (up | dn) {
                __threadfence_system();
                if (up) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pp->flags_out[0] + x), "r"(pp->epoch) : "memory");
                if (dn) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pp->flags_out[1] + x), "r"(pp->epoch) : "memory");
            }
        }
    }
};
// MINB = resident CTAs per SM the register allocation is capped for (2 x 128 or 4 x 64 threads: 255 registers)
template <class M, int WC, int MINB>
__global__ void __launch_bounds__(WC, MINB)
oc_k_twin(const __grid_constant__ OcConst c, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C,
          int ra, int rb, OcSeg2 seg, int x_halo, OcTwinMap map, const __grid_constant__ OcDep2 dep)
{
    asm volatile("griddepcontrol.launch_dependents;");      // the next step may be placed as soon as CTA slots free up
    if ((c.dbg & 8) && threadIdx.x == 0) oc_timeline_mark(c, 0);
    OcDevCtxT ctx;
    ctx.x = blockIdx.x % seg.nstrips; ctx.y = blockIdx.x / seg.nstrips;
    if (!oc_twin_body<M, WC, OcDevCtxT>(ctx, c, A, B, C, ra, rb, seg, x_halo, map, dep)) return;      // a dependency timed out: nothing published
    int by[2], bz[2], r0[2], r1[2];
    oc_twin_tiles(seg, map, ctx.x, ctx.y, blockIdx.z, ra, rb, by, bz, r0, r1);
    ctx.publish(dep, seg, by, bz, r0, r1);
}
#endif

// ---- host side (oc_march.cu) -------------------------------------------------------------------
int  oc_twin_configure(int device);
int  oc_twin_plan(const OcConst& c, bool exact, bool chained, bool linked, int ra, int rb, int sm_count, int occ_hint, OcMarchPlan* plan, OcSeg2* seg, OcTwinMap* map, int variant = 0);
int  oc_twin_nstrips(int nx);
#ifdef __CUDACC__
cudaError_t oc_twin_launch(const OcConst& c, bool exact, int ra, int rb, int sm_count,
                           const float4* A, const float4* B, float4* C, cudaStream_t stream, int* n_launches, OcChain2* chain,
                           const OcPeer2* peer = nullptr, int variant = 0);      // variant 1: oc_k_stream (oc_stream.cuh)
#endif
