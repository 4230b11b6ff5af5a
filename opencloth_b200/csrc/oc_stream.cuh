// oc_stream.cuh — kernel 6: the streaming GATHER kernel over twin tiles (one substep per launch).
//
// oc_k_march2 / oc_k_twin evaluate every spring once and hand the force to the partner: the minimum of arithmetic, paid
// for with two phases per row, a force exchange through shared memory, a large per-thread state (two rows of a column,
// the carried vertical forces, six spring results live across a barrier: ~250 registers) and therefore two warps per
// scheduler — which is what bounds them: no pipe is more than half busy.
//
// This kernel makes the opposite trade.  A thread owns one particle of the row (times two tiles, see oc_twin.cuh: the
// halves of every packed FP32x2 operation are the same particle position of two independent tiles) and evaluates ALL of
// its twelve springs itself, like the reference's force loop seen from the particle (and like oc_k_gather): twice the
// spring arithmetic, but no force exchange, no second phase, nothing carried from row to row except the shared-memory
// ring of rows, and a per-thread state of one particle.  The FMA pipe has the room (36 % busy in the fast kernel), the
// registers it frees buy resident warps.
//
//   * rows stream through an 8-row ring in shared memory (float2 = tile 0, tile 1 per column and component): row r+3
//     is requested with cp.async at the top of the iteration that computes row r and published at its end;
//   * row r is computed from rows r-2 .. r+2 of the ring: 12 neighbours x 6 components, all LDS.64; the per-row rest
//     lengths ride in the ring as well, the per-column ones live in registers;
//   * forces are accumulated in the order the reference's spring list touches the particle (oc_gather.cuh), so exact
//     mode is bit-identical; fast mode fuses F += s * dp and keeps X - X_last instead of V (1/dt folded into Kd);
//   * one SPLIT barrier per row: a thread arrives when it has published row r+3 and only waits just before the last
//     spring of the next row - the one that reads row r+3 (everything else it reads was published two or more
//     iterations ago); steady loop without predicates (edge columns: 0/1 multipliers), generic path with per-half
//     row predicates for edge rows, pipeline fill, pinned rows and the rows a linked band pushes to its neighbour.
// Tiles, launch chaining, linked bands and batches are those of oc_k_twin (OcSeg2 / OcDep2 / OcPeer2 / OcTwinMap).
#pragma once
#include "oc_core.cuh"
#include "oc_march.cuh"
#include "oc_march2.cuh"
#include "oc_twin.cuh"

#ifndef OC_SRING
#define OC_SRING 8                  // ring depth: rows r-2 .. r+3 are live.  (6 rows would fit a fourth CTA of 128 threads per SM: measured
                                    // equal, 75.0 against 75.0 G updates/s at 8192^2; so was requesting the first five rows of a tile together)
#endif
// slot of the row that iteration `it` + k touches; it + k >= 0
OC_HD int oc_sslot(int x) { return (OC_SRING & (OC_SRING - 1)) == 0 ? (x & (OC_SRING - 1)) : x % OC_SRING; }
#define OC_STREAM_AHEAD 3           // the row PUBLISHED at the end of the iteration that computes row r is r + 3
// Rows in flight: the row REQUESTED (cp.async) at the top of that iteration is r + 3 + OC_STREAM_DEPTH, into one of
// OC_STREAM_DEPTH + 1 landing zones; the thread then waits for all but the newest OC_STREAM_DEPTH groups.  With depth 0
// a row has one iteration to arrive - which is enough:
#ifndef OC_STREAM_DEPTH
#define OC_STREAM_DEPTH 0          // measured on B200: 0: 72.6, 1: 70.7, 2: 68.8, 4: 62 G updates/s (8192^2, fast): the rows are not what it waits for
#endif

// OC_STREAM_WIN (an experiment of round 2, off by default): in fast mode the steady loop keeps rows r-2 .. r+2 of the thread's
// OWN column in registers (the thread published them itself) and reads only the other columns from the ring: 48 instead of
// 78 LDS.64 per iteration, and the barrier wait moves from "before the last spring" to "before publishing" (rows r-2 / r+2
// are only read in the own column).  1: one copy of the loop, the window moves (48 MOV); 2: loop unrolled by five, the window
// rotates by renaming.  Same bits.  Measured on B200 (profiles/r2/stream_window_rates.log, r2_stream_fast_window_2048.md):
// 27 % fewer shared-memory wavefronts and NO gain - 1: 64.1 / 74.4 against 66.7 / 75.4 G updates/s (2048^2 / 8192^2);
// 2: 34.8 / 53.0 (a 54 KB loop body: instruction fetch).  So the shared-memory data pipe is not what bounds this kernel.
#ifndef OC_STREAM_WIN
#define OC_STREAM_WIN 0
#endif

template <int WC, bool kExact>
struct OcSmemS {
    float2 X[kExact ? 9 : 6][OC_SRING][WC + 4];      // x, y, z, then vx, vy, vz, dx, dy, dz (exact) or dx, dy, dz (fast); slot = iteration mod OC_SRING
    float2 RC[3][OC_SRING];                          // rv1, rv2, dz2 of the rows in this slot (tile 0, tile 1)
    float4 stage[OC_STREAM_DEPTH + 1][4][WC];        // landing zones of the asynchronous row loads: A, B of tile 0, A, B of tile 1
    unsigned long long bar;                          // split-phase barrier of the row pipeline (mbarrier)
};

// a * b + c.  Exact mode: product and sum rounded separately (see p_sump); fast mode: one explicit FFMA2, so that the steady
// and the generic path of the kernel round alike and a result never depends on the tiling
template <class M> OC_HD float2 oc_ma(float2 a, float2 b, float2 c, float one) { return M::kExact ? p_sump<M>(p_mul(a, b), c, one) : p_fma(a, b, c); }

// per-column rest lengths of a thread (exact: lengths / squared lengths; fast: lengths are pre-multiplied by -Ks)
struct OcStreamCol { float rh1m, rh1i, rh2m, rh2i, dx2m, dx2i; };
// per-row rest lengths of one tile
struct OcStreamRow { float rv1m, rv1, rv2m, rv2, dz2m, dz2; };

// neighbour k of the reference's order: 0 (-1,0) 1 (+1,0) 2 (0,-1) 3 (0,+1) 4 (-1,-1) 5 (+1,-1) 6 (-1,+1) 7 (+1,+1) 8 (-2,0) 9 (+2,0)
// 10 (+2,0) again [U-3] 11 (-2,0) again [U-1] 12 (0,-2) 13 (0,+2) 14 (0,+2) again [V-3] 15 (0,-2) again [V-1]      (dc, dr)
OC_HD void oc_stream_nbr(int k, int& dc, int& dr)
{
    const int dcs[16] = { -1, 1, 0, 0, -1, 1, -1, 1, -2, 2, 2, -2, 0, 0, 0, 0 };
    const int drs[16] = { 0, 0, -1, 1, -1, -1, 1, 1, 0, 0, 0, 0, -2, 2, 2, -2 };
    dc = dcs[k]; dr = drs[k];
}

// Cold path of exact mode: the whole force sum of ONE tile's particle with the IEEE intrinsics, from shared memory.
// `on`: bit k set = neighbour k acts (exists, particle not pinned).
template <class M, class S>
OC_COLD f3 oc_stream_force_slow(const OcConst* c, const S* s, int half, int it, int ci, unsigned on, bool pinned,
                                OcStreamCol K, OcStreamRow R)
{
#define OC_SX(comp, slot, col) (half ? s->X[comp][slot][col].y : s->X[comp][slot][col].x)
    const int s0 = oc_sslot(it + OC_SRING);
    const f3 xm = make_f3(OC_SX(0, s0, ci), OC_SX(1, s0, ci), OC_SX(2, s0, ci));
    const f3 vm = make_f3(OC_SX(3, s0, ci), OC_SX(4, s0, ci), OC_SX(5, s0, ci));
    f3 F = oc_base_force<M>(*c, vm, pinned);
    for (int k = 0; k < 16; ++k) {
        if (!((on >> k) & 1u)) continue;
        int dc, dr;
        oc_stream_nbr(k, dc, dr);
        const int sl = oc_sslot(it + dr + OC_SRING), cc = ci + dc;
        float rest, nks = c->nks_struct, kd = c->kd_struct;
        switch (k) {
        case 0: rest = K.rh1m; break;
        case 1: rest = K.rh1i; break;
        case 2: rest = R.rv1m; break;
        case 3: rest = R.rv1; break;
        case 4: rest = M::sqrt(M::add(K.dx2m, R.dz2m)); nks = c->nks_shear; kd = c->kd_shear; break;
        case 5: rest = M::sqrt(M::add(K.dx2i, R.dz2m)); nks = c->nks_shear; kd = c->kd_shear; break;
        case 6: rest = M::sqrt(M::add(K.dx2m, R.dz2)); nks = c->nks_shear; kd = c->kd_shear; break;
        case 7: rest = M::sqrt(M::add(K.dx2i, R.dz2)); nks = c->nks_shear; kd = c->kd_shear; break;
        case 8: case 11: rest = K.rh2m; nks = c->nks_bend; kd = c->kd_bend; break;
        case 9: case 10: rest = K.rh2i; nks = c->nks_bend; kd = c->kd_bend; break;
        case 12: case 15: rest = R.rv2m; nks = c->nks_bend; kd = c->kd_bend; break;
        default: rest = R.rv2; nks = c->nks_bend; kd = c->kd_bend; break;
        }
        const f3 f = oc_spring<M>(xm, vm, make_f3(OC_SX(0, sl, cc), OC_SX(1, sl, cc), OC_SX(2, sl, cc)),
                                  make_f3(OC_SX(3, sl, cc), OC_SX(4, sl, cc), OC_SX(5, sl, cc)), rest, nks, kd);
        F.x = M::add(F.x, f.x); F.y = M::add(F.y, f.y); F.z = M::add(F.z, f.z);
    }
#undef OC_SX
    return F;
}

template <class M, int WC, class Ctx>
struct OcStream {
    typedef OcSmemS<WC, M::kExact> Smem;
    Ctx& ctx;
    const OcConst& c;
    const float4* __restrict__ A; const float4* __restrict__ B;
    float4* __restrict__ C;
    Smem* sm;
    int i, ci, gi, U, V;
    int bz0, bz1;
    int lo0, hi0, in_lo0, in_hi0;            // rows of tile 0: produced [lo, hi), loaded [in_lo, in_hi)
    int lo1, hi1, in_lo1, in_hi1;
    int row0, dRow;                          // row of tile 0 computed at iteration 0; tile 1 is dRow rows further down
    int it_first;                            // first iteration of the loop
    bool ok, st;                             // column exists / column is stored by this CTA
    OcStreamCol K;                           // exact: rest lengths; fast: rest lengths times -Ks (dx2 plain)
    float mL1, mL2, mR1, mR2;                // 1 if the neighbour column -1 / -2 / +1 / +2 exists (and this one does), else 0
    float ydt, kdt_struct, kdt_shear, kdt_bend, damp_dt;
    long long goff0, dOff;
    const OcPeer2* peer;
    static constexpr bool kWin = !M::kExact && OC_STREAM_WIN;
    OcPV2 W0, W1, W2, W3, W4;                // register window of the own column (kWin): sub-iteration p keeps row r-2+k in W[(p+k) mod 5]
    template <int K> OC_HD OcPV2& win() { return K == 0 ? W0 : (K == 1 ? W1 : (K == 2 ? W2 : (K == 3 ? W3 : W4))); }

    OC_HD OcStream(Ctx& ctx_, const OcConst& c_) : ctx(ctx_), c(c_) {}

    OC_HD OcPV2 ld(int slot, int col) const
    {
        OcPV2 r;
        r.x.x = sm->X[0][slot][col]; r.x.y = sm->X[1][slot][col]; r.x.z = sm->X[2][slot][col];
        r.v.x = sm->X[3][slot][col]; r.v.y = sm->X[4][slot][col]; r.v.z = sm->X[5][slot][col];
        return r;
    }
    static OC_HD float4 benign(int ci_, int lrow) { return make_float4(1.0e3f + 8.0f * (float)ci_, 1.0e3f, 1.0e3f + 8.0f * (float)(lrow & 63), oc_u2f(OC_W_PLAIN)); }

    // One loaded row (both tiles) into ring slot sl: position, then X - X_last (fast) or V and X - X_last (exact)
    OC_HD OcPV2 publish(int sl, const float4 la0, const float4 lq0, const float4 la1, const float4 lq1)
    {
        Smem& s = *sm;
        OcPV2 r;
        OcPair3 d;
        d.x = make_float2(M::sub(la0.x, lq0.x), M::sub(la1.x, lq1.x));
        d.y = make_float2(M::sub(la0.y, lq0.y), M::sub(la1.y, lq1.y));
        d.z = make_float2(M::sub(la0.z, lq0.z), M::sub(la1.z, lq1.z));
        if (oc_hit(la0.w)) { d.x.x = 0.0f; d.y.x = 0.0f; d.z.x = 0.0f; }       // X_last == X (V:530)
        if (oc_hit(la1.w)) { d.x.y = 0.0f; d.y.y = 0.0f; d.z.y = 0.0f; }
        r.x.x = make_float2(la0.x, la1.x); r.x.y = make_float2(la0.y, la1.y); r.x.z = make_float2(la0.z, la1.z);
        r.v = d;
        s.X[0][sl][ci] = r.x.x;
        s.X[1][sl][ci] = r.x.y;
        s.X[2][sl][ci] = r.x.z;
        if (!M::kExact) { s.X[3][sl][ci] = d.x; s.X[4][sl][ci] = d.y; s.X[5][sl][ci] = d.z; return r; }
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
        constexpr int kE = M::kExact ? 1 : 0;      // (indices that exist in both layouts; reached in exact mode only)
        s.X[3][sl][ci] = v.x; s.X[4][sl][ci] = v.y; s.X[5][sl][ci] = v.z;
        s.X[6 * kE][sl][ci] = d.x; s.X[7 * kE][sl][ci] = d.y; s.X[8 * kE][sl][ci] = d.z;
        r.v = v;
        return r;
    }

    // spring of the particle pair `me` with neighbour pair n; adds the force to F.  p0 / p1: the spring acts on tile 0 / 1
    // (kAll: both, no predicates); mask: 0/1 multiplier for edge-column CTAs of the steady loop (kMask)
    template <bool kAll, bool kMask>
    OC_HD void spring(OcPair3& F, const OcPV2& me, const OcPV2& n, float2 rest, float2 nks, float2 kd, bool p0, bool p1, float mask, OcRange& rg)
    {
        if (M::kExact) {
            OcPair3 f = oc_spring_twin<M>(me.x, me.v, n.x, n.v, rest, nks, kd, c.one, rg);
            if (kMask) { const float2 m = p_bc(mask); f.x = p_mul(f.x, m); f.y = p_mul(f.y, m); f.z = p_mul(f.z, m); }
            oc_acc2<M, kAll>(F, f, p0, p1, false, c.one);
        } else {
            // fast mode: F += s * dp fused, 15 packed operations and two MUFU per spring pair; the generic path applies the
            // same fused operation per half under its predicate, so both paths round alike
            OcPair3 dp, dv;
            dp.x = p_sub(me.x.x, n.x.x); dp.y = p_sub(me.x.y, n.x.y); dp.z = p_sub(me.x.z, n.x.z);
            dv.x = p_sub(me.v.x, n.v.x); dv.y = p_sub(me.v.y, n.v.y); dv.z = p_sub(me.v.z, n.v.z);
            const float2 sqr  = p_fma(dp.z, dp.z, p_fma(dp.y, dp.y, p_mul(dp.x, dp.x)));
            const float2 rinv = p_rsq(sqr);
            const float2 dot  = p_fma(dv.z, dp.z, p_fma(dv.y, dp.y, p_mul(dv.x, dp.x)));
            const float2 u    = p_fma(p_mul(kd, dot), rinv, p_neg(rest));
            float2 sc         = p_fma(u, rinv, nks);
            if (kMask) sc = p_mul(sc, p_bc(mask));
            if (kAll) { F.x = p_fma(sc, dp.x, F.x); F.y = p_fma(sc, dp.y, F.y); F.z = p_fma(sc, dp.z, F.z); }
            else {
                if (p0) { F.x.x = MathFast::fma(sc.x, dp.x.x, F.x.x); F.y.x = MathFast::fma(sc.x, dp.y.x, F.y.x); F.z.x = MathFast::fma(sc.x, dp.z.x, F.z.x); }
                if (p1) { F.x.y = MathFast::fma(sc.y, dp.x.y, F.x.y); F.y.y = MathFast::fma(sc.y, dp.y.y, F.y.y); F.z.y = MathFast::fma(sc.y, dp.z.y, F.z.y); }
            }
        }
    }

    // kPh >= 0: sub-iteration kPh of the unrolled steady loop with the register window; -1: everything from the ring
    template <bool kSteady, bool kInterior, int kPh = -1>
    OC_HD void iter(int it)
    {
        static_assert(kPh < 0 || (kSteady && kWin), "the register window belongs to the steady loop of fast mode");
        constexpr bool kW = kPh >= 0;
        constexpr int kP = kW ? kPh : 0;
        Smem& s = *sm;
        const int row_0 = row0 + it, row_1 = row_0 + dRow;
        const int prow_0 = row_0 + OC_STREAM_AHEAD, prow_1 = row_1 + OC_STREAM_AHEAD;                 // rows published by this iteration
        const int lrow_0 = prow_0 + OC_STREAM_DEPTH, lrow_1 = prow_1 + OC_STREAM_DEPTH;               // rows requested by this iteration
        // ---- asynchronous global loads of rows lrow_0 / lrow_1 into this iteration's landing zone ---------------
        // (the steady loop only guarantees the PUBLISHED rows: near the end of a tile there is nothing left to request)
        const bool doL0 = lrow_0 >= in_lo0 && lrow_0 < in_hi0 && (kInterior || ok);
        const bool doL1 = lrow_1 >= in_lo1 && lrow_1 < in_hi1 && (kInterior || ok);
        const int zl = it % (OC_STREAM_DEPTH + 1);
        {
            const long long o = goff0 + (long long)lrow_0 * U;
            if (doL0) { oc_cp_async16(&s.stage[zl][0][i], A + o); oc_cp_async16(&s.stage[zl][1][i], B + o); }
            if (doL1) { oc_cp_async16(&s.stage[zl][2][i], A + o + dOff); oc_cp_async16(&s.stage[zl][3][i], B + o + dOff); }
            oc_cp_async_commit();
        }
        // ring slots of rows row-2 .. row+3: one modulo, five wrap-arounds
        const int s0 = oc_sslot(it);
        auto wrap = [](int x) { return x >= OC_SRING ? x - OC_SRING : x; };
        const int sp1 = wrap(s0 + 1), sp2 = wrap(s0 + 2), sp3 = wrap(s0 + OC_STREAM_AHEAD), sm2 = wrap(s0 + OC_SRING - 2), sm1 = wrap(s0 + OC_SRING - 1);

        // ---- the step of row `row` of both tiles ----------------------------------------------------------------
        const bool doG0 = kSteady || (row_0 >= lo0 && row_0 < hi0);
        const bool doG1 = kSteady || (row_1 >= lo1 && row_1 < hi1);
        bool waited = false;
        if (doG0 | doG1) {
            constexpr bool kAll = kSteady;
            constexpr bool kMask = kSteady && !kInterior;
            const OcPV2 me = kW ? win<(kP + 2) % 5>() : ld(s0, ci);
            OcPair3 dme;
            if (M::kExact) { constexpr int kE = M::kExact ? 1 : 0; dme.x = s.X[6 * kE][s0][ci]; dme.y = s.X[7 * kE][s0][ci]; dme.z = s.X[8 * kE][s0][ci]; }
            else dme = me.v;
            // per-row rest lengths of rows row-2 .. row (they ride in the ring)
            const float2 rv1m = s.RC[0][sm1], rv1 = s.RC[0][s0], rv2m = s.RC[1][sm2], rv2 = s.RC[1][s0], dz2m = s.RC[2][sm1], dz2 = s.RC[2][s0];
            const bool pin_0 = !kSteady && oc_pinned(c, bz0, gi, row_0), pin_1 = !kSteady && oc_pinned(c, bz1, gi, row_1);
            const bool e0 = !pin_0 && doG0, e1 = !pin_1 && doG1;
            const bool l1 = kAll || gi - 1 >= 0, l2 = kAll || gi - 2 >= 0, r1 = kAll || gi + 1 < U, r2 = kAll || gi + 2 < U;
            const bool u1_0 = kSteady || row_0 - 1 >= 0, u2_0 = kSteady || row_0 - 2 >= 0, d1_0 = kSteady || row_0 + 1 < V, d2_0 = kSteady || row_0 + 2 < V;
            const bool u1_1 = kSteady || row_1 - 1 >= 0, u2_1 = kSteady || row_1 - 2 >= 0, d1_1 = kSteady || row_1 + 1 < V, d2_1 = kSteady || row_1 + 2 < V;
            const float2 nS = p_bc(c.nks_struct), nB = p_bc(c.nks_bend), nSh = p_bc(c.nks_shear);
            const float2 kS = p_bc(M::kExact ? c.kd_struct : kdt_struct), kB = p_bc(M::kExact ? c.kd_bend : kdt_bend), kSh = p_bc(M::kExact ? c.kd_shear : kdt_shear);
            OcRange rg; rg.init();
            // shear rest lengths of the four cells round the particle
            float2 rUL = oc_sqrt2<M>(p_add(p_bc(K.dx2m), dz2m), rg), rUR = oc_sqrt2<M>(p_add(p_bc(K.dx2i), dz2m), rg);
            float2 rLL = oc_sqrt2<M>(p_add(p_bc(K.dx2m), dz2), rg),  rLR = oc_sqrt2<M>(p_add(p_bc(K.dx2i), dz2), rg);
            float2 tV1m = rv1m, tV1 = rv1, tV2m = rv2m, tV2 = rv2;
            if (!M::kExact) { rUL = p_mul(rUL, nSh); rUR = p_mul(rUR, nSh); rLL = p_mul(rLL, nSh); rLR = p_mul(rLR, nSh);
                              tV1m = p_mul(tV1m, nS); tV1 = p_mul(tV1, nS); tV2m = p_mul(tV2m, nB); tV2 = p_mul(tV2, nB); }
            // F = 0 + gravity*mass (unless pinned) + DEFAULT_DAMPING*V     V:451-459
            OcPair3 F;
            F.x = make_float2(pin_0 ? 0.0f : c.f0[0], pin_1 ? 0.0f : c.f0[0]);
            F.y = make_float2(pin_0 ? 0.0f : c.f0[1], pin_1 ? 0.0f : c.f0[1]);
            F.z = make_float2(pin_0 ? 0.0f : c.f0[2], pin_1 ? 0.0f : c.f0[2]);
            const float2 damp = p_bc(M::kExact ? c.damping : damp_dt);
            F.x = oc_ma<M>(damp, me.v.x, F.x, c.one);
            F.y = oc_ma<M>(damp, me.v.y, F.y, c.one);
            F.z = oc_ma<M>(damp, me.v.z, F.z, c.one);
            // the twelve springs in the order the reference's list touches the particle (V:286-320, oc_gather.cuh)
            spring<kAll, kMask>(F, me, ld(s0, ci - 1),  p_bc(K.rh1m), nS, kS, e0 && l1, e1 && l1, mL1, rg);                          // 1  (i-1, j)
            spring<kAll, kMask>(F, me, ld(s0, ci + 1),  p_bc(K.rh1i), nS, kS, e0 && r1, e1 && r1, mR1, rg);                          // 2  (i+1, j)
            spring<kAll, false>(F, me, kW ? win<(kP + 1) % 5>() : ld(sm1, ci), tV1m, nS, kS, e0 && u1_0, e1 && u1_1, 1.0f, rg);      // 3  (i, j-1)
            spring<kAll, false>(F, me, kW ? win<(kP + 3) % 5>() : ld(sp1, ci), tV1,  nS, kS, e0 && d1_0, e1 && d1_1, 1.0f, rg);      // 4  (i, j+1)
            spring<kAll, kMask>(F, me, ld(sm1, ci - 1), rUL, nSh, kSh, e0 && l1 && u1_0, e1 && l1 && u1_1, mL1, rg);                 // 5  (i-1, j-1)
            spring<kAll, kMask>(F, me, ld(sm1, ci + 1), rUR, nSh, kSh, e0 && r1 && u1_0, e1 && r1 && u1_1, mR1, rg);                 // 6  (i+1, j-1)
            spring<kAll, kMask>(F, me, ld(sp1, ci - 1), rLL, nSh, kSh, e0 && l1 && d1_0, e1 && l1 && d1_1, mL1, rg);                 // 7  (i-1, j+1)
            spring<kAll, kMask>(F, me, ld(sp1, ci + 1), rLR, nSh, kSh, e0 && r1 && d1_0, e1 && r1 && d1_1, mR1, rg);                 // 8  (i+1, j+1)
            {
                const OcPV2 nl = ld(s0, ci - 2), nr = ld(s0, ci + 2);
                spring<kAll, kMask>(F, me, nl, p_bc(K.rh2m), nB, kB, e0 && l2, e1 && l2, mL2, rg);                                   // 9  (i-2, j)
                spring<kAll, kMask>(F, me, nr, p_bc(K.rh2i), nB, kB, e0 && r2, e1 && r2, mR2, rg);                                   // 10 (i+2, j)
                if (!kAll) {                                                                                                    // 11 last bend spring of the row twice (V:313)
                    spring<false, false>(F, me, nr, p_bc(K.rh2i), nB, kB, e0 && gi == U - 3, e1 && gi == U - 3, 1.0f, rg);
                    spring<false, false>(F, me, nl, p_bc(K.rh2m), nB, kB, e0 && gi == U - 1, e1 && gi == U - 1, 1.0f, rg);
                } else if (kMask) {
                    spring<true, true>(F, me, nr, p_bc(K.rh2i), nB, kB, true, true, gi == U - 3 ? 1.0f : 0.0f, rg);
                    spring<true, true>(F, me, nl, p_bc(K.rh2m), nB, kB, true, true, gi == U - 1 ? 1.0f : 0.0f, rg);
                }
            }
            {
                const OcPV2 nu = kW ? win<kP>() : ld(sm2, ci);
                spring<kAll, false>(F, me, nu, tV2m, nB, kB, e0 && u2_0, e1 && u2_1, 1.0f, rg);                                      // 12 (i, j-2)
                // row + 2 was published at the end of the previous iteration: the only read that needs its barrier (with the
                // register window nothing does: rows +-2 are only read in the own column, and the wait moves down to the publish)
                if (!kW) { ctx.bar_wait(&s.bar, (unsigned)(it - it_first) & 1u); waited = true; }
                const OcPV2 nd = kW ? win<(kP + 4) % 5>() : ld(sp2, ci);
                spring<kAll, false>(F, me, nd, tV2,  nB, kB, e0 && d2_0, e1 && d2_1, 1.0f, rg);                                      // 13 (i, j+2)
                if (!kSteady) {                                                                                                 // 14 last bend spring of the column twice (V:319)
                    spring<false, false>(F, me, nd, tV2,  nB, kB, e0 && row_0 == V - 3, e1 && row_1 == V - 3, 1.0f, rg);
                    spring<false, false>(F, me, nu, tV2m, nB, kB, e0 && row_0 == V - 1, e1 && row_1 == V - 1, 1.0f, rg);
                }
            }
            if (__builtin_expect(M::kExact && rg.bad(), 0)) {
                // rare: an operand left the exact range of the branch-free sequences -> the whole sum again with the IEEE
                // intrinsics, per tile (cold, out of line; operands re-read from shared memory)
#ifdef __CUDA_ARCH__
                if (c.dbg & 4) {
                    atomicAdd(c.dbg_cnt, 1ull);
                    if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 1, 1ull);
                }
#endif
                unsigned on[2];
                for (int hh = 0; hh < 2; ++hh) {
                    const bool e = hh ? e1 : e0;
                    const bool u1 = hh ? u1_1 : u1_0, u2 = hh ? u2_1 : u2_0, d1 = hh ? d1_1 : d1_0, d2 = hh ? d2_1 : d2_0;
                    const int row = hh ? row_1 : row_0;
                    const bool exl1 = kAll ? mL1 != 0.0f || kInterior : l1, exl2 = kAll ? mL2 != 0.0f || kInterior : l2;
                    const bool exr1 = kAll ? mR1 != 0.0f || kInterior : r1, exr2 = kAll ? mR2 != 0.0f || kInterior : r2;
                    unsigned m = 0;
                    if (e) {
                        m |= (exl1 ? 1u : 0u) | (exr1 ? 2u : 0u) | (u1 ? 4u : 0u) | (d1 ? 8u : 0u);
                        m |= (exl1 && u1 ? 16u : 0u) | (exr1 && u1 ? 32u : 0u) | (exl1 && d1 ? 64u : 0u) | (exr1 && d1 ? 128u : 0u);
                        m |= (exl2 ? 256u : 0u) | (exr2 ? 512u : 0u) | (gi == U - 3 ? 1024u : 0u) | (gi == U - 1 ? 2048u : 0u);
                        m |= (u2 ? 4096u : 0u) | (d2 ? 8192u : 0u) | (row == V - 3 ? 16384u : 0u) | (row == V - 1 ? 32768u : 0u);
                    }
                    on[hh] = m;
                }
                OcStreamRow R0, R1;
                R0.rv1m = rv1m.x; R0.rv1 = rv1.x; R0.rv2m = rv2m.x; R0.rv2 = rv2.x; R0.dz2m = dz2m.x; R0.dz2 = dz2.x;
                R1.rv1m = rv1m.y; R1.rv1 = rv1.y; R1.rv2m = rv2m.y; R1.rv2 = rv2.y; R1.dz2m = dz2m.y; R1.dz2 = dz2.y;
                const f3 F0 = oc_stream_force_slow<M, Smem>(&c, sm, 0, it, ci, on[0], pin_0, K, R0);
                const f3 F1 = oc_stream_force_slow<M, Smem>(&c, sm, 1, it, ci, on[1], pin_1, K, R1);
                F.x = make_float2(F0.x, F1.x); F.y = make_float2(F0.y, F1.y); F.z = make_float2(F0.z, F1.z);
            }
            // ---- IntegrateVerlet (V:428-444) + EllipsoidCollision (V:509-533), both tiles ----------------
            OcPair3 n;
            n.x = oc_ma<M>(p_bc(c.dt2m), F.x, p_add(me.x.x, dme.x), c.one);
            n.y = oc_ma<M>(p_bc(c.dt2m), F.y, p_add(me.x.y, dme.y), c.one);
            n.z = oc_ma<M>(p_bc(c.dt2m), F.z, p_add(me.x.z, dme.z), c.one);
            if (n.y.x < 0.0f) n.y.x = 0.0f;
            if (n.y.y < 0.0f) n.y.y = 0.0f;
            bool hit_0 = false, hit_1 = false;
            const float2 ex = p_sub(n.x, p_bc(c.bs_c[0])), ey = p_sub(n.y, p_bc(c.bs_c[1])), ez = p_sub(n.z, p_bc(c.bs_c[2]));
            const float2 e2 = p_fma(ez, ez, p_fma(ey, ey, p_mul(ex, ex)));
            if ((e2.x <= c.bs_r2) | (e2.y <= c.bs_r2)) {
                OcPair3 p0;         // X_0 = inverse_ellipsoid * vec4(X,1) - center, rows x, y, z for (tile 0, tile 1)
                p0.x = p_sub(p_add(oc_ma<M>(p_bc(c.im[0][2]), n.z, oc_ma<M>(p_bc(c.im[0][1]), n.y, p_mul(p_bc(c.im[0][0]), n.x), c.one), c.one), p_bc(c.im[0][3])), p_bc(c.center[0]));
                p0.y = p_sub(p_add(oc_ma<M>(p_bc(c.im[1][2]), n.z, oc_ma<M>(p_bc(c.im[1][1]), n.y, p_mul(p_bc(c.im[1][0]), n.x), c.one), c.one), p_bc(c.im[1][3])), p_bc(c.center[1]));
                p0.z = p_sub(p_add(oc_ma<M>(p_bc(c.im[2][2]), n.z, oc_ma<M>(p_bc(c.im[2][1]), n.y, p_mul(p_bc(c.im[2][0]), n.x), c.one), c.one), p_bc(c.im[2][3])), p_bc(c.center[2]));
                const float2 sq = oc_ma<M>(p0.z, p0.z, oc_ma<M>(p0.y, p0.y, p_mul(p0.x, p0.x), c.one), c.one);
                hit_0 = sq.x < 1.0f; hit_1 = sq.y < 1.0f;                                           // V:513-514 (see oc_core.cuh)
#ifdef __CUDA_ARCH__
                if ((c.dbg & 4) && (hit_0 | hit_1)) { atomicAdd(c.dbg_cnt + 3, 1ull); if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 3, 1ull << 32); }
#endif
                if (__builtin_expect(hit_0 | hit_1, 0)) {
                    OcPair3 nn;
                    bool slow = false;
#ifdef __CUDA_ARCH__
                    if (M::kExact) {
                        OcRange rc; rc.init();
                        OcRangeStrict rn; rn.init();
                        const float2 distance = oc_sqrt2<M>(sq, rc);
                        const float2 sc = p_sub(p_bc(c.radius), distance);                                   // V:515
                        const float2 y0 = p_rcp(distance);
                        const float2 inv = p_fma(y0, p_fma(y0, p_neg(distance), p_bc(1.0f)), y0);
                        const float2 ax = p_mul(sc, p0.x), ay = p_mul(sc, p0.y), az = p_mul(sc, p0.z);
                        rn.add(ax.x); rn.add(ax.y); rn.add(ay.x); rn.add(ay.y); rn.add(az.x); rn.add(az.y);
                        float2 q0 = p_mul(ax, inv); const float2 dx = p_fma(inv, p_fma(q0, p_neg(distance), ax), q0);
                        q0 = p_mul(ay, inv);        const float2 dy = p_fma(inv, p_fma(q0, p_neg(distance), ay), q0);
                        q0 = p_mul(az, inv);        const float2 dz = p_fma(inv, p_fma(q0, p_neg(distance), az), q0);
                        nn.x = p_add(n.x, p_sump<M>(p_mul(dz, p_bc(c.tinv[0][2])), p_sump<M>(p_mul(dy, p_bc(c.tinv[0][1])), p_mul(dx, p_bc(c.tinv[0][0])), c.one), c.one));
                        nn.y = p_add(n.y, p_sump<M>(p_mul(dz, p_bc(c.tinv[1][2])), p_sump<M>(p_mul(dy, p_bc(c.tinv[1][1])), p_mul(dx, p_bc(c.tinv[1][0])), c.one), c.one));
                        nn.z = p_add(n.z, p_sump<M>(p_mul(dz, p_bc(c.tinv[2][2])), p_sump<M>(p_mul(dy, p_bc(c.tinv[2][1])), p_mul(dx, p_bc(c.tinv[2][0])), c.one), c.one));
                        slow = rc.bad() | rn.bad(OC_NUM_LO_BITS, OC_NUM_HI_BITS);
                    } else
#endif
                    if (!M::kExact) {
                        const float2 rinv = p_rsq(sq);
                        const float2 q = p_mul(p_sub(p_bc(c.radius), p_mul(sq, rinv)), rinv);
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
                // linked row bands: the first / last two rows of the band also go into the neighbour's halo (OcPeer2)
                const OcPeer2* pp = oc_opaque(peer);
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

        if (!waited) ctx.bar_wait(&s.bar, (unsigned)(it - it_first) & 1u);      // (iterations that compute nothing: keep the phases in step)
        // ---- publish rows prow (requested OC_STREAM_DEPTH iterations ago) and their per-row rest lengths ------------
        // (rows that do not exist or are not loaded become the benign far-away particle at rest)
        if (i < 3) {
            const float* t = i == 0 ? c.rv1 : (i == 1 ? c.rv2 : c.dz2);
            const int ra_ = prow_0 < 0 ? 0 : (prow_0 >= V ? V - 1 : prow_0), rb_ = prow_1 < 0 ? 0 : (prow_1 >= V ? V - 1 : prow_1);
            s.RC[i][sp3] = make_float2(OC_LDG(t + ra_), OC_LDG(t + rb_));
        }
        oc_cp_async_wait_n<OC_STREAM_DEPTH>();
        {
            const int zp = (it + 1) % (OC_STREAM_DEPTH + 1);          // == (it - OC_STREAM_DEPTH) mod (OC_STREAM_DEPTH + 1)
            const bool pub0 = kSteady ? (kInterior || ok) : (prow_0 >= in_lo0 && prow_0 < in_hi0 && ok);
            const bool pub1 = kSteady ? (kInterior || ok) : (prow_1 >= in_lo1 && prow_1 < in_hi1 && ok);
            const float4 la0 = pub0 ? s.stage[zp][0][i] : benign(ci, prow_0), lq0 = pub0 ? s.stage[zp][1][i] : benign(ci, prow_0);
            const float4 la1 = pub1 ? s.stage[zp][2][i] : benign(ci, prow_1), lq1 = pub1 ? s.stage[zp][3][i] : benign(ci, prow_1);
            const OcPV2 pv = publish(sp3, la0, lq0, la1, lq1);
            if (kW) {
                if (OC_STREAM_WIN == 2) win<kP>() = pv;                  // row r+3 takes the place of row r-2 (unrolled loop: renaming)
                else { W0 = W1; W1 = W2; W2 = W3; W3 = W4; W4 = pv; }    // one copy of the loop: the window moves
            }
        }
        ctx.bar_arrive(&s.bar);                   // phase it + 1 of the barrier: this thread has put its part of row prow into the ring
    }
};

template <class M, int WC, class Ctx>
OC_HD bool oc_stream_body(Ctx& ctx, const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                          float4* __restrict__ C, int ra, int rb, OcSeg2 seg, int x_halo, OcTwinMap map, const OcDep2& dep)
{
    OcStream<M, WC, Ctx> m(ctx, c);
    m.A = A; m.B = B; m.C = C;
    m.sm = reinterpret_cast<OcSmemS<WC, M::kExact>*>(ctx.smem());
    const int i = ctx.tid();
    const int U = c.U, V = c.V;
    const int W_out = WC - 2 * x_halo;
    const int cx0 = ctx.bx() * W_out - x_halo;
    const int gi = cx0 + i;
    int by[2], bz[2], r0[2], r1[2];
    oc_twin_tiles(seg, map, ctx.bx(), ctx.by(), ctx.bz(), ra, rb, by, bz, r0, r1);
    if (r0[0] >= r1[0] && r0[1] >= r1[1]) return ctx.wait_deps_twin(dep, c, seg, by, bz, r0, r1);
    m.i = i; m.ci = i + 2; m.gi = gi; m.U = U; m.V = V; m.bz0 = bz[0]; m.bz1 = bz[1];
    auto tile_rows = [&](int t0, int t1, int& lo, int& hi, int& in_lo, int& in_hi) {
        lo = t0; hi = t1;
        in_lo = lo - 2; if (in_lo < 0) in_lo = 0;
        in_hi = hi + 2; if (in_hi > V) in_hi = V;
        if (t0 >= t1) { in_lo = lo; in_hi = lo; }
    };
    tile_rows(r0[0], r1[0], m.lo0, m.hi0, m.in_lo0, m.in_hi0);
    tile_rows(r0[1], r1[1], m.lo1, m.hi1, m.in_lo1, m.in_hi1);
    const int n_rows = (r1[0] - r0[0]) > (r1[1] - r0[1]) ? (r1[0] - r0[0]) : (r1[1] - r0[1]);
    const bool empty0 = r0[0] >= r1[0], empty1 = r0[1] >= r1[1];
    const int f0 = empty0 ? r0[1] : r0[0], f1 = empty1 ? r0[0] : r0[1];
    // iteration `it` computes row row0 + it, publishes row row0 + it + 3 and requests row row0 + it + 3 + DEPTH: the first row a
    // tile needs is lo - 2
    m.row0 = f0 - 2 - OC_STREAM_AHEAD - OC_STREAM_DEPTH;
    m.dRow = f1 - f0;
    const int n_it = n_rows + 2 + OC_STREAM_AHEAD + OC_STREAM_DEPTH;
    m.peer = &dep.peer;
    m.ok = gi >= 0 && gi < U;
    m.st = m.ok && i >= x_halo && i < WC - x_halo;
    auto clampc = [&](int g) { return g < 0 ? 0 : (g >= U ? U - 1 : g); };
    m.K.rh1m = OC_LDG(c.rh1 + clampc(gi - 1)); m.K.rh1i = OC_LDG(c.rh1 + clampc(gi));
    m.K.rh2m = OC_LDG(c.rh2 + clampc(gi - 2)); m.K.rh2i = OC_LDG(c.rh2 + clampc(gi));
    m.K.dx2m = OC_LDG(c.dx2 + clampc(gi - 1)); m.K.dx2i = OC_LDG(c.dx2 + clampc(gi));
    if (!M::kExact) { m.K.rh1m *= c.nks_struct; m.K.rh1i *= c.nks_struct; m.K.rh2m *= c.nks_bend; m.K.rh2i *= c.nks_bend; }
    m.mL1 = (m.ok && gi - 1 >= 0) ? 1.0f : 0.0f; m.mL2 = (m.ok && gi - 2 >= 0) ? 1.0f : 0.0f;
    m.mR1 = (m.ok && gi + 1 < U) ? 1.0f : 0.0f;  m.mR2 = (m.ok && gi + 2 < U) ? 1.0f : 0.0f;
    m.ydt = oc_rcp_bf(c.dt);
    m.kdt_struct = c.kd_struct * c.inv_dt; m.kdt_shear = c.kd_shear * c.inv_dt; m.kdt_bend = c.kd_bend * c.inv_dt; m.damp_dt = c.damping * c.inv_dt;
    m.goff0 = (long long)bz[0] * c.cloth_stride - (long long)c.row_lo * U + gi;
    m.dOff = (long long)(bz[1] - bz[0]) * c.cloth_stride + (long long)m.dRow * U;

    // benign content for the whole ring (the pad columns are never written by a particle; the rows are, before they are read)
    {
        OcSmemS<WC, M::kExact>& s = *m.sm;
        const float2 z2 = make_float2(0.f, 0.f);
        constexpr int kComp = M::kExact ? 9 : 6;
        for (int e = i; e < OC_SRING * (WC + 4); e += WC) {
            const int slot = e / (WC + 4), col = e % (WC + 4);
            const float p = 1.0e3f + 8.0f * (float)col;
            for (int comp = 0; comp < kComp; ++comp)
                s.X[comp][slot][col] = comp == 0 ? make_float2(p, p) : (comp == 1 ? make_float2(1.0e3f, 1.0e3f) : (comp == 2 ? make_float2(1.0e3f + 8.0f * slot, 1.0e3f + 8.0f * slot) : z2));
        }
        for (int e = i; e < 3 * OC_SRING; e += WC) s.RC[e / OC_SRING][e % OC_SRING] = make_float2(1.0f, 1.0f);
        if (i == 0) ctx.bar_init(&s.bar, WC);
    }

    // steady range: interior rows of BOTH tiles
    int it_lo = 0, it_hi = n_it;
    auto steady_rows = [&](int bzh, int lo, int hi, int in_hi, int r00) {
        int st_lo = lo < 2 ? 2 : lo;
        int st_hi = hi < V - 3 ? hi : V - 3;
        if (st_hi > in_hi - OC_STREAM_AHEAD) st_hi = in_hi - OC_STREAM_AHEAD;
        if (dep.peer.c[0] && st_lo < dep.peer.ra + 2) st_lo = dep.peer.ra + 2;
        if (dep.peer.c[1] && st_hi > dep.peer.rb - 2) st_hi = dep.peer.rb - 2;
        if (st_lo < st_hi && !oc_rows_unpinned(c, bzh, st_lo, st_hi)) st_hi = st_lo;
        if (st_lo - r00 > it_lo) it_lo = st_lo - r00;
        if (st_hi - r00 < it_hi) it_hi = st_hi - r00;
    };
    steady_rows(bz[0], m.lo0, m.hi0, m.in_hi0, m.row0);
    steady_rows(bz[1], m.lo1, m.hi1, m.in_hi1, m.row0 + m.dRow);
    if (empty0 || empty1) it_lo = it_hi = n_it;
    if (it_lo < 0) it_lo = 0;
    if (it_hi > n_it) it_hi = n_it;
    if (it_hi <= it_lo) it_lo = it_hi = n_it;
    const bool interior = cx0 >= 2 && cx0 + WC + 2 <= U;

#ifdef __CUDA_ARCH__
    if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 1);
#endif
    ctx.sync();                                    // the ring's benign content is complete before anybody publishes into it
    if (!ctx.wait_deps_twin(dep, c, seg, by, bz, r0, r1)) return false;
    int it = 0;
    m.it_first = it;
    ctx.sync();
    ctx.bar_arrive(&m.sm->bar);                    // phase 0 of the barrier: the first iteration has nothing to wait for; iteration it waits for phase it - it_first
    for (int phase = 0; phase < 2; ++phase) {
        const int end = phase == 0 ? it_lo : n_it;
        for (; it < end; ++it) m.template iter<false, false>(it);
#ifdef __CUDA_ARCH__
        if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, phase == 0 ? 2 : 4);
#endif
        if (phase == 0) {
            if (OcStream<M, WC, Ctx>::kWin) {
                constexpr int kW0 = OcStream<M, WC, Ctx>::kWin ? 0 : -1, kW1 = kW0 < 0 ? -1 : 1, kW2 = kW0 < 0 ? -1 : 2, kW3 = kW0 < 0 ? -1 : 3, kW4 = kW0 < 0 ? -1 : 4;
                if (OC_STREAM_WIN == 2) {
                    // unrolled by five (the window rotates by renaming): the remainder first, from the ring
                    const int rem = (it_hi - it) % 5;
                    if (interior) { for (int e = it + rem; it < e; ++it) m.template iter<true, true>(it); }
                    else          { for (int e = it + rem; it < e; ++it) m.template iter<true, false>(it); }
                }
                if (it < it_hi) {
                    m.W0 = m.ld(oc_sslot(it + OC_SRING - 2), m.ci); m.W1 = m.ld(oc_sslot(it + OC_SRING - 1), m.ci); m.W2 = m.ld(oc_sslot(it), m.ci);
                    m.W3 = m.ld(oc_sslot(it + 1), m.ci);            m.W4 = m.ld(oc_sslot(it + 2), m.ci);
                }
                if (OC_STREAM_WIN == 2) {
                    if (interior) for (; it < it_hi; it += 5) { m.template iter<true, true, kW0>(it); m.template iter<true, true, kW1>(it + 1); m.template iter<true, true, kW2>(it + 2); m.template iter<true, true, kW3>(it + 3); m.template iter<true, true, kW4>(it + 4); }
                    else          for (; it < it_hi; it += 5) { m.template iter<true, false, kW0>(it); m.template iter<true, false, kW1>(it + 1); m.template iter<true, false, kW2>(it + 2); m.template iter<true, false, kW3>(it + 3); m.template iter<true, false, kW4>(it + 4); }
                } else {
                    if (interior) for (; it < it_hi; ++it) m.template iter<true, true, kW0>(it);
                    else          for (; it < it_hi; ++it) m.template iter<true, false, kW0>(it);
                }
            } else {
                if (interior) for (; it < it_hi; ++it) m.template iter<true, true>(it);
                else          for (; it < it_hi; ++it) m.template iter<true, false>(it);
            }
#ifdef __CUDA_ARCH__
            if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 3);
#endif
        }
    }
    return true;
}

#ifdef __CUDACC__
// MINB = resident CTAs per SM the register allocation is capped for
template <class M, int WC, int MINB>
__global__ void __launch_bounds__(WC, MINB)
oc_k_stream(const __grid_constant__ OcConst c, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C,
            int ra, int rb, OcSeg2 seg, int x_halo, OcTwinMap map, const __grid_constant__ OcDep2 dep)
{
    asm volatile("griddepcontrol.launch_dependents;");
    if ((c.dbg & 8) && threadIdx.x == 0) oc_timeline_mark(c, 0);
    OcDevCtxT ctx;
    ctx.x = blockIdx.x % seg.nstrips; ctx.y = blockIdx.x / seg.nstrips;
    if (!oc_stream_body<M, WC, OcDevCtxT>(ctx, c, A, B, C, ra, rb, seg, x_halo, map, dep)) return;
    int by[2], bz[2], r0[2], r1[2];
    oc_twin_tiles(seg, map, ctx.x, ctx.y, blockIdx.z, ra, rb, by, bz, r0, r1);
    ctx.publish(dep, seg, by, bz, r0, r1);
}
#endif
