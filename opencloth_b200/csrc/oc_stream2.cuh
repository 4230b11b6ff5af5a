// oc_stream2.cuh — kernel 7: oc_k_stream with TWO ADJACENT COLUMNS per thread (and two tiles, as there).
//
// oc_k_stream is bound by the shared-memory data pipe: a thread reads twelve neighbours of 48 bytes per particle pair and
// its neighbours in the warp read almost the same ones.  A thread that owns columns a = 2i and b = 2i+1
// shares them: of the 24 neighbour positions of its two particles (per tile) 6 are the other particle or a position the
// other one needs as well, loaded once and used from registers — 20 position loads per four particles instead of 26 per two.
// Everything else is oc_k_stream: the same ring, the same split-phase row barrier, the same per-particle arithmetic in
// the same (the reference's) order, so the results are bit-identical to it in both modes; WC / 2 threads per CTA.
#pragma once
#include "oc_stream.cuh"

template <int WC, bool kExact>
struct OcSmemS2 {
    float2 X[kExact ? 9 : 6][OC_SRING][WC + 4];      // as OcSmemS
    float2 RC[3][OC_SRING];
    float4 stage[8][WC / 2];                         // landing zone: A, B of tile 0, A, B of tile 1 for column a, then for column b
    unsigned long long bar;
};

template <class M, int WC, class Ctx>
struct OcStream2 {
    typedef OcSmemS2<WC, M::kExact> Smem;
    Ctx& ctx;
    const OcConst& c;
    const float4* __restrict__ A; const float4* __restrict__ B;
    float4* __restrict__ C;
    Smem* sm;
    int i, ci0, gi0, U, V;                   // thread, ring index and global index of column a (column b = + 1)
    int bz0, bz1;
    int lo0, hi0, in_lo0, in_hi0;
    int lo1, hi1, in_lo1, in_hi1;
    int row0, dRow, it_first;
    bool ok[2], st[2];
    float rh1s[3], rh2s[4], dxs[3];          // rh1[ga-1 .. gb], rh2[ga-2 .. gb], dx2[ga-1 .. gb]  (fast: rh1, rh2 times -Ks)
    float mL1[2], mL2[2], mR1[2], mR2[2];
    float ydt, kdt_struct, kdt_shear, kdt_bend, damp_dt;
    long long goff0, dOff;
    const OcPeer2* peer;

    OC_HD OcStream2(Ctx& ctx_, const OcConst& c_) : ctx(ctx_), c(c_) {}

    OC_HD OcPV2 ld(int slot, int col) const
    {
        OcPV2 r;
        r.x.x = sm->X[0][slot][col]; r.x.y = sm->X[1][slot][col]; r.x.z = sm->X[2][slot][col];
        r.v.x = sm->X[3][slot][col]; r.v.y = sm->X[4][slot][col]; r.v.z = sm->X[5][slot][col];
        return r;
    }
    static OC_HD float4 benign(int ci_, int lrow) { return make_float4(1.0e3f + 8.0f * (float)ci_, 1.0e3f, 1.0e3f + 8.0f * (float)(lrow & 63), oc_u2f(OC_W_PLAIN)); }

    // one loaded row of ring column `col` (both tiles) into slot sl (see OcStream::publish)
    OC_HD void publish(int sl, int col, const float4 la0, const float4 lq0, const float4 la1, const float4 lq1)
    {
        Smem& s = *sm;
        OcPair3 d;
        d.x = make_float2(M::sub(la0.x, lq0.x), M::sub(la1.x, lq1.x));
        d.y = make_float2(M::sub(la0.y, lq0.y), M::sub(la1.y, lq1.y));
        d.z = make_float2(M::sub(la0.z, lq0.z), M::sub(la1.z, lq1.z));
        if (oc_hit(la0.w)) { d.x.x = 0.0f; d.y.x = 0.0f; d.z.x = 0.0f; }       // X_last == X (V:530)
        if (oc_hit(la1.w)) { d.x.y = 0.0f; d.y.y = 0.0f; d.z.y = 0.0f; }
        s.X[0][sl][col] = make_float2(la0.x, la1.x);
        s.X[1][sl][col] = make_float2(la0.y, la1.y);
        s.X[2][sl][col] = make_float2(la0.z, la1.z);
        if (!M::kExact) { s.X[3][sl][col] = d.x; s.X[4][sl][col] = d.y; s.X[5][sl][col] = d.z; return; }
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
        constexpr int kE = M::kExact ? 1 : 0;
        s.X[3][sl][col] = v.x; s.X[4][sl][col] = v.y; s.X[5][sl][col] = v.z;
        s.X[6 * kE][sl][col] = d.x; s.X[7 * kE][sl][col] = d.y; s.X[8 * kE][sl][col] = d.z;
    }

    // as OcStream::spring
    template <bool kAll, bool kMask>
    OC_HD void spring(OcPair3& F, const OcPV2& me, const OcPV2& n, float2 rest, float2 nks, float2 kd, bool p0, bool p1, float mask, OcRange& rg)
    {
        if (M::kExact) {
            OcPair3 f = oc_spring_twin<M>(me.x, me.v, n.x, n.v, rest, nks, kd, c.one, rg);
            if (kMask) { const float2 m = p_bc(mask); f.x = p_mul(f.x, m); f.y = p_mul(f.y, m); f.z = p_mul(f.z, m); }
            oc_acc2<M, kAll>(F, f, p0, p1, false, c.one);
        } else {
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

    // IntegrateVerlet + EllipsoidCollision + store of column cc (the code of OcStream::iter)
    template <bool kSteady>
    OC_HD void finish(int cc, int g, bool stc, const OcPV2& me, const OcPair3& dme, const OcPair3& F, bool doG0, bool doG1, int row_0, int row_1)
    {
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
        const long long o = goff0 + (long long)row_0 * U + cc;
        const float4 out_0 = make_float4(n.x.x, n.y.x, n.z.x, oc_u2f(hit_0 ? OC_W_HIT : OC_W_PLAIN));
        const float4 out_1 = make_float4(n.x.y, n.y.y, n.z.y, oc_u2f(hit_1 ? OC_W_HIT : OC_W_PLAIN));
        if (stc && doG0) C[o] = out_0;
        if (stc && doG1) C[o + dOff] = out_1;
        if (!kSteady) {
            // linked row bands: the first / last two rows of the band also go into the neighbour's halo (OcPeer2)
            const OcPeer2* pp = oc_opaque(peer);
            if (stc && (pp->c[0] || pp->c[1])) {
                if (doG0) {
                    float4* pc = nullptr;
                    if (pp->c[0] && row_0 < pp->ra + 2) pc = pp->c[0];
                    if (pp->c[1] && row_0 >= pp->rb - 2) pc = pp->c[1];
                    if (pc) pc[(long long)row_0 * U + g] = out_0;
                }
                if (doG1) {
                    float4* pc = nullptr;
                    if (pp->c[0] && row_1 < pp->ra + 2) pc = pp->c[0];
                    if (pp->c[1] && row_1 >= pp->rb - 2) pc = pp->c[1];
                    if (pc) pc[(long long)row_1 * U + g] = out_1;
                }
            }
        }
    }

    template <bool kSteady, bool kInterior>
    OC_HD void iter(int it)
    {
        Smem& s = *sm;
        const int row_0 = row0 + it, row_1 = row_0 + dRow;
        const int prow_0 = row_0 + OC_STREAM_AHEAD, prow_1 = row_1 + OC_STREAM_AHEAD;                 // rows requested and published by this iteration
        const bool inL0 = prow_0 >= in_lo0 && prow_0 < in_hi0, inL1 = prow_1 >= in_lo1 && prow_1 < in_hi1;
        {
            const long long o = goff0 + (long long)prow_0 * U;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const bool okc = kInterior || ok[cc];
                if (inL0 && okc) { oc_cp_async16(&s.stage[4 * cc + 0][i], A + o + cc); oc_cp_async16(&s.stage[4 * cc + 1][i], B + o + cc); }
                if (inL1 && okc) { oc_cp_async16(&s.stage[4 * cc + 2][i], A + o + dOff + cc); oc_cp_async16(&s.stage[4 * cc + 3][i], B + o + dOff + cc); }
            }
            oc_cp_async_commit();
        }
        const int s0 = oc_sslot(it);
        auto wrap = [](int x) { return x >= OC_SRING ? x - OC_SRING : x; };
        const int sp1 = wrap(s0 + 1), sp2 = wrap(s0 + 2), sp3 = wrap(s0 + OC_STREAM_AHEAD), sm2 = wrap(s0 + OC_SRING - 2), sm1 = wrap(s0 + OC_SRING - 1);
        const int ca = ci0, cb = ci0 + 1, ga = gi0, gb = gi0 + 1;

        const bool doG0 = kSteady || (row_0 >= lo0 && row_0 < hi0);
        const bool doG1 = kSteady || (row_1 >= lo1 && row_1 < hi1);
        bool waited = false;
        if (doG0 | doG1) {
            constexpr bool kAll = kSteady;
            constexpr bool kMask = kSteady && !kInterior;
            const OcPV2 meA = ld(s0, ca), meB = ld(s0, cb);
            OcPair3 dmeA, dmeB;
            if (M::kExact) {
                constexpr int kE = M::kExact ? 1 : 0;
                dmeA.x = s.X[6 * kE][s0][ca]; dmeA.y = s.X[7 * kE][s0][ca]; dmeA.z = s.X[8 * kE][s0][ca];
                dmeB.x = s.X[6 * kE][s0][cb]; dmeB.y = s.X[7 * kE][s0][cb]; dmeB.z = s.X[8 * kE][s0][cb];
            } else { dmeA = meA.v; dmeB = meB.v; }
            const float2 rv1m = s.RC[0][sm1], rv1 = s.RC[0][s0], rv2m = s.RC[1][sm2], rv2 = s.RC[1][s0], dz2m = s.RC[2][sm1], dz2 = s.RC[2][s0];
            const bool pinA0 = !kSteady && oc_pinned(c, bz0, ga, row_0), pinA1 = !kSteady && oc_pinned(c, bz1, ga, row_1);
            const bool pinB0 = !kSteady && oc_pinned(c, bz0, gb, row_0), pinB1 = !kSteady && oc_pinned(c, bz1, gb, row_1);
            const bool eA0 = !pinA0 && doG0, eA1 = !pinA1 && doG1, eB0 = !pinB0 && doG0, eB1 = !pinB1 && doG1;
            // existence of the horizontal neighbours of a and of b (steady loop: the 0/1 masks do it)
            const bool al1 = kAll || ga - 1 >= 0, al2 = kAll || ga - 2 >= 0, ar1 = kAll || ga + 1 < U, ar2 = kAll || ga + 2 < U;
            const bool bl1 = kAll || gb - 1 >= 0, bl2 = kAll || gb - 2 >= 0, br1 = kAll || gb + 1 < U, br2 = kAll || gb + 2 < U;
            const bool u1_0 = kSteady || row_0 - 1 >= 0, u2_0 = kSteady || row_0 - 2 >= 0, d1_0 = kSteady || row_0 + 1 < V, d2_0 = kSteady || row_0 + 2 < V;
            const bool u1_1 = kSteady || row_1 - 1 >= 0, u2_1 = kSteady || row_1 - 2 >= 0, d1_1 = kSteady || row_1 + 1 < V, d2_1 = kSteady || row_1 + 2 < V;
            const float2 nS = p_bc(c.nks_struct), nB = p_bc(c.nks_bend), nSh = p_bc(c.nks_shear);
            const float2 kS = p_bc(M::kExact ? c.kd_struct : kdt_struct), kB = p_bc(M::kExact ? c.kd_bend : kdt_bend), kSh = p_bc(M::kExact ? c.kd_shear : kdt_shear);
            OcRange rgA, rgB; rgA.init(); rgB.init();
            // shear rest lengths of the cells of columns ga-1, ga, gb in the rows above and below
            float2 rU[3], rL[3];
            {
                OcRange rs; rs.init();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    rU[k] = oc_sqrt2<M>(p_add(p_bc(dxs[k]), dz2m), rs); rL[k] = oc_sqrt2<M>(p_add(p_bc(dxs[k]), dz2), rs);
                    if (!M::kExact) { rU[k] = p_mul(rU[k], nSh); rL[k] = p_mul(rL[k], nSh); }
                }
                rgA.sq = rs.sq; rgB.sq = rs.sq;
            }
            float2 tV1m = rv1m, tV1 = rv1, tV2m = rv2m, tV2 = rv2;
            if (!M::kExact) { tV1m = p_mul(tV1m, nS); tV1 = p_mul(tV1, nS); tV2m = p_mul(tV2m, nB); tV2 = p_mul(tV2, nB); }
            // F = 0 + gravity*mass (unless pinned) + DEFAULT_DAMPING*V     V:451-459
            OcPair3 FA, FB;
            const float2 damp = p_bc(M::kExact ? c.damping : damp_dt);
            FA.x = oc_ma<M>(damp, meA.v.x, make_float2(pinA0 ? 0.0f : c.f0[0], pinA1 ? 0.0f : c.f0[0]), c.one);
            FA.y = oc_ma<M>(damp, meA.v.y, make_float2(pinA0 ? 0.0f : c.f0[1], pinA1 ? 0.0f : c.f0[1]), c.one);
            FA.z = oc_ma<M>(damp, meA.v.z, make_float2(pinA0 ? 0.0f : c.f0[2], pinA1 ? 0.0f : c.f0[2]), c.one);
            FB.x = oc_ma<M>(damp, meB.v.x, make_float2(pinB0 ? 0.0f : c.f0[0], pinB1 ? 0.0f : c.f0[0]), c.one);
            FB.y = oc_ma<M>(damp, meB.v.y, make_float2(pinB0 ? 0.0f : c.f0[1], pinB1 ? 0.0f : c.f0[1]), c.one);
            FB.z = oc_ma<M>(damp, meB.v.z, make_float2(pinB0 ? 0.0f : c.f0[2], pinB1 ? 0.0f : c.f0[2]), c.one);
            // the twelve springs of a and of b, each particle's in the reference's order; a position both need is loaded once
            spring<kAll, kMask>(FA, meA, ld(s0, ca - 1), p_bc(rh1s[0]), nS, kS, eA0 && al1, eA1 && al1, mL1[0], rgA);               // 1  a: (i-1, j)
            spring<kAll, kMask>(FB, meB, meA,            p_bc(rh1s[1]), nS, kS, eB0 && bl1, eB1 && bl1, mL1[1], rgB);               //    b: a
            spring<kAll, kMask>(FA, meA, meB,            p_bc(rh1s[1]), nS, kS, eA0 && ar1, eA1 && ar1, mR1[0], rgA);               // 2  a: b
            spring<kAll, kMask>(FB, meB, ld(s0, cb + 1), p_bc(rh1s[2]), nS, kS, eB0 && br1, eB1 && br1, mR1[1], rgB);               //    b: (i+1, j)
            {
                const OcPV2 uA = ld(sm1, ca), uB = ld(sm1, cb), dA = ld(sp1, ca), dB = ld(sp1, cb);
                spring<kAll, false>(FA, meA, uA, tV1m, nS, kS, eA0 && u1_0, eA1 && u1_1, 1.0f, rgA);                                // 3  (i, j-1)
                spring<kAll, false>(FB, meB, uB, tV1m, nS, kS, eB0 && u1_0, eB1 && u1_1, 1.0f, rgB);
                spring<kAll, false>(FA, meA, dA, tV1,  nS, kS, eA0 && d1_0, eA1 && d1_1, 1.0f, rgA);                                // 4  (i, j+1)
                spring<kAll, false>(FB, meB, dB, tV1,  nS, kS, eB0 && d1_0, eB1 && d1_1, 1.0f, rgB);
                spring<kAll, kMask>(FA, meA, ld(sm1, ca - 1), rU[0], nSh, kSh, eA0 && al1 && u1_0, eA1 && al1 && u1_1, mL1[0], rgA); // 5  (i-1, j-1)
                spring<kAll, kMask>(FB, meB, uA,              rU[1], nSh, kSh, eB0 && bl1 && u1_0, eB1 && bl1 && u1_1, mL1[1], rgB);
                spring<kAll, kMask>(FA, meA, uB,              rU[1], nSh, kSh, eA0 && ar1 && u1_0, eA1 && ar1 && u1_1, mR1[0], rgA); // 6  (i+1, j-1)
                spring<kAll, kMask>(FB, meB, ld(sm1, cb + 1), rU[2], nSh, kSh, eB0 && br1 && u1_0, eB1 && br1 && u1_1, mR1[1], rgB);
                spring<kAll, kMask>(FA, meA, ld(sp1, ca - 1), rL[0], nSh, kSh, eA0 && al1 && d1_0, eA1 && al1 && d1_1, mL1[0], rgA); // 7  (i-1, j+1)
                spring<kAll, kMask>(FB, meB, dA,              rL[1], nSh, kSh, eB0 && bl1 && d1_0, eB1 && bl1 && d1_1, mL1[1], rgB);
                spring<kAll, kMask>(FA, meA, dB,              rL[1], nSh, kSh, eA0 && ar1 && d1_0, eA1 && ar1 && d1_1, mR1[0], rgA); // 8  (i+1, j+1)
                spring<kAll, kMask>(FB, meB, ld(sp1, cb + 1), rL[2], nSh, kSh, eB0 && br1 && d1_0, eB1 && br1 && d1_1, mR1[1], rgB);
            }
            {
                const OcPV2 l2 = ld(s0, ca - 2), l1 = ld(s0, ca - 1), r1 = ld(s0, cb + 1), r2 = ld(s0, cb + 2);
                spring<kAll, kMask>(FA, meA, l2, p_bc(rh2s[0]), nB, kB, eA0 && al2, eA1 && al2, mL2[0], rgA);                       // 9  (i-2, j)
                spring<kAll, kMask>(FB, meB, l1, p_bc(rh2s[1]), nB, kB, eB0 && bl2, eB1 && bl2, mL2[1], rgB);
                spring<kAll, kMask>(FA, meA, r1, p_bc(rh2s[2]), nB, kB, eA0 && ar2, eA1 && ar2, mR2[0], rgA);                       // 10 (i+2, j)
                spring<kAll, kMask>(FB, meB, r2, p_bc(rh2s[3]), nB, kB, eB0 && br2, eB1 && br2, mR2[1], rgB);
                if (!kAll) {                                                                                                   // 11 last bend spring of the row twice (V:313)
                    spring<false, false>(FA, meA, r1, p_bc(rh2s[2]), nB, kB, eA0 && ga == U - 3, eA1 && ga == U - 3, 1.0f, rgA);
                    spring<false, false>(FA, meA, l2, p_bc(rh2s[0]), nB, kB, eA0 && ga == U - 1, eA1 && ga == U - 1, 1.0f, rgA);
                    spring<false, false>(FB, meB, r2, p_bc(rh2s[3]), nB, kB, eB0 && gb == U - 3, eB1 && gb == U - 3, 1.0f, rgB);
                    spring<false, false>(FB, meB, l1, p_bc(rh2s[1]), nB, kB, eB0 && gb == U - 1, eB1 && gb == U - 1, 1.0f, rgB);
                } else if (kMask) {
                    spring<true, true>(FA, meA, r1, p_bc(rh2s[2]), nB, kB, true, true, ga == U - 3 ? 1.0f : 0.0f, rgA);
                    spring<true, true>(FA, meA, l2, p_bc(rh2s[0]), nB, kB, true, true, ga == U - 1 ? 1.0f : 0.0f, rgA);
                    spring<true, true>(FB, meB, r2, p_bc(rh2s[3]), nB, kB, true, true, gb == U - 3 ? 1.0f : 0.0f, rgB);
                    spring<true, true>(FB, meB, l1, p_bc(rh2s[1]), nB, kB, true, true, gb == U - 1 ? 1.0f : 0.0f, rgB);
                }
            }
            {
                const OcPV2 uuA = ld(sm2, ca), uuB = ld(sm2, cb);
                spring<kAll, false>(FA, meA, uuA, tV2m, nB, kB, eA0 && u2_0, eA1 && u2_1, 1.0f, rgA);                               // 12 (i, j-2)
                spring<kAll, false>(FB, meB, uuB, tV2m, nB, kB, eB0 && u2_0, eB1 && u2_1, 1.0f, rgB);
                // row + 2 was published at the end of the previous iteration: the only read that needs its barrier
                ctx.bar_wait(&s.bar, (unsigned)(it - it_first) & 1u); waited = true;
                const OcPV2 ddA = ld(sp2, ca), ddB = ld(sp2, cb);
                spring<kAll, false>(FA, meA, ddA, tV2, nB, kB, eA0 && d2_0, eA1 && d2_1, 1.0f, rgA);                                // 13 (i, j+2)
                spring<kAll, false>(FB, meB, ddB, tV2, nB, kB, eB0 && d2_0, eB1 && d2_1, 1.0f, rgB);
                if (!kSteady) {                                                                                                // 14 last bend spring of the column twice (V:319)
                    spring<false, false>(FA, meA, ddA, tV2,  nB, kB, eA0 && row_0 == V - 3, eA1 && row_1 == V - 3, 1.0f, rgA);
                    spring<false, false>(FA, meA, uuA, tV2m, nB, kB, eA0 && row_0 == V - 1, eA1 && row_1 == V - 1, 1.0f, rgA);
                    spring<false, false>(FB, meB, ddB, tV2,  nB, kB, eB0 && row_0 == V - 3, eB1 && row_1 == V - 3, 1.0f, rgB);
                    spring<false, false>(FB, meB, uuB, tV2m, nB, kB, eB0 && row_0 == V - 1, eB1 && row_1 == V - 1, 1.0f, rgB);
                }
            }
            if (__builtin_expect(M::kExact && (rgA.bad() | rgB.bad()), 0)) {
                // rare: an operand left the exact range of the branch-free sequences -> both particles' sums again with the
                // IEEE intrinsics, per tile (cold, out of line; operands re-read from shared memory)
#ifdef __CUDA_ARCH__
                if (c.dbg & 4) {
                    atomicAdd(c.dbg_cnt, 1ull);
                    if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(c.dbg_cnt + 1, 1ull);
                }
#endif
                OcStreamRow R0, R1;
                R0.rv1m = rv1m.x; R0.rv1 = rv1.x; R0.rv2m = rv2m.x; R0.rv2 = rv2.x; R0.dz2m = dz2m.x; R0.dz2 = dz2.x;
                R1.rv1m = rv1m.y; R1.rv1 = rv1.y; R1.rv2m = rv2m.y; R1.rv2 = rv2.y; R1.dz2m = dz2m.y; R1.dz2 = dz2.y;
                for (int cc = 0; cc < 2; ++cc) {
                    const int g = gi0 + cc;
                    OcStreamCol K;
                    K.rh1m = rh1s[cc]; K.rh1i = rh1s[cc + 1]; K.rh2m = rh2s[cc]; K.rh2i = rh2s[cc + 2]; K.dx2m = dxs[cc]; K.dx2i = dxs[cc + 1];
                    unsigned on[2];
                    for (int hh = 0; hh < 2; ++hh) {
                        const bool e = cc ? (hh ? eB1 : eB0) : (hh ? eA1 : eA0);
                        const bool u1 = hh ? u1_1 : u1_0, u2 = hh ? u2_1 : u2_0, d1 = hh ? d1_1 : d1_0, d2 = hh ? d2_1 : d2_0;
                        const int row = hh ? row_1 : row_0;
                        const bool exl1 = g - 1 >= 0 && g < U, exl2 = g - 2 >= 0 && g < U, exr1 = g >= 0 && g + 1 < U, exr2 = g >= 0 && g + 2 < U;
                        unsigned mk = 0;
                        if (e) {
                            mk |= (exl1 ? 1u : 0u) | (exr1 ? 2u : 0u) | (u1 ? 4u : 0u) | (d1 ? 8u : 0u);
                            mk |= (exl1 && u1 ? 16u : 0u) | (exr1 && u1 ? 32u : 0u) | (exl1 && d1 ? 64u : 0u) | (exr1 && d1 ? 128u : 0u);
                            mk |= (exl2 ? 256u : 0u) | (exr2 ? 512u : 0u) | (g == U - 3 ? 1024u : 0u) | (g == U - 1 ? 2048u : 0u);
                            mk |= (u2 ? 4096u : 0u) | (d2 ? 8192u : 0u) | (row == V - 3 ? 16384u : 0u) | (row == V - 1 ? 32768u : 0u);
                        }
                        on[hh] = mk;
                    }
                    const bool p0 = cc ? pinB0 : pinA0, p1 = cc ? pinB1 : pinA1;
                    const f3 F0 = oc_stream_force_slow<M, Smem>(&c, sm, 0, it, ci0 + cc, on[0], p0, K, R0);
                    const f3 F1 = oc_stream_force_slow<M, Smem>(&c, sm, 1, it, ci0 + cc, on[1], p1, K, R1);
                    OcPair3& F = cc ? FB : FA;
                    F.x = make_float2(F0.x, F1.x); F.y = make_float2(F0.y, F1.y); F.z = make_float2(F0.z, F1.z);
                }
            }
            finish<kSteady>(0, ga, st[0], meA, dmeA, FA, doG0, doG1, row_0, row_1);
            finish<kSteady>(1, gb, st[1], meB, dmeB, FB, doG0, doG1, row_0, row_1);
        }
        if (!waited) ctx.bar_wait(&s.bar, (unsigned)(it - it_first) & 1u);
        // ---- publish rows prow of both columns and their per-row rest lengths ------------------------------------------
        if (i < 3) {
            const float* t = i == 0 ? c.rv1 : (i == 1 ? c.rv2 : c.dz2);
            const int ra_ = prow_0 < 0 ? 0 : (prow_0 >= V ? V - 1 : prow_0), rb_ = prow_1 < 0 ? 0 : (prow_1 >= V ? V - 1 : prow_1);
            s.RC[i][sp3] = make_float2(OC_LDG(t + ra_), OC_LDG(t + rb_));
        }
        oc_cp_async_wait();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const bool pub0 = kSteady ? (kInterior || ok[cc]) : (inL0 && ok[cc]);
            const bool pub1 = kSteady ? (kInterior || ok[cc]) : (inL1 && ok[cc]);
            const float4 la0 = pub0 ? s.stage[4 * cc + 0][i] : benign(ci0 + cc, prow_0), lq0 = pub0 ? s.stage[4 * cc + 1][i] : benign(ci0 + cc, prow_0);
            const float4 la1 = pub1 ? s.stage[4 * cc + 2][i] : benign(ci0 + cc, prow_1), lq1 = pub1 ? s.stage[4 * cc + 3][i] : benign(ci0 + cc, prow_1);
            publish(sp3, ci0 + cc, la0, lq0, la1, lq1);
        }
        ctx.bar_arrive(&s.bar);
    }
};

template <class M, int WC, class Ctx>
OC_HD bool oc_stream2_body(Ctx& ctx, const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                           float4* __restrict__ C, int ra, int rb, OcSeg2 seg, int x_halo, OcTwinMap map, const OcDep2& dep)
{
    OcStream2<M, WC, Ctx> m(ctx, c);
    m.A = A; m.B = B; m.C = C;
    m.sm = reinterpret_cast<OcSmemS2<WC, M::kExact>*>(ctx.smem());
    constexpr int T = WC / 2;
    const int i = ctx.tid();
    const int U = c.U, V = c.V;
    const int W_out = WC - 2 * x_halo;
    const int cx0 = ctx.bx() * W_out - x_halo;
    const int gi0 = cx0 + 2 * i;
    int by[2], bz[2], r0[2], r1[2];
    oc_twin_tiles(seg, map, ctx.bx(), ctx.by(), ctx.bz(), ra, rb, by, bz, r0, r1);
    if (r0[0] >= r1[0] && r0[1] >= r1[1]) return ctx.wait_deps_twin(dep, c, seg, by, bz, r0, r1);
    m.i = i; m.ci0 = 2 * i + 2; m.gi0 = gi0; m.U = U; m.V = V; m.bz0 = bz[0]; m.bz1 = bz[1];
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
    m.row0 = f0 - 2 - OC_STREAM_AHEAD;
    m.dRow = f1 - f0;
    const int n_it = n_rows + 2 + OC_STREAM_AHEAD;
    m.peer = &dep.peer;
    auto clampc = [&](int g) { return g < 0 ? 0 : (g >= U ? U - 1 : g); };
    for (int cc = 0; cc < 2; ++cc) {
        const int g = gi0 + cc, w = 2 * i + cc;
        m.ok[cc] = g >= 0 && g < U;
        m.st[cc] = m.ok[cc] && w >= x_halo && w < WC - x_halo;
        m.mL1[cc] = (m.ok[cc] && g - 1 >= 0) ? 1.0f : 0.0f; m.mL2[cc] = (m.ok[cc] && g - 2 >= 0) ? 1.0f : 0.0f;
        m.mR1[cc] = (m.ok[cc] && g + 1 < U) ? 1.0f : 0.0f;  m.mR2[cc] = (m.ok[cc] && g + 2 < U) ? 1.0f : 0.0f;
    }
    for (int k = 0; k < 3; ++k) { m.rh1s[k] = OC_LDG(c.rh1 + clampc(gi0 - 1 + k)); m.dxs[k] = OC_LDG(c.dx2 + clampc(gi0 - 1 + k)); }
    for (int k = 0; k < 4; ++k) m.rh2s[k] = OC_LDG(c.rh2 + clampc(gi0 - 2 + k));
    if (!M::kExact) { for (int k = 0; k < 3; ++k) m.rh1s[k] *= c.nks_struct; for (int k = 0; k < 4; ++k) m.rh2s[k] *= c.nks_bend; }
    m.ydt = oc_rcp_bf(c.dt);
    m.kdt_struct = c.kd_struct * c.inv_dt; m.kdt_shear = c.kd_shear * c.inv_dt; m.kdt_bend = c.kd_bend * c.inv_dt; m.damp_dt = c.damping * c.inv_dt;
    m.goff0 = (long long)bz[0] * c.cloth_stride - (long long)c.row_lo * U + gi0;
    m.dOff = (long long)(bz[1] - bz[0]) * c.cloth_stride + (long long)m.dRow * U;

    {
        OcSmemS2<WC, M::kExact>& s = *m.sm;
        const float2 z2 = make_float2(0.f, 0.f);
        constexpr int kComp = M::kExact ? 9 : 6;
        for (int e = i; e < OC_SRING * (WC + 4); e += T) {
            const int slot = e / (WC + 4), col = e % (WC + 4);
            const float p = 1.0e3f + 8.0f * (float)col;
            for (int comp = 0; comp < kComp; ++comp)
                s.X[comp][slot][col] = comp == 0 ? make_float2(p, p) : (comp == 1 ? make_float2(1.0e3f, 1.0e3f) : (comp == 2 ? make_float2(1.0e3f + 8.0f * slot, 1.0e3f + 8.0f * slot) : z2));
        }
        for (int e = i; e < 3 * OC_SRING; e += T) s.RC[e / OC_SRING][e % OC_SRING] = make_float2(1.0f, 1.0f);
        if (i == 0) ctx.bar_init(&s.bar, T);
    }

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
    ctx.sync();
    if (!ctx.wait_deps_twin(dep, c, seg, by, bz, r0, r1)) return false;
    int it = 0;
    m.it_first = it;
    ctx.sync();
    ctx.bar_arrive(&m.sm->bar);
    for (int phase = 0; phase < 2; ++phase) {
        const int end = phase == 0 ? it_lo : n_it;
        for (; it < end; ++it) m.template iter<false, false>(it);
#ifdef __CUDA_ARCH__
        if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, phase == 0 ? 2 : 4);
#endif
        if (phase == 0) {
            if (interior) for (; it < it_hi; ++it) m.template iter<true, true>(it);
            else          for (; it < it_hi; ++it) m.template iter<true, false>(it);
#ifdef __CUDA_ARCH__
            if ((c.dbg & 8) && i == 0) oc_timeline_mark(c, 3);
#endif
        }
    }
    return true;
}

#ifdef __CUDACC__
template <class M, int WC, int MINB>
__global__ void __launch_bounds__(WC / 2, MINB)
oc_k_stream2(const __grid_constant__ OcConst c, const float4* __restrict__ A, const float4* __restrict__ B, float4* __restrict__ C,
             int ra, int rb, OcSeg2 seg, int x_halo, OcTwinMap map, const __grid_constant__ OcDep2 dep)
{
    asm volatile("griddepcontrol.launch_dependents;");
    if ((c.dbg & 8) && threadIdx.x == 0) oc_timeline_mark(c, 0);
    OcDevCtxT ctx;
    ctx.x = blockIdx.x % seg.nstrips; ctx.y = blockIdx.x / seg.nstrips;
    if (!oc_stream2_body<M, WC, OcDevCtxT>(ctx, c, A, B, C, ra, rb, seg, x_halo, map, dep)) return;
    int by[2], bz[2], r0[2], r1[2];
    oc_twin_tiles(seg, map, ctx.x, ctx.y, blockIdx.z, ra, rb, by, bz, r0, r1);
    ctx.publish(dep, seg, by, bz, r0, r1);
}
#endif
