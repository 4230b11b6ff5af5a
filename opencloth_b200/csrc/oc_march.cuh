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
#define OC_MARCH_LAG 4          // rows between consecutive stages
#define OC_RING 4               // ring depth (power of two)

template <int TW>
struct OcStageSmem {
    float4 P[OC_RING][TW + 4];      // x, y, z, vx        (2 pad columns either side)
    float2 Q[OC_RING][TW + 4];      // vy, vz
    float  D[OC_RING][3][TW + 4];   // X - X_last, read by the owning column only
    float4 FH[2][2][TW + 4];        // [row&1][0] = -f(+1,0), [row&1][1] = -f(+2,0)   forces on the partner
    float4 FD[OC_RING][2][TW + 4];  // [row&3][0] = -f(+1,+1), [row&3][1] = -f(-1,+1)
};

struct OcPV { f3 x, v; };

#ifdef __CUDA_ARCH__
#define OC_LDG(p) __ldg(p)
// Optimisation fence on a float4 held in registers: whatever consumes it is scheduled after this
// point.  Used to keep the consumers of a global load behind the barrier, so that the load's
// latency is covered by the spring phase instead of stalling it.
#define OC_KEEP4(v) asm volatile("" : "+f"((v).x), "+f"((v).y), "+f"((v).z), "+f"((v).w))
#else
#define OC_LDG(p) (*(p))
#define OC_KEEP4(v) ((void)0)
#endif

template <int TW>
OC_HD OcPV oc_ld_pv(const OcStageSmem<TW>& sm, int row, int col /* padded index */)
{
    float4 p = sm.P[row & (OC_RING - 1)][col];
    float2 q = sm.Q[row & (OC_RING - 1)][col];
    OcPV r; r.x = make_f3(p.x, p.y, p.z); r.v = make_f3(p.w, q.x, q.y);
    return r;
}
template <int TW>
OC_HD void oc_st_pvd(OcStageSmem<TW>& sm, int row, int col, f3 x, f3 v, f3 d)
{
    const int s = row & (OC_RING - 1);
    sm.P[s][col] = make_float4(x.x, x.y, x.z, v.x);
    sm.Q[s][col] = make_float2(v.y, v.z);
    sm.D[s][0][col] = d.x; sm.D[s][1][col] = d.y; sm.D[s][2][col] = d.z;
}
OC_HD float4 oc_neg4(f3 f) { return make_float4(-f.x, -f.y, -f.z, 0.0f); }
template <class M> OC_HD void oc_acc(f3& F, f3 g, bool on)
{
    if (on) { F.x = M::add(F.x, g.x); F.y = M::add(F.y, g.y); F.z = M::add(F.z, g.z); }
}
template <class M> OC_HD void oc_acc4(f3& F, float4 g, bool on)
{
    if (on) { F.x = M::add(F.x, g.x); F.y = M::add(F.y, g.y); F.z = M::add(F.z, g.z); }
}

// Ctx: tid(), bx(), by(), bz(), sync(), smem()  (see DevCtx below and tests/emu/oc_emu.cu)
template <class M, int S, int TW, class Ctx>
OC_HD void oc_march_body(Ctx& ctx, const OcConst& c,
                         const float4* __restrict__ A, const float4* __restrict__ B,
                         float4* __restrict__ C, float4* __restrict__ Dst,
                         int ra, int rb, int RS, int x_halo)
{
    typedef OcStageSmem<TW> Smem;
    Smem* rings = reinterpret_cast<Smem*>(ctx.smem());
    const int tid = ctx.tid();
    const int s = tid / TW;              // stage (substep s+1 of this launch)
    const int i = tid - s * TW;          // column lane
    const int ci = i + 2;                // padded smem column
    const int U = c.U, V = c.V;
    // x_halo = 2*S columns either side are recomputed by the neighbouring strips; 0 when one strip
    // spans the whole cloth width (both strip edges are cloth edges: nothing to recompute)
    const int W_out = TW - 2 * x_halo;
    const int cx0 = ctx.bx() * W_out - x_halo;
    const int gi = cx0 + i;              // global column
    const int b = ctx.bz();
    const int r0 = ra + ctx.by() * RS;
    const int r1 = (r0 + RS < rb) ? r0 + RS : rb;

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

    Smem& in = rings[s];
    const bool col_ok = gi >= 0 && gi < U;
    // columns this CTA stores (valid after S substeps)
    const bool col_store = col_ok && i >= x_halo && i < TW - x_halo;

    // per-column constants
    const int gic = gi < 0 ? 0 : (gi >= U ? U - 1 : gi);
    const int gim = gic > 0 ? gic - 1 : 0;
    const float rh1_i = OC_LDG(c.rh1 + gic), rh2_i = OC_LDG(c.rh2 + gic);
    const float dx2_i = OC_LDG(c.dx2 + gic), dx2_m = OC_LDG(c.dx2 + gim);
    const bool has_l1 = gi - 1 >= 0, has_l2 = gi - 2 >= 0, has_r1 = gi + 1 < U, has_r2 = gi + 2 < U;
    const bool dup_r = gi == U - 3, dup_l = gi == U - 1;

    const float ydt = oc_rcp_bf(c.dt);   // reciprocal of dt for the branch-free velocity division

    // row constants (rest lengths that depend on the row only) are fetched one iteration ahead
    float rv1_n, rv2_n, dz2_n;
    {
        int r = first - OC_MARCH_LAG * (s + 1);
        r = r < 0 ? 0 : (r >= V ? V - 1 : r);
        rv1_n = OC_LDG(c.rv1 + r); rv2_n = OC_LDG(c.rv2 + r); dz2_n = OC_LDG(c.dz2 + r);
    }

    OcPV n1, n2;                         // own column, rows c+1 and c+2
    n1.x = n1.v = n2.x = n2.v = make_f3(0.f, 0.f, 0.f);
    f3 k1 = make_f3(0.f, 0.f, 0.f), k2a = k1, k2b = k1;     // carried vertical forces (on me, from rows above)

    for (int it = 0; it < n_it; ++it) {
        const int row = first - OC_MARCH_LAG * (s + 1) + it;      // row this stage works on
        const int lrow = first + it;                              // row stage 0 loads
        // ---- stage 0: issue the global loads of row lrow early -------------------------------------
        float4 la = make_float4(0.f, 0.f, 0.f, 0.f), lq = la;
        const bool doL = (s == 0) && lrow >= in_lo && lrow < in_hi && col_ok;
        if (doL) {
            long long o = oc_index(c, b, gi, lrow);
            la = A[o]; lq = B[o];
        }
        const float rv1_j = rv1_n, rv2_j = rv2_n, dz2_j = dz2_n;
        {
            int r = row + 1;
            r = r < 0 ? 0 : (r >= V ? V - 1 : r);
            rv1_n = OC_LDG(c.rv1 + r); rv2_n = OC_LDG(c.rv2 + r); dz2_n = OC_LDG(c.dz2 + r);
        }

        // ---- P phase: forward springs of row `row` -------------------------------------------------
        const bool doP = row >= plo_s && row < hi_s;
        f3 gH1, gH2, gV1, gV2, gD, gA, dme, F0;
        OcPV me;
        gH1 = gH2 = gV1 = gV2 = gD = gA = dme = F0 = make_f3(0.f, 0.f, 0.f);
        me.x = me.v = make_f3(0.f, 0.f, 0.f);
        if (doP) {
            if (row == plo_s) { n1 = oc_ld_pv<TW>(in, row, ci); n2 = oc_ld_pv<TW>(in, row + 1, ci); }
            me = n1; n1 = n2; n2 = oc_ld_pv<TW>(in, row + 2, ci);
            const int sl = row & (OC_RING - 1);
            dme = make_f3(in.D[sl][0][ci], in.D[sl][1][ci], in.D[sl][2][ci]);
            const OcPV a1 = oc_ld_pv<TW>(in, row, ci + 1);
            const OcPV a2 = oc_ld_pv<TW>(in, row, ci + 2);
            const OcPV d1 = oc_ld_pv<TW>(in, row + 1, ci + 1);
            const OcPV d0 = oc_ld_pv<TW>(in, row + 1, ci - 1);
            bool bad = false;
            const float rD = oc_len_bf<M>(M::add(dx2_i, dz2_j), bad);
            const float rA = oc_len_bf<M>(M::add(dx2_m, dz2_j), bad);
            gH1 = oc_spring_bf<M>(me.x, me.v, a1.x, a1.v, rh1_i, c.nks_struct, c.kd_struct, bad);
            gV1 = oc_spring_bf<M>(me.x, me.v, n1.x, n1.v, rv1_j, c.nks_struct, c.kd_struct, bad);
            gA  = oc_spring_bf<M>(me.x, me.v, d0.x, d0.v, rA,    c.nks_shear,  c.kd_shear,  bad);
            gD  = oc_spring_bf<M>(me.x, me.v, d1.x, d1.v, rD,    c.nks_shear,  c.kd_shear,  bad);
            gH2 = oc_spring_bf<M>(me.x, me.v, a2.x, a2.v, rh2_i, c.nks_bend,   c.kd_bend,   bad);
            gV2 = oc_spring_bf<M>(me.x, me.v, n2.x, n2.v, rv2_j, c.nks_bend,   c.kd_bend,   bad);
            if (M::kExact && bad) {
                // an operand left the range of the branch-free sequences (or the neighbour does not
                // exist and the lane holds garbage): redo this lane with the IEEE intrinsics
                const float sD = M::sqrt(M::add(dx2_i, dz2_j)), sA = M::sqrt(M::add(dx2_m, dz2_j));
                gH1 = oc_spring<M>(me.x, me.v, a1.x, a1.v, rh1_i, c.nks_struct, c.kd_struct);
                gV1 = oc_spring<M>(me.x, me.v, n1.x, n1.v, rv1_j, c.nks_struct, c.kd_struct);
                gA  = oc_spring<M>(me.x, me.v, d0.x, d0.v, sA,    c.nks_shear,  c.kd_shear);
                gD  = oc_spring<M>(me.x, me.v, d1.x, d1.v, sD,    c.nks_shear,  c.kd_shear);
                gH2 = oc_spring<M>(me.x, me.v, a2.x, a2.v, rh2_i, c.nks_bend,   c.kd_bend);
                gV2 = oc_spring<M>(me.x, me.v, n2.x, n2.v, rv2_j, c.nks_bend,   c.kd_bend);
            }
            in.FH[row & 1][0][ci] = oc_neg4(gH1);
            in.FH[row & 1][1][ci] = oc_neg4(gH2);
            in.FD[sl][0][ci] = oc_neg4(gD);
            in.FD[sl][1][ci] = oc_neg4(gA);
            F0 = oc_base_force<M>(c, me.v, oc_pinned(c, gi, row));
        }

        ctx.sync();
        OC_KEEP4(la); OC_KEEP4(lq);

        // ---- G phase: gather in the reference's order, integrate, collide, hand on ------------------
        const bool doG = row >= lo_s && row < hi_s;
        if (doG) {
            f3 F = F0;
            if (!oc_pinned(c, gi, row)) {
                const bool up1 = row - 1 >= 0, up2 = row - 2 >= 0, dn1 = row + 1 < V, dn2 = row + 2 < V;
                const int su = (row - 1) & (OC_RING - 1);
                oc_acc4<M>(F, in.FH[row & 1][0][ci - 1], has_l1);                 // 1  (i-1, j)   structural
                oc_acc<M>(F, gH1, has_r1);                                        // 2  (i+1, j)
                oc_acc<M>(F, k1, up1);                                            // 3  (i, j-1)
                oc_acc<M>(F, gV1, dn1);                                           // 4  (i, j+1)
                oc_acc4<M>(F, in.FD[su][0][ci - 1], has_l1 && up1);               // 5  (i-1, j-1) shear
                oc_acc4<M>(F, in.FD[su][1][ci + 1], has_r1 && up1);               // 6  (i+1, j-1)
                oc_acc<M>(F, gA, has_l1 && dn1);                                  // 7  (i-1, j+1)
                oc_acc<M>(F, gD, has_r1 && dn1);                                  // 8  (i+1, j+1)
                const float4 hl2 = in.FH[row & 1][1][ci - 2];
                oc_acc4<M>(F, hl2, has_l2);                                       // 9  (i-2, j)   bend
                oc_acc<M>(F, gH2, has_r2);                                        // 10 (i+2, j)
                oc_acc<M>(F, gH2, dup_r);                                         // 11 duplicate of the row's last bend spring (V:313)
                oc_acc4<M>(F, hl2, dup_l);
                oc_acc<M>(F, k2b, up2);                                           // 12 (i, j-2)
                oc_acc<M>(F, gV2, dn2);                                           // 13 (i, j+2)
                oc_acc<M>(F, gV2, row == V - 3);                                  // 14 duplicate of the column's last bend spring (V:319)
                oc_acc<M>(F, k2b, row == V - 1);
            }
            bool hit;
            const f3 xn = oc_integrate_collide<M>(c, me.x, dme, F, &hit);
            const float4 out = make_float4(xn.x, xn.y, xn.z, oc_u2f(hit ? OC_W_HIT : OC_W_PLAIN));
            if (s == S - 1) {
                if (col_store) C[oc_index(c, b, gi, row)] = out;                  // X(t+S)
            } else {
                // new X_last is the old X (V:438) unless the collider moved the particle (V:530)
                const f3 dn = hit ? make_f3(0.f, 0.f, 0.f)
                                  : make_f3(M::sub(xn.x, me.x.x), M::sub(xn.y, me.x.y), M::sub(xn.z, me.x.z));
                bool badv = false;
                f3 vn = oc_velocity_bf<M>(dn, c, ydt, badv);
                if (M::kExact && badv) vn = M::velocity(dn, c);
                oc_st_pvd<TW>(rings[s + 1], row, ci, xn, vn, dn);
                if (s == S - 2 && col_store && row >= r0 && row < r1) Dst[oc_index(c, b, gi, row)] = out;   // X(t+S-1)
            }
        }
        if (doP) { k2b = k2a; k2a = make_f3(-gV2.x, -gV2.y, -gV2.z); k1 = make_f3(-gV1.x, -gV1.y, -gV1.z); }

        // ---- stage 0: publish the loaded row into its own ring --------------------------------------
        if (doL) {
            const f3 d = oc_delta<M>(la, lq);
            bool badv = false;
            f3 v = oc_velocity_bf<M>(d, c, ydt, badv);
            if (M::kExact && badv) v = M::velocity(d, c);
            oc_st_pvd<TW>(rings[0], lrow, ci, make_f3(la.x, la.y, la.z), v, d);
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
#define OC_MARCH_MIN_CTAS(threads) ((threads) <= 128 ? 4 : ((threads) <= 256 ? 2 : 1))

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
