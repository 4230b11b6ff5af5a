// oc_bandres.cuh — kernel 8: mid-size cloths RESIDENT in the shared memory of MANY SMs (one row band per CTA).
//
// Between the cloth that fits one CTA (oc_k_resident, <= 1536 particles) and the cloth that fills the machine with
// marching tiles (>= 10^6 particles) a step of the marching kernels is pure latency: a 512 x 512 cloth gives every CTA a
// tile of 8-16 rows behind a 6-row pipeline fill, one launch and one round of tile flags per substep - 21 us per step
// where the arithmetic needs 3.  Here the cloth is cut into one band of rows per CTA (at most one CTA per SM, launched
// cooperatively so that all of them are resident), each band keeps its rows - and two halo rows either side - in shared
// memory for ALL the substeps of an oc_step call, and per substep the bands exchange nothing but their two boundary
// rows, through a small global buffer and one flag word per band:
//   P1  every own particle gathers its twelve springs from shared memory (six packed pairs, oc_spring2; the reference's
//       order, see oc_gather.cuh), IntegrateVerlet, EllipsoidCollision -> X(t+1) into the other position buffer; boundary
//       rows also go to the exchange buffer of this substep's parity; barrier;
//   P2  every own particle: X - X_last and the velocity, in place;
//       the halo rows' new positions come from the exchange buffer, their velocities are derived from them like everybody
//       else's; barrier.
// The exchange (OC_BANDRES_LL, default): every coordinate travels as ONE 64-bit word {float, tag = epoch + substep} - a scalar
// 64-bit store is single-copy atomic, so a word whose tag matches carries its data; the collider flag rides in the top
// bit of z's tag.  The reader polls the words it needs: one store and one load through L2 per substep, no fence, no flag
// round trip (the "LL" protocol of collective libraries).  OC_BANDRES_LL=0: float4 rows (st.cg), fence, one flag word per
// band released by thread 0, acquired by threads 0 / 32 of the neighbours, then ld.cg - two L2 round trips more
// (256^2: 6.2 us per substep instead of the tagged words' figure in DESIGN.md 4.4).
// A band's exchange rows of parity p are overwritten two substeps later, after the band has received its neighbours' rows of
// the substep in between - which they send only after they have read parity p.  Arithmetic and order are those of the
// other kernels: bit-identical to the reference in exact mode.
// Limits: one whole cloth (no batch, no row band), Verlet, no Provot pass; the tallest band must fit shared memory
// (oc_bandres_plan: up to ~300 k particles, e.g. 512 x 576).
#pragma once
#include "oc_core.cuh"
#include "oc_march2.cuh"      // oc_flag_wait

#define OC_BANDRES_THREADS 512
#define OC_BANDRES_MAX_STEPS 4096           /* substeps per launch (bounds the run time of one launch) */
#define OC_BANDRES_MAX_BANDS 1024
#ifndef OC_BANDRES_LL
#define OC_BANDRES_LL 1
#endif
// bytes of the exchange buffer for a cloth of n particles (two parities; LL: three 64-bit words per particle)
#define OC_BANDRES_EX_BYTES(n) ((size_t)(n) * (OC_BANDRES_LL ? 48 : 32))
#define OC_BANDRES_SMEM_MAX (224 * 1024)    /* dynamic shared memory of one CTA (one CTA per SM) */

// shared memory (strides of the TALLEST band so that every CTA has the same layout): NL = (rmax + 4) * U local particles
// incl. halo rows, NO = rmax * U own particles.  Positions and velocities are float4 records - one LDS.128 and one address
// per neighbour and quantity instead of three scalar loads from three arrays (the gather is bound by its instruction
// count); w of a position is the collider flag, as in global memory.
struct OcBandresSmem {
    unsigned char* base; int NL, NO;
    OC_HD float4* X(int buf) const { return reinterpret_cast<float4*>(base) + (size_t)buf * NL; }     // positions, two buffers (t / t-1, swapped every substep)
    OC_HD float4* Vv() const { return reinterpret_cast<float4*>(base) + (size_t)2 * NL; }              // velocity (w unused)
    OC_HD float* D(int k) const { return reinterpret_cast<float*>(base + (size_t)48 * NL) + (size_t)k * NO; }   // X(t) - X_last(t), own rows
    static OC_HD size_t bytes(int U, int rmax) { return (size_t)48 * (rmax + 4) * U + (size_t)12 * rmax * U; }
};

// rows [r0, r1) of band b of nb over V rows: even cut, every band >= 2 rows when nb <= V / 2
OC_HD void oc_bandres_rows(int V, int nb, int b, int& r0, int& r1)
{
    r0 = (int)((long long)b * V / nb); r1 = (int)((long long)(b + 1) * V / nb);
}

// Bands and rows of the tallest band for a U x V cloth on a device of sm_count SMs; false if the state of the tallest band
// does not fit the shared memory of one SM.  At most one band per SM, every band at least two rows (the stencil reaches two
// rows: a band's halo then comes from its two neighbours only), no more bands than give every CTA's threads a particle.
OC_HD bool oc_bandres_plan(int U, int V, int sm_count, int* nb, int* rmax)
{
    if (U < 1 || V < 4 || sm_count < 1) return false;
    int n = sm_count < V / 2 ? sm_count : V / 2;
    if (n > OC_BANDRES_MAX_BANDS) n = OC_BANDRES_MAX_BANDS;
    const long long want = ((long long)U * V + OC_BANDRES_THREADS - 1) / OC_BANDRES_THREADS;
    if (n > want) n = want < 1 ? 1 : (int)want;
    const int r = (V + n - 1) / n;
    if (OcBandresSmem::bytes(U, r) > (size_t)OC_BANDRES_SMEM_MAX) return false;
    *nb = n; *rmax = r;
    return true;
}

#ifdef __CUDACC__
// the flag of a neighbour band: a short busy poll first (the neighbour is normally a few hundred cycles away), then the
// backed-off wait with its time-out and poison word (oc_flag_wait)
__device__ __forceinline__ bool oc_bandres_wait(const OcConst& c, const unsigned* p, unsigned want)
{
    for (int k = 0; k < 4096; ++k) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if ((int)(v - want) >= 0) return true;
    }
    return oc_flag_wait<false>(c, p, want);
}

// tagged 64-bit words: {float bits, tag} in one scalar store / load (single-copy atomic), strong at GPU scope (L2)
__device__ __forceinline__ void oc_bandres_put(unsigned long long* p, float v, unsigned tag)
{
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
}
// x, y, z of one particle (words p, p + stride, p + 2 stride) once all three carry `tag`; a.w = the collider flag.
// false: they never came (time-out / poison word, like oc_flag_wait)
__device__ __forceinline__ bool oc_bandres_get(const OcConst& c, const unsigned long long* p, int stride, unsigned tag, float4& a)
{
    unsigned long long w0, w1, w2;
    unsigned long long t0 = 0;
    for (unsigned spins = 1;; ++spins) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(p) : "memory");
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(p + stride) : "memory");
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w2) : "l"(p + 2 * stride) : "memory");
        if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag && ((unsigned)(w2 >> 32) & 0x7fffffffu) == tag) break;
        if ((spins & 1023u) != 0u) continue;
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        const bool poisoned = (*(volatile unsigned long long*)(c.dbg_cnt + 2) >> 40) != 0ull;
        if (poisoned || t - t0 > ((c.dbg & 32) ? 50000000ull : 2000000000ull)) {
            atomicAdd(c.dbg_cnt + 2, 1ull << 40);
            if (c.err) { *(volatile unsigned*)c.err = 1u; __threadfence_system(); }
            a = make_float4(0.0f, 0.0f, 0.0f, oc_u2f(OC_W_PLAIN));
            return false;
        }
    }
    a = make_float4(__uint_as_float((unsigned)w0), __uint_as_float((unsigned)w1), __uint_as_float((unsigned)w2),
                    oc_u2f(((unsigned)(w2 >> 32) & 0x80000000u) ? OC_W_HIT : OC_W_PLAIN));
    return true;
}

#endif      // __CUDACC__

// The kernel body over an execution context (the device's OcBandresCtx below; tests/emu/oc_emu.cu runs the same body on
// the CPU with every CTA's threads as fibers, so that the exchange protocol is checked without a GPU):
//   tid, nthreads, band (CTA index), nbands, smem, sync, sync_and, put / get (the tagged 64-bit words)
template <class M, class Ctx>
OC_HD void oc_bandres_body(Ctx& ctx, const OcConst& c, const float4* __restrict__ A, const float4* __restrict__ B,
                           float4* __restrict__ dst, float4* __restrict__ dst_prev, int n_steps,
                           void* __restrict__ ex_, unsigned* __restrict__ flags, unsigned epoch, int rmax)
{
    const int U = c.U, V = c.V;
    const int tid = ctx.tid(), T = ctx.nthreads(), b = ctx.band(), nb = ctx.nbands();
    int r0, r1;
    oc_bandres_rows(V, nb, b, r0, r1);
    const int R = r1 - r0;
    const int jbase = r0 - 2;                                        // global row of local row 0
    const int h0 = r0 - 2 < 0 ? 0 : r0 - 2, h1 = r1 + 2 > V ? V : r1 + 2;   // rows held locally
    OcBandresSmem s;
    s.base = ctx.smem(); s.NL = (rmax + 4) * U; s.NO = rmax * U;
    const long long goff = -(long long)c.row_lo * U;                 // storage offset of global row 0 (whole cloths: 0)
    const float ydt = oc_rcp_bf(c.dt);
    const size_t NG = (size_t)U * V;

    auto velocity = [&](f3 d) {
        bool bad = false;
        f3 v = oc_velocity_bf<M>(d, c, ydt, bad);
        if (M::kExact && bad) v = M::velocity(d, c);
        return v;
    };

    // ---- load: X(t) into buffer 0, X(t-1) into buffer 1, derived state --------------------------------------
    for (int lp = (h0 - jbase) * U + tid; lp < (h1 - jbase) * U; lp += T) {
        const long long g = goff + (long long)jbase * U + lp;
        const float4 a = A[g], q = B[g];
        s.X(0)[lp] = a; s.X(1)[lp] = q;
        const f3 d = oc_delta<M>(a, q);
        const f3 v = velocity(d);
        s.Vv()[lp] = make_float4(v.x, v.y, v.z, 0.0f);
        const int op = lp - 2 * U;
        if (op >= 0 && op < R * U) { s.D(0)[op] = d.x; s.D(1)[op] = d.y; s.D(2)[op] = d.z; }
    }
    ctx.sync();

    int cur = 0;
    for (int step = 1; step <= n_steps; ++step) {
        const int nxt = cur ^ 1;
        const float4* xc = s.X(cur); float4* xn = s.X(nxt); float4* vv = s.Vv();
        float4* exw = reinterpret_cast<float4*>(ex_) + (size_t)(step & 1) * NG;                                  // flag protocol: rows of float4
        (void)exw; (void)flags;
        unsigned long long* exl = reinterpret_cast<unsigned long long*>(ex_) + (size_t)(step & 1) * NG * 3;       // tagged words: [row][x, y, z][column]
        const unsigned tag = (epoch + (unsigned)step) & 0x7fffffffu;
        // ---- P1: gather, integrate, collide ------------------------------------------------------------------
        for (int op = tid; op < R * U; op += T) {
            const int oj = op / U, i = op - oj * U, j = r0 + oj, lp = op + 2 * U;
            const float4 xm4 = xc[lp], vm4 = vv[lp];
            const f3 xm = make_f3(xm4.x, xm4.y, xm4.z);
            const f3 vm = make_f3(vm4.x, vm4.y, vm4.z);
            const f3 d  = make_f3(s.D(0)[op], s.D(1)[op], s.D(2)[op]);
            const bool pinned = oc_pinned(c, 0, i, j);
            f3 F = oc_base_force<M>(c, vm, pinned);
            // two springs of the particle at once; a missing partner is a ghost at unit distance moving with the particle
            // (its force is not added).  `keep`: the pair's forces, for the springs the reference's list holds twice.
            auto pair = [&](int qa, bool ea, int qb, bool eb, float2 rest, float2 nks, float2 kd, OcPair3* keep) {
                if (!ea && !eb) return;
                const int a = ea ? qa : lp, bb = eb ? qb : lp;
                const float4 xa = xc[a], xb = xc[bb], va = vv[a], vb = vv[bb];
                OcPair3 qx, qv;
                qx.x = make_float2(ea ? xa.x : xm.x + 1.0f, eb ? xb.x : xm.x + 1.0f);
                qx.y = make_float2(xa.y, xb.y); qx.z = make_float2(xa.z, xb.z);
                qv.x = make_float2(va.x, vb.x); qv.y = make_float2(va.y, vb.y); qv.z = make_float2(va.z, vb.z);
                bool bad = false;
                OcPair3 f = oc_spring2<M>(xm, vm, qx, qv, M::kExact ? rest : p_mul(rest, nks), nks, kd, c.one, bad);
                if (M::kExact && bad) {                                      // rare: IEEE intrinsics
                    const f3 fa = oc_spring<M>(xm, vm, make_f3(qx.x.x, qx.y.x, qx.z.x), make_f3(qv.x.x, qv.y.x, qv.z.x), rest.x, nks.x, kd.x);
                    const f3 fb = oc_spring<M>(xm, vm, make_f3(qx.x.y, qx.y.y, qx.z.y), make_f3(qv.x.y, qv.y.y, qv.z.y), rest.y, nks.y, kd.y);
                    f.x = make_float2(fa.x, fb.x); f.y = make_float2(fa.y, fb.y); f.z = make_float2(fa.z, fb.z);
                }
                if (ea) { F.x = M::add(F.x, f.x.x); F.y = M::add(F.y, f.y.x); F.z = M::add(F.z, f.z.x); }
                if (eb) { F.x = M::add(F.x, f.x.y); F.y = M::add(F.y, f.y.y); F.z = M::add(F.z, f.z.y); }
                if (keep) *keep = f;
            };
            auto shear_rest = [&](int ia, int ib, int jj) {                  // sqrt(dx2[ia] + dz2[jj]), sqrt(dx2[ib] + dz2[jj])
                bool badr = false;
                const float sa = M::add(c.dx2[ia], c.dz2[jj]), sb = M::add(c.dx2[ib], c.dz2[jj]);
                float2 r = oc_sqrt2<M>(make_float2(sa, sb), badr);
                if (M::kExact && badr) r = make_float2(M::sqrt(sa), M::sqrt(sb));
                return r;
            };
            if (!pinned) {
                const bool l1 = i - 1 >= 0, r1e = i + 1 < U, l2 = i - 2 >= 0, r2e = i + 2 < U;
                const bool u1 = j - 1 >= 0, d1 = j + 1 < V, u2 = j - 2 >= 0, d2 = j + 2 < V;
                const int im1 = l1 ? i - 1 : 0, im2 = l2 ? i - 2 : 0, jm1 = u1 ? j - 1 : 0, jm2 = u2 ? j - 2 : 0;
                const float2 nS = p_bc(c.nks_struct), kS = p_bc(c.kd_struct), nSh = p_bc(c.nks_shear), kSh = p_bc(c.kd_shear);
                const float2 nB = p_bc(c.nks_bend), kB = p_bc(c.kd_bend);
                pair(lp - 1, l1, lp + 1, r1e, make_float2(c.rh1[im1], c.rh1[i]), nS, kS, nullptr);                            //  1  2   V:288-291
                pair(lp - U, u1, lp + U, d1,  make_float2(c.rv1[jm1], c.rv1[j]), nS, kS, nullptr);                            //  3  4   V:294-297
                pair(lp - U - 1, l1 && u1, lp - U + 1, r1e && u1, shear_rest(im1, i, jm1), nSh, kSh, nullptr);                //  5  6   V:301-305
                pair(lp + U - 1, l1 && d1, lp + U + 1, r1e && d1, shear_rest(im1, i, j),   nSh, kSh, nullptr);                //  7  8
                OcPair3 k;
                k.x = k.y = k.z = make_float2(0.0f, 0.0f);
                pair(lp - 2, l2, lp + 2, r2e, make_float2(c.rh2[im2], c.rh2[i]), nB, kB, &k);                                 //  9 10   V:309-314
                if (i == U - 3 && r2e) { F.x = M::add(F.x, k.x.y); F.y = M::add(F.y, k.y.y); F.z = M::add(F.z, k.z.y); }      // 11      the row's last bend spring twice (V:313)
                if (i == U - 1 && l2)  { F.x = M::add(F.x, k.x.x); F.y = M::add(F.y, k.y.x); F.z = M::add(F.z, k.z.x); }
                pair(lp - 2 * U, u2, lp + 2 * U, d2, make_float2(c.rv2[jm2], c.rv2[j]), nB, kB, &k);                          // 12 13   V:315-320
                if (j == V - 3 && d2) { F.x = M::add(F.x, k.x.y); F.y = M::add(F.y, k.y.y); F.z = M::add(F.z, k.z.y); }       // 14      the column's last bend spring twice (V:319)
                if (j == V - 1 && u2) { F.x = M::add(F.x, k.x.x); F.y = M::add(F.y, k.y.x); F.z = M::add(F.z, k.z.x); }
            }
            bool hit;
            const f3 n = oc_integrate_collide<M>(c, xm, d, F, &hit);
            const float w = oc_u2f(hit ? OC_W_HIT : OC_W_PLAIN);
            xn[lp] = make_float4(n.x, n.y, n.z, w);
            if ((oj < 2 && b > 0) || (oj >= R - 2 && b + 1 < nb)) {
#if OC_BANDRES_LL
                unsigned long long* e = exl + (size_t)j * 3 * U + i;
                ctx.put(e, n.x, tag); ctx.put(e + U, n.y, tag); ctx.put(e + 2 * U, n.z, tag | (hit ? 0x80000000u : 0u));
#else
                __stcg(exw + (size_t)j * U + i, make_float4(n.x, n.y, n.z, w));
#endif
            }
        }
        ctx.sync();
#if !OC_BANDRES_LL
        if (tid == 0) {                       // (fence + release by one thread after the barrier: cumulative over the CTA's stores, as in OcDevCtx2::publish)
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flags + b), "r"(epoch + (unsigned)step) : "memory");
        }
#endif
        // ---- P2: derived state of the own rows (in place: nobody reads another particle's in this phase) ---------
        for (int op = tid; op < R * U; op += T) {
            const int lp = op + 2 * U;
            const float4 nw = xn[lp], ol = xc[lp];
            f3 d = make_f3(0.0f, 0.0f, 0.0f);
            if (!oc_hit(nw.w)) d = make_f3(M::sub(nw.x, ol.x), M::sub(nw.y, ol.y), M::sub(nw.z, ol.z));
            const f3 v = velocity(d);
            s.D(0)[op] = d.x; s.D(1)[op] = d.y; s.D(2)[op] = d.z;
            vv[lp] = make_float4(v.x, v.y, v.z, 0.0f);
        }
        bool ok = true;
        if (step < n_steps) {
            // ---- the neighbours' boundary rows of this substep ---------------------------------------------------
#if !OC_BANDRES_LL
            if (tid == 0 && b > 0)       ok = oc_bandres_wait(c, flags + b - 1, epoch + (unsigned)step);
            if (tid == 32 && b + 1 < nb) ok = oc_bandres_wait(c, flags + b + 1, epoch + (unsigned)step);
            if (!ctx.sync_and(ok)) return;                               // a neighbour never arrived: error word set (oc_flag_wait)
#endif
            const int n_up = (r0 - h0) * U, n_dn = (h1 - r1) * U;
            for (int e = tid; e < n_up + n_dn; e += T) {
                const int lp = e < n_up ? (h0 - jbase) * U + e : (r1 - jbase) * U + (e - n_up);
                float4 a;
#if OC_BANDRES_LL
                const int lj = lp / U, i = lp - lj * U;
                ok &= ctx.get(c, exl + ((long long)(jbase + lj) * 3 * U + i), U, tag, a);
#else
                a = __ldcg(exw + ((long long)jbase * U + lp));
#endif
                const float4 ol = xc[lp];
                f3 d = make_f3(0.0f, 0.0f, 0.0f);
                if (!oc_hit(a.w)) d = make_f3(M::sub(a.x, ol.x), M::sub(a.y, ol.y), M::sub(a.z, ol.z));
                const f3 v = velocity(d);
                xn[lp] = a;
                vv[lp] = make_float4(v.x, v.y, v.z, 0.0f);
            }
        }
        if (!ctx.sync_and(ok)) return;                                   // a neighbour's rows never arrived: error word set (oc_bandres_get)
        cur = nxt;
    }
    // ---- store: X(t+n) and, for more than one substep, X(t+n-1) --------------------------------------------
    for (int op = tid; op < R * U; op += T) {
        const int lp = op + 2 * U;
        const long long g = goff + (long long)r0 * U + op;
        dst[g] = s.X(cur)[lp];
        if (n_steps > 1) dst_prev[g] = s.X(cur ^ 1)[lp];
    }
}

#ifdef __CUDACC__
struct OcBandresCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ int band() const { return blockIdx.x; }
    __device__ __forceinline__ int nbands() const { return gridDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ bool sync_and(bool ok) const { return __syncthreads_and(ok) != 0; }
    __device__ __forceinline__ unsigned char* smem() const { extern __shared__ __align__(16) unsigned char oc_dyn_smem[]; return oc_dyn_smem; }
    __device__ __forceinline__ void put(unsigned long long* p, float v, unsigned tag) const { oc_bandres_put(p, v, tag); }
    __device__ __forceinline__ bool get(const OcConst& c, const unsigned long long* p, int stride, unsigned tag, float4& a) const { return oc_bandres_get(c, p, stride, tag, a); }
};

template <class M, int TT>
__global__ void __launch_bounds__(TT, 1)
oc_k_bandres(const __grid_constant__ OcConst c, const float4* __restrict__ A, const float4* __restrict__ B,
             float4* __restrict__ dst, float4* __restrict__ dst_prev, int n_steps,
             void* __restrict__ ex_, unsigned* __restrict__ flags, unsigned epoch, int rmax)
{
    OcBandresCtx ctx;
    oc_bandres_body<M, OcBandresCtx>(ctx, c, A, B, dst, dst_prev, n_steps, ex_, flags, epoch, rmax);
}
#endif
