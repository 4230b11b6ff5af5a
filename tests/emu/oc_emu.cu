// tests/emu/oc_emu.cu — TEST INFRASTRUCTURE ONLY: runs the product's kernel BODIES on the CPU.
//
// The CUDA kernels of opencloth_b200/csrc are written as __host__ __device__ templates over an
// execution context.  This file provides a CPU context — one ucontext fiber per CUDA thread, a
// cooperative scheduler as __syncthreads, a malloc'd block as shared memory — and the same host
// sequencing (oc_host.h) as the C-ABI, so that `pytest -m "not gpu"` can check the kernels' logic
// (index arithmetic, pipeline hazards across barriers, accumulation order, band/halo protocol)
// bit for bit against the oracle without a GPU.  It is never loaded by the opencloth_b200 package
// and is not a CPU fallback: fibers make it thousands of times slower than the oracle.
//
// Host arithmetic is IEEE binary32 without contraction (-ffp-contract=off), which is exactly what
// the device intrinsics of MathExact compute, so exact-mode results must equal the GPU's.
// The file is compiled in PARTS (-DEMU_PART=k, tests/helpers.py builds them in parallel and links them): part 0 is the fiber
// machinery, the handle and the C entry points; parts 1-5 each hold the instantiations of one marching kernel's body, which
// dominate the compile time.  Without EMU_PART everything lands in one translation unit.
#ifndef EMU_PART
#define EMU_PART -1
#endif
#define EMU_HAS(p) (EMU_PART == -1 || EMU_PART == (p))
#if EMU_PART == -1
#define EMU_LOCAL static
#else
#define EMU_LOCAL
#endif

#include "../../opencloth_b200/csrc/oc_core.cuh"
#include "../../opencloth_b200/csrc/oc_host.h"
#if EMU_HAS(0)
#include "../../opencloth_b200/csrc/oc_gather.cuh"      // (holds non-template kernels: one part only)
#endif
#if EMU_HAS(0)
#include "../../opencloth_b200/csrc/oc_provot.cuh"      // (holds non-template kernels: one part only)
#endif
#if EMU_HAS(0)
#include "../../opencloth_b200/csrc/oc_normals.cuh"      // (holds non-template kernels: one part only)
#endif
#include "../../opencloth_b200/csrc/oc_march.cuh"
#include "../../opencloth_b200/csrc/oc_march2.cuh"
#include "../../opencloth_b200/csrc/oc_twin.cuh"
#include "../../opencloth_b200/csrc/oc_stream.cuh"
#include "../../opencloth_b200/csrc/oc_stream2.cuh"
#if EMU_HAS(0)
#include "../../opencloth_b200/csrc/oc_resident.cuh"      // (holds non-template kernels: one part only)
#include "../../opencloth_b200/csrc/oc_bandres.cuh"       // (host side only: the band plan; the kernel needs concurrently running CTAs)
#endif

#include <ucontext.h>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>
#include <functional>


// ------------------------------------------------------------------------------------------------
// fiber CTA
// ------------------------------------------------------------------------------------------------
struct EmuCta;
struct EmuCtx {
    int tid_, bx_, by_, bz_, nthreads_;
    unsigned char* smem_;
    EmuCta* cta;
    int tid() const { return tid_; }
    int nthreads() const { return nthreads_; }
    int bx() const { return bx_; }
    int by() const { return by_; }
    int bz() const { return bz_; }
    unsigned char* smem() const { return smem_; }
    bool wait_deps(const OcDep2&, const OcConst&, int, int) const { return true; }       // the emulator runs the tiles of a launch one after the other
    bool wait_deps_twin(const OcDep2&, const OcConst&, const OcSeg2&, const int*, const int*, const int*, const int*) const { return true; }
    // split-phase barrier (oc_stream.cuh): modelled as a barrier at the WAIT (all threads have arrived by then; stricter than the hardware)
    void bar_init(unsigned long long*, int) const {}
    void bar_arrive(unsigned long long*) const {}
    void bar_wait(unsigned long long*, unsigned) { sync(); }
    void sync();
};

struct EmuCta {
    ucontext_t sched;
    std::vector<ucontext_t> fib;
    std::vector<char*> stacks;
    std::vector<int> done;
    std::vector<long> nsync;
    std::function<void(EmuCtx&)> body;
    std::vector<EmuCtx> ctx;
    int cur;
};

#if EMU_HAS(0)
static EmuCta* g_cta = nullptr;
// Order in which the scheduler resumes the fibers between two barriers: 0 = ascending tid,
// 1 = descending, 2 = pseudo-random per round.  A kernel without intra-phase data races gives
// identical results under every order; tests run all three.
static int g_order = 0;
static unsigned g_rng = 12345u;
static const size_t kStack = 256 * 1024;

void EmuCtx::sync()
{
    EmuCta* c = cta;
    c->nsync[tid_]++;
    swapcontext(&c->fib[tid_], &c->sched);
}

static void fiber_main()
{
    EmuCta* c = g_cta;
    int t = c->cur;
    c->body(c->ctx[t]);
    c->done[t] = 1;
    swapcontext(&c->fib[t], &c->sched);
}

// run one CTA of nthreads threads; returns 0, or -1 if the threads disagree on the barrier count
EMU_LOCAL int run_cta(int nthreads, int bx, int by, int bz, size_t smem_bytes, const std::function<void(EmuCtx&)>& body)
{
    static EmuCta cta;              // stacks are reused between CTAs
    EmuCta* c = &cta;
    g_cta = c;
    c->body = body;
    if ((int)c->stacks.size() < nthreads) {
        size_t old = c->stacks.size();
        c->stacks.resize(nthreads);
        for (size_t t = old; t < (size_t)nthreads; ++t) c->stacks[t] = (char*)malloc(kStack);
    }
    c->fib.resize(nthreads); c->done.assign(nthreads, 0); c->nsync.assign(nthreads, 0); c->ctx.resize(nthreads);
    unsigned char* smem = (unsigned char*)aligned_alloc(128, (smem_bytes + 127) / 128 * 128 + 128);
    memset(smem, 0xCD, smem_bytes);            // garbage, like real shared memory
    for (int t = 0; t < nthreads; ++t) {
        c->ctx[t].tid_ = t; c->ctx[t].bx_ = bx; c->ctx[t].by_ = by; c->ctx[t].bz_ = bz; c->ctx[t].nthreads_ = nthreads;
        c->ctx[t].smem_ = smem; c->ctx[t].cta = c;
        getcontext(&c->fib[t]);
        c->fib[t].uc_stack.ss_sp = c->stacks[t];
        c->fib[t].uc_stack.ss_size = kStack;
        c->fib[t].uc_link = &c->sched;
        makecontext(&c->fib[t], fiber_main, 0);
    }
    int rc = 0;
    for (;;) {
        int alive = 0;
        unsigned off = 0, mul = 1;
        if (g_order == 2) {     // t -> (t*mul + off) mod n is a bijection when gcd(mul, n) = 1
            g_rng = g_rng * 1664525u + 1013904223u;
            off = (g_rng >> 8) % (unsigned)nthreads;
            static const unsigned odd[] = { 1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23 };
            for (;;) {
                g_rng = g_rng * 1664525u + 1013904223u; mul = odd[(g_rng >> 10) % 12];
                unsigned a = mul, b = (unsigned)nthreads;
                while (b) { unsigned r = a % b; a = b; b = r; }
                if (a == 1) break;
            }
        }
        for (int u = 0; u < nthreads; ++u) {
            int t = g_order == 0 ? u : (g_order == 1 ? nthreads - 1 - u : (int)(((unsigned)u * mul + off) % (unsigned)nthreads));
            if (c->done[t]) continue;
            c->cur = t;
            swapcontext(&c->sched, &c->fib[t]);
            if (!c->done[t]) alive++;
        }
        if (alive == 0) break;
        if (alive != nthreads) {               // some threads left the kernel while others wait at a barrier
            bool any_done = false;
            for (int t = 0; t < nthreads; ++t) any_done |= (c->done[t] != 0);
            if (any_done) {
                // legal in CUDA only if the finished threads never reach another barrier; our kernels
                // have uniform barrier counts, so flag it
                rc = -1;
            }
        }
    }
    for (int t = 1; t < nthreads; ++t) if (c->nsync[t] != c->nsync[0]) rc = -1;
    free(smem);
    return rc;
}

#else
int run_cta(int nthreads, int bx, int by, int bz, size_t smem_bytes, const std::function<void(EmuCtx&)>& body);
#endif

// ------------------------------------------------------------------------------------------------
// emulated handle: same state as the C-ABI handle, host memory
// ------------------------------------------------------------------------------------------------
struct EmuCloth {
    oc_params p;
    OcConst k;
    OcSeq q;
    OcHostTables T;
    std::vector<float4> buf[4];
    long long stored;
    int rows_own;
    long long barrier_errors;
    std::vector<unsigned> pins; std::vector<unsigned char> pin_rows;      // oc_set_pins
    EmuCloth* nb[2];            // linked row bands: upper / lower neighbour (same process, host memory)
    unsigned link_epoch;
};

#if EMU_HAS(1)
template <class M, int S, int TW>
static int emu_march(EmuCloth* e, const OcLaunch& L, int RS, int x_halo)
{
    const OcConst& k = e->k;
    int W_out = TW - 2 * x_halo;
    int nstrips = (k.U + W_out - 1) / W_out;
    int rows = L.rb - L.ra;
    if (RS <= 0 || RS > rows) RS = rows;
    int nseg = (rows + RS - 1) / RS;
    const float4* A = e->buf[L.src_a].data();
    const float4* B = e->buf[L.src_b].data();
    float4* C = e->buf[L.dst].data();
    float4* D = e->buf[L.dst_prev].data();
    int rc = 0;
    for (int bz = 0; bz < k.batch; ++bz)
        for (int by = 0; by < nseg; ++by)
            for (int bx = 0; bx < nstrips; ++bx) {
                int ra = L.ra, rb = L.rb;
                rc |= run_cta(S * TW, bx, by, bz, sizeof(OcStageSmem<TW>) * S, [&](EmuCtx& ctx) {
                    oc_march_body<M, S, TW, EmuCtx>(ctx, k, A, B, C, D, ra, rb, RS, x_halo);
                });
            }
    return rc;
}

template <class M>
static int emu_march_dispatch(EmuCloth* e, const OcLaunch& L, int TW, int RS)
{
    int x_halo = (e->k.U <= TW) ? 0 : 2 * L.S;
    if (TW - 2 * x_halo <= 0) return -2;
#define EMU_CASE(s, tw) if (L.S == s && TW == tw) return emu_march<M, s, tw>(e, L, RS, x_halo);
    EMU_CASE(1, 16) EMU_CASE(2, 16) EMU_CASE(3, 16)
    EMU_CASE(1, 32) EMU_CASE(2, 32) EMU_CASE(4, 32) EMU_CASE(8, 32)
    EMU_CASE(1, 64) EMU_CASE(2, 64) EMU_CASE(4, 64) EMU_CASE(8, 64)
    EMU_CASE(1, 128) EMU_CASE(2, 128) EMU_CASE(4, 128)
#undef EMU_CASE
    return -2;
}

int emu_disp_march(EmuCloth* e, const OcLaunch& L, int exact, int TW, int RS) { return exact ? emu_march_dispatch<MathExact>(e, L, TW, RS) : emu_march_dispatch<MathFast>(e, L, TW, RS); }
#else
int emu_disp_march(EmuCloth* e, const OcLaunch& L, int exact, int TW, int RS);
#endif

#if EMU_HAS(2)
template <class M, int WC>
static int emu_march2(EmuCloth* e, const OcLaunch& L, int RS)
{
    const OcConst& k = e->k;
    const int x_halo = (k.U <= WC) ? 0 : 2;
    const int W_out = WC - 2 * x_halo;
    const int nstrips = (k.U + W_out - 1) / W_out;
    const int rows = L.rb - L.ra;
    // RS = rows per segment | (rows per segment of the edge strips) << 8   (0: same; see OcSeg2)
    OcSeg2 seg;
    seg.rs = RS & 0xff; seg.rs_e = (RS >> 8) & 0xff; seg.nstrips = nstrips; seg.rev = (RS >> 17) & 1;      // bit 17: segments bottom to top
    if (seg.rs <= 0 || seg.rs > rows) seg.rs = rows;
    if (seg.rs_e <= 0) seg.rs_e = seg.rs;
    oc_seg2_finish(seg, rows);
    const float4* A = e->buf[L.src_a].data();
    const float4* B = e->buf[L.src_b].data();
    float4* C = e->buf[L.dst].data();
    // linked row bands: the same OcPeer2 the library passes (peer stores into the neighbours' halo rows); the flag
    // words stay null, the emulator runs bands and tiles one after the other
    OcDep2 dep = OcDep2();
    if (e->q.linked) {
        for (int sd = 0; sd < 2; ++sd)
            if (e->nb[sd]) dep.peer.c[sd] = e->nb[sd]->buf[L.dst].data() - (long long)e->nb[sd]->k.row_lo * k.U;
        dep.peer.epoch = ++e->link_epoch; dep.peer.ra = L.ra; dep.peer.rb = L.rb; dep.peer.nstrips = nstrips;
        if ((rows - 1) % seg.rs + 1 < 2) return -4;          // the library's planner never produces this (fix_last_segment)
    }
    int rc = 0;
    for (int bz = 0; bz < k.batch; ++bz)
        for (int t = 0; t < oc_seg2_tiles(seg); ++t) {
            int bx, by;
            oc_seg2_tile(seg, t, bx, by);
            int ra = L.ra, rb = L.rb;
            rc |= run_cta(WC / 2, bx, by, bz, sizeof(OcSmem2<WC>), [&](EmuCtx& ctx) {
                oc_march2_body<M, WC, EmuCtx>(ctx, k, A, B, C, ra, rb, seg, x_halo, dep);
            });
        }
    return rc;
}
template <class M>
static int emu_march2_dispatch(EmuCloth* e, const OcLaunch& L, int WC, int RS)
{
    if (WC == 16)  return emu_march2<M, 16>(e, L, RS);
    if (WC == 32)  return emu_march2<M, 32>(e, L, RS);
    if (WC == 64)  return emu_march2<M, 64>(e, L, RS);
    if (WC == 128) return emu_march2<M, 128>(e, L, RS);
    return -2;
}

int emu_disp_march2(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS) { return exact ? emu_march2_dispatch<MathExact>(e, L, WC, RS) : emu_march2_dispatch<MathFast>(e, L, WC, RS); }
#else
int emu_disp_march2(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS);
#endif


#if EMU_HAS(3) || EMU_HAS(4) || EMU_HAS(5)
// kernel 5: twin tiles (oc_twin.cuh).  RS = rows per segment (0: two segments); bit 16 of RS: pair the cloths of a batch
// (needs an even batch) instead of two segments of a strip.  The number of segments must be even when segments are paired.
template <class M, int WC, int kVariant>
static int emu_twin(EmuCloth* e, const OcLaunch& L, int RS)
{
    const OcConst& k = e->k;
    const int x_halo = (k.U <= WC) ? 0 : 2;
    const int W_out = WC - 2 * x_halo;
    const int nstrips = (k.U + W_out - 1) / W_out;
    const int rows = L.rb - L.ra;
    OcTwinMap map; map.pair_cloths = (RS >> 16) & 1;
    if (map.pair_cloths && k.batch % 2 != 0) return -2;
    OcSeg2 seg;
    seg.rs = RS & 0xffff; seg.nstrips = nstrips; seg.rev = (RS >> 17) & 1;      // bit 17: segments bottom to top
    if (seg.rs <= 0 || seg.rs > rows) seg.rs = map.pair_cloths ? rows : (rows + 1) / 2;
    seg.rs_e = seg.rs;
    oc_seg2_finish(seg, rows);
    if (!map.pair_cloths && seg.nseg_all % 2 != 0) return -2;
    const float4* A = e->buf[L.src_a].data();
    const float4* B = e->buf[L.src_b].data();
    float4* C = e->buf[L.dst].data();
    OcDep2 dep = OcDep2();
    if (e->q.linked) {
        for (int sd = 0; sd < 2; ++sd)
            if (e->nb[sd]) dep.peer.c[sd] = e->nb[sd]->buf[L.dst].data() - (long long)e->nb[sd]->k.row_lo * k.U;
        dep.peer.epoch = ++e->link_epoch; dep.peer.ra = L.ra; dep.peer.rb = L.rb; dep.peer.nstrips = nstrips;
        if ((rows - 1) % seg.rs + 1 < 2) return -4;
    }
    int rc = 0;
    const int nk = map.pair_cloths ? seg.nseg_all : seg.nseg_all / 2;
    const int nz = map.pair_cloths ? k.batch / 2 : k.batch;
    for (int bz = 0; bz < nz; ++bz)
        for (int by = 0; by < nk; ++by)
            for (int bx = 0; bx < nstrips; ++bx) {
                int ra = L.ra, rb = L.rb;
                if (kVariant == 2)
                    rc |= run_cta(WC / 2, bx, by, bz, sizeof(OcSmemS2<WC, M::kExact>), [&](EmuCtx& ctx) {
                        oc_stream2_body<M, WC, EmuCtx>(ctx, k, A, B, C, ra, rb, seg, x_halo, map, dep);
                    });
                else if (kVariant == 1)
                    rc |= run_cta(WC, bx, by, bz, sizeof(OcSmemS<WC, M::kExact>), [&](EmuCtx& ctx) {
                        oc_stream_body<M, WC, EmuCtx>(ctx, k, A, B, C, ra, rb, seg, x_halo, map, dep);
                    });
                else
                    rc |= run_cta(WC, bx, by, bz, sizeof(OcSmemT<WC, M::kExact>), [&](EmuCtx& ctx) {
                        oc_twin_body<M, WC, EmuCtx>(ctx, k, A, B, C, ra, rb, seg, x_halo, map, dep);
                    });
            }
    return rc;
}
template <class M, int kVariant>
static int emu_twin_dispatch(EmuCloth* e, const OcLaunch& L, int WC, int RS)
{
    if (WC == 8)   return emu_twin<M, 8, kVariant>(e, L, RS);
    if (WC == 16)  return emu_twin<M, 16, kVariant>(e, L, RS);
    if (WC == 32)  return emu_twin<M, 32, kVariant>(e, L, RS);
    if (WC == 64)  return emu_twin<M, 64, kVariant>(e, L, RS);
    if (WC == 128) return emu_twin<M, 128, kVariant>(e, L, RS);
    return -2;
}

#endif
#if EMU_HAS(3)
int emu_disp_twin(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS) { return exact ? emu_twin_dispatch<MathExact, 0>(e, L, WC, RS) : emu_twin_dispatch<MathFast, 0>(e, L, WC, RS); }
#else
int emu_disp_twin(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS);
#endif
#if EMU_HAS(4)
int emu_disp_stream(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS) { return exact ? emu_twin_dispatch<MathExact, 1>(e, L, WC, RS) : emu_twin_dispatch<MathFast, 1>(e, L, WC, RS); }
#else
int emu_disp_stream(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS);
#endif
#if EMU_HAS(5)
int emu_disp_stream2(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS) { return exact ? emu_twin_dispatch<MathExact, 2>(e, L, WC, RS) : emu_twin_dispatch<MathFast, 2>(e, L, WC, RS); }
#else
int emu_disp_stream2(EmuCloth* e, const OcLaunch& L, int exact, int WC, int RS);
#endif

#if EMU_HAS(0)
// kernel 4: one CTA per cloth, all the substeps of the launch inside; TW = threads of the CTA (0: like the library)
template <class M>
static int emu_resident(EmuCloth* e, const OcLaunch& L, int threads)
{
    const OcConst& k = e->k;
    const int N = k.U * k.V;
    if (e->q.band || N > OC_RESIDENT_MAX_PARTICLES) return -2;
    if (threads <= 0) { threads = (3 * N + 31) / 32 * 32; if (threads > OC_RESIDENT_THREADS) threads = OC_RESIDENT_THREADS; }
    const float4* A = e->buf[L.src_a].data();
    const float4* B = e->buf[L.src_b].data();
    float4* D = e->buf[L.dst].data();
    float4* Dp = e->buf[L.dst_prev].data();
    int rc = 0, S = L.S;
    for (int b = 0; b < k.batch; ++b)
        rc |= run_cta(threads, b, 0, 0, OcResidentSmem::bytes(N), [&](EmuCtx& ctx) {
            oc_resident_body<M, EmuCtx>(ctx, k, A, B, D, Dp, S);
        });
    return rc;
}

template <class M>
static void emu_gather(EmuCloth* e, const OcLaunch& L)
{
    const OcConst& k = e->k;
    const float4* A = e->buf[L.src_a].data();
    const float4* B = e->buf[L.src_b].data();
    float4* C = e->buf[L.dst].data();
    float4* D = e->buf[L.dst_prev].data();
    for (int b = 0; b < k.batch; ++b)
        for (int j = L.ra; j < L.rb; ++j)
            for (int i = 0; i < k.U; ++i) {
                const long long o = oc_index(k, b, i, j);
                if (e->q.xv) { float4 v; C[o] = oc_gather_particle<M, true>(k, A, B, b, i, j, &v); D[o] = v; }
                else C[o] = oc_gather_particle<M>(k, A, B, b, i, j);
            }
}

// ApplyProvotDynamicInverse with the library's kernel bodies: per-particle gather into V (Euler forms), or the in-place
// sweep of X by rows, columns and the skewed shear wavefront (Verlet form; `threads` = CTA size of the wavefront)
template <class M>
static int emu_provot(EmuCloth* e, int threads)
{
    const OcConst& k = e->k;
    float4* X = e->buf[e->q.ia].data();
    float4* S = e->buf[e->q.ib].data();
    int rc = 0;
    if (e->q.xv) {
        std::vector<float4> out((size_t)e->stored);
        for (int b = 0; b < k.batch; ++b)
            for (int j = 0; j < k.V; ++j)
                for (int i = 0; i < k.U; ++i) out[oc_index(k, b, i, j)] = oc_provot_v_particle<M>(k, X, S, b, i, j);
        memcpy(S, out.data(), out.size() * sizeof(float4));
        return 0;
    }
    for (long long t = 0; t < e->stored; ++t) oc_provot_materialize(X, S, t);
    for (int b = 0; b < k.batch; ++b) {
        for (int j = 0; j < k.V; ++j) oc_provot_row<M, 1>(k, X, b, j);
        for (int i = 0; i < k.U; ++i) oc_provot_col<M, 1>(k, X, b, i);
        rc |= run_cta(threads, b, 0, 0, 16, [&](EmuCtx& ctx) { oc_provot_shear_body<M, EmuCtx>(ctx, k, X); });
        for (int j = 0; j < k.V; ++j) oc_provot_row<M, 2>(k, X, b, j);
        for (int i = 0; i < k.U; ++i) oc_provot_col<M, 2>(k, X, b, i);
    }
    return rc;
}


#if EMU_HAS(0)
// ------------------------------------------------------------------------------------------------
// kernel 8 (oc_k_bandres): ALL the CTAs of the grid run concurrently - their threads are fibers of one scheduler - because
// the bands wait for each other's boundary rows inside the launch.  A thread is READY, AT a CTA BARRIER (released when all
// the live threads of its CTA are there), POLLING (a tagged word has not arrived: resumed in the next round) or DONE.
// ------------------------------------------------------------------------------------------------
struct EmuGrid;
struct EmuGridCtx {
    int tid_, nthreads_, band_, nbands_;
    unsigned char* smem_;
    EmuGrid* g; int fib;
    int tid() const { return tid_; }
    int nthreads() const { return nthreads_; }
    int band() const { return band_; }
    int nbands() const { return nbands_; }
    unsigned char* smem() const { return smem_; }
    void sync();
    bool sync_and(bool ok);
    void put(unsigned long long* p, float v, unsigned tag) const { *(volatile unsigned long long*)p = ((unsigned long long)tag << 32) | (unsigned long long)oc_f2u(v); }
    bool get(const OcConst&, const unsigned long long* p, int stride, unsigned tag, float4& a);
};
struct EmuGrid {
    ucontext_t sched;
    std::vector<ucontext_t> fib;
    std::vector<char*> stacks;
    std::vector<int> state;                 // 0 ready, 1 at barrier, 2 polling, 3 done
    std::vector<long> nsync;
    std::vector<EmuGridCtx> ctx;
    std::vector<int> vote;                  // per thread: its argument of the current sync_and
    std::function<void(EmuGridCtx&)> body;
    int cur; long polls;
};
static EmuGrid* g_grid = nullptr;
void EmuGridCtx::sync() { g->state[fib] = 1; g->nsync[fib]++; swapcontext(&g->fib[fib], &g->sched); }
bool EmuGridCtx::sync_and(bool ok)
{
    g->vote[fib] = ok ? 1 : 0;
    sync();
    bool r = true;
    for (int t = 0; t < nthreads_; ++t) r = r && g->vote[band_ * nthreads_ + t] != 0;
    sync();                                 // (everybody has read the votes before anybody casts the next one)
    return r;
}
bool EmuGridCtx::get(const OcConst&, const unsigned long long* p, int stride, unsigned tag, float4& a)
{
    for (;;) {
        const unsigned long long w0 = *(volatile const unsigned long long*)p, w1 = *(volatile const unsigned long long*)(p + stride), w2 = *(volatile const unsigned long long*)(p + 2 * stride);
        if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag && ((unsigned)(w2 >> 32) & 0x7fffffffu) == tag) {
            a = make_float4(oc_u2f((unsigned)w0), oc_u2f((unsigned)w1), oc_u2f((unsigned)w2), oc_u2f(((unsigned)(w2 >> 32) & 0x80000000u) ? OC_W_HIT : OC_W_PLAIN));
            return true;
        }
        if (++g->polls > 50000000L) { a = make_float4(0.f, 0.f, 0.f, oc_u2f(OC_W_PLAIN)); return false; }      // a protocol bug: never arrives
        g->state[fib] = 2;
        swapcontext(&g->fib[fib], &g->sched);
    }
}
static void grid_fiber_main()
{
    EmuGrid* g = g_grid;
    const int f = g->cur;
    g->body(g->ctx[f]);
    g->state[f] = 3;
    swapcontext(&g->fib[f], &g->sched);
}
// returns 0, -1 (barrier counts disagree inside a CTA / threads left a CTA whose others wait at a barrier), -5 (no progress)
static int run_grid(int nctas, int nthreads, size_t smem_bytes, const std::function<void(EmuGridCtx&)>& body)
{
    static EmuGrid grid;
    EmuGrid* g = &grid;
    g_grid = g;
    g->body = body;
    const int nf = nctas * nthreads;
    if ((int)g->stacks.size() < nf) {
        const size_t old = g->stacks.size();
        g->stacks.resize(nf);
        for (size_t t = old; t < (size_t)nf; ++t) g->stacks[t] = (char*)malloc(kStack);
    }
    g->fib.resize(nf); g->state.assign(nf, 0); g->nsync.assign(nf, 0); g->ctx.resize(nf); g->vote.assign(nf, 1); g->polls = 0;
    std::vector<unsigned char*> smem(nctas);
    for (int b = 0; b < nctas; ++b) {
        smem[b] = (unsigned char*)aligned_alloc(128, (smem_bytes + 127) / 128 * 128 + 128);
        memset(smem[b], 0xCD, smem_bytes);
        for (int t = 0; t < nthreads; ++t) {
            const int f = b * nthreads + t;
            EmuGridCtx& x = g->ctx[f];
            x.tid_ = t; x.nthreads_ = nthreads; x.band_ = b; x.nbands_ = nctas; x.smem_ = smem[b]; x.g = g; x.fib = f;
            getcontext(&g->fib[f]);
            g->fib[f].uc_stack.ss_sp = g->stacks[f];
            g->fib[f].uc_stack.ss_size = kStack;
            g->fib[f].uc_link = &g->sched;
            makecontext(&g->fib[f], grid_fiber_main, 0);
        }
    }
    int rc = 0;
    long rounds = 0;
    for (;;) {
        int alive = 0;
        for (int bb = 0; bb < nctas; ++bb) {
            const int b = g_order == 1 ? nctas - 1 - bb : bb;
            for (int u = 0; u < nthreads; ++u) {
                const int t = g_order == 0 ? u : (g_order == 1 ? nthreads - 1 - u : (int)(((unsigned)u * 7u + (unsigned)rounds * 3u) % (unsigned)nthreads));
                const int f = b * nthreads + t;
                if (g->state[f] == 3 || g->state[f] == 1) continue;
                g->state[f] = 0;
                g->cur = f;
                swapcontext(&g->sched, &g->fib[f]);
            }
            // barrier of this CTA: released when every live thread is at it
            int at = 0, live = 0, done = 0;
            for (int t = 0; t < nthreads; ++t) { const int st = g->state[b * nthreads + t]; at += st == 1; live += st != 3; done += st == 3; }
            if (live > 0 && at == live) {
                if (done > 0) rc = -1;
                long ns = -1;
                for (int t = 0; t < nthreads; ++t) if (g->state[b * nthreads + t] == 1) { const long v = g->nsync[b * nthreads + t]; if (ns < 0) ns = v; else if (v != ns) rc = -1; }
                for (int t = 0; t < nthreads; ++t) if (g->state[b * nthreads + t] == 1) g->state[b * nthreads + t] = 0;
            }
            alive += live;
        }
        if (alive == 0) break;
        if (++rounds > 20000000L) { rc = -5; break; }
    }
    for (int b = 0; b < nctas; ++b) free(smem[b]);
    return rc;
}

// kernel 8: TW = threads per CTA, RS = "SMs" the plan may use (bands <= RS); all the substeps of the launch in one grid
template <class M>
static int emu_bandres(EmuCloth* e, const OcLaunch& L, int TW, int RS)
{
    const OcConst& k = e->k;
    if (e->q.band || k.batch != 1 || e->p.provot || e->q.xv) return -2;
    int nb = 0, rmax = 0;
    if (!oc_bandres_plan(k.U, k.V, RS > 0 ? RS : 6, &nb, &rmax)) return -2;
    const size_t NG = (size_t)k.U * k.V;
    static std::vector<unsigned long long> ex;
    static unsigned epoch = 0;
    if (ex.size() != NG * 6) { ex.assign(NG * 6, 0ull); epoch = 0; }
    const float4* A = e->buf[L.src_a].data();
    const float4* B = e->buf[L.src_b].data();
    float4* d0 = e->buf[L.dst].data();
    float4* d1 = e->buf[L.dst_prev].data();
    const unsigned ep = epoch;
    const int rc = run_grid(nb, TW > 0 ? TW : 32, OcBandresSmem::bytes(k.U, rmax), [&](EmuGridCtx& ctx) {
        oc_bandres_body<M, EmuGridCtx>(ctx, k, A, B, d0, d1, L.S, ex.data(), nullptr, ep, rmax);
    });
    epoch += (unsigned)L.S;
    return rc;
}
#endif

extern "C" {

void* emu_create(const oc_params* p)
{
    if (!p || p->nx < 3 || p->ny < 3 || p->batch < 1) return nullptr;
    EmuCloth* e = new EmuCloth();
    e->p = *p;
    memset(&e->k, 0, sizeof(e->k));
    oc_host_geometry(e->p, e->k, e->q);
    e->rows_own = e->p.row_end - e->p.row_begin;
    e->stored = e->k.cloth_stride * p->batch;
    oc_host_derive_scalars(e->p, e->k);
    oc_host_build_tables(p->nx, p->ny, p->fullsize, e->T);
    oc_host_bind_tables(e->k, e->T.t.data(), e->T);
    const float* xs = &e->T.t[e->T.xs]; const float* zs = &e->T.t[e->T.zs];
    for (int b = 0; b < 4; ++b) e->buf[b].resize((size_t)e->stored);
    long long per = e->k.cloth_stride;
    for (long long t = 0; t < e->stored; ++t) {
        long long r = t % per;
        int j = (int)(r / e->k.U) + e->k.row_lo, i = (int)(r % e->k.U);
        float4 v = make_float4(xs[i], p->fullsize + 1, zs[j], oc_u2f(OC_W_PLAIN));
        e->buf[0][t] = e->buf[1][t] = e->buf[2][t] = e->buf[3][t] = v;
        if (e->q.xv) e->buf[1][t] = make_float4(0.0f, 0.0f, 0.0f, oc_u2f(OC_W_PLAIN));
    }
    e->barrier_errors = 0;
    e->nb[0] = e->nb[1] = nullptr; e->link_epoch = 0;
    return e;
}
void emu_destroy(void* h) { delete (EmuCloth*)h; }
void emu_set_order(int order) { g_order = order; }

int emu_set_params(void* h, const oc_params* p)
{
    EmuCloth* e = (EmuCloth*)h;
    oc_params q = *p;
    q.row_begin = e->p.row_begin; q.row_end = e->p.row_end; q.halo_rows = e->p.halo_rows;
    e->p = q;
    oc_host_derive_scalars(e->p, e->k);
    return 0;
}

int emu_upload(void* h, const float* X, const float* XL)
{
    EmuCloth* e = (EmuCloth*)h;
    const OcConst& k = e->k;
    long long t = 0;
    for (int b = 0; b < k.batch; ++b)
        for (int j = e->p.row_begin; j < e->p.row_end; ++j)
            for (int i = 0; i < k.U; ++i, ++t) {
                long long o = oc_index(k, b, i, j);
                e->buf[e->q.ia][o] = make_float4(X[t * 3], X[t * 3 + 1], X[t * 3 + 2], oc_u2f(OC_W_PLAIN));
                e->buf[e->q.ib][o] = make_float4(XL[t * 3], XL[t * 3 + 1], XL[t * 3 + 2], oc_u2f(OC_W_PLAIN));
            }
    if (e->q.band) e->q.fresh = e->q.kmax;
    return 0;
}

int emu_download(void* h, float* X, float* XL)
{
    EmuCloth* e = (EmuCloth*)h;
    const OcConst& k = e->k;
    long long t = 0;
    for (int b = 0; b < k.batch; ++b)
        for (int j = e->p.row_begin; j < e->p.row_end; ++j)
            for (int i = 0; i < k.U; ++i, ++t) {
                long long o = oc_index(k, b, i, j);
                float4 a = e->buf[e->q.ia][o];
                float4 q = oc_hit(a.w) ? a : e->buf[e->q.ib][o];
                if (X)  { X[t * 3] = a.x; X[t * 3 + 1] = a.y; X[t * 3 + 2] = a.z; }
                if (XL) { XL[t * 3] = q.x; XL[t * 3 + 1] = q.y; XL[t * 3 + 2] = q.z; }
            }
    return 0;
}

// kernel: 1 = gather, 2 = march, 3 = march2 (TW = window columns).  k = substeps per launch, TW = column window, RS = rows per segment
// (0 = one segment).  Returns 0, -1 on a barrier-count mismatch, -2 unsupported variant, -3 halo exhausted.
static int g_provot_threads = 32;
void emu_set_provot_threads(int t) { g_provot_threads = t > 0 ? t : 32; }
int emu_step(void* h, int n, int kernel, int exact, int k, int TW, int RS)
{
    EmuCloth* e = (EmuCloth*)h;
    if (e->q.xv) kernel = 1;
    if (e->p.provot || e->q.xv) k = 1;
    if (e->q.band && !e->q.linked && e->q.fresh + n > e->q.kmax) return -3;
    if (e->q.linked && kernel != 3 && kernel != 5 && kernel != 6 && kernel != 7) return -2;
    int rc = 0;
    while (n > 0) {
        OcLaunch L;
        // same stage-count choice as oc_step (TW = 16 exists only here and also has a 3-stage build)
        int kk = 1;
        if (kernel == 2) { int w = n < k ? n : k; kk = (TW == 16 && w <= 3) ? w : oc_host_pick_stages(w); }
        if (kernel == 4) kk = OC_RESIDENT_MAX_STEPS;
        if (kernel == 8) kk = OC_BANDRES_MAX_STEPS;
        if (e->p.provot || e->q.xv) kk = 1;
        oc_host_next_launch(e->q, n, kk, L);
        if (kernel == 4) {
            int r = exact ? emu_resident<MathExact>(e, L, TW) : emu_resident<MathFast>(e, L, TW);
            if (r == -2) return -2;
            rc |= r;
        } else if (kernel == 8) {
            int r = exact ? emu_bandres<MathExact>(e, L, TW, RS) : emu_bandres<MathFast>(e, L, TW, RS);
            if (r == -2 || r == -5) return r;
            rc |= r;
        } else if (kernel == 7) {
            int r = emu_disp_stream2(e, L, exact, TW, RS);
            if (r == -2 || r == -4) return r;
            rc |= r;
        } else if (kernel == 6) {
            int r = emu_disp_stream(e, L, exact, TW, RS);
            if (r == -2 || r == -4) return r;
            rc |= r;
        } else if (kernel == 5) {
            int r = emu_disp_twin(e, L, exact, TW, RS);
            if (r == -2 || r == -4) return r;
            rc |= r;
        } else if (kernel == 3) {
            int r = emu_disp_march2(e, L, exact, TW, RS);
            if (r == -2) return -2;
            rc |= r;
        } else if (kernel == 2) {
            int r = emu_disp_march(e, L, exact, TW, RS);
            if (r == -2) return -2;
            rc |= r;
        } else {
            if (exact) emu_gather<MathExact>(e, L); else emu_gather<MathFast>(e, L);
        }
        if (e->p.provot) rc |= exact ? emu_provot<MathExact>(e, g_provot_threads) : emu_provot<MathFast>(e, g_provot_threads);
    }
    return rc;
}

// copy the rows src sends towards `side` into dst's halo on the opposite side (both buffers)
int emu_halo_copy(void* hsrc, int side, void* hdst)
{
    EmuCloth* s = (EmuCloth*)hsrc; EmuCloth* d = (EmuCloth*)hdst;
    int r0s, ns, r0d, nd;
    if (!oc_host_halo_rows(s->p, s->q.band, side, true, &r0s, &ns)) return 0;
    if (!oc_host_halo_rows(d->p, d->q.band, 1 - side, false, &r0d, &nd)) return -1;
    if (ns != nd || r0s != r0d) return -1;
    const int U = s->k.U;
    for (int which = 0; which < 2; ++which) {
        const float4* src = s->buf[which == 0 ? s->q.ia : s->q.ib].data() + (long long)(r0s - s->k.row_lo) * U;
        float4* dst = d->buf[which == 0 ? d->q.ia : d->q.ib].data() + (long long)(r0d - d->k.row_lo) * U;
        memcpy(dst, src, (size_t)ns * U * sizeof(float4));
    }
    return 0;
}
// per-vertex normals of the current state (oc_normals.cuh), 3 floats per vertex
int emu_normals(void* h, float* out)
{
    EmuCloth* e = (EmuCloth*)h;
    const OcConst& k = e->k;
    if (e->q.band) return -1;
    const float4* X = e->buf[e->q.ia].data();
    for (int b = 0; b < k.batch; ++b)
        for (int j = 0; j < k.V; ++j)
            for (int i = 0; i < k.U; ++i) {
                const f3 n = oc_vertex_normal<MathExact>(k, X, b, i, j);
                float* o = out + (((long long)b * k.V + j) * k.U + i) * 3;
                o[0] = n.x; o[1] = n.y; o[2] = n.z;
            }
    return 0;
}
// oc_set_pins: same bitmap + row summary as the library
int emu_set_pins(void* h, int cloth, const int* idx, int n)
{
    EmuCloth* e = (EmuCloth*)h;
    const int U = e->k.U, V = e->k.V, B = e->k.batch;
    const long long per = (long long)U * V;
    if (e->pins.empty()) {
        e->pins.assign((size_t)((per * B + 31) / 32), 0u);
        e->pin_rows.assign((size_t)B * V, 0);
        for (int b = 0; b < B; ++b) {
            for (int q : { 0, U - 1 }) { long long bit = b * per + q; e->pins[bit >> 5] |= 1u << (bit & 31); }
            e->pin_rows[(size_t)b * V] = 1;
        }
    }
    for (int b = (cloth < 0 ? 0 : cloth); b < (cloth < 0 ? B : cloth + 1); ++b) {
        for (long long q = 0; q < per; ++q) { long long bit = b * per + q; e->pins[bit >> 5] &= ~(1u << (bit & 31)); }
        for (int j = 0; j < V; ++j) e->pin_rows[(size_t)b * V + j] = 0;
        for (int k = 0; k < n; ++k) {
            if (idx[k] < 0 || idx[k] >= per) return -1;
            long long bit = b * per + idx[k]; e->pins[bit >> 5] |= 1u << (bit & 31);
            e->pin_rows[(size_t)b * V + idx[k] / U] = 1;
        }
    }
    e->k.pins = e->pins.data(); e->k.pin_rows = e->pin_rows.data();
    return 0;
}
// Linked row bands (oc_band_link_local): h[0..n) are the bands of one cloth in row order.  Pulls the halos once; from
// then on every emu_step(band, 1, kernel 3, ...) pushes its boundary rows into the neighbours (step all bands in turn).
int emu_band_link(void** h, int n)
{
    for (int b = 0; b < n; ++b) {
        EmuCloth* e = (EmuCloth*)h[b];
        if (!e->q.band || e->p.halo_rows < 2 || e->rows_own < 4) return -1;
        e->nb[0] = b > 0 ? (EmuCloth*)h[b - 1] : nullptr;
        e->nb[1] = b + 1 < n ? (EmuCloth*)h[b + 1] : nullptr;
        if ((e->p.row_begin > 0) != (e->nb[0] != nullptr) || (e->p.row_end < e->p.ny) != (e->nb[1] != nullptr)) return -1;
        if (e->nb[0] && e->nb[0]->p.row_end != e->p.row_begin) return -1;
        e->q.linked = true; e->link_epoch = 0;
    }
    for (int b = 0; b < n; ++b) {
        EmuCloth* e = (EmuCloth*)h[b];
        const long long U = e->k.U;
        for (int sd = 0; sd < 2; ++sd) {
            EmuCloth* o = e->nb[sd];
            if (!o) continue;
            const int r0 = sd == 0 ? e->p.row_begin - 2 : e->p.row_end;
            for (int which = 0; which < 2; ++which) {
                const int bi = which == 0 ? e->q.ia : e->q.ib;
                if (o->q.ia != e->q.ia || o->q.ib != e->q.ib) return -2;
                memcpy(e->buf[bi].data() + (long long)(r0 - e->k.row_lo) * U, o->buf[bi].data() + (long long)(r0 - o->k.row_lo) * U, (size_t)2 * U * sizeof(float4));
            }
        }
    }
    return 0;
}
// host pointer + float4 count of a halo region (same meaning as oc_halo_send_region / oc_halo_recv_region)
int emu_halo_region(void* h, int side, int which, int send, void** ptr, size_t* count)
{
    EmuCloth* e = (EmuCloth*)h;
    *ptr = nullptr; *count = 0;
    int r0, rows;
    if (!oc_host_halo_rows(e->p, e->q.band, side, send != 0, &r0, &rows)) return 0;
    float4* base = e->buf[which == 0 ? e->q.ia : e->q.ib].data();
    *ptr = (void*)(base + (long long)(r0 - e->k.row_lo) * e->k.U);
    *count = (size_t)rows * e->k.U;
    return 0;
}
// Brute-force check of the tile arithmetic of OcSeg2 / OcDep2 (host only).  Previous launch: rows [pra, prb) cut into
// segments of prs (edge strips prs_e) rows; this launch likewise.  0: every row of every strip is covered exactly once and
// every previous tile that wrote rows [r0-2, r1+2) of a tile's own or neighbouring strips is among the <= 12 flags that
// tile waits for; 1: the host would not chain these two launches; negative: a violation.
int emu_check_tiling(int nstrips, int pra, int prb, int prs, int prs_e, int ra, int rb, int rs, int rs_e, int ignore_height)
{
    OcSeg2 pseg, seg;
    pseg.rs = prs; pseg.rs_e = prs_e; pseg.nstrips = nstrips; pseg.rev = 0; oc_seg2_finish(pseg, prb - pra);
    seg.rs = rs; seg.rs_e = rs_e; seg.nstrips = nstrips; seg.rev = 0; oc_seg2_finish(seg, rb - ra);
    // coverage of this launch
    std::vector<int> cover((size_t)nstrips * (rb - ra), 0);
    for (int t = 0; t < oc_seg2_tiles(seg); ++t) {
        int bx, by, r0, r1;
        oc_seg2_tile(seg, t, bx, by);
        if (bx < 0 || bx >= nstrips || oc_seg2_index(seg, bx, by) != t) return -3;
        oc_seg2_rows(seg, bx, by, ra, rb, r0, r1);
        for (int r = r0; r < r1; ++r) cover[(size_t)bx * (rb - ra) + (r - ra)]++;
    }
    for (size_t i = 0; i < cover.size(); ++i) if (cover[i] != 1) return -1;
    // the host's chaining condition; ignore_height drops its "tiles of at least 32 rows" part (a performance rule)
    if (!oc_dep2_chainable(seg, ra, rb, pseg, pra, prb, ignore_height != 0)) return 1;
    OcDep2 d = {};
    d.mode = 1; d.pra = pra; d.prb = prb; d.pseg = pseg;
    for (int t = 0; t < oc_seg2_tiles(seg); ++t) {
        int bx, by, r0, r1;
        oc_seg2_tile(seg, t, bx, by);
        oc_seg2_rows(seg, bx, by, ra, rb, r0, r1);
        if (r0 >= r1) r1 = r0;                                          // an empty tile still orders itself (and publishes)
        int deps[12];
        for (int k = 0; k < 12; ++k) deps[k] = oc_dep2_index(d, bx - 1 + k / 4, k % 4, r0, r1);
        {   // one flag word per tile index: the tile must wait for its namesake of the previous launch (oc_dep2_chainable (3))
            int px, py, p0, p1;
            oc_seg2_tile(pseg, t, px, py);
            oc_seg2_rows(pseg, px, py, pra, prb, p0, p1);
            if (t < oc_seg2_tiles(pseg) && p0 < p1) {
                bool found = false;
                for (int k = 0; k < 12; ++k) found |= deps[k] == t;
                if (!found) return -4;
            }
        }
        if (r0 >= r1) continue;
        for (int u = 0; u < oc_seg2_tiles(pseg); ++u) {
            int px, py, p0, p1;
            oc_seg2_tile(pseg, u, px, py);
            oc_seg2_rows(pseg, px, py, pra, prb, p0, p1);
            if (p0 >= p1 || px < bx - 1 || px > bx + 1) continue;
            if (p1 <= r0 - 2 || p0 >= r1 + 2) continue;                 // wrote nothing this tile reads
            bool found = false;
            for (int k = 0; k < 12; ++k) found |= deps[k] == u;
            if (!found) return -2;
        }
    }
    return 0;
}

// the collider's bounding sphere as the kernels use it (centre xyz, squared radius; +inf = shortcut off)
int emu_bounding_sphere(void* h, float out[4])
{
    const OcConst& k = ((EmuCloth*)h)->k;
    out[0] = k.bs_c[0]; out[1] = k.bs_c[1]; out[2] = k.bs_c[2]; out[3] = k.bs_r2;
    return 0;
}
int emu_halo_refreshed(void* h) { ((EmuCloth*)h)->q.fresh = 0; return 0; }
int emu_halo_budget(void* h) { EmuCloth* e = (EmuCloth*)h; return e->q.band ? e->q.kmax - e->q.fresh : 0x7fffffff; }

// oc_k_bandres: the host-side plan (bands, rows of the tallest band, shared memory) and the row cut
int emu_bandres_plan(int U, int V, int sm_count, int* nb, int* rmax, unsigned long long* smem)
{
    if (!oc_bandres_plan(U, V, sm_count, nb, rmax)) return 0;
    *smem = (unsigned long long)OcBandresSmem::bytes(U, *rmax);
    return 1;
}
void emu_bandres_rows(int V, int nb, int b, int* r0, int* r1) { oc_bandres_rows(V, nb, b, *r0, *r1); }
} // extern "C"

#endif      // EMU_HAS(0)
