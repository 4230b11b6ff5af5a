// oc_march.cu — host side of the marching kernel: variant table, launch geometry, launch.
#include "oc_march.cuh"
#include "oc_march2.cuh"
#include "oc_twin.cuh"
#include "oc_stream.cuh"
#include "oc_stream2.cuh"
#include <cstdlib>
#include <cstdio>

extern "C" const void* oc_march_fn_exact_32(int S);
extern "C" const void* oc_march_fn_exact_64(int S);
extern "C" const void* oc_march_fn_exact_128(int S);
extern "C" const void* oc_march_fn_fast_32(int S);
extern "C" const void* oc_march_fn_fast_64(int S);
extern "C" const void* oc_march_fn_fast_128(int S);

static const void* march_fn(bool exact, int TW, int S)
{
    switch (TW) {
    case 32:  return exact ? oc_march_fn_exact_32(S)  : oc_march_fn_fast_32(S);
    case 64:  return exact ? oc_march_fn_exact_64(S)  : oc_march_fn_fast_64(S);
    case 128: return exact ? oc_march_fn_exact_128(S) : oc_march_fn_fast_128(S);
    default:  return nullptr;
    }
}

static size_t stage_smem(int TW)
{
    switch (TW) {
    case 32:  return sizeof(OcStageSmem<32>);
    case 64:  return sizeof(OcStageSmem<64>);
    case 128: return sizeof(OcStageSmem<128>);
    default:  return 0;
    }
}

// resident CTAs per SM for each variant, filled by oc_march_configure: [exact][tw_idx][S]
static int g_occ[2][3][OC_MARCH_MAX_STAGES + 1];
static int tw_index(int TW) { return TW == 32 ? 0 : (TW == 64 ? 1 : 2); }

// width of the column window for a cloth of nx columns stepped S substeps per launch
static int pick_tw(int nx, int S)
{
    const char* env = getenv("OC_MARCH_TW");      // development override: 32 | 64 | 128
    if (env) {
        int tw = atoi(env);
        if ((tw == 32 || tw == 64 || (tw == 128 && S <= 4)) && tw - 4 * S > 0) return tw;
    }
    if (nx <= 32) return 32;
    if (nx <= 64) return 64;
    if (S <= 4) return 128;
    return 64;
}

int oc_march_configure(int device)
{
    (void)device;
    const int tws[3] = { 32, 64, 128 };
    for (int e = 0; e < 2; ++e)
        for (int t = 0; t < 3; ++t)
            for (int S = 1; S <= OC_MARCH_MAX_STAGES; ++S) {
                const void* fn = march_fn(e != 0, tws[t], S);
                g_occ[e][t][S] = 0;
                if (!fn) continue;
                size_t smem = stage_smem(tws[t]) * S;
                cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (err != cudaSuccess) return (int)err;
                int occ = 0;
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, tws[t] * S, smem);
                if (err != cudaSuccess) return (int)err;
                g_occ[e][t][S] = occ;
            }
    return 0;
}

// Launch geometry.  Rows are cut into nseg segments of RS rows; every segment costs
// RS + (pipeline fill) iterations, so few long segments are efficient but the grid must also fill
// sm_count * occupancy CTA slots in whole waves.  Pick the RS that maximises
//   useful row-iterations / (waves * slots * iterations per CTA).
int oc_march_plan(const OcConst& c, int S, int ra, int rb, int sm_count, int occ_hint, OcMarchPlan* pl)
{
    const int U = c.U;
    int TW = pick_tw(U, S);
    int x_halo = (U <= TW) ? 0 : 2 * S;
    int W_out = TW - 2 * x_halo;
    if (W_out <= 0) return -1;
    int nstrips = (U + W_out - 1) / W_out;
    int rows = rb - ra;
    if (rows <= 0) return -1;
    int occ = occ_hint > 0 ? occ_hint : 1;
    long long slots = (long long)sm_count * occ;
    const int fill = OC_MARCH_LAG * S + 2 * (S - 1) + 2;      // iterations beyond RS per segment
    int best_rs = rows; double best = -1.0;
    const char* env = getenv("OC_MARCH_RS");
    if (env && atoi(env) > 0) {
        best_rs = atoi(env);
        if (best_rs > rows) best_rs = rows;
    } else {
        for (int nseg = 1; nseg <= rows; ++nseg) {
            int rs = (rows + nseg - 1) / nseg;
            if (rs < 8 && nseg > 1) break;
            int ns = (rows + rs - 1) / rs;
            long long ctas = (long long)nstrips * ns * c.batch;
            long long waves = (ctas + slots - 1) / slots;
            double eff = (double)rows * nstrips * c.batch / ((double)waves * slots * (rs + fill));
            if (eff > best * 1.0001) { best = eff; best_rs = rs; }
        }
    }
    pl->TW = TW; pl->S = S; pl->x_halo = x_halo; pl->W_out = W_out; pl->nstrips = nstrips;
    pl->RS = best_rs; pl->nseg = (rows + best_rs - 1) / best_rs; pl->threads = TW * S;
    pl->smem = stage_smem(TW) * S;
    return 0;
}

cudaError_t oc_march_launch(const OcConst& c, bool exact, int S, int ra, int rb, int sm_count,
                            const float4* A, const float4* B, float4* C, float4* Dst,
                            cudaStream_t stream, int* n_launches)
{
    *n_launches = 0;
    OcMarchPlan pl;
    int TW = pick_tw(c.U, S);
    int occ = g_occ[exact ? 1 : 0][tw_index(TW)][S];
    if (oc_march_plan(c, S, ra, rb, sm_count, occ, &pl) != 0) return cudaErrorInvalidValue;
    const void* fn = march_fn(exact, pl.TW, S);
    if (!fn) return cudaErrorInvalidDeviceFunction;
    if (pl.nseg > 65535 || c.batch > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid(pl.nstrips, pl.nseg, c.batch), block(pl.threads, 1, 1);
    OcConst cc = c;
    int RS = pl.RS, xh = pl.x_halo;
    void* args[] = { &cc, (void*)&A, (void*)&B, (void*)&C, (void*)&Dst, &ra, &rb, &RS, &xh };
    cudaError_t e = cudaLaunchKernel(fn, grid, block, args, pl.smem, stream);
    if (e == cudaSuccess) *n_launches = 1;
    return e;
}

// ------------------------------------------------------------------------------------------------
// two-columns-per-thread kernel (oc_march2.cuh), one substep per launch
// ------------------------------------------------------------------------------------------------
extern "C" const void* oc_march2_fn_exact(int WC);
extern "C" const void* oc_march2_fn_fast(int WC);
static int g_occ2[2][2];      // [exact][WC == 128]
// tile-height ratios of the speed classes (percent), measured on B200 (profiles/ time line)
#define OC_MARCH2_EDGE_EXACT 24
#define OC_MARCH2_EDGE_FAST  10

static size_t smem2(int WC) { return WC == 64 ? sizeof(OcSmem2<64>) : sizeof(OcSmem2<128>); }
static int pick_wc(int nx)
{
    const char* env = getenv("OC_MARCH2_WC");
    if (env && (atoi(env) == 64 || atoi(env) == 128)) return atoi(env);
    return nx <= 64 ? 64 : 128;
}

int oc_march2_configure(int device)
{
    (void)device;
    for (int e = 0; e < 2; ++e)
        for (int w = 0; w < 2; ++w) {
            const int WC = w ? 128 : 64;
            const void* fn = e ? oc_march2_fn_exact(WC) : oc_march2_fn_fast(WC);
            cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2(WC));
            if (err != cudaSuccess) return (int)err;
            int occ = 0;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, WC / 2, smem2(WC));
            if (err != cudaSuccess) return (int)err;
            g_occ2[e][w] = occ;
        }
    return 0;
}

int oc_march2_nstrips(int nx)
{
    const int WC = pick_wc(nx);
    const int W_out = WC - 2 * ((nx <= WC) ? 0 : 2);
    return (nx + W_out - 1) / W_out;
}

// Linked row bands: the first and the last segment of a strip hold (at least) the two rows pushed to the neighbour,
// so that exactly one tile per strip produces, and reads, each halo (OcPeer2).  Segments are uniform with the
// remainder in the last one: grow rs until that remainder is at least 2 rows.
static int fix_last_segment(int rows, int rs)
{
    while (rs < rows && (rows - 1) % rs + 1 < 2) ++rs;
    return rs;
}

int oc_march2_plan(const OcConst& c, bool exact, bool chained, bool linked, int ra, int rb, int sm_count, int occ_hint, OcMarchPlan* pl, OcSeg2* seg)
{
    const int U = c.U;
    const int WC = pick_wc(U);
    const int x_halo = (U <= WC) ? 0 : 2;
    const int W_out = WC - 2 * x_halo;
    const int nstrips = (U + W_out - 1) / W_out;
    const int rows = rb - ra;
    if (rows <= 0) return -1;
    const long long slots = (long long)sm_count * (occ_hint > 0 ? occ_hint : 1);
    const int fill = OC_MARCH_LAG + 2;
    int best_rs = rows; double best = -1.0;
    const char* env = getenv("OC_MARCH_RS");
    if (env && atoi(env) > 0) { best_rs = atoi(env) < rows ? atoi(env) : rows; }
    else for (int nseg = 1; nseg <= rows; ++nseg) {
        int rs = (rows + nseg - 1) / nseg;
        if (rs < 8 && nseg > 1) break;
        int ns = (rows + rs - 1) / rs;
        long long ctas = (long long)nstrips * ns * c.batch;
        long long waves = (ctas + slots - 1) / slots;
        double eff = (double)rows * nstrips * c.batch / ((double)waves * slots * (rs + fill));
        if (eff > best * 1.0001) { best = eff; best_rs = rs; }
    }
    // Chained launches (OcDep2: the tiles of step e+1 start as those of step e finish) do not care about whole waves.
    // What pays (measured at 2048^2, 4096^2, 8192^2, profiles/) is the tallest tile that still leaves ~15 % more tiles
    // than CTA slots, so that a freed slot always finds a tile whose neighbourhood is done.
    bool oversub = false;
    if (chained && !(env && atoi(env) > 0)) {
        const long long per_seg = (long long)nstrips * c.batch;            // tiles per segment index (batched cloths: every cloth)
        const int nseg = (int)((slots * 115 / 100 + per_seg - 1) / per_seg);
        const int rs = (rows + nseg - 1) / nseg;
        if (nseg >= 1 && rs >= 32) { best_rs = rs; oversub = true; }
    }
    if (linked) best_rs = fix_last_segment(rows, best_rs);
    pl->TW = WC; pl->S = 1; pl->x_halo = x_halo; pl->W_out = W_out; pl->nstrips = nstrips;
    pl->RS = best_rs; pl->nseg = (rows + best_rs - 1) / best_rs; pl->threads = WC / 2; pl->smem = smem2(WC);
    seg->rs = seg->rs_e = best_rs; seg->nstrips = nstrips; seg->nseg_all = pl->nseg; seg->n_extra = 0; seg->rev = 0;
    // Single-wave launch of one cloth: shorter segments for the two edge strips (OcSeg2).
    //   OC_MARCH2_EDGE = interior / edge rows per segment - 1 in percent (0 = off).
    const char* ee = getenv("OC_MARCH2_EDGE");
    const int edge = ee ? atoi(ee) : (exact ? OC_MARCH2_EDGE_EXACT : OC_MARCH2_EDGE_FAST);
    const long long tiles = (long long)nstrips * pl->nseg;
    if (!linked && !oversub && !(env && atoi(env) > 0) && edge > 0 && c.batch == 1 && nstrips >= 3 && tiles <= slots) {
        for (int e = edge; e >= 4; e -= 2) {                          // the largest ratio whose extra edge tiles fit the same wave
            OcSeg2 g = *seg;
            g.rs_e = (int)((100.0 * g.rs) / (100.0 + e) + 0.5);
            if (g.rs_e < 8) continue;
            oc_seg2_finish(g, rows);
            if (oc_seg2_tiles(g) <= slots) { *seg = g; break; }
        }
    }
    pl->nseg = seg->nseg_all;
    return 0;
}

cudaError_t oc_march2_launch(const OcConst& c, bool exact, int ra, int rb, int sm_count,
                             const float4* A, const float4* B, float4* C, cudaStream_t stream, int* n_launches, OcChain2* chain,
                             const OcPeer2* peer)
{
    *n_launches = 0;
    OcMarchPlan pl;
    const int WC = pick_wc(c.U);
    OcSeg2 seg;
    static const bool pdl = !(getenv("OC_PDL") && atoi(getenv("OC_PDL")) == 0);
    static const bool tile_deps = !(getenv("OC_TILE_DEPS") && atoi(getenv("OC_TILE_DEPS")) == 0);
    const bool chained = pdl && tile_deps && chain && chain->flags;
    if (oc_march2_plan(c, exact, chained, peer != nullptr, ra, rb, sm_count, g_occ2[exact ? 1 : 0][WC == 128], &pl, &seg) != 0) return cudaErrorInvalidValue;
    const void* fn = exact ? oc_march2_fn_exact(WC) : oc_march2_fn_fast(WC);
    if (!fn) return cudaErrorInvalidDeviceFunction;
    if (c.batch > 65535) return cudaErrorInvalidConfiguration;
    if (c.dbg & 16) {        // development: print the segmentation once per distinct row range
        static int last_ra = -1, last_rb = -1;
        if (last_ra != ra || last_rb != rb) {
            last_ra = ra; last_rb = rb;
            fprintf(stderr, "[oc] march2 plan rows [%d,%d): strips %d x segs %d + %d extra; rows/segment %d, edge strips %d; exact %d\n",
                    ra, rb, seg.nstrips, seg.nseg_all, seg.n_extra, seg.rs, seg.rs_e, (int)exact);
        }
    }
    dim3 grid(oc_seg2_tiles(seg), 1, c.batch), block(pl.threads, 1, 1);
    OcConst cc = c;
    int xh = pl.x_halo;
    // dependencies on the previous launch (OcDep2)
    OcDep2 dep = {};
    const bool can_flag = chain && chain->flags && (long long)oc_seg2_tiles(seg) * c.batch <= chain->cap;      // one flag per tile and cloth
    if (peer && !can_flag) return cudaErrorInvalidConfiguration;      // linked bands publish through the same epilogue
    if (can_flag) {
        dep.flags = chain->flags; dep.epoch = ++chain->epoch;
        // <= 4 segments per strip to poll; short tiles gain nothing (measured: 16-row tiles of a 1024^2 cloth lose 15 %)
        if (chain->valid && chain->kind == 0 && pdl && tile_deps && oc_dep2_chainable(seg, ra, rb, chain->pseg, chain->pra, chain->prb)) {
            dep.mode = 1; dep.pra = chain->pra; dep.prb = chain->prb; dep.pseg = chain->pseg;
        }
    }
    if (peer) { dep.peer = *peer; dep.peer.ra = ra; dep.peer.rb = rb; dep.peer.nstrips = seg.nstrips; seg.rev = peer->rev; }
    void* args[] = { &cc, (void*)&A, (void*)&B, (void*)&C, &ra, &rb, &seg, &xh, &dep };
    // Programmatic dependent launch: consecutive steps are kernel -> kernel edges on one stream; the next launch's CTAs
    // are placed while this one drains and wait (griddepcontrol.wait) before they touch the state.  OC_PDL=0 turns it off.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = pl.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e == cudaSuccess) *n_launches = 1;
    if (chain) {
        chain->valid = can_flag && e == cudaSuccess;
        chain->kind = 0;
        chain->pra = ra; chain->prb = rb; chain->pseg = seg;
    }
    return e;
}

// ------------------------------------------------------------------------------------------------
// twin-tile kernel (oc_twin.cuh), one substep per launch: one column per thread, two tiles per CTA
// ------------------------------------------------------------------------------------------------
extern "C" const void* oc_twin_fn_exact(int WC, int occ);
extern "C" const void* oc_twin_fn_fast(int WC, int occ);
extern "C" const void* oc_stream_fn_exact(int WC, int occ);
extern "C" const void* oc_stream_fn_fast(int WC, int occ);
extern "C" const void* oc_stream2_fn_exact(int WC, int occ);
extern "C" const void* oc_stream2_fn_fast(int WC, int occ);
// variant: 0 = oc_k_twin, 1 = oc_k_stream, 2 = oc_k_stream2 (same tiles, same launch protocol)
static int g_occT[3][2][2];      // [variant][exact][WC == 128]

static size_t smemT(int WC, bool exact, int variant)
{
    if (variant == 2) return exact ? sizeof(OcSmemS2<128, true>) : sizeof(OcSmemS2<128, false>);
    if (variant == 1) {
        if (WC == 64) return exact ? sizeof(OcSmemS<64, true>) : sizeof(OcSmemS<64, false>);
        return exact ? sizeof(OcSmemS<128, true>) : sizeof(OcSmemS<128, false>);
    }
    if (WC == 64) return exact ? sizeof(OcSmemT<64, true>) : sizeof(OcSmemT<64, false>);
    return exact ? sizeof(OcSmemT<128, true>) : sizeof(OcSmemT<128, false>);
}
static int pick_wct(int nx, int variant)
{
    if (variant == 2) return 128;          // two columns per thread: 64 threads, the smallest CTA the dependency polling allows
    const char* env = getenv(variant == 1 ? "OC_STREAM_WC" : "OC_TWIN_WC");
    if (env && (atoi(env) == 64 || atoi(env) == 128)) return atoi(env);
    return nx <= 64 ? 64 : 128;
}
static const void* twin_fn_v(int WC, bool exact, int variant, int v)
{
    if (variant == 2) return exact ? oc_stream2_fn_exact(WC, v) : oc_stream2_fn_fast(WC, v);
    if (variant == 1) return exact ? oc_stream_fn_exact(WC, v) : oc_stream_fn_fast(WC, v);
    return exact ? oc_twin_fn_exact(WC, v) : oc_twin_fn_fast(WC, v);
}
// development: OC_TWIN_OCC / OC_STREAM_OCC = resident CTAs per SM the kernel variant is compiled for (see the *_inst.cu)
static const void* twin_fn(int WC, bool exact, int variant)
{
    const char* env = getenv(variant == 2 ? "OC_STREAM2_OCC" : (variant == 1 ? "OC_STREAM_OCC" : "OC_TWIN_OCC"));
    const int v = env ? atoi(env) : 0;
    const void* fn = v > 0 ? twin_fn_v(WC, exact, variant, v) : nullptr;
    return fn ? fn : twin_fn_v(WC, exact, variant, 0);
}

int oc_twin_configure(int device)
{
    (void)device;
    for (int variant = 0; variant < 3; ++variant)
        for (int e = 0; e < 2; ++e)
            for (int w = 0; w < 2; ++w) {
                const int WC = w ? 128 : 64;
                g_occT[variant][e][w] = 0;
                if (variant == 2 && WC != 128) continue;
                const void* fn = twin_fn(WC, e != 0, variant);
                cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemT(WC, e != 0, variant));
                if (err != cudaSuccess) return (int)err;
                int occ = 0;
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, variant == 2 ? WC / 2 : WC, smemT(WC, e != 0, variant));
                if (err != cudaSuccess) return (int)err;
                g_occT[variant][e][w] = occ;
            }
    return 0;
}

int oc_twin_nstrips(int nx)
{
    const int WC = pick_wct(nx, 0);
    const int W_out = WC - 2 * ((nx <= WC) ? 0 : 2);
    return (nx + W_out - 1) / W_out;
}

// Segment height for `rows` rows such that the number of segments is EVEN (a CTA takes segments 2k and 2k+1) and the
// last segment is (nearly) as tall as the others: the twins of a CTA run in lock step, so the steady loop of the last
// pair only covers the rows its shorter tile has (measured: a last segment of 20 rows against 52 cost 3x at 2048^2).
// Among the heights within 1/6 of `want`, the smallest shortfall of the last segment wins, then the nearest height;
// with `linked`, the last segment keeps at least two rows (fix_last_segment).
static int twin_even_rs(int rows, int want, bool linked)
{
    if (want < 1) want = 1;
    if (want > rows) want = rows;
    const int span = want / 6 + 1;
    int best = 0, best_short = 1 << 30, best_d = 1 << 30;
    for (int rs = want - span < 1 ? 1 : want - span; rs <= want + span && rs <= rows; ++rs) {
        const int n = (rows + rs - 1) / rs;
        if (n % 2 != 0) continue;
        if (linked && (rows - 1) % rs + 1 < 2) continue;
        const int shortfall = n * rs - rows, d = rs > want ? rs - want : want - rs;
        if (shortfall < best_short || (shortfall == best_short && d < best_d)) { best = rs; best_short = shortfall; best_d = d; }
    }
    if (best) return best;
    for (int d = 0; d <= rows; ++d)          // anything even
        for (int sgn = 0; sgn < 2; ++sgn) {
            const int rs = sgn ? want - d : want + d;
            if (rs < 1 || rs > rows) continue;
            const int n = (rows + rs - 1) / rs;
            if (n % 2 != 0) continue;
            if (linked && (rows - 1) % rs + 1 < 2) continue;
            return rs;
        }
    return 0;      // rows == 1: no even cut
}

int oc_twin_plan(const OcConst& c, bool exact, bool chained, bool linked, int ra, int rb, int sm_count, int occ_hint, OcMarchPlan* pl, OcSeg2* seg, OcTwinMap* map, int variant)
{
    const int U = c.U;
    const int WC = pick_wct(U, variant);
    const int x_halo = (U <= WC) ? 0 : 2;
    const int W_out = WC - 2 * x_halo;
    const int nstrips = (U + W_out - 1) / W_out;
    const int rows = rb - ra;
    if (rows <= 0) return -1;
    const long long slots = (long long)sm_count * (occ_hint > 0 ? occ_hint : 1);      // CTA slots; a CTA holds two tiles
    const int fill = OC_MARCH_LAG + 2;
    // Batches of an even number of cloths: the twins are the same tile of two cloths (no constraint on the segments);
    // otherwise two segments of one strip.
    const char* pe = getenv("OC_TWIN_PAIR");
    map->pair_cloths = (c.batch >= 2 && c.batch % 2 == 0 && !linked && !(pe && atoi(pe) == 0)) ? 1 : 0;
    const long long zs = map->pair_cloths ? c.batch / 2 : c.batch;
    const int per_cta = map->pair_cloths ? 1 : 2;      // segments of one strip per CTA
    int best_rs = 0; double best = -1.0;
    const char* env = getenv("OC_MARCH_RS");
    const bool forced = env && atoi(env) > 0;
    if (forced) best_rs = atoi(env) < rows ? atoi(env) : rows;
    else for (int nseg = per_cta; nseg <= rows; nseg += per_cta) {
        int rs = (rows + nseg - 1) / nseg;
        if (rs < 8 && nseg > per_cta) break;
        if (per_cta == 2) { rs = twin_even_rs(rows, rs, linked); if (rs == 0) continue; }
        const int ns = (rows + rs - 1) / rs;
        const long long ctas = (long long)nstrips * (ns / per_cta) * zs;
        const long long waves = (ctas + slots - 1) / slots;
        const double eff = (double)rows * nstrips * c.batch / ((double)waves * slots * 2 * (rs + fill));
        if (eff > best * 1.0001) { best = eff; best_rs = rs; }
    }
    // chained launches: the tallest tiles that still leave ~15 % more CTAs than CTA slots (see oc_march2_plan)
    if (chained && !forced) {
        const long long per_seg = (long long)nstrips * zs;
        int nseg = (int)((slots * 115 / 100 + per_seg - 1) / per_seg) * per_cta;
        if (nseg < per_cta) nseg = per_cta;
        const int rs = (rows + nseg - 1) / nseg;
        if (rs >= 32) best_rs = rs;
    }
    if (per_cta == 2) {
        best_rs = twin_even_rs(rows, best_rs > 0 ? best_rs : rows / 2, linked);
        if (best_rs == 0) return -1;
    } else {
        if (best_rs <= 0) best_rs = rows;
        if (linked) best_rs = fix_last_segment(rows, best_rs);
    }
    pl->TW = WC; pl->S = 1; pl->x_halo = x_halo; pl->W_out = W_out; pl->nstrips = nstrips;
    pl->RS = best_rs; pl->nseg = (rows + best_rs - 1) / best_rs; pl->threads = variant == 2 ? WC / 2 : WC; pl->smem = smemT(WC, exact, variant);
    seg->rs = seg->rs_e = best_rs; seg->nstrips = nstrips; seg->nseg_all = pl->nseg; seg->n_extra = 0; seg->rev = 0;
    return 0;
}

cudaError_t oc_twin_launch(const OcConst& c, bool exact, int ra, int rb, int sm_count,
                           const float4* A, const float4* B, float4* C, cudaStream_t stream, int* n_launches, OcChain2* chain,
                           const OcPeer2* peer, int variant)
{
    *n_launches = 0;
    OcMarchPlan pl;
    const int WC = pick_wct(c.U, variant);
    OcSeg2 seg;
    OcTwinMap map;
    static const bool pdl = !(getenv("OC_PDL") && atoi(getenv("OC_PDL")) == 0);
    static const bool tile_deps = !(getenv("OC_TILE_DEPS") && atoi(getenv("OC_TILE_DEPS")) == 0);
    const bool chained = pdl && tile_deps && chain && chain->flags;
    if (oc_twin_plan(c, exact, chained, peer != nullptr, ra, rb, sm_count, g_occT[variant][exact ? 1 : 0][WC == 128], &pl, &seg, &map, variant) != 0) return cudaErrorInvalidValue;
    const void* fn = twin_fn(WC, exact, variant);
    if (!fn) return cudaErrorInvalidDeviceFunction;
    if (c.batch > 65535) return cudaErrorInvalidConfiguration;
    if (c.dbg & 16) {        // development: print the segmentation once per distinct row range
        static int last_ra = -1, last_rb = -1;
        if (last_ra != ra || last_rb != rb) {
            last_ra = ra; last_rb = rb;
            fprintf(stderr, "[oc] %s plan rows [%d,%d): strips %d x segs %d; rows/segment %d; pair_cloths %d; CTAs/SM %d; exact %d\n", variant == 2 ? "stream2" : (variant ? "stream" : "twin"),
                    ra, rb, seg.nstrips, seg.nseg_all, seg.rs, map.pair_cloths, g_occT[variant][exact ? 1 : 0][WC == 128], (int)exact);
        }
    }
    const int nk = map.pair_cloths ? seg.nseg_all : seg.nseg_all / 2;
    dim3 grid(seg.nstrips * nk, 1, map.pair_cloths ? c.batch / 2 : c.batch), block(pl.threads, 1, 1);
    OcConst cc = c;
    int xh = pl.x_halo;
    // dependencies on the previous launch (OcDep2), one flag per tile and cloth as for oc_k_march2
    OcDep2 dep = {};
    const bool can_flag = chain && chain->flags && (long long)oc_seg2_tiles(seg) * c.batch <= chain->cap;
    if (peer && !can_flag) return cudaErrorInvalidConfiguration;
    if (can_flag) {
        dep.flags = chain->flags; dep.epoch = ++chain->epoch;
        if (chain->valid && chain->kind == 1 + variant && pdl && tile_deps && oc_dep2_chainable(seg, ra, rb, chain->pseg, chain->pra, chain->prb)) {
            dep.mode = 1; dep.pra = chain->pra; dep.prb = chain->prb; dep.pseg = chain->pseg;
        }
    }
    if (peer) { dep.peer = *peer; dep.peer.ra = ra; dep.peer.rb = rb; dep.peer.nstrips = seg.nstrips; seg.rev = peer->rev; }
    void* args[] = { &cc, (void*)&A, (void*)&B, (void*)&C, &ra, &rb, &seg, &xh, &map, &dep };
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = pl.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e == cudaSuccess) *n_launches = 1;
    if (chain) {
        chain->valid = can_flag && e == cudaSuccess;
        chain->kind = 1 + variant;
        chain->pra = ra; chain->prb = rb; chain->pseg = seg;
    }
    return e;
}
