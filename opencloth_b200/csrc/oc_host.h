// oc_host.h — host-only logic shared by the C-ABI (oc_api.cu) and the CPU kernel emulator
// (tests/emu/oc_emu.cu): derived constants, rest-length tables, band storage geometry and the
// splitting of n substeps into launches.  Pure functions, no CUDA calls.
//
// All fp32 arithmetic here must round exactly like the reference's set-up code
// (/root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp "V:"), so every translation unit that
// includes this file is compiled with -ffp-contract=off.
#pragma once
#include "../../include/opencloth.h"
#include "oc_core.cuh"
#include <vector>
#include <math.h>
#include <string.h>
#include <stdlib.h>

// ellipsoid = translate(0,2,0) * rotate(45 deg about x) * scale(1,1,0.5), inverse = glm::inverse(ellipsoid)
// (V:324-327).  The reference builds them with GLM 0.9.0.0 at start-up; these are the resulting fp32
// bit patterns (column-major), pinned against the verbatim reference build by tests/test_oracle.py.
static const float oc_k_ellipsoid[16] = {
    0x1.0p+0f, 0.0f, 0.0f, 0.0f,
    0.0f, 0x1.6a09e6p-1f, 0x1.6a09e6p-1f, 0.0f,
    0.0f, -0x1.6a09e6p-2f, 0x1.6a09e6p-2f, 0.0f,
    0.0f, 0x1.0p+1f, 0.0f, 0x1.0p+0f };
static const float oc_k_inv_ellipsoid[16] = {
    0x1.0p+0f, -0.0f, 0.0f, -0.0f,
    -0.0f, 0x1.6a09e8p-1f, -0x1.6a09e8p+0f, 0.0f,
    0.0f, 0x1.6a09e8p-1f, 0x1.6a09e8p+0f, -0.0f,
    -0.0f, -0x1.6a09e8p+0f, 0x1.6a09e8p+1f, 0x1.0p+0f };

static inline void oc_host_default_params(oc_params* p, int nx, int ny)
{
    memset(p, 0, sizeof(*p));
    p->nx = nx; p->ny = ny; p->batch = 1;
    p->row_begin = 0; p->row_end = 0; p->halo_rows = 0; p->device = -1;
    p->fullsize = 4.0f;                                   // V:61
    p->substeps_per_launch = 0; p->exact = 1; p->kernel = OC_KERNEL_AUTO;
    p->ks_struct = 50.75f; p->kd_struct = -0.25f;         // V:98
    p->ks_shear  = 50.75f; p->kd_shear  = -0.25f;         // V:99
    p->ks_bend   = 50.95f; p->kd_bend   = -0.25f;         // V:100
    p->damping = -0.0125f;                                // V:97
    p->gravity[0] = 0.0f; p->gravity[1] = -0.00981f; p->gravity[2] = 0.0f;   // V:101
    p->mass = 1.0f;                                       // V:102
    p->dt = 1 / 60.0f;                                    // V:104
    memcpy(p->ellipsoid, oc_k_ellipsoid, sizeof(oc_k_ellipsoid));
    memcpy(p->inv_ellipsoid, oc_k_inv_ellipsoid, sizeof(oc_k_inv_ellipsoid));
    p->center[0] = p->center[1] = p->center[2] = 0.0f;    // V:129
    p->radius = 1.0f;                                     // V:130
    p->integrator = OC_INTEGRATOR_VERLET; p->provot = 0;  // V:561 leaves the Provot pass commented out
}
// the globals of the sibling demos (E: OpenCloth_ExplicitEuler main.cpp:97-102, S: OpenCloth_SemiImplicit main.cpp:80-85)
static inline void oc_host_default_params_for(oc_params* p, int nx, int ny, int integrator)
{
    oc_host_default_params(p, nx, ny);
    p->integrator = integrator;
    if (integrator == OC_INTEGRATOR_EULER)         { p->ks_struct = 0.5f;  p->ks_shear = 0.5f;  p->ks_bend = 0.85f; p->mass = 0.5f; p->provot = 1; }
    if (integrator == OC_INTEGRATOR_SEMI_IMPLICIT) { p->ks_struct = 0.75f; p->ks_shear = 0.75f; p->ks_bend = 0.95f; p->mass = 0.5f; p->provot = 1; }
}

// Everything derived from the run-time scalars, with the reference's fp32 operations.
static inline void oc_host_derive_scalars(const oc_params& p, OcConst& k)
{
    k.dt = p.dt;
    k.inv_dt = 1.0f / p.dt;
    k.one = 1.0f;
    k.dt_bf = (p.dt >= 0x1.0p-20f && p.dt <= 0x1.0p+20f) ? 1 : 0;
    { const char* e = getenv("OC_DEBUG"); k.dbg = e ? atoi(e) : 0; }
    k.dt2m = (p.dt * p.dt) / p.mass;                                  // V:429
    k.dtm = p.dt / p.mass;                                            // E:470, S:465
    k.integ = p.integrator;
    k.damping = p.damping;
    for (int a = 0; a < 3; ++a)                                       // V:452, V:456; the Euler demos add gravity without the mass (E:441)
        k.f0[a] = p.integrator == OC_INTEGRATOR_VERLET ? 0.0f + p.gravity[a] * p.mass : 0.0f + p.gravity[a];
    k.nks_struct = -p.ks_struct; k.kd_struct = p.kd_struct;
    k.nks_shear  = -p.ks_shear;  k.kd_shear  = p.kd_shear;
    k.nks_bend   = -p.ks_bend;   k.kd_bend   = p.kd_bend;
    for (int r = 0; r < 3; ++r)
        for (int col = 0; col < 4; ++col) k.im[r][col] = p.inv_ellipsoid[col * 4 + r];
    for (int col = 0; col < 4; ++col) { k.imxy[col][0] = k.im[0][col]; k.imxy[col][1] = k.im[1][col]; }
    for (int a = 0; a < 3; ++a) k.center[a] = p.center[a];
    k.radius = p.radius;
    for (int a = 0; a < 3; ++a) {                                     // V:520-527
        float tx = p.ellipsoid[0 * 4 + a], ty = p.ellipsoid[1 * 4 + a], tz = p.ellipsoid[2 * 4 + a];
        float d = tx * tx + ty * ty + tz * tz;
        k.tinv[a][0] = tx / d; k.tinv[a][1] = ty / d; k.tinv[a][2] = tz / d;
    }
    // Bounding sphere of the collider, from the matrix the test really uses (V:511): delta0 = A X + t - c with A, t the
    // linear part and translation of inverse_ellipsoid.  |delta0| >= sigma_min(A) |X - Xc|, Xc = A^-1 (c - t), and
    // sigma_min(A) >= 1 / ||A^-1||_F: a particle with |X - Xc| > 1.05 ||A^-1||_F has |delta0| > 1.05 and is outside
    // whatever the fp32 rounding of V:511-514 (its error is ~1e-6 of the squared magnitudes below, bounded by 1e-2).
    // Anything unusual (singular A, large offsets, non-finite entries) switches the shortcut off.
    {
        double A[3][3], t[3];
        for (int r = 0; r < 3; ++r) { for (int col = 0; col < 3; ++col) A[r][col] = p.inv_ellipsoid[col * 4 + r]; t[r] = p.inv_ellipsoid[12 + r]; }
        const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        k.bs_c[0] = k.bs_c[1] = k.bs_c[2] = 0.0f; k.bs_r2 = INFINITY;
        double na = 0.0; for (int r = 0; r < 3; ++r) for (int col = 0; col < 3; ++col) na += A[r][col] * A[r][col];
        na = sqrt(na);
        if (det == det && fabs(det) > 1e-9 * na * na * na && na > 0.0) {
            double I[3][3];
            I[0][0] =  (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det; I[0][1] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]) / det; I[0][2] =  (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
            I[1][0] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]) / det; I[1][1] =  (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det; I[1][2] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]) / det;
            I[2][0] =  (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det; I[2][1] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]) / det; I[2][2] =  (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
            double ni = 0.0, xc[3], nt = 0.0, nc = 0.0, nx = 0.0;
            for (int r = 0; r < 3; ++r) for (int col = 0; col < 3; ++col) ni += I[r][col] * I[r][col];
            ni = sqrt(ni);
            for (int r = 0; r < 3; ++r) {
                xc[r] = I[r][0] * (p.center[0] - t[0]) + I[r][1] * (p.center[1] - t[1]) + I[r][2] * (p.center[2] - t[2]);
                nt += t[r] * t[r]; nc += (double)p.center[r] * p.center[r]; nx += xc[r] * xc[r];
            }
            const double rb = 1.05 * ni;
            const double mag = na * (sqrt(nx) + rb) + sqrt(nt) + sqrt(nc);       // magnitude of the terms of V:511-513 near the sphere
            if (rb == rb && mag == mag && mag < 100.0 && rb < 1e6) {
                for (int r = 0; r < 3; ++r) k.bs_c[r] = (float)xc[r];
                k.bs_r2 = (float)(rb * rb * 1.001 + 1e-6 * (nx + rb * rb));          // rounding of the centre and of the distance test itself
            }
        }
    }
}

static inline float oc_host_rest(float ax, float az, float bx, float bz)
{   // AddSpring V:141-142: deltaP = X[a]-X[b]; sqrt(dot(deltaP,deltaP)); the sheet is flat, dy = 0
    float dx = ax - bx, dy = 0.0f, dz = az - bz;
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

// Initial-sheet coordinates (V:254-260) and the rest-length tables of the implicit spring net
// (V:286-320 with V:141-142).  Layout of t: xs[U] zs[V] rh1[U] rh2[U] dx2[U] rv1[V] rv2[V] dz2[V]
struct OcHostTables {
    std::vector<float> t;
    size_t xs, zs, rh1, rh2, dx2, rv1, rv2, dz2;      // offsets into t
};
static inline void oc_host_build_tables(int U, int V, float fullsize, OcHostTables& T)
{
    T.t.assign((size_t)4 * U + 4 * V, 0.0f);
    T.xs = 0; T.zs = T.xs + U; T.rh1 = T.zs + V; T.rh2 = T.rh1 + U; T.dx2 = T.rh2 + U;
    T.rv1 = T.dx2 + U; T.rv2 = T.rv1 + V; T.dz2 = T.rv2 + V;
    float* xs = &T.t[T.xs]; float* zs = &T.t[T.zs];
    float halfsize = fullsize / 2.0f;                                                  // V:62
    for (int i = 0; i < U; ++i) xs[i] = (((float)i / (U - 1)) * 2 - 1) * halfsize;     // V:256
    for (int j = 0; j < V; ++j) zs[j] = (((float)j / (V - 1)) * fullsize);             // V:256
    for (int i = 0; i < U; ++i) {
        T.t[T.rh1 + i] = (i + 1 < U) ? oc_host_rest(xs[i], 0.0f, xs[i + 1], 0.0f) : 0.0f;
        T.t[T.rh2 + i] = (i + 2 < U) ? oc_host_rest(xs[i], 0.0f, xs[i + 2], 0.0f) : 0.0f;
        float d = (i + 1 < U) ? xs[i] - xs[i + 1] : 0.0f;
        T.t[T.dx2 + i] = d * d;
    }
    for (int j = 0; j < V; ++j) {
        T.t[T.rv1 + j] = (j + 1 < V) ? oc_host_rest(0.0f, zs[j], 0.0f, zs[j + 1]) : 0.0f;
        T.t[T.rv2 + j] = (j + 2 < V) ? oc_host_rest(0.0f, zs[j], 0.0f, zs[j + 2]) : 0.0f;
        float d = (j + 1 < V) ? zs[j] - zs[j + 1] : 0.0f;
        T.t[T.dz2 + j] = d * d;
    }
}
static inline void oc_host_bind_tables(OcConst& k, const float* base, const OcHostTables& T)
{
    k.rh1 = base + T.rh1; k.rh2 = base + T.rh2; k.dx2 = base + T.dx2;
    k.rv1 = base + T.rv1; k.rv2 = base + T.rv2; k.dz2 = base + T.dz2;
}

// ------------------------------------------------------------------------------------------------
// Storage geometry and launch sequencing
// ------------------------------------------------------------------------------------------------
struct OcSeq {
    // fixed
    int row_begin, row_end;   // rows owned
    int V;
    bool band;                // owns a strict sub-range of the rows
    bool linked;              // band whose halo rows are kept current by its neighbours (peer stores every substep): no shrink
    bool xv;                  // the state is (X, V) — Euler integrators: a step writes both free buffers, ia <- X, ib <- V
    int kmax;                 // band: substeps between halo exchanges (halo_rows / 2)
    // running
    int fresh;                // band: substeps taken since the halo rows were last current
    int ia, ib;               // buffer holding X(t), X(t-1)
};

// Normalises p (row band, halo) and fills the storage part of k.  Returns false on bad geometry.
static inline void oc_host_geometry(oc_params& p, OcConst& k, OcSeq& q)
{
    const int U = p.nx, V = p.ny;
    int rb = p.row_begin, re = p.row_end;
    if (rb == 0 && re == 0) re = V;
    p.row_begin = rb; p.row_end = re;
    q.band = (rb > 0 || re < V);
    int halo = q.band ? p.halo_rows : 0;
    p.halo_rows = halo;
    int lo = rb - halo; if (lo < 0) lo = 0;
    int hi = re + halo; if (hi > V) hi = V;
    k.U = U; k.V = V; k.row_lo = lo; k.srows = hi - lo; k.batch = p.batch;
    k.cloth_stride = (long long)k.srows * U;
    q.row_begin = rb; q.row_end = re; q.V = V;
    q.kmax = halo / 2; q.fresh = 0; q.ia = 0; q.ib = 1; q.linked = false;
    q.xv = p.integrator != OC_INTEGRATOR_VERLET;
}

// Stage counts (substeps per launch) the marching kernel is compiled for: 1, 2, 4, 8.
// Largest one not above `want`.
static inline int oc_host_pick_stages(int want)
{
    if (want >= 8) return 8;
    if (want >= 4) return 4;
    if (want >= 2) return 2;
    return 1;
}

struct OcLaunch {
    int S;                    // substeps in this launch
    int ra, rb;               // rows whose X(t+S) must be produced
    int src_a, src_b;         // buffers holding X(t), X(t-1)
    int dst, dst_prev;        // buffers receiving X(t+S) and (S >= 2) X(t+S-1)
};

// Next launch of a request for n more substeps; k = substeps per launch the kernel can take.
// Updates the sequence state (buffer rotation, halo shrink counter) and n.
// A band launch at shrink counter f recomputes rows [row_begin - 2(kmax-f-S), row_end + 2(kmax-f-S)):
// after the exchange every stored row is current; each substep loses 2 rows either side.
static inline void oc_host_next_launch(OcSeq& q, int& n, int k, OcLaunch& L)
{
    int S = n < k ? n : k;
    L.S = S;
    L.ra = q.row_begin; L.rb = q.row_end;
    if (q.band && !q.linked) {
        int grow = 2 * (q.kmax - q.fresh - S);
        L.ra -= grow; L.rb += grow;
        if (L.ra < 0) L.ra = 0;
        if (L.rb > q.V) L.rb = q.V;
    }
    int f[2], m = 0;
    for (int b = 0; b < 4; ++b) if (b != q.ia && b != q.ib) f[m++] = b;
    L.src_a = q.ia; L.src_b = q.ib; L.dst = f[0]; L.dst_prev = f[1];
    if (S == 1 && !q.xv) { q.ib = q.ia; q.ia = L.dst; }
    else                 { q.ia = L.dst; q.ib = L.dst_prev; }
    if (q.band && !q.linked) q.fresh += S;
    n -= S;
}

// Rows exchanged with a band neighbour.  side 0 = towards row 0, 1 = towards row ny-1.
// send: the first / last halo_rows OWNED rows;  recv: the halo rows beyond the owned range.
// Returns false when that side is the cloth boundary (nothing to exchange).
static inline bool oc_host_halo_rows(const oc_params& p, bool band, int side, bool send, int* r0, int* rows)
{
    *r0 = 0; *rows = 0;
    if (!band) return false;
    const int H = p.halo_rows;
    if (side == 0) {
        if (p.row_begin == 0) return false;
        *r0 = send ? p.row_begin : p.row_begin - H;
    } else {
        if (p.row_end == p.ny) return false;
        *r0 = send ? p.row_end - H : p.row_end;
    }
    *rows = H;
    return true;
}
