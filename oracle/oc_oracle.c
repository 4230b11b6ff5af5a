/* oracle/oc_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's Verlet cloth step
 *   StepPhysics = ComputeForces -> IntegrateVerlet -> EllipsoidCollision
 *   (/root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp:557-562, "V:" below)
 * in *gather* form: one pass per particle, no spring list, no scatter.  It exists because the
 * verbatim reference (oracle/_ref/libocref.so, built by oracle/build_ref.sh from the reference's
 * own source text) needs a 24 B/spring edge list (9.7 GB at 8192^2) and is single threaded.
 *
 * PARITY STATUS: pinned.  tests/test_oracle.py checks this restatement BITWISE against
 *   (a) the verbatim reference build, when oracle/_ref/libocref.so is present, on several grids
 *       for thousands of steps (collisions included), and
 *   (b) the committed golden vectors in tests/golden/ that were produced by the verbatim build
 *       (tests/golden/make_golden.py).
 * The reference itself has no tests or golden vectors.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (libopencloth_b200.so) never does and has no CPU fallback.
 *
 * All arithmetic is IEEE binary32, compiled with -ffp-contract=off, evaluation order matched to
 * the reference and to the GLM 0.9.0.0 inlines it calls (dep/glm/glm/core/func_geometric.inl:42-51
 * length, :139-149 dot, :220-230 normalize; func_exponential.inl:308-317 inversesqrt;
 * type_mat4x4.inl:561-572 mat4*vec4; type_vec3.inl operators are component-wise).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

typedef struct oco_params {
    int   nx, ny;               /* particles per row / per column (reference: numX+1, numY+1; V:59) */
    float fullsize;             /* V:61 */
    float ks_struct, kd_struct; /* V:98 */
    float ks_shear,  kd_shear;  /* V:99 */
    float ks_bend,   kd_bend;   /* V:100 */
    float damping;              /* V:97 DEFAULT_DAMPING */
    float gravity[3];           /* V:101 */
    float mass;                 /* V:102 */
    float dt;                   /* V:104 timeStep */
    float ellipsoid[16];        /* V:324-326, column-major as GLM stores it */
    float inv_ellipsoid[16];    /* V:327 */
    float center[3];            /* V:129 */
    float radius;               /* V:130 */
    int   integrator;           /* 0 Verlet (V:), 1 explicit Euler (E:), 2 semi-implicit / symplectic Euler (S:); see below */
    int   provot;               /* 1: ApplyProvotDynamicInverse after EllipsoidCollision (V:486-508, disabled at V:561;
                                   E:554-577 / S:437-462, enabled at E:639 / S:531) */
} oco_params;
/* The two explicit-integrator siblings of the Verlet demo, restated from
 *   E: = /root/reference/OpenCloth_ExplicitEuler/OpenCloth_ExplicitEuler/main.cpp   (StepPhysics E:626-641)
 *   S: = /root/reference/OpenCloth_SemiImplicit/OpenCloth_SemiImplicit/main.cpp     (StepPhysics S:527-532)
 * They keep X and V (here: the `xl` array holds V), use the same spring net and the same spring force on the stored
 * velocities (E:434-466 / S:402-436; gravity is added WITHOUT the mass factor, E:441), integrate with
 * IntegrateEuler (E:469-482: V += F*(dt/mass); X += dt*oldV) or IntegrateSemiImplicit (S:464-477: X += dt*newV),
 * zero V on collider contact (E:599 / S:499), and apply the Provot pass to V.  Pinned against the verbatim builds
 * oracle/_ref/libocref_euler.so / libocref_semi.so by tests/test_oracle.py. */
#define OCO_VERLET 0
#define OCO_EULER  1
#define OCO_SEMI   2

typedef struct oco_cloth {
    oco_params p;
    size_t n;
    float *x, *xl;              /* current X, X_last : n*3 floats each */
    float *x2, *xl2;            /* back buffers */
    float *xs, *zs;             /* initial-sheet coordinates per column / per row (V:256) */
    float *rh1, *rh2;           /* rest length of spring (i,j)-(i+1,j), (i,j)-(i+2,j)  [depends on i only] */
    float *rv1, *rv2;           /* rest length of spring (i,j)-(i,j+1), (i,j)-(i,j+2)  [depends on j only] */
    float *dx2, *dz2;           /* fl(dx_i*dx_i), fl(dz_j*dz_j) for the shear rest length */
    float tinv[3][3];           /* the three "transformInv" vectors of V:520-527 after the /= */
    unsigned char* pin;         /* NULL: the reference's literals (indices 0 and numX); else one byte per particle (oco_set_pins) */
} oco_cloth;
/* the reference's pin test `i==0 || i==numX` (V:455, V:479-482, V:498-501), generalised to a caller-supplied set for
 * the product's oc_set_pins extension (the verbatim reference can only check the default set) */
static inline int is_pinned(const oco_cloth* c, size_t idx)
{
    return c->pin ? c->pin[idx] : (idx == 0 || idx == (size_t)(c->p.nx - 1));
}

/* Default ellipsoid = translate(0,2,0) * rotate(45 deg, x) * scale(1,1,0.5) and its glm::inverse
 * (V:324-327).  Bit patterns taken from the verbatim build (tests/test_oracle.py re-checks them). */
static const float k_ellipsoid[16] = {
    0x1.0p+0f, 0.0f, 0.0f, 0.0f,
    0.0f, 0x1.6a09e6p-1f, 0x1.6a09e6p-1f, 0.0f,
    0.0f, -0x1.6a09e6p-2f, 0x1.6a09e6p-2f, 0.0f,
    0.0f, 0x1.0p+1f, 0.0f, 0x1.0p+0f };
static const float k_inv_ellipsoid[16] = {
    0x1.0p+0f, -0.0f, 0.0f, -0.0f,
    -0.0f, 0x1.6a09e8p-1f, -0x1.6a09e8p+0f, 0.0f,
    0.0f, 0x1.6a09e8p-1f, 0x1.6a09e8p+0f, -0.0f,
    -0.0f, -0x1.6a09e8p+0f, 0x1.6a09e8p+1f, 0x1.0p+0f };

void oco_default_params(oco_params* p, int nx, int ny)
{
    memset(p, 0, sizeof(*p));
    p->nx = nx; p->ny = ny;
    p->fullsize = 4.0f;
    p->ks_struct = 50.75f; p->kd_struct = -0.25f;
    p->ks_shear  = 50.75f; p->kd_shear  = -0.25f;
    p->ks_bend   = 50.95f; p->kd_bend   = -0.25f;
    p->damping = -0.0125f;
    p->gravity[0] = 0.0f; p->gravity[1] = -0.00981f; p->gravity[2] = 0.0f;
    p->mass = 1.0f;
    p->dt = 1 / 60.0f;
    memcpy(p->ellipsoid, k_ellipsoid, sizeof(k_ellipsoid));
    memcpy(p->inv_ellipsoid, k_inv_ellipsoid, sizeof(k_inv_ellipsoid));
    p->center[0] = p->center[1] = p->center[2] = 0.0f;
    p->radius = 1.0f;
    p->integrator = OCO_VERLET; p->provot = 0;
}
/* the globals of the sibling demos: E:97-102 / S:80-85 */
void oco_default_params_for(oco_params* p, int nx, int ny, int integrator)
{
    oco_default_params(p, nx, ny);
    p->integrator = integrator;
    if (integrator == OCO_EULER) { p->ks_struct = 0.5f;  p->ks_shear = 0.5f;  p->ks_bend = 0.85f; p->mass = 0.5f; p->provot = 1; }
    if (integrator == OCO_SEMI)  { p->ks_struct = 0.75f; p->ks_shear = 0.75f; p->ks_bend = 0.95f; p->mass = 0.5f; p->provot = 1; }
}

static float rest_from(float ax, float az, float bx, float bz)
{   /* AddSpring, V:141-142: deltaP = X[a]-X[b]; sqrt(dot(deltaP,deltaP)); y's are equal -> 0 */
    float dx = ax - bx, dy = 0.0f, dz = az - bz;
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

static void derive(oco_cloth* c)
{
    const oco_params* p = &c->p;
    int u = p->nx, v = p->ny;
    float halfsize = p->fullsize / 2.0f;                                    /* V:62 */
    for (int i = 0; i < u; ++i) c->xs[i] = (((float)i / (u - 1)) * 2 - 1) * halfsize;   /* V:256 */
    for (int j = 0; j < v; ++j) c->zs[j] = (((float)j / (v - 1)) * p->fullsize);        /* V:256 */
    for (int i = 0; i < u; ++i) {
        c->rh1[i] = (i + 1 < u) ? rest_from(c->xs[i], 0.0f, c->xs[i + 1], 0.0f) : 0.0f;
        c->rh2[i] = (i + 2 < u) ? rest_from(c->xs[i], 0.0f, c->xs[i + 2], 0.0f) : 0.0f;
        float d = (i + 1 < u) ? c->xs[i] - c->xs[i + 1] : 0.0f;
        c->dx2[i] = d * d;
    }
    for (int j = 0; j < v; ++j) {
        c->rv1[j] = (j + 1 < v) ? rest_from(0.0f, c->zs[j], 0.0f, c->zs[j + 1]) : 0.0f;
        c->rv2[j] = (j + 2 < v) ? rest_from(0.0f, c->zs[j], 0.0f, c->zs[j + 2]) : 0.0f;
        float d = (j + 1 < v) ? c->zs[j] - c->zs[j + 1] : 0.0f;
        c->dz2[j] = d * d;
    }
    /* V:520-527: transformInv = (E[0].c, E[1].c, E[2].c); transformInv /= dot(transformInv, transformInv) */
    for (int k = 0; k < 3; ++k) {
        float tx = p->ellipsoid[0 * 4 + k], ty = p->ellipsoid[1 * 4 + k], tz = p->ellipsoid[2 * 4 + k];
        float d = tx * tx + ty * ty + tz * tz;
        c->tinv[k][0] = tx / d; c->tinv[k][1] = ty / d; c->tinv[k][2] = tz / d;
    }
}

oco_cloth* oco_create(const oco_params* p)
{
    if (!p || p->nx < 3 || p->ny < 3) return NULL;
    oco_cloth* c = (oco_cloth*)calloc(1, sizeof(oco_cloth));
    c->p = *p;
    c->n = (size_t)p->nx * (size_t)p->ny;
    size_t b = c->n * 3 * sizeof(float);
    c->x = (float*)malloc(b); c->xl = (float*)malloc(b); c->x2 = (float*)malloc(b); c->xl2 = (float*)malloc(b);
    c->xs = (float*)malloc(p->nx * sizeof(float)); c->zs = (float*)malloc(p->ny * sizeof(float));
    c->rh1 = (float*)malloc(p->nx * sizeof(float)); c->rh2 = (float*)malloc(p->nx * sizeof(float));
    c->dx2 = (float*)malloc(p->nx * sizeof(float));
    c->rv1 = (float*)malloc(p->ny * sizeof(float)); c->rv2 = (float*)malloc(p->ny * sizeof(float));
    c->dz2 = (float*)malloc(p->ny * sizeof(float));
    derive(c);
    /* V:254-260 */
    size_t count = 0;
    for (int j = 0; j < p->ny; ++j)
        for (int i = 0; i < p->nx; ++i) {
            c->x[count * 3 + 0] = c->xs[i];
            c->x[count * 3 + 1] = p->fullsize + 1;
            c->x[count * 3 + 2] = c->zs[j];
            count++;
        }
    memcpy(c->xl, c->x, b);
    if (p->integrator != OCO_VERLET) memset(c->xl, 0, b);      /* V = 0, E:270 / S:237 */
    return c;
}

void oco_destroy(oco_cloth* c)
{
    if (!c) return;
    free(c->x); free(c->xl); free(c->x2); free(c->xl2); free(c->xs); free(c->zs);
    free(c->rh1); free(c->rh2); free(c->dx2); free(c->rv1); free(c->rv2); free(c->dz2); free(c->pin);
    free(c);
}

/* run-time changeable scalars (everything except nx, ny, fullsize which fix the spring net) */
int oco_set_params(oco_cloth* c, const oco_params* p)
{
    if (p->nx != c->p.nx || p->ny != c->p.ny || p->fullsize != c->p.fullsize || p->integrator != c->p.integrator) return -1;
    c->p = *p;
    derive(c);
    return 0;
}

/* n < 0: back to the reference's literals */
void oco_set_pins(oco_cloth* c, const int* idx, int n)
{
    free(c->pin); c->pin = NULL;
    if (n < 0) return;
    c->pin = (unsigned char*)calloc(c->n, 1);
    for (int k = 0; k < n; ++k) c->pin[idx[k]] = 1;
}
size_t oco_num_particles(const oco_cloth* c) { return c->n; }
void oco_get_state(const oco_cloth* c, float* x, float* xl)
{
    if (x)  memcpy(x,  c->x,  c->n * 3 * sizeof(float));
    if (xl) memcpy(xl, c->xl, c->n * 3 * sizeof(float));
}
void oco_set_state(oco_cloth* c, const float* x, const float* xl)
{
    memcpy(c->x,  x,  c->n * 3 * sizeof(float));
    memcpy(c->xl, xl, c->n * 3 * sizeof(float));
}
/* copy rows [j0,j1) of (X, X_last) out of / into the cloth: the halo exchange of the row-band tests */
void oco_get_rows(const oco_cloth* c, int j0, int j1, float* x, float* xl)
{
    size_t o = (size_t)j0 * c->p.nx * 3, b = (size_t)(j1 - j0) * c->p.nx * 3 * sizeof(float);
    memcpy(x, c->x + o, b); memcpy(xl, c->xl + o, b);
}
void oco_set_rows(oco_cloth* c, int j0, int j1, const float* x, const float* xl)
{
    size_t o = (size_t)j0 * c->p.nx * 3, b = (size_t)(j1 - j0) * c->p.nx * 3 * sizeof(float);
    memcpy(c->x + o, x, b); memcpy(c->xl + o, xl, b);
}

/* GetVerletVelocity, V:445-447 */
static inline void velocity(const float* x, const float* xl, float dt, float v[3])
{
    v[0] = (x[0] - xl[0]) / dt; v[1] = (x[1] - xl[1]) / dt; v[2] = (x[2] - xl[2]) / dt;
}

/* One spring of ComputeForces' second loop (V:463-477) with p1 = a, p2 = b; returns springForce.
 * f(b,a) == -f(a,b) bit for bit (every product/sum sees both operands negated or none), so a
 * particle that is p2 of a spring may evaluate the spring with itself as p1 and ADD the result
 * instead of subtracting f(p1,p2) (V:480-482). */
static inline void spring(const float* pa, const float* va, const float* pb, const float* vb,
                          float rest, float ks, float kd, float f[3])
{
    float dpx = pa[0] - pb[0], dpy = pa[1] - pb[1], dpz = pa[2] - pb[2];          /* V:471 */
    float dvx = va[0] - vb[0], dvy = va[1] - vb[1], dvz = va[2] - vb[2];          /* V:472 */
    float sqr = dpx * dpx + dpy * dpy + dpz * dpz;                                  /* glm::length */
    float dist = sqrtf(sqr);                                                        /* V:473 */
    float left = -ks * (dist - rest);                                               /* V:475 */
    float right = kd * ((dvx * dpx + dvy * dpy + dvz * dpz) / dist);                /* V:476 */
    float inv = 1.0f / sqrtf(sqr);                                                  /* glm::normalize -> inversesqrt */
    float nx = dpx * inv, ny = dpy * inv, nz = dpz * inv;
    float s = left + right;
    f[0] = s * nx; f[1] = s * ny; f[2] = s * nz;                                    /* V:477 */
}

#define PX(i, j) (c->x  + ((size_t)(j) * u + (i)) * 3)
#define PL(i, j) (c->xl + ((size_t)(j) * u + (i)) * 3)
#define ADD_SPRING(ni, nj, rest, ks, kd)                                   \
    do {                                                                   \
        float vb_[3], f_[3];                                               \
        if (verlet) velocity(PX(ni, nj), PL(ni, nj), dt, vb_);             \
        else { const float* q_ = PL(ni, nj); vb_[0] = q_[0]; vb_[1] = q_[1]; vb_[2] = q_[2]; }   /* stored V */ \
        spring(xm, vm, PX(ni, nj), vb_, (rest), (ks), (kd), f_);           \
        F[0] += f_[0]; F[1] += f_[1]; F[2] += f_[2];                       \
    } while (0)

static void particle_step(const oco_cloth* c, int i, int j, float* xo, float* xlo)
{
    const oco_params* p = &c->p;
    const int u = p->nx, v = p->ny;
    const float dt = p->dt;
    const float* xm = PX(i, j);
    const float* xlm = PL(i, j);
    const size_t idx = (size_t)j * u + i;
    const int pinned = is_pinned(c, idx);                         /* V:455, V:479-482: i!=0 && i!=numX */

    const int verlet = (p->integrator == OCO_VERLET);

    /* ---- ComputeForces, first loop (V:451-459; E:436-445 / S:406-415) ---- */
    float F[3] = { 0.0f, 0.0f, 0.0f };
    float vm[3];
    if (verlet) velocity(xm, xlm, dt, vm);
    else { vm[0] = xlm[0]; vm[1] = xlm[1]; vm[2] = xlm[2]; }
    if (!pinned) {
        if (verlet) { F[0] += p->gravity[0] * p->mass; F[1] += p->gravity[1] * p->mass; F[2] += p->gravity[2] * p->mass; }
        else        { F[0] += p->gravity[0]; F[1] += p->gravity[1]; F[2] += p->gravity[2]; }                  /* E:441 */
    }
    F[0] += p->damping * vm[0]; F[1] += p->damping * vm[1]; F[2] += p->damping * vm[2];

    /* ---- ComputeForces, second loop (V:462-483) re-ordered per particle; the order below is the
     *      order in which the spring list (V:286-320) touches particle (i,j) ---- */
    if (!pinned) {
        /* structural, horizontal (V:288-291) */
        if (i - 1 >= 0) ADD_SPRING(i - 1, j, c->rh1[i - 1], p->ks_struct, p->kd_struct);
        if (i + 1 <  u) ADD_SPRING(i + 1, j, c->rh1[i],     p->ks_struct, p->kd_struct);
        /* structural, vertical (V:294-297) */
        if (j - 1 >= 0) ADD_SPRING(i, j - 1, c->rv1[j - 1], p->ks_struct, p->kd_struct);
        if (j + 1 <  v) ADD_SPRING(i, j + 1, c->rv1[j],     p->ks_struct, p->kd_struct);
        /* shear (V:301-305): cells visited row-major, two springs per cell */
        if (i - 1 >= 0 && j - 1 >= 0) ADD_SPRING(i - 1, j - 1, sqrtf(c->dx2[i - 1] + c->dz2[j - 1]), p->ks_shear, p->kd_shear);
        if (i + 1 <  u && j - 1 >= 0) ADD_SPRING(i + 1, j - 1, sqrtf(c->dx2[i]     + c->dz2[j - 1]), p->ks_shear, p->kd_shear);
        if (i - 1 >= 0 && j + 1 <  v) ADD_SPRING(i - 1, j + 1, sqrtf(c->dx2[i - 1] + c->dz2[j]),     p->ks_shear, p->kd_shear);
        if (i + 1 <  u && j + 1 <  v) ADD_SPRING(i + 1, j + 1, sqrtf(c->dx2[i]     + c->dz2[j]),     p->ks_shear, p->kd_shear);
        /* bend, horizontal (V:309-314): the last spring of every row is added twice (V:313) */
        if (i - 2 >= 0) ADD_SPRING(i - 2, j, c->rh2[i - 2], p->ks_bend, p->kd_bend);
        if (i + 2 <  u) ADD_SPRING(i + 2, j, c->rh2[i],     p->ks_bend, p->kd_bend);
        if (i == u - 3) ADD_SPRING(i + 2, j, c->rh2[i],     p->ks_bend, p->kd_bend);
        if (i == u - 1) ADD_SPRING(i - 2, j, c->rh2[i - 2], p->ks_bend, p->kd_bend);
        /* bend, vertical (V:315-320): the last spring of every column is added twice (V:319) */
        if (j - 2 >= 0) ADD_SPRING(i, j - 2, c->rv2[j - 2], p->ks_bend, p->kd_bend);
        if (j + 2 <  v) ADD_SPRING(i, j + 2, c->rv2[j],     p->ks_bend, p->kd_bend);
        if (j == v - 3) ADD_SPRING(i, j + 2, c->rv2[j],     p->ks_bend, p->kd_bend);
        if (j == v - 1) ADD_SPRING(i, j - 2, c->rv2[j - 2], p->ks_bend, p->kd_bend);
    }

    float nx_, ny_, nz_, lx, ly, lz;
    if (verlet) {
        /* ---- IntegrateVerlet (V:428-444) ---- */
        float dt2m = (dt * dt) / p->mass;                                                   /* V:429 */
        nx_ = xm[0] + (xm[0] - xlm[0]) + dt2m * F[0];                                       /* V:436 */
        ny_ = xm[1] + (xm[1] - xlm[1]) + dt2m * F[1];
        nz_ = xm[2] + (xm[2] - xlm[2]) + dt2m * F[2];
        lx = xm[0]; ly = xm[1]; lz = xm[2];                                                 /* V:438 */
    } else {
        /* ---- IntegrateEuler (E:469-482) / IntegrateSemiImplicit (S:464-477); l* is the new V ---- */
        float dtm = dt / p->mass;                                                           /* E:470 */
        lx = vm[0] + F[0] * dtm; ly = vm[1] + F[1] * dtm; lz = vm[2] + F[2] * dtm;          /* E:475 V += F*deltaTimeMass */
        if (p->integrator == OCO_EULER) { nx_ = xm[0] + dt * vm[0]; ny_ = xm[1] + dt * vm[1]; nz_ = xm[2] + dt * vm[2]; }   /* E:476 oldV */
        else                            { nx_ = xm[0] + dt * lx;    ny_ = xm[1] + dt * ly;    nz_ = xm[2] + dt * lz; }      /* S:470 */
    }
    if (ny_ < 0) ny_ = 0;                                                               /* V:440-442 */

    /* ---- EllipsoidCollision (V:509-533) ---- */
    const float* m = p->inv_ellipsoid;   /* m[col*4+row] */
    float x0 = m[0] * nx_ + m[4] * ny_ + m[8]  * nz_ + m[12] * 1.0f;                    /* V:511 */
    float y0 = m[1] * nx_ + m[5] * ny_ + m[9]  * nz_ + m[13] * 1.0f;
    float z0 = m[2] * nx_ + m[6] * ny_ + m[10] * nz_ + m[14] * 1.0f;
    float d0x = x0 - p->center[0], d0y = y0 - p->center[1], d0z = z0 - p->center[2];    /* V:512 */
    float distance = sqrtf(d0x * d0x + d0y * d0y + d0z * d0z);                          /* V:513 */
    if (distance < 1.0f) {                                                              /* V:514 */
        float s = p->radius - distance;                                                 /* V:515 */
        d0x = (s * d0x) / distance; d0y = (s * d0y) / distance; d0z = (s * d0z) / distance;
        float dx = d0x * c->tinv[0][0] + d0y * c->tinv[0][1] + d0z * c->tinv[0][2];      /* V:520-522 */
        float dy = d0x * c->tinv[1][0] + d0y * c->tinv[1][1] + d0z * c->tinv[1][2];      /* V:523-525 */
        float dz = d0x * c->tinv[2][0] + d0y * c->tinv[2][1] + d0z * c->tinv[2][2];      /* V:526-528 */
        nx_ += dx; ny_ += dy; nz_ += dz;                                                /* V:529 */
        if (verlet) { lx = nx_; ly = ny_; lz = nz_; }                                   /* V:530 */
        else        { lx = 0.0f; ly = 0.0f; lz = 0.0f; }                                /* E:599 V[i] = vec3(0) */
    }
    xo[0] = nx_; xo[1] = ny_; xo[2] = nz_;
    xlo[0] = lx; xlo[1] = ly; xlo[2] = lz;
}

/* One StepPhysics restricted to rows [j0,j1): rows outside keep their values.  Reads rows
 * [j0-2, j1+2) of the previous state (the reach of the bend springs, V:311, V:317). */
void oco_step_rows(oco_cloth* c, int j0, int j1)
{
    const int u = c->p.nx;
    if (j0 < 0) j0 = 0;
    if (j1 > c->p.ny) j1 = c->p.ny;
    size_t b = c->n * 3 * sizeof(float);
    if (j0 > 0 || j1 < c->p.ny) { memcpy(c->x2, c->x, b); memcpy(c->xl2, c->xl, b); }
#pragma omp parallel for schedule(static)
    for (int j = j0; j < j1; ++j)
        for (int i = 0; i < u; ++i) {
            size_t o = ((size_t)j * u + i) * 3;
            particle_step(c, i, j, c->x2 + o, c->xl2 + o);
        }
    float* t;
    t = c->x;  c->x  = c->x2;  c->x2  = t;
    t = c->xl; c->xl = c->xl2; c->xl2 = t;
}

/* ApplyProvotDynamicInverse (V:486-508; E:554-577 / S:437-462) over the implicit spring list in the reference's list
 * order (V:286-320), sequentially like the reference: in the Verlet demo the corrections move X in place, so every
 * spring sees the positions its predecessors left (Gauss-Seidel); in the Euler demos they are added to V. */
static void provot_spring(oco_cloth* c, int i1, int j1, int i2, int j2, float rest)
{
    const int u = c->p.nx;
    float* p1 = PX(i1, j1); float* p2 = PX(i2, j2);
    float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];              /* V:491 */
    float sqr = dx * dx + dy * dy + dz * dz;
    float dist = sqrtf(sqr);                                                        /* V:492 */
    if (dist > rest) {                                                              /* V:493 */
        dist -= rest;                                                               /* V:494 */
        dist /= 2.0f;                                                               /* V:495 */
        float inv = 1.0f / sqrtf(sqr);                                              /* V:496 glm::normalize */
        dx = dx * inv; dy = dy * inv; dz = dz * inv;
        dx *= dist; dy *= dist; dz *= dist;                                         /* V:497 */
        const size_t a = (size_t)j1 * u + i1, b = (size_t)j2 * u + i2;
        const int pin1 = is_pinned(c, a), pin2 = is_pinned(c, b);
        float* t1 = c->p.integrator == OCO_VERLET ? p1 : PL(i1, j1);                /* X (V:498-505) or V (E:567-574) */
        float* t2 = c->p.integrator == OCO_VERLET ? p2 : PL(i2, j2);
        if (pin1)      { t2[0] += dx; t2[1] += dy; t2[2] += dz; }                   /* V:498-499 */
        else if (pin2) { t1[0] -= dx; t1[1] -= dy; t1[2] -= dz; }                   /* V:500-501 */
        else           { t1[0] -= dx; t1[1] -= dy; t1[2] -= dz; t2[0] += dx; t2[1] += dy; t2[2] += dz; }   /* V:503-504 */
    }
}
void oco_provot(oco_cloth* c)
{
    const int u = c->p.nx, v = c->p.ny;
    for (int j = 0; j < v; ++j) for (int i = 0; i < u - 1; ++i) provot_spring(c, i, j, i + 1, j, c->rh1[i]);     /* V:288-291 */
    for (int i = 0; i < u; ++i) for (int j = 0; j < v - 1; ++j) provot_spring(c, i, j, i, j + 1, c->rv1[j]);     /* V:294-297 */
    for (int j = 0; j < v - 1; ++j) for (int i = 0; i < u - 1; ++i) {                                            /* V:301-305 */
        float r = sqrtf(c->dx2[i] + c->dz2[j]);
        provot_spring(c, i, j, i + 1, j + 1, r);
        provot_spring(c, i, j + 1, i + 1, j, r);
    }
    for (int j = 0; j < v; ++j) {                                                                                /* V:309-314 */
        for (int i = 0; i < u - 2; ++i) provot_spring(c, i, j, i + 2, j, c->rh2[i]);
        provot_spring(c, u - 3, j, u - 1, j, c->rh2[u - 3]);
    }
    for (int i = 0; i < u; ++i) {                                                                                /* V:315-320 */
        for (int j = 0; j < v - 2; ++j) provot_spring(c, i, j, i, j + 2, c->rv2[j]);
        provot_spring(c, i, v - 3, i, v - 1, c->rv2[v - 3]);
    }
}

void oco_step(oco_cloth* c, int n)
{
    for (int s = 0; s < n; ++s) {
        oco_step_rows(c, 0, c->p.ny);
        if (c->p.provot) oco_provot(c);
    }
}

/* Spring energy diagnostic over the reference spring list, duplicates included (ours; the
 * reference has no energy function): sum 1/2 Ks (|p1-p2| - rest)^2, accumulated in double, in
 * the reference's list order (V:286-320) so that it matches oracle/ref_shim.cpp bit for bit. */
static double e_term(const oco_cloth* c, int i1, int j1, int i2, int j2, float rest, float ks)
{
    const int u = c->p.nx;
    const float* a = PX(i1, j1); const float* b = PX(i2, j2);
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    double len = sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
    double ext = len - (double)rest;
    return 0.5 * (double)ks * ext * ext;
}
double oco_spring_energy(const oco_cloth* c)
{
    const oco_params* p = &c->p;
    const int u = p->nx, v = p->ny;
    double e = 0.0;
    for (int j = 0; j < v; ++j) for (int i = 0; i < u - 1; ++i) e += e_term(c, i, j, i + 1, j, c->rh1[i], p->ks_struct);
    for (int i = 0; i < u; ++i) for (int j = 0; j < v - 1; ++j) e += e_term(c, i, j, i, j + 1, c->rv1[j], p->ks_struct);
    for (int j = 0; j < v - 1; ++j) for (int i = 0; i < u - 1; ++i) {
        float r = sqrtf(c->dx2[i] + c->dz2[j]);
        e += e_term(c, i, j, i + 1, j + 1, r, p->ks_shear);
        e += e_term(c, i, j + 1, i + 1, j, r, p->ks_shear);
    }
    for (int j = 0; j < v; ++j) {
        for (int i = 0; i < u - 2; ++i) e += e_term(c, i, j, i + 2, j, c->rh2[i], p->ks_bend);
        e += e_term(c, u - 3, j, u - 1, j, c->rh2[u - 3], p->ks_bend);
    }
    for (int i = 0; i < u; ++i) {
        for (int j = 0; j < v - 2; ++j) e += e_term(c, i, j, i, j + 2, c->rv2[j], p->ks_bend);
        e += e_term(c, i, v - 3, i, v - 1, c->rv2[v - 3], p->ks_bend);
    }
    return e;
}

/* rest-length tables, for the unit test against the verbatim spring list */
void oco_get_tables(const oco_cloth* c, float* rh1, float* rh2, float* rv1, float* rv2, float* dx2, float* dz2)
{
    memcpy(rh1, c->rh1, c->p.nx * sizeof(float)); memcpy(rh2, c->rh2, c->p.nx * sizeof(float));
    memcpy(dx2, c->dx2, c->p.nx * sizeof(float));
    memcpy(rv1, c->rv1, c->p.ny * sizeof(float)); memcpy(rv2, c->rv2, c->p.ny * sizeof(float));
    memcpy(dz2, c->dz2, c->p.ny * sizeof(float));
}
