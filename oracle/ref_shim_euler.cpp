// oracle/ref_shim_euler.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Headless shim around the *verbatim* physics of the reference's two explicit-integrator siblings of the Verlet demo
// (SURVEY.md 8(f)3), built twice by oracle/build_ref.sh:
//   -DOC_REF_EXPLICIT_EULER  /root/reference/OpenCloth_ExplicitEuler/OpenCloth_ExplicitEuler/main.cpp  ("E:")
//                            StepPhysics E:626-641 = ComputeForces E:434-466, IntegrateEuler E:469-482,
//                            EllipsoidCollision E:578-602, ApplyProvotDynamicInverse E:554-577
//   -DOC_REF_SEMI_IMPLICIT   /root/reference/OpenCloth_SemiImplicit/OpenCloth_SemiImplicit/main.cpp    ("S:")
//                            StepPhysics S:527-532 = ComputeForces S:402-436, IntegrateSemiImplicit S:464-477,
//                            EllipsoidCollision S:478-502, ApplyProvotDynamicInverse S:437-462
// Like ref_shim.cpp: the physics is the reference's own text, cut by line range into oracle/_ref/slices_*/ (git-ignored)
// and #included here; compiled against the reference's vendored GLM 0.9.0.0.  The shim adds the globals with a run-time
// grid size, the StepPhysics composition in the reference's order (its StepPhysics text also holds commented-out
// alternatives, so the four calls are restated here), and extern "C" entry points.
#include <vector>
#include <cmath>
#include <cstring>
#include <cstddef>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>

using namespace std;

int    numX = 20, numY = 20;                     // E:59 / S:40
size_t total_points = (numX + 1) * (numY + 1);   // E:60 / S:41 (const there)
float  fullsize = 4.0f;                          // E:61
float  halfsize = fullsize / 2.0f;               // E:62
#include "timestep.inc"                          // E:64 / S:45   timeStep

#include "spring_struct.inc"
vector<Spring>    springs;
vector<glm::vec3> X;
vector<glm::vec3> V;
vector<glm::vec3> F;

#include "constants.inc"
#include "params.inc"
#include "ellipsoid_globals.inc"
#include "add_spring.inc"
#include "forces.inc"
#include "integrate.inc"
#include "provot.inc"
#include "collision.inc"

static int g_provot = 1;

void StepPhysics(float dt)
{
    ComputeForces();                             // E:627 / S:528
#ifdef OC_REF_EXPLICIT_EULER
    IntegrateEuler(dt);                          // E:630
#else
    IntegrateSemiImplicit(dt);                   // S:529
#endif
    EllipsoidCollision();                        // E:638 / S:530
    if (g_provot) ApplyProvotDynamicInverse();   // E:639 / S:531
}

static void InitHeadless()
{
    int i = 0, j = 0, count = 0;
    int l1 = 0, l2 = 0;
    int v = numY + 1;
    int u = numX + 1;
    springs.clear();
    total_points = (size_t)(numX + 1) * (size_t)(numY + 1);
    X.resize(total_points);
    V.resize(total_points);
    F.resize(total_points);
#include "init_state.inc"
#include "init_springs.inc"
}

extern "C" {

int ref_init(int nx, int ny)
{
    if (nx < 3 || ny < 3) return -1;
    numX = nx - 1; numY = ny - 1;
    InitHeadless();
    return 0;
}
size_t ref_num_particles(void) { return total_points; }
size_t ref_num_springs(void)   { return springs.size(); }
void   ref_set_provot(int on)  { g_provot = on; }
void   ref_step(int n)         { for (int s = 0; s < n; ++s) StepPhysics(timeStep); }
void   ref_get_state(float* x, float* v)
{
    if (x) memcpy(x, &X[0], total_points * sizeof(glm::vec3));
    if (v) memcpy(v, &V[0], total_points * sizeof(glm::vec3));
}
void   ref_set_state(const float* x, const float* v)
{
    memcpy(&X[0], x, total_points * sizeof(glm::vec3));
    memcpy(&V[0], v, total_points * sizeof(glm::vec3));
}
void   ref_get_params(float* out /*[16]*/)
{
    out[0] = DEFAULT_DAMPING; out[1] = KsStruct; out[2] = KdStruct; out[3] = KsShear; out[4] = KdShear;
    out[5] = KsBend; out[6] = KdBend; out[7] = gravity.x; out[8] = gravity.y; out[9] = gravity.z;
    out[10] = mass; out[11] = timeStep; out[12] = fullsize; out[13] = radius;
    out[14] = center.x; out[15] = center.y;
}
void   ref_get_ellipsoid(float* m, float* inv)
{
    memcpy(m,   &ellipsoid[0][0],         16 * sizeof(float));
    memcpy(inv, &inverse_ellipsoid[0][0], 16 * sizeof(float));
}

} // extern "C"
