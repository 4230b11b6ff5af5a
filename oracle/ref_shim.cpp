// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Headless shim around the *verbatim* physics of the reference demo
//   /root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp   ("V:" below)
// The reference is a single-file GLUT/Win32 program and cannot be compiled as a whole
// (V:46-48 need GL/glew.h, GL/wglew.h, GL/freeglut.h; V:114-115 LARGE_INTEGER; V:564 `void main`).
// oracle/build_ref.sh therefore cuts the physics out of it by line range with `sed` into
// oracle/_ref/slices/*.inc (git-ignored; reference SOURCES are never copied into the repo)
// and this file #includes those slices where they lie.  Everything numerically relevant —
// Spring, AddSpring, IntegrateVerlet, GetVerletVelocity, ComputeForces, EllipsoidCollision,
// StepPhysics, the position/spring/ellipsoid set-up of InitGL — is the reference's own text,
// compiled against the reference's own vendored GLM 0.9.0.0 (dep/glm).
//
// What this shim adds (and nothing else):
//   * the file-scope globals of V:59-62, V:76-80, V:90-104, V:123-130, re-declared so that the
//     grid size is a run-time value (V:60 makes total_points a const initialised from numX/numY);
//   * extern "C" entry points so tests / bench can drive it through ctypes.
#include <vector>
#include <cmath>
#include <cstring>
#include <cstddef>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>

using namespace std;                       // V:55

// ---- globals, as V:59-62 (total_points made non-const) --------------------------------------
int    numX = 20, numY = 20;               // V:59
size_t total_points = (numX + 1) * (numY + 1);   // V:60 (const there)
float  fullsize = 4.0f;                    // V:61
float  halfsize = fullsize / 2.0f;         // V:62

#include "_ref/slices/spring_struct.inc"   // V:67-72   struct Spring

vector<Spring>    springs;                 // V:76
vector<glm::vec3> X;                       // V:78
vector<glm::vec3> X_last;                  // V:79
vector<glm::vec3> F;                       // V:80

#include "_ref/slices/constants.inc"       // V:90-93  spring type ids, spring_count
#include "_ref/slices/params.inc"          // V:97-104 damping, Ks/Kd, gravity, mass, timeStep
#include "_ref/slices/ellipsoid_globals.inc" // V:123-132 ellipsoid, inverse_ellipsoid, center, radius, StepPhysics decl

#include "_ref/slices/add_spring.inc"      // V:134-144 AddSpring
#include "_ref/slices/physics.inc"         // V:428-484 IntegrateVerlet, GetVerletVelocity, ComputeForces
#include "_ref/slices/provot.inc"          // V:486-508 ApplyProvotDynamicInverse
#include "_ref/slices/collision.inc"       // V:509-533 EllipsoidCollision
#include "_ref/slices/step.inc"            // V:557-562 StepPhysics

static void InitHeadless()
{
    int i = 0, j = 0, count = 0;           // V:242
    int l1 = 0, l2 = 0;                    // V:243
    int v = numY + 1;                      // V:244
    int u = numX + 1;                      // V:245
    springs.clear();
    total_points = (size_t)(numX + 1) * (size_t)(numY + 1);
    X.resize(total_points);                // V:249
    X_last.resize(total_points);           // V:250
    F.resize(total_points);                // V:251
#include "_ref/slices/init_positions.inc"  // V:253-260
#include "_ref/slices/init_springs.inc"    // V:286-327 (springs + ellipsoid matrices)
}

extern "C" {

// grid is nx x ny PARTICLES (reference: numX+1, numY+1)
int ref_init(int nx, int ny)
{
    if (nx < 3 || ny < 3) return -1;
    numX = nx - 1; numY = ny - 1;
    InitHeadless();
    return 0;
}
size_t ref_num_particles(void) { return total_points; }
size_t ref_num_springs(void)   { return springs.size(); }
void   ref_step(int n)         { for (int s = 0; s < n; ++s) StepPhysics(timeStep); }
void   ref_step_dt(int n, float dt) { for (int s = 0; s < n; ++s) StepPhysics(dt); }
// StepPhysics with the call the reference leaves commented out (V:561) enabled: the body of V:558-561, in its order
void   ref_step_provot(int n)
{
    for (int s = 0; s < n; ++s) { ComputeForces(timeStep); IntegrateVerlet(timeStep); EllipsoidCollision(); ApplyProvotDynamicInverse(); }
}
void   ref_provot_only(void) { ApplyProvotDynamicInverse(); }
void   ref_get_state(float* x, float* xl)
{
    if (x)  memcpy(x,  &X[0],      total_points * sizeof(glm::vec3));
    if (xl) memcpy(xl, &X_last[0], total_points * sizeof(glm::vec3));
}
void   ref_set_state(const float* x, const float* xl)
{
    memcpy(&X[0],      x,  total_points * sizeof(glm::vec3));
    memcpy(&X_last[0], xl, total_points * sizeof(glm::vec3));
}
// spring table (p1,p2 int32; rest,Ks,Kd float32) for unit tests and the energy diagnostic
void   ref_get_springs(int* p1, int* p2, float* rest, float* ks, float* kd, int* type)
{
    for (size_t s = 0; s < springs.size(); ++s) {
        if (p1) p1[s] = springs[s].p1;
        if (p2) p2[s] = springs[s].p2;
        if (rest) rest[s] = springs[s].rest_length;
        if (ks) ks[s] = springs[s].Ks;
        if (kd) kd[s] = springs[s].Kd;
        if (type) type[s] = springs[s].type;
    }
}
// column-major 4x4, as glm stores them
void   ref_get_ellipsoid(float* m, float* inv)
{
    memcpy(m,   &ellipsoid[0][0],         16 * sizeof(float));
    memcpy(inv, &inverse_ellipsoid[0][0], 16 * sizeof(float));
}
void   ref_get_params(float* out /*[16]*/)
{
    out[0] = DEFAULT_DAMPING; out[1] = KsStruct; out[2] = KdStruct; out[3] = KsShear; out[4] = KdShear;
    out[5] = KsBend; out[6] = KdBend; out[7] = gravity.x; out[8] = gravity.y; out[9] = gravity.z;
    out[10] = mass; out[11] = timeStep; out[12] = fullsize; out[13] = radius;
    out[14] = center.x; out[15] = center.y;
}
// Spring energy diagnostic (ours; the reference has none): sum 1/2 Ks (|p1-p2| - rest)^2 in double
double ref_spring_energy(void)
{
    double e = 0.0;
    for (size_t s = 0; s < springs.size(); ++s) {
        glm::vec3 d = X[springs[s].p1] - X[springs[s].p2];
        double len = std::sqrt((double)d.x * d.x + (double)d.y * d.y + (double)d.z * d.z);
        double ext = len - (double)springs[s].rest_length;
        e += 0.5 * (double)springs[s].Ks * ext * ext;
    }
    return e;
}

} // extern "C"
