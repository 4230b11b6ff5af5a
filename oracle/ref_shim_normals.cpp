// oracle/ref_shim_normals.cpp — TEST INFRASTRUCTURE ONLY.
//
// Headless shim around the verbatim UpdateNormals of the reference's lit demo
//   /root/reference/OpenCloth_ExplicitEuler_TextureMapped_Lit/OpenCloth_ExplicitEuler_TextureMapped_Lit/main.cpp ("L:")
// (L:684-707) and the triangle list it walks (InitGL, L:312-328), cut by line range by oracle/build_ref.sh into
// oracle/_ref/slices_lit/ and #included here.  Used to pin the product's oc_download_normals (oc_normals.cuh).
// The reference's index type is GLushort (L:77), so this checker is limited to grids of at most 65536 particles.
#include <vector>
#include <cmath>
#include <cstring>
#include <cstddef>
#include <glm/glm.hpp>

using namespace std;
typedef unsigned short GLushort;

int    numX = 20, numY = 20;                     // L:59
size_t total_points = (numX + 1) * (numY + 1);   // L:60 (const there)
vector<GLushort>  indices;                       // L:77
vector<glm::vec3> X;                             // L:80
#include "vertex_struct.inc"                     // L:89-90  struct Vertex, vertices
#include "update_normals.inc"                    // L:684-707

extern "C" int ref_normals(int nx, int ny, const float* x, float* n_out, int calls)
{
    if (nx < 2 || ny < 2 || (size_t)nx * ny > 65536) return -1;
    numX = nx - 1; numY = ny - 1;
    total_points = (size_t)nx * ny;
    int i = 0, j = 0;
    indices.resize(numX * numY * 2 * 3);         // L:253
    X.resize(total_points);
    vertices.assign(total_points, Vertex());     // value-initialised: n = 0, as after vertices.resize (L:255)
    memcpy(&X[0], x, total_points * sizeof(glm::vec3));
#include "init_indices.inc"                      // L:312-328
    for (int c = 0; c < calls; ++c) UpdateNormals();      // calls > 1 shows the reference's frame-to-frame accumulation
    for (size_t p = 0; p < total_points; ++p) memcpy(n_out + 3 * p, &vertices[p].n, sizeof(glm::vec3));
    return 0;
}
