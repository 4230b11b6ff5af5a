// oc_harness.cpp — headless host harness over the C-ABI (include/opencloth.h).
//
// The reference's physics lives inside a single-file GLUT program
// (/root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp, "V:"): main() V:564-584 opens a window,
// OnIdle V:536-555 calls StepPhysics(timeStep) once per idle tick, OnRender V:342-416 draws X and puts
// FPS / frame time into the window title (V:356-366).  This program is that loop with GL removed:
//
//     InitGL (V:232-328)            ->  oc_default_params + oc_create
//     OnIdle: StepPhysics (V:550)   ->  oc_step(handle, n)            (n substeps per "frame")
//     OnRender's title bar          ->  one line per frame: step, particle-updates/s, spring energy
//     glutMainLoop / OnShutdown     ->  --frames iterations, oc_destroy
//     mouse drag (V:184-210)        ->  --poke idx,x,y,z  (oc_set_particle before the first frame)
//
// It can also cut the cloth into row bands over several GPUs of ONE process (--gpus g [--devices d]): every band
// is a handle on device (band % d).  By default the bands are LINKED (oc_band_link_local): the step kernel stores
// the boundary rows into the neighbour's halo itself (peer memory) and no exchange step exists; --link 0 keeps the
// older host-driven exchange (oc_halo_exchange, cudaMemcpyPeerAsync every halo_rows/2 substeps).
// No oracle, no CPU path: without a CUDA device it prints the library's error and exits 2.
#include "../../include/opencloth.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// NumPy .npy (format 1.0) writer: a self-describing frame for tools that cannot guess the shape of a raw dump
static bool write_npy(const std::string& path, const float* data, size_t rows, size_t cols)
{
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    char dict[128];
    int n = snprintf(dict, sizeof(dict), "{'descr': '<f4', 'fortran_order': False, 'shape': (%zu, %zu), }", rows, cols);
    size_t total = 10 + (size_t)n + 1;                       // magic(6) + version(2) + header length(2) + dict + newline
    size_t pad = (64 - total % 64) % 64;
    unsigned short hlen = (unsigned short)(n + pad + 1);
    fwrite("\x93NUMPY\x01\x00", 1, 8, fp);
    fwrite(&hlen, 2, 1, fp);
    fwrite(dict, 1, (size_t)n, fp);
    for (size_t k = 0; k < pad; ++k) fputc(' ', fp);
    fputc('\n', fp);
    bool ok = fwrite(data, sizeof(float), rows * cols, fp) == rows * cols;
    fclose(fp);
    return ok;
}

static void die(const char* what)
{
    fprintf(stderr, "oc_harness: %s: %s\n", what, oc_last_error());
    exit(2);
}
#define CK(call) do { if ((call) != OC_OK) die(#call); } while (0)

int main(int argc, char** argv)
{
    int nx = 21, ny = 21, frames = 10, substeps = 100, k = 1, exact = 1, gpus = 1, batch = 1, halo = 16, energy = 1, link = 1, devices = 0;
    int poke_idx = -1; float poke[3] = { 0, 0, 0 };
    std::string dump, dump_npy;
    for (int a = 1; a < argc; ++a) {
        auto is = [&](const char* f) { return !strcmp(argv[a], f) && a + 1 < argc; };
        if (is("--nx")) nx = atoi(argv[++a]);
        else if (is("--ny")) ny = atoi(argv[++a]);
        else if (is("--frames")) frames = atoi(argv[++a]);
        else if (is("--substeps")) substeps = atoi(argv[++a]);
        else if (is("--k")) k = atoi(argv[++a]);
        else if (is("--exact")) exact = atoi(argv[++a]);
        else if (is("--gpus")) gpus = atoi(argv[++a]);
        else if (is("--batch")) batch = atoi(argv[++a]);
        else if (is("--halo")) halo = atoi(argv[++a]);
        else if (is("--link")) link = atoi(argv[++a]);
        else if (is("--devices")) devices = atoi(argv[++a]);
        else if (is("--energy")) energy = atoi(argv[++a]);
        else if (is("--dump")) dump = argv[++a];
        else if (is("--dump-npy")) dump_npy = argv[++a];
        else if (is("--poke")) { if (sscanf(argv[++a], "%d,%f,%f,%f", &poke_idx, &poke[0], &poke[1], &poke[2]) != 4) { fprintf(stderr, "--poke idx,x,y,z\n"); return 1; } }
        else {
            fprintf(stderr, "usage: oc_harness [--nx N --ny N] [--frames F] [--substeps S] [--k K] [--exact 0|1] [--gpus G [--devices D] [--link 0|1]]\n"
                            "                  [--batch B] [--halo ROWS] [--energy 0|1] [--poke idx,x,y,z] [--dump file.f32] [--dump-npy prefix]\n");
            return 1;
        }
    }
    printf("%s\n", oc_version());

    std::vector<oc_cloth*> bands(gpus, nullptr);
    for (int g = 0; g < gpus; ++g) {
        oc_params p;
        CK(oc_default_params(&p, nx, ny));
        p.batch = batch; p.substeps_per_launch = k; p.exact = exact;
        if (gpus > 1) {
            p.row_begin = (int)((long long)ny * g / gpus); p.row_end = (int)((long long)ny * (g + 1) / gpus);
            p.halo_rows = link ? 2 : halo; p.device = g % (devices > 0 ? devices : gpus);
            if (link) { p.substeps_per_launch = 1; p.kernel = OC_KERNEL_AUTO; }
        }
        CK(oc_create(&bands[g], &p));
    }
    if (poke_idx >= 0) for (auto* b : bands) CK(oc_set_particle(b, 0, poke_idx, poke));      // V:203-208
    const bool linked = gpus > 1 && link;
    if (linked) CK(oc_band_link_local(bands.data(), gpus));

    const double particles = (double)nx * ny * batch;
    int step = 0;
    for (int f = 0; f < frames; ++f) {
        auto t0 = std::chrono::steady_clock::now();
        int left = substeps;
        while (left > 0) {
            int n = left;
            if (linked) n = 1;          // interleave the bands' launches: each band's edge tiles wait for the neighbours' previous step
            else if (gpus > 1) {
                if (oc_halo_budget(bands[0]) == 0) CK(oc_halo_exchange(bands.data(), gpus));
                n = oc_halo_budget(bands[0]) < left ? oc_halo_budget(bands[0]) : left;
            }
            for (auto* b : bands) CK(oc_step(b, n));
            left -= n;
        }
        for (auto* b : bands) CK(oc_sync(b));
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        step += substeps;
        double e = 0.0;
        if (energy && gpus == 1) CK(oc_spring_energy(bands[0], 0, &e));
        printf("{\"step\": %d, \"frame_ms\": %.3f, \"particle_updates_per_s\": %.4e, \"spring_energy\": %.9g, \"gpus\": %d, \"k\": %d, \"exact\": %d}\n",
               step, s * 1e3, particles * substeps / s, e, gpus, k, exact);
    }
    if (!dump.empty()) {            // render hand-off: X as float4 (x,y,z,1), the layout of the reference's pos_vbo
        FILE* fp = fopen(dump.c_str(), "wb");
        if (!fp) { perror("dump"); return 1; }
        for (int g = 0; g < gpus; ++g) {
            oc_params p; CK(oc_get_params(bands[g], &p));
            size_t n = (size_t)(p.row_end - p.row_begin) * nx * batch;
            std::vector<float> x(n * 4);
            CK(oc_download(bands[g], x.data(), nullptr, 4));
            fwrite(x.data(), sizeof(float), x.size(), fp);
        }
        fclose(fp);
        printf("wrote %s (%d x %d x %d float4)\n", dump.c_str(), batch, ny, nx);
    }
    if (!dump_npy.empty()) {        // render hand-off as self-describing frames: <prefix>_X.npy (n x 4: x,y,z,1), <prefix>_N.npy (n x 3 vertex normals)
        const size_t total = (size_t)nx * ny * batch;
        std::vector<float> x(total * 4);
        size_t off = 0;
        for (int g = 0; g < gpus; ++g) {
            oc_params p; CK(oc_get_params(bands[g], &p));
            CK(oc_download(bands[g], x.data() + off, nullptr, 4));
            off += (size_t)(p.row_end - p.row_begin) * nx * batch * 4;
        }
        if (!write_npy(dump_npy + "_X.npy", x.data(), total, 4)) { perror("dump-npy"); return 1; }
        // normals need the whole mesh on one handle: with row bands, a whole-cloth handle receives the gathered state
        oc_cloth* whole = bands[0];
        if (gpus > 1) {
            oc_params p; CK(oc_default_params(&p, nx, ny)); p.batch = batch;
            CK(oc_create(&whole, &p));
            std::vector<float> x3(total * 3);
            for (size_t q = 0; q < total; ++q) { x3[3 * q] = x[4 * q]; x3[3 * q + 1] = x[4 * q + 1]; x3[3 * q + 2] = x[4 * q + 2]; }
            CK(oc_upload(whole, x3.data(), x3.data(), 3));
        }
        std::vector<float> nrm(total * 3);
        CK(oc_download_normals(whole, nrm.data(), 3));
        if (gpus > 1) oc_destroy(whole);
        if (!write_npy(dump_npy + "_N.npy", nrm.data(), total, 3)) { perror("dump-npy"); return 1; }
        printf("wrote %s_X.npy (%zu x 4) and %s_N.npy (%zu x 3)\n", dump_npy.c_str(), total, dump_npy.c_str(), total);
    }
    for (auto* b : bands) oc_destroy(b);
    return 0;
}
