/* opencloth.h — C-ABI of libopencloth_b200.so
 *
 * A B200-native (sm_100a) drop-in for ONE path of mmmovania/opencloth: the Verlet mass-spring
 * cloth step  StepPhysics = ComputeForces -> IntegrateVerlet -> EllipsoidCollision  of
 * OpenCloth_Verlet/OpenCloth_Verlet/main.cpp ("V:" below; StepPhysics is V:557-562).
 *
 * The reference has no library interface: the step is a free function over file-scope globals
 * (V:76-80 X, X_last, F, springs; V:97-104 and V:123-130 parameters), initialised by InitGL
 * (V:249-327) and driven by OnIdle (V:548-552).  Its own GPU back ends (5-mode demo "C:" =
 * OpenCloth_Verlet_CUDA_GLSL_OPENCL_CPU/OpenCloth_Verlet_CUDA/main.cpp:145-154, verlet.cu "H:",
 * verlet_cl.cpp "LH:") expose the de-facto operator interface Init / Upload / Verlet / ReadBuffer /
 * Shutdown.  Each entry point below says which of those it replaces.
 *
 * Conventions: every function returns OC_OK (0) or a negative oc_status and never aborts
 * (the reference's cutilSafeCall / oclCheckError abort, H:21, LH:30); oc_last_error() gives the text of
 * the last failure on the calling thread.  A handle is not thread safe; distinct handles are
 * independent.  The caller owns every host buffer.  oc_step is asynchronous; oc_sync / oc_download
 * wait.  There is NO CPU fallback: without a CUDA device every compute call fails with
 * OC_ERR_NO_DEVICE.
 *
 * Particle (i,j), i = column 0..nx-1 (x direction), j = row 0..ny-1 (z direction), linear index
 * j*nx + i, exactly as V:254-259.  Host-side state is float[3] (stride 3, the reference's
 * vector<glm::vec3>) or float[4] (stride 4, the float4 of the reference's GPU back ends, w = 1).
 */
#ifndef OPENCLOTH_H
#define OPENCLOTH_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_ABI_VERSION 2

typedef enum oc_status {
    OC_OK = 0,
    OC_ERR_INVALID = -1,     /* bad argument */
    OC_ERR_NO_DEVICE = -2,   /* no CUDA device / driver: there is no CPU fallback */
    OC_ERR_CUDA = -3,        /* a CUDA runtime call failed (text in oc_last_error) */
    OC_ERR_NOMEM = -4,
    OC_ERR_UNSUPPORTED = -5
} oc_status;

typedef enum oc_kernel {
    OC_KERNEL_AUTO = 0,      /* resident for small whole cloths; else, one substep per launch: march2 in exact mode, stream in fast
                                mode from about three million particles per handle (march2 below); march for k > 1 */
    OC_KERNEL_GATHER = 1,    /* one thread per particle, 12-neighbour gather from global memory */
    OC_KERNEL_MARCH = 2,     /* fused shared-memory marching stencil, one column per thread, k substeps per launch */
    OC_KERNEL_MARCH2 = 3,    /* the same with two columns per thread (one substep per launch) */
    OC_KERNEL_RESIDENT = 4,  /* small whole cloths (<= 1536 particles): one CTA per cloth keeps the state in shared memory
                                and takes all the substeps of an oc_step call in one launch; larger cloths and row
                                bands are served by OC_KERNEL_MARCH2 */
    OC_KERNEL_TWIN = 5,      /* marching stencil, one column per thread, every packed FP32x2 operation spans the same column of
                                TWO independent tiles (two row segments of a strip, or two cloths of a batch); one substep
                                per launch; whole cloths, batches, row bands and linked row bands */
    OC_KERNEL_STREAM = 6,    /* streaming gather over twin tiles: every particle evaluates all of its twelve springs itself
                                from a shared-memory ring of rows (no force exchange, small per-thread state, many resident
                                warps); same coverage as OC_KERNEL_TWIN */
    OC_KERNEL_STREAM2 = 7,   /* the same with two adjacent columns per thread: a third of the neighbour loads is shared */
    OC_KERNEL_BANDRES = 8    /* mid-size whole cloths (about 10^3 .. 3*10^5 particles, single cloth): one row band per CTA, at most
                                one CTA per SM (cooperative launch), state resident in shared memory for all the substeps of
                                an oc_step call, two boundary rows per band exchanged through global memory per substep.
                                AUTO between OC_KERNEL_RESIDENT and the marching kernels; falls back to OC_KERNEL_MARCH2
                                where it does not apply (batches, row bands, Provot pass, bands too tall for shared memory) */
} oc_kernel;

/* The reference's explicit integrators on this spring net.  "E:" =
 * OpenCloth_ExplicitEuler/OpenCloth_ExplicitEuler/main.cpp, "S:" = OpenCloth_SemiImplicit/OpenCloth_SemiImplicit/main.cpp.
 * The Euler demos keep X and V instead of X and X_last: for them every "X_last" argument / result of this API is V. */
typedef enum oc_integrator {
    OC_INTEGRATOR_VERLET = 0,          /* IntegrateVerlet V:428-444 (the hot path)                                       */
    OC_INTEGRATOR_EULER = 1,           /* ComputeForces E:434-466, IntegrateEuler E:469-482 (X += dt * old V)             */
    OC_INTEGRATOR_SEMI_IMPLICIT = 2    /* ComputeForces S:402-436, IntegrateSemiImplicit S:464-477 (X += dt * new V)      */
} oc_integrator;

typedef struct oc_cloth oc_cloth;      /* opaque; owns all device memory of one simulation */

/* All defaults (oc_default_params) are the reference's values. */
typedef struct oc_params {
    /* ---- fixed at oc_create ------------------------------------------------------------- */
    int   nx, ny;              /* particles per row / per column; reference 21 x 21 = (numX+1, numY+1), V:59 */
    int   batch;               /* independent cloths stepped together (1 = the reference's single cloth)  */
    int   row_begin, row_end;  /* row band [row_begin,row_end) of the ny rows this handle owns.
                                  0,0 = whole cloth.  Used by the multi-process row-band driver:
                                  one handle per GPU, see oc_halo_* below.                                  */
    int   halo_rows;           /* band only: rows of neighbour state kept either side (multiple of 2)    */
    int   device;              /* CUDA device ordinal, -1 = current                                        */
    float fullsize;            /* 4.0f, V:61 (halfsize = fullsize/2, V:62); cloth spans x[-2,2] z[0,4] y=5 */
    /* ---- changeable with oc_set_params ---------------------------------------------------- */
    int   substeps_per_launch; /* k: temporal blocking, 1..8 (march kernel); 0 = library default        */
    int   exact;               /* 1 = IEEE op-for-op order of the reference (bitwise equal to its CPU
                                  path); 0 = fast (FMA contraction, MUFU rsqrt; within the stated tolerance) */
    int   kernel;              /* oc_kernel                                                                */
    float ks_struct, kd_struct;/* 50.75f, -0.25f  V:98  */
    float ks_shear,  kd_shear; /* 50.75f, -0.25f  V:99  */
    float ks_bend,   kd_bend;  /* 50.95f, -0.25f  V:100 */
    float damping;             /* -0.0125f  V:97 DEFAULT_DAMPING */
    float gravity[3];          /* 0, -0.00981f, 0  V:101 */
    float mass;                /* 1.0f  V:102 */
    float dt;                  /* 1/60.0f  V:104 timeStep */
    float ellipsoid[16];       /* column-major; translate(0,2,0)*rotate(45deg,x)*scale(1,1,.5)  V:324-326 */
    float inv_ellipsoid[16];   /* glm::inverse(ellipsoid)  V:327 */
    float center[3];           /* 0,0,0  V:129 */
    float radius;              /* 1.0f   V:130 */
    /* ---- the step's optional parts ---------------------------------------------------------------------------- */
    int   integrator;          /* oc_integrator; fixed at oc_create.  Euler variants: whole cloths only (no row bands),
                                  served by the gather kernel                                                       */
    int   provot;              /* 1: ApplyProvotDynamicInverse after EllipsoidCollision, bit-exact in list order
                                  (V:486-508, which the Verlet demo leaves disabled at V:561 — default 0; E:554-577 /
                                  S:437-462, enabled at E:639 / S:531 — default 1 there).  Changeable; whole cloths only */
} oc_params;

/* Fill *p with the reference's values for an nx x ny cloth (replaces the globals V:59-62, V:97-104,
 * V:123-130 and the ellipsoid set-up V:324-327). */
int oc_default_params(oc_params* p, int nx, int ny);
/* The same for one of the sibling demos: its own spring constants, mass 0.5 and Provot pass (E:97-102, S:80-85). */
int oc_default_params_for(oc_params* p, int nx, int ny, int integrator);

/* Allocate device state and initialise X = X_last = the flat sheet of InitGL (V:249-260); the
 * spring net of V:286-320 is implicit in (i,j), its rest lengths are derived from the same fp32
 * expressions (V:141-142).  Replaces InitGL's physics part, InitCUDA (H:16-25) and the one-time
 * upload in UploadCUDA (H:53-79). */
int oc_create(oc_cloth** out, const oc_params* p);

/* Change run-time scalars (spring constants, damping, gravity, mass, dt, collider, k, exact,
 * kernel).  nx, ny, batch, band and fullsize are fixed at create and must match.  Replaces editing
 * the globals V:97-104 / V:123-130 (and the per-call scalar arguments of VerletCUDA, H:81). */
int oc_set_params(oc_cloth* c, const oc_params* p);
int oc_get_params(const oc_cloth* c, oc_params* p);

/* Overwrite the state with host arrays (stride_floats 3 or 4), batch*rows*nx particles of this
 * handle's band, cloth-major then row-major.  Replaces UploadCUDA(positions, positions_old, size)
 * (H:53) / UploadOpenCL (LH:155). */
int oc_upload(oc_cloth* c, const float* X, const float* X_last, int stride_floats);
/* (Pinned host arrays make the copies asynchronous.  For a whole single cloth the sequence oc_upload, oc_step(c, 1),
 * oc_download is pipelined in row chunks inside the library — H2D of later rows, the step of the middle ones and D2H
 * of the early ones overlap — so a host-resident simulation pays PCIe once, not twice, per step.) */

/* n x StepPhysics(dt) (V:557-562; the body of the OnIdle step V:548-552 and of VerletCUDA H:81-100).
 * Asynchronous on the handle's stream. */
int oc_step(oc_cloth* c, int n);

/* Wait for all queued work (the reference's cudaThreadSynchronize H:96 / clFinish LH:213). */
int oc_sync(oc_cloth* c);

/* Synchronise and copy the state to host arrays (either pointer may be NULL).  Replaces the mapped
 * VBO read (C:603-605) / ReadBuffer (LH:219). */
int oc_download(oc_cloth* c, float* X, float* X_last, int stride_floats);

/* Render hand-off.  oc_download(..., stride 4) is the float4 (x,y,z,1) vertex buffer of the
 * reference's GPU back ends (pos_vbo, verlet_kernel.cu:193).  oc_download_normals adds the per-vertex normals the
 * reference's lit demo computes (UpdateNormals, OpenCloth_ExplicitEuler_TextureMapped_Lit/.../main.cpp:684-707,
 * over its triangle list :313-327): cross(p2-p1, p3-p1)/3 of every triangle summed into its vertices in list order,
 * then normalised — history-free (the reference keeps adding onto the previous frame's normals; this is its first
 * call).  stride 3 or 4 floats per vertex (w = 0); whole-cloth handles. */
int oc_download_normals(oc_cloth* c, float* N, int stride_floats);

/* ShutdownCUDA (H:27) / OnShutdown (V:418-424). */
void oc_destroy(oc_cloth* c);

/* Text of the last error on this thread ("" if none). */
const char* oc_last_error(void);

/* ---- interaction write-back (mouse drag, V:203-208): X[idx] = X_last[idx] = xyz ---------------- */
int oc_set_particle(oc_cloth* c, int cloth, int idx, const float xyz[3]);
/* n write-backs in one launch — the per-environment actions of a batch: particle idx[k] of cloth cloth[k] <- xyz[3k..3k+2] */
int oc_set_particles(oc_cloth* c, int n, const int* cloth, const int* idx, const float* xyz);
/* oc_upload / oc_download for ONE cloth of a batch (reset / observe one environment); arrays of rows*nx particles */
int oc_upload_cloth(oc_cloth* c, int cloth, const float* X, const float* X_last, int stride_floats);
int oc_download_cloth(oc_cloth* c, int cloth, float* X, float* X_last, int stride_floats);

/* ---- run-time pin set ----------------------------------------------------------------------
 * The reference pins particles 0 and numX by literal index tests (V:455 no gravity term, V:479-482 no spring force
 * applied; the Provot pass leaves them where they are, V:498-501).  oc_set_pins replaces that set for one cloth of
 * the batch (cloth = -1: every cloth) with the n linear indices idx[] (n = 0: nothing pinned); "pinned" keeps the
 * reference's meaning.  A dragged particle (oc_set_particle) that should stay where it is put is pinned the same way.
 * Rows holding a custom pin are stepped on the kernels' general (edge-row) path.  oc_reset_pins returns to the
 * reference's two corners.  Row bands: the indices are those of the whole cloth; call it on every band. */
int oc_set_pins(oc_cloth* c, int cloth, const int* idx, int n);
int oc_reset_pins(oc_cloth* c);

/* ---- stream / timing plumbing (the host side owns streams; torch passes its current stream) --- */
/* Launch on the caller's stream (cudaStream_t).  NULL is CUDA's legacy default stream, as everywhere in the runtime
 * API (so torch's default stream, whose handle is 0, binds correctly).  oc_reset_stream returns to the handle's own
 * non-blocking stream, which is what a new handle uses.  Both wait for the work queued so far. */
int oc_set_stream(oc_cloth* c, void* cuda_stream);
int oc_reset_stream(oc_cloth* c);
/* kernels launched by this handle since create (bench.py's gpu_launches) */
long long oc_launch_count(const oc_cloth* c);
/* time n substeps on the device with CUDA events on the handle's stream; returns milliseconds */
int oc_step_timed(oc_cloth* c, int n, float* ms);

/* ---- row-band halo plumbing (multi-GPU) ----------------------------------
 * A band handle stores rows [row_begin-halo_rows, row_end+halo_rows) clipped to [0,ny).  After the
 * neighbours' halo rows are current, oc_step(c, n) with n <= halo_rows/2 advances the band by n
 * substeps, recomputing a shrinking part of the halo (communication-avoiding: one exchange per
 * halo_rows/2 substeps).  The exchange itself is done by the host: device pointers of the rows to
 * send and of the halo rows to fill are exposed here, each a contiguous run of
 * rows*nx float4 (x,y,z,w).  Two buffers (X at t and X at t-1) are exchanged.
 *   side: 0 = towards row 0 (upper neighbour), 1 = towards row ny-1 (lower neighbour)
 *   which: 0 = X(t) buffer, 1 = X(t-1) buffer
 * Returns the pointer in *dev_ptr and the element count (float4s) in *count (0 if that side is
 * the cloth boundary). */
int oc_halo_send_region(oc_cloth* c, int side, int which, void** dev_ptr, size_t* count);
int oc_halo_recv_region(oc_cloth* c, int side, int which, void** dev_ptr, size_t* count);
/* tell the handle that the halo rows are current again (resets the shrink counter) */
int oc_halo_refreshed(oc_cloth* c);
/* substeps that can still be taken before the next exchange is required */
int oc_halo_budget(const oc_cloth* c);
/* Like oc_step, for overlapping the halo exchange with compute.  If the n-th substep exhausts the halo budget
 * (it then covers exactly the owned rows), it is issued as: the first and last halo_rows owned rows — the rows
 * the neighbours are waiting for —, an event, then the interior; `exchange_stream` (cudaStream_t) is made to wait
 * for that event, so an exchange enqueued on it runs concurrently with the interior launch.  *did_split tells
 * whether the split happened (it does not for k > 1 launches or very small bands). */
int oc_step_split(oc_cloth* c, int n, void* exchange_stream, int* did_split);
/* Single-process exchange: bands[0..n) are the bands of ONE cloth in row order (same process, same
 * or different devices).  Copies every send region into the neighbour's halo (cudaMemcpyPeerAsync
 * across devices, i.e. NVLink on an HGX box), ordered after each band's queued work and before its
 * next step, without blocking the host, then marks all bands refreshed. */
int oc_halo_exchange(oc_cloth* const* bands, int n);

/* ---- linked row bands: halo rows pushed by the kernel over peer memory (NVLink), no exchange step ----------------
 * Multi-GPU path of choice.  One band handle per GPU (oc_params.row_begin/row_end,
 * halo_rows >= 2), in one process or one process per GPU.  Linking maps each neighbour's position buffers and a small
 * array of flag words into this handle (CUDA IPC across processes, peer access inside one).  From then on oc_step on
 * a linked band computes exactly its owned rows every substep; the tiles at the band edge store the two rows the
 * neighbour's bend springs reach (V:311, V:317) straight into the neighbour's halo and release one flag word per
 * column strip in the neighbour's memory, and the neighbour's edge tiles of the next substep wait for those words.
 * No host round trip, no exchange period, no recomputed rows, and the chained launches of a band never break.
 * Every band must take the same steps (same n, same order) — the flag words count linked substeps.
 *
 * Protocol (the caller provides the barriers when the bands live in different processes):
 *   1. every band: oc_sync, oc_band_endpoint(c, blob)              -> exchange the blobs (any transport)
 *   2. [barrier]   every band: oc_band_link(c, upper blob or NULL, lower blob or NULL), oc_band_pull_halo, oc_sync
 *   3. [barrier]   oc_step(c, n) on every band, freely; oc_sync / oc_download as usual
 * After oc_upload / oc_set_particle on any band (call oc_set_particle on EVERY band: each updates the copies it
 * stores): every band oc_sync, [barrier], oc_band_pull_halo, oc_sync, [barrier].
 * oc_band_link_local does steps 1-3 for bands living in one process (the C++ harness; tests).
 * A band that waits more than 30 s for a neighbour makes oc_sync / oc_step / oc_download fail with OC_ERR_CUDA. */
#define OC_BAND_ENDPOINT_BYTES 512
int oc_band_endpoint(oc_cloth* c, void* blob, size_t bytes);
int oc_band_link(oc_cloth* c, const void* upper_blob, const void* lower_blob);
int oc_band_pull_halo(oc_cloth* c);
int oc_band_unlink(oc_cloth* c);
int oc_band_link_local(oc_cloth* const* bands, int n);

/* ---- environment switches (development / measurement; read by the library, defaults in brackets) -------------
 *   OC_PDL=0          launch oc_k_march2 without programmatic stream serialization              [on]
 *   OC_TILE_DEPS=0    consecutive steps wait for the whole previous grid instead of per-tile flags [on]
 *   OC_MARCH2_EDGE=p  un-chained single-wave launches: interior/edge rows per segment - 1, percent [24 exact, 10 fast]
 *   OC_MARCH_RS=n, OC_MARCH_TW=n, OC_MARCH2_WC=n   force rows per segment / window width of the marching kernels
 *   OC_DEBUG=bits     4: count fallbacks (oc_debug_counters)  8: CTA time line (oc_debug_timeline)
 *                     16: print the launch plan to stderr  (1, 2: force / suppress the fallback of oc_k_march)
 *                     32: test hook, tile dependencies time out  64: time the upload/step/download pipeline (oc_debug_pipeline)
 * None of them changes a result. */

/* ---- diagnostics ----------------------------------------------------------------------------- */
/* Spring energy  sum 1/2 Ks (|p1-p2| - rest)^2  over the reference's spring list (duplicated edge
 * bend springs included), reduced in double on the device; whole-cloth handles only. */
int oc_spring_energy(oc_cloth* c, int cloth, double* energy);
/* Device self-test of the branch-free IEEE sequences the exact-mode kernels use for sqrt, 1/x and
 * a/b: n random operands (seeded) inside the accepted exponent ranges are compared with the
 * correctly rounded intrinsics (sqrt.rn, rcp.rn, div.rn); *mismatches receives the number of
 * differing results (must be 0). */
int oc_selftest_math(unsigned long long n, unsigned int seed, unsigned long long* mismatches);
/* Development counters (only counted when the environment has OC_DEBUG=4 at oc_create): lanes and warps
 * that took the IEEE-intrinsic fallback of the exact-mode spring phase, velocity fallbacks; reset on read. */
int oc_debug_counters(oc_cloth* c, unsigned long long out[4]);
/* out[0], out[1]: lanes / warps that redid their springs with the IEEE intrinsics; out[2]: low 40 bits velocity
 * fallbacks, bits 40..: tile-dependency waits that timed out (must be 0); out[3]: low 32 bits threads, high 32 bits
 * warps that resolved a collider contact (oc_k_march2) */
/* Development CTA time line of the last oc_k_march2 launch (only recorded with OC_DEBUG=8 at oc_create): 8 words
 * per CTA (linear index x + gridDim.x * (y + gridDim.y * z), first 4096 CTAs): %globaltimer at entry, set-up
 * done, lead-in done, steady loop done, exit; SM id; 2 unused.  Copies min(n_words, 8*4096) words. */
int oc_debug_timeline(oc_cloth* c, unsigned long long* out, size_t n_words);
/* Development time line of the host <-> device pipeline (only recorded with OC_DEBUG=64 at oc_create): out[5*chunk+s] =
 * milliseconds after the start of the last oc_upload at which chunk's H2D copy (s=0), unpack (1), step (2), pack (3) and
 * D2H copy (4) finished.  Returns the number of chunks. */
int oc_debug_pipeline(oc_cloth* c, float* out, int max_chunks);
/* sizeof(oc_params) as the library was compiled, so that FFI bindings can verify their mirror */
size_t oc_sizeof_params(void);
/* library / device info string: "opencloth_b200 abi=1 sm=100 device=NVIDIA B200 ..." */
const char* oc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OPENCLOTH_H */
