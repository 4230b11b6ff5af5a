// oc_api.cu — the C-ABI of include/opencloth.h: handle, device memory, launches.
//
// Host side of the hot path.  Replaces, for the Verlet demo of the reference
// (/root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp, "V:"), the physics part of InitGL
// (V:249-327), the OnIdle step (V:548-552) and OnShutdown (V:418-424); and for the reference's own
// GPU back end (…/OpenCloth_Verlet_CUDA/verlet.cu, "H:") InitCUDA/UploadCUDA/VerletCUDA/ShutdownCUDA
// (H:16-100).  No torch types, no CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/opencloth.h"
#include "oc_core.cuh"
#include "oc_host.h"
#include "oc_gather.cuh"
#include "oc_provot.cuh"
#include "oc_normals.cuh"
#include "oc_resident.cuh"
#include "oc_bandres.cuh"
#include "oc_march.cuh"
#include "oc_march2.cuh"
#include "oc_twin.cuh"
#include "oc_stream.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <new>

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int oc_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define OC_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return oc_fail((e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)        \
                               ? OC_ERR_NO_DEVICE : (e_ == cudaErrorMemoryAllocation ? OC_ERR_NOMEM : OC_ERR_CUDA), \
                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* oc_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct oc_cloth {
    oc_params p;
    OcConst   k;                 // kernel constants (device pointers inside)
    int       dev;
    int       sm_count;
    float4*   buf[4];            // rotating position buffers
    OcSeq     q;                 // buffer rotation + band shrink counter (oc_host.h)
    float*    tables;            // device: xs[U] zs[V] rh1[U] rh2[U] dx2[U] rv1[V] rv2[V] dz2[V]
    float*    d_xs; float* d_zs;
    long long stored;            // float4 elements per buffer (batch * srows * U)
    int       rows_own;          // row_end - row_begin
    cudaStream_t own_stream, stream;
    cudaEvent_t  ev0, ev1;
    cudaEvent_t  ev_ready, ev_filled;   // band exchange: my rows are final / my halo has been written
    long long launches;
    float*    stage[2];          // device staging for upload/download
    size_t    stage_bytes;
    cudaStream_t s_in, s_out;    // copy streams of the host <-> device pipeline (upload_impl): copies only, so that the
    cudaStream_t s_un, s_pk;     // link never waits for a kernel; the unpack / pack kernels have streams of their own
    struct {
        cudaEvent_t ev_tail, ev_cp[32], ev_in[32], ev_step[32], ev_pk[32];
        cudaEvent_t ev_t0, ev_out[32];      // development (OC_DEBUG=64): start of an upload, D2H of a chunk done
        bool timed; int last_chunks;
        int up_chunks;           // an upload is queued in this many row chunks (events ev_in); 0: none pending
        int step_chunks;         // the last substep was launched per chunk (events ev_step)
    } pipe;
    double*   d_energy;
    unsigned long long* d_dbg;   // development counters (OC_DEBUG & 4)
    OcChain2  chain;             // tile-level dependencies between consecutive oc_k_march2 launches (oc_march2.cuh)
    unsigned* h_err;             // sticky error word written by the kernels (pinned, mapped): see OcConst::err
    // linked row bands (OcPeer2): the neighbours' buffers and flag words, mapped into this process / device
    struct {
        bool      on;
        bool      has[2];               // [0] upper, [1] lower neighbour
        float4*   buf[2][4];            // the neighbour's four position buffers
        unsigned* flags_out[2];         // the neighbour's incoming-flag array for my side
        int       row_lo[2];            // global row held in the neighbour's storage row 0
        int       row_begin[2], row_end[2];
        void*     opened[2][5];         // cudaIpcOpenMemHandle mappings to close (nullptr: same process)
        unsigned  epoch;                // linked steps taken
        int       rev;                  // this band launches its segments bottom to top (OcPeer2::rev): odd bands of the chain
    } link;
    unsigned* in_flags;          // device: 2 x OC_LINK_STRIPS words released by the neighbours' boundary tiles
    // oc_k_bandres (mid-size cloths resident in shared memory, one row band per CTA): exchange rows of two parities, one
    // flag word per band, substeps taken so far (the flags count on from launch to launch); allocated at the first use
    struct { void* ex; unsigned* flags; unsigned epoch; int coop; } bres;
    // run-time pin sets (oc_set_pins): bitmap over batch x ny x nx particles + per-row summary; empty = reference default
    std::vector<unsigned>* h_pins;
    std::vector<unsigned char>* h_pin_rows;
    unsigned* d_pins;
    unsigned char* d_pin_rows;
};
#define OC_LINK_STRIPS 4096
#define OC_CHAIN_CAP (1 << 16)
// anything that writes the state other than the chained kernel itself, or reorders the stream, breaks the chain
static inline void chain_break(oc_cloth* c) { c->chain.valid = false; }

static int free_handle(oc_cloth* c)
{
    if (!c) return OC_OK;
    for (int i = 0; i < 4; ++i) if (c->buf[i]) cudaFree(c->buf[i]);
    if (c->tables) cudaFree(c->tables);
    if (c->stage[0]) cudaFree(c->stage[0]);
    if (c->stage[1]) cudaFree(c->stage[1]);
    if (c->d_energy) cudaFree(c->d_energy);
    if (c->d_dbg) cudaFree(c->d_dbg);
    if (c->chain.flags) cudaFree(c->chain.flags);
    if (c->in_flags) cudaFree(c->in_flags);
    if (c->bres.ex) cudaFree(c->bres.ex);
    if (c->bres.flags) cudaFree(c->bres.flags);
    if (c->d_pins) cudaFree(c->d_pins);
    if (c->d_pin_rows) cudaFree(c->d_pin_rows);
    delete c->h_pins; delete c->h_pin_rows;
    if (c->h_err) cudaFreeHost(c->h_err);
    for (int sd = 0; sd < 2; ++sd) for (int b = 0; b < 5; ++b) if (c->link.opened[sd][b]) cudaIpcCloseMemHandle(c->link.opened[sd][b]);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_filled) cudaEventDestroy(c->ev_filled);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_un) cudaStreamDestroy(c->s_un);
    if (c->s_pk) cudaStreamDestroy(c->s_pk);
    if (c->pipe.ev_tail) cudaEventDestroy(c->pipe.ev_tail);
    for (int k = 0; k < 32; ++k) {
        if (c->pipe.ev_in[k]) cudaEventDestroy(c->pipe.ev_in[k]);
        if (c->pipe.ev_step[k]) cudaEventDestroy(c->pipe.ev_step[k]);
        if (c->pipe.ev_cp[k]) cudaEventDestroy(c->pipe.ev_cp[k]);
        if (c->pipe.ev_pk[k]) cudaEventDestroy(c->pipe.ev_pk[k]);
    }
    delete c;
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// defaults: the reference's globals (oc_host.h)
// ------------------------------------------------------------------------------------------------
extern "C" int oc_default_params(oc_params* p, int nx, int ny)
{
    if (!p) return oc_fail(OC_ERR_INVALID, "oc_default_params: null");
    oc_host_default_params(p, nx, ny);
    return OC_OK;
}

extern "C" int oc_default_params_for(oc_params* p, int nx, int ny, int integrator)
{
    if (!p) return oc_fail(OC_ERR_INVALID, "oc_default_params_for: null");
    if (integrator < OC_INTEGRATOR_VERLET || integrator > OC_INTEGRATOR_SEMI_IMPLICIT) return oc_fail(OC_ERR_INVALID, "bad integrator id %d", integrator);
    oc_host_default_params_for(p, nx, ny, integrator);
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// small kernels: initial sheet, pack / unpack, particle write-back, energy
// ------------------------------------------------------------------------------------------------
__global__ void oc_k_init(OcConst c, const float* __restrict__ xs, const float* __restrict__ zs, float y,
                          float4* __restrict__ A, float4* __restrict__ B)
{
    long long per = (long long)c.srows * c.U;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per * c.batch) return;
    long long r = t % per;
    int j = (int)(r / c.U) + c.row_lo, i = (int)(r % c.U);
    float4 v = make_float4(xs[i], y, zs[j], oc_u2f(OC_W_PLAIN));       // V:256-257
    A[t] = v;
    B[t] = c.integ == 0 ? v : make_float4(0.0f, 0.0f, 0.0f, oc_u2f(OC_W_PLAIN));      // X_last = X (V:259), or V = 0 (E:270)
}

// host layout (stride 3 or 4; cloths [hcloth0, ...) x owned rows, row-major) <-> float4 storage (owned rows inside the
// halo'd store).  One launch covers rows [row0, row0 + rows) of cloths [cloth0, cloth0 + ncloth): the copies are cut
// into row chunks that travel over PCIe while their neighbours are unpacked, stepped or packed (oc_upload).
struct OcXfer { int cloth0, hcloth0, ncloth, row_own0, rows_own, row0, rows, stride; };
__global__ void oc_k_unpack(OcConst c, OcXfer x, const float* __restrict__ X, const float* __restrict__ XL,
                            float4* __restrict__ A, float4* __restrict__ B)
{
    const long long per = (long long)x.rows * c.U;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per * x.ncloth) return;
    const int b = x.cloth0 + (int)(t / per); const long long r = t % per;
    const int j = (int)(r / c.U) + x.row0, i = (int)(r % c.U);
    const long long o = oc_index(c, b, i, j);
    const long long h = (((long long)(b - x.hcloth0) * x.rows_own + (j - x.row_own0)) * c.U + i) * x.stride;
    A[o] = make_float4(X[h], X[h + 1], X[h + 2], oc_u2f(OC_W_PLAIN));
    B[o] = make_float4(XL[h], XL[h + 1], XL[h + 2], oc_u2f(OC_W_PLAIN));
}
__global__ void oc_k_pack(OcConst c, OcXfer x, const float4* __restrict__ A, const float4* __restrict__ B,
                          float* __restrict__ X, float* __restrict__ XL)
{
    const long long per = (long long)x.rows * c.U;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per * x.ncloth) return;
    const int b = x.cloth0 + (int)(t / per); const long long r = t % per;
    const int j = (int)(r / c.U) + x.row0, i = (int)(r % c.U);
    const long long o = oc_index(c, b, i, j);
    const long long h = (((long long)(b - x.hcloth0) * x.rows_own + (j - x.row_own0)) * c.U + i) * x.stride;
    const float4 a = A[o];
    if (X) { X[h] = a.x; X[h + 1] = a.y; X[h + 2] = a.z; if (x.stride == 4) X[h + 3] = 1.0f; }
    if (XL) {
        const float4 q = oc_hit(a.w) ? a : B[o];                        // V:530 vs V:438 (Euler integrators: B is V, never flagged)
        XL[h] = q.x; XL[h + 1] = q.y; XL[h + 2] = q.z; if (x.stride == 4) XL[h + 3] = 1.0f;
    }
}

struct OcPoke { long long o; float x, y, z; int pad; };
__global__ void oc_k_set_particles(float4* A, float4* B, const OcPoke* __restrict__ p, int n, int xv)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const OcPoke q = p[t];
    const float4 v = make_float4(q.x, q.y, q.z, oc_u2f(OC_W_PLAIN));
    A[q.o] = v; B[q.o] = xv ? make_float4(0.0f, 0.0f, 0.0f, oc_u2f(OC_W_PLAIN)) : v;
}

__global__ void oc_k_set_particle(float4* A, float4* B, long long o, float x, float y, float z, int xv)
{
    float4 v = make_float4(x, y, z, oc_u2f(OC_W_PLAIN));                // V:203-208: X_last = X; Euler demos E:201-210: V = 0
    A[o] = v; B[o] = xv ? make_float4(0.0f, 0.0f, 0.0f, oc_u2f(OC_W_PLAIN)) : v;
}

// sum over springs of 1/2 Ks (|p1-p2| - rest)^2 in double; every particle owns its forward springs
__device__ double oc_e_term(const OcConst& c, const float4* A, int b, int i, int j, int ni, int nj, float rest, float ks)
{
    float4 p = A[oc_index(c, b, i, j)], q = A[oc_index(c, b, ni, nj)];
    float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    double len = sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
    double ext = len - (double)rest;
    return 0.5 * (double)ks * ext * ext;
}
__global__ void oc_k_energy(OcConst c, const float4* __restrict__ A, int b, float ks_struct, float ks_shear, float ks_bend,
                            double* __restrict__ out)
{
    long long per = (long long)c.V * c.U;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (t < per) {
        int j = (int)(t / c.U), i = (int)(t % c.U);
        const int U = c.U, V = c.V;
        if (i + 1 < U) e += oc_e_term(c, A, b, i, j, i + 1, j, c.rh1[i], ks_struct);
        if (j + 1 < V) e += oc_e_term(c, A, b, i, j, i, j + 1, c.rv1[j], ks_struct);
        if (i + 1 < U && j + 1 < V) {
            float r = __fsqrt_rn(__fadd_rn(c.dx2[i], c.dz2[j]));
            e += oc_e_term(c, A, b, i, j, i + 1, j + 1, r, ks_shear);
            e += oc_e_term(c, A, b, i, j + 1, i + 1, j, r, ks_shear);
        }
        if (i + 2 < U) e += oc_e_term(c, A, b, i, j, i + 2, j, c.rh2[i], ks_bend) * (i == U - 3 ? 2.0 : 1.0);
        if (j + 2 < V) e += oc_e_term(c, A, b, i, j, i, j + 2, c.rv2[j], ks_bend) * (j == V - 3 ? 2.0 : 1.0);
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    __shared__ double s[8];
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s[w] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) tot += s[q];
        atomicAdd(out, tot);
    }
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
static int validate(const oc_params* p)
{
    if (!p) return oc_fail(OC_ERR_INVALID, "null params");
    if (p->nx < 3 || p->ny < 3) return oc_fail(OC_ERR_INVALID, "nx, ny must be >= 3 (bend springs reach 2; got %d x %d)", p->nx, p->ny);
    if (p->batch < 1) return oc_fail(OC_ERR_INVALID, "batch must be >= 1");
    if (p->row_begin != 0 || p->row_end != 0) {
        if (p->row_begin < 0 || p->row_end > p->ny || p->row_end <= p->row_begin)
            return oc_fail(OC_ERR_INVALID, "bad row band [%d,%d) of %d rows", p->row_begin, p->row_end, p->ny);
        bool sub = p->row_begin > 0 || p->row_end < p->ny;
        if (sub) {
            if (p->batch != 1) return oc_fail(OC_ERR_UNSUPPORTED, "row bands need batch == 1");
            if (p->halo_rows < 2 || (p->halo_rows & 1)) return oc_fail(OC_ERR_INVALID, "halo_rows must be an even number >= 2");
            if (p->row_end - p->row_begin < p->halo_rows) return oc_fail(OC_ERR_INVALID, "band has fewer rows than halo_rows");
        }
    }
    if (!(p->dt > 0.0f) || !(p->mass > 0.0f)) return oc_fail(OC_ERR_INVALID, "dt and mass must be positive");
    if (p->substeps_per_launch < 0 || p->substeps_per_launch > OC_MARCH_MAX_STAGES)
        return oc_fail(OC_ERR_INVALID, "substeps_per_launch must be 0..%d", OC_MARCH_MAX_STAGES);
    if (p->kernel < OC_KERNEL_AUTO || p->kernel > OC_KERNEL_BANDRES) return oc_fail(OC_ERR_INVALID, "bad kernel id");
    if (p->integrator < OC_INTEGRATOR_VERLET || p->integrator > OC_INTEGRATOR_SEMI_IMPLICIT) return oc_fail(OC_ERR_INVALID, "bad integrator id %d", p->integrator);
    if (p->provot != 0 && p->provot != 1) return oc_fail(OC_ERR_INVALID, "provot must be 0 or 1");
    if ((p->integrator != OC_INTEGRATOR_VERLET || p->provot) && (p->row_begin != 0 || p->row_end != 0) && (p->row_begin > 0 || p->row_end < p->ny))
        return oc_fail(OC_ERR_UNSUPPORTED, "the Euler integrators and the Provot pass need a whole-cloth handle (no row bands)");
    return OC_OK;
}

extern "C" int oc_create(oc_cloth** out, const oc_params* p)
{
    if (!out) return oc_fail(OC_ERR_INVALID, "oc_create: null out");
    *out = nullptr;
    int rc = validate(p);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return oc_fail(OC_ERR_NO_DEVICE, "no CUDA device (%s); libopencloth_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    oc_cloth* c = new (std::nothrow) oc_cloth();
    if (!c) return oc_fail(OC_ERR_NOMEM, "host allocation failed");
    memset(c, 0, sizeof(*c));
    c->p = *p;
#define OC_CREATE_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { free_handle(c); \
        return oc_fail(e_ == cudaErrorMemoryAllocation ? OC_ERR_NOMEM : OC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
    if (p->device >= 0) c->dev = p->device; else { OC_CREATE_CUDA(cudaGetDevice(&c->dev)); }
    if (c->dev >= ndev) { free_handle(c); return oc_fail(OC_ERR_INVALID, "device %d of %d", p->device, ndev); }
    OC_CREATE_CUDA(cudaSetDevice(c->dev));
    OC_CREATE_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->dev));

    const int U = p->nx, V = p->ny;
    OcConst& k = c->k;
    oc_host_geometry(c->p, k, c->q);
    c->rows_own = c->p.row_end - c->p.row_begin;
    c->stored = k.cloth_stride * p->batch;
    oc_host_derive_scalars(c->p, k);

    // rest-length tables from the initial sheet (V:254-260 positions, V:141-142 rest lengths)
    OcHostTables T;
    oc_host_build_tables(U, V, p->fullsize, T);
    OC_CREATE_CUDA(cudaMalloc(&c->tables, T.t.size() * sizeof(float)));
    OC_CREATE_CUDA(cudaMemcpy(c->tables, T.t.data(), T.t.size() * sizeof(float), cudaMemcpyHostToDevice));
    c->d_xs = c->tables + T.xs; c->d_zs = c->tables + T.zs;
    oc_host_bind_tables(k, c->tables, T);

    for (int b = 0; b < 4; ++b) OC_CREATE_CUDA(cudaMalloc(&c->buf[b], (size_t)c->stored * sizeof(float4)));
    OC_CREATE_CUDA(cudaMalloc(&c->d_energy, sizeof(double)));
    OC_CREATE_CUDA(cudaMalloc(&c->d_dbg, OC_DBG_WORDS * sizeof(unsigned long long)));
    OC_CREATE_CUDA(cudaMemset(c->d_dbg, 0, OC_DBG_WORDS * sizeof(unsigned long long)));
    k.dbg_cnt = c->d_dbg;
    OC_CREATE_CUDA(cudaMalloc(&c->chain.flags, OC_CHAIN_CAP * sizeof(unsigned)));
    OC_CREATE_CUDA(cudaMemset(c->chain.flags, 0, OC_CHAIN_CAP * sizeof(unsigned)));
    c->chain.cap = OC_CHAIN_CAP; c->chain.epoch = 0; c->chain.valid = false; c->chain.kind = 0;
    OC_CREATE_CUDA(cudaMalloc(&c->in_flags, 2 * OC_LINK_STRIPS * sizeof(unsigned)));
    OC_CREATE_CUDA(cudaMemset(c->in_flags, 0, 2 * OC_LINK_STRIPS * sizeof(unsigned)));
    OC_CREATE_CUDA(cudaHostAlloc(&c->h_err, sizeof(unsigned), cudaHostAllocMapped));
    *c->h_err = 0u;
    OC_CREATE_CUDA(cudaHostGetDevicePointer(&k.err, c->h_err, 0));
    OC_CREATE_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    OC_CREATE_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    OC_CREATE_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    OC_CREATE_CUDA(cudaStreamCreateWithFlags(&c->s_un, cudaStreamNonBlocking));
    OC_CREATE_CUDA(cudaStreamCreateWithFlags(&c->s_pk, cudaStreamNonBlocking));
    OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_tail, cudaEventDisableTiming));
    c->pipe.timed = (k.dbg & 64) != 0;
    const unsigned evf = c->pipe.timed ? cudaEventDefault : cudaEventDisableTiming;
    OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_t0, evf));
    for (int q = 0; q < 32; ++q) {
        OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_in[q], evf));
        OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_step[q], evf));
        OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_cp[q], evf));
        OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_pk[q], evf));
        OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->pipe.ev_out[q], evf));
    }
    OC_CREATE_CUDA(cudaEventCreate(&c->ev0));
    OC_CREATE_CUDA(cudaEventCreate(&c->ev1));
    OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    OC_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_filled, cudaEventDisableTiming));
    {
        long long n = c->stored;
        oc_k_init<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(k, c->d_xs, c->d_zs, p->fullsize + 1, c->buf[0], c->buf[1]);
        c->launches++;
        OC_CREATE_CUDA(cudaGetLastError());
        // the two spare buffers start as copies so that never-computed halo rows hold finite values
        OC_CREATE_CUDA(cudaMemcpyAsync(c->buf[2], c->buf[0], (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        OC_CREATE_CUDA(cudaMemcpyAsync(c->buf[3], c->buf[0], (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    }
    rc = oc_march_configure(c->dev);
    if (rc == 0) rc = oc_march2_configure(c->dev);
    if (rc == 0) rc = oc_twin_configure(c->dev);
    // function attributes are per device: set them at every create, for the device of this handle
    if (rc == 0) rc = (int)cudaFuncSetAttribute((const void*)&oc_k_resident<MathExact>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OcResidentSmem::bytes(OC_RESIDENT_MAX_PARTICLES));
    if (rc == 0) rc = (int)cudaFuncSetAttribute((const void*)&oc_k_resident<MathFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OcResidentSmem::bytes(OC_RESIDENT_MAX_PARTICLES));
    if (rc == 0) rc = (int)cudaFuncSetAttribute((const void*)&oc_k_bandres<MathExact, OC_BANDRES_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, OC_BANDRES_SMEM_MAX);
    if (rc == 0) rc = (int)cudaFuncSetAttribute((const void*)&oc_k_bandres<MathFast, OC_BANDRES_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, OC_BANDRES_SMEM_MAX);
    if (rc == 0) rc = (int)cudaDeviceGetAttribute(&c->bres.coop, cudaDevAttrCooperativeLaunch, c->dev);
    if (rc != 0) { free_handle(c); return oc_fail(OC_ERR_CUDA, "oc_march_configure failed: %s", cudaGetErrorString((cudaError_t)rc)); }
#undef OC_CREATE_CUDA
    *out = c;
    return OC_OK;
}

extern "C" void oc_destroy(oc_cloth* c)
{
    if (!c) return;
    cudaSetDevice(c->dev);
    cudaStreamSynchronize(c->stream);
    free_handle(c);
}

extern "C" int oc_set_params(oc_cloth* c, const oc_params* p)
{
    if (!c || !p) return oc_fail(OC_ERR_INVALID, "oc_set_params: null");
    oc_params q = *p;
    if (q.row_begin == 0 && q.row_end == 0) q.row_end = q.ny;
    if (q.integrator != c->p.integrator) return oc_fail(OC_ERR_INVALID, "oc_set_params: the integrator is fixed at oc_create (it decides what the second state buffer holds)");
    if (q.nx != c->p.nx || q.ny != c->p.ny || q.batch != c->p.batch || q.fullsize != c->p.fullsize ||
        q.row_begin != c->p.row_begin || q.row_end != c->p.row_end || (c->q.band && q.halo_rows != c->p.halo_rows))
        return oc_fail(OC_ERR_INVALID, "oc_set_params: nx, ny, batch, row band, halo_rows and fullsize are fixed at oc_create");
    int rc = validate(&q);
    if (rc) return rc;
    q.device = c->p.device; q.halo_rows = c->p.halo_rows;
    c->p = q;
    oc_host_derive_scalars(c->p, c->k);
    return OC_OK;
}

extern "C" int oc_get_params(const oc_cloth* c, oc_params* p)
{
    if (!c || !p) return oc_fail(OC_ERR_INVALID, "oc_get_params: null");
    *p = c->p;
    return OC_OK;
}

static int bind_stream(oc_cloth* c, cudaStream_t s)
{
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = s;
    chain_break(c);
    return OC_OK;
}
extern "C" int oc_set_stream(oc_cloth* c, void* s)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_set_stream: null");
    return bind_stream(c, (cudaStream_t)s);          // NULL is the legacy default stream, like everywhere in CUDA
}
extern "C" int oc_reset_stream(oc_cloth* c)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_reset_stream: null");
    return bind_stream(c, c->own_stream);
}

// a kernel that gave up on a tile dependency has written why into the handle's error word and trapped
static int check_err_word(const oc_cloth* c)
{
    const unsigned e = c->h_err ? *(volatile unsigned*)c->h_err : 0u;
    if (e == 0u) return OC_OK;
    return oc_fail(OC_ERR_CUDA, e == 2u ? "a tile waited more than 30 s for a boundary tile of a neighbour GPU (linked row bands): a band fell behind, failed, or took a different number of steps"
                                        : "a tile dependency between consecutive launches timed out; the state is invalid");
}

extern "C" long long oc_launch_count(const oc_cloth* c) { return c ? c->launches : 0; }

extern "C" int oc_sync(oc_cloth* c)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_sync: null");
    OC_CUDA(cudaSetDevice(c->dev));
    cudaError_t e = cudaStreamSynchronize(c->s_in);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_un);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_pk);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_out);
    c->pipe.up_chunks = 0;                 // everything has landed: nothing left to wait for chunk by chunk
    const int rc = check_err_word(c);
    if (rc) return rc;
    OC_CUDA(e);
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// upload / download / set_particle
// ------------------------------------------------------------------------------------------------
static int ensure_stage(oc_cloth* c, size_t bytes)
{
    if (c->stage_bytes >= bytes) return OC_OK;
    OC_CUDA(cudaStreamSynchronize(c->s_in));
    OC_CUDA(cudaStreamSynchronize(c->s_un));
    OC_CUDA(cudaStreamSynchronize(c->s_pk));
    OC_CUDA(cudaStreamSynchronize(c->s_out));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->stage[0]) cudaFree(c->stage[0]);
    if (c->stage[1]) cudaFree(c->stage[1]);
    c->stage[0] = c->stage[1] = nullptr; c->stage_bytes = 0;
    OC_CUDA(cudaMalloc(&c->stage[0], bytes));
    OC_CUDA(cudaMalloc(&c->stage[1], bytes));
    c->stage_bytes = bytes;
    return OC_OK;
}

// The host <-> device pipeline.  A cloth that arrives from the host, takes one step and goes back
// (oc_upload, oc_step(1), oc_download — the end-to-end pattern of bench.py and of a host application that owns the
// state) would spend its time on PCIe one direction at a time.  The three calls therefore work on row CHUNKS:
//   oc_upload    per chunk, on a copy stream: H2D of the chunk's X and X_last rows, unpack into the float4 buffers, event
//   oc_step      the first substep after an upload is launched chunk by chunk; chunk c waits for the upload event of
//                chunk c+1 (its stencil reaches two rows into it) and records its own event
//   oc_download  per chunk, on a second copy stream: waits for the chunk's step event, packs, D2H
// so that the H2D of the later chunks, the step of the middle ones and the D2H of the early ones overlap (PCIe is
// full duplex).  Whole single cloths only; other handles, and any other order of calls, take the plain path.
#define OC_PIPE_MAX_CHUNKS 32
static int pipe_chunks(const oc_cloth* c)
{
    static const int env = getenv("OC_PIPE_CHUNKS") ? atoi(getenv("OC_PIPE_CHUNKS")) : 0;
    if (c->p.batch != 1 || c->q.band) return 1;
    int n = env > 0 ? env : 10;            // measured at 2048^2 (tools/run_e2e.sh), ms per round trip: equal chunks 4: 3.15, 8: 3.02, 16: 3.15, 32: 3.21;
                                           // tapered (chunk_rows) 8: 3.10, 10: 2.95, 12: 2.96, 16: 3.11
    if (n > OC_PIPE_MAX_CHUNKS) n = OC_PIPE_MAX_CHUNKS;
    while (n > 1 && c->rows_own / n < 64) n = n > 8 ? 8 : n / 2;          // chunks of 64 rows on average (a tapered end chunk: >= 8)
    return n;
}
// Rows of chunk k.  The chunks are TAPERED: the round trip is the H2D time of the whole cloth plus the time the last chunk
// still needs afterwards (unpack, step, pack, its own D2H), and symmetrically the D2H link idles until the first chunk is
// through; small chunks at both ends shorten that fill and drain, large ones in the middle keep the per-copy overhead low.
// Weights 1, 2, 4, 8, 8, ..., 8, 4, 2, 1 (OC_PIPE_TAPER=0: equal chunks).
static int chunk_weight(int nch, int k)
{
    static const bool taper = !(getenv("OC_PIPE_TAPER") && atoi(getenv("OC_PIPE_TAPER")) == 0);
    if (!taper) return 1;
    const int d = k < nch - 1 - k ? k : nch - 1 - k;          // distance from the nearer end
    return d >= 3 ? 8 : (1 << d);
}
static void chunk_rows(const oc_cloth* c, int nch, int k, int* r0, int* r1)
{
    long long tot = 0, lo = 0, hi = 0;
    for (int j = 0; j < nch; ++j) { const int w = chunk_weight(nch, j); if (j < k) lo += w; if (j <= k) hi += w; tot += w; }
    *r0 = c->p.row_begin + (int)((long long)c->rows_own * lo / tot);
    *r1 = c->p.row_begin + (int)((long long)c->rows_own * hi / tot);
}
// every queued chunk of an upload has landed as far as the compute stream is concerned
static int join_upload(oc_cloth* c)
{
    if (c->pipe.up_chunks > 0) {
        OC_CUDA(cudaStreamWaitEvent(c->stream, c->pipe.ev_in[c->pipe.up_chunks - 1], 0));
        c->pipe.up_chunks = 0;
    }
    c->pipe.step_chunks = 0;
    return OC_OK;
}

static int upload_impl(oc_cloth* c, int cloth0, int ncloth, const float* X, const float* X_last, int stride)
{
    if (stride != 3 && stride != 4) return oc_fail(OC_ERR_INVALID, "stride_floats must be 3 or 4");
    OC_CUDA(cudaSetDevice(c->dev));
    chain_break(c);
    int rc = join_upload(c);
    if (rc) return rc;
    const long long n = (long long)ncloth * c->rows_own * c->p.nx;
    rc = ensure_stage(c, (size_t)c->p.batch * c->rows_own * c->p.nx * 4 * sizeof(float));
    if (rc) return rc;
    const int nch = (ncloth == c->p.batch) ? pipe_chunks(c) : 1;
    // the copy stream starts after everything queued so far (the buffers may still be in use)
    OC_CUDA(cudaEventRecord(c->pipe.ev_tail, c->stream));
    OC_CUDA(cudaStreamWaitEvent(c->s_in, c->pipe.ev_tail, 0));
    if (c->pipe.timed) OC_CUDA(cudaEventRecord(c->pipe.ev_t0, c->s_in));
    c->pipe.last_chunks = nch;
    for (int k = 0; k < nch; ++k) {
        int r0, r1;
        chunk_rows(c, nch, k, &r0, &r1);
        const size_t off = (size_t)(r0 - c->p.row_begin) * c->p.nx * stride;            // floats; nch > 1 only for one cloth
        const size_t cnt = (nch == 1) ? (size_t)n * stride : (size_t)(r1 - r0) * c->p.nx * stride;
        OC_CUDA(cudaMemcpyAsync(c->stage[0] + off, X + off, cnt * sizeof(float), cudaMemcpyHostToDevice, c->s_in));
        OC_CUDA(cudaMemcpyAsync(c->stage[1] + off, X_last + off, cnt * sizeof(float), cudaMemcpyHostToDevice, c->s_in));
        OC_CUDA(cudaEventRecord(c->pipe.ev_cp[k], c->s_in));
        OC_CUDA(cudaStreamWaitEvent(c->s_un, c->pipe.ev_cp[k], 0));
        const OcXfer x = { cloth0, cloth0, ncloth, c->p.row_begin, c->rows_own, r0, r1 - r0, stride };
        const long long m = (long long)ncloth * (r1 - r0) * c->p.nx;
        oc_k_unpack<<<(unsigned)((m + 255) / 256), 256, 0, c->s_un>>>(c->k, x, c->stage[0], c->stage[1], c->buf[c->q.ia], c->buf[c->q.ib]);
        c->launches++;
        OC_CUDA(cudaGetLastError());
        OC_CUDA(cudaEventRecord(c->pipe.ev_in[k], c->s_un));
    }
    c->pipe.up_chunks = nch;
    if (c->q.band && !c->q.linked) c->q.fresh = c->q.kmax;     // halos are stale until the host exchanges them
    return OC_OK;
}

extern "C" int oc_upload(oc_cloth* c, const float* X, const float* X_last, int stride)
{
    if (!c || !X || !X_last) return oc_fail(OC_ERR_INVALID, "oc_upload: null");
    return upload_impl(c, 0, c->p.batch, X, X_last, stride);
}
extern "C" int oc_upload_cloth(oc_cloth* c, int cloth, const float* X, const float* X_last, int stride)
{
    if (!c || !X || !X_last) return oc_fail(OC_ERR_INVALID, "oc_upload_cloth: null");
    if (cloth < 0 || cloth >= c->p.batch) return oc_fail(OC_ERR_INVALID, "oc_upload_cloth: cloth %d of %d", cloth, c->p.batch);
    return upload_impl(c, cloth, 1, X, X_last, stride);
}

static int download_impl(oc_cloth* c, int cloth0, int ncloth, float* X, float* X_last, int stride)
{
    if (stride != 3 && stride != 4) return oc_fail(OC_ERR_INVALID, "stride_floats must be 3 or 4");
    OC_CUDA(cudaSetDevice(c->dev));
    int rc = ensure_stage(c, (size_t)c->p.batch * c->rows_own * c->p.nx * 4 * sizeof(float));
    if (rc) return rc;
    const long long n = (long long)ncloth * c->rows_own * c->p.nx;
    // chunk events of a step that was launched chunk by chunk (oc_step right after oc_upload): follow them; otherwise
    // wait for the whole compute stream once
    const bool piped = c->pipe.step_chunks > 1 && ncloth == c->p.batch;
    const int nch = piped ? c->pipe.step_chunks : ((ncloth == c->p.batch) ? pipe_chunks(c) : 1);
    if (!piped) {
        rc = join_upload(c);
        if (rc) return rc;
        OC_CUDA(cudaEventRecord(c->pipe.ev_tail, c->stream));
        OC_CUDA(cudaStreamWaitEvent(c->s_pk, c->pipe.ev_tail, 0));
    }
    for (int k = 0; k < nch; ++k) {
        int r0, r1;
        chunk_rows(c, nch, k, &r0, &r1);
        if (piped) OC_CUDA(cudaStreamWaitEvent(c->s_pk, c->pipe.ev_step[k], 0));
        const OcXfer x = { cloth0, cloth0, ncloth, c->p.row_begin, c->rows_own, r0, r1 - r0, stride };
        const long long m = (long long)ncloth * (r1 - r0) * c->p.nx;
        oc_k_pack<<<(unsigned)((m + 255) / 256), 256, 0, c->s_pk>>>(c->k, x, c->buf[c->q.ia], c->buf[c->q.ib],
                                                                    X ? c->stage[0] : nullptr, X_last ? c->stage[1] : nullptr);
        c->launches++;
        OC_CUDA(cudaGetLastError());
        OC_CUDA(cudaEventRecord(c->pipe.ev_pk[k], c->s_pk));
        OC_CUDA(cudaStreamWaitEvent(c->s_out, c->pipe.ev_pk[k], 0));
        const size_t off = (size_t)(r0 - c->p.row_begin) * c->p.nx * stride;
        const size_t cnt = (nch == 1) ? (size_t)n * stride : (size_t)(r1 - r0) * c->p.nx * stride;
        if (X) OC_CUDA(cudaMemcpyAsync(X + off, c->stage[0] + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->s_out));
        if (X_last) OC_CUDA(cudaMemcpyAsync(X_last + off, c->stage[1] + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->s_out));
        if (c->pipe.timed) OC_CUDA(cudaEventRecord(c->pipe.ev_out[k], c->s_out));
    }
    c->pipe.step_chunks = 0;
    // the compute stream must not run ahead of the packing (the next step overwrites what it reads)
    OC_CUDA(cudaEventRecord(c->pipe.ev_tail, c->s_pk));
    OC_CUDA(cudaStreamWaitEvent(c->stream, c->pipe.ev_tail, 0));
    return oc_sync(c);
}
extern "C" int oc_download(oc_cloth* c, float* X, float* X_last, int stride)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_download: null");
    return download_impl(c, 0, c->p.batch, X, X_last, stride);
}
extern "C" int oc_download_cloth(oc_cloth* c, int cloth, float* X, float* X_last, int stride)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_download_cloth: null");
    if (cloth < 0 || cloth >= c->p.batch) return oc_fail(OC_ERR_INVALID, "oc_download_cloth: cloth %d of %d", cloth, c->p.batch);
    return download_impl(c, cloth, 1, X, X_last, stride);
}

// Render hand-off: per-vertex normals of the current state (oc_normals.cuh), to host memory.
extern "C" int oc_download_normals(oc_cloth* c, float* N, int stride)
{
    if (!c || !N) return oc_fail(OC_ERR_INVALID, "oc_download_normals: null");
    if (stride != 3 && stride != 4) return oc_fail(OC_ERR_INVALID, "stride_floats must be 3 or 4");
    if (c->q.band) return oc_fail(OC_ERR_UNSUPPORTED, "oc_download_normals: whole-cloth handles only (a band lacks its neighbours' rows)");
    OC_CUDA(cudaSetDevice(c->dev));
    int rc = join_upload(c);
    if (rc) return rc;
    rc = ensure_stage(c, (size_t)c->p.batch * c->rows_own * c->p.nx * 4 * sizeof(float));
    if (rc) return rc;
    dim3 blk(128, 1, 1), grd((c->p.nx + 127) / 128, c->p.ny, c->p.batch);
    oc_k_normals<<<grd, blk, 0, c->stream>>>(c->k, c->buf[c->q.ia], c->stage[0], stride);
    c->launches++;
    OC_CUDA(cudaGetLastError());
    OC_CUDA(cudaMemcpyAsync(N, c->stage[0], (size_t)c->p.batch * c->p.ny * c->p.nx * stride * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    return oc_sync(c);
}

// Many write-backs in one launch: the per-environment actions of a batch (cloth[k], idx[k]) <- xyz[3k..3k+2], V:203-208 each.
extern "C" int oc_set_particles(oc_cloth* c, int n, const int* cloth, const int* idx, const float* xyz)
{
    if (!c || n < 0 || (n > 0 && (!cloth || !idx || !xyz))) return oc_fail(OC_ERR_INVALID, "oc_set_particles: null");
    if (n == 0) return OC_OK;
    OC_CUDA(cudaSetDevice(c->dev));
    std::vector<OcPoke> pk;
    pk.reserve((size_t)n);
    for (int k = 0; k < n; ++k) {
        if (cloth[k] < 0 || cloth[k] >= c->p.batch || idx[k] < 0 || idx[k] >= c->p.nx * c->p.ny)
            return oc_fail(OC_ERR_INVALID, "oc_set_particles: entry %d: cloth %d / index %d out of range", k, cloth[k], idx[k]);
        const int j = idx[k] / c->p.nx, i = idx[k] % c->p.nx;
        if (j < c->k.row_lo || j >= c->k.row_lo + c->k.srows) continue;            // not stored by this band
        OcPoke q = { oc_index(c->k, cloth[k], i, j), xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], 0 };
        pk.push_back(q);
    }
    chain_break(c);
    int rc = join_upload(c);
    if (rc) return rc;
    if (pk.empty()) return OC_OK;
    OcPoke* d = nullptr;
    OC_CUDA(cudaMallocAsync(&d, pk.size() * sizeof(OcPoke), c->stream));
    OC_CUDA(cudaMemcpyAsync(d, pk.data(), pk.size() * sizeof(OcPoke), cudaMemcpyHostToDevice, c->stream));
    oc_k_set_particles<<<(unsigned)((pk.size() + 127) / 128), 128, 0, c->stream>>>(c->buf[c->q.ia], c->buf[c->q.ib], d, (int)pk.size(), c->q.xv ? 1 : 0);
    c->launches++;
    OC_CUDA(cudaGetLastError());
    OC_CUDA(cudaFreeAsync(d, c->stream));
    OC_CUDA(cudaStreamSynchronize(c->stream));                                    // pk is pageable host memory
    return OC_OK;
}

extern "C" int oc_set_particle(oc_cloth* c, int cloth, int idx, const float xyz[3])
{
    if (!c || !xyz) return oc_fail(OC_ERR_INVALID, "oc_set_particle: null");
    if (cloth < 0 || cloth >= c->p.batch || idx < 0 || idx >= c->p.nx * c->p.ny)
        return oc_fail(OC_ERR_INVALID, "oc_set_particle: cloth %d / index %d out of range", cloth, idx);
    chain_break(c);
    int j = idx / c->p.nx, i = idx % c->p.nx;
    if (j < c->k.row_lo || j >= c->k.row_lo + c->k.srows) return OC_OK;      // not stored by this band
    OC_CUDA(cudaSetDevice(c->dev));
    { const int rcj = join_upload(c); if (rcj) return rcj; }
    oc_k_set_particle<<<1, 1, 0, c->stream>>>(c->buf[c->q.ia], c->buf[c->q.ib], oc_index(c->k, cloth, i, j), xyz[0], xyz[1], xyz[2], c->q.xv ? 1 : 0);
    c->launches++;
    OC_CUDA(cudaGetLastError());
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// run-time pin sets
// ------------------------------------------------------------------------------------------------
static void pins_set_bit(oc_cloth* c, int cloth, int idx, bool on)
{
    const long long per = (long long)c->p.nx * c->p.ny;
    const long long bit = cloth * per + idx;
    if (on) (*c->h_pins)[bit >> 5] |= 1u << (bit & 31); else (*c->h_pins)[bit >> 5] &= ~(1u << (bit & 31));
}
extern "C" int oc_set_pins(oc_cloth* c, int cloth, const int* idx, int n)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_set_pins: null");
    if (cloth < -1 || cloth >= c->p.batch) return oc_fail(OC_ERR_INVALID, "oc_set_pins: cloth %d of %d", cloth, c->p.batch);
    if (n < 0 || (n > 0 && !idx)) return oc_fail(OC_ERR_INVALID, "oc_set_pins: bad index list");
    const int U = c->p.nx, V = c->p.ny, B = c->p.batch;
    const long long per = (long long)U * V;
    for (int k = 0; k < n; ++k) if (idx[k] < 0 || idx[k] >= per) return oc_fail(OC_ERR_INVALID, "oc_set_pins: index %d out of range", idx[k]);
    OC_CUDA(cudaSetDevice(c->dev));
    const size_t words = (size_t)((per * B + 31) / 32);
    if (!c->h_pins) {
        c->h_pins = new std::vector<unsigned>(words, 0u);
        c->h_pin_rows = new std::vector<unsigned char>((size_t)B * V, 0);
        for (int b = 0; b < B; ++b) { pins_set_bit(c, b, 0, true); pins_set_bit(c, b, U - 1, true); (*c->h_pin_rows)[(size_t)b * V] = 1; }   // V:455
        OC_CUDA(cudaMalloc(&c->d_pins, words * sizeof(unsigned)));
        OC_CUDA(cudaMalloc(&c->d_pin_rows, (size_t)B * V));
    }
    for (int b = (cloth < 0 ? 0 : cloth); b < (cloth < 0 ? B : cloth + 1); ++b) {
        for (long long q = 0; q < per; ++q) pins_set_bit(c, b, (int)q, false);
        for (int j = 0; j < V; ++j) (*c->h_pin_rows)[(size_t)b * V + j] = 0;
        for (int k = 0; k < n; ++k) { pins_set_bit(c, b, idx[k], true); (*c->h_pin_rows)[(size_t)b * V + idx[k] / U] = 1; }
    }
    chain_break(c);
    OC_CUDA(cudaStreamSynchronize(c->stream));                      // kernels in flight still read the old set
    OC_CUDA(cudaMemcpy(c->d_pins, c->h_pins->data(), words * sizeof(unsigned), cudaMemcpyHostToDevice));
    OC_CUDA(cudaMemcpy(c->d_pin_rows, c->h_pin_rows->data(), (size_t)B * V, cudaMemcpyHostToDevice));
    c->k.pins = c->d_pins; c->k.pin_rows = c->d_pin_rows;
    return OC_OK;
}
extern "C" int oc_reset_pins(oc_cloth* c)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_reset_pins: null");
    OC_CUDA(cudaSetDevice(c->dev));
    chain_break(c);
    OC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_pins) cudaFree(c->d_pins);
    if (c->d_pin_rows) cudaFree(c->d_pin_rows);
    delete c->h_pins; delete c->h_pin_rows;
    c->d_pins = nullptr; c->d_pin_rows = nullptr; c->h_pins = nullptr; c->h_pin_rows = nullptr;
    c->k.pins = nullptr; c->k.pin_rows = nullptr;
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------------
static bool oc_stream_pays(const oc_cloth* c) { return (long long)c->p.nx * c->rows_own * c->p.batch >= (3LL << 20); }
// oc_k_bandres: bands and rows of the tallest band for this cloth; false if the kernel does not apply
static bool bandres_plan(const oc_cloth* c, int* nb, int* rmax)
{
    if (c->q.band || c->link.on || c->p.batch != 1 || c->p.integrator != OC_INTEGRATOR_VERLET || c->p.provot || !c->bres.coop) return false;
    return oc_bandres_plan(c->p.nx, c->p.ny, c->sm_count, nb, rmax);
}

static int pick_kernel(const oc_cloth* c)
{
    const bool can_reside = !c->q.band && (long long)c->p.nx * c->p.ny <= OC_RESIDENT_MAX_PARTICLES;
    int bres_nb = 0, bres_r = 0;
    const bool can_band = bandres_plan(c, &bres_nb, &bres_r);
    if (c->link.on) return (c->p.kernel == OC_KERNEL_TWIN || c->p.kernel == OC_KERNEL_STREAM || c->p.kernel == OC_KERNEL_STREAM2) ? c->p.kernel : ((c->p.kernel == OC_KERNEL_AUTO && !c->p.exact && oc_stream_pays(c)) ? OC_KERNEL_STREAM : OC_KERNEL_MARCH2);      // linked row bands: the kernels that push their boundary rows
    if (c->p.integrator != OC_INTEGRATOR_VERLET) return OC_KERNEL_GATHER;      // state (X, V): oc_k_gather_xv
    if (c->p.provot && (c->p.kernel == OC_KERNEL_RESIDENT || c->p.kernel == OC_KERNEL_BANDRES || c->p.kernel == OC_KERNEL_AUTO)) return OC_KERNEL_MARCH2;   // a pass after EVERY substep
    if (c->p.kernel == OC_KERNEL_RESIDENT) return can_reside ? OC_KERNEL_RESIDENT : OC_KERNEL_MARCH2;   // a cloth that does not fit one CTA's shared memory: the fastest general kernel
    if (c->p.kernel == OC_KERNEL_BANDRES) return can_band ? OC_KERNEL_BANDRES : OC_KERNEL_MARCH2;       // (batches, row bands, Provot, bands too tall for shared memory)
    if (c->p.kernel != OC_KERNEL_AUTO) return c->p.kernel;
    // small whole cloths (the reference's own 21 x 21): state resident in shared memory, all substeps in one launch.
    // One CTA per cloth: worth it while the CTA's 1024 threads cover the cloth's springs in a few passes (beyond that
    // the gather kernel, which spreads one cloth over many SMs, or the marching kernel is quicker).
    if (can_reside && c->p.substeps_per_launch <= 1 && (long long)c->p.nx * c->p.ny <= 1024) return OC_KERNEL_RESIDENT;
    // mid-size whole cloths: one row band per SM resident in shared memory, all substeps in one launch (oc_bandres.cuh)
    if (can_band && c->p.substeps_per_launch <= 1) return OC_KERNEL_BANDRES;
    // one substep per launch: the two-columns-per-thread kernel (fastest); k > 1: the staged one-column kernel
    if (c->p.substeps_per_launch > 1) return OC_KERNEL_MARCH;
    // one substep per launch.  Exact mode: every spring once, two columns per thread (oc_k_march2; the bit-exact arithmetic is
    // FMA-pipe heavy, so halving it wins).  Fast mode: the streaming gather kernel (oc_k_stream; twice the spring arithmetic,
    // no force exchange, more resident warps) is 1-7 % ahead from about three million particles per handle (2048^2: 66.8
    // against 63.0 G updates/s; 512 x 128^2: 70.8 against 66.9); below that its CTAs of two tiles leave too few rows per tile
    // (1536^2: 49.0 against 51.7; 1024^2: 31 against 41).  DESIGN.md 4.2-4.3.
    if (c->p.exact) return OC_KERNEL_MARCH2;
    return oc_stream_pays(c) ? OC_KERNEL_STREAM : OC_KERNEL_MARCH2;
}

// one launch: rows [ra, rb) of the launch descriptor's destination, S substeps
static int launch_rows(oc_cloth* c, int kern, const OcLaunch& L, int ra, int rb)
{
    if (rb <= ra) return OC_OK;
    if ((kern == OC_KERNEL_TWIN || kern == OC_KERNEL_STREAM || kern == OC_KERNEL_STREAM2) && rb - ra < 2) kern = OC_KERNEL_MARCH2;      // (a single row cannot be cut into two tiles)
    if (kern == OC_KERNEL_MARCH2 || kern == OC_KERNEL_TWIN || kern == OC_KERNEL_STREAM || kern == OC_KERNEL_STREAM2) {
        int nl = 0;
        OcPeer2 peer = {};
        if (c->link.on) {
            const long long U = c->p.nx;
            for (int sd = 0; sd < 2; ++sd) {
                if (!c->link.has[sd]) continue;
                peer.c[sd] = c->link.buf[sd][L.dst] - (long long)c->link.row_lo[sd] * U;
                peer.flags_out[sd] = c->link.flags_out[sd];
                peer.flags_in[sd] = c->in_flags + sd * OC_LINK_STRIPS;
            }
            peer.epoch = ++c->link.epoch;
            peer.rev = c->link.rev;
        }
        cudaError_t e = kern == OC_KERNEL_MARCH2
            ? oc_march2_launch(c->k, c->p.exact != 0, ra, rb, c->sm_count, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->stream, &nl, &c->chain,
                               c->link.on ? &peer : nullptr)
            : oc_twin_launch(c->k, c->p.exact != 0, ra, rb, c->sm_count, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->stream, &nl, &c->chain,
                             c->link.on ? &peer : nullptr, kern == OC_KERNEL_STREAM2 ? 2 : (kern == OC_KERNEL_STREAM ? 1 : 0));
        c->launches += nl;
        if (e != cudaSuccess) return oc_fail(OC_ERR_CUDA, "march2 kernel launch failed: %s", cudaGetErrorString(e));
    } else if (kern == OC_KERNEL_RESIDENT) {
        chain_break(c);
        const int N = c->p.nx * c->p.ny;
        const size_t smem = OcResidentSmem::bytes(N);
        int threads = (3 * N + 31) / 32 * 32; if (threads > OC_RESIDENT_THREADS) threads = OC_RESIDENT_THREADS;
        if (c->p.exact) oc_k_resident<MathExact><<<c->p.batch, threads, smem, c->stream>>>(c->k, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->buf[L.dst_prev], L.S);
        else            oc_k_resident<MathFast><<<c->p.batch, threads, smem, c->stream>>>(c->k, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->buf[L.dst_prev], L.S);
        c->launches++;
        OC_CUDA(cudaGetLastError());
    } else if (kern == OC_KERNEL_BANDRES) {
        chain_break(c);
        int nb = 0, rmax = 0;
        if (!bandres_plan(c, &nb, &rmax)) return oc_fail(OC_ERR_INVALID, "oc_k_bandres does not apply to this cloth");
        const size_t NG = (size_t)c->p.nx * c->p.ny;
        if (!c->bres.ex) {
            OC_CUDA(cudaMalloc(&c->bres.ex, OC_BANDRES_EX_BYTES(NG)));
            OC_CUDA(cudaMemsetAsync(c->bres.ex, 0, OC_BANDRES_EX_BYTES(NG), c->stream));      // tag 0: nothing sent yet
            c->bres.epoch = 0;
            OC_CUDA(cudaMalloc(&c->bres.flags, OC_BANDRES_MAX_BANDS * sizeof(unsigned)));
            OC_CUDA(cudaMemsetAsync(c->bres.flags, 0, OC_BANDRES_MAX_BANDS * sizeof(unsigned), c->stream));
            c->bres.epoch = 0;
        }
        const float4* a = c->buf[L.src_a]; const float4* b = c->buf[L.src_b];
        float4* d0 = c->buf[L.dst]; float4* d1 = c->buf[L.dst_prev];
        int S = L.S;
        unsigned epoch = c->bres.epoch;
        void* args[] = { (void*)&c->k, (void*)&a, (void*)&b, (void*)&d0, (void*)&d1, (void*)&S, (void*)&c->bres.ex, (void*)&c->bres.flags, (void*)&epoch, (void*)&rmax };
        // (768 threads per CTA - 24 warps at 80 registers, three instead of four rounds of particles at 512^2 - measured equal
        // within +-5 %, profiles/r2/bandres_threads.log: the gather is bound by its instruction count, not by the rounds)
        const int threads = OC_BANDRES_THREADS;
        const void* fn = c->p.exact ? (const void*)&oc_k_bandres<MathExact, OC_BANDRES_THREADS> : (const void*)&oc_k_bandres<MathFast, OC_BANDRES_THREADS>;
        if (c->k.dbg & 16) fprintf(stderr, "[oc] bandres: %d bands of <= %d rows, %zu bytes of shared memory, %d substeps\n", nb, rmax, OcBandresSmem::bytes(c->p.nx, rmax), S);
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(threads), args, OcBandresSmem::bytes(c->p.nx, rmax), c->stream);
        if (e != cudaSuccess) return oc_fail(OC_ERR_CUDA, "oc_k_bandres launch failed: %s", cudaGetErrorString(e));
        c->bres.epoch += (unsigned)S;
        c->launches++;
    } else if (kern == OC_KERNEL_MARCH) {
        chain_break(c);
        int nl = 0;
        cudaError_t e = oc_march_launch(c->k, c->p.exact != 0, L.S, ra, rb, c->sm_count,
                                        c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->buf[L.dst_prev], c->stream, &nl);
        c->launches += nl;
        if (e != cudaSuccess) return oc_fail(OC_ERR_CUDA, "march kernel launch failed: %s", cudaGetErrorString(e));
    } else {
        chain_break(c);
        dim3 blk(128, 1, 1), grd((c->p.nx + 127) / 128, rb - ra, c->p.batch);
        if (c->q.xv) {
            if (c->p.exact) oc_k_gather_xv<MathExact><<<grd, blk, 0, c->stream>>>(c->k, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->buf[L.dst_prev], ra);
            else            oc_k_gather_xv<MathFast><<<grd, blk, 0, c->stream>>>(c->k, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], c->buf[L.dst_prev], ra);
        }
        else if (c->p.exact) oc_k_gather<MathExact><<<grd, blk, 0, c->stream>>>(c->k, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], ra);
        else                 oc_k_gather<MathFast><<<grd, blk, 0, c->stream>>>(c->k, c->buf[L.src_a], c->buf[L.src_b], c->buf[L.dst], ra);
        c->launches++;
        OC_CUDA(cudaGetLastError());
    }
    return OC_OK;
}

// ApplyProvotDynamicInverse on the state the last substep produced (oc_provot.cuh), bit-exact in list order.
template <class M>
static int provot_pass_t(oc_cloth* c)
{
    const int U = c->p.nx, V = c->p.ny, B = c->p.batch;
    float4* X = c->buf[c->q.ia];
    float4* S = c->buf[c->q.ib];          // X_last (Verlet) or V (Euler)
    chain_break(c);
    if (c->q.xv) {
        dim3 blk(128, 1, 1), grd((U + 127) / 128, V, B);
        oc_k_provot_v<M><<<grd, blk, 0, c->stream>>>(c->k, X, S, S);
        c->launches++;
    } else {
        const long long n = c->stored;
        oc_k_provot_materialize<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->k, X, S);
        oc_k_provot_rows<M, 1><<<dim3((V + 127) / 128, B), 128, 0, c->stream>>>(c->k, X);          // V:288-291
        oc_k_provot_cols<M, 1><<<dim3((U + 127) / 128, B), 128, 0, c->stream>>>(c->k, X);          // V:294-297
        int t = V - 1 < 1024 ? (V - 1 + 31) / 32 * 32 : 1024;
        oc_k_provot_shear<M><<<B, t, 0, c->stream>>>(c->k, X);                                     // V:301-305
        oc_k_provot_rows<M, 2><<<dim3((V + 127) / 128, B), 128, 0, c->stream>>>(c->k, X);          // V:309-314
        oc_k_provot_cols<M, 2><<<dim3((U + 127) / 128, B), 128, 0, c->stream>>>(c->k, X);          // V:315-320
        c->launches += 6;
    }
    OC_CUDA(cudaGetLastError());
    return OC_OK;
}
static int provot_pass(oc_cloth* c) { return c->p.exact ? provot_pass_t<MathExact>(c) : provot_pass_t<MathFast>(c); }

// n substeps; if split_stream != nullptr and the last substep is a single-substep launch over exactly the
// owned rows of a band, that launch is issued as boundary rows first (the rows the neighbours need), an
// event, then the interior, and split_stream is made to wait for the event.
static int step_impl(oc_cloth* c, int n, cudaStream_t split_stream, bool want_split, int* did_split)
{
    if (did_split) *did_split = 0;
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_step: null");
    if (n < 0) return oc_fail(OC_ERR_INVALID, "oc_step: n < 0");
    { const int rc = check_err_word(c); if (rc) return rc; }
    if (c->q.band && !c->q.linked && c->q.fresh + n > c->q.kmax)
        return oc_fail(OC_ERR_INVALID, "oc_step: %d substeps requested but only %d remain before the halo rows must be exchanged "
                                       "(halo_rows=%d)", n, c->q.kmax - c->q.fresh, c->p.halo_rows);
    OC_CUDA(cudaSetDevice(c->dev));
    const int kern = pick_kernel(c);
    const int kdef = c->p.substeps_per_launch > 0 ? c->p.substeps_per_launch : 1;
    // host <-> device pipeline (upload_impl): the first substep after a chunked upload follows the chunks
    c->pipe.step_chunks = 0;
    const bool pipe_first = c->pipe.up_chunks > 1 && n >= 1 && !c->p.provot && !c->q.band && c->p.batch == 1 &&
                            (kern == OC_KERNEL_MARCH2 || kern == OC_KERNEL_TWIN || kern == OC_KERNEL_STREAM || kern == OC_KERNEL_STREAM2 || kern == OC_KERNEL_GATHER);
    if (!pipe_first) { const int rcj = join_upload(c); if (rcj) return rcj; }
    if (pipe_first) {
        const int nch = c->pipe.up_chunks, n_call = n;
        OcLaunch L;
        oc_host_next_launch(c->q, n, 1, L);
        for (int k = 0; k < nch; ++k) {
            int r0, r1;
            chunk_rows(c, nch, k, &r0, &r1);
            OC_CUDA(cudaStreamWaitEvent(c->stream, c->pipe.ev_in[k + 1 < nch ? k + 1 : k], 0));      // the stencil reaches 2 rows into the next chunk
            chain_break(c);
            const int rc = launch_rows(c, kern, L, r0, r1);
            if (rc) return rc;
            OC_CUDA(cudaEventRecord(c->pipe.ev_step[k], c->stream));
        }
        chain_break(c);
        c->pipe.up_chunks = 0;
        c->pipe.step_chunks = (n_call == 1) ? nch : 0;         // a download that follows directly may go chunk by chunk
    }
    while (n > 0) {
        int kmaxS = (kern == OC_KERNEL_MARCH) ? oc_host_pick_stages(n < kdef ? n : kdef) : (kern == OC_KERNEL_RESIDENT ? OC_RESIDENT_MAX_STEPS : (kern == OC_KERNEL_BANDRES ? OC_BANDRES_MAX_STEPS : 1));
        if (c->p.provot || c->q.xv) kmaxS = 1;                    // the Provot pass follows every substep; (X, V) steps are single
        OcLaunch L;
        oc_host_next_launch(c->q, n, kmaxS, L);
        const int H = c->p.halo_rows;
        const bool split = want_split && n == 0 && c->q.band && !c->q.linked && L.S == 1 && L.ra == c->p.row_begin && L.rb == c->p.row_end &&
                           (L.rb - L.ra) > 2 * H + 8;
        int rc;
        if (!split) {
            rc = launch_rows(c, kern, L, L.ra, L.rb);
            if (rc) return rc;
            if (c->p.provot) { rc = provot_pass(c); if (rc) return rc; }
        } else {
            // three launches of ONE step over parts of the rows: each waits for the whole grid before it (no chaining)
            chain_break(c);
            rc = launch_rows(c, kern, L, L.ra, L.ra + H);          if (rc) return rc;
            chain_break(c);
            rc = launch_rows(c, kern, L, L.rb - H, L.rb);          if (rc) return rc;
            OC_CUDA(cudaEventRecord(c->ev_ready, c->stream));
            if (split_stream) OC_CUDA(cudaStreamWaitEvent(split_stream, c->ev_ready, 0));
            chain_break(c);
            rc = launch_rows(c, kern, L, L.ra + H, L.rb - H);      if (rc) return rc;
            chain_break(c);
            if (did_split) *did_split = 1;
        }
    }
    return OC_OK;
}

extern "C" int oc_step(oc_cloth* c, int n) { return step_impl(c, n, nullptr, false, nullptr); }

extern "C" int oc_step_split(oc_cloth* c, int n, void* exchange_stream, int* did_split)
{
    return step_impl(c, n, (cudaStream_t)exchange_stream, true, did_split);
}

extern "C" int oc_step_timed(oc_cloth* c, int n, float* ms)
{
    if (!c || !ms) return oc_fail(OC_ERR_INVALID, "oc_step_timed: null");
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    OC_CUDA(cudaEventRecord(c->ev0, c->stream));
    int rc = oc_step(c, n);
    if (rc) return rc;
    OC_CUDA(cudaEventRecord(c->ev1, c->stream));
    OC_CUDA(cudaEventSynchronize(c->ev1));
    OC_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// row-band halo plumbing
// ------------------------------------------------------------------------------------------------
static int halo_region(oc_cloth* c, int side, int which, bool send, void** ptr, size_t* count)
{
    if (!c || !ptr || !count) return oc_fail(OC_ERR_INVALID, "oc_halo_*: null");
    if ((side != 0 && side != 1) || (which != 0 && which != 1)) return oc_fail(OC_ERR_INVALID, "oc_halo_*: side/which must be 0 or 1");
    *ptr = nullptr; *count = 0;
    int r0, rows;
    if (!oc_host_halo_rows(c->p, c->q.band, side, send, &r0, &rows)) return OC_OK;
    const int U = c->p.nx;
    float4* base = c->buf[which == 0 ? c->q.ia : c->q.ib];
    *ptr = (void*)(base + (long long)(r0 - c->k.row_lo) * U);
    *count = (size_t)rows * U;
    return OC_OK;
}
extern "C" int oc_halo_send_region(oc_cloth* c, int side, int which, void** p, size_t* n) { return halo_region(c, side, which, true, p, n); }
extern "C" int oc_halo_recv_region(oc_cloth* c, int side, int which, void** p, size_t* n) { return halo_region(c, side, which, false, p, n); }
extern "C" int oc_halo_refreshed(oc_cloth* c)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_halo_refreshed: null");
    chain_break(c);              // the halo rows were written by something else (peer copy / NCCL receive)
    c->q.fresh = 0;
    return OC_OK;
}
extern "C" int oc_halo_budget(const oc_cloth* c)
{
    if (!c) return 0;
    return (c->q.band && !c->q.linked) ? c->q.kmax - c->q.fresh : 0x7fffffff;
}

extern "C" int oc_halo_exchange(oc_cloth* const* bands, int n)
{
    if (!bands || n < 1) return oc_fail(OC_ERR_INVALID, "oc_halo_exchange: no bands");
    for (int b = 0; b < n; ++b) {
        if (!bands[b]) return oc_fail(OC_ERR_INVALID, "oc_halo_exchange: null band");
        chain_break(bands[b]);
        if (b > 0 && (bands[b]->p.row_begin != bands[b - 1]->p.row_end || bands[b]->p.nx != bands[0]->p.nx ||
                      bands[b]->p.ny != bands[0]->p.ny || bands[b]->p.halo_rows != bands[0]->p.halo_rows))
            return oc_fail(OC_ERR_INVALID, "oc_halo_exchange: bands must be consecutive row bands of one cloth with equal halo_rows");
    }
    // 0. direct peer access between neighbouring bands' devices (NVLink on an HGX box); without it the
    //    peer copies are staged through the host.  "already enabled" / "not supported" are not errors.
    for (int b = 0; b + 1 < n; ++b) {
        const int d0 = bands[b]->dev, d1 = bands[b + 1]->dev;
        if (d0 == d1) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, d0, d1) == cudaSuccess && can) { cudaSetDevice(d0); cudaDeviceEnablePeerAccess(d1, 0); }
        if (cudaDeviceCanAccessPeer(&can, d1, d0) == cudaSuccess && can) { cudaSetDevice(d1); cudaDeviceEnablePeerAccess(d0, 0); }
        cudaGetLastError();
    }
    // 1. every band announces that its owned rows are final
    for (int b = 0; b < n; ++b) {
        OC_CUDA(cudaSetDevice(bands[b]->dev));
        OC_CUDA(cudaEventRecord(bands[b]->ev_ready, bands[b]->stream));
    }
    // 2. every band pulls its two halos on its own stream, after the producer is ready
    for (int b = 0; b < n; ++b) {
        oc_cloth* me = bands[b];
        OC_CUDA(cudaSetDevice(me->dev));
        for (int side = 0; side < 2; ++side) {
            oc_cloth* nb = (side == 0) ? (b > 0 ? bands[b - 1] : nullptr) : (b + 1 < n ? bands[b + 1] : nullptr);
            if (!nb) continue;
            OC_CUDA(cudaStreamWaitEvent(me->stream, nb->ev_ready, 0));
            for (int which = 0; which < 2; ++which) {
                void *src, *dst; size_t ns, nd;
                int rc = halo_region(nb, 1 - side, which, true, &src, &ns);
                if (rc) return rc;
                rc = halo_region(me, side, which, false, &dst, &nd);
                if (rc) return rc;
                if (ns != nd) return oc_fail(OC_ERR_INVALID, "oc_halo_exchange: halo size mismatch");
                if (ns == 0) continue;
                if (nb->dev == me->dev) OC_CUDA(cudaMemcpyAsync(dst, src, ns * sizeof(float4), cudaMemcpyDeviceToDevice, me->stream));
                else                    OC_CUDA(cudaMemcpyPeerAsync(dst, me->dev, src, nb->dev, ns * sizeof(float4), me->stream));
            }
        }
        OC_CUDA(cudaEventRecord(me->ev_filled, me->stream));
    }
    // 3. nobody overwrites rows a neighbour is still reading: wait for the neighbours' pulls
    for (int b = 0; b < n; ++b) {
        oc_cloth* me = bands[b];
        OC_CUDA(cudaSetDevice(me->dev));
        if (b > 0)     OC_CUDA(cudaStreamWaitEvent(me->stream, bands[b - 1]->ev_filled, 0));
        if (b + 1 < n) OC_CUDA(cudaStreamWaitEvent(me->stream, bands[b + 1]->ev_filled, 0));
        me->q.fresh = 0;
    }
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// linked row bands: peer-memory halos and cross-GPU tile flags (OcPeer2 in oc_march2.cuh)
// ------------------------------------------------------------------------------------------------
#include <unistd.h>
struct OcEndpoint {
    unsigned magic, abi;
    long long pid;
    int dev;
    int nx, ny, row_begin, row_end, row_lo, srows, halo_rows, ia, ib;
    int ipc_ok;
    unsigned long long buf_ptr[4], flags_ptr;        // addresses in the exporting process
    cudaIpcMemHandle_t buf_h[4], flags_h;
};
static_assert(sizeof(OcEndpoint) <= OC_BAND_ENDPOINT_BYTES, "OC_BAND_ENDPOINT_BYTES too small");
#define OC_ENDPOINT_MAGIC 0x4f43424eu

extern "C" int oc_band_endpoint(oc_cloth* c, void* blob, size_t bytes)
{
    if (!c || !blob) return oc_fail(OC_ERR_INVALID, "oc_band_endpoint: null");
    if (bytes < OC_BAND_ENDPOINT_BYTES) return oc_fail(OC_ERR_INVALID, "oc_band_endpoint: buffer of %zu bytes, need OC_BAND_ENDPOINT_BYTES = %d", bytes, OC_BAND_ENDPOINT_BYTES);
    if (!c->q.band) return oc_fail(OC_ERR_INVALID, "oc_band_endpoint: the handle owns the whole cloth");
    OC_CUDA(cudaSetDevice(c->dev));
    memset(blob, 0, OC_BAND_ENDPOINT_BYTES);
    OcEndpoint e;
    memset(&e, 0, sizeof(e));
    e.magic = OC_ENDPOINT_MAGIC; e.abi = OC_ABI_VERSION; e.pid = (long long)getpid(); e.dev = c->dev;
    e.nx = c->p.nx; e.ny = c->p.ny; e.row_begin = c->p.row_begin; e.row_end = c->p.row_end;
    e.row_lo = c->k.row_lo; e.srows = c->k.srows; e.halo_rows = c->p.halo_rows; e.ia = c->q.ia; e.ib = c->q.ib;
    e.ipc_ok = 1;
    for (int b = 0; b < 4; ++b) {
        e.buf_ptr[b] = (unsigned long long)(uintptr_t)c->buf[b];
        if (cudaIpcGetMemHandle(&e.buf_h[b], c->buf[b]) != cudaSuccess) e.ipc_ok = 0;
    }
    e.flags_ptr = (unsigned long long)(uintptr_t)c->in_flags;
    if (cudaIpcGetMemHandle(&e.flags_h, c->in_flags) != cudaSuccess) e.ipc_ok = 0;
    cudaGetLastError();                    // without IPC support the endpoint still serves bands of the same process
    memcpy(blob, &e, sizeof(e));
    return OC_OK;
}

static void unlink_band(oc_cloth* c)
{
    for (int sd = 0; sd < 2; ++sd)
        for (int b = 0; b < 5; ++b)
            if (c->link.opened[sd][b]) { cudaIpcCloseMemHandle(c->link.opened[sd][b]); c->link.opened[sd][b] = nullptr; }
    memset(&c->link, 0, sizeof(c->link));
    c->q.linked = false;
    chain_break(c);
}

extern "C" int oc_band_unlink(oc_cloth* c)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_band_unlink: null");
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    unlink_band(c);
    if (c->q.band) c->q.fresh = c->q.kmax;          // halo rows are stale until an exchange
    return OC_OK;
}

extern "C" int oc_band_link(oc_cloth* c, const void* up_blob, const void* down_blob)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_band_link: null");
    if (!c->q.band) return oc_fail(OC_ERR_INVALID, "oc_band_link: the handle owns the whole cloth");
    if (c->p.halo_rows < 2 || c->rows_own < 4) return oc_fail(OC_ERR_UNSUPPORTED, "oc_band_link: needs halo_rows >= 2 and at least 4 owned rows");
    if (oc_march2_nstrips(c->p.nx) > OC_LINK_STRIPS) return oc_fail(OC_ERR_UNSUPPORTED, "oc_band_link: cloth too wide (%d strips)", oc_march2_nstrips(c->p.nx));
    const bool need[2] = { c->p.row_begin > 0, c->p.row_end < c->p.ny };
    const void* blobs[2] = { up_blob, down_blob };
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    unlink_band(c);
    for (int sd = 0; sd < 2; ++sd) {
        if (!need[sd]) continue;
        if (!blobs[sd]) return oc_fail(OC_ERR_INVALID, "oc_band_link: rows [%d,%d) of %d have a neighbour %s but no endpoint was given",
                                       c->p.row_begin, c->p.row_end, c->p.ny, sd ? "below" : "above");
        OcEndpoint e;
        memcpy(&e, blobs[sd], sizeof(e));
        if (e.magic != OC_ENDPOINT_MAGIC || e.abi != OC_ABI_VERSION) { unlink_band(c); return oc_fail(OC_ERR_INVALID, "oc_band_link: not an endpoint of this library version"); }
        const bool adjacent = sd == 0 ? e.row_end == c->p.row_begin : e.row_begin == c->p.row_end;
        if (e.nx != c->p.nx || e.ny != c->p.ny || !adjacent || e.halo_rows < 2 || e.row_end - e.row_begin < 4) {
            unlink_band(c);
            return oc_fail(OC_ERR_INVALID, "oc_band_link: endpoint rows [%d,%d) of a %d x %d cloth are not the %s neighbour of rows [%d,%d) of a %d x %d cloth",
                           e.row_begin, e.row_end, e.nx, e.ny, sd ? "lower" : "upper", c->p.row_begin, c->p.row_end, c->p.nx, c->p.ny);
        }
        if (e.ia != c->q.ia || e.ib != c->q.ib) {
            unlink_band(c);
            return oc_fail(OC_ERR_INVALID, "oc_band_link: the bands have taken different numbers of steps (buffer rotation %d/%d vs %d/%d)", e.ia, e.ib, c->q.ia, c->q.ib);
        }
        void* ptr[5];
        if (e.pid == (long long)getpid()) {
            for (int b = 0; b < 4; ++b) ptr[b] = (void*)(uintptr_t)e.buf_ptr[b];
            ptr[4] = (void*)(uintptr_t)e.flags_ptr;
            if (e.dev != c->dev) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, c->dev, e.dev) != cudaSuccess || !can) {
                    unlink_band(c);
                    return oc_fail(OC_ERR_UNSUPPORTED, "oc_band_link: device %d cannot access device %d's memory", c->dev, e.dev);
                }
                const cudaError_t pe = cudaDeviceEnablePeerAccess(e.dev, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { unlink_band(c); OC_CUDA(pe); }
                cudaGetLastError();
            }
        } else {
            if (!e.ipc_ok) { unlink_band(c); return oc_fail(OC_ERR_UNSUPPORTED, "oc_band_link: the neighbour could not export CUDA IPC handles"); }
            for (int b = 0; b < 5; ++b) {
                const cudaError_t oe = cudaIpcOpenMemHandle(&ptr[b], b < 4 ? e.buf_h[b] : e.flags_h, cudaIpcMemLazyEnablePeerAccess);
                if (oe != cudaSuccess) { unlink_band(c); OC_CUDA(oe); }
                c->link.opened[sd][b] = ptr[b];
            }
        }
        for (int b = 0; b < 4; ++b) c->link.buf[sd][b] = (float4*)ptr[b];
        // the neighbour's words for ITS side that faces me: I am the lower neighbour (side 1) of my upper neighbour
        c->link.flags_out[sd] = (unsigned*)ptr[4] + (1 - sd) * OC_LINK_STRIPS;
        c->link.row_lo[sd] = e.row_lo; c->link.row_begin[sd] = e.row_begin; c->link.row_end[sd] = e.row_end;
        c->link.has[sd] = true;
    }
    OC_CUDA(cudaMemsetAsync(c->in_flags, 0, 2 * OC_LINK_STRIPS * sizeof(unsigned), c->stream));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    c->link.epoch = 0;
    {
        // neighbouring bands launch their segments in opposite directions (OcPeer2::rev).  The position of the band in the
        // chain is taken from its rows (exact for equal bands; unequal cuts may put two equal directions side by side,
        // which costs slack, not correctness).  OC_LINK_REV=0 / 1 forces one direction everywhere (measurements).
        const char* env = getenv("OC_LINK_REV");
        const int rows = c->p.row_end - c->p.row_begin;
        c->link.rev = env ? (atoi(env) == 1 ? 1 : 0) : (int)(((2LL * c->p.row_begin + rows) / (2LL * rows)) & 1);
    }
    c->link.on = true;
    c->q.linked = true;
    return OC_OK;
}

extern "C" int oc_band_pull_halo(oc_cloth* c)
{
    if (!c) return oc_fail(OC_ERR_INVALID, "oc_band_pull_halo: null");
    if (!c->link.on) return oc_fail(OC_ERR_INVALID, "oc_band_pull_halo: the band is not linked");
    OC_CUDA(cudaSetDevice(c->dev));
    chain_break(c);
    const long long U = c->p.nx;
    for (int sd = 0; sd < 2; ++sd) {
        if (!c->link.has[sd]) continue;
        const int r0 = sd == 0 ? c->p.row_begin - 2 : c->p.row_end;         // the two rows next to my band, owned by the neighbour
        for (int which = 0; which < 2; ++which) {
            const int b = which == 0 ? c->q.ia : c->q.ib;
            const float4* src = c->link.buf[sd][b] + (long long)(r0 - c->link.row_lo[sd]) * U;
            float4* dst = c->buf[b] + (long long)(r0 - c->k.row_lo) * U;
            OC_CUDA(cudaMemcpyAsync(dst, src, (size_t)2 * U * sizeof(float4), cudaMemcpyDefault, c->stream));
        }
    }
    return OC_OK;
}

extern "C" int oc_band_link_local(oc_cloth* const* bands, int n)
{
    if (!bands || n < 1) return oc_fail(OC_ERR_INVALID, "oc_band_link_local: no bands");
    std::vector<std::vector<unsigned char>> ep((size_t)n, std::vector<unsigned char>(OC_BAND_ENDPOINT_BYTES));
    for (int b = 0; b < n; ++b) {
        if (!bands[b]) return oc_fail(OC_ERR_INVALID, "oc_band_link_local: null band");
        int rc = oc_sync(bands[b]);                                         // every band's state is final ...
        if (rc == 0) rc = oc_band_endpoint(bands[b], ep[b].data(), ep[b].size());
        if (rc) return rc;
    }
    for (int b = 0; b < n; ++b) {
        const int rc = oc_band_link(bands[b], b > 0 ? ep[b - 1].data() : nullptr, b + 1 < n ? ep[b + 1].data() : nullptr);
        if (rc) return rc;
    }
    for (int b = 0; b < n; ++b) { const int rc = oc_band_pull_halo(bands[b]); if (rc) return rc; }
    for (int b = 0; b < n; ++b) { const int rc = oc_sync(bands[b]); if (rc) return rc; }      // ... and every halo current before anyone steps
    return OC_OK;
}

// ------------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------------
// ---- self-test of the branch-free IEEE sequences (oc_core.cuh) against the rounding intrinsics ----
__device__ __forceinline__ unsigned oc_rng(unsigned long long& st)
{
    st = st * 6364136223846793005ULL + 1442695040888963407ULL;
    return (unsigned)(st >> 32);
}
// random float with a uniformly random exponent in [elo, ehi] (unbiased), random mantissa and sign s
__device__ __forceinline__ float oc_rand_float(unsigned long long& st, int elo, int ehi, bool neg_ok)
{
    unsigned r = oc_rng(st);
    unsigned e = (unsigned)(elo + 127) + (oc_rng(st) % (unsigned)(ehi - elo + 1));
    unsigned bits = (e << 23) | (r & 0x7fffffu);
    if (neg_ok && (r & 0x80000000u)) bits |= 0x80000000u;
    return __uint_as_float(bits);
}
__global__ void oc_k_selftest(unsigned long long per_thread, unsigned seed, float dt, unsigned long long* out, float one)
{
    unsigned long long st = ((unsigned long long)seed << 32) ^ (0x9E3779B97F4A7C15ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1));
    unsigned long long bad_count = 0;
    const float ydt = oc_rcp_bf(dt);
    for (unsigned long long it = 0; it < per_thread; ++it) {
        bool bad = false;
        // sqrt over [2^-94, 2^94], rcp of the result, spring-style division
        float x = oc_rand_float(st, -94, 93, false);
        float s = oc_sqrt_bf(x, bad);
        if (__float_as_uint(s) != __float_as_uint(__fsqrt_rn(x))) bad_count++;
        float y = oc_rcp_bf(s);
        if (__float_as_uint(y) != __float_as_uint(__frcp_rn(s))) bad_count++;
        float a = oc_rand_float(st, -70, 69, true);
        if ((it & 63) == 0) a = (it & 64) ? -0.0f : 0.0f;
        float q = oc_div_bf(a, s, y, OC_NUM_LO, OC_NUM_HI, bad);
        if (__float_as_uint(q) != __float_as_uint(__fdiv_rn(a, s))) bad_count++;
        // clustered operands: a/b close to 1 and to powers of two (hard rounding cases)
        float b2 = oc_rand_float(st, -40, 40, false);
        bool bad2 = false;
        float sb = oc_sqrt_bf(b2 * b2, bad2);
        float yb = oc_rcp_bf(sb);
        float a2 = __uint_as_float(__float_as_uint(sb) + (oc_rng(st) & 7u) - 3u);
        float q2 = oc_div_bf(a2, sb, yb, OC_NUM_LO, OC_NUM_HI, bad2);
        if (__float_as_uint(q2) != __float_as_uint(__fdiv_rn(a2, sb))) bad_count++;
        // velocity-style division by dt over [2^-100, 2^100]
        float d = oc_rand_float(st, -100, 99, true);
        float v = oc_div_bf(d, dt, ydt, OC_VEL_LO, OC_VEL_HI, bad);
        if (__float_as_uint(v) != __float_as_uint(__fdiv_rn(d, dt))) bad_count++;
        if (bad || bad2) bad_count += 1000000;      // operands were generated inside the accepted ranges
        // accumulated range tests (OcRange / OcRangeVel) against the per-operand tests, on arbitrary bit patterns
        // biased towards the interval ends, the zeros and NaN / Inf
        {
            auto pattern = [&](unsigned lo, unsigned hi) -> float {
                const unsigned r = oc_rng(st), k = r & 15u;
                unsigned b;
                if (k == 0) b = 0u; else if (k == 1) b = 0x80000000u;
                else if (k == 2) b = lo + ((r >> 8) & 3u) - 2u; else if (k == 3) b = hi + ((r >> 8) & 3u) - 2u;
                else if (k == 4) b = 0x7f800000u | ((r >> 8) & 0x400001u);
                else if (k < 8) b = oc_rng(st);
                else b = lo + (unsigned)(((unsigned long long)(hi - lo) * (oc_rng(st) >> 4)) >> 28);
                if ((r & 0x100000u) && k >= 2) b ^= 0x80000000u;
                return __uint_as_float(b);
            };
            OcRange rg; rg.init(); OcRangeStrict rv; rv.init();
            bool ref_sq = false, ref_num = false, ref_vel = false;
            const int n_ops = 1 + (int)(oc_rng(st) & 3u);
            for (int k = 0; k < n_ops; ++k) {
                const float fs = pattern(0x10800000u, 0x6e800000u), fn = pattern(OC_NUM_LO_BITS, OC_NUM_HI_BITS), fv = pattern(OC_VEL_LO_BITS, OC_VEL_HI_BITS);
                rg.sqr(fs); rg.num(fn); rv.add(fv);
                ref_sq |= oc_bad_sqr(fs); ref_num |= oc_bad_num(fn, OC_NUM_LO_BITS, OC_NUM_HI_BITS); ref_vel |= oc_bad_vel(fv, OC_VEL_LO_BITS, OC_VEL_HI_BITS);
            }
            if (rg.bad() != (ref_sq | ref_num)) bad_count += 1ull << 48;
            if (rv.bad(OC_VEL_LO_BITS, OC_VEL_HI_BITS) != ref_vel) bad_count += 1ull << 48;
        }
        // packed FP32x2 forms: primitive ops and the pair sequences of the marching kernel
        {
            const float u0 = oc_rand_float(st, -30, 30, true), u1 = oc_rand_float(st, -30, 30, true);
            const float w0 = oc_rand_float(st, -30, 30, true), w1 = oc_rand_float(st, -30, 30, true);
            const float z0 = oc_rand_float(st, -30, 30, true), z1 = oc_rand_float(st, -30, 30, true);
            const float2 pa = p_add(make_float2(u0, u1), make_float2(w0, w1));
            const float2 pm = p_mul(make_float2(u0, u1), make_float2(w0, w1));
            const float2 pf = p_fma(make_float2(u0, u1), p_neg(make_float2(w0, w1)), make_float2(z0, z1));
            if (__float_as_uint(pa.x) != __float_as_uint(__fadd_rn(u0, w0)) || __float_as_uint(pa.y) != __float_as_uint(__fadd_rn(u1, w1))) bad_count += 1ull << 20;
            if (__float_as_uint(pm.x) != __float_as_uint(__fmul_rn(u0, w0)) || __float_as_uint(pm.y) != __float_as_uint(__fmul_rn(u1, w1))) bad_count += 1ull << 24;
            if (__float_as_uint(pf.x) != __float_as_uint(__fmaf_rn(u0, -w0, z0)) || __float_as_uint(pf.y) != __float_as_uint(__fmaf_rn(u1, -w1, z1))) bad_count += 1ull << 28;
            bool b3 = false;
            const float x2 = oc_rand_float(st, -94, 93, false);
            const float2 sq = oc_sqrt2<MathExact>(make_float2(x, x2), b3);
            if (__float_as_uint(sq.x) != __float_as_uint(__fsqrt_rn(x)) || __float_as_uint(sq.y) != __float_as_uint(__fsqrt_rn(x2))) bad_count += 1ull << 32;
            const float2 y0 = p_rcp(sq);
            const float2 inv = p_fma(y0, p_fma(y0, p_neg(sq), p_bc(1.0f)), y0);
            if (__float_as_uint(inv.x) != __float_as_uint(__frcp_rn(sq.x)) || __float_as_uint(inv.y) != __float_as_uint(__frcp_rn(sq.y))) bad_count += 1ull << 36;
            const float a3 = oc_rand_float(st, -70, 69, true);
            const float2 aa = make_float2(a, a3);
            const float2 q0 = p_mul(aa, inv);
            const float2 qq = p_fma(inv, p_fma(q0, p_neg(sq), aa), q0);
            if (a != 0.0f && __float_as_uint(qq.x) != __float_as_uint(__fdiv_rn(a, sq.x))) bad_count += 1ull << 40;
            if (__float_as_uint(qq.y) != __float_as_uint(__fdiv_rn(a3, sq.y))) bad_count += 1ull << 40;
        }
        // whole spring pair on cloth-like operands against the scalar intrinsic formula
        {
            auto jitter = [&](float base, float amp) { return base + amp * ((float)(oc_rng(st) & 0xffff) / 65536.0f - 0.5f); };
            const f3 mx = make_f3(jitter(1.6f, 1e-3f), jitter(4.99f, 1e-3f), jitter(0.0f, 1e-6f));
            const f3 mv = make_f3(jitter(0.0f, 1e-4f), jitter(-0.01f, 1e-4f), jitter(0.0f, 1e-5f));
            const f3 ax = make_f3(mx.x + jitter(0.2f, 1e-3f), mx.y + jitter(0.0f, 1e-3f), mx.z + jitter(0.2f, 1e-3f));
            const f3 bx = make_f3(mx.x - jitter(0.2f, 1e-3f), mx.y + jitter(0.0f, 1e-3f), mx.z + jitter(0.2f, 1e-3f));
            const f3 av = make_f3(jitter(0.0f, 1e-4f), jitter(-0.01f, 1e-4f), jitter(0.0f, 1e-5f));
            const f3 bv = make_f3(jitter(0.0f, 1e-4f), jitter(-0.01f, 1e-4f), jitter(0.0f, 1e-5f));
            OcPair3 qx, qv;
            qx.x = make_float2(ax.x, bx.x); qx.y = make_float2(ax.y, bx.y); qx.z = make_float2(ax.z, bx.z);
            qv.x = make_float2(av.x, bv.x); qv.y = make_float2(av.y, bv.y); qv.z = make_float2(av.z, bv.z);
            const float ra = jitter(0.2828f, 1e-4f), rb = jitter(0.2828f, 1e-4f);
            bool b4 = false;
            const OcPair3 g = oc_spring2<MathExact>(mx, mv, qx, qv, make_float2(ra, rb), p_bc(-50.75f), p_bc(-0.25f), one, b4);
            const f3 fa = oc_spring<MathExact>(mx, mv, ax, av, ra, -50.75f, -0.25f);
            const f3 fb = oc_spring<MathExact>(mx, mv, bx, bv, rb, -50.75f, -0.25f);
            if (!b4) {
                if (__float_as_uint(g.x.x) != __float_as_uint(fa.x) || __float_as_uint(g.y.x) != __float_as_uint(fa.y) || __float_as_uint(g.z.x) != __float_as_uint(fa.z)) bad_count += 1ull << 44;
                if (__float_as_uint(g.x.y) != __float_as_uint(fb.x) || __float_as_uint(g.y.y) != __float_as_uint(fb.y) || __float_as_uint(g.z.y) != __float_as_uint(fb.z)) bad_count += 1ull << 44;
            }
        }
    }
    if (bad_count) atomicAdd(out, bad_count);
}
extern "C" int oc_selftest_math(unsigned long long n, unsigned int seed, unsigned long long* mismatches)
{
    if (!mismatches) return oc_fail(OC_ERR_INVALID, "oc_selftest_math: null");
    unsigned long long* d = nullptr;
    OC_CUDA(cudaMalloc(&d, sizeof(*d)));
    OC_CUDA(cudaMemset(d, 0, sizeof(*d)));
    const int blocks = 1184, threads = 256;
    unsigned long long per = (n + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
    const float dts[4] = { 1 / 60.0f, 1 / 90.0f, 0.001f, 0.37f };
    for (int t = 0; t < 4; ++t) {
        oc_k_selftest<<<blocks, threads>>>((per + 3) / 4, seed + t, dts[t], d, 1.0f);
        OC_CUDA(cudaGetLastError());
    }
    OC_CUDA(cudaMemcpy(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost));
    OC_CUDA(cudaFree(d));
    return OC_OK;
}

extern "C" int oc_debug_counters(oc_cloth* c, unsigned long long out[4])
{
    if (!c || !out) return oc_fail(OC_ERR_INVALID, "oc_debug_counters: null");
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    OC_CUDA(cudaMemcpy(out, c->d_dbg, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    OC_CUDA(cudaMemset(c->d_dbg, 0, 4 * sizeof(unsigned long long)));
    return OC_OK;
}

// Development (OC_DEBUG=64 at oc_create): milliseconds after the start of the last oc_upload at which, per chunk, the
// H2D copy, the unpack, the step, the pack and the D2H copy finished: out[5 * chunk + stage]; returns the chunk count.
extern "C" int oc_debug_pipeline(oc_cloth* c, float* out, int max_chunks)
{
    if (!c || !out || !c->pipe.timed) return 0;
    cudaSetDevice(c->dev);
    int n = c->pipe.last_chunks < max_chunks ? c->pipe.last_chunks : max_chunks;
    for (int k = 0; k < n; ++k) {
        cudaEvent_t ev[5] = { c->pipe.ev_cp[k], c->pipe.ev_in[k], c->pipe.ev_step[k], c->pipe.ev_pk[k], c->pipe.ev_out[k] };
        for (int q = 0; q < 5; ++q) { float ms = -1.0f; if (cudaEventElapsedTime(&ms, c->pipe.ev_t0, ev[q]) != cudaSuccess) { ms = -1.0f; cudaGetLastError(); } out[5 * k + q] = ms; }
    }
    return n;
}

extern "C" int oc_debug_timeline(oc_cloth* c, unsigned long long* out, size_t n_words)
{
    if (!c || !out) return oc_fail(OC_ERR_INVALID, "oc_debug_timeline: null");
    if (n_words > (size_t)8 * OC_DBG_TL_CTAS) n_words = (size_t)8 * OC_DBG_TL_CTAS;
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    OC_CUDA(cudaMemcpy(out, c->d_dbg + OC_DBG_TL_BASE, n_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return OC_OK;
}

extern "C" size_t oc_sizeof_params(void) { return sizeof(oc_params); }

extern "C" int oc_spring_energy(oc_cloth* c, int cloth, double* energy)
{
    if (!c || !energy) return oc_fail(OC_ERR_INVALID, "oc_spring_energy: null");
    if (c->q.band) return oc_fail(OC_ERR_UNSUPPORTED, "oc_spring_energy: whole-cloth handles only");
    if (cloth < 0 || cloth >= c->p.batch) return oc_fail(OC_ERR_INVALID, "oc_spring_energy: cloth out of range");
    OC_CUDA(cudaSetDevice(c->dev));
    OC_CUDA(cudaMemsetAsync(c->d_energy, 0, sizeof(double), c->stream));
    long long n = (long long)c->p.nx * c->p.ny;
    oc_k_energy<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->k, c->buf[c->q.ia], cloth, c->p.ks_struct, c->p.ks_shear, c->p.ks_bend, c->d_energy);
    c->launches++;
    OC_CUDA(cudaGetLastError());
    OC_CUDA(cudaMemcpyAsync(energy, c->d_energy, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    OC_CUDA(cudaStreamSynchronize(c->stream));
    return OC_OK;
}

extern "C" const char* oc_version(void)
{
    static thread_local char buf[256];
    int ndev = 0, dev = 0;
    char name[128] = "none";
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && cudaGetDevice(&dev) == cudaSuccess) {
        cudaDeviceProp pr;
        if (cudaGetDeviceProperties(&pr, dev) == cudaSuccess) snprintf(name, sizeof(name), "%s sm_%d%d x%d", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
    } else {
        cudaGetLastError();
    }
    snprintf(buf, sizeof(buf), "opencloth_b200 abi=%d built-for=sm_100a devices=%d device=%s", OC_ABI_VERSION, ndev, name);
    return buf;
}
