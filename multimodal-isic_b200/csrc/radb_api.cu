// C-ABI of the radb engine (include/radb.h) over the sm_100a kernels in radb_kernels.cuh.
// Drop-in boundary for RadiomicExtractor.py:14-48 of rbuler/multimodal-isic; see include/radb.h
// for the reference interface each entry point replaces.  No CPU path: every call either
// launches the CUDA kernel on the caller's stream or returns an error code.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "radb_host.h"
#include "radb_kernels.cuh"

static thread_local std::string g_err;

#define RADB_TAB_NINV 4096   // entries of the device 1/k^2 table (beyond it the kernels divide)
#define RADB_CHUNK_DEFAULT 65536  // patches per pass through the kernels (bounds the workspace; the thread-level reduction kernels want >= 5 waves)
// Default chunk of a handle, fixed when the handle is created (no process-global state): RADB_CHUNK (tuning knob,
// multiples of 4) or RADB_CHUNK_DEFAULT; radb_set_chunk overrides it per handle.
static int64_t radb_chunk_from_env()
{
    const char* e = getenv("RADB_CHUNK");
    const long v = e ? atol(e) : 0;
    return (v >= 4) ? (v - v % 4) : RADB_CHUNK_DEFAULT;
}

struct radb_handle {
    radb::Plan plan;
    int device;
    int smem_optin;              // max dynamic shared memory per block the device allows
    int smem_set[40];             // configured MaxDynamicSharedMemorySize per kernel
    int64_t launches;
    struct Ws { void* stream; unsigned char* p; size_t bytes; long long* meta; size_t meta_bytes; };
    std::vector<Ws> ws;          // per-patch records of one chunk, one workspace per CUDA stream
                                 // (launches on different streams may overlap; each owns its records)
    double* d_inv2;
    double* d_tlog;
    int64_t chunk;               // patches per chunk set by radb_set_chunk (0: chunk_default)
    int64_t chunk_default;       // RADB_CHUNK env / RADB_CHUNK_DEFAULT, read once at radb_create
    cudaStream_t red_stream;     // high-priority side stream of the GLCM / GLRLM reduction kernels
    cudaStream_t red_stream3;    // third side stream: the warp-per-angle kernel next to the Lanczos MCC kernel (many gray levels)
    cudaStream_t red_stream2;    // second side stream: first-order / GLDM / NGTDM / GLSZM / shape reductions (concurrent with the first)
    std::vector<cudaEvent_t> sync_events;  // build-done / reduce-done events of the two-stream pipeline (re-used)
    std::vector<cudaEvent_t> chunk_events;  // caller's completion events for the chunks of the next call (radb_set_chunk_events)
    bool profiling;              // record CUDA events around every kernel (radb_set_profiling)
    std::vector<cudaEvent_t> events;  // 4 per chunk: start, after build, after angle, after misc (pool, re-used across calls)
    size_t events_used;               // marks recorded since the last radb_kernel_ms
};

// Every entry point that launches work selects the handle's device and restores the caller's on return
// (a multi-GPU process must not find its current device changed by a library call).
struct DeviceGuard {
    int prev;
    bool switched;
    explicit DeviceGuard(int dev) : prev(-1), switched(false)
    {
        cudaGetDevice(&prev);
        if (prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
};

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return RADB_E_CUDA;
}

extern "C" const char* radb_last_error(void) { return g_err.c_str(); }
extern "C" const char* radb_version(void) { return "radb 0.1 (sm_100a)"; }

extern "C" void radb_destroy(radb_handle* h);

extern "C" int radb_create(const radb_settings* s, radb_handle** out)
{
    if (!s || !out) return fail(RADB_E_INVALID, "null argument");
    radb_handle* h = new radb_handle();
    std::string err;
    int rc = radb::make_plan(*s, h->plan, err);
    if (rc) { delete h; return fail(rc, err); }
    h->device = s->device;
    h->launches = 0;
    for (int i = 0; i < 40; i++) h->smem_set[i] = 0;
    h->d_inv2 = h->d_tlog = nullptr;
    h->profiling = false;
    h->red_stream = nullptr;
    h->red_stream2 = nullptr;
    h->red_stream3 = nullptr;
    h->chunk = 0;
    h->chunk_default = radb_chunk_from_env();
    h->events_used = 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        delete h;
        return cuda_fail(e != cudaSuccess ? e : cudaErrorNoDevice, "radb_create: no CUDA device (there is no CPU fallback)");
    }
    if (s->device < 0 || s->device >= ndev) { delete h; return fail(RADB_E_INVALID, "device ordinal out of range"); }
    e = cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaDeviceGetAttribute"); }
    {   // device tables for the reduction kernels
        int cur = -1;
        cudaGetDevice(&cur);
        cudaSetDevice(s->device);
        std::vector<double> inv2, tlog;
        radb::make_tables(RADB_TAB_NINV, inv2, tlog);
        e = cudaMalloc(&h->d_inv2, inv2.size() * 8);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_tlog, tlog.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_inv2, inv2.data(), inv2.size() * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_tlog, tlog.data(), tlog.size() * 8, cudaMemcpyHostToDevice);
        if (cur >= 0) cudaSetDevice(cur);
        if (e != cudaSuccess) { radb_destroy(h); return cuda_fail(e, "radb_create: table allocation"); }
    }
    *out = h;
    return RADB_OK;
}

extern "C" void radb_destroy(radb_handle* h)
{
    if (!h) return;
    for (auto& w : h->ws) {
        if (w.p) cudaFree(w.p);
        if (w.meta) cudaFree(w.meta);
    }
    if (h->d_inv2) cudaFree(h->d_inv2);
    if (h->d_tlog) cudaFree(h->d_tlog);
    for (auto ev : h->sync_events) cudaEventDestroy(ev);
    for (auto ev : h->events) cudaEventDestroy(ev);
    if (h->red_stream) cudaStreamDestroy(h->red_stream);
    if (h->red_stream2) cudaStreamDestroy(h->red_stream2);
    if (h->red_stream3) cudaStreamDestroy(h->red_stream3);
    delete h;
}

// Patches per pass through the kernels: bounds the workspace (records + wide-mode scratch).  Batches larger
// than one chunk are cut into equal chunks so that the two-stream pipeline below has balanced stages.
static int64_t chunk_for(const radb_handle* h, const RadbParams& p, int64_t B)
{
    const size_t per = (size_t)p.rec_bytes + (size_t)p.scr_bytes;
    // <= 1 GiB of records per slot; 4 GiB when a record is large (many gray levels: 1 MB per patch at 256 levels),
    // so that a chunk still fills the 148 SMs several times over
    int64_t n = (int64_t)(((size_t)(per > (256u << 10) ? 4 : 1) << 30) / per);
    const int64_t want = h->chunk > 0 ? h->chunk : h->chunk_default;
    if (n > want) n = want;
    n -= n % 4;  // keeps the 4 planes of an image (shared mask) in one chunk
    if (n < 4) n = 4;
    if (B <= n) return B;
    const int64_t k = (B + n - 1) / n;        // number of chunks
    int64_t eq = (B + k - 1) / k;             // equalised
    eq = (eq + 3) / 4 * 4;
    return eq < n ? eq : n;
}

// Chunk sizes of a dense call.  Default: equal chunks (chunk_for).  A decreasing schedule (fractions of the batch,
// RADB_CHUNK_SCHED="0.5,0.3,0.2" or the built-in one for large batches) keeps the early chunks large -- the
// thread-level reduction kernels want several waves -- and the LAST chunk small: its reductions are the only ones
// that do not run under a later chunk's build kernel.  Only when the caller has not fixed the chunk size
// (radb_set_chunk: equal chunks, which radb_chunk_rows / radb_set_chunk_events describe).  Returns the largest chunk.
static int64_t chunk_plan(const radb_handle* h, const RadbParams& p, int64_t B, std::vector<int64_t>& sizes)
{
    sizes.clear();
    const int64_t cap = chunk_for(h, p, B > 0 ? (int64_t)1 << 40 : 0);  // the cap itself (a huge batch is cut at the cap)
    static const char* env = getenv("RADB_CHUNK_SCHED");
    std::vector<double> fr;
    if (h->chunk <= 0 && B >= 32768 && !p.rows) {
        if (env && *env) {
            for (const char* q = env; *q;) {
                char* e;
                double f = strtod(q, &e);
                if (e == q) break;
                if (f > 0) fr.push_back(f);
                q = (*e == ',') ? e + 1 : e;
                if (*e && *e != ',') break;
            }
        }
    }
    if (fr.size() >= 2) {
        int64_t left = B;
        for (size_t i = 0; i < fr.size() && left > 0; i++) {
            int64_t n = (i + 1 == fr.size()) ? left : (int64_t)(B * fr[i] + 0.5);
            n = (n + 3) / 4 * 4;
            while (n > cap) { sizes.push_back(cap); left -= cap; n -= cap; }
            if (n > left) n = left;
            if (n > 0) { sizes.push_back(n); left -= n; }
        }
        while (left > 0) { int64_t n = left < cap ? left : cap; sizes.push_back(n); left -= n; }
    } else {
        const int64_t n = chunk_for(h, p, B);
        for (int64_t done = 0; done < B; done += n) sizes.push_back(B - done < n ? B - done : n);
    }
    int64_t mx = 0;
    for (int64_t n : sizes) mx = n > mx ? n : mx;
    return mx;
}

// Grow-only workspace: one record (+ scratch) per patch of a chunk, keyed by the stream the kernels run on.
// Batches of several chunks get two such slots: chunk i+1 is built while chunk i is being reduced.
static int ensure_ws(radb_handle* h, const RadbParams& p, int64_t B, void* stream, unsigned char** out)
{
    std::vector<int64_t> sizes;
    const int64_t n = chunk_plan(h, p, B, sizes);
    const size_t slots = B > n ? 2 : 1;
    const size_t need = slots * (size_t)n * ((size_t)p.rec_bytes + (size_t)p.scr_bytes);
    radb_handle::Ws* w = nullptr;
    for (auto& e : h->ws)
        if (e.stream == stream) w = &e;
    if (!w) {
        h->ws.push_back({stream, nullptr, 0, nullptr, 0});
        w = &h->ws.back();
    }
    if (need > w->bytes) {
        if (w->p) cudaFree(w->p);  // synchronises the device: nothing is still reading it
        w->p = nullptr;
        w->bytes = 0;
        cudaError_t e = cudaMalloc(&w->p, need);
        if (e != cudaSuccess) return cuda_fail(e, "workspace allocation");
        w->bytes = need;
    }
    *out = w->p;
    return RADB_OK;
}

extern "C" int radb_reserve(radb_handle* h, int H, int W, int dtype, int64_t max_batch, void* cuda_stream)
{
    if (!h || max_batch < 0) return fail(RADB_E_INVALID, "bad argument");
    RadbParams p;
    std::string err;
    int rc = radb::fill_params(h->plan, H, W, dtype, p, err);
    if (rc) return fail(rc, err);
    DeviceGuard guard(h->device);
    unsigned char* unused = nullptr;
    return ensure_ws(h, p, max_batch, cuda_stream, &unused);
}
extern "C" int radb_feature_count(const radb_handle* h) { return h ? h->plan.F : RADB_E_INVALID; }
extern "C" const char* radb_feature_name(const radb_handle* h, int i)
{
    if (!h || i < 0 || i >= h->plan.F) return nullptr;
    return h->plan.names[i].c_str();
}
extern "C" int radb_max_ng(const radb_handle* h) { return h ? h->plan.max_ng : RADB_E_INVALID; }
extern "C" int64_t radb_launch_count(const radb_handle* h) { return h ? h->launches : 0; }

extern "C" int64_t radb_chunk_rows(const radb_handle* h, int H, int W, int dtype, int64_t B)
{
    if (!h || B < 0) return RADB_E_INVALID;
    RadbParams p;
    std::string err;
    int rc = radb::fill_params(h->plan, H, W, dtype, p, err);
    if (rc) return fail(rc, err);
    return chunk_for(h, p, B);
}

extern "C" int radb_set_chunk_events(radb_handle* h, void* const* cuda_events, int n)
{
    if (!h || n < 0 || (n > 0 && !cuda_events)) return fail(RADB_E_INVALID, "radb_set_chunk_events: bad arguments");
    h->chunk_events.assign((const cudaEvent_t*)cuda_events, (const cudaEvent_t*)cuda_events + n);
    return RADB_OK;
}

extern "C" int radb_smem_bytes(const radb_handle* h, int H, int W, int dtype)
{
    if (!h) return RADB_E_INVALID;
    RadbParams p;
    std::string err;
    int rc = radb::fill_params(h->plan, H, W, dtype, p, err);
    if (rc) return fail(rc, err);
    return p.smem_total;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel (per device), not to a handle: two handles with
// different gray-level bounds share the same kernel instances, so the configured maximum is tracked per (device,
// kernel) for the whole process and only ever raised.
static int g_smem_set[64][40];
template <typename K>
static int set_smem(radb_handle* h, K kernel, int which, int bytes)
{
    if (bytes > h->smem_optin) return fail(RADB_E_SMEM, "shared memory request exceeds the device opt-in limit");
    int* cur = &g_smem_set[h->device & 63][which];
    if (bytes > *cur) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        *cur = bytes;
    }
    return RADB_OK;
}

#define RADB_CHECK_LAUNCH(name)                                                                        \
    do {                                                                                               \
        cudaError_t le_ = cudaPeekAtLastError();                                                       \
        if (le_ != cudaSuccess) { cudaGetLastError(); return cuda_fail(le_, name " launch"); }        \
    } while (0)

// One pass = three kernels over a chunk of patches, stream-ordered, sharing the record workspace.
static int launch(radb_handle* h, RadbParams& p, int dtype, void* stream)
{
    std::vector<cudaEvent_t> user_events;
    user_events.swap(h->chunk_events);  // one-shot: they belong to this launch, whatever its outcome
    cudaError_t e;
    DeviceGuard guard(h->device);
    const bool dbg = p.dbg_levels || p.dbg_glcm || p.dbg_glrlm || p.dbg_glszm || p.dbg_gldm || p.dbg_ng;
    typedef void (*build_fn)(const RadbParams);
    build_fn build = nullptr;
#define RADB_PICK(PT) (p.wide ? (dbg ? radb_build_kernel<PT, true, true> : radb_build_kernel<PT, false, true>) \
                              : (dbg ? radb_build_kernel<PT, true, false> : radb_build_kernel<PT, false, false>))
#define RADB_PICK16(PT) (dbg ? radb_build_kernel<PT, true, true, true> : radb_build_kernel<PT, false, true, true>)
    const bool l16 = p.lev_bytes == 2;  // > 255 gray levels: u16 level image (big mode, never uint8 pixels)
    switch (dtype) {
        case RADB_DTYPE_U8: build = RADB_PICK(unsigned char); break;
        case RADB_DTYPE_U16: build = l16 ? RADB_PICK16(unsigned short) : RADB_PICK(unsigned short); break;
        case RADB_DTYPE_F32: build = l16 ? RADB_PICK16(float) : RADB_PICK(float); break;
        case RADB_DTYPE_F64: build = l16 ? RADB_PICK16(double) : RADB_PICK(double); break;
        default: return fail(RADB_E_INVALID, "unknown pixel dtype");
    }
#undef RADB_PICK
#undef RADB_PICK16
    // the headline configuration has a compile-time specialisation of the build kernel (radb_kernels.cuh: FAST)
    static const bool no_fast = getenv("RADB_NO_FAST") != nullptr;  // A/B switch
    const bool fast = !no_fast && dtype == RADB_DTYPE_U8 && !dbg && !p.wide && !p.big && !l16 && p.vec4 && p.lev4 && p.use_tma &&
                      p.xo == 4 && p.bw_int != 0 && p.bin_count <= 0 && p.n_angles == 4 && p.symmetric &&
                      p.alpha == 0 && p.glcm_pad && p.glrlm_dense == 16 && p.off_glcm >= 0 && p.off_gldm >= 0 && p.off_glrlm >= 0 &&
                      p.off_glszm >= 0 && p.off_ngtdm >= 0 && p.ang_y[0] == 1 && p.ang_x[0] == 1 && p.ang_y[1] == 0 &&
                      p.ang_x[1] == 1 && p.ang_y[2] == -1 && p.ang_x[2] == 1 && p.ang_y[3] == 1 && p.ang_x[3] == 0;
    if (fast) build = p.o_runs >= 0 ? radb_build_kernel<unsigned char, false, false, false, 1> : radb_build_kernel<unsigned char, false, false, false, 2>;
    int rc = set_smem(h, build, fast ? (p.o_runs >= 0 ? 39 : 38) : (l16 ? 25 + dtype * 2 + (dbg ? 1 : 0) : 9 + dtype * 4 + (p.wide ? 2 : 0) + (dbg ? 1 : 0)), p.smem_total);
    static const bool no_lane = getenv("RADB_NO_LANE") != nullptr;  // A/B switch: force the warp-per-angle kernel
    if (no_lane) p.use_lane = 0;
    if (!rc) rc = p.use_lane ? set_smem(h, radb_angle_lane_kernel, 3, p.l_smem_total) : set_smem(h, radb_angle_kernel, 1, p.a_smem_total);
    if (!rc && p.use_lane == 2) rc = set_smem(h, radb_mcc_g8_kernel, 5, p.g8_smem_total);
    if (p.use_lane) p.use_lanczos = 0;  // (RADB_NO_LANE forces the warp-per-angle kernel; its layout was made for Lanczos or dense)
    if (!rc && p.use_lanczos) rc = set_smem(h, radb_mcc_lanczos_kernel, 6, p.z_smem_total);
    if (!rc && p.off_shape >= 0) rc = set_smem(h, radb_shape_kernel, 8, p.s_smem_total);
    if (!rc) rc = set_smem(h, radb_misc_kernel, 2, p.m_smem_total);
    p.only_big_ovf = (!no_lane && p.ml_smem_total <= 96 * 1024) ? 1 : 0;
    if (!rc && p.only_big_ovf) rc = set_smem(h, radb_misc_lane_kernel, 4, p.ml_smem_total);
    unsigned char* wsp = nullptr;
    if (!rc) rc = ensure_ws(h, p, p.B, stream, &wsp);
    if (rc) return rc;
    std::vector<int64_t> sizes;
    const long long chunk = chunk_plan(h, p, p.B, sizes);  // largest chunk = slot size
    const size_t slot_bytes = (size_t)chunk * ((size_t)p.rec_bytes + (size_t)p.scr_bytes);
    p.g_inv2 = h->d_inv2;
    p.g_tlog = h->d_tlog;
    if (p.ninv > RADB_TAB_NINV) p.ninv = RADB_TAB_NINV;
    cudaStream_t st = (cudaStream_t)stream;
    // Two-stream pipeline for batches of several chunks: the build kernel of chunk i+1 (issue-bound) runs on
    // the caller's stream while the reduction kernels of chunk i (latency-bound, few warps) run on a
    // high-priority side stream and fill the SMs' idle issue slots.  Records alternate between two workspace
    // slots; events order build -> reduce per chunk and reduce(i) -> build(i+2) per slot; the caller's stream
    // waits for the last reduction, so the call stays stream-ordered for the caller.  Per-kernel timing
    // (radb_set_profiling) needs the kernels back to back: profiling runs serialise on the caller's stream.
    static const bool no_overlap = getenv("RADB_NO_OVERLAP") != nullptr;
    const long long nchunks = (long long)sizes.size();
    // (single-chunk calls of at least 4096 patches fork too: the two reduction families are latency-bound kernels with
    // few resident warps and run side by side)
    const bool piped = (nchunks > 1 || p.B >= 4096) && !h->profiling && !no_overlap;
    if (piped) {
        if (!h->red_stream) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            e = cudaStreamCreateWithPriority(&h->red_stream, cudaStreamNonBlocking, hi);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->red_stream2, cudaStreamNonBlocking, hi);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->red_stream3, cudaStreamNonBlocking, hi);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithPriority");
        }
        while ((long long)h->sync_events.size() < 4 * nchunks) {
            cudaEvent_t ev;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
            h->sync_events.push_back(ev);
        }
    }
    cudaStream_t rs = piped ? h->red_stream : st;    // GLCM / GLRLM (+ MCC) reductions
    cudaStream_t ms = piped ? h->red_stream2 : st;   // first-order / GLDM / NGTDM / GLSZM / shape reductions
    cudaStream_t as = (piped && p.use_lanczos) ? h->red_stream3 : rs;  // warp-per-angle kernel beside the Lanczos kernel
    long long done = 0;
    for (long long c = 0; done < p.B; c++) {
        const long long n = sizes[(size_t)c];
        RadbParams q = p;
        q.ws = wsp + (size_t)(c & 1) * (nchunks > 1 ? slot_bytes : 0);
        q.ws_scr = q.ws + (size_t)chunk * (size_t)p.rec_bytes;
        if (p.rows) {  // ragged group: the index lists advance, the pools and the output stay put
            q.img_off = p.img_off + done;
            q.mask_off = p.mask_off + done;
            q.rows = p.rows + done;
        } else {
            q.img = (const unsigned char*)p.img + done * p.img_stride;
            q.mask = p.mask + (done / p.mask_group) * p.mask_stride;  // chunks are multiples of mask_group
            q.out = p.out + done * p.F;
            q.status = p.status + done;
        }
        q.B = n;
        if (done) {  // debug buffers (parity tests) advance with the chunk
            const long long HW = p.HW, NA = p.n_angles, NG = p.max_ng;
            if (q.dbg_levels) q.dbg_levels += done * HW;
            if (q.dbg_glcm) q.dbg_glcm += done * NA * NG * NG;
            if (q.dbg_glrlm) q.dbg_glrlm += done * NA * NG * p.nr;
            if (q.dbg_glszm) q.dbg_glszm += done * NG * HW;
            if (q.dbg_gldm) q.dbg_gldm += done * NG * (2 * NA + 1);
            if (q.dbg_ngn) q.dbg_ngn += done * NG;
            if (q.dbg_ngs) q.dbg_ngs += done * NG;
            if (q.dbg_ng) q.dbg_ng += done;
        }
        auto mark = [&]() {
            if (!h->profiling) return;
            if (h->events_used == h->events.size()) {  // grow the pool; events are re-used by later calls
                cudaEvent_t ev;
                if (cudaEventCreate(&ev) != cudaSuccess) return;
                h->events.push_back(ev);
            }
            cudaEventRecord(h->events[h->events_used++], st);
        };
        if (piped && c >= 2) {  // slot free again: both reduction families of chunk c - 2 are done
            cudaStreamWaitEvent(st, h->sync_events[4 * (c - 2) + 1], 0);
            cudaStreamWaitEvent(st, h->sync_events[4 * (c - 2) + 2], 0);
        }
        mark();
        build<<<(unsigned)n, RADB_NTB, p.smem_total, st>>>(q);
        RADB_CHECK_LAUNCH("radb_build_kernel");
        mark();
        if (piped) {
            cudaEventRecord(h->sync_events[4 * c], st);
            cudaStreamWaitEvent(rs, h->sync_events[4 * c], 0);
            cudaStreamWaitEvent(ms, h->sync_events[4 * c], 0);
        }
        const bool angle_classes = p.off_glcm >= 0 || p.off_glrlm >= 0;  // nothing to reduce per angle otherwise
        if (p.use_lane == 2 && p.off_glcm >= 0) {
            const int na = p.n_angles;
            const long long g8_warps = (na == 1 || na == 2 || na == 4) ? (n * na + 3) / 4 : n;  // 4 (patch, angle) tasks per warp
            radb_mcc_g8_kernel<<<(unsigned)((g8_warps + RADB_NTM / 32 - 1) / (RADB_NTM / 32)), RADB_NTM, p.g8_smem_total, rs>>>(q);
            RADB_CHECK_LAUNCH("radb_mcc_g8_kernel");
            h->launches += 1;
        }
        if (p.use_lanczos && p.off_glcm >= 0) {
            radb_mcc_lanczos_kernel<<<(unsigned)(n * p.n_angles), RADB_NTZ, p.z_smem_total, rs>>>(q);
            RADB_CHECK_LAUNCH("radb_mcc_lanczos_kernel");
            h->launches += 1;
        }
        if (as != rs) cudaStreamWaitEvent(as, h->sync_events[4 * c], 0);
        if (!angle_classes)
            h->launches -= 1;
        else if (p.use_lane)
            radb_angle_lane_kernel<<<(unsigned)((n * p.l_nap + RADB_NTL - 1) / RADB_NTL), RADB_NTL, p.l_smem_total, rs>>>(q);
        else
            radb_angle_kernel<<<(unsigned)n, RADB_NT, p.a_smem_total, as>>>(q);
        RADB_CHECK_LAUNCH(p.use_lane ? "radb_angle_lane_kernel" : "radb_angle_kernel");
        if (as != rs) {  // rs joins the third stream: everything after this point on rs follows the angle kernel too
            cudaEventRecord(h->sync_events[4 * c + 3], as);
            cudaStreamWaitEvent(rs, h->sync_events[4 * c + 3], 0);
        }
        if (p.use_lanczos && p.off_glcm >= 0) {  // MCC column: after the Lanczos kernel (rs) and the angle kernel (as)
            radb_mcc_combine_kernel<<<(unsigned)((n + 127) / 128), 128, 0, rs>>>(q);
            RADB_CHECK_LAUNCH("radb_mcc_combine_kernel");
            h->launches += 1;
        }
        mark();
        if (p.only_big_ovf) {
            radb_misc_lane_kernel<<<(unsigned)((n + RADB_NT - 1) / RADB_NT), RADB_NT, p.ml_smem_total, ms>>>(q);
            RADB_CHECK_LAUNCH("radb_misc_lane_kernel");
            h->launches += 1;
        }
        radb_misc_kernel<<<(unsigned)n, RADB_NT, p.m_smem_total, ms>>>(q);
        RADB_CHECK_LAUNCH("radb_misc_kernel");
        if (p.off_shape >= 0) {
            radb_shape_kernel<<<(unsigned)n, RADB_NT, p.s_smem_total, ms>>>(q);
            h->launches += 1;
        }
        mark();
        if (piped) {
            cudaEventRecord(h->sync_events[4 * c + 1], rs);
            cudaEventRecord(h->sync_events[4 * c + 2], ms);
        }
        if (c < (long long)user_events.size() && user_events[c]) {
            // rows [done, done + n) are final once both reduction families of this chunk are: the second family's
            // stream joins the first one's event, then carries the caller's event
            if (piped) cudaStreamWaitEvent(ms, h->sync_events[4 * c + 1], 0);
            cudaEventRecord(user_events[c], ms);
        }
        h->launches += 3;
        done += n;
    }
    if (piped) {  // the side streams are in order: joining their last events joins everything
        cudaStreamWaitEvent(st, h->sync_events[4 * (nchunks - 1) + 1], 0);
        cudaStreamWaitEvent(st, h->sync_events[4 * (nchunks - 1) + 2], 0);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radb kernel launch");
    return RADB_OK;
}

static int setup(radb_handle* h, const void* img, int dtype, const uint8_t* mask, int64_t B, int H, int W,
                 int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status, RadbParams& p, int mask_bits = 0)
{
    if (!h || !img || !mask || !out || !status) return fail(RADB_E_INVALID, "null argument");
    if (B < 0) return fail(RADB_E_INVALID, "negative batch size");
    std::string err;
    int rc = radb::fill_params(h->plan, H, W, dtype, p, err);
    if (rc) return fail(rc, err);
    const int64_t mask_bytes = mask_bits ? ((int64_t)H * W + 7) / 8 : (int64_t)H * W;
    if (img_stride_b < (int64_t)H * W * p.pix_bytes || mask_stride_b < mask_bytes)
        return fail(RADB_E_INVALID, "patch stride smaller than the patch");
    p.mask_bits = mask_bits ? 1 : 0;
    p.img = img;
    p.mask = mask;
    p.img_stride = img_stride_b;
    p.mask_stride = mask_stride_b;
    p.out = out;
    p.status = status;
    p.B = B;
    // TMA bulk copies need 16-byte aligned sources and sizes
    p.use_tma = !p.wide && ((uintptr_t)img % 16 == 0) && ((uintptr_t)mask % 16 == 0) && (img_stride_b % 16 == 0) &&
                (mask_stride_b % 16 == 0) && (p.HW % 16 == 0) && (mask_bytes % 16 == 0) && (img_stride_b % p.pix_bytes == 0);
    return RADB_OK;
}

extern "C" int radb_extract(radb_handle* h, const void* img, int dtype, const uint8_t* mask, int64_t B, int H,
                            int W, int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status,
                            void* cuda_stream)
{
    RadbParams p;
    int rc = setup(h, img, dtype, mask, B, H, W, img_stride_b, mask_stride_b, out, status, p);
    if (rc) return rc;
    if (B == 0) return RADB_OK;
    return launch(h, p, dtype, cuda_stream);
}

// Same as radb_extract with bit-packed masks: `mask_bits` holds, per patch, ceil(H*W/8) bytes whose bit i (LSB
// first) says whether pixel i belongs to the ROI (what `mask == label` would give, RadiomicExtractor.py:38);
// `mask_stride_b` = bytes between the streams of consecutive patches.  The kernels read the bits directly: a mask
// crosses the host link and HBM at 1/8 of the bytes (radb_pack_masks_host produces this layout).
extern "C" int radb_extract_packed(radb_handle* h, const void* img, int dtype, const uint8_t* mask_bits, int64_t B, int H,
                                   int W, int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status,
                                   void* cuda_stream)
{
    RadbParams p;
    int rc = setup(h, img, dtype, mask_bits, B, H, W, img_stride_b, mask_stride_b, out, status, p, 1);
    if (rc) return rc;
    if (B == 0) return RADB_OK;
    return launch(h, p, dtype, cuda_stream);
}

// Variable-size batches: patches are grouped by (H, W) and every group runs the same kernels as radb_extract,
// reading its patches through per-patch byte offsets and writing rows in input order.
extern "C" int radb_extract_ragged(radb_handle* h, const void* img_pool, int dtype, const uint8_t* mask_pool, int64_t n,
                                   const int64_t* img_off, const int64_t* mask_off, const int32_t* hw, double* out,
                                   int32_t* status, void* cuda_stream)
{
    if (!h || !img_pool || !mask_pool || !out || !status) return fail(RADB_E_INVALID, "null argument");
    if (n < 0) return fail(RADB_E_INVALID, "negative batch size");
    if (n == 0) return RADB_OK;
    if (!img_off || !mask_off || !hw) return fail(RADB_E_INVALID, "null argument");
    std::vector<radb::RaggedGroup> groups;
    std::string err;
    int rc = radb::group_ragged(n, hw, groups, err);
    if (rc) return fail(rc, err);
    DeviceGuard guard(h->device);
    // device copy of the index lists: [img_off | mask_off | rows] per group, group after group
    radb_handle::Ws* w = nullptr;
    for (auto& e : h->ws)
        if (e.stream == cuda_stream) w = &e;
    if (!w) {
        h->ws.push_back({cuda_stream, nullptr, 0, nullptr, 0});
        w = &h->ws.back();
    }
    const size_t need = (size_t)n * 3 * sizeof(long long);
    if (need > w->meta_bytes) {
        if (w->meta) cudaFree(w->meta);
        w->meta = nullptr;
        w->meta_bytes = 0;
        cudaError_t e = cudaMalloc(&w->meta, need);
        if (e != cudaSuccess) return cuda_fail(e, "ragged index allocation");
        w->meta_bytes = need;
    }
    std::vector<long long> host((size_t)n * 3);
    {
        size_t o = 0;
        for (const auto& g : groups) {
            const size_t k = g.idx.size();
            for (size_t j = 0; j < k; j++) {
                host[o + j] = img_off[g.idx[j]];
                host[o + k + j] = mask_off[g.idx[j]];
                host[o + 2 * k + j] = g.idx[j];
            }
            o += 3 * k;
        }
    }
    long long* meta = w->meta;  // (ensure_ws below may reallocate h->ws entries, not this buffer)
    cudaError_t e = cudaMemcpyAsync(meta, host.data(), need, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return cuda_fail(e, "ragged index upload");
    // pageable source: the call returns once `host` has been staged, so the vector may die with this scope
    size_t o = 0;
    for (const auto& g : groups) {
        const long long k = (long long)g.idx.size();
        RadbParams p;
        const long long HWb = (long long)g.H * g.W;
        rc = setup(h, img_pool, dtype, mask_pool, k, g.H, g.W, HWb * (dtype == RADB_DTYPE_U8 ? 1 : dtype == RADB_DTYPE_U16 ? 2 : dtype == RADB_DTYPE_F32 ? 4 : 8),
                   HWb, out, status, p);
        if (rc) return rc;
        p.img_off = meta + o;
        p.mask_off = meta + o + k;
        p.rows = meta + o + 2 * k;
        bool aligned = p.use_tma;
        for (long long j = 0; j < k && aligned; j++)
            aligned = (img_off[g.idx[j]] % 16 == 0) && (mask_off[g.idx[j]] % 16 == 0);
        p.use_tma = aligned ? 1 : 0;
        for (long long j = 0; j < k; j++)
            if (img_off[g.idx[j]] % p.pix_bytes != 0) return fail(RADB_E_INVALID, "ragged batch: pixel offset not aligned to the pixel size");
        rc = launch(h, p, dtype, cuda_stream);
        if (rc) return rc;
        o += 3 * k;
    }
    return RADB_OK;
}

extern "C" int radb_debug_matrices(radb_handle* h, const void* img, int dtype, const uint8_t* mask, int64_t B,
                                   int H, int W, int64_t img_stride_b, int64_t mask_stride_b, double* out,
                                   int32_t* status, int32_t* levels, int32_t* glcm, int32_t* glrlm,
                                   int32_t* glszm, int32_t* gldm, int32_t* ngtdm_n, double* ngtdm_s, int32_t* ng,
                                   void* cuda_stream)
{
    RadbParams p;
    int rc = setup(h, img, dtype, mask, B, H, W, img_stride_b, mask_stride_b, out, status, p);
    if (rc) return rc;
    if (B == 0) return RADB_OK;
    p.dbg_levels = levels;
    p.dbg_glcm = glcm;
    p.dbg_glrlm = glrlm;
    p.dbg_glszm = glszm;
    p.dbg_gldm = gldm;
    p.dbg_ngn = ngtdm_n;
    p.dbg_ngs = ngtdm_s;
    p.dbg_ng = ng;
    return launch(h, p, dtype, cuda_stream);
}

extern "C" int radb_set_chunk(radb_handle* h, int64_t patches)
{
    if (!h || patches < 0) return fail(RADB_E_INVALID, "bad argument");
    h->chunk = patches ? (patches + 3) / 4 * 4 : 0;
    return RADB_OK;
}

extern "C" int radb_set_profiling(radb_handle* h, int on)
{
    if (!h) return fail(RADB_E_INVALID, "null handle");
    h->events_used = 0;
    h->profiling = on != 0;
    return RADB_OK;
}

extern "C" int radb_kernel_ms(radb_handle* h, double* ms3)
{
    if (!h || !ms3) return fail(RADB_E_INVALID, "null argument");
    ms3[0] = ms3[1] = ms3[2] = 0.0;
    for (size_t i = 0; i + 3 < h->events_used; i += 4) {
        cudaError_t e = cudaEventSynchronize(h->events[i + 3]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
        for (int k = 0; k < 3; k++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->events[i + k], h->events[i + k + 1]);
            ms3[k] += ms;
        }
    }
    h->events_used = 0;
    return RADB_OK;
}

// RadiomicExtractor.extract_radiomics (RadiomicExtractor.py:23-55) for a batch of decoded records:
// interleaved BGR images + one mask each -> gray/R/G/B planes (device scratch `planes`, 4*H*W bytes
// per image) -> 4 executes per image sharing the mask.  out: [n_images*4][F] in gray, R, G, B order.
extern "C" int radb_extract_bgr(radb_handle* h, const uint8_t* bgr, const uint8_t* mask, int64_t n_images, int H,
                                int W, uint8_t* planes, double* out, int32_t* status, void* cuda_stream)
{
    if (!h || !bgr || !mask || !planes || !out || !status) return fail(RADB_E_INVALID, "null argument");
    if (n_images < 0) return fail(RADB_E_INVALID, "negative batch size");
    if (n_images == 0) return RADB_OK;
    RadbParams p;
    const int64_t HW = (int64_t)H * W;
    int rc = setup(h, planes, RADB_DTYPE_U8, mask, n_images * 4, H, W, HW, HW, out, status, p);
    if (rc) return rc;
    p.mask_group = 4;
    DeviceGuard guard(h->device);
    const long long threads = n_images * ((HW + 3) / 4);
    radb_bgr_planes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(bgr, planes, n_images, HW);
    h->launches += 1;
    return launch(h, p, RADB_DTYPE_U8, cuda_stream);
}

// RadiomicExtractor.py:34-35 on the device: n masks [sH][sW] -> [dH][dW], bit-exact with cv2.resize(INTER_NEAREST).
extern "C" int radb_resize_mask(radb_handle* h, const uint8_t* src, int64_t n, int sH, int sW, uint8_t* dst, int dH, int dW,
                                void* cuda_stream)
{
    if (!h || !src || !dst || n < 0) return fail(RADB_E_INVALID, "bad argument");
    if (sH < 1 || sW < 1 || dH < 1 || dW < 1) return fail(RADB_E_INVALID, "mask sizes must be positive");
    if (n == 0) return RADB_OK;
    DeviceGuard guard(h->device);
    // cv2 (resize -> resizeNN): inv_scale = (double)dsize / ssize, source index = floor(dst * (1. / inv_scale))
    const double ify = 1.0 / ((double)dH / (double)sH), ifx = 1.0 / ((double)dW / (double)sW);
    const long long threads = n * (long long)dH * dW;
    radb_resize_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(src, sH, sW, dst, dH, dW, n, ify, ifx);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radb_resize_mask launch");
    return RADB_OK;
}

// Device half of the packed-mask transfer path (host half: radb_hostpack.cpp).
extern "C" int radb_unpack_mask(radb_handle* h, const uint8_t* packed, int64_t n_bytes, uint8_t* mask, void* cuda_stream)
{
    if (!h || !packed || !mask || n_bytes < 0) return fail(RADB_E_INVALID, "bad argument");
    if (n_bytes == 0) return RADB_OK;
    DeviceGuard guard(h->device);
    const long long threads = (n_bytes + 15) / 16;
    radb_unpack_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(packed, n_bytes, h->plan.s.label, mask);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radb_unpack_mask launch");
    return RADB_OK;
}

// imageType filters of the pyradiomics parameter file (params.yml:141-144) for uint8 images:
// type 1 Square, 2 SquareRoot, 3 Logarithm, 4 Exponential.  img [n][H*W] uint8 -> out [n][H*W] float64;
// `mx` is an int32 [n] device scratch (per-image maximum).
extern "C" int radb_derive_image(radb_handle* h, const uint8_t* img, int64_t n_images, int64_t HW, int type,
                                 double* out, int32_t* mx, void* cuda_stream)
{
    if (!h || !img || !out || !mx) return fail(RADB_E_INVALID, "null argument");
    if (type < 1 || type > 4) return fail(RADB_E_UNSUPPORTED, "image type not implemented (1 Square, 2 SquareRoot, 3 Logarithm, 4 Exponential)");
    if (n_images <= 0 || HW <= 0) return RADB_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DeviceGuard guard(h->device);
    cudaError_t e = cudaMemsetAsync(mx, 0, (size_t)n_images * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    const long long t1 = n_images * ((HW + 3) / 4), t2 = n_images * HW;
    radb_image_max_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(img, n_images, HW, mx);
    radb_derive_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(img, n_images, HW, mx, type, out);
    h->launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radb_derive_image launch");
    return RADB_OK;
}

// Filtered image types of the parameter file (params.yml:138-140,145; pyradiomics imageoperations.getGradientImage,
// getLoGImage, getWaveletImage) for uint8 images [n][H][W] (DEVICE pointers, stream-ordered):
//   type 5 Gradient          -> out float32 [n][H][W]                              (scratch unused)
//   type 6 LoG, param=sigma  -> out float32 [n][H][W]; scratch >= n*H*W*12 bytes   (needs H, W >= 4)
//   type 7 Wavelet (coif1, level 1): flags bit 0 set = transform along x only (force2D on a 2-D array)
//                              -> out float64 [n][2][H][W] = wavelet-H, wavelet-L;
//                            flags 0 -> out float64 [n][4][H][W] = wavelet-LH, -HL, -HH, -LL;
//                              scratch >= n*2*(H+1)*(W+1)*8 bytes
extern "C" int radb_filter_image(radb_handle* h, const uint8_t* img, int64_t n_images, int H, int W, int type, double param,
                                 int flags, void* out, void* scratch, void* cuda_stream)
{
    if (!h || !img || !out) return fail(RADB_E_INVALID, "null argument");
    if (n_images < 0 || H < 1 || W < 1) return fail(RADB_E_INVALID, "bad image size");
    if (n_images == 0) return RADB_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DeviceGuard guard(h->device);
    const long long HW = (long long)H * W, npx = n_images * HW;
    const int T = 256;
    if (type == RADB_IT_GRADIENT) {
        radb_gradient_kernel<<<(unsigned)((npx + T - 1) / T), T, 0, st>>>(img, n_images, H, W, (float*)out);
        h->launches += 1;
    } else if (type == RADB_IT_WAVELET) {
        if (flags & 1) {
            radb_wavelet_x_kernel<<<(unsigned)((npx + T - 1) / T), T, 0, st>>>(img, n_images, H, W, (double*)out);
            h->launches += 1;
        } else {
            if (!scratch) return fail(RADB_E_INVALID, "wavelet: scratch buffer required");
            const long long npp = n_images * (long long)(H + (H & 1)) * (W + (W & 1));
            radb_wavelet_rows_kernel<<<(unsigned)((npp + T - 1) / T), T, 0, st>>>(img, n_images, H, W, (double*)scratch);
            radb_wavelet_cols_kernel<<<(unsigned)((npx + T - 1) / T), T, 0, st>>>((const double*)scratch, n_images, H, W, (double*)out);
            h->launches += 2;
        }
    } else if (type == RADB_IT_LOG) {
        if (!scratch) return fail(RADB_E_INVALID, "LoG: scratch buffer required");
        if (H < 4 || W < 4) return fail(RADB_E_INVALID, "LoG: ITK's recursive Gaussian needs at least 4 pixels per axis");
        if (!(param > 0)) return fail(RADB_E_INVALID, "LoG: sigma must be > 0");
        RadbIir c0, c2;
        radb::deriche_coefficients(param, 0, c0.N, c0.D, c0.M, c0.BN, c0.BM);
        radb::deriche_coefficients(param, 2, c2.N, c2.D, c2.M, c2.BN, c2.BM);
        double* scr = (double*)scratch;            // anti-causal pass of the line being filtered
        float* tmp = (float*)(scr + npx);          // second-derivative image between the two passes of a dimension
        const long long rows = n_images * H, cols = n_images * W;
        const int TL = 64;
        // ITK dimension 0 (x): d2/dx2 along rows, smoothing along columns -> out
        radb_iir_u8_kernel<<<(unsigned)((rows + TL - 1) / TL), TL, 0, st>>>(img, tmp, scr, n_images, H, W, 0, c2, 0);
        radb_iir_f32_kernel<<<(unsigned)((cols + TL - 1) / TL), TL, 0, st>>>(tmp, (float*)out, scr, n_images, H, W, 1, c0, 0);
        // ITK dimension 1 (y): d2/dy2 along columns, smoothing along rows, accumulated in float32
        radb_iir_u8_kernel<<<(unsigned)((cols + TL - 1) / TL), TL, 0, st>>>(img, tmp, scr, n_images, H, W, 1, c2, 0);
        radb_iir_f32_kernel<<<(unsigned)((rows + TL - 1) / TL), TL, 0, st>>>(tmp, (float*)out, scr, n_images, H, W, 0, c0, 1);
        h->launches += 4;
    } else {
        return fail(RADB_E_UNSUPPORTED, "radb_filter_image: type must be 5 (Gradient), 6 (LoG) or 7 (Wavelet)");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radb_filter_image launch");
    return RADB_OK;
}
