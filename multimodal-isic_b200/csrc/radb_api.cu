// C-ABI of the radb engine (include/radb.h) over the sm_100a kernels in radb_kernels.cuh.
// Drop-in boundary for RadiomicExtractor.py:14-48 of rbuler/multimodal-isic; see include/radb.h
// for the reference interface each entry point replaces.  No CPU path: every call either
// launches the CUDA kernel on the caller's stream or returns an error code.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include "radb_host.h"
#include "radb_kernels.cuh"

static thread_local std::string g_err;

struct radb_handle {
    radb::Plan plan;
    int device;
    int smem_optin;     // max dynamic shared memory per block the device allows
    int smem_set;       // currently configured attribute value
    int64_t launches;
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

extern "C" int radb_create(const radb_settings* s, radb_handle** out)
{
    if (!s || !out) return fail(RADB_E_INVALID, "null argument");
    radb_handle* h = new radb_handle();
    std::string err;
    int rc = radb::make_plan(*s, h->plan, err);
    if (rc) { delete h; return fail(rc, err); }
    h->device = s->device;
    h->launches = 0;
    h->smem_set = 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        delete h;
        return cuda_fail(e != cudaSuccess ? e : cudaErrorNoDevice, "radb_create: no CUDA device (there is no CPU fallback)");
    }
    if (s->device < 0 || s->device >= ndev) { delete h; return fail(RADB_E_INVALID, "device ordinal out of range"); }
    e = cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaDeviceGetAttribute"); }
    *out = h;
    return RADB_OK;
}

extern "C" void radb_destroy(radb_handle* h) { delete h; }
extern "C" int radb_feature_count(const radb_handle* h) { return h ? h->plan.F : RADB_E_INVALID; }
extern "C" const char* radb_feature_name(const radb_handle* h, int i)
{
    if (!h || i < 0 || i >= h->plan.F) return nullptr;
    return h->plan.names[i].c_str();
}
extern "C" int radb_max_ng(const radb_handle* h) { return h ? h->plan.max_ng : RADB_E_INVALID; }
extern "C" int64_t radb_launch_count(const radb_handle* h) { return h ? h->launches : 0; }

extern "C" int radb_smem_bytes(const radb_handle* h, int H, int W, int dtype)
{
    if (!h) return RADB_E_INVALID;
    RadbParams p;
    std::string err;
    int rc = radb::fill_params(h->plan, H, W, dtype, p, err);
    if (rc) return fail(rc, err);
    return p.smem_total;
}

static int launch(radb_handle* h, RadbParams& p, int dtype, void* stream)
{
    (void)dtype;
    cudaError_t e;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != h->device) {
        e = cudaSetDevice(h->device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    }
    if (p.smem_total > h->smem_optin) return fail(RADB_E_SMEM, "shared memory request exceeds the device opt-in limit");
    if (p.smem_total > h->smem_set) {
        e = cudaFuncSetAttribute(radb_extract_kernel<unsigned char>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 p.smem_total);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        cudaFuncSetAttribute(radb_extract_kernel<unsigned char>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        h->smem_set = p.smem_total;
    }
    const long long maxgrid = 0x7fffffffLL;
    long long done = 0;
    while (done < p.B) {
        long long n = p.B - done < maxgrid ? p.B - done : maxgrid;
        RadbParams q = p;
        q.img = (const unsigned char*)p.img + done * p.img_stride;
        q.mask = p.mask + done * p.mask_stride;
        q.out = p.out + done * p.F;
        q.status = p.status + done;
        // debug buffers are only used with small batches (one launch)
        radb_extract_kernel<unsigned char><<<(unsigned)n, RADB_NT, p.smem_total, (cudaStream_t)stream>>>(q);
        h->launches++;
        done += n;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radb_extract_kernel launch");
    return RADB_OK;
}

static int setup(radb_handle* h, const void* img, int dtype, const uint8_t* mask, int64_t B, int H, int W,
                 int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status, RadbParams& p)
{
    if (!h || !img || !mask || !out || !status) return fail(RADB_E_INVALID, "null argument");
    if (B < 0) return fail(RADB_E_INVALID, "negative batch size");
    std::string err;
    int rc = radb::fill_params(h->plan, H, W, dtype, p, err);
    if (rc) return fail(rc, err);
    if (img_stride_b < (int64_t)H * W || mask_stride_b < (int64_t)H * W)
        return fail(RADB_E_INVALID, "patch stride smaller than the patch");
    p.img = img;
    p.mask = mask;
    p.img_stride = img_stride_b;
    p.mask_stride = mask_stride_b;
    p.out = out;
    p.status = status;
    p.B = B;
    // TMA bulk copies need 16-byte aligned sources and sizes
    p.use_tma = ((uintptr_t)img % 16 == 0) && ((uintptr_t)mask % 16 == 0) && (img_stride_b % 16 == 0) &&
                (mask_stride_b % 16 == 0) && (p.HW % 16 == 0);
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
    if (B > 0x7fffffffLL) return fail(RADB_E_INVALID, "debug batches must fit one launch");
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
