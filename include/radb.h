/* radb -- C ABI of the B200-native radiomic feature engine (libradb_b200.so).
 *
 * This is the drop-in boundary for the radiomic-extraction hot path of rbuler/multimodal-isic.
 * The reference has no FFI of its own on this path: it reaches its native code through the
 * third-party pyradiomics package.  Each entry point below names the reference-side interface
 * it stands in for (file:line under /root/reference) and the pyradiomics call behind it.
 *
 * Conventions: plain pointers and sizes only; `img`, `mask`, `out`, `status` and every debug
 * buffer are DEVICE pointers owned by the caller; calls are asynchronous and ordered on the
 * given CUDA stream (a `cudaStream_t` passed as void*); the only allocation after radb_create
 * is the grow-only record workspace (see radb_reserve).  Every function returns 0 on success and a negative code on misuse / CUDA
 * errors (text via radb_last_error); nothing aborts.  A handle is not thread-safe: one handle
 * per (thread, device).
 */
#ifndef RADB_H
#define RADB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct radb_handle radb_handle;

/* feature classes; output order = shape2D (if enabled), then the order below (params.yml:164-171) */
enum {
    RADB_CLASS_FIRSTORDER = 1u << 0,
    RADB_CLASS_GLCM = 1u << 1,
    RADB_CLASS_GLDM = 1u << 2,
    RADB_CLASS_GLRLM = 1u << 3,
    RADB_CLASS_GLSZM = 1u << 4,
    RADB_CLASS_NGTDM = 1u << 5,
    RADB_CLASS_SHAPE2D = 1u << 6, /* mask-only; its 9 columns come first (pyradiomics computes shape first) */
    RADB_CLASS_ALL = 0x7fu
};

/* pixel types of `img` */
enum { RADB_DTYPE_U8 = 0, RADB_DTYPE_U16 = 1, RADB_DTYPE_F32 = 2, RADB_DTYPE_F64 = 3 };

/* per-patch status written to `status[b]`; rows with status != 0 are NaN.
 * 1-3 are the ValueErrors pyradiomics' imageoperations.checkMask raises through
 * RadiomicExtractor.py:38 (no try/except there: the reference aborts the run). */
enum {
    RADB_ST_OK = 0,
    RADB_ST_LABEL_ABSENT = 1,   /* "Label (255) not present in mask" */
    RADB_ST_SINGLE_VOXEL = 2,   /* "mask only contains 1 segmented voxel" */
    RADB_ST_TOO_FEW_DIMS = 3,   /* ROI spans < minimumROIDimensions (2) axes */
    RADB_ST_NG_OVERFLOW = 4     /* more gray levels than the handle was sized for */
};

/* error codes (function return values) */
enum {
    RADB_OK = 0,
    RADB_E_INVALID = -1,      /* bad argument / unsupported setting */
    RADB_E_CUDA = -2,         /* CUDA runtime error */
    RADB_E_UNSUPPORTED = -3,  /* valid pyradiomics setting this build does not implement yet */
    RADB_E_SMEM = -4          /* patch size x gray levels exceed 227 KB of shared memory */
};

/* Resolved extraction settings.  Stands in for the `setting:` section of the pyradiomics
 * parameter file the reference loads at RadiomicExtractor.py:15 (params.yml:62-119) and that
 * RadiomicsFeatureExtractor.execute forwards to every feature class. */
typedef struct radb_settings {
    double bin_width;          /* params.yml:96  binWidth (10); pyradiomics default 25 */
    int32_t bin_count;         /* binCount (params.yml:97): > 0 splits the ROI range into this many bins (numpy.histogram
                                  edges, last edge + 1) and takes precedence over bin_width; 0 = fixed-width binning */
    int32_t label;             /* params.yml:93  label (255) */
    int32_t n_angles;          /* unidirectional offsets resolved on the host from force2D /
                                  force2Ddimension (params.yml:100): 1 (row-only, the literal
                                  2-D + force2D case) or 4 (in-plane) */
    int8_t angles[8][2];       /* (dy, dx) per unidirectional angle; the bidirectional set used
                                  by GLSZM/GLDM/NGTDM is these plus their negations */
    int32_t symmetrical_glcm;  /* params.yml:119 symmetricalGLCM (True) */
    double gldm_alpha;         /* gldm_a (0) */
    double voxel_array_shift;  /* voxelArrayShift (0) */
    uint32_t class_mask;       /* RADB_CLASS_* bits: params.yml:164-171 featureClass */
    int32_t max_ng;            /* sizing bound on gray levels (<= 256; <= 255 for uint8 pixels); 0 = derive from bin_width for
                                  uint8 pixels; required for the other pixel types */
    int32_t device;            /* CUDA device ordinal */
} radb_settings;

/* RadiomicsExtractor.__init__ (RadiomicExtractor.py:14-15): build an extractor from settings. */
int radb_create(const radb_settings* s, radb_handle** out);
void radb_destroy(radb_handle* h);

/* get_enabled_features (RadiomicExtractor.py:20-21) at feature granularity: number and names
 * ("original_glcm_Contrast", ...) of the columns of one output row, in output order. */
int radb_feature_count(const radb_handle* h);
const char* radb_feature_name(const radb_handle* h, int i);

/* Pre-allocates the per-chunk record workspace used by launches on `cuda_stream` for batches of
 * up to `max_batch` HxW patches, so that later radb_extract calls on that stream allocate
 * nothing.  Without it the workspace grows on demand (cudaMalloc, grow-only) the first time a
 * larger batch or record size is seen.  The handle keeps one workspace per stream, so launches
 * issued on different streams may overlap. */
int radb_reserve(radb_handle* h, int H, int W, int dtype, int64_t max_batch, void* cuda_stream);

/* Patches per chunk (rounded up to a multiple of 4; 0 restores the default: RADB_CHUNK in the environment, else
 * 65536).  A batch larger than one chunk is cut into equal chunks and pipelined over two streams: the build
 * kernel of chunk i+1 overlaps the reduction kernels of chunk i.  The workspace holds two chunks of records. */
int radb_set_chunk(radb_handle* h, int64_t patches);

/* Rows per chunk a dense call with B patches of HxW `dtype` will use (the chunks are equal but for the last one). */
int64_t radb_chunk_rows(const radb_handle* h, int H, int W, int dtype, int64_t B);

/* Completion events (cudaEvent_t[n]) for the chunks of the NEXT dense extraction call on this handle (one-shot):
 * cuda_events[c] is recorded once the output rows of chunk c -- rows [c * radb_chunk_rows, ...) -- are final, so that
 * a consumer on another stream (the multi-GPU driver's all-gather of a slice of rows) can start while later chunks are
 * still being extracted.  The reference has no such hook: its fan-out returns whole lists
 * (/root/reference/RadiomicExtractor.py:63-65). */
int radb_set_chunk_events(radb_handle* h, void* const* cuda_events, int n);

/* Dynamic shared memory (bytes) one CTA of the build kernel needs for HxW patches of `dtype`; < 0 if it cannot fit. */
int radb_smem_bytes(const radb_handle* h, int H, int W, int dtype);

/* extractor.execute(image, mask, label=label) (RadiomicExtractor.py:38,42,45,48), batched:
 * B (image, mask) pairs of H x W pixels -> out[B][F] float64 rows + status[B].
 * img/mask strides are BYTES between consecutive patches.  Asynchronous on `cuda_stream`. */
int radb_extract(radb_handle* h, const void* img, int dtype, const uint8_t* mask, int64_t B, int H, int W,
                 int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status, void* cuda_stream);

/* Variable-size batches -- the reference's records are whole images of differing sizes, each with its own
 * mask (RadiomicExtractor.py:29-36), fanned out one execute() per record (:58-65).  `img_pool` / `mask_pool`
 * are DEVICE buffers holding n patches back to back in any order; the HOST arrays img_off[n] / mask_off[n]
 * give each patch's byte offset into its pool and hw[n][2] its (H, W).  Patches of equal size are launched
 * together; row i of out / status belongs to patch i whatever its size.  Offsets that are multiples of 16
 * bytes keep the TMA staging path.  The index lists are copied to a grow-only device buffer of the handle. */
int radb_extract_ragged(radb_handle* h, const void* img_pool, int dtype, const uint8_t* mask_pool, int64_t n,
                        const int64_t* img_off, const int64_t* mask_off, const int32_t* hw, double* out,
                        int32_t* status, void* cuda_stream);

/* RadiomicsExtractor.extract_radiomics (RadiomicExtractor.py:23-55) for a batch of decoded records:
 * `bgr` = interleaved BGR uint8 images [n][H][W][3] exactly as cv2.imread returns them (:29), `mask`
 * = one uint8 mask per image [n][H][W] (:33-36).  A front-end kernel writes the gray (cv2 BGR2GRAY
 * fixed point, :30), R, G, B planes (:41-47) into the caller's device scratch `planes`
 * ([n][4][H][W] bytes) and the four executes of every image share its mask.
 * out: [n*4][F], status: [n*4], rows in gray, R, G, B order per image. */
int radb_extract_bgr(radb_handle* h, const uint8_t* bgr, const uint8_t* mask, int64_t n_images, int H, int W,
                     uint8_t* planes, double* out, int32_t* status, void* cuda_stream);

/* RadiomicExtractor.py:34-35: when the mask's size differs from the image's, the reference nearest-resizes it with
 * cv2.resize(mask, (W, H), interpolation=cv2.INTER_NEAREST).  Device twin, bit-exact with cv2 4.x: n masks
 * src [n][sH][sW] -> dst [n][dH][dW] (DEVICE pointers, stream-ordered). */
int radb_resize_mask(radb_handle* h, const uint8_t* src, int64_t n, int sH, int sW, uint8_t* dst, int dH, int dW,
                     void* cuda_stream);

/* Packed-mask transfer path.  The reference hands over uint8 masks (RadiomicExtractor.py:33-36) of which only
 * `mask == label` matters (params.yml:93): one bit per pixel.  End to end the extraction is bound by the
 * host-to-device link, so the host may pack a mask buffer to 1 bit per pixel (bit i of the byte stream <->
 * byte i of `mask`; HOST pointers; AVX2 over `threads` host threads, memory bound), copy n_bytes / 8 bytes
 * instead of n_bytes, and expand it on the device with radb_unpack_mask (DEVICE pointers, stream-ordered;
 * `mask` gets `label` where the bit is set and a different value elsewhere) before radb_extract. */
int radb_pack_mask_host(const uint8_t* mask, int64_t n_bytes, int label, uint8_t* packed, int threads);

/* radb_extract with bit-packed masks consumed directly by the kernels (no expansion pass, 1/8 of the mask bytes in
 * HBM and over the link): per patch ceil(H*W/8) bytes, bit i (LSB first) <=> pixel i is in the ROI, i.e. what
 * `mask == label` gives at RadiomicExtractor.py:38; `mask_stride_b` = bytes between the bit streams of consecutive
 * patches (>= ceil(H*W/8); 16-byte multiples keep the TMA staging path).  radb_pack_mask_host over a contiguous
 * [B][H][W] mask buffer produces exactly this layout when H*W is a multiple of 8 (stride H*W/8). */
int radb_pack_masks_host(const uint8_t* mask, int64_t n_patches, int64_t hw, int label, uint8_t* packed,
                         int64_t stride_b, int threads); /* per-patch streams at `stride_b` bytes: any H*W (HOST pointers) */
int radb_extract_packed(radb_handle* h, const void* img, int dtype, const uint8_t* mask_bits, int64_t B, int H, int W,
                        int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status, void* cuda_stream);
int radb_unpack_mask(radb_handle* h, const uint8_t* packed, int64_t n_bytes, uint8_t* mask, void* cuda_stream);

/* imageType filters of the parameter file (params.yml:141-144; pyradiomics imageoperations.getSquareImage,
 * getSquareRootImage, getLogarithmImage, getExponentialImage) for uint8 images, computed in float64:
 * type 1 Square, 2 SquareRoot, 3 Logarithm, 4 Exponential.  img [n][H*W] uint8 -> out [n][H*W] float64
 * (feed it to radb_extract with RADB_DTYPE_F64); mx = int32 [n] device scratch. */
int radb_derive_image(radb_handle* h, const uint8_t* img, int64_t n_images, int64_t HW, int type, double* out,
                      int32_t* mx, void* cuda_stream);

/* Filtered image types of the parameter file (params.yml:138-140,145; pyradiomics imageoperations.getGradientImage ->
 * sitk.GradientMagnitude, getLoGImage -> sitk.LaplacianRecursiveGaussian with NormalizeAcrossScale, getWaveletImage
 * -> level-1 stationary coif1 transform) for uint8 images [n][H][W]; DEVICE pointers, stream-ordered:
 *   type 5 Gradient          -> out float32 [n][H][W]                                  (scratch may be null)
 *   type 6 LoG, param=sigma  -> out float32 [n][H][W]; scratch >= n*H*W*12 bytes       (H, W >= 4)
 *   type 7 Wavelet: flags bit 0 set = transform along x only (pyradiomics' axis removal for force2D on a 2-D
 *                   array) -> out float64 [n][2][H][W] = wavelet-H, wavelet-L; flags 0 -> out float64 [n][4][H][W]
 *                   = wavelet-LH, -HL, -HH, -LL (first letter <-> x); scratch >= n*2*(H+1)*(W+1)*8 bytes
 * Feed the result to radb_extract with RADB_DTYPE_F32 / RADB_DTYPE_F64. */
int radb_filter_image(radb_handle* h, const uint8_t* img, int64_t n_images, int H, int W, int type, double param,
                      int flags, void* out, void* scratch, void* cuda_stream);

/* Same launch as radb_extract, additionally dumping the discretised image and the integer
 * texture matrices the features were reduced from (what pyradiomics' cMatrices.calculate_*
 * return) for bit-exact parity tests.  Any debug pointer may be NULL.  All buffers must be
 * zero-filled by the caller.  Shapes (Ng = radb_max_ng(h), Na = n_angles, Nr = max(H, W)):
 *   levels int32 [B][H][W]      glcm  int32 [B][Na][Ng][Ng]   glrlm int32 [B][Na][Ng][Nr]
 *   glszm  int32 [B][Ng][H*W]   gldm  int32 [B][Ng][2*Na+1]   ngtdm_n int32 [B][Ng]
 *   ngtdm_s float64 [B][Ng]     ng    int32 [B] (max gray level of each patch) */
int radb_debug_matrices(radb_handle* h, const void* img, int dtype, const uint8_t* mask, int64_t B, int H,
                        int W, int64_t img_stride_b, int64_t mask_stride_b, double* out, int32_t* status,
                        int32_t* levels, int32_t* glcm, int32_t* glrlm, int32_t* glszm, int32_t* gldm,
                        int32_t* ngtdm_n, double* ngtdm_s, int32_t* ng, void* cuda_stream);

int radb_max_ng(const radb_handle* h);

/* Number of kernel launches this handle has issued (bench.py's gpu_launches evidence): four
 * kernels (build, angle-level reductions, misc thread-level, misc warp-level residual) per chunk of
 * 65536 patches (+1 with shape2D, +1 for the BGR front-end). */
int64_t radb_launch_count(const radb_handle* h);

/* Per-kernel device timing for bench.py's roofline: with profiling on, every launch records CUDA
 * events on its stream around the build / angle / misc kernels; radb_kernel_ms waits for them,
 * returns the accumulated milliseconds per kernel in ms3[0..2] and clears the record. */
int radb_set_profiling(radb_handle* h, int on);
int radb_kernel_ms(radb_handle* h, double* ms3);

const char* radb_last_error(void);
const char* radb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RADB_H */
