/*
 * cvtt_b200.h -- C ABI of libcvtt_b200.so, the B200 (sm_100a) implementation of the per-4x4-block encode
 * hot path of elasota/ConvectionKernels.
 *
 * Every entry point replaces one piece of the reference's public interface
 * (reference ConvectionKernels.h:236-277, implemented in ConvectionKernels_API.cpp):
 *
 *   cvttb200_encode(CVTTB200_BC7, ...)        <- cvtt::Kernels::EncodeBC7           ConvectionKernels.h:252, API.cpp:41-54
 *   cvttb200_bc7_plan_default                 <- cvtt::BC7EncodingPlan()            ConvectionKernels.h:166-198
 *   cvttb200_bc7_plan_from_quality            <- ConfigureBC7EncodingPlanFromQuality        ConvectionKernels.h:262, BC67.cpp:3291
 *   cvttb200_bc7_plan_from_fine_tuning        <- ConfigureBC7EncodingPlanFromFineTuningParams ConvectionKernels.h:265, BC67.cpp:3355
 *   cvttb200_options_default                  <- cvtt::Options()                    ConvectionKernels.h:89-101
 *
 * The reference encodes exactly cvtt::NumParallelBlocks = 8 blocks per call (ConvectionKernels.h:71) and a
 * few decisions are taken jointly for those 8 blocks, so a block's bytes depend on which 8-block group it is
 * in.  cvttb200_encode takes any multiple of 8 blocks and treats blocks [8k, 8k+8) as the k-th reference
 * call; the result equals calling the reference nBlocks/8 times.
 *
 * Structs are plain-old-data with the exact layout of their cvtt:: counterparts (sizes are checked at
 * library load), so a C++ caller may pass &cvtt::Options / &cvtt::BC7EncodingPlan directly.
 *
 * There is no CPU fallback: every encode call fails with CVTTB200_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef CVTT_B200_H
#define CVTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Output formats.  Input block size / output block size in bytes are given for each. */
typedef enum cvttb200_format
{
    CVTTB200_BC1 = 1,            /* PixelBlockU8 64 B -> 8 B   EncodeBC1   ConvectionKernels.h:243 */
    CVTTB200_BC2 = 2,            /* 64 -> 16                   EncodeBC2   :244 */
    CVTTB200_BC3 = 3,            /* 64 -> 16                   EncodeBC3   :245 */
    CVTTB200_BC4U = 4,           /* 64 -> 8                    EncodeBC4U  :246 */
    CVTTB200_BC4S = 5,           /* PixelBlockS8 64 -> 8       EncodeBC4S  :247 */
    CVTTB200_BC5U = 6,           /* 64 -> 16                   EncodeBC5U  :248 */
    CVTTB200_BC5S = 7,           /* 64 -> 16                   EncodeBC5S  :249 */
    CVTTB200_BC6HU = 8,          /* PixelBlockF16 128 -> 16    EncodeBC6HU :250 */
    CVTTB200_BC6HS = 9,          /* 128 -> 16                  EncodeBC6HS :251 */
    CVTTB200_BC7 = 10,           /* 64 -> 16                   EncodeBC7   :252 */
    CVTTB200_ETC1 = 11,          /* 64 -> 8                    EncodeETC1  :253 */
    CVTTB200_ETC2 = 12,          /* 64 -> 8                    EncodeETC2  :254 */
    CVTTB200_ETC2_RGBA = 13,     /* 64 -> 16                   EncodeETC2RGBA :255 */
    CVTTB200_ETC2_PUNCHTHROUGH = 14, /* 64 -> 8                EncodeETC2PunchthroughAlpha :256 */
    CVTTB200_ETC2_ALPHA = 15,    /* 64 -> 8                    EncodeETC2Alpha :258 */
    CVTTB200_EAC_R11U = 16,      /* PixelBlockScalarS16 32 -> 8  EncodeETC2Alpha11(isSigned=false) :259 */
    CVTTB200_EAC_R11S = 17       /* 32 -> 8                    EncodeETC2Alpha11(isSigned=true)  :259 */
} cvttb200_format;

typedef enum cvttb200_status
{
    CVTTB200_OK = 0,
    CVTTB200_ERR_BAD_ARGUMENT = -1,   /* null pointer, nBlocks not a multiple of 8, missing plan, ... */
    CVTTB200_ERR_UNSUPPORTED = -2,    /* format or flag not implemented by this build (never silently ignored) */
    CVTTB200_ERR_NO_DEVICE = -3,      /* no usable CUDA device / wrong architecture */
    CVTTB200_ERR_CUDA = -4            /* a CUDA call failed; see cvttb200_last_error() */
} cvttb200_status;

/* cvtt::Options, ConvectionKernels.h:73-103 (44 bytes) */
typedef struct cvttb200_options
{
    uint32_t flags;
    float threshold;
    float redWeight, greenWeight, blueWeight, alphaWeight;
    int refineRoundsBC7, refineRoundsBC6H, refineRoundsIIC, refineRoundsS3TC;
    int seedPoints;
} cvttb200_options;

/* cvtt::Flags, ConvectionKernels.h:33-69 */
#define CVTTB200_FLAG_BC7_FAST_INDEXING         0x008u
#define CVTTB200_FLAG_BC7_TRY_SINGLE_COLOR      0x010u
#define CVTTB200_FLAG_BC7_RESPECT_PUNCH_THROUGH 0x020u
#define CVTTB200_FLAG_BC6H_FAST_INDEXING        0x040u
#define CVTTB200_FLAG_S3TC_EXHAUSTIVE           0x080u
#define CVTTB200_FLAG_S3TC_PARANOID             0x100u
#define CVTTB200_FLAG_UNIFORM                   0x200u
#define CVTTB200_FLAG_ETC_USE_FAKE_BT709        0x400u
#define CVTTB200_FLAG_ETC_FAKE_BT709_ACCURATE   0x800u

/* cvtt::BC7EncodingPlan, ConvectionKernels.h:142-199 (808 bytes) */
typedef struct cvttb200_bc7_plan
{
    uint64_t mode1PartitionEnabled;
    uint64_t mode2PartitionEnabled;
    uint64_t mode3PartitionEnabled;
    uint16_t mode0PartitionEnabled;
    uint64_t mode7RGBAPartitionEnabled;
    uint64_t mode7RGBPartitionEnabled;
    uint8_t mode4SP[4][2];
    uint8_t mode5SP[4];
    uint8_t mode6Enabled;
    uint8_t seedPointsForShapeRGB[243];
    uint8_t seedPointsForShapeRGBA[129];
    uint8_t rgbaShapeList[129];
    uint8_t rgbaNumShapesToEvaluate;
    uint8_t rgbShapeList[243];
    uint8_t rgbNumShapesToEvaluate;
} cvttb200_bc7_plan;

/* cvtt::BC7FineTuningParams, ConvectionKernels.h:105-140 (285 bytes) */
typedef struct cvttb200_bc7_fine_tuning
{
    uint8_t mode0SP[16];
    uint8_t mode1SP[64];
    uint8_t mode2SP[64];
    uint8_t mode3SP[64];
    uint8_t mode4SP[4][2];
    uint8_t mode5SP[4];
    uint8_t mode6SP;
    uint8_t mode7SP[64];
} cvttb200_bc7_fine_tuning;

/* ---- lifetime ------------------------------------------------------------------------------------------ */

/* Prepares `device` (uploads the constant tables, derives the reciprocal table from this host's
 * _mm_rcp_ps: the reference's EndpointRefiner uses that instruction (ConvectionKernels_ParallelMath.h:569-575,
 * ConvectionKernels_EndpointRefiner.h:106) and its result differs between CPU vendors, so bit-exactness is
 * defined against the reference running on the same host).  Idempotent; encode calls initialise the current
 * device on first use. */
int cvttb200_init(int device);
void cvttb200_shutdown(void);

/* Overrides the 17-entry table rcp[n] ~ 1/n (n = 0..16) used in place of _mm_rcp_ps((float)n); NULL restores
 * the host's own.  Used to replay golden vectors recorded on a different CPU model. */
int cvttb200_set_rcp_table(const float *rcp17);
int cvttb200_get_rcp_table(float *rcp17);

/* Thread-local description of the last failure. */
const char *cvttb200_last_error(void);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t cvttb200_launch_count(void);

/* Device self-test of the arithmetic building blocks the kernels rely on for bit-exactness: the two-lane fp32
 * division used by the BC7 search is compared with the compiler's IEEE division on `samples` pseudo-random operand
 * pairs (log-uniform magnitudes in 2^-40 .. 2^40, both signs, zero numerators).  *mismatches receives the number of
 * quotients whose bits differ (a differing sign of a zero quotient is not counted).  Returns a cvttb200_status. */
int cvttb200_selftest(uint64_t samples, uint64_t seed, uint64_t *mismatches);

/* ---- configuration (pure host code) ---------------------------------------------------------------------- */

void cvttb200_options_default(cvttb200_options *options);
void cvttb200_bc7_plan_default(cvttb200_bc7_plan *plan);
void cvttb200_bc7_plan_from_quality(cvttb200_bc7_plan *plan, int quality);
int cvttb200_bc7_plan_from_fine_tuning(cvttb200_bc7_plan *plan, const cvttb200_bc7_fine_tuning *params);   /* returns 1 like the reference's `true` */
void cvttb200_bc7_fine_tuning_default(cvttb200_bc7_fine_tuning *params);

size_t cvttb200_input_block_bytes(int format);
size_t cvttb200_output_block_bytes(int format);

/* ---- image <-> block array (device memory; the loops of the reference's sample packer, etc2packer/etc2packer.cpp:215-248
 *      and :277-284, which callers otherwise run on the CPU) ------------------------------------------------------------- */

/* Number of blocks cvttb200_tile_image produces: ceil(height / 4) rows of ceil(width / 32) groups of 8 blocks. */
size_t cvttb200_tiled_block_count(int width, int height);

/* Cuts a linear image (pixelBytes = 4: RGBA8 -> PixelBlockU8, 8: RGBA16F -> PixelBlockF16; rows rowPitchBytes apart) into the
 * block array the encoders take: row-major 4x4 blocks, 8 horizontally consecutive blocks per reference call, coordinates past
 * the right / bottom edge clamped to the last column / row (so the padding blocks of a row's last group repeat the edge, exactly
 * like the sample packer).  `image` and `blocks` are device pointers; the work is enqueued on `stream`. */
int cvttb200_tile_image(int pixelBytes, const void *image, int width, int height, size_t rowPitchBytes, void *blocks, void *stream);

/* Inverse bookkeeping for encoded data: drops the padding blocks of every row (rows of ceil(width / 32) * 8 encoded blocks ->
 * rows of ceil(width / 4) blocks, the order a KTX / DDS payload uses).  blockBytes is 8 or 16.  Device pointers. */
int cvttb200_untile_blocks(const void *encoded, int width, int height, size_t blockBytes, void *out, void *stream);

/* KTX 1.1 container header for the payload cvttb200_untile_blocks produces, as the reference's sample packer writes it
 * (etc2packer/ktxheader.h; etc2packer/etc2packer.cpp:114-141 header fields, :143-180 GL enums per target, :190-193 the imageSize
 * word): fills 68 bytes of HOST memory, the 64-byte header followed by the 32-bit byte count of the one mip level; the file is
 * those 68 bytes followed by ceil(w/4) * ceil(h/4) encoded blocks in row-major order.  Formats: the sample's six targets --
 * ETC1, ETC2, ETC2_RGBA, ETC2_PUNCHTHROUGH, EAC_R11U, EAC_R11S; anything else answers CVTTB200_ERR_UNSUPPORTED. */
#define CVTTB200_KTX_HEADER_BYTES 68
int cvttb200_ktx_header(int format, int width, int height, void *header68);

/* ---- the hot path ---------------------------------------------------------------------------------------- */

/* Encodes nBlocks (a multiple of 8) blocks.  `blocks` and `out` may each be host memory (pageable or pinned) or
 * device memory of the current CUDA device; the kind is detected per pointer.  With device pointers the work is
 * enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call returns without synchronising; if
 * either pointer is host memory the call copies through device staging buffers and returns when `out` is
 * complete (both pointers host memory and stream == NULL: the work runs on a stream of the library's own instead of
 * queueing behind the default stream, so that such calls from several host threads overlap on the device).  `plan` is required for CVTTB200_BC7 and ignored otherwise.
 * Thread safety: like the reference (README.md:57) every entry point may be called from any number of host threads at once;
 * calls on different streams overlap on the device (staging buffers and kernel scratch are per call).
 * Returns a cvttb200_status. */
int cvttb200_encode(int format, const void *blocks, size_t nBlocks, void *out,
                    const cvttb200_options *options, const cvttb200_bc7_plan *plan, void *stream);

/* cvttb200_encode with the second Options object the reference's ETC2 entry points see: cvtt::Kernels::AllocETC2Data(alloc,
 * context, options) (ConvectionKernels.h:268, ConvectionKernels_ETC.cpp:3117-3145) derives the chroma side axes of the T/H-mode
 * search from THOSE options once, and EncodeETC2 / EncodeETC2RGBA / EncodeETC2PunchthroughAlpha read them from the
 * ETC2CompressionData (ETC.cpp:1773) while taking flags and error weights from their own `options` argument.
 * etc2AllocOptions = the Options given to AllocETC2Data; NULL means "the same as options" (what cvttb200_encode does).  Ignored
 * by every other format. */
int cvttb200_encode_ex(int format, const void *blocks, size_t nBlocks, void *out,
                       const cvttb200_options *options, const cvttb200_bc7_plan *plan,
                       const cvttb200_options *etc2AllocOptions, void *stream);

/* The same call over several GPUs of this process (the reference itself is single threaded and leaves
 * parallelism to its caller, README.md:57).  `blocks` and `out` are HOST buffers.  The blocks are split into nDevices contiguous
 * ranges of whole 8-block groups -- a group is one reference call (cvtt::NumParallelBlocks, ConvectionKernels.h:71) and is never
 * split, because the reference decides some things jointly for its 8 blocks -- and every device copies its range in, encodes it
 * and copies the result back on a stream of its own; the call returns when `out` is complete.  The result is byte-identical to
 * cvttb200_encode on one device.  devices == NULL means devices 0 .. nDevices - 1; nDevices <= 0 means every visible device.
 * The current device of the calling thread is preserved.  Returns a cvttb200_status. */
int cvttb200_encode_multi(int format, const void *blocks, size_t nBlocks, void *out,
                          const cvttb200_options *options, const cvttb200_bc7_plan *plan, const int *devices, int nDevices);

/* ---- decoders ---------------------------------------------------------------------------------------------
 * cvtt::Kernels::DecodeBC7 / DecodeBC6HU / DecodeBC6HS (ConvectionKernels.h:273-275, ConvectionKernels_API.cpp:288-322 ->
 * BC7Computer::UnpackOne BC67.cpp:2206-2423, BC6HComputer::UnpackOne :3059-3289).  format: CVTTB200_BC7 (-> PixelBlockU8, 64 B
 * per block) or CVTTB200_BC6HU / CVTTB200_BC6HS (-> PixelBlockF16, 128 B per block, alpha = 1.0).  Any block count; blocks are
 * independent.  Pointer kinds, stream and return value as for cvttb200_encode.  Invalid blocks decode like the reference's:
 * BC7 without a mode bit -> all zero, reserved BC6H modes -> (0, 0, 0, 1.0). */
int cvttb200_decode(int format, const void *encoded, size_t nBlocks, void *pixelBlocks, void *stream);

#ifdef __cplusplus
}
#endif

#endif
