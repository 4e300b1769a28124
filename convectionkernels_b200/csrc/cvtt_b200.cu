// libcvtt_b200.so: CUDA kernels (sm_100a) + the C ABI declared in include/cvtt_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false ... (convectionkernels_b200/build.py).
// -fmad=false is part of the numerical contract: the reference's fp32 expressions must round after every
// operation.
#include <cuda_runtime.h>
#include <xmmintrin.h>

#include <algorithm>
#include <deque>
#include <mutex>
#include <string>
#include <vector>
#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cvtt_b200.h"
#include "bc7_host.h"
#include "bc6h_host.h"
#include "etc_host.h"
#include "s3tc_host.h"
#include "decode_core.cuh"
#include "cvtt_internal.h"

using namespace cvttb200;

static_assert(sizeof(cvttb200_options) == 44 && sizeof(OptionsPOD) == 44, "cvtt::Options layout");
static_assert(sizeof(cvttb200_bc7_plan) == 808 && sizeof(BC7PlanPOD) == 808, "cvtt::BC7EncodingPlan layout");
static_assert(sizeof(cvttb200_bc7_fine_tuning) == 285 && sizeof(BC7FineTuningPOD) == 285, "cvtt::BC7FineTuningParams layout");

// =========================================================================================================
// Kernels

// ---------------------------------------------------------------------------------------------------------
// Image <-> block array (the step before / after the encode path; the reference's sample does it on the CPU,
// etc2packer/etc2packer.cpp:215-248 and :277-284).  Pure data movement: this is the HBM-bound part of the pipeline.

namespace
{
    // One thread per block row (4 pixels), kTileUnroll rows in flight per thread.  PIXEL_BYTES = 4 (RGBA8 -> PixelBlockU8) or 8
    // (RGBA16F -> PixelBlockF16).  Blocks are laid out like the sample packer does: rows of ceil(width / 32) groups of 8 blocks;
    // coordinates past the edge are clamped, so the padding blocks of the last group of a row repeat the last column (they take
    // part in the group semantics of the encoders).  Thread t of a warp handles row (t & 3) of block (t >> 2): the warp's
    // stores are one contiguous 512 B (1 KB) run of the block array, its loads are four 128 B (256 B) row segments -- whole
    // 32-byte sectors on both sides.  A thread issues the loads of all its rows (one CTA-width apart, so every access keeps
    // that shape) before the first store: with a single 16-byte load in flight per thread the RGBA8 case stayed at 88 % of
    // the copy bandwidth (long_scoreboard, profiles/r02_misc_kernels.csv).
    constexpr int kTileUnroll = 4;

    template<int PIXEL_BYTES>
    __global__ void __launch_bounds__(256)
    tile_image_kernel(const unsigned char *__restrict__ image, int width, int height, size_t pitch, int blocksPerRow, uint32_t nBlocks, uint4 *__restrict__ blocks, int vectorOK)
    {
        constexpr int kRowVec = PIXEL_BYTES / 4;      // uint4 per block row
        const uint64_t first = (uint64_t)blockIdx.x * (blockDim.x * kTileUnroll) + threadIdx.x;
        uint4 v[kTileUnroll][kRowVec];
        bool fast[kTileUnroll];
#pragma unroll
        for (int u = 0; u < kTileUnroll; u++)
        {
            const uint64_t i = first + (uint64_t)u * blockDim.x;
            const uint32_t b = (uint32_t)(i >> 2);
            const int r = (int)(i & 3);
            fast[u] = false;
            if (b >= nBlocks)
                continue;
            const int by = (int)(b / (uint32_t)blocksPerRow), bx = (int)(b % (uint32_t)blocksPerRow);
            const int x0 = bx * 4, y = min(by * 4 + r, height - 1);
            const unsigned char *rowPtr = image + (size_t)y * pitch;
            if (vectorOK && x0 + 4 <= width)
            {
                fast[u] = true;
                const uint4 *src = reinterpret_cast<const uint4 *>(rowPtr + (size_t)x0 * PIXEL_BYTES);
#pragma unroll
                for (int k = 0; k < kRowVec; k++)
                    v[u][k] = __ldg(src + k);
            }
        }
#pragma unroll
        for (int u = 0; u < kTileUnroll; u++)
        {
            const uint64_t i = first + (uint64_t)u * blockDim.x;
            const uint32_t b = (uint32_t)(i >> 2);
            const int r = (int)(i & 3);
            if (b >= nBlocks)
                continue;
            uint4 *dst = blocks + ((size_t)b * 4 + r) * kRowVec;
            if (fast[u])
            {
#pragma unroll
                for (int k = 0; k < kRowVec; k++)
                    dst[k] = v[u][k];
                continue;
            }
            // edge blocks (clamped columns) and unaligned images: pixel by pixel
            const int by = (int)(b / (uint32_t)blocksPerRow), bx = (int)(b % (uint32_t)blocksPerRow);
            const int x0 = bx * 4, y = min(by * 4 + r, height - 1);
            const unsigned char *rowPtr = image + (size_t)y * pitch;
            uint32_t row[4 * (PIXEL_BYTES / 4)];
            for (int c = 0; c < 4; c++)
            {
                const int x = min(x0 + c, width - 1);
                const uint32_t *px = reinterpret_cast<const uint32_t *>(rowPtr + (size_t)x * PIXEL_BYTES);
                for (int k = 0; k < PIXEL_BYTES / 4; k++)
                    row[c * (PIXEL_BYTES / 4) + k] = px[k];
            }
            for (int k = 0; k < kRowVec; k++)
                dst[k] = make_uint4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
        }
    }

    // Drops the padding blocks: encoded rows of blocksPerRow blocks -> rows of ceil(width / 4) blocks (the payload order of the
    // sample's KTX writer).  One thread per output block (VEC = uint2 for 8-byte blocks, uint4 for 16-byte blocks).
    template<class VEC>
    __global__ void __launch_bounds__(256)
    untile_blocks_kernel(const VEC *__restrict__ encoded, int blocksPerRow, int realBlocksPerRow, uint64_t nOut, VEC *__restrict__ out)
    {
        const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= nOut)
            return;
        const uint64_t row = i / (uint64_t)realBlocksPerRow, col = i % (uint64_t)realBlocksPerRow;
        out[i] = __ldg(encoded + row * (uint64_t)blocksPerRow + col);
    }
}

// ---------------------------------------------------------------------------------------------------------
// BC7 / BC6H decoders (DecodeBC7 / DecodeBC6HU / DecodeBC6HS): 16 bytes in, 64 or 128 bytes out per block

namespace
{
    constexpr int kDecodeThreads = 256;

    struct DecodeTables
    {
        BC7PackTables bc7;
        BC6HTables bc6h;
    };
    __device__ DecodeTables g_decodeTables;         // global-memory copy: the decoders index the tables per lane

    // The warp's output tile in shared memory as 16-byte chunks, CHUNKS per block.  A lane writes its own block's chunks, the
    // warp then streams the tile out with consecutive lanes on consecutive chunks; the XOR keeps both phases conflict-free.
    template<int CHUNKS>
    struct TileSink
    {
        uint4 *tile;
        uint32_t lane;
        __device__ __forceinline__ static uint32_t slot(uint32_t t, uint32_t q)
        {
            return t * CHUNKS + (q ^ (CHUNKS == 4 ? ((t >> 1) & 3u) : (t & 7u)));
        }
        __device__ __forceinline__ void put4(int q, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const
        {
            tile[slot(lane, (uint32_t)q)] = make_uint4(a, b, c, d);
        }
    };

    // KIND: 0 BC7, 1 BC6H unsigned, 2 BC6H signed.  One thread per block, warps walk 32-block slices (persistent grid).
    template<int KIND>
    __global__ void __launch_bounds__(kDecodeThreads)
    decode_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nBlocks)
    {
        constexpr int CHUNKS = (KIND == 0) ? 4 : 8;
        __shared__ __align__(16) DecodeTables sTables;
        __shared__ __align__(16) uint4 sTile[kDecodeThreads * CHUNKS];

        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&g_decodeTables);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&sTables);
            for (uint32_t i = threadIdx.x; i < sizeof(DecodeTables) / 4; i += kDecodeThreads)
                dst[i] = __ldg(src + i);
        }
        __syncthreads();

        const uint32_t lane = threadIdx.x & 31, warpInCta = threadIdx.x >> 5;
        TileSink<CHUNKS> sink;
        sink.tile = sTile + warpInCta * 32 * CHUNKS;
        sink.lane = lane;

        const uint32_t warpsPerGrid = gridDim.x * (kDecodeThreads / 32);
        const uint32_t nSlices = (nBlocks + 31) / 32;
        for (uint32_t slice = blockIdx.x * (kDecodeThreads / 32) + warpInCta; slice < nSlices; slice += warpsPerGrid)
        {
            const uint32_t block = slice * 32 + lane;
            if (block < nBlocks)
            {
                const uint4 v = __ldg(in + block);
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                if (KIND == 0)
                    bc7_decode_block(sTables.bc7, w, sink);
                else
                    bc6h_decode_block(sTables.bc6h, w, KIND == 2, sink);
            }
            __syncwarp();
            const uint32_t valid = min(32u, nBlocks - slice * 32) * CHUNKS;
            uint4 *dst = out + (size_t)slice * 32 * CHUNKS;
#pragma unroll
            for (int k = 0; k < CHUNKS; k++)
            {
                const uint32_t c = k * 32 + lane;
                if (c < valid)
                    dst[c] = sink.tile[TileSink<CHUNKS>::slot(c / CHUNKS, c % CHUNKS)];
            }
            __syncwarp();
        }
    }
}

// =========================================================================================================
// Host state
//
// Like the reference (README.md:57) the library may be called from any number of host threads at once.  What is shared:
//   * the registry of per-device contexts and the reciprocal table        -> g_mutex, held for look-ups only
//   * a context's BC7 plan cache                                           -> DeviceContext::planMutex (bc7_kernels.cu)
//   * a context's pair of pipeline streams / its encode_multi stream       -> DeviceContext::pipeMutex / g_multiMutex
// Everything a call needs on the device (staging buffers of host-pointer calls, kernel scratch) is allocated per call,
// stream-ordered, from the context's private memory pool, so that calls on different streams overlap on the device and no
// lock is held across a copy, a launch or a synchronisation.

namespace
{
    thread_local std::string t_lastError;

    std::mutex g_mutex;                               // registry of contexts, reciprocal table
    std::mutex g_multiMutex;                          // cvttb200_encode_multi calls take turns (each uses every listed device)
    std::deque<DeviceContext> g_contexts;             // references stay valid while contexts are added
    float g_rcpN[17];
    bool g_rcpOverridden = false, g_rcpReady = false;

    // restores the calling thread's current device on every exit path
    struct DeviceRestore
    {
        int prev = -1;
        ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
    };
}

namespace cvttb200
{
    std::atomic<uint64_t> g_launches(0);

    int fail(int code, const std::string &msg)
    {
        t_lastError = msg;
        return code;
    }

    int fail_cuda(cudaError_t e, const char *what)
    {
        return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? CVTTB200_ERR_NO_DEVICE : CVTTB200_ERR_CUDA,
                    std::string(what) + ": " + cudaGetErrorString(e));
    }

    int pool_alloc(DeviceContext &ctx, void **p, size_t bytes, cudaStream_t stream)
    {
        *p = nullptr;
        CVTT_CUDA(cudaMallocFromPoolAsync(p, bytes ? bytes : 16, ctx.pool, stream));
        return CVTTB200_OK;
    }
}

namespace
{
    void host_rcp_table(float *t)
    {
        for (int n = 0; n < 17; n++)
            t[n] = _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    }

    // caller holds g_mutex
    int create_context(int device, DeviceContext **out)
    {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            return fail(CVTTB200_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + " (libcvtt_b200 has no CPU fallback)");
        if (device < 0 || device >= count)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "device index out of range");
        cudaDeviceProp prop;
        CVTT_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return fail(CVTTB200_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100; this library contains sm_100a code only");

        DeviceRestore restore;
        CVTT_CUDA(cudaGetDevice(&restore.prev));
        CVTT_CUDA(cudaSetDevice(device));

        // The library's own stream-ordered pool: staging buffers and kernel scratch are cached here between calls (release
        // threshold = never) without changing the behaviour of the device's default pool, which the rest of the process uses.
        cudaMemPool_t pool = nullptr;
        {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            CVTT_CUDA(cudaMemPoolCreate(&pool, &props));
            uint64_t threshold = ~(uint64_t)0;
            CVTT_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        }
        {
            int rc = bc7_device_setup();
            if (rc == CVTTB200_OK) rc = bc6h_device_setup();
            if (rc == CVTTB200_OK) rc = etc_device_setup();
            if (rc != CVTTB200_OK)
            {
                cudaMemPoolDestroy(pool);
                return rc;
            }
        }
        {
            DecodeTables dt;
            dt.bc7 = bc7_pack_tables();
            dt.bc6h = bc6h_tables();
            CVTT_CUDA(cudaMemcpyToSymbol(g_decodeTables, &dt, sizeof(dt)));
        }
        CVTT_CUDA(cudaDeviceSynchronize());

        g_contexts.emplace_back();
        DeviceContext &c = g_contexts.back();
        c.numSMs = prop.multiProcessorCount;
        c.device = device;
        c.pool = pool;
        c.ready = true;
        *out = &c;
        return CVTTB200_OK;
    }

    // the context of `device`, created on first use; rcpN (may be null) receives a snapshot of the reciprocal table
    int get_context(int device, DeviceContext **out, float *rcpN = nullptr)
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (!g_rcpReady)
        {
            host_rcp_table(g_rcpN);
            g_rcpReady = true;
        }
        if (rcpN)
            memcpy(rcpN, g_rcpN, sizeof(g_rcpN));
        for (size_t i = 0; i < g_contexts.size(); i++)
            if (g_contexts[i].device == device && g_contexts[i].ready)
            {
                *out = &g_contexts[i];
                return CVTTB200_OK;
            }
        return create_context(device, out);
    }

    int current_context(DeviceContext **out, float *rcpN = nullptr)
    {
        int device = 0;
        const cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return fail_cuda(e, "cudaGetDevice");
        return get_context(device, out, rcpN);
    }

    bool is_device_pointer(const void *p)
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, p) != cudaSuccess)
        {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    }

    // a stream-ordered allocation from the context's pool that is returned (stream-ordered) when the scope ends
    struct PoolBuffer
    {
        void *p = nullptr;
        cudaStream_t stream = nullptr;
        int alloc(DeviceContext &ctx, size_t bytes, cudaStream_t s)
        {
            stream = s;
            return pool_alloc(ctx, &p, bytes, s);
        }
        ~PoolBuffer() { if (p) cudaFreeAsync(p, stream); }
    };

    // A call whose buffers are both host memory and that names no stream is complete when it returns and has no ordering
    // with any device work of the caller, so it does not have to queue behind the legacy default stream: it borrows one of the
    // context's non-blocking streams.  Unmodified multi-threaded callers of the reference interface then overlap on the device.
    struct StreamLease
    {
        DeviceContext *ctx = nullptr;
        cudaStream_t stream = nullptr;
        int acquire(DeviceContext &c)
        {
            ctx = &c;
            {
                std::lock_guard<std::mutex> lock(c.streamMutex);
                if (!c.idleStreams.empty())
                {
                    stream = c.idleStreams.back();
                    c.idleStreams.pop_back();
                    return CVTTB200_OK;
                }
            }
            CVTT_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            return CVTTB200_OK;
        }
        ~StreamLease()
        {
            if (stream)
            {
                std::lock_guard<std::mutex> lock(ctx->streamMutex);
                ctx->idleStreams.push_back(stream);
            }
        }
    };

    int launch_decode(DeviceContext &ctx, int format, const void *dIn, size_t nBlocks, void *dOut, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        const unsigned grid = (unsigned)std::min<size_t>((nBlocks + kDecodeThreads - 1) / kDecodeThreads, (size_t)ctx.numSMs * 8);
        const uint4 *in = (const uint4 *)dIn;
        uint4 *out = (uint4 *)dOut;
        if (format == CVTTB200_BC7)
            decode_kernel<0><<<grid, kDecodeThreads, 0, stream>>>(in, out, (uint32_t)nBlocks);
        else if (format == CVTTB200_BC6HU)
            decode_kernel<1><<<grid, kDecodeThreads, 0, stream>>>(in, out, (uint32_t)nBlocks);
        else
            decode_kernel<2><<<grid, kDecodeThreads, 0, stream>>>(in, out, (uint32_t)nBlocks);
        g_launches++;
        CVTT_CUDA(cudaGetLastError());
        return CVTTB200_OK;
    }

    // what an encode call needs besides its buffers
    struct EncodeRequest
    {
        int format;
        OptionsPOD options;
        OptionsPOD etc2AllocOptions;          // the Options AllocETC2Data was called with (chroma side axes, ETC.cpp:3117-3145)
        BC7PlanPOD plan;
        float rcpN[17];
    };

    // format -> kernel launch; the context's device is current
    int dispatch_encode(DeviceContext &ctx, const EncodeRequest &rq, const void *dIn, size_t nBlocks, void *dOut, cudaStream_t stream)
    {
        const int format = rq.format;
        if (format == CVTTB200_BC7)
            return launch_bc7(ctx, dIn, nBlocks, dOut, rq.options, rq.plan, rq.rcpN, stream);
        if (format <= CVTTB200_BC5S)
            return launch_s3tc(format, dIn, nBlocks, dOut, rq.options, rq.rcpN, stream);
        if (format == CVTTB200_BC6HU || format == CVTTB200_BC6HS)
            return launch_bc6h(ctx, dIn, nBlocks, dOut, rq.options, format == CVTTB200_BC6HS, rq.rcpN, stream);
        return launch_etc(ctx, format, dIn, nBlocks, dOut, rq.options, rq.etc2AllocOptions, stream);
    }

    int check_encode_arguments(int format, const void *blocks, size_t nBlocks, const void *out, const cvttb200_options *options, const cvttb200_bc7_plan *plan)
    {
        if (!blocks || !out || !options)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "null argument");
        if (nBlocks % 8 != 0)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "nBlocks must be a multiple of 8 (cvtt::NumParallelBlocks)");
        if (!cvttb200_input_block_bytes(format))
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "unknown format");
        if (format == CVTTB200_BC7 && !plan)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "CVTTB200_BC7 needs an encoding plan");
        return CVTTB200_OK;
    }

    void fill_request(EncodeRequest &rq, int format, const cvttb200_options *options, const cvttb200_bc7_plan *plan, const cvttb200_options *etc2AllocOptions)
    {
        rq.format = format;
        memcpy(&rq.options, options, sizeof(rq.options));
        memcpy(&rq.etc2AllocOptions, etc2AllocOptions ? etc2AllocOptions : options, sizeof(rq.etc2AllocOptions));
        if (plan)
            memcpy(&rq.plan, plan, sizeof(rq.plan));
        else
            memset(&rq.plan, 0, sizeof(rq.plan));
    }
}

// =========================================================================================================
// C ABI

extern "C"
{

int cvttb200_init(int device)
{
    DeviceContext *ctx = nullptr;
    return get_context(device, &ctx);
}

void cvttb200_shutdown(void)
{
    std::lock_guard<std::mutex> multi(g_multiMutex);
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceRestore restore;
    if (cudaGetDevice(&restore.prev) != cudaSuccess)
    {
        cudaGetLastError();
        restore.prev = -1;
        g_contexts.clear();
        return;
    }
    for (size_t i = 0; i < g_contexts.size(); i++)
    {
        DeviceContext &c = g_contexts[i];
        if (cudaSetDevice(c.device) != cudaSuccess)
            continue;
        cudaDeviceSynchronize();
        for (size_t k = 0; k < c.plans.size(); k++)
            cudaFree(c.plans[k].dCmds);
        if (c.setupStream) cudaStreamDestroy(c.setupStream);
        if (c.multiStream) cudaStreamDestroy(c.multiStream);
        for (int k = 0; k < 2; k++)
            if (c.pipeStream[k]) cudaStreamDestroy(c.pipeStream[k]);
        for (size_t k = 0; k < c.idleStreams.size(); k++)
            cudaStreamDestroy(c.idleStreams[k]);
        if (c.pool) cudaMemPoolDestroy(c.pool);
    }
    g_contexts.clear();
}

int cvttb200_set_rcp_table(const float *rcp17)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (rcp17)
    {
        memcpy(g_rcpN, rcp17, sizeof(g_rcpN));
        g_rcpOverridden = true;
    }
    else
    {
        host_rcp_table(g_rcpN);
        g_rcpOverridden = false;
    }
    g_rcpReady = true;
    return CVTTB200_OK;
}

int cvttb200_get_rcp_table(float *rcp17)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!rcp17)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "null table");
    if (!g_rcpReady)
    {
        host_rcp_table(g_rcpN);
        g_rcpReady = true;
    }
    memcpy(rcp17, g_rcpN, sizeof(g_rcpN));
    return CVTTB200_OK;
}

const char *cvttb200_last_error(void) { return t_lastError.c_str(); }

// etc2packer/etc2packer.cpp:114-193: what the sample packer puts in front of the encoded blocks
int cvttb200_ktx_header(int format, int width, int height, void *header68)
{
    if (!header68 || width <= 0 || height <= 0)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "bad KTX header request");
    uint32_t glInternalFormat, glBaseInternalFormat, blockBytes = 8;
    switch (format)
    {
    case CVTTB200_ETC1: glInternalFormat = 0x8D64; glBaseInternalFormat = 0x1907; break;
    case CVTTB200_ETC2: glInternalFormat = 0x9274; glBaseInternalFormat = 0x1907; break;
    case CVTTB200_ETC2_RGBA: glInternalFormat = 0x9278; glBaseInternalFormat = 0x1908; blockBytes = 16; break;
    case CVTTB200_ETC2_PUNCHTHROUGH: glInternalFormat = 0x9276; glBaseInternalFormat = 0x1908; break;
    case CVTTB200_EAC_R11U: glInternalFormat = 0x9270; glBaseInternalFormat = 0x1903; break;
    case CVTTB200_EAC_R11S: glInternalFormat = 0x9271; glBaseInternalFormat = 0x1903; break;
    default:
        return fail(CVTTB200_ERR_UNSUPPORTED, "the KTX writer covers the sample packer's targets: ETC1, ETC2, ETC2_RGBA, ETC2_PUNCHTHROUGH, EAC_R11U, EAC_R11S");
    }
    static const uint8_t identifier[12] = { 0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A };
    const uint32_t words[14] = {
        0x04030201u,                 // endianness
        0u, 1u, 0u,                  // glType, glTypeSize, glFormat (compressed data)
        glInternalFormat, glBaseInternalFormat,
        (uint32_t)width, (uint32_t)height, 0u,      // pixelDepth
        0u, 1u, 1u,                  // array elements, faces, mip levels
        0u,                          // bytesOfKeyValueData
        (uint32_t)((width + 3) / 4) * (uint32_t)((height + 3) / 4) * blockBytes       // imageSize of the one level
    };
    memcpy(header68, identifier, 12);
    memcpy((uint8_t *)header68 + 12, words, sizeof(words));
    return CVTTB200_OK;
}

size_t cvttb200_tiled_block_count(int width, int height)
{
    if (width <= 0 || height <= 0)
        return 0;
    return (size_t)((height + 3) / 4) * (size_t)((width + 31) / 32) * 8;
}

int cvttb200_tile_image(int pixelBytes, const void *image, int width, int height, size_t rowPitchBytes, void *blocks, void *streamPtr)
{
    if (!image || !blocks || width <= 0 || height <= 0 || (pixelBytes != 4 && pixelBytes != 8) || rowPitchBytes < (size_t)width * (size_t)pixelBytes)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "bad image description");
    if (!is_device_pointer(image) || !is_device_pointer(blocks))
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "cvttb200_tile_image works on device memory");
    DeviceContext *ctx = nullptr;
    int rc = current_context(&ctx);
    if (rc != CVTTB200_OK)
        return rc;
    const int blocksPerRow = ((width + 31) / 32) * 8;
    const size_t nBlocks = cvttb200_tiled_block_count(width, height);
    if (nBlocks > 0xffffff00u)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "image too large for one call");
    const int vectorOK = ((uintptr_t)image % 16 == 0) && (rowPitchBytes % 16 == 0);
    const unsigned grid = (unsigned)((nBlocks * 4 + 256 * kTileUnroll - 1) / (256 * kTileUnroll));
    cudaStream_t stream = (cudaStream_t)streamPtr;
    if (pixelBytes == 4)
        tile_image_kernel<4><<<grid, 256, 0, stream>>>((const unsigned char *)image, width, height, rowPitchBytes, blocksPerRow, (uint32_t)nBlocks, (uint4 *)blocks, vectorOK);
    else
        tile_image_kernel<8><<<grid, 256, 0, stream>>>((const unsigned char *)image, width, height, rowPitchBytes, blocksPerRow, (uint32_t)nBlocks, (uint4 *)blocks, vectorOK);
    g_launches++;
    CVTT_CUDA(cudaGetLastError());
    return CVTTB200_OK;
}

int cvttb200_untile_blocks(const void *encoded, int width, int height, size_t blockBytes, void *out, void *streamPtr)
{
    if (!encoded || !out || width <= 0 || height <= 0 || (blockBytes != 8 && blockBytes != 16))
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "bad arguments");
    if (!is_device_pointer(encoded) || !is_device_pointer(out))
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "cvttb200_untile_blocks works on device memory");
    const int blocksPerRow = ((width + 31) / 32) * 8, realBlocksPerRow = (width + 3) / 4;
    const uint64_t nOut = (uint64_t)((height + 3) / 4) * (uint64_t)realBlocksPerRow;
    const unsigned grid = (unsigned)((nOut + 255) / 256);
    if (blockBytes == 16 && (uintptr_t)encoded % 16 == 0 && (uintptr_t)out % 16 == 0)
        untile_blocks_kernel<uint4><<<grid, 256, 0, (cudaStream_t)streamPtr>>>((const uint4 *)encoded, blocksPerRow, realBlocksPerRow, nOut, (uint4 *)out);
    else if (blockBytes == 16)
        untile_blocks_kernel<uint2><<<(unsigned)((nOut * 2 + 255) / 256), 256, 0, (cudaStream_t)streamPtr>>>((const uint2 *)encoded, blocksPerRow * 2, realBlocksPerRow * 2, nOut * 2, (uint2 *)out);
    else
        untile_blocks_kernel<uint2><<<grid, 256, 0, (cudaStream_t)streamPtr>>>((const uint2 *)encoded, blocksPerRow, realBlocksPerRow, nOut, (uint2 *)out);
    g_launches++;
    CVTT_CUDA(cudaGetLastError());
    return CVTTB200_OK;
}

int cvttb200_selftest(uint64_t samples, uint64_t seed, uint64_t *mismatches)
{
    if (!mismatches)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "null argument");
    DeviceContext *ctx = nullptr;
    int rc = current_context(&ctx);
    if (rc != CVTTB200_OK)
        return rc;
    return bc7_selftest_div(samples, seed, mismatches);
}

uint64_t cvttb200_launch_count(void) { return g_launches.load(); }

void cvttb200_options_default(cvttb200_options *o)
{
    // cvtt::Options::Options(), ConvectionKernels.h:89-101
    o->flags = CVTTB200_FLAG_BC7_FAST_INDEXING | CVTTB200_FLAG_S3TC_PARANOID;
    o->threshold = 0.5f;
    o->redWeight = 0.2125f / 0.7154f;
    o->greenWeight = 1.0f;
    o->blueWeight = 0.0721f / 0.7154f;
    o->alphaWeight = 1.0f;
    o->refineRoundsBC7 = 2;
    o->refineRoundsBC6H = 3;
    o->refineRoundsIIC = 8;
    o->refineRoundsS3TC = 2;
    o->seedPoints = 4;
}

void cvttb200_bc7_plan_default(cvttb200_bc7_plan *plan) { bc7_plan_default(*reinterpret_cast<BC7PlanPOD *>(plan)); }

void cvttb200_bc7_plan_from_quality(cvttb200_bc7_plan *plan, int quality) { bc7_plan_from_quality(*reinterpret_cast<BC7PlanPOD *>(plan), quality); }

int cvttb200_bc7_plan_from_fine_tuning(cvttb200_bc7_plan *plan, const cvttb200_bc7_fine_tuning *params)
{
    return bc7_plan_from_fine_tuning(*reinterpret_cast<BC7PlanPOD *>(plan), *reinterpret_cast<const BC7FineTuningPOD *>(params)) ? 1 : 0;
}

void cvttb200_bc7_fine_tuning_default(cvttb200_bc7_fine_tuning *params)
{
    memset(params, 4, sizeof(*params));   // BC7FineTuningParams(): every seed-point count is 4, ConvectionKernels.h:117-139
}

size_t cvttb200_input_block_bytes(int format)
{
    switch (format)
    {
    case CVTTB200_BC6HU: case CVTTB200_BC6HS: return 128;
    case CVTTB200_EAC_R11U: case CVTTB200_EAC_R11S: return 32;
    default: return (format >= CVTTB200_BC1 && format <= CVTTB200_EAC_R11S) ? 64 : 0;
    }
}

size_t cvttb200_output_block_bytes(int format)
{
    switch (format)
    {
    case CVTTB200_BC1: case CVTTB200_BC4U: case CVTTB200_BC4S: case CVTTB200_ETC1: case CVTTB200_ETC2:
    case CVTTB200_ETC2_PUNCHTHROUGH: case CVTTB200_ETC2_ALPHA: case CVTTB200_EAC_R11U: case CVTTB200_EAC_R11S:
        return 8;
    default:
        return (format >= CVTTB200_BC1 && format <= CVTTB200_EAC_R11S) ? 16 : 0;
    }
}

int cvttb200_encode_ex(int format, const void *blocks, size_t nBlocks, void *out, const cvttb200_options *options, const cvttb200_bc7_plan *plan,
                       const cvttb200_options *etc2AllocOptions, void *streamPtr)
{
    int rc = check_encode_arguments(format, blocks, nBlocks, out, options, plan);
    if (rc != CVTTB200_OK)
        return rc;
    if (nBlocks == 0)
        return CVTTB200_OK;
    const size_t inBytes = cvttb200_input_block_bytes(format), outBytes = cvttb200_output_block_bytes(format);

    EncodeRequest rq;
    fill_request(rq, format, options, plan, etc2AllocOptions);
    DeviceContext *ctx = nullptr;
    rc = current_context(&ctx, rq.rcpN);
    if (rc != CVTTB200_OK)
        return rc;

    cudaStream_t stream = (cudaStream_t)streamPtr;
    const bool inOnDevice = is_device_pointer(blocks), outOnDevice = is_device_pointer(out);
    if (inOnDevice && outOnDevice)
        return dispatch_encode(*ctx, rq, blocks, nBlocks, out, stream);       // enqueued; no synchronisation

    const bool fastFormat = format <= CVTTB200_BC5S || format == CVTTB200_ETC2_ALPHA || format == CVTTB200_EAC_R11U || format == CVTTB200_EAC_R11S;
    const size_t kChunkBlocks = 131072;
    const bool pipelined = !inOnDevice && !outOnDevice && fastFormat && nBlocks >= 2 * kChunkBlocks;

    // Host buffers in and out, and a format whose kernel is shorter than its PCIe transfers (BC1-BC5, EAC: >= 100 Mblocks/s on
    // the device): chunks of whole groups alternate between two streams, so that the copy-in of one chunk, the kernel of the
    // previous one and the copy-out of the one before overlap (both copy engines and the SMs busy).  The search formats
    // (BC7, BC6H, ETC colour) spend > 97 % of an end-to-end call in the kernel and lose more to the extra partial waves of
    // chunked launches than the overlap returns, so they stay one launch.
    if (pipelined)
    {
        std::lock_guard<std::mutex> pipe(ctx->pipeMutex);         // one pipelined call per device at a time: they are PCIe-bound
        for (int k = 0; k < 2; k++)
            if (!ctx->pipeStream[k])
                CVTT_CUDA(cudaStreamCreateWithFlags(&ctx->pipeStream[k], cudaStreamNonBlocking));
        PoolBuffer stageIn, stageOut;
        rc = stageIn.alloc(*ctx, nBlocks * inBytes, ctx->pipeStream[0]);
        if (rc == CVTTB200_OK)
            rc = stageOut.alloc(*ctx, nBlocks * outBytes, ctx->pipeStream[0]);
        if (rc != CVTTB200_OK)
            return rc;
        CVTT_CUDA(cudaStreamSynchronize(ctx->pipeStream[0]));     // the allocations are usable from the second stream as well
        int chunk = 0;
        for (size_t first = 0; first < nBlocks; first += kChunkBlocks, chunk++)
        {
            const size_t n = std::min(kChunkBlocks, nBlocks - first);
            cudaStream_t s = ctx->pipeStream[chunk & 1];
            unsigned char *cIn = (unsigned char *)stageIn.p + first * inBytes, *cOut = (unsigned char *)stageOut.p + first * outBytes;
            cudaError_t e = cudaMemcpyAsync(cIn, (const unsigned char *)blocks + first * inBytes, n * inBytes, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess)
            {
                rc = fail_cuda(e, "cudaMemcpyAsync (host to device)");
                break;
            }
            rc = dispatch_encode(*ctx, rq, cIn, n, cOut, s);
            if (rc != CVTTB200_OK)
                break;
            e = cudaMemcpyAsync((unsigned char *)out + first * outBytes, cOut, n * outBytes, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess)
            {
                rc = fail_cuda(e, "cudaMemcpyAsync (device to host)");
                break;
            }
        }
        for (int k = 0; k < 2; k++)
        {
            const cudaError_t e = cudaStreamSynchronize(ctx->pipeStream[k]);
            if (e != cudaSuccess && rc == CVTTB200_OK)
                rc = fail_cuda(e, "cudaStreamSynchronize");
        }
        return rc;          // the staging buffers go back to the pool on pipeStream[0], which is idle now
    }

    StreamLease lease;              // declared before the buffers: they are returned to the pool on this stream first
    if (!inOnDevice && !outOnDevice && !stream)
    {
        rc = lease.acquire(*ctx);
        if (rc != CVTTB200_OK)
            return rc;
        stream = lease.stream;
    }
    PoolBuffer stageIn, stageOut;
    const void *dIn = blocks;
    void *dOut = out;
    if (!inOnDevice)
    {
        rc = stageIn.alloc(*ctx, nBlocks * inBytes, stream);
        if (rc != CVTTB200_OK)
            return rc;
        CVTT_CUDA(cudaMemcpyAsync(stageIn.p, blocks, nBlocks * inBytes, cudaMemcpyHostToDevice, stream));
        dIn = stageIn.p;
    }
    if (!outOnDevice)
    {
        rc = stageOut.alloc(*ctx, nBlocks * outBytes, stream);
        if (rc != CVTTB200_OK)
            return rc;
        dOut = stageOut.p;
    }
    rc = dispatch_encode(*ctx, rq, dIn, nBlocks, dOut, stream);
    if (rc != CVTTB200_OK)
        return rc;
    if (!outOnDevice)
        CVTT_CUDA(cudaMemcpyAsync(out, dOut, nBlocks * outBytes, cudaMemcpyDeviceToHost, stream));
    CVTT_CUDA(cudaStreamSynchronize(stream));
    return CVTTB200_OK;
}

int cvttb200_encode(int format, const void *blocks, size_t nBlocks, void *out, const cvttb200_options *options, const cvttb200_bc7_plan *plan, void *streamPtr)
{
    return cvttb200_encode_ex(format, blocks, nBlocks, out, options, plan, nullptr, streamPtr);
}

int cvttb200_encode_multi(int format, const void *blocks, size_t nBlocks, void *out, const cvttb200_options *options, const cvttb200_bc7_plan *plan,
                          const int *devices, int nDevices)
{
    int rc = check_encode_arguments(format, blocks, nBlocks, out, options, plan);
    if (rc != CVTTB200_OK)
        return rc;
    const size_t inBytes = cvttb200_input_block_bytes(format), outBytes = cvttb200_output_block_bytes(format);
    if (is_device_pointer(blocks) || is_device_pointer(out))
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "cvttb200_encode_multi takes host buffers (use cvttb200_encode per device for device memory)");

    int count = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            return fail(CVTTB200_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + " (libcvtt_b200 has no CPU fallback)");
    }
    if (nDevices <= 0)
        nDevices = count;
    std::vector<int> ids((size_t)nDevices);
    for (int i = 0; i < nDevices; i++)
    {
        ids[(size_t)i] = devices ? devices[i] : i;
        if (ids[(size_t)i] < 0 || ids[(size_t)i] >= count)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "device index out of range");
        for (int k = 0; k < i; k++)
            if (ids[(size_t)k] == ids[(size_t)i])
                return fail(CVTTB200_ERR_BAD_ARGUMENT, "a device is listed twice");
    }
    if (nBlocks == 0)
        return CVTTB200_OK;

    std::lock_guard<std::mutex> multi(g_multiMutex);
    DeviceRestore restore;
    CVTT_CUDA(cudaGetDevice(&restore.prev));

    EncodeRequest rq;
    fill_request(rq, format, options, plan, nullptr);

    // Contiguous ranges of whole 8-block groups (a group is one reference call and is never split).  Two passes, so that the
    // devices work at the same time whatever kind of host memory the caller has: a copy to or from PAGEABLE memory only
    // returns when its data has moved, so a copy-out issued right behind a device's kernel would hold the loop until that
    // kernel is done, and the next device would not even have started.  Pass 1 gives every device its input and its kernel;
    // pass 2 collects the results (the copy-out of device i waits for kernel i while the other kernels keep running).
    struct Part
    {
        DeviceContext *ctx;
        size_t first, n;
        void *dIn, *dOut;
    };
    std::vector<Part> parts;
    const size_t nGroups = nBlocks / 8;
    for (int i = 0; i < nDevices && rc == CVTTB200_OK; i++)
    {
        Part part;
        part.first = nGroups * (size_t)i / (size_t)nDevices * 8;
        part.n = nGroups * (size_t)(i + 1) / (size_t)nDevices * 8 - part.first;
        part.dIn = part.dOut = nullptr;
        if (part.n == 0)
            continue;
        cudaError_t e = cudaSetDevice(ids[(size_t)i]);
        if (e != cudaSuccess)
        {
            rc = fail_cuda(e, "cudaSetDevice");
            break;
        }
        rc = get_context(ids[(size_t)i], &part.ctx, rq.rcpN);
        if (rc != CVTTB200_OK)
            break;
        DeviceContext *ctx = part.ctx;
        if (!ctx->multiStream && (e = cudaStreamCreateWithFlags(&ctx->multiStream, cudaStreamNonBlocking)) != cudaSuccess)
        {
            rc = fail_cuda(e, "cudaStreamCreateWithFlags");
            break;
        }
        rc = pool_alloc(*ctx, &part.dIn, part.n * inBytes, ctx->multiStream);
        if (rc == CVTTB200_OK)
            rc = pool_alloc(*ctx, &part.dOut, part.n * outBytes, ctx->multiStream);
        parts.push_back(part);                  // from here on the buffers are released below
        if (rc != CVTTB200_OK)
            break;
        if ((e = cudaMemcpyAsync(part.dIn, (const unsigned char *)blocks + part.first * inBytes, part.n * inBytes, cudaMemcpyHostToDevice, ctx->multiStream)) != cudaSuccess)
        {
            rc = fail_cuda(e, "cudaMemcpyAsync (host to device)");
            break;
        }
        rc = dispatch_encode(*ctx, rq, part.dIn, part.n, part.dOut, ctx->multiStream);
    }
    for (size_t k = 0; k < parts.size() && rc == CVTTB200_OK; k++)
    {
        cudaSetDevice(parts[k].ctx->device);
        const cudaError_t e = cudaMemcpyAsync((unsigned char *)out + parts[k].first * outBytes, parts[k].dOut, parts[k].n * outBytes, cudaMemcpyDeviceToHost, parts[k].ctx->multiStream);
        if (e != cudaSuccess)
            rc = fail_cuda(e, "cudaMemcpyAsync (device to host)");
    }
    for (size_t k = 0; k < parts.size(); k++)
    {
        cudaSetDevice(parts[k].ctx->device);
        const cudaError_t e = cudaStreamSynchronize(parts[k].ctx->multiStream);
        if (e != cudaSuccess && rc == CVTTB200_OK)
            rc = fail_cuda(e, "cudaStreamSynchronize");
        if (parts[k].dIn) cudaFreeAsync(parts[k].dIn, parts[k].ctx->multiStream);
        if (parts[k].dOut) cudaFreeAsync(parts[k].dOut, parts[k].ctx->multiStream);
    }
    return rc;
}

int cvttb200_decode(int format, const void *encoded, size_t nBlocks, void *pixelBlocks, void *streamPtr)
{
    if (!encoded || !pixelBlocks)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "null argument");
    if (format != CVTTB200_BC7 && format != CVTTB200_BC6HU && format != CVTTB200_BC6HS)
        return fail(CVTTB200_ERR_UNSUPPORTED, "the reference decodes BC7, BC6HU and BC6HS only");
    if (nBlocks == 0)
        return CVTTB200_OK;
    const size_t inBytes = 16, outBytes = cvttb200_input_block_bytes(format);      // a decoded block is the encoder's input block

    DeviceContext *ctx = nullptr;
    int rc = current_context(&ctx);
    if (rc != CVTTB200_OK)
        return rc;

    cudaStream_t stream = (cudaStream_t)streamPtr;
    const bool inOnDevice = is_device_pointer(encoded), outOnDevice = is_device_pointer(pixelBlocks);
    StreamLease lease;
    if (!inOnDevice && !outOnDevice && !stream)
    {
        rc = lease.acquire(*ctx);
        if (rc != CVTTB200_OK)
            return rc;
        stream = lease.stream;
    }
    PoolBuffer stageIn, stageOut;
    const void *dIn = encoded;
    void *dOut = pixelBlocks;
    if (!inOnDevice)
    {
        rc = stageIn.alloc(*ctx, nBlocks * inBytes, stream);
        if (rc != CVTTB200_OK)
            return rc;
        CVTT_CUDA(cudaMemcpyAsync(stageIn.p, encoded, nBlocks * inBytes, cudaMemcpyHostToDevice, stream));
        dIn = stageIn.p;
    }
    if (!outOnDevice)
    {
        rc = stageOut.alloc(*ctx, nBlocks * outBytes, stream);
        if (rc != CVTTB200_OK)
            return rc;
        dOut = stageOut.p;
    }
    rc = launch_decode(*ctx, format, dIn, nBlocks, dOut, stream);
    if (rc != CVTTB200_OK)
        return rc;
    if (!outOnDevice)
        CVTT_CUDA(cudaMemcpyAsync(pixelBlocks, dOut, nBlocks * outBytes, cudaMemcpyDeviceToHost, stream));
    if (!inOnDevice || !outOnDevice)
        CVTT_CUDA(cudaStreamSynchronize(stream));
    return CVTTB200_OK;
}

} // extern "C"
