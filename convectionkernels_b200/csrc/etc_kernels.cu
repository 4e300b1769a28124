// ETC1 / ETC2 / EAC translation unit of libcvtt_b200.so: encode kernels, scratch allocation, launch, set-up.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cvtt_internal.h"
#include "etc_host.h"
#include "cvtt_segment.cuh"

using namespace cvttb200;

namespace
{
    constexpr int kETCThreads = 256;      // 8 warps = 32 reference groups per CTA in phase lock-step, two CTAs per SM (10.6 Mblocks/s; 1 x 16 warps: 10.25, 4 x 4: 9.0)
    constexpr int kETCCtasPerSM = 2;
    constexpr size_t kETCSmemBytes = (size_t)kETCThreads * 16 * sizeof(F4);

    __constant__ ETCTables c_etcTables;

    // EAC alpha of a block whose 16 alphas are all `a`: the search is a pure function of the 16 values, so its result for the 256
    // constant blocks is computed once per device by the search itself (eac_flat_lut_kernel) and looked up when a whole warp
    // holds such blocks (opaque regions).  The reference runs the search every time.
    __device__ uint2 g_eacFlatLut[256];

    __global__ void eac_flat_lut_kernel()
    {
        int a[16];
        for (int px = 0; px < 16; px++)
            a[px] = (int)threadIdx.x;
        uint32_t o[2];
        etc_alpha_encode_block(c_etcTables, a, false, false, o);
        g_eacFlatLut[threadIdx.x] = make_uint2(o[0], o[1]);
    }

    // 8-bit EAC alpha for one thread's block; a warp whose blocks are all constant in alpha takes the table
    __device__ __forceinline__ void eac_alpha8(const int *alpha, uint32_t out[2])
    {
        bool flat = true;
#pragma unroll
        for (int px = 1; px < 16; px++)
            flat = flat && alpha[px] == alpha[0];
        if (__all_sync(0xffffffffu, flat))
        {
            const uint2 v = g_eacFlatLut[alpha[0]];
            out[0] = v.x;
            out[1] = v.y;
            return;
        }
        etc_alpha_encode_block(c_etcTables, alpha, false, false, out);
    }



    enum { kETCKindETC1 = 0, kETCKindETC2 = 1, kETCKindETC2RGBA = 2, kETCKindETC2Punchthrough = 3 };

    // Persistent kernel: the grid is sized to the device (SMs x resident CTAs), every warp walks 32-block slices of the
    // input.  One thread per block; the per-thread scratch of the differential / H-mode searches (the reference's
    // ETC2CompressionData) is a slice of one global allocation, laid out [entry][thread].
    template<int KIND, bool UNIFORM, bool BT709>
    __global__ void __launch_bounds__(kETCThreads, kETCCtasPerSM)
    etc_encode_kernel(const __grid_constant__ ETCParams P, const uint4 *__restrict__ in, uint32_t *__restrict__ out, uint32_t nBlocks, ETCScratch scratch)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        F4 *sPw = reinterpret_cast<F4 *>(smem);
        const uint32_t tid = threadIdx.x;
        const uint32_t gthread = blockIdx.x * kETCThreads + tid;

        ETCScratch S = scratch;
        S.drsErr += gthread;
        S.drsMeta += gthread;
        S.hErr += gthread;
        S.hMeta += gthread;

        ETCLane<kETCThreads> L;
        L.pw = sPw + tid;
        SegmentMax vote;

        // CTA-uniform trip count: every thread of the CTA takes part in the phase barriers of the encode functions
        for (uint32_t tileBase = blockIdx.x * kETCThreads; tileBase < nBlocks; tileBase += gridDim.x * kETCThreads)
        {
            const uint32_t block = tileBase + tid;
            const bool active = block < nBlocks;
            int alpha[16];
            uint32_t transparentMask = 0;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (active)
                    v = __ldg(in + (size_t)block * 4 + q);
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    F4 p;
                    const float r = (float)(w[k] & 0xffu), g = (float)((w[k] >> 8) & 0xffu), b = (float)((w[k] >> 16) & 0xffu);
                    if (BT709)
                    {
                        // ExtractBlocks with Flags::ETC_UseFakeBT709: the "pre-weighted" pixel is its fake-BT.709 YUV (ETC.cpp:2142-2143)
                        float yuv[3];
                        etc_to_bt709(r, g, b, yuv);
                        p.x = yuv[0];
                        p.y = yuv[1];
                        p.z = yuv[2];
                    }
                    else
                    {
                        p.x = UNIFORM ? r : r * P.w[0];
                        p.y = UNIFORM ? g : g * P.w[1];
                        p.z = UNIFORM ? b : b * P.w[2];
                    }
                    p.w = __uint_as_float(w[k]);
                    alpha[q * 4 + k] = (int)(w[k] >> 24);
                    if (KIND == kETCKindETC2Punchthrough && alpha[q * 4 + k] < P.punchThreshold)
                    {
                        // CompressETC2Block zeroes the transparent pixels, ETC.cpp:1705-1718
                        transparentMask |= 1u << (q * 4 + k);
                        p.x = p.y = p.z = 0.0f;
                        p.w = __uint_as_float(w[k] & 0xff000000u);
                    }
                    sPw[(q * 4 + k) * kETCThreads + tid] = p;
                }
            }
            __syncwarp();

            uint32_t color[2];
            if (KIND == kETCKindETC1)
                etc1_encode_block<UNIFORM, BT709, kETCThreads>(P, c_etcTables, L, S, color);
            else if (KIND == kETCKindETC2Punchthrough)
                etc2_punchthrough_encode_block<UNIFORM, BT709, kETCThreads>(P, c_etcTables, L, S, vote, transparentMask, color);
            else
                etc2_encode_block<UNIFORM, BT709, kETCThreads>(P, c_etcTables, L, S, vote, color);

            if (KIND == kETCKindETC2RGBA)
            {
                uint32_t a[2];
                eac_alpha8(alpha, a);
                if (active)
                    reinterpret_cast<uint4 *>(out)[block] = make_uint4(etc_bswap(a[0]), etc_bswap(a[1]), etc_bswap(color[0]), etc_bswap(color[1]));
            }
            else if (active)
                reinterpret_cast<uint2 *>(out)[block] = make_uint2(etc_bswap(color[0]), etc_bswap(color[1]));
            __syncwarp();
        }
    }

    // EncodeETC2Alpha (8-bit alpha of PixelBlockU8) and EncodeETC2Alpha11 (PixelBlockScalarS16): pure integer, one thread per block
    // kind: 0 = 8-bit alpha, 1 = EAC R11 unsigned, 2 = EAC R11 signed
    template<int KIND>
    __global__ void __launch_bounds__(128)
    eac_encode_kernel(const void *__restrict__ in, uint2 *__restrict__ out, uint32_t nBlocks)
    {
        const uint32_t block = ::min(blockIdx.x * blockDim.x + threadIdx.x, nBlocks - 1);     // surplus threads repeat the last block
        int a[16];
        if (KIND == 0)
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(in) + (size_t)block * 4;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const uint4 v = __ldg(src + q);
                a[q * 4 + 0] = (int)(v.x >> 24);
                a[q * 4 + 1] = (int)(v.y >> 24);
                a[q * 4 + 2] = (int)(v.z >> 24);
                a[q * 4 + 3] = (int)(v.w >> 24);
            }
        }
        else
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(in) + (size_t)block * 2;
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                const uint4 v = __ldg(src + q);
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int k = 0; k < 4; k++)
                    for (int h = 0; h < 2; h++)
                    {
                        // CompressEACBlock, ETC.cpp:2087-2110
                        int px = (int)(int16_t)(uint16_t)(w[k] >> (16 * h));
                        if (KIND == 2)
                            px = ::max(1, ::min(px, 1023) + 1024);
                        else
                            px = ::max(0, ::min(px, 2047));
                        a[q * 8 + k * 2 + h] = px;
                    }
            }
        }
        uint32_t o[2];
        if (KIND == 0)
            eac_alpha8(a, o);
        else
            etc_alpha_encode_block(c_etcTables, a, true, KIND == 2, o);
        out[block] = make_uint2(etc_bswap(o[0]), etc_bswap(o[1]));
    }
}

namespace cvttb200
{
    int etc_device_setup()
    {
        CVTT_CUDA(cudaMemcpyToSymbol(c_etcTables, &etc_tables(), sizeof(ETCTables)));
        eac_flat_lut_kernel<<<1, 256>>>();
        CVTT_CUDA(cudaGetLastError());
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<0, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(etc_encode_kernel<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kETCSmemBytes));
        return CVTTB200_OK;
    }

    template<int KIND>
    static int launch_etc_color(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const ETCParams &P, bool uniform, bool bt709, cudaStream_t stream)
    {
        // resident threads: the whole device, or fewer for small inputs
        const size_t maxCtas = (size_t)ctx.numSMs * kETCCtasPerSM;
        const unsigned grid = (unsigned)std::min(maxCtas, (nBlocks + kETCThreads - 1) / kETCThreads);
        const size_t threads = (size_t)grid * kETCThreads;
        void *dScratch = nullptr;
        {
            const int rc = pool_alloc(ctx, &dScratch, etc_scratch_bytes(threads), stream);
            if (rc != CVTTB200_OK)
                return rc;
        }
        ETCScratch S;
        etc_scratch_layout(S, dScratch, threads);
        const uint4 *in = (const uint4 *)dIn;
        uint32_t *out = (uint32_t *)dOut;
        if (bt709)
        {
            if (uniform)
                etc_encode_kernel<KIND, true, true><<<grid, kETCThreads, kETCSmemBytes, stream>>>(P, in, out, (uint32_t)nBlocks, S);
            else
                etc_encode_kernel<KIND, false, true><<<grid, kETCThreads, kETCSmemBytes, stream>>>(P, in, out, (uint32_t)nBlocks, S);
        }
        else if (uniform)
            etc_encode_kernel<KIND, true, false><<<grid, kETCThreads, kETCSmemBytes, stream>>>(P, in, out, (uint32_t)nBlocks, S);
        else
            etc_encode_kernel<KIND, false, false><<<grid, kETCThreads, kETCSmemBytes, stream>>>(P, in, out, (uint32_t)nBlocks, S);
        g_launches++;
        CVTT_CUDA(cudaGetLastError());
        CVTT_CUDA(cudaFreeAsync(dScratch, stream));
        return CVTTB200_OK;
    }

    int launch_etc(DeviceContext &ctx, int format, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const OptionsPOD &allocOptions, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        const unsigned grid = (unsigned)((nBlocks + 127) / 128);
        if (format == CVTTB200_ETC2_ALPHA || format == CVTTB200_EAC_R11U || format == CVTTB200_EAC_R11S)
        {
            if (format == CVTTB200_ETC2_ALPHA)
                eac_encode_kernel<0><<<grid, 128, 0, stream>>>(dIn, (uint2 *)dOut, (uint32_t)nBlocks);
            else if (format == CVTTB200_EAC_R11U)
                eac_encode_kernel<1><<<grid, 128, 0, stream>>>(dIn, (uint2 *)dOut, (uint32_t)nBlocks);
            else
                eac_encode_kernel<2><<<grid, 128, 0, stream>>>(dIn, (uint2 *)dOut, (uint32_t)nBlocks);
            g_launches++;
            CVTT_CUDA(cudaGetLastError());
            return CVTTB200_OK;
        }
        ETCParams P;
        etc_fill_params(P, options, allocOptions);
        const bool uniform = (options.flags & kFlag_Uniform) != 0, bt709 = (options.flags & kFlag_ETC_UseFakeBT709) != 0;
        switch (format)
        {
        case CVTTB200_ETC1: return launch_etc_color<kETCKindETC1>(ctx, dIn, nBlocks, dOut, P, uniform, bt709, stream);
        case CVTTB200_ETC2: return launch_etc_color<kETCKindETC2>(ctx, dIn, nBlocks, dOut, P, uniform, bt709, stream);
        case CVTTB200_ETC2_RGBA: return launch_etc_color<kETCKindETC2RGBA>(ctx, dIn, nBlocks, dOut, P, uniform, bt709, stream);
        case CVTTB200_ETC2_PUNCHTHROUGH: return launch_etc_color<kETCKindETC2Punchthrough>(ctx, dIn, nBlocks, dOut, P, uniform, bt709, stream);
        default: return fail(CVTTB200_ERR_BAD_ARGUMENT, "not an ETC format");
        }
    }
}
