// BC6H translation unit of libcvtt_b200.so: encode kernels (signed / unsigned, slow / fast indexing), launch, set-up.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cvtt_internal.h"
#include "bc6h_host.h"
#include "cvtt_segment.cuh"

using namespace cvttb200;

namespace
{
    constexpr int kBC6HThreads = 128;
    // per thread: 48 pre-weighted floats, 32 words of raw pixels, and the 24-word interpolator table of the slow index search
    constexpr size_t kBC6HSmemBytesSlow = (size_t)kBC6HThreads * 104 * 4, kBC6HSmemBytesFast = (size_t)kBC6HThreads * 80 * 4;

    __constant__ BC6HTables c_bc6hTables;

    // One thread per block, warp = 4 reference groups.  Input: PixelBlockF16 = int16 [16][4] (128 B, alpha ignored), read
    // with eight 128-bit loads per thread; converted once into [word][thread] planes in shared memory (416 B per thread with
    // the interpolator table, so that four CTAs = 16 warps fit an SM).
    template<bool SIGNED, bool FAST>
    __global__ void __launch_bounds__(kBC6HThreads, 4)
    bc6h_encode_kernel(const __grid_constant__ BC6HParams P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nBlocks)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        float *sPw = reinterpret_cast<float *>(smem);
        uint32_t *sRaw = reinterpret_cast<uint32_t *>(sPw + 48 * kBC6HThreads);

        const uint32_t tid = threadIdx.x;
        const uint32_t block = blockIdx.x * kBC6HThreads + tid;
        const bool active = block < nBlocks;

        BC6HLane<kBC6HThreads> L;
        L.pw = sPw + tid;
        L.raw = sRaw + tid;
        L.tab = sRaw + 32 * kBC6HThreads + tid;           // only touched with slow indexing
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (active)
                v = __ldg(in + (size_t)block * 8 + q);
            bc6h_load_pixel<SIGNED>(P, L, 2 * q, (int)(v.x & 0xffffu), (int)(v.x >> 16), (int)(v.y & 0xffffu));
            bc6h_load_pixel<SIGNED>(P, L, 2 * q + 1, (int)(v.z & 0xffffu), (int)(v.z >> 16), (int)(v.w & 0xffffu));
        }
        __syncwarp();

        SegmentVote vote;
        vote.segMask = 0xffu << (tid & 24);
        uint32_t o[4];
        bc6h_encode_block<SIGNED, FAST, kBC6HThreads>(P, c_bc6hTables, L, vote, o);
        if (active)
            out[block] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

namespace cvttb200
{
    int bc6h_device_setup()
    {
        CVTT_CUDA(cudaMemcpyToSymbol(c_bc6hTables, &bc6h_tables(), sizeof(BC6HTables)));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        return CVTTB200_OK;
    }

    int launch_bc6h(const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, bool isSigned, const float *rcpN, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        BC6HParams P;
        bc6h_fill_params(P, options, rcpN);
        const unsigned grid = (unsigned)((nBlocks + kBC6HThreads - 1) / kBC6HThreads);
        const bool fast = (options.flags & kFlag_BC6H_FastIndexing) != 0;
        const uint4 *in = (const uint4 *)dIn;
        uint4 *out = (uint4 *)dOut;
        if (isSigned)
        {
            if (fast) bc6h_encode_kernel<true, true><<<grid, kBC6HThreads, kBC6HSmemBytesFast, stream>>>(P, in, out, (uint32_t)nBlocks);
            else bc6h_encode_kernel<true, false><<<grid, kBC6HThreads, kBC6HSmemBytesSlow, stream>>>(P, in, out, (uint32_t)nBlocks);
        }
        else
        {
            if (fast) bc6h_encode_kernel<false, true><<<grid, kBC6HThreads, kBC6HSmemBytesFast, stream>>>(P, in, out, (uint32_t)nBlocks);
            else bc6h_encode_kernel<false, false><<<grid, kBC6HThreads, kBC6HSmemBytesSlow, stream>>>(P, in, out, (uint32_t)nBlocks);
        }
        g_launches++;
        CVTT_CUDA(cudaGetLastError());
        return CVTTB200_OK;
    }
}
