// BC1 - BC5 translation unit of libcvtt_b200.so: encode kernels and their launch.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cvtt_internal.h"
#include "s3tc_host.h"
#include "cvtt_segment.cuh"

using namespace cvttb200;

namespace
{
    constexpr int kS3TCThreads = 128;

    // One thread per block.  FMT is the cvttb200_format id (BC1 .. BC5S).  Pixels are expanded once to fp32 in shared memory
    // ([pixel][thread]); signed inputs are biased like Util::BiasSignedInput (Util.cpp:47-60).
    template<int FMT>
    __global__ void __launch_bounds__(kS3TCThreads)
    s3tc_encode_kernel(const __grid_constant__ S3TCParams P, const uint4 *__restrict__ in, uint32_t *__restrict__ out, uint32_t nBlocks)
    {
        __shared__ F4 sPx[16 * kS3TCThreads];
        const uint32_t tid = threadIdx.x;
        const uint32_t block = blockIdx.x * kS3TCThreads + tid;
        const bool active = block < nBlocks;      // whole warps stay alive: the exhaustive search uses a segment maximum
        constexpr bool isSigned = (FMT == CVTTB200_BC4S || FMT == CVTTB200_BC5S);
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
                int c[4];
#pragma unroll
                for (int ch = 0; ch < 4; ch++)
                {
                    c[ch] = (int)((w[k] >> (8 * ch)) & 0xffu);
                    if (isSigned)
                        c[ch] = ::max((int)(int8_t)c[ch], -127) + 127;
                }
                F4 p;
                p.x = (float)c[0]; p.y = (float)c[1]; p.z = (float)c[2]; p.w = (float)c[3];
                sPx[(q * 4 + k) * kS3TCThreads + tid] = p;
            }
        }
        S3TCLane<kS3TCThreads> L;
        L.px = sPx + tid;

        uint32_t w[4] = { 0, 0, 0, 0 };
        SegmentMax vote;
        if (FMT == CVTTB200_BC1)
            s3tc_pack_rgb<kS3TCThreads>(P, L, true, vote, w);
        else if (FMT == CVTTB200_BC2)
        {
            s3tc_pack_explicit_alpha<kS3TCThreads>(L, 3, w);
            s3tc_pack_rgb<kS3TCThreads>(P, L, false, vote, w + 2);
        }
        else if (FMT == CVTTB200_BC3)
        {
            s3tc_pack_interpolated_alpha<kS3TCThreads>(P, L, 3, false, w);
            s3tc_pack_rgb<kS3TCThreads>(P, L, false, vote, w + 2);
        }
        else if (FMT == CVTTB200_BC4U || FMT == CVTTB200_BC4S)
            s3tc_pack_interpolated_alpha<kS3TCThreads>(P, L, 0, isSigned, w);
        else
        {
            s3tc_pack_interpolated_alpha<kS3TCThreads>(P, L, 0, isSigned, w);
            s3tc_pack_interpolated_alpha<kS3TCThreads>(P, L, 1, isSigned, w + 2);
        }

        if (!active)
            return;
        if (FMT == CVTTB200_BC1 || FMT == CVTTB200_BC4U || FMT == CVTTB200_BC4S)
            reinterpret_cast<uint2 *>(out)[block] = make_uint2(w[0], w[1]);
        else
            reinterpret_cast<uint4 *>(out)[block] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

namespace cvttb200
{
    int launch_s3tc(int format, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const float *rcpN, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        S3TCParams P;
        s3tc_fill_params(P, options, rcpN);
        const unsigned grid = (unsigned)((nBlocks + kS3TCThreads - 1) / kS3TCThreads);
        const uint4 *in = (const uint4 *)dIn;
        uint32_t *out = (uint32_t *)dOut;
        switch (format)
        {
        case CVTTB200_BC1: s3tc_encode_kernel<CVTTB200_BC1><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        case CVTTB200_BC2: s3tc_encode_kernel<CVTTB200_BC2><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        case CVTTB200_BC3: s3tc_encode_kernel<CVTTB200_BC3><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        case CVTTB200_BC4U: s3tc_encode_kernel<CVTTB200_BC4U><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        case CVTTB200_BC4S: s3tc_encode_kernel<CVTTB200_BC4S><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        case CVTTB200_BC5U: s3tc_encode_kernel<CVTTB200_BC5U><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        default: s3tc_encode_kernel<CVTTB200_BC5S><<<grid, kS3TCThreads, 0, stream>>>(P, in, out, (uint32_t)nBlocks); break;
        }
        g_launches++;
        CVTT_CUDA(cudaGetLastError());
        return CVTTB200_OK;
    }
}
