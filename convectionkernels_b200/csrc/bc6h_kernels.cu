// BC6H translation unit of libcvtt_b200.so: encode kernels (signed / unsigned, slow / fast indexing), launch, set-up.
#include <cuda_runtime.h>

#include <stdlib.h>

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

    // PixelBlockF16 = int16 [16][4] (128 B, alpha ignored), read with eight 128-bit loads per thread and converted once into
    // [word][thread] planes in shared memory; an inactive thread holds a block of zeros
    template<bool SIGNED>
    __device__ __forceinline__ void bc6h_load_block(const BC6HParams &P, const BC6HLane<kBC6HThreads> &L, const uint4 *__restrict__ in, uint32_t block, bool active)
    {
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
    }

    __device__ __forceinline__ BC6HLane<kBC6HThreads> bc6h_lane(unsigned char *smem, uint32_t tid)
    {
        float *sPw = reinterpret_cast<float *>(smem);
        uint32_t *sRaw = reinterpret_cast<uint32_t *>(sPw + 48 * kBC6HThreads);
        BC6HLane<kBC6HThreads> L;
        L.pw = sPw + tid;
        L.raw = sRaw + tid;
        L.tab = sRaw + 32 * kBC6HThreads + tid;           // only touched with slow indexing
        return L;
    }

    // One thread per block, warp = 4 reference groups; 416 B of shared memory per thread with the interpolator table, so that
    // four CTAs = 16 warps fit an SM.
    template<bool SIGNED, bool FAST>
    __global__ void __launch_bounds__(kBC6HThreads, 4)
    bc6h_encode_kernel(const __grid_constant__ BC6HParams P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nBlocks)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        const uint32_t tid = threadIdx.x;
        const uint32_t block = blockIdx.x * kBC6HThreads + tid;
        const bool active = block < nBlocks;
        const BC6HLane<kBC6HThreads> L = bc6h_lane(smem, tid);
        bc6h_load_block<SIGNED>(P, L, in, block, active);

        SegmentVote vote;
        vote.segMask = 0xffu << (tid & 24);
        uint32_t o[4];
        bc6h_encode_block<SIGNED, FAST, kBC6HThreads>(P, c_bc6hTables, L, vote, o);
        if (active)
            out[block] = make_uint4(o[0], o[1], o[2], o[3]);
    }

    // Small-call launch, first kernel (bc6h_core.cuh, "The search as a numbered sequence of CALLS"): blockIdx.y = a range of
    // callsPerSlice calls of the search, run from a fresh best for the CTA's blocks; only the error history leaves the kernel:
    // history[call * nBlocks + block].  (Ranges that start from the error of the four one-subset calls, done by a launch of
    // their own first, are exact too -- bc6h_search_calls, tests/test_bc6h_host.py -- and were measured: BC6HU 4 % slower at
    // 512-8192 blocks, BC6HS 12 % faster at 4096-16 384; what prunes in the sequential search is the best PARTITION found so
    // far, not the one-subset modes.)
    template<bool SIGNED, bool FAST>
    __global__ void __launch_bounds__(kBC6HThreads, 4)
    bc6h_search_kernel(const __grid_constant__ BC6HParams P, const uint4 *__restrict__ in, uint32_t nBlocks, int callsPerSlice, float *__restrict__ history)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        const uint32_t tid = threadIdx.x;
        const uint32_t block = blockIdx.x * kBC6HThreads + tid;
        const bool active = block < nBlocks;
        const BC6HLane<kBC6HThreads> L = bc6h_lane(smem, tid);
        bc6h_load_block<SIGNED>(P, L, in, block, active);

        SegmentVote vote;
        vote.segMask = 0xffu << (tid & 24);
        const int callBegin = (int)blockIdx.y * callsPerSlice, callEnd = ::min((int)kBC6HCalls, callBegin + callsPerSlice);
        bc6h_search_calls<SIGNED, FAST, kBC6HThreads>(P, c_bc6hTables, L, vote, callBegin, callEnd, FLT_MAX, history + (size_t)callBegin * nBlocks + (active ? block : 0), nBlocks, active);
    }

    // Small-call launch, second kernel: warp = (group, k).  Its first eight lanes hold the group's blocks; they re-run the k-th
    // distinct winner call of the group from the lanes' true entry errors, and the lanes whose winner it is pack their block.
    template<bool SIGNED, bool FAST>
    __global__ void __launch_bounds__(kBC6HThreads, 4)
    bc6h_resolve_kernel(const __grid_constant__ BC6HParams P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nBlocks, const float *__restrict__ history)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        const uint32_t tid = threadIdx.x, lane = tid & 31;
        const uint32_t warp = blockIdx.x * (kBC6HThreads / 32) + (tid >> 5);
        const uint32_t group = warp >> 3, k = warp & 7;
        const uint32_t block = group * 8 + lane;
        const bool active = lane < 8 && block < nBlocks;
        const BC6HLane<kBC6HThreads> L = bc6h_lane(smem, tid);
        bc6h_load_block<SIGNED>(P, L, in, block, active);

        const float *hist = history + (active ? block : 0);
        int winner = -1;
        if (active)
            bc6h_history(hist, nBlocks, kBC6HCalls, winner);
        // k-th distinct winner of the group, in lane order
        int call = -1, count = 0;
        int seen[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const int w = __shfl_sync(0xffffffffu, winner, j);
            bool dup = w < 0;
#pragma unroll
            for (int i = 0; i < 8; i++)
                dup = dup || (i < count && seen[i] == w);
            if (!dup)
            {
                if (count == (int)k)
                    call = w;
                seen[count++] = w;
            }
        }
        BC6HBest best;
        bc6h_best_reset(best);
        if (call < 0)
        {
            // nothing to re-run for this warp (warp-uniform); a block that never committed leaves as the reference's untouched state
            if (k == 0 && active && winner < 0)
            {
                uint32_t o[4];
                bc6h_pack_block<SIGNED, FAST, kBC6HThreads>(P, c_bc6hTables, L, best, o);
                out[block] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            return;
        }
        int unused;
        // the lanes without a block can never be better than their best: they add no work to the warp's votes
        best.error = active ? bc6h_history(hist, nBlocks, call, unused) : -FLT_MAX;

        SegmentVote vote;
        vote.segMask = 0xffu << (tid & 24);
        bc6h_run_call<SIGNED, FAST, kBC6HThreads>(P, c_bc6hTables, L, vote, call, best);
        if (active && (winner == call || (k == 0 && winner < 0)))
        {
            if (winner < 0)
                bc6h_best_reset(best);
            uint32_t o[4];
            bc6h_pack_block<SIGNED, FAST, kBC6HThreads>(P, c_bc6hTables, L, best, o);
            out[block] = make_uint4(o[0], o[1], o[2], o[3]);
        }
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
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_search_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_search_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_search_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_search_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_resolve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_resolve_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_resolve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_resolve_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        return CVTTB200_OK;
    }

    template<bool SIGNED, bool FAST>
    static int launch_bc6h_split(DeviceContext &ctx, const BC6HParams &P, const uint4 *in, uint4 *out, size_t nBlocks, int callsPerSlice, cudaStream_t stream)
    {
        const size_t smem = FAST ? kBC6HSmemBytesFast : kBC6HSmemBytesSlow;
        float *dHistory = nullptr;
        const int rc = pool_alloc(ctx, (void **)&dHistory, (size_t)kBC6HCalls * nBlocks * sizeof(float), stream);
        if (rc != CVTTB200_OK)
            return rc;
        const unsigned ctas = (unsigned)((nBlocks + kBC6HThreads - 1) / kBC6HThreads);
        const unsigned slices = (unsigned)((kBC6HCalls + callsPerSlice - 1) / callsPerSlice);
        bc6h_search_kernel<SIGNED, FAST><<<dim3(ctas, slices), kBC6HThreads, smem, stream>>>(P, in, (uint32_t)nBlocks, callsPerSlice, dHistory);
        // one warp per (group, distinct winner call): eight per group, the surplus ones leave at once
        const unsigned resolveCtas = (unsigned)((nBlocks / 8 * 8 + kBC6HThreads / 32 - 1) / (kBC6HThreads / 32));
        bc6h_resolve_kernel<SIGNED, FAST><<<resolveCtas, kBC6HThreads, smem, stream>>>(P, in, out, (uint32_t)nBlocks, dHistory);
        g_launches += 2;
        CVTT_CUDA(cudaFreeAsync(dHistory, stream));
        CVTT_CUDA(cudaGetLastError());
        return CVTTB200_OK;
    }

    int launch_bc6h(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, bool isSigned, const float *rcpN, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        BC6HParams P;
        bc6h_fill_params(P, options, rcpN);
        const unsigned grid = (unsigned)((nBlocks + kBC6HThreads - 1) / kBC6HThreads);
        const bool fast = (options.flags & kFlag_BC6H_FastIndexing) != 0;
        const uint4 *in = (const uint4 *)dIn;
        uint4 *out = (uint4 *)dOut;

        // Small calls (the reference's own call is 8 blocks).  One warp walks the 196 calls of the search in 10.8 ms whatever
        // the size of the call.  When the call's CTAs leave room on the device (four resident CTAs per SM), the calls are dealt
        // out to that many times as many CTAs, which only record error histories, and a second kernel re-runs each block's
        // winner call (exactness: bc6h_core.cuh, "The search as a numbered sequence of CALLS").  8 blocks: 196 CTAs of one
        // call each, 0.43 ms; 4096 blocks: 18 ranges, 3.9 ms against 20.8; the gain ends at 37 888 blocks.  A range starts
        // without an error to prune against, so the split costs work: it is a latency device, not a throughput one.
        static const long splitOverride = getenv("CVTTB200_BC6H_SPLIT") ? atol(getenv("CVTTB200_BC6H_SPLIT")) : -1;     // A/B: 0 = never, n = n calls per slice
        int callsPerSlice = 0;
        if (splitOverride > 0)
            callsPerSlice = (int)splitOverride;
        else if (splitOverride < 0 && nBlocks > 0)
        {
            const size_t resident = (size_t)ctx.numSMs * 4;
            const size_t slices = std::min<size_t>(resident / grid, (size_t)kBC6HCalls);
            if (slices >= 2)
                callsPerSlice = (int)((kBC6HCalls + slices - 1) / slices);
        }
        if (callsPerSlice > 0)
        {
            if (isSigned)
                return fast ? launch_bc6h_split<true, true>(ctx, P, in, out, nBlocks, callsPerSlice, stream) : launch_bc6h_split<true, false>(ctx, P, in, out, nBlocks, callsPerSlice, stream);
            return fast ? launch_bc6h_split<false, true>(ctx, P, in, out, nBlocks, callsPerSlice, stream) : launch_bc6h_split<false, false>(ctx, P, in, out, nBlocks, callsPerSlice, stream);
        }

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
