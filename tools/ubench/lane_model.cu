// Lane-model A/B for the BC7 search: the same work -- BC7 mode 3, all 64 partitions, four seed points,
// two refine rounds, with exactly the product's arithmetic (bc7_core.cuh) -- mapped onto the warp in two ways:
//   block_lanes      thread = block (the product's model): a warp holds 32 blocks, every lane walks the 64 partitions of its
//                    own block; pixel subsets are warp-uniform, all loops have uniform trip counts
//   candidate_lanes  lanes = candidates: a warp works on ONE block at a time, lane l searches partitions l and l + 32 of that
//                    block and the warp reduces (error, partition) with shuffles; 32 blocks per warp, one after the other
// Both produce, per block, the winning (error, partition, endpoints of both subsets); the program checks that they are
// identical and prints the time per block of each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I convectionkernels_b200/csrc \
//        -o lane_model tools/ubench/lane_model.cu convectionkernels_b200/csrc/bc7_host.cpp
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "bc7_host.h"

using namespace cvttb200;

namespace
{
    constexpr int kThreads = 384;
    constexpr size_t kSmemBytes = (size_t)kThreads * 16 * (sizeof(uint32_t) + 2 * sizeof(F4));
    __constant__ BC7PackTables c_tables;

    struct Result
    {
        float error;
        uint32_t partition, e00, e01, e10, e11;
    };

    // one (block, partition): both subsets of mode 3; the pixels come from column `rawCol` of the CTA's raw array
    __device__ __forceinline__ void search_partition(const BC7Params &P, const uint32_t *rawCol, F4 *gv, F4 *gw, int partition, Result &r)
    {
        BC7Lane<kThreads> L;
        L.raw = rawCol;
        L.gv = gv;
        L.gw = gw;
        float total = 0.0f;
        uint32_t ep[2][2];
#pragma unroll 1
        for (int s = 0; s < 2; s++)
        {
            const uint32_t m2 = c_tables.partitionMask2[partition];
            const uint32_t mask = s ? m2 : (~m2 & 0xffffu);
            const int n = __popc(mask);
            float sumV[4], accA;
            bc7_gather<kThreads>(L, mask, 0, P.w, sumV, accA);
            const float staticAlphaError = fmul(accA, P.wSq[3]);
            float baseRGB[3], offsRGB[3], baseRGBA[4], offsRGBA[4];
            bc7_shape_fits<kThreads>(P, gw, n, true, false, false, true, false, true, false, baseRGB, offsRGB, baseRGBA, offsRGBA);
            BC7ShapeBest best;
            bc7_shape_trials<3, true, kThreads>(P, gv, gw, n, 4, baseRGB, offsRGB, sumV, staticAlphaError, best);
            total = fadd(total, best.err);
            ep[s][0] = best.e0;
            ep[s][1] = best.e1;
        }
        if (total < r.error || (total == r.error && (uint32_t)partition < r.partition))
        {
            r.error = total;
            r.partition = (uint32_t)partition;
            r.e00 = ep[0][0]; r.e01 = ep[0][1]; r.e10 = ep[1][0]; r.e11 = ep[1][1];
        }
    }

    __device__ __forceinline__ void load_raw(const uint4 *in, uint32_t block, uint32_t nBlocks, uint32_t *sRaw, uint32_t tid)
    {
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            uint4 v = make_uint4(0xff000000u, 0xff000000u, 0xff000000u, 0xff000000u);
            if (block < nBlocks)
                v = __ldg(in + (size_t)block * 4 + q);
            sRaw[(q * 4 + 0) * kThreads + tid] = v.x;
            sRaw[(q * 4 + 1) * kThreads + tid] = v.y;
            sRaw[(q * 4 + 2) * kThreads + tid] = v.z;
            sRaw[(q * 4 + 3) * kThreads + tid] = v.w;
        }
        __syncwarp();
    }

    __global__ void __launch_bounds__(kThreads, 1) block_lanes(const __grid_constant__ BC7Params P, const uint4 *in, Result *out, uint32_t nBlocks)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        F4 *sGv = reinterpret_cast<F4 *>(smem), *sGw = sGv + 16 * kThreads;
        uint32_t *sRaw = reinterpret_cast<uint32_t *>(sGw + 16 * kThreads);
        const uint32_t tid = threadIdx.x, block = blockIdx.x * kThreads + tid;
        load_raw(in, block, nBlocks, sRaw, tid);
        Result r;
        r.error = FLT_MAX;
        r.partition = 0xffffffffu;
        r.e00 = r.e01 = r.e10 = r.e11 = 0;
#pragma unroll 1
        for (int p = 0; p < 64; p++)
            search_partition(P, sRaw + tid, sGv + tid, sGw + tid, p, r);
        if (block < nBlocks)
            out[block] = r;
    }

    __global__ void __launch_bounds__(kThreads, 1) candidate_lanes(const __grid_constant__ BC7Params P, const uint4 *in, Result *out, uint32_t nBlocks)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        F4 *sGv = reinterpret_cast<F4 *>(smem), *sGw = sGv + 16 * kThreads;
        uint32_t *sRaw = reinterpret_cast<uint32_t *>(sGw + 16 * kThreads);
        const uint32_t tid = threadIdx.x, lane = tid & 31, warpBase = tid & ~31u, block = blockIdx.x * kThreads + tid;
        load_raw(in, block, nBlocks, sRaw, tid);
#pragma unroll 1
        for (uint32_t b = 0; b < 32; b++)           // the warp's 32 blocks, one after the other
        {
            Result r;
            r.error = FLT_MAX;
            r.partition = 0xffffffffu;
            r.e00 = r.e01 = r.e10 = r.e11 = 0;
#pragma unroll 1
            for (int half = 0; half < 2; half++)
                search_partition(P, sRaw + warpBase + b, sGv + tid, sGw + tid, (int)lane + 32 * half, r);
            // arg-min over the lanes by (error, partition)
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1)
            {
                Result o;
                o.error = __shfl_xor_sync(0xffffffffu, r.error, step);
                o.partition = __shfl_xor_sync(0xffffffffu, r.partition, step);
                o.e00 = __shfl_xor_sync(0xffffffffu, r.e00, step);
                o.e01 = __shfl_xor_sync(0xffffffffu, r.e01, step);
                o.e10 = __shfl_xor_sync(0xffffffffu, r.e10, step);
                o.e11 = __shfl_xor_sync(0xffffffffu, r.e11, step);
                if (o.error < r.error || (o.error == r.error && o.partition < r.partition))
                    r = o;
            }
            const uint32_t ob = blockIdx.x * kThreads + warpBase + b;
            if (lane == 0 && ob < nBlocks)
                out[ob] = r;
        }
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv)
{
    const uint32_t nBlocks = argc > 1 ? (uint32_t)atoi(argv[1]) : 148u * kThreads * 2;
    // synthetic content: gradients + noise + an edge per block, opaque
    std::vector<uint8_t> h((size_t)nBlocks * 64);
    uint32_t s = 12345;
    for (uint32_t b = 0; b < nBlocks; b++)
        for (int px = 0; px < 16; px++)
        {
            const int x = px & 3, y = px >> 2;
            for (int ch = 0; ch < 3; ch++)
            {
                s = s * 1664525u + 1013904223u;
                int v = (int)((b * 7 + ch * 40) & 255) / 2 + x * (8 + ch * 3) + y * (5 + (b & 7)) + (int)((s >> 24) % ((b >> 3) % 48 + 1));
                if (((x + y * (b & 3)) & 7) > 4)
                    v = 255 - v;
                h[((size_t)b * 16 + px) * 4 + ch] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
            h[((size_t)b * 16 + px) * 4 + 3] = 255;
        }
    OptionsPOD o;
    o.flags = kFlag_BC7_FastIndexing | kFlag_S3TC_Paranoid;
    o.threshold = 0.5f;
    o.redWeight = 0.2125f / 0.7154f; o.greenWeight = 1.0f; o.blueWeight = 0.0721f / 0.7154f; o.alphaWeight = 1.0f;
    o.refineRoundsBC7 = 2; o.refineRoundsBC6H = 3; o.refineRoundsIIC = 8; o.refineRoundsS3TC = 2; o.seedPoints = 4;
    BC7PlanPOD plan;
    bc7_plan_from_quality(plan, 100);
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = n ? 1.0f / (float)n : 0.0f;          // any table: both kernels use the same
    BC7Params P;
    bc7_fill_params(P, o, plan, rcpN);
    P.cmds = nullptr;
    CK(cudaMemcpyToSymbol(c_tables, &bc7_pack_tables(), sizeof(BC7PackTables)));
    CK(cudaFuncSetAttribute(block_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    CK(cudaFuncSetAttribute(candidate_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    uint4 *dIn;
    Result *dA, *dB;
    CK(cudaMalloc(&dIn, h.size()));
    CK(cudaMalloc(&dA, nBlocks * sizeof(Result)));
    CK(cudaMalloc(&dB, nBlocks * sizeof(Result)));
    CK(cudaMemcpy(dIn, h.data(), h.size(), cudaMemcpyHostToDevice));
    const unsigned grid = (nBlocks + kThreads - 1) / kThreads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float msA = 0, msB = 0;
    for (int rep = 0; rep < 3; rep++)
    {
        cudaEventRecord(e0);
        block_lanes<<<grid, kThreads, kSmemBytes>>>(P, dIn, dA, nBlocks);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&msA, e0, e1);
        cudaEventRecord(e0);
        candidate_lanes<<<grid, kThreads, kSmemBytes>>>(P, dIn, dB, nBlocks);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&msB, e0, e1);
    }
    std::vector<Result> a(nBlocks), b(nBlocks);
    CK(cudaMemcpy(a.data(), dA, nBlocks * sizeof(Result), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), dB, nBlocks * sizeof(Result), cudaMemcpyDeviceToHost));
    uint32_t differing = 0;
    for (uint32_t i = 0; i < nBlocks; i++)
        differing += memcmp(&a[i], &b[i], sizeof(Result)) != 0;
    printf("{\"workload\": \"BC7 mode 3, 64 partitions x 2 subsets, 4 seed points, 2 refine rounds\", \"blocks\": %u, \"block_lanes_ms\": %.3f, \"candidate_lanes_ms\": %.3f, "
           "\"block_lanes_mblocks_per_s\": %.3f, \"candidate_lanes_mblocks_per_s\": %.3f, \"candidate_over_block_time\": %.3f, \"results_differing\": %u}\n",
           nBlocks, msA, msB, nBlocks / msA / 1e3, nBlocks / msB / 1e3, msB / msA, differing);
    return differing ? 2 : 0;
}
