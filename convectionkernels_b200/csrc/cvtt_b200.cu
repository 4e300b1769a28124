// libcvtt_b200.so: CUDA kernels (sm_100a) + the C ABI declared in include/cvtt_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false ... (convectionkernels_b200/build.py).
// -fmad=false is part of the numerical contract: the reference's fp32 expressions must round after every
// operation.
#include <cuda_runtime.h>
#include <xmmintrin.h>

#include <algorithm>
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

using namespace cvttb200;

static_assert(sizeof(cvttb200_options) == 44 && sizeof(OptionsPOD) == 44, "cvtt::Options layout");
static_assert(sizeof(cvttb200_bc7_plan) == 808 && sizeof(BC7PlanPOD) == 808, "cvtt::BC7EncodingPlan layout");
static_assert(sizeof(cvttb200_bc7_fine_tuning) == 285 && sizeof(BC7FineTuningPOD) == 285, "cvtt::BC7FineTuningParams layout");

// =========================================================================================================
// Kernels

namespace
{
    constexpr int kBC7Threads = 384;     // 12 warps = 48 reference groups per CTA, one CTA per SM
    constexpr int kBC7CtasPerSM = 1;
    // per thread: 16 packed pixels + 16 gathered biased pixels + 16 gathered pre-weighted pixels
    constexpr size_t kBC7SmemBytes = (size_t)kBC7Threads * 16 * (sizeof(uint32_t) + 2 * sizeof(F4));

    __constant__ BC7PackTables c_bc7PackTables;

    // Pre-pass: sorts the reference groups (8 consecutive blocks = one reference call) into three classes by the two
    // group-wide votes of BC7Computer::TrySinglePlane (BC67.cpp:1069-1072), so that every warp of the encode kernel
    // holds four groups that walk the same set of modes.  Pure scheduling: the encode kernel recomputes the votes.
    //   class 0: opaque group (RGB modes, no 4-channel fits, mode 7 only if the plan asks for it on RGB)
    //   class 1: some block has alpha and some block is (nearly) opaque: every mode runs
    //   class 2: every block has alpha <= 250 somewhere: RGB modes 0-3 are off
    // lists[c * nGroups + i] = i-th group of class c (order within a class is not deterministic and does not matter).
    __global__ void __launch_bounds__(256)
    bc7_classify_kernel(const uint4 *__restrict__ in, uint32_t nBlocks, uint32_t nGroups, uint32_t *__restrict__ counts, uint32_t *__restrict__ lists)
    {
        const uint32_t block = blockIdx.x * blockDim.x + threadIdx.x;
        const bool active = block < nBlocks;
        uint32_t minAlpha = 255;
        if (active)
        {
            const uint4 *src = in + (size_t)block * 4;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const uint4 v = __ldg(src + q);
                minAlpha = min(minAlpha, min(min(v.x >> 24, v.y >> 24), min(v.z >> 24, v.w >> 24)));
            }
        }
        const uint32_t segMask = 0xffu << (threadIdx.x & 24);
        const bool anyAlpha = (__ballot_sync(0xffffffffu, active && minAlpha < 255) & segMask) != 0;
        const bool allowRGB = (__ballot_sync(0xffffffffu, active && minAlpha > 250) & segMask) != 0;
        if (active && (threadIdx.x & 7) == 0)
        {
            const int cls = !anyAlpha ? 0 : (allowRGB ? 1 : 2);
            const uint32_t pos = atomicAdd(counts + cls, 1u);
            lists[(size_t)cls * nGroups + pos] = block >> 3;
        }
    }

    // cvttb200_selftest: f2_div against the compiler's IEEE division
    __device__ __forceinline__ uint32_t selftest_hash(uint64_t x)
    {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return (uint32_t)x;
    }

    __device__ __forceinline__ float selftest_operand(uint32_t h)
    {
        // sign | exponent in [127 - 40, 127 + 40] | 23 random mantissa bits
        const uint32_t e = 127u - 40u + (h >> 23) % 81u;
        return __uint_as_float((h & 0x80000000u) | (e << 23) | (h & 0x007fffffu));
    }

    __global__ void selftest_div_kernel(uint64_t samples, uint64_t seed, unsigned long long *mismatches)
    {
        unsigned long long bad = 0;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < samples; i += (uint64_t)gridDim.x * blockDim.x)
        {
            const uint32_t h0 = selftest_hash(seed + 4 * i), h1 = selftest_hash(seed + 4 * i + 1), h2 = selftest_hash(seed + 4 * i + 2), h3 = selftest_hash(seed + 4 * i + 3);
            f2 a = f2_make(selftest_operand(h0), selftest_operand(h1));
            const f2 b = f2_make(selftest_operand(h2), selftest_operand(h3));
            if ((h0 & 0xff) == 0)
                a.x = 0.0f;
            if ((h1 & 0xff) == 1)       // small integers over small integers, the shape of maxV / lenSq
            {
                a.y = (float)(1 + (h1 >> 8) % 15);
            }
            const f2 q = f2_div(a, b);
            const float wx = __fdiv_rn(a.x, b.x), wy = __fdiv_rn(a.y, b.y);
            if (__float_as_uint(q.x) != __float_as_uint(wx) && !(q.x == 0.0f && wx == 0.0f))
                bad++;
            if (__float_as_uint(q.y) != __float_as_uint(wy) && !(q.y == 0.0f && wy == 0.0f))
                bad++;
        }
        if (bad)
            atomicAdd(mismatches, bad);
    }

    // The reference's AnySet / AllSet over the 8 lanes of one call (ParallelMath.h:1260-1278): ballots restricted to
    // the lane's 8-lane segment.  Every lane of the warp executes every vote (control flow around votes is uniform).
    struct SegmentVote
    {
        uint32_t segMask;
        __device__ __forceinline__ bool any(bool x) const { return (__ballot_sync(0xffffffffu, x) & segMask) != 0; }
        __device__ __forceinline__ bool all(bool x) const { return (__ballot_sync(0xffffffffu, x) & segMask) == segMask; }
        __device__ __forceinline__ bool warp_any(bool x) const { return __any_sync(0xffffffffu, x) != 0; }
    };

    // One thread per block, warp = 4 reference groups of one class; see cvtt_common.cuh / bc7_core.cuh.
    //  * input: each thread reads its own 64-byte PixelBlockU8 with four 128-bit loads (512 B contiguous per group) and
    //    keeps it packed in shared memory, laid out [pixel][thread] (conflict-free)
    //  * per pixel subset the search gathers the subset's pixels once into two [index][thread] arrays of fp32x4 (biased
    //    value, pre-weighted value); every trial then streams them with 128-bit conflict-free loads
    //  * output: one 128-bit store per thread
    template<bool FAST, bool PUNCH>
    __global__ void __launch_bounds__(kBC7Threads, kBC7CtasPerSM)
    bc7_encode_kernel(const __grid_constant__ BC7Params P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nGroups,
                      const uint32_t *__restrict__ counts, const uint32_t *__restrict__ lists)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        F4 *sGv = reinterpret_cast<F4 *>(smem);
        F4 *sGw = sGv + 16 * kBC7Threads;
        uint32_t *sRaw = reinterpret_cast<uint32_t *>(sGw + 16 * kBC7Threads);

        const uint32_t tid = threadIdx.x, lane = tid & 31;

        // warp -> (class, four groups of that class); the expensive classes go first so that the tail of the launch is
        // filled by the cheap opaque warps
        const uint32_t n0 = counts[0], n1 = counts[1], n2 = counts[2];
        const uint32_t w1 = (n1 + 3) >> 2, w2 = (n2 + 3) >> 2, w0 = (n0 + 3) >> 2;
        uint32_t warp = blockIdx.x * (kBC7Threads / 32) + (tid >> 5);
        uint32_t cls, clsCount;
        if (warp < w1) { cls = 1; clsCount = n1; }
        else if (warp < w1 + w2) { cls = 2; clsCount = n2; warp -= w1; }
        else if (warp < w1 + w2 + w0) { cls = 0; clsCount = n0; warp -= w1 + w2; }
        else { cls = 0; clsCount = 0; }      // surplus warp of the last CTA: no work, but it keeps the CTA's barriers company
        const uint32_t slot = warp * 4 + (lane >> 3);
        const bool active = slot < clsCount;
        const uint32_t block = active ? lists[(size_t)cls * nGroups + slot] * 8 + (lane & 7) : 0;

        BC7Lane<kBC7Threads> L;
        L.raw = sRaw + tid;
        L.gv = sGv + tid;
        L.gw = sGw + tid;

        uint32_t minAlpha = 255, maxAlpha = 0;
        bool isPunchThrough = true;
        if (active)
        {
            const uint4 *src = in + (size_t)block * 4;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const uint4 v = __ldg(src + q);
                minAlpha = min(minAlpha, min(min(v.x >> 24, v.y >> 24), min(v.z >> 24, v.w >> 24)));
                maxAlpha = max(maxAlpha, max(max(v.x >> 24, v.y >> 24), max(v.z >> 24, v.w >> 24)));
                const uint32_t a[4] = { v.x >> 24, v.y >> 24, v.z >> 24, v.w >> 24 };
#pragma unroll
                for (int k = 0; k < 4; k++)
                    isPunchThrough = isPunchThrough && (a[k] == 0 || a[k] == 255);
                sRaw[(q * 4 + 0) * kBC7Threads + tid] = v.x;
                sRaw[(q * 4 + 1) * kBC7Threads + tid] = v.y;
                sRaw[(q * 4 + 2) * kBC7Threads + tid] = v.z;
                sRaw[(q * 4 + 3) * kBC7Threads + tid] = v.w;
            }
        }
        else
        {
#pragma unroll
            for (int px = 0; px < 16; px++)
                sRaw[px * kBC7Threads + tid] = 0xff000000u;
        }
        __syncwarp();

        // group votes (reference AnySet over the 8 lanes of one call, BC67.cpp:1069-1072) and warp-level skips
        const uint32_t segMask = 0xffu << (lane & 24);
        const uint32_t hasAlphaBallot = __ballot_sync(0xffffffffu, active && minAlpha < 255);
        const uint32_t allowRGBBallot = __ballot_sync(0xffffffffu, active && minAlpha > 250);
        BC7LaneFlags lf;
        lf.anyBlockHasAlpha = (hasAlphaBallot & segMask) != 0;
        lf.allowRGBModes = (allowRGBBallot & segMask) != 0;
        lf.blockHasNonMaxAlpha = minAlpha < 255;
        lf.blockHasNonZeroAlpha = maxAlpha > 0;
        lf.isPunchThrough = isPunchThrough;
        const bool usePCA4 = lf.anyBlockHasAlpha || !lf.allowRGBModes;
        const bool mode7 = lf.anyBlockHasAlpha || P.mode7RGBPartitionEnabled != 0;
        lf.warpAnyRGB = __any_sync(0xffffffffu, active && lf.allowRGBModes);
        lf.warpAnyPCA4 = __any_sync(0xffffffffu, active && usePCA4);
        lf.warpAnyExpand = true;
        lf.warpAnyMode7 = __any_sync(0xffffffffu, active && mode7);

        uint32_t o[4];
        if (PUNCH)
        {
            SegmentVote vote;
            vote.segMask = segMask;
            bc7_encode_block<FAST, kBC7Threads, true>(P, c_bc7PackTables, L, lf, vote, o);
        }
        else
        {
            BC7NoVote vote;
            bc7_encode_block<FAST, kBC7Threads, false>(P, c_bc7PackTables, L, lf, vote, o);
        }

        if (active)
            out[block] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// BC6H

namespace
{
    constexpr int kBC6HThreads = 128;
    // per thread: 48 + 48 floats, plus 32 words of raw pixels for the fast-indexing kernels
    constexpr size_t kBC6HSmemBytesSlow = (size_t)kBC6HThreads * 96 * 4, kBC6HSmemBytesFast = (size_t)kBC6HThreads * 128 * 4;

    __constant__ BC6HTables c_bc6hTables;

    // One thread per block, warp = 4 reference groups.  Input: PixelBlockF16 = int16 [16][4] (128 B, alpha ignored), read
    // with eight 128-bit loads per thread; converted once into [word][thread] planes in shared memory (384 B per thread, so
    // that four CTAs = 16 warps fit an SM).
    template<bool SIGNED, bool FAST>
    __global__ void __launch_bounds__(kBC6HThreads, 4)
    bc6h_encode_kernel(const __grid_constant__ BC6HParams P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nBlocks)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        float *sLin = reinterpret_cast<float *>(smem);
        float *sPw = sLin + 48 * kBC6HThreads;

        const uint32_t tid = threadIdx.x;
        const uint32_t block = blockIdx.x * kBC6HThreads + tid;
        const bool active = block < nBlocks;

        BC6HLane<kBC6HThreads, FAST> L;
        L.lin = sLin + tid;
        L.pw = sPw + tid;
        L.pix = reinterpret_cast<uint32_t *>(sPw + 48 * kBC6HThreads) + tid;
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

// ---------------------------------------------------------------------------------------------------------
// ETC1 / ETC2 / EAC

namespace
{
    constexpr int kETCThreads = 512;      // 16 warps = 64 reference groups per CTA, one CTA per SM, phases in lock-step
    constexpr int kETCCtasPerSM = 1;
    constexpr size_t kETCSmemBytes = (size_t)kETCThreads * 16 * sizeof(F4);

    __constant__ ETCTables c_etcTables;

    // group maximum (the reference's per-call maximum over its 8 lanes): butterfly over the lane's 8-lane segment
    struct SegmentMax
    {
        __device__ __forceinline__ int max(int v) const
        {
            v = ::max(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = ::max(v, __shfl_xor_sync(0xffffffffu, v, 2));
            v = ::max(v, __shfl_xor_sync(0xffffffffu, v, 4));
            return v;
        }
    };

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
                etc_alpha_encode_block(c_etcTables, alpha, false, false, a);
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
        const uint32_t block = blockIdx.x * blockDim.x + threadIdx.x;
        if (block >= nBlocks)
            return;
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
        etc_alpha_encode_block(c_etcTables, a, KIND != 0, KIND == 2, o);
        out[block] = make_uint2(etc_bswap(o[0]), etc_bswap(o[1]));
    }
}

// ---------------------------------------------------------------------------------------------------------
// BC1 - BC5

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

// ---------------------------------------------------------------------------------------------------------
// Image <-> block array (the step before / after the encode path; the reference's sample does it on the CPU,
// etc2packer/etc2packer.cpp:215-248 and :277-284).  Pure data movement: this is the HBM-bound part of the pipeline.

namespace
{
    // One thread per block row (4 pixels).  PIXEL_BYTES = 4 (RGBA8 -> PixelBlockU8) or 8 (RGBA16F -> PixelBlockF16).  Blocks are
    // laid out like the sample packer does: rows of ceil(width / 32) groups of 8 blocks; coordinates past the edge are clamped,
    // so the padding blocks of the last group of a row repeat the last column (they take part in the group semantics of the
    // encoders).  Thread t of a warp handles row (t & 3) of block (t >> 2): the warp's stores are one contiguous 512 B (1 KB)
    // run of the block array, its loads are four 128 B (256 B) row segments -- whole 32-byte sectors on both sides.
    template<int PIXEL_BYTES>
    __global__ void __launch_bounds__(256)
    tile_image_kernel(const unsigned char *__restrict__ image, int width, int height, size_t pitch, int blocksPerRow, uint32_t nBlocks, uint4 *__restrict__ blocks, int vectorOK)
    {
        const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t b = (uint32_t)(i >> 2);
        const int r = (int)(i & 3);
        if (b >= nBlocks)
            return;
        const int by = (int)(b / (uint32_t)blocksPerRow), bx = (int)(b % (uint32_t)blocksPerRow);
        const int x0 = bx * 4, y = min(by * 4 + r, height - 1);
        constexpr int kRowVec = PIXEL_BYTES / 4;      // uint4 per block row
        uint4 *dst = blocks + ((size_t)b * 4 + r) * kRowVec;
        const unsigned char *rowPtr = image + (size_t)y * pitch;
        if (vectorOK && x0 + 4 <= width)
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(rowPtr + (size_t)x0 * PIXEL_BYTES);
#pragma unroll
            for (int k = 0; k < kRowVec; k++)
                dst[k] = __ldg(src + k);
        }
        else
        {
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
    // sample's KTX writer).  One thread per 8 output bytes.
    __global__ void __launch_bounds__(256)
    untile_blocks_kernel(const uint2 *__restrict__ encoded, int blocksPerRow, int realBlocksPerRow, int wordsPerBlock, uint64_t nWords, uint2 *__restrict__ out)
    {
        const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= nWords)
            return;
        const uint64_t block = i / (uint64_t)wordsPerBlock, word = i % (uint64_t)wordsPerBlock;
        const uint64_t row = block / (uint64_t)realBlocksPerRow, col = block % (uint64_t)realBlocksPerRow;
        out[i] = __ldg(encoded + (row * (uint64_t)blocksPerRow + col) * (uint64_t)wordsPerBlock + word);
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

namespace
{
    thread_local std::string t_lastError;
    std::atomic<uint64_t> g_launches(0);

    struct PlanCacheEntry
    {
        BC7PlanPOD plan;
        uint32_t *dCmds;
    };

    struct DeviceContext
    {
        int device = -1;
        bool ready = false;
        int numSMs = 0;
        std::vector<PlanCacheEntry> plans;
        void *stageIn = nullptr, *stageOut = nullptr;
        size_t stageInBytes = 0, stageOutBytes = 0;
    };

    std::mutex g_mutex;
    std::vector<DeviceContext> g_contexts;
    float g_rcpN[17];
    bool g_rcpOverridden = false, g_rcpReady = false;

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

#define CVTT_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(e_, #call); } while (0)

    void host_rcp_table(float *t)
    {
        for (int n = 0; n < 17; n++)
            t[n] = _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    }

    // caller holds g_mutex
    int get_context(int device, DeviceContext **out)
    {
        if (!g_rcpReady)
        {
            host_rcp_table(g_rcpN);
            g_rcpReady = true;
        }
        for (size_t i = 0; i < g_contexts.size(); i++)
            if (g_contexts[i].device == device && g_contexts[i].ready)
            {
                *out = &g_contexts[i];
                return CVTTB200_OK;
            }

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

        int prev = 0;
        CVTT_CUDA(cudaGetDevice(&prev));
        CVTT_CUDA(cudaSetDevice(device));
        {
            // scratch of the BC7 / ETC launches comes from the stream-ordered pool; keep it cached between calls instead of
            // returning it to the driver at every synchronisation
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
            {
                uint64_t threshold = ~(uint64_t)0;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
            }
            cudaGetLastError();
        }
        CVTT_CUDA(cudaMemcpyToSymbol(c_bc7PackTables, &bc7_pack_tables(), sizeof(BC7PackTables)));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaMemcpyToSymbol(c_bc6hTables, &bc6h_tables(), sizeof(BC6HTables)));
        {
            DecodeTables dt;
            dt.bc7 = bc7_pack_tables();
            dt.bc6h = bc6h_tables();
            CVTT_CUDA(cudaMemcpyToSymbol(g_decodeTables, &dt, sizeof(dt)));
        }
        CVTT_CUDA(cudaMemcpyToSymbol(c_etcTables, &etc_tables(), sizeof(ETCTables)));
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
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesSlow));
        CVTT_CUDA(cudaFuncSetAttribute(bc6h_encode_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC6HSmemBytesFast));
        CVTT_CUDA(cudaDeviceSynchronize());
        CVTT_CUDA(cudaSetDevice(prev));

        g_contexts.emplace_back();
        g_contexts.back().numSMs = prop.multiProcessorCount;
        g_contexts.back().device = device;
        g_contexts.back().ready = true;
        *out = &g_contexts.back();
        return CVTTB200_OK;
    }

    // caller holds g_mutex and has made ctx.device current
    int get_plan_commands(DeviceContext &ctx, const BC7PlanPOD &plan, const uint32_t **dCmds)
    {
        for (size_t i = 0; i < ctx.plans.size(); i++)
            if (memcmp(&ctx.plans[i].plan, &plan, sizeof(plan)) == 0)
            {
                *dCmds = ctx.plans[i].dCmds;
                return CVTTB200_OK;
            }
        std::vector<uint32_t> cmds;
        const int slots = bc7_compile_plan(plan, cmds);
        if (slots > kBC7MaxSlots)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "BC7 plan needs more result slots than the kernel provides");
        if (ctx.plans.size() >= 16)
        {
            cudaFree(ctx.plans.front().dCmds);
            ctx.plans.erase(ctx.plans.begin());
        }
        PlanCacheEntry entry;
        entry.plan = plan;
        entry.dCmds = nullptr;
        CVTT_CUDA(cudaMalloc(&entry.dCmds, cmds.size() * sizeof(uint32_t)));
        CVTT_CUDA(cudaMemcpy(entry.dCmds, cmds.data(), cmds.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        ctx.plans.push_back(entry);
        *dCmds = entry.dCmds;
        return CVTTB200_OK;
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

    int ensure_stage(void **buf, size_t *have, size_t need)
    {
        if (*have >= need)
            return CVTTB200_OK;
        if (*buf)
            cudaFree(*buf);
        *buf = nullptr;
        *have = 0;
        CVTT_CUDA(cudaMalloc(buf, need));
        *have = need;
        return CVTTB200_OK;
    }

    int launch_bc6h(const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, bool isSigned, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        BC6HParams P;
        bc6h_fill_params(P, options, g_rcpN);
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

    int launch_s3tc(int format, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");
        S3TCParams P;
        s3tc_fill_params(P, options, g_rcpN);
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

    template<int KIND>
    int launch_etc_color(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const ETCParams &P, bool uniform, bool bt709, cudaStream_t stream)
    {
        // resident threads: the whole device, or fewer for small inputs
        const size_t maxCtas = (size_t)ctx.numSMs * kETCCtasPerSM;
        const unsigned grid = (unsigned)std::min(maxCtas, (nBlocks + kETCThreads - 1) / kETCThreads);
        const size_t threads = (size_t)grid * kETCThreads;
        void *dScratch = nullptr;
        CVTT_CUDA(cudaMallocAsync(&dScratch, etc_scratch_bytes(threads), stream));
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

    int launch_etc(DeviceContext &ctx, int format, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, cudaStream_t stream)
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
        etc_fill_params(P, options);
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

    int launch_bc7(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const BC7PlanPOD &plan, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");

        BC7Params P;
        bc7_fill_params(P, options, plan, g_rcpN);
        const uint32_t *dCmds = nullptr;
        int rc = get_plan_commands(ctx, plan, &dCmds);
        if (rc != CVTTB200_OK)
            return rc;
        P.cmds = dCmds;

        // stream-ordered scratch for the group classification: counts[4] then lists[3][nGroups]
        const uint32_t nGroups = (uint32_t)(nBlocks / 8);
        uint32_t *dScratch = nullptr;
        CVTT_CUDA(cudaMallocAsync((void **)&dScratch, (4 + 3 * (size_t)nGroups) * sizeof(uint32_t), stream));
        CVTT_CUDA(cudaMemsetAsync(dScratch, 0, 4 * sizeof(uint32_t), stream));
        bc7_classify_kernel<<<(unsigned)((nBlocks + 255) / 256), 256, 0, stream>>>((const uint4 *)dIn, (uint32_t)nBlocks, nGroups, dScratch, dScratch + 4);
        g_launches++;

        // at most three partially filled warps (one per class)
        const unsigned warps = nGroups / 4 + 3;
        const unsigned grid = (warps + kBC7Threads / 32 - 1) / (kBC7Threads / 32);
        const bool fast = (options.flags & kFlag_BC7_FastIndexing) != 0, punch = (options.flags & kFlag_BC7_RespectPunchThrough) != 0;
        if (fast && !punch)
            bc7_encode_kernel<true, false><<<grid, kBC7Threads, kBC7SmemBytes, stream>>>(P, (const uint4 *)dIn, (uint4 *)dOut, nGroups, dScratch, dScratch + 4);
        else if (!fast && !punch)
            bc7_encode_kernel<false, false><<<grid, kBC7Threads, kBC7SmemBytes, stream>>>(P, (const uint4 *)dIn, (uint4 *)dOut, nGroups, dScratch, dScratch + 4);
        else if (fast)
            bc7_encode_kernel<true, true><<<grid, kBC7Threads, kBC7SmemBytes, stream>>>(P, (const uint4 *)dIn, (uint4 *)dOut, nGroups, dScratch, dScratch + 4);
        else
            bc7_encode_kernel<false, true><<<grid, kBC7Threads, kBC7SmemBytes, stream>>>(P, (const uint4 *)dIn, (uint4 *)dOut, nGroups, dScratch, dScratch + 4);
        g_launches++;
        CVTT_CUDA(cudaFreeAsync(dScratch, stream));
        CVTT_CUDA(cudaGetLastError());
        return CVTTB200_OK;
    }
}


namespace
{
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
}

// =========================================================================================================
// C ABI

extern "C"
{

int cvttb200_init(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceContext *ctx = nullptr;
    return get_context(device, &ctx);
}

void cvttb200_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess)
    {
        cudaGetLastError();
        g_contexts.clear();
        return;
    }
    for (size_t i = 0; i < g_contexts.size(); i++)
    {
        DeviceContext &c = g_contexts[i];
        if (cudaSetDevice(c.device) != cudaSuccess)
            continue;
        for (size_t k = 0; k < c.plans.size(); k++)
            cudaFree(c.plans[k].dCmds);
        if (c.stageIn) cudaFree(c.stageIn);
        if (c.stageOut) cudaFree(c.stageOut);
    }
    g_contexts.clear();
    cudaSetDevice(prev);
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
    std::lock_guard<std::mutex> lock(g_mutex);
    int device = 0;
    {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return fail_cuda(e, "cudaGetDevice");
    }
    DeviceContext *ctx = nullptr;
    int rc = get_context(device, &ctx);
    if (rc != CVTTB200_OK)
        return rc;
    const int blocksPerRow = ((width + 31) / 32) * 8;
    const size_t nBlocks = cvttb200_tiled_block_count(width, height);
    if (nBlocks > 0xffffff00u)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "image too large for one call");
    const int vectorOK = ((uintptr_t)image % 16 == 0) && (rowPitchBytes % 16 == 0);
    const unsigned grid = (unsigned)((nBlocks * 4 + 255) / 256);
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
    std::lock_guard<std::mutex> lock(g_mutex);
    const int blocksPerRow = ((width + 31) / 32) * 8, realBlocksPerRow = (width + 3) / 4;
    const int wordsPerBlock = (int)(blockBytes / 8);
    const uint64_t nWords = (uint64_t)((height + 3) / 4) * (uint64_t)realBlocksPerRow * (uint64_t)wordsPerBlock;
    const unsigned grid = (unsigned)((nWords + 255) / 256);
    untile_blocks_kernel<<<grid, 256, 0, (cudaStream_t)streamPtr>>>((const uint2 *)encoded, blocksPerRow, realBlocksPerRow, wordsPerBlock, nWords, (uint2 *)out);
    g_launches++;
    CVTT_CUDA(cudaGetLastError());
    return CVTTB200_OK;
}

int cvttb200_selftest(uint64_t samples, uint64_t seed, uint64_t *mismatches)
{
    if (!mismatches)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(g_mutex);
    int device = 0;
    {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return fail_cuda(e, "cudaGetDevice");
    }
    DeviceContext *ctx = nullptr;
    int rc = get_context(device, &ctx);
    if (rc != CVTTB200_OK)
        return rc;
    unsigned long long *dBad = nullptr;
    CVTT_CUDA(cudaMalloc((void **)&dBad, sizeof(unsigned long long)));
    CVTT_CUDA(cudaMemset(dBad, 0, sizeof(unsigned long long)));
    selftest_div_kernel<<<148 * 8, 256>>>(samples, seed, dBad);
    g_launches++;
    unsigned long long bad = 0;
    cudaError_t e = cudaMemcpy(&bad, dBad, sizeof(bad), cudaMemcpyDeviceToHost);
    cudaFree(dBad);
    if (e != cudaSuccess)
        return fail_cuda(e, "selftest_div_kernel");
    *mismatches = bad;
    return CVTTB200_OK;
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

int cvttb200_encode(int format, const void *blocks, size_t nBlocks, void *out, const cvttb200_options *options, const cvttb200_bc7_plan *plan, void *streamPtr)
{
    if (!blocks || !out || !options)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "null argument");
    if (nBlocks % 8 != 0)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "nBlocks must be a multiple of 8 (cvtt::NumParallelBlocks)");
    const size_t inBytes = cvttb200_input_block_bytes(format), outBytes = cvttb200_output_block_bytes(format);
    if (!inBytes)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "unknown format");
    if (format == CVTTB200_BC7 && !plan)
        return fail(CVTTB200_ERR_BAD_ARGUMENT, "CVTTB200_BC7 needs an encoding plan");
    if (nBlocks == 0)
        return CVTTB200_OK;

    std::lock_guard<std::mutex> lock(g_mutex);

    int device = 0;
    {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return fail_cuda(e, "cudaGetDevice");
    }
    DeviceContext *ctx = nullptr;
    int rc = get_context(device, &ctx);
    if (rc != CVTTB200_OK)
        return rc;

    cudaStream_t stream = (cudaStream_t)streamPtr;
    const bool inOnDevice = is_device_pointer(blocks), outOnDevice = is_device_pointer(out);
    const void *dIn = blocks;
    void *dOut = out;
    if (!inOnDevice)
    {
        rc = ensure_stage(&ctx->stageIn, &ctx->stageInBytes, nBlocks * inBytes);
        if (rc != CVTTB200_OK)
            return rc;
        CVTT_CUDA(cudaMemcpyAsync(ctx->stageIn, blocks, nBlocks * inBytes, cudaMemcpyHostToDevice, stream));
        dIn = ctx->stageIn;
    }
    if (!outOnDevice)
    {
        rc = ensure_stage(&ctx->stageOut, &ctx->stageOutBytes, nBlocks * outBytes);
        if (rc != CVTTB200_OK)
            return rc;
        dOut = ctx->stageOut;
    }

    OptionsPOD opt;
    memcpy(&opt, options, sizeof(opt));
    if (format == CVTTB200_BC7)
    {
        BC7PlanPOD planPOD;
        memcpy(&planPOD, plan, sizeof(planPOD));
        rc = launch_bc7(*ctx, dIn, nBlocks, dOut, opt, planPOD, stream);
    }
    else if (format <= CVTTB200_BC5S)
        rc = launch_s3tc(format, dIn, nBlocks, dOut, opt, stream);
    else if (format == CVTTB200_BC6HU || format == CVTTB200_BC6HS)
        rc = launch_bc6h(dIn, nBlocks, dOut, opt, format == CVTTB200_BC6HS, stream);
    else
        rc = launch_etc(*ctx, format, dIn, nBlocks, dOut, opt, stream);
    if (rc != CVTTB200_OK)
        return rc;

    if (!outOnDevice)
        CVTT_CUDA(cudaMemcpyAsync(out, dOut, nBlocks * outBytes, cudaMemcpyDeviceToHost, stream));
    if (!inOnDevice || !outOnDevice)
        CVTT_CUDA(cudaStreamSynchronize(stream));
    return CVTTB200_OK;
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

    std::lock_guard<std::mutex> lock(g_mutex);
    int device = 0;
    {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return fail_cuda(e, "cudaGetDevice");
    }
    DeviceContext *ctx = nullptr;
    int rc = get_context(device, &ctx);
    if (rc != CVTTB200_OK)
        return rc;

    cudaStream_t stream = (cudaStream_t)streamPtr;
    const bool inOnDevice = is_device_pointer(encoded), outOnDevice = is_device_pointer(pixelBlocks);
    const void *dIn = encoded;
    void *dOut = pixelBlocks;
    if (!inOnDevice)
    {
        rc = ensure_stage(&ctx->stageIn, &ctx->stageInBytes, nBlocks * inBytes);
        if (rc != CVTTB200_OK)
            return rc;
        CVTT_CUDA(cudaMemcpyAsync(ctx->stageIn, encoded, nBlocks * inBytes, cudaMemcpyHostToDevice, stream));
        dIn = ctx->stageIn;
    }
    if (!outOnDevice)
    {
        rc = ensure_stage(&ctx->stageOut, &ctx->stageOutBytes, nBlocks * outBytes);
        if (rc != CVTTB200_OK)
            return rc;
        dOut = ctx->stageOut;
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
