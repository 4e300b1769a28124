// BC7 translation unit of libcvtt_b200.so: classification pre-pass, encode kernels, their launch and device set-up.
// Built with -fmad=false (numerical contract); see build.py.
#include <cuda_runtime.h>

#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cvtt_internal.h"
#include "bc7_host.h"
#include "cvtt_segment.cuh"

using namespace cvttb200;

namespace
{
    constexpr int kBC7Threads = 384;     // 12 warps = 48 reference groups per CTA, one CTA per SM
    constexpr int kBC7CtasPerSM = 1;
    constexpr int kBC7FinishThreads = 128;
    // Small-call launch: slice counts the streams are compiled for.  The largest one whose grid fits one wave of the 148 SMs
    // is taken; calls too large for the smallest take the normal launch.
    constexpr int kBC7SliceChoices[] = { 144, 48, 24, 12, 6, 3, 2 };
    // per thread: 16 packed pixels + 16 gathered biased pixels + 16 gathered pre-weighted pixels
    constexpr size_t kBC7SmemBytes = (size_t)kBC7Threads * 16 * (sizeof(uint32_t) + 2 * sizeof(F4));

    __constant__ BC7PackTables c_bc7PackTables;

    // The exchange of a group of PAIR2 commands (bc7_core.cuh) inside one CTA.  compact() turns the per-thread "wanted classes"
    // masks into one dense list of owner threads per class, each padded to whole 32-slot chunks; chunk k goes to warp k mod 12,
    // so only as many warps as there are chunks walk the second subsets' trials.  Results travel through entries 8..13 of the
    // owner's gathered-pixel arrays (gv for the first command of the group, gw for the second): subset B is the smaller subset
    // of its partition (at most 8 pixels), so during the task phase no gather touches entries 8..15 of anybody's arrays.
    struct BC7CtaExchange
    {
        // a thread wants at most 5 classes per command (one unit of mode 1, two each of modes 3 and 7)
        enum { kWarps = kBC7Threads / 32, kListSlots = kBC7Threads * 5 * kBC7PairGroup + kBC7PairClasses * 32 };
        F4 *gvBase, *gwBase;
        const uint32_t *rawBase;
        uint16_t *list;             // [kListSlots]
        uint16_t *warpCounts;       // [kWarps][kBC7PairClasses]
        int *classBase;             // [kBC7PairClasses + 1], then [kBC7PairClasses] counts
        uint8_t *flags;             // [kBC7Threads]
        uint32_t tid;

        __device__ __forceinline__ void publish(uint32_t f) { flags[tid] = (uint8_t)f; }
        __device__ __forceinline__ int compact(uint32_t wantMask)
        {
            const uint32_t lane = tid & 31, warp = tid >> 5;
#pragma unroll
            for (int c = 0; c < kBC7PairClasses; c++)
            {
                const uint32_t ballot = __ballot_sync(0xffffffffu, (wantMask >> c) & 1u);
                if (lane == 0)
                    warpCounts[warp * kBC7PairClasses + c] = (uint16_t)__popc(ballot);
            }
            __syncthreads();
            int pos = 0;
#pragma unroll
            for (int c = 0; c < kBC7PairClasses; c++)
            {
                uint32_t before = 0, total = 0;
#pragma unroll
                for (uint32_t k = 0; k < kWarps; k++)
                {
                    const uint32_t n = warpCounts[k * kBC7PairClasses + c];
                    before += (k < warp) ? n : 0u;
                    total += n;
                }
                if (tid == 0)
                {
                    classBase[c] = pos;
                    classBase[kBC7PairClasses + 1 + c] = (int)total;
                }
                const uint32_t ballot = __ballot_sync(0xffffffffu, (wantMask >> c) & 1u);
                if ((wantMask >> c) & 1u)
                    list[pos + before + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)tid;
                pos += (int)((total + 31u) & ~31u);
            }
            if (tid == 0)
                classBase[kBC7PairClasses] = pos;
            __syncthreads();
            return pos;
        }
        __device__ __forceinline__ int first_slot() const { return (int)tid; }
        __device__ __forceinline__ int slot_stride() const { return kBC7Threads; }
        __device__ __forceinline__ bool chunk_in_range(int slot, int total) const { return slot - (int)(tid & 31) < total; }
        __device__ __forceinline__ void task(int slot, int &cls, int &owner) const
        {
            const int chunk = slot - (int)(tid & 31);        // warp-uniform
            cls = 0;
#pragma unroll
            for (int c = 1; c < kBC7PairClasses; c++)
                cls = (chunk >= classBase[c]) ? c : cls;
            const int base = classBase[cls], count = classBase[kBC7PairClasses + 1 + cls];
            owner = (slot - base < count) ? (int)list[slot] : -1;
        }
        __device__ __forceinline__ uint32_t owner_flags(int owner) const { return flags[owner]; }
        __device__ __forceinline__ const uint32_t *owner_raw(int owner) const { return rawBase + owner; }
        __device__ __forceinline__ bool task_any(bool x) const { return __any_sync(0xffffffffu, x) != 0; }
        __device__ __forceinline__ F4 *slot_of(int owner, int cls) const
        {
            return (cls < kBC7PairClassesPerCommand ? gvBase + (8 + cls) * kBC7Threads : gwBase + (8 + cls - kBC7PairClassesPerCommand) * kBC7Threads) + owner;
        }
        __device__ __forceinline__ void post(int owner, int cls, const F4 &r) { *slot_of(owner, cls) = r; }
        __device__ __forceinline__ void sync() { __syncthreads(); }
        __device__ __forceinline__ F4 result(int cls) const { return *slot_of((int)tid, cls); }
        // TRIPLE commands: two words per class, classes 0 .. 7 -> entries 8 .. 15 of gv, then of gw (the subsets a task gathers
        // have at most 7 pixels)
        __device__ __forceinline__ F4 *slot2_of(int owner, int word) const
        {
            return (word < 8 ? gvBase + (8 + word) * kBC7Threads : gwBase + word * kBC7Threads) + owner;
        }
        __device__ __forceinline__ void post2(int owner, int cls, const F4 &a, const F4 &b)
        {
            *slot2_of(owner, 2 * cls) = a;
            *slot2_of(owner, 2 * cls + 1) = b;
        }
        __device__ __forceinline__ void result2(int cls, F4 &a, F4 &b) const
        {
            a = *slot2_of((int)tid, 2 * cls);
            b = *slot2_of((int)tid, 2 * cls + 1);
        }
    };

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


    // warp -> four groups: in call order (counts == nullptr, the smallest calls skip the classification), else (class, four groups
    // of that class) with the expensive classes first, so that the tail of the launch is filled by the cheap opaque warps
    __device__ __forceinline__ void bc7_map_block(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ lists, uint32_t nGroups, uint32_t warp, uint32_t lane,
                                                  bool &active, uint32_t &block)
    {
        if (counts == nullptr)
        {
            const uint32_t group = warp * 4 + (lane >> 3);
            active = group < nGroups;
            block = active ? group * 8 + (lane & 7) : 0;
            return;
        }
        const uint32_t n0 = counts[0], n1 = counts[1], n2 = counts[2];
        const uint32_t w1 = (n1 + 3) >> 2, w2 = (n2 + 3) >> 2, w0 = (n0 + 3) >> 2;
        uint32_t cls, clsCount;
        if (warp < w1) { cls = 1; clsCount = n1; }
        else if (warp < w1 + w2) { cls = 2; clsCount = n2; warp -= w1; }
        else if (warp < w1 + w2 + w0) { cls = 0; clsCount = n0; warp -= w1 + w2; }
        else { cls = 0; clsCount = 0; }      // surplus warp of the last CTA: no work, but it keeps the CTA's barriers company
        const uint32_t slot = warp * 4 + (lane >> 3);
        active = slot < clsCount;
        block = active ? lists[(size_t)cls * nGroups + slot] * 8 + (lane & 7) : 0;
    }

    // One thread per block, warp = 4 reference groups of one class; see cvtt_common.cuh / bc7_core.cuh.
    //  * input: each thread reads its own 64-byte PixelBlockU8 with four 128-bit loads (512 B contiguous per group) and
    //    keeps it packed in shared memory, laid out [pixel][thread] (conflict-free)
    //  * per pixel subset the search gathers the subset's pixels once into two [index][thread] arrays of fp32x4 (biased
    //    value, pre-weighted value); every trial then streams them with 128-bit conflict-free loads
    //  * output: one 128-bit store per thread
    template<bool FAST, bool PUNCH>
    __global__ void __launch_bounds__(kBC7Threads, kBC7CtasPerSM)
    bc7_encode_kernel(const __grid_constant__ BC7Params P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nGroups,
                      const uint32_t *__restrict__ counts, const uint32_t *__restrict__ lists, uint4 *__restrict__ candidates, uint32_t warpBase, uint32_t candStride,
                      uint32_t warpsPerCta)
    {
        extern __shared__ __align__(16) unsigned char smem[];
        F4 *sGv = reinterpret_cast<F4 *>(smem);
        F4 *sGw = sGv + 16 * kBC7Threads;
        uint32_t *sRaw = reinterpret_cast<uint32_t *>(sGw + 16 * kBC7Threads);

        const uint32_t tid = threadIdx.x, lane = tid & 31;

        // warp -> (class, four groups of that class); the expensive classes go first so that the tail of the launch is
        // filled by the cheap opaque warps
        // Sliced launch (P.splitSlices > 0; launch_bc7: small calls, and the partial second wave of a call of one to one and a
        // third waves): blockIdx.y = the slice of the search this CTA runs for its blocks; the winners of the slices go to
        // `candidates` ([slice][thread of the launch]) and bc7_finish_kernel reduces them.  All warps of a CTA still walk one
        // stream, so they stay in step.
        const bool split = P.splitSlices > 0;
        bool active;
        uint32_t block;
        // warpsPerCta < 12 (calls of less than one wave): the CTA's first warps hold blocks, the others only take part in the
        // barriers and in the PAIR2 task phases, so that every SM gets a CTA
        const uint32_t warpInCta = tid >> 5;
        bc7_map_block(counts, lists, nGroups, warpInCta < warpsPerCta ? warpBase + blockIdx.x * warpsPerCta + warpInCta : 0xffffff00u, lane, active, block);

        BC7Lane<kBC7Threads> L;
        L.raw = sRaw + tid;
        L.gv = sGv + tid;
        L.gw = sGw + tid;

        __shared__ uint16_t sTaskList[BC7CtaExchange::kListSlots];
        __shared__ uint16_t sWarpCounts[BC7CtaExchange::kWarps * kBC7PairClasses];
        __shared__ int sClassBase[2 * kBC7PairClasses + 1];
        __shared__ uint8_t sOwnerFlags[kBC7Threads];
        BC7CtaExchange ex;
        ex.gvBase = sGv;
        ex.gwBase = sGw;
        ex.rawBase = sRaw;
        ex.list = sTaskList;
        ex.warpCounts = sWarpCounts;
        ex.classBase = sClassBase;
        ex.flags = sOwnerFlags;
        ex.tid = tid;

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
        lf.warpHasWork = __any_sync(0xffffffffu, active);
        lf.warpAnyMode7 = __any_sync(0xffffffffu, active && mode7);

        BC7Work work;
        const uint32_t *pc = split ? P.cmds + __ldg(P.cmds + blockIdx.y) : P.cmds;
        if (PUNCH)
        {
            SegmentVote vote;
            vote.segMask = segMask;
            bc7_search_block<FAST, kBC7Threads, true>(P, L, lf, vote, ex, pc, work);
        }
        else
        {
            BC7NoVote vote;
            bc7_search_block<FAST, kBC7Threads, false>(P, L, lf, vote, ex, pc, work);
        }

        if (split)
        {
            if (active)
            {
                BC7Candidate c;
                bc7_candidate_pack(work, c);
                uint4 *dst = candidates + ((size_t)blockIdx.y * candStride + blockIdx.x * kBC7Threads + tid) * 2;
                dst[0] = make_uint4(c.a[0], c.a[1], c.a[2], c.a[3]);
                dst[1] = make_uint4(c.b[0], c.b[1], c.b[2], c.b[3]);
            }
            return;
        }
        uint32_t o[4];
        bc7_finish_block<FAST, kBC7Threads>(P, c_bc7PackTables, L, work, o);
        if (active)
            out[block] = make_uint4(o[0], o[1], o[2], o[3]);
    }

    // Second kernel of a sliced launch: thread t of the launch (same warp -> blocks mapping as the encode kernel) takes the
    // (error, reference key) minimum of the slices' winners for its block, selects the winner's indexes and packs the block.
    template<bool FAST>
    __global__ void __launch_bounds__(kBC7FinishThreads)
    bc7_finish_kernel(const __grid_constant__ BC7Params P, const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t nGroups,
                      const uint32_t *__restrict__ counts, const uint32_t *__restrict__ lists, const uint4 *__restrict__ candidates, uint32_t warpBase, uint32_t candStride)
    {
        __shared__ uint32_t sRaw[16 * kBC7FinishThreads];
        const uint32_t tid = threadIdx.x, t = blockIdx.x * kBC7FinishThreads + tid;
        if (t >= candStride)
            return;
        bool active;
        uint32_t block;
        bc7_map_block(counts, lists, nGroups, warpBase + (t >> 5), t & 31, active, block);
        if (!active)
            return;
        const uint4 *src = in + (size_t)block * 4;
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const uint4 v = __ldg(src + q);
            sRaw[(q * 4 + 0) * kBC7FinishThreads + tid] = v.x;
            sRaw[(q * 4 + 1) * kBC7FinishThreads + tid] = v.y;
            sRaw[(q * 4 + 2) * kBC7FinishThreads + tid] = v.z;
            sRaw[(q * 4 + 3) * kBC7FinishThreads + tid] = v.w;
        }
        BC7Lane<kBC7FinishThreads> L;
        L.raw = sRaw + tid;
        L.gv = nullptr;
        L.gw = nullptr;
        BC7Work work;
        bc7_work_reset(work);
        for (int s = 0; s < P.splitSlices; s++)
        {
            const uint4 *c4 = candidates + ((size_t)s * candStride + t) * 2;
            const uint4 a = __ldg(c4), b = __ldg(c4 + 1);
            BC7Candidate c;
            c.a[0] = a.x; c.a[1] = a.y; c.a[2] = a.z; c.a[3] = a.w;
            c.b[0] = b.x; c.b[1] = b.y; c.b[2] = b.z; c.b[3] = b.w;
            bc7_candidate_merge(work, c);
        }
        uint32_t o[4];
        bc7_finish_block<FAST, kBC7FinishThreads>(P, c_bc7PackTables, L, work, o);
        out[block] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

namespace cvttb200
{
    int bc7_device_setup()
    {
        CVTT_CUDA(cudaMemcpyToSymbol(c_bc7PackTables, &bc7_pack_tables(), sizeof(BC7PackTables)));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7SmemBytes));
        CVTT_CUDA(cudaFuncSetAttribute(bc7_encode_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        return CVTTB200_OK;
    }

    // f2_div against __fdiv_rn on `samples` operand pairs (cvttb200_selftest)
    int bc7_selftest_div(uint64_t samples, uint64_t seed, uint64_t *mismatches)
    {
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

    // caller holds ctx.planMutex and has made ctx.device current
    // form: kBC7Stream* | slices << 8
    static int get_plan_commands(DeviceContext &ctx, const BC7PlanPOD &plan, int form, const uint32_t **dCmds)
    {
        for (size_t i = 0; i < ctx.plans.size(); i++)
            if (ctx.plans[i].form == form && memcmp(&ctx.plans[i].plan, &plan, sizeof(plan)) == 0)
            {
                // most recently used last: a launch that looks up two streams must not lose the first to the second's eviction
                std::rotate(ctx.plans.begin() + (ptrdiff_t)i, ctx.plans.begin() + (ptrdiff_t)i + 1, ctx.plans.end());
                *dCmds = ctx.plans.back().dCmds;
                return CVTTB200_OK;
            }
        std::vector<uint32_t> cmds;
        const int slots = bc7_compile_plan(plan, cmds, form & 0xff, form >> 8);
        if (slots > kBC7MaxSlots)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "BC7 plan needs more result slots than the kernel provides");
        if (ctx.plans.size() >= 32)
        {
            // every launch that reads a cached plan was enqueued under planMutex, so after this synchronisation no kernel can
            // still be reading the evicted command stream
            CVTT_CUDA(cudaDeviceSynchronize());
            cudaFree(ctx.plans.front().dCmds);
            ctx.plans.erase(ctx.plans.begin());
        }
        PlanCacheEntry entry;
        entry.plan = plan;
        entry.form = form;
        entry.dCmds = nullptr;
        if (!ctx.setupStream)
            CVTT_CUDA(cudaStreamCreateWithFlags(&ctx.setupStream, cudaStreamNonBlocking));
        CVTT_CUDA(cudaMalloc(&entry.dCmds, cmds.size() * sizeof(uint32_t)));
        // The command stream must have LANDED in device memory before any stream may launch with it: a plain cudaMemcpy from
        // pageable memory returns once the data is staged, and the caller's stream (possibly non-blocking) has no ordering
        // with the copy.  `cmds` stays alive until the synchronisation.
        cudaError_t e = cudaMemcpyAsync(entry.dCmds, cmds.data(), cmds.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx.setupStream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(ctx.setupStream);
        if (e != cudaSuccess)
        {
            cudaFree(entry.dCmds);
            return fail_cuda(e, "upload of the compiled BC7 plan");
        }
        ctx.plans.push_back(entry);
        *dCmds = entry.dCmds;
        return CVTTB200_OK;
    }

    int launch_bc7(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const BC7PlanPOD &plan, const float *rcpN, cudaStream_t stream)
    {
        if (nBlocks > 0xffffff00u)
            return fail(CVTTB200_ERR_BAD_ARGUMENT, "too many blocks for one call");

        BC7Params P;
        bc7_fill_params(P, options, plan, rcpN);
        std::lock_guard<std::mutex> planLock(ctx.planMutex);      // until the launches below are enqueued (see the eviction above)
        // PAIR2 commands hand a block's second subset to another thread; the variants whose trials vote inside the block's
        // group (BC7_RespectPunchThrough) or return more than endpoints (BC7_TrySingleColor) keep the plain command stream
        const bool pairCommands = (options.flags & (kFlag_BC7_RespectPunchThrough | kFlag_BC7_TrySingleColor)) == 0;
        const uint32_t nGroups = (uint32_t)(nBlocks / 8);
        const unsigned ctaWarps = kBC7Threads / 32;
        const bool fast = (options.flags & kFlag_BC7_FastIndexing) != 0, punch = (options.flags & kFlag_BC7_RespectPunchThrough) != 0;

        // Small calls (the reference's own call is 8 blocks).  In the normal launch one warp walks the whole search for its 32
        // blocks, 3.6 ms whatever the size of the call, and a call of fewer than 148 x 12 x 32 blocks leaves SMs empty.  When
        // the call's CTAs times a slice count fit one wave, the search of every block is instead dealt out to that many
        // CTAs (blockIdx.y; kBC7StreamSplit sub-streams, same blocks, disjoint trials) and bc7_finish_kernel takes the
        // (error, reference key) minimum of their winners -- the same block, since the order of the trials is free.
        static const long splitOverride = getenv("CVTTB200_BC7_SPLIT") ? atol(getenv("CVTTB200_BC7_SPLIT")) : -1;      // A/B: 0 = never, n = n slices
        static const long tailOverride = getenv("CVTTB200_BC7_TAIL") ? atol(getenv("CVTTB200_BC7_TAIL")) : -1;         // A/B: 0 = no sliced second wave
        // The small-call launch takes its groups in call order: the classification pre-pass brings up to three more partially
        // filled warps, which costs more slices at the CTA-count steps than same-class warps gain (measured both ways).
        const unsigned plainWarps = (nGroups + 3) / 4, classWarps = nGroups / 4 + 3;
        const unsigned numSMs = (unsigned)ctx.numSMs;
        int slices = 0;
        if (pairCommands && nGroups > 0)
        {
            if (splitOverride > 0)
                slices = (int)splitOverride;
            else if (splitOverride < 0)
                for (int k = 0; k < (int)(sizeof(kBC7SliceChoices) / sizeof(kBC7SliceChoices[0])) && !slices; k++)
                    if ((plainWarps + ctaWarps - 1) / ctaWarps * (unsigned)kBC7SliceChoices[k] <= numSMs)
                        slices = kBC7SliceChoices[k];
        }
        const bool classify = !slices;
        const unsigned warps = classify ? classWarps : plainWarps;
        const unsigned ctas = (warps + ctaWarps - 1) / ctaWarps;

        // Calls of one to one and a third waves (one CTA per SM; a 1024 x 1024 texture is 1.15): the CTAs behind the whole wave
        // would hold a few SMs for a whole CTA time while the others idle, so they are launched sliced as well -- the same warps
        // (the tail of the class order, the cheap opaque groups), with the slice count that still fits the wave, three at
        // least.  65 536 blocks: 11.3 -> 10.0 ms.  Two slices lose (81 920 blocks: 11.7 -> 12.2 ms), and with two or more whole
        // waves in front the last CTAs fill in dynamically and slicing them is neutral or worse (524 288 blocks: 9.46 -> 9.20
        // Mblocks/s), so it stays with the one case.
        unsigned mainCtas = slices ? 0u : ctas, tailCtas = 0;
        int tailSlices = 0;
        if (!slices && pairCommands && tailOverride != 0 && ctas > numSMs && ctas < 2 * numSMs && numSMs / (ctas - numSMs) >= 3)
        {
            tailCtas = ctas - numSMs;
            mainCtas = numSMs;
            tailSlices = (int)std::min<unsigned>(numSMs / tailCtas, 48u);
        }
        const int sliced = slices ? slices : tailSlices;            // slice count of the sliced launch, if there is one
        const unsigned slicedCtas = slices ? ctas : tailCtas;
        const uint32_t slicedWarpBase = slices ? 0u : mainCtas * ctaWarps, candStride = slicedCtas * kBC7Threads;

        const uint32_t *dCmds = nullptr, *dCmdsSliced = nullptr;
        int rc = CVTTB200_OK;
        if (mainCtas)
            rc = get_plan_commands(ctx, plan, pairCommands ? kBC7StreamPair : kBC7StreamPlain, &dCmds);
        if (rc == CVTTB200_OK && sliced)
            rc = get_plan_commands(ctx, plan, kBC7StreamSplit | (sliced << 8), &dCmdsSliced);
        if (rc != CVTTB200_OK)
            return rc;

        // stream-ordered scratch: group classification (counts[4] then lists[3][nGroups]), then the slices' winners
        const size_t classifyBytes = classify ? (4 + 3 * (size_t)nGroups) * sizeof(uint32_t) : 0;
        const size_t candOffset = (classifyBytes + 15) & ~(size_t)15;
        unsigned char *dScratchBytes = nullptr;
        rc = pool_alloc(ctx, (void **)&dScratchBytes, candOffset + (size_t)sliced * candStride * 2 * sizeof(uint4) + 16, stream);
        if (rc != CVTTB200_OK)
            return rc;
        uint32_t *dScratch = classify ? (uint32_t *)dScratchBytes : nullptr;
        uint4 *dCand = (uint4 *)(dScratchBytes + candOffset);
        if (classify)
        {
            CVTT_CUDA(cudaMemsetAsync(dScratch, 0, 4 * sizeof(uint32_t), stream));
            bc7_classify_kernel<<<(unsigned)((nBlocks + 255) / 256), 256, 0, stream>>>((const uint4 *)dIn, (uint32_t)nBlocks, nGroups, dScratch, dScratch + 4);
            g_launches++;
        }
        const uint32_t *dCounts = dScratch, *dLists = classify ? dScratch + 4 : nullptr;
        const uint4 *in = (const uint4 *)dIn;
        uint4 *out = (uint4 *)dOut;

        // at most three partially filled warps (one per class).  Two other ways of avoiding a partial last wave were measured
        // and dropped: 11 or 12 warps per CTA over whole waves (7.57 against 7.80 Mblocks/s: warps are bound to schedulers and
        // three of the four still carry three warps), and a last wave of light CTAs with 3-6 working warps each (no change at
        // 1 048 576 blocks, 3 % slower at 524 288).
        // Calls of half a wave to one wave that are not sliced: CTAs of fewer working warps, one CTA on every SM (32 768 blocks: 148
        // CTAs of 7 working warps instead of 86 of 12).  Not below half a wave: a CTA takes a whole SM whatever its working
        // warps, so spreading a small call would keep concurrent callers' launches from running side by side
        // (tests/test_concurrency_gpu.py), and with fewer than six warps per SM a warp is no faster anyway.
        static const long spreadOverride = getenv("CVTTB200_BC7_SPREAD") ? atol(getenv("CVTTB200_BC7_SPREAD")) : -1;    // A/B: 0 = always 12 warps
        unsigned mainWarpsPerCta = ctaWarps;
        if (mainCtas && !sliced && mainCtas < numSMs && 2 * mainCtas > numSMs && spreadOverride != 0)
        {
            mainWarpsPerCta = std::max(1u, (warps + numSMs - 1) / numSMs);
            mainCtas = (warps + mainWarpsPerCta - 1) / mainWarpsPerCta;
        }
        if (mainCtas)
        {
            P.splitSlices = 0;
            P.cmds = dCmds;
            if (fast && !punch)
                bc7_encode_kernel<true, false><<<mainCtas, kBC7Threads, kBC7SmemBytes, stream>>>(P, in, out, nGroups, dCounts, dLists, nullptr, 0u, 0u, mainWarpsPerCta);
            else if (!fast && !punch)
                bc7_encode_kernel<false, false><<<mainCtas, kBC7Threads, kBC7SmemBytes, stream>>>(P, in, out, nGroups, dCounts, dLists, nullptr, 0u, 0u, mainWarpsPerCta);
            else if (fast)
                bc7_encode_kernel<true, true><<<mainCtas, kBC7Threads, kBC7SmemBytes, stream>>>(P, in, out, nGroups, dCounts, dLists, nullptr, 0u, 0u, mainWarpsPerCta);
            else
                bc7_encode_kernel<false, true><<<mainCtas, kBC7Threads, kBC7SmemBytes, stream>>>(P, in, out, nGroups, dCounts, dLists, nullptr, 0u, 0u, mainWarpsPerCta);
            g_launches++;
        }
        if (sliced)
        {
            P.splitSlices = sliced;
            P.cmds = dCmdsSliced;
            const dim3 grid(slicedCtas, (unsigned)sliced);
            const unsigned finishGrid = (candStride + kBC7FinishThreads - 1) / kBC7FinishThreads;
            if (fast)
            {
                bc7_encode_kernel<true, false><<<grid, kBC7Threads, kBC7SmemBytes, stream>>>(P, in, nullptr, nGroups, dCounts, dLists, dCand, slicedWarpBase, candStride, ctaWarps);
                bc7_finish_kernel<true><<<finishGrid, kBC7FinishThreads, 0, stream>>>(P, in, out, nGroups, dCounts, dLists, dCand, slicedWarpBase, candStride);
            }
            else
            {
                bc7_encode_kernel<false, false><<<grid, kBC7Threads, kBC7SmemBytes, stream>>>(P, in, nullptr, nGroups, dCounts, dLists, dCand, slicedWarpBase, candStride, ctaWarps);
                bc7_finish_kernel<false><<<finishGrid, kBC7FinishThreads, 0, stream>>>(P, in, out, nGroups, dCounts, dLists, dCand, slicedWarpBase, candStride);
            }
            g_launches += 2;
        }
        CVTT_CUDA(cudaFreeAsync(dScratchBytes, stream));
        CVTT_CUDA(cudaGetLastError());
        return CVTTB200_OK;
    }
}
