// Group-of-8 cross-lane primitives shared by the encode kernels (device only).
//
// The reference takes some decisions jointly for the 8 blocks of one call (AnySet / AllSet and per-call maxima over its
// 8 SIMD lanes).  Here a warp holds four such groups, one per 8-lane segment.
#pragma once
#include <stdint.h>

namespace
{
    // The reference's AnySet / AllSet over the 8 lanes of one call (ParallelMath.h:1260-1278): ballots restricted to
    // the lane's 8-lane segment.  Every lane of the warp executes every vote (control flow around votes is uniform).
    struct SegmentVote
    {
        uint32_t segMask;
        __device__ __forceinline__ bool any(bool x) const { return (__ballot_sync(0xffffffffu, x) & segMask) != 0; }
        __device__ __forceinline__ bool all(bool x) const { return (__ballot_sync(0xffffffffu, x) & segMask) == segMask; }
        __device__ __forceinline__ bool warp_any(bool x) const { return __any_sync(0xffffffffu, x) != 0; }
    };

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
}
