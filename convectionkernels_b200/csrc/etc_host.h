// Host-side ETC support: per-launch constants and the constant tables of the kernels.
#pragma once

#include "etc_core.cuh"

namespace cvttb200
{
    // weights and the scalar part of EncodePlanar's solve from the per-call options; the chroma side axes from allocOptions, the
    // Options the caller's AllocETC2Data received (ETC2CompressionDataInternal ctor, ETC.cpp:3117-3145, read at :1773)
    void etc_fill_params(ETCParams &P, const OptionsPOD &options, const OptionsPOD &allocOptions);
    const ETCTables &etc_tables();
    // bytes of per-thread scratch for `threads` resident threads (the reference's ETC1/ETC2CompressionData, ETC.h:36-78)
    size_t etc_scratch_bytes(size_t threads);
    // carves the scratch block into the arrays of ETCScratch; thread t of `threads` uses base pointers + t
    void etc_scratch_layout(ETCScratch &S, void *base, size_t threads);
}
