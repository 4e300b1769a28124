// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI shim around the UNMODIFIED reference (elasota/ConvectionKernels).  It is
// compiled by oracle/Makefile together with the reference's own .cpp files, read
// in place from /root/reference (never copied into this repository); the output
// goes to oracle/_ref/libcvtt_ref.so.  tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs are the only callers.
//
// Every entry point forwards to the reference's public API
// (ConvectionKernels.h:236-277) in batches of cvtt::NumParallelBlocks = 8 blocks,
// row-major block order, exactly like etc2packer/etc2packer.cpp:215-282 does.
#include "ConvectionKernels.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include <algorithm>
#include <xmmintrin.h>

namespace
{
    // Mirrors include/cvtt_b200.h: cvttb200_format
    enum Format
    {
        F_BC1 = 1, F_BC2, F_BC3, F_BC4U, F_BC4S, F_BC5U, F_BC5S, F_BC6HU, F_BC6HS, F_BC7,
        F_ETC1, F_ETC2, F_ETC2_RGBA, F_ETC2_PUNCHTHROUGH, F_ETC2_ALPHA, F_EAC_R11U, F_EAC_R11S
    };

    size_t InBytes(int fmt)
    {
        switch (fmt)
        {
        case F_BC6HU: case F_BC6HS: return 128;
        case F_EAC_R11U: case F_EAC_R11S: return 32;
        default: return 64;
        }
    }

    size_t OutBytes(int fmt)
    {
        switch (fmt)
        {
        case F_BC1: case F_BC4U: case F_BC4S: case F_ETC1: case F_ETC2: case F_ETC2_PUNCHTHROUGH:
        case F_ETC2_ALPHA: case F_EAC_R11U: case F_EAC_R11S:
            return 8;
        default:
            return 16;
        }
    }

    void *AllocShim(void *, size_t size)
    {
        void *p = NULL;
        if (posix_memalign(&p, 64, size) != 0)
            return NULL;
        return p;
    }

    void FreeShim(void *, void *ptr, size_t) { free(ptr); }

    // allocOpt: the Options handed to AllocETC2Data (the reference keeps the chroma side axes derived from them)
    void EncodeRange(int fmt, const uint8_t *in, size_t firstGroup, size_t lastGroup, uint8_t *out, const cvtt::Options &opt, const cvtt::BC7EncodingPlan *plan, const cvtt::Options &allocOpt)
    {
        using namespace cvtt;
        const size_t ib = InBytes(fmt) * 8, ob = OutBytes(fmt) * 8;
        ETC1CompressionData *etc1 = NULL;
        ETC2CompressionData *etc2 = NULL;
        if (fmt == F_ETC1)
            etc1 = Kernels::AllocETC1Data(AllocShim, NULL);
        if (fmt == F_ETC2 || fmt == F_ETC2_RGBA || fmt == F_ETC2_PUNCHTHROUGH)
            etc2 = Kernels::AllocETC2Data(AllocShim, NULL, allocOpt);

        for (size_t g = firstGroup; g < lastGroup; g++)
        {
            const uint8_t *src = in + g * ib;
            uint8_t *dst = out + g * ob;
            switch (fmt)
            {
            case F_BC1: Kernels::EncodeBC1(dst, reinterpret_cast<const PixelBlockU8*>(src), opt); break;
            case F_BC2: Kernels::EncodeBC2(dst, reinterpret_cast<const PixelBlockU8*>(src), opt); break;
            case F_BC3: Kernels::EncodeBC3(dst, reinterpret_cast<const PixelBlockU8*>(src), opt); break;
            case F_BC4U: Kernels::EncodeBC4U(dst, reinterpret_cast<const PixelBlockU8*>(src), opt); break;
            case F_BC4S: Kernels::EncodeBC4S(dst, reinterpret_cast<const PixelBlockS8*>(src), opt); break;
            case F_BC5U: Kernels::EncodeBC5U(dst, reinterpret_cast<const PixelBlockU8*>(src), opt); break;
            case F_BC5S: Kernels::EncodeBC5S(dst, reinterpret_cast<const PixelBlockS8*>(src), opt); break;
            case F_BC6HU: Kernels::EncodeBC6HU(dst, reinterpret_cast<const PixelBlockF16*>(src), opt); break;
            case F_BC6HS: Kernels::EncodeBC6HS(dst, reinterpret_cast<const PixelBlockF16*>(src), opt); break;
            case F_BC7: Kernels::EncodeBC7(dst, reinterpret_cast<const PixelBlockU8*>(src), opt, *plan); break;
            case F_ETC1: Kernels::EncodeETC1(dst, reinterpret_cast<const PixelBlockU8*>(src), opt, etc1); break;
            case F_ETC2: Kernels::EncodeETC2(dst, reinterpret_cast<const PixelBlockU8*>(src), opt, etc2); break;
            case F_ETC2_RGBA: Kernels::EncodeETC2RGBA(dst, reinterpret_cast<const PixelBlockU8*>(src), opt, etc2); break;
            case F_ETC2_PUNCHTHROUGH: Kernels::EncodeETC2PunchthroughAlpha(dst, reinterpret_cast<const PixelBlockU8*>(src), opt, etc2); break;
            case F_ETC2_ALPHA: Kernels::EncodeETC2Alpha(dst, reinterpret_cast<const PixelBlockU8*>(src), opt); break;
            case F_EAC_R11U: Kernels::EncodeETC2Alpha11(dst, reinterpret_cast<const PixelBlockScalarS16*>(src), false, opt); break;
            case F_EAC_R11S: Kernels::EncodeETC2Alpha11(dst, reinterpret_cast<const PixelBlockScalarS16*>(src), true, opt); break;
            default: break;
            }
        }

        if (etc1)
            Kernels::ReleaseETC1Data(etc1, FreeShim);
        if (etc2)
            Kernels::ReleaseETC2Data(etc2, FreeShim);
    }
}

extern "C"
{
    size_t cvttref_sizeof_options(void) { return sizeof(cvtt::Options); }
    size_t cvttref_sizeof_plan(void) { return sizeof(cvtt::BC7EncodingPlan); }
    size_t cvttref_sizeof_finetune(void) { return sizeof(cvtt::BC7FineTuningParams); }

    void cvttref_default_options(void *opt) { new (opt) cvtt::Options(); }
    void cvttref_default_plan(void *plan) { new (plan) cvtt::BC7EncodingPlan(); }
    void cvttref_default_finetune(void *ft) { new (ft) cvtt::BC7FineTuningParams(); }

    void cvttref_plan_from_quality(void *plan, int quality)
    {
        cvtt::Kernels::ConfigureBC7EncodingPlanFromQuality(*static_cast<cvtt::BC7EncodingPlan*>(plan), quality);
    }

    int cvttref_plan_from_finetune(void *plan, const void *ft)
    {
        return cvtt::Kernels::ConfigureBC7EncodingPlanFromFineTuningParams(*static_cast<cvtt::BC7EncodingPlan*>(plan), *static_cast<const cvtt::BC7FineTuningParams*>(ft)) ? 1 : 0;
    }

    // Encodes nBlocks (multiple of 8) blocks with nThreads host threads; each thread owns a
    // contiguous range of 8-block groups.  Returns 0, or -1 on a bad argument.
    int cvttref_encode_alloc(int fmt, const void *blocks, size_t nBlocks, void *out, const void *options, const void *plan, const void *etc2AllocOptions, int nThreads);

    int cvttref_encode(int fmt, const void *blocks, size_t nBlocks, void *out, const void *options, const void *plan, int nThreads)
    {
        return cvttref_encode_alloc(fmt, blocks, nBlocks, out, options, plan, options, nThreads);
    }

    // The same with AllocETC2Data receiving its own Options (etc2AllocOptions), as a caller of the reference may do.
    int cvttref_encode_alloc(int fmt, const void *blocks, size_t nBlocks, void *out, const void *options, const void *plan, const void *etc2AllocOptions, int nThreads)
    {
        if (!blocks || !out || !options || !etc2AllocOptions || (nBlocks % cvtt::NumParallelBlocks) != 0)
            return -1;
        if (fmt == F_BC7 && !plan)
            return -1;
        if (fmt < F_BC1 || fmt > F_EAC_R11S)
            return -1;

        cvtt::Options opt, allocOpt;
        memcpy(&opt, options, sizeof(opt));
        memcpy(&allocOpt, etc2AllocOptions, sizeof(allocOpt));
        const cvtt::BC7EncodingPlan *p = static_cast<const cvtt::BC7EncodingPlan*>(plan);

        const size_t nGroups = nBlocks / cvtt::NumParallelBlocks;
        if (nThreads < 1)
            nThreads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
        if (static_cast<size_t>(nThreads) > nGroups)
            nThreads = static_cast<int>(std::max<size_t>(1, nGroups));

        const uint8_t *in = static_cast<const uint8_t*>(blocks);
        uint8_t *dst = static_cast<uint8_t*>(out);

        if (nThreads == 1)
        {
            EncodeRange(fmt, in, 0, nGroups, dst, opt, p, allocOpt);
            return 0;
        }

        std::vector<std::thread> threads;
        for (int t = 0; t < nThreads; t++)
        {
            size_t first = nGroups * t / nThreads;
            size_t last = nGroups * (t + 1) / nThreads;
            threads.emplace_back(EncodeRange, fmt, in, first, last, dst, opt, p, allocOpt);
        }
        for (size_t t = 0; t < threads.size(); t++)
            threads[t].join();
        return 0;
    }

    // nBlocks must be a multiple of 8.  fmt: F_BC7, F_BC6HU, F_BC6HS.
    int cvttref_decode(int fmt, const void *bc, size_t nBlocks, void *pixels)
    {
        if ((nBlocks % 8) != 0)
            return -1;
        const uint8_t *src = static_cast<const uint8_t*>(bc);
        for (size_t g = 0; g < nBlocks / 8; g++)
        {
            switch (fmt)
            {
            case F_BC7: cvtt::Kernels::DecodeBC7(static_cast<cvtt::PixelBlockU8*>(pixels) + g * 8, src + g * 128); break;
            case F_BC6HU: cvtt::Kernels::DecodeBC6HU(static_cast<cvtt::PixelBlockF16*>(pixels) + g * 8, src + g * 128); break;
            case F_BC6HS: cvtt::Kernels::DecodeBC6HS(static_cast<cvtt::PixelBlockF16*>(pixels) + g * 8, src + g * 128); break;
            default: return -1;
            }
        }
        return 0;
    }

    // The host's _mm_rcp_ps, exposed so tests can compare it with the table the product library derives.
    float cvttref_rcp(float v)
    {
        return _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(v)));
    }

    unsigned cvttref_hardware_threads(void) { return std::thread::hardware_concurrency(); }
}
