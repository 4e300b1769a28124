// cvtt_b200_dropin.h -- C++ source-compatibility shim: the `cvtt::` names of the reference's public header
// (reference ConvectionKernels.h:31-277) implemented over the C ABI of libcvtt_b200.so (cvtt_b200.h).
//
// A caller written against the reference replaces
//     #include "ConvectionKernels.h"        + link ConvectionKernels.lib
// by  #include "cvtt_b200_dropin.h"         + link -lcvtt_b200
// and keeps calling cvtt::Kernels::EncodeBC7(pBC, pBlocks, options, plan) with 8 blocks per call.  For throughput,
// call cvtt::Kernels::B200::Encode* with a whole image's blocks instead (same result, one launch).
//
// Types derive from the C PODs, so their layout is the reference's by construction (checked by static_assert below).
// Error behaviour: the reference's functions return void and only assert; this shim prints cvttb200_last_error() and
// aborts when the GPU path fails -- there is no CPU fallback to continue on.
#ifndef CVTT_B200_DROPIN_H
#define CVTT_B200_DROPIN_H

#include <stdio.h>
#include <stdlib.h>

#include "cvtt_b200.h"

namespace cvtt
{
    namespace Flags
    {
        const uint32_t BC7_FastIndexing = CVTTB200_FLAG_BC7_FAST_INDEXING;
        const uint32_t BC7_TrySingleColor = CVTTB200_FLAG_BC7_TRY_SINGLE_COLOR;
        const uint32_t BC7_RespectPunchThrough = CVTTB200_FLAG_BC7_RESPECT_PUNCH_THROUGH;
        const uint32_t BC6H_FastIndexing = CVTTB200_FLAG_BC6H_FAST_INDEXING;
        const uint32_t S3TC_Exhaustive = CVTTB200_FLAG_S3TC_EXHAUSTIVE;
        const uint32_t S3TC_Paranoid = CVTTB200_FLAG_S3TC_PARANOID;
        const uint32_t Uniform = CVTTB200_FLAG_UNIFORM;
        const uint32_t ETC_UseFakeBT709 = CVTTB200_FLAG_ETC_USE_FAKE_BT709;
        const uint32_t ETC_FakeBT709Accurate = CVTTB200_FLAG_ETC_FAKE_BT709_ACCURATE;

        const uint32_t Fastest = BC6H_FastIndexing | BC7_FastIndexing | S3TC_Paranoid;
        const uint32_t Faster = Fastest;
        const uint32_t Fast = BC7_FastIndexing | S3TC_Paranoid;
        const uint32_t Default = Fast;
        const uint32_t Better = S3TC_Paranoid | S3TC_Exhaustive;
        const uint32_t Ultra = BC7_TrySingleColor | S3TC_Paranoid | S3TC_Exhaustive | ETC_FakeBT709Accurate;
    }

    const unsigned int NumParallelBlocks = 8;

    struct Options : cvttb200_options { Options() { cvttb200_options_default(this); } };
    struct BC7FineTuningParams : cvttb200_bc7_fine_tuning { BC7FineTuningParams() { cvttb200_bc7_fine_tuning_default(this); } };
    struct BC7EncodingPlan : cvttb200_bc7_plan
    {
        static const int kNumRGBAShapes = 129;
        static const int kNumRGBShapes = 243;
        BC7EncodingPlan() { cvttb200_bc7_plan_default(this); }
    };

    struct PixelBlockU8 { uint8_t m_pixels[16][4]; };
    struct PixelBlockS8 { int8_t m_pixels[16][4]; };
    struct PixelBlockScalarS16 { int16_t m_pixels[16]; };
    struct PixelBlockF16 { int16_t m_pixels[16][4]; };

    static_assert(sizeof(Options) == 44 && sizeof(BC7EncodingPlan) == 808 && sizeof(BC7FineTuningParams) == 285, "layout must match the reference");

    namespace Kernels
    {
        namespace B200
        {
            inline void Check(int status, const char *what)
            {
                if (status != CVTTB200_OK)
                {
                    fprintf(stderr, "cvtt (B200): %s failed with status %d: %s\n", what, status, cvttb200_last_error());
                    abort();
                }
            }

            // whole-image entry points: numBlocks is any multiple of NumParallelBlocks; pointers may be host or device memory
            inline void EncodeBC7(uint8_t *pBC, const PixelBlockU8 *pBlocks, size_t numBlocks, const Options &options, const BC7EncodingPlan &encodingPlan, void *cudaStream = NULL)
            {
                Check(cvttb200_encode(CVTTB200_BC7, pBlocks, numBlocks, pBC, &options, &encodingPlan, cudaStream), "EncodeBC7");
            }
        }

        inline void EncodeBC7(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options, const BC7EncodingPlan &encodingPlan)
        {
            B200::EncodeBC7(pBC, pBlocks, NumParallelBlocks, options, encodingPlan);
        }

        inline void ConfigureBC7EncodingPlanFromQuality(BC7EncodingPlan &encodingPlan, int quality)
        {
            cvttb200_bc7_plan_from_quality(&encodingPlan, quality);
        }

        inline bool ConfigureBC7EncodingPlanFromFineTuningParams(BC7EncodingPlan &encodingPlan, const BC7FineTuningParams &params)
        {
            return cvttb200_bc7_plan_from_fine_tuning(&encodingPlan, &params) != 0;
        }
    }
}

#endif
