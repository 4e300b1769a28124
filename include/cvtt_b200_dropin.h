// cvtt_b200_dropin.h -- C++ source-compatibility shim: the `cvtt::` names of the reference's public header
// (reference ConvectionKernels.h:31-277) implemented over the C ABI of libcvtt_b200.so (cvtt_b200.h).
//
// A caller written against the reference replaces
//     #include "ConvectionKernels.h"        + link ConvectionKernels.lib
// by  #include "cvtt_b200_dropin.h"         + link -lcvtt_b200
// and keeps calling cvtt::Kernels::EncodeBC7 / EncodeBC1..5 / EncodeBC6H* / EncodeETC* with 8 blocks per call.  For
// throughput, call cvtt::Kernels::B200::Encode with a whole image's blocks instead (same result, one launch).  The decoders
// (DecodeBC7 / DecodeBC6HU / DecodeBC6HS) are provided the same way (B200::Decode for whole images).
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

    // The reference's ETC compression data is CPU scratch memory (ConvectionKernels_ETC.h:36-78).  The device equivalent is
    // allocated by the library per call, so these objects only keep the reference's allocation protocol alive: Alloc* calls
    // allocFunc(context, size) once and returns the pointer, Release* hands it back to freeFunc(context, ptr, size).
    // ETC2CompressionData also carries the Options given to AllocETC2Data: the reference derives the chroma side axes of the
    // T / H search from them at allocation time (ConvectionKernels_ETC.cpp:3117-3145), not from the per-call options.
    class ETC2CompressionData
    {
    public:
        void *m_context;
        Options m_options;
    };

    class ETC1CompressionData
    {
    public:
        void *m_context;
    };

    namespace Kernels
    {
        typedef void *allocFunc_t(void *context, size_t size);
        typedef void freeFunc_t(void *context, void *ptr, size_t size);

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

            // whole-image entry point: numBlocks is any multiple of NumParallelBlocks; pointers may be host or device memory.
            // format is a cvttb200_format; plan is only read for CVTTB200_BC7.
            inline void Encode(int format, uint8_t *pBC, const void *pBlocks, size_t numBlocks, const Options &options, const BC7EncodingPlan *encodingPlan = NULL, void *cudaStream = NULL)
            {
                Check(cvttb200_encode(format, pBlocks, numBlocks, pBC, &options, encodingPlan, cudaStream), "Encode");
            }

            // ETC2 colour formats: compressionData carries the allocation-time options (must not be NULL, as in the reference)
            inline void EncodeETC2Family(int format, uint8_t *pBC, const void *pBlocks, size_t numBlocks, const Options &options, const ETC2CompressionData *compressionData, void *cudaStream = NULL)
            {
                if (!compressionData)
                {
                    fprintf(stderr, "cvtt (B200): ETC2 encode called without ETC2CompressionData (AllocETC2Data)\n");
                    abort();
                }
                Check(cvttb200_encode_ex(format, pBlocks, numBlocks, pBC, &options, NULL, &compressionData->m_options, cudaStream), "EncodeETC2");
            }

            inline void EncodeBC7(uint8_t *pBC, const PixelBlockU8 *pBlocks, size_t numBlocks, const Options &options, const BC7EncodingPlan &encodingPlan, void *cudaStream = NULL)
            {
                Check(cvttb200_encode(CVTTB200_BC7, pBlocks, numBlocks, pBC, &options, &encodingPlan, cudaStream), "EncodeBC7");
            }

            // whole-image decode: format is CVTTB200_BC7 (PixelBlockU8 out) or CVTTB200_BC6HU / _BC6HS (PixelBlockF16 out)
            inline void Decode(int format, void *pBlocks, const uint8_t *pBC, size_t numBlocks, void *cudaStream = NULL)
            {
                Check(cvttb200_decode(format, pBC, numBlocks, pBlocks, cudaStream), "Decode");
            }
        }

        // The reference's entry points (ConvectionKernels.h:242-259): NumParallelBlocks blocks in, NumParallelBlocks blocks out.
        inline void EncodeBC1(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC1, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC2(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC2, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC3(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC3, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC4U(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC4U, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC4S(uint8_t *pBC, const PixelBlockS8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC4S, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC5U(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC5U, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC5S(uint8_t *pBC, const PixelBlockS8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC5S, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC6HU(uint8_t *pBC, const PixelBlockF16 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC6HU, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC6HS(uint8_t *pBC, const PixelBlockF16 *pBlocks, const Options &options) { B200::Encode(CVTTB200_BC6HS, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeBC7(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options, const BC7EncodingPlan &encodingPlan)
        {
            B200::EncodeBC7(pBC, pBlocks, NumParallelBlocks, options, encodingPlan);
        }
        inline void EncodeETC1(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options, ETC1CompressionData *) { B200::Encode(CVTTB200_ETC1, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeETC2(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options, ETC2CompressionData *compressionData) { B200::EncodeETC2Family(CVTTB200_ETC2, pBC, pBlocks, NumParallelBlocks, options, compressionData); }
        inline void EncodeETC2RGBA(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options, ETC2CompressionData *compressionData) { B200::EncodeETC2Family(CVTTB200_ETC2_RGBA, pBC, pBlocks, NumParallelBlocks, options, compressionData); }
        inline void EncodeETC2PunchthroughAlpha(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options, ETC2CompressionData *compressionData) { B200::EncodeETC2Family(CVTTB200_ETC2_PUNCHTHROUGH, pBC, pBlocks, NumParallelBlocks, options, compressionData); }
        inline void EncodeETC2Alpha(uint8_t *pBC, const PixelBlockU8 *pBlocks, const Options &options) { B200::Encode(CVTTB200_ETC2_ALPHA, pBC, pBlocks, NumParallelBlocks, options); }
        inline void EncodeETC2Alpha11(uint8_t *pBC, const PixelBlockScalarS16 *pBlocks, bool isSigned, const Options &options)
        {
            B200::Encode(isSigned ? CVTTB200_EAC_R11S : CVTTB200_EAC_R11U, pBC, pBlocks, NumParallelBlocks, options);
        }

        // ConvectionKernels.h:273-275
        inline void DecodeBC6HU(PixelBlockF16 *pBlocks, const uint8_t *pBC) { B200::Decode(CVTTB200_BC6HU, pBlocks, pBC, NumParallelBlocks); }
        inline void DecodeBC6HS(PixelBlockF16 *pBlocks, const uint8_t *pBC) { B200::Decode(CVTTB200_BC6HS, pBlocks, pBC, NumParallelBlocks); }
        inline void DecodeBC7(PixelBlockU8 *pBlocks, const uint8_t *pBC) { B200::Decode(CVTTB200_BC7, pBlocks, pBC, NumParallelBlocks); }

        inline ETC2CompressionData *AllocETC2Data(allocFunc_t allocFunc, void *context, const Options &options)
        {
            void *buffer = allocFunc(context, sizeof(ETC2CompressionData));
            if (!buffer)
                return NULL;
            ETC2CompressionData *data = static_cast<ETC2CompressionData *>(buffer);
            data->m_context = context;
            data->m_options = options;
            return data;
        }

        inline void ReleaseETC2Data(ETC2CompressionData *compressionData, freeFunc_t freeFunc)
        {
            freeFunc(compressionData->m_context, compressionData, sizeof(ETC2CompressionData));
        }

        inline ETC1CompressionData *AllocETC1Data(allocFunc_t allocFunc, void *context)
        {
            void *buffer = allocFunc(context, sizeof(ETC1CompressionData));
            if (!buffer)
                return NULL;
            ETC1CompressionData *data = static_cast<ETC1CompressionData *>(buffer);
            data->m_context = context;
            return data;
        }

        inline void ReleaseETC1Data(ETC1CompressionData *compressionData, freeFunc_t freeFunc)
        {
            freeFunc(compressionData->m_context, compressionData, sizeof(ETC1CompressionData));
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
