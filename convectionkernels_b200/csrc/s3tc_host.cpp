// Host-side BC1-BC5 support (see s3tc_host.h).
#include "s3tc_host.h"

#include <algorithm>
#include <math.h>
#include <string.h>

namespace cvttb200
{
    static void tweak_factors(int tweak, int range, float *out)
    {
        // Util::ComputeTweakFactors, ConvectionKernels_Util.cpp:75-85
        const int totalUnits = range - 1;
        const int minOutsideUnits = (tweak >> 1) & 1, maxOutsideUnits = tweak & 1;
        const int insideUnits = totalUnits - minOutsideUnits - maxOutsideUnits;
        out[0] = -(float)minOutsideUnits / (float)insideUnits;
        out[1] = (float)maxOutsideUnits / (float)insideUnits + 1.0f;
    }

    void s3tc_fill_params(S3TCParams &P, const OptionsPOD &options, const float rcpN[17])
    {
        memset(&P, 0, sizeof(P));
        // Util::FillWeights, ConvectionKernels_Util.cpp:62-73
        const bool uniform = (options.flags & kFlag_Uniform) != 0;
        P.w[0] = uniform ? 1.0f : options.redWeight;
        P.w[1] = uniform ? 1.0f : options.greenWeight;
        P.w[2] = uniform ? 1.0f : options.blueWeight;
        P.w[3] = uniform ? 1.0f : options.alphaWeight;
        for (int ch = 0; ch < 3; ch++)
        {
            P.wSq[ch] = P.w[ch] * P.w[ch];
            P.rcpW[ch] = (P.w[ch] != 0.0f) ? 1.0f / P.w[ch] : 1.0f;      // EndpointRefiner.h:52-57
        }
        for (int n = 0; n < 17; n++)
            P.rcpN[n] = rcpN[n];
        for (int t = 0; t < 3; t++)
            tweak_factors(t, 3, P.tweak3[t]);
        for (int t = 0; t < 4; t++)
        {
            tweak_factors(t, 4, P.tweak4[t]);
            tweak_factors(t, 8, P.tweak8[t]);
        }
        P.flags = options.flags;
        P.alphaThreshold = (int)(uint16_t)floor(options.threshold * 255.0f + 0.5f);      // S3TC.cpp:746
        P.seedPoints = std::max(1, options.seedPoints);
        P.refineRoundsS3TC = std::max(1, options.refineRoundsS3TC);
        P.refineRoundsIIC = std::max(1, options.refineRoundsIIC);
    }
}
