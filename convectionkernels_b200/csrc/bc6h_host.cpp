// Host-side BC6H support (see bc6h_host.h).
#include "bc6h_host.h"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "bc6h_tables.inc"
#include "bc7_tables.inc"

namespace cvttb200
{
    void bc6h_fill_params(BC6HParams &P, const OptionsPOD &options, const float rcpN[17])
    {
        memset(&P, 0, sizeof(P));
        // Util::FillWeights, ConvectionKernels_Util.cpp:62-73
        const bool uniform = (options.flags & kFlag_Uniform) != 0;
        const float w[3] = { uniform ? 1.0f : options.redWeight, uniform ? 1.0f : options.greenWeight, uniform ? 1.0f : options.blueWeight };
        for (int ch = 0; ch < 3; ch++)
        {
            P.w[ch] = w[ch];
            P.wSq[ch] = w[ch] * w[ch];
            P.rcpW[ch] = (w[ch] != 0.0f) ? 1.0f / w[ch] : 1.0f;        // EndpointRefiner.h:52-57
        }
        for (int n = 0; n < 17; n++)
            P.rcpN[n] = rcpN[n];
        for (int r = 0; r < 2; r++)
            for (int tweak = 0; tweak < 4; tweak++)
            {
                // Util::ComputeTweakFactors, ConvectionKernels_Util.cpp:75-85
                const int range = r ? 16 : 8;
                const int totalUnits = range - 1;
                const int minOutsideUnits = (tweak >> 1) & 1, maxOutsideUnits = tweak & 1;
                const int insideUnits = totalUnits - minOutsideUnits - maxOutsideUnits;
                P.tweak[r][tweak][0] = -(float)minOutsideUnits / (float)insideUnits;
                P.tweak[r][tweak][1] = (float)maxOutsideUnits / (float)insideUnits + 1.0f;
            }
        P.flags = options.flags;
        P.tweakRounds = std::min(4, std::max(1, options.seedPoints));
        P.refineRounds = std::min(3, std::max(1, options.refineRoundsBC6H));
        const char *noPrune = getenv("CVTTB200_BC6H_NO_PRUNE");          // A/B timing only: results are identical either way
        P.prune = (noPrune && noPrune[0] == '1') ? 0 : 1;
    }

    const BC6HTables &bc6h_tables()
    {
        static BC6HTables T;
        static bool ready = false;
        if (!ready)
        {
            memset(&T, 0, sizeof(T));
            for (int m = 0; m < 14; m++)
            {
                for (int k = 0; k < 7; k++)
                    T.modes[m][k] = kBC6HModes[m][k];
                for (int i = 0; i < 82; i++)
                    T.headerBits[m][i] = kBC6HHeaderBits[m][i];
            }
            for (int p = 0; p < 32; p++)
            {
                T.partitionMask[p] = kBC7PartitionMask2[p];
                T.fixup[p] = kBC7Fixup2[p];
            }
            ready = true;
        }
        return T;
    }
}
