// Host-side ETC support (see etc_host.h).
#include "etc_host.h"

#include <math.h>
#include <algorithm>
#include <string.h>

#include "etc_tables.inc"

namespace cvttb200
{
    void etc_fill_params(ETCParams &P, const OptionsPOD &options, const OptionsPOD &allocOptions)
    {
        memset(&P, 0, sizeof(P));
        P.flags = options.flags;
        {
            // CompressETC2Block, ETC.cpp:1672-1675 (std::min / std::max operand order kept for NaN thresholds)
            const float fThreshold = std::max<float>(std::min<float>(1.0f, options.threshold), 0.0f) * 255.0f;
            P.punchThreshold = (int)static_cast<uint16_t>(floorf(fThreshold + 1.0f));
        }
        const float cd[3] = { options.redWeight, options.greenWeight, options.blueWeight };
        for (int ch = 0; ch < 3; ch++)
        {
            P.w[ch] = cd[ch];
            P.wSq[ch] = cd[ch] * cd[ch];
        }

        // ETC2CompressionDataInternal::ETC2CompressionDataInternal, ETC.cpp:3117-3145: from the allocation-time weights
        {
            const float cd[3] = { allocOptions.redWeight, allocOptions.greenWeight, allocOptions.blueWeight };
            const float rotCD[3] = { cd[1], cd[2], cd[0] };
            const float offs = -(rotCD[0] * cd[0] + rotCD[1] * cd[1] + rotCD[2] * cd[2]) / (cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
            const float chromaAxis0[3] = { rotCD[0] + cd[0] * offs, rotCD[1] + cd[1] * offs, rotCD[2] + cd[2] * offs };
            const float un[3] =
            {
                chromaAxis0[1] * cd[2] - chromaAxis0[2] * cd[1],
                chromaAxis0[2] * cd[0] - chromaAxis0[0] * cd[2],
                chromaAxis0[0] * cd[1] - chromaAxis0[1] * cd[0]
            };
            const float ca0LengthSq = (chromaAxis0[0] * chromaAxis0[0] + chromaAxis0[1] * chromaAxis0[1] + chromaAxis0[2] * chromaAxis0[2]);
            const float ca1UNLengthSq = (un[0] * un[0] + un[1] * un[1] + un[2] * un[2]);
            const float lengthRatio = sqrtf(ca0LengthSq / ca1UNLengthSq);
            for (int i = 0; i < 3; i++)
            {
                P.chromaAxis0[i] = chromaAxis0[i];
                P.chromaAxis1[i] = un[i] * lengthRatio;
            }
        }

        // EncodePlanar, ETC.cpp:1294-1386: the coefficient sums do not depend on the pixels.  The reference accumulates
        // through aliased references (foh is fho, fvh is fhv, fvo is fov), so those three sums are doubled.
        {
            float fhh = 0.f, fho = 0.f, fhv = 0.f, foo = 0.f, fov = 0.f, fvv = 0.f;
            for (int px = 0; px < 16; px++)
            {
                const float x = (float)(px % 4), y = (float)(px / 4);
                fhh += x * x;
                fhv += x * y;
                fho += x;
                fhv += y * x;
                fvv += y * y;
                fov += y;
                fho += x;
                fov += y;
                foo += 1;
            }
            const float d = 2.0f * fhh, e = fho, f = fhv;
            const float i = fhv, j = fov, k = 2.0f * fvv;
            const float m = fho, n = 2.0f * foo, p = fov;
            const float r0to1 = -i / d, r0to2 = -m / d;
            const float j1 = j + r0to1 * e, k1 = k + r0to1 * f;
            const float n1 = n + r0to2 * e, p1 = p + r0to2 * f;
            const float r1to2 = -p1 / k1;
            const float n2 = n1 + r1to2 * j1;
            const float r2to1 = -j1 / n2;
            const float elim2 = -f / k1, elim1 = -e / n2;
            P.pl_r0to1 = r0to1;
            P.pl_r0to2 = r0to2;
            P.pl_r1to2 = r1to2;
            P.pl_n2 = n2;
            P.pl_r2to1 = r2to1;
            P.pl_elim2 = elim2;
            P.pl_elim1 = elim1;
            P.pl_d = d;
            P.pl_k1 = k1;
        }
    }

    const ETCTables &etc_tables()
    {
        static ETCTables T;
        static bool ready = false;
        if (!ready)
        {
            memset(&T, 0, sizeof(T));
            int src = 0, dst = 0;
            for (int table = 0; table < 8; table++)
            {
                const int count = kETCPotentialOffsets[src];
                for (int i = 0; i <= count; i++)
                    T.potentialOffsets[dst++] = kETCPotentialOffsets[src++];
            }
            for (int i = 0; i < 8; i++)
            {
                T.thModifier[i] = kETCThModifier[i];
                for (int s = 0; s < 4; s++)
                    T.etc1Modifiers[i][s] = kETC1Modifiers[i][s];
            }
            for (int t = 0; t < 16; t++)
            {
                for (int i = 0; i < 4; i++)
                    T.alphaModifier[t][i] = kETCAlphaModifier[t][i];
                for (int i = 0; i < 13; i++)
                    T.alphaRounding[t][i] = kETCAlphaRounding[t][i];
            }
            ready = true;
        }
        return T;
    }

    size_t etc_scratch_bytes(size_t threads)
    {
        return threads * (size_t)(2 * kETCMaxAttempts * 8 + kETCHColors * 16 * 4 + kETCHColors * 4);
    }

    void etc_scratch_layout(ETCScratch &S, void *base, size_t threads)
    {
        char *p = static_cast<char *>(base);
        S.stride = threads;
        S.drsErr = reinterpret_cast<float *>(p);
        p += threads * (size_t)(2 * kETCMaxAttempts * 4);
        S.drsMeta = reinterpret_cast<uint32_t *>(p);
        p += threads * (size_t)(2 * kETCMaxAttempts * 4);
        S.hErr = reinterpret_cast<float *>(p);
        p += threads * (size_t)(kETCHColors * 16 * 4);
        S.hMeta = reinterpret_cast<uint32_t *>(p);
    }
}
