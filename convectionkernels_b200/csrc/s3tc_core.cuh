// BC1-BC5 (S3TC / RGTC) encode search, one 4x4 block per thread (lane = block; see cvtt_common.cuh).
//
// What it reproduces (reference elasota/ConvectionKernels, file:line):
//   S3TCComputer::PackRGB                 ConvectionKernels_S3TC.cpp:717-1051 (non-exhaustive path)
//   S3TCComputer::TestEndpoints           :190-258, QuantizeTo565 :52-69, ParanoidDiff :71-81
//   S3TCComputer::PackExplicitAlpha       :303-341
//   S3TCComputer::PackInterpolatedAlpha   :343-715
//   EndpointSelector<3,8> with pixel weights  ConvectionKernels_EndpointSelector.h:13-150
//   IndexSelector<N>::Init / SelectIndexLDR / ReconstructLDRPrecise  ConvectionKernels_IndexSelector.h:27-131
//   EndpointRefiner<N>                    ConvectionKernels_EndpointRefiner.h:38-152
//
// Every AnySet / AllSet of the default path only skips work whose result cannot change a lane, so lanes are
// independent there.  The exhaustive search (Flags::S3TC_Exhaustive) has one coupling: TestCounts stops feeding its refiner once
// the position inside a cluster reaches the largest pixel count of the reference call (S3TC.cpp:279-283), expressed by Vote::max.
#pragma once

#include "cvtt_common.cuh"

namespace cvttb200
{
#ifndef CVTT_F4_DEFINED
#define CVTT_F4_DEFINED
    struct alignas(16) F4 { float x, y, z, w; };
#endif

    struct S3TCParams
    {
        float w[4], wSq[3], rcpW[3];       // Util::FillWeights
        float rcpN[17];                    // _mm_rcp_ps((float)n) of the host the library was initialised on
        float tweak3[3][2], tweak4[4][2], tweak8[4][2];   // Util::ComputeTweakFactors(tweak, range)
        uint32_t flags;
        int alphaThreshold;                // floor(options.threshold * 255 + 0.5), S3TC.cpp:746
        int seedPoints, refineRoundsS3TC, refineRoundsIIC;   // clamped to >= 1
    };

    // pixels of one lane: px[i * STRIDE] = (r, g, b, a) as floats 0..255 (already biased for the signed formats)
    template<int STRIDE>
    struct S3TCLane
    {
        F4 *px;
    };

    CVTT_HD float chan(const F4 &p, int ch) { return ch == 0 ? p.x : (ch == 1 ? p.y : (ch == 2 ? p.z : p.w)); }
    CVTT_HD int s3_round(float v) { return (int)(unbias(round_biased(v))); }       // RoundAndConvertToU15 of a value in 0..255, round to nearest even

    // QuantizeTo5Bits / QuantizeTo6Bits, S3TC.cpp:52-62
    CVTT_HD int s3_quant5(int v) { const int r = wrap_u16(v * 249 + 1024) >> 11; return (r << 3) | (r >> 2); }
    CVTT_HD int s3_quant6(int v) { const int r = wrap_u16(v * 253 + 512) >> 10; return (r << 2) | (r >> 4); }

    struct S3TCBestRGB
    {
        float error;
        int ep[2][3];
        uint32_t idx;       // 16 x 2 bits
        int range;
    };

    // TestEndpoints (S3TC.cpp:190-258) followed, when `refine`, by EndpointRefiner::GetRefinedEndpointsLDR into ec
    template<int STRIDE>
    CVTT_HD void s3tc_test_endpoints(const S3TCParams &P, const S3TCLane<STRIDE> &L, int ec[2][3], int range, bool refine, S3TCBestRGB &best)
    {
        int ep[2][3];
        for (int e = 0; e < 2; e++)
        {
            ep[e][0] = s3_quant5(ec[e][0]);
            ep[e][1] = s3_quant6(ec[e][1]);
            ep[e][2] = s3_quant5(ec[e][2]);
        }
        const float maxV = (float)(range - 1);
        const int recip = (range == 4) ? 10923 : 16384;       // g_weightReciprocals[range]
        float origin[3], dW[3], axis[3], paranoid[3];
        for (int ch = 0; ch < 3; ch++)
        {
            origin[ch] = (float)ep[0][ch];
            dW[ch] = fmul(fsub((float)ep[1][ch], origin[ch]), P.w[ch]);
            paranoid[ch] = fmul(fabsf((float)(ep[0][ch] - ep[1][ch])), 0.03f);
        }
        float lenSq = fmul(dW[0], dW[0]);
        lenSq = fadd(lenSq, fmul(dW[1], dW[1]));
        lenSq = fadd(lenSq, fmul(dW[2], dW[2]));
        safe_denominator(lenSq);
        const float mdl = fdiv(maxV, lenSq);
        for (int ch = 0; ch < 3; ch++)
            axis[ch] = fmul(fmul(dW[ch], P.w[ch]), mdl);

        float tv[3] = { 0.0f, 0.0f, 0.0f }, sv[3] = { 0.0f, 0.0f, 0.0f }, tt = 0.0f, ts = 0.0f;
        const float rcpMaxIndex = 1.0f / maxV;
        float error = 0.0f;
        int agg[3] = { 0, 0, 0 };
        uint32_t indexes = 0;
        const bool paranoidMetric = (P.flags & kFlag_S3TC_Paranoid) != 0;
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.px[px * STRIDE];
            const float pv[3] = { p.x, p.y, p.z };
            float dist = fmul(fsub(pv[0], origin[0]), axis[0]);
            dist = fadd(dist, fmul(fsub(pv[1], origin[1]), axis[1]));
            dist = fadd(dist, fmul(fsub(pv[2], origin[2]), axis[2]));
            const int index = s3_round(clamp_for_round(dist, 0.0f, maxV));
            indexes |= (uint32_t)index << (2 * px);

            if (refine)
            {
                const float t = fmul((float)index, rcpMaxIndex);
                for (int ch = 0; ch < 3; ch++)
                {
                    const float v = fmul(pv[ch], P.w[ch]);
                    tv[ch] = fadd(tv[ch], fmul(t, v));
                    sv[ch] = fadd(sv[ch], v);
                }
                tt = fadd(tt, fmul(t, t));
                ts = fadd(ts, t);
            }

            // ReconstructLDRPrecise (IndexSelector.h:102-112)
            const int weight = wrap_u16(recip * index + 64) >> 7;
            for (int ch = 0; ch < 3; ch++)
            {
                const int rec = wrap_u16(wrap_u16((256 - weight) * ep[0][ch]) + wrap_u16(weight * ep[1][ch]) + 128) >> 8;
                const int orig = (int)pv[ch];
                if (paranoidMetric)
                {
                    const float absDiff = fadd(fabsf((float)(rec - orig)), paranoid[ch]);
                    error = fadd(error, fmul(fmul(absDiff, absDiff), P.wSq[ch]));
                }
                else
                    agg[ch] += wrap_u16((rec - orig) * (rec - orig));
            }
        }
        if (!paranoidMetric)
        {
            // AggregatedError::Finalize (AggregatedError.h:25-46)
            if (P.flags & kFlag_Uniform)
                error = (float)(agg[0] + agg[1] + agg[2]);
            else
                error = fadd(fadd(fmul((float)agg[0], P.wSq[0]), fmul((float)agg[1], P.wSq[1])), fmul((float)agg[2], P.wSq[2]));
        }

        if (error < best.error)
        {
            best.error = error;
            for (int e = 0; e < 2; e++)
                for (int ch = 0; ch < 3; ch++)
                    best.ep[e][ch] = ep[e][ch];
            best.idx = indexes;
            best.range = range;
        }

        if (refine)
        {
            // EndpointRefiner::GetRefinedEndpoints (EndpointRefiner.h:99-142): all 16 pixels contribute with unit weight
            const float wN = 16.0f, wRcp = P.rcpN[16];
            float adenom = fmul(fsub(fmul(tt, wN), fmul(ts, ts)), wRcp);
            const bool adenomZero = (adenom == 0.0f);
            if (adenomZero)
                adenom = 1.0f;
            for (int ch = 0; ch < 3; ch++)
            {
                const float a = fdiv(fsub(tv[ch], fmul(fmul(ts, sv[ch]), wRcp)), adenom);
                const float b = fmul(fsub(sv[ch], fmul(a, ts)), wRcp);
                float p1 = b, p2 = fadd(a, b);
                if (adenomZero)
                    p1 = p2 = fmul(sv[ch], wRcp);
                ec[0][ch] = s3_round(clamp_for_round(fmul(p1, P.rcpW[ch]), 0.0f, 255.0f));
                ec[1][ch] = s3_round(clamp_for_round(fmul(p2, P.rcpW[ch]), 0.0f, 255.0f));
            }
        }
    }

#ifndef CVTT_SC_QUALIFIER
#if defined(__CUDACC__)
#define CVTT_SC_QUALIFIER static __device__ const
#else
#define CVTT_SC_QUALIFIER static const
#endif
#endif
#if defined(__CUDACC__) || defined(CVTT_HOSTSIM)
#include "s3tc_sc_tables.inc"
#define CVTT_HAVE_S3TC_SC_TABLES 1
#endif

    // no cross-lane coupling: the vote used by callers that never take the exhaustive path
    struct NoVote
    {
        CVTT_HD int max(int v) const { return v; }
    };

    // TestSingleColor (S3TC.cpp:83-188), Flags::S3TC_Exhaustive only
    template<int STRIDE>
    CVTT_HD void s3tc_test_single_color(const S3TCParams &P, const S3TCLane<STRIDE> &L, int range, S3TCBestRGB &best)
    {
#if defined(CVTT_HAVE_S3TC_SC_TABLES) && (defined(__CUDA_ARCH__) || defined(CVTT_HOSTSIM))
        int totals[3] = { 0, 0, 0 };
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.px[px * STRIDE];
            totals[0] += (int)p.x;
            totals[1] += (int)p.y;
            totals[2] += (int)p.z;
        }
        const bool paranoidMetric = (P.flags & kFlag_S3TC_Paranoid) != 0;
        int eps[2][3], interpolated[3];
        float factors[3];
        for (int ch = 0; ch < 3; ch++)
        {
            const int average = (totals[ch] + 8) >> 4;
            const unsigned char *entry = kS3TCSCTables[paranoidMetric ? 1 : 0][range == 3 ? 1 : 0][ch == 1 ? 1 : 0] + 4 * average;
            eps[0][ch] = entry[0];
            eps[1][ch] = entry[1];
            interpolated[ch] = entry[2];
            factors[ch] = fmul(fabsf((float)(int)entry[3]), 0.03f);      // ParanoidFactorForSpan
        }
        float error = 0.0f;
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.px[px * STRIDE];
            const int pv[3] = { (int)p.x, (int)p.y, (int)p.z };
            for (int ch = 0; ch < 3; ch++)
            {
                if (paranoidMetric)
                {
                    const float absDiff = fadd(fabsf((float)(interpolated[ch] - pv[ch])), factors[ch]);
                    error = fadd(error, fmul(fmul(absDiff, absDiff), P.wSq[ch]));
                }
                else
                    error = fadd(error, fmul((float)wrap_u16((interpolated[ch] - pv[ch]) * (interpolated[ch] - pv[ch])), P.wSq[ch]));
            }
        }
        if (error < best.error)
        {
            best.error = error;
            for (int e = 0; e < 2; e++)
                for (int ch = 0; ch < 3; ch++)
                    best.ep[e][ch] = eps[e][ch];
            best.idx = 0x55555555u;       // every index is 1
            best.range = range;
        }
#endif
    }

    // TestCounts (S3TC.cpp:260-301): least-squares endpoints for "the counts[i] next sorted pixels take index i", then TestEndpoints.
    // sortedOrder: 16 x 4 bits, entry e = block pixel of the e-th sorted input (only the first numElements are pixels, the rest
    // of the reference's array is zero); groupMaxElements: the largest numElements of the reference call (see the escape, :279-283).
    template<int STRIDE>
    CVTT_HD void s3tc_test_counts(const S3TCParams &P, const S3TCLane<STRIDE> &L, const int *counts, int nCounts, int numElements, int groupMaxElements,
        uint64_t sortedOrder, S3TCBestRGB &best)
    {
        const float rcpMaxIndex = 1.0f / (float)(nCounts - 1);
        float tv[3] = { 0.0f, 0.0f, 0.0f }, sv[3] = { 0.0f, 0.0f, 0.0f }, tt = 0.0f, ts = 0.0f;
        int contributed = 0, e = 0;
        bool escape = false;
        for (int i = 0; i < nCounts && !escape; i++)
            for (int n = 0; n < counts[i]; n++)
            {
                if (n >= groupMaxElements)
                {
                    escape = true;
                    break;
                }
                if (n < numElements)
                {
                    float v[3] = { 0.0f, 0.0f, 0.0f };
                    if (e < numElements)
                    {
                        const F4 p = L.px[(int)((sortedOrder >> (4 * e)) & 15u) * STRIDE];
                        v[0] = fmul(p.x, P.w[0]);
                        v[1] = fmul(p.y, P.w[1]);
                        v[2] = fmul(p.z, P.w[2]);
                    }
                    const float t = fmul((float)i, rcpMaxIndex);
                    for (int ch = 0; ch < 3; ch++)
                    {
                        tv[ch] = fadd(tv[ch], fmul(t, v[ch]));
                        sv[ch] = fadd(sv[ch], v[ch]);
                    }
                    tt = fadd(tt, fmul(t, t));
                    ts = fadd(ts, t);
                    contributed++;
                }
                e++;
            }

        // EndpointRefiner::GetRefinedEndpointsLDR (EndpointRefiner.h:99-152)
        float wN = (float)contributed;
        safe_denominator(wN);
        const float wRcp = P.rcpN[(int)wN];
        float adenom = fmul(fsub(fmul(tt, wN), fmul(ts, ts)), wRcp);
        const bool adenomZero = (adenom == 0.0f);
        if (adenomZero)
            adenom = 1.0f;
        int ec[2][3];
        for (int ch = 0; ch < 3; ch++)
        {
            const float a = fdiv(fsub(tv[ch], fmul(fmul(ts, sv[ch]), wRcp)), adenom);
            const float b = fmul(fsub(sv[ch], fmul(a, ts)), wRcp);
            float p1 = b, p2 = fadd(a, b);
            if (adenomZero)
                p1 = p2 = fmul(sv[ch], wRcp);
            ec[0][ch] = s3_round(clamp_for_round(fmul(p1, P.rcpW[ch]), 0.0f, 255.0f));
            ec[1][ch] = s3_round(clamp_for_round(fmul(p2, P.rcpW[ch]), 0.0f, 255.0f));
        }
        s3tc_test_endpoints<STRIDE>(P, L, ec, nCounts, false, best);
    }

    // PackRGB (S3TC.cpp:717-1051).  out: the 8 bytes as two little-endian words.  vote.max is the maximum over the 8 blocks of the
    // reference call, needed by the exhaustive search only.
    template<int STRIDE, class Vote>
    CVTT_HD void s3tc_pack_rgb(const S3TCParams &P, const S3TCLane<STRIDE> &L, bool alphaTest, Vote &vote, uint32_t out[2])
    {
        // alpha test: alpha becomes 0 / 255, transparent pixels get weight 0 in the endpoint fit (S3TC.cpp:744-773)
        uint32_t transparentMask = 0;
        if (alphaTest)
            for (int px = 0; px < 16; px++)
                if ((int)L.px[px * STRIDE].w < P.alphaThreshold)
                    transparentMask |= 1u << px;

        // EndpointSelector<3, 8> with pixel weights; the extent pass takes every pixel (EndpointSelector.h:33-41, 132-140)
        float centroid[3] = { 0.0f, 0.0f, 0.0f }, weightTotal = 0.0f, cov[6] = { 0, 0, 0, 0, 0, 0 };
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.px[px * STRIDE];
            const float wgt = ((transparentMask >> px) & 1) ? 0.0f : 1.0f;
            centroid[0] = fadd(centroid[0], fmul(fmul(p.x, P.w[0]), wgt));
            centroid[1] = fadd(centroid[1], fmul(fmul(p.y, P.w[1]), wgt));
            centroid[2] = fadd(centroid[2], fmul(fmul(p.z, P.w[2]), wgt));
            weightTotal = fadd(weightTotal, wgt);
        }
        {
            float denom = weightTotal;
            safe_denominator(denom);
            for (int ch = 0; ch < 3; ch++)
                centroid[ch] = fdiv(centroid[ch], denom);
        }
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.px[px * STRIDE];
            const float wgt = ((transparentMask >> px) & 1) ? 0.0f : 1.0f;
            const float d[3] = { fsub(fmul(p.x, P.w[0]), centroid[0]), fsub(fmul(p.y, P.w[1]), centroid[1]), fsub(fmul(p.z, P.w[2]), centroid[2]) };
            int index = 0;
#pragma unroll
            for (int row = 0; row < 3; row++)
#pragma unroll
                for (int col = 0; col <= row; col++)
                {
                    cov[index] = fadd(cov[index], fmul(fmul(d[row], d[col]), wgt));
                    index++;
                }
        }
        float approx[3] = { 1.0f, 1.0f, 1.0f };
#pragma unroll 1
        for (int it = 0; it < 8; it++)
        {
            float product[3];
#pragma unroll
            for (int row = 0; row < 3; row++)
            {
                float sum = 0.0f;
#pragma unroll
                for (int col = 0; col < 3; col++)
                {
                    const int hi = (row > col) ? row : col, lo = (row > col) ? col : row;
                    sum = fadd(sum, fmul(approx[col], cov[hi * (hi + 1) / 2 + lo]));
                }
                product[row] = sum;
            }
            float largest = sse_max(sse_max(product[0], product[1]), product[2]);
            safe_denominator(largest);
            for (int ch = 0; ch < 3; ch++)
                approx[ch] = fdiv(product[ch], largest);
        }
        float approxLen = fadd(fadd(fadd(0.0f, fmul(approx[0], approx[0])), fmul(approx[1], approx[1])), fmul(approx[2], approx[2]));
        approxLen = sqrtf(approxLen);
        safe_denominator(approxLen);
        const float dir[3] = { fdiv(approx[0], approxLen), fdiv(approx[1], approxLen), fdiv(approx[2], approxLen) };
        float minDist = FLT_MAX, maxDist = -FLT_MAX;
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.px[px * STRIDE];
            float dist = fadd(0.0f, fmul(dir[0], fsub(fmul(p.x, P.w[0]), centroid[0])));
            dist = fadd(dist, fmul(dir[1], fsub(fmul(p.y, P.w[1]), centroid[1])));
            dist = fadd(dist, fmul(dir[2], fsub(fmul(p.z, P.w[2]), centroid[2])));
            minDist = sse_min(minDist, dist);
            maxDist = sse_max(maxDist, dist);
        }
        float base[3], offs[3];
        for (int ch = 0; ch < 3; ch++)
        {
            const float mn = fadd(centroid[ch], fmul(dir[ch], minDist));
            const float mx = fadd(centroid[ch], fmul(dir[ch], maxDist));
            base[ch] = fdiv(mn, P.w[ch]);
            offs[ch] = fdiv(fsub(mx, mn), P.w[ch]);
        }

        S3TCBestRGB best;
        best.error = FLT_MAX;
        best.idx = 0;
        best.range = 0;
        for (int e = 0; e < 2; e++)
            best.ep[e][0] = best.ep[e][1] = best.ep[e][2] = 0;

        if (P.flags & kFlag_S3TC_Exhaustive)
        {
            // sort the pixels along the fitted axis: 11-bit index << 4 | pixel number, transparent pixels get -16 + px (S3TC.cpp:805-842)
            int sortBins[16];
            {
                int sortEP[2][3];
                for (int ch = 0; ch < 3; ch++)
                {
                    // UnfinishedEndpoints::FinishLDR(0, 11): tweak 0 has the factors (-0.0, 1.0)
                    sortEP[0][ch] = s3_round(clamp_for_round(fadd(base[ch], fmul(offs[ch], -0.0f)), 0.0f, 255.0f));
                    sortEP[1][ch] = s3_round(clamp_for_round(fadd(base[ch], fmul(offs[ch], 1.0f)), 0.0f, 255.0f));
                }
                float origin[3], dW[3], axis[3];
                for (int ch = 0; ch < 3; ch++)
                {
                    origin[ch] = (float)sortEP[0][ch];
                    dW[ch] = fmul(fsub((float)sortEP[1][ch], origin[ch]), P.w[ch]);
                }
                float lenSq = fmul(dW[0], dW[0]);
                lenSq = fadd(lenSq, fmul(dW[1], dW[1]));
                lenSq = fadd(lenSq, fmul(dW[2], dW[2]));
                safe_denominator(lenSq);
                const float mdl = fdiv(2047.0f, lenSq);
                for (int ch = 0; ch < 3; ch++)
                    axis[ch] = fmul(fmul(dW[ch], P.w[ch]), mdl);
                for (int px = 0; px < 16; px++)
                {
                    const F4 p = L.px[px * STRIDE];
                    float dist = fmul(fsub(p.x, origin[0]), axis[0]);
                    dist = fadd(dist, fmul(fsub(p.y, origin[1]), axis[1]));
                    dist = fadd(dist, fmul(fsub(p.z, origin[2]), axis[2]));
                    int bin = s3_round(clamp_for_round(dist, 0.0f, 2047.0f)) << 4;
                    if ((transparentMask >> px) & 1)
                        bin = -16;
                    sortBins[px] = bin + px;
                }
            }
            for (int sortEnd = 1; sortEnd < 16; sortEnd++)
                for (int sortLoc = sortEnd; sortLoc > 0; sortLoc--)
                {
                    const int a = sortBins[sortLoc], b = sortBins[sortLoc - 1];
                    sortBins[sortLoc] = a > b ? a : b;
                    sortBins[sortLoc - 1] = a > b ? b : a;
                }
            int firstElement = 0;
            for (int e = 0; e < 16; e++)
                if (sortBins[e] < 0)
                    firstElement = e + 1;
            const int numElements = 16 - firstElement;
            uint64_t sortedOrder = 0;       // sortedInputs[15 - e] = pixels[sortBins[e] & 15]
            for (int e = firstElement; e < 16; e++)
                sortedOrder |= (uint64_t)(sortBins[e] & 15) << (4 * (15 - e));
            const int groupMaxElements = vote.max(numElements);

            for (int n0 = 0; n0 <= 15; n0++)
            {
                const int remainingFor1 = (16 - n0 == 16) ? 15 : 16 - n0;
                for (int n1 = 0; n1 <= remainingFor1; n1++)
                {
                    const int remainingFor2 = (16 - n1 - n0 == 16) ? 15 : 16 - n1 - n0;
                    for (int n2 = 0; n2 <= remainingFor2; n2++)
                    {
                        const int n3 = 16 - n2 - n1 - n0;
                        if (n3 == 16)
                            continue;
                        const int counts[4] = { n0, n1, n2, n3 };
                        s3tc_test_counts<STRIDE>(P, L, counts, 4, numElements, groupMaxElements, sortedOrder, best);
                    }
                }
            }
            s3tc_test_single_color<STRIDE>(P, L, 4, best);
            if (alphaTest)
            {
                for (int n0 = 0; n0 <= 15; n0++)
                {
                    const int remainingFor1 = (16 - n0 == 16) ? 15 : 16 - n0;
                    for (int n1 = 0; n1 <= remainingFor1; n1++)
                    {
                        const int n2 = 16 - n1 - n0;
                        if (n2 == 16)
                            continue;
                        const int counts[4] = { n0, n1, n2, 0 };
                        s3tc_test_counts<STRIDE>(P, L, counts, 3, numElements, groupMaxElements, sortedOrder, best);
                    }
                }
                s3tc_test_single_color<STRIDE>(P, L, 3, best);
            }
        }
        else
        for (int range = alphaTest ? 3 : 4; range <= 4; range++)
        {
            int tweakRounds = (range == 3) ? 3 : 4;               // BCCommon::TweakRoundsForRange
            if (tweakRounds > P.seedPoints)
                tweakRounds = P.seedPoints;
            for (int tweak = 0; tweak < tweakRounds; tweak++)
            {
                const float tf0 = (range == 3) ? P.tweak3[tweak][0] : P.tweak4[tweak][0], tf1 = (range == 3) ? P.tweak3[tweak][1] : P.tweak4[tweak][1];
                int ec[2][3];
                for (int ch = 0; ch < 3; ch++)
                {
                    ec[0][ch] = s3_round(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf0)), 0.0f, 255.0f));
                    ec[1][ch] = s3_round(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf1)), 0.0f, 255.0f));
                }
                for (int refine = 0; refine < P.refineRoundsS3TC; refine++)
                    s3tc_test_endpoints<STRIDE>(P, L, ec, range, refine != P.refineRoundsS3TC - 1, best);
            }
        }

        // 565 packing with the endpoint-order trick (S3TC.cpp:967-1048)
        uint32_t c[2];
        for (int e = 0; e < 2; e++)
            c[e] = (uint32_t)(((best.ep[e][0] & 0xf8) << 8) | ((best.ep[e][1] & 0xfc) << 3) | ((best.ep[e][2] & 0xf8) >> 3));
        uint32_t order;     // indexOrder[i] at bits 2i
        if (best.range == 4)
        {
            if (c[0] == c[1])
                order = 0;
            else if (c[0] < c[1])
            {
                const uint32_t t = c[0]; c[0] = c[1]; c[1] = t;
                order = 1u | (3u << 2) | (2u << 4) | (0u << 6);
            }
            else
                order = 0u | (2u << 2) | (3u << 4) | (1u << 6);
        }
        else
        {
            if (c[0] > c[1])
            {
                const uint32_t t = c[0]; c[0] = c[1]; c[1] = t;
                order = 1u | (2u << 2) | (0u << 4) | (3u << 6);
            }
            else
                order = 0u | (2u << 2) | (1u << 4) | (3u << 6);
        }
        uint32_t packed = 0;
        for (int px = 0; px < 16; px++)
        {
            const uint32_t index = (best.idx >> (2 * px)) & 3u;
            packed |= ((order >> (2 * index)) & 3u) << (2 * px);
        }
        out[0] = c[0] | (c[1] << 16);
        out[1] = packed;
    }

    // PackExplicitAlpha (S3TC.cpp:303-341): 4-bit alpha of BC2
    template<int STRIDE>
    CVTT_HD void s3tc_pack_explicit_alpha(const S3TCLane<STRIDE> &L, int channel, uint32_t out[2])
    {
        // IndexSelector<1>::Init with endpoints 0 and 255, range 16
        const float dW = fmul(fsub(255.0f, 0.0f), 1.0f);
        float lenSq = fmul(dW, dW);
        safe_denominator(lenSq);
        const float axis = fmul(fmul(dW, 1.0f), fdiv(15.0f, lenSq));
        uint32_t w[2] = { 0, 0 };
        for (int px = 0; px < 16; px++)
        {
            const float v = chan(L.px[px * STRIDE], channel);
            const int index = s3_round(clamp_for_round(fmul(fsub(v, 0.0f), axis), 0.0f, 15.0f));
            w[px >> 3] |= (uint32_t)index << (4 * (px & 7));
        }
        out[0] = w[0];
        out[1] = w[1];
    }

    // PackInterpolatedAlpha (S3TC.cpp:343-715): BC3 alpha, BC4, BC5.  out: the 8 bytes as two little-endian words.
    template<int STRIDE>
    CVTT_HD void s3tc_pack_interpolated_alpha(const S3TCParams &P, const S3TCLane<STRIDE> &L, int channel, bool isSigned, uint32_t out[2])
    {
        const int highTerminal = isSigned ? 254 : 255;
        int pixels[16], sorted[16];
        for (int px = 0; px < 16; px++)
        {
            int v = (int)chan(L.px[px * STRIDE], channel);
            if (isSigned)
                v = imin(v, highTerminal);
            pixels[px] = v;
            sorted[px] = v;
        }
        for (int sortEnd = 15; sortEnd > 0; sortEnd--)
            for (int i = 0; i < sortEnd; i++)
            {
                const int a = sorted[i], b = sorted[i + 1];
                sorted[i] = a < b ? a : b;
                sorted[i + 1] = a < b ? b : a;
            }

        bool bestIsFullRange = false;
        float bestError = FLT_MAX;
        int bestEP[2] = { 0, 0 };
        uint32_t bestIdx[2] = { 0, 0 };     // 16 x 4 bits

        const int numRefine = P.refineRoundsIIC;
        const int numTweak = (4 > P.seedPoints) ? P.seedPoints : 4;     // TweakRoundsForRange(8) == TweakRoundsForRange(6) == 4

        // one search over a base / offset pair: range 8 (all eight values interpolated) or 6 (0 and highTerminal are fixed codes)
        auto search = [&](float base, float offset, int range)
        {
            const float maxV = (float)(range - 1), rcpMaxIndex = 1.0f / maxV;
            const int recip = (range == 8) ? 4681 : 6554;         // g_weightReciprocals[range]
            for (int tweak = 0; tweak < numTweak; tweak++)
            {
                // UnfinishedEndpoints<1>::FinishLDR(tweak, 8, ...): both searches use the range-8 tweak factors
                int ep[2];
                ep[0] = s3_round(clamp_for_round(fadd(base, fmul(offset, P.tweak8[tweak][0])), 0.0f, 255.0f));
                ep[1] = s3_round(clamp_for_round(fadd(base, fmul(offset, P.tweak8[tweak][1])), 0.0f, 255.0f));
                for (int refinePass = 0; refinePass < numRefine; refinePass++)
                {
                    const bool refine = refinePass != numRefine - 1;
                    if (isSigned)
                    {
                        ep[0] = imin(ep[0], highTerminal);
                        ep[1] = imin(ep[1], highTerminal);
                    }
                    const float origin = (float)ep[0];
                    const float dW = fmul(fsub((float)ep[1], origin), 1.0f);
                    float lenSq = fmul(dW, dW);
                    safe_denominator(lenSq);
                    const float axis = fmul(fmul(dW, 1.0f), fdiv(maxV, lenSq));

                    float tv = 0.0f, sv = 0.0f, tt = 0.0f, ts = 0.0f, wsum = 0.0f;
                    uint32_t idx[2] = { 0, 0 };
                    int aggTotal = 0;
                    float error = 0.0f;
                    for (int px = 0; px < 16; px++)
                    {
                        const float fp = (float)pixels[px];
                        const int selected = s3_round(clamp_for_round(fmul(fsub(fp, origin), axis), 0.0f, maxV));
                        const int weight = wrap_u16(recip * selected + 64) >> 7;
                        const int rec = wrap_u16(wrap_u16((256 - weight) * ep[0]) + wrap_u16(weight * ep[1]) + 128) >> 8;
                        const int dSel = wrap_u16((rec - pixels[px]) * (rec - pixels[px]));
                        int index = selected;
                        bool contribute = true;
                        if (range == 8)
                            aggTotal += dSel;
                        else
                        {
                            // codes 6 and 7 are the exact values 0 and highTerminal (S3TC.cpp:620-652)
                            const float zeroError = (float)wrap_u16(pixels[px] * pixels[px]);
                            const float highError = (float)wrap_u16((highTerminal - pixels[px]) * (highTerminal - pixels[px]));
                            const float selectedError = (float)dSel;
                            float bestPixelError = zeroError;
                            index = 6;
                            if (highError < bestPixelError)
                                index = 7;
                            bestPixelError = sse_min(bestPixelError, highError);
                            const bool selectedBetter = selectedError < bestPixelError;
                            contribute = selectedBetter;
                            if (selectedBetter)
                                index = selected;
                            bestPixelError = sse_min(bestPixelError, selectedError);
                            error = fadd(error, bestPixelError);
                        }
                        if (refine && contribute)
                        {
                            const float t = fmul((float)selected, rcpMaxIndex);
                            tv = fadd(tv, fmul(t, fp));
                            sv = fadd(sv, fp);
                            tt = fadd(tt, fmul(t, t));
                            ts = fadd(ts, t);
                            wsum = fadd(wsum, 1.0f);
                        }
                        idx[px >> 3] |= (uint32_t)index << (4 * (px & 7));
                    }
                    if (range == 8)
                        error = (float)aggTotal;          // AggregatedError<1>::Finalize with Flags::Uniform

                    if (error < bestError)
                    {
                        bestError = error;
                        bestIsFullRange = (range == 8);
                        bestIdx[0] = idx[0];
                        bestIdx[1] = idx[1];
                        bestEP[0] = ep[0];
                        bestEP[1] = ep[1];
                    }

                    if (refine)
                    {
                        float wN = wsum;
                        safe_denominator(wN);
                        const float wRcp = P.rcpN[(int)wN];
                        float adenom = fmul(fsub(fmul(tt, wN), fmul(ts, ts)), wRcp);
                        const bool adenomZero = (adenom == 0.0f);
                        if (adenomZero)
                            adenom = 1.0f;
                        const float a = fdiv(fsub(tv, fmul(fmul(ts, sv), wRcp)), adenom);
                        const float b = fmul(fsub(sv, fmul(a, ts)), wRcp);
                        float p1 = b, p2 = fadd(a, b);
                        if (adenomZero)
                            p1 = p2 = fmul(sv, wRcp);
                        ep[0] = s3_round(clamp_for_round(fmul(p1, 1.0f), 0.0f, 255.0f));
                        ep[1] = s3_round(clamp_for_round(fmul(p2, 1.0f), 0.0f, 255.0f));
                    }
                }
            }
        };

        // full precision
        search((float)sorted[0], (float)(sorted[15] - sorted[0]), 8);

        // reduced precision with the two reserved codes
        {
            int heuristicMin = sorted[0], heuristicMax = sorted[15];
            {
                const int largestPossibleRange = heuristicMax - heuristicMin;
                const int lowestPossibleClearance = imin(heuristicMin, highTerminal - heuristicMax);
                const int times10 = (lowestPossibleClearance << 2) + (lowestPossibleClearance << 4);
                // ParallelMath::LessOrEqual on int16 lanes is a strict "<" (ParallelMath.h:740-745)
                if (times10 < largestPossibleRange)
                {
                    for (int firstIndex = 0; firstIndex < 16; firstIndex++)
                    {
                        const int lowClearance = (firstIndex == 0) ? 0 : sorted[firstIndex - 1];
                        for (int lastIndex = firstIndex; lastIndex < 16; lastIndex++)
                        {
                            const int numSkippedHigh = 15 - lastIndex, numSkipped = firstIndex + numSkippedHigh;
                            if (!(0 < numSkipped))      // bestSkipCount is never updated by the reference
                                continue;
                            const int highClearance = (numSkippedHigh == 0) ? 0 : (highTerminal - sorted[16 - numSkippedHigh]);
                            const int clearance = highClearance > lowClearance ? highClearance : lowClearance;
                            const int clearanceTimes10 = (clearance << 2) + (clearance << 4);
                            const int rng = sorted[lastIndex] - sorted[firstIndex];
                            if (clearanceTimes10 < rng)
                            {
                                heuristicMin = sorted[firstIndex];
                                heuristicMax = sorted[lastIndex];
                            }
                        }
                    }
                }
            }
            int simpleMin = 1, simpleMax = highTerminal - 1;
            for (int px = 0; px < 16; px++)
            {
                if (0 < sorted[15 - px])
                    simpleMin = sorted[15 - px];
                if (sorted[px] < highTerminal)
                    simpleMax = sorted[px];
            }
            const int minEPs[2] = { simpleMin, heuristicMin }, maxEPs[2] = { simpleMax, heuristicMax };
            // the reference drops the second candidate when it repeats the first on all lanes; a repeated candidate can never win
            for (int mi = 0; mi < 2; mi++)
                for (int xi = 0; xi < 2; xi++)
                    search((float)minEPs[mi], (float)wrap_u16(maxEPs[xi] - minEPs[mi]), 6);
        }

        // packing (S3TC.cpp:656-714)
        int ep0 = bestEP[0], ep1 = bestEP[1];
        if (isSigned)
        {
            ep0 -= 127;
            ep1 -= 127;
        }
        const bool swapEndpoints = bestIsFullRange != (ep0 > ep1);
        if (swapEndpoints)
        {
            const int t = ep0; ep0 = ep1; ep1 = t;
        }
        const int maxValue = bestIsFullRange ? 7 : 5;
        uint64_t bits = (uint64_t)(ep0 & 0xff) | ((uint64_t)(ep1 & 0xff) << 8);
        for (int px = 0; px < 16; px++)
        {
            int index = (int)((bestIdx[px >> 3] >> (4 * (px & 7))) & 15u);
            if (swapEndpoints && index <= maxValue)
                index = maxValue - index;
            if (index != 0)
            {
                if (index == maxValue)
                    index = 1;
                else if (index < maxValue)
                    index++;
            }
            bits |= (uint64_t)index << (16 + 3 * px);
        }
        out[0] = (uint32_t)bits;
        out[1] = (uint32_t)(bits >> 32);
    }
}
