// BC7 encode search, one 4x4 block per thread (lane = block; see cvtt_common.cuh).
//
// What it reproduces (reference elasota/ConvectionKernels, file:line):
//   BC7Computer::Pack            ConvectionKernels_BC67.cpp:1975-2204
//   BC7Computer::TrySinglePlane  ConvectionKernels_BC67.cpp:1042-1662   (modes 0,1,2,3,6,7)
//   BC7Computer::TryDualPlane    ConvectionKernels_BC67.cpp:1664-1965   (modes 4,5)
//   CompressEndpoints0-7 / Quantize / QuantizeP / Unquantize             :829-938
//   EndpointSelector<N,8>        ConvectionKernels_EndpointSelector.h:13-150
//   UnfinishedEndpoints::FinishLDR ConvectionKernels_UnfinishedEndpoints.h:75-91
//   IndexSelector<N>             ConvectionKernels_IndexSelector.h:27-131
//   EndpointRefiner<N>           ConvectionKernels_EndpointRefiner.h:38-152
//   AggregatedError<N>           ConvectionKernels_AggregatedError.h:11-50
//
// How it differs in structure (not in results):
//   * The plan is compiled on the host into a flat command stream (bc7_plan.cpp) that every lane of every
//     warp walks in lock-step: SHAPE (endpoint fit + trials of one pixel subset for the modes that use it),
//     EVAL (sum the subsets of one (mode, partition) and keep it if better), DUAL (one mode-4/5 rotation).
//   * The reference's "strictly better, in iteration order" rule is implemented as a lexicographic
//     (error, sequence-number) minimum, so commands may be executed in any order.
//   * The reference's 16-bit integer work (endpoint quantise/unquantise, 6-bit interpolation, squared
//     differences) is evaluated *exactly* in fp32 (all values are integers below 2^24), using the
//     round-to-integer magic constant instead of cvt instructions; pixels live in shared memory as
//     (value + 1.5*2^23) so one 128-bit load feeds index selection, error and refinement.
#pragma once

#include "cvtt_common.cuh"

namespace cvttb200
{
    struct alignas(16) F4 { float x, y, z, w; };

    // Per-mode constants of the endpoint quantisers in exact-fp32 form.
    //   QuantizeP(bits, p): ((c << (bits+1)) - c + addend(p)) >> 9, then (v << 1) | p      BC67.cpp:835-851
    //   Quantize(bits):     ((c << bits) - c + 127 + (1 << (7-bits))) >> 8                 BC67.cpp:829-833
    //   Unquantize(bits):   (v << (8-bits)) | (v >> (2*bits-8))                            BC67.cpp:853-861
    // floor(N / 2^k) is computed as rne(N / 2^k - (2^k - 1) / 2^(k+1)), which never hits a tie.
    struct QuantConst
    {
        float qMul;          // (2^b - 1) / 2^k
        float qAdd[2];       // (2 * addend(p) - (2^k - 1)) / 2^(k+1)
        float pMul, pAdd;    // v2 = v * pMul + p * pAdd
        float uMul, uScale, uOff;   // out = v2 * uMul + rne(v2 * uScale + uOff)
        int hasUnq;
    };

    struct BC7ModeConst
    {
        QuantConst q;
        int parityBitMax;    // 1, 2 or 4                                                   BC67.cpp:1183-1189
        int sharedP;         // mode 1: both endpoints take p[0]                            BC67.cpp:871-878
        int indexBits;
    };

    struct IndexConst        // per index precision (2, 3, 4 bits), [bits - 2]
    {
        float maxValue;      // range - 1
        float wScale;        // 64 / (range - 1): rne(index * wScale) == (g_weightReciprocals[range] * index + 256) >> 9
        float rcpMaxIndex;   // 1.0f / (range - 1)                                          EndpointRefiner.h:50
        float tweak[4][2];   // Util::ComputeTweakFactors(tweak, range)                     Util.cpp:75-85
    };

    struct BC7Params
    {
        float w[4], wSq[4], rcpW[4];
        float rcpN[17];                    // _mm_rcp_ps((float)n) of the host the library was initialised on
        IndexConst ic[3];
        BC7ModeConst mc[8];
        QuantConst alphaQ4;                // mode 4 alpha: Quantize(6) + Unquantize(6)
        uint64_t mode7RGBPartitionEnabled;
        uint32_t flags;
        int refineRounds;
        const uint32_t *cmds;
    };

    // Command stream opcodes (built by bc7_plan.cpp)
    enum { kCmdEnd = 0, kCmdShape = 1, kCmdEval = 2, kCmdDual = 3 };
    enum { kBC7MaxSlots = 184 };

    // lexicographic order of the reference's commit sequence: modes 0,1,2,3,6,7 (TrySinglePlane) then 4,5 (TryDualPlane)
    CVTT_HD int bc7_mode_order(int mode) { return (mode < 4) ? mode : (mode == 6 ? 4 : (mode == 7 ? 5 : (mode == 4 ? 6 : 7))); }

    struct BC7LaneFlags
    {
        bool anyBlockHasAlpha;    // group vote, BC67.cpp:1069
        bool allowRGBModes;       // group vote, BC67.cpp:1072
        bool blockHasNonMaxAlpha; // this block
        // warp-level "does any lane need this path" (pure work skipping, never changes a lane's result)
        bool warpAnyRGB, warpAnyPCA4, warpAnyExpand, warpAnyMode7;
    };

    struct BC7Work   // BC67::WorkInfo, BC67.cpp:59-76, in packed form
    {
        float error;
        int key;              // bc7_mode_order(mode) * 64 + (partition | rotation*2+indexSelector)
        int mode, sub;        // sub = partition, or rotation | indexSelector << 2
        uint32_t ep[3][2];    // per subset, per endpoint: r | g << 8 | b << 16 | a << 24
        uint32_t idx[2];      // 16 x 4 bits, pixel order
        uint32_t idx2[2];
    };

    CVTT_HD float quant_one(const QuantConst &q, float c, int p)
    {
        float v = rne(xfma(c, q.qMul, q.qAdd[p]));
        float v2 = xfma(v, q.pMul, p ? q.pAdd : 0.0f);
        if (q.hasUnq)
        {
            float fl = rne(xfma(v2, q.uScale, q.uOff));
            v2 = xfma(v2, q.uMul, fl);
        }
        return v2;
    }

    CVTT_HD uint32_t pack_ep_bytes(const float *e, int nch)
    {
        uint32_t r = 0;
        for (int ch = 0; ch < nch; ch++)
            r |= (as_uint(e[ch] + kMagic) & 0xffu) << (8 * ch);
        return r;
    }

    // ---------------------------------------------------------------------------------------------------------
    // EndpointSelector<NCH, 8> over the pixels of `mask` (ascending pixel order), with unit pixel weights.
    // Pixels are read as (value + kMagic); wv are the channel weights (possibly rotated for modes 4/5).
    template<int NCH>
    CVTT_HD void bc7_endpoint_selector(const F4 *pix, int stride, uint32_t mask, int n, const float *wv, float *base, float *offs)
    {
        float centroid[NCH], cov[NCH * (NCH + 1) / 2];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            centroid[ch] = 0.0f;
#pragma unroll
        for (int i = 0; i < NCH * (NCH + 1) / 2; i++)
            cov[i] = 0.0f;

        // pass 0: centroid (EndpointSelector.h:73-86)
        for (uint32_t m = mask; m; m &= m - 1)
        {
            int px = ctz32(m);
            F4 p = pix[px * stride];
            const float pv[4] = { p.x, p.y, p.z, p.w };
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                centroid[ch] = fadd(centroid[ch], fmul(fsub(pv[ch], kMagic), wv[ch]));
        }
        {
            float denom = (float)n;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                centroid[ch] = fdiv(centroid[ch], denom);
        }

        // pass 1: covariance (EndpointSelector.h:88-95, PackedCovarianceMatrix.h:29-40)
        for (uint32_t m = mask; m; m &= m - 1)
        {
            int px = ctz32(m);
            F4 p = pix[px * stride];
            const float pv[4] = { p.x, p.y, p.z, p.w };
            float diff[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                diff[ch] = fsub(fmul(fsub(pv[ch], kMagic), wv[ch]), centroid[ch]);
            int index = 0;
#pragma unroll
            for (int row = 0; row < NCH; row++)
#pragma unroll
                for (int col = 0; col <= row; col++)
                {
                    cov[index] = fadd(cov[index], fmul(diff[row], diff[col]));
                    index++;
                }
        }

        // power iteration (EndpointSelector.h:97-130, PackedCovarianceMatrix.h:42-60)
        float approx[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            approx[ch] = 1.0f;

        for (int it = 0; it < 8; it++)
        {
            float product[NCH];
#pragma unroll
            for (int row = 0; row < NCH; row++)
            {
                float sum = 0.0f;
#pragma unroll
                for (int col = 0; col < NCH; col++)
                {
                    const int hi = (row > col) ? row : col, lo = (row > col) ? col : row;
                    sum = fadd(sum, fmul(approx[col], cov[hi * (hi + 1) / 2 + lo]));
                }
                product[row] = sum;
            }
            float largest = product[0];
#pragma unroll
            for (int ch = 1; ch < NCH; ch++)
                largest = sse_max(largest, product[ch]);
            safe_denominator(largest);
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                approx[ch] = fdiv(product[ch], largest);
        }

        float approxLen = 0.0f;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            approxLen = fadd(approxLen, fmul(approx[ch], approx[ch]));
        approxLen = sqrtf(approxLen);
        safe_denominator(approxLen);
        float direction[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            direction[ch] = fdiv(approx[ch], approxLen);

        // pass 2: extent along the axis (EndpointSelector.h:132-140)
        float minDist = FLT_MAX, maxDist = -FLT_MAX;
        for (uint32_t m = mask; m; m &= m - 1)
        {
            int px = ctz32(m);
            F4 p = pix[px * stride];
            const float pv[4] = { p.x, p.y, p.z, p.w };
            float dist = 0.0f;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                dist = fadd(dist, fmul(direction[ch], fsub(fmul(fsub(pv[ch], kMagic), wv[ch]), centroid[ch])));
            minDist = sse_min(minDist, dist);
            maxDist = sse_max(maxDist, dist);
        }

        // GetEndpoints (EndpointSelector.h:51-70): divides by the raw channel weight
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
        {
            float mn = fadd(centroid[ch], fmul(direction[ch], minDist));
            float mx = fadd(centroid[ch], fmul(direction[ch], maxDist));
            base[ch] = fdiv(mn, wv[ch]);
            offs[ch] = fdiv(fsub(mx, mn), wv[ch]);
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // The inner search of TrySinglePlane for one (mode, shape): tweaks x parity bits x refine rounds
    // (BC67.cpp:1298-1432).  NCH = numRealChannels.  Result: best error / endpoints / indexes of the shape.
    struct BC7ShapeBest
    {
        float err;
        uint32_t e0, e1;      // packed endpoint bytes
        uint32_t idxLo, idxHi;
    };

    template<int NCH, bool FAST>
    CVTT_HD void bc7_shape_trials(const BC7Params &P, const BC7ModeConst &mc, const F4 *pix, int stride, uint32_t mask, int n,
        int seeds, const float *base, const float *offs, const float *sumV, float staticAlphaError, BC7ShapeBest &out)
    {
        const IndexConst &ic = P.ic[mc.indexBits - 2];
        const float maxV = ic.maxValue, wScale = ic.wScale, rcpMaxIndex = ic.rcpMaxIndex;
        const int R = P.refineRounds;
        const bool uniform = (P.flags & kFlag_Uniform) != 0;
        const float wN = (float)n, wRcp = P.rcpN[n];

        float bestErr = FLT_MAX;
        int bestSeq = 0x7fffffff;
        float bE0[NCH], bE1[NCH];
        uint32_t bLo = 0, bHi = 0;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            bE0[ch] = bE1[ch] = 0.0f;

        for (int tweak = 0; tweak < seeds; tweak++)
        {
            // UnfinishedEndpoints::FinishLDR (UnfinishedEndpoints.h:75-91)
            const float tf0 = ic.tweak[tweak][0], tf1 = ic.tweak[tweak][1];
            float u0[NCH], u1[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
            {
                u0[ch] = rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf0)), 0.0f, 255.0f));
                u1[ch] = rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf1)), 0.0f, 255.0f));
            }

            for (int pIter = 0; pIter < mc.parityBitMax; pIter++)
            {
                const int p0 = pIter & 1;
                const int p1 = mc.sharedP ? p0 : ((pIter >> 1) & 1);

                float e0[NCH], e1[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    e0[ch] = u0[ch];
                    e1[ch] = u1[ch];
                }

                for (int refine = 0; refine < R; refine++)
                {
                    const int seq = ((pIter * 4 + tweak) << 16) + refine;   // reference order: pIter, tweak, refine
                    const bool lastRound = (refine == R - 1);

                    // CompressEndpointsN (BC67.cpp:862-938)
                    float q0[NCH], q1[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                    {
                        q0[ch] = quant_one(mc.q, e0[ch], p0);
                        q1[ch] = quant_one(mc.q, e1[ch], p1);
                    }

                    // IndexSelector<4>::Init (IndexSelector.h:27-78).  For NCH == 3 the alpha endpoints are both 255, so
                    // the fourth channel contributes exactly +0 to every sum below and is left out.
                    float dW[NCH], axis[NCH], om[NCH], d64[NCH], bq[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                        dW[ch] = fmul(fsub(q1[ch], q0[ch]), P.w[ch]);
                    float lenSq = fmul(dW[0], dW[0]);
#pragma unroll
                    for (int ch = 1; ch < NCH; ch++)
                        lenSq = fadd(lenSq, fmul(dW[ch], dW[ch]));
                    safe_denominator(lenSq);
                    const float mdl = fdiv(maxV, lenSq);
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                    {
                        axis[ch] = fmul(fmul(dW[ch], P.w[ch]), mdl);
                        om[ch] = q0[ch] + kMagic;                               // exact
                        d64[ch] = (q1[ch] - q0[ch]) * 0.015625f;                // exact
                        bq[ch] = q0[ch] + 0.0078125f;                           // exact
                    }

                    float acc[NCH], tv[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                        acc[ch] = tv[ch] = 0.0f;
                    float tt = 0.0f, ts = 0.0f, slowErr = 0.0f;
                    uint32_t iLo = 0, iHi = 0;

                    for (uint32_t m = mask; m; m &= m - 1)
                    {
                        const int px = ctz32(m);
                        const F4 p = pix[px * stride];
                        const float pv[4] = { p.x, p.y, p.z, p.w };

                        // SelectIndexLDR (IndexSelector.h:124-131)
                        float dist = fmul(fsub(pv[0], om[0]), axis[0]);
#pragma unroll
                        for (int ch = 1; ch < NCH; ch++)
                            dist = fadd(dist, fmul(fsub(pv[ch], om[ch]), axis[ch]));
                        float idxf = rne(clamp_for_round(dist, 0.0f, maxV));

                        // ReconstructLDR_BC7 (IndexSelector.h:90-100) + ComputeErrorLDR (BCCommon.h:24-43)
                        float wf = xfma(idxf, wScale, kMagic) - kMagic;
                        float d2[NCH];
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++)
                        {
                            const float df = (xfma(wf, d64[ch], bq[ch]) + kMagic) - pv[ch];
                            if (FAST)
                                acc[ch] = xfma(df, df, acc[ch]);                // exact (< 2^24)
                            else
                                d2[ch] = df * df;                               // exact (< 2^16)
                        }

                        if (!FAST)
                        {
                            // BC67.cpp:1364-1386: probe index-1 and index+1 in weighted float error
                            float error;
                            if (uniform)
                            {
                                error = d2[0];
#pragma unroll
                                for (int ch = 1; ch < NCH; ch++)
                                    error = error + d2[ch];
                            }
                            else
                            {
                                error = fmul(d2[0], P.wSq[0]);
#pragma unroll
                                for (int ch = 1; ch < NCH; ch++)
                                    error = fadd(error, fmul(d2[ch], P.wSq[ch]));
                            }
                            const float alt[2] = { fmaxf(idxf, 1.0f) - 1.0f, fminf(idxf + 1.0f, maxV) };
#pragma unroll
                            for (int ii = 0; ii < 2; ii++)
                            {
                                float awf = xfma(alt[ii], wScale, kMagic) - kMagic;
                                float altError;
#pragma unroll
                                for (int ch = 0; ch < NCH; ch++)
                                {
                                    float df = (xfma(awf, d64[ch], bq[ch]) + kMagic) - pv[ch];
                                    float sq = df * df;
                                    if (uniform)
                                        altError = (ch == 0) ? sq : altError + sq;
                                    else
                                        altError = (ch == 0) ? fmul(sq, P.wSq[0]) : fadd(altError, fmul(sq, P.wSq[ch]));
                                }
                                const bool better = altError < error;
                                error = sse_min(error, altError);
                                if (better)
                                    idxf = alt[ii];
                            }
                            slowErr = fadd(slowErr, error);
                        }

                        // EndpointRefiner::ContributeUnweightedPW (EndpointRefiner.h:78-92); the sum of v is the
                        // per-shape constant sumV
                        if (!lastRound)
                        {
                            const float t = fmul(idxf, rcpMaxIndex);
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++)
                                tv[ch] = fadd(tv[ch], fmul(t, fmul(fsub(pv[ch], kMagic), P.w[ch])));
                            tt = fadd(tt, fmul(t, t));
                            ts = fadd(ts, t);
                        }

                        const uint32_t nib = as_uint(idxf + kMagic) & 15u;
                        if (px < 8)
                            iLo |= nib << (4 * px);
                        else
                            iHi |= nib << (4 * (px - 8));
                    }

                    // AggregatedError::Finalize (AggregatedError.h:25-46)
                    float shapeError;
                    if (FAST)
                    {
                        if (uniform)
                        {
                            shapeError = acc[0];
#pragma unroll
                            for (int ch = 1; ch < NCH; ch++)
                                shapeError = shapeError + acc[ch];
                        }
                        else
                        {
                            shapeError = fmul(acc[0], P.wSq[0]);
#pragma unroll
                            for (int ch = 1; ch < NCH; ch++)
                                shapeError = fadd(shapeError, fmul(acc[ch], P.wSq[ch]));
                        }
                    }
                    else
                        shapeError = slowErr;
                    if (NCH == 3)
                        shapeError = fadd(shapeError, staticAlphaError);

                    if (shapeError < bestErr || (shapeError == bestErr && seq < bestSeq))
                    {
                        bestErr = shapeError;
                        bestSeq = seq;
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++)
                        {
                            bE0[ch] = q0[ch];
                            bE1[ch] = q1[ch];
                        }
                        bLo = iLo;
                        bHi = iHi;
                    }

                    // EndpointRefiner::GetRefinedEndpointsLDR (EndpointRefiner.h:99-152)
                    if (!lastRound)
                    {
                        float adenom = fmul(fsub(fmul(tt, wN), fmul(ts, ts)), wRcp);
                        const bool adenomZero = (adenom == 0.0f);
                        if (adenomZero)
                            adenom = 1.0f;
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++)
                        {
                            float a = fdiv(fsub(tv[ch], fmul(fmul(ts, sumV[ch]), wRcp)), adenom);
                            float b = fmul(fsub(sumV[ch], fmul(a, ts)), wRcp);
                            float p1v = b, p2v = fadd(a, b);
                            if (adenomZero)
                                p1v = p2v = fmul(sumV[ch], wRcp);
                            e0[ch] = rne(clamp_for_round(fmul(p1v, P.rcpW[ch]), 0.0f, 255.0f));
                            e1[ch] = rne(clamp_for_round(fmul(p2v, P.rcpW[ch]), 0.0f, 255.0f));
                        }
                    }
                }
            }
        }

        out.err = bestErr;
        out.e0 = pack_ep_bytes(bE0, NCH);
        out.e1 = pack_ep_bytes(bE1, NCH);
        out.idxLo = bLo;
        out.idxHi = bHi;
    }

    // ---------------------------------------------------------------------------------------------------------
    // One (mode, rotation, indexSelector) of TryDualPlane (BC67.cpp:1664-1965).  The caller has swapped the
    // rotation's colour channel with alpha in the pixel store, so .xyz is the rotated RGB and .w the scalar
    // plane; wr/wSqr/rcpWr are the weights permuted the same way.
    template<bool FAST>
    CVTT_HD void bc7_dual_plane(const BC7Params &P, const F4 *pix, int stride, int mode, int rotation, int indexSelector, int seeds,
        const float *wr, const float *wSqr, const float *rcpWr, BC7Work &work)
    {
        const int R = P.refineRounds;
        const bool uniform = (P.flags & kFlag_Uniform) != 0;
        const int rgbPrec = (mode == 4 && indexSelector) ? 3 : 2;
        const int alphaPrec = (mode == 4 && !indexSelector) ? 3 : 2;
        const IndexConst &icRGB = P.ic[rgbPrec - 2], &icA = P.ic[alphaPrec - 2];
        const QuantConst &qRGB = P.mc[mode].q;
        const float wN = 16.0f, wRcp = P.rcpN[16];

        float base[3], offs[3];
        bc7_endpoint_selector<3>(pix, stride, 0xffffu, 16, wr, base, offs);

        // alpha range, sums of the refiner's v terms (identical for every trial)
        float aMin, aMax, sumV[3] = { 0.0f, 0.0f, 0.0f }, sumA = 0.0f;
        for (int px = 0; px < 16; px++)
        {
            const F4 p = pix[px * stride];
            const float a = p.w - kMagic;
            if (px == 0)
                aMin = aMax = a;
            else
            {
                aMin = fminf(aMin, a);
                aMax = fmaxf(aMax, a);
            }
            sumV[0] = fadd(sumV[0], fmul(p.x - kMagic, wr[0]));
            sumV[1] = fadd(sumV[1], fmul(p.y - kMagic, wr[1]));
            sumV[2] = fadd(sumV[2], fmul(p.z - kMagic, wr[2]));
            sumA = fadd(sumA, a);
        }

        float bestRGBError = FLT_MAX, bestAlphaError = FLT_MAX;
        float bRGB0[3] = { 0, 0, 0 }, bRGB1[3] = { 0, 0, 0 }, bA0 = 0.0f, bA1 = 0.0f;
        uint32_t bRGBIdx[2] = { 0, 0 }, bAIdx[2] = { 0, 0 };

        for (int tweak = 0; tweak < seeds; tweak++)
        {
            float e0[3], e1[3], a0, a1;
            {
                const float tf0 = icRGB.tweak[tweak][0], tf1 = icRGB.tweak[tweak][1];
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
                {
                    e0[ch] = rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf0)), 0.0f, 255.0f));
                    e1[ch] = rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf1)), 0.0f, 255.0f));
                }
                // TweakAlpha (BC67.cpp:815-827)
                const float af0 = icA.tweak[tweak][0], af1 = icA.tweak[tweak][1];
                const float aoffs = fsub(aMax, aMin);
                a0 = rne(clamp_for_round(fadd(aMin, fmul(aoffs, af0)), 0.0f, 255.0f));
                a1 = rne(clamp_for_round(fadd(aMin, fmul(aoffs, af1)), 0.0f, 255.0f));
            }

            for (int refine = 0; refine < R; refine++)
            {
                const bool lastRound = (refine == R - 1);

                // CompressEndpoints4 / CompressEndpoints5 (BC67.cpp:896-920)
                float q0[3], q1[3], qa0 = a0, qa1 = a1;
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
                {
                    q0[ch] = quant_one(qRGB, e0[ch], 0);
                    q1[ch] = quant_one(qRGB, e1[ch], 0);
                }
                if (mode == 4)
                {
                    qa0 = quant_one(P.alphaQ4, a0, 0);
                    qa1 = quant_one(P.alphaQ4, a1, 0);
                }

                // alpha IndexSelector<1> with unit weight, RGB IndexSelector<3> with the rotated weights
                float lenSqA = fmul(fsub(qa1, qa0), fsub(qa1, qa0));
                safe_denominator(lenSqA);
                const float axisA = fmul(fsub(qa1, qa0), fdiv(icA.maxValue, lenSqA));
                const float omA = qa0 + kMagic, d64A = (qa1 - qa0) * 0.015625f, bqA = qa0 + 0.0078125f;

                float dW[3], axis[3], om[3], d64[3], bq[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
                    dW[ch] = fmul(fsub(q1[ch], q0[ch]), wr[ch]);
                float lenSq = fmul(dW[0], dW[0]);
                lenSq = fadd(lenSq, fmul(dW[1], dW[1]));
                lenSq = fadd(lenSq, fmul(dW[2], dW[2]));
                safe_denominator(lenSq);
                const float mdl = fdiv(icRGB.maxValue, lenSq);
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
                {
                    axis[ch] = fmul(fmul(dW[ch], wr[ch]), mdl);
                    om[ch] = q0[ch] + kMagic;
                    d64[ch] = (q1[ch] - q0[ch]) * 0.015625f;
                    bq[ch] = q0[ch] + 0.0078125f;
                }

                float acc[3] = { 0, 0, 0 }, accA = 0.0f, tv[3] = { 0, 0, 0 }, tt = 0.0f, ts = 0.0f, tvA = 0.0f, ttA = 0.0f, tsA = 0.0f;
                float slowRGB = 0.0f, slowA = 0.0f;
                uint32_t rgbIdx[2] = { 0, 0 }, aIdx[2] = { 0, 0 };

                for (int px = 0; px < 16; px++)
                {
                    const F4 p = pix[px * stride];
                    const float pv[4] = { p.x, p.y, p.z, p.w };

                    float dist = fmul(fsub(pv[0], om[0]), axis[0]);
                    dist = fadd(dist, fmul(fsub(pv[1], om[1]), axis[1]));
                    dist = fadd(dist, fmul(fsub(pv[2], om[2]), axis[2]));
                    float rgbIndex = rne(clamp_for_round(dist, 0.0f, icRGB.maxValue));
                    float alphaIndex = rne(clamp_for_round(fmul(fsub(pv[3], omA), axisA), 0.0f, icA.maxValue));

                    float wf = xfma(rgbIndex, icRGB.wScale, kMagic) - kMagic;
                    float wfA = xfma(alphaIndex, icA.wScale, kMagic) - kMagic;
                    float d2[3], d2A;
#pragma unroll
                    for (int ch = 0; ch < 3; ch++)
                    {
                        const float df = (xfma(wf, d64[ch], bq[ch]) + kMagic) - pv[ch];
                        if (FAST)
                            acc[ch] = xfma(df, df, acc[ch]);
                        else
                            d2[ch] = df * df;
                    }
                    {
                        const float df = (xfma(wfA, d64A, bqA) + kMagic) - pv[3];
                        if (FAST)
                            accA = xfma(df, df, accA);
                        else
                            d2A = df * df;
                    }

                    if (!FAST)
                    {
                        // BC67.cpp:1822-1868
                        float rgbError, alphaError;
                        if (uniform)
                        {
                            rgbError = d2[0] + d2[1] + d2[2];
                            alphaError = d2A;
                        }
                        else
                        {
                            rgbError = fadd(fadd(fmul(d2[0], wSqr[0]), fmul(d2[1], wSqr[1])), fmul(d2[2], wSqr[2]));
                            alphaError = fmul(d2A, wSqr[3]);
                        }
                        const float altRGB[2] = { fmaxf(rgbIndex, 1.0f) - 1.0f, fminf(rgbIndex + 1.0f, icRGB.maxValue) };
                        const float altA[2] = { fmaxf(alphaIndex, 1.0f) - 1.0f, fminf(alphaIndex + 1.0f, icA.maxValue) };
#pragma unroll
                        for (int ii = 0; ii < 2; ii++)
                        {
                            float awf = xfma(altRGB[ii], icRGB.wScale, kMagic) - kMagic;
                            float awfA = xfma(altA[ii], icA.wScale, kMagic) - kMagic;
                            float s[3];
#pragma unroll
                            for (int ch = 0; ch < 3; ch++)
                            {
                                float df = (xfma(awf, d64[ch], bq[ch]) + kMagic) - pv[ch];
                                s[ch] = df * df;
                            }
                            float dfA = (xfma(awfA, d64A, bqA) + kMagic) - pv[3];
                            float sA = dfA * dfA;
                            float altRGBError, altAlphaError;
                            if (uniform)
                            {
                                altRGBError = s[0] + s[1] + s[2];
                                altAlphaError = sA;
                            }
                            else
                            {
                                altRGBError = fadd(fadd(fmul(s[0], wSqr[0]), fmul(s[1], wSqr[1])), fmul(s[2], wSqr[2]));
                                altAlphaError = fmul(sA, wSqr[3]);
                            }
                            const bool rgbBetter = altRGBError < rgbError, alphaBetter = altAlphaError < alphaError;
                            rgbError = sse_min(altRGBError, rgbError);
                            alphaError = sse_min(altAlphaError, alphaError);
                            if (rgbBetter)
                                rgbIndex = altRGB[ii];
                            if (alphaBetter)
                                alphaIndex = altA[ii];
                        }
                        slowRGB = fadd(slowRGB, rgbError);
                        slowA = fadd(slowA, alphaError);
                    }

                    if (!lastRound)
                    {
                        const float t = fmul(rgbIndex, icRGB.rcpMaxIndex);
#pragma unroll
                        for (int ch = 0; ch < 3; ch++)
                            tv[ch] = fadd(tv[ch], fmul(t, fmul(fsub(pv[ch], kMagic), wr[ch])));
                        tt = fadd(tt, fmul(t, t));
                        ts = fadd(ts, t);
                        const float tA = fmul(alphaIndex, icA.rcpMaxIndex);
                        tvA = fadd(tvA, fmul(tA, fsub(pv[3], kMagic)));
                        ttA = fadd(ttA, fmul(tA, tA));
                        tsA = fadd(tsA, tA);
                    }

                    const uint32_t nibRGB = as_uint(rgbIndex + kMagic) & 15u, nibA = as_uint(alphaIndex + kMagic) & 15u;
                    rgbIdx[px >> 3] |= nibRGB << (4 * (px & 7));
                    aIdx[px >> 3] |= nibA << (4 * (px & 7));
                }

                float errorRGB, errorA;
                if (FAST)
                {
                    if (uniform)
                    {
                        errorRGB = acc[0] + acc[1] + acc[2];
                        errorA = accA;
                    }
                    else
                    {
                        errorRGB = fadd(fadd(fmul(acc[0], wSqr[0]), fmul(acc[1], wSqr[1])), fmul(acc[2], wSqr[2]));
                        errorA = fmul(accA, wSqr[3]);
                    }
                }
                else
                {
                    errorRGB = slowRGB;
                    errorA = slowA;
                }

                if (errorRGB < bestRGBError)
                {
                    bestRGBError = errorRGB;
                    bRGBIdx[0] = rgbIdx[0];
                    bRGBIdx[1] = rgbIdx[1];
#pragma unroll
                    for (int ch = 0; ch < 3; ch++)
                    {
                        bRGB0[ch] = q0[ch];
                        bRGB1[ch] = q1[ch];
                    }
                }
                if (errorA < bestAlphaError)
                {
                    bestAlphaError = errorA;
                    bAIdx[0] = aIdx[0];
                    bAIdx[1] = aIdx[1];
                    bA0 = qa0;
                    bA1 = qa1;
                }

                if (!lastRound)
                {
                    {
                        float adenom = fmul(fsub(fmul(tt, wN), fmul(ts, ts)), wRcp);
                        const bool adenomZero = (adenom == 0.0f);
                        if (adenomZero)
                            adenom = 1.0f;
#pragma unroll
                        for (int ch = 0; ch < 3; ch++)
                        {
                            float a = fdiv(fsub(tv[ch], fmul(fmul(ts, sumV[ch]), wRcp)), adenom);
                            float b = fmul(fsub(sumV[ch], fmul(a, ts)), wRcp);
                            float p1v = b, p2v = fadd(a, b);
                            if (adenomZero)
                                p1v = p2v = fmul(sumV[ch], wRcp);
                            e0[ch] = rne(clamp_for_round(fmul(p1v, rcpWr[ch]), 0.0f, 255.0f));
                            e1[ch] = rne(clamp_for_round(fmul(p2v, rcpWr[ch]), 0.0f, 255.0f));
                        }
                    }
                    {
                        float adenom = fmul(fsub(fmul(ttA, wN), fmul(tsA, tsA)), wRcp);
                        const bool adenomZero = (adenom == 0.0f);
                        if (adenomZero)
                            adenom = 1.0f;
                        float a = fdiv(fsub(tvA, fmul(fmul(tsA, sumA), wRcp)), adenom);
                        float b = fmul(fsub(sumA, fmul(a, tsA)), wRcp);
                        float p1v = b, p2v = fadd(a, b);
                        if (adenomZero)
                            p1v = p2v = fmul(sumA, wRcp);
                        a0 = rne(clamp_for_round(p1v, 0.0f, 255.0f));
                        a1 = rne(clamp_for_round(p2v, 0.0f, 255.0f));
                    }
                }
            }
        }

        const float combinedError = fadd(bestRGBError, bestAlphaError);
        const int sub = rotation * 2 + indexSelector;
        const int key = bc7_mode_order(mode) * 64 + sub;
        if (combinedError < work.error || (combinedError == work.error && key < work.key))
        {
            work.error = combinedError;
            work.key = key;
            work.mode = mode;
            work.sub = rotation | (indexSelector << 2);
            const float c0[4] = { bRGB0[0], bRGB0[1], bRGB0[2], bA0 }, c1[4] = { bRGB1[0], bRGB1[1], bRGB1[2], bA1 };
            work.ep[0][0] = pack_ep_bytes(c0, 4);
            work.ep[0][1] = pack_ep_bytes(c1, 4);
            for (int h = 0; h < 2; h++)
            {
                work.idx[h] = indexSelector ? bAIdx[h] : bRGBIdx[h];
                work.idx2[h] = indexSelector ? bRGBIdx[h] : bAIdx[h];
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // Bit packing tail of BC7Computer::Pack (BC67.cpp:2003-2203)
    struct BC7PackTables
    {
        uint16_t partitionMask2[64];
        uint32_t partitionMap3[64];
        uint8_t fixup2[64];
        uint8_t fixup3[128];
    };

    struct Bits128
    {
        uint64_t lo, hi;
        int offset;
    };

    CVTT_HD void put_bits(Bits128 &b, uint32_t value, int bits)
    {
        if (bits == 0)
            return;
        const uint64_t v = value;
        if (b.offset < 64)
        {
            b.lo |= v << b.offset;
            if (b.offset + bits > 64)
                b.hi |= v >> (64 - b.offset);
        }
        else
            b.hi |= v << (b.offset - 64);
        b.offset += bits;
    }

    CVTT_HD void bc7_pack_block(const BC7Work &work, const BC7PackTables &T, uint32_t out[4])
    {
        // g_modes, BC67.cpp:108-119
        const int mode = work.mode;
        const int numSubsets = (mode == 0 || mode == 2) ? 3 : ((mode == 1 || mode == 3 || mode == 7) ? 2 : 1);
        const int rgbBitsT[8] = { 4, 6, 5, 7, 5, 7, 7, 5 }, alphaBitsT[8] = { 0, 0, 0, 0, 6, 8, 7, 5 };
        const int partBitsT[8] = { 4, 6, 6, 6, 0, 0, 0, 6 }, idxBitsT[8] = { 3, 3, 2, 2, 2, 2, 4, 2 }, aIdxBitsT[8] = { 0, 0, 0, 0, 3, 2, 0, 0 };
        const int pbitT[8] = { 2, 1, 0, 2, 0, 0, 2, 2 };   // 0 none, 1 per subset, 2 per endpoint
        const int rgbBits = rgbBitsT[mode], alphaBits = alphaBitsT[mode], partBits = partBitsT[mode];
        const int indexBits = idxBitsT[mode], alphaIndexBits = aIdxBitsT[mode], pbitMode = pbitT[mode];
        const bool separateAlpha = (mode == 4 || mode == 5);

        uint32_t idx[2] = { work.idx[0], work.idx[1] }, idx2[2] = { work.idx2[0], work.idx2[1] };
        uint32_t ep[3][2];
        for (int s = 0; s < 3; s++)
        {
            ep[s][0] = work.ep[s][0];
            ep[s][1] = work.ep[s][1];
        }

        int fixups[3] = { 0, 0, 0 };
        int partition = 0, rotation = 0, indexSelector = 0;

        if (separateAlpha)
        {
            rotation = work.sub & 3;
            indexSelector = (work.sub >> 2) & 1;
            bool flipRGB = ((idx[0] & 15u) & (1u << (indexBits - 1))) != 0;
            bool flipAlpha = ((idx2[0] & 15u) & (1u << (alphaIndexBits - 1))) != 0;
            if (flipRGB)
            {
                const uint32_t hiIdx = ((1u << indexBits) - 1u) * 0x11111111u;
                idx[0] = hiIdx - idx[0];
                idx[1] = hiIdx - idx[1];
            }
            if (flipAlpha)
            {
                const uint32_t hiIdx = ((1u << alphaIndexBits) - 1u) * 0x11111111u;
                idx2[0] = hiIdx - idx2[0];
                idx2[1] = hiIdx - idx2[1];
            }
            if (indexSelector)
            {
                bool t = flipRGB;
                flipRGB = flipAlpha;
                flipAlpha = t;
            }
            if (flipRGB)
            {
                const uint32_t a = ep[0][0], b = ep[0][1];
                ep[0][0] = (a & 0xff000000u) | (b & 0x00ffffffu);
                ep[0][1] = (b & 0xff000000u) | (a & 0x00ffffffu);
            }
            if (flipAlpha)
            {
                const uint32_t a = ep[0][0], b = ep[0][1];
                ep[0][0] = (b & 0xff000000u) | (a & 0x00ffffffu);
                ep[0][1] = (a & 0xff000000u) | (b & 0x00ffffffu);
            }
        }
        else
        {
            partition = work.sub;
            if (numSubsets == 2)
                fixups[1] = T.fixup2[partition];
            else if (numSubsets == 3)
            {
                fixups[1] = T.fixup3[partition * 2];
                fixups[2] = T.fixup3[partition * 2 + 1];
            }
            const uint32_t hiIdx = (1u << indexBits) - 1u;
            bool flip[3] = { false, false, false };
            for (int s = 0; s < numSubsets; s++)
            {
                const int fx = fixups[s];
                const uint32_t v = (idx[fx >> 3] >> (4 * (fx & 7))) & 15u;
                flip[s] = (v & (1u << (indexBits - 1))) != 0;
            }
            if (flip[0] || flip[1] || flip[2])
            {
                for (int px = 0; px < 16; px++)
                {
                    int subset = 0;
                    if (numSubsets == 2)
                        subset = (T.partitionMask2[partition] >> px) & 1;
                    else if (numSubsets == 3)
                        subset = (T.partitionMap3[partition] >> (px * 2)) & 3;
                    if (flip[subset])
                    {
                        const int sh = 4 * (px & 7);
                        const uint32_t v = (idx[px >> 3] >> sh) & 15u;
                        idx[px >> 3] = (idx[px >> 3] & ~(15u << sh)) | ((hiIdx - v) << sh);
                    }
                }
                for (int s = 0; s < numSubsets; s++)
                    if (flip[s])
                    {
                        // Alpha_Combined swaps 4 channels, Alpha_None 3 (alpha is not stored there)
                        const uint32_t t = ep[s][0];
                        ep[s][0] = ep[s][1];
                        ep[s][1] = t;
                    }
            }
        }

        Bits128 pv;
        pv.lo = pv.hi = 0;
        pv.offset = 0;
        put_bits(pv, 1u << mode, mode + 1);
        put_bits(pv, (uint32_t)partition, partBits);
        if (separateAlpha)
            put_bits(pv, (uint32_t)rotation, 2);
        if (mode == 4)
            put_bits(pv, (uint32_t)indexSelector, 1);

        for (int ch = 0; ch < 3; ch++)
            for (int s = 0; s < numSubsets; s++)
                for (int e = 0; e < 2; e++)
                    put_bits(pv, ((ep[s][e] >> (8 * ch)) & 0xffu) >> (8 - rgbBits), rgbBits);
        if (alphaBits)
            for (int s = 0; s < numSubsets; s++)
                for (int e = 0; e < 2; e++)
                    put_bits(pv, (ep[s][e] >> 24) >> (8 - alphaBits), alphaBits);

        if (pbitMode == 1)
        {
            for (int s = 0; s < numSubsets; s++)
                put_bits(pv, ((ep[s][0] & 0xffu) >> (7 - rgbBits)) & 1u, 1);
        }
        else if (pbitMode == 2)
        {
            for (int s = 0; s < numSubsets; s++)
                for (int e = 0; e < 2; e++)
                    put_bits(pv, ((ep[s][e] & 0xffu) >> (7 - rgbBits)) & 1u, 1);
        }

        for (int px = 0; px < 16; px++)
        {
            int bits = indexBits;
            if (px == 0 || px == fixups[1] || px == fixups[2])
                bits--;
            put_bits(pv, (idx[px >> 3] >> (4 * (px & 7))) & 15u, bits);
        }
        if (separateAlpha)
            for (int px = 0; px < 16; px++)
            {
                int bits = alphaIndexBits;
                if (px == 0)
                    bits--;
                put_bits(pv, (idx2[px >> 3] >> (4 * (px & 7))) & 15u, bits);
            }

        out[0] = (uint32_t)pv.lo;
        out[1] = (uint32_t)(pv.lo >> 32);
        out[2] = (uint32_t)pv.hi;
        out[3] = (uint32_t)(pv.hi >> 32);
    }

    // ---------------------------------------------------------------------------------------------------------
    // The whole search for one block.  `pix` holds the block's 16 pixels as (value + kMagic), element px at
    // pix[px * stride]; it is modified in place during the dual-plane rotations and restored.
    template<bool FAST>
    CVTT_HD void bc7_encode_block(const BC7Params &P, const BC7PackTables &T, F4 *pix, int stride, const BC7LaneFlags &lf, uint32_t out[4])
    {
        BC7Work work;
        work.error = FLT_MAX;
        work.key = -1;
        work.mode = 0;
        work.sub = 0;
        for (int s = 0; s < 3; s++)
            work.ep[s][0] = work.ep[s][1] = 0;
        work.idx[0] = work.idx[1] = work.idx2[0] = work.idx2[1] = 0;

        // per-(mode, shape) results, indexed by the slot numbers the host assigned
        uint32_t res[kBC7MaxSlots][5];

        const bool usePCA4 = lf.anyBlockHasAlpha || !lf.allowRGBModes;                   // BC67.cpp:1121
        const bool allowMode7 = lf.anyBlockHasAlpha || (P.mode7RGBPartitionEnabled != 0); // BC67.cpp:1078
        const bool uniform = (P.flags & kFlag_Uniform) != 0;

        const uint32_t *pc = P.cmds;
        for (;;)
        {
            const uint32_t w0 = pc[0];
            const int op = w0 & 0xff;
            if (op == kCmdEnd)
                break;

            if (op == kCmdShape)
            {
                const int nRuns = (w0 >> 8) & 0xff;
                const bool listedRGB = (w0 >> 16) & 1, listedRGBA = (w0 >> 17) & 1, needRGBA = (w0 >> 18) & 1;
                const uint32_t w1 = pc[1];
                const uint32_t mask = w1 & 0xffffu;
                const int n = (w1 >> 16) & 0xff;

                // per-shape constants: sum of pre-weighted pixels (the refiner's v sums, EndpointRefiner.h:85-88)
                // and the error of replacing alpha by 255 (BC67.cpp:1250-1264)
                float sumV[4] = { 0.0f, 0.0f, 0.0f, 0.0f }, accA = 0.0f;
                for (uint32_t m = mask; m; m &= m - 1)
                {
                    const int px = ctz32(m);
                    const F4 p = pix[px * stride];
                    sumV[0] = fadd(sumV[0], fmul(p.x - kMagic, P.w[0]));
                    sumV[1] = fadd(sumV[1], fmul(p.y - kMagic, P.w[1]));
                    sumV[2] = fadd(sumV[2], fmul(p.z - kMagic, P.w[2]));
                    sumV[3] = fadd(sumV[3], fmul(p.w - kMagic, P.w[3]));
                    const float da = (255.0f + kMagic) - p.w;
                    accA = xfma(da, da, accA);
                }
                const float staticAlphaError = uniform ? accA : fmul(accA, P.wSq[3]);

                // endpoint fits.  Shapes the plan does not list keep the all-zero "unfinished" endpoints the
                // zero-initialised reference build has (SinglePlaneTemporaries, BC67.cpp:803-811).
                float baseRGB[3] = { 0, 0, 0 }, offsRGB[3] = { 0, 0, 0 };
                if (listedRGB && lf.warpAnyRGB)
                {
                    float b3[3], o3[3];
                    bc7_endpoint_selector<3>(pix, stride, mask, n, P.w, b3, o3);
                    if (lf.allowRGBModes)                                       // BC67.cpp:1085
                        for (int ch = 0; ch < 3; ch++)
                        {
                            baseRGB[ch] = b3[ch];
                            offsRGB[ch] = o3[ch];
                        }
                }
                float baseRGBA[4] = { 0, 0, 0, 0 }, offsRGBA[4] = { 0, 0, 0, 0 };
                if (needRGBA && listedRGBA)
                {
                    // ExpandTo<4>(255), UnfinishedEndpoints.h:93-114
                    for (int ch = 0; ch < 3; ch++)
                    {
                        baseRGBA[ch] = baseRGB[ch];
                        offsRGBA[ch] = offsRGB[ch];
                    }
                    baseRGBA[3] = 255.0f;
                    offsRGBA[3] = 0.0f;
                    if (lf.warpAnyPCA4)
                    {
                        float b4[4], o4[4];
                        bc7_endpoint_selector<4>(pix, stride, mask, n, P.w, b4, o4);
                        if (usePCA4)
                            for (int ch = 0; ch < 4; ch++)
                            {
                                baseRGBA[ch] = b4[ch];
                                offsRGBA[ch] = o4[ch];
                            }
                    }
                }

                for (int r = 0; r < nRuns; r++)
                {
                    const uint32_t rw = pc[2 + r];
                    const int mode = rw & 0xf, seeds = (rw >> 4) & 0xf, slot = (rw >> 8) & 0xff;
                    BC7ShapeBest best;
                    if (mode < 4)
                    {
                        if (!lf.warpAnyRGB)
                            continue;
                        bc7_shape_trials<3, FAST>(P, P.mc[mode], pix, stride, mask, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best);
                    }
                    else
                    {
                        if (mode == 7 && !lf.warpAnyMode7)
                            continue;
                        bc7_shape_trials<4, FAST>(P, P.mc[mode], pix, stride, mask, n, seeds, baseRGBA, offsRGBA, sumV, 0.0f, best);
                    }
                    res[slot][0] = as_uint(best.err);
                    res[slot][1] = best.e0;
                    res[slot][2] = best.e1;
                    res[slot][3] = best.idxLo;
                    res[slot][4] = best.idxHi;
                }
                pc += 2 + nRuns;
            }
            else if (op == kCmdEval)
            {
                // partition scan body, BC67.cpp:1602-1660
                const int mode = (w0 >> 8) & 0xff, partition = (w0 >> 16) & 0xff, numSubsets = (w0 >> 24) & 0xff;
                const uint32_t w1 = pc[1];
                pc += 2;
                if ((mode < 4 && !lf.warpAnyRGB) || (mode == 7 && !lf.warpAnyMode7))
                    continue;
                const int slots[3] = { (int)(w1 & 0xff), (int)((w1 >> 8) & 0xff), (int)((w1 >> 16) & 0xff) };
                float totalError = as_float(res[slots[0]][0]);
                for (int s = 1; s < numSubsets; s++)
                    totalError = fadd(totalError, as_float(res[slots[s]][0]));

                const int key = bc7_mode_order(mode) * 64 + partition;
                bool better = totalError < work.error || (totalError == work.error && key < work.key);
                if (mode < 4 && !lf.allowRGBModes)
                    better = false;
                if (mode == 7)
                {
                    if (!allowMode7)
                        better = false;
                    if (lf.anyBlockHasAlpha && ((P.mode7RGBPartitionEnabled >> partition) & 1) == 0 && !lf.blockHasNonMaxAlpha)
                        better = false;                                         // BC67.cpp:1625-1634
                }
                if (better)
                {
                    work.error = totalError;
                    work.key = key;
                    work.mode = mode;
                    work.sub = partition;
                    uint32_t lo = 0, hi = 0;
                    for (int s = 0; s < numSubsets; s++)
                    {
                        work.ep[s][0] = res[slots[s]][1];
                        work.ep[s][1] = res[slots[s]][2];
                        lo |= res[slots[s]][3];
                        hi |= res[slots[s]][4];
                    }
                    work.idx[0] = lo;
                    work.idx[1] = hi;
                }
            }
            else // kCmdDual
            {
                const int mode = (w0 >> 8) & 0xff, rotation = (w0 >> 16) & 0xf, indexSelector = (w0 >> 20) & 0xf, seeds = (w0 >> 24) & 0xf;
                pc += 1;

                // swap the rotation's channel with alpha (BC67.cpp:1690-1716)
                float wr[4] = { P.w[0], P.w[1], P.w[2], P.w[3] }, wSqr[4] = { P.wSq[0], P.wSq[1], P.wSq[2], P.wSq[3] };
                float rcpWr[4] = { P.rcpW[0], P.rcpW[1], P.rcpW[2], P.rcpW[3] };
                if (rotation)
                {
                    const int c = rotation - 1;
                    for (int px = 0; px < 16; px++)
                    {
                        F4 p = pix[px * stride];
                        float t = p.w;
                        if (c == 0) { p.w = p.x; p.x = t; }
                        else if (c == 1) { p.w = p.y; p.y = t; }
                        else { p.w = p.z; p.z = t; }
                        pix[px * stride] = p;
                    }
                    float t;
                    t = wr[3]; wr[3] = wr[c]; wr[c] = t;
                    t = wSqr[3]; wSqr[3] = wSqr[c]; wSqr[c] = t;
                    t = rcpWr[3]; rcpWr[3] = rcpWr[c]; rcpWr[c] = t;
                }

                bc7_dual_plane<FAST>(P, pix, stride, mode, rotation, indexSelector, seeds, wr, wSqr, rcpWr, work);

                if (rotation)
                {
                    const int c = rotation - 1;
                    for (int px = 0; px < 16; px++)
                    {
                        F4 p = pix[px * stride];
                        float t = p.w;
                        if (c == 0) { p.w = p.x; p.x = t; }
                        else if (c == 1) { p.w = p.y; p.y = t; }
                        else { p.w = p.z; p.z = t; }
                        pix[px * stride] = p;
                    }
                }
            }
        }

        bc7_pack_block(work, T, out);
    }
}
