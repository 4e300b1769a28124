// BC6H encode search, one 4x4 block per thread (lane = block; see cvtt_common.cuh).
//
// What it reproduces (reference elasota/ConvectionKernels, file:line):
//   BC6HComputer::Pack                       ConvectionKernels_BC67.cpp:2665-3051
//   QuantizeSingleEndpointElement*           :2425-2446   (fp32 under MXCSR round-up)
//   UnquantizeSingleEndpointElement*         :2448-2501
//   QuantizeEndpointsSigned/Unsigned         :2503-2595   (fix-up index inversion)
//   EvaluatePartitionedLegality/Single       :2597-2663   (delta coding feasibility, int16 wrap)
//   IndexSelectorHDR<3>                      ConvectionKernels_IndexSelectorHDR.h:16-151
//   UnscaleHDRValueSigned/Unsigned           ConvectionKernels_BC67.cpp:766-787
//   TwosCLHalfToFloat, SqDiff2CL, SqDiffSInt16  ConvectionKernels_ParallelMath.h:996-1058
//   UnfinishedEndpoints::FinishHDR*          ConvectionKernels_UnfinishedEndpoints.h:39-73
//   EndpointRefiner<3>                       ConvectionKernels_EndpointRefiner.h:38-175
//   BC6H_IO::WriteMode0-13                   ConvectionKernels_BC6H_IO.cpp:43-139 (as a bit table, bc6h_tables.inc)
//
// The eight blocks of one reference call are not independent here: a meta round is skipped only
// when *all eight* lanes repeat earlier endpoints (BC67.cpp:2868-2876), and the mode commit loop keeps visiting
// modes while *any* lane still needs a commit, letting a later legal mode overwrite an earlier one
// (BC67.cpp:2936-2983).  Both are expressed through the Vote parameter: a ballot over the lane's 8-lane segment in
// the kernel, a barrier-synchronised OR across eight host threads in tests/hostsim.
#pragma once

#include "cvtt_common.cuh"

#if !defined(__CUDA_ARCH__)
#include <xmmintrin.h>
#endif

namespace cvttb200
{
#ifndef CVTT_F4_DEFINED
#define CVTT_F4_DEFINED
    struct alignas(16) F4 { float x, y, z, w; };
#endif

    struct BC6HParams
    {
        float w[3], wSq[3], rcpW[3];
        float rcpN[17];                    // _mm_rcp_ps((float)n) of the host the library was initialised on
        float tweak[2][4][2];              // Util::ComputeTweakFactors(tweak, range) for range 8 ([0]) and 16 ([1])
        uint32_t flags;
        int tweakRounds, refineRounds;     // clamped to 1..4 and 1..3 (BC67.cpp:2667-2675)
    };

    struct BC6HTables
    {
        uint8_t modes[14][8];              // modeID, partitioned, transformed, aPrec, bPrec[3], pad   (g_hdrModes, BC67.cpp:151-167)
        uint8_t headerBits[14][84];        // see bc6h_tables.inc
        uint16_t partitionMask[32];        // g_partitionMap[0..31]
        uint8_t fixup[32];                 // g_fixupIndexes2[0..31]
    };

    // ---- fp32 operations under a directed rounding mode (the reference switches MXCSR, ParallelMath.h:71-102) ----
#if defined(__CUDA_ARCH__)
    CVTT_HD float fmul_ru(float a, float b) { return __fmul_ru(a, b); }
    CVTT_HD float fdiv_ru(float a, float b) { return __fdiv_ru(a, b); }
    CVTT_HD float fadd_ru(float a, float b) { return __fadd_ru(a, b); }
    CVTT_HD int f2i_ru(float a) { return __float2int_ru(a); }
    CVTT_HD int f2i_rn(float a) { return __float2int_rn(a); }
#else
    inline float host_directed(int op, float a, float b)
    {
        const unsigned csr = _mm_getcsr();
        _mm_setcsr((csr & ~_MM_ROUND_MASK) | _MM_ROUND_UP);
        volatile float va = a, vb = b;
        volatile float r = (op == 0) ? va * vb : ((op == 1) ? va / vb : va + vb);
        _mm_setcsr(csr);
        return r;
    }
    inline float fmul_ru(float a, float b) { return host_directed(0, a, b); }
    inline float fdiv_ru(float a, float b) { return host_directed(1, a, b); }
    inline float fadd_ru(float a, float b) { return host_directed(2, a, b); }
    inline int f2i_ru(float a)
    {
        const unsigned csr = _mm_getcsr();
        _mm_setcsr((csr & ~_MM_ROUND_MASK) | _MM_ROUND_UP);
        volatile float va = a;
        const int r = _mm_cvtss_si32(_mm_set_ss(va));
        _mm_setcsr(csr);
        return r;
    }
    inline int f2i_rn(float a) { return _mm_cvtss_si32(_mm_set_ss(a)); }
#endif

    // _mm_max_ps(_mm_min_ps(v, hi), lo), ParallelMath.h:561-567
    CVTT_HD float sse_clamp(float v, float lo, float hi) { return sse_max(sse_min(v, hi), lo); }

    // ParallelMath::TwosCLHalfToFloat (ParallelMath.h:1012-1041): works on the raw 16 bits (sign | exponent | mantissa)
    CVTT_HD float twoscl_half_to_float(int v)
    {
        const uint32_t u = (uint32_t)v & 0xffffu;
        const uint32_t sign = (u & 0x8000u) << 16;
        const float f = as_float(sign | (((u & 0x7fffu) << 13) + 0x38000000u));
        return ((u & 0x7c00u) == 0) ? fsub(f, as_float(sign | 0x38000000u)) : f;
    }

    // UnscaleHDRValueUnsigned (BC67.cpp:784-787) of a 16-bit interpolated value
    CVTT_HD int unscale_hdr_unsigned(int v16) { return packs_s16((v16 * 31) >> 6); }

    // UnscaleHDRValueSigned (BC67.cpp:766-782): result is sign | magnitude
    CVTT_HD int unscale_hdr_signed(int v)
    {
        const bool negative = v < 0;
        const int absComp = wrap_u16(negative ? -v : v);
        const int scaled = packs_s16((absComp * 31) >> 5);
        return wrap_s16(scaled | (negative ? 0x8000 : 0));
    }

    // QuantizeSingleEndpointElementSigned / Unsigned (BC67.cpp:2425-2446).  The reference evaluates elem * 32 / 31 (signed) or
    // elem * 64 / 31 (unsigned) in fp32 under MXCSR round-up and converts with round-up again (RoundAndConvertToU15 / U16,
    // ParallelMath.h:923-945).  For the admissible inputs (0..31743) that is exactly ceil(N / 31) with N = elem * 32 or
    // elem * 64: N < 2^21, so a non-integer quotient is at least 1/31 below the next integer while the two upward roundings
    // add less than 2^-7.  tests/test_bc6h_host.py checks all 31744 inputs against the fp32 formulation
    // (bc6h_quantize_element_reference), which is what the reference executes.
    template<bool SIGNED>
    CVTT_HD int bc6h_quantize_element(int elem, int precision)
    {
        if (SIGNED)
        {
            const bool negative = elem < 0;
            const int absElem = negative ? -elem : elem;
            const int q = ((absElem * 32 + 30) / 31) >> (16 - precision);
            return negative ? -q : q;
        }
        else
            return ((elem * 64 + 30) / 31) >> (16 - precision);
    }

    template<bool SIGNED>
    CVTT_HD int bc6h_quantize_element_reference(int elem, int precision)
    {
        if (SIGNED)
        {
            const bool negative = elem < 0;
            const int absElem = negative ? -elem : elem;
            const float f = fdiv_ru(fmul_ru((float)absElem, 32.0f), 31.0f);
            const int q = wrap_u16(packs_s16(f2i_ru(f))) >> (16 - precision);
            return negative ? -q : q;
        }
        else
        {
            const float f = sse_min(fdiv_ru(fmul_ru((float)elem, 64.0f), 31.0f), 65535.0f);
            const int expanded = wrap_u16(packs_s16(f2i_ru(fadd_ru(f, -32768.0f)))) ^ 0x8000;
            return expanded >> (16 - precision);
        }
    }

    // unq: the interpolation endpoint (int16 for signed, uint16 for unsigned); fin: the colour-space endpoint
    template<bool SIGNED>
    CVTT_HD void bc6h_unquantize_element(int comp, int precision, int &unq, int &fin)
    {
        if (SIGNED)
        {
            // UnquantizeSingleEndpointElementSigned, BC67.cpp:2448-2482
            const bool negative = comp < 0;
            const int absComp = wrap_u16(negative ? -comp : comp);
            int absUnq;
            if (precision >= 16)
            {
                unq = comp;
                absUnq = absComp;
            }
            else
            {
                const int maxCompMinusOne = (1 << (precision - 1)) - 2;
                absUnq = wrap_u16((absComp << (16 - precision)) + (0x4000 >> (precision - 1)));
                if (comp == 0)
                    absUnq = 0;
                if (maxCompMinusOne < comp)
                    absUnq = 0x7fff;
                unq = negative ? wrap_s16(-absUnq) : wrap_s16(absUnq);
            }
            const int funq = packs_s16((absUnq * 31) >> 5);
            fin = negative ? wrap_s16(-funq) : funq;
        }
        else
        {
            // UnquantizeSingleEndpointElementUnsigned, BC67.cpp:2484-2501
            unq = wrap_u16(comp);
            if (precision < 15)
            {
                const int maxCompMinusOne = (1 << precision) - 2;
                unq = wrap_u16((comp << (16 - precision)) + (0x8000 >> precision));
                if (comp == 0)
                    unq = 0;
                if (maxCompMinusOne < wrap_s16(comp))
                    unq = 0xffff;
            }
            fin = wrap_u16((unq * 31) >> 6);
        }
    }

    // IndexSelectorHDR::ReconstructHDR*Uninverted (IndexSelectorHDR.h:34-66): e0/e1 are the interpolation endpoints
    template<bool SIGNED>
    CVTT_HD int bc6h_reconstruct(int weight, int e0, int e1)
    {
        if (SIGNED)
        {
            const int pixel32 = ((64 - weight) * e0 + weight * e1 + 32) >> 6;
            return unscale_hdr_signed(packs_s16(pixel32));
        }
        else
        {
            const int pixel31 = ((64 - weight) * e0 + weight * e1 + 32) >> 6;
            return unscale_hdr_unsigned(wrap_u16(pixel31));
        }
    }

    // EvaluatePartitionedLegality / EvaluateSingleLegality (BC67.cpp:2597-2663).  q[s][e][ch] are the quantised
    // endpoints as 16-bit patterns; enc receives what is written to the block.
    CVTT_HD bool bc6h_legality(int numSubsets, const int q[2][2][3], int aPrec, const uint8_t *bPrec, bool transformed, int enc[2][2][3])
    {
        bool legal = true;
        const int aMask = (1 << aPrec) - 1;
        for (int ch = 0; ch < 3; ch++)
        {
            for (int s = 0; s < numSubsets; s++)
                for (int e = 0; e < 2; e++)
                    enc[s][e][ch] = wrap_u16(q[s][e][ch]);
            if (transformed)
                for (int s = 0; s < numSubsets; s++)
                    for (int e = 0; e < 2; e++)
                    {
                        if (e == 0 && s == 0)
                            continue;
                        const int bReduced = enc[s][e][ch] & aMask;
                        const int lost = 16 - bPrec[ch];
                        const int diff = wrap_s16(enc[s][e][ch] - enc[0][0][ch]);
                        const int delta = wrap_s16(wrap_s16(diff << lost) >> lost);      // TruncateToPrecisionSigned
                        enc[s][e][ch] = wrap_u16(delta);
                        const int reconstructed = wrap_u16(delta + enc[0][0][ch]) & aMask;
                        legal = legal && (reconstructed == bReduced);
                    }
        }
        return legal;
    }

    // ---------------------------------------------------------------------------------------------------------
    // EndpointSelector<3, 8> (ConvectionKernels_EndpointSelector.h:13-150) over the pixels of `mask`, ascending
    // pixel order, unit pixel weights; L.pw_at(px) is the pre-weighted pixel.
    template<class Lane>
    CVTT_HD void endpoint_selector3_masked(const Lane &L, uint32_t mask, int n, const float *wv, float *base, float *offs)
    {
        float centroid[3] = { 0.0f, 0.0f, 0.0f }, cov[6] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
        for (uint32_t m = mask; m; m &= m - 1)
        {
            const F4 p = L.pw_at(ctz32(m));
            centroid[0] = fadd(centroid[0], p.x);
            centroid[1] = fadd(centroid[1], p.y);
            centroid[2] = fadd(centroid[2], p.z);
        }
        {
            const float denom = (float)n;
            for (int ch = 0; ch < 3; ch++)
                centroid[ch] = fdiv(centroid[ch], denom);
        }
        for (uint32_t m = mask; m; m &= m - 1)
        {
            const F4 p = L.pw_at(ctz32(m));
            const float d[3] = { fsub(p.x, centroid[0]), fsub(p.y, centroid[1]), fsub(p.z, centroid[2]) };
            int index = 0;
#pragma unroll
            for (int row = 0; row < 3; row++)
#pragma unroll
                for (int col = 0; col <= row; col++)
                {
                    cov[index] = fadd(cov[index], fmul(d[row], d[col]));
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
        for (uint32_t m = mask; m; m &= m - 1)
        {
            const F4 p = L.pw_at(ctz32(m));
            float dist = fadd(0.0f, fmul(dir[0], fsub(p.x, centroid[0])));
            dist = fadd(dist, fmul(dir[1], fsub(p.y, centroid[1])));
            dist = fadd(dist, fmul(dir[2], fsub(p.z, centroid[2])));
            minDist = sse_min(minDist, dist);
            maxDist = sse_max(maxDist, dist);
        }
        for (int ch = 0; ch < 3; ch++)
        {
            const float mn = fadd(centroid[ch], fmul(dir[ch], minDist));
            const float mx = fadd(centroid[ch], fmul(dir[ch], maxDist));
            base[ch] = fdiv(mn, wv[ch]);
            offs[ch] = fdiv(fsub(mx, mn), wv[ch]);
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // Per-lane pixel storage: planes of 32-bit words, element i of a plane at [i * STRIDE] (conflict-free in shared memory).
    //   lin[px * 3 + ch] = TwosCLHalfToFloat(pixel)           (the value the errors compare against)
    //   pw[px * 3 + ch]  = (float)pixel * channelWeight       (EndpointSelector / EndpointRefiner input)
    //   pix[px * 2 + k]  = pixel ch0 | ch1 << 16, pixel ch2   (only read, and only allocated, with BC6H_FastIndexing)
    template<int STRIDE, bool WITH_PIX>
    struct BC6HLane
    {
        float *lin;
        float *pw;
        uint32_t *pix;

        CVTT_HD F4 lin_at(int px) const
        {
            F4 r;
            r.x = lin[(px * 3 + 0) * STRIDE];
            r.y = lin[(px * 3 + 1) * STRIDE];
            r.z = lin[(px * 3 + 2) * STRIDE];
            r.w = WITH_PIX ? as_float(pix[(px * 2 + 0) * STRIDE]) : 0.0f;
            return r;
        }
        CVTT_HD F4 pw_at(int px) const
        {
            F4 r;
            r.x = pw[(px * 3 + 0) * STRIDE];
            r.y = pw[(px * 3 + 1) * STRIDE];
            r.z = pw[(px * 3 + 2) * STRIDE];
            r.w = WITH_PIX ? as_float(pix[(px * 2 + 1) * STRIDE]) : 0.0f;
            return r;
        }
        CVTT_HD void store(int px, const F4 &l, const F4 &q) const
        {
            lin[(px * 3 + 0) * STRIDE] = l.x;
            lin[(px * 3 + 1) * STRIDE] = l.y;
            lin[(px * 3 + 2) * STRIDE] = l.z;
            pw[(px * 3 + 0) * STRIDE] = q.x;
            pw[(px * 3 + 1) * STRIDE] = q.y;
            pw[(px * 3 + 2) * STRIDE] = q.z;
            if (WITH_PIX)
            {
                pix[(px * 2 + 0) * STRIDE] = as_uint(l.w);
                pix[(px * 2 + 1) * STRIDE] = as_uint(q.w);
            }
        }
    };

    // Loads one PixelBlockF16 (int16 [16][4], alpha ignored) and converts it as BC6HComputer::Pack does (BC67.cpp:2691-2715)
    template<bool SIGNED, class Lane>
    CVTT_HD void bc6h_load_pixel(const BC6HParams &P, const Lane &L, int px, int r, int g, int b)
    {
        int c[3] = { r, g, b };
        F4 lin, pw;
        float lf[3], pf[3];
        for (int ch = 0; ch < 3; ch++)
        {
            int v = wrap_s16(c[ch]);
            if (SIGNED)
            {
                if (v < 0)
                    v = -(v & 32767);
                v = (v > -31743) ? v : -31743;
            }
            else
                v = (v > 0) ? v : 0;
            v = (v < 31743) ? v : 31743;
            c[ch] = v;
            lf[ch] = twoscl_half_to_float(v);
            pf[ch] = fmul((float)v, P.w[ch]);
        }
        lin.x = lf[0]; lin.y = lf[1]; lin.z = lf[2];
        lin.w = as_float(((uint32_t)c[0] & 0xffffu) | ((uint32_t)c[1] << 16));
        pw.x = pf[0]; pw.y = pf[1]; pw.z = pf[2];
        pw.w = as_float((uint32_t)c[2] & 0xffffu);
        L.store(px, lin, pw);
    }

    struct BC6HBest
    {
        float error;
        int mode, partition;
        int ep[2][2][3];          // what goes into the block (16-bit patterns)
        uint32_t idx[2];          // 16 x 4 bits
    };

    // One meta round's outcome for one subset
    struct BC6HSelector
    {
        int e[2][3];              // interpolation endpoints (unquantised)
        bool inverted;
    };

    // ---------------------------------------------------------------------------------------------------------
    // All trials of one (precision, partition): fills the meta arrays, then the commit scan (BC67.cpp:2790-2988).
    template<bool SIGNED, bool FAST, int RANGE, int STRIDE, class Vote>
    CVTT_HD void bc6h_partition(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE, FAST> &L, Vote &vote, bool partitioned, int aPrec, int p,
        const float *ufepBase /* [2][3] */, const float *ufepOffs /* [2][3] */, BC6HBest &best)
    {
        enum { kMaxTweak = 4, kMaxRefine = 3, kMeta = 12 };
        const int numSubsets = partitioned ? 2 : 1;
        const uint32_t partitionMask = partitioned ? T.partitionMask[p] : 0u;
        const int weightRecip = (RANGE == 8) ? 4681 : 2185;           // g_weightReciprocals[range], IndexSelector.cpp:43-62
        const float maxV = (float)(RANGE - 1);

        // zero-initialised like the reference build's automatics
        uint32_t metaEP[kMeta][2][3];      // [round][subset][ch]: quantised ep0 | ep1 << 16
        uint32_t metaIdx[kMeta][2];
        float metaErr[kMeta][2];
        for (int r = 0; r < kMeta; r++)
            for (int s = 0; s < 2; s++)
            {
                metaEP[r][s][0] = metaEP[r][s][1] = metaEP[r][s][2] = 0;
                metaIdx[r][s] = 0;
                metaErr[r][s] = 0.0f;
            }
        uint32_t roundValid = 0xffffffu;   // bit (round * 2 + subset); uniform over the 8 lanes of a group

        for (int subset = 0; subset < numSubsets; subset++)
        {
            const uint32_t mask = subset ? partitionMask : (~partitionMask & 0xffffu);
            int n = 0;
            float sumV[3] = { 0.0f, 0.0f, 0.0f };
            for (uint32_t m = mask; m; m &= m - 1)
            {
                const F4 q = L.pw_at(ctz32(m));
                sumV[0] = fadd(sumV[0], q.x);
                sumV[1] = fadd(sumV[1], q.y);
                sumV[2] = fadd(sumV[2], q.z);
                n++;
            }
            const int fixupIndex = (subset == 0) ? 0 : T.fixup[p];

            for (int tweak = 0; tweak < kMaxTweak; tweak++)
            {
                // EndpointRefiner sums of the previous pass; `contributed` is false after a skipped round (BC67.cpp:2843)
                float tv[3] = { 0.0f, 0.0f, 0.0f }, tt = 0.0f, ts = 0.0f;
                bool contributed = false;

                for (int refinePass = 0; refinePass < kMaxRefine; refinePass++)
                {
                    const int metaRound = tweak * kMaxRefine + refinePass;
                    if (tweak >= P.tweakRounds || refinePass >= P.refineRounds)
                    {
                        roundValid &= ~(1u << (metaRound * 2 + subset));
                        continue;
                    }

                    int ec[2][3];
                    if (refinePass == 0)
                    {
                        // UnfinishedEndpoints::FinishHDRSigned / Unsigned
                        const float tf[2] = { P.tweak[RANGE == 16][tweak][0], P.tweak[RANGE == 16][tweak][1] };
                        for (int ch = 0; ch < 3; ch++)
                            for (int epi = 0; epi < 2; epi++)
                            {
                                const float f = sse_clamp(fadd(ufepBase[subset * 3 + ch], fmul(ufepOffs[subset * 3 + ch], tf[epi])), SIGNED ? -31743.0f : 0.0f, 31743.0f);
                                ec[epi][ch] = packs_s16(f2i_rn(f));
                            }
                    }
                    else
                    {
                        // EndpointRefiner::GetRefinedEndpoints (EndpointRefiner.h:99-142) + GetRefinedEndpointsHDR (:160-175)
                        float wN = contributed ? (float)n : 0.0f;
                        safe_denominator(wN);
                        const float wRcp = contributed ? P.rcpN[n] : P.rcpN[1];
                        const float sv[3] = { contributed ? sumV[0] : 0.0f, contributed ? sumV[1] : 0.0f, contributed ? sumV[2] : 0.0f };
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
                            const float f0 = fmul(p1, P.rcpW[ch]), f1 = fmul(p2, P.rcpW[ch]);
                            ec[0][ch] = packs_s16(f2i_rn(sse_clamp(f0, SIGNED ? -31743.0f : 0.0f, 31743.0f)));
                            ec[1][ch] = packs_s16(f2i_rn(sse_clamp(f1, SIGNED ? -31743.0f : 0.0f, 31743.0f)));
                        }
                    }
                    // refiners[subset].Init
                    tv[0] = tv[1] = tv[2] = 0.0f;
                    tt = ts = 0.0f;
                    contributed = false;

                    // QuantizeEndpointsSigned / Unsigned (BC67.cpp:2503-2595)
                    int q[2][3], unq[2][3], fin[2][3];
                    for (int epi = 0; epi < 2; epi++)
                        for (int ch = 0; ch < 3; ch++)
                        {
                            q[epi][ch] = bc6h_quantize_element<SIGNED>(ec[epi][ch], aPrec);
                            bc6h_unquantize_element<SIGNED>(q[epi][ch], aPrec, unq[epi][ch], fin[epi][ch]);
                        }

                    // IndexSelector::Init (fast indexing only needs it) and IndexSelectorHDR::InitHDR
                    float origin[3] = { 0, 0, 0 }, axis[3] = { 0, 0, 0 };
                    float reconW[RANGE][3];       // m_reconstructedInterpolators
                    float reconLin[RANGE][3];     // TwosCLHalfToFloat of the reconstructed colours (for the error)
                    if (FAST)
                    {
                        float dW[3];
                        for (int ch = 0; ch < 3; ch++)
                        {
                            origin[ch] = (float)fin[0][ch];
                            dW[ch] = fmul(fsub((float)fin[1][ch], origin[ch]), P.w[ch]);
                        }
                        float lenSq = fmul(dW[0], dW[0]);
                        lenSq = fadd(lenSq, fmul(dW[1], dW[1]));
                        lenSq = fadd(lenSq, fmul(dW[2], dW[2]));
                        safe_denominator(lenSq);
                        const float mdl = fdiv(maxV, lenSq);
                        for (int ch = 0; ch < 3; ch++)
                            axis[ch] = fmul(fmul(dW[ch], P.w[ch]), mdl);
                    }
                    else
                    {
#pragma unroll
                        for (int i = 0; i < RANGE; i++)
                        {
                            const int weight = wrap_u16(weightRecip * i + 256) >> 9;
#pragma unroll
                            for (int ch = 0; ch < 3; ch++)
                            {
                                const float f = twoscl_half_to_float(bc6h_reconstruct<SIGNED>(weight, unq[0][ch], unq[1][ch]));
                                reconLin[i][ch] = f;
                                reconW[i][ch] = fmul(f, P.w[ch]);
                            }
                        }
                    }

                    // index selection; returns the uninverted interpolator number
                    auto selectIndex = [&](int px, float *linOut) -> int
                    {
                        const F4 lp = L.lin_at(px);
                        if (FAST)
                        {
                            const F4 pq = L.pw_at(px);
                            const uint32_t w0 = as_uint(lp.w), w1 = as_uint(pq.w);
                            const float c0 = (float)wrap_s16((int)(w0 & 0xffffu)), c1 = (float)wrap_s16((int)(w0 >> 16)), c2 = (float)wrap_s16((int)(w1 & 0xffffu));
                            float dist = fmul(fsub(c0, origin[0]), axis[0]);
                            dist = fadd(dist, fmul(fsub(c1, origin[1]), axis[1]));
                            dist = fadd(dist, fmul(fsub(c2, origin[2]), axis[2]));
                            return packs_s16(f2i_rn(sse_clamp(dist, 0.0f, maxV)));
                        }
                        else
                        {
                            // SelectIndexHDRSlow (IndexSelectorHDR.h:125-139)
                            const float pl[3] = { fmul(lp.x, P.w[0]), fmul(lp.y, P.w[1]), fmul(lp.z, P.w[2]) };
                            int index = 0;
                            float bestError = 0.0f, l0 = reconLin[0][0], l1 = reconLin[0][1], l2 = reconLin[0][2];
#pragma unroll
                            for (int i = 0; i < RANGE; i++)
                            {
                                const float d0 = fsub(pl[0], reconW[i][0]), d1 = fsub(pl[1], reconW[i][1]), d2 = fsub(pl[2], reconW[i][2]);
                                const float error = fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2));
                                if (i == 0)
                                    bestError = error;
                                else
                                {
                                    const bool better = error < bestError;
                                    index = better ? i : index;
                                    l0 = better ? reconLin[i][0] : l0;      // selects, so the tables stay in registers
                                    l1 = better ? reconLin[i][1] : l1;
                                    l2 = better ? reconLin[i][2] : l2;
                                    bestError = sse_min(bestError, error);
                                }
                            }
                            linOut[0] = l0;
                            linOut[1] = l1;
                            linOut[2] = l2;
                            return index;
                        }
                    };

                    // fix-up index and conditional inversion
                    float fixLin[3] = { 0, 0, 0 };
                    int fixRaw = selectIndex(fixupIndex, fixLin);
                    const bool invert = (RANGE / 2 - 1) < fixRaw;
                    int fixIndexStored = invert ? (RANGE - 1 - fixRaw) : fixRaw;
                    if (invert)
                        for (int ch = 0; ch < 3; ch++)
                        {
                            const int t = q[0][ch];
                            q[0][ch] = q[1][ch];
                            q[1][ch] = t;
                        }
                    const uint32_t qp[3] = { (uint32_t)wrap_u16(q[0][0]) | ((uint32_t)wrap_u16(q[1][0]) << 16),
                                             (uint32_t)wrap_u16(q[0][1]) | ((uint32_t)wrap_u16(q[1][1]) << 16),
                                             (uint32_t)wrap_u16(q[0][2]) | ((uint32_t)wrap_u16(q[1][2]) << 16) };
                    metaEP[metaRound][subset][0] = qp[0];
                    metaEP[metaRound][subset][1] = qp[1];
                    metaEP[metaRound][subset][2] = qp[2];
                    // indexes[fixupIndex] = index (the array is shared by both subsets of the round)
                    uint32_t roundIdx[2] = { 0, 0 }, roundMask[2] = { 0, 0 };
                    {
                        const int sh = 4 * (fixupIndex & 7);
                        roundIdx[fixupIndex >> 3] = (uint32_t)fixIndexStored << sh;
                        roundMask[fixupIndex >> 3] = 15u << sh;
                    }

                    // a round that repeats an earlier round's endpoints on all eight lanes is dropped (BC67.cpp:2853-2877)
                    if (metaRound > 0)
                    {
                        bool anySame = false;
                        for (int prev = 0; prev < metaRound; prev++)
                            anySame = anySame || (metaEP[prev][subset][0] == qp[0] && metaEP[prev][subset][1] == qp[1] && metaEP[prev][subset][2] == qp[2]);
                        if (vote.all(anySame))
                        {
                            roundValid &= ~(1u << (metaRound * 2 + subset));
                            for (int h = 0; h < 2; h++)
                                metaIdx[metaRound][h] = (metaIdx[metaRound][h] & ~roundMask[h]) | roundIdx[h];
                            continue;
                        }
                    }

                    float subsetError = 0.0f;
                    const bool refineNext = (refinePass != P.refineRounds - 1);
                    for (uint32_t m = mask; m; m &= m - 1)
                    {
                        const int px = ctz32(m);
                        float rl[3] = { fixLin[0], fixLin[1], fixLin[2] };
                        int raw, index;
                        if (px == fixupIndex)
                        {
                            raw = fixRaw;
                            index = fixIndexStored;
                        }
                        else
                        {
                            raw = selectIndex(px, rl);
                            index = invert ? (RANGE - 1 - raw) : raw;
                            const int sh = 4 * (px & 7);
                            if (px < 8)
                            {
                                roundIdx[0] |= (uint32_t)index << sh;
                                roundMask[0] |= 15u << sh;
                            }
                            else
                            {
                                roundIdx[1] |= (uint32_t)index << sh;
                                roundMask[1] |= 15u << sh;
                            }
                        }

                        const F4 lp = L.lin_at(px);
                        float error = 0.0f;
                        if (FAST)
                        {
                            // ReconstructHDR* + ComputeErrorHDRFast (BCCommon.h:45-61, SqDiffSInt16 ParallelMath.h:996-1010)
                            const F4 pq = L.pw_at(px);
                            const uint32_t w0 = as_uint(lp.w), w1 = as_uint(pq.w);
                            const int orig[3] = { wrap_s16((int)(w0 & 0xffffu)), wrap_s16((int)(w0 >> 16)), wrap_s16((int)(w1 & 0xffffu)) };
                            const int weight = wrap_u16(weightRecip * raw + 256) >> 9;
                            for (int ch = 0; ch < 3; ch++)
                            {
                                const int rc = bc6h_reconstruct<SIGNED>(weight, unq[0][ch], unq[1][ch]);
                                const int hi = rc > orig[ch] ? rc : orig[ch], lo = rc > orig[ch] ? orig[ch] : rc;
                                const uint32_t diffU = (uint32_t)wrap_u16(hi - lo);
                                const float sq = (float)(int32_t)(diffU * diffU);
                                error = (P.flags & kFlag_Uniform) ? fadd(error, sq) : fadd(error, fmul(sq, P.wSq[ch]));
                            }
                        }
                        else
                        {
                            // ComputeErrorHDRSlow (BCCommon.h:63-79): SqDiff2CL(reconstructed, original)
                            const float ol[3] = { lp.x, lp.y, lp.z };
                            for (int ch = 0; ch < 3; ch++)
                            {
                                const float diff = fsub(rl[ch], ol[ch]);
                                const float sq = fmul(diff, diff);
                                error = (P.flags & kFlag_Uniform) ? fadd(error, sq) : fadd(error, fmul(sq, P.wSq[ch]));
                            }
                        }
                        subsetError = fadd(subsetError, error);

                        if (refineNext)
                        {
                            // EndpointRefiner::ContributeUnweightedPW (EndpointRefiner.h:78-92)
                            const F4 pq = L.pw_at(px);
                            const float t = fmul((float)index, 1.0f / maxV);
                            tv[0] = fadd(tv[0], fmul(t, pq.x));
                            tv[1] = fadd(tv[1], fmul(t, pq.y));
                            tv[2] = fadd(tv[2], fmul(t, pq.z));
                            tt = fadd(tt, fmul(t, t));
                            ts = fadd(ts, t);
                            contributed = true;
                        }
                    }
                    metaErr[metaRound][subset] = subsetError;
                    for (int h = 0; h < 2; h++)
                        metaIdx[metaRound][h] = (metaIdx[metaRound][h] & ~roundMask[h]) | roundIdx[h];
                }
            }
        }

        // Combine the rounds of the two subsets; a combination that improves on the best so far is committed with the
        // modes of this (partitioned, precision) that can encode it (BC67.cpp:2915-2985).
        const int numMeta1 = partitioned ? kMeta : 1;

        // Most of the 12 x 12 combinations cannot improve on the best so far.  fl(a + b) is monotonic in b, so a subset-0 round
        // whose error plus the SMALLEST valid subset-1 error is not below the lane's best has no partner that is; best.error
        // only falls during the scan, which keeps the test conservative.  Rows that no lane of the warp can use are skipped
        // with their twelve votes (combinations without a candidate lane are no-ops in the reference's loop as well).
        float minErr1 = partitioned ? FLT_MAX : 0.0f;
        if (partitioned)
            for (int meta1 = 0; meta1 < kMeta; meta1++)
                if ((roundValid >> (meta1 * 2 + 1)) & 1)
                    minErr1 = sse_min(minErr1, metaErr[meta1][1]);

        for (int meta0 = 0; meta0 < kMeta; meta0++)
        {
            // roundValid is uniform over a group but not over the warp: only skip what every group of the warp skips, the
            // votes below must be executed by all lanes
            const bool valid0 = ((roundValid >> (meta0 * 2)) & 1) != 0;
            const bool rowPossible = valid0 && ((partitioned ? fadd(metaErr[meta0][0], minErr1) : metaErr[meta0][0]) < best.error);
            if (!vote.warp_any(rowPossible))
                continue;
            for (int meta1 = 0; meta1 < numMeta1; meta1++)
            {
                float combinedError = metaErr[meta0][0];
                bool valid = valid0;
                if (partitioned)
                {
                    valid = valid && ((roundValid >> (meta1 * 2 + 1)) & 1) != 0;
                    combinedError = fadd(combinedError, metaErr[meta1][1]);
                }
                const bool errorBetter = valid && (combinedError < best.error);
                bool needsCommit = errorBetter;
                if (!vote.warp_any(errorBetter))        // no lane of the warp: no group either, one vote instead of two
                    continue;
                bool groupActive = vote.any(errorBetter);

                int q[2][2][3];
                for (int ch = 0; ch < 3; ch++)
                {
                    q[0][0][ch] = (int)(metaEP[meta0][0][ch] & 0xffffu);
                    q[0][1][ch] = (int)(metaEP[meta0][0][ch] >> 16);
                    q[1][0][ch] = (int)(metaEP[meta1][1][ch] & 0xffffu);
                    q[1][1][ch] = (int)(metaEP[meta1][1][ch] >> 16);
                }

                for (int mode = 0; mode < 14; mode++)
                {
                    const uint8_t *mi = T.modes[mode];
                    if ((mi[1] != 0) != partitioned || (int)mi[3] != aPrec)
                        continue;
                    bool legalAndBetter = false;
                    if (groupActive && errorBetter)
                    {
                        int enc[2][2][3];
                        for (int s = 0; s < 2; s++)
                            for (int e = 0; e < 2; e++)
                                enc[s][e][0] = enc[s][e][1] = enc[s][e][2] = 0;
                        if (bc6h_legality(numSubsets, q, aPrec, mi + 4, mi[2] != 0, enc))
                        {
                            legalAndBetter = true;
                            best.error = combinedError;
                            best.mode = mode;
                            best.partition = p;
                            for (int s = 0; s < numSubsets; s++)
                                for (int e = 0; e < 2; e++)
                                    for (int ch = 0; ch < 3; ch++)
                                        best.ep[s][e][ch] = enc[s][e][ch];
                            // pixels of subset 0 take the indexes of meta0, pixels of subset 1 those of meta1
                            uint32_t sel[2];
                            for (int h = 0; h < 2; h++)
                            {
                                uint32_t nib = 0;
                                const uint32_t pm = (partitionMask >> (8 * h)) & 0xffu;
                                for (int k = 0; k < 8; k++)
                                    if ((pm >> k) & 1)
                                        nib |= 15u << (4 * k);
                                sel[h] = nib;
                                best.idx[h] = (metaIdx[meta0][h] & ~sel[h]) | (metaIdx[meta1][h] & sel[h]);
                            }
                        }
                    }
                    needsCommit = needsCommit && !legalAndBetter;
                    const bool groupNeeds = vote.any(needsCommit);
                    groupActive = groupActive && groupNeeds;
                    if (!vote.warp_any(groupActive))
                        break;
                }
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // The whole search for one block and the bit packing tail (BC67.cpp:2990-3050).
    template<bool SIGNED, bool FAST, int STRIDE, class Vote>
    CVTT_HD void bc6h_encode_block(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE, FAST> &L, Vote &vote, uint32_t out[4])
    {
        BC6HBest best;
        best.error = FLT_MAX;
        best.mode = 0;
        best.partition = 0;
        for (int s = 0; s < 2; s++)
            for (int e = 0; e < 2; e++)
                best.ep[s][e][0] = best.ep[s][e][1] = best.ep[s][e][2] = 0;
        best.idx[0] = best.idx[1] = 0;

        // endpoint fits of the 32 partitions and of the whole block (BC67.cpp:2739-2774)
        float ufepBase[33][6], ufepOffs[33][6];
        for (int p = 0; p < 32; p++)
            for (int subset = 0; subset < 2; subset++)
            {
                const uint32_t mask = subset ? T.partitionMask[p] : ((uint32_t)~T.partitionMask[p] & 0xffffu);
                int n = 0;
                for (uint32_t m = mask; m; m &= m - 1)
                    n++;
                endpoint_selector3_masked(L, mask, n, P.w, ufepBase[p] + subset * 3, ufepOffs[p] + subset * 3);
            }
        endpoint_selector3_masked(L, 0xffffu, 16, P.w, ufepBase[32], ufepOffs[32]);
        for (int ch = 0; ch < 3; ch++)
            ufepBase[32][3 + ch] = ufepOffs[32][3 + ch] = 0.0f;

        // g_hdrModesExistForPrecision (BC67.cpp:144-149)
        const uint32_t existSingle = (1u << 10) | (1u << 11) | (1u << 12) | (1u << 16);
        const uint32_t existPartitioned = (1u << 6) | (1u << 7) | (1u << 8) | (1u << 9) | (1u << 10) | (1u << 11);
        for (int partitionedInt = 0; partitionedInt < 2; partitionedInt++)
            for (int aPrec = 16; aPrec >= 0; aPrec--)
            {
                if (!(((partitionedInt ? existPartitioned : existSingle) >> aPrec) & 1))
                    continue;
                if (partitionedInt)
                {
                    for (int p = 0; p < 32; p++)
                        bc6h_partition<SIGNED, FAST, 8, STRIDE>(P, T, L, vote, true, aPrec, p, ufepBase[p], ufepOffs[p], best);
                }
                else
                    bc6h_partition<SIGNED, FAST, 16, STRIDE>(P, T, L, vote, false, aPrec, 0, ufepBase[32], ufepOffs[32], best);
            }

        // header: scatter the endpoint fields into the mode's bit layout, then the indexes (PackingVector::Pack does not
        // mask its argument; the fix-up indexes have their top bit clear by construction)
        const uint8_t *mi = T.modes[best.mode];
        const bool partitioned = mi[1] != 0;
        const int headerBits = partitioned ? 82 : 65;
        uint32_t fields[14];
        fields[0] = mi[0];
        fields[1] = (uint32_t)best.partition;
        for (int ch = 0; ch < 3; ch++)
        {
            fields[2 + ch * 4 + 0] = (uint32_t)best.ep[0][0][ch] & 0xffffu;
            fields[2 + ch * 4 + 1] = (uint32_t)best.ep[0][1][ch] & 0xffffu;
            fields[2 + ch * 4 + 2] = (uint32_t)best.ep[1][0][ch] & 0xffffu;
            fields[2 + ch * 4 + 3] = (uint32_t)best.ep[1][1][ch] & 0xffffu;
        }
        uint32_t v[4] = { 0, 0, 0, 0 };
        for (int i = 0; i < headerBits; i++)
        {
            const uint32_t code = T.headerBits[best.mode][i];
            const uint32_t bit = (fields[code >> 4] >> (code & 15u)) & 1u;
            v[i >> 5] |= bit << (i & 31);
        }
        int offset = headerBits;
        const int fixupIndex1 = partitioned ? T.fixup[best.partition] : 0;
        const int indexBits = partitioned ? 3 : 4;
        for (int px = 0; px < 16; px++)
        {
            const uint32_t index = (best.idx[px >> 3] >> (4 * (px & 7))) & 15u;
            const int bits = (px == 0 || px == fixupIndex1) ? indexBits - 1 : indexBits;
            const int vOffset = offset >> 5, bitOffset = offset & 31;
            v[vOffset] |= index << bitOffset;
            const int overflowBits = bitOffset + bits - 32;
            if (overflowBits > 0)
                v[vOffset + 1] |= index >> (bits - overflowBits);
            offset += bits;
        }
        out[0] = v[0];
        out[1] = v[1];
        out[2] = v[2];
        out[3] = v[3];
    }
}
