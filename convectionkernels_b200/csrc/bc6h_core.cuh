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

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif
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
        int prune;                         // skip work that provably cannot change the result (bc6h_partition); 0 only for A/B timing
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
            // e0, e1 <= 65535, so pixel31 <= 65535 and pixel31 * 31 >> 6 <= 31743: neither the 16-bit wrap nor the saturating
            // pack of UnscaleHDRValueUnsigned can trigger
            const int pixel31 = ((64 - weight) * e0 + weight * e1 + 32) >> 6;
            return (pixel31 * 31) >> 6;
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
    // ParallelMath::TwosCLHalfToFloat (twoscl_half_to_float above) through the hardware half -> float conversion, for 16-bit
    // sign | exponent | mantissa patterns with exponent below 31 (unsigned pixels, and every reconstructed interpolator:
    // BC67.cpp:766-787 clamp them to 31743).  For normal numbers the reference's bit arithmetic is the IEEE conversion.  For
    // exponent 0 it is not: it builds 2^-15 (1 + m / 1024) and subtracts 2^-15, i.e. m 2^-25 -- HALF the IEEE value m 2^-24 --
    // and the encoder's errors are computed with that value, so it is reproduced (tools/ubench/half_cvt_check.cu compares all
    // patterns on the device).  The pattern 0x8000 gives -0 here and +0 there, which no caller can observe: the values are only
    // multiplied by channel weights, subtracted from other values and squared.
    CVTT_HD float ieee_half_bits_to_float(uint32_t u)
    {
#if defined(__CUDA_ARCH__)
        return __half2float(__ushort_as_half((unsigned short)u));
#else
        const uint32_t sign = (u & 0x8000u) << 16, e = (u >> 10) & 31u, m = u & 0x3ffu;
        if (e == 0)
            return as_float(as_uint((float)m * 5.9604644775390625e-8f) | sign);       // m 2^-24, exact
        return as_float(sign | ((e + 112u) << 23) | (m << 13));
#endif
    }

    CVTT_HD float half_bits_to_float(uint32_t u)
    {
        const float f = ieee_half_bits_to_float(u);
        return (u & 0x7c00u) ? f : fmul(f, 0.5f);
    }

    // Per-lane storage: planes of 32-bit words, element i of a plane at [i * STRIDE] (conflict-free in shared memory).
    //   pw[px * 3 + ch]  = (float)pixel * channelWeight       (EndpointSelector / EndpointRefiner input)
    //   raw[px * 2 + k]  = pixel ch0 | ch1 << 16, pixel ch2   (16-bit patterns: the integers of the fast-indexing path and,
    //                                                          read as half floats, the values the errors compare against)
    //   tab[24]          = linear colours of the current round's interpolators (slow indexing; layout in BC6HIndexer)
    template<int STRIDE>
    struct BC6HLane
    {
        enum { kStride = STRIDE };
        float *pw;
        uint32_t *raw;
        uint32_t *tab;

        CVTT_HD F4 pw_at(int px) const
        {
            F4 r;
            r.x = pw[(px * 3 + 0) * STRIDE];
            r.y = pw[(px * 3 + 1) * STRIDE];
            r.z = pw[(px * 3 + 2) * STRIDE];
            r.w = 0.0f;
            return r;
        }
        CVTT_HD void raw_at(int px, uint32_t &w0, uint32_t &w1) const
        {
            w0 = raw[(px * 2 + 0) * STRIDE];
            w1 = raw[(px * 2 + 1) * STRIDE];
        }
    };

    // The pixel's three values as the reference's TwosCLHalfToFloat sees them.  Signed pixels are two's complement integers
    // (BC67.cpp:2691-2715), which that function reads as sign | 15 raw bits -- for small negative values the "exponent" is 31
    // and the result a large finite number, not what a half conversion gives -- so the signed encoder keeps the literal formula.
    template<bool SIGNED>
    CVTT_HD void bc6h_raw_to_lin(uint32_t w0, uint32_t w1, float *lin)
    {
        if (!SIGNED)
        {
            lin[0] = half_bits_to_float(w0 & 0xffffu);
            lin[1] = half_bits_to_float(w0 >> 16);
            lin[2] = half_bits_to_float(w1 & 0xffffu);
            return;
        }
        lin[0] = twoscl_half_to_float((int)(w0 & 0xffffu));
        lin[1] = twoscl_half_to_float((int)(w0 >> 16));
        lin[2] = twoscl_half_to_float((int)(w1 & 0xffffu));
    }

    CVTT_HD void bc6h_raw_to_int(uint32_t w0, uint32_t w1, int *c)
    {
        c[0] = wrap_s16((int)(w0 & 0xffffu));
        c[1] = wrap_s16((int)(w0 >> 16));
        c[2] = wrap_s16((int)(w1 & 0xffffu));
    }

    // Loads one PixelBlockF16 pixel (int16, alpha ignored) and converts it as BC6HComputer::Pack does (BC67.cpp:2691-2715)
    template<bool SIGNED, class Lane>
    CVTT_HD void bc6h_load_pixel(const BC6HParams &P, const Lane &L, int px, int r, int g, int b)
    {
        int c[3] = { r, g, b };
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
            L.pw[(px * 3 + ch) * Lane::kStride] = fmul((float)v, P.w[ch]);
        }
        L.raw[(px * 2 + 0) * Lane::kStride] = ((uint32_t)c[0] & 0xffffu) | ((uint32_t)c[1] << 16);
        L.raw[(px * 2 + 1) * Lane::kStride] = (uint32_t)c[2] & 0xffffu;
    }

    struct BC6HBest
    {
        float error;
        int mode, partition;
        int ep[2][2][3];          // what goes into the block (16-bit patterns)
        uint32_t q[2][3];         // the committed rounds' quantised endpoints per subset and channel: ep0 | ep1 << 16
        uint32_t inverted;        // bit s: subset s was committed with its endpoints swapped
        bool committed;
    };

    // ---------------------------------------------------------------------------------------------------------
    // IndexSelectorHDR of one round: set up from the quantised endpoints, then asked for the interpolator of single pixels.
    // Indexes are not carried through the search: they are a pure function of (quantised endpoints, pixel), so the winner's
    // are derived once per block by the same code (bc6h_derive_indexes).
    template<bool SIGNED, bool FAST, int RANGE, int STRIDE>
    struct BC6HIndexer
    {
        enum { kPairs = FAST ? 1 : RANGE / 2 };
        f2 rw[kPairs][3];            // slow: m_reconstructedInterpolators, interpolators 2j and 2j + 1 side by side
        int unq[2][3];               // interpolation endpoints
        float origin[3], axis[3];    // fast: IndexSelector::Init

        static CVTT_HD int weight_of(int i)
        {
            return wrap_u16(((RANGE == 8) ? 4681 : 2185) * i + 256) >> 9;     // g_weightReciprocals[range], IndexSelector.cpp:43-62
        }

        // q[epi][ch]: quantised endpoints as the quantiser returned them (signed values for SIGNED)
        CVTT_HD void init(const BC6HParams &P, const BC6HLane<STRIDE> &L, const int q[2][3], int aPrec)
        {
            int fin[2][3];
            for (int epi = 0; epi < 2; epi++)
                for (int ch = 0; ch < 3; ch++)
                    bc6h_unquantize_element<SIGNED>(q[epi][ch], aPrec, unq[epi][ch], fin[epi][ch]);
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
                const float mdl = fdiv((float)(RANGE - 1), lenSq);
                for (int ch = 0; ch < 3; ch++)
                    axis[ch] = fmul(fmul(dW[ch], P.w[ch]), mdl);
            }
            else
            {
                // InitHDR (IndexSelectorHDR.h:85-108).  The linear colours go to the lane's table: as floats [i][ch] for
                // range 8, as pairs of half patterns [ch][i / 2] for range 16 (the table has 24 words).
#pragma unroll
                for (int j = 0; j < RANGE / 2; j++)
#pragma unroll
                    for (int ch = 0; ch < 3; ch++)
                    {
                        const uint32_t r0 = (uint32_t)bc6h_reconstruct<SIGNED>(weight_of(2 * j), unq[0][ch], unq[1][ch]) & 0xffffu;
                        const uint32_t r1 = (uint32_t)bc6h_reconstruct<SIGNED>(weight_of(2 * j + 1), unq[0][ch], unq[1][ch]) & 0xffffu;
                        const float f0 = half_bits_to_float(r0), f1 = half_bits_to_float(r1);
                        rw[j][ch] = f2_make(fmul(f0, P.w[ch]), fmul(f1, P.w[ch]));
                        if (RANGE == 8)
                        {
                            L.tab[((2 * j) * 3 + ch) * STRIDE] = as_uint(f0);
                            L.tab[((2 * j + 1) * 3 + ch) * STRIDE] = as_uint(f1);
                        }
                        else
                            L.tab[(ch * 8 + j) * STRIDE] = r0 | (r1 << 16);
                    }
            }
        }

        // Uninverted interpolator number for the pixel with raw words (w0, w1) / linear colour lin; with slow indexing rl
        // receives that interpolator's linear colour (SelectIndexHDRSlow / SelectIndexHDRFast, IndexSelectorHDR.h:125-144).
        CVTT_HD int select(const BC6HParams &P, const BC6HLane<STRIDE> &L, uint32_t w0, uint32_t w1, const float *lin, float *rl) const
        {
            if (FAST)
            {
                int c[3];
                bc6h_raw_to_int(w0, w1, c);
                float dist = fmul(fsub((float)c[0], origin[0]), axis[0]);
                dist = fadd(dist, fmul(fsub((float)c[1], origin[1]), axis[1]));
                dist = fadd(dist, fmul(fsub((float)c[2], origin[2]), axis[2]));
                return packs_s16(f2i_rn(sse_clamp(dist, 0.0f, (float)(RANGE - 1))));
            }
            else
            {
                // errors of two interpolators per packed instruction, every lane with the reference's operation sequence
                const float pl0 = fmul(lin[0], P.w[0]), pl1 = fmul(lin[1], P.w[1]), pl2 = fmul(lin[2], P.w[2]);
                float e[RANGE];
#pragma unroll
                for (int j = 0; j < RANGE / 2; j++)
                {
                    const f2 d0 = f2_sub(pl0, rw[j][0]), d1 = f2_sub(pl1, rw[j][1]), d2 = f2_sub(pl2, rw[j][2]);
                    const f2 err = f2_add(f2_add(f2_mul(d0, d0), f2_mul(d1, d1)), f2_mul(d2, d2));
                    e[2 * j] = err.x;
                    e[2 * j + 1] = err.y;
                }
                // "first strictly smaller in ascending order" is the lexicographic minimum of (error, index): a tournament
                // gives the same index as the reference's sequential scan with a dependency chain of log2(RANGE) steps
                int ix[RANGE];
#pragma unroll
                for (int i = 0; i < RANGE; i++)
                    ix[i] = i;
#pragma unroll
                for (int step = 1; step < RANGE; step *= 2)
#pragma unroll
                    for (int i = 0; i + step < RANGE; i += 2 * step)
                    {
                        const bool better = e[i + step] < e[i];
                        ix[i] = better ? ix[i + step] : ix[i];
                        e[i] = better ? e[i + step] : e[i];
                    }
                const int index = ix[0];
                if (RANGE == 8)
                {
                    rl[0] = as_float(L.tab[(index * 3 + 0) * STRIDE]);
                    rl[1] = as_float(L.tab[(index * 3 + 1) * STRIDE]);
                    rl[2] = as_float(L.tab[(index * 3 + 2) * STRIDE]);
                }
                else
                {
                    const int word = index >> 1, sh = (index & 1) * 16;
                    for (int ch = 0; ch < 3; ch++)
                        rl[ch] = half_bits_to_float((L.tab[(ch * 8 + word) * STRIDE] >> sh) & 0xffffu);
                }
                return index;
            }
        }

        // error of the pixel against interpolator `raw` (ComputeErrorHDRFast / ComputeErrorHDRSlow, BCCommon.h:45-79)
        CVTT_HD float pixel_error(const BC6HParams &P, uint32_t w0, uint32_t w1, const float *lin, int raw, const float *rl) const
        {
            float error = 0.0f;
            const bool uniform = (P.flags & kFlag_Uniform) != 0;
            if (FAST)
            {
                // ReconstructHDR* + SqDiffSInt16 (ParallelMath.h:996-1010)
                int orig[3];
                bc6h_raw_to_int(w0, w1, orig);
                const int weight = wrap_u16(((RANGE == 8) ? 4681 : 2185) * raw + 256) >> 9;
                for (int ch = 0; ch < 3; ch++)
                {
                    const int rc = bc6h_reconstruct<SIGNED>(weight, unq[0][ch], unq[1][ch]);
                    const int hi = rc > orig[ch] ? rc : orig[ch], lo = rc > orig[ch] ? orig[ch] : rc;
                    const uint32_t diffU = (uint32_t)wrap_u16(hi - lo);
                    const float sq = (float)(int32_t)(diffU * diffU);
                    error = uniform ? fadd(error, sq) : fadd(error, fmul(sq, P.wSq[ch]));
                }
            }
            else
            {
                // SqDiff2CL(reconstructed, original)
                for (int ch = 0; ch < 3; ch++)
                {
                    const float diff = fsub(rl[ch], lin[ch]);
                    const float sq = fmul(diff, diff);
                    error = uniform ? fadd(error, sq) : fadd(error, fmul(sq, P.wSq[ch]));
                }
            }
            return error;
        }
    };

    // 8-bit digest (never 0) of a round's quantised endpoints, and "does any byte of (a, b, c) equal it"
    CVTT_HD uint32_t bc6h_digest(uint32_t q0, uint32_t q1, uint32_t q2)
    {
        const uint32_t h = (q0 * 0x9E3779B1u) ^ (q1 * 0x85EBCA77u) ^ (q2 * 0xC2B2AE3Du);
        return (h >> 24) | 1u;
    }
    CVTT_HD uint32_t bc6h_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }
    CVTT_HD uint32_t bc6h_byte_flags(uint32_t z) { return (((z >> 7) * 0x00204081u) >> 21) & 15u; }      // 0x80 flags of four bytes -> four bits

    // ---------------------------------------------------------------------------------------------------------
    // All trials of one (precision, partition): the meta rounds of each subset, then the commit scan (BC67.cpp:2790-2988).
    //
    // Control flow is warp-uniform wherever a vote follows: a round that the group drops (below) is carried as a lane
    // predicate (`live`) through the pixel loop instead of a group-divergent `continue`.
    //
    // Work that cannot change the result is skipped when no lane of the warp can use it (P.prune; the reference evaluates and
    // rejects it).  All three tests rest on the same facts: errors are sums of non-negative terms (see `prune` below for the one
    // configuration where they are not), fl(a + b) is monotonic in both operands, and the lane's best error only falls.
    //   (1) after subset 0: if even its smallest round error is not below the lane's best, no combination is;
    //   (2) in the last refinement pass (whose only product is the round's error) the pixel loop stops as soon as the partial
    //       sum, plus the smallest subset-0 error for a subset-1 round, reaches the best -- the round is then kept with the
    //       partial sum as its error, which cannot pass the commit test either;
    //   (3) the commit scan is skipped when the two smallest subset errors together are not below the best.
    template<bool SIGNED, bool FAST, int RANGE, int STRIDE, class Vote>
    CVTT_HD void bc6h_partition(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE> &L, Vote &vote, bool partitioned, int aPrec, int p,
        const float *ufepBase /* [2][3] */, const float *ufepOffs /* [2][3] */, BC6HBest &best)
    {
        enum { kMaxTweak = 4, kMaxRefine = 3, kMeta = 12 };
        const int numSubsets = partitioned ? 2 : 1;
        const uint32_t partitionMask = partitioned ? T.partitionMask[p] : 0u;
        const float maxV = (float)(RANGE - 1);
        // not for signed fast indexing: SqDiffSInt16 (ParallelMath.h:996-1010) converts the 32-bit square as a SIGNED integer, and a
        // sign | magnitude interpolator read as two's complement can be more than 46340 away from the pixel -- error terms can be
        // negative there, which breaks the monotonicity the tests rely on
        const bool prune = P.prune != 0 && !(SIGNED && FAST);

        // rounds that are never evaluated read as zero, like the automatics of the reference build;
        // they are written where a round is aborted, every other round writes its own entry before anything reads it
        uint32_t metaEP[kMeta][2][3];      // [round][subset][ch]: quantised ep0 | ep1 << 16
        float metaErr[kMeta][2];
        uint32_t roundValid = 0xffffffu;   // bit (round * 2 + subset); uniform over the 8 lanes of a group
        uint32_t roundInverted = 0;        // same numbering
        float minErr0 = FLT_MAX, minErr1 = partitioned ? FLT_MAX : 0.0f;     // smallest error of a valid round per subset

        for (int subset = 0; subset < numSubsets; subset++)
        {
            if (prune && subset == 1 && !vote.warp_any(minErr0 < best.error))
                return;

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
            uint32_t fixW0, fixW1;
            L.raw_at(fixupIndex, fixW0, fixW1);
            float fixPixelLin[3] = { 0.0f, 0.0f, 0.0f };
            if (!FAST)
                bc6h_raw_to_lin<SIGNED>(fixW0, fixW1, fixPixelLin);

            uint32_t dg0 = 0, dg1 = 0, dg2 = 0;     // digests of this subset's rounds so far, one byte per round (0 = none yet)
            float minErrS = FLT_MAX;

            for (int tweak = 0; tweak < kMaxTweak; tweak++)
            {
                // EndpointRefiner sums of the previous pass; `contributed` is false after a dropped round (BC67.cpp:2843)
                float tv[3] = { 0.0f, 0.0f, 0.0f }, tt = 0.0f, ts = 0.0f;
                bool contributed = false;

                for (int refinePass = 0; refinePass < kMaxRefine; refinePass++)
                {
                    const int metaRound = tweak * kMaxRefine + refinePass;
                    const uint32_t roundBit = 1u << (metaRound * 2 + subset);
                    const uint32_t dgShift = (uint32_t)(metaRound & 3) * 8u;
                    uint32_t digest;
                    bool live = true;
                    int ec[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } };

                    if (tweak >= P.tweakRounds || refinePass >= P.refineRounds)
                    {
                        roundValid &= ~roundBit;
                        metaEP[metaRound][subset][0] = metaEP[metaRound][subset][1] = metaEP[metaRound][subset][2] = 0;
                        digest = bc6h_digest(0, 0, 0);
                        live = false;
                    }
                    else
                    {
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
                    }
                    // refiners[subset].Init
                    tv[0] = tv[1] = tv[2] = 0.0f;
                    tt = ts = 0.0f;
                    contributed = false;

                    BC6HIndexer<SIGNED, FAST, RANGE, STRIDE> ix;
                    int fixRaw = 0;
                    float fixLin[3] = { 0.0f, 0.0f, 0.0f };
                    bool invert = false;
                    if (live)       // warp-uniform here: P.tweakRounds / P.refineRounds
                    {
                        // QuantizeEndpointsSigned / Unsigned (BC67.cpp:2503-2595)
                        int q[2][3];
                        for (int epi = 0; epi < 2; epi++)
                            for (int ch = 0; ch < 3; ch++)
                                q[epi][ch] = bc6h_quantize_element<SIGNED>(ec[epi][ch], aPrec);
                        ix.init(P, L, q, aPrec);

                        // fix-up index and conditional inversion
                        fixRaw = ix.select(P, L, fixW0, fixW1, fixPixelLin, fixLin);
                        invert = (RANGE / 2 - 1) < fixRaw;
                        const uint32_t qa[3] = { (uint32_t)wrap_u16(q[0][0]), (uint32_t)wrap_u16(q[0][1]), (uint32_t)wrap_u16(q[0][2]) };
                        const uint32_t qb[3] = { (uint32_t)wrap_u16(q[1][0]), (uint32_t)wrap_u16(q[1][1]), (uint32_t)wrap_u16(q[1][2]) };
                        uint32_t qp[3];
                        for (int ch = 0; ch < 3; ch++)
                            qp[ch] = invert ? (qb[ch] | (qa[ch] << 16)) : (qa[ch] | (qb[ch] << 16));
                        metaEP[metaRound][subset][0] = qp[0];
                        metaEP[metaRound][subset][1] = qp[1];
                        metaEP[metaRound][subset][2] = qp[2];
                        if (invert)
                            roundInverted |= roundBit;
                        digest = bc6h_digest(qp[0], qp[1], qp[2]);

                        // A round that repeats an earlier round's endpoints on all eight lanes is dropped (BC67.cpp:2853-2877).
                        // The digests of the earlier rounds are in registers; the exact comparison against the stored endpoints
                        // only runs for a group whose eight lanes all have a digest match.
                        if (metaRound > 0)
                        {
                            const uint32_t x = digest * 0x01010101u;
                            // bit r: round r's digest equals this one.  The zero-byte test can flag the byte above a true match
                            // as well: harmless for an earlier round (it is compared exactly), masked off for rounds to come.
                            uint32_t candidates = bc6h_byte_flags(bc6h_zero_byte(dg0 ^ x)) | (bc6h_byte_flags(bc6h_zero_byte(dg1 ^ x)) << 4) | (bc6h_byte_flags(bc6h_zero_byte(dg2 ^ x)) << 8);
                            candidates &= (1u << metaRound) - 1u;
                            const bool needExact = vote.all(candidates != 0);
                            if (vote.warp_any(needExact))
                            {
                                bool anySame = false;
                                if (needExact)
                                    for (; candidates; candidates &= candidates - 1)
                                    {
                                        const int prev = ctz32(candidates);
                                        const uint32_t p0 = metaEP[prev][subset][0], p1 = metaEP[prev][subset][1], p2 = metaEP[prev][subset][2];
                                        anySame = anySame || (p0 == qp[0] && p1 == qp[1] && p2 == qp[2]);
                                    }
                                if (vote.all(needExact && anySame))
                                {
                                    roundValid &= ~roundBit;
                                    live = false;
                                }
                            }
                        }
                    }
                    {
                        const uint32_t v = digest << dgShift;
                        if (metaRound < 4) dg0 |= v;
                        else if (metaRound < 8) dg1 |= v;
                        else dg2 |= v;
                    }
                    if (!vote.warp_any(live))
                        continue;

                    float subsetError = 0.0f;
                    const bool refineNext = (refinePass != P.refineRounds - 1);
                    const bool pruneLoop = prune && !refineNext;
                    for (uint32_t m = mask; m; m &= m - 1)
                    {
                        const int px = ctz32(m);
                        if (live)
                        {
                            uint32_t w0, w1;
                            L.raw_at(px, w0, w1);
                            float lin[3] = { 0.0f, 0.0f, 0.0f };
                            if (!FAST)
                                bc6h_raw_to_lin<SIGNED>(w0, w1, lin);
                            float rl[3] = { fixLin[0], fixLin[1], fixLin[2] };
                            int raw = fixRaw;
                            if (px != fixupIndex)
                                raw = ix.select(P, L, w0, w1, lin, rl);
                            subsetError = fadd(subsetError, ix.pixel_error(P, w0, w1, lin, raw, rl));

                            if (refineNext)
                            {
                                // EndpointRefiner::ContributeUnweightedPW (EndpointRefiner.h:78-92)
                                const int index = invert ? (RANGE - 1 - raw) : raw;
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
                        if (pruneLoop)
                        {
                            const float bound = (subset == 0) ? subsetError : fadd(minErr0, subsetError);
                            if (!vote.warp_any(live && bound < best.error))
                                break;
                        }
                    }
                    if (live)
                    {
                        metaErr[metaRound][subset] = subsetError;
                        minErrS = sse_min(minErrS, subsetError);
                    }
                }
            }
            if (subset == 0)
                minErr0 = minErrS;
            else
                minErr1 = minErrS;
        }

        // Combine the rounds of the two subsets; a combination that improves on the best so far is committed with the
        // modes of this (partitioned, precision) that can encode it (BC67.cpp:2915-2985).
        if (prune && !vote.warp_any(fadd(minErr0, minErr1) < best.error))
            return;
        const int numMeta1 = partitioned ? kMeta : 1;

        for (int meta0 = 0; meta0 < kMeta; meta0++)
        {
            // roundValid is uniform over a group but not over the warp: only skip what every group of the warp skips, the
            // votes below must be executed by all lanes.  A subset-0 round whose error plus the smallest subset-1 error is
            // not below the lane's best has no partner that is.
            const bool valid0 = ((roundValid >> (meta0 * 2)) & 1) != 0;
            const float err0 = valid0 ? metaErr[meta0][0] : FLT_MAX;
            const bool rowPossible = valid0 && (fadd(err0, minErr1) < best.error);
            if (!vote.warp_any(rowPossible))
                continue;
            for (int meta1 = 0; meta1 < numMeta1; meta1++)
            {
                float combinedError = err0;
                bool valid = valid0;
                if (partitioned)
                {
                    const bool valid1 = ((roundValid >> (meta1 * 2 + 1)) & 1) != 0;
                    valid = valid && valid1;
                    combinedError = fadd(combinedError, valid1 ? metaErr[meta1][1] : FLT_MAX);
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
                    q[1][0][ch] = partitioned ? (int)(metaEP[meta1][1][ch] & 0xffffu) : 0;
                    q[1][1][ch] = partitioned ? (int)(metaEP[meta1][1][ch] >> 16) : 0;
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
                            best.committed = true;
                            for (int s = 0; s < numSubsets; s++)
                                for (int e = 0; e < 2; e++)
                                    for (int ch = 0; ch < 3; ch++)
                                        best.ep[s][e][ch] = enc[s][e][ch];
                            // pixels of subset 0 take the indexes of meta0, pixels of subset 1 those of meta1: remembered as
                            // the rounds' endpoints and inversion, from which bc6h_derive_indexes recomputes them
                            for (int ch = 0; ch < 3; ch++)
                            {
                                best.q[0][ch] = metaEP[meta0][0][ch];
                                best.q[1][ch] = partitioned ? metaEP[meta1][1][ch] : 0u;
                            }
                            best.inverted = ((roundInverted >> (meta0 * 2)) & 1u) | (partitioned ? (((roundInverted >> (meta1 * 2 + 1)) & 1u) << 1) : 0u);
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

    // Indexes of the committed rounds: IndexSelectorHDR again on the winner's endpoints (the search keeps no indexes).
    template<bool SIGNED, bool FAST, int RANGE, int STRIDE>
    CVTT_HD void bc6h_derive_indexes(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE> &L, const BC6HBest &best, int aPrec, uint32_t idx[2])
    {
        const bool partitioned = (RANGE == 8);
        const uint32_t partitionMask = partitioned ? T.partitionMask[best.partition] : 0u;
        for (int subset = 0; subset < (partitioned ? 2 : 1); subset++)
        {
            const uint32_t mask = subset ? partitionMask : (~partitionMask & 0xffffu);
            const bool inverted = ((best.inverted >> subset) & 1u) != 0;
            int q[2][3];
            for (int ch = 0; ch < 3; ch++)
            {
                // the stored pair is in block order; the round searched with the quantiser's order
                const int a = (int)(best.q[subset][ch] & 0xffffu), b = (int)(best.q[subset][ch] >> 16);
                q[0][ch] = inverted ? b : a;
                q[1][ch] = inverted ? a : b;
                if (SIGNED)
                {
                    q[0][ch] = wrap_s16(q[0][ch]);
                    q[1][ch] = wrap_s16(q[1][ch]);
                }
            }
            BC6HIndexer<SIGNED, FAST, RANGE, STRIDE> ix;
            ix.init(P, L, q, aPrec);
            for (uint32_t m = mask; m; m &= m - 1)
            {
                const int px = ctz32(m);
                uint32_t w0, w1;
                L.raw_at(px, w0, w1);
                float lin[3] = { 0.0f, 0.0f, 0.0f }, rl[3];
                if (!FAST)
                    bc6h_raw_to_lin<SIGNED>(w0, w1, lin);
                const int raw = ix.select(P, L, w0, w1, lin, rl);
                const int index = inverted ? (RANGE - 1 - raw) : raw;
                idx[px >> 3] |= (uint32_t)index << (4 * (px & 7));
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    CVTT_HD void bc6h_best_reset(BC6HBest &best)
    {
        best.error = FLT_MAX;
        best.mode = 0;
        best.partition = 0;
        best.inverted = 0;
        best.committed = false;
        for (int s = 0; s < 2; s++)
            for (int e = 0; e < 2; e++)
                best.ep[s][e][0] = best.ep[s][e][1] = best.ep[s][e][2] = 0;
        for (int s = 0; s < 2; s++)
            best.q[s][0] = best.q[s][1] = best.q[s][2] = 0;
    }

    // The bit packing tail of BC6HComputer::Pack (BC67.cpp:2990-3050) for whatever `best` holds
    template<bool SIGNED, bool FAST, int STRIDE>
    CVTT_HD void bc6h_pack_block(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE> &L, const BC6HBest &best, uint32_t out[4])
    {
        // header: scatter the endpoint fields into the mode's bit layout, then the indexes (PackingVector::Pack does not
        // mask its argument; the fix-up indexes have their top bit clear by construction)
        const uint8_t *mi = T.modes[best.mode];
        const bool partitioned = mi[1] != 0;
        uint32_t idx[2] = { 0, 0 };
        if (best.committed)
        {
            if (partitioned)
                bc6h_derive_indexes<SIGNED, FAST, 8, STRIDE>(P, T, L, best, (int)mi[3], idx);
            else
                bc6h_derive_indexes<SIGNED, FAST, 16, STRIDE>(P, T, L, best, (int)mi[3], idx);
        }
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
            const uint32_t index = (idx[px >> 3] >> (4 * (px & 7))) & 15u;
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

    // ---------------------------------------------------------------------------------------------------------
    // The search as a numbered sequence of CALLS of bc6h_partition, in the reference's order (BC67.cpp:2776-2788): the four
    // one-subset precisions 16, 12, 11, 10, then for each two-subset precision 11, 10, 9, 8, 7, 6 the 32 partitions.
    //
    // What the small-call launch (bc6h_kernels.cu) rests on.  A call changes a lane's best only through commits of
    // combinations that are strictly better than its best error, so (1) the lane's best ERROR after a call is min(error
    // before, smallest legal combination of the call) whatever the other lanes of its group do -- a lane that is better keeps
    // the mode loop of its group going until it has committed -- and (2) the lane's final mode / endpoints are those it held
    // at the end of its WINNER call, the first call that reaches its final error.  What the other lanes do decides only which
    // of the legal modes a commit ends on (the group-wide mode loop), and that depends on their best errors at the START of
    // the winner call.  So the calls can be searched in any grouping for their error histories alone (bc6h_search_calls),
    // and a lane's block is then reproduced exactly by running its winner call once more for the whole group with every
    // lane's true entry error (bc6h_history + bc6h_run_call).
    enum { kBC6HCalls = 4 + 6 * 32 };

    CVTT_HD void bc6h_call_info(int call, bool &partitioned, int &aPrec, int &p)
    {
        if (call < 4)
        {
            partitioned = false;
            aPrec = call == 0 ? 16 : 13 - call;
            p = 0;
        }
        else
        {
            partitioned = true;
            aPrec = 11 - ((call - 4) >> 5);
            p = (call - 4) & 31;
        }
    }

    // One call: the endpoint fits of its partition (BC67.cpp:2739-2774), then the trials
    template<bool SIGNED, bool FAST, int STRIDE, class Vote>
    CVTT_HD void bc6h_run_call(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE> &L, Vote &vote, int call, BC6HBest &best)
    {
        bool partitioned;
        int aPrec, p;
        bc6h_call_info(call, partitioned, aPrec, p);
        float base[6] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f }, offs[6] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
        if (partitioned)
        {
            for (int subset = 0; subset < 2; subset++)
            {
                const uint32_t mask = subset ? T.partitionMask[p] : ((uint32_t)~T.partitionMask[p] & 0xffffu);
                int n = 0;
                for (uint32_t m = mask; m; m &= m - 1)
                    n++;
                endpoint_selector3_masked(L, mask, n, P.w, base + subset * 3, offs + subset * 3);
            }
            bc6h_partition<SIGNED, FAST, 8, STRIDE>(P, T, L, vote, true, aPrec, p, base, offs, best);
        }
        else
        {
            endpoint_selector3_masked(L, 0xffffu, 16, P.w, base, offs);
            bc6h_partition<SIGNED, FAST, 16, STRIDE>(P, T, L, vote, false, aPrec, 0, base, offs, best);
        }
    }

    // Calls [callBegin, callEnd) from a best that holds nothing but the error `startError`; running[(call - callBegin) * stride]
    // receives the lane's best error after each call, a running minimum local to this range.  startError is FLT_MAX or the
    // lane's best error after calls that PRECEDE the range (something to prune against): such calls are part of every prefix
    // the history is asked for, so the minima stay exact.
    template<bool SIGNED, bool FAST, int STRIDE, class Vote>
    CVTT_HD void bc6h_search_calls(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE> &L, Vote &vote, int callBegin, int callEnd, float startError, float *running, size_t stride, bool store)
    {
        BC6HBest best;
        bc6h_best_reset(best);
        best.error = startError;
        for (int call = callBegin; call < callEnd; call++)
        {
            bc6h_run_call<SIGNED, FAST, STRIDE>(P, T, L, vote, call, best);
            if (store)
                running[(size_t)(call - callBegin) * stride] = best.error;
        }
    }

    // The lane's error history from the ranges' running minima, history[call * stride]: every entry is the minimum over SOME
    // calls up to and including its own, so the running minimum of the entries is the best error after each call whatever the
    // ranges were.  Returns the best error before call `upto` (FLT_MAX for none) and the first call that reached it (-1:
    // nothing committed).
    CVTT_HD float bc6h_history(const float *history, size_t stride, int upto, int &winner)
    {
        float cur = FLT_MAX;
        winner = -1;
        for (int call = 0; call < upto; call++)
        {
            const float r = history[(size_t)call * stride];
            if (r < cur)
            {
                cur = r;
                winner = call;
            }
        }
        return cur;
    }

    // ---------------------------------------------------------------------------------------------------------
    // The whole search for one block and the bit packing tail (BC67.cpp:2990-3050).  The endpoint fits of a call's partition
    // (BC67.cpp:2739-2774) are made when the call runs, not kept in a table of all 33 across the search: a fit is ~1 % of a
    // call, the table was 1.6 KB of local memory per thread (more than the L2 holds for the resident threads).
    template<bool SIGNED, bool FAST, int STRIDE, class Vote>
    CVTT_HD void bc6h_encode_block(const BC6HParams &P, const BC6HTables &T, const BC6HLane<STRIDE> &L, Vote &vote, uint32_t out[4])
    {
        BC6HBest best;
        bc6h_best_reset(best);
        for (int call = 0; call < kBC6HCalls; call++)
            bc6h_run_call<SIGNED, FAST, STRIDE>(P, T, L, vote, call, best);
        bc6h_pack_block<SIGNED, FAST, STRIDE>(P, T, L, best, out);
    }
}
