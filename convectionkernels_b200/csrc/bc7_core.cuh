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
#ifndef CVTT_F4_DEFINED
#define CVTT_F4_DEFINED
    struct alignas(16) F4 { float x, y, z, w; };
#endif

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
        int splitSlices;                   // > 0: small-call launch, cmds starts with this many offsets of independent sub-streams
        const uint32_t *cmds;
    };

    // Command stream opcodes (built by bc7_plan.cpp)
    enum { kCmdEnd = 0, kCmdShape = 1, kCmdEval = 2, kCmdDual = 3, kCmdPair2 = 4, kCmdTriple = 5 };
    enum { kBC7MaxSlots = 184 };

    // lexicographic order of the reference's commit sequence: modes 0,1,2,3,6,7 (TrySinglePlane) then 4,5 (TryDualPlane)
    CVTT_HD int bc7_mode_order(int mode) { return (mode < 4) ? mode : (mode == 6 ? 4 : (mode == 7 ? 5 : (mode == 4 ? 6 : 7))); }

    struct BC7LaneFlags
    {
        bool anyBlockHasAlpha;    // group vote, BC67.cpp:1069
        bool allowRGBModes;       // group vote, BC67.cpp:1072
        bool blockHasNonMaxAlpha; // this block
        bool blockHasNonZeroAlpha, isPunchThrough;   // this block: max alpha > 0; every alpha is 0 or 255 (BC67.cpp:1056-1067)
        // warp-level "does any lane need this path" (pure work skipping, never changes a lane's result)
        bool warpAnyRGB, warpAnyPCA4, warpAnyExpand, warpAnyMode7;
        bool warpHasWork;           // false for a warp that holds no block at all
    };

    struct BC7Work   // BC67::WorkInfo, BC67.cpp:59-76, in packed form
    {
        float error;
        int key;              // bc7_mode_order(mode) * 64 + (partition | rotation*2+indexSelector)
        int mode, sub;        // sub = partition, or rotation | indexSelector << 2
        uint32_t ep[3][2];    // per subset, per endpoint: r | g << 8 | b << 16 | a << 24
        uint32_t idx[2];      // 16 x 4 bits, pixel order
        uint32_t idx2[2];
        uint32_t sc[3];       // per subset: 0, or 0x80000000 | index when the subset's winner is a single-colour candidate
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
    // Per-lane pixel storage.  STRIDE is the element stride between consecutive entries of one lane (the CTA
    // size in the kernel, 1 on the CPU), so that every unrolled access has a compile-time offset.
    //   raw[px]  the block's 16 pixels as packed RGBA8 (never modified)
    //   gv[i]    the i-th pixel of the current pixel subset as (value + kMagic), channels possibly rotated
    //   gw[i]    the same pixel pre-weighted: (value * channelWeight), BCCommon::PreWeightPixelsLDR (BCCommon.h:81-99)
    template<int STRIDE>
    struct BC7Lane
    {
        const uint32_t *raw;
        F4 *gv;
        F4 *gw;
    };

    CVTT_HD F4 bc7_expand_pixel(uint32_t w)
    {
        F4 p;
#if defined(__CUDA_ARCH__)
        p.x = __uint_as_float(__byte_perm(w, kMagicBits, 0x7650));
        p.y = __uint_as_float(__byte_perm(w, kMagicBits, 0x7651));
        p.z = __uint_as_float(__byte_perm(w, kMagicBits, 0x7652));
        p.w = __uint_as_float(__byte_perm(w, kMagicBits, 0x7653));
#else
        p.x = as_float(kMagicBits | (w & 0xffu));
        p.y = as_float(kMagicBits | ((w >> 8) & 0xffu));
        p.z = as_float(kMagicBits | ((w >> 16) & 0xffu));
        p.w = as_float(kMagicBits | (w >> 24));
#endif
        return p;
    }

    // swaps the colour channel (rotation - 1) with alpha, BC67.cpp:1690-1716
    CVTT_HD void bc7_rotate(F4 &p, int rotation)
    {
        const float t = p.w;
        if (rotation == 1) { p.w = p.x; p.x = t; }
        else if (rotation == 2) { p.w = p.y; p.y = t; }
        else if (rotation == 3) { p.w = p.z; p.z = t; }
    }

    // ---------------------------------------------------------------------------------------------------------
    // EndpointSelector<NCH, 8> over the n gathered pixels, with unit pixel weights.  gw holds the pre-weighted
    // pixels; wv are the channel weights they were weighted with (GetEndpoints divides by them again).
    template<int NCH, int STRIDE>
    CVTT_HD void bc7_endpoint_selector(const F4 *gw, int n, const float *wv, float *base, float *offs)
    {
        float centroid[NCH], cov[NCH * (NCH + 1) / 2];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            centroid[ch] = 0.0f;
#pragma unroll
        for (int i = 0; i < NCH * (NCH + 1) / 2; i++)
            cov[i] = 0.0f;

        // pass 0: centroid (EndpointSelector.h:73-86)
#pragma unroll 2
        for (int i = 0; i < n; i++)
        {
            const F4 p = gw[i * STRIDE];
            const float pv[4] = { p.x, p.y, p.z, p.w };
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                centroid[ch] = fadd(centroid[ch], pv[ch]);
        }
        {
            const float denom = (float)n;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                centroid[ch] = fdiv(centroid[ch], denom);
        }

        // pass 1: covariance (EndpointSelector.h:88-95, PackedCovarianceMatrix.h:29-40)
#pragma unroll 2
        for (int i = 0; i < n; i++)
        {
            const F4 p = gw[i * STRIDE];
            const float pv[4] = { p.x, p.y, p.z, p.w };
            float diff[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                diff[ch] = fsub(pv[ch], centroid[ch]);
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

#pragma unroll 1
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
#pragma unroll 2
        for (int i = 0; i < n; i++)
        {
            const F4 p = gw[i * STRIDE];
            const float pv[4] = { p.x, p.y, p.z, p.w };
            float dist = 0.0f;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                dist = fadd(dist, fmul(direction[ch], fsub(pv[ch], centroid[ch])));
            minDist = sse_min(minDist, dist);
            maxDist = sse_max(maxDist, dist);
        }

        // GetEndpoints (EndpointSelector.h:51-70): divides by the raw channel weight
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
        {
            const float mn = fadd(centroid[ch], fmul(direction[ch], minDist));
            const float mx = fadd(centroid[ch], fmul(direction[ch], maxDist));
            base[ch] = fdiv(mn, wv[ch]);
            offs[ch] = fdiv(fsub(mx, mn), wv[ch]);
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // Compile-time description of the single-plane modes: g_modes (BC67.cpp:108-119), the parity-bit loop bounds
    // (BC67.cpp:1183-1189) and CompressEndpoints0-7 (BC67.cpp:862-938) in the exact-fp32 form of QuantConst.
    template<int MODE> struct BC7ModeT;
    template<> struct BC7ModeT<0> { enum { NCH = 3, BITS = 4, WITHP = 1, UNQ = 5, PMAX = 4, SHAREDP = 0, IB = 3 }; };
    template<> struct BC7ModeT<1> { enum { NCH = 3, BITS = 6, WITHP = 1, UNQ = 7, PMAX = 2, SHAREDP = 1, IB = 3 }; };
    template<> struct BC7ModeT<2> { enum { NCH = 3, BITS = 5, WITHP = 0, UNQ = 5, PMAX = 1, SHAREDP = 0, IB = 2 }; };
    template<> struct BC7ModeT<3> { enum { NCH = 3, BITS = 7, WITHP = 1, UNQ = 0, PMAX = 4, SHAREDP = 0, IB = 2 }; };
    template<> struct BC7ModeT<6> { enum { NCH = 4, BITS = 7, WITHP = 1, UNQ = 0, PMAX = 4, SHAREDP = 0, IB = 4 }; };
    template<> struct BC7ModeT<7> { enum { NCH = 4, BITS = 5, WITHP = 1, UNQ = 6, PMAX = 4, SHAREDP = 0, IB = 2 }; };

    // Quantises the integer-valued endpoint channel c to the mode's precision with parity bit p and expands it
    // back to 8 bits; the result is returned biased (value + kMagic).  qAddP = qAdd[p], pM = p + kMagic.  Both lanes.
    template<int MODE>
    CVTT_HD f2 bc7_quant_biased(f2 c, f2 qAddP, f2 p, f2 pM)
    {
        typedef BC7ModeT<MODE> M;
        const float qMul = M::WITHP ? (float)((1 << (M::BITS + 1)) - 1) / 512.0f : (float)((1 << M::BITS) - 1) / 256.0f;
        const float m = magic_in_register();          // register operand, see bc7_trial_pixels
        const f2 vb = f2_add(f2_fma(c, qMul, qAddP), m);
        if (M::UNQ)
        {
            const int s = M::UNQ ? 2 * M::UNQ - 8 : 0;
            const float uMul = (float)(1 << (8 - M::UNQ)), uScale = 1.0f / (float)(1 << s), uOff = -(float)((1 << s) - 1) / (float)(1 << (s + 1));
            const f2 v = f2_sub(vb, m);
            const f2 v2 = M::WITHP ? f2_fma(v, 2.0f, p) : v;
            const f2 flb = f2_add(f2_fma(v2, uScale, uOff), m);
            return f2_fma(v2, uMul, flb);
        }
        else if (M::WITHP)
            return f2_fma(f2_sub(vb, m), 2.0f, pM);
        else
            return vb;
    }

    template<int MODE>
    CVTT_HD float bc7_quant_add(int p)
    {
        typedef BC7ModeT<MODE> M;
        if (M::WITHP)
        {
            const float a0 = (float)(2 * 255 - 511) / 1024.0f, a1 = (float)(2 * ((1 << (8 - M::BITS)) - 1) - 511) / 1024.0f;
            return p ? a1 : a0;
        }
        return (float)(2 * (127 + (1 << (7 - M::BITS))) - 255) / 512.0f;
    }

    // ---------------------------------------------------------------------------------------------------------
    // One pair of trials of the inner search (one per fp32 lane): index selection, reconstruction error and
    // (REFINE) the refiner's sums over the n gathered pixels (BC67.cpp:1355-1392).
    //   nom = -(q0 + kMagic), nd64 = -((q1 - q0) / 64 + 2^-20): negated so that the loop only adds.
    //
    // Reconstruction in one rounding.  ReconstructLDR_BC7 is floor(x + 1/2) with x = q0 + w (q1 - q0) / 64, a multiple of
    // 1/64.  fma(w, nd64, nom) evaluates -(x + w 2^-20) - kMagic exactly and rounds it once, at kMagic's scale, to an
    // integer (round-to-nearest-even).  The w 2^-20 <= 2^-14 term only matters when x is exactly half way (it is then the
    // tie-break towards x + 1/2, which is what the floor does); any other x is at least 1/64 away from a half.  For w = 0,
    // x = q0 is an integer.  (q1 - q0) / 64 + 2^-20 needs 22 significant bits, so nd64 is exact.
    //
    // Index -> weight in one rounding from the *biased* index.  g_weights2/3/4 (BC67.cpp:121-132) are rne(i * 64 / maxIndex);
    // any multiplier close enough to 64 / maxIndex gives the same integers, and a short dyadic one, s = 683/32, 585/64,
    // 273/64, makes kMagic * s and kMagic - kMagic * s exactly representable, so fma(kMagic + i, s, kMagic - kMagic * s)
    // = rne(i * s) + kMagic with no intermediate rounding (no product i * s lands on a half: checked for every index).
    // The unbiased index is then only needed by the refiner.
    //
    // Instruction forms (tools/ubench/pipe_rates.cu, cycles per warp instruction per scheduler at this kernel's occupancy):
    // FADD2 reg + scalar register 2.1, FADD2 reg + reg 2.25, FFMA2 2.6-2.8, FADD2 with an *immediate* 3.4, FMNMX free next
    // to packed ops, LOP3 / PRMT +2.4.  Hence kMagic is passed in a register (m) and the weight stays on the fp32 pipe.
    template<int IB> struct BC7WeightScale;
    template<> struct BC7WeightScale<2> { static constexpr float s = 21.34375f; };
    template<> struct BC7WeightScale<3> { static constexpr float s = 9.140625f; };
    template<> struct BC7WeightScale<4> { static constexpr float s = 4.265625f; };

    template<int NCH, int IB, bool FAST, bool REFINE, int STRIDE>
    CVTT_HD f2 bc7_trial_pixels(const BC7Params &P, const F4 *gv, const F4 *gw, int n, const f2 *nom, const f2 *axis, const f2 *nd64,
        f2 *tv, f2 &tt, f2 &ts)
    {
        const float maxV = (float)((1 << IB) - 1), wScale = 64.0f / (float)((1 << IB) - 1), rcpMaxIndex = 1.0f / (float)((1 << IB) - 1);
        const float ws = BC7WeightScale<IB>::s, wsBias = kMagic - kMagic * ws;           // both exact
        const float m = magic_in_register();
        f2 acc[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            acc[ch] = f2_splat(0.0f);
        f2 slowErr = f2_splat(0.0f);

#pragma unroll 4
        for (int i = 0; i < n; i++)
        {
            const F4 p = gv[i * STRIDE];
            const float pv[4] = { p.x, p.y, p.z, p.w };

            // SelectIndexLDR (IndexSelector.h:124-131)
            f2 dist = f2_mul(f2_add(nom[0], pv[0]), axis[0]);
#pragma unroll
            for (int ch = 1; ch < NCH; ch++)
                dist = f2_add(dist, f2_mul(f2_add(nom[ch], pv[ch]), axis[ch]));
            const f2 idxb = f2_add(f2_clamp_for_round(dist, 0.0f, maxV), m);           // kMagic + index
            f2 idxf = f2_splat(0.0f);
            if (REFINE || !FAST)
                idxf = f2_sub(idxb, m);

            // ReconstructLDR_BC7 (IndexSelector.h:90-100) + ComputeErrorLDR (BCCommon.h:24-43); df is the negated difference
            const f2 wf = f2_sub(f2_fma(idxb, ws, wsBias), m);
            f2 d2[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
            {
                const f2 df = f2_add(f2_fma(wf, nd64[ch], nom[ch]), pv[ch]);
                if (FAST)
                    acc[ch] = f2_fma(df, df, acc[ch]);              // exact (< 2^24)
                else
                    d2[ch] = f2_mul(df, df);                        // exact (< 2^16)
            }

            if (!FAST)
            {
                // BC67.cpp:1364-1386: probe index-1 and index+1 in weighted float error (wSq is 1 under Flags::Uniform)
                f2 error = f2_mul(d2[0], P.wSq[0]);
#pragma unroll
                for (int ch = 1; ch < NCH; ch++)
                    error = f2_add(error, f2_mul(d2[ch], P.wSq[ch]));
                f2 alt[2];
                alt[0] = f2_sub(f2_make(fmaxf(idxf.x, 1.0f), fmaxf(idxf.y, 1.0f)), 1.0f);
                alt[1] = f2_add(idxf, 1.0f);
                alt[1] = f2_make(fminf(alt[1].x, maxV), fminf(alt[1].y, maxV));
#pragma unroll
                for (int ii = 0; ii < 2; ii++)
                {
                    const f2 awf = f2_sub(f2_fma(alt[ii], wScale, kMagic), kMagic);
                    f2 altError = f2_splat(0.0f);
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                    {
                        const f2 df = f2_add(f2_fma(awf, nd64[ch], nom[ch]), pv[ch]);
                        const f2 sq = f2_mul(df, df);
                        altError = (ch == 0) ? f2_mul(sq, P.wSq[0]) : f2_add(altError, f2_mul(sq, P.wSq[ch]));
                    }
                    const bool betterX = altError.x < error.x, betterY = altError.y < error.y;
                    error = f2_make(sse_min(error.x, altError.x), sse_min(error.y, altError.y));
                    if (betterX)
                        idxf.x = alt[ii].x;
                    if (betterY)
                        idxf.y = alt[ii].y;
                }
                slowErr = f2_add(slowErr, error);
            }

            // EndpointRefiner::ContributeUnweightedPW (EndpointRefiner.h:78-92)
            if (REFINE)
            {
                const F4 q = gw[i * STRIDE];
                const float qv[4] = { q.x, q.y, q.z, q.w };
                const f2 t = f2_mul(idxf, rcpMaxIndex);
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                    tv[ch] = f2_add(tv[ch], f2_mul(t, qv[ch]));
                tt = f2_add(tt, f2_mul(t, t));
                ts = f2_add(ts, t);
            }
        }

        // AggregatedError::Finalize (AggregatedError.h:25-46)
        if (!FAST)
            return slowErr;
        f2 shapeError;
        if (P.flags & kFlag_Uniform)
        {
            shapeError = acc[0];
#pragma unroll
            for (int ch = 1; ch < NCH; ch++)
                shapeError = f2_add(shapeError, acc[ch]);
        }
        else
        {
            shapeError = f2_mul(acc[0], P.wSq[0]);
#pragma unroll
            for (int ch = 1; ch < NCH; ch++)
                shapeError = f2_add(shapeError, f2_mul(acc[ch], P.wSq[ch]));
        }
        return shapeError;
    }

    // ---------------------------------------------------------------------------------------------------------
    // The inner search of TrySinglePlane for one (mode, shape): tweaks x parity bits x refine rounds
    // (BC67.cpp:1298-1432).  Result: best error and endpoints of the shape; the indexes of the overall winner are
    // re-derived from its endpoints at the end (bc7_derive_indices), they are a pure function of them.
    //
    // Two trial chains run side by side in the two fp32 lanes: the two values of the first parity bit (modes with
    // parity bits) or two tweaks (mode 2).  Each lane keeps its own best; the reference's "first strictly better in
    // (pIter, tweak, refine) order" is the lexicographic minimum of (error, sequence number) over both lanes.
    struct BC7ShapeBest
    {
        float err;
        uint32_t e0, e1;      // packed endpoint bytes
        int seq;              // reference sequence number of the winning trial (bc7_shape_trials), for merging partial searches
    };

    template<int NCH>
    struct BC7PairBest
    {
        f2 err;
        int seqX, seqY;
        f2 e0[NCH], e1[NCH];  // biased

        // one finished refine round of both chains: each lane keeps its (error, sequence number) minimum
        CVTT_HD void trial(int sx, int sy, const f2 &shapeError, const f2 *q0b, const f2 *q1b)
        {
            if (shapeError.x < err.x || (shapeError.x == err.x && sx < seqX))
            {
                err.x = shapeError.x;
                seqX = sx;
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    e0[ch].x = q0b[ch].x;
                    e1[ch].x = q1b[ch].x;
                }
            }
            if (shapeError.y < err.y || (shapeError.y == err.y && sy < seqY))
            {
                err.y = shapeError.y;
                seqY = sy;
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    e0[ch].y = q0b[ch].y;
                    e1[ch].y = q1b[ch].y;
                }
            }
        }
    };

    // Flags::BC7_RespectPunchThrough, modes 6 and 7 (BC67.cpp:1281-1303,1400-1428).  A parity combination that would move
    // a punch-through block's alpha off 0 / 255 is "invalid" for that block.  What the reference does with that:
    //   * a parity combination invalid for all 8 blocks of the call is skipped (:1300)
    //   * if it is invalid for none, trials commit normally
    //   * otherwise (:1409-1416) a trial that is better for at least one block of the call is committed under
    //     AndNot(invalid, better), and ParallelMath::AndNot(a, b) is a & ~b (ParallelMath.h:901-906): the blocks that take
    //     the trial are the INVALID ones for which it is NOT better; valid blocks take nothing.
    // Commits are therefore neither monotonic nor independent of the neighbours, so the trials run one chain at a time in
    // the reference's (pIter, tweak, refine) order (fp32 lane x; lane y repeats it) with one group vote per trial.
    template<int NCH, class Vote>
    struct BC7PunchCommit
    {
        Vote *vote;
        float err;
        uint32_t e0, e1;                 // packed endpoint bytes
        bool invalid, groupAnyInvalid, groupAllInvalid;

        CVTT_HD void trial(int, int, const f2 &shapeError, const f2 *q0b, const f2 *q1b)
        {
            const bool better = !groupAllInvalid && shapeError.x < err;
            if (vote->any(better))
            {
                const bool take = groupAnyInvalid ? (invalid && !better) : better;
                if (take)
                {
                    err = shapeError.x;
                    e0 = e1 = 0;
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++)
                    {
                        e0 |= (as_uint(q0b[ch].x) & 0xffu) << (8 * ch);
                        e1 |= (as_uint(q1b[ch].x) & 0xffu) << (8 * ch);
                    }
                }
            }
        }
    };

    // refine rounds of one pair of trial chains starting from the tweaked endpoints u0/u1
    template<int MODE, bool FAST, int STRIDE, class Commit>
    CVTT_HD void bc7_trial_pair(const BC7Params &P, const F4 *gv, const F4 *gw, int n, const float *sumV, float staticAlphaError,
        const f2 *u0, const f2 *u1, int p0x, int p0y, int p1x, int p1y, int seqX, int seqY, Commit &best)
    {
        typedef BC7ModeT<MODE> M;
        enum { NCH = M::NCH };
        const float maxV = (float)((1 << M::IB) - 1);
        const int R = P.refineRounds;
        const float wN = (float)n, wRcp = P.rcpN[n];

        const f2 qA0 = f2_make(bc7_quant_add<MODE>(p0x), bc7_quant_add<MODE>(p0y)), qA1 = f2_make(bc7_quant_add<MODE>(p1x), bc7_quant_add<MODE>(p1y));
        const f2 pf0 = f2_make((float)p0x, (float)p0y), pf1 = f2_make((float)p1x, (float)p1y);
        const f2 pM0 = f2_add(pf0, magic_in_register()), pM1 = f2_add(pf1, magic_in_register());

        f2 e0[NCH], e1[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
        {
            e0[ch] = u0[ch];
            e1[ch] = u1[ch];
        }

#pragma unroll 1
        for (int refine = 0; refine < R; refine++)
        {
            const bool lastRound = (refine == R - 1);

            // CompressEndpointsN (BC67.cpp:862-938); q0b/q1b are biased
            f2 q0b[NCH], q1b[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
            {
                q0b[ch] = bc7_quant_biased<MODE>(e0[ch], qA0, pf0, pM0);
                q1b[ch] = bc7_quant_biased<MODE>(e1[ch], qA1, pf1, pM1);
            }

            // IndexSelector<4>::Init (IndexSelector.h:27-78).  For NCH == 3 the alpha endpoints are both 255, so
            // the fourth channel contributes exactly +0 to every sum below and is left out.
            f2 dq[NCH], dW[NCH], axis[NCH], nom[NCH], nd64[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
            {
                dq[ch] = f2_sub(q1b[ch], q0b[ch]);                      // exact
                dW[ch] = f2_mul(dq[ch], P.w[ch]);
            }
            f2 lenSq = f2_mul(dW[0], dW[0]);
#pragma unroll
            for (int ch = 1; ch < NCH; ch++)
                lenSq = f2_add(lenSq, f2_mul(dW[ch], dW[ch]));
            safe_denominator(lenSq.x);
            safe_denominator(lenSq.y);
            const f2 mdl = f2_div(f2_splat(maxV), lenSq);
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
            {
                axis[ch] = f2_mul(f2_mul(dW[ch], P.w[ch]), mdl);
                nom[ch] = f2_neg(q0b[ch]);
                nd64[ch] = f2_fma(dq[ch], -0.015625f, f2_splat(-9.5367431640625e-07f));     // -(dq / 64 + 2^-20), exact
            }

            f2 tv[NCH], tt = f2_splat(0.0f), ts = f2_splat(0.0f);
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
                tv[ch] = f2_splat(0.0f);

            f2 shapeError;
            if (lastRound)
                shapeError = bc7_trial_pixels<NCH, M::IB, FAST, false, STRIDE>(P, gv, gw, n, nom, axis, nd64, tv, tt, ts);
            else
                shapeError = bc7_trial_pixels<NCH, M::IB, FAST, true, STRIDE>(P, gv, gw, n, nom, axis, nd64, tv, tt, ts);
            if (NCH == 3)
                shapeError = f2_add(shapeError, staticAlphaError);

            best.trial(seqX + refine, seqY + refine, shapeError, q0b, q1b);

            // EndpointRefiner::GetRefinedEndpointsLDR (EndpointRefiner.h:99-152)
            if (!lastRound)
            {
                f2 adenom = f2_mul(f2_sub(f2_mul(tt, wN), f2_mul(ts, ts)), wRcp);
                const bool zeroX = (adenom.x == 0.0f), zeroY = (adenom.y == 0.0f);
                if (zeroX)
                    adenom.x = 1.0f;
                if (zeroY)
                    adenom.y = 1.0f;
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    const f2 a = f2_div(f2_sub(tv[ch], f2_mul(f2_mul(ts, sumV[ch]), wRcp)), adenom);
                    const f2 b = f2_mul(f2_sub(sumV[ch], f2_mul(a, ts)), wRcp);
                    f2 p1v = b, p2v = f2_add(a, b);
                    const float flat = fmul(sumV[ch], wRcp);
                    if (zeroX)
                        p1v.x = p2v.x = flat;
                    if (zeroY)
                        p1v.y = p2v.y = flat;
                    e0[ch] = f2_rne(f2_clamp_for_round(f2_mul(p1v, P.rcpW[ch]), 0.0f, 255.0f));
                    e1[ch] = f2_rne(f2_clamp_for_round(f2_mul(p2v, P.rcpW[ch]), 0.0f, 255.0f));
                }
            }
        }
    }

    // ppFirst / ppEnd restrict the search to the parity-pair iterations [ppFirst, ppEnd) (modes with four parity combinations:
    // 0 .. 2); the results of disjoint parts merge by (err, seq).
    template<int MODE, bool FAST, int STRIDE>
    CVTT_HD void bc7_shape_trials(const BC7Params &P, const F4 *gv, const F4 *gw, int n, int seeds, const float *base, const float *offs,
        const float *sumV, float staticAlphaError, BC7ShapeBest &out, int ppFirst = 0, int ppEnd = 2)
    {
        typedef BC7ModeT<MODE> M;
        enum { NCH = M::NCH };
        const IndexConst &ic = P.ic[M::IB - 2];

        BC7PairBest<NCH> best;
        best.err = f2_splat(FLT_MAX);
        best.seqX = best.seqY = 0x7fffffff;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            best.e0[ch] = best.e1[ch] = f2_splat(kMagic);

        // sequence number of a trial: reference order is pIter, tweak, refine
        if (M::PMAX >= 2)
        {
#pragma unroll 1
            for (int tweak = 0; tweak < seeds; tweak++)
            {
                // UnfinishedEndpoints::FinishLDR (UnfinishedEndpoints.h:75-91), shared by both lanes
                const float tf0 = ic.tweak[tweak][0], tf1 = ic.tweak[tweak][1];
                f2 u0[NCH], u1[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    u0[ch] = f2_splat(rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf0)), 0.0f, 255.0f)));
                    u1[ch] = f2_splat(rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf1)), 0.0f, 255.0f)));
                }
#pragma unroll 1
                for (int pp = ppFirst; pp < imin(ppEnd, M::PMAX / 2); pp++)
                {
                    // lanes: pIter = 2 pp and 2 pp + 1, i.e. first parity bit 0 and 1
                    const int p1 = M::SHAREDP ? 0 : pp;
                    const int seqX = ((pp * 2) * 4 + tweak) << 16, seqY = ((pp * 2 + 1) * 4 + tweak) << 16;
                    bc7_trial_pair<MODE, FAST, STRIDE>(P, gv, gw, n, sumV, staticAlphaError, u0, u1, 0, 1, M::SHAREDP ? 0 : p1, M::SHAREDP ? 1 : p1, seqX, seqY, best);
                }
            }
        }
        else
        {
#pragma unroll 1
            for (int tweak = 0; tweak < seeds; tweak += 2)
            {
                // lanes: tweak and tweak + 1; an odd tail repeats the last tweak with a sequence number that loses every tie
                const bool pairFull = tweak + 1 < seeds;
                const int tweakY = pairFull ? tweak + 1 : tweak;
                const f2 tf0 = f2_make(ic.tweak[tweak][0], ic.tweak[tweakY][0]), tf1 = f2_make(ic.tweak[tweak][1], ic.tweak[tweakY][1]);
                f2 u0[NCH], u1[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    u0[ch] = f2_rne(f2_clamp_for_round(f2_add(f2_mul(tf0, offs[ch]), base[ch]), 0.0f, 255.0f));
                    u1[ch] = f2_rne(f2_clamp_for_round(f2_add(f2_mul(tf1, offs[ch]), base[ch]), 0.0f, 255.0f));
                }
                bc7_trial_pair<MODE, FAST, STRIDE>(P, gv, gw, n, sumV, staticAlphaError, u0, u1, 0, 0, 0, 0, tweak << 16, pairFull ? (tweakY << 16) : 0x7ff00000, best);
            }
        }

        const bool takeY = best.err.y < best.err.x || (best.err.y == best.err.x && best.seqY < best.seqX);
        out.err = takeY ? best.err.y : best.err.x;
        out.seq = takeY ? best.seqY : best.seqX;
        uint32_t r0 = 0, r1 = 0;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
        {
            r0 |= (as_uint(takeY ? best.e0[ch].y : best.e0[ch].x) & 0xffu) << (8 * ch);
            r1 |= (as_uint(takeY ? best.e1[ch].y : best.e1[ch].x) & 0xffu) << (8 * ch);
        }
        out.e0 = r0;
        out.e1 = r1;
    }

    // bc7_shape_trials for modes 6 / 7 under Flags::BC7_RespectPunchThrough, see BC7PunchCommit.  punchInvalid: bit pIter.
    template<int MODE, bool FAST, int STRIDE, class Vote>
    CVTT_HD void bc7_shape_trials_punch(const BC7Params &P, const F4 *gv, const F4 *gw, int n, int seeds, const float *base, const float *offs,
        const float *sumV, uint32_t punchInvalid, Vote &vote, BC7ShapeBest &out)
    {
        typedef BC7ModeT<MODE> M;
        enum { NCH = M::NCH };
        const IndexConst &ic = P.ic[M::IB - 2];

        BC7PunchCommit<NCH, Vote> commit;
        commit.vote = &vote;
        commit.err = FLT_MAX;
        commit.e0 = commit.e1 = 0;

#pragma unroll 1
        for (int pIter = 0; pIter < 4; pIter++)
        {
            commit.invalid = ((punchInvalid >> pIter) & 1) != 0;
            commit.groupAllInvalid = vote.all(commit.invalid);
            commit.groupAnyInvalid = vote.any(commit.invalid);
            if (!vote.warp_any(!commit.groupAllInvalid))
                continue;       // no group of the warp tries this parity combination
#pragma unroll 1
            for (int tweak = 0; tweak < seeds; tweak++)
            {
                const float tf0 = ic.tweak[tweak][0], tf1 = ic.tweak[tweak][1];
                f2 u0[NCH], u1[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ch++)
                {
                    u0[ch] = f2_splat(rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf0)), 0.0f, 255.0f)));
                    u1[ch] = f2_splat(rne(clamp_for_round(fadd(base[ch], fmul(offs[ch], tf1)), 0.0f, 255.0f)));
                }
                bc7_trial_pair<MODE, FAST, STRIDE>(P, gv, gw, n, sumV, 0.0f, u0, u1, pIter & 1, pIter & 1, pIter >> 1, pIter >> 1, 0, 0, commit);
            }
        }
        out.err = commit.err;
        out.e0 = commit.e0;
        out.e1 = commit.e1;
    }

    template<int MODE, bool FAST, int STRIDE, bool PUNCH, class Vote>
    CVTT_HD void bc7_shape_trials_67(const BC7Params &P, const F4 *gv, const F4 *gw, int n, int seeds, const float *base, const float *offs,
        const float *sumV, uint32_t punchInvalid, Vote &vote, BC7ShapeBest &out)
    {
        if (PUNCH)
            bc7_shape_trials_punch<MODE, FAST, STRIDE>(P, gv, gw, n, seeds, base, offs, sumV, punchInvalid, vote, out);
        else
            bc7_shape_trials<MODE, FAST, STRIDE>(P, gv, gw, n, seeds, base, offs, sumV, 0.0f, out);
    }

    // ---------------------------------------------------------------------------------------------------------
    // TrySingleColorRGBAMultiTable (BC67.cpp:940-1040) for one (mode, shape), evaluated after the shape's trials when
    // Flags::BC7_TrySingleColor is set (BC67.cpp:1436-1570).
    //
    // What the reference actually computes: its table loop accepts a candidate under
    //     better = ParallelMath::AndNot(pti, better)           (BC67.cpp:998)
    // and AndNot(a, b) is a & ~b (ParallelMath.h:901-906), i.e. "punch-through-invalid and NOT closer to the average".
    // Without punch-through invalidation pti is false, and with it the first comparison against FLT_MAX is always
    // "better", so no table entry is ever accepted: the single-colour tables (ConvectionKernels_BC7_SingleColor.h) are dead
    // data and the candidate that reaches the error test is always the initial one -- both endpoints (0, 0, 0[, 255]),
    // reconstructed colour (0, 0, 0[, 255]), index 0 for every pixel of the shape.  Bit-exactness means reproducing that:
    // the flag adds an "all black at index 0" candidate per (mode, shape).  (The average the reference computes over
    // pixels[0..n-1] instead of the shape's pixels, :1444-1446, therefore has no observable effect either.)
    // Returns true when the candidate replaces the shape's best; scIndex receives the index every pixel of the shape takes.
    template<int MODE, int STRIDE>
    CVTT_HD bool bc7_try_single_color(const BC7Params &P, const BC7Lane<STRIDE> &L, int n, float staticAlphaError, BC7ShapeBest &best, uint32_t &scIndex)
    {
        typedef BC7ModeT<MODE> M;
        enum { NCH = M::NCH };
        const int reconstructed[4] = { 0, 0, 0, 255 };
        int agg[4] = { 0, 0, 0, 0 };
        for (int i = 0; i < n; i++)
        {
            const F4 p = L.gv[i * STRIDE];
            const int pv[4] = { (int)(as_uint(p.x) & 0xffu), (int)(as_uint(p.y) & 0xffu), (int)(as_uint(p.z) & 0xffu), (int)(as_uint(p.w) & 0xffu) };
            for (int ch = 0; ch < NCH; ch++)
                agg[ch] += (reconstructed[ch] - pv[ch]) * (reconstructed[ch] - pv[ch]);
        }
        // AggregatedError<4>::Finalize over all four accumulators (the unused one is zero), then the static alpha error
        float error;
        if (P.flags & kFlag_Uniform)
            error = (float)(agg[0] + agg[1] + agg[2] + agg[3]);
        else
        {
            error = fmul((float)agg[0], P.wSq[0]);
            for (int ch = 1; ch < 4; ch++)
                error = fadd(error, fmul((float)agg[ch], P.wSq[ch]));
        }
        error = fadd(error, staticAlphaError);

        if (error < best.err)
        {
            best.err = error;
            best.e0 = best.e1 = (NCH == 4) ? 0xff000000u : 0u;
            scIndex = 0x80000000u;
            return true;
        }
        return false;
    }

    // ---------------------------------------------------------------------------------------------------------
    // One (mode, rotation, indexSelector) of TryDualPlane (BC67.cpp:1664-1965).  The caller has gathered the 16
    // pixels with the rotation's colour channel swapped with alpha, so .xyz is the rotated RGB and .w the scalar
    // plane; wr/wSqr/rcpWr are the weights permuted the same way (gw is weighted with wr).
    template<bool FAST, int STRIDE>
    CVTT_HD void bc7_dual_plane(const BC7Params &P, const F4 *gv, const F4 *gw, int mode, int rotation, int indexSelector, int seeds,
        const float *wr, const float *wSqr, const float *rcpWr, BC7Work &work)
    {
        const int R = P.refineRounds;
        const int rgbPrec = (mode == 4 && indexSelector) ? 3 : 2;
        const int alphaPrec = (mode == 4 && !indexSelector) ? 3 : 2;
        const IndexConst &icRGB = P.ic[rgbPrec - 2], &icA = P.ic[alphaPrec - 2];
        const QuantConst &qRGB = P.mc[mode].q;
        const float wN = 16.0f, wRcp = P.rcpN[16];

        float base[3], offs[3];
        bc7_endpoint_selector<3, STRIDE>(gw, 16, wr, base, offs);

        // alpha range, sums of the refiner's v terms (identical for every trial)
        float aMin = 0.0f, aMax = 0.0f, sumV[3] = { 0.0f, 0.0f, 0.0f }, sumA = 0.0f;
#pragma unroll 2
        for (int px = 0; px < 16; px++)
        {
            const F4 p = gv[px * STRIDE];
            const F4 q = gw[px * STRIDE];
            const float a = p.w - kMagic;
            if (px == 0)
                aMin = aMax = a;
            else
            {
                aMin = fminf(aMin, a);
                aMax = fmaxf(aMax, a);
            }
            sumV[0] = fadd(sumV[0], q.x);
            sumV[1] = fadd(sumV[1], q.y);
            sumV[2] = fadd(sumV[2], q.z);
            sumA = fadd(sumA, a);
        }

        float bestRGBError = FLT_MAX, bestAlphaError = FLT_MAX;
        float bRGB0[3] = { 0, 0, 0 }, bRGB1[3] = { 0, 0, 0 }, bA0 = 0.0f, bA1 = 0.0f;

#pragma unroll 1
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

#pragma unroll 1
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

#pragma unroll 2
                for (int px = 0; px < 16; px++)
                {
                    const F4 p = gv[px * STRIDE];
                    const float pv[4] = { p.x, p.y, p.z, p.w };

                    float dist = fmul(fsub(pv[0], om[0]), axis[0]);
                    dist = fadd(dist, fmul(fsub(pv[1], om[1]), axis[1]));
                    dist = fadd(dist, fmul(fsub(pv[2], om[2]), axis[2]));
                    float rgbIndex = rne(clamp_for_round(dist, 0.0f, icRGB.maxValue));
                    float alphaIndex = rne(clamp_for_round(fmul(fsub(pv[3], omA), axisA), 0.0f, icA.maxValue));

                    const float wf = xfma(rgbIndex, icRGB.wScale, kMagic) - kMagic;
                    const float wfA = xfma(alphaIndex, icA.wScale, kMagic) - kMagic;
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
                        // BC67.cpp:1822-1868 (wSqr is 1 under Flags::Uniform)
                        float rgbError = fadd(fadd(fmul(d2[0], wSqr[0]), fmul(d2[1], wSqr[1])), fmul(d2[2], wSqr[2]));
                        float alphaError = fmul(d2A, wSqr[3]);
                        const float altRGB[2] = { fmaxf(rgbIndex, 1.0f) - 1.0f, fminf(rgbIndex + 1.0f, icRGB.maxValue) };
                        const float altA[2] = { fmaxf(alphaIndex, 1.0f) - 1.0f, fminf(alphaIndex + 1.0f, icA.maxValue) };
#pragma unroll
                        for (int ii = 0; ii < 2; ii++)
                        {
                            const float awf = xfma(altRGB[ii], icRGB.wScale, kMagic) - kMagic;
                            const float awfA = xfma(altA[ii], icA.wScale, kMagic) - kMagic;
                            float s[3];
#pragma unroll
                            for (int ch = 0; ch < 3; ch++)
                            {
                                const float df = (xfma(awf, d64[ch], bq[ch]) + kMagic) - pv[ch];
                                s[ch] = df * df;
                            }
                            const float dfA = (xfma(awfA, d64A, bqA) + kMagic) - pv[3];
                            const float sA = dfA * dfA;
                            const float altRGBError = fadd(fadd(fmul(s[0], wSqr[0]), fmul(s[1], wSqr[1])), fmul(s[2], wSqr[2]));
                            const float altAlphaError = fmul(sA, wSqr[3]);
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
                        const F4 q = gw[px * STRIDE];
                        const float qv[3] = { q.x, q.y, q.z };
                        const float t = fmul(rgbIndex, icRGB.rcpMaxIndex);
#pragma unroll
                        for (int ch = 0; ch < 3; ch++)
                            tv[ch] = fadd(tv[ch], fmul(t, qv[ch]));
                        tt = fadd(tt, fmul(t, t));
                        ts = fadd(ts, t);
                        const float tA = fmul(alphaIndex, icA.rcpMaxIndex);
                        tvA = fadd(tvA, fmul(tA, fsub(pv[3], kMagic)));
                        ttA = fadd(ttA, fmul(tA, tA));
                        tsA = fadd(tsA, tA);
                    }
                }

                float errorRGB, errorA;
                if (FAST)
                {
                    if (P.flags & kFlag_Uniform)
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
                            const float a = fdiv(fsub(tv[ch], fmul(fmul(ts, sumV[ch]), wRcp)), adenom);
                            const float b = fmul(fsub(sumV[ch], fmul(a, ts)), wRcp);
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
                        const float a = fdiv(fsub(tvA, fmul(fmul(tsA, sumA), wRcp)), adenom);
                        const float b = fmul(fsub(sumA, fmul(a, tsA)), wRcp);
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
            work.sc[0] = work.sc[1] = work.sc[2] = 0;
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // Indexes of the winning configuration.  Index selection (IndexSelector.h:124-131, and the +-1 probe of
    // BC67.cpp:1364-1386 / 1822-1868 without FastIndexing) is a pure function of the quantised endpoints and the
    // pixel, so the trial loops above do not carry indexes; they are re-derived here, once per block, with exactly
    // the operations of the trial that produced the endpoints.
    //   first/nch: the channels of `p` this selector covers (RGB: 0..2, RGBA: 0..3, the dual-plane scalar: 3)
    struct BC7IndexSelector
    {
        float om[4], axis[4], d64[4], bq[4], wSq[4];
        float maxV, wScale;
        int first, nch;
    };

    CVTT_HD void bc7_selector_init(BC7IndexSelector &S, int first, int nch, int indexBits, uint32_t ep0, uint32_t ep1, const float *w, const float *wSq)
    {
        S.first = first;
        S.nch = nch;
        S.maxV = (float)((1 << indexBits) - 1);
        S.wScale = 64.0f / (float)((1 << indexBits) - 1);
        float dW[4], q0[4], q1[4];
        for (int k = 0; k < nch; k++)
        {
            const int ch = first + k;
            q0[k] = (float)((ep0 >> (8 * ch)) & 0xffu);
            q1[k] = (float)((ep1 >> (8 * ch)) & 0xffu);
            dW[k] = fmul(fsub(q1[k], q0[k]), w[ch]);
            S.wSq[k] = wSq[ch];
        }
        float lenSq = fmul(dW[0], dW[0]);
        for (int k = 1; k < nch; k++)
            lenSq = fadd(lenSq, fmul(dW[k], dW[k]));
        safe_denominator(lenSq);
        const float mdl = fdiv(S.maxV, lenSq);
        for (int k = 0; k < nch; k++)
        {
            const int ch = first + k;
            // the scalar plane's selector is IndexSelector<1> with unit weight: axis = d * (maxV / d^2)
            S.axis[k] = fmul(fmul(dW[k], w[ch]), mdl);
            S.om[k] = q0[k] + kMagic;
            S.d64[k] = (q1[k] - q0[k]) * 0.015625f;
            S.bq[k] = q0[k] + 0.0078125f;
        }
    }

    template<bool FAST>
    CVTT_HD uint32_t bc7_select_index(const BC7IndexSelector &S, const F4 &p)
    {
        const float pAll[4] = { p.x, p.y, p.z, p.w };
        float pv[4];
        for (int k = 0; k < S.nch; k++)
            pv[k] = pAll[S.first + k];
        float dist = fmul(fsub(pv[0], S.om[0]), S.axis[0]);
        for (int k = 1; k < S.nch; k++)
            dist = fadd(dist, fmul(fsub(pv[k], S.om[k]), S.axis[k]));
        float idxf = rne(clamp_for_round(dist, 0.0f, S.maxV));
        if (!FAST)
        {
            float error = 0.0f;
            {
                const float wf = xfma(idxf, S.wScale, kMagic) - kMagic;
                for (int k = 0; k < S.nch; k++)
                {
                    const float df = (xfma(wf, S.d64[k], S.bq[k]) + kMagic) - pv[k];
                    const float sq = df * df;
                    error = (k == 0) ? fmul(sq, S.wSq[0]) : fadd(error, fmul(sq, S.wSq[k]));
                }
            }
            const float alt[2] = { fmaxf(idxf, 1.0f) - 1.0f, fminf(idxf + 1.0f, S.maxV) };
            for (int ii = 0; ii < 2; ii++)
            {
                const float awf = xfma(alt[ii], S.wScale, kMagic) - kMagic;
                float altError = 0.0f;
                for (int k = 0; k < S.nch; k++)
                {
                    const float df = (xfma(awf, S.d64[k], S.bq[k]) + kMagic) - pv[k];
                    const float sq = df * df;
                    altError = (k == 0) ? fmul(sq, S.wSq[0]) : fadd(altError, fmul(sq, S.wSq[k]));
                }
                const bool better = altError < error;
                error = sse_min(error, altError);
                if (better)
                    idxf = alt[ii];
            }
        }
        return as_uint(idxf + kMagic) & 15u;
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
    // Gathers the pixels of `mask` (ascending pixel order = the reference's iteration order) into gv/gw, optionally
    // rotated, weighted with wv.  Returns the sums the trials share: the refiner's per-channel sum of v
    // (EndpointRefiner.h:85-88) and the squared error of replacing alpha by 255 (BC67.cpp:1250-1264).
    template<int STRIDE>
    CVTT_HD void bc7_gather(const BC7Lane<STRIDE> &L, uint32_t mask, int rotation, const float *wv, float *sumV, float &accA)
    {
        sumV[0] = sumV[1] = sumV[2] = sumV[3] = 0.0f;
        accA = 0.0f;
        int i = 0;
        for (uint32_t m = mask; m; m &= m - 1, i++)
        {
            const int px = ctz32(m);
            F4 p = bc7_expand_pixel(L.raw[px * STRIDE]);
            if (rotation)
                bc7_rotate(p, rotation);
            F4 q;
            q.x = fmul(p.x - kMagic, wv[0]);
            q.y = fmul(p.y - kMagic, wv[1]);
            q.z = fmul(p.z - kMagic, wv[2]);
            q.w = fmul(p.w - kMagic, wv[3]);
            L.gv[i * STRIDE] = p;
            L.gw[i * STRIDE] = q;
            sumV[0] = fadd(sumV[0], q.x);
            sumV[1] = fadd(sumV[1], q.y);
            sumV[2] = fadd(sumV[2], q.z);
            sumV[3] = fadd(sumV[3], q.w);
            const float da = (255.0f + kMagic) - p.w;
            accA = xfma(da, da, accA);
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // Endpoint fits of one gathered pixel subset (the first half of a SHAPE command).  Shapes the plan does not list keep the
    // all-zero "unfinished" endpoints the zero-initialised reference build has (SinglePlaneTemporaries, BC67.cpp:803-811).
    // allowRGBModes / usePCA4 are the votes of the block the pixels belong to; anyRGB / anyPCA4 only say whether some lane of
    // the warp needs the fit at all (pure work skipping).
    template<int STRIDE>
    CVTT_HD void bc7_shape_fits(const BC7Params &P, const F4 *gw, int n, bool listedRGB, bool listedRGBA, bool needRGBA,
        bool allowRGBModes, bool usePCA4, bool anyRGB, bool anyPCA4, float *baseRGB, float *offsRGB, float *baseRGBA, float *offsRGBA)
    {
        for (int ch = 0; ch < 3; ch++)
            baseRGB[ch] = offsRGB[ch] = 0.0f;
        if (listedRGB && anyRGB)
        {
            float b3[3], o3[3];
            bc7_endpoint_selector<3, STRIDE>(gw, n, P.w, b3, o3);
            if (allowRGBModes)                                       // BC67.cpp:1085
                for (int ch = 0; ch < 3; ch++)
                {
                    baseRGB[ch] = b3[ch];
                    offsRGB[ch] = o3[ch];
                }
        }
        for (int ch = 0; ch < 4; ch++)
            baseRGBA[ch] = offsRGBA[ch] = 0.0f;
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
            if (anyPCA4)
            {
                float b4[4], o4[4];
                bc7_endpoint_selector<4, STRIDE>(gw, n, P.w, b4, o4);
                if (usePCA4)
                    for (int ch = 0; ch < 4; ch++)
                    {
                        baseRGBA[ch] = b4[ch];
                        offsRGBA[ch] = o4[ch];
                    }
            }
        }
    }

    // the trials of one (mode, gathered subset) of the two-subset modes, without the punch-through and single-colour variants
    template<bool FAST, int STRIDE>
    CVTT_HD void bc7_run_pair_mode(const BC7Params &P, int mode, const F4 *gv, const F4 *gw, int n, int seeds, const float *baseRGB, const float *offsRGB,
        const float *baseRGBA, const float *offsRGBA, const float *sumV, float staticAlphaError, BC7ShapeBest &best, int ppFirst = 0, int ppEnd = 2)
    {
        if (mode == 1)
            bc7_shape_trials<1, FAST, STRIDE>(P, gv, gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best, ppFirst, ppEnd);
        else if (mode == 3)
            bc7_shape_trials<3, FAST, STRIDE>(P, gv, gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best, ppFirst, ppEnd);
        else
            bc7_shape_trials<7, FAST, STRIDE>(P, gv, gw, n, seeds, baseRGBA, offsRGBA, sumV, 0.0f, best, ppFirst, ppEnd);
    }

    // ---------------------------------------------------------------------------------------------------------
    // PAIR2 commands (two-subset modes 1 / 3 / 7, one partition): subset A is searched by every block's own thread; subset B
    // only for the (block, run) pairs that can still win, as TASKS that are dealt out to the warps of the CTA, so that the
    // warps without a task do not walk those trials at all.  A task is one class of work for one block: class = run * 2 +
    // unit, where a run with four parity combinations may be split into two units (its two parity-pair iterations).  All
    // tasks of a 32-slot chunk have the same class, so a warp executes one mode's trials at a time.
    //
    // The exchange between a block's owner thread and the threads that run its tasks goes through this interface.  The
    // kernel implements it with shared memory (bc7_kernels.cu); a single lane (tests/hostsim, and kernels whose command
    // streams hold no PAIR2) runs its own tasks one after the other.
    // Up to kBC7PairGroup consecutive PAIR2 commands are searched as one group: their first halves back to back, ONE exchange
    // for the tasks of all of them.  A partition leaves 4-6 chunks of tasks for the 12 warps of a CTA, so two partitions'
    // tasks run in the time of one (a block's best is then one partition stale when the second one's wants are decided).
    enum { kBC7PairGroup = 2, kBC7PairClassesPerCommand = 6, kBC7PairClasses = kBC7PairGroup * kBC7PairClassesPerCommand };
    // TRIPLE commands (three-subset modes 0 / 2, one partition each) form groups of up to eight commands with up to eight runs
    // between them: class = the run's number in the group (a nibble map gives command and run back)
    enum { kBC7TripleGroup = 8, kBC7TripleClasses = 8 };

    struct BC7SoloExchange
    {
        F4 posted[kBC7PairClasses];
        uint32_t flags, want;
        const uint32_t *raw;
        // the owner's group votes, for whoever fits its pixels: bit 0 allowRGBModes, bit 1 usePCA4
        CVTT_HD void publish(uint32_t f) { flags = f; }
        // wantMask: bit c = this block wants class c searched.  Returns the number of task slots (padded to whole chunks).
        CVTT_HD int compact(uint32_t wantMask) { want = wantMask; return kBC7PairClasses; }
        CVTT_HD int first_slot() const { return 0; }
        CVTT_HD int slot_stride() const { return 1; }
        CVTT_HD bool chunk_in_range(int slot, int total) const { return slot < total; }
        // the task in `slot`: its class (the same for the whole chunk) and the owner thread, -1 for a padding slot
        CVTT_HD void task(int slot, int &cls, int &owner) const { cls = slot; owner = ((want >> slot) & 1u) ? 0 : -1; }
        CVTT_HD uint32_t owner_flags(int) const { return flags; }
        CVTT_HD const uint32_t *owner_raw(int) const { return raw; }
        CVTT_HD bool task_any(bool x) const { return x; }
        CVTT_HD void post(int, int cls, const F4 &r) { posted[cls] = r; }
        CVTT_HD void sync() {}
        CVTT_HD F4 result(int cls) const { return posted[cls]; }
        // TRIPLE commands post two words per class (classes 0 .. kBC7TripleClasses - 1)
        F4 posted2[kBC7PairClasses][2];
        CVTT_HD void post2(int, int cls, const F4 &a, const F4 &b) { posted2[cls][0] = a; posted2[cls][1] = b; }
        CVTT_HD void result2(int cls, F4 &a, F4 &b) const { a = posted2[cls][0]; b = posted2[cls][1]; }
    };

    // Second half of a group of PAIR2 commands: runs the task slots dealt to this thread.  pcs point at the commands; class =
    // command * 6 + run * 2 + unit.
    template<bool FAST, int STRIDE, class Exchange>
    CVTT_HD void bc7_pair2_tasks(const BC7Params &P, const BC7Lane<STRIDE> &L, Exchange &ex, const uint32_t *const *pcs, int total)
    {
        for (int slot = ex.first_slot(); ex.chunk_in_range(slot, total); slot += ex.slot_stride())
        {
            int cls, owner;
            ex.task(slot, cls, owner);
            const bool hasTask = owner >= 0;
            if (!ex.task_any(hasTask))
                continue;
            const int command = cls >= kBC7PairClassesPerCommand ? 1 : 0, local = cls - command * kBC7PairClassesPerCommand;
            const uint32_t *pc = pcs[command];
            const uint32_t w0 = pc[0], w2 = pc[2];
            const bool listedRGB = (w0 >> 19) & 1, listedRGBA = (w0 >> 20) & 1, split = (w0 >> 22) & 1;
            const uint32_t maskB = w2 & 0xffffu;
            const int nB = (w2 >> 16) & 0xff;
            const int r = local >> 1, unit = local & 1;
            const uint32_t rw = pc[3 + r];
            const int mode = rw & 0xf, seeds = (rw >> 8) & 0xf;
            const int o = hasTask ? owner : 0;
            const uint32_t of = hasTask ? ex.owner_flags(o) : 0u;
            const bool allowRGBModes = (of & 1u) != 0, usePCA4 = (of & 2u) != 0;

            BC7Lane<STRIDE> LB = L;
            LB.raw = ex.owner_raw(o);
            float sumV[4], accA;
            bc7_gather<STRIDE>(LB, maskB, 0, P.w, sumV, accA);
            const float staticAlphaError = (P.flags & kFlag_Uniform) ? accA : fmul(accA, P.wSq[3]);
            // the fits this mode reads: RGB modes the 3-channel one; mode 7 the 4-channel one, or the expanded 3-channel one
            // for a block whose group has no alpha (bc7_shape_fits)
            const bool rgba = (mode == 7);
            const bool anyRGB = ex.task_any(hasTask && (rgba ? !usePCA4 : allowRGBModes));
            const bool anyPCA4 = rgba && ex.task_any(hasTask && usePCA4);
            float baseRGB[3], offsRGB[3], baseRGBA[4], offsRGBA[4];
            bc7_shape_fits<STRIDE>(P, L.gw, nB, listedRGB, listedRGBA, rgba, allowRGBModes, usePCA4, anyRGB, anyPCA4, baseRGB, offsRGB, baseRGBA, offsRGBA);

            BC7ShapeBest best;
            const bool wide = (mode != 1) && split;          // four parity combinations, searched as two units
            bc7_run_pair_mode<FAST, STRIDE>(P, mode, L.gv, L.gw, nB, seeds, baseRGB, offsRGB, baseRGBA, offsRGBA, sumV, staticAlphaError, best,
                                            wide ? unit : 0, wide ? unit + 1 : 2);
            if (hasTask)
            {
                F4 res;
                res.x = best.err;
                res.y = as_float((uint32_t)best.seq);
                res.z = as_float(best.e0);
                res.w = as_float(best.e1);
                ex.post(owner, cls, res);
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // TRIPLE commands (three-subset modes 0 / 2, one partition): the partition's LARGEST subset (the anchor) has been searched
    // for every block by a SHAPE command; the other two subsets are searched only for the blocks whose anchor error leaves
    // room below their best (total = sum of three non-negative errors >= the anchor's, fl(a + b) monotonic), as tasks dealt out
    // to the warps of the CTA exactly like the second halves of PAIR2 commands.  Per block 95-99 % of them are dead.
    //   w0: TRIPLE | runs << 8 | anchor subset << 16 | B listed << 18 | C listed << 19 | partition << 24
    //   w1: mask of subset B | pixels << 16 | subset number << 24;   w2: the same for subset C
    //   run word: mode | seeds of B << 4 | seeds of C << 8 | result slot of the anchor << 16
    // A task posts (err B, err C) and (B endpoints, C endpoints).
    template<bool FAST, int STRIDE, class Exchange>
    CVTT_HD void bc7_triple_tasks(const BC7Params &P, const BC7Lane<STRIDE> &L, Exchange &ex, const uint32_t *const *pcs, uint32_t classMap, int total)
    {
        for (int slot = ex.first_slot(); ex.chunk_in_range(slot, total); slot += ex.slot_stride())
        {
            int cls, owner;
            ex.task(slot, cls, owner);
            const bool hasTask = owner >= 0;
            if (!ex.task_any(hasTask))
                continue;
            const uint32_t entry = (classMap >> (4 * (cls & (kBC7TripleClasses - 1)))) & 15u;      // command << 1 | run
            const uint32_t *pc = pcs[entry >> 1];
            const uint32_t w0 = pc[0], rw = pc[3 + (entry & 1u)];
            const int mode = rw & 0xf;
            BC7Lane<STRIDE> LB = L;
            LB.raw = ex.owner_raw(hasTask ? owner : 0);
            BC7ShapeBest found[2];
            for (int half = 0; half < 2; half++)
            {
                const uint32_t w = pc[1 + half];
                const int n = (w >> 16) & 0xff, seeds = (rw >> (4 + 4 * half)) & 0xf;
                const bool listed = (w0 >> (18 + half)) & 1;
                float sumV[4], accA;
                bc7_gather<STRIDE>(LB, w & 0xffffu, 0, P.w, sumV, accA);
                const float staticAlphaError = (P.flags & kFlag_Uniform) ? accA : fmul(accA, P.wSq[3]);
                float baseRGB[3], offsRGB[3], baseRGBA[4], offsRGBA[4];
                // only blocks whose group allows the RGB modes have tasks
                bc7_shape_fits<STRIDE>(P, L.gw, n, listed, false, false, true, false, true, false, baseRGB, offsRGB, baseRGBA, offsRGBA);
                if (mode == 0)
                    bc7_shape_trials<0, FAST, STRIDE>(P, L.gv, L.gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, found[half]);
                else
                    bc7_shape_trials<2, FAST, STRIDE>(P, L.gv, L.gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, found[half]);
            }
            if (hasTask)
            {
                F4 e, ep;
                e.x = found[0].err;
                e.y = found[1].err;
                e.z = e.w = 0.0f;
                ep.x = as_float(found[0].e0);
                ep.y = as_float(found[0].e1);
                ep.z = as_float(found[1].e0);
                ep.w = as_float(found[1].e1);
                ex.post2(owner, cls, e, ep);
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // The whole search for one block.
    struct BC7NoVote      // the search without Flags::BC7_RespectPunchThrough has no per-trial group votes
    {
        CVTT_HD bool any(bool x) const { return x; }
        CVTT_HD bool all(bool x) const { return x; }
        CVTT_HD bool warp_any(bool x) const { return x; }
    };

    // A block's best candidate so far as two 128-bit words (the small-call launch searches a block in several CTAs and
    // reduces their winners; streams with single-colour candidates never take that launch, so `sc` does not travel).
    struct BC7Candidate
    {
        uint32_t a[4], b[4];      // a: error bits, key | mode << 12 | sub << 16 (bit 31: nothing committed), ep[0][0], ep[0][1]; b: ep[1][0..1], ep[2][0..1]
    };

    CVTT_HD void bc7_work_reset(BC7Work &work)
    {
        work.error = FLT_MAX;
        work.key = -1;
        work.mode = 0;
        work.sub = 0;
        for (int s = 0; s < 3; s++)
            work.ep[s][0] = work.ep[s][1] = 0;
        work.idx[0] = work.idx[1] = work.idx2[0] = work.idx2[1] = 0;
        work.sc[0] = work.sc[1] = work.sc[2] = 0;
    }

    CVTT_HD void bc7_candidate_pack(const BC7Work &work, BC7Candidate &c)
    {
        c.a[0] = as_uint(work.error);
        c.a[1] = (work.key < 0) ? 0x80000000u : ((uint32_t)work.key | ((uint32_t)work.mode << 12) | ((uint32_t)work.sub << 16));
        c.a[2] = work.ep[0][0]; c.a[3] = work.ep[0][1];
        c.b[0] = work.ep[1][0]; c.b[1] = work.ep[1][1]; c.b[2] = work.ep[2][0]; c.b[3] = work.ep[2][1];
    }

    // lexicographic (error, reference key) minimum: the same winner as one thread walking every trial in any order
    CVTT_HD void bc7_candidate_merge(BC7Work &work, const BC7Candidate &c)
    {
        if (c.a[1] & 0x80000000u)
            return;
        const float error = as_float(c.a[0]);
        const int key = (int)(c.a[1] & 0xfffu);
        if (error < work.error || (error == work.error && key < work.key))
        {
            work.error = error;
            work.key = key;
            work.mode = (int)((c.a[1] >> 12) & 0xfu);
            work.sub = (int)((c.a[1] >> 16) & 0xffu);
            work.ep[0][0] = c.a[2]; work.ep[0][1] = c.a[3];
            work.ep[1][0] = c.b[0]; work.ep[1][1] = c.b[1]; work.ep[2][0] = c.b[2]; work.ep[2][1] = c.b[3];
            work.sc[0] = work.sc[1] = work.sc[2] = 0;
        }
    }

    // The search: walks the command stream at `pc` and leaves the block's best (mode, partition / rotation, endpoints) in work.
    // PUNCH: Flags::BC7_RespectPunchThrough; vote = the reference's AnySet / AllSet over the 8 blocks of one call
    template<bool FAST, int STRIDE, bool PUNCH, class Vote, class Exchange>
    CVTT_HD void bc7_search_block(const BC7Params &P, const BC7Lane<STRIDE> &L, const BC7LaneFlags &lf, Vote &vote, Exchange &ex, const uint32_t *pc, BC7Work &work)
    {
        bc7_work_reset(work);

        // per-(mode, shape) results, indexed by the slot numbers the host assigned: error, endpoint 0, endpoint 1, single-colour marker
        uint32_t res[kBC7MaxSlots][4];

        const bool trySingleColor = (P.flags & kFlag_BC7_TrySingleColor) != 0;
        // BC67.cpp:1286-1295: parity 0 cannot keep alpha 255, parity 3 cannot keep alpha 0, mixed parities keep neither
        const uint32_t punchInvalid67 = !lf.isPunchThrough ? 0u : ((lf.blockHasNonZeroAlpha ? 1u : 0u) | 6u | (lf.blockHasNonMaxAlpha ? 8u : 0u));

        const bool usePCA4 = lf.anyBlockHasAlpha || !lf.allowRGBModes;                   // BC67.cpp:1121
        const bool allowMode7 = lf.anyBlockHasAlpha || (P.mode7RGBPartitionEnabled != 0); // BC67.cpp:1078
        const bool uniform = (P.flags & kFlag_Uniform) != 0;

        for (;;)
        {
            // Every warp of the CTA walks the same command stream; keeping them on the same command keeps the code they
            // execute (one mode's trial loops, a few KB) resident in the SM's 32 KB instruction cache.
            const uint32_t w0 = pc[0];
            const int op = w0 & 0xff;
            if (op != kCmdEval)          // EVAL is a handful of instructions; the stream is uniform, so every thread agrees
                cta_sync();
            if (op == kCmdEnd)
                break;
            if (!lf.warpHasWork)
            {
                // a warp without blocks (the dealt-out tail of a CTA) only walks the stream for the barriers
                if (op == kCmdPair2)
                {
                    // no block, no task to offer, but the exchange's barriers and the task warps need every warp of the CTA
                    const uint32_t *pcs[kBC7PairGroup];
                    int commands = 0;
                    while (commands < kBC7PairGroup && (pc[0] & 0xff) == kCmdPair2)
                    {
                        if (commands)
                            cta_sync();             // the rendezvous the working warps keep before a command of the group
                        pcs[commands++] = pc;
                        pc += 3 + (int)((pc[0] >> 8) & 0xff);
                    }
                    for (int k = commands; k < kBC7PairGroup; k++)
                        pcs[k] = pcs[0];
                    ex.publish(0u);
                    const int total = ex.compact(0u);
                    bc7_pair2_tasks<FAST, STRIDE>(P, L, ex, pcs, total);
                    ex.sync();
                    continue;
                }
                if (op == kCmdTriple)
                {
                    const uint32_t *pcs[kBC7TripleGroup];
                    uint32_t classMap = 0;
                    int commands = 0, classes = 0;
                    while (commands < kBC7TripleGroup && (pc[0] & 0xff) == kCmdTriple && classes + (int)((pc[0] >> 8) & 0xff) <= kBC7TripleClasses)
                    {
                        const int nRuns = (pc[0] >> 8) & 0xff;
                        for (int r = 0; r < nRuns; r++)
                            classMap |= (uint32_t)(commands * 2 + r) << (4 * classes++);
                        pcs[commands++] = pc;
                        pc += 3 + nRuns;
                    }
                    for (int k = commands; k < kBC7TripleGroup; k++)
                        pcs[k] = pcs[0];
                    ex.publish(0u);
                    const int total = ex.compact(0u);
                    bc7_triple_tasks<FAST, STRIDE>(P, L, ex, pcs, classMap, total);
                    ex.sync();
                    continue;
                }
                pc += (op == kCmdShape) ? 2 + (int)((w0 >> 8) & 0xff) : (op == kCmdEval ? 2 : 1);
                continue;
            }

            if (op == kCmdShape)
            {
                const int nRuns = (w0 >> 8) & 0xff;
                const bool listedRGB = (w0 >> 16) & 1, listedRGBA = (w0 >> 17) & 1, needRGBA = (w0 >> 18) & 1;
                const uint32_t w1 = pc[1];
                const uint32_t mask = w1 & 0xffffu;
                const int n = (w1 >> 16) & 0xff;

                float sumV[4], accA;
                bc7_gather<STRIDE>(L, mask, 0, P.w, sumV, accA);
                const float staticAlphaError = uniform ? accA : fmul(accA, P.wSq[3]);

                float baseRGB[3], offsRGB[3], baseRGBA[4], offsRGBA[4];
                bc7_shape_fits<STRIDE>(P, L.gw, n, listedRGB, listedRGBA, needRGBA, lf.allowRGBModes, usePCA4, lf.warpAnyRGB, lf.warpAnyPCA4, baseRGB, offsRGB, baseRGBA, offsRGBA);

                for (int r = 0; r < nRuns; r++)
                {
                    const uint32_t rw = pc[2 + r];
                    const int mode = rw & 0xf, seeds = (rw >> 4) & 0xf, slot = (rw >> 8) & 0xff;
                    BC7ShapeBest best;
                    uint32_t scIndex = 0;
                    if (mode < 4)
                    {
                        if (!lf.warpAnyRGB)
                            continue;
                        switch (mode)
                        {
                        case 0:
                            bc7_shape_trials<0, FAST, STRIDE>(P, L.gv, L.gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best);
                            if (trySingleColor)
                                bc7_try_single_color<0, STRIDE>(P, L, n, staticAlphaError, best, scIndex);
                            break;
                        case 1:
                            bc7_shape_trials<1, FAST, STRIDE>(P, L.gv, L.gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best);
                            if (trySingleColor)
                                bc7_try_single_color<1, STRIDE>(P, L, n, staticAlphaError, best, scIndex);
                            break;
                        case 2:
                            bc7_shape_trials<2, FAST, STRIDE>(P, L.gv, L.gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best);
                            if (trySingleColor)
                                bc7_try_single_color<2, STRIDE>(P, L, n, staticAlphaError, best, scIndex);
                            break;
                        default:
                            bc7_shape_trials<3, FAST, STRIDE>(P, L.gv, L.gw, n, seeds, baseRGB, offsRGB, sumV, staticAlphaError, best);
                            if (trySingleColor)
                                bc7_try_single_color<3, STRIDE>(P, L, n, staticAlphaError, best, scIndex);
                            break;
                        }
                    }
                    else if (mode == 6)
                    {
                        bc7_shape_trials_67<6, FAST, STRIDE, PUNCH>(P, L.gv, L.gw, n, seeds, baseRGBA, offsRGBA, sumV, punchInvalid67, vote, best);
                        if (trySingleColor)
                            bc7_try_single_color<6, STRIDE>(P, L, n, 0.0f, best, scIndex);
                    }
                    else
                    {
                        if (!lf.warpAnyMode7)
                            continue;
                        bc7_shape_trials_67<7, FAST, STRIDE, PUNCH>(P, L.gv, L.gw, n, seeds, baseRGBA, offsRGBA, sumV, punchInvalid67, vote, best);
                        if (trySingleColor)
                            bc7_try_single_color<7, STRIDE>(P, L, n, 0.0f, best, scIndex);
                    }
                    res[slot][0] = as_uint(best.err);
                    res[slot][1] = best.e0;
                    res[slot][2] = best.e1;
                    res[slot][3] = scIndex;
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
                    for (int s = 0; s < numSubsets; s++)
                    {
                        work.ep[s][0] = res[slots[s]][1];
                        work.ep[s][1] = res[slots[s]][2];
                        work.sc[s] = res[slots[s]][3];
                    }
                }
            }
            else if (op == kCmdPair2)
            {
                // Two-subset modes, one partition per command: subset A (the larger one) for every block, subset B only where A's
                // error leaves room below the block's best.  total = err(A) + err(B) >= err(A) (errors are sums of non-negative
                // terms and fl(a + b) is monotonic), so a block whose err(A) already exceeds its best cannot take this
                // (mode, partition) whatever B gives -- the reference evaluates it and rejects it.  Consecutive commands form a
                // group (kBC7PairGroup): first halves back to back, one exchange for the second halves of all of them.
                // The first halves' results wait in result slots 0 .. 5 (free in this stream form once mode 6 has been scanned):
                // nothing but the wanted classes is carried in registers across the task phase.
                const uint32_t *pcs[kBC7PairGroup];
                uint32_t wantMask = 0;                              // bit (command * 6 + run * 2 + unit)
                int commands = 0;
                while (commands < kBC7PairGroup && (pc[0] & 0xff) == kCmdPair2)
                {
                    if (commands)
                        cta_sync();
                    const int k = commands++;
                    pcs[k] = pc;
                    const uint32_t c0 = pc[0];
                    const int nRuns = (c0 >> 8) & 0xff, partition = (c0 >> 24) & 0x3f;
                    const bool listedRGB = (c0 >> 16) & 1, listedRGBA = (c0 >> 17) & 1, needRGBA = (c0 >> 18) & 1;
                    const uint32_t w1 = pc[1];
                    const uint32_t maskA = w1 & 0xffffu;
                    const int nA = (w1 >> 16) & 0xff;

                    float sumV[4], accA;
                    bc7_gather<STRIDE>(L, maskA, 0, P.w, sumV, accA);
                    const float staticAlphaError = uniform ? accA : fmul(accA, P.wSq[3]);
                    float baseRGB[3], offsRGB[3], baseRGBA[4], offsRGBA[4];
                    bc7_shape_fits<STRIDE>(P, L.gw, nA, listedRGB, listedRGBA, needRGBA, lf.allowRGBModes, usePCA4, lf.warpAnyRGB, lf.warpAnyPCA4, baseRGB, offsRGB, baseRGBA, offsRGBA);

                    const bool split = (c0 >> 22) & 1;
                    for (int r = 0; r < nRuns; r++)
                    {
                        const uint32_t rw = pc[3 + r];
                        const int mode = rw & 0xf, seeds = (rw >> 4) & 0xf;
                        if ((mode < 4 && !lf.warpAnyRGB) || (mode == 7 && !lf.warpAnyMode7))
                            continue;
                        BC7ShapeBest best;
                        bc7_run_pair_mode<FAST, STRIDE>(P, mode, L.gv, L.gw, nA, seeds, baseRGB, offsRGB, baseRGBA, offsRGBA, sumV, staticAlphaError, best);
                        bool eligible = true;                              // the lane conditions of the partition scan, BC67.cpp:1602-1634
                        if (mode < 4 && !lf.allowRGBModes)
                            eligible = false;
                        if (mode == 7)
                        {
                            if (!allowMode7)
                                eligible = false;
                            if (lf.anyBlockHasAlpha && ((P.mode7RGBPartitionEnabled >> partition) & 1) == 0 && !lf.blockHasNonMaxAlpha)
                                eligible = false;
                        }
                        if (eligible && !(best.err > work.error))
                        {
                            res[k * 3 + r][0] = as_uint(best.err);
                            res[k * 3 + r][1] = best.e0;
                            res[k * 3 + r][2] = best.e1;
                            wantMask |= ((mode != 1 && split) ? 3u : 1u) << (k * kBC7PairClassesPerCommand + 2 * r);
                        }
                    }
                    pc += 3 + nRuns;
                }
                for (int k = commands; k < kBC7PairGroup; k++)
                    pcs[k] = pcs[0];
                ex.publish((lf.allowRGBModes ? 1u : 0u) | (usePCA4 ? 2u : 0u));
                const int total = ex.compact(wantMask);
                bc7_pair2_tasks<FAST, STRIDE>(P, L, ex, pcs, total);
                ex.sync();
                for (int k = 0; k < commands; k++)
                {
                    const uint32_t c0 = pcs[k][0];
                    const int nRuns = (c0 >> 8) & 0xff, partition = (c0 >> 24) & 0x3f;
                    const bool aIsSubset1 = (c0 >> 21) & 1;
                    for (int r = 0; r < nRuns; r++)
                    {
                        const int cls = k * kBC7PairClassesPerCommand + 2 * r;
                        if (!((wantMask >> cls) & 1u))
                            continue;
                        const int mode = pcs[k][3 + r] & 0xf;
                        F4 got = ex.result(cls);
                        if ((wantMask >> (cls + 1)) & 1u)
                        {
                            // the run was searched as two units: "first strictly better in sequence order" over both
                            const F4 other = ex.result(cls + 1);
                            if (other.x < got.x || (other.x == got.x && (int)as_uint(other.y) < (int)as_uint(got.y)))
                                got = other;
                        }
                        const float errA = as_float(res[k * 3 + r][0]);
                        const float totalError = aIsSubset1 ? fadd(got.x, errA) : fadd(errA, got.x);       // subset 0 + subset 1
                        const int key = bc7_mode_order(mode) * 64 + partition;
                        if (totalError < work.error || (totalError == work.error && key < work.key))
                        {
                            work.error = totalError;
                            work.key = key;
                            work.mode = mode;
                            work.sub = partition;
                            const int sA = aIsSubset1 ? 1 : 0, sB = 1 - sA;
                            work.ep[sA][0] = res[k * 3 + r][1];
                            work.ep[sA][1] = res[k * 3 + r][2];
                            work.ep[sB][0] = as_uint(got.z);
                            work.ep[sB][1] = as_uint(got.w);
                            work.sc[0] = work.sc[1] = work.sc[2] = 0;
                        }
                    }
                }
            }
            else if (op == kCmdTriple)
            {
                // the anchors' results stay in their slots: nothing but the wanted classes is carried across the task phase
                const uint32_t *pcs[kBC7TripleGroup];
                uint32_t wantMask = 0, classMap = 0;                // bit (class); nibble (class) = command << 1 | run
                int commands = 0, classes = 0;
                while (commands < kBC7TripleGroup && (pc[0] & 0xff) == kCmdTriple && classes + (int)((pc[0] >> 8) & 0xff) <= kBC7TripleClasses)
                {
                    const int nRuns = (pc[0] >> 8) & 0xff;
                    for (int r = 0; r < nRuns; r++)
                    {
                        const float e = as_float(res[(pc[3 + r] >> 16) & 0xff][0]);
                        // the lane condition of the partition scan (BC67.cpp:1602-1634) and room below the best
                        if (lf.warpAnyRGB && lf.allowRGBModes && !(e > work.error))
                            wantMask |= 1u << classes;
                        classMap |= (uint32_t)(commands * 2 + r) << (4 * classes++);
                    }
                    pcs[commands++] = pc;
                    pc += 3 + nRuns;
                }
                for (int k = commands; k < kBC7TripleGroup; k++)
                    pcs[k] = pcs[0];
                ex.publish((lf.allowRGBModes ? 1u : 0u) | (usePCA4 ? 2u : 0u));
                const int total = ex.compact(wantMask);
                bc7_triple_tasks<FAST, STRIDE>(P, L, ex, pcs, classMap, total);
                ex.sync();
                for (uint32_t m = wantMask; m; m &= m - 1)
                {
                    const int cls = ctz32(m);
                    const uint32_t entry = (classMap >> (4 * cls)) & 15u;
                    const int k = (int)(entry >> 1), r = (int)(entry & 1u);
                    const uint32_t c0 = pcs[k][0], rw = pcs[k][3 + r];
                    const int partition = (c0 >> 24) & 0x3f, sA = (c0 >> 16) & 3, sB = (pcs[k][1] >> 24) & 3, sC = (pcs[k][2] >> 24) & 3;
                    const int mode = rw & 0xf, slotA = (rw >> 16) & 0xff;
                    F4 e, ep;
                    ex.result2(cls, e, ep);
                    float errs[3];
                    errs[sA] = as_float(res[slotA][0]);
                    errs[sB] = e.x;
                    errs[sC] = e.y;
                    const float totalError = fadd(fadd(errs[0], errs[1]), errs[2]);       // EVAL's order: subset 0 + 1 + 2
                    const int key = bc7_mode_order(mode) * 64 + partition;
                    if (totalError < work.error || (totalError == work.error && key < work.key))
                    {
                        work.error = totalError;
                        work.key = key;
                        work.mode = mode;
                        work.sub = partition;
                        work.ep[sA][0] = res[slotA][1];
                        work.ep[sA][1] = res[slotA][2];
                        work.ep[sB][0] = as_uint(ep.x);
                        work.ep[sB][1] = as_uint(ep.y);
                        work.ep[sC][0] = as_uint(ep.z);
                        work.ep[sC][1] = as_uint(ep.w);
                        work.sc[0] = work.sc[1] = work.sc[2] = 0;
                    }
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
                    float t;
                    t = wr[3]; wr[3] = wr[c]; wr[c] = t;
                    t = wSqr[3]; wSqr[3] = wSqr[c]; wSqr[c] = t;
                    t = rcpWr[3]; rcpWr[3] = rcpWr[c]; rcpWr[c] = t;
                }
                {
                    float sumV[4], accA;
                    bc7_gather<STRIDE>(L, 0xffffu, rotation, wr, sumV, accA);
                }
                bc7_dual_plane<FAST, STRIDE>(P, L.gv, L.gw, mode, rotation, indexSelector, seeds, wr, wSqr, rcpWr, work);
            }
        }
    }

    // Index selection for the winner and the block's 128 bits
    template<bool FAST, int STRIDE>
    CVTT_HD void bc7_finish_block(const BC7Params &P, const BC7PackTables &T, const BC7Lane<STRIDE> &L, BC7Work &work, uint32_t out[4])
    {
        // indexes of the winner
        {
            const int mode = work.mode;
            uint32_t idx[2] = { 0, 0 }, idx2[2] = { 0, 0 };
            if (mode == 4 || mode == 5)
            {
                const int rotation = work.sub & 3, indexSelector = (work.sub >> 2) & 1;
                const int rgbPrec = (mode == 4 && indexSelector) ? 3 : 2;
                const int alphaPrec = (mode == 4 && !indexSelector) ? 3 : 2;
                float wr[4] = { P.w[0], P.w[1], P.w[2], P.w[3] }, wSqr[4] = { P.wSq[0], P.wSq[1], P.wSq[2], P.wSq[3] };
                if (rotation)
                {
                    const int c = rotation - 1;
                    float t;
                    t = wr[3]; wr[3] = wr[c]; wr[c] = t;
                    t = wSqr[3]; wSqr[3] = wSqr[c]; wSqr[c] = t;
                }
                const float unit[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
                BC7IndexSelector sRGB, sA;
                bc7_selector_init(sRGB, 0, 3, rgbPrec, work.ep[0][0], work.ep[0][1], wr, wSqr);
                bc7_selector_init(sA, 3, 1, alphaPrec, work.ep[0][0], work.ep[0][1], unit, wSqr);
                uint32_t rgbIdx[2] = { 0, 0 }, aIdx[2] = { 0, 0 };
                for (int px = 0; px < 16; px++)
                {
                    F4 p = bc7_expand_pixel(L.raw[px * STRIDE]);
                    bc7_rotate(p, rotation);
                    rgbIdx[px >> 3] |= bc7_select_index<FAST>(sRGB, p) << (4 * (px & 7));
                    aIdx[px >> 3] |= bc7_select_index<FAST>(sA, p) << (4 * (px & 7));
                }
                for (int h = 0; h < 2; h++)
                {
                    idx[h] = indexSelector ? aIdx[h] : rgbIdx[h];
                    idx2[h] = indexSelector ? rgbIdx[h] : aIdx[h];
                }
            }
            else
            {
                const int numSubsets = (mode == 0 || mode == 2) ? 3 : ((mode == 1 || mode == 3 || mode == 7) ? 2 : 1);
                const int nch = (mode < 4) ? 3 : 4;
                const int indexBits = (mode < 2) ? 3 : (mode == 6 ? 4 : 2);
                const int partition = work.sub;
                for (int s = 0; s < numSubsets; s++)
                {
                    uint32_t mask = 0xffffu;
                    if (numSubsets == 2)
                        mask = s ? T.partitionMask2[partition] : (uint32_t)(~T.partitionMask2[partition]) & 0xffffu;
                    else if (numSubsets == 3)
                    {
                        mask = 0;
                        for (int px = 0; px < 16; px++)
                            if ((int)((T.partitionMap3[partition] >> (px * 2)) & 3) == s)
                                mask |= 1u << px;
                    }
                    BC7IndexSelector S;
                    bc7_selector_init(S, 0, nch, indexBits, work.ep[s][0], work.ep[s][1], P.w, P.wSq);
                    for (uint32_t m = mask; m; m &= m - 1)
                    {
                        const int px = ctz32(m);
                        const F4 p = bc7_expand_pixel(L.raw[px * STRIDE]);
                        // a single-colour winner gives every pixel of the subset the table's index (BC67.cpp:1035-1036)
                        const uint32_t index = work.sc[s] ? (work.sc[s] & 15u) : bc7_select_index<FAST>(S, p);
                        idx[px >> 3] |= index << (4 * (px & 7));
                    }
                }
            }
            work.idx[0] = idx[0];
            work.idx[1] = idx[1];
            work.idx2[0] = idx2[0];
            work.idx2[1] = idx2[1];
        }

        bc7_pack_block(work, T, out);
    }

    // The whole of BC7Computer::Pack for one block (BC67.cpp:1756-1878)
    template<bool FAST, int STRIDE, bool PUNCH, class Vote, class Exchange>
    CVTT_HD void bc7_encode_block(const BC7Params &P, const BC7PackTables &T, const BC7Lane<STRIDE> &L, const BC7LaneFlags &lf, Vote &vote, Exchange &ex, uint32_t out[4])
    {
        BC7Work work;
        bc7_search_block<FAST, STRIDE, PUNCH>(P, L, lf, vote, ex, P.cmds, work);
        bc7_finish_block<FAST, STRIDE>(P, T, L, work, out);
    }
}
