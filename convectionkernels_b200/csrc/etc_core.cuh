// ETC1 / ETC2 RGB / ETC2 punch-through / EAC alpha encode search, one 4x4 block per thread (lane = block; see cvtt_common.cuh).
//
// What it reproduces (reference elasota/ConvectionKernels, file:line):
//   ETCComputer::CompressETC2Block            ConvectionKernels_ETC.cpp:1664-1887 (opaque and punch-through)
//   EncodeVirtualTModePunchthrough            :887-1262, CompressETC1PunchthroughBlockInternal :2885-3080,
//   TestHalfBlockPunchthrough                 :151-217
//   ETCComputer::EncodePlanar                 :1274-1662
//   ETCComputer::EncodeTMode / EncodeHMode    :396-647 / :649-885
//   ETCComputer::CompressETC1BlockInternal    :2624-2882, TestHalfBlock :94-149,
//   FindBestDifferentialCombination           :219-362 (the scalar sorted search, here without a sort and warp-cooperative)
//   CompressETC2AlphaBlockInternal            :1902-2085, QuantizeETC2Alpha :2366-2411 (8-bit alpha and EAC R11)
//   EmitTModeBlock / EmitHModeBlock / EmitETC1Block  :2414-2622
//   tables                                    ConvectionKernels_ETC1.h, ConvectionKernels_ETC2.h, ConvectionKernels_ETC2_Rounding.h
//
//   Flags::ETC_UseFakeBT709 / ETC_FakeBT709Accurate  ConvertToFakeBT709 / ResolveHalfBlockFakeBT709Rounding* / ResolveTHFakeBT709Rounding
//                                             (etc_to_bt709, etc_resolve_half_block_bt709, etc_resolve_th_bt709)
// Every format and flag of the reference's ETC entry points is implemented.
//
// Cross-lane semantics: every AnySet guard of these functions is idempotent except one.  In T
// mode the candidate list of a lane with fewer unique line colours than the largest count in its group of 8 has a
// never-written slot (ETC.cpp:570-577: the fill loop starts at numUnique + 1), which the zero-initialised
// reference build evaluates as colour 0.  The kernel therefore needs the group maximum of the unique-colour count
// (Vote::max) and evaluates that extra candidate.
//
// Scratch (the reference's DifferentialResolveStorage and HModeEval, ETC.h:36-55, "AllocETC2Data") lives in global
// memory, laid out [entry][thread] so that a warp's accesses to one entry are contiguous.
#pragma once

#include "cvtt_common.cuh"

namespace cvttb200
{
#ifndef CVTT_F4_DEFINED
#define CVTT_F4_DEFINED
    struct alignas(16) F4 { float x, y, z, w; };
#endif

    enum { kETCMaxAttempts = 624, kETCHColors = 62 };

    struct ETCParams
    {
        float w[3];                 // options.redWeight / greenWeight / blueWeight (the ETC path does not use FillWeights)
        float wSq[3];               // w * w, the per-channel factors of EncodePlanar (ETC.cpp:1561-1577)
        float chromaAxis0[3], chromaAxis1[3];       // ETC2CompressionDataInternal ctor, ETC.cpp:3117-3145
        // scalar constants of EncodePlanar's 3x3 solve (ETC.cpp:1294-1386), identical for every block and channel
        float pl_r0to1, pl_r0to2, pl_r1to2, pl_n2, pl_r2to1, pl_elim2, pl_elim1, pl_d, pl_k1;
        uint32_t flags;
        int punchThreshold;         // CompressETC2Block, ETC.cpp:1672-1675: pixels with alpha below this are transparent
    };

    struct ETCTables
    {
        int16_t potentialOffsets[8 * 82];           // per table: count, then the offsets (Tables::ETC1::g_potentialOffsets4)
        int16_t thModifier[8];                      // g_thModifierTable
        int16_t alphaModifier[16][4];               // Tables::ETC2::g_alphaModifierTablePositive
        uint8_t alphaRounding[16][16];              // Tables::ETC2::g_alphaRoundingTables (13 wide)
        int16_t etc1Modifiers[8][4];                // modifierTables of CompressETC1BlockInternal, ETC.cpp:2665-2675
    };

    template<int STRIDE>
    struct ETCLane
    {
        const F4 *pw;               // pw[px * STRIDE]: .xyz = pre-weighted pixel (ExtractBlocks, ETC.cpp:2128-2155), .w = bits r | g << 8 | b << 16 | a << 24
    };

    // per-thread scratch; element i of an array is at [i * stride]
    struct ETCScratch
    {
        float *drsErr;              // [2 * kETCMaxAttempts]   DifferentialResolveStorage::diffErrors
        uint32_t *drsMeta;          // [2 * kETCMaxAttempts]   packed colour | table << 15
        float *hErr;                // [kETCHColors * 16]      HModeEval::errors
        uint32_t *hMeta;            // [kETCHColors]           signBits | uniqueQuantizedColor << 16
        size_t stride;
    };

    struct ETCBest
    {
        float error;
        uint32_t hi, lo;            // the 8 output bytes as two big-endian words
    };

    CVTT_HD int etc_px(const F4 &p, int ch) { return (int)((as_uint(p.w) >> (8 * ch)) & 0xffu); }

#ifndef CVTT_SC_QUALIFIER
#if defined(__CUDACC__)
#define CVTT_SC_QUALIFIER static __device__ const
#else
#define CVTT_SC_QUALIFIER static const
#endif
#endif
#if defined(__CUDACC__) || defined(CVTT_HOSTSIM)
#include "etc_bt709_table.inc"
#define CVTT_HAVE_BT709_TABLE 1
#endif

    // ConvertToFakeBT709, ETC.cpp:2344-2353
    CVTT_HD void etc_to_bt709(float r, float g, float b, float *yuv)
    {
        yuv[0] = fadd(fadd(fmul(r, 0.368233989135369f), fmul(g, 1.23876274963149f)), fmul(b, 0.125054068802017f));
        yuv[1] = fsub(fsub(fmul(r, 0.5f), fmul(g, 0.4541529f)), fmul(b, 0.04584709f));
        yuv[2] = fadd(fsub(fmul(r, -0.081014709086133f), fmul(g, 0.272538676238785f)), fmul(b, 0.353553390593274f));
    }

    // ComputeErrorUniform (ETC.cpp:59-71) / ComputeErrorWeighted (:73-80) / ComputeErrorFakeBT709 (:82-92).
    // cw = what etc_weigh made of the candidate colour: (float)colour * weight, or its fake-BT.709 YUV.
    template<bool UNIFORM, bool BT709>
    CVTT_HD float etc_error(const F4 &p, const int *c, const float *cw)
    {
        if (BT709)
        {
            const float dy = fsub(cw[0], p.x), du = fsub(cw[1], p.y), dv = fsub(cw[2], p.z);
            return fadd(fadd(fmul(dy, dy), fmul(du, du)), fmul(dv, dv));
        }
        else if (UNIFORM)
        {
            const float d0 = (float)(etc_px(p, 0) - c[0]), d1 = (float)(etc_px(p, 1) - c[1]), d2 = (float)(etc_px(p, 2) - c[2]);
            return fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2));
        }
        else
        {
            const float dr = fsub(cw[0], p.x), dg = fsub(cw[1], p.y), db = fsub(cw[2], p.z);
            return fadd(fadd(fmul(dr, dr), fmul(dg, dg)), fmul(db, db));
        }
    }

    // etc_error (weighted form) for two colours at once in packed fp32: lane x measures cwA, lane y measures cwB.  Each lane
    // performs exactly the scalar sequence (cw - p == cw + (-p) in IEEE arithmetic).
    CVTT_HD f2 etc_error_pair(const F4 &p, const f2 *cw2)
    {
        const f2 dr = f2_sub(cw2[0], p.x), dg = f2_sub(cw2[1], p.y), db = f2_sub(cw2[2], p.z);
        return f2_add(f2_add(f2_mul(dr, dr), f2_mul(dg, dg)), f2_mul(db, db));
    }

    template<bool UNIFORM, bool BT709>
    CVTT_HD void etc_weigh(const ETCParams &P, const int *c, float *cw)
    {
        if (BT709)
            etc_to_bt709((float)c[0], (float)c[1], (float)c[2], cw);
        else if (!UNIFORM)
            for (int ch = 0; ch < 3; ch++)
                cw[ch] = fmul((float)c[ch], P.w[ch]);
    }

    // ResolveTHFakeBT709Rounding (ETC.cpp:2301-2342) and the shared tail of ResolveHalfBlockFakeBT709RoundingAccurate
    // (:2197-2237): pick the octant of (low, high) unquantised values whose YUV is closest to the target's.  The reference's
    // error expression adds delta[1] twice instead of squaring it; reproduced.
    CVTT_HD int etc_bt709_best_octant(const float *low, const float *high, const float *targetYUV)
    {
        float bestError = FLT_MAX;
        int bestOctant = 0;
        for (int octant = 0; octant < 8; octant++)
        {
            float yuv[3];
            etc_to_bt709((octant & 1) ? high[0] : low[0], (octant & 2) ? high[1] : low[1], (octant & 4) ? high[2] : low[2], yuv);
            const float d0 = fsub(yuv[0], targetYUV[0]), d1 = fsub(yuv[1], targetYUV[1]), d2 = fsub(yuv[2], targetYUV[2]);
            const float error = fadd(fadd(fadd(fmul(d0, d0), d1), d1), fmul(d2, d2));
            if (error < bestError)
                bestOctant = octant;
            bestError = sse_min(error, bestError);
        }
        return bestOctant;
    }

    CVTT_HD void etc_resolve_th_bt709(int *quantized, const int *targets, int granularity)
    {
        float low[3], high[3], targetYUV[3];
        for (int ch = 0; ch < 3; ch++)
        {
            const int unq = (quantized[ch] << 4) | quantized[ch];
            const int unqNext = imin(255, unq + 17);
            low[ch] = (float)wrap_u16(wrap_u16(unq * granularity) << 1);
            high[ch] = (float)wrap_u16(wrap_u16(unqNext * granularity) << 1);
        }
        etc_to_bt709((float)targets[0], (float)targets[1], (float)targets[2], targetYUV);
        const int octant = etc_bt709_best_octant(low, high, targetYUV);
        for (int ch = 0; ch < 3; ch++)
            quantized[ch] += (octant >> ch) & 1;
    }

    // ResolveHalfBlockFakeBT709RoundingAccurate / Fast (ETC.cpp:2157-2299): cu = the clamped cumulative colour of the half block
    CVTT_HD void etc_resolve_half_block_bt709(int *quantized, const int *cu, bool differential, bool accurate)
    {
        if (accurate)
        {
            float low[3], high[3], targetYUV[3];
            for (int ch = 0; ch < 3; ch++)
            {
                const int c = cu[ch];
                quantized[ch] = differential ? (wrap_u16((c << 5) - c + (c >> 3)) >> 11) : (wrap_u16((c << 5) - (c << 1) + (c >> 3)) >> 12);
                int unq, unqNext;
                if (differential)
                {
                    unq = (quantized[ch] << 3) | (quantized[ch] >> 2);
                    const int qn = imin(31, quantized[ch] + 1);
                    unqNext = (qn << 3) | (qn >> 2);
                }
                else
                {
                    unq = (quantized[ch] << 4) | quantized[ch];
                    unqNext = imin(255, unq + 17);
                }
                low[ch] = (float)(unq << 3);
                high[ch] = (float)(unqNext << 3);
            }
            etc_to_bt709((float)cu[0], (float)cu[1], (float)cu[2], targetYUV);
            const int octant = etc_bt709_best_octant(low, high, targetYUV);
            for (int ch = 0; ch < 3; ch++)
                quantized[ch] += (octant >> ch) & 1;
        }
        else
        {
#if defined(CVTT_HAVE_BT709_TABLE) && (defined(__CUDA_ARCH__) || defined(CVTT_HOSTSIM))
            int fill[3];
            for (int ch = 0; ch < 3; ch++)
                fill[ch] = cu[ch] + (cu[ch] >> 8);
            int lookup, base[3], upper;
            if (differential)
            {
                lookup = ((fill[0] << 6) & 0xf00) | ((fill[1] << 4) & 0x0f0) | ((fill[2] >> 2) & 0x00f);
                for (int ch = 0; ch < 3; ch++)
                    base[ch] = fill[ch] >> 6;
                upper = 31;
            }
            else
            {
                lookup = ((fill[0] << 5) & 0xf00) | ((fill[1] << 1) & 0x0f0) | ((fill[2] >> 3) & 0x00f);
                for (int ch = 0; ch < 3; ch++)
                    base[ch] = fill[ch] >> 7;
                upper = 15;
            }
            const int octant = kETCFakeBT709Rounding16[lookup];
            for (int ch = 0; ch < 3; ch++)
                quantized[ch] = imin(base[ch] + ((octant >> ch) & 1), upper);
#endif
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // emitters (ETC.cpp:2414-2622)
    CVTT_HD void etc_emit_t(ETCBest &best, const int *lineColor, const int *isolatedColor, uint32_t packedSelectors, int table, bool opaque)
    {
        uint32_t lowBits = 0, highBits = 0;
        const int rh = (isolatedColor[0] >> 2) & 3, rl = isolatedColor[0] & 3;
        if (rh + rl < 4)
            highBits |= 1u << (58 - 32);
        else
            highBits |= 7u << (61 - 32);
        highBits |= (uint32_t)rh << (59 - 32);
        highBits |= (uint32_t)rl << (56 - 32);
        highBits |= (uint32_t)isolatedColor[1] << (52 - 32);
        highBits |= (uint32_t)isolatedColor[2] << (48 - 32);
        highBits |= (uint32_t)lineColor[0] << (44 - 32);
        highBits |= (uint32_t)lineColor[1] << (40 - 32);
        highBits |= (uint32_t)lineColor[2] << (36 - 32);
        highBits |= (uint32_t)((table >> 1) & 3) << (34 - 32);
        if (opaque)
            highBits |= 1u << (33 - 32);
        highBits |= (uint32_t)(table & 1);
        for (int px = 0; px < 16; px++)
        {
            const int order = (px & 3) * 4 + (px >> 2);       // selectorOrder
            const uint32_t sel = (packedSelectors >> (2 * order)) & 3u;
            if (sel & 1u)
                lowBits |= 1u << px;
            if (sel & 2u)
                lowBits |= 1u << (16 + px);
        }
        best.hi = highBits;
        best.lo = lowBits;
    }

    CVTT_HD void etc_emit_h(ETCBest &best, const int *blockColors, uint32_t sectorBits, uint32_t signBits, int table, bool opaque)
    {
        if (blockColors[0] == blockColors[1])
        {
            int lineColor[3], isolatedColor[3];
            lineColor[0] = isolatedColor[0] = (blockColors[0] >> 10) & 0x1f;
            lineColor[1] = isolatedColor[1] = (blockColors[0] >> 5) & 0x1f;
            lineColor[2] = isolatedColor[2] = blockColors[0] & 0x1f;
            uint32_t packedSelectors = 0x55555555u;
            for (int px = 0; px < 16; px++)
                packedSelectors |= ((signBits >> px) & 1u) << (px * 2 + 1);
            etc_emit_t(best, lineColor, isolatedColor, packedSelectors, table, opaque);
            return;
        }
        int colors[2][3];
        for (int sector = 0; sector < 2; sector++)
            for (int ch = 0; ch < 3; ch++)
                colors[sector][ch] = (blockColors[sector] >> ((2 - ch) * 5)) & 15;
        uint32_t lowBits = 0, highBits = 0;
        if (((table & 1) == 1) != (blockColors[0] > blockColors[1]))
        {
            for (int ch = 0; ch < 3; ch++)
            {
                const int t = colors[0][ch];
                colors[0][ch] = colors[1][ch];
                colors[1][ch] = t;
            }
            sectorBits ^= 0xffffu;
        }
        const int r1 = colors[0][0], g1a = colors[0][1] >> 1, g1b = colors[0][1] & 1, b1a = colors[0][2] >> 3, b1b = colors[0][2] & 7;
        const int r2 = colors[1][0], g2 = colors[1][1], b2 = colors[1][2];
        if ((g1a & 4) != 0 && r1 + g1a < 8)
            highBits |= 1u << (63 - 32);
        const int fakeDG = b1b >> 1, fakeG = b1a | (g1b << 1);
        if (fakeG + fakeDG < 4)
            highBits |= 1u << (50 - 32);
        else
            highBits |= 7u << (53 - 32);
        const int da = (table >> 2) & 1, db = (table >> 1) & 1;
        highBits |= (uint32_t)r1 << (59 - 32);
        highBits |= (uint32_t)g1a << (56 - 32);
        highBits |= (uint32_t)g1b << (52 - 32);
        highBits |= (uint32_t)b1a << (51 - 32);
        highBits |= (uint32_t)b1b << (47 - 32);
        highBits |= (uint32_t)r2 << (43 - 32);
        highBits |= (uint32_t)g2 << (39 - 32);
        highBits |= (uint32_t)b2 << (35 - 32);
        highBits |= (uint32_t)da << (34 - 32);
        if (opaque)
            highBits |= 1u << (33 - 32);
        highBits |= (uint32_t)db;
        for (int px = 0; px < 16; px++)
        {
            const int order = (px & 3) * 4 + (px >> 2);
            lowBits |= ((signBits >> order) & 1u) << px;
            lowBits |= ((sectorBits >> order) & 1u) << (16 + px);
        }
        best.hi = highBits;
        best.lo = lowBits;
    }

    // block pixel of (flip, sector, i): g_flipTables, ETC.cpp:47-57
    CVTT_HD int etc_flip_pixel(int flip, int sector, int i)
    {
        return flip ? (sector * 8 + i) : ((i >> 1) * 4 + sector * 2 + (i & 1));
    }

    CVTT_HD void etc_emit_etc1(ETCBest &best, int flip, int d, const int colors[2][3], const int *tables, const uint32_t *selectors, bool transparent = false)
    {
        uint32_t highBits = 0, lowBits = 0;
        if (d == 0)
        {
            highBits |= (uint32_t)colors[0][0] << 28;
            highBits |= (uint32_t)colors[1][0] << 24;
            highBits |= (uint32_t)colors[0][1] << 20;
            highBits |= (uint32_t)colors[1][1] << 16;
            highBits |= (uint32_t)colors[0][2] << 12;
            highBits |= (uint32_t)colors[1][2] << 8;
        }
        else
        {
            highBits |= (uint32_t)colors[0][0] << 27;
            highBits |= (uint32_t)((colors[1][0] - colors[0][0]) & 7) << 24;
            highBits |= (uint32_t)colors[0][1] << 19;
            highBits |= (uint32_t)((colors[1][1] - colors[0][1]) & 7) << 16;
            highBits |= (uint32_t)colors[0][2] << 11;
            highBits |= (uint32_t)((colors[1][2] - colors[0][2]) & 7) << 8;
        }
        highBits |= (uint32_t)tables[0] << 5;
        highBits |= (uint32_t)tables[1] << 2;
        if (!transparent)
            highBits |= (uint32_t)d << 1;
        highBits |= (uint32_t)flip;
        uint32_t codes = 0;     // 2 bits per block pixel
        for (int sector = 0; sector < 2; sector++)
            for (int px = 0; px < 8; px++)
            {
                const uint32_t selector = (selectors[sector] >> (2 * px)) & 3u;
                const uint32_t code = (0x4B >> (2 * selector)) & 3u;          // modifierCodes { 3, 2, 0, 1 }
                codes |= code << (2 * etc_flip_pixel(flip, sector, px));
            }
        for (int sb = 0; sb < 2; sb++)
            for (int px = 0; px < 16; px++)
            {
                const int order = (px & 3) * 4 + (px >> 2);
                lowBits |= ((codes >> (2 * order + sb)) & 1u) << (px + sb * 16);
            }
        best.hi = highBits;
        best.lo = lowBits;
    }

    // ---------------------------------------------------------------------------------------------------------
    // EncodePlanar, ETC.cpp:1274-1662 (RGB path)
    template<bool UNIFORM, bool BT709, int STRIDE>
    CVTT_HD void etc_planar(const ETCParams &P, const ETCLane<STRIDE> &L, ETCBest &best)
    {
        float totalError = 0.0f;
        int bestCoeffs[3][3];
        float oAll[3], hAll[3], vAll[3];
        for (int ch = 0; ch < 3; ch++)
        {
            float fc = 0.0f, fh = 0.0f, fv = 0.0f, fo = 0.0f;
            for (int px = 0; px < 16; px++)
            {
                const float x = (float)(px % 4), y = (float)(px / 4);
                const F4 pp = L.pw[px * STRIDE];
                const float c = BT709 ? (ch == 0 ? pp.x : (ch == 1 ? pp.y : pp.z)) : (float)etc_px(pp, ch);
                fh = fsub(fh, fmul(c, x));
                fv = fsub(fv, fmul(c, y));
                fo = fsub(fo, c);
                fh = fsub(fh, fmul(c, x));
                fv = fsub(fv, fmul(c, y));
                fo = fsub(fo, c);
                fc = fadd(fc, fmul(c, c));
            }
            const float gD = fh, lD = fv, qD = fo;
            const float l1D = fadd(lD, fmul(gD, P.pl_r0to1));
            const float q1D = fadd(qD, fmul(gD, P.pl_r0to2));
            const float q2D = fadd(q1D, fmul(l1D, P.pl_r1to2));
            const float o = fdiv(fsub(0.0f, q2D), P.pl_n2);
            const float l2D = fadd(l1D, fmul(q2D, P.pl_r2to1));
            const float g2D = fadd(fadd(gD, fmul(l2D, P.pl_elim2)), fmul(q2D, P.pl_elim1));
            float h = fdiv(fsub(0.0f, g2D), P.pl_d);
            float v = fdiv(fsub(0.0f, l2D), P.pl_k1);
            oAll[ch] = o;
            hAll[ch] = fadd(fmul(h, 4.0f), o);
            vAll[ch] = fadd(fmul(v, 4.0f), o);
        }

        if (BT709)
        {
            // the fit was done in fake-BT.709 YUV; back to RGB, round to nearest, measure in YUV (ETC.cpp:1395-1455)
            float rgbO[3], rgbH[3], rgbV[3];
            const float *src[3] = { oAll, hAll, vAll };
            float *dstp[3] = { rgbO, rgbH, rgbV };
            for (int k = 0; k < 3; k++)
            {
                // ConvertFromFakeBT709, ETC.cpp:2355-2364
                const float yy = fmul(src[k][0], 0.57735026466774571071f), u = src[k][1], vv = src[k][2];
                dstp[k][0] = fadd(yy, fmul(u, 1.5748000207960953486f));
                dstp[k][1] = fsub(fsub(yy, fmul(u, 0.46812425854364753669f)), fmul(vv, 0.26491652528157560861f));
                dstp[k][2] = fadd(yy, fmul(vv, 2.6242146882856944069f));
            }
            int rec[16][3];
            for (int ch = 0; ch < 3; ch++)
            {
                const float fcoeffsIn[3] = { rgbO[ch], rgbH[ch], rgbV[ch] };
                for (int c = 0; c < 3; c++)
                {
                    float coeff = sse_max(0.0f, fcoeffsIn[c]);
                    if (ch == 1)
                        coeff = sse_min(127.0f, fmul(coeff, 127.0f / 255.0f));
                    else
                        coeff = sse_min(63.0f, fmul(coeff, 63.0f / 255.0f));
                    bestCoeffs[ch][c] = (int)rne(coeff);
                }
                const int cO = bestCoeffs[ch][0], cH = bestCoeffs[ch][1], cV = bestCoeffs[ch][2];
                const int dO = (ch == 1) ? ((cO << 1) | (cO >> 6)) : ((cO << 2) | (cO >> 4));
                const int dH = (ch == 1) ? ((cH << 1) | (cH >> 6)) : ((cH << 2) | (cH >> 4));
                const int dV = (ch == 1) ? ((cV << 1) | (cV >> 6)) : ((cV << 2) | (cV >> 4));
                const int hMinusO = dH - dO, vMinusO = dV - dO, addend = (dO << 2) + 2;
                for (int px = 0; px < 16; px++)
                    rec[px][ch] = imin(255, imax(0, ((px & 3) * hMinusO + (px >> 2) * vMinusO + addend) >> 2));
            }
            totalError = 0.0f;
            for (int px = 0; px < 16; px++)
            {
                float yuv[3];
                etc_weigh<UNIFORM, true>(P, rec[px], yuv);
                totalError = fadd(totalError, etc_error<UNIFORM, true>(L.pw[px * STRIDE], rec[px], yuv));
            }
        }
        else
        for (int ch = 0; ch < 3; ch++)
        {
            const float o = oAll[ch], h = hAll[ch], v = vAll[ch];

            const float fcoeffsIn[3] = { o, h, v };
            int ranges[3][2];
            for (int c = 0; c < 3; c++)
            {
                float coeff = sse_max(0.0f, fcoeffsIn[c]);
                if (ch == 1)
                    coeff = sse_min(127.0f, fmul(coeff, 127.0f / 255.0f));
                else
                    coeff = sse_min(63.0f, fmul(coeff, 63.0f / 255.0f));
                // RoundAndConvertToU15 under round-down / round-up (the value is in 0..127)
                ranges[c][0] = (int)floorf(coeff);
                ranges[c][1] = (int)ceilf(coeff);
            }

            float bestChannelError = FLT_MAX;
            for (int io = 0; io < 2; io++)
                for (int ih = 0; ih < 2; ih++)
                    for (int iv = 0; iv < 2; iv++)
                    {
                        const int cO = ranges[0][io], cH = ranges[1][ih], cV = ranges[2][iv];
                        // DecodePlanarCoeff, ETC.cpp:1264-1270
                        const int dO = (ch == 1) ? ((cO << 1) | (cO >> 6)) : ((cO << 2) | (cO >> 4));
                        const int dH = (ch == 1) ? ((cH << 1) | (cH >> 6)) : ((cH << 2) | (cH >> 4));
                        const int dV = (ch == 1) ? ((cV << 1) | (cV >> 6)) : ((cV << 2) | (cV >> 4));
                        const int hMinusO = dH - dO, vMinusO = dV - dO, addend = (dO << 2) + 2;
                        float error = 0.0f;
                        for (int px = 0; px < 16; px++)
                        {
                            const int interpolated = ((px & 3) * hMinusO + (px >> 2) * vMinusO + addend) >> 2;
                            const int dec = imin(255, imax(0, interpolated));
                            const float deltaF = (float)(etc_px(L.pw[px * STRIDE], ch) - dec);
                            error = fadd(error, fmul(deltaF, deltaF));
                        }
                        if (error < bestChannelError)
                        {
                            bestChannelError = error;
                            bestCoeffs[ch][0] = cO;
                            bestCoeffs[ch][1] = cH;
                            bestCoeffs[ch][2] = cV;
                        }
                    }
            if (!UNIFORM)
                bestChannelError = fmul(bestChannelError, P.wSq[ch]);
            totalError = fadd(totalError, bestChannelError);
        }

        if (totalError < best.error)
        {
            best.error = totalError;
            const int ro = bestCoeffs[0][0], rh = bestCoeffs[0][1], rv = bestCoeffs[0][2];
            const int go = bestCoeffs[1][0], gh = bestCoeffs[1][1], gv = bestCoeffs[1][2];
            const int bo = bestCoeffs[2][0], bh = bestCoeffs[2][1], bv = bestCoeffs[2][2];
            const int go1 = go >> 6, go2 = go & 63, bo1 = bo >> 5, bo2 = (bo >> 3) & 3, bo3 = bo & 7, rh1 = rh >> 1, rh2 = rh & 1;
            const int fakeR = ro >> 2, fakeDR = go1 | ((ro & 3) << 1);
            const int fakeG = go2 >> 2, fakeDG = ((go2 & 3) << 1) | bo1;
            const int fakeB = bo2, fakeDB = bo3 >> 1;
            uint32_t highBits = 0, lowBits = 0;
            if ((fakeDR & 4) != 0 && fakeR + fakeDR < 8)
                highBits |= 1u << (63 - 32);
            if ((fakeDG & 4) != 0 && fakeG + fakeDG < 8)
                highBits |= 1u << (55 - 32);
            if (fakeB + fakeDB < 4)
                highBits |= 1u << (42 - 32);
            else
                highBits |= 7u << (45 - 32);
            highBits |= (uint32_t)ro << (57 - 32);
            highBits |= (uint32_t)go1 << (56 - 32);
            highBits |= (uint32_t)go2 << (49 - 32);
            highBits |= (uint32_t)bo1 << (48 - 32);
            highBits |= (uint32_t)bo2 << (43 - 32);
            highBits |= (uint32_t)bo3 << (39 - 32);
            highBits |= (uint32_t)rh1 << (34 - 32);
            highBits |= 1u << (33 - 32);
            highBits |= (uint32_t)rh2;
            lowBits |= (uint32_t)gh << 25;
            lowBits |= (uint32_t)bh << 19;
            lowBits |= (uint32_t)rv << 13;
            lowBits |= (uint32_t)gv << 6;
            lowBits |= (uint32_t)bv;
            best.hi = highBits;
            best.lo = lowBits;
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // EncodeTMode, ETC.cpp:396-647.  isolatedMask: bit px set = pixel is in the isolated cluster.
    // rendezvous: whether the CTA's warps meet before the table loop (see etc2_encode_block)
    template<bool UNIFORM, bool BT709, int STRIDE, class Vote>
    CVTT_HD void etc_t_mode(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, Vote &vote, uint32_t isolatedMask, ETCBest &best, bool rendezvous = true)
    {
        int isolatedTotal[3] = { 0, 0, 0 }, lineTotal[3] = { 0, 0, 0 }, numIsolated = 0;
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.pw[px * STRIDE];
            const bool iso = (isolatedMask >> px) & 1;
            for (int ch = 0; ch < 3; ch++)
            {
                const int v = etc_px(p, ch);
                lineTotal[ch] += v;
                if (iso)
                    isolatedTotal[ch] += v;
            }
            numIsolated += iso ? 1 : 0;
        }
        for (int ch = 0; ch < 3; ch++)
            lineTotal[ch] -= isolatedTotal[ch];
        const int numLine = 16 - numIsolated;
        const int lineDivisorV = numLine * 34, lineAddendV = (numLine << 4) | numLine;
        const UDivisor lineDivide = udiv_prepare((uint32_t)lineDivisorV);     // up to 33 x 3 x 8 divisions by it below

        int isolatedQ[3], isolatedColor[3];
        {
            const int divisor = numIsolated * 34, addend = (numIsolated << 4) | numIsolated;
            int targets[3];
            for (int ch = 0; ch < 3; ch++)
            {
                const int numerator = isolatedTotal[ch] + isolatedTotal[ch] + (BT709 ? 0 : addend);
                isolatedQ[ch] = (divisor == 0) ? 0 : (numerator / divisor);
                targets[ch] = numerator;
            }
            if (BT709)
                etc_resolve_th_bt709(isolatedQ, targets, numIsolated);
            for (int ch = 0; ch < 3; ch++)
                isolatedColor[ch] = wrap_u16(isolatedQ[ch] | (isolatedQ[ch] << 4));
        }
        float isoW[3];
        etc_weigh<UNIFORM, BT709>(P, isolatedColor, isoW);

        // packed line colour for one offset step (ETC.cpp:507-548)
        auto lineColorAt = [&](int offs, int modifierOffset) -> int
        {
            int q[3], targets[3];
            for (int ch = 0; ch < 3; ch++)
            {
                const int numerator = imax(0, wrap_s16(lineTotal[ch] + lineTotal[ch] + (BT709 ? 0 : lineAddendV) + offs * modifierOffset));
                const int divided = (lineDivisorV == 0) ? 0 : (int)udiv((uint32_t)numerator, lineDivide);
                q[ch] = imin(15, divided);
                targets[ch] = numerator;
            }
            if (BT709)
                etc_resolve_th_bt709(q, targets, numLine);
            return q[0] | (q[1] << 5) | (q[2] << 10);
        };

        bool bestIsThisMode = false;
        uint32_t bestSelectors = 0;
        int bestTable = 0, bestLineColor = 0;

        if (rendezvous)
            cta_sync();     // keeps the warps of the CTA in the same code region (instruction cache), see DESIGN.md
        for (int table = 0; table < 8; table++)
        {
            const int modifier = T.thModifier[table];
            const int modifierOffset = modifier + modifier;

            // unique line colours of this lane, in order (at most 2 * numLine + 1 <= 33, the reference keeps 31)
            int numUnique = 0, lastColor = -1;
            // first pass counts, second pass evaluates: the group maximum of the count is needed before the extra candidate
            for (int offs = -numLine; offs <= numLine; offs++)
            {
                const int packed = lineColorAt(offs, modifierOffset);
                if (numUnique == 0 || packed != lastColor)
                {
                    numUnique++;
                    lastColor = packed;
                }
            }
            const int maxUnique = vote.max(numUnique);
            const int numCandidates = numUnique + ((numUnique < maxUnique) ? 1 : 0);

            int offs = -numLine;
            lastColor = -1;
            for (int ci = 0; ci < numCandidates; ci++)
            {
                int packedColor = 0;        // the never-written slot of the reference's candidate array reads as colour 0
                if (ci < numUnique)
                {
                    for (;;)
                    {
                        const int packed = lineColorAt(offs, modifierOffset);
                        offs++;
                        if (packed != lastColor)
                        {
                            lastColor = packed;
                            packedColor = packed;
                            break;
                        }
                    }
                }

                int lineColors[3][3];
                float lineW[3][3];
                for (int ch = 0; ch < 3; ch++)
                {
                    const int q = (packedColor >> (ch * 5)) & 15;
                    const int unq = (q << 4) | q;
                    lineColors[0][ch] = imin(255, unq + modifier);
                    lineColors[1][ch] = unq;
                    lineColors[2][ch] = imax(0, unq - modifier);
                }
                for (int i = 0; i < 3; i++)
                    etc_weigh<UNIFORM, false>(P, lineColors[i], lineW[i]);      // the line colours are never measured in YUV (ETC.cpp:603)

                uint32_t selectors = 0;
                float error = 0.0f;
                const f2 pairA[3] = { f2_make(isoW[0], lineW[0][0]), f2_make(isoW[1], lineW[0][1]), f2_make(isoW[2], lineW[0][2]) };
                const f2 pairB[3] = { f2_make(lineW[1][0], lineW[2][0]), f2_make(lineW[1][1], lineW[2][1]), f2_make(lineW[1][2], lineW[2][2]) };
                for (int px = 0; px < 16; px++)
                {
                    const F4 p = L.pw[px * STRIDE];
                    float e4[4];
                    if (!UNIFORM && !BT709)
                    {
                        const f2 a = etc_error_pair(p, pairA), b = etc_error_pair(p, pairB);    // isolated | line 0, line 1 | line 2
                        e4[0] = a.x;
                        e4[1] = a.y;
                        e4[2] = b.x;
                        e4[3] = b.y;
                    }
                    else
                    {
                        e4[0] = etc_error<UNIFORM, BT709>(p, isolatedColor, isoW);
#pragma unroll
                        for (int i = 0; i < 3; i++)
                            e4[i + 1] = etc_error<UNIFORM, false>(p, lineColors[i], lineW[i]);
                    }
                    float pixelError = e4[0];
                    uint32_t pixelBestSelector = 0;
#pragma unroll
                    for (int i = 0; i < 3; i++)
                    {
                        const float e = e4[i + 1];
                        if (e < pixelError)
                            pixelBestSelector = (uint32_t)(i + 1);
                        pixelError = sse_min(e, pixelError);
                    }
                    error = fadd(error, pixelError);
                    selectors |= pixelBestSelector << (px * 2);
                }
                if (error < best.error)
                {
                    best.error = error;
                    bestLineColor = packedColor;
                    bestSelectors = selectors;
                    bestTable = table;
                    bestIsThisMode = true;
                }
            }
        }

        if (bestIsThisMode)
        {
            int lineColor[3];
            for (int ch = 0; ch < 3; ch++)
                lineColor[ch] = (bestLineColor >> (ch * 5)) & 15;
            etc_emit_t(best, lineColor, isolatedQ, bestSelectors, bestTable, true);
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // EncodeHMode, ETC.cpp:649-885.  groupMask: bit px set = pixel belongs to sector 1.
    template<bool UNIFORM, bool BT709, int STRIDE>
    CVTT_HD void etc_h_mode(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, const ETCScratch &S, uint32_t groupMask, ETCBest &best)
    {
        int totals[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } }, counts[2] = { 0, 0 };
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.pw[px * STRIDE];
            const bool g = (groupMask >> px) & 1;
            for (int ch = 0; ch < 3; ch++)
            {
                const int v = etc_px(p, ch);
                totals[0][ch] += v;
                if (g)
                    totals[1][ch] += v;
            }
            counts[1] += g ? 1 : 0;
        }
        for (int ch = 0; ch < 3; ch++)
            totals[0][ch] -= totals[1][ch];
        counts[0] = 16 - counts[1];

        bool bestIsThisMode = false;
        uint32_t bestSectorBits = 0, bestSignBits = 0;
        int bestColors[2] = { 0, 0 }, bestTable = 0;

        cta_sync();
        for (int table = 0; table < 8; table++)
        {
            const int modifier = T.thModifier[table];
            const int modifierOffset = modifier * 2;
            int numUnique[2] = { 0, 0 };

            // unique colours per sector, each evaluated once against all 16 pixels (errors of the better of +-modifier)
            int total = 0;
            float floorErr[16];
#pragma unroll
            for (int px = 0; px < 16; px++)
                floorErr[px] = FLT_MAX;
            for (int sector = 0; sector < 2; sector++)
            {
                const int count = counts[sector];
                const UDivisor countDivide = udiv_prepare((uint32_t)(count * 34));
                int lastColor = -1;
                for (int offs = -count; offs <= count; offs++)
                {
                    int packed = 0;
                    for (int ch = 0; ch < 3; ch++)
                    {
                        int q = 0;
                        if (count != 0)
                            q = imin(15, (int)udiv((uint32_t)imax(0, wrap_s16(totals[sector][ch] * 2 + count * 17 + modifierOffset * offs)), countDivide));
                        packed |= q << ((2 - ch) * 5);
                    }
                    if (numUnique[sector] != 0 && packed == lastColor)
                        continue;
                    lastColor = packed;
                    numUnique[sector]++;

                    int colors[2][3];
                    float cw[2][3];
                    for (int ch = 0; ch < 3; ch++)
                    {
                        const int q = (packed >> ((2 - ch) * 5)) & 15;
                        const int unq = (q << 4) | q;
                        colors[0][ch] = imin(255, unq + modifier);
                        colors[1][ch] = imax(0, unq - modifier);
                    }
                    etc_weigh<UNIFORM, BT709>(P, colors[0], cw[0]);
                    etc_weigh<UNIFORM, BT709>(P, colors[1], cw[1]);
                    uint32_t signBits = 0;
                    const f2 cw2[3] = { f2_make(cw[0][0], cw[1][0]), f2_make(cw[0][1], cw[1][1]), f2_make(cw[0][2], cw[1][2]) };
#pragma unroll
                    for (int px = 0; px < 16; px++)
                    {
                        const F4 p = L.pw[px * STRIDE];
                        float e0, e1;
                        if (!UNIFORM && !BT709)
                        {
                            const f2 e = etc_error_pair(p, cw2);        // +modifier and -modifier colours side by side
                            e0 = e.x;
                            e1 = e.y;
                        }
                        else
                        {
                            e0 = etc_error<UNIFORM, BT709>(p, colors[0], cw[0]);
                            e1 = etc_error<UNIFORM, BT709>(p, colors[1], cw[1]);
                        }
                        if (e1 < e0)
                            signBits |= 1u << px;
                        const float em = sse_min(e0, e1);
                        S.hErr[(size_t)(total * 16 + px) * S.stride] = em;
                        floorErr[px] = fminf(floorErr[px], em);
                    }
                    S.hMeta[(size_t)total * S.stride] = signBits | ((uint32_t)packed << 16);
                    total++;
                }
            }

            // colour pairs in the reference's stepping order (ETC.cpp:800-822); a lane only needs its own n0 * n1 steps
            const int n0 = numUnique[0], n1 = numUnique[1];
            // No pair can do better, pixel by pixel, than the smallest error any candidate of this table reaches there, and a
            // sequential fp32 sum is monotonic in its terms: if even that floor does not beat the lane's best, none of the
            // n0 * n1 pair sums can, and the lane skips them (the reference would evaluate and reject every one).
            float floorTotal = 0.0f;
#pragma unroll
            for (int px = 0; px < 16; px++)
                floorTotal = fadd(floorTotal, floorErr[px]);
            const int combos = (floorTotal < best.error) ? n0 * n1 : 0;
            // The second colour only changes every n0 steps: its 16 per-pixel errors stay in registers in between, which halves
            // the scratch loads of this loop (they were 10 % of the kernel's instructions and 12 % of its stall samples).
            int index0 = 0, index1 = 0, loaded1 = -1;
            float row1[16];
            uint32_t m1 = 0;
#pragma unroll
            for (int px = 0; px < 16; px++)
                row1[px] = 0.0f;
            for (int combo = 0; combo < combos; combo++)
            {
                index0++;
                const bool overflow = (n0 - 1) < index0;
                if (overflow)
                    index0 = 0;
                index1 = imin(n1 - 1, index1 + (overflow ? 1 : 0));
                const int ci0 = index0, ci1 = index1 + n0;
                const uint32_t m0 = S.hMeta[(size_t)ci0 * S.stride];
                if (ci1 != loaded1)
                {
                    loaded1 = ci1;
                    m1 = S.hMeta[(size_t)ci1 * S.stride];
#pragma unroll
                    for (int px = 0; px < 16; px++)
                        row1[px] = S.hErr[(size_t)(ci1 * 16 + px) * S.stride];
                }

                float totalError = 0.0f;
                uint32_t sectorBits = 0;
#pragma unroll
                for (int px = 0; px < 16; px++)
                {
                    const float e0 = S.hErr[(size_t)(ci0 * 16 + px) * S.stride], e1 = row1[px];
                    totalError = fadd(totalError, sse_min(e0, e1));
                    if (e1 < e0)
                        sectorBits |= 1u << px;
                }
                if (totalError < best.error)
                {
                    best.error = totalError;
                    bestIsThisMode = true;
                    bestTable = table;
                    bestColors[0] = (int)(m0 >> 16);
                    bestColors[1] = (int)(m1 >> 16);
                    bestSectorBits = sectorBits;
                    bestSignBits = ((m1 & sectorBits) | (m0 & ~sectorBits)) & 0xffffu;
                }
            }
        }

        if (bestIsThisMode)
            etc_emit_h(best, bestColors, bestSectorBits, bestSignBits, bestTable, true);
    }

    // ---------------------------------------------------------------------------------------------------------
    // TestHalfBlock, ETC.cpp:94-149
    template<bool UNIFORM, bool BT709, int STRIDE>
    CVTT_HD float etc_test_half_block(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, int flip, int sector, int packedColor, int table, bool differential, uint32_t &outSelectors)
    {
        int mod[4][3];
        float modW[4][3];
        for (int ch = 0; ch < 3; ch++)
        {
            const int q = (packedColor >> (ch * 5)) & 31;
            const int unq = differential ? ((q << 3) | (q >> 2)) : ((q << 4) | q);
            for (int s = 0; s < 4; s++)
                mod[s][ch] = imin(imax(unq + T.etc1Modifiers[table][s], 0), 255);
        }
        for (int s = 0; s < 4; s++)
            etc_weigh<UNIFORM, BT709>(P, mod[s], modW[s]);

        uint32_t selectors = 0;
        float totalError = 0.0f;
        if (!UNIFORM && !BT709)
        {
            // The weighted squared distances of two selectors' colours run side by side in packed fp32 (each lane performs
            // the reference's sequence of individually rounded operations, cw - p == cw + (-p) exactly); the scan for the
            // first smallest one stays scalar.  This function is 40 % of the kernel's instructions.
            f2 mw[2][3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
            {
                mw[0][ch] = f2_make(modW[0][ch], modW[1][ch]);
                mw[1][ch] = f2_make(modW[2][ch], modW[3][ch]);
            }
#pragma unroll
            for (int px = 0; px < 8; px++)
            {
                const F4 p = L.pw[etc_flip_pixel(flip, sector, px) * STRIDE];
                float e4[4];
#pragma unroll
                for (int h = 0; h < 2; h++)
                {
                    const f2 e = etc_error_pair(p, mw[h]);
                    e4[2 * h] = e.x;
                    e4[2 * h + 1] = e.y;
                }
                float bestError = FLT_MAX;
                uint32_t bestSelector = 0;
#pragma unroll
                for (int s = 0; s < 4; s++)
                {
                    if (e4[s] < bestError)
                        bestSelector = (uint32_t)s;
                    bestError = sse_min(e4[s], bestError);
                }
                totalError = fadd(totalError, bestError);
                selectors |= bestSelector << (px * 2);
            }
            outSelectors = selectors;
            return totalError;
        }
#pragma unroll
        for (int px = 0; px < 8; px++)
        {
            const F4 p = L.pw[etc_flip_pixel(flip, sector, px) * STRIDE];
            float bestError = FLT_MAX;
            uint32_t bestSelector = 0;
#pragma unroll
            for (int s = 0; s < 4; s++)
            {
                const float e = etc_error<UNIFORM, BT709>(p, mod[s], modW[s]);
                if (e < bestError)
                    bestSelector = (uint32_t)s;
                bestError = sse_min(e, bestError);
            }
            totalError = fadd(totalError, bestError);
            selectors |= bestSelector << (px * 2);
        }
        outSelectors = selectors;
        return totalError;
    }

    CVTT_HD bool etc_differential_legal(int a, int b)
    {
        for (int ch = 0; ch < 3; ch++)
        {
            const int diff = ((b >> (ch * 5)) & 31) - ((a >> (ch * 5)) & 31);
            if (diff < -4 || diff > 3)
                return false;
        }
        return true;
    }

    // ---- warp-cooperative helpers: in the kernel the 32 lanes of a warp, on the CPU (tests/hostsim) one lane on its own ----
    CVTT_HD int coop_lane()
    {
#if defined(__CUDA_ARCH__)
        return (int)(threadIdx.x & 31u);
#else
        return 0;
#endif
    }
    CVTT_HD int coop_width()
    {
#if defined(__CUDA_ARCH__)
        return 32;
#else
        return 1;
#endif
    }
    CVTT_HD uint32_t coop_ballot(bool x)
    {
#if defined(__CUDA_ARCH__)
        return __ballot_sync(0xffffffffu, x);
#else
        return x ? 1u : 0u;
#endif
    }
    CVTT_HD float coop_bcast(float v, int src)
    {
#if defined(__CUDA_ARCH__)
        return __shfl_sync(0xffffffffu, v, src);
#else
        (void)src;
        return v;
#endif
    }
    CVTT_HD int coop_bcast(int v, int src)
    {
#if defined(__CUDA_ARCH__)
        return __shfl_sync(0xffffffffu, v, src);
#else
        (void)src;
        return v;
#endif
    }
    // lexicographic minimum of (e, i) over the lanes; i < 0 marks "no entry" and loses against everything
    CVTT_HD void coop_min_entry(float &e, int &i)
    {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
        {
            const float oe = __shfl_xor_sync(0xffffffffu, e, step);
            const int oi = __shfl_xor_sync(0xffffffffu, i, step);
            const bool take = oi >= 0 && (i < 0 || oe < e || (oe == e && oi < i));
            e = take ? oe : e;
            i = take ? oi : i;
        }
#else
        (void)e;
        (void)i;
#endif
    }

    // FindBestDifferentialCombination, ETC.cpp:219-362.  winMeta receives the pair to encode (colour | table << 15 per
    // sector), selMeta the attempts whose selectors go with it (they differ only when a transparent sector 0 borrows sector
    // 1's colour); both stay -1 when no legal pair beats bestErrorIn.
    //
    // The attempts arrive already filtered: per sector the smallest error of ALL attempts with its colour / table (the first
    // one on ties, in generation order) and, in the scratch arrays, only the kept[sector] attempts with error < bestErrorIn,
    // in generation order -- the only ones the pair search can use.  The differential stages filter while they generate (the
    // block's best error cannot change during that stage), which leaves a fraction of the reference's
    // DifferentialResolveStorage stores and no second pass over them.
    //
    // Most blocks are settled by their two best attempts.  The rest -- a few per cent, but with up to 624 attempts per sector
    // -- need the reference's pair scan: its lists sorted by (error, index), sector 0 walked in that order, the partner of an
    // entry being the first legal one of sector 1 below the error still allowed.  Done by one lane that is a chain of tens of
    // thousands of dependent scratch loads on which the whole CTA waits at its next rendezvous, so the warp does it TOGETHER,
    // one such block after the other: each step of the walk is a "next larger key" selection over sector 0 and a "smallest
    // legal key" selection over sector 1, both as strided scans with a lexicographic shuffle reduction.  Every lane of the
    // warp must call this function (the control flow around it is warp-uniform).
    CVTT_HD void etc_find_best_differential_kept(const ETCScratch &S, const int *kept, const float *bestDiffErrors, const uint32_t *bestDiffMeta, bool canIgnore0,
        float bestErrorIn, int *winMeta, int *selMeta, float &winTotal)
    {
        const float blockBestTotalError = bestErrorIn;
        bool needScan = false;
        if (fadd(bestDiffErrors[0], bestDiffErrors[1]) < blockBestTotalError)
        {
            // with punch-through a fully transparent sector 0 takes the colour of sector 1 and makes any pair legal (ETC.cpp:251-260)
            uint32_t pairMeta0 = bestDiffMeta[0];
            if (canIgnore0)
                pairMeta0 = (bestDiffMeta[0] & ~0x7fffu) | (bestDiffMeta[1] & 0x7fffu);
            if (canIgnore0 || etc_differential_legal((int)(pairMeta0 & 0x7fffu), (int)(bestDiffMeta[1] & 0x7fffu)))
            {
                winMeta[0] = (int)pairMeta0;
                winMeta[1] = (int)bestDiffMeta[1];
                selMeta[0] = (int)bestDiffMeta[0];
                selMeta[1] = (int)bestDiffMeta[1];
                winTotal = fadd(bestDiffErrors[0], bestDiffErrors[1]);
            }
            else
                needScan = true;
        }

        const int lane = coop_lane(), width = coop_width();
        for (uint32_t todo = coop_ballot(needScan); todo; todo &= todo - 1)
        {
            // the block of lane `owner`: its scratch entries sit (owner - lane) elements from this lane's
            const int owner = ctz32(todo);
            const ptrdiff_t shift = (ptrdiff_t)owner - (ptrdiff_t)lane;
            const float *err = S.drsErr + shift;
            const uint32_t *meta = S.drsMeta + shift;
            const int kept0 = coop_bcast(kept[0], owner), kept1 = coop_bcast(kept[1], owner);
            const float bestDiff1 = coop_bcast(bestDiffErrors[1], owner);
            float current = coop_bcast(blockBestTotalError, owner);
            int found0 = -1, found1 = -1;           // metas of the pair found so far (identical on every lane)
            float lastErr = -1.0f;
            int lastIdx = -1;
            for (;;)
            {
                // next entry of sector 0 in (error, index) order
                float e0 = 0.0f;
                int i0 = -1;
                for (int i = lane; i < kept0; i += width)
                {
                    const float e = err[(size_t)i * S.stride];
                    const bool after = (e > lastErr) || (e == lastErr && i > lastIdx);
                    if (after && (i0 < 0 || e < e0))
                    {
                        i0 = i;
                        e0 = e;
                    }
                }
                coop_min_entry(e0, i0);
                if (i0 < 0)
                    break;
                lastErr = e0;
                lastIdx = i0;
                if (e0 >= current)
                    break;
                const float maxError1 = fsub(current, e0);
                if (maxError1 < bestDiff1)
                    break;
                const uint32_t m0 = meta[(size_t)i0 * S.stride];
                // its partner: the smallest (error, index) of sector 1 below maxError1 whose colour makes a legal pair
                float e1 = 0.0f;
                int j1 = -1;
                for (int j = lane; j < kept1; j += width)
                {
                    const size_t slot = (size_t)(kETCMaxAttempts + j) * S.stride;
                    const float e = err[slot];
                    if (e < maxError1 && (j1 < 0 || e < e1) && etc_differential_legal((int)(m0 & 0x7fffu), (int)(meta[slot] & 0x7fffu)))
                    {
                        j1 = j;
                        e1 = e;
                    }
                }
                coop_min_entry(e1, j1);
                if (j1 >= 0)
                {
                    current = fadd(e0, e1);
                    found0 = (int)m0;
                    found1 = (int)meta[(size_t)(kETCMaxAttempts + j1) * S.stride];
                }
            }
            if (lane == owner && found0 >= 0)
            {
                winMeta[0] = selMeta[0] = found0;
                winMeta[1] = selMeta[1] = found1;
                winTotal = current;
            }
        }
    }

    // CompressETC1BlockInternal, ETC.cpp:2624-2882.  MIN_D = 1 is the ETC2 call (differential only), 0 is ETC1.
    template<bool UNIFORM, bool BT709, int MIN_D, int STRIDE>
    CVTT_HD void etc_etc1(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, const ETCScratch &S, ETCBest &best)
    {
        bool bestIsThisMode = false;
        int bestColors[2] = { 0, 0 }, bestTables[2] = { 0, 0 }, bestFlip = 0, bestD = 0;
        uint32_t bestSelectors[2] = { 0, 0 };

#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int flip = 0; flip < 2; flip++)
        {
            int cumulative[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } };
            for (int sector = 0; sector < 2; sector++)
                for (int px = 0; px < 8; px++)
                {
                    const F4 p = L.pw[etc_flip_pixel(flip, sector, px) * STRIDE];
                    for (int ch = 0; ch < 3; ch++)
                        cumulative[sector][ch] += etc_px(p, ch);
                }

            int kept[2] = { 0, 0 };
            float bestDiffErrors[2] = { FLT_MAX, FLT_MAX };
            uint32_t bestDiffMeta[2] = { 0, 0 };
            float bestIndError[2] = { FLT_MAX, FLT_MAX };
            uint32_t bestIndSelectors[2] = { 0, 0 };
            int bestIndColors[2] = { 0, 0 }, bestIndTable[2] = { 0, 0 };

            for (int d = MIN_D; d < 2; d++)
            {
                for (int sector = 0; sector < 2; sector++)
                {
                    const int16_t *potentialOffsets = T.potentialOffsets;
                    // One rendezvous where the CTA enters this stage: flips and sectors run the same loop body (the flip loop is
                    // not unrolled), so the warps stay in the same code without meeting again; measured +1 % against a
                    // rendezvous per (flip, sector) and before the pair search.
                    if (flip == 0 && sector == 0 && d == MIN_D)
                        cta_sync();
                    for (int table = 0; table < 8; table++)
                    {
                        const int numOffsets = *potentialOffsets++;
                        int lastColor = -1;
                        for (int oi = 0; oi < numOffsets; oi++)
                        {
                            int packed = 0;
                            if (BT709)
                            {
                                int cu[3], q[3];
                                for (int ch = 0; ch < 3; ch++)
                                    cu[ch] = imin(2040, imax(0, cumulative[sector][ch] + potentialOffsets[oi]));
                                etc_resolve_half_block_bt709(q, cu, d == 1, (P.flags & kFlag_ETC_FakeBT709Accurate) != 0);
                                packed = q[0] | (q[1] << 5) | (q[2] << 10);
                            }
                            else
                            for (int ch = 0; ch < 3; ch++)
                            {
                                const int cu = imin(2040, imax(0, cumulative[sector][ch] + potentialOffsets[oi]));
                                const int q = (d == 1) ? (((cu << 5) - cu + (cu >> 3) + 1024) >> 11) : (((cu << 5) - (cu << 1) + (cu >> 3) + 2048) >> 12);
                                packed |= q << (ch * 5);
                            }
                            if (oi != 0 && packed == lastColor)
                                continue;       // adjacent duplicates are dropped (ETC.cpp:2756-2768)
                            lastColor = packed;

                            uint32_t selectors;
                            const float error = etc_test_half_block<UNIFORM, BT709, STRIDE>(P, T, L, flip, sector, packed, table, d == 1, selectors);
                            if (d == 0)
                            {
                                if (error < bestIndError[sector])
                                {
                                    bestIndError[sector] = error;
                                    bestIndSelectors[sector] = selectors;
                                    bestIndColors[sector] = packed;
                                    bestIndTable[sector] = table;
                                }
                            }
                            else
                            {
                                // filtered as it is generated, see etc_find_best_differential_kept (best.error is fixed here)
                                const uint32_t meta = (uint32_t)packed | ((uint32_t)table << 15);
                                if (error < bestDiffErrors[sector])
                                {
                                    bestDiffErrors[sector] = error;
                                    bestDiffMeta[sector] = meta;
                                }
                                if (error < best.error)
                                {
                                    const size_t slot = (size_t)(sector * kETCMaxAttempts + kept[sector]) * S.stride;
                                    S.drsErr[slot] = error;
                                    S.drsMeta[slot] = meta;
                                    kept[sector]++;
                                }
                            }
                        }
                        potentialOffsets += numOffsets;
                    }
                }

                if (d == 0)
                {
                    const float total = fadd(bestIndError[0], bestIndError[1]);
                    if (total < best.error)
                    {
                        bestIsThisMode = true;
                        best.error = total;
                        bestFlip = flip;
                        bestD = 0;
                        for (int sector = 0; sector < 2; sector++)
                        {
                            bestColors[sector] = bestIndColors[sector];
                            bestSelectors[sector] = bestIndSelectors[sector];
                            bestTables[sector] = bestIndTable[sector];
                        }
                    }
                }
                else
                {
                    int winMeta[2] = { -1, -1 }, selMeta[2] = { -1, -1 };
                    float winTotal = 0.0f;
                    etc_find_best_differential_kept(S, kept, bestDiffErrors, bestDiffMeta, false, best.error, winMeta, selMeta, winTotal);
                    if (winMeta[0] >= 0)
                    {
                        bestIsThisMode = true;
                        best.error = winTotal;
                        bestFlip = flip;
                        bestD = 1;
                        for (int sector = 0; sector < 2; sector++)
                        {
                            bestColors[sector] = winMeta[sector] & 0x7fff;
                            bestTables[sector] = (winMeta[sector] >> 15) & 7;
                            // the selectors are a function of (colour, table); recomputed instead of stored per attempt
                            etc_test_half_block<UNIFORM, BT709, STRIDE>(P, T, L, flip, sector, selMeta[sector] & 0x7fff, bestTables[sector], true, bestSelectors[sector]);
                        }
                    }
                }
            }
        }

        if (bestIsThisMode)
        {
            int colors[2][3];
            for (int sector = 0; sector < 2; sector++)
                for (int ch = 0; ch < 3; ch++)
                    colors[sector][ch] = (bestColors[sector] >> (ch * 5)) & 31;
            etc_emit_etc1(best, bestFlip, bestD, colors, bestTables, bestSelectors);
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // Punch-through alpha (EncodeETC2PunchthroughAlpha).  transparentMask: bit px set = alpha below the threshold; the lane's
    // stored pixels are already zero there (CompressETC2Block, ETC.cpp:1706-1719).

    // TestHalfBlockPunchthrough, ETC.cpp:151-217
    template<bool UNIFORM, bool BT709, int STRIDE>
    CVTT_HD float etc_test_half_block_punchthrough(const ETCParams &P, const ETCLane<STRIDE> &L, int flip, int sector, int packedColor, int modifier, uint32_t transparentMask, uint32_t &outSelectors)
    {
        int mod[3][3];
        float modW[3][3];
        for (int ch = 0; ch < 3; ch++)
        {
            const int q = (packedColor >> (ch * 5)) & 31;
            const int unq = (q << 3) | (q >> 2);
            mod[0][ch] = imax(unq, modifier) - modifier;
            mod[1][ch] = unq;
            mod[2][ch] = imin(unq + modifier, 255);
        }
        for (int s = 0; s < 3; s++)
            etc_weigh<UNIFORM, BT709>(P, mod[s], modW[s]);
        uint32_t selectors = 0;
        float totalError = 0.0f;
        for (int px = 0; px < 8; px++)
        {
            const int bp = etc_flip_pixel(flip, sector, px);
            const F4 p = L.pw[bp * STRIDE];
            float bestError = FLT_MAX;
            uint32_t bestSelector = 0;
            for (int s = 0; s < 3; s++)
            {
                const float e = etc_error<UNIFORM, BT709>(p, mod[s], modW[s]);
                if (e < bestError)
                    bestSelector = (uint32_t)s;
                bestError = sse_min(e, bestError);
            }
            // table order vs encoding order: the transparent code is 1, the colours are 0, 2, 3 (ETC.cpp:199-208)
            bestSelector = (bestSelector << 1) < 3u ? (bestSelector << 1) : 3u;
            if ((transparentMask >> bp) & 1)
            {
                bestError = 0.0f;
                bestSelector = 1;
            }
            totalError = fadd(totalError, bestError);
            selectors |= bestSelector << (px * 2);
        }
        outSelectors = selectors;
        return totalError;
    }

    // CompressETC1PunchthroughBlockInternal, ETC.cpp:2885-3080.  Quirks kept: the per-sector count the reference calls
    // "sectorNumOpaque" counts the *transparent* pixels (:2954-2956), and only sector 0 can ever be ignored (:2944).
    template<bool UNIFORM, bool BT709, int STRIDE>
    CVTT_HD void etc_etc1_punchthrough(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, const ETCScratch &S, uint32_t transparentMask, ETCBest &best)
    {
        bool bestIsThisMode = false;
        int bestColors[2] = { 0, 0 }, bestTables[2] = { 0, 0 }, bestFlip = 0;
        uint32_t bestSelectors[2] = { 0, 0 };

        for (int flip = 0; flip < 2; flip++)
        {
            int cumulative[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } }, numT[2] = { 0, 0 };
            for (int sector = 0; sector < 2; sector++)
                for (int px = 0; px < 8; px++)
                {
                    const int bp = etc_flip_pixel(flip, sector, px);
                    const F4 p = L.pw[bp * STRIDE];
                    for (int ch = 0; ch < 3; ch++)
                        cumulative[sector][ch] += etc_px(p, ch);
                    numT[sector] += (int)((transparentMask >> bp) & 1);
                }
            const bool canIgnore0 = numT[0] == 8;

            int kept[2] = { 0, 0 };
            float bestDiffErrors[2] = { FLT_MAX, FLT_MAX };
            uint32_t bestDiffMeta[2] = { 0, 0 };
            for (int sector = 0; sector < 2; sector++)
            {
                const int n = numT[sector];
                const int denominator = imax(1, n) << 8, addend = n << 7, cumulativeMax = wrap_u16(255 * n);
                const UDivisor denominatorDivide = udiv_prepare((uint32_t)denominator);
                for (int table = 0; table < 8; table++)
                {
                    cta_sync();
                    const int modifier = -T.etc1Modifiers[table][0];       // 8, 17, 29, 42, 60, 80, 106, 183
                    int lastColor = -1;
                    for (int om = -n; om <= n; om++)
                    {
                        const int offset = wrap_s16(om * modifier);
                        int packed = 0;
                        for (int ch = 0; ch < 3; ch++)
                        {
                            const int cu = imin(cumulativeMax, imax(0, wrap_s16(cumulative[sector][ch] + offset)));
                            const int numerator = wrap_u16(wrap_u16((cu << 5) - cu) + wrap_u16((cu >> 3) + addend));
                            packed |= (int)udiv((uint32_t)numerator, denominatorDivide) << (ch * 5);
                        }
                        if (om != -n && packed == lastColor)
                            continue;
                        lastColor = packed;
                        uint32_t selectors;
                        const float error = etc_test_half_block_punchthrough<UNIFORM, BT709, STRIDE>(P, L, flip, sector, packed, modifier, transparentMask, selectors);
                        // filtered as it is generated, see etc_find_best_differential_kept (best.error is fixed in this stage)
                        const uint32_t meta = (uint32_t)packed | ((uint32_t)table << 15);
                        if (error < bestDiffErrors[sector])
                        {
                            bestDiffErrors[sector] = error;
                            bestDiffMeta[sector] = meta;
                        }
                        if (error < best.error)
                        {
                            const size_t slot = (size_t)(sector * kETCMaxAttempts + kept[sector]) * S.stride;
                            S.drsErr[slot] = error;
                            S.drsMeta[slot] = meta;
                            kept[sector]++;
                        }
                    }
                }
            }

            cta_sync();
            int winMeta[2] = { -1, -1 }, selMeta[2] = { -1, -1 };
            float winTotal = 0.0f;
            etc_find_best_differential_kept(S, kept, bestDiffErrors, bestDiffMeta, canIgnore0, best.error, winMeta, selMeta, winTotal);
            if (winMeta[0] >= 0)
            {
                bestIsThisMode = true;
                best.error = winTotal;
                bestFlip = flip;
                for (int sector = 0; sector < 2; sector++)
                {
                    bestColors[sector] = winMeta[sector] & 0x7fff;
                    bestTables[sector] = (winMeta[sector] >> 15) & 7;
                    etc_test_half_block_punchthrough<UNIFORM, BT709, STRIDE>(P, L, flip, sector, selMeta[sector] & 0x7fff, -T.etc1Modifiers[bestTables[sector]][0], transparentMask, bestSelectors[sector]);
                }
            }
        }

        if (bestIsThisMode)
        {
            int colors[2][3];
            for (int sector = 0; sector < 2; sector++)
                for (int ch = 0; ch < 3; ch++)
                    colors[sector][ch] = (bestColors[sector] >> (ch * 5)) & 31;
            etc_emit_etc1(best, bestFlip, 1, colors, bestTables, bestSelectors, true);
        }
    }

    // EncodeVirtualTModePunchthrough, ETC.cpp:887-1262: T and H mode share their colour set once code 2 means "transparent".
    // Cross-lane: the offset walk starts at minus the largest line-pixel count of the call and steps by two (:1038), and the
    // candidate list has the same never-written slot as the opaque T mode (:1109-1116).
    template<bool UNIFORM, bool BT709, int STRIDE, class Vote>
    CVTT_HD void etc_virtual_t_punchthrough(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, Vote &vote, uint32_t isolatedBaseMask, uint32_t transparentMask, ETCBest &best)
    {
        const uint32_t opaqueMask = ~transparentMask & 0xffffu;
        const uint32_t isolatedMask = isolatedBaseMask & opaqueMask, lineMask = ~isolatedBaseMask & opaqueMask;
        int isolatedTotal[3] = { 0, 0, 0 }, lineTotal[3] = { 0, 0, 0 }, numIsolated = 0, numLine = 0;
        for (int px = 0; px < 16; px++)
        {
            const F4 p = L.pw[px * STRIDE];
            for (int ch = 0; ch < 3; ch++)
            {
                const int v = etc_px(p, ch);
                if ((isolatedMask >> px) & 1)
                    isolatedTotal[ch] += v;
                if ((lineMask >> px) & 1)
                    lineTotal[ch] += v;
            }
            numIsolated += (int)((isolatedMask >> px) & 1);
            numLine += (int)((lineMask >> px) & 1);
        }

        int isolatedQ[3], hQ[8][3], isolatedColor[3];
        {
            const int divisor = numIsolated * 34, addend = (numIsolated << 4) | numIsolated;
            int targets[3];
            for (int ch = 0; ch < 3; ch++)
            {
                const int numerator = wrap_u16(isolatedTotal[ch] + isolatedTotal[ch] + (BT709 ? 0 : addend));
                isolatedQ[ch] = (divisor == 0) ? 0 : (numerator / divisor);
                targets[ch] = numerator;
                for (int table = 0; table < 8; table++)
                {
                    const int offsetTotal = wrap_u16(isolatedTotal[ch] + wrap_u16(T.thModifier[table] * numIsolated));
                    const int hNumerator = wrap_u16(offsetTotal + offsetTotal + addend);
                    hQ[table][ch] = (divisor == 0) ? 0 : (hNumerator / divisor);
                }
            }
            if (BT709)
                etc_resolve_th_bt709(isolatedQ, targets, numIsolated);
            for (int table = 0; table < 8; table++)
                for (int ch = 0; ch < 3; ch++)
                    hQ[table][ch] = imin(15, hQ[table][ch]);
            for (int ch = 0; ch < 3; ch++)
                isolatedColor[ch] = wrap_u16(isolatedQ[ch] | (isolatedQ[ch] << 4));
        }

        float isolatedError[16];
        {
            float isoW[3];
            etc_weigh<UNIFORM, BT709>(P, isolatedColor, isoW);
            for (int px = 0; px < 16; px++)
                isolatedError[px] = ((transparentMask >> px) & 1) ? 0.0f : etc_error<UNIFORM, BT709>(L.pw[px * STRIDE], isolatedColor, isoW);
        }

        bool bestIsThisMode = false, bestIsHMode = false;
        uint32_t bestSelectors = 0;
        int bestTable = 0, bestLineColor = 0, bestHModeColor2 = 0;

        const int lineDivisor = numLine * 34, lineAddend = (numLine << 4) | numLine;
        const UDivisor lineDivide = udiv_prepare((uint32_t)lineDivisor);
        const int clusterMaxLine = vote.max(numLine);

        for (int table = 0; table < 8; table++)
        {
            cta_sync();
            const int modifier = T.thModifier[table];
            const int modifierOffset = modifier + modifier;

            auto lineColorAt = [&](int offs) -> int
            {
                const int clamped = imax(-numLine, imin(numLine, offs));
                int q[3], targets[3];
                for (int ch = 0; ch < 3; ch++)
                {
                    const int numerator = imax(0, wrap_s16(wrap_s16(lineTotal[ch] + lineTotal[ch] + (BT709 ? 0 : lineAddend)) + wrap_s16(clamped * modifierOffset)));
                    const int divided = (lineDivisor == 0) ? 0 : (int)udiv((uint32_t)numerator, lineDivide);
                    q[ch] = imin(15, divided);
                    targets[ch] = numerator;
                }
                if (BT709)
                    etc_resolve_th_bt709(q, targets, numLine);
                return (q[0] << 10) | (q[1] << 5) | q[2];
            };

            int numUnique = 0, lastColor = -1;
            for (int offs = -clusterMaxLine; offs <= clusterMaxLine; offs += 2)
            {
                const int packed = lineColorAt(offs);
                if (numUnique == 0 || packed != lastColor)
                {
                    numUnique++;
                    lastColor = packed;
                }
            }
            const int maxUnique = vote.max(numUnique);
            const int numCandidates = numUnique + ((numUnique < maxUnique) ? 1 : 0);

            int hModeColor[3];
            float hW[3], hModeErrors[16];
            for (int ch = 0; ch < 3; ch++)
            {
                const int q = hQ[table][ch];
                hModeColor[ch] = imax(0, ((q << 4) | q) - modifier);
            }
            etc_weigh<UNIFORM, false>(P, hModeColor, hW);
            for (int px = 0; px < 16; px++)
                hModeErrors[px] = ((transparentMask >> px) & 1) ? 0.0f : etc_error<UNIFORM, false>(L.pw[px * STRIDE], hModeColor, hW);
            const int packedHModeColor2 = (hQ[table][0] << 10) | (hQ[table][1] << 5) | hQ[table][2];
            const bool tableLowBitIsZero = (table & 1) == 0;

            int offs = -clusterMaxLine;
            lastColor = -1;
            for (int ci = 0; ci < numCandidates; ci++)
            {
                int packedColor = 0;
                if (ci < numUnique)
                {
                    for (;;)
                    {
                        const int packed = lineColorAt(offs);
                        offs += 2;
                        if (packed != lastColor)
                        {
                            lastColor = packed;
                            packedColor = packed;
                            break;
                        }
                    }
                }

                int lineColors[2][3];
                float lineW[2][3];
                for (int ch = 0; ch < 3; ch++)
                {
                    const int q = (packedColor >> (10 - ch * 5)) & 15;
                    const int unq = (q << 4) | q;
                    lineColors[0][ch] = imin(255, unq + modifier);
                    lineColors[1][ch] = imax(0, unq - modifier);
                }
                etc_weigh<UNIFORM, false>(P, lineColors[0], lineW[0]);
                etc_weigh<UNIFORM, false>(P, lineColors[1], lineW[1]);

                float bestLineError[16];
                uint32_t lineSelectors = 0;
                float tModeError = 0.0f, hModeError = 0.0f;
                for (int px = 0; px < 16; px++)
                {
                    const F4 p = L.pw[px * STRIDE];
                    const float e0 = etc_error<UNIFORM, false>(p, lineColors[0], lineW[0]), e1 = etc_error<UNIFORM, false>(p, lineColors[1], lineW[1]);
                    lineSelectors |= ((e0 <= e1) ? 1u : 3u) << (px * 2);
                    float e = sse_min(e0, e1);
                    if ((transparentMask >> px) & 1)
                        e = 0.0f;
                    bestLineError[px] = e;
                    tModeError = fadd(tModeError, sse_min(e, isolatedError[px]));
                    hModeError = fadd(hModeError, sse_min(e, hModeErrors[px]));
                }

                const bool hLessError = hModeError < tModeError;
                const bool hModeTableLowBitMustBeZero = packedColor < packedHModeColor2;
                const bool useHMode = hLessError && (hModeTableLowBitMustBeZero == tableLowBitIsZero);
                const float roundBestError = useHMode ? hModeError : tModeError;
                if (roundBestError < best.error)
                {
                    uint32_t selectors = 0;
                    for (int px = 0; px < 16; px++)
                    {
                        uint32_t selector = (lineSelectors >> (px * 2)) & 3u;
                        const float isolatedPixelError = useHMode ? hModeErrors[px] : isolatedError[px];
                        if (isolatedPixelError < bestLineError[px])
                            selector = 0;
                        if ((transparentMask >> px) & 1)
                            selector = 2;
                        selectors |= selector << (px * 2);
                    }
                    best.error = roundBestError;
                    bestLineColor = packedColor;
                    bestSelectors = selectors;
                    bestTable = table;
                    bestIsHMode = useHMode;
                    bestHModeColor2 = packedHModeColor2;
                    bestIsThisMode = true;
                }
            }
        }

        if (bestIsThisMode)
        {
            if (bestIsHMode)
            {
                // T mode: C1, C2+M, transparent, C2-M;  H mode: C1+M, C1-M, transparent, C2-M  (ETC.cpp:1236-1256)
                uint32_t signBits = 0, sectorBits = 0;
                for (int px = 0; px < 16; px++)
                {
                    const uint32_t selector = (bestSelectors >> (px * 2)) & 3u;
                    sectorBits |= ((0x5u >> selector) & 1u) << px;        // selectorRemapSector { 1, 0, 1, 0 }
                    signBits |= ((0x9u >> selector) & 1u) << px;          // selectorRemapSign   { 1, 0, 0, 1 }
                }
                const int blockColors[2] = { bestLineColor, bestHModeColor2 };
                etc_emit_h(best, blockColors, sectorBits, signBits, bestTable, false);
            }
            else
            {
                int lineColor[3];
                for (int ch = 0; ch < 3; ch++)
                    lineColor[ch] = (bestLineColor >> (10 - ch * 5)) & 15;
                etc_emit_t(best, lineColor, isolatedQ, bestSelectors, bestTable, false);
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // CompressETC2Block without punch-through (ETC.cpp:1664-1887): chroma split, then planar, T, T, H, differential
    // chroma-plane split of CompressETC2Block (ETC.cpp:1723-1848); numOpaque is 16 without punch-through
    template<bool UNIFORM, int STRIDE>
    CVTT_HD uint32_t etc2_chroma_split(const ETCParams &P, const ETCLane<STRIDE> &L, int numOpaque)
    {
        float chromaDelta[16][2];
        if (UNIFORM)
        {
            int coords[16][2], centroid[2] = { 0, 0 };
            for (int px = 0; px < 16; px++)
            {
                const F4 p = L.pw[px * STRIDE];
                const int r = etc_px(p, 0), g = etc_px(p, 1), b = etc_px(p, 2);
                coords[px][0] = r - b;
                coords[px][1] = r - (g << 1) + b;
                centroid[0] += coords[px][0];
                centroid[1] += coords[px][1];
            }
            for (int px = 0; px < 16; px++)
            {
                chromaDelta[px][0] = (float)(coords[px][0] * numOpaque - centroid[0]);
                chromaDelta[px][1] = fmul((float)(coords[px][1] * numOpaque - centroid[1]), 0.57735026918962576450914878050196f);
            }
        }
        else
        {
            float coords[16][2], centroid[2] = { 0.0f, 0.0f };
            const float numOpaqueF = (float)numOpaque;
            for (int px = 0; px < 16; px++)
            {
                const F4 p = L.pw[px * STRIDE];
                coords[px][0] = fadd(fadd(fmul(p.x, P.chromaAxis0[0]), fmul(p.y, P.chromaAxis0[1])), fmul(p.z, P.chromaAxis0[2]));
                coords[px][1] = fadd(fadd(fmul(p.x, P.chromaAxis1[0]), fmul(p.y, P.chromaAxis1[1])), fmul(p.z, P.chromaAxis1[2]));
            }
            for (int px = 0; px < 16; px++)
                for (int ch = 0; ch < 2; ch++)
                    centroid[ch] = fadd(centroid[ch], coords[px][ch]);
            for (int px = 0; px < 16; px++)
                for (int ch = 0; ch < 2; ch++)
                    chromaDelta[px][ch] = fsub(fmul(coords[px][ch], numOpaqueF), centroid[ch]);
        }

        float covXX = 0.0f, covYY = 0.0f, covXY = 0.0f;
        for (int px = 0; px < 16; px++)
        {
            const float nx = chromaDelta[px][0], ny = chromaDelta[px][1];
            covXX = fadd(covXX, fmul(nx, nx));
            covYY = fadd(covYY, fmul(ny, ny));
            covXY = fadd(covXY, fmul(nx, ny));
        }
        const float halfTrace = fmul(fadd(covXX, covYY), 0.5f);
        const float det = fsub(fmul(covXX, covYY), fmul(covXY, covXY));
        const float mm = sqrtf(sse_max(0.0f, fsub(fmul(halfTrace, halfTrace), det)));
        const float ev = fadd(halfTrace, mm);
        float dx = fadd(fsub(covYY, ev), covXY);
        const float dy = fsub(0.0f, fadd(fsub(covXX, ev), covXY));
        if (dx == 0.0f && dy == 0.0f)
            dx = 1.0f;
        uint32_t sectorMask = 0;
        for (int px = 0; px < 16; px++)
            if (fadd(fmul(chromaDelta[px][0], dx), fmul(chromaDelta[px][1], dy)) < 0.0f)
                sectorMask |= 1u << px;
        return sectorMask;
    }

    template<bool UNIFORM, bool BT709, int STRIDE, class Vote>
    CVTT_HD void etc2_encode_block(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, const ETCScratch &S, Vote &vote, uint32_t out[2])
    {
        ETCBest best;
        best.error = FLT_MAX;
        best.hi = best.lo = 0;

        cta_sync();
        etc_planar<UNIFORM, BT709, STRIDE>(P, L, best);

        // The two T-mode passes (isolated colour = one chroma sector, then the other) are ONE loop body: the warps of the CTA
        // then need no rendezvous between them to stay in the same code, and a pass's cost grows with its number of line
        // pixels, which the two passes split between them -- per warp the sum varies far less than either pass.
        uint32_t sectorMask = etc2_chroma_split<UNIFORM, STRIDE>(P, L, 16);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int pass = 0; pass < 2; pass++)
        {
            if (pass)
                sectorMask ^= 0xffffu;
            etc_t_mode<UNIFORM, BT709, STRIDE>(P, T, L, vote, sectorMask, best, pass == 0);
        }
        etc_h_mode<UNIFORM, BT709, STRIDE>(P, T, L, S, sectorMask, best);
        etc_etc1<UNIFORM, BT709, 1, STRIDE>(P, T, L, S, best);

        out[0] = best.hi;
        out[1] = best.lo;
    }

    // CompressETC2Block with punch-through alpha (ETC.cpp:1664-1887).  The lane's pixels below the threshold are already zero
    // (transparentMask).  Cross-lane: the opaque stages run for a group unless all of its lanes are fully transparent, the
    // punch-through stages run if any lane has a transparent pixel; a lane with a transparent pixel restarts from FLT_MAX
    // before the punch-through stages, so only fully opaque lanes keep their opaque-stage result.  The stage barriers are
    // CTA-wide, so a stage runs for the whole CTA when any of its groups needs it and the groups that do not need it drop
    // what it produced.
    template<bool UNIFORM, bool BT709, int STRIDE, class Vote>
    CVTT_HD void etc2_punchthrough_encode_block(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, const ETCScratch &S, Vote &vote, uint32_t transparentMask, uint32_t out[2])
    {
        ETCBest best;
        best.error = FLT_MAX;
        best.hi = best.lo = 0;

        const bool anyTransparent = transparentMask != 0, allTransparent = transparentMask == 0xffffu;
        const bool groupOpaqueStages = vote.max(allTransparent ? 0 : 1) != 0;
        const bool groupPunchStages = vote.max(anyTransparent ? 1 : 0) != 0;
        int numOpaque = 16;
        for (int px = 0; px < 16; px++)
            numOpaque -= (int)((transparentMask >> px) & 1);

        uint32_t sectorMask = etc2_chroma_split<UNIFORM, STRIDE>(P, L, numOpaque);

        if (cta_any(groupOpaqueStages))
        {
            etc_planar<UNIFORM, BT709, STRIDE>(P, L, best);
            etc_t_mode<UNIFORM, BT709, STRIDE>(P, T, L, vote, sectorMask, best);
            etc_t_mode<UNIFORM, BT709, STRIDE>(P, T, L, vote, sectorMask ^ 0xffffu, best);
            etc_h_mode<UNIFORM, BT709, STRIDE>(P, T, L, S, sectorMask ^ 0xffffu, best);
            etc_etc1<UNIFORM, BT709, 1, STRIDE>(P, T, L, S, best);
            if (!groupOpaqueStages)
            {
                best.error = FLT_MAX;
                best.hi = best.lo = 0;
            }
        }

        if (cta_any(groupPunchStages))
        {
            const ETCBest opaqueBest = best;
            if (anyTransparent)
                best.error = FLT_MAX;
            etc_virtual_t_punchthrough<UNIFORM, BT709, STRIDE>(P, T, L, vote, sectorMask, transparentMask, best);
            etc_virtual_t_punchthrough<UNIFORM, BT709, STRIDE>(P, T, L, vote, sectorMask ^ 0xffffu, transparentMask, best);
            etc_etc1_punchthrough<UNIFORM, BT709, STRIDE>(P, T, L, S, transparentMask, best);
            if (!groupPunchStages)
                best = opaqueBest;
        }

        out[0] = best.hi;
        out[1] = best.lo;
    }

    // CompressETC1Block, ETC.cpp:2112-2126
    template<bool UNIFORM, bool BT709, int STRIDE>
    CVTT_HD void etc1_encode_block(const ETCParams &P, const ETCTables &T, const ETCLane<STRIDE> &L, const ETCScratch &S, uint32_t out[2])
    {
        ETCBest best;
        best.error = FLT_MAX;
        best.hi = best.lo = 0;
        etc_etc1<UNIFORM, BT709, 0, STRIDE>(P, T, L, S, best);
        out[0] = best.hi;
        out[1] = best.lo;
    }

    // ---------------------------------------------------------------------------------------------------------
    // CompressETC2AlphaBlockInternal (ETC.cpp:1902-2085): 16 tables x 10 ranges x 2 multipliers, pure integer.
    // pixels: 0..255 (8-bit alpha) or the shifted 11-bit range of CompressEACBlock.  Output: the 8 bytes as two
    // big-endian words like the colour emitters.
    CVTT_HD void etc_alpha_encode_block(const ETCTables &T, const int *pixels, bool is11Bit, bool isSigned, uint32_t out[2])
    {
        int minAlpha = is11Bit ? 2047 : 255, maxAlpha = 0;
        for (int px = 0; px < 16; px++)
        {
            minAlpha = imin(minAlpha, pixels[px]);
            maxAlpha = imax(maxAlpha, pixels[px]);
        }
        const int alphaSpan = maxAlpha - minAlpha, alphaSpanMidpointTimes2 = maxAlpha + minAlpha;

        int bestTotalError = 0x7fffffff, bestTableIndex = 0, bestBaseCodeword = 0, bestMultiplier = 0;
        uint32_t bestIndexes[2] = { 0, 0 };     // 16 x 3 bits in pixel order: pixels 0-9 in [0], 10-15 in [1]

        for (int tableIndex = 0; tableIndex < 16; tableIndex++)
            for (int r = 0; r < 10; r++)
            {
                const int subrange = r % 3, mainRange = r / 3;
                const int maxOffset = T.alphaModifier[tableIndex][3 - mainRange - (subrange & 1)];
                const int minOffset = -T.alphaModifier[tableIndex][3 - mainRange - ((subrange >> 1) & 1)] - 1;
                const int offsetSpan = (maxOffset - minOffset) & 0xffff;

                int minMultiplier = alphaSpan / offsetSpan;
                if (is11Bit)
                    minMultiplier = imin(minMultiplier, 112) & 120;
                else
                    minMultiplier = imax(imin(minMultiplier, 14), 1);

                for (int multiplierOffset = 0; multiplierOffset < 2; multiplierOffset++)
                {
                    int multiplier = minMultiplier;
                    if (is11Bit)
                    {
                        if (multiplierOffset == 1)
                            multiplier += 8;
                        else
                            multiplier = imax(multiplier, 1);
                    }
                    else if (multiplierOffset == 1)
                        multiplier += 1;

                    const int multipliedMinOffset = wrap_s16(multiplier * minOffset);
                    const int multipliedMaxOffset = wrap_s16(multiplier * maxOffset);
                    int unclampedBaseAlphaTimes2 = wrap_s16(alphaSpanMidpointTimes2 - multipliedMaxOffset - multipliedMinOffset);

                    int baseAlpha;
                    if (is11Bit)
                    {
                        if (isSigned)
                            unclampedBaseAlphaTimes2 = wrap_s16(unclampedBaseAlphaTimes2 + 8);
                        const int minBaseAlphaTimes2 = isSigned ? 16 : 0;
                        const int clamped = imin(imax(unclampedBaseAlphaTimes2, minBaseAlphaTimes2), 4095);
                        baseAlpha = (clamped >> 1) & 2040;
                        if (!isSigned)
                            baseAlpha += 4;
                    }
                    else
                    {
                        const int clamped = imin(imax(unclampedBaseAlphaTimes2, 0), 510);
                        baseAlpha = (clamped + 1) >> 1;
                    }

                    uint32_t indexes[2] = { 0, 0 };
                    int totalError = 0;
                    const UDivisor multiplierDivide = udiv_prepare((uint32_t)multiplier);
                    for (int px = 0; px < 16; px++)
                    {
                        // QuantizeETC2Alpha, ETC.cpp:2366-2411
                        const int offset = wrap_s16(pixels[px] - baseAlpha);
                        const int aboutReflectorTimes2 = wrap_s16(offset + offset + multiplier);
                        const int absTimes2 = (aboutReflectorTimes2 < 0 ? -aboutReflectorTimes2 : aboutReflectorTimes2) & 0xffff;
                        int lookup = (int)udiv((uint32_t)(absTimes2 >> 1), multiplierDivide);
                        if (lookup >= 13)
                            lookup = 12;
                        const int positiveIndex = T.alphaRounding[tableIndex][lookup];
                        const int positiveOffset = T.alphaModifier[tableIndex][positiveIndex];
                        const int signBits = aboutReflectorTimes2 >> 15;           // 0 or -1
                        const int offsetUnmultiplied = wrap_s16(positiveOffset ^ signBits);
                        const int quantizedOffset = wrap_s16(offsetUnmultiplied * multiplier);
                        const int offsetValue = wrap_s16(baseAlpha + quantizedOffset);
                        int q;
                        if (is11Bit)
                            q = imin(2047, imax(isSigned ? 1 : 0, offsetValue));
                        else
                            q = imin(255, imax(0, offsetValue));
                        const int index = positiveIndex + 4 - (signBits & 4);
                        if (px < 10)
                            indexes[0] |= (uint32_t)index << (3 * px);
                        else
                            indexes[1] |= (uint32_t)index << (3 * (px - 10));
                        const int delta = q - pixels[px];
                        totalError += is11Bit ? (delta * delta) : ((delta * delta) & 0xffff);
                    }
                    if (totalError < bestTotalError)
                    {
                        bestTotalError = totalError;
                        bestTableIndex = tableIndex;
                        bestBaseCodeword = baseAlpha;
                        bestMultiplier = multiplier;
                        bestIndexes[0] = indexes[0];
                        bestIndexes[1] = indexes[1];
                    }
                }
            }

        if (is11Bit)
        {
            bestMultiplier >>= 3;
            if (isSigned)
                bestBaseCodeword ^= 0x80;
        }

        // byte 0: base codeword, byte 1: multiplier << 4 | table, then the sixteen 3-bit indexes, column-major, MSB first
        uint64_t bits = 0;
        for (int s = 0; s < 16; s++)
        {
            const int px = (s & 3) * 4 + (s >> 2);      // indexes[pixelSelectorOrder[px]] = bestIndexes[px]
            const uint32_t index = (px < 10) ? ((bestIndexes[0] >> (3 * px)) & 7u) : ((bestIndexes[1] >> (3 * (px - 10))) & 7u);
            bits = (bits << 3) | index;
        }
        const uint32_t b0 = (uint32_t)bestBaseCodeword & 0xffu, b1 = (uint32_t)((bestMultiplier << 4) | bestTableIndex) & 0xffu;
        out[0] = (b0 << 24) | (b1 << 16) | (uint32_t)((bits >> 32) & 0xffffu);
        out[1] = (uint32_t)(bits & 0xffffffffu);
    }

    // the byte order of the block: both words are written most significant byte first
    CVTT_HD uint32_t etc_bswap(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
}
