// Per-thread BC7 / BC6H block decoders (device code, also compiled for the CPU by tests/hostsim).
//
// Reference: cvtt::Kernels::DecodeBC7 / DecodeBC6HU / DecodeBC6HS (ConvectionKernels_API.cpp:288-331) ->
// BC7Computer::UnpackOne (ConvectionKernels_BC67.cpp:2206-2423) and BC6HComputer::UnpackOne (:3059-3289).  The formats are the
// standard ones; what is reference-specific and reproduced here is the handling of invalid blocks (BC7 mode byte 0 -> all zero,
// reserved BC6H modes -> (0, 0, 0, 1.0)), the endpoint bit replication and the BC6H sign handling of the final scale.
//
// One thread decodes one block: 16 bytes in, a PixelBlockU8 (64 B) or PixelBlockF16 (128 B) out.  The kernels around these
// functions are HBM-bound; see cvtt_b200.cu (decode_kernel) for the staging that keeps loads and stores coalesced.
#pragma once

#include "bc7_core.cuh"
#include "bc6h_core.cuh"

namespace cvttb200
{
    // LSB-first reader over the 128 bits of a block (UnpackingVector, BC67.cpp:700-764)
    struct BitReader128
    {
        uint64_t lo, hi;
        CVTT_HD uint32_t take(int bits)
        {
            const uint32_t v = (uint32_t)lo & ((1u << bits) - 1u);
            lo = (lo >> bits) | (hi << (64 - bits));      // bits is 1..31 wherever this is called with a non-zero count
            hi >>= bits;
            return v;
        }
    };

    // interpolation weights of the 2, 3 and 4 bit index sets (g_weight2/3/4, BC67.cpp:121-123), packed 7 bits each
    CVTT_HD int bc_decode_weight(int indexBits, int index)
    {
        if (indexBits == 2)
            return (index * 64 + 1) / 3;                  // 0 21 43 64
        if (indexBits == 3)
            return (index * 64 + 3) / 7;                  // 0 9 18 27 37 46 55 64
        return (index * 64 + 7) / 15;                     // 0 4 9 13 17 21 26 30 34 38 43 47 51 55 60 64
    }

    // g_modes, BC67.cpp:108-119: p-bit mode (0 per endpoint, 1 per subset, 2 none), alpha mode (0 none, 1 combined, 2 separate),
    // rgb bits, alpha bits, partition bits, subsets, index bits, alpha index bits, index selector
    struct BC7DecodeMode
    {
        uint8_t pbits, alpha, rgbBits, alphaBits, partitionBits, subsets, indexBits, alphaIndexBits, hasIndexSelector;
    };

    CVTT_HD BC7DecodeMode bc7_decode_mode(int mode)
    {
        BC7DecodeMode m;
        switch (mode)
        {
        case 0:  m = { 0, 0, 4, 0, 4, 3, 3, 0, 0 }; break;
        case 1:  m = { 1, 0, 6, 0, 6, 2, 3, 0, 0 }; break;
        case 2:  m = { 2, 0, 5, 0, 6, 3, 2, 0, 0 }; break;
        case 3:  m = { 0, 0, 7, 0, 6, 2, 2, 0, 0 }; break;
        case 4:  m = { 2, 2, 5, 6, 0, 1, 2, 3, 1 }; break;
        case 5:  m = { 2, 2, 7, 8, 0, 1, 2, 2, 0 }; break;
        case 6:  m = { 0, 1, 7, 7, 0, 1, 4, 0, 0 }; break;
        default: m = { 0, 1, 5, 5, 6, 2, 2, 0, 0 }; break;
        }
        return m;
    }

    // plain array sink of the decoders: 16-byte chunk q of the block's pixel data
    struct ArraySink
    {
        uint32_t *words;
        CVTT_HD void put4(int q, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const
        {
            words[4 * q] = a;
            words[4 * q + 1] = b;
            words[4 * q + 2] = c;
            words[4 * q + 3] = d;
        }
    };

    // BC7Computer::UnpackOne.  in: the 16 bytes as four little-endian words; out: four chunks of four pixels r | g << 8 | b << 16 | a << 24
    template<class Sink>
    CVTT_HD void bc7_decode_block(const BC7PackTables &T, const uint32_t in[4], const Sink &out)
    {
        BitReader128 br;
        br.lo = (uint64_t)in[0] | ((uint64_t)in[1] << 32);
        br.hi = (uint64_t)in[2] | ((uint64_t)in[3] << 32);

        const uint32_t modeByte = in[0] & 0xffu;
        if (modeByte == 0)
        {
            // no mode bit in the first byte (BC67.cpp:2221-2228)
            for (int q = 0; q < 4; q++)
                out.put4(q, 0, 0, 0, 0);
            return;
        }
        const int mode = ctz32(modeByte);
        br.take(mode + 1);
        const BC7DecodeMode M = bc7_decode_mode(mode);

        const int partition = M.partitionBits ? (int)br.take(M.partitionBits) : 0;
        const int rotation = (M.alpha == 2) ? (int)br.take(2) : 0;
        const int indexSelector = M.hasIndexSelector ? (int)br.take(1) : 0;

        int fixup1 = 0, fixup2 = 0;
        uint32_t subsetMap = 0;                 // 2 bits per pixel
        if (M.subsets == 2)
        {
            fixup1 = T.fixup2[partition];
            const uint32_t mask = T.partitionMask2[partition];
            for (int px = 0; px < 16; px++)
                subsetMap |= ((mask >> px) & 1u) << (2 * px);
        }
        else if (M.subsets == 3)
        {
            fixup1 = T.fixup3[partition * 2];
            fixup2 = T.fixup3[partition * 2 + 1];
            subsetMap = T.partitionMap3[partition];
        }

        // end points: channel-major, then subset, then end point (BC67.cpp:2262-2288)
        int ep[3][2][4];
        for (int ch = 0; ch < 3; ch++)
            for (int s = 0; s < M.subsets; s++)
                for (int e = 0; e < 2; e++)
                    ep[s][e][ch] = (int)(br.take(M.rgbBits) << (8 - M.rgbBits));
        for (int s = 0; s < M.subsets; s++)
            for (int e = 0; e < 2; e++)
                ep[s][e][3] = M.alpha ? (int)(br.take(M.alphaBits) << (8 - M.alphaBits)) : 255;

        int parityBits = 0;
        if (M.pbits != 2)
        {
            parityBits = 1;
            for (int s = 0; s < M.subsets; s++)
            {
                int p = 0;
                if (M.pbits == 1)
                    p = (int)br.take(1);
                for (int e = 0; e < 2; e++)
                {
                    if (M.pbits == 0)
                        p = (int)br.take(1);
                    for (int ch = 0; ch < 3; ch++)
                        ep[s][e][ch] |= p << (7 - M.rgbBits);
                    if (M.alpha)
                        ep[s][e][3] |= p << (7 - M.alphaBits);
                }
            }
        }
        // replicate the top bits into the bits the mode does not store (BC67.cpp:2330-2340)
        for (int s = 0; s < M.subsets; s++)
            for (int e = 0; e < 2; e++)
            {
                for (int ch = 0; ch < 3; ch++)
                    ep[s][e][ch] |= ep[s][e][ch] >> (M.rgbBits + parityBits);
                if (M.alpha)
                    ep[s][e][3] |= ep[s][e][3] >> (M.alphaBits + parityBits);
            }

        int idx[16], idx2[16];
        for (int px = 0; px < 16; px++)
        {
            const bool anchor = (px == 0) || (px == fixup1) || (px == fixup2);
            idx[px] = (int)br.take(M.indexBits - (anchor ? 1 : 0));
        }
        for (int px = 0; px < 16; px++)
            idx2[px] = (M.alpha == 2) ? (int)br.take(M.alphaIndexBits - (px == 0 ? 1 : 0)) : 0;

        uint32_t row[4];
#pragma unroll
        for (int px = 0; px < 16; px++)
        {
            int rgbWeight = bc_decode_weight(M.indexBits, idx[px]), alphaWeight = 0;
            if (M.alpha == 1)
                alphaWeight = rgbWeight;
            else if (M.alpha == 2)
                alphaWeight = bc_decode_weight(M.alphaIndexBits, idx2[px]);
            if (indexSelector)
            {
                const int t = rgbWeight;
                rgbWeight = alphaWeight;
                alphaWeight = t;
            }
            const int s = (int)((subsetMap >> (2 * px)) & 3u);
            int pixel[4];
            for (int ch = 0; ch < 3; ch++)
                pixel[ch] = ((64 - rgbWeight) * ep[s][0][ch] + rgbWeight * ep[s][1][ch] + 32) >> 6;
            pixel[3] = M.alpha ? (((64 - alphaWeight) * ep[s][0][3] + alphaWeight * ep[s][1][3] + 32) >> 6) : 255;
            if (rotation)
            {
                const int t = pixel[rotation - 1];
                pixel[rotation - 1] = pixel[3];
                pixel[3] = t;
            }
            row[px & 3] = (uint32_t)(pixel[0] & 0xff) | ((uint32_t)(pixel[1] & 0xff) << 8) | ((uint32_t)(pixel[2] & 0xff) << 16) | ((uint32_t)(pixel[3] & 0xff) << 24);
            if ((px & 3) == 3)
                out.put4(px >> 2, row[0], row[1], row[2], row[3]);
        }
    }

    CVTT_HD int bc6h_sign_extend(int v, int bits)
    {
        return (v & (1 << (bits - 1))) ? (v | -(1 << bits)) : v;
    }

    // BC6HComputer::UnpackOne.  out: eight chunks of two pixels, a pixel = 2 words (r | g << 16, b | 0x3c00 << 16): PixelBlockF16
    template<class Sink>
    CVTT_HD void bc6h_decode_block(const BC6HTables &T, const uint32_t in[4], bool isSigned, const Sink &out)
    {
        int modeBits = (int)(in[0] & 3u);
        if (modeBits > 1)
            modeBits = (int)(in[0] & 0x1fu);
        int mode = -1;
        for (int m = 0; m < 14; m++)
            if (T.modes[m][0] == modeBits)
            {
                mode = m;
                break;
            }
        if (mode < 0)
        {
            // reserved mode ids (BC67.cpp:3079-3088)
            for (int q = 0; q < 8; q++)
                out.put4(q, 0, 0x3c000000u, 0, 0x3c000000u);
            return;
        }
        const bool partitioned = T.modes[mode][1] != 0, transformed = T.modes[mode][2] != 0;
        const int aPrec = T.modes[mode][3];
        const int headerBits = partitioned ? 82 : 65;

        // gather the header fields: header bit i is bit (v & 15) of field (v >> 4); fields m d rw rx ry rz gw gx gy gz bw bx by bz
        int field[14];
        for (int f = 0; f < 14; f++)
            field[f] = 0;
        for (int i = 0; i < headerBits; i++)
        {
            const int v = T.headerBits[mode][i];
            const uint32_t bit = (in[i >> 5] >> (i & 31)) & 1u;
            field[v >> 4] |= (int)(bit << (v & 15));
        }
        const int partition = partitioned ? field[1] : 0;
        int eps[2][2][3];
        for (int ch = 0; ch < 3; ch++)
        {
            eps[0][0][ch] = field[2 + ch * 4];
            eps[0][1][ch] = field[3 + ch * 4];
            eps[1][0][ch] = field[4 + ch * 4];
            eps[1][1][ch] = field[5 + ch * 4];
        }

        BitReader128 br;
        br.lo = (uint64_t)in[0] | ((uint64_t)in[1] << 32);
        br.hi = (uint64_t)in[2] | ((uint64_t)in[3] << 32);
        // skip the header (65 or 82 bits)
        br.lo = br.hi;
        br.hi = 0;
        br.take(headerBits - 64);

        const int fixup = partitioned ? T.fixup[partition] : 0;
        const int indexBits = partitioned ? 3 : 4, numSubsets = partitioned ? 2 : 1;
        int idx[16];
        for (int px = 0; px < 16; px++)
            idx[px] = (int)br.take(indexBits - ((px == 0 || px == fixup) ? 1 : 0));

        for (int ch = 0; ch < 3; ch++)
        {
            const int bPrec = T.modes[mode][4 + ch];
            if (isSigned)
                eps[0][0][ch] = bc6h_sign_extend(eps[0][0][ch], aPrec);
            if (transformed || isSigned)
            {
                eps[0][1][ch] = bc6h_sign_extend(eps[0][1][ch], bPrec);
                if (partitioned)
                {
                    eps[1][0][ch] = bc6h_sign_extend(eps[1][0][ch], bPrec);
                    eps[1][1][ch] = bc6h_sign_extend(eps[1][1][ch], bPrec);
                }
            }
            if (transformed)
            {
                const int wrapMask = (1 << aPrec) - 1;
                for (int k = 1; k < (partitioned ? 4 : 2); k++)
                {
                    int &e = eps[k >> 1][k & 1][ch];
                    e = (eps[0][0][ch] + e) & wrapMask;
                    if (isSigned)
                        e = bc6h_sign_extend(e, aPrec);
                }
            }
        }

        // unquantise the end points (BC67.cpp:3192-3250)
        for (int s = 0; s < numSubsets; s++)
            for (int e = 0; e < 2; e++)
                for (int ch = 0; ch < 3; ch++)
                {
                    int v = eps[s][e][ch];
                    if (isSigned)
                    {
                        if (aPrec < 16)
                        {
                            const bool neg = v < 0;
                            const int comp = neg ? -v : v;
                            int unq;
                            if (comp == 0)
                                unq = 0;
                            else if (comp >= ((1 << (aPrec - 1)) - 1))
                                unq = 0x7fff;
                            else
                                unq = ((comp << 15) + 0x4000) >> (aPrec - 1);
                            v = neg ? -unq : unq;
                        }
                    }
                    else if (aPrec < 15 && v != 0)
                        v = (v == ((1 << aPrec) - 1)) ? 0xffff : (((v << 16) + 0x8000) >> aPrec);
                    eps[s][e][ch] = v;
                }

        const uint32_t mask = partitioned ? T.partitionMask[partition] : 0u;
        uint32_t pair[4];
#pragma unroll
        for (int px = 0; px < 16; px++)
        {
            const int s = (int)((mask >> px) & 1u);
            const int w = bc_decode_weight(indexBits, idx[px]);
            uint32_t h[3];
            for (int ch = 0; ch < 3; ch++)
            {
                int comp = ((64 - w) * eps[s][0][ch] + w * eps[s][1][ch] + 32) >> 6;
                if (isSigned)
                {
                    comp = (comp < 0) ? -(((-comp) * 31) >> 5) : ((comp * 31) >> 5);
                    uint32_t sign = 0;
                    if (comp < 0)
                    {
                        sign = 0x8000u;
                        comp = -comp;
                    }
                    h[ch] = (sign | (uint32_t)comp) & 0xffffu;
                }
                else
                    h[ch] = (uint32_t)((comp * 31) >> 6) & 0xffffu;
            }
            pair[(px & 1) * 2] = h[0] | (h[1] << 16);
            pair[(px & 1) * 2 + 1] = h[2] | 0x3c000000u;
            if (px & 1)
                out.put4(px >> 1, pair[0], pair[1], pair[2], pair[3]);
        }
    }
}
