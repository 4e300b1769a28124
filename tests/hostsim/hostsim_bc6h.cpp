// TEST INFRASTRUCTURE ONLY.  CPU build of the BC6H per-thread device code (convectionkernels_b200/csrc/bc6h_core.cuh).
// The eight lanes of a reference group run as eight host threads; the group votes of the kernel (ballots over an
// 8-lane segment) become a spin barrier with an OR reduction.  Not part of the product library.
#include <thread>
#include <vector>
#include <xmmintrin.h>

#include "../../convectionkernels_b200/csrc/bc6h_host.h"
#include "host_vote.h"

using namespace cvttb200;

namespace
{
    // lanesPerWarp = 8: one group at a time (warp_any == group any); 32: four groups share the warp-scope votes like the kernel's warps
    template<bool SIGNED, bool FAST>
    void run_lane(int lane, int lanesPerWarp, GroupShared *groupShared, GroupShared *warpShared, const BC6HParams *P, const int16_t *blocks, size_t nBlocks, uint8_t *out)
    {
        HostWarpVote vote;
        vote.group.g = groupShared;
        vote.warp.g = warpShared;
        const BC6HTables &T = bc6h_tables();
        static const int16_t zeroBlock[64] = { 0 };
        for (size_t base = 0; base < nBlocks; base += lanesPerWarp)
        {
            const bool active = base + lane < nBlocks;          // the kernel's ragged last warp encodes zero blocks and drops them
            const int16_t *src = active ? blocks + (base + lane) * 64 : zeroBlock;
            float pw[48];
            uint32_t raw[32], tab[24];
            BC6HLane<1> L;
            L.pw = pw;
            L.raw = raw;
            L.tab = tab;
            for (int px = 0; px < 16; px++)
                bc6h_load_pixel<SIGNED>(*P, L, px, src[px * 4 + 0], src[px * 4 + 1], src[px * 4 + 2]);
            uint32_t o[4];
            bc6h_encode_block<SIGNED, FAST, 1>(*P, T, L, vote, o);
            if (active)
                memcpy(out + (base + lane) * 16, o, 16);
        }
    }
}

static int encode_bc6h(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable, int lanesPerWarp)
{
    if (nBlocks % 8)
        return -1;
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = rcpTable ? rcpTable[n] : _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    BC6HParams P;
    bc6h_fill_params(P, *options, rcpN);
    const bool fast = (options->flags & kFlag_BC6H_FastIndexing) != 0;
    GroupShared groups[4], warp;
    warp.size = lanesPerWarp;
    std::vector<std::thread> threads;
    for (int lane = 0; lane < lanesPerWarp; lane++)
    {
        if (isSigned)
            threads.emplace_back(fast ? run_lane<true, true> : run_lane<true, false>, lane, lanesPerWarp, &groups[lane / 8], &warp, &P, blocks, nBlocks, out);
        else
            threads.emplace_back(fast ? run_lane<false, true> : run_lane<false, false>, lane, lanesPerWarp, &groups[lane / 8], &warp, &P, blocks, nBlocks, out);
    }
    for (auto &t : threads)
        t.join();
    return 0;
}

extern "C" int hostsim_encode_bc6h(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable)
{
    return encode_bc6h(blocks, nBlocks, out, options, isSigned, rcpTable, 8);
}

// four groups per simulated warp, like the kernel (the exact pruning votes over the warp)
extern "C" int hostsim_encode_bc6h_warp(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable)
{
    return encode_bc6h(blocks, nBlocks, out, options, isSigned, rcpTable, 32);
}

// number of inputs (0..31743, every precision the modes use) for which the integer form of the endpoint quantiser differs from
// the fp32 directed-rounding form the reference executes
extern "C" int hostsim_bc6h_quantizer_mismatches(void)
{
    int bad = 0;
    const int precisions[] = { 6, 7, 8, 9, 10, 11, 12, 16 };
    for (int e = 0; e <= 31743; e++)
        for (int p : precisions)
        {
            bad += bc6h_quantize_element<false>(e, p) != bc6h_quantize_element_reference<false>(e, p);
            bad += bc6h_quantize_element<true>(e, p) != bc6h_quantize_element_reference<true>(e, p);
            bad += bc6h_quantize_element<true>(-e, p) != bc6h_quantize_element_reference<true>(-e, p);
        }
    return bad;
}

// number of 16-bit patterns with exponent below 31 for which half_bits_to_float (the conversion the kernels use: IEEE value, halved
// for exponent 0) differs from the reference's TwosCLHalfToFloat bit arithmetic
extern "C" int hostsim_bc6h_half_conversion_mismatches(void)
{
    int bad = 0;
    for (uint32_t u = 0; u < 65536; u++)
        if ((u & 0x7c00u) != 0x7c00u)
            bad += !(half_bits_to_float(u) == twoscl_half_to_float((int)u));
    return bad;
}
